// Minimal stand-in for boost/program_options.hpp — TEST INFRASTRUCTURE ONLY. Entirely ours (no Boost code, and not the
// Boost library): just the subset the reference's main.cpp uses, so that its unmodified main() compiles into
// oracle/_ref/SLAM_ref. Long options only ("--name value", "--name=value", unambiguous prefixes), positional arguments
// collected under the name given to positional_options_description::add, typed values with default_value and an optional
// bound variable, variables_map with count() and operator[]().as<T>().
#pragma once
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

struct value_base {
  virtual ~value_base() {}
  virtual void parse(const std::string &text) = 0;      // one occurrence on the command line
  virtual bool apply_default() = 0;                      // true when a default exists (and is now the value)
  virtual void notify() {}
  virtual std::string default_text() const { return std::string(); }
};
template <class T> struct value_parser {
  static void parse(T &out, const std::string &text) {
    std::istringstream in(text);
    in >> out;
    if (in.fail() || !in.eof()) throw std::runtime_error("the argument ('" + text + "') is invalid");
  }
};
template <> struct value_parser<std::string> { static void parse(std::string &out, const std::string &text) { out = text; } };
template <class T> struct typed_value : value_base {
  T current{}; bool has_default = false; T def{}; T *bound = nullptr; std::string def_text;
  explicit typed_value(T *b) : bound(b) {}
  typed_value *default_value(const T &v) { has_default = true; def = v; std::ostringstream o; o << v; def_text = o.str(); return this; }
  void parse(const std::string &text) override { value_parser<T>::parse(current, text); }
  bool apply_default() override { if (has_default) current = def; return has_default; }
  void notify() override { if (bound) *bound = current; }
  std::string default_text() const override { return def_text; }
};
template <class T> struct typed_value<std::vector<T>> : value_base {
  std::vector<T> current; std::vector<T> *bound = nullptr;
  explicit typed_value(std::vector<T> *b) : bound(b) {}
  void parse(const std::string &text) override { T v; value_parser<T>::parse(v, text); current.push_back(v); }
  bool apply_default() override { return false; }
  void notify() override { if (bound) *bound = current; }
};
template <class T> typed_value<T> *value() { return new typed_value<T>(nullptr); }
template <class T> typed_value<T> *value(T *bound) { return new typed_value<T>(bound); }

struct option_description {
  std::string name, text;
  std::shared_ptr<value_base> semantic;                  // null: a flag without argument
};
class options_description;
class options_description_easy_init {
  options_description *owner_;
 public:
  explicit options_description_easy_init(options_description *o) : owner_(o) {}
  options_description_easy_init &operator()(const char *name, const char *text);
  options_description_easy_init &operator()(const char *name, value_base *semantic, const char *text);
};
class options_description {
 public:
  std::string caption;
  std::vector<option_description> options;
  explicit options_description(const std::string &c = std::string()) : caption(c) {}
  options_description_easy_init add_options() { return options_description_easy_init(this); }
  options_description &add(const options_description &other) { options.insert(options.end(), other.options.begin(), other.options.end()); return *this; }
};
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *text) {
  owner_->options.push_back(option_description{name, text, nullptr});
  return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, value_base *semantic, const char *text) {
  owner_->options.push_back(option_description{name, text, std::shared_ptr<value_base>(semantic)});
  return *this;
}
inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
  os << d.caption << ":\n";
  for (const auto &o : d.options) {
    std::string left = "  --" + o.name;
    if (o.semantic) { left += " arg"; if (!o.semantic->default_text().empty()) left += " (=" + o.semantic->default_text() + ")"; }
    os << left << "  " << o.text << "\n";
  }
  return os;
}

class positional_options_description {
 public:
  std::string name;
  positional_options_description &add(const char *n, int) { name = n; return *this; }
};

struct variable_value {
  std::shared_ptr<value_base> semantic;
  template <class T> const T &as() const {
    auto *tv = dynamic_cast<typed_value<T> *>(semantic.get());
    if (!tv) throw std::runtime_error("bad any_cast");
    return tv->current;
  }
};
class variables_map {
 public:
  std::map<std::string, variable_value> values;
  std::vector<std::shared_ptr<value_base>> to_notify;
  size_t count(const std::string &name) const { return values.count(name); }
  const variable_value &operator[](const std::string &name) const {
    static const variable_value empty;
    auto it = values.find(name);
    return it == values.end() ? empty : it->second;
  }
};

struct parsed_options {
  const options_description *desc = nullptr;
  std::vector<std::pair<std::string, std::string>> found;   // (option name, text; empty text for flags)
};
class command_line_parser {
  std::vector<std::string> args_;
  const options_description *desc_ = nullptr;
  std::string positional_;
 public:
  command_line_parser(int argc, const char *const argv[]) { for (int i = 1; i < argc; i++) args_.push_back(argv[i]); }
  command_line_parser &options(const options_description &d) { desc_ = &d; return *this; }
  command_line_parser &positional(const positional_options_description &p) { positional_ = p.name; return *this; }
  parsed_options run() {
    parsed_options out;
    out.desc = desc_;
    for (size_t i = 0; i < args_.size(); i++) {
      const std::string &a = args_[i];
      if (a.size() < 3 || a.compare(0, 2, "--") != 0) { out.found.push_back({positional_, a}); continue; }
      std::string name = a.substr(2), text;
      bool has_text = false;
      const size_t eq = name.find('=');
      if (eq != std::string::npos) { text = name.substr(eq + 1); name = name.substr(0, eq); has_text = true; }
      const option_description *match = nullptr;
      int n_match = 0;
      for (const auto &o : desc_->options) {
        if (o.name == name) { match = &o; n_match = 1; break; }
        if (o.name.compare(0, name.size(), name) == 0) { match = &o; n_match++; }
      }
      if (n_match == 0) throw std::runtime_error("unrecognised option '--" + name + "'");
      if (n_match > 1) throw std::runtime_error("option '--" + name + "' is ambiguous");
      if (match->semantic && !has_text) {
        if (i + 1 >= args_.size()) throw std::runtime_error("the required argument for option '--" + match->name + "' is missing");
        text = args_[++i];
      }
      out.found.push_back({match->name, match->semantic ? text : std::string()});
    }
    return out;
  }
};
inline void store(const parsed_options &parsed, variables_map &vm) {
  for (const auto &f : parsed.found) {
    const option_description *od = nullptr;
    for (const auto &o : parsed.desc->options) if (o.name == f.first) od = &o;
    if (!od) throw std::runtime_error("too many positional options have been specified on the command line");
    variable_value &v = vm.values[od->name];
    if (od->semantic) { v.semantic = od->semantic; od->semantic->parse(f.second); }
  }
  for (const auto &o : parsed.desc->options)
    if (o.semantic) {
      if (!vm.values.count(o.name) && o.semantic->apply_default()) vm.values[o.name].semantic = o.semantic;
      vm.to_notify.push_back(o.semantic);
    }
}
inline void notify(variables_map &vm) { for (auto &s : vm.to_notify) if (s) s->notify(); }

}  // namespace program_options
}  // namespace boost
