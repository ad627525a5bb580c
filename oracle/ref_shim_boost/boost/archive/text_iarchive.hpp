// boost/archive/text_iarchive.hpp backed by the REAL Boost.Serialization library — TEST INFRASTRUCTURE ONLY.
//
// The reading side of text_oarchive.hpp: the reference's unmodified getIndexFromBoostSerial (GenbankTools.h:336-344:
// `boost::archive::text_iarchive ia(ifs); ia >> index;`) gets a text_iarchive whose archive object is the library's own
// text_iarchive_impl (its init() reads and checks the header), whose class objects go through the library's
// basic_iarchive::load_object (which reads the class preambles), and whose strings and item versions are read by the
// library's load(). Restated from Boost's header templates: the dispatch (a class type loads through a serializer with the
// default traits whose body calls the class's own serialize(); a std::vector is count, item_version, items; arithmetic
// types are operator>>).
#pragma once
#include <istream>
#include <stdexcept>

#include "kref_boost_decls.hpp"

namespace boost {
namespace archive {
namespace detail {
class basic_iarchive;
class basic_pointer_iserializer;
class basic_iserializer : public basic_serializer {
  basic_pointer_iserializer *m_bpis;
 protected:
  explicit basic_iserializer(const boost::serialization::extended_type_info &type);
  virtual ~basic_iserializer();
 public:
  virtual void load_object_data(basic_iarchive &ar, void *x, const unsigned int file_version) const = 0;
  virtual bool class_info() const = 0;
  virtual bool tracking(const unsigned int) const = 0;
  virtual version_type version() const = 0;
  virtual bool is_polymorphic() const = 0;
  virtual void destroy(void *address) const = 0;
};
class basic_iarchive {
 public:
  void load_object(void *t, const basic_iserializer &bis);
};
}  // namespace detail
class text_iarchive;
template <class Archive> class text_iarchive_impl {
 public:
  text_iarchive_impl(std::istream &is, unsigned int flags);
  ~text_iarchive_impl();
  void init();
  void load(std::string &s);
  void load(boost::serialization::item_version_type &t);
};

class text_iarchive {
  alignas(64) unsigned char raw_[8192];                     // the library's text_iarchive_impl<text_iarchive>, at offset 0
  std::istream &is_;
  typedef text_iarchive_impl<text_iarchive> impl_t;
  detail::basic_iarchive *base() { return reinterpret_cast<detail::basic_iarchive *>(raw_); }

  template <class T> struct serializer : detail::basic_iserializer {
    serializer() : detail::basic_iserializer(key_ref()) {}
    static const kref_detail::type_key &key_ref() { static kref_detail::type_key k(kref_detail::next_key()); return k; }
    void load_object_data(detail::basic_iarchive &ar, void *x, const unsigned int file_version) const override {
      reinterpret_cast<text_iarchive *>(&ar)->body(*static_cast<T *>(x), file_version);
    }
    bool class_info() const override { return true; }
    bool tracking(const unsigned int) const override { return false; }
    version_type version() const override { return version_type(0); }
    bool is_polymorphic() const override { return false; }
    void destroy(void *address) const override { delete static_cast<T *>(address); }
    static const serializer &instance() { static serializer s; return s; }
  };
  template <class T> typename std::enable_if<!kref_detail::is_vector<T>::value>::type body(T &t, unsigned version) { t.serialize(*this, version); }
  template <class T> typename std::enable_if<kref_detail::is_vector<T>::value>::type body(T &v, unsigned) {   // boost/serialization/vector.hpp
    std::size_t count = 0;
    load(count);
    boost::serialization::item_version_type item_version;
    reinterpret_cast<impl_t *>(raw_)->load(item_version);   // library versions > 3 store it
    v.clear();
    v.resize(count);
    for (auto &item : v) load(item);
  }
  void load(std::string &s) { reinterpret_cast<impl_t *>(raw_)->load(s); }
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type load(T &v) {
    if (!(is_ >> v)) throw std::runtime_error("input stream error");
  }
  template <class T> typename std::enable_if<std::is_class<T>::value && !std::is_same<T, std::string>::value>::type load(T &t) {
    base()->load_object(&t, serializer<T>::instance());
  }

 public:
  explicit text_iarchive(std::istream &is) : is_(is) {
    std::memset(raw_, 0, sizeof raw_);
    new (raw_) impl_t(is, 0);
    reinterpret_cast<impl_t *>(raw_)->init();
  }
  ~text_iarchive() { reinterpret_cast<impl_t *>(raw_)->~impl_t(); }
  template <class T> text_iarchive &operator>>(T &t) { load(t); return *this; }
  template <class T> text_iarchive &operator&(T &t) { load(t); return *this; }
};
}  // namespace archive
}  // namespace boost
