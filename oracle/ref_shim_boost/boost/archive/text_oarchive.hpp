// boost/archive/text_oarchive.hpp backed by the REAL Boost.Serialization library — TEST INFRASTRUCTURE ONLY.
//
// This image has no Boost headers, but libboost_serialization.so.1.78.0 ships inside Nsight Compute. This header gives the
// reference's unmodified writeIndexToBoostSerial (GenbankTools.h:201-205: `boost::archive::text_oarchive oa(ofs); oa << *this;`)
// a text_oarchive that forwards to that library: the archive object is built by the library's own text_oarchive_impl
// constructor, every class object goes through the library's basic_oarchive::save_object (which writes the class preamble),
// strings and item versions through the library's save(), tokens through its newtoken(). What this header restates from
// Boost's header templates is the dispatch only: a class type is saved through a serializer with Boost's default traits
// (class info on, tracking off, version 0) whose body calls the class's own serialize(); a std::vector is count,
// item_version, items; arithmetic types are a token and operator<<. The ORDER of the members comes from the reference's code.
#pragma once
#include <cstdint>
#include <cstring>
#include <new>
#include <ostream>
#include <string>
#include <type_traits>
#include <vector>

#include "kref_boost_decls.hpp"

namespace boost {
namespace archive {
namespace detail {
class basic_oarchive;
class basic_pointer_oserializer;
class basic_oserializer : public basic_serializer {
  basic_pointer_oserializer *m_bpos;
 protected:
  explicit basic_oserializer(const boost::serialization::extended_type_info &type_);
  virtual ~basic_oserializer();
 public:
  virtual void save_object_data(basic_oarchive &ar, const void *x) const = 0;
  virtual bool class_info() const = 0;
  virtual bool tracking(const unsigned int flags) const = 0;
  virtual version_type version() const = 0;
  virtual bool is_polymorphic() const = 0;
};
class basic_oarchive {
 public:
  void save_object(const void *x, const basic_oserializer &bos);
  void end_preamble();
};
}  // namespace detail
class text_oarchive;
template <class Archive> class basic_text_oarchive {
 public:
  void init();
  void newtoken();
};
template <class Archive> class text_oarchive_impl {
 public:
  text_oarchive_impl(std::ostream &os, unsigned int flags);
  ~text_oarchive_impl();
  void save(const std::string &s);
  void save(const boost::serialization::item_version_type &t);
};


class text_oarchive {
  alignas(64) unsigned char raw_[8192];                     // the library's text_oarchive_impl<text_oarchive>, at offset 0
  std::ostream &os_;
  typedef text_oarchive_impl<text_oarchive> impl_t;
  typedef basic_text_oarchive<text_oarchive> text_t;
  detail::basic_oarchive *base() { return reinterpret_cast<detail::basic_oarchive *>(raw_); }

  template <class T> struct serializer : detail::basic_oserializer {
    serializer() : detail::basic_oserializer(key_ref()) {}
    // the key object must exist before the base is constructed: a function-local static per type
    static const kref_detail::type_key &key_ref() { static kref_detail::type_key k(kref_detail::next_key()); return k; }
    void save_object_data(detail::basic_oarchive &ar, const void *x) const override {
      reinterpret_cast<text_oarchive *>(&ar)->body(*static_cast<const T *>(x));
    }
    bool class_info() const override { return true; }
    bool tracking(const unsigned int) const override { return false; }
    version_type version() const override { return version_type(0); }
    bool is_polymorphic() const override { return false; }
    static const serializer &instance() { static serializer s; return s; }
  };
  // body of a class object: its own serialize(); of a vector: count, item_version, items (boost/serialization/vector.hpp)
  template <class T> typename std::enable_if<!kref_detail::is_vector<T>::value>::type body(const T &t) { const_cast<T &>(t).serialize(*this, 0u); }
  template <class T> typename std::enable_if<kref_detail::is_vector<T>::value>::type body(const T &v) {
    save(static_cast<std::size_t>(v.size()));
    reinterpret_cast<impl_t *>(raw_)->save(boost::serialization::item_version_type(0));
    for (const auto &item : v) save(item);
  }
  void save(const std::string &s) { base()->end_preamble(); reinterpret_cast<impl_t *>(raw_)->save(s); }
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type save(const T &v) {
    base()->end_preamble();
    reinterpret_cast<text_t *>(raw_)->newtoken();
    os_ << v;
  }
  template <class T> typename std::enable_if<std::is_class<T>::value && !std::is_same<T, std::string>::value>::type save(const T &t) {
    base()->save_object(&t, serializer<T>::instance());
  }

 public:
  explicit text_oarchive(std::ostream &os) : os_(os) {
    std::memset(raw_, 0, sizeof raw_);
    new (raw_) impl_t(os, 0);
    reinterpret_cast<text_t *>(raw_)->init();
  }
  ~text_oarchive() { reinterpret_cast<impl_t *>(raw_)->~impl_t(); }
  template <class T> text_oarchive &operator<<(const T &t) { save(t); return *this; }
  template <class T> text_oarchive &operator&(const T &t) { save(t); return *this; }
};
}  // namespace archive
}  // namespace boost
