// Declarations shared by the real-library-backed text_oarchive / text_iarchive stand-ins — TEST INFRASTRUCTURE ONLY.
// Re-declared with Boost.Serialization 1.78's own names and layouts so that the constructors / destructors exported by
// libboost_serialization.so link; see text_oarchive.hpp.
#pragma once
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

namespace boost {
namespace serialization {
class extended_type_info {
 private:
  virtual bool is_less_than(const extended_type_info &) const = 0;
  virtual bool is_equal(const extended_type_info &) const = 0;
  const unsigned int m_type_info_key;
  const char *m_key;
 protected:
  extended_type_info(const unsigned int type_info_key, const char *key);
  virtual ~extended_type_info();
 public:
  virtual const char *get_debug_info() const = 0;
  virtual void *construct(unsigned int = 0, ...) const = 0;
  virtual void destroy(void const *const) const = 0;
};
class item_version_type {
  unsigned int t;
 public:
  item_version_type() : t(0) {}
  explicit item_version_type(const unsigned int &t_) : t(t_) {}
  item_version_type(const item_version_type &o) : t(o.t) {}
};
}  // namespace serialization
namespace archive {
class version_type {
  uint_least32_t t;
 public:
  version_type() : t(0) {}
  explicit version_type(const unsigned int &t_) : t(t_) {}
  version_type(const version_type &o) : t(o.t) {}
};
namespace detail {
class basic_serializer {
  const boost::serialization::extended_type_info *m_eti;
 protected:
  explicit basic_serializer(const boost::serialization::extended_type_info &eti) : m_eti(&eti) {}
};
}  // namespace detail
namespace kref_detail {
struct type_key : boost::serialization::extended_type_info {   // one key per serialised type: the library orders by key first
  explicit type_key(unsigned key) : extended_type_info(key, nullptr) {}
  bool is_less_than(const extended_type_info &) const override { return false; }
  bool is_equal(const extended_type_info &) const override { return false; }
  const char *get_debug_info() const override { return "kref"; }
  void *construct(unsigned int, ...) const override { return nullptr; }
  void destroy(void const *const) const override {}
};
inline unsigned next_key() { static unsigned k = 1000; return ++k; }
template <class T> struct is_vector : std::false_type {};
template <class T, class A> struct is_vector<std::vector<T, A>> : std::true_type {};
}  // namespace kref_detail
}  // namespace archive
}  // namespace boost
