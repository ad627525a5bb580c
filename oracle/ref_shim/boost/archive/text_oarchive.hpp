// Stand-in for Boost.Serialization (no Boost headers in this image). The reference's
// (de)serialisation is outside the matching path; these archives write nothing and only let its
// headers compile unmodified for the oracle/_ref harness. The one thing they do: hand the object
// given to operator<< to a hook, so the harness can look at the GenbankIndex the reference's
// database builders (createIndexFromGBFF / createIndexFromFASTA) construct before "serialising" it.
#pragma once
#include <ostream>
namespace kref_shim {
typedef void (*archive_hook_fn)(const void *object);
inline archive_hook_fn &archive_hook() { static archive_hook_fn hook = nullptr; return hook; }
}
namespace boost { namespace archive {
struct text_oarchive {
  text_oarchive(std::ostream&) {}
  template <class T> text_oarchive& operator<<(const T& t) { if (kref_shim::archive_hook()) kref_shim::archive_hook()(&t); return *this; }
  template <class T> text_oarchive& operator&(const T&) { return *this; }
};
}}
