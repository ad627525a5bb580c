// Stand-in for Boost.Serialization (no Boost headers in this image). The reference's
// (de)serialisation is outside the matching path; these no-op archives only let its
// headers compile unmodified for the oracle/_ref harness.
#pragma once
#include <ostream>
namespace boost { namespace archive {
struct text_oarchive {
  text_oarchive(std::ostream&) {}
  template <class T> text_oarchive& operator<<(const T&) { return *this; }
  template <class T> text_oarchive& operator&(const T&) { return *this; }
};
}}
