// See text_oarchive.hpp: no-op stand-in so the reference headers compile.
#pragma once
#include <istream>
namespace boost { namespace archive {
struct text_iarchive {
  text_iarchive(std::istream&) {}
  template <class T> text_iarchive& operator>>(T&) { return *this; }
  template <class T> text_iarchive& operator&(T&) { return *this; }
};
}}
