// No-op stand-in for boost::progress_display (only used by the DB builders).
#pragma once
namespace boost {
struct progress_display {
  progress_display(unsigned long) {}
  progress_display& operator++() { return *this; }
};
}
