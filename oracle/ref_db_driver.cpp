// oracle/_ref/libkslam_refdb.so — TEST INFRASTRUCTURE ONLY.
// The reference's OWN database builders, createIndexFromGBFF / createIndexFromFASTA (GenbankTools.h:224-260,481-527), writing
// a real archive: its unmodified writeIndexToBoostSerial runs over oracle/ref_shim_boost's text_oarchive, which forwards to the
// real Boost.Serialization library found in this image (see that header). Nothing of the reference is copied; it is
// #included where it lies.
#include <vector>
#include <string>
#include <fstream>
#include <iostream>
#include <sstream>
#include <algorithm>
#include <stdexcept>
#include <limits>
#include <memory>
#include <set>
#include <climits>
#include <cmath>
#include <numeric>
#include <unordered_map>
#include <array>
#include <map>
#include <tuple>
#include <cstring>
#include <cstdint>
#include <omp.h>
#include "SLAM.h"

extern "C" int kref_write_database(int kind, const char *const *paths, uint64_t n, const char *out_path) {
  std::vector<std::string> names(paths, paths + n);
  try {
    if (kind == 0) SLAM::createIndexFromGBFF(names, out_path); else SLAM::createIndexFromFASTA(names, out_path);
    return 0;
  } catch (...) { return -1; }
}

// The reference's getIndexFromBoostSerial (GenbankTools.h:336-344) reading `path` through the real library; the index it
// loaded comes back in the dump format of ref_driver.cpp (fields separated by 0x1f, records by 0x1e). Returns the dump
// length, or (uint64_t)-1 when the reference / the library throws.
extern "C" uint64_t kref_read_database(const char *path, char *buf, uint64_t cap) {
  try {
    SLAM::GenbankIndex index = SLAM::getIndexFromBoostSerial(path);
    std::string t;
    const char F = 0x1f, R = 0x1e;
    for (auto &e : index.entries) {
      t += "E"; t += F; t += e.locusTag; t += F; t += std::to_string(e.taxonomyID); t += F; t += std::to_string(e.genbankID); t += F;
      t += std::to_string((int)e.isPlasmid); t += F; t += std::to_string((int)e.is16S); t += F; t += e.bases; t += R;
      for (auto &g : e.genes) {
        t += "G"; t += F; t += g.geneName; t += F; t += g.locusTag; t += F; t += g.proteinID; t += F; t += g.product; t += F;
        t += g.referenceSequence; t += F; t += std::to_string(g.geneID); t += F; t += std::to_string(g.codingSequence.start); t += F;
        t += std::to_string(g.codingSequence.stop); t += F; t += std::to_string((int)g.codingSequence.complement); t += R;
      }
    }
    if (buf && cap >= t.size()) memcpy(buf, t.data(), t.size());
    return t.size();
  } catch (...) { return (uint64_t)-1; }
}
