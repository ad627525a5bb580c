// host_stages.h — types shared by the host stages that follow pairing (sam.cu: screens, pseudo-assembly, SAM records;
// taxon.cu: LCA, genes, XML). Host code only.
#pragma once
#include "../../include/kslam.h"
#include <algorithm>
#include <stdint.h>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

namespace kslam_host {

struct POv {                       // PairedOverlap, /root/reference/src/PairedOverlap.h:32-58, overlaps held as indices
  uint32_t combinedScore = 0, entry = 0;
  int refStart = 0, refEnd = 0;
  uint32_t insertSize = 0;
  bool hasR1 = false, hasR2 = false;
  int32_t r1 = -1, r2 = -1;        // index into sorted_overlaps
};
struct ReadPair { uint32_t r1Pos = 0, r2Pos = 0; std::vector<POv> pairs; };   // ReadPairAndOverlaps, PairedOverlap.h:60-66

// contiguous ranges of [0, n) on `threads` host threads (every stage is independent per read pair or per entry)
template <class F> void parallel_ranges(uint32_t threads, size_t n, F f) {
  if (threads <= 1 || n < 2 * (size_t)threads) { f(0, (size_t)0, n); return; }
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < threads; t++) th.emplace_back([=] { f(t, n * t / threads, n * (t + 1) / threads); });
  for (auto &x : th) x.join();
}

// f(t) once on each of `threads` host threads
template <class F> void parallel_threads(uint32_t threads, F f) {
  if (threads <= 1) { f(0u); return; }
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < threads; t++) th.emplace_back([=] { f(t); });
  for (auto &x : th) x.join();
}

struct Ctx {
  const kslam_sam_params *prm; const kslam_sam_db *db; const kslam_read_batch *reads; const kslam_pairs *in;
  bool paired;                     // Globals.h pairedData
  // compact input (kslam_pairs_compact, runs without --sam-file): `in` is null, POv::r1 / r2 hold the PAIR index, and the
  // mates of the pairs beyond the insert-size limit come from `far` (ascending pair_index)
  const kslam_far_mates *far = nullptr; uint64_t n_far = 0;
  std::string read_id(uint32_t i) const { return std::string(reads->ids + reads->id_offs[i], reads->ids + reads->id_offs[i + 1]); }
  std::string locus(uint32_t e) const { return std::string(db->locus_tags + db->locus_offs[e], db->locus_tags + db->locus_offs[e + 1]); }
};

inline std::string_view gene_str(const kslam_sam_db *db, const kslam_gene *g, int which) {
  return std::string_view(db->gene_strings + g->str_offs[which], (size_t)(g->str_offs[which + 1] - g->str_offs[which]));
}
enum { GENE_NAME = 0, GENE_LOCUS = 1, GENE_PROTEIN = 2, GENE_PRODUCT = 3, GENE_REFERENCE = 4 };

}  // namespace kslam_host
struct kslam_gene_index {
  std::vector<uint8_t> sorted;       // per entry: genes ordered by cds_start (as int), every start / stop <= INT32_MAX
  std::vector<int32_t> max_stop;     // per gene: largest cds_stop among the entry's genes up to and including this one
};
namespace kslam_host {

// GenbankEntry::getGene, GenbankTools.h:170-185: the gene with the largest (strictly positive) overlap with
// [startPos, endPos], first one on ties; nullptr when the entry has no gene table or nothing overlaps.
inline const kslam_gene *best_gene(const kslam_sam_db *db, uint32_t entry, int32_t startPos, int32_t endPos) {
  if (!db->genes || !db->gene_offs) return nullptr;
  const kslam_gene *bestMatch = nullptr;
  int32_t largestOverlap = 0;
  const uint64_t g0 = db->gene_offs[entry], g1 = db->gene_offs[entry + 1];
  if (db->gene_index && db->gene_index->sorted[entry]) {
    // genes that can overlap start before endPos: binary search for the first start >= endPos, then walk back while some
    // gene at or before the position still ends after startPos. Walking back with >= keeps the FIRST gene among equals.
    uint64_t lo = g0, hi = g1;
    while (lo < hi) { const uint64_t mid = (lo + hi) / 2; if ((int32_t)db->genes[mid].cds_start < endPos) lo = mid + 1; else hi = mid; }
    const int32_t *max_stop = db->gene_index->max_stop.data();
    for (uint64_t gi = lo; gi > g0 && max_stop[gi - 1] > startPos; gi--) {
      const kslam_gene *g = db->genes + (gi - 1);
      const int32_t numBasesOverlap = std::min<int>(endPos, (int)g->cds_stop) - std::max<int>(startPos, (int)g->cds_start);
      if (numBasesOverlap > 0 && numBasesOverlap >= largestOverlap) { bestMatch = g; largestOverlap = numBasesOverlap; }
    }
    return bestMatch;
  }
  for (uint64_t gi = g0; gi < g1; gi++) {
    const kslam_gene *g = db->genes + gi;
    const int32_t numBasesOverlap = std::min<int>(endPos, (int)g->cds_stop) - std::max<int>(startPos, (int)g->cds_start);
    if (numBasesOverlap > largestOverlap) { bestMatch = g; largestOverlap = numBasesOverlap; }
  }
  return bestMatch;
}

// taxon.cu: getResultFromPairedOverlaps over the batch's per-read records, appended to the run's results
int taxa_add_batch(kslam_taxa *taxa, const kslam_taxdb *taxdb, const Ctx &c, const std::vector<ReadPair> &rp, uint32_t threads);

char *dup_text(const std::string &s);

}  // namespace kslam_host
