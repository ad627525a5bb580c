// oracle/_ref harness — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// This translation unit is OUR code. It #includes the reference's headers where they
// lie under /root/reference/src (nothing is copied into this repo) and exposes the
// reference's own matching path through a small C interface so that tests and the
// CPU-baseline leg of bench.py can run the UNMODIFIED reference implementation:
//
//   getKMersFromReads            /root/reference/src/KMer.h:373-381
//   GenbankIndex::getKMers       /root/reference/src/GenbankTools.h:211-219
//   sortKMers                    /root/reference/src/KMer.h:388-398
//   findOverlaps[_parallel]      /root/reference/src/Overlap.h:230-246,277-295
//   alignToDatabase              /root/reference/src/SLAM.h:60-79
//   screenOverlapsByScoreThreshold /root/reference/src/Overlap.h:329-341
//   getPairedOverlaps            /root/reference/src/PairedOverlap.h:243-272
//   StripedSmithWaterman::Aligner::Align  /root/reference/src/ssw_cpp.cpp:234-283
//   host stages up to SAM / XML: PairedOverlap.h:314-576, SAM.h, MetagenomicResults.h, TaxonomyDatabase.h
//   createIndexFromGBFF / createIndexFromFASTA   /root/reference/src/GenbankTools.h:224-260,481-527
//
// Built by oracle/Makefile into oracle/_ref/libkslam_ref.so (git-ignored, travels to the
// GPU box as a prebuilt binary). Boost is absent from this image; oracle/ref_shim/
// holds six no-op stand-in headers so the reference headers compile.
// The reference headers rely on transitive includes (normally via Boost); pre-include them.
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

using namespace SLAM;

namespace {

typedef KMerAndData<KMerInt, k> KM;

struct KrefCtx {
  GenbankIndex idx;
  std::vector<MetagenomicFASTQSequence> reads;
  std::vector<KM> kmers;
  std::vector<OverlapTemp> seeds;
  std::vector<Overlap> overlaps;
  std::vector<PairedOverlap> pairs;
  std::vector<IdentifiedTaxonomy> taxa;   // grows batch by batch like SLAM.h:243-249
};
GenbankIndex g_parsed_index;              // what the reference's database builders handed to the (stub) archive
void capture_index(const void *obj) { g_parsed_index = *(const GenbankIndex *)obj; }

// Records shared with oracle/kslam_oracle.h and include/kslam.h (same layout).
struct KmerRec { uint64_t kmer; uint32_t id_flags; uint32_t offset; };
struct SeedRec { uint32_t read; uint32_t entry; int32_t rel; uint32_t rev_comp; };
struct OverlapRec {
  uint32_t read, entry; int32_t rel; uint32_t rev_comp;
  int32_t ref_begin, ref_end, query_begin, query_end;
  uint32_t sw_score, cigar_off, cigar_len, pad;
};
struct PairRec {
  uint32_t combined_score, entry; int32_t ref_start, ref_end;
  uint32_t insert_size; int32_t r1_idx, r2_idx; uint32_t pad;
};

void fill_overlap(OverlapRec &o, const Overlap &v, uint32_t cigar_off) {
  o.read = v.readPosInArray; o.entry = v.entryPosInArray; o.rel = v.relativePosition;
  o.rev_comp = v.revComp;
  o.ref_begin = v.alignment.ref_begin; o.ref_end = v.alignment.ref_end;
  o.query_begin = v.alignment.query_begin; o.query_end = v.alignment.query_end;
  o.sw_score = v.alignment.sw_score; o.cigar_off = cigar_off;
  o.cigar_len = v.alignment.cigar ? (uint32_t)v.alignment.cigarLen : 0; o.pad = 0;
}

}  // namespace

extern "C" {

void *kref_create() { forceParallel(); return new KrefCtx(); }
void kref_destroy(void *h) { delete (KrefCtx *)h; }
int kref_num_threads() { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1; the CPU baseline must be able to ask for every host core explicitly.
void kref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// Globals.h:27-42 — main.cpp:36-58 normally fills these.
void kref_set_params(uint32_t m, uint32_t x, uint32_t go, uint32_t ge, uint32_t thr,
                     int report_cigar) {
  match = m; misMatch = x; gapOpen = go; gapExtend = ge; scoreThreshold = thr;
  reportCigar = report_cigar != 0;
}

void kref_set_genomes(void *h, uint64_t n, const char *bases, const uint64_t *offs) {
  KrefCtx *c = (KrefCtx *)h;
  c->idx.entries.clear();
  c->idx.entries.resize(n);
  for (uint64_t i = 0; i < n; i++) {
    c->idx.entries[i].bases.assign(bases + offs[i], bases + offs[i + 1]);
    c->idx.entries[i].locusTag = "g" + std::to_string(i);
  }
}

void kref_set_reads(void *h, uint64_t n, const char *bases, const uint64_t *offs) {
  KrefCtx *c = (KrefCtx *)h;
  c->reads.clear();
  c->reads.reserve(n);
  for (uint64_t i = 0; i < n; i++) {
    std::string b(bases + offs[i], bases + offs[i + 1]);
    std::string q(b.size(), 'I');
    c->reads.emplace_back("@r" + std::to_string(i), b, q);
  }
}

// which: 1 = reads only, 2 = genomes only (gap k/2), 3 = both, in alignToDatabase order.
uint64_t kref_extract_kmers(void *h, int which) {
  KrefCtx *c = (KrefCtx *)h;
  c->kmers.clear();
  if (which & 1) getKMersFromReads(c->reads, c->kmers);
  if (which & 2) c->idx.getKMers<KMerInt, k>(c->kmers, k / 2);
  return c->kmers.size();
}
void kref_sort_kmers(void *h) { sortKMers(((KrefCtx *)h)->kmers); }
void kref_get_kmers(void *h, void *out) {
  KrefCtx *c = (KrefCtx *)h;
  KmerRec *o = (KmerRec *)out;
  for (size_t i = 0; i < c->kmers.size(); i++) {
    o[i].kmer = c->kmers[i].kMerInt;
    o[i].id_flags = c->kmers[i].kMerData.ID_isFromGB_RC;
    o[i].offset = c->kmers[i].kMerData.offset;
  }
}

// Raw pile walk over the whole sorted list as one chunk (Overlap.h:230-246).
uint64_t kref_find_seeds_raw(void *h) {
  KrefCtx *c = (KrefCtx *)h;
  // findOverlaps may peek at *last (Overlap.h:243); keep one spare element past the end.
  c->kmers.reserve(c->kmers.size() + 1);
  c->seeds = findOverlaps(c->kmers.begin(), c->kmers.end(), c->reads);
  return c->seeds.size();
}
// Chunked walk + sort + fuzzy unique (Overlap.h:277-295).
uint64_t kref_find_seeds(void *h) {
  KrefCtx *c = (KrefCtx *)h;
  c->kmers.reserve(c->kmers.size() + 1);
  c->seeds = findOverlaps_parallel(c->kmers.begin(), c->kmers.end(), c->reads);
  return c->seeds.size();
}
void kref_get_seeds(void *h, void *out) {
  KrefCtx *c = (KrefCtx *)h;
  SeedRec *o = (SeedRec *)out;
  for (size_t i = 0; i < c->seeds.size(); i++) {
    o[i].read = c->seeds[i].readPosInArray; o[i].entry = c->seeds[i].entryPosInArray;
    o[i].rel = c->seeds[i].relativePosition; o[i].rev_comp = c->seeds[i].revComp;
  }
}

uint64_t kref_align_to_database(void *h) {
  KrefCtx *c = (KrefCtx *)h;
  c->overlaps = alignToDatabase(c->reads, c->idx);
  return c->overlaps.size();
}
uint64_t kref_screen(void *h) {
  KrefCtx *c = (KrefCtx *)h;
  screenOverlapsByScoreThreshold(c->overlaps, scoreThreshold);
  return c->overlaps.size();
}
uint64_t kref_num_overlaps(void *h) { return ((KrefCtx *)h)->overlaps.size(); }
uint64_t kref_cigar_total(void *h) {
  KrefCtx *c = (KrefCtx *)h;
  uint64_t t = 0;
  for (auto &o : c->overlaps) if (o.alignment.cigar) t += o.alignment.cigarLen;
  return t;
}
void kref_get_overlaps(void *h, void *out, uint32_t *cigar_pool) {
  KrefCtx *c = (KrefCtx *)h;
  OverlapRec *o = (OverlapRec *)out;
  uint32_t off = 0;
  for (size_t i = 0; i < c->overlaps.size(); i++) {
    fill_overlap(o[i], c->overlaps[i], off);
    for (uint32_t j = 0; j < o[i].cigar_len; j++) cigar_pool[off + j] = c->overlaps[i].alignment.cigar[j];
    off += o[i].cigar_len;
  }
}

// getPairedOverlaps sorts c->overlaps in place; r1_idx/r2_idx index that sorted order.
uint64_t kref_pair(void *h) {
  KrefCtx *c = (KrefCtx *)h;
  c->pairs = getPairedOverlaps(c->overlaps.begin(), c->overlaps.end(), c->reads);
  return c->pairs.size();
}
void kref_get_pairs(void *h, void *out) {
  KrefCtx *c = (KrefCtx *)h;
  PairRec *o = (PairRec *)out;
  std::map<std::tuple<uint32_t, uint32_t, int32_t>, int32_t> where;
  for (size_t i = 0; i < c->overlaps.size(); i++)
    where[std::make_tuple(c->overlaps[i].readPosInArray, c->overlaps[i].entryPosInArray,
                          c->overlaps[i].relativePosition)] = (int32_t)i;
  for (size_t i = 0; i < c->pairs.size(); i++) {
    const PairedOverlap &p = c->pairs[i];
    o[i].combined_score = p.combinedScore; o[i].entry = p.entryPosInArray;
    o[i].ref_start = p.refStart; o[i].ref_end = p.refEnd; o[i].insert_size = p.insertSize;
    o[i].r1_idx = p.hasR1 ? where[std::make_tuple(p.r1Overlap.readPosInArray, p.r1Overlap.entryPosInArray,
                                                  p.r1Overlap.relativePosition)] : -1;
    o[i].r2_idx = p.hasR2 ? where[std::make_tuple(p.r2Overlap.readPosInArray, p.r2Overlap.entryPosInArray,
                                                  p.r2Overlap.relativePosition)] : -1;
    o[i].pad = 0;
  }
}

// Direct batched Aligner::Align (ssw_cpp.cpp:234-283) over n (query, ref) pairs, all host
// threads. Output coordinates are SSW's own (no un-flip, no refStart offset). cigar_pool
// has cigar_cap u32 per alignment. Returns number of alignments.
uint64_t kref_ssw_batch(uint64_t n, const char *q, const uint64_t *qoffs, const char *r,
                        const uint64_t *roffs, int report_cigar, uint32_t score_filter,
                        void *out, uint32_t *cigar_pool, uint32_t cigar_cap, int threads) {
  OverlapRec *o = (OverlapRec *)out;
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
  {
    const StripedSmithWaterman::Aligner aligner(match, misMatch, gapOpen, gapExtend);
    StripedSmithWaterman::Filter filter;
    filter.report_begin_position = true;
    filter.report_cigar = report_cigar != 0;
    filter.score_filter = score_filter;
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++) {
      std::string qs(q + qoffs[i], q + qoffs[i + 1]);
      std::string rs(r + roffs[i], r + roffs[i + 1]);
      Overlap ov;
      aligner.Align(qs.c_str(), rs.c_str(), (int)rs.size(), filter, &ov.alignment);
      fill_overlap(o[i], ov, (uint32_t)(i * cigar_cap));
      if (o[i].cigar_len > cigar_cap) o[i].cigar_len = cigar_cap | 0x80000000u;
      for (uint32_t j = 0; j < (o[i].cigar_len & 0x7fffffffu) && j < cigar_cap; j++)
        cigar_pool[i * (uint64_t)cigar_cap + j] = ov.alignment.cigar[j];
    }
  }
  return n;
}

// Replace the qualities of the reads set by kref_set_reads (same lengths as the bases).
void kref_set_read_quals(void *h, const char *quals, const uint64_t *offs) {
  KrefCtx *c = (KrefCtx *)h;
  for (size_t i = 0; i < c->reads.size(); i++) c->reads[i].quality.assign(quals + offs[i], quals + offs[i + 1]);
}
// The rest of the batch loop for a --sam-file run, SLAM.h:215-239, on the pairs kref_pair left in the ctx:
// getPerReadOverlaps, getMaxAllowedInsertSize, both screens, pseudoAssembly (optional), writeSAMOutputPairs.
// Returns the SAM text length; the text is copied into buf when it is large enough.
uint64_t kref_sam(void *h, uint32_t num_alignments, double fraction, int pseudo, int sam_xa, const char *tmp_path,
                  char *buf, uint64_t cap, uint32_t *max_insert) {
  KrefCtx *c = (KrefCtx *)h;
  numSAMAlignments = num_alignments; scoreFractionThreshold = fraction; SAMXA = sam_xa != 0; pairedData = true;
  auto rp = getPerReadOverlaps(c->pairs.begin(), c->pairs.end(), c->reads.size() / 2);
  c->pairs.clear();
  uint32_t maxInsertSize = getMaxAllowedInsertSize(rp);
  if (max_insert) *max_insert = maxInsertSize;
  screenPairedAlignmentsByInsertSize(rp, maxInsertSize, true);
  screenPairedAlignmentsByScore(rp, scoreFractionThreshold);
  if (pseudo) {
    pseudoAssembly(rp, c->reads, c->idx);
    screenPairedAlignmentsByScore(rp, scoreFractionThreshold);
  }
  {
    std::ofstream sam(tmp_path);
    for (auto &read : rp) writeSAMOutputPairs(sam, read, c->reads, c->idx);
  }
  std::ifstream in(tmp_path, std::ios::binary);
  std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if (buf && cap >= text.size()) memcpy(buf, text.data(), text.size());
  return text.size();
}
// single-end flavour of the same loop body (SLAM.h:223-228) on the overlaps kref_align_to_database + kref_screen left
uint64_t kref_sam_single(void *h, uint32_t num_alignments, double fraction, int pseudo, int sam_xa, const char *tmp_path,
                         char *buf, uint64_t cap) {
  KrefCtx *c = (KrefCtx *)h;
  numSAMAlignments = num_alignments; scoreFractionThreshold = fraction; SAMXA = sam_xa != 0; pairedData = false;
  auto perRead = getPerReadOverlaps(c->overlaps.begin(), c->overlaps.end());
  auto rp = getDummyAlignmentPairsFromSingleEndReads(perRead, c->reads);
  screenPairedAlignmentsByScore(rp, scoreFractionThreshold);
  if (pseudo) {
    pseudoAssembly(rp, c->reads, c->idx);
    screenPairedAlignmentsByScore(rp, scoreFractionThreshold);
  }
  {
    std::ofstream sam(tmp_path);
    for (auto &read : rp) writeSAMOutputPairs(sam, read, c->reads, c->idx);
  }
  pairedData = true;
  std::ifstream in(tmp_path, std::ios::binary);
  std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if (buf && cap >= text.size()) memcpy(buf, text.data(), text.size());
  return text.size();
}
uint64_t kref_sam_header(void *h, const char *cmd, char *buf, uint64_t cap) {
  KrefCtx *c = (KrefCtx *)h;
  commandLine = cmd;
  std::string t = getHeader(c->idx);
  if (buf && cap >= t.size()) memcpy(buf, t.data(), t.size());
  return t.size();
}

// ---- the reference's own FASTQ reader (FASTQsequence.h:110-165) over real files, batch by batch ------------------
// Streams stay open across calls like the ifstreams of the batch loop (SLAM.h:194-208). Returns the number of reads
// of the batch, or (uint64_t)-1 when the reference throws ("mismatch in R1 and R2 size").
struct KrefFastq { std::ifstream r1, r2; bool paired; std::vector<MetagenomicFASTQSequence> reads; };
void *kref_fastq_open(const char *r1, const char *r2) {
  KrefFastq *f = new KrefFastq();
  f->r1.open(r1); f->paired = r2 != nullptr;
  if (f->paired) f->r2.open(r2);
  return f;
}
uint64_t kref_fastq_next(void *h, unsigned max_reads) {
  KrefFastq *f = (KrefFastq *)h;
  f->reads.clear();
  try {
    if (f->paired) getPairedSequencesFromFASTQFiles(f->r1, f->r2, f->reads, max_reads);
    else getSequencesFromFASTQFile(f->r1, f->reads, max_reads);
  } catch (const std::runtime_error &) { return (uint64_t)-1; }
  return f->reads.size();
}
// which: 0 = bases, 1 = quality, 2 = sequenceIdentifier. offs gets n+1 entries; buf may be NULL to size it.
uint64_t kref_fastq_get(void *h, int which, char *buf, uint64_t *offs) {
  KrefFastq *f = (KrefFastq *)h;
  uint64_t pos = 0;
  for (size_t i = 0; i < f->reads.size(); i++) {
    const std::string &s = which == 0 ? f->reads[i].bases : (which == 1 ? f->reads[i].quality : f->reads[i].sequenceIdentifier);
    if (offs) offs[i] = pos;
    if (buf) memcpy(buf + pos, s.data(), s.size());
    pos += s.size();
  }
  if (offs) offs[f->reads.size()] = pos;
  return pos;
}
void kref_fastq_close(void *h) { delete (KrefFastq *)h; }

// ---- genes / taxonomy ids / read ids of the in-memory index, the taxonomy database and the metagenomic outputs -----------
void kref_set_read_ids(void *h, const char *ids, const uint64_t *offs) {
  KrefCtx *c = (KrefCtx *)h;
  for (size_t i = 0; i < c->reads.size(); i++) c->reads[i].sequenceIdentifier.assign(ids + offs[i], ids + offs[i + 1]);
}
void kref_set_entry_meta(void *h, uint64_t e, uint32_t taxonomy_id, const char *locus_tag) {
  KrefCtx *c = (KrefCtx *)h;
  c->idx.entries[e].taxonomyID = taxonomy_id;
  if (locus_tag) c->idx.entries[e].locusTag = locus_tag;
}
void kref_add_gene(void *h, uint64_t e, const char *name, const char *locus, const char *protein, const char *product,
                   const char *reference, uint32_t gene_id, uint32_t start, uint32_t stop) {
  KrefCtx *c = (KrefCtx *)h;
  Gene g(name, locus, protein, product, reference, CDS(start, stop, false));
  g.geneID = gene_id;
  c->idx.entries[e].genes.push_back(g);
}
void *kref_taxdb_open(const char *path) {
  try { return new TaxonomyDB(path); } catch (...) { return nullptr; }
}
void kref_taxdb_close(void *t) { delete (TaxonomyDB *)t; }
int kref_taxdb_build(const char *names, const char *nodes, const char *out) {       // --parse-taxonomy, main.cpp:131-141
  try { TaxonomyDB t; t.writeTaxonomyIndex(out, names, nodes); return 0; } catch (...) { return -1; }
}
uint64_t kref_taxdb_size(void *t) { return ((TaxonomyDB *)t)->taxIDsAndEntries.size(); }
uint32_t kref_lca(void *t, const uint32_t *ids, uint64_t n) {
  return ((TaxonomyDB *)t)->getLowestCommonAncestor(std::vector<uint32_t>(ids, ids + n));
}
uint64_t kref_lineage(void *t, uint32_t id, int which, char *buf, uint64_t cap) {
  std::string s = which == 0 ? ((TaxonomyDB *)t)->getLineage(id) : ((TaxonomyDB *)t)->getScientificName(id);
  if (buf && cap >= s.size()) memcpy(buf, s.data(), s.size());
  return s.size();
}
// One pass of the batch loop after pairing, SLAM.h:215-249, on what kref_pair (paired) or kref_align_to_database +
// kref_screen (single-end) left in the ctx: screens, pseudo-assembly, SAM records if wanted, then
// convertAlignmentsToIdentifiedTaxonomies_parallel appended to the run's results. Returns the SAM text length.
uint64_t kref_meta_batch(void *h, void *taxdb, uint32_t num_alignments, double fraction, int pseudo, int sam_xa, int paired,
                         int want_sam, const char *tmp_path, char *buf, uint64_t cap) {
  KrefCtx *c = (KrefCtx *)h;
  numSAMAlignments = num_alignments; scoreFractionThreshold = fraction; SAMXA = sam_xa != 0; pairedData = paired != 0;
  std::vector<ReadPairAndOverlaps> rp;
  if (paired) {
    rp = getPerReadOverlaps(c->pairs.begin(), c->pairs.end(), c->reads.size() / 2);
    c->pairs.clear();
    uint32_t maxInsertSize = getMaxAllowedInsertSize(rp);
    screenPairedAlignmentsByInsertSize(rp, maxInsertSize, true);
    screenPairedAlignmentsByScore(rp, scoreFractionThreshold);
  } else {
    auto perRead = getPerReadOverlaps(c->overlaps.begin(), c->overlaps.end());
    rp = getDummyAlignmentPairsFromSingleEndReads(perRead, c->reads);
    screenPairedAlignmentsByScore(rp, scoreFractionThreshold);
  }
  if (pseudo) {
    pseudoAssembly(rp, c->reads, c->idx);
    screenPairedAlignmentsByScore(rp, scoreFractionThreshold);
  }
  std::string text;
  if (want_sam) {
    {
      std::ofstream sam(tmp_path);
      for (auto &read : rp) writeSAMOutputPairs(sam, read, c->reads, c->idx);
    }
    std::ifstream in(tmp_path, std::ios::binary);
    text.assign((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (buf && cap >= text.size()) memcpy(buf, text.data(), text.size());
  }
  if (taxdb) {
    auto fresh = convertAlignmentsToIdentifiedTaxonomies_parallel(rp.begin(), rp.end(), c->reads, c->idx, *(TaxonomyDB *)taxdb);
    c->taxa.insert(c->taxa.end(), fresh.begin(), fresh.end());
  }
  pairedData = true;
  return text.size();
}
// End of the run, SLAM.h:256-265: <prefix>_PerRead, <prefix> (XML), <prefix>_abbreviated.
void kref_meta_finish(void *h, void *taxdb, uint32_t num_reads, const char *prefix) {
  KrefCtx *c = (KrefCtx *)h;
  TaxonomyDB &taxDB = *(TaxonomyDB *)taxdb;
  std::string outFileName(prefix);
  std::ofstream perReadout(outFileName + "_PerRead");
  writePerReadResults(c->taxa, perReadout);
  c->taxa = combineTaxonomies(c->taxa);
  std::ofstream outFile(outFileName);
  writeResults(c->taxa, outFile, taxDB, num_reads);
  writeAbbreviatedResultsFile(c->taxa, outFileName + "_abbreviated", taxDB, num_reads);
  c->taxa.clear();
}

// ---- the reference's database builders; the GenbankIndex they build is captured from the stub archive -------------------
// kind 0 = createIndexFromGBFF (needs a file "taxDB" in the CWD, GenbankTools.h:483), 1 = createIndexFromFASTA.
// Returns the number of entries or -1 when the reference throws.
int64_t kref_parse_index(int kind, const char *const *paths, uint64_t n, const char *out_path) {
  std::vector<std::string> names(paths, paths + n);
  g_parsed_index = GenbankIndex();
  kref_shim::archive_hook() = capture_index;
  int64_t rc;
  try {
    if (kind == 0) createIndexFromGBFF(names, out_path); else createIndexFromFASTA(names, out_path);
    rc = (int64_t)g_parsed_index.entries.size();
  } catch (...) { rc = -1; }
  kref_shim::archive_hook() = nullptr;
  return rc;
}
// Text dump of the captured index, fields separated by 0x1f, records by 0x1e:
//   E locusTag taxonomyID genbankID isPlasmid is16S bases | G geneName locusTag proteinID product referenceSequence geneID start stop complement
uint64_t kref_parsed_index_dump(char *buf, uint64_t cap) {
  std::string t;
  const char F = 0x1f, R = 0x1e;
  for (auto &e : g_parsed_index.entries) {
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
}
// Use the captured index as the ctx's database (so alignment, SAM gene tags and taxonomy run on what the reference parsed).
void kref_use_parsed_index(void *h) { ((KrefCtx *)h)->idx = g_parsed_index; }

}  // extern "C"
