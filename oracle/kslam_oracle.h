/* oracle/kslam_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded-by-default CPU restatement of the reference's read-to-genome
 * matching path (k-SLAM, /root/reference/src). It exists ONLY as the checker for tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg. Nothing under k-slam_b200/
 * may include, link, load or call it; the product fails loudly without its CUDA library.
 *
 * Parity status: PINNED — every function here is checked (tests/test_oracle_vs_ref.py)
 * against the reference's own code compiled unmodified into oracle/_ref/libkslam_ref.so
 * (oracle/ref_driver.cpp), and against golden vectors generated from that build and
 * committed under tests/golden/ (generator: tests/golden/make_golden.py). The reference
 * itself ships no golden vectors or runnable tests for this path.
 *
 * Record layouts are byte-identical to include/kslam.h so buffers can be compared raw.
 */
#ifndef KSLAM_ORACLE_H_
#define KSLAM_ORACLE_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KO_K 32 /* Globals.h:25 */

/* KMerAndData, KMer.h:58-116.  id_flags: bits 0-29 id, bit 30 revComp, bit 31 isFromGB. */
typedef struct { uint64_t kmer; uint32_t id_flags; uint32_t offset; } ko_kmer;
/* OverlapTemp, Overlap.h:36-52 */
typedef struct { uint32_t read; uint32_t entry; int32_t rel; uint32_t rev_comp; } ko_seed;
/* Overlap + StripedSmithWaterman::Alignment, Overlap.h:53-74, ssw_cpp.h:10-18 */
typedef struct {
  uint32_t read, entry; int32_t rel; uint32_t rev_comp;
  int32_t ref_begin, ref_end, query_begin, query_end;
  uint32_t sw_score, cigar_off, cigar_len, flags;
} ko_overlap;
/* PairedOverlap, PairedOverlap.h:32-58; r1_idx/r2_idx index the pair-sorted overlap array, -1 = none */
typedef struct {
  uint32_t combined_score, entry; int32_t ref_start, ref_end;
  uint32_t insert_size; int32_t r1_idx, r2_idx; uint32_t pad;
} ko_pair;

/* ko_overlap.flags */
#define KO_FLAG_UNDEFINED 1u /* reference behaviour is undefined here (score 0 with cigar, or
                                traceback left the written band): result not comparable */
#define KO_FLAG_CIGAR_OVERFLOW 2u

typedef struct {
  int32_t match, mismatch, gap_open, gap_extend; /* Globals.h:27-30; narrowed to u8 at ssw_cpp.cpp:114-117 */
  uint32_t score_threshold;                      /* Globals.h:31 */
  int32_t report_cigar;                          /* Globals.h:36 */
} ko_params;

/* KMer.h:160-181,246-280: canonical 32-mers of each sequence, gap 1 (reads) or 16 (genomes).
 * Returns the number of records written to out (call with out==NULL to count). */
uint64_t ko_extract_kmers(uint64_t n, const char *bases, const uint64_t *offs, int is_gb,
                          uint32_t gap, ko_kmer *out);
/* KMer.h:388-398: kmer ascending, id_flags DESCENDING (offset ascending as a fixed tie-break). */
void ko_sort_kmers(ko_kmer *recs, uint64_t n);
/* Overlap.h:153-246: pile walk over a sorted combined list. read_lens[id] = length of read id.
 * out==NULL counts. */
uint64_t ko_find_seeds_raw(const ko_kmer *recs, uint64_t n, const uint32_t *read_lens, ko_seed *out);
/* Overlap.h:87-98,79-85,289-291: sort by (read, entry, rel[, rev_comp]) then unique vs last kept. */
uint64_t ko_sort_unique_seeds(ko_seed *seeds, uint64_t n);

/* ssw_cpp.cpp:234-283 + ssw.c:841-951 for one (query, ref) pair; SSW's own coordinates.
 * cigar (capacity cigar_cap u32) may be NULL when !report_cigar. */
void ko_ssw_align(const char *q, int32_t qlen, const char *r, int32_t rlen, const ko_params *p,
                  ko_overlap *out, uint32_t *cigar, uint32_t cigar_cap);
/* Cross-check: the same alignment through plain (un-striped) Gotoh with SSW's tie rules; equal to ko_ssw_align's score and
 * coordinates when gap_extend < gap_open and mismatch <= 2 * gap_extend. */
void ko_ssw_align_gotoh(const char *q, int32_t qlen, const char *r, int32_t rlen, const ko_params *p, ko_overlap *out);
/* Batched form used by tests and the CPU baseline; cigar pool has cigar_cap u32 per pair. */
void ko_ssw_batch(uint64_t n, const char *q, const uint64_t *qoffs, const char *r,
                  const uint64_t *roffs, const ko_params *p, ko_overlap *out, uint32_t *cigar_pool,
                  uint32_t cigar_cap, int threads);
/* SmithWaterman.h:184-233: window build, align, un-flip for each seed (in place on ov[]). */
void ko_align_seeds(uint64_t n_seeds, ko_overlap *ov, const char *read_bases, const uint64_t *read_offs,
                    const char *gen_bases, const uint64_t *gen_offs, const ko_params *p,
                    uint32_t *cigar_pool, uint32_t cigar_cap, int threads);
/* PairedOverlap.h:243-257: sort overlaps by (read % mid, entry, rel) in place. */
void ko_sort_for_pairing(ko_overlap *ov, uint64_t n, uint32_t mid);
/* PairedOverlap.h:107-242,258-272 on pair-sorted overlaps. out==NULL counts. */
uint64_t ko_pair_overlaps(const ko_overlap *ov, uint64_t n, uint32_t mid, const uint32_t *read_lens, ko_pair *out);

#ifdef __cplusplus
}
#endif
#endif
