/* include/kslam.h — C ABI of the B200-native k-SLAM matching path (libkslam.so).
 *
 * The reference (aindj/k-SLAM) has no plugin/FFI interface for this path: the batch loop calls two
 * C++ templates directly. This header is therefore the boundary a maintainer would
 * bind INSTEAD of those two calls (see INTEGRATION.md for the stub):
 *
 *   kslam_load_genomes   replaces GenbankIndex::getKMers + the per-batch re-sort of genome k-mers
 *                        /root/reference/src/GenbankTools.h:211-219, /root/reference/src/SLAM.h:65-66
 *   kslam_align_batch    replaces alignToDatabase            /root/reference/src/SLAM.h:60-79
 *   kslam_pair_batch     replaces screenOverlapsByScoreThreshold + getPairedOverlaps
 *                        /root/reference/src/Overlap.h:329-341, /root/reference/src/PairedOverlap.h:243-272
 *   kslam_ssw_batch      replaces StripedSmithWaterman::Aligner::Align, batched
 *                        /root/reference/src/ssw_cpp.cpp:234-283 (C ABI below it: ssw.h:97-192)
 *
 * Plain pointers and sizes only; no exceptions cross the boundary (0 = ok, negative = error, text
 * from kslam_last_error). Sequences are passed as ONE concatenated byte array plus n+1 offsets
 * (sequence i = bases[offs[i] .. offs[i+1])), exactly the bytes of std::string `bases` in the
 * reference's FASTQSequence / GenbankEntry. (`bases` is copied with cudaMemcpyDefault: a pointer into device or managed
 * memory of the same process works too — a 20 Gbp database that was produced on the GPU need not visit the host.) Result buffers are owned by the ctx and stay valid until
 * the next call of the same function or kslam_destroy. One in-flight batch per ctx; use one ctx per
 * GPU (and two per GPU to double-buffer). There is NO CPU fallback: every entry point that computes
 * fails with KSLAM_ERR_CUDA if no sm_100 device is usable.
 *
 * All arithmetic on this path is integer; results are bit-exact with the reference for every scoring
 * parameter set (kslam_params_fast() tells which kernels run).
 */
#ifndef KSLAM_H_
#define KSLAM_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KSLAM_K 32u /* Globals.h:25 */

#define KSLAM_OK 0
#define KSLAM_ERR_ARG (-1)
#define KSLAM_ERR_CUDA (-2)
#define KSLAM_ERR_NOMEM (-3)
#define KSLAM_ERR_STATE (-4)

typedef struct kslam_ctx kslam_ctx;

/* Globals.h:27-42 as filled by main.cpp:36-58; narrowed to u8 as at ssw_cpp.cpp:114-117. */
typedef struct {
  uint8_t match, mismatch, gap_open, gap_extend; /* defaults 2 3 5 2 */
  uint16_t score_threshold;                      /* --min-alignment-score, default 0 */
  uint8_t report_cigar;                          /* reportCigar: true iff --sam-file (SLAM.h:169) */
  uint8_t reserved0;
  int32_t device;                                /* CUDA device ordinal */
  uint32_t genome_gap;                           /* k/2 = 16 (SLAM.h:65); 0 = default */
  uint32_t max_cigar_ops;                        /* per-alignment cigar capacity, 0 = default 32 */
  uint32_t stream_priority;                      /* 0 = default; 1 = the ctx stream gets the highest CUDA stream priority: of two
                                                    contexts double-buffering on one GPU, give it to one so that its kernels run
                                                    first and the other fills the gaps its PCIe copies leave */
} kslam_params;

/* KMerAndData, KMer.h:58-116. id_flags: bits 0-29 id, bit 30 revComp, bit 31 isFromGB. */
typedef struct { uint64_t kmer; uint32_t id_flags; uint32_t offset; } kslam_kmer;
/* OverlapTemp, Overlap.h:36-52. */
typedef struct { uint32_t read; uint32_t entry; int32_t rel; uint32_t rev_comp; } kslam_seed;
/* Overlap (Overlap.h:53-74) with its StripedSmithWaterman::Alignment (ssw_cpp.h:10-18). The cigar of
 * overlap i is cigar_pool[cigar_off .. cigar_off+cigar_len), BAM encoding len<<4|op (M0 I1 D2). */
typedef struct {
  uint32_t read, entry; int32_t rel; uint32_t rev_comp;
  int32_t ref_begin, ref_end, query_begin, query_end;
  uint32_t sw_score, cigar_off, cigar_len, flags;
} kslam_overlap;
#define KSLAM_FLAG_UNDEFINED 1u      /* the reference's behaviour is undefined for this input (score 0 with
                                        cigar requested, or its traceback leaves the written band) */
#define KSLAM_FLAG_CIGAR_OVERFLOW 2u /* more than max_cigar_ops ops; cigar truncated */
/* PairedOverlap, PairedOverlap.h:32-58. r1_idx / r2_idx index `sorted_overlaps`; -1 = absent mate. */
typedef struct {
  uint32_t combined_score, entry; int32_t ref_start, ref_end;
  uint32_t insert_size; int32_t r1_idx, r2_idx; uint32_t pad;
} kslam_pair;

typedef struct {
  uint64_t n_overlaps;            /* == alignToDatabase(...).size(), same order */
  const kslam_overlap *overlaps;  /* host, pinned, ctx-owned */
  uint64_t n_cigar_words;
  const uint32_t *cigar_pool;     /* host, pinned, ctx-owned */
} kslam_alignments;

typedef struct {
  uint64_t n_sorted;                    /* overlaps surviving the score screen */
  const kslam_overlap *sorted_overlaps; /* in getPairedOverlaps' sorted order */
  uint64_t n_cigar_words;
  const uint32_t *cigar_pool;
  uint64_t n_pairs;
  const kslam_pair *pairs;              /* == getPairedOverlaps(...) order */
} kslam_pairs;

/* Device time of the stages of the last kslam_align_batch / kslam_pair_batch (CUDA events on the
 * ctx stream) and the unit counts the roofline figures are computed from. */
typedef struct {
  float ms_h2d, ms_pack, ms_extract, ms_sort, ms_join, ms_seed_sort, ms_unique;
  float ms_sw_prepare, ms_sw_forward, ms_sw_reverse, ms_sw_traceback, ms_sw_slow, ms_d2h, ms_pair, ms_total;
  uint64_t n_read_kmers, n_sorted_kmers, n_genome_kmers, n_raw_seeds, n_seeds, n_sort_passes;
  uint64_t sw_cells_forward, sw_cells_reverse, sw_cells_computed, n_sw_fast, n_sw_slow, n_sw_band, n_sw_band64, n_sw_band_rev, n_traceback_dp, n_pairs;
  uint64_t n_sw_tier8, n_sw_tier16, n_sw_tier32, n_sw_tier48, n_sw_tier64, n_sw_sweep32; /* forward work-list tiers */
  uint64_t n_sw_tier96, n_sw_tier128;  /* forward tiers swept by three / four lanes of 32 diagonals */
  uint64_t n_sw_fwd_tier[12];          /* alignments per forward band tier: 8 16 24 32 40 48 56 64 72 80 96 128 diagonals */
  uint64_t n_sw_rev_tier[12];          /* ... and per reverse band tier */
  uint64_t n_sw_rev_diagonal;          /* reverse sweeps replaced by a diagonal score (bands one diagonal wide) */
  uint64_t sw_alu_ops;                 /* ALU-pipe thread-ops the computed cells need (3 per cell, 2.5 where the sweep runs without tracking) */
  uint64_t kernel_launches;
} kslam_timings;

int kslam_create(const kslam_params *params, kslam_ctx **out);
void kslam_destroy(kslam_ctx *ctx);
const char *kslam_last_error(const kslam_ctx *ctx); /* ctx may be NULL: last create error */
/* Results are bit-exact with the reference for EVERY scoring parameter set (any u8 match / mismatch / gap_open /
 * gap_extend, main.cpp:45-52): kslam_params_exact returns 1 for any non-NULL params and is kept for callers of the first
 * release. kslam_params_fast says which kernels do the work: 1 = the packed band / wavefront kernels (SSW's striped
 * kernels equal plain Gotoh there: gap_extend < gap_open and mismatch <= 2 * gap_extend; the defaults 2/3/5/2 are inside),
 * 0 = the literal lane-for-lane restatement of SSW's striped byte / word kernels (ssw.c:143-592), slower. */
int kslam_params_exact(const kslam_params *params);
/* The offsets [lo, hi] (column - row in the reversed matrix) that the reverse pass of an alignment scoring `score` sweeps
 * when its reversed prefixes are rows x cols (ssw.c:905-923: every alignment reaching the forward score starts at the
 * reversed origin). Pure arithmetic on the host, for parameters inside kslam_params_fast; KSLAM_ERR_ARG
 * otherwise. */
int kslam_reverse_band(int32_t rows, int32_t cols, int32_t score, const kslam_params *params, int32_t *lo, int32_t *hi);
int kslam_params_fast(const kslam_params *params);
const char *kslam_version(void);
/* HBM of a device (cudaMemGetInfo): a host that must choose between a replicated and a partitioned index asks here. */
int kslam_device_memory(int32_t device, uint64_t *free_bytes, uint64_t *total_bytes);

/* Pack the genomes, extract every genome_gap-th canonical 32-mer (KMer.h:160-181), sort once
 * (KMer.h:388-398) and keep everything resident in HBM. */
int kslam_load_genomes(kslam_ctx *ctx, uint64_t n_entries, const char *bases, const uint64_t *offs);

/* alignToDatabase (SLAM.h:60-79) for one batch: reads[0..n_reads), paired data laid out R1 block then
 * R2 block (FASTQsequence.h:110-123). */
int kslam_align_batch(kslam_ctx *ctx, uint64_t n_reads, const char *bases, const uint64_t *offs,
                      kslam_alignments *out);
/* Same work, split for measurement with HBM-resident inputs: upload once, run many times. */
int kslam_upload_reads(kslam_ctx *ctx, uint64_t n_reads, const char *bases, const uint64_t *offs);
int kslam_align_resident(kslam_ctx *ctx, int fetch_results, kslam_alignments *out /* may be NULL */);

/* screenOverlapsByScoreThreshold (Overlap.h:329-341) + getPairedOverlaps (PairedOverlap.h:243-272) on
 * the alignments of the last batch (device-resident). */
int kslam_pair_batch(kslam_ctx *ctx, int fetch_results, kslam_pairs *out /* may be NULL */);

/* Copy the results of the last kslam_pair_batch(fetch_results = 0) to the host. Splitting a batch into kslam_upload_reads
 * (H2D), kslam_align_resident + kslam_pair_batch (kernels) and kslam_fetch_pairs (D2H) lets a caller that double-buffers
 * with two contexts hand the GPU from one to the other with a host mutex around the kernel phase, so that the copies
 * of one batch always run under the kernels of the other (bench.py's e2e leg does exactly that). */
int kslam_fetch_pairs(kslam_ctx *ctx, kslam_pairs *out);

/* ---- compact results for runs WITHOUT --sam-file -------------------------------------------------------
 * The host stages of an XML-only run (insert-size screen, score screens, pseudo-assembly, per-read LCA and genes; SLAM.h:215-249)
 * read per pair record only: the read pair it belongs to, entry, reference span, insert size, score and which mates it has.
 * kslam_fetch_pairs_compact ships exactly that — 24 bytes per pair instead of 32 + the 48-byte alignment records — after
 * kslam_pair_batch(fetch_results = 0). The one place the mates themselves are looked at is the insert-size screen, which
 * splits a pair beyond the batch's limit into its two single-ended records (PairedOverlap.h:396-436): the call computes the
 * limit on the host (getMaxAllowedInsertSize, PairedOverlap.h:314-360, from the insert sizes it just received), then fetches
 * the mates of the pairs beyond it, in pair order. kslam_batch_outputs_compact continues from there: same XML / _PerRead /
 * _abbreviated as kslam_batch_outputs(want_sam = 0) on the full records. */
typedef struct {
  uint32_t pair_id;                /* index of the read pair (R1's read index) */
  uint32_t entry; int32_t ref_start, ref_end; uint32_t insert_size;
  uint32_t score_flags;            /* bits 0-29 combinedScore, bit 30 hasR1, bit 31 hasR2 */
} kslam_pair_compact;
typedef struct {
  uint32_t pair_index;             /* index into kslam_pairs_compact.pairs */
  uint32_t score1; int32_t ref_begin1, ref_end1;   /* the R1 alignment: sw_score, ref_begin, ref_end */
  uint32_t score2; int32_t ref_begin2, ref_end2;   /* the R2 alignment */
  uint32_t pad;
} kslam_far_mates;
typedef struct {
  uint64_t n_pairs; const kslam_pair_compact *pairs;       /* getPairedOverlaps order; host, pinned, ctx-owned */
  uint32_t insert_size_limit;                              /* getMaxAllowedInsertSize of this batch */
  uint64_t n_far; const kslam_far_mates *far;              /* mates of the pairs with insert_size > insert_size_limit, ascending pair_index */
} kslam_pairs_compact;
int kslam_fetch_pairs_compact(kslam_ctx *ctx, uint32_t host_threads /* 0 = all cores */, kslam_pairs_compact *out);
uint32_t kslam_insert_size_limit_compact(const kslam_pair_compact *pairs, uint64_t n, uint32_t host_threads);
/* The same statistic from value counts (counts[v] = pair records with insert size v, v = 0 .. top; counts[0] is ignored). */
uint32_t kslam_insert_size_limit_counts(const uint64_t *counts, uint32_t top);
/* The far-mates table for a limit given by the caller: a batch sharded over several contexts has ONE limit (a statistic of the
 * whole batch, PairedOverlap.h:314-360), computed from the merged compact records. */
int kslam_fetch_far_mates(kslam_ctx *ctx, uint32_t insert_size_limit, uint64_t *n_far, const kslam_far_mates **far);

/* The body of the reference's batch loop in one call (SLAM.h:209-214: alignToDatabase, score screen, getPairedOverlaps):
 * same results as kslam_align_batch + kslam_pair_batch, but the unsorted alignment vector — which the loop discards
 * once it is sorted — is never copied to the host. */
int kslam_align_pair_batch(kslam_ctx *ctx, uint64_t n_reads, const char *bases, const uint64_t *offs, kslam_pairs *out);

/* Aligner::Align (ssw_cpp.cpp:234-283) for n independent (query, ref) pairs, SSW's own coordinates
 * (no window un-flip). out[n] and cigar_pool[n * max_cigar_ops] are caller buffers (host). */
int kslam_ssw_batch(kslam_ctx *ctx, uint64_t n, const char *q, const uint64_t *qoffs, const char *r,
                    const uint64_t *roffs, kslam_overlap *out, uint32_t *cigar_pool);
int kslam_ssw_upload(kslam_ctx *ctx, uint64_t n, const char *q, const uint64_t *qoffs, const char *r,
                     const uint64_t *roffs);
int kslam_ssw_resident(kslam_ctx *ctx, kslam_overlap *out /* may be NULL */, uint32_t *cigar_pool /* may be NULL */);

/* ---- k-mer-range partitioned database (BASELINE config 4: the genome k-mer list exceeds one GPU) ----
 * Rank `part` of `n_parts` keeps the packed genomes and the prefilter bitmap (replicated) plus the slice of the sorted
 * genome k-mer list (KMer.h:388-398) whose kMerInt lies in [splitters[part], splitters[part+1]); splitters are the
 * quantiles of a sorted sample of the genome k-mers, a pure function of the database, so every rank derives the same
 * ones. Equal k-mers share an owner, so no pile (Overlap.h:153-199) is split — the invariant of Overlap.h:285-287.
 * One batch then takes two exchanges of 16-byte records, carried by the caller (NCCL all-to-all in
 * k-slam_b200/dist.py); the pointers handed out below are DEVICE pointers into ctx-owned buffers:
 *   1. kslam_upload_reads, kslam_part_route_kmers: read k-mer records (job-global read id = read_id_base + index,
 *      all ids < 2^30) grouped by key owner; counts[p] records go to rank p.
 *   2. receiver: kslam_part_recv_buffer(total) -> fill it -> kslam_part_join: sort, merge-join against the local
 *      range, raw matches {read record, genome record} grouped by read owner (id_bases[n_parts+1] = first global
 *      read id of every rank); counts[p] matches go back to rank p.
 *   3. read owner: kslam_part_match_buffer(total) -> fill it -> kslam_part_finish: match -> seed (Overlap.h:185-193
 *      needs the read length), seed sort + fuzzy unique, Smith-Waterman. Results as kslam_align_batch, bit-identical
 *      to the unpartitioned path on the same reads; kslam_pair_batch follows as usual. */
int kslam_load_genomes_part(kslam_ctx *ctx, uint64_t n_entries, const char *bases, const uint64_t *offs,
                            uint32_t part, uint32_t n_parts /* 1..64 */);
int kslam_get_partition(const kslam_ctx *ctx, uint32_t *part, uint32_t *n_parts, uint64_t *splitters /* n_parts+1, may be NULL */,
                        uint64_t *n_genome_kmers_total /* may be NULL */);
int kslam_part_route_kmers(kslam_ctx *ctx, uint32_t read_id_base, const void **dev_records, uint64_t *counts /* n_parts */);
int kslam_part_recv_buffer(kslam_ctx *ctx, uint64_t n_records, void **dev_ptr);
int kslam_part_join(kslam_ctx *ctx, uint64_t n_records, const uint32_t *id_bases /* n_parts+1 */,
                    const void **dev_matches, uint64_t *counts /* n_parts */);
int kslam_part_match_buffer(kslam_ctx *ctx, uint64_t n_matches, void **dev_ptr);
int kslam_part_finish(kslam_ctx *ctx, uint64_t n_matches, uint32_t read_id_base, int fetch_results,
                      kslam_alignments *out /* may be NULL */);

/* ---- the partitioned path driven by the library itself over NCCL (csrc/comm.cu) ------------------------------------------
 * kslam_comm = one rank of a job whose genome k-mer list is range-partitioned over its GPUs: the ctx holds this rank's key
 * range (kslam_load_genomes_part(part = rank, n_parts = ranks)); kslam_comm_align_resident is alignToDatabase (SLAM.h:60-79)
 * for the reads this rank uploaded (kslam_upload_reads) — the three stages above with the two all-to-alls in between issued
 * as ncclSend / ncclRecv groups on the ctx's stream (NVLink / NVSwitch). COLLECTIVE: every rank calls it once per batch.
 * Results as kslam_align_resident, bit-identical to the unpartitioned path on the same reads; kslam_pair_batch follows.
 * NCCL is opened at run time (dlopen libnccl.so.2); KSLAM_ERR_STATE when it is absent.
 *   one process, several GPUs:  kslam_comm_init_all(n, ctxs, comms)   one ctx per DISTINCT device, one host thread per rank
 *   one process per GPU:        rank 0: kslam_comm_unique_id(id) -> the launcher hands id to every rank ->
 *                               kslam_comm_init_rank(ctx, rank, n_ranks, id, &comm) */
typedef struct kslam_comm kslam_comm;
typedef struct {
  float ms_route, ms_exchange_kmers, ms_sort, ms_join, ms_exchange_matches, ms_finish;   /* device time of the stages of the last batch */
  uint64_t kmers_sent, kmers_received, matches_sent, matches_received;                    /* 16-byte records, this rank */
  uint64_t bytes_sent_kmers, bytes_sent_matches;                                          /* bytes that left this GPU (self excluded) */
  float ms_bucket_kmers, ms_bucket_matches;   /* the part of ms_route / ms_join spent grouping records by destination (the work the partition adds) */
} kslam_comm_stats;
int kslam_comm_unique_id(void *id128 /* 128 bytes */);
int kslam_comm_init_rank(kslam_ctx *ctx, uint32_t rank, uint32_t n_ranks, const void *id128, kslam_comm **out);
int kslam_comm_init_all(uint32_t n, kslam_ctx *const *ctxs, kslam_comm **out /* n entries */);
void kslam_comm_destroy(kslam_comm *comm);
int kslam_comm_rank(const kslam_comm *comm, uint32_t *rank, uint32_t *n_ranks);
int kslam_comm_align_resident(kslam_comm *comm, int fetch_results, kslam_alignments *out /* may be NULL */);
/* A rank that cannot take part in a batch (its kslam_upload_reads failed, say) calls this INSTEAD of
 * kslam_comm_align_resident: the other ranks' calls then return KSLAM_ERR_STATE instead of waiting for it. The same
 * holds inside kslam_comm_align_resident: a stage that fails on one rank ends the batch with an error on every rank. */
int kslam_comm_abort_batch(kslam_comm *comm);
int kslam_comm_get_stats(const kslam_comm *comm, kslam_comm_stats *out);

/* ---- FASTQ ingest (host side) ------------------------------------------------------------------
 * Chunk-parallel restatement of the reference's reader: getSequencesFromFASTQFile / getPairedSequencesFromFASTQFiles
 * (FASTQsequence.h:110-165) over safeGetline (sequenceTools.h:45-73: "\n", "\r\n" and lone "\r" end a line) with the
 * read-id rule of FASTQsequence.h:61-71 (drop '@', cut at the first space, then at the first '/'). Records are four
 * lines, never validated, exactly like the reference. One call returns the next max_reads records of R1 followed by
 * the next max_reads records of R2 (R1 block then R2 block, the layout kslam_align_batch expects;
 * --num-reads-at-once, SLAM.h:194-208); n_reads == 0 at end of input. Buffers are owned by the reader, valid until the
 * next call, page-locked when a CUDA device is present. KSLAM_ERR_STATE = the reference's "mismatch in R1 and R2 size". */
typedef struct kslam_fastq kslam_fastq;
typedef struct {
  uint64_t n_reads, n_r1;          /* reads in this batch; how many of them came from R1 */
  const char *bases; const uint64_t *offs;         /* read i = bases[offs[i] .. offs[i+1]) */
  const char *quals; const uint64_t *qual_offs;    /* quality line i, stored as is */
  const char *ids; const uint64_t *id_offs;        /* sequenceIdentifier i */
} kslam_read_batch;
int kslam_fastq_open(const char *r1_path, const char *r2_path /* NULL: single-end */, uint32_t threads /* 0: all cores */,
                     kslam_fastq **out);
/* Before the first batch: keep n buffer sets (1..16) and fill them round-robin, so a batch stays valid until n more
 * batches have been read — a pipelined caller (ingest | GPU | SAM) needs no copies. Default 1. */
int kslam_fastq_set_ring(kslam_fastq *reader, uint32_t n_buffer_sets);
int kslam_fastq_next(kslam_fastq *reader, uint64_t max_reads, kslam_read_batch *out);
const char *kslam_fastq_error(const kslam_fastq *reader);
void kslam_fastq_close(kslam_fastq *reader);

/* ---- host stages after pairing, up to the SAM text (host code, no GPU involved) -----------
 * kslam_sam_batch restates, literally and with the same std::sort calls so that ties fall as in the reference:
 * getPerReadOverlaps, getMaxAllowedInsertSize, screenPairedAlignmentsByInsertSize(replace), screenPairedAlignmentsByScore,
 * pseudoAssembly (+ second score screen) and writeSAMOutputPairs (PairedOverlap.h:314-576, SAM.h:101-517) — the rest
 * of the reference's batch loop for a --sam-file run (SLAM.h:215-239) — on the output of kslam_pair_batch (paired
 * data) or kslam_align_batch (kslam_sam_batch_single). With a gene table in kslam_sam_db the records carry XG / XP / XR of the gene
 * with the largest overlap (GenbankEntry::getGene, GenbankTools.h:170-185; SAM.h:361-370). */
typedef struct {
  uint32_t num_alignments;          /* --num-alignments (numSAMAlignments), default 10 */
  uint8_t pseudo_assembly;          /* 1 unless --no-pseudo-assembly (Globals.h:36) */
  uint8_t report_cigar;             /* reportCigar: cigar + MD columns */
  uint8_t sam_xa;                   /* --sam-xa: primary line only */
  uint8_t threads;                  /* host threads (stages are independent per read pair / per entry); 0 = all cores */
  double score_fraction_threshold;  /* --score-fraction-threshold, default 0.95 */
} kslam_sam_params;
/* Gene (GenbankTools.h:67-110): string i of {geneName, locusTag, proteinID, product, referenceSequence} =
 * gene_strings[str_offs[i] .. str_offs[i+1]). */
typedef struct kslam_gene {
  uint32_t cds_start, cds_stop;     /* CDS::start / stop as parsed (GenbankTools.h:389-413) */
  uint32_t gene_id, complement;
  uint64_t str_offs[6];
} kslam_gene;
typedef struct {
  uint64_t n_entries;
  const char *bases; const uint64_t *offs;             /* GenbankEntry::bases, as given to kslam_load_genomes */
  const char *locus_tags; const uint64_t *locus_offs;  /* GenbankEntry::locusTag of entry e = locus_tags[locus_offs[e] .. locus_offs[e+1]) */
  const uint32_t *taxonomy_ids;                        /* GenbankEntry::taxonomyID, may be NULL (all 0) */
  /* GenbankEntry::genes of GenBank databases (GenbankTools.h:67-110,149); all three NULL for FASTA databases.
   * Entry e owns genes[gene_offs[e] .. gene_offs[e+1]), in the entry's stored order. */
  const kslam_gene *genes; const uint64_t *gene_offs; const char *gene_strings;
  /* Optional (NULL = scan every gene of the entry, as GenbankEntry::getGene does): lookup structure over the gene table
   * from kslam_gene_index_build, same answers. kslam_index_db fills it in. */
  const struct kslam_gene_index *gene_index;
} kslam_sam_db;
/* Per entry whose genes are ordered by CDS start (what the database builders produce): the running maximum of the CDS
 * stops, so that the best gene of an alignment is found from a binary search and a short backward scan instead of a pass
 * over the entry's whole gene list (thousands of genes per genome, once per alignment). Entries in any other order keep the
 * full scan. The db's gene arrays must stay valid and unchanged while the index is in use. */
typedef struct kslam_gene_index kslam_gene_index;
int kslam_gene_index_build(const kslam_sam_db *db, kslam_gene_index **out);
void kslam_gene_index_free(kslam_gene_index *index);
int kslam_sam_header(const kslam_sam_db *db, const char *command_line, char **text, uint64_t *len);   /* getHeader, SAM.h:518-531 */
int kslam_sam_batch(const kslam_sam_params *params, const kslam_sam_db *db, const kslam_read_batch *reads,
                    const kslam_pairs *pairs, char **text, uint64_t *len, uint32_t *max_insert_size /* may be NULL */);
/* Single-end reads (SLAM.h:223-228): takes kslam_align_batch's output directly; score screen, per-read grouping, one
 * R1-only record per alignment, score-fraction screen, pseudo-assembly, SAM lines without mate fields. */
int kslam_sam_batch_single(const kslam_sam_params *params, const kslam_sam_db *db, const kslam_read_batch *reads,
                           const kslam_alignments *alignments, uint32_t score_threshold, char **text, uint64_t *len);
void kslam_sam_free(char *text);

/* ---- taxonomy and the metagenomic outputs (host side; the part of the batch loop after the SAM records, SLAM.h:243-265) --
 * kslam_taxdb  = TaxonomyDB (TaxonomyDatabase.h:45-348): the `taxDB` file of a --db directory (four lines per node: id,
 *                parent id, scientific name, rank — readTaxonomyIndex :166-183) or NCBI names.dmp / nodes.dmp
 *                (--parse-taxonomy, writeTaxonomyIndex :153-164).
 * kslam_taxa   = the run's std::vector<IdentifiedTaxonomy> (MetagenomicResults.h:32-42), grown batch by batch.
 * kslam_batch_outputs = kslam_sam_batch (SAM text optional) followed by convertAlignmentsToIdentifiedTaxonomies_parallel
 *                (MetagenomicResults.h:88-111,182-197) on the same per-read records: per read pair the LCA
 *                (getLowestCommonAncestor, TaxonomyDatabase.h:185-223) of the entries it still aligns to, plus the best
 *                gene of every alignment. db (bases, gene table) must stay valid until kslam_taxa_results.
 * kslam_taxa_results  = writePerReadResults, combineTaxonomies, writeResults (XML) and writeAbbreviatedResultsFile
 *                (MetagenomicResults.h:149-176,213-275,302-369,455-463): the texts of <out>_PerRead, <out> and
 *                <out>_abbreviated. combineTaxonomies orders reads by taxon with the sequential std::sort, i.e. the
 *                reference run with one OpenMP thread (its __gnu_parallel::sort is thread-count dependent on ties). */
typedef struct kslam_taxdb kslam_taxdb;
typedef struct kslam_taxa kslam_taxa;
int kslam_taxdb_open(const char *taxdb_path, kslam_taxdb **out);
int kslam_taxdb_build(const char *names_dmp, const char *nodes_dmp, const char *out_path);       /* --parse-taxonomy */
uint64_t kslam_taxdb_size(const kslam_taxdb *db);
uint32_t kslam_taxdb_lca(const kslam_taxdb *db, const uint32_t *tax_ids, uint64_t n);
int kslam_taxdb_lineage(const kslam_taxdb *db, uint32_t tax_id, char **text, uint64_t *len);     /* getLineage :249-265; free with kslam_sam_free */
int kslam_taxdb_name(const kslam_taxdb *db, uint32_t tax_id, char **text, uint64_t *len);        /* getScientificName :233-239 */
void kslam_taxdb_close(kslam_taxdb *db);
int kslam_taxa_create(kslam_taxa **out);
void kslam_taxa_destroy(kslam_taxa *taxa);
int kslam_batch_outputs(const kslam_sam_params *params, const kslam_sam_db *db, const kslam_read_batch *reads,
                        const kslam_pairs *pairs, int want_sam, char **sam_text /* may be NULL when !want_sam */, uint64_t *sam_len,
                        uint32_t *max_insert_size /* may be NULL */, const kslam_taxdb *taxdb, kslam_taxa *taxa);
int kslam_batch_outputs_single(const kslam_sam_params *params, const kslam_sam_db *db, const kslam_read_batch *reads,
                               const kslam_alignments *alignments, uint32_t score_threshold, int want_sam, char **sam_text,
                               uint64_t *sam_len, const kslam_taxdb *taxdb, kslam_taxa *taxa);
/* kslam_batch_outputs(want_sam = 0) on the compact records of kslam_fetch_pairs_compact */
int kslam_batch_outputs_compact(const kslam_sam_params *params, const kslam_sam_db *db, const kslam_read_batch *reads,
                                const kslam_pairs_compact *pairs, uint32_t *max_insert_size /* may be NULL */,
                                const kslam_taxdb *taxdb, kslam_taxa *taxa);
int kslam_taxa_results(kslam_taxa *taxa, const kslam_taxdb *taxdb, uint32_t num_reads, char **per_read, uint64_t *per_read_len,
                       char **xml, uint64_t *xml_len, char **abbreviated, uint64_t *abbreviated_len);

/* ---- database builders (host side): --parse-genbank / --parse-fasta / DIR/database ----------------------------------
 * kslam_index = GenbankIndex (GenbankTools.h:189-207). kslam_index_parse_genbank = createIndexFromGBFF (:481-527) with
 * parseSection (:348-476); kslam_index_parse_fasta = createIndexFromFASTA (:224-260); kslam_index_read / _write = the Boost
 * text archive of getIndexFromBoostSerial / writeIndexToBoostSerial (:201-205,336-344; grammar in SURVEY.md App. B.1 —
 * the bytes are pinned against the real Boost.Serialization library, tests/test_database_format.py). kslam_index_db fills a kslam_sam_db view whose
 * pointers stay valid until kslam_index_free. */
typedef struct kslam_index kslam_index;
int kslam_index_parse_genbank(const char *const *paths, uint64_t n_paths, kslam_index **out);
int kslam_index_parse_fasta(const char *const *paths, uint64_t n_paths, kslam_index **out);
int kslam_index_read(const char *database_path, kslam_index **out);
int kslam_index_write(const kslam_index *index, const char *database_path);
int kslam_index_db(const kslam_index *index, kslam_sam_db *out);
const char *kslam_index_error(void);
void kslam_index_free(kslam_index *index);

/* Stage taps for parity tests (results of the last batch; copy to caller buffers; pass NULL to query
 * the count). Returns the count or a negative error. */
int64_t kslam_get_genome_kmers(kslam_ctx *ctx, kslam_kmer *out, uint64_t cap);      /* sorted, resident */
int64_t kslam_get_read_kmers(kslam_ctx *ctx, kslam_kmer *out, uint64_t cap);        /* sorted by kmer */
int64_t kslam_get_raw_seeds(kslam_ctx *ctx, kslam_seed *out, uint64_t cap);         /* after the join */
int64_t kslam_get_seeds(kslam_ctx *ctx, kslam_seed *out, uint64_t cap);             /* after sort+unique */
/* Generic hand-written LSD radix sort of 16-byte records by bits [lo_bit,hi_bit) of their first u64
 * (exposed for parity tests and the sort microbenchmark). In place on a host buffer. */
int kslam_sort_records(kslam_ctx *ctx, kslam_kmer *recs, uint64_t n, uint32_t lo_bit, uint32_t hi_bit,
                       float *device_ms /* may be NULL */);

int kslam_get_timings(const kslam_ctx *ctx, kslam_timings *out);
/* Issue-rate microbenchmark of the packed-int16 DPX op the SW sweeps are made of (VIADDMNMX.S16x2), in thread-ops
 * per second: the denominator of the integer-pipe roofline. */
int kslam_measure_int_peak(kslam_ctx *ctx, double *ops_per_s);
/* Prefilter (default on): read k-mers whose hash misses a bitmap of the genome k-mers are dropped while they are
 * extracted — they cannot seed (Overlap.h:157,236-239) — so only the survivors are written, sorted and joined.
 * Results are identical either way; with the filter off the read k-mer tap holds every record (KMer.h:160-181). */
int kslam_set_prefilter(kslam_ctx *ctx, int on);
/* Read k-mer records are radix-sorted on their leading `bits` bits only before the merge-join (it binary-searches each
 * record inside its tile's genome sub-range, so a total order is not needed): 0 (default) = log2(genome k-mers) + 2
 * rounded up to whole 8-bit digits, 64 = total order as KMer.h:388-398. The seed multiset is identical either way;
 * the read k-mer tap is ordered on those bits. */
int kslam_set_kmer_sort_bits(kslam_ctx *ctx, uint32_t bits);
int kslam_get_kmer_sort_bits(const kslam_ctx *ctx);
/* Banded Smith-Waterman tiers: 0 = full-matrix kernel only; 1 = alignments whose optimum is provably inside a
 * 32-diagonal band run in the banded kernel (sweep, then verify); 2 = additionally a 64-diagonal tier for what that
 * sweep bounded but could not prove; 3 (default) = additionally bands of 8 / 16 / 32 / 48 / 64 diagonals placed directly
 * from a lower bound of the score (best ungapped segment on the seed's diagonal; the forward score for the reverse
 * pass). Whatever cannot be proven falls through to the full-matrix kernel. Results are identical at every level
 * (DESIGN.md §3.4). */
int kslam_set_sw_band(kslam_ctx *ctx, int level);
/* reportCigar (Globals.h:36; true iff --sam-file, SLAM.h:169) for the batches that follow: kslam_params.report_cigar, changeable
 * between batches (one upload of a read / window set can then be aligned with and without CIGARs). */
int kslam_set_report_cigar(kslam_ctx *ctx, int on);
/* Keep (1, default) or drop (0) stage-tap buffers between stages; dropping saves HBM on big batches. */
int kslam_set_debug_taps(kslam_ctx *ctx, int keep);

#ifdef __cplusplus
}
#endif
#endif /* KSLAM_H_ */
