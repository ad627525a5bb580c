// api.cu — the C ABI of libkslam.so (include/kslam.h) and the per-batch orchestration that restates
// alignToDatabase (/root/reference/src/SLAM.h:60-79) as a chain of kernel launches on one stream.
#include "common.cuh"
#include <string.h>
#include <exception>

static thread_local std::string g_create_err;

cudaEvent_t tm_mark(kslam_ctx *c) {
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    c->ev_pool.push_back(e);
  }
  cudaEvent_t e = c->ev_pool[c->ev_used++];
  CUDA_TRY(cudaEventRecord(e, c->stream));
  return e;
}
float tm_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  if (a == b) return 0;
  cudaEventSynchronize(b);
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

__global__ void k_read_small(uint32_t *__restrict__ dst_host, const uint32_t *__restrict__ src, uint32_t n_words) {
  for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) dst_host[i] = src[i];
  __threadfence_system();
}
void read_small(kslam_ctx *c, void *host_pinned, const void *dev, size_t bytes) {
  k_read_small<<<1, 32, 0, c->stream>>>((uint32_t *)host_pinned, (const uint32_t *)dev, (uint32_t)(bytes / 4));
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

int api_fail(kslam_ctx *c, int code, const std::string &msg) {
  if (c) c->err = msg; else g_create_err = msg;
  return code;
}
static int fail(kslam_ctx *c, int code, const std::string &msg) { return api_fail(c, code, msg); }

extern "C" {

const char *kslam_version(void) { return "kslam-b200 0.1 (sm_100a)"; }

const char *kslam_last_error(const kslam_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int kslam_params_fast(const kslam_params *p) {
  if (!p) return 0;
  // SSW's striped kernels compute plain Gotoh when gap_extend < gap_open (its lazy-F loops are then a full
  // F) and mismatch <= 2 * gap_extend (its "E before lazy-F" shortcut then only drops dominated paths). Inside that domain
  // the packed band / wavefront kernels run; outside it k_sw_striped restates the striped kernels lane for lane.
  return p->match >= 1 && p->match <= 127 && p->mismatch <= 128 && p->gap_extend < p->gap_open && p->mismatch <= 2 * p->gap_extend;
}
int kslam_params_exact(const kslam_params *p) { return p != nullptr; }   // every parameter set is bit-exact now

int kslam_create(const kslam_params *params, kslam_ctx **out) {
  if (!params || !out) return fail(nullptr, KSLAM_ERR_ARG, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, KSLAM_ERR_CUDA, "no CUDA device: libkslam has no CPU fallback");
  }
  if (params->device < 0 || params->device >= ndev) return fail(nullptr, KSLAM_ERR_ARG, "device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, params->device) != cudaSuccess) return fail(nullptr, KSLAM_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) {
    char buf[160];
    snprintf(buf, sizeof buf, "device %d is sm_%d%d; libkslam is built for sm_100a only", params->device, prop.major, prop.minor);
    return fail(nullptr, KSLAM_ERR_CUDA, buf);
  }
  kslam_ctx *c = nullptr;
  try {
    c = new kslam_ctx();
    c->prm = *params;
    if (c->prm.genome_gap == 0) c->prm.genome_gap = KSLAM_K / 2;
    if (c->prm.max_cigar_ops == 0) c->prm.max_cigar_ops = 32;
    c->device = params->device;
    c->num_sms = prop.multiProcessorCount;
    memset(&c->tm, 0, sizeof c->tm);
    CUDA_TRY(cudaSetDevice(c->device));
    if (params->stream_priority) {
      int lo = 0, hi = 0;
      CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // numerically lower = higher priority
      CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
    } else CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->counters.reserve(64 * 8);
    c->h_counters.reserve(64 * 8);
  } catch (const CudaError &e) {
    delete c;
    return fail(nullptr, KSLAM_ERR_CUDA, std::string("create: ") + e.what + ": " + cudaGetErrorString(e.e));
  }
  *out = c;
  return KSLAM_OK;
}

void kslam_destroy(kslam_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  c->genomes.release(); c->reads.release(); c->swq.release(); c->swr.release();
  DevBuf *bufs[] = {&c->g_keys, &c->g_vals, &c->recA, &c->recB, &c->sort_hist, &c->scan_tmp, &c->counters,
                    &c->raw_seeds, &c->seedA, &c->seedB, &c->seed_keep, &c->seeds, &c->ov, &c->cig, &c->cig_dense, &c->pair_keys,
                    &c->pair_keys2, &c->ov_sorted, &c->cig_sorted, &c->pair_cnt, &c->pairs, &c->bitmap, &c->d_bounds,
                    &c->part_send, &c->part_recv, &c->part_tmp, &c->part_m, &c->part_msend, &c->part_mrecv, &c->pairs_compact, &c->far_mates, &c->insert_hist};
  for (DevBuf *b : bufs) b->release();
  HostBuf *hb[] = {&c->h_stage, &c->h_counters, &c->h_ov, &c->h_cig, &c->h_ov_sorted, &c->h_cig_sorted, &c->h_pairs, &c->h_pairs_compact, &c->h_far_mates, &c->h_insert_hist};
  for (HostBuf *b : hb) b->release();
  sw_workspace_free(c);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int kslam_set_prefilter(kslam_ctx *ctx, int on) {
  if (!ctx) return KSLAM_ERR_ARG;
  ctx->prefilter = on != 0;
  return KSLAM_OK;
}

int kslam_device_memory(int32_t device, uint64_t *free_bytes, uint64_t *total_bytes) {
  size_t f = 0, t = 0;
  if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) { cudaGetLastError(); return KSLAM_ERR_CUDA; }
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return KSLAM_OK;
}

int kslam_set_report_cigar(kslam_ctx *ctx, int on) {
  if (!ctx) return KSLAM_ERR_ARG;
  ctx->prm.report_cigar = on != 0;
  ctx->aligned = false; ctx->paired = false;      // results of the other mode are not mixed with this one
  return KSLAM_OK;
}

int kslam_set_sw_band(kslam_ctx *ctx, int on) {
  if (!ctx) return KSLAM_ERR_ARG;
  ctx->sw_band = on != 0;          // 1: 32-wide sweep tier only; 2: + 64-wide tier; >= 3 (default): + direct tiers
  ctx->sw_band64 = on >= 2;
  ctx->sw_tiers = on >= 3;
  return KSLAM_OK;
}

int kslam_set_kmer_sort_bits(kslam_ctx *ctx, uint32_t bits) {
  if (!ctx || bits > 64) return KSLAM_ERR_ARG;
  ctx->sort_bits = bits;
  return KSLAM_OK;
}

int kslam_get_kmer_sort_bits(const kslam_ctx *ctx) { return ctx ? (int)kmer_sort_bits(ctx) : KSLAM_ERR_ARG; }

int kslam_set_debug_taps(kslam_ctx *ctx, int keep) {
  if (!ctx) return KSLAM_ERR_ARG;
  ctx->keep_taps = keep != 0;
  return KSLAM_OK;
}

int kslam_measure_int_peak(kslam_ctx *c, double *ops_per_s) {
  API_BEGIN(c)
  if (!ops_per_s) return fail(c, KSLAM_ERR_ARG, "null output");
  *ops_per_s = sw_measure_int_peak(c);
  return KSLAM_OK;
  API_END(c)
}

int kslam_get_timings(const kslam_ctx *ctx, kslam_timings *out) {
  if (!ctx || !out) return KSLAM_ERR_ARG;
  *out = ctx->tm;
  out->kernel_launches = ctx->launches;
  return KSLAM_OK;
}

__global__ void __launch_bounds__(256) k_split_recs(const Rec16 *__restrict__ in, uint64_t n, uint64_t *__restrict__ keys,
                                                    uint64_t *__restrict__ vals) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Rec16 r = in[i]; keys[i] = r.key; vals[i] = r.val;
  }
}
// The reference orders equal k-mers by ID_isFromGB_RC descending (KMer.h:392-396). Inside a pile only the split
// "genome records first" matters and the genome list holds genome records only, but the tap promises the
// reference's order, so the genome list is sorted on ~id_flags as the secondary key.
__global__ void __launch_bounds__(256) k_flip_idflags(Rec16 *__restrict__ r, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    r[i].val ^= 0xffffffffull;
}

}  // extern "C"

void finish_genome_index(kslam_ctx *c, DevBuf &a, DevBuf &b, uint64_t n) {
  uint64_t blocks = (n + 255) / 256, maxb = (uint64_t)c->num_sms * 16;
  if (blocks > maxb) blocks = maxb;
  k_flip_idflags<<<(unsigned)blocks, 256, 0, c->stream>>>(a.as<Rec16>(), n);
  uint64_t passes = 0;
  Rec16 *cur = radix_sort(c, a.as<Rec16>(), b.as<Rec16>(), n, 1, 0, 32, &passes);   // ~id_flags ascending
  cur = radix_sort(c, cur, cur == a.as<Rec16>() ? b.as<Rec16>() : a.as<Rec16>(), n, 0, 0, 64, &passes);
  k_flip_idflags<<<(unsigned)blocks, 256, 0, c->stream>>>(cur, n);
  c->g_keys.reserve((size_t)n * 8 + 64); c->g_vals.reserve((size_t)n * 8 + 64);
  k_split_recs<<<(unsigned)blocks, 256, 0, c->stream>>>(cur, n, c->g_keys.as<uint64_t>(), c->g_vals.as<uint64_t>());
  c->launches += 3;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(c->stream));
}

extern "C" {

int kslam_load_genomes(kslam_ctx *c, uint64_t n, const char *bases, const uint64_t *offs) {
  API_BEGIN(c)
  if (!offs || (n && !bases && offs[n] != offs[0])) return fail(c, KSLAM_ERR_ARG, "null genome buffers");
  if (n >= (1ull << 30)) return fail(c, KSLAM_ERR_ARG, "more than 2^30 entries (KMer.h:65-66)");
  for (uint64_t i = 0; i < n; i++)
    if (offs[i + 1] - offs[i] >= (1ull << 31)) return fail(c, KSLAM_ERR_ARG, "entry longer than 2^31 bases");
  c->genomes_loaded = false; c->aligned = false;
  pack_sequences(c, c->genomes, n, bases, offs, c->prm.genome_gap, false);
  c->max_genome_len = c->genomes.max_len;
  c->n_gk = c->genomes.n_kmers;
  if (c->n_gk) {
    DevBuf a, b;
    a.reserve((size_t)c->n_gk * sizeof(Rec16)); b.reserve((size_t)c->n_gk * sizeof(Rec16));
    extract_kmers(c, c->genomes, true, c->prm.genome_gap, a.as<Rec16>());
    finish_genome_index(c, a, b, c->n_gk);
    a.release(); b.release();
  }
  c->part = 0; c->n_parts = 1; c->n_gk_total = c->n_gk;
  c->splitters.assign({0ull, ~0ull});
  build_prefilter(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->tm.n_genome_kmers = c->n_gk;
  c->genomes_loaded = true;
  return KSLAM_OK;
  API_END(c)
}

static void fetch_alignments(kslam_ctx *c, kslam_alignments *out) {
  const uint64_t n = c->n_seeds;
  cudaEvent_t e0 = tm_mark(c);
  c->h_ov.reserve((size_t)n * sizeof(kslam_overlap) + 64);
  if (n) CUDA_TRY(cudaMemcpyAsync(c->h_ov.p, c->ov.p, (size_t)n * sizeof(kslam_overlap), cudaMemcpyDeviceToHost, c->stream));
  const bool with_cig = c->prm.report_cigar && n;
  if (with_cig) {
    c->h_cig.reserve((size_t)c->n_cig_words * 4 + 64);
    if (c->n_cig_words) CUDA_TRY(cudaMemcpyAsync(c->h_cig.p, c->cig_dense.p, (size_t)c->n_cig_words * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  cudaEvent_t e1 = tm_mark(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->tm.ms_d2h = tm_ms(e0, e1);
  if (out) {
    out->n_overlaps = n; out->overlaps = c->h_ov.as<kslam_overlap>();
    out->n_cigar_words = with_cig ? c->n_cig_words : 0; out->cigar_pool = with_cig ? c->h_cig.as<uint32_t>() : nullptr;
  }
}

// device part of alignToDatabase on reads that are already packed in HBM
static void align_device(kslam_ctx *c) {
  cudaStream_t st = c->stream;
  c->ev_used = 0;
  cudaEvent_t e0 = tm_mark(c);
  c->n_rk = c->reads.n_kmers;
  c->tm.n_read_kmers = c->n_rk; c->tm.n_sort_passes = 0;
  c->sorted_rk = nullptr;
  if (c->n_rk) {
    if (c->prefilter && c->filter_bits) {
      // fused extract + prefilter: n_rk becomes the number of records that can still match a genome k-mer
      c->n_rk = extract_read_kmers_filtered(c, c->reads, c->recA);
    } else {
      c->recA.reserve((size_t)c->n_rk * sizeof(Rec16));
      extract_kmers(c, c->reads, false, 1, c->recA.as<Rec16>());
    }
    c->recB.reserve((size_t)c->n_rk * sizeof(Rec16) + 64);
  }
  c->tm.n_sorted_kmers = c->n_rk;
  cudaEvent_t e1 = tm_mark(c);
  if (c->n_rk) {
    uint64_t passes = 0;
    c->sorted_rk = radix_sort(c, c->recA.as<Rec16>(), c->recB.as<Rec16>(), c->n_rk, 0, 64 - kmer_sort_bits(c), 64, &passes);
    c->tm.n_sort_passes += passes;
  }
  cudaEvent_t e2 = tm_mark(c);
  join_and_unique(c);
  sw_align_seeds(c);
  cudaEvent_t e3 = tm_mark(c);
  CUDA_TRY(cudaStreamSynchronize(st));
  c->tm.ms_extract = tm_ms(e0, e1);
  c->tm.ms_sort = tm_ms(e1, e2);
  c->tm.ms_total = tm_ms(e0, e3);
  c->aligned = true;
}

int kslam_upload_reads(kslam_ctx *c, uint64_t n, const char *bases, const uint64_t *offs) {
  API_BEGIN(c)
  if (!c->genomes_loaded) return fail(c, KSLAM_ERR_STATE, "kslam_load_genomes first");
  if (!offs || (n && !bases && offs[n] != offs[0])) return fail(c, KSLAM_ERR_ARG, "null read buffers");
  if (n >= (1ull << 30)) return fail(c, KSLAM_ERR_ARG, "more than 2^30 reads (KMer.h:65-66)");
  c->reads_loaded = false; c->aligned = false; c->paired = false;
  c->ev_used = 0;
  cudaEvent_t e0 = tm_mark(c);
  pack_sequences(c, c->reads, n, bases, offs, 1, true);
  cudaEvent_t e1 = tm_mark(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->tm.ms_h2d = 0; c->tm.ms_pack = tm_ms(e0, e1);   // copy + pack (the copy is inside pack_sequences)
  c->reads_loaded = true;
  return KSLAM_OK;
  API_END(c)
}

int kslam_align_resident(kslam_ctx *c, int fetch, kslam_alignments *out) {
  API_BEGIN(c)
  if (!c->reads_loaded) return fail(c, KSLAM_ERR_STATE, "kslam_upload_reads first");
  align_device(c);
  if (fetch) fetch_alignments(c, out);
  else if (out) { out->n_overlaps = c->n_seeds; out->overlaps = nullptr; out->n_cigar_words = 0; out->cigar_pool = nullptr; }
  return KSLAM_OK;
  API_END(c)
}

int kslam_align_batch(kslam_ctx *c, uint64_t n, const char *bases, const uint64_t *offs, kslam_alignments *out) {
  int rc = kslam_upload_reads(c, n, bases, offs);
  if (rc != KSLAM_OK) return rc;
  float pack = c->tm.ms_pack;
  rc = kslam_align_resident(c, 1, out);
  c->tm.ms_pack = pack;
  return rc;
}

int kslam_part_finish(kslam_ctx *c, uint64_t n_matches, uint32_t read_id_base, int fetch, kslam_alignments *out) {
  API_BEGIN(c)
  if (!c->reads_loaded) return fail(c, KSLAM_ERR_STATE, "kslam_upload_reads first");
  if (n_matches * sizeof(Rec16) > c->part_mrecv.cap) return fail(c, KSLAM_ERR_ARG, "more matches than kslam_part_match_buffer reserved");
  c->ev_used = 0;
  cudaEvent_t e0 = tm_mark(c);
  part_matches_to_seeds(c, n_matches, read_id_base);
  seed_sort_unique(c);
  sw_align_seeds(c);
  cudaEvent_t e3 = tm_mark(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->tm.ms_total = tm_ms(e0, e3);
  c->aligned = true;
  if (fetch) fetch_alignments(c, out);
  else if (out) { out->n_overlaps = c->n_seeds; out->overlaps = nullptr; out->n_cigar_words = 0; out->cigar_pool = nullptr; }
  return KSLAM_OK;
  API_END(c)
}

// D2H of the pair stage's results (pair-sorted overlaps, dense CIGAR pool, pairs) into the ctx's pinned buffers
static void fetch_pairs(kslam_ctx *c, kslam_pairs *out) {
  const bool with_cig = c->prm.report_cigar && c->n_sorted && c->cig_dense.p;
  c->h_ov_sorted.reserve((size_t)c->n_sorted * sizeof(kslam_overlap) + 64);
  c->h_pairs.reserve((size_t)c->n_pairs * sizeof(kslam_pair) + 64);
  if (c->n_sorted) CUDA_TRY(cudaMemcpyAsync(c->h_ov_sorted.p, c->ov_sorted.p, (size_t)c->n_sorted * sizeof(kslam_overlap), cudaMemcpyDeviceToHost, c->stream));
  if (c->n_pairs) CUDA_TRY(cudaMemcpyAsync(c->h_pairs.p, c->pairs.p, (size_t)c->n_pairs * sizeof(kslam_pair), cudaMemcpyDeviceToHost, c->stream));
  if (with_cig) {
    c->h_cig_sorted.reserve((size_t)c->n_cig_words * 4 + 64);
    if (c->n_cig_words) CUDA_TRY(cudaMemcpyAsync(c->h_cig_sorted.p, c->cig_dense.p, (size_t)c->n_cig_words * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (out) {
    out->n_sorted = c->n_sorted; out->n_pairs = c->n_pairs;
    out->sorted_overlaps = c->h_ov_sorted.as<kslam_overlap>(); out->pairs = c->h_pairs.as<kslam_pair>();
    out->n_cigar_words = with_cig ? c->n_cig_words : 0;
    out->cigar_pool = with_cig ? c->h_cig_sorted.as<uint32_t>() : nullptr;
  }
}

int kslam_pair_batch(kslam_ctx *c, int fetch, kslam_pairs *out) {
  API_BEGIN(c)
  if (!c->aligned) return fail(c, KSLAM_ERR_STATE, "kslam_align_batch first");
  c->ev_used = 0;
  c->paired = false;
  cudaEvent_t e0 = tm_mark(c);
  pair_overlaps(c);
  cudaEvent_t e1 = tm_mark(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->tm.ms_pair = tm_ms(e0, e1);
  c->tm.n_pairs = c->n_pairs;
  c->paired = true;
  if (fetch) fetch_pairs(c, out);
  else if (out) {
    out->n_sorted = c->n_sorted; out->n_pairs = c->n_pairs; out->sorted_overlaps = nullptr; out->pairs = nullptr;
    out->n_cigar_words = 0; out->cigar_pool = nullptr;
  }
  return KSLAM_OK;
  API_END(c)
}

int kslam_fetch_pairs(kslam_ctx *c, kslam_pairs *out) {
  API_BEGIN(c)
  if (!c->paired) return fail(c, KSLAM_ERR_STATE, "kslam_pair_batch first");
  fetch_pairs(c, out);
  return KSLAM_OK;
  API_END(c)
}

int kslam_fetch_pairs_compact(kslam_ctx *c, uint32_t host_threads, kslam_pairs_compact *out) {
  API_BEGIN(c)
  if (!c->paired) return fail(c, KSLAM_ERR_STATE, "kslam_pair_batch first");
  if (!out) return fail(c, KSLAM_ERR_ARG, "null output");
  const uint64_t n = c->n_pairs;
  if (n >> 32) return fail(c, KSLAM_ERR_ARG, "more than 2^32 pair records in one batch");
  c->pairs_compact.reserve((size_t)n * sizeof(kslam_pair_compact) + 64);
  c->h_pairs_compact.reserve((size_t)n * sizeof(kslam_pair_compact) + 64);
  // The batch's insert-size limit is a function of the value counts: they are taken on the device (k_insert_hist) and only
  // the few hundred counters in use cross PCIe before the records do, so the mates of the pairs beyond the limit are
  // picked while the records are still on their way. (Values outside (0, 2^22): counted from the records on the host.)
  const unsigned long long *hist = nullptr; uint32_t top = 0;
  const bool counted = insert_hist_device(c, &hist, &top);
  pairs_compact_device(c, c->pairs_compact.as<kslam_pair_compact>());
  if (n) CUDA_TRY(cudaMemcpyAsync(c->h_pairs_compact.p, c->pairs_compact.p, (size_t)n * sizeof(kslam_pair_compact), cudaMemcpyDeviceToHost, c->stream));
  out->n_pairs = n; out->pairs = c->h_pairs_compact.as<kslam_pair_compact>();
  if (counted) out->insert_size_limit = kslam_insert_size_limit_counts((const uint64_t *)hist, top);
  else { CUDA_TRY(cudaStreamSynchronize(c->stream)); out->insert_size_limit = kslam_insert_size_limit_compact(out->pairs, n, host_threads); }
  const uint64_t n_far = far_mates_device(c, out->insert_size_limit, c->far_mates);
  c->h_far_mates.reserve((size_t)n_far * sizeof(kslam_far_mates) + 64);
  if (n_far) CUDA_TRY(cudaMemcpyAsync(c->h_far_mates.p, c->far_mates.p, (size_t)n_far * sizeof(kslam_far_mates), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  out->n_far = n_far; out->far = c->h_far_mates.as<kslam_far_mates>();
  return KSLAM_OK;
  API_END(c)
}

int kslam_fetch_far_mates(kslam_ctx *c, uint32_t insert_size_limit, uint64_t *n_far, const kslam_far_mates **far) {
  API_BEGIN(c)
  if (!c->paired) return fail(c, KSLAM_ERR_STATE, "kslam_pair_batch first");
  if (!n_far || !far) return fail(c, KSLAM_ERR_ARG, "null output");
  const uint64_t n = far_mates_device(c, insert_size_limit, c->far_mates);
  c->h_far_mates.reserve((size_t)n * sizeof(kslam_far_mates) + 64);
  if (n) CUDA_TRY(cudaMemcpyAsync(c->h_far_mates.p, c->far_mates.p, (size_t)n * sizeof(kslam_far_mates), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  *n_far = n; *far = c->h_far_mates.as<kslam_far_mates>();
  return KSLAM_OK;
  API_END(c)
}

// alignToDatabase + screen + getPairedOverlaps in one call (the body of the reference's batch loop, SLAM.h:209-214):
// the unsorted alignment vector never leaves the GPU, only what the loop keeps (pair-sorted overlaps, CIGARs, pairs).
int kslam_align_pair_batch(kslam_ctx *c, uint64_t n, const char *bases, const uint64_t *offs, kslam_pairs *out) {
  int rc = kslam_upload_reads(c, n, bases, offs);
  if (rc != KSLAM_OK) return rc;
  const float pack = c->tm.ms_pack;
  rc = kslam_align_resident(c, 0, nullptr);
  c->tm.ms_pack = pack;
  if (rc != KSLAM_OK) return rc;
  return kslam_pair_batch(c, 1, out);
}

// ---- Aligner::Align batch ------------------------------------------------------------------------
int kslam_ssw_upload(kslam_ctx *c, uint64_t n, const char *q, const uint64_t *qoffs, const char *r, const uint64_t *roffs) {
  API_BEGIN(c)
  if (!qoffs || !roffs) return fail(c, KSLAM_ERR_ARG, "null offsets");
  if (n >= (1ull << 31)) return fail(c, KSLAM_ERR_ARG, "too many pairs");
  c->sw_loaded = false;
  pack_sequences(c, c->swq, n, q, qoffs, 1, true);
  pack_sequences(c, c->swr, n, r, roffs, 1, true);
  c->sw_loaded = true;
  return KSLAM_OK;
  API_END(c)
}

int kslam_ssw_resident(kslam_ctx *c, kslam_overlap *out, uint32_t *cigar_pool) {
  API_BEGIN(c)
  if (!c->sw_loaded) return fail(c, KSLAM_ERR_STATE, "kslam_ssw_upload first");
  const uint64_t n = c->swq.n;
  const uint32_t cap = c->prm.max_cigar_ops;
  c->ev_used = 0;
  cudaEvent_t e0 = tm_mark(c);
  c->ov.reserve((size_t)n * sizeof(kslam_overlap) + 64);
  if (c->prm.report_cigar) c->cig.reserve((size_t)n * cap * 4 + 64);
  sw_align_pairs(c, n, c->ov.as<kslam_overlap>(), c->prm.report_cigar ? c->cig.as<uint32_t>() : nullptr);
  cudaEvent_t e1 = tm_mark(c);
  if (out && n) CUDA_TRY(cudaMemcpyAsync(out, c->ov.p, (size_t)n * sizeof(kslam_overlap), cudaMemcpyDeviceToHost, c->stream));
  if (cigar_pool && n && c->prm.report_cigar)
    CUDA_TRY(cudaMemcpyAsync(cigar_pool, c->cig.p, (size_t)n * cap * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->tm.ms_total = tm_ms(e0, e1);
  c->aligned = false;
  return KSLAM_OK;
  API_END(c)
}

int kslam_ssw_batch(kslam_ctx *c, uint64_t n, const char *q, const uint64_t *qoffs, const char *r, const uint64_t *roffs,
                    kslam_overlap *out, uint32_t *cigar_pool) {
  int rc = kslam_ssw_upload(c, n, q, qoffs, r, roffs);
  if (rc != KSLAM_OK) return rc;
  return kslam_ssw_resident(c, out, cigar_pool);
}

// ---- stage taps -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_merge_recs(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals,
                                                    uint64_t n, Rec16 *__restrict__ out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Rec16 r; r.key = keys[i]; r.val = vals[i]; out[i] = r;
  }
}

int64_t kslam_get_genome_kmers(kslam_ctx *c, kslam_kmer *out, uint64_t cap) {
  API_BEGIN(c)
  if (!c->genomes_loaded) return fail(c, KSLAM_ERR_STATE, "no genomes");
  if (!out) return (int64_t)c->n_gk;
  if (cap < c->n_gk) return fail(c, KSLAM_ERR_ARG, "buffer too small");
  if (c->n_gk) {
    DevBuf tmp; tmp.reserve((size_t)c->n_gk * sizeof(Rec16));
    uint64_t blocks = (c->n_gk + 255) / 256, maxb = (uint64_t)c->num_sms * 16;
    if (blocks > maxb) blocks = maxb;
    k_merge_recs<<<(unsigned)blocks, 256, 0, c->stream>>>(c->g_keys.as<uint64_t>(), c->g_vals.as<uint64_t>(), c->n_gk, tmp.as<Rec16>());
    CUDA_TRY(cudaMemcpyAsync(out, tmp.p, (size_t)c->n_gk * sizeof(Rec16), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    tmp.release();
  }
  return (int64_t)c->n_gk;
  API_END(c)
}

int64_t kslam_get_read_kmers(kslam_ctx *c, kslam_kmer *out, uint64_t cap) {
  API_BEGIN(c)
  if (!c->aligned) return fail(c, KSLAM_ERR_STATE, "no batch");
  if (!out) return (int64_t)c->n_rk;
  if (cap < c->n_rk) return fail(c, KSLAM_ERR_ARG, "buffer too small");
  if (c->n_rk) {
    CUDA_TRY(cudaMemcpyAsync(out, c->sorted_rk, (size_t)c->n_rk * sizeof(Rec16), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return (int64_t)c->n_rk;
  API_END(c)
}

int64_t kslam_get_raw_seeds(kslam_ctx *c, kslam_seed *out, uint64_t cap) {
  API_BEGIN(c)
  if (!c->aligned) return fail(c, KSLAM_ERR_STATE, "no batch");
  if (!c->keep_taps) return fail(c, KSLAM_ERR_STATE, "taps disabled");
  if (!out) return (int64_t)c->n_raw;
  if (cap < c->n_raw) return fail(c, KSLAM_ERR_ARG, "buffer too small");
  if (c->n_raw) {
    std::vector<Rec16> tmp(c->n_raw);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), c->raw_seeds.p, (size_t)c->n_raw * sizeof(Rec16), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const uint32_t bias = c->reads.max_len;
    for (uint64_t i = 0; i < c->n_raw; i++) {   // unpack the sort representation (join.cu: pack_seed)
      out[i].read = (uint32_t)(tmp[i].key >> 32); out[i].entry = (uint32_t)tmp[i].key;
      out[i].rel = (int32_t)((uint32_t)(tmp[i].val >> 1) - bias); out[i].rev_comp = (uint32_t)(tmp[i].val & 1);
    }
  }
  return (int64_t)c->n_raw;
  API_END(c)
}

int64_t kslam_get_seeds(kslam_ctx *c, kslam_seed *out, uint64_t cap) {
  API_BEGIN(c)
  if (!c->aligned) return fail(c, KSLAM_ERR_STATE, "no batch");
  if (!out) return (int64_t)c->n_seeds;
  if (cap < c->n_seeds) return fail(c, KSLAM_ERR_ARG, "buffer too small");
  if (c->n_seeds) {
    CUDA_TRY(cudaMemcpyAsync(out, c->seeds.p, (size_t)c->n_seeds * sizeof(kslam_seed), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  return (int64_t)c->n_seeds;
  API_END(c)
}

int kslam_sort_records(kslam_ctx *c, kslam_kmer *recs, uint64_t n, uint32_t lo_bit, uint32_t hi_bit, float *device_ms) {
  API_BEGIN(c)
  if (!recs && n) return fail(c, KSLAM_ERR_ARG, "null records");
  if (n) {
    DevBuf a, b;
    a.reserve((size_t)n * sizeof(Rec16)); b.reserve((size_t)n * sizeof(Rec16));
    CUDA_TRY(cudaMemcpyAsync(a.p, recs, (size_t)n * sizeof(Rec16), cudaMemcpyHostToDevice, c->stream));
    c->ev_used = 0;
    cudaEvent_t e0 = tm_mark(c);
    uint64_t passes = 0;
    Rec16 *cur = radix_sort(c, a.as<Rec16>(), b.as<Rec16>(), n, 0, lo_bit, hi_bit, &passes);
    cudaEvent_t e1 = tm_mark(c);
    CUDA_TRY(cudaMemcpyAsync(recs, cur, (size_t)n * sizeof(Rec16), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (device_ms) *device_ms = tm_ms(e0, e1);
    c->tm.n_sort_passes = passes;
    a.release(); b.release();
  } else if (device_ms) *device_ms = 0;
  return KSLAM_OK;
  API_END(c)
}

}  // extern "C"
