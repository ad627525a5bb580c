// dist.cu — the matching path over a k-mer-RANGE PARTITIONED genome index (BASELINE config 4).
//
// When the sorted genome k-mer list (KMer.h:388-398; 16 B x 1.25 G records for 5,000 genomes) does not fit one GPU it
// is cut into n_parts contiguous kMerInt ranges. Equal k-mers always share an owner, so no pile (Overlap.h:153-199)
// is ever split — the same invariant the reference's own chunking keeps (Overlap.h:285-287). Per batch there is one
// exchange step each way, carried by the host side over NCCL all-to-all (k-slam_b200/dist.py):
//
//   read owner   kslam_part_route_kmers  extract + prefilter read k-mers (job-global read ids), bucket by key owner
//                 ---- all-to-all of 16 B k-mer records ---->
//   key owner    kslam_part_join         LSD radix sort, merge-join against the local key range, bucket the raw
//                                        matches {read record, genome record} by read owner
//                 <--- all-to-all of 16 B match records -----
//   read owner   kslam_part_finish       match -> seed (needs the read length), seed sort, fuzzy unique, SW, CIGAR
//
// Genome BASES (bit planes) and the prefilter bitmap are replicated: windows are gathered locally. Splitters come
// from a sorted sample of the genome k-mers and are a pure function of the database, so every rank computes the
// same ones without talking. Everything after the second exchange is the single-GPU path unchanged, which is why
// the result per rank is bit-identical to kslam_align_batch on that rank's reads.
#include "common.cuh"

#define DIST_MAX_PARTS 64

// bucket of a record: MODE 0 = key owner (number of splitters[1..n_parts) <= kmer), MODE 1 = read owner
// (number of id_bases[1..n_parts) <= global read id, bits 0-29 of the record's first word)
template <int MODE>
__device__ __forceinline__ uint32_t bucket_of(const uint64_t *s_bounds, uint32_t n_parts, uint64_t first_word) {
  const uint64_t v = MODE == 0 ? first_word : (first_word & 0x3FFFFFFFull);
  uint32_t lo = 0, hi = n_parts - 1;            // answer = count of bounds[1..n_parts-1] <= v
  while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (s_bounds[mid] <= v) lo = mid; else hi = mid - 1; }
  return lo;
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_bucket_count(const Rec16 *__restrict__ in, uint64_t n, const uint64_t *__restrict__ bounds, uint32_t n_parts,
               unsigned long long *__restrict__ counts) {
  __shared__ uint64_t s_bounds[DIST_MAX_PARTS];
  __shared__ uint32_t s_hist[DIST_MAX_PARTS];
  if (threadIdx.x < DIST_MAX_PARTS) { s_bounds[threadIdx.x] = threadIdx.x < n_parts ? bounds[threadIdx.x] : ~0ull; s_hist[threadIdx.x] = 0; }
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t b = bucket_of<MODE>(s_bounds, n_parts, __ldg(&in[i].key));
    atomicAdd(&s_hist[b], 1u);
  }
  __syncthreads();
  if (threadIdx.x < n_parts && s_hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
}

// cursors[b] starts at the exclusive prefix of the counts; each warp reserves one run per bucket it holds
template <int MODE>
__global__ void __launch_bounds__(256)
k_bucket_scatter(const Rec16 *__restrict__ in, uint64_t n, const uint64_t *__restrict__ bounds, uint32_t n_parts,
                 unsigned long long *__restrict__ cursors, Rec16 *__restrict__ out) {
  __shared__ uint64_t s_bounds[DIST_MAX_PARTS];
  if (threadIdx.x < DIST_MAX_PARTS) s_bounds[threadIdx.x] = threadIdx.x < n_parts ? bounds[threadIdx.x] : ~0ull;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t n_round = (n + 31) & ~31ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * blockDim.x) {
    const bool live = i < n;
    ulonglong2 r = make_ulonglong2(0, 0);
    uint32_t b = DIST_MAX_PARTS;                 // dead lanes form their own group and write nothing
    if (live) { r = __ldg(reinterpret_cast<const ulonglong2 *>(in + i)); b = bucket_of<MODE>(s_bounds, n_parts, r.x); }
    const uint32_t peers = __match_any_sync(0xffffffffu, b);
    const uint32_t leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (live && lane == leader) base = atomicAdd(&cursors[b], (unsigned long long)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) *reinterpret_cast<ulonglong2 *>(out + base + __popc(peers & ((1u << lane) - 1u))) = r;
  }
}

// Groups n records by bucket into `out` (order inside a bucket is arbitrary: they are sorted next); counts_host gets
// the n_parts bucket sizes. `bounds_host` holds n_parts lower bounds (bounds[0] is ignored: bucket 0 starts at 0).
template <int MODE>
static void bucket_records(kslam_ctx *c, const Rec16 *in, uint64_t n, const uint64_t *bounds_host, Rec16 *out,
                           uint64_t *counts_host) {
  cudaStream_t st = c->stream;
  const uint32_t P = c->n_parts;
  c->d_bounds.reserve(DIST_MAX_PARTS * 8 * 3);
  uint64_t *d_bounds = c->d_bounds.as<uint64_t>();
  unsigned long long *d_counts = c->d_bounds.as<unsigned long long>() + DIST_MAX_PARTS, *d_cursors = d_counts + DIST_MAX_PARTS;
  c->h_stage.reserve(DIST_MAX_PARTS * 8 * 3);
  uint64_t *h = c->h_stage.as<uint64_t>();
  for (uint32_t p = 0; p < P; p++) h[p] = bounds_host[p];
  CUDA_TRY(cudaMemcpyAsync(d_bounds, h, P * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(d_counts, 0, DIST_MAX_PARTS * 8, st));
  uint64_t blocks = (n + 255) / 256, maxb = (uint64_t)c->num_sms * 8;
  if (blocks > maxb) blocks = maxb;
  if (n) { k_bucket_count<MODE><<<(unsigned)blocks, 256, 0, st>>>(in, n, d_bounds, P, d_counts); c->launches++; }
  uint64_t *h_counts = h + DIST_MAX_PARTS;
  read_small(c, h_counts, d_counts, P * 8);
  CUDA_TRY(cudaStreamSynchronize(st));
  uint64_t *h_cur = h + 2 * DIST_MAX_PARTS, run = 0;
  for (uint32_t p = 0; p < P; p++) { counts_host[p] = h_counts[p]; h_cur[p] = run; run += h_counts[p]; }
  if (n) {
    CUDA_TRY(cudaMemcpyAsync(d_cursors, h_cur, P * 8, cudaMemcpyHostToDevice, st));
    k_bucket_scatter<MODE><<<(unsigned)blocks, 256, 0, st>>>(in, n, d_bounds, P, d_cursors, out);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaStreamSynchronize(st));   // h_stage is reused by the next call
}

// ---- index build: one pass over chunks of the flat genome k-mer list -----------------------------------------
// PASS 0: set every k-mer's bit in the (replicated) prefilter bitmap and count the records of this part's key range;
// PASS 1: append the records of the range to `out` (order arbitrary, sorted next).
template <int PASS>
__global__ void __launch_bounds__(256)
k_part_scan(const Rec16 *__restrict__ chunk, uint64_t n, uint64_t lo, uint64_t hi, uint32_t last_part, uint32_t bits,
            uint32_t *__restrict__ bitmap, unsigned long long *__restrict__ counter, Rec16 *__restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t n_round = (n + 31) & ~31ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * blockDim.x) {
    bool mine = false;
    ulonglong2 r = make_ulonglong2(0, 0);
    if (i < n) {
      r = __ldg(reinterpret_cast<const ulonglong2 *>(chunk + i));
      mine = r.x >= lo && (last_part || r.x < hi);
      if (PASS == 0 && r.x != 0) { const uint64_t h = kmer_hash(r.x, bits); atomicOr(&bitmap[h >> 5], 1u << (h & 31)); }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, mine);
    if (!m) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
    if (PASS == 1) {
      base = __shfl_sync(0xffffffffu, base, 0);
      if (mine) *reinterpret_cast<ulonglong2 *>(out + base + __popc(m & ((1u << lane) - 1u))) = r;
    }
  }
}

extern "C" {

int kslam_load_genomes_part(kslam_ctx *c, uint64_t n, const char *bases, const uint64_t *offs, uint32_t part, uint32_t n_parts) {
  API_BEGIN(c)
  if (!offs || (n && !bases && offs[n] != offs[0])) return api_fail(c, KSLAM_ERR_ARG, "null genome buffers");
  if (n_parts < 1 || n_parts > DIST_MAX_PARTS || part >= n_parts) return api_fail(c, KSLAM_ERR_ARG, "part / n_parts out of range (1..64)");
  if (n >= (1ull << 30)) return api_fail(c, KSLAM_ERR_ARG, "more than 2^30 entries (KMer.h:65-66)");
  for (uint64_t i = 0; i < n; i++)
    if (offs[i + 1] - offs[i] >= (1ull << 31)) return api_fail(c, KSLAM_ERR_ARG, "entry longer than 2^31 bases");
  cudaStream_t st = c->stream;
  c->genomes_loaded = false; c->aligned = false;
  pack_sequences(c, c->genomes, n, bases, offs, c->prm.genome_gap, false);
  c->max_genome_len = c->genomes.max_len;
  const uint64_t total = c->genomes.n_kmers;
  c->n_gk_total = total; c->n_gk = 0; c->part = part; c->n_parts = n_parts;
  c->splitters.assign(n_parts + 1, 0ull);
  c->splitters[n_parts] = ~0ull;
  unsigned long long *d_cnt = c->counters.as<unsigned long long>() + 4;
  unsigned long long *h_cnt = c->h_counters.as<unsigned long long>() + 4;
  c->filter_bits = 0;
  if (total) {
    // splitters: quantiles of a sorted sample (every stride-th record of the flat list)
    if (n_parts > 1) {
      const uint64_t stride = total / (4ull << 20) + 1, ns = (total + stride - 1) / stride;
      DevBuf a, b;
      a.reserve((size_t)ns * sizeof(Rec16)); b.reserve((size_t)ns * sizeof(Rec16));
      extract_genome_kmers_range(c, c->genomes, c->prm.genome_gap, 0, stride, ns, a.as<Rec16>());
      uint64_t passes = 0;
      Rec16 *sorted = radix_sort(c, a.as<Rec16>(), b.as<Rec16>(), ns, 0, 0, 64, &passes);
      c->h_stage.reserve(DIST_MAX_PARTS * 8 * 3);
      uint64_t *h = c->h_stage.as<uint64_t>();
      for (uint32_t p = 1; p < n_parts; p++)
        CUDA_TRY(cudaMemcpyAsync(h + p, &sorted[(uint64_t)p * ns / n_parts].key, 8, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      for (uint32_t p = 1; p < n_parts; p++) c->splitters[p] = h[p];
      a.release(); b.release();
    }
    const uint64_t lo = c->splitters[part], hi = c->splitters[part + 1];
    const uint32_t last = part == n_parts - 1;
    const uint32_t bits = prefilter_bits(total);
    c->bitmap.reserve((size_t)1 << (bits - 3));
    CUDA_TRY(cudaMemsetAsync(c->bitmap.p, 0, (size_t)1 << (bits - 3), st));
    const uint64_t CH = 32ull << 20;   // records per chunk (512 MB)
    DevBuf chunk; chunk.reserve((size_t)(total < CH ? total : CH) * sizeof(Rec16));
    DevBuf a, b;
    for (int pass = 0; pass < 2; pass++) {
      CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 8, st));
      for (uint64_t i0 = 0; i0 < total; i0 += CH) {
        const uint64_t cn = total - i0 < CH ? total - i0 : CH;
        extract_genome_kmers_range(c, c->genomes, c->prm.genome_gap, i0, 1, cn, chunk.as<Rec16>());
        uint64_t blocks = (cn + 255) / 256, maxb = (uint64_t)c->num_sms * 16;
        if (blocks > maxb) blocks = maxb;
        if (pass == 0) k_part_scan<0><<<(unsigned)blocks, 256, 0, st>>>(chunk.as<Rec16>(), cn, lo, hi, last, bits, c->bitmap.as<uint32_t>(), d_cnt, nullptr);
        else k_part_scan<1><<<(unsigned)blocks, 256, 0, st>>>(chunk.as<Rec16>(), cn, lo, hi, last, bits, nullptr, d_cnt, a.as<Rec16>());
        c->launches++;
        CUDA_TRY(cudaGetLastError());
      }
      if (pass == 0) {
        read_small(c, h_cnt, d_cnt, 8);
        CUDA_TRY(cudaStreamSynchronize(st));
        c->n_gk = h_cnt[0];
        if (!c->n_gk) break;
        a.reserve((size_t)c->n_gk * sizeof(Rec16)); b.reserve((size_t)c->n_gk * sizeof(Rec16));
      }
    }
    chunk.release();
    if (c->n_gk) finish_genome_index(c, a, b, c->n_gk);
    a.release(); b.release();
    c->filter_bits = bits;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  c->tm.n_genome_kmers = c->n_gk;
  c->genomes_loaded = true;
  return KSLAM_OK;
  API_END(c)
}

int kslam_get_partition(const kslam_ctx *c, uint32_t *part, uint32_t *n_parts, uint64_t *splitters, uint64_t *n_genome_kmers_total) {
  if (!c) return KSLAM_ERR_ARG;
  if (!c->genomes_loaded) return KSLAM_ERR_STATE;
  if (part) *part = c->part;
  if (n_parts) *n_parts = c->n_parts;
  if (splitters) for (uint32_t p = 0; p <= c->n_parts; p++) splitters[p] = c->splitters[p];
  if (n_genome_kmers_total) *n_genome_kmers_total = c->n_gk_total;
  return KSLAM_OK;
}

int kslam_part_route_kmers(kslam_ctx *c, uint32_t read_id_base, const void **dev_records, uint64_t *counts) {
  API_BEGIN(c)
  if (!c->reads_loaded) return api_fail(c, KSLAM_ERR_STATE, "kslam_upload_reads first");
  if (!dev_records || !counts) return api_fail(c, KSLAM_ERR_ARG, "null output");
  if ((uint64_t)read_id_base + c->reads.n > (1ull << 30)) return api_fail(c, KSLAM_ERR_ARG, "job-global read ids exceed 2^30 (KMer.h:65-66)");
  c->ev_used = 0; c->aligned = false;
  cudaEvent_t e0 = tm_mark(c);
  uint64_t n = 0;
  c->tm.n_read_kmers = c->reads.n_kmers; c->tm.n_sort_passes = 0;
  if (c->reads.n_kmers && c->filter_bits) {
    n = extract_read_kmers_filtered(c, c->reads, c->recA, read_id_base);
  }
  c->part_send.reserve((size_t)n * sizeof(Rec16) + 64);
  cudaEvent_t em = tm_mark(c);
  bucket_records<0>(c, c->recA.as<Rec16>(), n, c->splitters.data(), c->part_send.as<Rec16>(), counts);
  cudaEvent_t e1 = tm_mark(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->tm.ms_extract = tm_ms(e0, e1);
  c->part_ms_bucket = tm_ms(em, e1);
  c->tm.n_sorted_kmers = n;
  *dev_records = c->part_send.p;
  return KSLAM_OK;
  API_END(c)
}

int kslam_part_recv_buffer(kslam_ctx *c, uint64_t n_records, void **dev_ptr) {
  API_BEGIN(c)
  if (!dev_ptr) return api_fail(c, KSLAM_ERR_ARG, "null output");
  c->part_recv.reserve((size_t)n_records * sizeof(Rec16) + 64);
  c->part_tmp.reserve((size_t)n_records * sizeof(Rec16) + 64);
  *dev_ptr = c->part_recv.p;
  return KSLAM_OK;
  API_END(c)
}

int kslam_part_join(kslam_ctx *c, uint64_t n_records, const uint32_t *id_bases, const void **dev_matches, uint64_t *counts) {
  API_BEGIN(c)
  if (!c->genomes_loaded) return api_fail(c, KSLAM_ERR_STATE, "kslam_load_genomes_part first");
  if (!id_bases || !dev_matches || !counts) return api_fail(c, KSLAM_ERR_ARG, "null argument");
  if (n_records * sizeof(Rec16) > c->part_recv.cap) return api_fail(c, KSLAM_ERR_ARG, "more records than kslam_part_recv_buffer reserved");
  c->ev_used = 0;
  cudaEvent_t e0 = tm_mark(c);
  uint64_t passes = 0;
  const Rec16 *sorted = nullptr;
  if (n_records) sorted = radix_sort(c, c->part_recv.as<Rec16>(), c->part_tmp.as<Rec16>(), n_records, 0, 64 - kmer_sort_bits(c), 64, &passes);
  cudaEvent_t e1 = tm_mark(c);
  const uint64_t n_m = run_join(c, sorted, n_records, true, c->part_m);
  c->part_msend.reserve((size_t)n_m * sizeof(Rec16) + 64);
  uint64_t bounds[DIST_MAX_PARTS];
  for (uint32_t p = 0; p < c->n_parts; p++) bounds[p] = id_bases[p];
  cudaEvent_t em = tm_mark(c);
  bucket_records<1>(c, c->part_m.as<Rec16>(), n_m, bounds, c->part_msend.as<Rec16>(), counts);
  cudaEvent_t e2 = tm_mark(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->tm.ms_sort = tm_ms(e0, e1); c->tm.ms_join = tm_ms(e1, e2);
  c->part_ms_bucket_matches = tm_ms(em, e2);
  c->tm.n_sort_passes = passes;
  *dev_matches = c->part_msend.p;
  return KSLAM_OK;
  API_END(c)
}

int kslam_part_match_buffer(kslam_ctx *c, uint64_t n_matches, void **dev_ptr) {
  API_BEGIN(c)
  if (!dev_ptr) return api_fail(c, KSLAM_ERR_ARG, "null output");
  c->part_mrecv.reserve((size_t)n_matches * sizeof(Rec16) + 64);
  *dev_ptr = c->part_mrecv.p;
  return KSLAM_OK;
  API_END(c)
}

// kslam_part_finish lives in api.cu (it shares the result-fetch code of kslam_align_batch)

}  // extern "C"

// seeds of this GPU's reads from the matches every key owner sent back; leaves them in c->seedA for seed_sort_unique
void part_matches_to_seeds(kslam_ctx *c, uint64_t n_matches, uint32_t read_id_base) {
  matches_to_seeds(c, c->part_mrecv.as<Rec16>(), n_matches, read_id_base);
}
