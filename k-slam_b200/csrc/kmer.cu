// kmer.cu — canonical 32-mer extraction from 2-bit packed sequences (kernel K1).
//
// Restates /root/reference/src/KMer.h:160-181 (splitIntoKMersAndAddToVector) + :272-280
// (addBaseToKMers) without the rolling loop: the forward k-mer at base p is a 64-bit funnel shift of
// two packed words, its reverse complement is a bit-reversal + pair swap + complement mask, so every
// k-mer position is independent and one lane owns one record. Output is the reference's 16-byte
// record (KMer.h:58-116) written with one 128-bit store per lane, 512 B contiguous per warp.
//
// Algorithmic bytes: 16 B written per record (+0.25 B read per base) — HBM-write bound.
#include "common.cuh"

__device__ __forceinline__ uint64_t revcomp32(uint64_t f) {
  // reverse the order of the 32 two-bit groups, then complement (code ^ 2, KMer.h:279)
  uint64_t r = __brevll(f);                                   // reverses bits: groups reversed AND bit-swapped
  r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
  return r ^ 0xAAAAAAAAAAAAAAAAull;
}

__device__ __forceinline__ void emit_kmer(Rec16 *out, uint64_t f, uint32_t id, uint32_t pos, uint32_t len,
                                          bool is_gb) {
  uint64_t rc = revcomp32(f);
  Rec16 r;
  uint32_t flags = (id & 0x3FFFFFFFu) | (is_gb ? 0x80000000u : 0u);
  if (f < rc) {                       // forward wins only if strictly smaller (KMer.h:173)
    r.key = f; r.val = (uint64_t)flags | ((uint64_t)pos << 32);
  } else {                            // palindromes take the rc branch; read offsets are measured on the
    uint32_t off = is_gb ? pos : len - KSLAM_K - pos;   // reverse strand: len-1-i with i = pos+31 (KMer.h:176)
    r.key = rc; r.val = (uint64_t)(flags | 0x40000000u) | ((uint64_t)off << 32);
  }
  *reinterpret_cast<ulonglong2 *>(out) = make_ulonglong2(r.key, r.val);
}

__device__ __forceinline__ uint64_t kmer_at(const uint64_t *__restrict__ kbits, uint64_t woff, uint32_t p) {
  uint32_t wi = p >> 5, r = p & 31;
  uint64_t w0 = __ldg(&kbits[woff + wi]);
  if (r == 0) return w0;
  uint64_t w1 = __ldg(&kbits[woff + wi + 1]);   // guard words past the end make this always readable
  return (w0 << (2 * r)) | (w1 >> (64 - 2 * r));
}

// Reads (gap 1): one warp per sequence, lanes stride over k-mer positions.
__global__ void __launch_bounds__(256)
k_extract_reads(const uint64_t *__restrict__ kbits, const uint64_t *__restrict__ offs,
                const uint64_t *__restrict__ word_off, const uint64_t *__restrict__ kmer_off,
                uint64_t n_seqs, Rec16 *__restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t seq = warp; seq < n_seqs; seq += nwarps) {
    uint64_t len64 = __ldg(&offs[seq + 1]) - __ldg(&offs[seq]);
    if (len64 < KSLAM_K) continue;              // KMer.h:167
    uint32_t len = (uint32_t)len64;
    uint64_t woff = __ldg(&word_off[seq]), koff = __ldg(&kmer_off[seq]);
    uint32_t nk = len - KSLAM_K + 1;
    for (uint32_t p = lane; p < nk; p += 32)
      emit_kmer(out + koff + p, kmer_at(kbits, woff, p), (uint32_t)seq, p, len, false);
  }
}

// Genomes (gap 16 by default): few long sequences; one thread per record, owner found by binary
// search over kmer_off. Runs once per database load.
__global__ void __launch_bounds__(256)
k_extract_flat(const uint64_t *__restrict__ kbits, const uint64_t *__restrict__ offs,
               const uint64_t *__restrict__ word_off, const uint64_t *__restrict__ kmer_off,
               uint64_t n_seqs, uint64_t n_recs, uint32_t gap, bool is_gb, Rec16 *__restrict__ out,
               uint64_t i_begin, uint64_t i_stride) {
  // output slot t holds flat record i = i_begin + t * i_stride (whole list: 0 / 1; chunks and samples: dist path)
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n_recs;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = i_begin + t * i_stride;
    uint64_t lo = 0, hi = n_seqs;
    while (hi - lo > 1) {
      uint64_t mid = (lo + hi) >> 1;
      if (__ldg(&kmer_off[mid]) <= i) lo = mid; else hi = mid;
    }
    uint64_t seq = lo;
    uint64_t len = __ldg(&offs[seq + 1]) - __ldg(&offs[seq]);
    uint64_t p = (i - __ldg(&kmer_off[seq])) * gap;
    emit_kmer(out + t, kmer_at(kbits, __ldg(&word_off[seq]), (uint32_t)p), (uint32_t)seq, (uint32_t)p,
              (uint32_t)len, is_gb);
  }
}

// ---- prefilter: a bitmap over hashed genome k-mers ---------------------------------------------------
// Only read k-mers that equal some genome k-mer can ever seed (Overlap.h:157: a pile must start with a genome
// record), and zero k-mers never do (Overlap.h:236-239). A read record whose hashed k-mer misses the bitmap of
// genome k-mers therefore cannot contribute to any pile and is dropped before it is ever written to HBM; false
// positives are harmless (the join drops them). The seed multiset is unchanged; the sort shrinks ~10x.

__global__ void __launch_bounds__(256)
k_bitmap_build(const uint64_t *__restrict__ gkeys, uint64_t n, uint32_t bits, uint32_t *__restrict__ bitmap) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t k = __ldg(&gkeys[i]);
    if (k == 0) continue;
    uint64_t h = kmer_hash(k, bits);
    atomicOr(&bitmap[h >> 5], 1u << (h & 31));
  }
}

#define XF_WBUF 160   // per-warp staging records (flushed with one global atomic once > 128 are pending)

// Reads (gap 1) with the prefilter fused in: one warp per read; survivors are staged per warp in shared memory
// and appended to the output in contiguous bursts. Record order is arbitrary (they are sorted next).
__global__ void __launch_bounds__(256)
k_extract_reads_filtered(const uint64_t *__restrict__ kbits, const uint64_t *__restrict__ offs,
                         const uint64_t *__restrict__ word_off, uint64_t n_seqs, const uint32_t *__restrict__ bitmap,
                         uint32_t bits, Rec16 *__restrict__ out, unsigned long long *__restrict__ counter,
                         uint32_t id_base, uint64_t cap) {
  __shared__ __align__(16) Rec16 s_buf[8][XF_WBUF];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  Rec16 *buf = s_buf[wib];
  uint32_t pending = 0;   // warp-uniform
  for (uint64_t seq = warp; seq < n_seqs; seq += nwarps) {
    uint64_t len64 = __ldg(&offs[seq + 1]) - __ldg(&offs[seq]);
    if (len64 < KSLAM_K) continue;
    uint32_t len = (uint32_t)len64;
    uint64_t woff = __ldg(&word_off[seq]);
    uint32_t nk = len - KSLAM_K + 1;
    // four rounds of 32 positions at a time: the four bitmap probes of a lane are independent loads in flight together
    // (one round at a time left every probe waiting out its own L2 / HBM latency: 26 ms per config-2 batch)
    for (uint32_t p0 = 0; p0 < nk; p0 += 128) {
      uint64_t key[4], val[4];
      uint32_t word[4], bit[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t p = p0 + 32 * j + lane;
        key[j] = 0; val[j] = 0; word[j] = 0; bit[j] = 0;
        if (p < nk) {
          uint64_t f = kmer_at(kbits, woff, p), rc = revcomp32(f);
          uint32_t flags = ((uint32_t)seq + id_base) & 0x3FFFFFFFu;   // id_base != 0: job-global read ids (dist.cu)
          if (f < rc) { key[j] = f; val[j] = (uint64_t)flags | ((uint64_t)p << 32); }
          else { key[j] = rc; val[j] = (uint64_t)(flags | 0x40000000u) | ((uint64_t)(len - KSLAM_K - p) << 32); }
          if (key[j] != 0) { const uint64_t h = kmer_hash(key[j], bits); word[j] = __ldg(&bitmap[h >> 5]); bit[j] = (uint32_t)h & 31u; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (p0 + 32 * j >= nk) break;                                  // warp-uniform
        const bool keep = (word[j] >> bit[j]) & 1u;                    // (zero k-mers and lanes past the end probe nothing: word 0)
        uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) *reinterpret_cast<ulonglong2 *>(&buf[pending + __popc(m & ((1u << lane) - 1))]) = make_ulonglong2(key[j], val[j]);
        pending += __popc(m);
        if (pending > XF_WBUF - 32) {
          __syncwarp();
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(counter, (unsigned long long)pending);
          base = __shfl_sync(0xffffffffu, base, 0);
          if (base + pending <= cap)       // overflow: the counter still ends at the exact total and the host re-runs
            for (uint32_t i = lane; i < pending; i += 32)
              *reinterpret_cast<ulonglong2 *>(out + base + i) = *reinterpret_cast<const ulonglong2 *>(&buf[i]);
          __syncwarp();
          pending = 0;
        }
      }
    }
  }
  if (pending) {
    __syncwarp();
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)pending);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + pending <= cap)
      for (uint32_t i = lane; i < pending; i += 32)
        *reinterpret_cast<ulonglong2 *>(out + base + i) = *reinterpret_cast<const ulonglong2 *>(&buf[i]);
  }
}

void build_prefilter(kslam_ctx *c) {
  c->filter_bits = 0;
  if (!c->n_gk) return;
  uint32_t bits = prefilter_bits(c->n_gk);
  c->bitmap.reserve((size_t)1 << (bits - 3));
  CUDA_TRY(cudaMemsetAsync(c->bitmap.p, 0, (size_t)1 << (bits - 3), c->stream));
  uint64_t blocks = (c->n_gk + 255) / 256, maxb = (uint64_t)c->num_sms * 16;
  if (blocks > maxb) blocks = maxb;
  k_bitmap_build<<<(unsigned)blocks, 256, 0, c->stream>>>(c->g_keys.as<uint64_t>(), c->n_gk, bits, c->bitmap.as<uint32_t>());
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  c->filter_bits = bits;
}

// Extracts the read k-mers that pass the prefilter into `outbuf` and returns their number. The buffer starts at an
// eighth of the worst case (the filter keeps ~7 % on unrelated data) and is regrown to the exact size — the kernel's
// counter keeps counting past the capacity — if a batch needs more.
uint64_t extract_read_kmers_filtered(kslam_ctx *c, const PackedSeqs &s, DevBuf &outbuf, uint32_t id_base) {
  if (!s.n_kmers) return 0;
  unsigned long long *d_cnt = c->counters.as<unsigned long long>() + 2;
  unsigned long long *h_cnt = c->h_counters.as<unsigned long long>() + 2;
  uint64_t want = s.n_kmers / 8 + (1u << 20);
  if (want > s.n_kmers) want = s.n_kmers;
  if (outbuf.cap < want * sizeof(Rec16)) outbuf.reserve((size_t)want * sizeof(Rec16));
  uint64_t blocks = (s.n * 32 + 255) / 256, maxb = (uint64_t)c->num_sms * 8;
  if (blocks > maxb) blocks = maxb;
  for (int attempt = 0; attempt < 2; attempt++) {
    const uint64_t cap = outbuf.cap / sizeof(Rec16);
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 8, c->stream));
    k_extract_reads_filtered<<<(unsigned)blocks, 256, 0, c->stream>>>(s.kbits.as<uint64_t>(), s.offs.as<uint64_t>(),
        s.word_off.as<uint64_t>(), s.n, c->bitmap.as<uint32_t>(), c->filter_bits, outbuf.as<Rec16>(), d_cnt, id_base, cap);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    read_small(c, h_cnt, d_cnt, 8);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h_cnt[0] <= cap) break;
    if (attempt == 1) throw CudaError{cudaErrorMemoryAllocation, "read k-mer buffer overflow after regrow", __FILE__, __LINE__};
    outbuf.reserve((size_t)h_cnt[0] * sizeof(Rec16));
  }
  return h_cnt[0];
}

void extract_kmers(kslam_ctx *c, const PackedSeqs &s, bool is_gb, uint32_t gap, Rec16 *out) {
  if (!s.n_kmers) return;
  if (!is_gb && gap == 1) {
    uint64_t warps = s.n;
    uint64_t blocks = (warps * 32 + 255) / 256;
    uint64_t maxb = (uint64_t)c->num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    k_extract_reads<<<(unsigned)blocks, 256, 0, c->stream>>>(s.kbits.as<uint64_t>(), s.offs.as<uint64_t>(),
                                                             s.word_off.as<uint64_t>(), s.kmer_off.as<uint64_t>(),
                                                             s.n, out);
  } else {
    uint64_t blocks = (s.n_kmers + 255) / 256;
    uint64_t maxb = (uint64_t)c->num_sms * 16;
    if (blocks > maxb) blocks = maxb;
    k_extract_flat<<<(unsigned)blocks, 256, 0, c->stream>>>(s.kbits.as<uint64_t>(), s.offs.as<uint64_t>(),
                                                            s.word_off.as<uint64_t>(), s.kmer_off.as<uint64_t>(),
                                                            s.n, s.n_kmers, gap, is_gb, out, 0, 1);
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

// n_out genome records starting at flat record i_begin, every i_stride-th one (chunked index build and splitter
// sampling of the k-mer-range partitioned database, dist.cu)
void extract_genome_kmers_range(kslam_ctx *c, const PackedSeqs &s, uint32_t gap, uint64_t i_begin, uint64_t i_stride,
                                uint64_t n_out, Rec16 *out) {
  if (!n_out) return;
  uint64_t blocks = (n_out + 255) / 256, maxb = (uint64_t)c->num_sms * 16;
  if (blocks > maxb) blocks = maxb;
  k_extract_flat<<<(unsigned)blocks, 256, 0, c->stream>>>(s.kbits.as<uint64_t>(), s.offs.as<uint64_t>(), s.word_off.as<uint64_t>(),
                                                          s.kmer_off.as<uint64_t>(), s.n, n_out, gap, true, out, i_begin, i_stride);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}
