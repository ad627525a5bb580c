// pack.cu — raw sequence bytes -> bit planes in HBM.
//
// Three alphabets exist on this path and packing must not lose any of them:
//   k-mer codec  /root/reference/src/KMer.h:246-266      A0 C1 T2 G3, every other byte 0
//   SSW codec    /root/reference/src/ssw_cpp.cpp:11-23   A/a0 C/c1 G/g2 T/t3 U/u0, every other byte 4
//   complement  /root/reference/src/sequenceTools.h:98-116 only swaps UPPER-CASE A/C/G/T (window rev-comp)
// so each 32-base word is stored as: kbits (u64, k-mer codes, FIRST base in the MOST significant
// bit pair so a 32-mer is a funnel shift of two words), sbits (u64, SSW codes 0-3, base b at bits
// 2b..2b+1), nmask (u32, bit b set where the SSW code is 4) and xmask (u32, bit b set where the byte
// is not upper-case ACGT but still has an SSW code < 4: a c g t U u — such bases are NOT complemented).
#include "common.cuh"

__constant__ uint8_t c_kcode[256];  // k-mer alphabet
__constant__ uint8_t c_scode[256];  // SSW alphabet (0..4) | 8 when the byte is a/c/g/t/U/u

static bool g_tables_ready[64] = {false};

static void ensure_tables(int device) {
  if (device < 64 && g_tables_ready[device]) return;
  uint8_t k[256], s[256];
  for (int i = 0; i < 256; i++) { k[i] = 0; s[i] = 4; }
  k['A'] = 0; k['C'] = 1; k['T'] = 2; k['G'] = 3;
  s['A'] = s['a'] = s['U'] = s['u'] = 0;
  s['C'] = s['c'] = 1; s['G'] = s['g'] = 2; s['T'] = s['t'] = 3;
  s['a'] |= 8; s['c'] |= 8; s['g'] |= 8; s['t'] |= 8; s['U'] |= 8; s['u'] |= 8;
  CUDA_TRY(cudaMemcpyToSymbol(c_kcode, k, 256));
  CUDA_TRY(cudaMemcpyToSymbol(c_scode, s, 256));
  if (device < 64) g_tables_ready[device] = true;
}

// One thread per 32-base word. The sequence owning a word is found by binary search in word_off
// (L2-resident); bytes are read through the read-only path (a warp covers 1 KB of contiguous input).
__global__ void __launch_bounds__(256)
k_pack(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ offs,
       const uint64_t *__restrict__ word_off, uint64_t n_seqs, uint64_t n_words,
       uint64_t *__restrict__ kbits, uint64_t *__restrict__ sbits, uint32_t *__restrict__ nmask,
       uint32_t *__restrict__ xmask) {
  __shared__ uint8_t s_k[256], s_s[256];
  s_k[threadIdx.x] = c_kcode[threadIdx.x];
  s_s[threadIdx.x] = c_scode[threadIdx.x];
  __syncthreads();
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
       w += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t lo = 0, hi = n_seqs;  // last seq with word_off[seq] <= w  (empty sequences own no word)
    while (hi - lo > 1) {
      uint64_t mid = (lo + hi) >> 1;
      if (__ldg(&word_off[mid]) <= w) lo = mid; else hi = mid;
    }
    uint64_t seq = lo;
    uint64_t base0 = __ldg(&offs[seq]) + (w - __ldg(&word_off[seq])) * 32;
    uint64_t end = __ldg(&offs[seq + 1]);
    uint32_t cnt = (uint32_t)((end - base0) < 32 ? (end - base0) : 32);
    uint64_t kb = 0, sb = 0; uint32_t nm = 0, xm = 0;
#pragma unroll 8
    for (uint32_t b = 0; b < 32; b++) {
      if (b < cnt) {
        uint8_t ch = __ldg(&raw[base0 + b]);
        uint32_t sc = s_s[ch];
        kb |= (uint64_t)s_k[ch] << (62 - 2 * b);
        sb |= (uint64_t)(sc & 3) << (2 * b);
        nm |= ((sc >> 2) & 1) << b;
        xm |= (sc >> 3) << b;
      }
    }
    kbits[w] = kb; sbits[w] = sb; nmask[w] = nm; xmask[w] = xm;
  }
}

// offsets of a batch in which every sequence has the same length: generated on the device, nothing to upload
__global__ void __launch_bounds__(256)
k_uniform_offsets(uint64_t n, uint64_t len, uint64_t wpl, uint64_t kpl, uint64_t *__restrict__ offs,
                  uint64_t *__restrict__ word_off, uint64_t *__restrict__ kmer_off) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
    offs[i] = i * len; word_off[i] = i * wpl; kmer_off[i] = i * kpl;
  }
}

// has_n[seq] = 1 when any nmask word of the sequence is non-zero: one warp per sequence (genomes: few, long)
__global__ void __launch_bounds__(256)
k_seq_has_n(const uint32_t *__restrict__ nmask, const uint64_t *__restrict__ word_off, uint64_t n, uint8_t *__restrict__ has_n) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t seq = warp; seq < n; seq += nwarps) {
    uint32_t any = 0;
    for (uint64_t w = word_off[seq] + lane; w < word_off[seq + 1]; w += 32) any |= nmask[w];
    any = __reduce_or_sync(0xffffffffu, any);
    if (lane == 0) has_n[seq] = any ? 1 : 0;
  }
}

void pack_sequences(kslam_ctx *c, PackedSeqs &s, uint64_t n, const char *bases, const uint64_t *offs,
                    uint32_t kmer_gap, bool keep_raw) {
  ensure_tables(c->device);
  cudaStream_t st = c->stream;
  s.n = n;
  s.n_bases = offs[n] - offs[0];
  // one pass over the offsets: uniform-length batches (the usual FASTQ case) need no per-sequence tables from the host
  const uint64_t len0 = n ? offs[1] - offs[0] : 0;
  bool uniform = n > 0;
  for (uint64_t i = 0; i < n && uniform; i++) uniform = (offs[i + 1] - offs[i]) == len0;
  s.offs.reserve((n + 1) * 8); s.word_off.reserve((n + 1) * 8); s.kmer_off.reserve((n + 1) * 8);
  uint64_t w = 0, k = 0; uint32_t max_len = 0;
  if (uniform) {
    const uint64_t wpl = (len0 + 31) / 32, kpl = len0 >= KSLAM_K ? (len0 - KSLAM_K) / kmer_gap + 1 : 0;
    w = wpl * n; k = kpl * n; max_len = (uint32_t)len0;
    uint64_t blocks = (n + 256) / 256, maxb = (uint64_t)c->num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    k_uniform_offsets<<<(unsigned)blocks, 256, 0, st>>>(n, len0, wpl, kpl, s.offs.as<uint64_t>(), s.word_off.as<uint64_t>(),
                                                        s.kmer_off.as<uint64_t>());
    c->launches++;
  } else {
    // ragged batch: build the three tables in pinned staging memory owned by the ctx, then copy asynchronously
    c->h_stage.reserve((n + 1) * 24);
    uint64_t *ro = c->h_stage.as<uint64_t>(), *wo = ro + (n + 1), *ko = wo + (n + 1);
    for (uint64_t i = 0; i < n; i++) {
      uint64_t len = offs[i + 1] - offs[i];
      ro[i] = offs[i] - offs[0]; wo[i] = w; ko[i] = k;
      w += (len + 31) / 32;
      if (len >= KSLAM_K) k += (len - KSLAM_K) / kmer_gap + 1;  // KMer.h:202
      if (len > max_len) max_len = (uint32_t)(len > 0xffffffffull ? 0xffffffffull : len);
    }
    ro[n] = offs[n] - offs[0]; wo[n] = w; ko[n] = k;
    CUDA_TRY(cudaMemcpyAsync(s.offs.p, ro, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s.word_off.p, wo, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s.kmer_off.p, ko, (n + 1) * 8, cudaMemcpyHostToDevice, st));
  }
  s.n_words = w; s.n_kmers = k; s.max_len = max_len;
  s.raw.reserve(s.n_bases + 16);
  s.kbits.reserve((w + 2) * 8); s.sbits.reserve((w + 2) * 8); s.nmask.reserve((w + 2) * 4);
  s.xmask.reserve((w + 2) * 4);
  if (s.n_bases) CUDA_TRY(cudaMemcpyAsync(s.raw.p, bases + offs[0], s.n_bases, cudaMemcpyDefault, st));   // host or (UVA) device source
  // two guard words past the end so funnel shifts / window gathers may read one word ahead
  CUDA_TRY(cudaMemsetAsync((char *)s.kbits.p + w * 8, 0, 16, st));
  CUDA_TRY(cudaMemsetAsync((char *)s.sbits.p + w * 8, 0, 16, st));
  CUDA_TRY(cudaMemsetAsync((char *)s.nmask.p + w * 4, 0, 8, st));
  CUDA_TRY(cudaMemsetAsync((char *)s.xmask.p + w * 4, 0, 8, st));
  if (w) {
    uint64_t blocks = (w + 255) / 256;
    uint64_t maxb = (uint64_t)c->num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    k_pack<<<(unsigned)blocks, 256, 0, st>>>(s.raw.as<uint8_t>(), s.offs.as<uint64_t>(),
                                             s.word_off.as<uint64_t>(), n, w, s.kbits.as<uint64_t>(),
                                             s.sbits.as<uint64_t>(), s.nmask.as<uint32_t>(), s.xmask.as<uint32_t>());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  if (!keep_raw && n) {      // genomes: which of them hold code-4 bases at all (their windows need no scan for them otherwise)
    s.has_n.reserve(n + 64);
    uint64_t blocks = (n * 32 + 255) / 256, maxb = (uint64_t)c->num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    k_seq_has_n<<<(unsigned)blocks, 256, 0, st>>>(s.nmask.as<uint32_t>(), s.word_off.as<uint64_t>(), n, s.has_n.as<uint8_t>());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  // the caller's buffers and the pinned staging tables must stay untouched until the copies have landed
  CUDA_TRY(cudaStreamSynchronize(st));
  if (!keep_raw) s.raw.release();   // genomes: bytes are not needed again; reads: keep the allocation for the next batch
}
