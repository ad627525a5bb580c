// scan.cu — device-wide exclusive prefix sum (u32 in, u32 out, u64 total), three launches:
// per-tile reduce, single-block scan of tile sums, per-tile downsweep. Used for stream compaction
// (seed unique, pairing). Warp-shuffle scans inside a tile; HBM traffic 2 reads + 1 write per item.
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total, uint32_t *s_warp /*[32]*/) {
  uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v);
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t nw = blockDim.x >> 5;
    uint32_t w = lane < nw ? s_warp[lane] : 0;
    uint32_t winc = warp_incl_scan(w);
    s_warp[lane] = winc - w;
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  uint32_t r = s_warp[wid] + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_total;
  uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint64_t idx = base + (uint64_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) sum += in[idx];
  }
  block_excl_scan(sum, &s_total, s_warp);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(1024)
k_scan_tiles(uint32_t *__restrict__ tile_sums, uint64_t n_tiles, uint64_t *__restrict__ total_out) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_total;
  uint64_t carry = 0;   // tile offsets are kept in u32 (callers guarantee totals < 2^32); total is u64
  for (uint64_t base = 0; base < n_tiles; base += 1024) {
    uint64_t idx = base + threadIdx.x;
    uint32_t v = idx < n_tiles ? tile_sums[idx] : 0;
    uint32_t ex = block_excl_scan(v, &s_total, s_warp);
    if (idx < n_tiles) tile_sums[idx] = (uint32_t)(carry + ex);
    carry += s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_down(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n,
            const uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_total;
  // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
  uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = base + i < n ? in[base + i] : 0; sum += v[i]; }
  uint32_t ex = block_excl_scan(sum, &s_total, s_warp) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
}

void exclusive_scan_u32(kslam_ctx *c, const uint32_t *in, uint32_t *out, uint64_t n, uint64_t *total_dev) {
  if (n == 0) { if (total_dev) CUDA_TRY(cudaMemsetAsync(total_dev, 0, 8, c->stream)); return; }
  uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  c->scan_tmp.reserve(tiles * 4 + 64);
  uint32_t *ts = c->scan_tmp.as<uint32_t>();
  k_scan_reduce<<<(unsigned)tiles, SCAN_THREADS, 0, c->stream>>>(in, n, ts);
  k_scan_tiles<<<1, 1024, 0, c->stream>>>(ts, tiles, total_dev);
  k_scan_down<<<(unsigned)tiles, SCAN_THREADS, 0, c->stream>>>(in, out, n, ts);
  c->launches += 3;
  CUDA_TRY(cudaGetLastError());
}
