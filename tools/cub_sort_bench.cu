// tools/cub_sort_bench.cu — NOT part of the product. Times CUB's library radix sort (u64 key + u64 value,
// the same 16 bytes per record) on the same sizes as tools/prof_sort.py, as an outside reference point for
// the hand-written sort in k-slam_b200/csrc/radix_sort.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cub_sort_bench tools/cub_sort_bench.cu
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
int main(int argc, char **argv) {
  size_t n = argc > 1 ? atoll(argv[1]) : 64000000;
  std::vector<unsigned long long> h(n);
  std::mt19937_64 rng(1);
  for (auto &x : h) x = rng();
  unsigned long long *k0, *k1, *v0, *v1;
  cudaMalloc(&k0, n * 8); cudaMalloc(&k1, n * 8); cudaMalloc(&v0, n * 8); cudaMalloc(&v1, n * 8);
  cudaMemcpy(k0, h.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemset(v0, 0, n * 8);
  void *tmp = nullptr; size_t tmp_bytes = 0;
  cub::DoubleBuffer<unsigned long long> dk(k0, k1), dv(v0, v1);
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)n);
  cudaMalloc(&tmp, tmp_bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 4; rep++) {
    cudaMemcpy(k0, h.data(), n * 8, cudaMemcpyHostToDevice);
    cub::DoubleBuffer<unsigned long long> k(k0, k1), v(v0, v1);
    cudaEventRecord(a);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (int)n);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("cub SortPairs u64+u64 n=%zu: %.3f ms -> %.1f GB/s moved (8 passes x 32 B), %.1f GB/s algorithmic\n", n, ms,
           32.0 * n * 8 / ms / 1e6, 32.0 * n / ms / 1e6);
  }
  return 0;
}
