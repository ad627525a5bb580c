// pair.cu — score screen + R1/R2 pairing (kernel K9).
//
// Reference: screenOverlapsByScoreThreshold (/root/reference/src/Overlap.h:329-341) then getPairedOverlaps
// (/root/reference/src/PairedOverlap.h:243-272): sort by (read % mid, entry, rel) and run the four-slot state
// machine getPairsFromRead (:132-242, makePair :107-123) over every (pair, entry) run.
//
// The alignment array is already ordered by (read, entry, rel) with all R1 reads before all R2 reads, so a
// STABLE radix sort on (pair id, entry, rel) yields the reference's order with R1 first on exact ties
// (SURVEY.md App. C H3). One thread walks one (pair, entry) run twice (count, then emit after a prefix sum).
#include "common.cuh"

__global__ void __launch_bounds__(256)
k_pair_keys(const kslam_overlap *__restrict__ ov, uint32_t n, uint32_t mid, uint32_t thr, uint32_t bias,
            Rec16 *__restrict__ keys, uint32_t *__restrict__ n_pass) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool pass = false;
  if (i < n) {
    const kslam_overlap o = ov[i];
    Rec16 k;
    pass = o.sw_score >= thr;           // Overlap.h:335: removed when sw_score < scoreThreshold
    if (pass) {
      k.key = ((uint64_t)(o.read % mid) << 32) | o.entry;
      k.val = ((uint64_t)(uint32_t)(o.rel + (int32_t)bias) << 32) | i;
    } else { k.key = ~0ull; k.val = ((uint64_t)0xffffffffu << 32) | i; }
    keys[i] = k;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, pass);     // one atomic per warp
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_pass, (uint32_t)__popc(m));
}

__global__ void __launch_bounds__(256)
k_pair_gather(const Rec16 *__restrict__ sorted, uint32_t n_sorted, const kslam_overlap *__restrict__ ov,
              kslam_overlap *__restrict__ ov_out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_sorted) return;
  ov_out[k] = ov[(uint32_t)sorted[k].val];      // cigar_off keeps pointing into the batch's dense CIGAR pool
}

// getPairsFromRead over the run starting at `first`; emits into out (or only counts when out == nullptr)
__device__ uint32_t pair_run(const kslam_overlap *__restrict__ ov, uint32_t first, uint32_t n, uint32_t mid,
                             const uint64_t *__restrict__ read_offs, kslam_pair *out) {
  const uint32_t pid = ov[first].read % mid, entry = ov[first].entry;
  int32_t l1 = -1, l2 = -1, l1rc = -1, l2rc = -1;
  bool u1 = false, u2 = false, u1rc = false, u2rc = false;
  uint32_t cnt = 0;
  auto single = [&](int32_t idx, bool is_r1) {
    if (out) {
      kslam_pair p;
      p.combined_score = ov[idx].sw_score & 0xffffu; p.entry = ov[idx].entry;
      p.ref_start = ov[idx].ref_begin; p.ref_end = ov[idx].ref_end; p.insert_size = 0;
      p.r1_idx = is_r1 ? idx : -1; p.r2_idx = is_r1 ? -1 : idx; p.pad = 0;
      out[cnt] = p;
    }
    cnt++;
  };
  auto both = [&](int32_t r1, int32_t r2, bool orientation) {   // makePair, PairedOverlap.h:107-123
    if (out) {
      kslam_pair p;
      p.combined_score = (ov[r1].sw_score + ov[r2].sw_score) & 0xffffu;   // u16 ctor parameter
      p.entry = ov[r2].entry;
      p.ref_start = ov[r1].ref_begin < ov[r2].ref_begin ? ov[r1].ref_begin : ov[r2].ref_begin;
      p.ref_end = ov[r1].ref_end > ov[r2].ref_end ? ov[r1].ref_end : ov[r2].ref_end;
      const uint32_t ra = orientation ? ov[r2].read : ov[r1].read;
      const uint32_t len = (uint32_t)(read_offs[ra + 1] - read_offs[ra]);
      p.insert_size = orientation ? (uint32_t)(ov[r2].rel - ov[r1].rel) + len : (uint32_t)(ov[r1].rel - ov[r2].rel) + len;
      p.r1_idx = r1; p.r2_idx = r2; p.pad = 0;
      out[cnt] = p;
    }
    cnt++;
  };
  uint32_t cur = first;
  while (cur < n && ov[cur].read % mid == pid && ov[cur].entry == entry) {
    const int32_t c = (int32_t)cur;
    if (ov[cur].read < mid) {
      if (ov[cur].rev_comp) {
        if (!u1rc && l1rc >= 0) single(l1rc, true);
        l1rc = c; u1rc = false;
        if (l2 >= 0) { both(c, l2, false); u1rc = true; u2 = true; }
      } else {
        if (!u1 && l1 >= 0) single(l1, true);
        l1 = c; u1 = false;
        if (l2rc >= 0) { both(c, l2rc, false); u1 = true; u2rc = true; }
      }
    } else {
      if (ov[cur].rev_comp) {
        if (!u2rc && l2rc >= 0) single(l2rc, false);
        l2rc = c; u2rc = false;
        if (l1 >= 0) { both(l1, c, true); u1 = true; u2rc = true; }
      } else {
        if (!u2 && l2 >= 0) single(l2, false);
        l2 = c; u2 = false;
        if (l1rc >= 0) { both(l1rc, c, true); u1rc = true; u2 = true; }
      }
    }
    cur++;
  }
  if (!u2 && l2 >= 0) single(l2, false);          // flush order: PairedOverlap.h:217-240
  if (!u2rc && l2rc >= 0) single(l2rc, false);
  if (!u1 && l1 >= 0) single(l1, true);
  if (!u1rc && l1rc >= 0) single(l1rc, true);
  return cnt;
}

__global__ void __launch_bounds__(256)
k_pair_count(const kslam_overlap *__restrict__ ov, uint32_t n, uint32_t mid, const uint64_t *__restrict__ read_offs,
             uint32_t *__restrict__ cnt) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const bool head = k == 0 || ov[k - 1].read % mid != ov[k].read % mid || ov[k - 1].entry != ov[k].entry;
  cnt[k] = head ? pair_run(ov, k, n, mid, read_offs, nullptr) : 0u;
}

__global__ void __launch_bounds__(256)
k_pair_emit(const kslam_overlap *__restrict__ ov, uint32_t n, uint32_t mid, const uint64_t *__restrict__ read_offs,
            const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ pos, kslam_pair *__restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || cnt[k] == 0) return;
  pair_run(ov, k, n, mid, read_offs, out + pos[k]);
}

void pair_overlaps(kslam_ctx *c) {
  cudaStream_t st = c->stream;
  const uint32_t n = (uint32_t)c->n_seeds;
  c->n_sorted = 0; c->n_pairs = 0;
  if (!n) return;
  const uint32_t mid = (uint32_t)(c->reads.n / 2);
  if (mid == 0) return;
  uint32_t *d_cnt = c->counters.as<uint32_t>() + 48;
  uint32_t *h_cnt = c->h_counters.as<uint32_t>() + 48;
  c->pair_keys.reserve((size_t)n * sizeof(Rec16) + 64);
  c->pair_keys2.reserve((size_t)n * sizeof(Rec16) + 64);
  CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 16, st));
  const unsigned nb = (n + 255) / 256;
  k_pair_keys<<<nb, 256, 0, st>>>(c->ov.as<kslam_overlap>(), n, mid, c->prm.score_threshold, c->reads.max_len,
                                  c->pair_keys.as<Rec16>(), d_cnt);
  c->launches++;
  uint64_t passes = 0;
  Rec16 *a = c->pair_keys.as<Rec16>(), *b = c->pair_keys2.as<Rec16>();
  Rec16 *cur = radix_sort(c, a, b, n, 1, 32, 64, &passes);       // rel (stable: index order kept on ties)
  cur = radix_sort(c, cur, cur == a ? b : a, n, 0, 0, 64, &passes);  // entry, then pair id
  read_small(c, h_cnt, d_cnt, 4);
  CUDA_TRY(cudaStreamSynchronize(st));
  const uint32_t ns = h_cnt[0];
  c->n_sorted = ns;
  if (!ns) return;
  c->ov_sorted.reserve((size_t)ns * sizeof(kslam_overlap) + 64);
  const unsigned nbs = (ns + 255) / 256;
  k_pair_gather<<<nbs, 256, 0, st>>>(cur, ns, c->ov.as<kslam_overlap>(), c->ov_sorted.as<kslam_overlap>());
  c->pair_cnt.reserve((size_t)ns * 8 + 64);
  uint32_t *cnt = c->pair_cnt.as<uint32_t>(), *pos = cnt + ns;
  k_pair_count<<<nbs, 256, 0, st>>>(c->ov_sorted.as<kslam_overlap>(), ns, mid, c->reads.offs.as<uint64_t>(), cnt);
  c->launches += 2;
  unsigned long long *d_tot = c->counters.as<unsigned long long>() + 30;
  exclusive_scan_u32(c, cnt, pos, ns, (uint64_t *)d_tot);
  unsigned long long *h_tot = c->h_counters.as<unsigned long long>() + 30;
  read_small(c, h_tot, d_tot, 8);
  CUDA_TRY(cudaStreamSynchronize(st));
  c->n_pairs = h_tot[0];
  if (c->n_pairs) {
    c->pairs.reserve((size_t)c->n_pairs * sizeof(kslam_pair) + 64);
    k_pair_emit<<<nbs, 256, 0, st>>>(c->ov_sorted.as<kslam_overlap>(), ns, mid, c->reads.offs.as<uint64_t>(), cnt, pos,
                                     c->pairs.as<kslam_pair>());
    c->launches++;
  }
  CUDA_TRY(cudaGetLastError());
}
