// sw_band.cuh — banded forward / reverse Smith-Waterman sweep, exact by construction (included by sw.cu).
//
// Why a band can be EXACT here. An alignment that scores S needs at least a = ceil(S / match) diagonal steps.
// A path that touches matrix offset d = j - i consumes d extra columns (d >= 0) or -d extra rows, so inside an
// m x n matrix it has at most n - d (resp. m + d) diagonal steps. Hence EVERY alignment scoring >= S lies inside
// the offsets [-(m - a), n - a].
//   MODE 0 (forward, nothing known): sweep a W-diagonal band placed on that interval's centre, take the best score
//     S' found in it (a true alignment score, so S' <= S); if [-(m - a'), n - a'] fits in the band then all
//     alignments scoring >= S' — every optimal one, hence every cell SSW's tie rules (ssw.c:316-342) look at — are
//     inside the band and were computed with their true H: exact. Otherwise S' is kept as a lower bound and the
//     alignment moves to the next tier (a 64-wide band placed exactly on [-(m - a'), ...] if that is <= 64 wide,
//     else the full-matrix kernel k_sw_fast).
//   MODE 2 (forward, lower bound S' known): band placed at c0 = -(m - a'); it contains every alignment scoring
//     >= S', so the sweep is exact without a further check (the check is still evaluated; a failure would fall back).
//   MODE 1 (reverse pass, ssw.c:905-923): S is known, the band is exactly [-(rows - a), cols - a].
//
// Mapping: ONE THREAD owns two alignments (halves of s16x2 registers) and keeps the band's previous row (H), the
// vertical-gap state (V) and W PRMT column selectors in registers; it walks the rows, W cells per row fully
// unrolled, 7 DPX/PRMT ops per cell pair, no shuffles and no shared memory. Column selectors slide by one
// register per row (moves go to the FMA pipe; the ALU pipe only sees the recurrences). Band cells that fall
// outside the matrix use "replicate-sign" selectors that can only produce 0 or -1, so they stay at H = 0 on the
// left and can never reach a maximum on the right. Per-alignment inputs (column selectors and query codes) come
// from a byte plane written by k_band_bytes, one coalesced 16-bit load per row and pair.
#pragma once

#define SWB_MAXROWS 160
#define SWB_BLOCK 128
#define SWB_MAXW 64
#define SWB_PLANE_ROWS (SWB_MAXROWS + SWB_MAXW)   // band indices k of one plane (multiple of 32)

__device__ __forceinline__ int32_t ceil_div_pos(int32_t a, int32_t b) { return (a + b - 1) / b; }

// geometry of one alignment inside the band sweep
struct BandGeo { int32_t rows, cols, c0; };

template <int MODE, int W>
__device__ __forceinline__ BandGeo band_geo(const SwTask &t, const SwRes &r, const SwScore &sc) {
  BandGeo g;
  if (MODE == 1) {
    g.rows = r.read_end + 1; g.cols = r.ref_end + 1;
    g.c0 = -(g.rows - ceil_div_pos(r.score, sc.match));
  } else {
    g.rows = (int32_t)t.m; g.cols = (int32_t)t.n;
    if (MODE == 0) g.c0 = ((g.cols - g.rows) >> 1) - W / 2;   // centre of [-(m - a), n - a] is (n - m) / 2 whatever a is
    else g.c0 = -(g.rows - ceil_div_pos(r.score, sc.match));   // r.score = proven lower bound from the 32-wide sweep
  }
  return g;
}

// Band byte plane: bytes[k * stride + slot] for band index k in [0, rows + W) of the alignment in list slot `slot`.
// low nibble = selector of matrix column j = k + c0 (SSW code 0-3, or 8 = outside the matrix), high nibble = SSW
// code of query row k (0-4, 5 = past the query). One thread fills 32 consecutive k of one alignment, so the packed
// words are fetched once and consecutive lanes (slots) write consecutive bytes. gridDim.y = (SWB_MAXROWS + W) / 32.
template <int MODE, int W>
__global__ void __launch_bounds__(256)
k_band_bytes(const SwTask *__restrict__ tasks, const uint32_t *__restrict__ list, uint32_t n_list, SwPlanes pl, SwScore sc,
             const SwRes *__restrict__ res, uint8_t *__restrict__ bytes, uint32_t stride) {
  constexpr bool REVERSE = MODE == 1;
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_list) return;
  const int32_t k0 = (int32_t)blockIdx.y * 32;
  const uint32_t idx = list[slot];
  const SwTask t = tasks[idx];
  SwRes r; if (MODE != 0) r = res[idx];
  const BandGeo g = band_geo<MODE, W>(t, r, sc);
  const bool rev = (t.flags & SWT_REV) != 0;
  // genome position of matrix column j is P0 + sg * j (window reversal and the reverse sweep both flip the sign)
  int32_t P0, sg;
  if (!REVERSE) { P0 = rev ? (int32_t)(t.w_start + t.n - 1) : (int32_t)t.w_start; sg = rev ? -1 : 1; }
  else { P0 = rev ? (int32_t)(t.w_start + t.n - 1) - r.ref_end : (int32_t)t.w_start + r.ref_end; sg = rev ? 1 : -1; }
  const uint64_t *wb = pl.w_sbits + t.w_word; const uint32_t *wx = pl.w_xmask + t.w_word;
  const uint64_t *qb = pl.q_sbits + t.q_word; const uint32_t *qn = pl.q_nmask + t.q_word;
  int32_t curw = -1, curq = -1; uint64_t bits = 0, qbits = 0; uint32_t xm = 0, qnm = 0;
  uint8_t *dst = bytes + (size_t)k0 * stride + slot;
#pragma unroll 4
  for (int32_t kk = 0; kk < 32; kk++) {
    const int32_t k = k0 + kk, j = k + g.c0;
    uint32_t lo = 8u;
    if (j >= 0 && j < g.cols) {
      const int32_t pos = P0 + sg * j, wi = pos >> 5;
      if (wi != curw) { curw = wi; bits = __ldg(wb + wi); xm = rev ? ~__ldg(wx + wi) : 0u; }
      lo = (uint32_t)(bits >> (2 * (pos & 31))) & 3u;
      if ((xm >> (pos & 31)) & 1u) lo ^= 3u;               // complement (3 - c) unless the base is a/c/g/t/U/u
    }
    uint32_t hi = 5u;
    if (k < g.rows) {
      const int32_t qi = REVERSE ? g.rows - 1 - k : k, wq = qi >> 5;
      if (wq != curq) { curq = wq; qbits = __ldg(qb + wq); qnm = __ldg(qn + wq); }
      hi = ((qnm >> (qi & 31)) & 1u) ? 4u : ((uint32_t)(qbits >> (2 * (qi & 31))) & 3u);
    }
    dst[(size_t)kk * stride] = (uint8_t)(lo | (hi << 4));
  }
}

// `list`/`bytes` cover one chunk of a band list; slot pairs (2p, 2p+1) share a thread. stride is even, so the two
// bytes of a pair are one aligned 16-bit load. Failures go to next_list (the 64-wide tier; score lower bound left in
// res[].score) when the interval they need is <= SWB_MAXW wide, else to fb_keys (full-matrix kernel, key = columns).
template <int MODE, int W>
__global__ void __launch_bounds__(SWB_BLOCK, (W > 32 ? 2 : (W == 32 ? 3 : 6)))
k_sw_band(const SwTask *__restrict__ tasks, const uint32_t *__restrict__ list, uint32_t n_list, SwScore sc,
          SwRes *__restrict__ res, const uint8_t *__restrict__ bytes, uint32_t stride, Rec16 *__restrict__ fb_keys,
          uint32_t *__restrict__ fb_count, uint32_t *__restrict__ next_list, uint32_t *__restrict__ next_count) {
  constexpr bool REVERSE = MODE == 1;
  const uint32_t p = blockIdx.x * SWB_BLOCK + threadIdx.x;
  if (2 * p >= n_list) return;
  const bool single = 2 * p + 1 >= n_list;
  const uint32_t ia = list[2 * p], ib = single ? ia : list[2 * p + 1];
  const SwTask ta = tasks[ia], tb = tasks[ib];
  SwRes ra, rb;
  if (MODE != 0) { ra = res[ia]; rb = res[ib]; }
  const BandGeo ga = band_geo<MODE, W>(ta, ra, sc), gb = band_geo<MODE, W>(tb, rb, sc);
  const int32_t rows[2] = {ga.rows, gb.rows}, cols[2] = {ga.cols, gb.cols}, c0[2] = {ga.c0, gb.c0};
  const int32_t rows_max = rows[0] > rows[1] ? rows[0] : rows[1];
  const uint8_t *bp = bytes + 2 * (size_t)p;
  // low byte = alignment A, high byte = alignment B (for an unpaired last slot the B half computes on whatever the
  // neighbour byte holds and its result is dropped)
  auto ld2 = [&](int32_t k) -> uint32_t { return __ldg(reinterpret_cast<const uint16_t *>(bp + (size_t)k * stride)); };

  const uint32_t mis_b = (uint32_t)(-(sc.mismatch * 32)) & 0xffu, mat_b = (uint32_t)(sc.match * 32) & 0xffu;
  const uint32_t MIS4 = mis_b * 0x01010101u, DIFF = mis_b ^ mat_b;
  const uint32_t NEG_GO = pack2(-sc.gap_open * 32), NEG_GE = pack2(-sc.gap_extend * 32), MIN2 = 0x80008000u;

  // selector halves: alignment A copies byte w of PA (sign-replicated into the high byte), B byte w of PB;
  // outside the matrix both nibbles replicate a sign (score 0 or -1, never positive)
  auto mk_sel = [](uint32_t ab) -> uint32_t {       // ab = byte A | byte B << 8
    const uint32_t a = ab & 15u, b = (ab >> 8) & 15u;
    return (a | ((a & 3u) << 4) | 0x80u) | ((b | ((b & 3u) << 4) | 0xC4u) << 8);
  };
  uint32_t H[W], V[W], sel[W];
#pragma unroll
  for (int t = 0; t < W; t++) {
    H[t] = 0; V[t] = 0;
    sel[t] = mk_sel(ld2(t));
  }
  uint32_t rowbytes = ld2(0);                                // query codes of row 0 (prefetched one row ahead)
  // forward: key = score << 12 | (4095 - column); reverse: key = 4095 - scan column of the first hit (0 = none)
  uint32_t bestA = 0, bestB = 0, rowA = 0, rowB = 0;
  const uint32_t thrA = REVERSE ? (uint32_t)ra.score * 32u : 0u, thrB = REVERSE ? (uint32_t)rb.score * 32u : 0u;
  // tracking keys carry (31 - slot) in the low bits; scores are scaled by 32 (5 free bits) so for W = 64 the key is
  // kept per 32-slot half and the halves are compared explicitly
  for (int32_t i = 0; i < rows_max; i++) {
    // row profiles [s(q,A) s(q,C) s(q,G) s(q,T)] x 32; code-4 rows score 0; rows past the query mismatch everything
    const uint32_t qa = (rowbytes >> 4) & 15u, qb = rowbytes >> 12;
    const uint32_t PA = qa == 4u ? 0u : (qa == 5u ? MIS4 : (MIS4 ^ (DIFF << (8 * qa))));
    const uint32_t PB = qb == 4u ? 0u : (qb == 5u ? MIS4 : (MIS4 ^ (DIFF << (8 * qb))));
    // issue next row's loads now; they complete under the cell body (indices stay below rows_max + W)
    const uint32_t next = ld2(i + 1), ent = ld2(i + W);
    constexpr int NH = W > 32 ? W / 32 : 1;      // tracking keys per 32-slot half
    uint32_t e = 0, acc[NH];
#pragma unroll
    for (int h = 0; h < NH; h++) acc[h] = 0;
#pragma unroll
    for (int t = 0; t < W; t++) {
      const uint32_t s = prmt(PA, PB, sel[t]);
      const uint32_t v = V[t];
      uint32_t h = __viaddmax_s16x2_relu(H[t], s, v);           // max(H[i-1][j-1] + s, vertical gap, 0)
      h = __vimax3_s16x2(h, e, e);                                // ... and the horizontal gap
      H[t] = h;
      const uint32_t hgo = __viaddmax_s16x2(h, NEG_GO, MIN2);
      e = __viaddmax_s16x2(e, NEG_GE, hgo);                      // horizontal gap into (i, j+1)
      if (t > 0) V[t - 1] = __viaddmax_s16x2(v, NEG_GE, hgo);    // vertical gap into (i+1, j): band slot t-1 next row
      acc[t / 32] = __viaddmax_s16x2(h, (uint32_t)(31 - (t & 31)) * 0x10001u, acc[t / 32]);   // H*32 + (31 - slot): smallest column wins ties
    }
    // slide the column selectors: next row's slot t is this row's slot t+1; the last slot takes the entering column
#pragma unroll
    for (int t = 0; t < W - 1; t++) sel[t] = sel[t + 1];
    sel[W - 1] = mk_sel(ent);
    rowbytes = next;
    // row winners -> running best (first column, then smallest row: rows only grow, so ties keep the old one)
#pragma unroll
    for (int h = 0; h < NH; h++) {
      const uint32_t aA = acc[h] & 0xffffu, aB = acc[h] >> 16;
      const uint32_t jA = (uint32_t)(i + 32 * h + (int32_t)(31u - (aA & 31u)) + c0[0]);
      const uint32_t jB = (uint32_t)(i + 32 * h + (int32_t)(31u - (aB & 31u)) + c0[1]);
      if (REVERSE) {
        const uint32_t kA = 4095u - jA, kB = 4095u - jB;
        if (aA >= thrA && kA > bestA && jA < 4096u) { bestA = kA; rowA = (uint32_t)i; }
        if (aB >= thrB && kB > bestB && jB < 4096u) { bestB = kB; rowB = (uint32_t)i; }
      } else {
        const uint32_t kA = ((aA >> 5) << 12) | (4095u - (jA & 4095u)), kB = ((aB >> 5) << 12) | (4095u - (jB & 4095u));
        if (aA >= 32u && kA > bestA) { bestA = kA; rowA = (uint32_t)i; }
        if (aB >= 32u && kB > bestB) { bestB = kB; rowB = (uint32_t)i; }
      }
    }
  }

#pragma unroll
  for (int al = 0; al < 2; al++) {
    if (al == 1 && single) break;
    const uint32_t idx = al ? ib : ia;
    const uint32_t best = al ? bestB : bestA, brow = al ? rowB : rowA;
    if (REVERSE) {
      const SwRes &r0 = al ? rb : ra;
      if (best) {
        res[idx].ref_begin = r0.ref_end - (int32_t)(4095u - best);
        res[idx].read_begin = r0.read_end - (int32_t)brow;
        res[idx].flags = r0.flags | SWR_REV_TIER(SWR_TIER_OF_W(W));
      } else {                      // cannot happen when the bound holds; never guess: hand over to the full kernel
        const uint32_t k = atomicAdd(fb_count, 1u);
        fb_keys[k].key = (uint64_t)(r0.ref_end + 1); fb_keys[k].val = idx;
      }
    } else {
      const int32_t S = (int32_t)(best >> 12);
      const int32_t a = ceil_div_pos(S, sc.match);
      const bool proven = S > 0 && c0[al] <= -(rows[al] - a) && (cols[al] - a) <= c0[al] + (W - 1);
      if (proven) {
        SwRes o;
        o.flags = SWR_FWD_TIER(SWR_TIER_OF_W(W)); o.pad0 = o.pad1 = 0; o.ref_begin = -1; o.read_begin = 0;
        o.score = S; o.ref_end = (int32_t)(4095u - (best & 4095u)); o.read_end = (int32_t)brow;
        res[idx] = o;
      } else if (MODE == 0 && next_list && S > 0 && rows[al] + cols[al] - 2 * a + 1 <= SWB_MAXW) {
        res[idx].score = S;         // lower bound: every alignment scoring >= S lies in [-(m - a), n - a]
        next_list[atomicAdd(next_count, 1u)] = idx;
      } else {
        const uint32_t k = atomicAdd(fb_count, 1u);
        fb_keys[k].key = (uint64_t)(al ? tb.n : ta.n); fb_keys[k].val = idx;
      }
    }
  }
}
