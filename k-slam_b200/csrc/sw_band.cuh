// sw_band.cuh — banded forward / reverse Smith-Waterman sweep, exact by construction (included by sw.cu).
//
// Why a band can be EXACT here. An alignment that scores S needs at least a = ceil(S / match) diagonal steps.
// A path that touches matrix offset d = j - i consumes d extra columns (d >= 0) or -d extra rows, so inside an
// m x n matrix it has at most n - d (resp. m + d) diagonal steps. Hence EVERY alignment scoring >= S lies inside
// the offsets [-(m - a), n - a]. Forward pass: sweep a 32-diagonal band placed on that interval's centre, take
// the best score S' found in it (a true alignment score, so S' <= S); if [-(m - a'), n - a'] fits in the band then
// all alignments scoring >= S' — in particular every optimal one, and therefore every cell SSW's tie rules
// (ssw.c:316-342) look at — are inside the band and were computed with their true H: the result is exact.
// Otherwise the alignment is appended to the fallback list and the full-matrix kernel (k_sw_fast) runs it.
// Reverse pass (ssw.c:905-923): S is known, the band is exactly [-(rows - a), cols - a]; used when <= 32 wide.
//
// Mapping: ONE THREAD owns two alignments (halves of s16x2 registers) and keeps the band's previous row (H), the
// vertical-gap state (V) and 32 PRMT column selectors in registers; it walks the rows, 32 cells per row fully
// unrolled, 7 DPX/PRMT ops per cell pair, no shuffles and no shared-memory traffic in the recurrences. Column
// selectors slide by one register per row (moves go to the FMA pipe; the ALU pipe only sees the recurrences).
// Band cells that fall outside the matrix use "replicate-sign" selectors that can only produce 0 or -1, so they
// stay at H = 0 on the left and can never reach a maximum on the right.
#pragma once

#define SWB_W 32
#define SWB_MAXROWS 160
#define SWB_COLS (SWB_MAXROWS + SWB_W)
#define SWB_BLOCK 128

__device__ __forceinline__ int32_t ceil_div_pos(int32_t a, int32_t b) { return (a + b - 1) / b; }

// selector byte for one alignment's column: (copy byte w, replicate sign of byte w) of its profile register.
// `second` selects the second PRMT source (alignment B). Out-of-matrix columns use two replicate nibbles.
__device__ __forceinline__ uint32_t band_sel_byte(uint32_t w, bool in_range, bool second) {
  const uint32_t base = second ? 4u : 0u;
  if (!in_range) return (8u | base) | ((8u | base) << 4);
  return (base + w) | ((8u | (base + w)) << 4);
}

template <bool REVERSE>
__global__ void __launch_bounds__(SWB_BLOCK, 3)
k_sw_band(const SwTask *__restrict__ tasks, const uint32_t *__restrict__ list, uint32_t n_list, SwPlanes pl, SwScore sc,
          SwRes *__restrict__ res, Rec16 *__restrict__ fb_keys, uint32_t *__restrict__ fb_count) {
  extern __shared__ uint8_t s_selb[];   // [2][SWB_COLS][SWB_BLOCK]
  const uint32_t tid = threadIdx.x;
  const uint32_t p = blockIdx.x * SWB_BLOCK + tid;
  if (2 * p >= n_list) return;
  const uint32_t ia = list[2 * p], ib = (2 * p + 1 < n_list) ? list[2 * p + 1] : ia;
  const SwTask ta = tasks[ia], tb = tasks[ib];
  SwRes ra, rb;
  int32_t rows[2], cols[2], c0[2];
  if (REVERSE) {
    ra = res[ia]; rb = res[ib];
    rows[0] = ra.read_end + 1; cols[0] = ra.ref_end + 1; rows[1] = rb.read_end + 1; cols[1] = rb.ref_end + 1;
    c0[0] = -(rows[0] - ceil_div_pos(ra.score, sc.match)); c0[1] = -(rows[1] - ceil_div_pos(rb.score, sc.match));
  } else {
    rows[0] = (int32_t)ta.m; cols[0] = (int32_t)ta.n; rows[1] = (int32_t)tb.m; cols[1] = (int32_t)tb.n;
    // centre of [-(m - a), n - a] is (n - m) / 2 whatever a is; band = [c0, c0 + 31]
    c0[0] = ((cols[0] - rows[0]) >> 1) - SWB_W / 2; c0[1] = ((cols[1] - rows[1]) >> 1) - SWB_W / 2;
  }
  const int32_t rows_max = rows[0] > rows[1] ? rows[0] : rows[1];

  // ---- column selector bytes for band column index k (matrix column j = k + c0), k in [0, rows_max + 32)
#pragma unroll
  for (int al = 0; al < 2; al++) {
    const SwTask &t = al ? tb : ta;
    const bool rev = (t.flags & SWT_REV) != 0;
    const int32_t ref_end = REVERSE ? (al ? rb.ref_end : ra.ref_end) : 0;
    uint64_t curw = ~0ull, bits = 0; uint32_t xm = 0;
    uint8_t *dst = s_selb + (size_t)al * SWB_COLS * SWB_BLOCK + tid;
    for (int32_t k = 0; k < rows_max + SWB_W; k++) {
      const int32_t j = k + c0[al];
      uint32_t byte;
      if (j < 0 || j >= cols[al]) byte = band_sel_byte(0, false, al);
      else {
        const uint32_t x = REVERSE ? (uint32_t)(ref_end - j) : (uint32_t)j;
        const uint32_t pos = rev ? t.w_start + t.n - 1 - x : t.w_start + x;
        const uint64_t w = t.w_word + (pos >> 5);
        if (w != curw) { curw = w; bits = __ldg(&pl.w_sbits[w]); xm = rev ? __ldg(&pl.w_xmask[w]) : 0u; }
        uint32_t c = (uint32_t)(bits >> (2 * (pos & 31))) & 3u;
        if (rev && !((xm >> (pos & 31)) & 1u)) c = 3u - c;
        byte = band_sel_byte(c, true, al);
      }
      dst[(size_t)k * SWB_BLOCK] = (uint8_t)byte;
    }
  }
  // (each thread reads back only what it wrote: no barrier needed)

  const uint32_t mis_b = (uint32_t)(-(sc.mismatch * 32)) & 0xffu, mat_b = (uint32_t)(sc.match * 32) & 0xffu;
  const uint32_t MIS4 = mis_b * 0x01010101u, DIFF = mis_b ^ mat_b;
  const uint32_t NEG_GO = pack2(-sc.gap_open * 32), NEG_GE = pack2(-sc.gap_extend * 32), MIN2 = 0x80008000u;
  const uint8_t *selA = s_selb + tid, *selB = s_selb + (size_t)SWB_COLS * SWB_BLOCK + tid;

  uint32_t H[SWB_W], V[SWB_W], sel[SWB_W];
#pragma unroll
  for (int t = 0; t < SWB_W; t++) {
    H[t] = 0; V[t] = 0;
    sel[t] = (uint32_t)selA[(size_t)t * SWB_BLOCK] | ((uint32_t)selB[(size_t)t * SWB_BLOCK] << 8);
  }
  // forward: key = score << 12 | (4095 - column); reverse: key = 4095 - scan column of the first hit (0 = none)
  uint32_t bestA = 0, bestB = 0, rowA = 0, rowB = 0;
  const uint32_t thrA = REVERSE ? (uint32_t)ra.score * 32u : 0u, thrB = REVERSE ? (uint32_t)rb.score * 32u : 0u;

  for (int32_t i = 0; i < rows_max; i++) {
    // row profiles [s(q,A) s(q,C) s(q,G) s(q,T)] x 32; code-4 rows score 0; rows past the query mismatch everything
    uint32_t PA = MIS4, PB = MIS4;
    if (i < rows[0]) { const uint32_t c = q_code(pl, ta, REVERSE ? (uint32_t)(rows[0] - 1 - i) : (uint32_t)i);
                       PA = c == 4 ? 0u : (MIS4 ^ (DIFF << (8 * c))); }
    if (i < rows[1]) { const uint32_t c = q_code(pl, tb, REVERSE ? (uint32_t)(rows[1] - 1 - i) : (uint32_t)i);
                       PB = c == 4 ? 0u : (MIS4 ^ (DIFF << (8 * c))); }
    uint32_t e = 0, acc = 0;
#pragma unroll
    for (int t = 0; t < SWB_W; t++) {
      const uint32_t s = prmt(PA, PB, sel[t]);
      const uint32_t v = V[t];
      uint32_t h = __viaddmax_s16x2_relu(H[t], s, v);           // max(H[i-1][j-1] + s, vertical gap, 0)
      h = __vimax3_s16x2(h, e, e);                                // ... and the horizontal gap
      H[t] = h;
      const uint32_t hgo = __viaddmax_s16x2(h, NEG_GO, MIN2);
      e = __viaddmax_s16x2(e, NEG_GE, hgo);                      // horizontal gap into (i, j+1)
      if (t > 0) V[t - 1] = __viaddmax_s16x2(v, NEG_GE, hgo);    // vertical gap into (i+1, j): band slot t-1 next row
      acc = __viaddmax_s16x2(h, (uint32_t)(31 - t) * 0x10001u, acc);   // max of H*32 + (31 - t): smallest column wins ties
    }
    // slide the column selectors: next row's slot t is this row's slot t+1; slot 31 takes the entering column
#pragma unroll
    for (int t = 0; t < SWB_W - 1; t++) sel[t] = sel[t + 1];
    {
      const size_t k = (size_t)(i + SWB_W) * SWB_BLOCK;
      sel[SWB_W - 1] = (uint32_t)selA[k] | ((uint32_t)selB[k] << 8);
    }
    // row winners -> running best (first column, then smallest row: rows only grow, so ties keep the old one)
    const uint32_t aA = acc & 0xffffu, aB = acc >> 16;
    const uint32_t jA = (uint32_t)(i + (int32_t)(31u - (aA & 31u)) + c0[0]), jB = (uint32_t)(i + (int32_t)(31u - (aB & 31u)) + c0[1]);
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

#pragma unroll
  for (int al = 0; al < 2; al++) {
    if (al == 1 && ib == ia) break;
    const uint32_t idx = al ? ib : ia;
    const uint32_t best = al ? bestB : bestA, brow = al ? rowB : rowA;
    if (REVERSE) {
      const SwRes &r0 = al ? rb : ra;
      if (best) {
        res[idx].ref_begin = r0.ref_end - (int32_t)(4095u - best);
        res[idx].read_begin = r0.read_end - (int32_t)brow;
      } else {                      // cannot happen when the bound holds; never guess: hand over to the full kernel
        const uint32_t k = atomicAdd(fb_count, 1u);
        fb_keys[k].key = (uint64_t)(r0.ref_end + 1); fb_keys[k].val = idx;
      }
    } else {
      const int32_t S = (int32_t)(best >> 12);
      const int32_t a = ceil_div_pos(S, sc.match);
      const bool proven = S > 0 && c0[al] <= -(rows[al] - a) && (cols[al] - a) <= c0[al] + (SWB_W - 1);
      if (proven) {
        SwRes o;
        o.flags = 0; o.pad0 = o.pad1 = 0; o.ref_begin = -1; o.read_begin = 0;
        o.score = S; o.ref_end = (int32_t)(4095u - (best & 4095u)); o.read_end = (int32_t)brow;
        res[idx] = o;
      } else {
        const uint32_t k = atomicAdd(fb_count, 1u);
        fb_keys[k].key = (uint64_t)(al ? tb.n : ta.n); fb_keys[k].val = idx;
      }
    }
  }
}
