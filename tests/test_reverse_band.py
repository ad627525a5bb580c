"""CPU tier: the band of the reverse Smith-Waterman sweeps (kslam_reverse_band, csrc/sw_band.cuh: reverse_band) is exact.

SSW's reverse pass (ssw.c:905-923) scans the reversed prefixes read[0..read_end], ref[0..ref_end] and stops at the first
column holding a cell with H == the forward score. The GPU path sweeps only the offsets an alignment ANCHORED at the reversed
origin can reach; this test restates plain Gotoh in Python and checks, over random sequence pairs (mutated copies with
indels, tandem repeats, two-letter alphabets) and random scoring parameters of the plain-Gotoh domain, that the band-
restricted scan finds the same cell as the scan of the full reversed matrix — no GPU, no oracle involved."""
import ctypes as C
import random

import _lib as T


def gotoh_scan(q, r, M, X, go, ge, lo=None, hi=None, thr=None):
    """Column-major local Gotoh like SSW: (best score, first column attaining it, smallest row in it); with `thr` also the
    first column whose maximum reaches thr and the smallest row holding that maximum. Cells outside offsets [lo, hi] are 0."""
    m, n = len(q), len(r)
    H = [[0] * (n + 1) for _ in range(m + 1)]
    E = [[0] * (n + 1) for _ in range(m + 1)]
    F = [[0] * (n + 1) for _ in range(m + 1)]
    best, hit = (0, -1, 0), None
    for j in range(1, n + 1):
        colmax, colrow = 0, 0
        for i in range(1, m + 1):
            d = j - i
            if lo is not None and (d < lo or d > hi):
                continue
            e = max(E[i][j - 1] - ge, H[i][j - 1] - go, 0)
            f = max(F[i - 1][j] - ge, H[i - 1][j] - go, 0)
            h = max(0, H[i - 1][j - 1] + (M if q[i - 1] == r[j - 1] else -X), e, f)
            H[i][j], E[i][j], F[i][j] = h, e, f
            if h > colmax:
                colmax, colrow = h, i - 1
        if colmax > best[0]:
            best = (colmax, j - 1, colrow)
        if thr is not None and hit is None and colmax >= thr:
            hit = (j - 1, colrow)
    return best, hit


def band_of(L, rows, cols, S, M, X, go, ge):
    P = L.Params(M, X, go, ge, 0, 0, 0, 0, 16, 32, 0)
    lo, hi = C.c_int32(), C.c_int32()
    assert L.lib().kslam_reverse_band(rows, cols, S, C.byref(P), C.byref(lo), C.byref(hi)) == 0
    return lo.value, hi.value


def test_reverse_band_is_the_anchored_inequality():
    """hi = the largest d with match * min(rows, cols - d) - gapOpen - (d - 1) gapExtend >= S (lo: rows and cols swapped)."""
    pkg = T.load_pkg()
    pkg.lib().kslam_reverse_band.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = random.Random(5)
    for _ in range(20000):
        ge = rng.randint(0, 5); go = rng.randint(ge + 1, ge + 8); M = rng.randint(1, 3); X = rng.randint(0, 2 * ge)
        rows, cols = rng.randint(1, 200), rng.randint(1, 200)
        S = rng.randint(1, M * min(rows, cols))

        def reach(other, shrinking):
            d = 0
            while d + 1 <= shrinking - 1 and M * min(other, shrinking - (d + 1)) - (go + d * ge) >= S:
                d += 1
            return d
        assert band_of(pkg, rows, cols, S, M, X, go, ge) == (-reach(cols, rows), reach(rows, cols)), (rows, cols, S, M, go, ge)


def test_reverse_band_finds_what_the_full_reverse_pass_finds():
    pkg = T.load_pkg()
    pkg.lib().kslam_reverse_band.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = random.Random(11)
    checked = narrower = 0
    for _ in range(700):
        ge = rng.randint(0, 3); go = rng.randint(ge + 1, ge + 6); X = rng.randint(0, 2 * ge); M = rng.randint(1, 3)
        if rng.random() < 0.5:
            M, X, go, ge = 2, 3, 5, 2
        alpha = "AC" if rng.random() < 0.3 else "ACGT"
        m = rng.randint(5, 40)
        q = [rng.choice(alpha) for _ in range(m)]
        r = []
        for c in q:                                          # a mutated copy of the read with insertions and deletions
            u = rng.random()
            if u < 0.06:
                continue
            if u < 0.12:
                r.append(rng.choice(alpha))
            r.append(rng.choice(alpha) if u < 0.2 else c)
        r = [rng.choice(alpha) for _ in range(rng.randint(0, 8))] + r + [rng.choice(alpha) for _ in range(rng.randint(0, 8))]
        if rng.random() < 0.3:                               # tandem repeats: many equal-score alignments, many begin positions
            unit = [rng.choice(alpha) for _ in range(rng.randint(1, 4))]
            q, r = (unit * 30)[:m], (unit * 30)[:len(r)]
        (S, er, eq), _ = gotoh_scan(q, r, M, X, go, ge)
        if S <= 0:
            continue
        qq, rr = q[:eq + 1][::-1], r[:er + 1][::-1]          # the reversed prefixes (ssw.c:905-915)
        _, full = gotoh_scan(qq, rr, M, X, go, ge, thr=S)
        lo, hi = band_of(pkg, eq + 1, er + 1, S, M, X, go, ge)
        _, band = gotoh_scan(qq, rr, M, X, go, ge, lo=lo, hi=hi, thr=S)
        assert full == band, (q, r, (M, X, go, ge), S, full, band, lo, hi)
        a = -(-S // M)
        narrower += (hi - lo + 1) < (eq + 1) + (er + 1) - 2 * a + 1
        checked += 1
    assert checked > 500 and narrower > checked // 3
