"""CPU tier: the oracle restatement (oracle/kslam_oracle.c) against the reference's own code
(oracle/_ref/libkslam_ref.so, built from /root/reference by oracle/Makefile). Skipped where the
prebuilt _ref is absent; the golden-vector tests cover the oracle there."""
import numpy as np
import pytest

import _lib as T

pytestmark = pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built (no /root/reference here)")

FIELDS = ["read", "entry", "rel", "rev_comp", "ref_begin", "ref_end", "query_begin", "query_end", "sw_score", "cigar_len"]


def canon(a):
    return np.sort(a, order=list(a.dtype.names))


def assert_overlaps_equal(a, pa, b, pb):
    assert len(a) == len(b)
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), f
    assert T.cigars_of(a, pa) == T.cigars_of(b, pb)


@pytest.mark.parametrize("cigar,thr", [(1, 0), (0, 0), (1, 70)])
def test_pipeline_adversarial(pkg, cigar, thr):
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=11, n_genomes=10, glen=12_000, n_pairs=1500)
    P = T.default_params(report_cigar=cigar, score_threshold=thr)
    R = T.Ref(gb, go, rb, ro, P)
    got = T.ko_pipeline(gb, go, rb, ro, P)
    assert np.array_equal(R.kmers(1), got["read_kmers"])
    assert np.array_equal(R.kmers(2), got["genome_kmers"])
    allk = R.kmers(3, sort=True)
    assert np.array_equal(allk["kmer"], got["sorted_kmers"]["kmer"])
    assert np.array_equal(allk["id_flags"], got["sorted_kmers"]["id_flags"])
    assert np.array_equal(canon(R.seeds(raw=True)), canon(got["raw_seeds"]))
    assert np.array_equal(R.seeds(raw=False), got["seeds"])
    ov, pool = R.align_to_database()
    assert (got["overlaps"]["flags"] & 1).sum() == 0
    assert_overlaps_equal(ov, pool, got["overlaps"], got["cigar_pool"])
    ovs, pools, pairs = R.screen_and_pair()
    assert_overlaps_equal(ovs, pools, got["pair_sorted_overlaps"], got["cigar_pool"])
    assert np.array_equal(pairs, got["pairs"])


def test_pipeline_config1_shape(pkg):
    gb, go = pkg.synth.random_genomes(5, 100_000, seed=1)
    rb, ro, _ = pkg.synth.paired_reads(gb, go, 4000, seed=2)
    P = T.default_params(report_cigar=1)
    R = T.Ref(gb, go, rb, ro, P)
    got = T.ko_pipeline(gb, go, rb, ro, P)
    ov, pool = R.align_to_database()
    assert len(ov) > 7000
    assert_overlaps_equal(ov, pool, got["overlaps"], got["cigar_pool"])
    ovs, pools, pairs = R.screen_and_pair()
    assert np.array_equal(pairs, got["pairs"])
    # Tests.h:136,323 property: a perfect overlap scores 2 x overlap length
    assert got["overlaps"]["sw_score"].max() == 300


@pytest.mark.parametrize("shape", [(150, 150), (150, 300), (101, 140), (60, 60)])
@pytest.mark.parametrize("cigar", [0, 1])
def test_ssw_config3(pkg, shape, cigar):
    q, qo, r, ro = pkg.synth.sw_pairs(6000, shape[0], shape[1], seed=shape[0] + shape[1] + cigar)
    P = T.default_params(report_cigar=cigar)
    a, pa = T.ref_ssw_batch(q, qo, r, ro, P)
    b, pb = T.ko_ssw_batch(q, qo, r, ro, P)
    assert (b["flags"] & 1).sum() == 0
    assert_overlaps_equal(a, pa, b, pb)


def test_ssw_tandem_repeats(pkg):
    rng = np.random.default_rng(5)
    ACGT = pkg.synth.ACGT
    qs, rs = [], []
    for _ in range(8000):
        per = int(rng.integers(2, 25)); unit = ACGT[rng.integers(0, 4, size=per)]
        L = int(rng.integers(60, 181)); w = np.tile(unit, L // per + 2)[:L].copy()
        m = rng.random(L) < 0.03; w[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
        ql = int(rng.integers(40, 161)); st = int(rng.integers(0, max(1, L - ql))); qq = w[st:st + ql].copy()
        m = rng.random(len(qq)) < 0.03; qq[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
        if rng.random() < 0.3 and len(qq) > 30:
            p = int(rng.integers(5, len(qq) - 5)); k = int(rng.integers(1, 6)); qq = np.concatenate([qq[:p], qq[p + k:]])
        qs.append(qq); rs.append(w)
    q, qo = T.concat(qs); r, ro = T.concat(rs)
    for P in (T.default_params(report_cigar=1), T.default_params(report_cigar=1, score_threshold=100)):
        a, pa = T.ref_ssw_batch(q, qo, r, ro, P)
        b, pb = T.ko_ssw_batch(q, qo, r, ro, P)
        ok = (b["flags"] & 1) == 0
        assert ok.mean() > 0.999
        assert_overlaps_equal(a[ok], pa, b[ok], pb)


COORDS = ["ref_begin", "ref_end", "query_begin", "query_end", "sw_score"]


def test_ssw_other_scoring_inside_the_gotoh_domain(pkg):
    """Parameters with gap_extend < gap_open and mismatch <= 2*gap_extend: the striped restatement equals
    the reference, and plain Gotoh (what the packed CUDA kernels compute) equals both."""
    q, qo, r, ro = pkg.synth.sw_pairs(3000, 120, 160, seed=77)
    for (m, x, go, ge) in [(1, 2, 6, 1), (2, 4, 6, 2), (3, 4, 5, 3), (1, 1, 3, 2), (3, 2, 9, 1), (2, 3, 5, 2)]:
        P = T.default_params(report_cigar=1, match=m, mismatch=x, gap_open=go, gap_extend=ge)
        a, pa = T.ref_ssw_batch(q, qo, r, ro, P, cigar_cap=256)
        b, pb = T.ko_ssw_batch(q, qo, r, ro, P, cigar_cap=256)
        assert_overlaps_equal(a, pa, b, pb)
        g = T.ko_ssw_gotoh(q[:600 * 120], qo[:601], r[:600 * 160], ro[:601], P)
        for f in COORDS:
            assert np.array_equal(g[f], a[f][:600]), (f, (m, x, go, ge))


@pytest.mark.parametrize("prm", [(5, 4, 10, 10), (1, 1, 1, 1), (2, 8, 3, 3), (3, 1, 1, 4), (2, 9, 2, 1), (4, 6, 4, 4), (10, 2, 3, 3), (2, 3, 2, 5)])
def test_ssw_scoring_outside_the_gotoh_domain(pkg, prm):
    """gap_extend >= gap_open or mismatch > 2 * gap_extend: SSW's answer depends on its striping (E before lazy-F, the two
    lazy-F loops, ssw.c:257-305,514-524). The oracle restates the striped kernels lane by lane and must equal the reference:
    score, the four coordinates and the CIGAR, with and without the score filter."""
    m, x, go, ge = prm
    q, qo, r, ro = pkg.synth.sw_pairs(2500, 120, 160, seed=78)
    q2, qo2, r2, ro2 = pkg.synth.sw_pairs(1500, 150, 150, seed=79)
    for (qq, qqo, rr, rro) in ((q, qo, r, ro), (q2, qo2, r2, ro2)):
        for thr in (0, 90):
            P = T.default_params(report_cigar=1, match=m, mismatch=x, gap_open=go, gap_extend=ge, score_threshold=thr)
            a, pa = T.ref_ssw_batch(qq, qqo, rr, rro, P, cigar_cap=512)
            b, pb = T.ko_ssw_batch(qq, qqo, rr, rro, P, cigar_cap=512)
            ok = (b["flags"] & 1) == 0
            assert ok.mean() > 0.999
            assert_overlaps_equal(a[ok], pa, b[ok], pb)


def test_ssw_random_parameter_sweep(pkg):
    """40 random (match, mismatch, gap_open, gap_extend) over ragged lengths, repeats and N runs."""
    rng = np.random.default_rng(1234)
    q, qo, r, ro = pkg.synth.sw_pairs(700, 100, 130, seed=80)
    for _ in range(40):
        m, x, go, ge = int(rng.integers(1, 12)), int(rng.integers(0, 12)), int(rng.integers(0, 14)), int(rng.integers(0, 14))
        P = T.default_params(report_cigar=1, match=m, mismatch=x, gap_open=go, gap_extend=ge)
        a, pa = T.ref_ssw_batch(q, qo, r, ro, P, cigar_cap=512)
        b, pb = T.ko_ssw_batch(q, qo, r, ro, P, cigar_cap=512)
        ok = (b["flags"] & 1) == 0
        assert ok.mean() > 0.99, (m, x, go, ge)
        for f in COORDS + ["cigar_len"]:
            assert np.array_equal(a[f][ok], b[f][ok]), (f, (m, x, go, ge))
        ca, cb = T.cigars_of(a, pa), T.cigars_of(b, pb)
        assert all(ca[i] == cb[i] for i in range(len(ca)) if ok[i]), (m, x, go, ge)
