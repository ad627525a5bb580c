"""CPU tier: the oracle against golden vectors produced by the reference's own code
(tests/golden/make_golden.py). These travel with the repo, so they pin the oracle on the GPU box too."""
import ctypes as C

import numpy as np
import pytest

import _lib as T

FIELDS = ["read", "entry", "rel", "rev_comp", "ref_begin", "ref_end", "query_begin", "query_end", "sw_score", "cigar_len"]


def params_of(g):
    p = g["params"]
    return T.default_params(match=int(p[0]), mismatch=int(p[1]), gap_open=int(p[2]), gap_extend=int(p[3]),
                            score_threshold=int(p[4]), report_cigar=int(p[5]))


@pytest.mark.parametrize("name", ["pipeline_adversarial_cigar.npz", "pipeline_adversarial_thr60.npz",
                                  "pipeline_config1_mini.npz"])
def test_pipeline_golden(golden, name):
    g = golden(name)
    P = params_of(g)
    got = T.ko_pipeline(g["gen_bases"], g["gen_offs"], g["read_bases"], g["read_offs"], P)
    assert np.array_equal(got["read_kmers"], g["read_kmers"])
    assert np.array_equal(got["genome_kmers"], g["genome_kmers"])
    raw = np.sort(got["raw_seeds"], order=["read", "entry", "rel", "rev_comp"])
    assert np.array_equal(raw, g["raw_seeds_sorted"])
    assert np.array_equal(got["seeds"], g["seeds"])
    for f in FIELDS:
        assert np.array_equal(got["overlaps"][f], g["overlaps"][f]), f
    assert T.cigars_of(got["overlaps"], got["cigar_pool"]) == T.cigars_of(g["overlaps"], g["cigar_pool"])
    for f in FIELDS:
        assert np.array_equal(got["pair_sorted_overlaps"][f], g["pair_sorted_overlaps"][f]), f
    assert np.array_equal(got["pairs"], g["pairs"])


@pytest.mark.parametrize("name", ["ssw_150x150.npz", "ssw_150x300.npz", "ssw_101x140_nocigar.npz",
                                  "ssw_params_5_4_10_10.npz", "ssw_params_2_8_3_3.npz", "ssw_params_1_1_1_1.npz"])
def test_ssw_golden(golden, name):
    g = golden(name)
    P = params_of(g)
    b, pb = T.ko_ssw_batch(g["q"], g["qoffs"], g["r"], g["roffs"], P, cigar_cap=len(g["cigar_pool"]) // len(g["expect"]))
    a = g["expect"]
    for f in FIELDS[4:]:
        assert np.array_equal(a[f], b[f]), f
    assert T.cigars_of(a, g["cigar_pool"]) == T.cigars_of(b, pb)


def test_kmer_codec_known_answers():
    """KMer.h:23-27 worked example style: A0 C1 T2 G3, first base most significant; palindromes take rc."""
    s = b"ACGT" * 8  # its own reverse complement
    k = T.ko_extract(s, np.array([0, 32], dtype=np.uint64), False, 1)
    f = 0
    for ch in s:
        f = (f << 2) | {65: 0, 67: 1, 84: 2, 71: 3}[ch]
    assert len(k) == 1 and int(k[0]["kmer"]) == f and (int(k[0]["id_flags"]) >> 30) & 1 == 1
    # lower case / N bases encode as A (KMer.h:261-263); poly-A 32-mer is 0 and its rc (poly-T) is larger
    k = T.ko_extract(b"n" * 32 + b"a", np.array([0, 33], dtype=np.uint64), True, 1)
    assert len(k) == 2 and (k["kmer"] == 0).all() and ((k["id_flags"] >> 30) & 1 == 0).all()
    # shorter than k: nothing (KMer.h:167); gap 16 count = (n-32)/16+1 (KMer.h:202)
    assert len(T.ko_extract(b"ACGT" * 7, np.array([0, 28], dtype=np.uint64), True, 16)) == 0
    assert len(T.ko_extract(b"ACGT" * 25, np.array([0, 100], dtype=np.uint64), True, 16)) == (100 - 32) // 16 + 1


def test_unique_is_against_last_kept():
    """Overlap.h:79-85,290: rel 0,2,4 keeps 0 and 4 (|4-0| >= 3 against the last KEPT element)."""
    s = np.zeros(5, dtype=T.SEED_DT)
    s["rel"] = [4, 0, 2, 9, 7]
    out = T.ko_sort_unique(s)
    assert out["rel"].tolist() == [0, 4, 7]


def test_perfect_overlap_scores_twice_length():
    """Tests.h:136,323: sw_score == 2 x overlapping length."""
    rng = np.random.default_rng(1)
    w = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=150)]
    P = T.default_params(report_cigar=1)
    out, pool = T.ko_ssw_batch(w[20:120], np.array([0, 100], np.uint64), w, np.array([0, 150], np.uint64), P)
    assert out[0]["sw_score"] == 200 and out[0]["ref_begin"] == 20 and out[0]["ref_end"] == 119
    assert T.cigars_of(out, pool)[0] == (100 << 4,)
