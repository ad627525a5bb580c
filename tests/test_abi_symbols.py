"""CPU tier: the C-ABI library loads and exports every symbol include/kslam.h declares; compute entry
points fail loudly (never fall back) when no GPU is present."""
import ctypes as C
import os

import numpy as np
import pytest


def test_header_symbols_exported(pkg):
    syms = pkg.declared_symbols()
    assert len(syms) >= 20 and "kslam_align_batch" in syms and "kslam_pair_batch" in syms
    L = pkg.lib()   # raises if the .so is missing or lacks a declared symbol
    for s in syms:
        assert hasattr(L, s), s
    assert b"sm_100a" in L.kslam_version()


def test_record_layouts_match_header(pkg):
    assert pkg.KMER_DT.itemsize == 16 and pkg.SEED_DT.itemsize == 16
    assert pkg.OVERLAP_DT.itemsize == 48 and pkg.PAIR_DT.itemsize == 32
    assert C.sizeof(pkg.Params) == 24
    assert C.sizeof(pkg.Timings) == 16 * 4 + (23 + 2 + 2 * 12 + 2) * 8      # + tier96 / tier128, forward and reverse counts of the 12 band tiers


def test_exact_domain(pkg):
    """Every scoring parameter set is bit-exact; kslam_params_fast tells which kernels run (packed vs literal striped)."""
    L = pkg.lib()
    ok = pkg.Params(2, 3, 5, 2, 0, 0, 0, 0, 16, 32, 0)
    assert L.kslam_params_exact(C.byref(ok)) == 1 and L.kslam_params_fast(C.byref(ok)) == 1
    for slow in [(5, 4, 10, 10), (1, 1, 1, 1), (2, 8, 3, 3)]:   # gap_extend >= gap_open or mismatch > 2*gap_extend
        p = pkg.Params(*slow, 0, 0, 0, 0, 16, 32, 0)
        assert L.kslam_params_exact(C.byref(p)) == 1 and L.kslam_params_fast(C.byref(p)) == 0


def test_no_cpu_fallback(pkg):
    """Without a usable sm_100 device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tier")
    with pytest.raises(pkg.KslamError, match="no CUDA device|sm_"):
        pkg.Aligner()


def test_product_never_touches_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "k-slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "kslam_oracle" not in text and "libkslam_ref" not in text, f"{f} references the oracle"
