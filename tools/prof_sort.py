"""Radix-sort microbenchmark / ncu target: sorts N random 16-byte records by the full 64-bit key."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_pkg()
n = int(os.environ.get("N", str(64_000_000)))
rng = np.random.default_rng(1)
recs = np.zeros(n, dtype=pkg.KMER_DT)
recs["kmer"] = rng.integers(0, 2**63, size=n, dtype=np.uint64)
recs["id_flags"] = np.arange(n, dtype=np.uint32)
al = pkg.Aligner()
for _ in range(int(os.environ.get("REPS", "3"))):
    got, ms = al.sort_records(recs, 0, 64)
    print(f"n={n} sort {ms:.3f} ms  -> {32*n*8/ms/1e6:.1f} GB/s moved over 8 passes, {32*n/ms/1e6:.1f} GB/s algorithmic (one logical sort)")
assert (np.diff(got["kmer"].astype(np.uint64)) >= 0).all() if n < 70_000_000 else True
al.close()
