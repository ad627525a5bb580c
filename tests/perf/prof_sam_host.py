"""Host-stage profile without a GPU: alignments of a config-1-like batch from the reference (oracle/_ref, test infrastructure),
then kslam_sam_batch timed with KSLAM_SAM_TRACE=1. Usage: python tests/perf/prof_sam_host.py [pairs] [threads]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import _lib as T  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pkg = T.load_pkg()
gb, go = pkg.synth.random_genomes(10, 1_000_000, seed=1)
rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
cache = f"/tmp/prof_sam_{pairs}.npz"
if os.path.exists(cache):
    z = np.load(cache); ov, pool, pr = z["ov"], z["pool"], z["pr"]
else:
    R = T.Ref(gb, go, rb, ro, T.default_params(report_cigar=1))
    t0 = time.time(); R.align_to_database(); ov, pool, pr = R.screen_and_pair(); R.close()
    print(f"reference alignToDatabase + pairing: {time.time() - t0:.1f}s")
    np.savez(cache, ov=ov, pool=pool, pr=pr)
n = len(ro) - 1
quals = np.full(len(rb), ord("I"), np.uint8)
ids = [b"r%d" % (i % pairs) for i in range(n)]
idb = np.frombuffer(b"".join(ids), np.uint8); ido = np.zeros(n + 1, np.uint64); ido[1:] = np.cumsum([len(x) for x in ids])
os.environ["KSLAM_SAM_TRACE"] = "1"
w = pkg.SamWriter(gb, go, [f"g{i}" for i in range(len(go) - 1)], report_cigar=True, threads=threads)
for _ in range(3):
    t0 = time.time()
    text, mi = w.batch(rb, ro, quals, ro, idb, ido, ov, pool, pr)
    print(f"kslam_sam_batch: {time.time() - t0:.3f}s for {pairs} pairs, {len(text) / 1e6:.1f} MB, {pairs / (time.time() - t0) / 1e6:.2f} M pairs/s")
