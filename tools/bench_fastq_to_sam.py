"""End-to-end --sam-file run at config-1 scale: FASTQ text + FASTA -> SAM text, per-stage wall times.
usage: python tools/bench_fastq_to_sam.py [pairs] [reads_at_once]"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_pkg()
from kslam_b200 import slam
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
at_once = int(sys.argv[2]) if len(sys.argv) > 2 else pairs
gb, go = pkg.synth.random_genomes(50, 3_000_000, seed=1)
rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
d = tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp"))
fa = os.path.join(d, "db.fa")
with open(fa, "wb") as f:
    for i in range(len(go) - 1):
        f.write(b">g%02d synthetic\n" % i + gb[int(go[i]):int(go[i + 1])].tobytes() + b"\n")
L = 150
paths = []
for k in range(2):
    rows = rb.reshape(-1, L)[k * pairs:(k + 1) * pairs]
    p = os.path.join(d, f"R{k + 1}.fq"); paths.append(p)
    with open(p, "wb") as f:
        for lo in range(0, pairs, 100_000):
            blk = rows[lo:lo + 100_000]
            f.write(b"".join(b"@r%d/%d\n" % (lo + i, k + 1) + blk[i].tobytes() + b"\n+\n" + b"I" * L + b"\n" for i in range(len(blk))))
for rep in range(2):
    t0 = time.time()
    st = slam.align_to_sam(pkg, [fa], paths[0], paths[1], os.path.join(d, "out.sam"), reads_at_once=at_once)
    dt = time.time() - t0
    print(f"run {rep}: {pairs} pairs in {dt:.2f}s ({pairs / dt * 60 / 1e6:.1f} M pairs/min incl. FASTA parse + index build), {st['batches']} batches, stage busy times {st['seconds']}")
