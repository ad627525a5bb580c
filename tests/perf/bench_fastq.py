"""FASTQ ingest throughput: kslam_fastq_* (chunk-parallel) next to the reference's own reader (oracle/_ref) on the same
files. Host-side measurement; run anywhere: python tests/perf/bench_fastq.py [pairs]"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as ge
import _lib as T
pkg = ge.load_pkg()
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
gb, go = pkg.synth.random_genomes(4, 1_000_000, seed=1)
rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
L = 150
d = tempfile.mkdtemp()
paths = []
for k in range(2):
    rows = rb.reshape(-1, L)[k * pairs:(k + 1) * pairs]
    p = os.path.join(d, f"R{k + 1}.fq"); paths.append(p)
    with open(p, "wb") as f:
        for lo in range(0, pairs, 100_000):
            blk = rows[lo:lo + 100_000]
            f.write(b"".join(b"@r%d/%d\n" % (lo + i, k + 1) + blk[i].tobytes() + b"\n+\n" + b"I" * L + b"\n" for i in range(len(blk))))
size = sum(os.path.getsize(p) for p in paths)
for threads in (1, os.cpu_count() or 1):
    t0 = time.time()
    with pkg.FastqReader(paths[0], paths[1], threads=threads) as rd:
        n = 0
        while True:
            b = rd.next(250_000, copy=False)
            if b is None:
                break
            n += len(b)
    dt = time.time() - t0
    print(f"kslam_fastq threads={threads}: {n} reads, {size / dt / 1e9:.2f} GB/s of FASTQ text, {n / dt / 1e6:.1f} M reads/s")
if T.have_ref():
    L_ = T.ref()
    t0 = time.time()
    h = L_.kref_fastq_open(paths[0].encode(), paths[1].encode()); n = 0
    while True:
        k = L_.kref_fastq_next(h, 250_000)
        if k == 0:
            break
        n += k
    L_.kref_fastq_close(h)
    dt = time.time() - t0
    print(f"reference reader (1 thread, as it is): {n} reads, {size / dt / 1e9:.2f} GB/s, {n / dt / 1e6:.1f} M reads/s")
