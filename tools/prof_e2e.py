"""End-to-end (host buffers in, results out) timing of one config-1 batch with a per-phase wall-clock split."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_pkg()
pairs = int(os.environ.get("PAIRS", "1000000"))
gb, go = pkg.synth.random_genomes(50, 3_000_000, seed=1)
rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
rb = torch.from_numpy(rb).pin_memory().numpy()
al = pkg.Aligner(report_cigar=bool(int(os.environ.get("CIGAR", "0")))); al.set_debug_taps(False)
al.load_genomes(gb, go)
for i in range(5):
    t0 = time.perf_counter(); al.upload_reads(rb, ro)
    t1 = time.perf_counter(); al.align_resident(fetch=False)
    t2 = time.perf_counter(); res = al.align_resident(fetch=True, copy=False)
    t3 = time.perf_counter(); pr = al.pair_batch(fetch=True, copy=False)
    t4 = time.perf_counter()
    tm = al.timings()
    print(f"upload {1e3*(t1-t0):.1f} ms | align {1e3*(t2-t1):.1f} | align+fetch {1e3*(t3-t2):.1f} (d2h {tm['ms_d2h']:.1f}) | pair+fetch {1e3*(t4-t3):.1f} (dev {tm['ms_pair']:.1f})")
t0 = time.perf_counter()
for i in range(5):
    res = al.align_batch(rb, ro, copy=False); pr = al.pair_batch(fetch=True, copy=False)
dt = (time.perf_counter() - t0) / 5
print(f"e2e {dt*1e3:.1f} ms/step -> {pairs/dt*60/1e6:.1f} M pairs/min")
al.close()
