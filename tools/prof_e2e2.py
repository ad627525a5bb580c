"""Diagnostic: per-call wall times of the two-context e2e pipeline (config-1 shape)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_pkg()
pairs = int(os.environ.get("PAIRS", "1000000"))
gb, go = pkg.synth.random_genomes(50, 3_000_000, seed=1)
rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
rbp = torch.from_numpy(rb).pin_memory().numpy()
als = []
ND = int(os.environ.get('DEPTH', '3'))
for k in range(ND):
    a = pkg.Aligner(report_cigar=False, stream_priority=1 if (k == 1 and os.environ.get('PRIO')) else 0); a.set_debug_taps(False); a.load_genomes(gb, go); als.append(a)
for depth in range(1, ND + 1):
    log = []
    def worker(k, n):
        for b in range(k, n, depth):
            t0 = time.perf_counter(); als[k].align_pair_batch(rbp, ro, copy=False); t1 = time.perf_counter()
            tm = als[k].timings()
            log.append((k, b, round((t0 - T0) * 1e3, 2), round((t1 - t0) * 1e3, 2), round(tm["ms_pack"], 2), round(tm["ms_total"], 2), round(tm["ms_pair"], 2)))
    for rep in range(2):
        log.clear()
        T0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(k, 12)) for k in range(depth)]
        [t.start() for t in th]; [t.join() for t in th]
        tot = time.perf_counter() - T0
    print(f"depth {depth}: 12 batches in {tot*1e3:.1f} ms -> {tot*1e3/12:.2f} ms/batch")
    if os.environ.get("VERBOSE"):
        for r in sorted(log, key=lambda r: r[2]):
            print("   ctx %d batch %d start %.2f dur %.2f pack %.2f total %.2f pair %.2f" % r)
