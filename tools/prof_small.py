"""Small single-batch run for ncu captures: config-1 shape, PAIRS read pairs (default 200k), no cigar."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_pkg()
pairs = int(os.environ.get("PAIRS", "200000"))
cigar = bool(int(os.environ.get("CIGAR", "0")))
if os.environ.get("WORKLOAD", "config1") == "tree":      # config-2-like: related genomes, multi-genome piles, all SW tiers
    gb, go = pkg.synth.tree_genomes(100, 1_000_000, seed=1)
else:
    gb, go = pkg.synth.random_genomes(50, 3_000_000, seed=1)
rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
al = pkg.Aligner(report_cigar=cigar); al.set_debug_taps(False)
al.load_genomes(gb, go)
al.upload_reads(rb, ro)
for _ in range(int(os.environ.get("REPS", "2"))):
    al.align_resident(); al.pair_batch(fetch=False)
print({k: v for k, v in al.timings().items()})
al.close()
