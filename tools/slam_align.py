#!/usr/bin/env python
"""SLAM --just-align --sam-file, through libkslam.so:  python tools/slam_align.py --fasta DB.fa [...] --sam-file out.sam R1.fq R2.fq"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--fasta", nargs="+", help="database as FASTA files (parsed as --parse-fasta does)")
ap.add_argument("--db", help="SLAM database directory (DIR/database)")
ap.add_argument("--sam-file", required=True)
ap.add_argument("--num-reads-at-once", type=int, default=10_000_000)
ap.add_argument("--num-alignments", type=int, default=10)
ap.add_argument("--score-fraction-threshold", type=float, default=0.95)
ap.add_argument("--min-alignment-score", type=int, default=0)
ap.add_argument("--no-pseudo-assembly", action="store_true")
ap.add_argument("--sam-xa", action="store_true")
ap.add_argument("reads", nargs=2)
a = ap.parse_args()
if not a.fasta and not a.db:
    ap.error("give --fasta or --db")
pkg = ge.load_pkg()
from kslam_b200 import slam
st = slam.align_to_sam(pkg, a.fasta, a.reads[0], a.reads[1], a.sam_file, a.num_reads_at_once, a.num_alignments, a.score_fraction_threshold,
                       not a.no_pseudo_assembly, a.sam_xa, a.min_alignment_score, " ".join(sys.argv), log=lambda m: print(m, file=sys.stderr),
                       db_dir=a.db)
print(st, file=sys.stderr)
