"""CPU tier: the N>1 path (read-pair sharding, gather, merge) with world_size 2 over gloo. No collective sits on
the data path; the merged result must equal the single-process result record for record."""
import os
import subprocess
import sys

import numpy as np

import _lib as T

FIELDS = ["read", "entry", "rel", "rev_comp", "ref_begin", "ref_end", "query_begin", "query_end", "sw_score", "cigar_len"]


def test_pair_range_and_slices(pkg):
    from kslam_b200 import shard
    assert [shard.pair_range(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert [shard.pair_range(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    seqs = [bytes([65 + i]) * (i + 1) for i in range(8)]          # 4 pairs, ragged
    b, o = T.concat([np.frombuffer(s, np.uint8) for s in seqs])
    sb, so = shard.slice_reads(b, o, 1, 3)
    got = [bytes(sb[int(so[i]):int(so[i + 1])]) for i in range(4)]
    assert got == [seqs[1], seqs[2], seqs[5], seqs[6]]
    r = shard.globalize_reads(np.array([0, 1, 2, 3], np.uint32), lo=1, cnt=2, mid=4)
    assert r.tolist() == [1, 2, 5, 6]


def test_world2_gloo_matches_single_process(pkg, tmp_path):
    out = str(tmp_path / "merged.npz")
    env = dict(os.environ, KSLAM_DIST_OUT=out, OMP_NUM_THREADS="2")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dist_worker.py")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", "29631", worker], check=True, env=env, timeout=600)
    got = np.load(out)
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=21, n_genomes=8, glen=8000, n_pairs=901)
    want = T.ko_pipeline(gb, go, rb, ro, T.default_params(report_cigar=1), cigar_cap=32)
    for f in FIELDS:
        assert np.array_equal(got["ov"][f], want["overlaps"][f]), f
    assert T.cigars_of(got["ov"], got["pool"]) == T.cigars_of(want["overlaps"], want["cigar_pool"])
    for f in FIELDS:
        assert np.array_equal(got["so"][f], want["pair_sorted_overlaps"][f]), f
    # the pair-sorted overlaps follow their CIGARs into the merged pool (their cigar_off is rebased, not left rank-local)
    assert T.cigars_of(got["so"], got["pool"]) == T.cigars_of(want["pair_sorted_overlaps"], want["cigar_pool"])
    assert np.array_equal(got["pr"], want["pairs"])
