"""CPU tier: the k-mer-range partitioned path. The protocol of k-slam_b200/dist.py is run with
an oracle-backed engine over (a) in-process loopback ranks and (b) world_size-2 gloo; each rank's seeds must equal the
unpartitioned oracle's seeds for that rank's reads — partitioning the genome k-mer list by key range never changes a
result because equal k-mers share an owner (Overlap.h:285-287)."""
import os
import subprocess
import sys
import threading

import numpy as np

import _lib as T
from _part_engine import OracleEngine, sample_splitters


def test_ownership_helpers(pkg):
    from kslam_b200 import dist as kd
    spl = np.array([0, 10, 10, 50, 2**64 - 1], dtype=np.uint64)          # an empty range (10..10) is legal
    assert kd.key_owner(np.array([0, 9, 10, 49, 50, 2**64 - 1], np.uint64), spl).tolist() == [0, 0, 2, 2, 3, 3]
    idb = kd.id_bases_of([3, 0, 4])
    assert idb.tolist() == [0, 3, 3, 7]
    assert kd.read_owner(np.array([0, 2, 3, 6]), idb).tolist() == [0, 0, 2, 2]
    try:
        kd.id_bases_of([1 << 29, 1 << 29, 1])
        assert False, "2^30 id overflow not detected"
    except ValueError:
        pass


def _expected_seeds(gb, go, sb, so):
    return T.ko_pipeline(gb, go, sb, so, T.default_params(report_cigar=0), paired=False)["seeds"]


def test_loopback_three_ranks_matches_unpartitioned(pkg):
    from kslam_b200 import dist as kd, shard
    world = 3
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=31, n_genomes=6, glen=6000, n_pairs=300)
    grp = kd.LoopbackGroup(world)
    out, errs = [None] * world, []

    def run(rank):
        try:
            lo, hi = shard.pair_range(300, world, rank)
            sb, so = shard.slice_reads(rb, ro, lo, hi)
            eng = OracleEngine(kd, gb, go, sb, so, rank, world)
            seeds, stats = kd.align_partitioned(eng, grp.exchange(rank), len(so) - 1)
            out[rank] = (seeds, stats, _expected_seeds(gb, go, sb, so))
        except Exception as e:   # noqa: BLE001
            errs.append(e); grp.barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    sent = sum(o[1]["kmers_sent"] for o in out); recv = sum(o[1]["kmers_received"] for o in out)
    assert sent == recv and sent > 0
    assert sum(o[1]["matches_sent"] for o in out) == sum(o[1]["matches_received"] for o in out)
    for seeds, _, want in out:
        assert len(want) > 0 and np.array_equal(seeds, want)


def test_skewed_splitters_and_empty_rank(pkg):
    """All keys owned by one rank, and a rank without reads: the protocol must not care."""
    from kslam_b200 import dist as kd
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=32, n_genomes=6, glen=5000, n_pairs=120)
    world = 2
    spl = np.array([0, 2**64 - 1, 2**64 - 1], dtype=np.uint64)             # rank 1 owns (almost) nothing
    grp = kd.LoopbackGroup(world)
    empty_b, empty_o = np.zeros(0, np.uint8), np.zeros(1, np.uint64)
    reads = [(rb, ro), (empty_b, empty_o)]
    out, errs = [None] * world, []

    def run(rank):
        try:
            sb, so = reads[rank]
            eng = OracleEngine(kd, gb, go, sb, so, rank, world, splitters=spl)
            out[rank] = kd.align_partitioned(eng, grp.exchange(rank), len(so) - 1)[0]
        except Exception as e:   # noqa: BLE001
            errs.append(e); grp.barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    assert np.array_equal(out[0], _expected_seeds(gb, go, rb, ro))
    assert len(out[1]) == 0


def test_world2_gloo_partitioned(pkg, tmp_path):
    from kslam_b200 import shard
    out = str(tmp_path / "part")
    env = dict(os.environ, KSLAM_DIST_OUT=out, OMP_NUM_THREADS="2")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dist_part_worker.py")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", "29641", worker], check=True, env=env, timeout=600)
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=33, n_genomes=6, glen=6000, n_pairs=401)
    for rank in range(2):
        got = np.load(out + f".{rank}.npz")
        lo, hi = shard.pair_range(401, 2, rank)
        sb, so = shard.slice_reads(rb, ro, lo, hi)
        assert np.array_equal(got["seeds"], _expected_seeds(gb, go, sb, so))
        assert got["stats"].sum() > 0
