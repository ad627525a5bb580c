"""GPU tier: the k-mer-range partitioned path (include/kslam.h "partitioned") on ONE GPU.

`world` logical ranks run as threads of this process, each with its own kslam_ctx holding one key range of the genome
k-mer list; the two all-to-alls are device-to-device copies (LoopbackExchange). Every rank's alignments, CIGARs and
pair records must be bit-identical to the oracle on that rank's reads, and the key ranges must tile the sorted
genome k-mer list exactly."""
import threading

import numpy as np
import pytest

import _lib as T
from test_gpu_parity import FIELDS, check_overlaps

pytestmark = pytest.mark.gpu


def run_partitioned(pkg, gb, go, reads, world, report_cigar=True):
    """reads[r] = (bases, offs) of logical rank r. Returns per-rank (Alignments, Pairs, genome k-mer tap, partition, stats)."""
    import torch
    from kslam_b200 import dist as kd
    grp = kd.LoopbackGroup(world)
    out, errs = [None] * world, []

    def run(rank):
        try:
            torch.cuda.set_device(0)
            with pkg.Aligner(report_cigar=report_cigar) as al:
                al.load_genomes_part(gb, go, rank, world)
                sb, so = reads[rank]
                al.upload_reads(sb, so)
                res, stats = kd.align_partitioned(kd.CudaEngine(al, 0), grp.exchange(rank), len(so) - 1)
                pairs = al.pair_batch()
                out[rank] = (res, pairs, al.genome_kmers(), al.partition(), stats, al.seeds())
        except Exception as e:   # noqa: BLE001
            errs.append(e); grp.barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    return out


@pytest.mark.parametrize("world", [1, 2, 3])
def test_partitioned_matches_oracle(pkg, world):
    from kslam_b200 import shard
    n_pairs = 600
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=41, n_genomes=8, glen=8000, n_pairs=n_pairs)
    reads = [shard.slice_reads(rb, ro, *shard.pair_range(n_pairs, world, r)) for r in range(world)]
    out = run_partitioned(pkg, gb, go, reads, world)
    P = T.default_params(report_cigar=1)
    # the key ranges tile the reference's sorted genome list (KMer.h:388-398) exactly
    gk_all = np.concatenate([o[2] for o in out])
    want_gk = T.ko_sort_kmers(T.ko_extract(gb, go, True, 16))
    assert np.array_equal(gk_all, want_gk)
    spl = out[0][3]["splitters"]
    for r, o in enumerate(out):
        assert np.array_equal(o[3]["splitters"], spl), "ranks disagree on the splitters"
        k = o[2]["kmer"]
        assert (k >= spl[r]).all() and (r == world - 1 or (k < spl[r + 1]).all())
    assert sum(o[4]["kmers_sent"] for o in out) == sum(o[4]["kmers_received"] for o in out)
    for r, (res, pairs, _, _, stats, seeds) in enumerate(out):
        sb, so = reads[r]
        want = T.ko_pipeline(gb, go, sb, so, P)
        assert np.array_equal(seeds, want["seeds"])
        check_overlaps(res.overlaps, res.cigar_pool, want["overlaps"], want["cigar_pool"])
        check_overlaps(pairs.sorted_overlaps, pairs.cigar_pool, want["pair_sorted_overlaps"], want["cigar_pool"], cigars=False)
        assert np.array_equal(pairs.pairs, want["pairs"])


def test_partitioned_equals_replicated_at_volume(pkg):
    """20k pairs against 16 x 300 kbp genomes over 4 key ranges: identical to the replicated-index path."""
    from kslam_b200 import shard
    world, n_pairs = 4, 20000
    gb, go = pkg.synth.random_genomes(16, 300_000, seed=5)
    rb, ro, _ = pkg.synth.paired_reads(gb, go, n_pairs, seed=6)
    reads = [shard.slice_reads(rb, ro, *shard.pair_range(n_pairs, world, r)) for r in range(world)]
    out = run_partitioned(pkg, gb, go, reads, world, report_cigar=False)
    sizes = [len(o[2]) for o in out]
    assert max(sizes) < 1.2 * (sum(sizes) / world), f"unbalanced key ranges {sizes}"      # sampled quantiles balance
    with pkg.Aligner(report_cigar=False) as al:
        al.load_genomes(gb, go)
        for r in range(world):
            sb, so = reads[r]
            want = al.align_batch(sb, so)
            wp = al.pair_batch()
            got, gp = out[r][0], out[r][1]
            assert len(want.overlaps) > 1000
            for f in FIELDS:
                assert np.array_equal(got.overlaps[f], want.overlaps[f]), (r, f)
            assert np.array_equal(gp.pairs, wp.pairs)


def test_partitioned_empty_rank_and_errors(pkg):
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=42, n_genomes=6, glen=5000, n_pairs=100)
    empty = (np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    out = run_partitioned(pkg, gb, go, [(rb, ro), empty], 2)
    want = T.ko_pipeline(gb, go, rb, ro, T.default_params(report_cigar=1))
    check_overlaps(out[0][0].overlaps, out[0][0].cigar_pool, want["overlaps"], want["cigar_pool"])
    assert len(out[1][0].overlaps) == 0 and len(out[1][1].pairs) == 0
    with pkg.Aligner() as al:
        with pytest.raises(pkg.KslamError):
            al.load_genomes_part(gb, go, 2, 2)          # part out of range
        with pytest.raises(pkg.KslamError):
            al.load_genomes_part(gb, go, 0, 65)         # more than 64 parts
        al.load_genomes_part(gb, go, 0, 2)
        with pytest.raises(pkg.KslamError):
            al.part_route_kmers(0)                      # no reads uploaded


def test_comm_single_rank_equals_unpartitioned(pkg):
    """kslam_comm (the exchanges issued by the C++ host over NCCL, csrc/comm.cu) with one rank: every NCCL call of the
    protocol runs (all-gathers, send / recv to self) and the result equals kslam_align_batch + the oracle."""
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=43, n_genomes=8, glen=8000, n_pairs=500)
    want = T.ko_pipeline(gb, go, rb, ro, T.default_params(report_cigar=1))
    with pkg.Aligner(report_cigar=True) as al:
        al.load_genomes_part(gb, go, 0, 1)
        comm = pkg.Comm.init_rank(al, 0, 1, pkg.Comm.unique_id())
        for _ in range(2):                                   # a second batch reuses every buffer
            al.upload_reads(rb, ro)
            res = comm.align_resident()
            pairs = al.pair_batch()
            check_overlaps(res.overlaps, res.cigar_pool, want["overlaps"], want["cigar_pool"])
            assert np.array_equal(pairs.pairs, want["pairs"])
        st = comm.stats()
        assert st["kmers_sent"] == st["kmers_received"] > 0 and st["matches_sent"] == st["matches_received"] > 0
        assert st["bytes_sent_kmers"] == 0                   # nothing leaves the only GPU
        comm.close()


def test_comm_failed_rank_ends_the_batch_for_everyone(pkg):
    """A rank that cannot take part must not leave the others waiting in a collective: with one rank, a call before any
    upload and an aborted batch both return errors and the communicator stays usable for the next batch."""
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=44, n_genomes=6, glen=6000, n_pairs=200)
    want = T.ko_pipeline(gb, go, rb, ro, T.default_params(report_cigar=1))
    with pkg.Aligner(report_cigar=True) as al:
        al.load_genomes_part(gb, go, 0, 1)
        comm = pkg.Comm.init_rank(al, 0, 1, pkg.Comm.unique_id())
        with pytest.raises(pkg.KslamError, match="kslam_upload_reads first"):
            comm.align_resident()                            # took part in the first gather, then reported its own state
        comm.abort_batch()
        al.upload_reads(rb, ro)
        res = comm.align_resident()
        check_overlaps(res.overlaps, res.cigar_pool, want["overlaps"], want["cigar_pool"])
        comm.close()


def test_comm_two_gpus_abort_reaches_the_peer(pkg):
    """Two ranks in one process: rank 1 aborts the batch, rank 0's kslam_comm_align_resident returns an error naming it
    (instead of waiting for ever); the next batch runs normally on both."""
    import torch
    from kslam_b200 import shard
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n_pairs = 2, 2000
    gb, go = pkg.synth.random_genomes(6, 100_000, seed=17)
    rb, ro, _ = pkg.synth.paired_reads(gb, go, n_pairs, seed=18)
    reads = [shard.slice_reads(rb, ro, *shard.pair_range(n_pairs, world, r)) for r in range(world)]
    als = [pkg.Aligner(report_cigar=True, device=r) for r in range(world)]
    for r, al in enumerate(als):
        al.load_genomes_part(gb, go, r, world)
    comms = pkg.Comm.init_all(als)
    seen, out = [None] * world, [None] * world

    def run(r):
        als[r].upload_reads(*reads[r])
        try:                                                 # batch 1: rank 1 cannot take part
            if r == 1:
                comms[r].abort_batch()
            else:
                comms[r].align_resident()
        except pkg.KslamError as e:
            seen[r] = str(e)
        out[r] = comms[r].align_resident()                   # batch 2: both ranks
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join(timeout=120) for t in th]
    assert not any(t.is_alive() for t in th), "a rank is still waiting in a collective"
    assert seen[0] and "rank 1" in seen[0] and seen[1] is None
    P = T.default_params(report_cigar=1)
    for r in range(world):
        want = T.ko_pipeline(gb, go, *reads[r], P)
        check_overlaps(out[r].overlaps, out[r].cigar_pool, want["overlaps"], want["cigar_pool"])
    [c.close() for c in comms]; [a.close() for a in als]


def test_comm_two_gpus_one_process(pkg):
    """kslam_comm_init_all: one process, one ctx per GPU, one host thread per rank (what `SLAM --devices 0,1` does)."""
    import torch
    from kslam_b200 import shard
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n_pairs = 2, 6000
    gb, go = pkg.synth.random_genomes(12, 200_000, seed=7)
    rb, ro, _ = pkg.synth.paired_reads(gb, go, n_pairs, seed=8)
    reads = [shard.slice_reads(rb, ro, *shard.pair_range(n_pairs, world, r)) for r in range(world)]
    als = [pkg.Aligner(report_cigar=True, device=r) for r in range(world)]
    for r, al in enumerate(als):
        al.load_genomes_part(gb, go, r, world)
    comms = pkg.Comm.init_all(als)
    out, errs = [None] * world, []

    def run(r):
        try:
            als[r].upload_reads(*reads[r])
            res = comms[r].align_resident()
            out[r] = (res, als[r].pair_batch(), comms[r].stats())
        except Exception as e:   # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    assert sum(o[2]["bytes_sent_kmers"] for o in out) > 0
    P = T.default_params(report_cigar=1)
    for r in range(world):
        want = T.ko_pipeline(gb, go, *reads[r], P)
        check_overlaps(out[r][0].overlaps, out[r][0].cigar_pool, want["overlaps"], want["cigar_pool"])
        assert np.array_equal(out[r][1].pairs, want["pairs"])
    [c.close() for c in comms]; [a.close() for a in als]


def test_partitioned_nccl_two_gpus():
    """Real exchange: one process per GPU, NCCL all_to_all_single over NVLink (skipped on a single-GPU box)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dist_nccl_worker.py")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", "29651", worker], check=True, timeout=900)
