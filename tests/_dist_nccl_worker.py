"""Worker for tests/test_gpu_dist.py::test_partitioned_nccl_two_gpus: one rank per GPU, NCCL all-to-all both ways.
Every rank compares the partitioned result on its reads with the replicated-index path on the same GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _lib as T  # noqa: E402

FIELDS = ["read", "entry", "rel", "rev_comp", "ref_begin", "ref_end", "query_begin", "query_end", "sw_score", "cigar_len"]


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = T.load_pkg()
    from kslam_b200 import dist as kd, shard
    n_pairs = 30000
    gb, go = pkg.synth.random_genomes(12, 400_000, seed=7)
    rb, ro, _ = pkg.synth.paired_reads(gb, go, n_pairs, seed=8)
    sb, so = shard.slice_reads(rb, ro, *shard.pair_range(n_pairs, world, rank))
    with pkg.Aligner(report_cigar=True, device=local) as al:
        al.load_genomes_part(gb, go, rank, world)
        al.upload_reads(sb, so)
        got, stats = kd.align_partitioned(kd.CudaEngine(al, local), kd.TorchExchange(device=torch.device("cuda", local)), len(so) - 1)
        gp = al.pair_batch()
    with pkg.Aligner(report_cigar=True, device=local) as al:
        al.load_genomes(gb, go)
        want = al.align_batch(sb, so)
        wp = al.pair_batch()
    assert len(want.overlaps) > 1000 and stats["kmers_sent"] > 0
    for f in FIELDS:
        assert np.array_equal(got.overlaps[f], want.overlaps[f]), (rank, f)
    assert T.cigars_of(got.overlaps, got.cigar_pool) == T.cigars_of(want.overlaps, want.cigar_pool)
    assert np.array_equal(gp.pairs, wp.pairs)
    print(f"rank {rank}: {len(got.overlaps)} alignments identical, exchange {stats}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
