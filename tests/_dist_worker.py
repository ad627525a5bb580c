"""Worker for tests/test_dist_gloo.py: world_size-2 gloo run of the read-pair sharding path on CPU.
The per-rank compute is the ORACLE (test infrastructure) so that shard / gather / merge are covered without a GPU."""
import os
import sys

import numpy as np
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _lib as T  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = T.load_pkg()
    from kslam_b200 import shard
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=21, n_genomes=8, glen=8000, n_pairs=901)   # odd count: uneven shards
    P = T.default_params(report_cigar=1)

    def align_fn(sb, so):
        r = T.ko_pipeline(gb, go, sb, so, P, cigar_cap=32)
        ov = r["overlaps"]
        return ov, r["cigar_pool"][:len(ov) * 32], r["pair_sorted_overlaps"], r["pairs"]

    def gather_fn(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    m_ov, m_pool, m_so, m_pr = shard.align_sharded(align_fn, rb, ro, rank, world, gather_fn)
    if rank == 0:
        np.savez(os.environ["KSLAM_DIST_OUT"], ov=m_ov, pool=m_pool, so=m_so, pr=m_pr)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
