"""Worker for tests/test_dist_part.py: world_size-2 gloo run of the k-mer-range partitioned protocol on CPU
(oracle-backed engine; the all-to-alls are real torch.distributed collectives)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _lib as T  # noqa: E402
from _part_engine import OracleEngine  # noqa: E402


class TorchCpuEngine(OracleEngine):
    """numpy buffers of the oracle engine seen as flat uint8 torch tensors (what TorchExchange moves)."""

    def route_kmers(self, id_base):
        b, c = super().route_kmers(id_base)
        return torch.from_numpy(b.copy()), c

    def recv_buffer(self, n):
        return torch.from_numpy(super().recv_buffer(n))

    def join(self, n, idb):
        b, c = super().join(n, idb)
        return torch.from_numpy(b.copy()), c

    def match_buffer(self, n):
        return torch.from_numpy(super().match_buffer(n))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = T.load_pkg()
    from kslam_b200 import dist as kd, shard
    gb, go, rb, ro = pkg.synth.adversarial_set(seed=33, n_genomes=6, glen=6000, n_pairs=401)
    lo, hi = shard.pair_range(401, world, rank)
    sb, so = shard.slice_reads(rb, ro, lo, hi)
    eng = TorchCpuEngine(kd, gb, go, sb, so, rank, world)
    seeds, stats = kd.align_partitioned(eng, kd.TorchExchange(), len(so) - 1)
    np.savez(os.environ["KSLAM_DIST_OUT"] + f".{rank}.npz", seeds=seeds, stats=np.array([stats[k] for k in sorted(stats)]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
