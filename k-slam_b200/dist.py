"""k-mer-range partitioned matching across ranks (SURVEY.md §8e "DB exceeds one GPU", BASELINE config 4).

The sorted genome k-mer list (KMer.h:388-398) is cut into `world` contiguous kMerInt ranges, one per GPU; genome
bases are replicated. Equal k-mers always share an owner, so no pile (Overlap.h:153-199) is split — the invariant
the reference's own chunking keeps (Overlap.h:285-287). One batch costs ONE exchange each way:

    read owner   route_kmers : extract + prefilter read k-mers (job-global read ids), group by key owner
                   ---- all-to-all (16 B k-mer records) ---->
    key owner    join        : radix sort, merge-join against the local range, group raw matches by read owner
                   <--- all-to-all (16 B match records) -----
    read owner   finish      : match -> seed, seed sort, fuzzy unique, Smith-Waterman (the single-GPU path unchanged)

`align_partitioned` is that protocol, written once against two small interfaces:
  * an ENGINE with route_kmers / recv_buffer / join / match_buffer / finish — the product's is `CudaEngine` (the C ABI of
    include/kslam.h, device pointers in ctx-owned HBM); the CPU tests plug in an engine built on the oracle so the
    routing logic is covered with gloo and no GPU;
  * an EXCHANGE with allgather_int / all_to_all_records — `TorchExchange` (torch.distributed: NCCL over NVLink on
    GPUs, gloo on CPU) or `LoopbackExchange` (several logical ranks as threads of one process, e.g. several contexts on
    ONE GPU, which is how the single-GPU test tier exercises the partitioned kernels).
There is no CPU fallback on the product side: CudaEngine fails without libkslam.so and an sm_100 device.
"""
from __future__ import annotations

import threading

import numpy as np

REC_BYTES = 16


def id_bases_of(n_reads_per_rank):
    """First job-global read id of every rank (+ the end): KMerData keeps 30 bits of id (KMer.h:65-66)."""
    b = np.zeros(len(n_reads_per_rank) + 1, dtype=np.uint64)
    b[1:] = np.cumsum(np.asarray(n_reads_per_rank, dtype=np.uint64))
    if int(b[-1]) > (1 << 30):
        raise ValueError(f"{int(b[-1])} reads in one job-wide batch exceed the 2^30 ids of a k-mer record")
    return b.astype(np.uint32)


def key_owner(kmers: np.ndarray, splitters: np.ndarray) -> np.ndarray:
    """Rank owning each k-mer: the last p with splitters[p] <= kmer (splitters[0] = 0). Host mirror of dist.cu's
    bucket_of<0>, used by the CPU engine and the tests."""
    inner = np.asarray(splitters[1:-1], dtype=np.uint64)
    return np.searchsorted(inner, np.asarray(kmers, dtype=np.uint64), side="right").astype(np.int64)


def read_owner(global_ids: np.ndarray, id_bases: np.ndarray) -> np.ndarray:
    inner = np.asarray(id_bases[1:-1], dtype=np.uint64)
    return np.searchsorted(inner, np.asarray(global_ids, dtype=np.uint64), side="right").astype(np.int64)


def align_partitioned(engine, exch, n_reads: int, fetch=True):
    """One batch through the partitioned path on this rank. The engine already holds this rank's reads
    (upload_reads) and key range (load_genomes_part). Returns what engine.finish returns."""
    id_bases = id_bases_of(exch.allgather_int(n_reads))
    my_base = int(id_bases[exch.rank])
    send, counts = engine.route_kmers(my_base)
    recv_counts = exch.exchange_counts(counts)
    n_recv = int(recv_counts.sum())
    recv = engine.recv_buffer(n_recv)
    exch.all_to_all_records(send, counts, recv, recv_counts)
    msend, mcounts = engine.join(n_recv, id_bases)
    mrecv_counts = exch.exchange_counts(mcounts)
    n_m = int(mrecv_counts.sum())
    mrecv = engine.match_buffer(n_m)
    exch.all_to_all_records(msend, mcounts, mrecv, mrecv_counts)
    stats = dict(kmers_sent=int(np.sum(counts)), kmers_received=n_recv, matches_sent=int(np.sum(mcounts)), matches_received=n_m)
    return engine.finish(n_m, my_base, fetch), stats


# ----------------------------------------------------------------------------------------------- engines
class CudaEngine:
    """The product engine: one kslam_ctx (Aligner) holding this rank's key range; buffers are device pointers."""

    def __init__(self, aligner, device):
        self.al = aligner
        self.device = device

    def _tensor(self, ptr, n_records):
        import torch
        nbytes = int(n_records) * REC_BYTES
        if nbytes == 0 or not ptr:
            return torch.empty(0, dtype=torch.uint8, device=f"cuda:{self.device}")

        class _Dev:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(_Dev(), device=f"cuda:{self.device}")

    def route_kmers(self, id_base):
        ptr, counts = self.al.part_route_kmers(id_base)
        return self._tensor(ptr, counts.sum()), counts

    def recv_buffer(self, n):
        return self._tensor(self.al.part_recv_buffer(n), n)

    def join(self, n_recv, id_bases):
        ptr, counts = self.al.part_join(n_recv, id_bases)
        return self._tensor(ptr, counts.sum()), counts

    def match_buffer(self, n):
        return self._tensor(self.al.part_match_buffer(n), n)

    def finish(self, n_m, id_base, fetch):
        return self.al.part_finish(n_m, id_base, fetch=fetch)


# ----------------------------------------------------------------------------------------------- exchanges
class TorchExchange:
    """all-to-all over torch.distributed (NCCL on GPUs: NVLink / NVSwitch; gloo on CPU for the tests).
    Buffers are flat uint8 torch tensors (device tensors for NCCL)."""

    def __init__(self, device=None, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device

    def _int_tensor(self, vals):
        import torch
        t = torch.tensor([int(v) for v in vals], dtype=torch.int64)
        return t.to(self.device) if self.device is not None else t

    def allgather_int(self, v):
        import torch
        mine = self._int_tensor([v])
        out = torch.empty(self.world, dtype=torch.int64, device=mine.device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group) if mine.is_cuda else \
            self.dist.all_gather(list(out.split(1)), mine, group=self.group)
        return [int(x) for x in out.cpu().tolist()]

    def exchange_counts(self, counts):
        import torch
        send = self._int_tensor(counts)
        recv = torch.empty_like(send)
        self.dist.all_to_all_single(recv, send, group=self.group)
        return recv.cpu().numpy().astype(np.uint64)

    def all_to_all_records(self, send, counts, recv, recv_counts):
        import torch
        if send.is_cuda:
            torch.cuda.current_stream().synchronize()
        self.dist.all_to_all_single(recv, send, output_split_sizes=[int(c) * REC_BYTES for c in recv_counts],
                                    input_split_sizes=[int(c) * REC_BYTES for c in counts], group=self.group)
        if send.is_cuda:
            torch.cuda.current_stream().synchronize()   # the library works on its own stream: finish before it reads


class LoopbackGroup:
    """`world` logical ranks inside one process (threads). Used to run the partitioned kernels on ONE GPU."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world

    def exchange(self, rank):
        return LoopbackExchange(self, rank)


class LoopbackExchange:
    def __init__(self, group, rank):
        self.g, self.rank, self.world = group, rank, group.world

    def _swap(self, item):
        self.g.slots[self.rank] = item
        self.g.barrier.wait()
        got = list(self.g.slots)
        self.g.barrier.wait()
        return got

    def allgather_int(self, v):
        return [int(x) for x in self._swap(int(v))]

    def exchange_counts(self, counts):
        allc = self._swap(np.asarray(counts, dtype=np.uint64))
        return np.array([allc[src][self.rank] for src in range(self.world)], dtype=np.uint64)

    def all_to_all_records(self, send, counts, recv, recv_counts):
        is_torch = hasattr(send, "is_cuda")
        self.g.slots[self.rank] = (send, np.asarray(counts, dtype=np.uint64))
        self.g.barrier.wait()
        off = 0
        for src in range(self.world):
            s_buf, s_counts = self.g.slots[src]
            start = int(s_counts[:self.rank].sum()) * REC_BYTES
            n = int(s_counts[self.rank]) * REC_BYTES
            assert n == int(recv_counts[src]) * REC_BYTES
            if n:
                if is_torch:
                    recv[off:off + n].copy_(s_buf[start:start + n])
                else:
                    recv[off:off + n] = s_buf[start:start + n]
            off += n
        if is_torch and recv.is_cuda:
            import torch
            torch.cuda.synchronize()
        self.g.barrier.wait()   # nobody reuses its send buffer before every peer has copied out of it
