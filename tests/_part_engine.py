"""CPU stand-in for the partitioned path's ENGINE (k-slam_b200/dist.py) built on the oracle — TEST INFRASTRUCTURE.

It runs the same protocol (route by key owner -> join -> route matches back -> finish) with numpy buffers so that the
host logic of the N>1 path (splitters, ownership, id bases, all-to-all bookkeeping) is covered without a GPU. The
k-mer extraction and the seed sort / unique come from oracle/ (KMer.h:160-181, Overlap.h:277-295); the pile cross
product that emits raw matches restates Overlap.h:175-197 in numpy.
"""
import numpy as np

import _lib as T

REC = np.dtype([("a", "<u8"), ("b", "<u8")])       # 16-byte wire record: k-mer record or {read val, genome val}


def sample_splitters(genome_kmers, world):
    """Quantiles of the sorted genome k-mers (the CUDA side samples; any monotone splitters are valid)."""
    keys = np.sort(genome_kmers["kmer"])
    spl = np.zeros(world + 1, dtype=np.uint64)
    for p in range(1, world):
        spl[p] = keys[p * len(keys) // world]
    spl[world] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return spl


class OracleEngine:
    def __init__(self, dist_mod, gb, go, rb, ro, rank, world, splitters=None):
        self.d = dist_mod
        self.rank, self.world = rank, world
        gk = T.ko_extract(gb, go, True, 16)
        self.splitters = sample_splitters(gk, world) if splitters is None else splitters
        own = self.d.key_owner(gk["kmer"], self.splitters) == rank
        g = gk[own]
        order = np.lexsort((-(g["id_flags"].astype(np.int64)), g["kmer"]))     # KMer.h:392-396
        self.g = g[order]
        self.rb, self.ro = T.u8(rb), np.ascontiguousarray(ro, dtype=np.uint64)
        self.read_lens = (self.ro[1:] - self.ro[:-1]).astype(np.uint32)

    # -- read owner
    def route_kmers(self, id_base):
        rk = T.ko_extract(self.rb, self.ro, False, 1)
        rk = rk[rk["kmer"] != 0]                                 # Overlap.h:236-239 (the CUDA prefilter drops them too)
        rk["id_flags"] = rk["id_flags"] + np.uint32(id_base)
        owner = self.d.key_owner(rk["kmer"], self.splitters)
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=self.world).astype(np.uint64)
        return np.ascontiguousarray(rk[order]).view(np.uint8), counts

    def recv_buffer(self, n):
        self._recv = np.zeros(n * 16, dtype=np.uint8)
        return self._recv

    # -- key owner
    def join(self, n_recv, id_bases):
        r = self._recv[:n_recv * 16].view(T.KMER_DT)
        g = self.g
        lo = np.searchsorted(g["kmer"], r["kmer"], side="left")
        hi = np.searchsorted(g["kmer"], r["kmer"], side="right")
        cnt = hi - lo
        ridx = np.repeat(np.arange(len(r)), cnt)
        gidx = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)]) if len(r) else np.zeros(0, np.int64)
        m = np.zeros(len(ridx), dtype=REC)
        m["a"] = r["id_flags"][ridx].astype(np.uint64) | (r["offset"][ridx].astype(np.uint64) << np.uint64(32))
        m["b"] = g["id_flags"][gidx].astype(np.uint64) | (g["offset"][gidx].astype(np.uint64) << np.uint64(32))
        owner = self.d.read_owner(m["a"] & np.uint64(0x3FFFFFFF), id_bases)
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=self.world).astype(np.uint64)
        return np.ascontiguousarray(m[order]).view(np.uint8), counts

    def match_buffer(self, n):
        self._mrecv = np.zeros(n * 16, dtype=np.uint8)
        return self._mrecv

    # -- read owner again
    def finish(self, n_m, id_base, fetch):
        m = self._mrecv[:n_m * 16].view(REC)
        idf = (m["a"] & np.uint64(0xFFFFFFFF)).astype(np.uint32); r_off = (m["a"] >> np.uint64(32)).astype(np.uint32)
        gf = (m["b"] & np.uint64(0xFFFFFFFF)).astype(np.uint32); g_off = (m["b"] >> np.uint64(32)).astype(np.uint32)
        rid = (idf & np.uint32(0x3FFFFFFF)) - np.uint32(id_base)
        r_rc = (idf >> np.uint32(30)) & np.uint32(1); g_rc = (gf >> np.uint32(30)) & np.uint32(1)
        rlen = self.read_lens[rid]
        off = np.where(g_rc == 1, rlen - r_off - np.uint32(32), r_off).astype(np.uint32)     # Overlap.h:185-189
        seeds = np.zeros(n_m, dtype=T.SEED_DT)
        seeds["read"] = rid; seeds["entry"] = gf & np.uint32(0x3FFFFFFF)
        seeds["rel"] = (g_off - off).astype(np.uint32).view(np.int32)
        seeds["rev_comp"] = (g_rc != r_rc).astype(np.uint32)
        return T.ko_sort_unique(seeds)
