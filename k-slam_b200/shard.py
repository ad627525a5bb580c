"""Read-pair sharding across ranks ("DB fits one GPU").

Units are independent per read pair once the genome index is available (Overlap.h:175-197 never pairs
read x read; unique is per (read, genome); pairing is per pair, PairedOverlap.h:265-267). So a batch is split into
contiguous pair ranges that keep both mates together (R1 index i and R2 index i + mid,
PairedOverlap.h:247-256), every rank runs the whole path on its slice against a replicated index, and there is
NO collective on the data path — only a gather of results. Concatenating the per-rank results in rank order,
R1 blocks first, reproduces exactly the order a single GPU (and the reference) would produce.

The compute itself is pluggable (`align_fn`): the product passes Aligner.align_batch / pair_batch; the CPU
gloo tests pass the oracle so the sharding / gather / merge logic is covered without a GPU.
"""
from __future__ import annotations

import numpy as np


def pair_range(n_pairs: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) of pair indices for `rank`."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_reads(bases: np.ndarray, offs: np.ndarray, lo: int, hi: int):
    """Sub-batch holding pairs [lo, hi): its R1 block then its R2 block (FASTQsequence.h:110-123 layout)."""
    n = len(offs) - 1
    mid = n // 2
    offs = np.asarray(offs, dtype=np.uint64)
    parts, lens = [], []
    for a, b in ((lo, hi), (mid + lo, mid + hi)):
        parts.append(bases[int(offs[a]):int(offs[b])])
        lens.append((offs[a + 1:b + 1] - offs[a:b]).astype(np.uint64))
    sub_bases = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    sub_offs = np.zeros(2 * (hi - lo) + 1, dtype=np.uint64)
    sub_offs[1:] = np.cumsum(np.concatenate(lens))
    return np.ascontiguousarray(sub_bases), sub_offs


def globalize_reads(read_local: np.ndarray, lo: int, cnt: int, mid: int) -> np.ndarray:
    """Local read index (R1: [0,cnt), R2: [cnt,2cnt)) -> index in the whole batch."""
    r = read_local.astype(np.int64)
    return np.where(r < cnt, r + lo, r - cnt + mid + lo).astype(np.uint32)


def merge_alignments(parts, ranges, n_pairs: int, cigar_cap: int = 0, moved=None):
    """parts[r] = (overlaps, cigar_pool) of rank r (local read ids); returns the batch-order arrays.

    Global order is (read, entry, rel) with all R1 reads before all R2 reads, so the R1 segments of every rank
    come first (rank order), then the R2 segments. `moved` (a list, one entry per rank) receives, per rank, the pair
    (old cigar_off, new cigar_off) of every alignment, so that other records pointing into a rank's pool (the pair-sorted
    overlaps) can follow the CIGARs into the merged pool."""
    mid = n_pairs
    segs_ov, segs_cg, base = [], [], 0
    if moved is not None:
        moved[:] = [([], []) for _ in parts]
    for want_r2 in (False, True):
        for rank, ((ov, pool), (lo, hi)) in enumerate(zip(parts, ranges)):
            cnt = hi - lo
            split = int(np.searchsorted(ov["read"], cnt, side="left"))
            seg = ov[split:] if want_r2 else ov[:split]
            seg = seg.copy()
            seg["read"] = globalize_reads(seg["read"], lo, cnt, mid)
            if len(pool) and len(seg):
                # gather every CIGAR of the segment (cigar_off / cigar_len index the rank's pool, strided or dense)
                lens = seg["cigar_len"].astype(np.int64)
                starts = np.cumsum(lens) - lens
                idx = np.repeat(seg["cigar_off"].astype(np.int64) - starts, lens) + np.arange(int(lens.sum()), dtype=np.int64)
                segs_cg.append(pool[idx])
                if moved is not None:
                    moved[rank][0].append(seg["cigar_off"].astype(np.int64)); moved[rank][1].append(starts + base)
                seg["cigar_off"] = (starts + base).astype(np.uint32)
                base += int(lens.sum())
            segs_ov.append(seg)
    ov = np.concatenate(segs_ov) if segs_ov else np.zeros(0, dtype=parts[0][0].dtype)
    pool = np.concatenate(segs_cg) if segs_cg else np.zeros(0, dtype=np.uint32)
    return ov, pool


def merge_pairs(parts, ranges, n_pairs: int, moved=None):
    """parts[r] = (sorted_overlaps, pairs) of rank r; pair ids are contiguous per rank so rank order is the
    global (pair id, entry, rel) order; r1_idx / r2_idx are rebased onto the concatenated overlap array. With `moved`
    (from merge_alignments) the cigar_off of every pair-sorted overlap is rebased onto the merged CIGAR pool too; without
    it they keep indexing the rank's own pool."""
    mid = n_pairs
    ovs, prs, base = [], [], 0
    for rank, ((so, pr), (lo, hi)) in enumerate(zip(parts, ranges)):
        so = so.copy(); pr = pr.copy()
        if moved is not None and len(so) and moved[rank][0]:
            old = np.concatenate(moved[rank][0]); new = np.concatenate(moved[rank][1])
            order = np.argsort(old, kind="stable")
            old, new = old[order], new[order]
            has = so["cigar_len"] > 0                      # (records without a CIGAR carry no meaningful offset)
            at = np.searchsorted(old, so["cigar_off"][has].astype(np.int64))
            assert (at < len(old)).all() and (old[at] == so["cigar_off"][has]).all(), "a pair-sorted overlap points outside its rank's CIGARs"
            off = so["cigar_off"].copy(); off[has] = new[at].astype(np.uint32); so["cigar_off"] = off
        so["read"] = globalize_reads(so["read"], lo, hi - lo, mid)
        for f in ("r1_idx", "r2_idx"):
            pr[f] = np.where(pr[f] >= 0, pr[f] + base, -1)
        base += len(so)
        ovs.append(so); prs.append(pr)
    return np.concatenate(ovs), np.concatenate(prs)


def align_sharded(align_fn, bases, offs, rank: int, world: int, gather_fn):
    """Run `align_fn(sub_bases, sub_offs) -> (overlaps, cigar_pool, sorted_overlaps, pairs)` on this rank's slice
    and gather everything with `gather_fn(obj) -> list over ranks`. Every rank returns the merged result."""
    n_pairs = (len(offs) - 1) // 2
    ranges = [pair_range(n_pairs, world, r) for r in range(world)]
    lo, hi = ranges[rank]
    sub_b, sub_o = slice_reads(bases, offs, lo, hi)
    ov, pool, so, pr = align_fn(sub_b, sub_o)
    cap = (len(pool) // max(1, len(ov))) if len(ov) and len(pool) else 0
    gathered = gather_fn((ov, pool, so, pr, cap))
    cap = max(g[4] for g in gathered)
    moved = []
    m_ov, m_pool = merge_alignments([(g[0], g[1]) for g in gathered], ranges, n_pairs, cap, moved)
    m_so, m_pr = merge_pairs([(g[2], g[3]) for g in gathered], ranges, n_pairs, moved)      # (m_so's CIGARs index m_pool as well)
    return m_ov, m_pool, m_so, m_pr
