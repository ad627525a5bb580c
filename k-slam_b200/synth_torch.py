"""The synthetic workloads of synth.py generated with torch on whatever device is at hand (bench.py: the GPU).

Same shapes and distributions as the numpy generators, not the same random streams: 10 M read pairs and
1 Gbp of related genomes take ~90 s in numpy on the bench box's host cores and ~2 s here, and that time is better spent
measuring. Results come back as host numpy arrays in the layout the C ABI takes (one byte array + n+1 offsets); pass
pin=True for a page-locked copy (the e2e leg hands plain host pointers to the library).
"""
from __future__ import annotations

import numpy as np
import torch

_ASCII = (65, 67, 71, 84)  # A C G T


def _device(device=None):
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _gen(seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    return g


def _acgt(shape, g, dev):
    """iid uniform bases as ASCII bytes"""
    c = torch.randint(0, 4, shape, generator=g, device=dev, dtype=torch.uint8)
    return (65 + 2 * c + 2 * (c == 2).to(torch.uint8) + 13 * (c == 3).to(torch.uint8)).to(torch.uint8)   # 0 -> A(65) 1 -> C(67) 2 -> G(71) 3 -> T(84)


def _host(t, pin=False):
    if t.device.type == "cpu":
        out = t
    else:
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=pin)
        out.copy_(t)
    return out.numpy()


def _comp(x):
    """complement of upper-case ACGT bytes (A<->T, C<->G)"""
    return torch.where(x == 65, 84, torch.where(x == 84, 65, torch.where(x == 67, 71, torch.where(x == 71, 67, x)))).to(torch.uint8)


def random_genomes(n_genomes, length, seed=1, device=None, keep_device=False):
    dev = _device(device)
    bases = _acgt((n_genomes * length,), _gen(seed, dev), dev)
    offs = np.arange(n_genomes + 1, dtype=np.uint64) * np.uint64(length)
    return (bases if keep_device else _host(bases)), offs


def tree_genomes(n_strains=500, length=2_000_000, seed=1, divergence=(0.25, 0.15, 0.05, 0.01), device=None, keep_device=False):
    """root -> phyla -> genera -> species -> strains; every child = its parent with a fraction of positions redrawn."""
    dev = _device(device)
    g = _gen(seed, dev)

    def mutate(parent, d):
        m = torch.rand(length, generator=g, device=dev) < d
        return torch.where(m, _acgt((length,), g, dev), parent)

    root = _acgt((length,), g, dev)
    n_species = max(1, n_strains // 5); n_genera = max(1, n_species // 4); n_phyla = max(1, n_genera // 5)
    phyla = [mutate(root, divergence[0]) for _ in range(n_phyla)]
    genera = [mutate(phyla[i % n_phyla], divergence[1]) for i in range(n_genera)]
    species = [mutate(genera[i % n_genera], divergence[2]) for i in range(n_species)]
    out = torch.empty((n_strains, length), dtype=torch.uint8, device=dev)
    for i in range(n_strains):
        out[i] = mutate(species[i % n_species], divergence[3])
    offs = np.arange(n_strains + 1, dtype=np.uint64) * np.uint64(length)
    out = out.reshape(-1)
    return (out if keep_device else _host(out)), offs


def paired_reads(gen_bases, gen_offs, n_pairs, read_len=150, seed=2, sub_rate=0.01, indel_frac=0.05, frag_mean=350.0, frag_sd=35.0,
                 chunk=500_000, device=None, pin=False, out=None):
    """FR pairs as synth.paired_reads: fragment ~ round(N(350, 35)) in [read_len + 1, 600], uniform genome / position /
    strand, 1 % substitutions (always to another base), 5 % of the reads with one 1-3 bp indel at 40-110.
    gen_bases may be a device tensor (tree_genomes(..., keep_device=True)) or a host array. out: a (2, n_pairs, read_len) uint8
    host tensor to fill (a page-locked buffer that is reused from batch to batch: allocating one costs ~1 s per GB)."""
    dev = _device(device)
    g = _gen(seed, dev)
    G = gen_bases if isinstance(gen_bases, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(gen_bases))
    G = G.to(dev)
    offs = torch.from_numpy(gen_offs.astype(np.int64)).to(dev)
    glen = offs[1:] - offs[:-1]
    n_gen = len(gen_offs) - 1
    L, pad = read_len, 4
    if out is None:
        out = torch.empty((2, n_pairs, L), dtype=torch.uint8, pin_memory=pin and dev.type == "cuda")
    col = torch.arange(L, device=dev, dtype=torch.int64)[None, :]
    for c0 in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - c0)
        gi = torch.randint(0, n_gen, (n,), generator=g, device=dev)
        frag = torch.clamp(torch.round(torch.randn(n, generator=g, device=dev) * frag_sd + frag_mean), L + 1, 600).to(torch.int64)
        frag = torch.minimum(frag, glen[gi] - pad)
        pos = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (glen[gi] - frag - pad + 1).to(torch.float64)).to(torch.int64)
        base = offs[gi] + pos
        last = (offs[gi + 1] - 1)[:, None]
        mates = []
        for mate in range(2):
            start = base if mate == 0 else base + frag - L
            has = torch.rand(n, generator=g, device=dev) < indel_frac
            ipos = torch.randint(40, 111, (n,), generator=g, device=dev)[:, None]
            ilen = torch.randint(1, 4, (n,), generator=g, device=dev)[:, None]
            is_del = torch.rand(n, generator=g, device=dev) < 0.5
            dsel = (has & is_del)[:, None] & (col >= ipos)
            isel = (has & ~is_del)[:, None]
            idx = start[:, None] + col + dsel * ilen - isel * torch.clamp(col - ipos, 0).clamp(max=ilen)
            seq = G[torch.minimum(idx, last)]
            ins = isel & (col >= ipos) & (col < ipos + ilen)
            seq = torch.where(ins, _acgt((n, L), g, dev), seq)
            sub = torch.rand((n, L), generator=g, device=dev) < sub_rate
            # another base: rotate within ACGT by 1..3
            code = (seq == 67).to(torch.uint8) + 2 * (seq == 71).to(torch.uint8) + 3 * (seq == 84).to(torch.uint8)
            rot = (code + torch.randint(1, 4, (n, L), generator=g, device=dev, dtype=torch.uint8)) % 4
            seq = torch.where(sub, (65 + 2 * rot + 2 * (rot == 2).to(torch.uint8) + 13 * (rot == 3).to(torch.uint8)).to(torch.uint8), seq)
            mates.append(seq)
        left, right = mates[0], _comp(mates[1].flip(1))
        swap = (torch.rand(n, generator=g, device=dev) < 0.5)[:, None]
        out[0, c0:c0 + n].copy_(torch.where(swap, right, left))
        out[1, c0:c0 + n].copy_(torch.where(swap, left, right))
    offs_out = np.arange(2 * n_pairs + 1, dtype=np.uint64) * np.uint64(L)
    return out.numpy().reshape(-1), offs_out


def sw_pairs(n, read_len=150, window_len=150, seed=3, device=None, pin=False, keep_device=False):
    """Config 3 mix (synth.sw_pairs): 70 % ~1 % substitutions, 20 % one 1-5 bp indel, 5 % unrelated, 5 % with an N run; the
    read is planted at a uniform offset of the window."""
    dev = _device(device)
    g = _gen(seed, dev)
    W = _acgt((n, window_len), g, dev)
    off = torch.randint(0, window_len - read_len + 1, (n,), generator=g, device=dev)[:, None]
    kind = torch.rand(n, generator=g, device=dev)
    col = torch.arange(read_len, device=dev, dtype=torch.int64)[None, :]
    indel = (kind >= 0.70) & (kind < 0.90)
    lo = min(30, read_len // 3)
    ipos = torch.randint(lo, read_len - lo, (n,), generator=g, device=dev)[:, None]
    ilen = torch.randint(1, 6, (n,), generator=g, device=dev)[:, None]
    is_del = torch.rand(n, generator=g, device=dev) < 0.5
    idx = off + col + ((indel & is_del)[:, None] & (col >= ipos)) * ilen - (indel & ~is_del)[:, None] * torch.clamp(col - ipos, 0).clamp(max=ilen)
    Q = torch.gather(W, 1, idx.clamp(0, window_len - 1))
    ins = (indel & ~is_del)[:, None] & (col >= ipos) & (col < ipos + ilen)
    Q = torch.where(ins, _acgt((n, read_len), g, dev), Q)
    sub = torch.rand((n, read_len), generator=g, device=dev) < 0.01
    Q = torch.where(sub, _acgt((n, read_len), g, dev), Q)
    unrelated = ((kind >= 0.90) & (kind < 0.95))[:, None]
    Q = torch.where(unrelated, _acgt((n, read_len), g, dev), Q)
    withn = kind >= 0.95
    in_read = torch.rand(n, generator=g, device=dev) < 0.5
    for arr, ln, sel in ((Q, read_len, withn & in_read), (W, window_len, withn & ~in_read)):
        st = torch.randint(0, ln - 10, (n,), generator=g, device=dev)[:, None]
        run = torch.randint(1, 10, (n,), generator=g, device=dev)[:, None]
        c = torch.arange(ln, device=dev, dtype=torch.int64)[None, :]
        arr.masked_fill_(sel[:, None] & (c >= st) & (c < st + run), 78)   # 'N'
    qoffs = np.arange(n + 1, dtype=np.uint64) * np.uint64(read_len)
    roffs = np.arange(n + 1, dtype=np.uint64) * np.uint64(window_len)
    if keep_device:
        return Q.reshape(-1), qoffs, W.reshape(-1), roffs
    return _host(Q.reshape(-1), pin), qoffs, _host(W.reshape(-1), pin), roffs
