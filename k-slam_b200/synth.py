"""Seeded synthetic workloads of the shapes named in BASELINE.json / SURVEY.md §8(d).

Everything is produced as one concatenated uint8 byte array plus a uint64 offsets array
(n+1 entries) — the layout the C ABI (include/kslam.h) takes. No file I/O, no network.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def revcomp_rows(x: np.ndarray) -> np.ndarray:
    """Reverse-complement each row of a 2-D uint8 array (upper-case ACGT swapped, rest kept)."""
    return _COMP[x[:, ::-1]]


def random_genomes(n_genomes: int, length: int, seed: int = 1):
    """`n_genomes` iid-uniform ACGT genomes of `length` bp (config 1: 50 x 3 Mbp)."""
    rng = np.random.default_rng(seed)
    bases = ACGT[rng.integers(0, 4, size=n_genomes * length, dtype=np.uint8)]
    offs = np.arange(n_genomes + 1, dtype=np.uint64) * np.uint64(length)
    return bases, offs


def related_genomes(n_genomes: int, length: int, seed: int = 1, n_roots: int = 5, divergence: float = 0.03):
    """Config-2-like: genomes derived from a few roots by point mutation, so conserved k-mers
    create multi-genome piles."""
    rng = np.random.default_rng(seed)
    roots = ACGT[rng.integers(0, 4, size=(n_roots, length), dtype=np.uint8)]
    out = np.empty((n_genomes, length), dtype=np.uint8)
    for g in range(n_genomes):
        src = roots[g % n_roots].copy()
        mut = rng.random(length) < divergence
        src[mut] = ACGT[rng.integers(0, 4, size=int(mut.sum()), dtype=np.uint8)]
        out[g] = src
    offs = np.arange(n_genomes + 1, dtype=np.uint64) * np.uint64(length)
    return out.reshape(-1), offs


def tree_genomes(n_strains: int = 500, length: int = 2_000_000, seed: int = 1,
                 divergence=(0.25, 0.15, 0.05, 0.01)):
    """Config 2: strains in a synthetic phylogeny root -> 5 phyla -> 25 genera -> 100 species ->
    `n_strains` strains; every child is its parent with a fraction of positions redrawn (per-level `divergence`),
    so conserved 32-mers give multi-genome piles and reads seed on several related strains."""
    rng = np.random.default_rng(seed)

    def mutate(parent, d):
        child = parent.copy()
        m = rng.random(length) < d
        child[m] = ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        return child

    root = ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
    n_species = max(1, n_strains // 5); n_genera = max(1, n_species // 4); n_phyla = max(1, n_genera // 5)
    phyla = [mutate(root, divergence[0]) for _ in range(n_phyla)]
    genera = [mutate(phyla[i % n_phyla], divergence[1]) for i in range(n_genera)]
    species = [mutate(genera[i % n_genera], divergence[2]) for i in range(n_species)]
    out = np.empty((n_strains, length), dtype=np.uint8)
    for i in range(n_strains):
        out[i] = mutate(species[i % n_species], divergence[3])
    offs = np.arange(n_strains + 1, dtype=np.uint64) * np.uint64(length)
    return out.reshape(-1), offs


def paired_reads(gen_bases: np.ndarray, gen_offs: np.ndarray, n_pairs: int, read_len: int = 150,
                 seed: int = 2, sub_rate: float = 0.01, indel_frac: float = 0.05,
                 frag_mean: float = 350.0, frag_sd: float = 35.0, chunk: int = 200_000):
    """FR read pairs simulated from the genomes.

    Returns (bases, offs, truth) with reads laid out R1[0..n_pairs) then R2[0..n_pairs)
    (FASTQsequence.h:110-123 order). Fragment length ~ round(N(frag_mean, frag_sd)) clipped to
    [read_len+1, 600]; strand of the fragment uniform; per-base substitution `sub_rate`; with
    probability `indel_frac` a read carries one indel of 1-3 bp at position 40-110.
    """
    rng = np.random.default_rng(seed)
    n_gen = len(gen_offs) - 1
    glen = (gen_offs[1:] - gen_offs[:-1]).astype(np.int64)
    L = read_len
    r1 = np.empty((n_pairs, L), dtype=np.uint8)
    r2 = np.empty((n_pairs, L), dtype=np.uint8)
    truth = np.zeros(n_pairs, dtype=[("genome", "u4"), ("pos", "i8"), ("frag", "i4"), ("swapped", "u1")])
    pad = 4  # spare template bases so deletions still yield L read bases
    for c0 in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - c0)
        g = rng.integers(0, n_gen, size=n)
        frag = np.clip(np.rint(rng.normal(frag_mean, frag_sd, size=n)), L + 1, 600).astype(np.int64)
        frag = np.minimum(frag, glen[g] - pad)
        pos = (rng.random(n) * (glen[g] - frag - pad + 1)).astype(np.int64)
        base = gen_offs[g].astype(np.int64) + pos
        mates = []
        for mate in range(2):
            # forward-strand template start: left mate at pos, right mate ends at pos+frag
            start = base if mate == 0 else base + frag - L
            idx = start[:, None] + np.arange(L, dtype=np.int64)[None, :]
            has_indel = rng.random(n) < indel_frac
            ipos = rng.integers(40, 111, size=n)
            ilen = rng.integers(1, 4, size=n)
            is_del = rng.random(n) < 0.5
            col = np.arange(L, dtype=np.int64)[None, :]
            # deletion: skip ilen template bases after ipos
            dsel = (has_indel & is_del)[:, None] & (col >= ipos[:, None])
            idx = idx + dsel * ilen[:, None]
            # insertion: template stalls for ilen columns starting at ipos
            isel = (has_indel & ~is_del)[:, None]
            stall = np.clip(col - ipos[:, None], 0, ilen[:, None])
            idx = idx - isel * stall
            idx = np.minimum(idx, (gen_offs[g + 1].astype(np.int64) - 1)[:, None])
            seq = gen_bases[idx]
            ins_cols = isel & (col >= ipos[:, None]) & (col < (ipos + ilen)[:, None])
            seq[ins_cols] = ACGT[rng.integers(0, 4, size=int(ins_cols.sum()), dtype=np.uint8)]
            sub = rng.random((n, L)) < sub_rate
            shift = rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)
            cur = seq[sub]
            code = np.searchsorted(ACGT_SORTED, cur)
            seq[sub] = ACGT_SORTED[(code + shift) % 4]
            mates.append(seq)
        left, right = mates[0], revcomp_rows(mates[1])
        swap = rng.random(n) < 0.5
        r1[c0:c0 + n] = np.where(swap[:, None], right, left)
        r2[c0:c0 + n] = np.where(swap[:, None], left, right)
        truth["genome"][c0:c0 + n] = g
        truth["pos"][c0:c0 + n] = pos
        truth["frag"][c0:c0 + n] = frag
        truth["swapped"][c0:c0 + n] = swap
    bases = np.concatenate([r1.reshape(-1), r2.reshape(-1)])
    offs = np.arange(2 * n_pairs + 1, dtype=np.uint64) * np.uint64(L)
    return bases, offs, truth


ACGT_SORTED = np.frombuffer(b"ACGT", dtype=np.uint8)  # already ascending in ASCII


def sw_pairs(n: int, read_len: int = 150, window_len: int = 150, seed: int = 3):
    """Config 3 (SW microbench): n (read, window) pairs. Mix: 70 % ~1 % substitutions, 20 % one
    1-5 bp indel, 5 % unrelated random, 5 % containing N runs. window_len >= read_len; the read is
    planted at a uniform offset in the window."""
    rng = np.random.default_rng(seed)
    W = ACGT[rng.integers(0, 4, size=(n, window_len), dtype=np.uint8)]
    off = rng.integers(0, window_len - read_len + 1, size=n)
    kind = rng.random(n)
    col = np.arange(read_len, dtype=np.int64)[None, :]
    idx = off[:, None] + col
    indel = (kind >= 0.70) & (kind < 0.90)
    lo = min(30, read_len // 3)
    ipos = rng.integers(lo, read_len - lo, size=n)
    ilen = rng.integers(1, 6, size=n)
    is_del = rng.random(n) < 0.5
    idx = idx + ((indel & is_del)[:, None] & (col >= ipos[:, None])) * ilen[:, None]
    idx = idx - (indel & ~is_del)[:, None] * np.clip(col - ipos[:, None], 0, ilen[:, None])
    idx = np.clip(idx, 0, window_len - 1)
    Q = np.take_along_axis(W, idx, axis=1)
    ins_cols = (indel & ~is_del)[:, None] & (col >= ipos[:, None]) & (col < (ipos + ilen)[:, None])
    Q[ins_cols] = ACGT[rng.integers(0, 4, size=int(ins_cols.sum()), dtype=np.uint8)]
    sub = rng.random((n, read_len)) < 0.01
    Q[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
    unrelated = (kind >= 0.90) & (kind < 0.95)
    Q[unrelated] = ACGT[rng.integers(0, 4, size=(int(unrelated.sum()), read_len), dtype=np.uint8)]
    withn = kind >= 0.95
    nidx = np.flatnonzero(withn)
    for t, (arr, ln) in enumerate(((Q, read_len), (W, window_len))):
        sel = nidx[t::2]
        st = rng.integers(0, ln - 10, size=len(sel))
        run = rng.integers(1, 10, size=len(sel))
        for r, s, k in zip(sel, st, run):
            arr[r, s:s + k] = ord("N")
    qoffs = np.arange(n + 1, dtype=np.uint64) * np.uint64(read_len)
    roffs = np.arange(n + 1, dtype=np.uint64) * np.uint64(window_len)
    return Q.reshape(-1), qoffs, W.reshape(-1), roffs


def adversarial_set(seed: int = 7, n_genomes: int = 12, glen: int = 20_000, n_pairs: int = 3000):
    """Small hostile data set in the spirit of SURVEY.md App. A.4's validation: genomes sharing a
    diverged block, poly-A and N runs, a reverse-palindrome, ragged read lengths (some < 32), lower-case
    and random-byte reads, indels, both strands. Reads come back as R1 block then R2 block."""
    rng = np.random.default_rng(seed)
    gens = [ACGT[rng.integers(0, 4, size=glen + int(rng.integers(0, 500)), dtype=np.uint8)] for _ in range(n_genomes)]
    block = ACGT[rng.integers(0, 4, size=3000, dtype=np.uint8)]
    for g in range(0, n_genomes, 2):  # shared block, 2.5 % diverged
        b = block.copy()
        m = rng.random(len(b)) < 0.025
        b[m] = ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        p = int(rng.integers(0, glen - 3000))
        gens[g][p:p + 3000] = b
    gens[1][500:800] = ord("A")          # poly-A: zero k-mers never seed (Overlap.h:236-239)
    gens[2][900:1200] = ord("T")
    gens[3][1500:1550] = ord("N")
    half = ACGT[rng.integers(0, 4, size=20, dtype=np.uint8)]
    gens[4][2000:2040] = np.concatenate([half, _COMP[half[::-1]]])  # reverse palindrome
    gens[5][-10:] = ord("N")
    gens.append(ACGT[rng.integers(0, 4, size=40, dtype=np.uint8)])   # tiny genome (3 k-mers at gap 16)
    gens.append(ACGT[rng.integers(0, 4, size=20, dtype=np.uint8)])   # shorter than k: no k-mers
    gen_offs = np.zeros(len(gens) + 1, dtype=np.uint64)
    gen_offs[1:] = np.cumsum([len(g) for g in gens])
    gen_bases = np.concatenate(gens)

    def one_read(glist):
        g = int(rng.integers(0, len(glist)))
        L = int(rng.integers(100, 161)) if rng.random() > 0.05 else int(rng.integers(10, 60))
        src = glist[g]
        p = int(rng.integers(-30, len(src) - L + 30))  # may hang off either end
        lo, hi = max(p, 0), min(p + L, len(src))
        s = src[lo:hi].copy()
        if len(s) == 0:
            s = ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        m = rng.random(len(s)) < 0.02
        s[m] = ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        if rng.random() < 0.15 and len(s) > 80:  # indel
            q = int(rng.integers(30, len(s) - 30)); k = int(rng.integers(1, 4))
            if rng.random() < 0.5:
                s = np.concatenate([s[:q], s[q + k:]])
            else:
                s = np.concatenate([s[:q], ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)], s[q:]])
        if rng.random() < 0.5:
            s = _COMP[s[::-1]]
        u = rng.random()
        if u < 0.01:
            s = np.frombuffer(bytes(s).lower(), dtype=np.uint8).copy()
        elif u < 0.03:
            k = int(rng.integers(1, 6)); q = int(rng.integers(0, max(1, len(s) - k)))
            s[q:q + k] = ord("N")
        elif u < 0.04:
            s = rng.integers(33, 127, size=len(s), dtype=np.uint8)
        return s

    r1 = [one_read(gens[:n_genomes]) for _ in range(n_pairs)]
    r2 = [one_read(gens[:n_genomes]) for _ in range(n_pairs)]
    # make half the pairs proper FR pairs from one fragment so pairing has work to do
    for i in range(0, n_pairs, 2):
        g = int(rng.integers(0, n_genomes)); src = gens[g]
        frag = int(rng.integers(200, 500)); p = int(rng.integers(0, len(src) - frag))
        a = src[p:p + 150].copy(); b = _COMP[src[p + frag - 150:p + frag][::-1]]
        m = rng.random(150) < 0.01
        a[m] = ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        if rng.random() < 0.5:
            a, b = b, a
        r1[i], r2[i] = a, b
    reads = r1 + r2
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    return gen_bases, gen_offs, np.concatenate(reads), offs


# ---------------------------------------------------------------- config 2: taxonomy + GenBank flat files

def tree_taxonomy(n_strains: int):
    """The NCBI-style taxonomy of tree_genomes(n_strains): root 1 -> superkingdom 2 -> phyla -> genera -> species -> one
    strain node per genome (same index arithmetic as tree_genomes). -> (nodes [(id, parent, rank, name)], strain tax ids)."""
    n_species = max(1, n_strains // 5); n_genera = max(1, n_species // 4); n_phyla = max(1, n_genera // 5)
    nodes = [(1, 1, "no rank", "root"), (131567, 1, "no rank", "cellular organisms"), (2, 131567, "superkingdom", "Bacteria")]
    ph = [1000 + i for i in range(n_phyla)]; ge = [2000 + i for i in range(n_genera)]; sp = [10000 + i for i in range(n_species)]
    nodes += [(ph[i], 2, "phylum", f"Phylum{i} <synthetic>") for i in range(n_phyla)]
    nodes += [(ge[i], ph[i % n_phyla], "genus", f"Genus{i}") for i in range(n_genera)]
    nodes += [(sp[i], ge[i % n_genera], "species", f"Genus{i % n_genera} species{i}") for i in range(n_species)]
    strains = [100000 + i for i in range(n_strains)]
    nodes += [(strains[i], sp[i % n_species], "no rank", f"Genus{(i % n_species) % n_genera} species{i % n_species} str. S{i} & co") for i in range(n_strains)]
    return nodes, np.array(strains, dtype=np.uint32)


def write_taxonomy_dumps(nodes, names_path, nodes_path):
    """names.dmp / nodes.dmp in NCBI's `\\t|\\t` layout (what --parse-taxonomy reads, TaxonomyDatabase.h:95-151); every node
    also gets a synonym line, which the parser must skip."""
    with open(nodes_path, "w") as f:
        for tid, parent, rank, _ in nodes:
            f.write(f"{tid}\t|\t{parent}\t|\t{rank}\t|\t\t|\t0\t|\t1\t|\t11\t|\t1\t|\t0\t|\t1\t|\t1\t|\t0\t|\t\t|\n")
    with open(names_path, "w") as f:
        for tid, _, _, name in nodes:
            f.write(f"{tid}\t|\told name of {tid}\t|\t\t|\tsynonym\t|\n")
            f.write(f"{tid}\t|\t{name}\t|\t\t|\tscientific name\t|\n")


def genbank_text(bases: bytes, accession: str, gi: int, taxon: int, organism: str, strain_idx: int, species_idx: int, seed: int = 0,
                 cds_every: int = 1000) -> bytes:
    """One GenBank flat-file record: LOCUS / DEFINITION / VERSION ACC.1 GI:n / source with
    /db_xref="taxon:T" / a gene + CDS feature about every `cds_every` bases (qualifiers /gene /locus_tag /product /protein_id
    /db_xref="GeneID:n", some products wrapped over two lines, some CDS on the complement strand, a tRNA now and then) /
    ORIGIN in 60-base lines / `//`. Protein ids are per SPECIES, so strains of one species share proteins."""
    rng = np.random.default_rng(seed)
    n = len(bases)
    out = [f"LOCUS       {accession:<16} {n:>10} bp    DNA     circular BCT 01-JAN-2020",
           f"DEFINITION  {organism}, complete genome.",
           f"ACCESSION   {accession}",
           f"VERSION     {accession}.1  GI:{gi}",
           "KEYWORDS    .",
           f"SOURCE      {organism}",
           f"  ORGANISM  {organism}",
           "            Bacteria; Synthetic.",
           "FEATURES             Location/Qualifiers",
           f"     source          1..{n}",
           f"                     /organism=\"{organism}\"",
           "                     /mol_type=\"genomic DNA\"",
           f"                     /db_xref=\"taxon:{taxon}\""]
    pos, k = 50, 0
    while pos + 400 < n:
        length = int(rng.integers(300, min(900, n - pos - 10)))
        a, b = pos + 1, pos + length
        loc = f"complement({a}..{b})" if k % 3 == 2 else f"{a}..{b}"
        tag = f"S{strain_idx}_{k:04d}"
        if k % 7 == 6:
            out += [f"     tRNA            {loc}", f"                     /locus_tag=\"{tag}\"", f"                     /product=\"tRNA-Ala\""]
        else:
            gene = f"gen{k % 50}{'AB'[k % 2]}"
            product = f"protein <{k}> of species {species_idx}"
            out += [f"     gene            {loc}", f"                     /gene=\"{gene}\"", f"                     /locus_tag=\"{tag}\""]
            out += [f"     CDS             {loc}", f"                     /gene=\"{gene}\"", f"                     /locus_tag=\"{tag}\""]
            if k % 4:
                out += [f"                     /product=\"{product}\""]
            else:
                out += [f"                     /product=\"very long hypothetical membrane transporter", f"                     component number {k} of species {species_idx}\""]
            if k % 5:
                out += [f"                     /protein_id=\"WP_{species_idx:03d}{k:05d}.1\""]
            out += [f"                     /db_xref=\"GeneID:{5_000_000 + species_idx * 10_000 + k}\"",
                    "                     /translation=\"MKVLAAGIVGLCAQEPTW\""]
        pos += int(rng.integers(cds_every // 2, cds_every * 3 // 2)); k += 1
    out.append("ORIGIN      ")
    low = bases.lower().decode()
    for i in range(0, n, 60):
        chunk = low[i:i + 60]
        out.append(f"{i + 1:>9} " + " ".join(chunk[j:j + 10] for j in range(0, len(chunk), 10)))
    out.append("//")
    return ("\n".join(out) + "\n").encode()
