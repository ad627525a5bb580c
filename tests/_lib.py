"""ctypes bindings for the CHECKERS (oracle/ and oracle/_ref) and shared test helpers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module: the product path (k-slam_b200/) never touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

KMER_DT = np.dtype([("kmer", "<u8"), ("id_flags", "<u4"), ("offset", "<u4")])
SEED_DT = np.dtype([("read", "<u4"), ("entry", "<u4"), ("rel", "<i4"), ("rev_comp", "<u4")])
OVERLAP_DT = np.dtype([("read", "<u4"), ("entry", "<u4"), ("rel", "<i4"), ("rev_comp", "<u4"),
                       ("ref_begin", "<i4"), ("ref_end", "<i4"), ("query_begin", "<i4"), ("query_end", "<i4"),
                       ("sw_score", "<u4"), ("cigar_off", "<u4"), ("cigar_len", "<u4"), ("flags", "<u4")])
PAIR_DT = np.dtype([("combined_score", "<u4"), ("entry", "<u4"), ("ref_start", "<i4"), ("ref_end", "<i4"),
                    ("insert_size", "<u4"), ("r1_idx", "<i4"), ("r2_idx", "<i4"), ("pad", "<u4")])
assert KMER_DT.itemsize == 16 and SEED_DT.itemsize == 16 and OVERLAP_DT.itemsize == 48 and PAIR_DT.itemsize == 32


class KoParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32),
                ("gap_extend", C.c_int32), ("score_threshold", C.c_uint32), ("report_cigar", C.c_int32)]


def default_params(report_cigar=1, score_threshold=0, match=2, mismatch=3, gap_open=5, gap_extend=2):
    return KoParams(match, mismatch, gap_open, gap_extend, score_threshold, report_cigar)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def u8(a):
    return np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a, dtype=np.uint8)


def offsets_of(seqs):
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    return offs


def concat(seqs):
    if not seqs:
        return np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64)
    return np.concatenate([u8(s) for s in seqs]) if sum(len(s) for s in seqs) else np.zeros(0, np.uint8), offsets_of(seqs)


def build_oracle():
    """Compile oracle/libkslam_oracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libkslam_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.ko_extract_kmers.restype = C.c_uint64
        L.ko_extract_kmers.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
        L.ko_sort_kmers.argtypes = [C.c_void_p, C.c_uint64]
        L.ko_find_seeds_raw.restype = C.c_uint64
        L.ko_find_seeds_raw.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ko_sort_unique_seeds.restype = C.c_uint64
        L.ko_sort_unique_seeds.argtypes = [C.c_void_p, C.c_uint64]
        L.ko_ssw_batch.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(KoParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        L.ko_ssw_align_gotoh.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(KoParams), C.c_void_p]
        L.ko_align_seeds.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(KoParams), C.c_void_p, C.c_uint32, C.c_int]
        L.ko_sort_for_pairing.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.ko_pair_overlaps.restype = C.c_uint64
        L.ko_pair_overlaps.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libkslam_ref.so"))


def ref():
    """The reference's own code (oracle/_ref/libkslam_ref.so). It writes log.txt into the CWD
    (sequenceTools.h:176), so the first call happens inside a scratch directory."""
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libkslam_ref.so"))
        L.kref_create.restype = C.c_void_p
        for name in ("kref_extract_kmers", "kref_find_seeds_raw", "kref_find_seeds", "kref_align_to_database",
                     "kref_screen", "kref_num_overlaps", "kref_cigar_total", "kref_pair", "kref_ssw_batch"):
            getattr(L, name).restype = C.c_uint64
        L.kref_destroy.argtypes = [C.c_void_p]
        L.kref_set_threads.argtypes = [C.c_int]
        L.kref_set_params.argtypes = [C.c_uint32] * 5 + [C.c_int]
        L.kref_set_genomes.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.kref_set_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.kref_extract_kmers.argtypes = [C.c_void_p, C.c_int]
        for name in ("kref_sort_kmers", "kref_find_seeds_raw", "kref_find_seeds", "kref_align_to_database",
                     "kref_screen", "kref_num_overlaps", "kref_cigar_total", "kref_pair"):
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("kref_get_kmers", "kref_get_seeds", "kref_get_pairs"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        L.kref_get_overlaps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.kref_ssw_batch.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        L.kref_set_read_quals.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.kref_sam.restype = C.c_uint64
        L.kref_sam.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
        L.kref_sam_single.restype = C.c_uint64
        L.kref_sam_single.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_uint64]
        L.kref_sam_header.restype = C.c_uint64
        L.kref_sam_header.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64]
        L.kref_fastq_open.restype = C.c_void_p
        L.kref_fastq_open.argtypes = [C.c_char_p, C.c_char_p]
        L.kref_fastq_next.restype = C.c_uint64
        L.kref_fastq_next.argtypes = [C.c_void_p, C.c_uint]
        L.kref_fastq_get.restype = C.c_uint64
        L.kref_fastq_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.kref_fastq_close.argtypes = [C.c_void_p]
        L.kref_set_read_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.kref_set_entry_meta.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_char_p]
        L.kref_add_gene.argtypes = [C.c_void_p, C.c_uint64] + [C.c_char_p] * 5 + [C.c_uint32] * 3
        L.kref_taxdb_open.restype = C.c_void_p
        L.kref_taxdb_open.argtypes = [C.c_char_p]
        L.kref_taxdb_close.argtypes = [C.c_void_p]
        L.kref_taxdb_build.argtypes = [C.c_char_p] * 3
        L.kref_taxdb_size.restype = C.c_uint64
        L.kref_taxdb_size.argtypes = [C.c_void_p]
        L.kref_lca.restype = C.c_uint32
        L.kref_lca.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.kref_lineage.restype = C.c_uint64
        L.kref_lineage.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64]
        L.kref_meta_batch.restype = C.c_uint64
        L.kref_meta_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                      C.c_void_p, C.c_uint64]
        L.kref_meta_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p]
        L.kref_parse_index.restype = C.c_int64
        L.kref_parse_index.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_uint64, C.c_char_p]
        L.kref_parsed_index_dump.restype = C.c_uint64
        L.kref_parsed_index_dump.argtypes = [C.c_void_p, C.c_uint64]
        L.kref_use_parsed_index.argtypes = [C.c_void_p]
        cwd = os.getcwd()
        scratch = tempfile.mkdtemp(prefix="kref_")
        os.chdir(scratch)
        try:
            h = L.kref_create()
            L.kref_set_params(2, 3, 5, 2, 0, 0)
            z = np.zeros(1, dtype=np.uint64)
            L.kref_set_reads(h, 0, None, _p(z))
            L.kref_extract_kmers(h, 1)  # first log() call opens log.txt here
            L.kref_destroy(h)
        finally:
            os.chdir(cwd)
        _ref = L
    return _ref


# ---------------------------------------------------------------- oracle wrappers

def ko_extract(bases, offs, is_gb, gap):
    L = oracle()
    bases = u8(bases); offs = np.ascontiguousarray(offs, dtype=np.uint64)
    n = len(offs) - 1
    cnt = L.ko_extract_kmers(n, _p(bases), _p(offs), int(is_gb), gap, None)
    out = np.zeros(cnt, dtype=KMER_DT)
    L.ko_extract_kmers(n, _p(bases), _p(offs), int(is_gb), gap, _p(out))
    return out


def ko_sort_kmers(recs):
    recs = recs.copy()
    oracle().ko_sort_kmers(_p(recs), len(recs))
    return recs


def ko_seeds_raw(sorted_recs, read_lens):
    L = oracle()
    read_lens = np.ascontiguousarray(read_lens, dtype=np.uint32)
    cnt = L.ko_find_seeds_raw(_p(sorted_recs), len(sorted_recs), _p(read_lens), None)
    out = np.zeros(cnt, dtype=SEED_DT)
    L.ko_find_seeds_raw(_p(sorted_recs), len(sorted_recs), _p(read_lens), _p(out))
    return out


def ko_sort_unique(seeds):
    seeds = seeds.copy()
    n = oracle().ko_sort_unique_seeds(_p(seeds), len(seeds))
    return seeds[:n].copy()


def ko_ssw_batch(q, qoffs, r, roffs, params, cigar_cap=64, threads=8):
    L = oracle()
    q = u8(q); r = u8(r)
    qoffs = np.ascontiguousarray(qoffs, dtype=np.uint64); roffs = np.ascontiguousarray(roffs, dtype=np.uint64)
    n = len(qoffs) - 1
    out = np.zeros(n, dtype=OVERLAP_DT)
    pool = np.zeros(n * cigar_cap, dtype=np.uint32)
    L.ko_ssw_batch(n, _p(q), _p(qoffs), _p(r), _p(roffs), C.byref(params), _p(out), _p(pool), cigar_cap, threads)
    return out, pool


def ko_ssw_gotoh(q, qoffs, r, roffs, params):
    """Score and coordinates through the un-striped (plain Gotoh) scan: the cross-check of the striped restatement."""
    L = oracle()
    q = u8(q); r = u8(r)
    n = len(qoffs) - 1
    out = np.zeros(n, dtype=OVERLAP_DT)
    for i in range(n):
        L.ko_ssw_align_gotoh(q.ctypes.data + int(qoffs[i]), int(qoffs[i + 1] - qoffs[i]), r.ctypes.data + int(roffs[i]),
                             int(roffs[i + 1] - roffs[i]), C.byref(params), out.ctypes.data + i * OVERLAP_DT.itemsize)
    return out


def ko_pipeline(gen_bases, gen_offs, read_bases, read_offs, params, cigar_cap=64, threads=8, paired=True):
    """Oracle restatement of alignToDatabase (+ screen + getPairedOverlaps). Returns a dict of every stage."""
    L = oracle()
    gen_bases = u8(gen_bases); read_bases = u8(read_bases)
    gen_offs = np.ascontiguousarray(gen_offs, dtype=np.uint64); read_offs = np.ascontiguousarray(read_offs, dtype=np.uint64)
    read_lens = (read_offs[1:] - read_offs[:-1]).astype(np.uint32)
    rk = ko_extract(read_bases, read_offs, False, 1)
    gk = ko_extract(gen_bases, gen_offs, True, 16)
    allk = ko_sort_kmers(np.concatenate([rk, gk]))
    raw = ko_seeds_raw(allk, read_lens)
    seeds = ko_sort_unique(raw)
    ov = np.zeros(len(seeds), dtype=OVERLAP_DT)
    for f in ("read", "entry", "rel", "rev_comp"):
        ov[f] = seeds[f]
    pool = np.zeros(max(1, len(ov) * cigar_cap), dtype=np.uint32)
    L.ko_align_seeds(len(ov), _p(ov), _p(read_bases), _p(read_offs), _p(gen_bases), _p(gen_offs),
                     C.byref(params), _p(pool), cigar_cap, threads)
    res = dict(read_kmers=rk, genome_kmers=gk, sorted_kmers=allk, raw_seeds=raw, seeds=seeds,
               overlaps=ov, cigar_pool=pool, cigar_cap=cigar_cap)
    if paired:
        mid = (len(read_offs) - 1) // 2
        kept = ov[ov["sw_score"] >= params.score_threshold].copy()  # Overlap.h:329-341
        L.ko_sort_for_pairing(_p(kept), len(kept), mid)
        cnt = L.ko_pair_overlaps(_p(kept), len(kept), mid, _p(read_lens), None)
        pairs = np.zeros(cnt, dtype=PAIR_DT)
        L.ko_pair_overlaps(_p(kept), len(kept), mid, _p(read_lens), _p(pairs))
        res.update(pair_sorted_overlaps=kept, pairs=pairs)
    return res


def cigars_of(ov, pool):
    """List of tuples of cigar words per overlap (pool indexed by cigar_off)."""
    return [tuple(pool[o["cigar_off"]:o["cigar_off"] + o["cigar_len"]].tolist()) for o in ov]


# ---------------------------------------------------------------- reference (_ref) wrappers

class Ref:
    """One reference context (GenbankIndex + reads) driving the reference's own functions."""

    def __init__(self, gen_bases, gen_offs, read_bases, read_offs, params):
        self.L = ref()
        self.h = self.L.kref_create()
        self.L.kref_set_params(params.match, params.mismatch, params.gap_open, params.gap_extend,
                               params.score_threshold, params.report_cigar)
        gb = u8(gen_bases); go = np.ascontiguousarray(gen_offs, dtype=np.uint64)
        rb = u8(read_bases); ro = np.ascontiguousarray(read_offs, dtype=np.uint64)
        self.L.kref_set_genomes(self.h, len(go) - 1, _p(gb), _p(go))
        self.L.kref_set_reads(self.h, len(ro) - 1, _p(rb), _p(ro))

    def close(self):
        if self.h:
            self.L.kref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def kmers(self, which, sort=False):
        n = self.L.kref_extract_kmers(self.h, which)
        if sort:
            self.L.kref_sort_kmers(self.h)
        out = np.zeros(n, dtype=KMER_DT)
        self.L.kref_get_kmers(self.h, _p(out))
        return out

    def seeds(self, raw):
        """Needs kmers(3, sort=True) first."""
        n = self.L.kref_find_seeds_raw(self.h) if raw else self.L.kref_find_seeds(self.h)
        out = np.zeros(n, dtype=SEED_DT)
        self.L.kref_get_seeds(self.h, _p(out))
        return out

    def _overlaps(self):
        n = self.L.kref_num_overlaps(self.h)
        out = np.zeros(n, dtype=OVERLAP_DT)
        pool = np.zeros(max(1, self.L.kref_cigar_total(self.h)), dtype=np.uint32)
        self.L.kref_get_overlaps(self.h, _p(out), _p(pool))
        return out, pool

    def align_to_database(self):
        self.L.kref_align_to_database(self.h)
        return self._overlaps()

    def screen_and_pair(self):
        self.L.kref_screen(self.h)
        n = self.L.kref_pair(self.h)
        pairs = np.zeros(n, dtype=PAIR_DT)
        self.L.kref_get_pairs(self.h, _p(pairs))
        ov, pool = self._overlaps()
        return ov, pool, pairs


def ref_sam(R, quals=None, qual_offs=None, num_alignments=10, fraction=0.95, pseudo=True, sam_xa=False):
    """The reference's host stages after kref_pair up to the SAM text (SLAM.h:215-239) on a Ref context that has run
    align_to_database() and screen_and_pair(). Returns (text, max insert size). Consumes the context's pairs."""
    L = R.L
    if quals is not None:
        q = u8(quals); qo = np.ascontiguousarray(qual_offs, dtype=np.uint64)
        L.kref_set_read_quals(R.h, _p(q), _p(qo))
    tmp = tempfile.NamedTemporaryFile(suffix=".sam", delete=False); tmp.close()
    mi = C.c_uint32()
    buf = np.zeros(1 << 26, dtype=np.uint8)
    n = L.kref_sam(R.h, num_alignments, fraction, int(pseudo), int(sam_xa), tmp.name.encode(), _p(buf), len(buf), C.byref(mi))
    os.unlink(tmp.name)
    assert n <= len(buf)
    return bytes(buf[:n]), mi.value


def ref_sam_single(R, quals=None, qual_offs=None, num_alignments=10, fraction=0.95, pseudo=True, sam_xa=False):
    """Single-end flavour (SLAM.h:223-228) on a Ref context that has run align_to_database() and kref_screen."""
    L = R.L
    if quals is not None:
        q = u8(quals); qo = np.ascontiguousarray(qual_offs, dtype=np.uint64)
        L.kref_set_read_quals(R.h, _p(q), _p(qo))
    tmp = tempfile.NamedTemporaryFile(suffix=".sam", delete=False); tmp.close()
    buf = np.zeros(1 << 26, dtype=np.uint8)
    n = L.kref_sam_single(R.h, num_alignments, fraction, int(pseudo), int(sam_xa), tmp.name.encode(), _p(buf), len(buf))
    os.unlink(tmp.name)
    return bytes(buf[:n])


def ref_sam_header(R, cmd=""):
    buf = np.zeros(1 << 20, dtype=np.uint8)
    n = R.L.kref_sam_header(R.h, cmd.encode(), _p(buf), len(buf))
    return bytes(buf[:n])


def ref_ssw_batch(q, qoffs, r, roffs, params, cigar_cap=64, threads=0):
    L = ref()
    L.kref_set_params(params.match, params.mismatch, params.gap_open, params.gap_extend,
                      params.score_threshold, params.report_cigar)
    q = u8(q); r = u8(r)
    qoffs = np.ascontiguousarray(qoffs, dtype=np.uint64); roffs = np.ascontiguousarray(roffs, dtype=np.uint64)
    n = len(qoffs) - 1
    out = np.zeros(n, dtype=OVERLAP_DT)
    pool = np.zeros(n * cigar_cap, dtype=np.uint32)
    L.kref_ssw_batch(n, _p(q), _p(qoffs), _p(r), _p(roffs), params.report_cigar, params.score_threshold,
                     _p(out), _p(pool), cigar_cap, threads)
    return out, pool


def ref_read_fastq(r1, r2, max_reads):
    """The reference's own reader (FASTQsequence.h:110-165) over files, batch by batch. Yields per batch a list of
    (id, bases, quality) byte strings, or the string "mismatch" when the reference throws."""
    L = ref()
    h = L.kref_fastq_open(r1.encode(), r2.encode() if r2 else None)
    out = []
    try:
        while True:
            n = L.kref_fastq_next(h, max_reads)
            if n == 2**64 - 1:
                out.append("mismatch"); break
            if n == 0:
                break
            fields = []
            for which in (2, 0, 1):
                offs = np.zeros(n + 1, dtype=np.uint64)
                size = L.kref_fastq_get(h, which, None, _p(offs))
                buf = np.zeros(max(1, size), dtype=np.uint8)
                L.kref_fastq_get(h, which, _p(buf), _p(offs))
                fields.append([bytes(buf[int(offs[i]):int(offs[i + 1])]) for i in range(n)])
            out.append(list(zip(*fields)))
    finally:
        L.kref_fastq_close(h)
    return out


def load_pkg():
    """Import the product package (directory name has a hyphen, so go through __graft_entry__)."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    return ge.load_pkg()


# ---------------------------------------------------------------- reference: taxonomy, metagenomic outputs, database builders

def ref_meta_batch(R, taxdb, quals=None, qual_offs=None, ids=None, id_offs=None, num_alignments=10, fraction=0.95, pseudo=True,
                   sam_xa=False, paired=True, want_sam=True):
    """SLAM.h:215-249 on a Ref context after screen_and_pair() (paired) or align_to_database() + kref_screen (single-end).
    taxdb: handle from ref().kref_taxdb_open or None. Returns the SAM text (b"" when not wanted)."""
    L = R.L
    if quals is not None:
        q = u8(quals); qo = np.ascontiguousarray(qual_offs, dtype=np.uint64)
        L.kref_set_read_quals(R.h, _p(q), _p(qo))
    if ids is not None:
        i = u8(ids); io = np.ascontiguousarray(id_offs, dtype=np.uint64)
        L.kref_set_read_ids(R.h, _p(i), _p(io))
    tmp = tempfile.NamedTemporaryFile(suffix=".sam", delete=False); tmp.close()
    buf = np.zeros(1 << 26, dtype=np.uint8)
    n = L.kref_meta_batch(R.h, taxdb, num_alignments, fraction, int(pseudo), int(sam_xa), int(paired), int(want_sam), tmp.name.encode(),
                          _p(buf), len(buf))
    os.unlink(tmp.name)
    assert n <= len(buf)
    return bytes(buf[:n])


def ref_meta_finish(R, taxdb, num_reads):
    """SLAM.h:256-265 -> (text of _PerRead, XML, _abbreviated)."""
    d = tempfile.mkdtemp(prefix="kref_meta_")
    prefix = os.path.join(d, "out")
    R.L.kref_meta_finish(R.h, taxdb, num_reads, prefix.encode())
    res = []
    for suffix in ("_PerRead", "", "_abbreviated"):
        with open(prefix + suffix, "rb") as f:
            res.append(f.read())
        os.unlink(prefix + suffix)
    os.rmdir(d)
    return tuple(res)


def ref_parse_index(kind, paths, taxdb_path=None):
    """The reference's createIndexFromGBFF (kind 0; needs ./taxDB) or createIndexFromFASTA (kind 1) -> list of entries
    (dicts like database.read_database's), decoded from the harness's dump of the GenbankIndex it built."""
    L = ref()
    cwd = os.getcwd()
    d = tempfile.mkdtemp(prefix="kref_idx_")
    try:
        os.chdir(d)
        if taxdb_path:
            with open(taxdb_path, "rb") as f, open("taxDB", "wb") as g:
                g.write(f.read())
        arr = (C.c_char_p * len(paths))(*[os.fsencode(os.path.join(cwd, p) if not os.path.isabs(p) else p) for p in paths])
        n = L.kref_parse_index(kind, arr, len(paths), b"database")
    finally:
        os.chdir(cwd)
        for f in os.listdir(d):
            os.unlink(os.path.join(d, f))
        os.rmdir(d)
    if n < 0:
        return None
    size = L.kref_parsed_index_dump(None, 0)
    buf = np.zeros(max(1, size), dtype=np.uint8)
    L.kref_parsed_index_dump(_p(buf), len(buf))
    return decode_index_dump(bytes(buf[:size]))


def decode_index_dump(text):
    entries = []
    for rec in text.split(b"\x1e"):
        if not rec:
            continue
        f = rec.split(b"\x1f")
        if f[0] == b"E":
            entries.append(dict(locus_tag=f[1], taxonomy_id=int(f[2]), genbank_id=int(f[3]), is_plasmid=int(f[4]), is_16s=int(f[5]), bases=f[6], genes=[]))
        else:
            entries[-1]["genes"].append(dict(gene_name=f[1], locus_tag=f[2], protein_id=f[3], product=f[4], reference_sequence=f[5],
                                             gene_id=int(f[6]), start=int(f[7]), stop=int(f[8]), complement=int(f[9])))
    return entries


def index_entries(ix):
    """The same list-of-dicts view of a kslam_b200.Index."""
    out = []
    b = ix.bases.tobytes()
    for e in range(ix.n_entries):
        out.append(dict(locus_tag=ix.locus_tags[e], taxonomy_id=int(ix.taxonomy_ids[e]), bases=b[int(ix.offs[e]):int(ix.offs[e + 1])],
                        genes=ix.gene_records(e)))
    return out
