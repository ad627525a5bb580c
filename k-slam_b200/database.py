"""`DIR/database` — the reference's index file.

The reference stores `GenbankIndex` with `boost::archive::text_oarchive` (GenbankTools.h:197-205) and reads it back with
`text_iarchive` (:336-344). Serialised members, in order (the `serialize` methods at GenbankTools.h:57-62,100-109,154-163,
197-200):
    GenbankIndex { vector<GenbankEntry> entries }
    GenbankEntry { string bases; u32 taxonomyID; u32 genbankID; bool isPlasmid; bool is16S; string locusTag; vector<Gene> genes }
    Gene         { string geneName, locusTag, proteinID, product, referenceSequence; u32 geneID; CDS codingSequence }
    CDS          { u32 start; u32 stop; bool complement }
Text-archive grammar: space-separated tokens after the header `22 serialization::archive <libver>`; the FIRST object of
every class type is preceded by its class information `<tracking> <version>` = `0 0`; `std::vector<T>` is
`<count> <item_version>` then the items; `std::string` is `<len>`, one space, then exactly `len` raw bytes; bool is 0/1.

Parity: this image has no Boost headers, so the reference's own writer cannot be compiled here; the format is pinned
against the REAL Boost.Serialization library instead — oracle/boost_archive_probe.cpp links the header-less
libboost_serialization.so (1.78) that ships inside Nsight Compute and lets its save_object / text_oarchive machinery write
the same index, and oracle/ref_shim_boost lets the reference's OWN writeIndexToBoostSerial / getIndexFromBoostSerial run on
that library: tests/test_database_format.py compares the bytes (only the library-version token differs) and has the
reference's reader load our files.
"""
from __future__ import annotations

import numpy as np

HEADER = b"22 serialization::archive"
CLASSES = ("index", "entries", "entry", "genes", "gene", "cds")


class ArchiveError(ValueError):
    pass


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.p, self.seen = data, 0, set()

    def token(self) -> bytes:
        d, p, n = self.d, self.p, len(self.d)
        while p < n and d[p] in b" \n\r\t":
            p += 1
        q = p
        while q < n and d[q] not in b" \n\r\t":
            q += 1
        if q == p:
            raise ArchiveError("unexpected end of archive")
        self.p = q
        return d[p:q]

    def uint(self) -> int:
        try:
            return int(self.token())
        except ValueError as e:
            raise ArchiveError(f"expected a number at byte {self.p}") from e

    def string(self) -> bytes:
        n = self.uint()
        self.p += 1                      # exactly one separator, then n raw bytes (they may contain spaces)
        if self.p + n > len(self.d):
            raise ArchiveError("string runs past the end of the archive")
        s = self.d[self.p:self.p + n]
        self.p += n
        return s

    def class_info(self, name):
        if name not in self.seen:        # first object of this class: <tracking> <version>
            self.seen.add(name)
            self.uint(); self.uint()

    def vector(self, name):
        self.class_info(name)
        count = self.uint()
        self.uint()                      # item_version
        return count


def read_database(path):
    """-> list of entries: dict(bases, taxonomy_id, genbank_id, is_plasmid, is_16s, locus_tag, genes=[dict(...)])."""
    with open(path, "rb") as f:
        data = f.read()
    if not data.startswith(HEADER):
        raise ArchiveError("not a Boost text archive (header missing)")
    r = _Reader(data)
    r.p = len(HEADER)
    if r.uint() < 4:
        raise ArchiveError("archive library version < 4 is not supported")
    r.class_info("index")
    entries = []
    for _ in range(r.vector("entries")):
        r.class_info("entry")
        e = dict(bases=r.string(), taxonomy_id=r.uint(), genbank_id=r.uint(), is_plasmid=bool(r.uint()), is_16s=bool(r.uint()),
                 locus_tag=r.string(), genes=[])
        for _ in range(r.vector("genes")):
            r.class_info("gene")
            g = dict(gene_name=r.string(), locus_tag=r.string(), protein_id=r.string(), product=r.string(),
                     reference_sequence=r.string(), gene_id=r.uint())
            r.class_info("cds")
            g.update(start=r.uint(), stop=r.uint(), complement=bool(r.uint()))
            e["genes"].append(g)
        entries.append(e)
    return entries


def write_database(path, entries, libver=17):
    """Inverse of read_database (what `--parse-fasta` / `--parse-genbank` would write)."""
    out, seen = [HEADER, b" %d" % libver], set()

    def info(name):
        if name not in seen:
            seen.add(name); out.append(b" 0 0")

    def s(x):
        x = x if isinstance(x, bytes) else str(x).encode()
        out.append(b" %d " % len(x)); out.append(x)

    info("index"); info("entries")
    out.append(b" %d 0" % len(entries))
    for e in entries:
        info("entry")
        s(e["bases"])
        out.append(b" %d %d %d %d" % (e.get("taxonomy_id", 0), e.get("genbank_id", 0), int(e.get("is_plasmid", False)), int(e.get("is_16s", False))))
        s(e.get("locus_tag", b""))
        genes = e.get("genes", [])
        info("genes")
        out.append(b" %d 0" % len(genes))
        for g in genes:
            info("gene")
            for k in ("gene_name", "locus_tag", "protein_id", "product", "reference_sequence"):
                s(g.get(k, b""))
            out.append(b" %d" % g.get("gene_id", 0))
            info("cds")
            out.append(b" %d %d %d" % (g.get("start", 0), g.get("stop", 0), int(g.get("complement", False))))
    out.append(b"\n")
    with open(path, "wb") as f:
        f.write(b"".join(out))


def flatten(entries):
    """-> (bases u8, offs u64, locus tags, taxonomy ids u32): what Aligner.load_genomes and SamWriter take."""
    offs = np.zeros(len(entries) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(e["bases"]) for e in entries])
    bases = np.frombuffer(b"".join(e["bases"] for e in entries), dtype=np.uint8).copy() if entries else np.zeros(0, np.uint8)
    return bases, offs, [e["locus_tag"] for e in entries], np.array([e["taxonomy_id"] for e in entries], dtype=np.uint32)
