"""CPU tier: DIR/database (Boost text archive of GenbankIndex). PARITY UNPINNED — no Boost in this image;
the vectors below are SURVEY's worked example and this module's own round trip."""
import numpy as np
import pytest


def test_survey_worked_example(pkg, tmp_path):
    from kslam_b200 import database as db
    p = tmp_path / "database"
    p.write_bytes(b"22 serialization::archive 12 0 0 0 0 1 0 0 0 8 ACGT ACG 0 0 0 0 2 g0 0 0 0 0\n")   # bases may hold a space: parsed by length
    e = db.read_database(p)
    assert len(e) == 1 and e[0]["bases"] == b"ACGT ACG" and e[0]["locus_tag"] == b"g0" and e[0]["genes"] == [] and e[0]["taxonomy_id"] == 0


def test_round_trip_with_genes(pkg, tmp_path):
    from kslam_b200 import database as db
    entries = [dict(bases=b"ACGT" * 10, taxonomy_id=562, genbank_id=7, is_plasmid=False, is_16s=True, locus_tag=b"NC_1",
                    genes=[dict(gene_name=b"dnaA", locus_tag=b"b0001", protein_id=b"NP_1.1", product=b"chromosomal replication initiator",
                                reference_sequence=b"", gene_id=944, start=10, stop=30, complement=True)]),
               dict(bases=b"GGGCCC", taxonomy_id=0, genbank_id=0, is_plasmid=True, is_16s=False, locus_tag=b"", genes=[]),
               dict(bases=b"TTTT", taxonomy_id=1, genbank_id=2, is_plasmid=False, is_16s=False, locus_tag=b"x y", genes=[
                   dict(gene_name=b"a", locus_tag=b"", protein_id=b"", product=b"p q", reference_sequence=b"AC", gene_id=1, start=0, stop=3, complement=False)] * 2)]
    p = tmp_path / "database"
    db.write_database(p, entries)
    got = db.read_database(p)
    assert got == entries
    text = p.read_bytes()
    assert text.startswith(b"22 serialization::archive 17 0 0 0 0 3 0 0 0 40 ACGT") and text.count(b" 0 0") >= 6
    bases, offs, tags, tax = db.flatten(got)
    assert offs.tolist() == [0, 40, 46, 50] and tags == [b"NC_1", b"", b"x y"] and tax.tolist() == [562, 0, 1]
    assert bytes(bases[40:46]) == b"GGGCCC"


def test_rejects_garbage(pkg, tmp_path):
    from kslam_b200 import database as db
    p = tmp_path / "database"
    for blob in (b"not an archive", b"22 serialization::archive 12 0 0 0 0 1 0 0 0 99 ACGT", b"22 serialization::archive 12 0 0 0 0 2 0 0 0 1 A 0 0 0 0 1 g 0 0 0 0"):
        p.write_bytes(blob)
        with pytest.raises(db.ArchiveError):
            db.read_database(p)


def test_archive_header_constants_match_a_boost_library(pkg, tmp_path):
    """The one piece of Boost.Serialization present in this image is a header-less libboost_serialization.so (1.78, bundled
    with Nsight Compute). It exports the archive signature and library version: the header our writers emit is
    `<len(signature)> <signature> <version>` with that signature, and the readers accept that library's version."""
    import ctypes as C
    import glob
    from kslam_b200 import database
    libs = glob.glob("/opt/nvidia/nsight-compute/*/host/*/libboost_serialization.so.1.78.0")
    if not libs:
        pytest.skip("no Boost serialization library in this image")
    L = C.CDLL(libs[0])
    sig = L._ZN5boost7archive23BOOST_ARCHIVE_SIGNATUREEv
    sig.restype = C.c_char_p
    ver = L._ZN5boost7archive21BOOST_ARCHIVE_VERSIONEv          # returns library_version_type through a hidden pointer
    ver.restype, ver.argtypes = C.c_void_p, [C.c_void_p]
    buf = (C.c_uint16 * 8)()
    ver(C.byref(buf))
    signature, version = sig(), int(buf[0])
    assert database.HEADER == b"%d %s" % (len(signature), signature) and version >= 17
    path = str(tmp_path / "database")
    database.write_database(path, [dict(bases=b"ACGT", locus_tag=b"x")], libver=version)
    assert open(path, "rb").read().startswith(b"22 serialization::archive %d 0 0" % version)
    assert database.read_database(path)[0]["bases"] == b"ACGT"
    ix = pkg.Index.read(path)
    assert ix.n_entries == 1 and ix.bases.tobytes() == b"ACGT" and ix.locus_tags == [b"x"]
