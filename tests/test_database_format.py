"""CPU tier: DIR/database (Boost text archive of GenbankIndex): SURVEY's worked example, round trips, and the
bytes the REAL Boost.Serialization library writes for the same index (oracle/boost_archive_probe.cpp over the header-less
libboost_serialization.so 1.78 that ships inside Nsight Compute — this image has no Boost headers)."""
import os

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


def _probe():
    import os
    import _lib as T
    p = os.path.join(T.ORACLE_DIR, "_ref", "boost_archive_probe")
    return p if os.path.exists(p) else None


def _dump(entries):
    """The index dump format of oracle/ref_driver.cpp / boost_archive_probe.cpp from database.read_database's dicts."""
    F, R = b"\x1f", b"\x1e"
    out = []
    for e in entries:
        out.append(F.join([b"E", e["locus_tag"], b"%d" % e["taxonomy_id"], b"%d" % e["genbank_id"], b"%d" % int(e["is_plasmid"]),
                           b"%d" % int(e["is_16s"]), e["bases"]]) + R)
        for g in e["genes"]:
            out.append(F.join([b"G", g["gene_name"], g["locus_tag"], g["protein_id"], g["product"], g["reference_sequence"], b"%d" % g["gene_id"],
                               b"%d" % g["start"], b"%d" % g["stop"], b"%d" % int(g["complement"])]) + R)
    return b"".join(out)


@pytest.mark.skipif(_probe() is None, reason="needs oracle/_ref/boost_archive_probe (built where the Boost serialization library exists)")
def test_archive_bytes_equal_the_real_boost_library(pkg, tmp_path):
    """DIR/database pinned: the same index written by kslam_index_write / database.write_database and by the REAL
    Boost.Serialization 1.78 library (oracle/boost_archive_probe.cpp drives its save_object / text_oarchive machinery) —
    byte for byte, the library version in the header aside — and the real library's archive read back by both readers.
    GenBank database (genes, taxonomy ids, GI numbers, products with spaces), FASTA database (no genes), empty database."""
    import subprocess
    import _lib as T
    from kslam_b200 import database
    from test_taxon_host import make_db
    *_, paths = make_db(pkg, tmp_path, n_strains=4, length=3000)
    fa = tmp_path / "g.fa"
    fa.write_bytes(b">a one\nacgt\nNNAC\n>b two\nGGGG\n>c x\nAC GT\n")
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b"")
    for name, ix in (("genbank", pkg.Index.parse_genbank(paths)), ("fasta", pkg.Index.parse_fasta([str(fa)])), ("empty", pkg.Index.parse_fasta([str(empty)]))):
        ours, real, dump = tmp_path / (name + ".ours"), tmp_path / (name + ".real"), tmp_path / (name + ".dump")
        ix.write(str(ours))
        entries = database.read_database(str(ours))
        dump.write_bytes(_dump(entries))
        subprocess.run([_probe(), str(dump), str(real)], check=True)
        want = real.read_bytes()
        assert want.startswith(b"22 serialization::archive 19 ")
        assert ours.read_bytes().replace(b"archive 17 ", b"archive 19 ", 1) == want, name
        py = tmp_path / (name + ".py")
        database.write_database(str(py), entries, libver=19)
        assert py.read_bytes() == want, name
        assert database.read_database(str(real)) == entries
        assert T.index_entries(pkg.Index.read(str(real))) == T.index_entries(ix)
        if name == "genbank":
            assert sum(len(e["genes"]) for e in entries) > 4 and any(e["genbank_id"] for e in entries) and any(b" " in g["product"] for e in entries for g in e["genes"])


def test_archive_golden_from_the_boost_library(pkg, golden, tmp_path):
    """The same check from a fixture (tests/golden/make_golden.py boost): an archive the real Boost 1.78 library wrote for a
    small GenBank database travels with the repo; parsing the GenBank text and writing the index must give those bytes."""
    g = golden("database_boost178.npz")
    gbff = tmp_path / "db.gbff"
    gbff.write_bytes(g["gbff"].tobytes())
    ix = pkg.Index.parse_genbank([str(gbff)])
    out = tmp_path / "database"
    ix.write(str(out))
    want = g["archive"].tobytes()
    assert out.read_bytes().replace(b"archive 17 ", b"archive 19 ", 1) == want and len(want) > 5000
    import _lib as T
    real = tmp_path / "real"
    real.write_bytes(want)
    assert T.index_entries(pkg.Index.read(str(real))) == T.index_entries(ix)


def _refdb():
    import os
    import _lib as T
    p = os.path.join(T.ORACLE_DIR, "_ref", "libkslam_refdb.so")
    return p if os.path.exists(p) else None


@pytest.mark.skipif(_refdb() is None, reason="needs oracle/_ref/libkslam_refdb.so (reference sources + the Boost serialization library)")
def test_database_file_equals_the_reference_writing_through_real_boost(pkg, tmp_path):
    """The reference's OWN `--parse-genbank` / `--parse-fasta` code path end to end — createIndexFromGBFF / createIndexFromFASTA
    and its unmodified writeIndexToBoostSerial, with boost::archive::text_oarchive forwarding to the real Boost 1.78 library
    (oracle/ref_shim_boost) — against the file `SLAM --parse-genbank` / `--parse-fasta` writes: byte for byte, the
    library-version token aside. Member order, class nesting and every parsed field come from the reference's code here."""
    import ctypes as C
    import os
    import subprocess
    import _lib as T
    from test_taxon_host import make_db
    L = C.CDLL(_refdb())
    L.kref_write_database.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_uint64, C.c_char_p]
    *_, taxdb, paths = make_db(pkg, tmp_path, n_strains=6, length=5000)
    fa1, fa2 = tmp_path / "a.fa", tmp_path / "b.fa"
    fa1.write_bytes(b"acgtn\n>g1 first genome\nACGT\nacgtnn\n\n>nospace\nTTTT\n>g3 \n>g4 empty above\nGG\n")
    fa2.write_bytes(b">c1 crlf\r\nACGT\r\nAC\r\n>c2 cr only\rGGCC\rTT")
    exe = os.path.join(T.ROOT, "k-slam_b200", "SLAM")
    cwd = os.getcwd()
    work = tmp_path / "refcwd"
    work.mkdir()
    (work / "taxDB").write_bytes(open(taxdb, "rb").read())          # createIndexFromGBFF opens ./taxDB (GenbankTools.h:483)
    for kind, flag, files in ((0, "--parse-genbank", paths), (1, "--parse-fasta", [str(fa1), str(fa2)])):
        ref_out, our_out = str(tmp_path / f"ref{kind}"), str(tmp_path / f"ours{kind}")
        arr = (C.c_char_p * len(files))(*[os.fsencode(f) for f in files])
        os.chdir(work)
        try:
            assert L.kref_write_database(kind, arr, len(files), ref_out.encode()) == 0
        finally:
            os.chdir(cwd)
        assert subprocess.run([exe, flag, "--output-file", our_out, *files], cwd=tmp_path).returncode == 0
        want, got = open(ref_out, "rb").read(), open(our_out, "rb").read()
        assert want.startswith(b"22 serialization::archive 19 ") and len(want) > 100
        assert got.replace(b"archive 17 ", b"archive 19 ", 1) == want, flag
        assert T.index_entries(pkg.Index.read(ref_out)) == T.index_entries(pkg.Index.read(our_out))


@pytest.mark.skipif(_refdb() is None, reason="needs oracle/_ref/libkslam_refdb.so (reference sources + the Boost serialization library)")
def test_the_reference_reads_our_database_through_real_boost(pkg, tmp_path):
    """The other direction of the drop-in: the reference's OWN getIndexFromBoostSerial — boost::archive::text_iarchive forwarding
    to the real Boost 1.78 library, whose init() checks the header and whose load_object reads the class preambles — loads the
    file `SLAM --parse-genbank` wrote (library version 17 in the header) and ends up with the same GenbankIndex, genes included;
    a truncated file makes it throw."""
    import ctypes as C
    import os
    import subprocess
    import numpy as np
    import _lib as T
    from test_taxon_host import make_db
    L = C.CDLL(_refdb())
    L.kref_read_database.restype = C.c_uint64
    L.kref_read_database.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
    *_, paths = make_db(pkg, tmp_path, n_strains=5, length=4000)
    exe = os.path.join(T.ROOT, "k-slam_b200", "SLAM")
    db = str(tmp_path / "database")
    assert subprocess.run([exe, "--parse-genbank", "--output-file", db, *paths], cwd=tmp_path).returncode == 0
    cwd = os.getcwd()
    os.chdir(tmp_path)                                       # the reference logs to ./log.txt
    try:
        n = L.kref_read_database(db.encode(), None, 0)
        assert n != 2**64 - 1 and n > 20_000
        buf = np.zeros(n, np.uint8)
        L.kref_read_database(db.encode(), T._p(buf), n)
        got = T.decode_index_dump(bytes(buf))
        ours = T.index_entries(pkg.Index.read(db))
        assert len(got) == len(ours) == 5
        for g, w in zip(got, ours):
            assert g["locus_tag"] == w["locus_tag"] and g["taxonomy_id"] == w["taxonomy_id"] and g["bases"] == w["bases"] and g["genes"] == w["genes"]
        assert sum(len(e["genes"]) for e in got) > 10 and all(e["genbank_id"] > 0 for e in got)
        cut = str(tmp_path / "cut")
        open(cut, "wb").write(open(db, "rb").read()[:3000])
        assert L.kref_read_database(cut.encode(), None, 0) == 2**64 - 1
    finally:
        os.chdir(cwd)


def test_side_car_cache_of_the_database(pkg, tmp_path):
    """DIR/database.kslam: written after the first parse, used while its key (size + mtime of the archive)
    matches, ignored afterwards; an index that came from the cache is the same GenbankIndex (it writes the same archive)."""
    import time
    from test_taxon_host import make_db
    _, _, _, _, _, _, _, paths = make_db(pkg, tmp_path, n_strains=6, length=5000)
    ix = pkg.Index.parse_genbank(paths)
    db = str(tmp_path / "database")
    ix.write(db)
    want = open(db, "rb").read()
    cache = db + ".kslam"
    assert not os.path.exists(cache)
    a = pkg.Index.read(db)                                  # parses the text archive, leaves the side-car behind
    assert os.path.exists(cache) and os.path.getsize(cache) > 6 * 5000
    b = pkg.Index.read(db)                                  # comes from the side-car (mapped)
    for k, got in enumerate((a, b)):
        out = str(tmp_path / f"again{k}")
        got.write(out)
        assert open(out, "rb").read() == want, k
        assert got.gene_records(0) == ix.gene_records(0)
    a.close(); b.close()
    # a stale side-car (the archive changed) is ignored and replaced
    sub = tmp_path / "other"; sub.mkdir()
    ix2 = pkg.Index.parse_genbank(make_db(pkg, sub, n_strains=4, length=3000)[7])
    time.sleep(0.01)
    ix2.write(db)
    want2 = open(db, "rb").read()
    c = pkg.Index.read(db)
    out = str(tmp_path / "again2"); c.write(out)
    assert open(out, "rb").read() == want2 != want
    c.close()
    # a truncated side-car is ignored too
    open(cache, "r+b").truncate(100)
    d = pkg.Index.read(db)
    out = str(tmp_path / "again3"); d.write(out)
    assert open(out, "rb").read() == want2
    d.close()
    # a side-car whose key still matches but whose offset tables are damaged is dropped, not trusted
    good = open(cache, "rb").read()
    for at in (64 + 8, 64 + 8 * 3):                         # entry offsets of the bases table (header: 8 + 7 x 8 bytes)
        bad = bytearray(good); bad[at:at + 8] = (2**62).to_bytes(8, "little")
        open(cache, "wb").write(bytes(bad))
        e = pkg.Index.read(db)
        out = str(tmp_path / "again4"); e.write(out)
        assert open(out, "rb").read() == want2
        e.close()
    os.environ["KSLAM_NO_INDEX_CACHE"] = "1"
    try:
        os.unlink(cache)
        pkg.Index.read(db).close()
        assert not os.path.exists(cache)
    finally:
        del os.environ["KSLAM_NO_INDEX_CACHE"]
