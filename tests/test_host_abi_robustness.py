"""CPU tier: the host-side entry points added around the path (taxonomy, database index, batch outputs) must refuse bad
input with an error code — never crash, never read past a buffer."""
import ctypes as C

import numpy as np
import pytest

import _lib as T
from test_taxon_host import make_db


def test_archive_reader_survives_truncation_and_garbage(pkg, tmp_path):
    *_, paths = make_db(pkg, tmp_path, n_strains=3, length=1500)
    good = tmp_path / "database"
    pkg.Index.parse_genbank(paths).write(str(good))
    data = good.read_bytes()
    want = T.index_entries(pkg.Index.read(str(good)))
    rng = np.random.default_rng(9)
    cuts = sorted(set(rng.integers(0, len(data), size=150).tolist() + [0, 1, 25, 26, len(data) - 1]))
    bad = tmp_path / "cut"
    ok = 0
    for c in cuts:
        bad.write_bytes(data[:c])
        try:
            got = T.index_entries(pkg.Index.read(str(bad)))
            ok += 1
            assert len(got) <= len(want)
        except pkg.KslamError:
            pass
    assert ok < len(cuts)                                   # most truncations are detected
    for k in range(60):                                     # one byte replaced: either an error or a parse, never a crash
        b = bytearray(data)
        pos = int(rng.integers(0, min(len(b), 4000)))
        b[pos] = int(rng.integers(0, 256))
        bad.write_bytes(bytes(b))
        try:
            pkg.Index.read(str(bad))
        except pkg.KslamError:
            pass
    bad.write_bytes(b"22 serialization::archive 17 0 0 0 0 1 0 0 0 18446744073709551615 ACGT")    # a length that wraps
    with pytest.raises(pkg.KslamError):
        pkg.Index.read(str(bad))


def test_taxonomy_database_errors(pkg, tmp_path):
    bad = tmp_path / "taxDB"
    bad.write_bytes(b"not a number\n1\nroot\nno rank\n")
    with pytest.raises(pkg.KslamError):
        pkg.TaxDb(str(bad))
    bad.write_bytes(b"")                                    # empty file: an empty tree, every query answers "unknown"
    db = pkg.TaxDb(str(bad))
    assert len(db) == 0 and db.lca([5, 6]) == 0 and db.lca([7, 7]) == 7 and db.name(7) == b"" and db.lineage(7) == b""
    with pytest.raises(pkg.KslamError):
        pkg.TaxDb.build(str(tmp_path / "no_names"), str(tmp_path / "no_nodes"), str(tmp_path / "out"))
    loop = tmp_path / "loop"
    loop.write_bytes(b"5\n6\na\nspecies\n6\n5\nb\ngenus\n")    # a parent cycle would hang the reference; here the walk is bounded
    db = pkg.TaxDb(str(loop))
    assert db.lca([5, 6]) in (0, 5, 6) and isinstance(db.lineage(5), bytes)


def test_batch_outputs_argument_checks(pkg, tmp_path):
    L = pkg.lib()
    gb, go = pkg.synth.random_genomes(2, 1000, seed=1)
    w = pkg.SamWriter(gb, go, ["a", "b"])
    z8, z1 = np.zeros(0, np.uint8), np.zeros(1, np.uint64)
    args = (z8, z1, z8, z1, z8, z1, np.zeros(0, pkg.OVERLAP_DT), np.zeros(0, np.uint32), np.zeros(0, pkg.PAIR_DT))
    (tmp_path / "taxDB").write_bytes(b"1\n1\nroot\nno rank\n")
    db, taxa = pkg.TaxDb(str(tmp_path / "taxDB")), pkg.Taxa()
    with pytest.raises(pkg.KslamError):
        w.batch(*args, taxdb=db)                            # a taxonomy database without a result accumulator
    with pytest.raises(pkg.KslamError):
        w.batch(*args, taxa=taxa)
    text, _ = w.batch(*args, taxdb=db, taxa=taxa)
    assert text == b"" and taxa.results(db, 0)[0] == b""
    w.db.genes = 1                                          # a gene table without its offsets
    with pytest.raises(pkg.KslamError):
        w.batch(*args)
    assert L.kslam_taxdb_open(None, None) != 0 and L.kslam_index_read(None, None) != 0 and L.kslam_taxa_create(None) != 0
    assert L.kslam_taxdb_lca(None, None, 0) == 0 and L.kslam_taxdb_size(None) == 0
    L.kslam_taxdb_close(None); L.kslam_taxa_destroy(None); L.kslam_index_free(None)
