"""CPU tier: the metagenomic half of the batch loop and the database builders (host code in libkslam.so) against the
reference's OWN functions run here through oracle/_ref:
  * TaxonomyDB (--parse-taxonomy file, LCA, lineage)                      TaxonomyDatabase.h
  * createIndexFromGBFF / createIndexFromFASTA                             GenbankTools.h:224-260,348-527
  * SAM records with gene tags, per-read taxon assignment, XML / _PerRead / _abbreviated    SLAM.h:215-265, MetagenomicResults.h
and, for boxes without /root/reference, against golden files made by tests/golden/make_golden.py."""
import os

import numpy as np
import pytest

import _lib as T

needs_ref = pytest.mark.skipif(not T.have_ref(), reason="needs oracle/_ref (built where /root/reference exists)")


def make_db(pkg, tmp_path, n_strains=10, length=20_000, seed=5, files=2):
    """tree genomes + their taxonomy as GenBank flat files, names.dmp / nodes.dmp and our taxDB file."""
    gb, go = pkg.synth.tree_genomes(n_strains, length, seed=seed)
    nodes, strain_tax = pkg.synth.tree_taxonomy(n_strains)
    n_species = max(1, n_strains // 5)
    names, nodesf, taxdb = str(tmp_path / "names.dmp"), str(tmp_path / "nodes.dmp"), str(tmp_path / "taxDB")
    pkg.synth.write_taxonomy_dumps(nodes, names, nodesf)
    pkg.TaxDb.build(names, nodesf, taxdb)
    paths = [str(tmp_path / f"part{k}.gbff") for k in range(files)]
    handles = [open(p, "wb") for p in paths]
    for i in range(n_strains):
        rec = pkg.synth.genbank_text(gb[int(go[i]):int(go[i + 1])].tobytes(), f"NC_{i:06d}", 7_000_000 + i, int(strain_tax[i]),
                                     f"Synthetic strain {i}", i, i % n_species, seed=seed + i % n_species)   # same gene layout within a species
        handles[i * files // n_strains].write(rec)
    [h.close() for h in handles]
    return gb, go, nodes, strain_tax, names, nodesf, taxdb, paths


def make_reads(pkg, gb, go, n_pairs, seed):
    rb, ro, _ = pkg.synth.paired_reads(gb, go, n_pairs, seed=seed)
    rng = np.random.default_rng(seed)
    quals = rng.integers(35, 75, size=len(rb), dtype=np.uint8)
    n = len(ro) - 1
    ids = [b"read<%d>&x" % (i % (n // 2)) for i in range(n)]            # both mates share the id, with characters XML must escape
    idb = np.frombuffer(b"".join(ids), np.uint8); ido = np.zeros(n + 1, np.uint64); ido[1:] = np.cumsum([len(x) for x in ids])
    return rb, ro, quals, idb, ido


@needs_ref
def test_taxonomy_database_equals_reference(pkg, tmp_path):
    _, _, nodes, strain_tax, names, nodesf, taxdb, _ = make_db(pkg, tmp_path, n_strains=40, length=1000, files=1)
    L = T.ref()
    ref_file = str(tmp_path / "taxDB_ref")
    assert L.kref_taxdb_build(names.encode(), nodesf.encode(), ref_file.encode()) == 0
    assert open(taxdb, "rb").read() == open(ref_file, "rb").read()          # same node order: same container, same inserts
    rt = L.kref_taxdb_open(ref_file.encode())
    db = pkg.TaxDb(taxdb)
    assert len(db) == L.kref_taxdb_size(rt) == len(nodes)
    ids = [n[0] for n in nodes] + [0, 999_999]

    def ref_text(i, which):
        buf = np.zeros(4096, np.uint8)
        n = L.kref_lineage(rt, i, which, T._p(buf), len(buf))
        return bytes(buf[:n])
    for i in ids:
        assert db.lineage(i) == ref_text(i, 0) and db.name(i) == ref_text(i, 1), i
    rng = np.random.default_rng(3)
    for _ in range(3000):
        k = int(rng.integers(1, 6))
        pick = rng.choice(ids, size=k).astype(np.uint32) if rng.random() < 0.3 else rng.choice(strain_tax, size=k).astype(np.uint32)
        assert db.lca(pick) == L.kref_lca(rt, T._p(pick), len(pick)), pick
    assert db.lca(np.zeros(0, np.uint32)) == 0
    assert db.lca(strain_tax[:1]) == int(strain_tax[0])
    L.kref_taxdb_close(rt)


def test_taxonomy_known_answers(pkg, tmp_path):
    """Runs without the reference: LCA / lineage on the synthetic tree, by construction."""
    _, _, nodes, strain_tax, _, _, taxdb, _ = make_db(pkg, tmp_path, n_strains=40, length=1000, files=1)
    db = pkg.TaxDb(taxdb)
    n_species, n_genera = 8, 2
    assert db.lca([strain_tax[0], strain_tax[8]]) == 10000                  # two strains of species 0
    assert db.lca([strain_tax[0], strain_tax[2]]) == 2000                   # species 0 and 2 share genus 0
    assert db.lca([strain_tax[0], strain_tax[1]]) == 1000                   # genera 0 and 1 share the only phylum
    assert db.lca([strain_tax[0], 999_999]) == 0                            # an unknown id has a path of its own
    assert db.name(2000) == b"Genus0" and db.name(5) == b""
    assert db.lineage(int(strain_tax[0])) == b"Bacteria; Phylum0 <synthetic>; Genus0."    # cleared at the species node, 131567 skipped
    with pytest.raises(pkg.KslamError):
        pkg.TaxDb(str(tmp_path / "missing"))


def same_entries(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in ("locus_tag", "taxonomy_id", "bases"):
            assert g[k] == w[k], k
        assert len(g["genes"]) == len(w["genes"])
        for a, b in zip(g["genes"], w["genes"]):
            assert a == b


@needs_ref
def test_genbank_and_fasta_parsers_equal_reference(pkg, tmp_path):
    gb, go, _, strain_tax, _, _, taxdb, paths = make_db(pkg, tmp_path, n_strains=6, length=8000)
    odd = tmp_path / "odd.gbff"                                              # CRLF record, feature without numbers, no GI, no taxon, junk
    odd.write_bytes(b"LOCUS       ODD1\r\nVERSION     ODD1.2\r\nFEATURES             Location/Qualifiers\r\n     source          1..30\r\n"
                    b"     CDS             join(5..12,20..28)\r\n                     /product=\"split\r\n                     product\"\r\n"
                    b"     gene            <1..>9\r\n                     /gene=\"g1\"\r\n     CDS             order\r\n"
                    b"ORIGIN\r\n        1 acgtacgtnn acgtacgtac ryacgtacgt\r\n//\r\n"
                    b"LOCUS       ODD2\nVERSION     ODD2.1  GI:42 extra\n\n     source          1..10\n                     /db_xref=\"taxon:77\"\n"
                    b"                     /db_xref=\"taxon:88\"\nORIGIN\n        1 aaaaacccccgggggttttt\n//\ntrailing junk\n")
    all_paths = paths + [str(odd)]
    want = T.ref_parse_index(0, all_paths, taxdb)
    ix = pkg.Index.parse_genbank(all_paths)
    got = T.index_entries(ix)
    same_entries(got, want)
    # "//\r" is not "//": the CRLF record never closes and runs into the next one, in the reference and here
    assert len(got) == 7 and sum(len(e["genes"]) for e in got) > 30
    assert [e["taxonomy_id"] for e in got[:6]] == strain_tax.tolist() and got[6]["taxonomy_id"] == 77 and got[6]["locus_tag"] == b"ODD2.1"
    assert got[6]["bases"].endswith(b"AAAAACCCCCGGGGGTTTTT") and b"\r" in got[6]["bases"]
    assert got[0]["bases"] == gb[:8000].tobytes()
    # FASTA: LF / CRLF / CR line ends, lower case, header without a space, empty lines, bases before the first header
    fa1, fa2 = tmp_path / "a.fa", tmp_path / "b.fa"
    fa1.write_bytes(b"acgtn\n>g1 first genome\nACGT\nacgtnn\n\n>nospace\nTTTT\n>g3 \n>g4 empty above\nGG\n")
    fa2.write_bytes(b">c1 crlf\r\nACGT\r\nAC\r\n>c2 cr only\rGGCC\rTT")
    want = T.ref_parse_index(1, [str(fa1), str(fa2)])
    fx = pkg.Index.parse_fasta([str(fa1), str(fa2)])
    got = T.index_entries(fx)
    same_entries(got, want)
    assert [e["locus_tag"] for e in got] == [b"", b"g1", b"", b"g4", b"c1", b"c2"]
    with pytest.raises(pkg.KslamError):
        pkg.Index.parse_genbank([str(tmp_path / "missing.gbff")])


def test_database_archive_round_trip(pkg, tmp_path):
    """DIR/database: the C++ writer / reader agree with database.py (the bytes are pinned against the real Boost library in tests/test_database_format.py)."""
    from kslam_b200 import database
    *_, paths = make_db(pkg, tmp_path, n_strains=4, length=3000)
    ix = pkg.Index.parse_genbank(paths)
    f1, f2 = str(tmp_path / "database"), str(tmp_path / "database_py")
    ix.write(f1)
    back = pkg.Index.read(f1)
    assert T.index_entries(back) == T.index_entries(ix)
    py = database.read_database(f1)
    assert [e["bases"] for e in py] == [e["bases"] for e in T.index_entries(ix)]
    assert [[g["product"] for g in e["genes"]] for e in py] == [[g["product"] for g in e["genes"]] for e in T.index_entries(ix)]
    database.write_database(f2, py)
    assert open(f1, "rb").read() == open(f2, "rb").read()
    bad = tmp_path / "bad"
    bad.write_bytes(b"22 serialization::archive 17 0 0 0 0 1 0 0 0 99 ACGT")
    with pytest.raises(pkg.KslamError):
        pkg.Index.read(str(bad))


def run_ours_compact(pkg, ix, taxdb_path, batches, **kw):
    """The same run without --sam-file on COMPACT pair records (kslam_batch_outputs_compact)."""
    w = pkg.SamWriter(index=ix, num_alignments=kw.get("num_alignments", 10), score_fraction_threshold=kw.get("fraction", 0.95),
                      pseudo_assembly=kw.get("pseudo", True), report_cigar=False)
    db, taxa = pkg.TaxDb(taxdb_path), pkg.Taxa()
    n_reads, limits = 0, []
    for (rb, ro, quals, idb, ido, ov, pool, pairs) in batches:
        compact, limit, far = pkg.compact_pairs_host(ov, pairs, (len(ro) - 1) // 2)
        limits.append((limit, w.batch_compact(ro, idb, ido, compact, limit, far, taxdb=db, taxa=taxa), len(far)))
        n_reads += (len(ro) - 1) // 2
    return taxa.results(db, n_reads), limits


def run_ours(pkg, ix, taxdb_path, batches, want_sam, paired=True, **kw):
    w = pkg.SamWriter(index=ix, num_alignments=kw.get("num_alignments", 10), score_fraction_threshold=kw.get("fraction", 0.95),
                      pseudo_assembly=kw.get("pseudo", True), report_cigar=True, sam_xa=kw.get("sam_xa", False))
    db, taxa = pkg.TaxDb(taxdb_path), pkg.Taxa()
    sam, n_reads = [], 0
    for (rb, ro, quals, idb, ido, ov, pool, pairs) in batches:
        if paired:
            text, _ = w.batch(rb, ro, quals, ro, idb, ido, ov, pool, pairs, want_sam=want_sam, taxdb=db, taxa=taxa)
            n_reads += (len(ro) - 1) // 2
        else:
            text = w.batch_single(rb, ro, quals, ro, idb, ido, ov, pool, want_sam=want_sam, taxdb=db, taxa=taxa)
            n_reads += len(ro) - 1
        sam.append(text)
    return b"".join(sam), taxa.results(db, n_reads)


@needs_ref
@pytest.mark.parametrize("want_sam,paired,par_min", [(True, True, None), (False, True, None), (True, False, None), (False, False, None),
                                                     (True, True, 8), (False, False, 8)])
def test_metagenomic_outputs_equal_reference(pkg, tmp_path, want_sam, paired, par_min, monkeypatch):
    """Two batches through SLAM.h:209-265 in the reference (database = what ITS createIndexFromGBFF parsed) and through
    kslam_batch_outputs + kslam_taxa_results on the same alignments: SAM text with XG / XP / XR / XT, _PerRead, XML,
    _abbreviated byte for byte. One OpenMP thread in the reference: combineTaxonomies' parallel sort is thread-count
    dependent on ties (include/kslam.h). par_min forces the all-threads forms of the host stages and of the end-of-run merge
    (parallel read-name sorts, XML blocks) onto this small run."""
    if par_min is not None:
        monkeypatch.setenv("KSLAM_HOST_PAR_MIN", str(par_min))
    gb, go, _, _, _, _, taxdb, paths = make_db(pkg, tmp_path, n_strains=20, length=12_000)
    L = T.ref()
    assert T.ref_parse_index(0, paths, taxdb) is not None
    ix = pkg.Index.parse_genbank(paths)
    assert ix.bases.tobytes() == gb.tobytes()
    rt = L.kref_taxdb_open(taxdb.encode())
    L.kref_set_threads(1)
    try:
        for kw in (dict(), dict(pseudo=False, fraction=0.5), dict(num_alignments=2, fraction=0.0)):
            R, batches, want_sam_text, n_reads = None, [], [], 0
            for b in range(2):
                rb, ro, quals, idb, ido = make_reads(pkg, gb, go, 300, seed=20 + b)
                if R is None:
                    R = T.Ref(gb, go, rb, ro, T.default_params(report_cigar=1))
                    L.kref_use_parsed_index(R.h)
                else:
                    L.kref_set_reads(R.h, len(ro) - 1, T._p(T.u8(rb)), T._p(np.ascontiguousarray(ro, dtype=np.uint64)))
                ov_all, pool_all = R.align_to_database()
                if paired:
                    ov, pool, pairs = R.screen_and_pair()
                    n_reads += (len(ro) - 1) // 2
                else:
                    L.kref_screen(R.h)
                    ov, pool, pairs = ov_all, pool_all, None
                    n_reads += len(ro) - 1
                want_sam_text.append(T.ref_meta_batch(R, rt, quals, ro, idb, ido, paired=paired, want_sam=want_sam, **kw))
                batches.append((rb, ro, quals, idb, ido, ov, pool, pairs))
            want = T.ref_meta_finish(R, rt, n_reads)
            R.close()
            got_sam, got = run_ours(pkg, ix, taxdb, batches, want_sam, paired=paired, **kw)
            assert got_sam == b"".join(want_sam_text)
            if want_sam:
                assert b"\tXG:Z:" in got_sam and b"\tXP:Z:WP_" in got_sam and b"\tXR:Z:\"" in got_sam and b"\tXT:i:1000" in got_sam
            for name, g, w in zip(("_PerRead", "xml", "_abbreviated"), got, want):
                if g != w:
                    a, b_ = g.split(b"\n"), w.split(b"\n")
                    bad = [(i, x, y) for i, (x, y) in enumerate(zip(a, b_)) if x != y][:3]
                    raise AssertionError((name, kw, len(a), len(b_), bad))
            assert got[1].count(b"<taxon>") >= 3 and b"<gene protein=\"WP_" in got[1] and b"&lt;" in got[1] and len(got[0]) > 1000
            if paired and not want_sam:
                got_c, limits = run_ours_compact(pkg, ix, taxdb, batches, **kw)
                assert got_c == got, "compact pair records give other result files"
                assert all(a == b for a, b, _ in limits)
                # pairs far beyond the insert-size limit are split into their mates (PairedOverlap.h:396-436): the full
                # records read them from the alignment vector, the compact ones from the far-mates table
                far_batches = []
                for (rb, ro, quals, idb, ido, ov, pool, pairs) in batches:
                    pr = pairs.copy()
                    both = np.flatnonzero((pr["r1_idx"] >= 0) & (pr["r2_idx"] >= 0))
                    pr["insert_size"][both[::7]] = 40_000
                    far_batches.append((rb, ro, quals, idb, ido, ov, pool, pr))
                _, want_far = run_ours(pkg, ix, taxdb, far_batches, False, paired=True, **kw)
                got_far, limits = run_ours_compact(pkg, ix, taxdb, far_batches, **kw)
                assert got_far == want_far and all(n_far > 0 for _, _, n_far in limits)
    finally:
        L.kref_set_threads(os.cpu_count() or 1)
        L.kref_taxdb_close(rt)


def test_metagenomic_golden(pkg, golden, tmp_path):
    """Inputs + the reference's outputs from tests/golden/make_golden.py (travels to boxes without /root/reference)."""
    g = golden("meta_mini.npz")
    gbff, taxdb = tmp_path / "db.gbff", tmp_path / "taxDB"
    gbff.write_bytes(g["gbff"].tobytes()); taxdb.write_bytes(g["taxdb"].tobytes())
    ix = pkg.Index.parse_genbank([str(gbff)])
    batch = (g["rb"], g["ro"], g["quals"], g["ids"], g["id_offs"], g["ov"], g["pool"], g["pairs"])
    sam, (per_read, xml, abbreviated) = run_ours(pkg, ix, str(taxdb), [batch], True)
    assert sam == g["sam"].tobytes()
    assert per_read == g["per_read"].tobytes() and xml == g["xml"].tobytes() and abbreviated == g["abbreviated"].tobytes()


def test_gene_lookup_index_equals_full_scan(pkg, golden):
    """kslam_gene_index (binary search + running maximum of the CDS stops) against GenbankEntry::getGene's scan over every
    gene, on gene tables made to be hostile: nested and abutting genes, many equal overlaps (the FIRST gene must win), an
    entry whose genes are not ordered by start (keeps the scan), an entry without genes."""
    g = golden("meta_mini.npz")
    gbff = g["gbff"].tobytes()
    n_entries = gbff.count(b"\n//\n")
    rng = np.random.default_rng(4)
    genes, offs, strings = [], [0], bytearray()
    for e in range(n_entries):
        n = 0 if e == 1 else 120
        starts = np.sort(rng.integers(0, 5900, size=n)) if e != 2 else rng.integers(0, 5900, size=n)   # entry 2: unordered
        for k in range(n):
            length = int(rng.choice([30, 30, 30, 90, 400, 2500]))          # equal lengths give equal overlaps
            rec = np.zeros(1, pkg.GENE_DT)[0]
            rec["cds_start"], rec["cds_stop"], rec["gene_id"] = int(starts[k]), int(starts[k]) + length, 1000 * e + k
            so = []
            for txt in (b"g%d_%d" % (e, k), b"l%d_%d" % (e, k), b"P%d_%d" % (e, k), b"prod %d" % k, b"ref%d" % e):
                so.append(len(strings)); strings += txt
            so.append(len(strings))
            rec["str_offs"] = so
            genes.append(rec)
        offs.append(len(genes))
    genes = np.array(genes, dtype=pkg.GENE_DT)
    import tempfile, pathlib
    d = pathlib.Path(tempfile.mkdtemp())
    (d / "db.gbff").write_bytes(gbff)
    base = pkg.Index.parse_genbank([str(d / "db.gbff")])
    tags = base.locus_tags
    outs = []
    for use_index in (True, False):
        w = pkg.SamWriter(base.bases, base.offs, tags, taxonomy_ids=base.taxonomy_ids, report_cigar=True,
                          genes=(genes, np.array(offs, np.uint64), np.frombuffer(bytes(strings), np.uint8)), gene_index=use_index)
        assert bool(w.db.gene_index) == use_index
        text, _ = w.batch(g["rb"], g["ro"], g["quals"], g["ro"], g["ids"], g["id_offs"], g["ov"], g["pool"], g["pairs"])
        outs.append(text)
    assert outs[0] == outs[1] and outs[0].count(b"\tXG:Z:") > 500
