"""CPU tier: the reference's OWN executable, built here from its unmodified main.cpp (oracle/_ref/SLAM_ref: boost::archive forwards to
the real Boost.Serialization library found in the image, boost::program_options is a small stand-in of ours) next to our `SLAM`:
  * the three --parse-* modes write the same files;
  * the reference executable aligns against a database directory written by OUR executable and gives exactly what the in-memory
    harness of the reference's functions gives (the chain every GPU parity test compares the CUDA path with) — so the harness is the
    reference's process behaviour, and our database directory is a drop-in for it.
The GPU tier (tests/test_gpu_zz_executables.py) runs both executables on the same command lines."""
import os
import subprocess

import numpy as np
import pytest

import _lib as T
from test_gpu_cli import reference_run, write_fastq
from test_taxon_host import make_db, make_reads

REF = os.path.join(T.ORACLE_DIR, "_ref", "SLAM_ref")
OURS = os.path.join(T.ROOT, "k-slam_b200", "SLAM")


def ref_binary_ok():
    try:
        return subprocess.run([REF, "--version"], capture_output=True, timeout=60).stdout == b"1.0\n"
    except OSError:
        return False


pytestmark = pytest.mark.skipif(not (T.have_ref() and ref_binary_ok()),
                                reason="needs oracle/_ref/SLAM_ref (reference sources + Boost serialization library)")


def run(exe, *args, cwd, threads=1):
    return subprocess.run([exe, *map(str, args)], cwd=cwd, capture_output=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS=str(threads)))


def test_parse_modes_write_the_same_files(pkg, tmp_path):
    _, _, _, _, names, nodesf, taxdb, paths = make_db(pkg, tmp_path, n_strains=6, length=5000)
    fa = tmp_path / "g.fa"
    fa.write_bytes(b"acgtn\n>g1 first genome\nACGT\nacgtnn\n\n>nospace\nTTTT\n>g4 x\nGG\r\n>c2 cr only\rGGCC\rTT")
    (tmp_path / "taxDB").write_bytes(open(taxdb, "rb").read())             # createIndexFromGBFF opens ./taxDB
    for flag, files in (("--parse-genbank", paths), ("--parse-fasta", [fa]), ("--parse-taxonomy", [names, nodesf])):
        a, b = tmp_path / ("ref" + flag), tmp_path / ("ours" + flag)
        assert run(REF, flag, "--output-file", a, *files, cwd=tmp_path).returncode == 0
        assert run(OURS, flag, "--output-file=" + str(b), *files, cwd=tmp_path).returncode == 0
        assert a.read_bytes() == b.read_bytes().replace(b"archive 17 ", b"archive 19 ", 1) and a.stat().st_size > 50, flag
    assert run(REF, "--version", cwd=tmp_path).returncode == run(OURS, "--version", cwd=tmp_path).returncode == 1


def test_reference_executable_on_our_database_equals_the_harness(pkg, tmp_path):
    gb, go, _, _, names, nodesf, taxdb, paths = make_db(pkg, tmp_path, n_strains=20, length=12_000)
    db = tmp_path / "db"
    db.mkdir()
    assert run(OURS, "--parse-genbank", "--output-file", db / "database", *paths, cwd=tmp_path).returncode == 0
    assert run(OURS, "--parse-taxonomy", "--output-file", db / "taxDB", names, nodesf, cwd=tmp_path).returncode == 0
    n_pairs = 400
    rb, ro, quals, _, _ = make_reads(pkg, gb, go, n_pairs, seed=77)
    ids = [b"p%d" % (i % n_pairs) for i in range(2 * n_pairs)]
    r1, r2 = tmp_path / "R1.fq", tmp_path / "R2.fq"
    write_fastq(r1, rb, ro, quals, ids, 0, n_pairs, 1)
    write_fastq(r2, rb, ro, quals, ids, n_pairs, 2 * n_pairs, 2)
    L = T.ref()
    assert T.ref_parse_index(0, paths, taxdb) is not None
    rt = L.kref_taxdb_open(taxdb.encode())
    L.kref_set_threads(1)
    try:
        r = run(REF, "--db", db, "--sam-file", "ref.sam", "--output-file", "ref.xml", "--num-reads-at-once", 150, r1, r2, cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        hdr, sam, outs = reference_run(L, rt, gb, go, rb, ro, quals, ids, n_pairs, 150, True, True, True)
        got = (tmp_path / "ref.sam").read_bytes()
        assert got[got.index(b"@PG"):].split(b"\n", 1)[1] == sam and got[:got.index(b"@PG")] == hdr[:hdr.index(b"@PG")]
        for suffix, want in zip(("_PerRead", "", "_abbreviated"), outs):
            assert (tmp_path / ("ref.xml" + suffix)).read_bytes() == want, suffix
        # taxonomy only, XML to stdout, --num-reads cut (batches of 100, 100, 50), other screens
        r = run(REF, "--db=" + str(db), "--num-reads", 250, "--num-reads-at-once", 100, "--no-pseudo-assembly", "--score-fraction-threshold", 0.5,
                r1, r2, cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        keep = list(range(250)) + [n_pairs + i for i in range(250)]
        k_seqs = [rb[int(ro[i]):int(ro[i + 1])] for i in keep]
        k_rb, k_ro = np.concatenate(k_seqs), T.offsets_of(k_seqs)
        k_q = np.concatenate([quals[int(ro[i]):int(ro[i + 1])] for i in keep])
        _, _, outs = reference_run(L, rt, gb, go, k_rb, k_ro, k_q, [ids[i] for i in keep], 250, 100, True, False, True, pseudo=False, fraction=0.5)
        assert r.stdout == outs[1] and (tmp_path / "_PerRead").read_bytes() == outs[0]
        # single-end, --just-align, other options
        r = run(REF, "--db", db, "--just-align", "--sam-file", "s.sam", "--num-reads-at-once", 300, "--num-alignments", 3, r1, cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        seqs = [rb[int(ro[i]):int(ro[i + 1])] for i in range(n_pairs)]
        s_rb, s_ro = np.concatenate(seqs), T.offsets_of(seqs)
        _, sam, _ = reference_run(L, rt, gb, go, s_rb, s_ro, quals[:len(s_rb)], ids[:n_pairs], n_pairs // 2, 300, False, True, False, num_alignments=3)
        got = (tmp_path / "s.sam").read_bytes()
        assert got[got.index(b"@PG"):].split(b"\n", 1)[1] == sam
    finally:
        L.kref_set_threads(os.cpu_count() or 1)
        L.kref_taxdb_close(rt)
