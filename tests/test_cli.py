"""CPU tier: the `SLAM` executable (k-slam_b200/csrc/slam_main.cpp) — the reference's command line (main.cpp:24-169) over the
C ABI. The database-building modes are host-only and run here; the alignment modes need the GPU (tests/test_gpu_cli.py) and
must stop loudly without one."""
import os
import subprocess

import numpy as np
import pytest

import _lib as T
from test_taxon_host import make_db

BIN = os.path.join(T.ROOT, "k-slam_b200", "SLAM")


def slam(*args, cwd=None):
    return subprocess.run([BIN, *map(str, args)], cwd=cwd, capture_output=True, timeout=600)


def test_version_help_and_bad_options(tmp_path):
    r = slam("--version", cwd=tmp_path)
    assert r.returncode == 1 and r.stdout == b"1.0\n"                       # main.cpp:89-92 returns 1
    for args in ((), ("--help",)):
        r = slam(*args, cwd=tmp_path)
        assert r.returncode == 1 and r.stdout.startswith(b"Usage\tSLAM [option] --db=DATABASE R1FILE R2FILE\n") and b"--num-reads-at-once" in r.stdout
    assert slam("--bogus", cwd=tmp_path).returncode == 2
    assert slam("--num-reads", "abc", "x.fq", cwd=tmp_path).returncode == 2
    assert b"ambiguous" in slam("--num", "3", cwd=tmp_path).stderr          # --num-reads / --num-reads-at-once / --num-alignments
    r = slam("--parse-taxonomy", "only_one", cwd=tmp_path)
    assert r.returncode == 1 and r.stdout == b"Provide names.dmp and nodes.dmp\n"


def test_database_building_modes(pkg, tmp_path):
    from kslam_b200 import database
    _, _, _, _, names, nodesf, taxdb, paths = make_db(pkg, tmp_path, n_strains=5, length=4000)
    out = tmp_path / "db"
    out.mkdir()
    assert slam("--parse-genbank", "--output-file", out / "database", *paths, cwd=tmp_path).returncode == 0
    assert T.index_entries(pkg.Index.read(str(out / "database"))) == T.index_entries(pkg.Index.parse_genbank(paths))
    assert slam("--parse-taxonomy", "--output-file=" + str(out / "taxDB"), names, nodesf, cwd=tmp_path).returncode == 0
    assert (out / "taxDB").read_bytes() == open(taxdb, "rb").read()
    fa = tmp_path / "g.fa"
    fa.write_bytes(b">a one\nacgt\nNNAC\n>b two\r\nGGGG\r\n")
    assert slam("--parse-fasta", "--output", out / "fasta_db", fa, cwd=tmp_path).returncode == 0     # unambiguous prefix of --output-file
    from kslam_b200 import slam as slam_py
    gb, go, tags = slam_py.parse_fasta([str(fa)])
    got = database.read_database(str(out / "fasta_db"))
    assert [e["bases"] for e in got] == [b"ACGTNNAC", b"GGGG"] == [gb[int(go[i]):int(go[i + 1])].tobytes() for i in range(2)]
    assert [e["locus_tag"] for e in got] == [b"a", b"b"] == tags
    assert (tmp_path / "log.txt").read_bytes().count(b"Parsing") == 2       # the reference's log file, same place, rewritten by every run
    assert slam("--parse-genbank", "--output-file", out / "x", tmp_path / "missing.gbff", cwd=tmp_path).returncode == 2


def test_alignment_without_gpu_fails_loudly(pkg, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by tests/test_gpu_cli.py")
    (tmp_path / "db").mkdir()
    fa = tmp_path / "g.fa"
    fa.write_bytes(b">a one\n" + b"ACGT" * 40 + b"\n")
    assert slam("--parse-fasta", "--output-file", tmp_path / "db" / "database", fa, cwd=tmp_path).returncode == 0
    fq = tmp_path / "r.fq"
    fq.write_bytes(b"@r0\n" + b"ACGT" * 20 + b"\n+\n" + b"I" * 80 + b"\n")
    r = slam("--db", tmp_path / "db", "--just-align", "--sam-file", tmp_path / "o.sam", fq, cwd=tmp_path)
    assert r.returncode == 3 and b"no CPU fallback" in r.stderr
    r = slam("--db", tmp_path / "db", fq, cwd=tmp_path)                      # metagenomic mode needs DB/taxDB before anything else
    assert r.returncode == 2 and b"taxDB" in r.stderr
