"""GPU tier, last file of the tier on purpose (a failure here must not stop the parity tests before it under `pytest -x`): the
reference's OWN executable next to ours on the same command lines. See tests/test_reference_binary.py for the CPU half."""
import os
import subprocess

import pytest

import _lib as T
from test_gpu_cli import BIN, write_fastq
from test_taxon_host import make_db, make_reads

pytestmark = pytest.mark.gpu

def _ref_executable():
    exe = os.path.join(T.ORACLE_DIR, "_ref", "SLAM_ref")
    try:
        return exe if subprocess.run([exe, "--version"], capture_output=True, timeout=60).stdout == b"1.0\n" else None
    except OSError:
        return None


@pytest.mark.skipif(_ref_executable() is None, reason="needs oracle/_ref/SLAM_ref (the reference's own executable, prebuilt; travels with the snapshot)")
def test_both_executables_same_command_lines(pkg, tmp_path):
    """Process interface against process interface: the reference's own executable (oracle/_ref/SLAM_ref, its unmodified main.cpp;
    one OpenMP thread, see include/kslam.h on combineTaxonomies) and ours, same database directory (written by ours), same FASTQ
    files, same command lines — every output file byte for byte; the SAM header apart from the program path in its @PG line."""
    ref_exe = _ref_executable()
    gb, go, _, _, names, nodesf, taxdb, paths = make_db(pkg, tmp_path, n_strains=20, length=12_000)
    db = tmp_path / "db"
    db.mkdir()

    def run(exe, *a, cwd):
        return subprocess.run([exe, *map(str, a)], cwd=cwd, capture_output=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert run(BIN, "--parse-genbank", "--output-file", db / "database", *paths, cwd=tmp_path).returncode == 0
    assert run(BIN, "--parse-taxonomy", "--output-file", db / "taxDB", names, nodesf, cwd=tmp_path).returncode == 0
    n_pairs = 400
    rb, ro, quals, _, _ = make_reads(pkg, gb, go, n_pairs, seed=77)
    ids = [b"p%d" % (i % n_pairs) for i in range(2 * n_pairs)]
    r1, r2 = tmp_path / "R1.fq", tmp_path / "R2.fq"
    write_fastq(r1, rb, ro, quals, ids, 0, n_pairs, 1)
    write_fastq(r2, rb, ro, quals, ids, n_pairs, 2 * n_pairs, 2)
    commands = [
        (["--db", db, "--sam-file", "o.sam", "--output-file", "o.xml", "--num-reads-at-once", 150, r1, r2], ["o.sam", "o.xml", "o.xml_PerRead", "o.xml_abbreviated"]),
        (["--db=" + str(db), "--num-reads", 250, "--num-reads-at-once", 100, "--no-pseudo-assembly", "--score-fraction-threshold", 0.5, r1, r2], ["_PerRead"]),
        (["--db", db, "--just-align", "--sam-file", "s.sam", "--num-reads-at-once", 300, "--num-alignments", 3, r1], ["s.sam"]),
        # everything in one go (main.cpp:146-166 takes metagenomicAnalysis then, whose per-read file has no underscore, SLAM.h:142)
        (["--db", db, "--sam-file", "a.sam", "--output-file", "a.xml", "--num-reads-at-once", 4294967295, "--sam-xa", r1, r2],
         ["a.sam", "a.xml", "a.xmlPerRead", "a.xml_abbreviated"]),
    ]
    for k, (args, files) in enumerate(commands):
        outs = []
        for exe in (ref_exe, BIN):
            d = tmp_path / f"run{k}_{'ref' if exe == ref_exe else 'ours'}"
            d.mkdir()
            r = run(exe, *args, cwd=d)
            assert r.returncode == 0, (exe, r.stderr)
            got = {"stdout": r.stdout}
            for f in files:
                text = (d / f).read_bytes()
                if f.endswith(".sam"):                         # @PG CL:"<program path> ..." differs by the program path only
                    head, body = text[:text.index(b"@PG")], text[text.index(b"@PG"):].split(b"\n", 1)[1]
                    got[f] = (head, body)
                else:
                    got[f] = text
            outs.append(got)
        assert outs[0] == outs[1], (k, [name for name in outs[0] if outs[0][name] != outs[1][name]])
        assert any(len(v[1] if isinstance(v, tuple) else v) > 1000 for v in outs[0].values())
