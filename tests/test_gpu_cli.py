"""GPU tier: the whole drop-in, process interface to process interface. `SLAM --db DB --sam-file .. --output-file ..
R1.fq R2.fq` (database built by `SLAM --parse-genbank` / `--parse-taxonomy`) against the reference's own chain run on the
same FASTQ records batch by batch (oracle/_ref: its createIndexFromGBFF, alignToDatabase, getPairedOverlaps, host stages,
SAM.h, MetagenomicResults.h): SAM, XML, _PerRead and _abbreviated byte for byte."""
import os
import subprocess

import numpy as np
import pytest

import _lib as T
from test_taxon_host import make_db, make_reads

pytestmark = pytest.mark.gpu
BIN = os.path.join(T.ROOT, "k-slam_b200", "SLAM")


def write_fastq(path, rb, ro, quals, ids, lo, hi, mate):
    with open(path, "wb") as f:
        for i in range(lo, hi):
            a, b = int(ro[i]), int(ro[i + 1])
            f.write(b"@" + ids[i] + b"/%d extra words\n" % mate + rb[a:b].tobytes() + b"\n+\n" + quals[a:b].tobytes() + b"\n")


def reference_run(L, rt, gb, go, rb, ro, quals, ids, n_pairs, at_once, paired, want_sam, taxonomy, **kw):
    R, sam, n_reads = None, [], 0
    total = n_pairs if paired else 2 * n_pairs
    for lo in range(0, total, at_once):
        hi = min(total, lo + at_once)
        sel = list(range(lo, hi)) + ([n_pairs + i for i in range(lo, hi)] if paired else [])
        seqs = [rb[int(ro[i]):int(ro[i + 1])] for i in sel]
        b_rb = np.concatenate(seqs); b_ro = T.offsets_of(seqs)
        b_q = np.concatenate([quals[int(ro[i]):int(ro[i + 1])] for i in sel])
        b_ids = [ids[i] for i in sel]
        idb = np.frombuffer(b"".join(b_ids), np.uint8); ido = T.offsets_of(b_ids)
        if R is None:
            R = T.Ref(gb, go, b_rb, b_ro, T.default_params(report_cigar=int(want_sam)))
            L.kref_use_parsed_index(R.h)
        else:
            L.kref_set_reads(R.h, len(b_ro) - 1, T._p(T.u8(b_rb)), T._p(b_ro))
        R.align_to_database()
        if paired:
            R.screen_and_pair()
        else:
            L.kref_screen(R.h)
        n_reads += hi - lo
        sam.append(T.ref_meta_batch(R, rt if taxonomy else None, b_q, b_ro, idb, ido, paired=paired, want_sam=want_sam, **kw))
    outs = T.ref_meta_finish(R, rt, n_reads) if taxonomy else None
    hdr = T.ref_sam_header(R, "CMD")
    R.close()
    return hdr, b"".join(sam), outs


@pytest.mark.skipif(not T.have_ref(), reason="needs oracle/_ref (prebuilt, travels with the repo snapshot)")
def test_slam_executable_equals_reference(pkg, tmp_path):
    gb, go, _, _, names, nodesf, taxdb, paths = make_db(pkg, tmp_path, n_strains=20, length=12_000)
    db = tmp_path / "db"
    db.mkdir()
    run = lambda *a: subprocess.run([BIN, *map(str, a)], cwd=tmp_path, capture_output=True, timeout=900)   # noqa: E731
    assert run("--parse-genbank", "--output-file", db / "database", *paths).returncode == 0
    assert run("--parse-taxonomy", "--output-file", db / "taxDB", names, nodesf).returncode == 0
    L = T.ref()
    assert T.ref_parse_index(0, paths, taxdb) is not None
    rt = L.kref_taxdb_open(taxdb.encode())
    L.kref_set_threads(1)
    n_pairs = 400
    rb, ro, quals, _, _ = make_reads(pkg, gb, go, n_pairs, seed=77)
    ids = [b"p%d" % (i % n_pairs) for i in range(2 * n_pairs)]
    r1, r2 = tmp_path / "R1.fq", tmp_path / "R2.fq"
    write_fastq(r1, rb, ro, quals, ids, 0, n_pairs, 1)
    write_fastq(r2, rb, ro, quals, ids, n_pairs, 2 * n_pairs, 2)
    try:
        # paired, three batches, SAM + taxonomy
        r = run("--db", db, "--sam-file", "o.sam", "--output-file", "o.xml", "--num-reads-at-once", 150, r1, r2)
        assert r.returncode == 0, r.stderr
        hdr, sam, outs = reference_run(L, rt, gb, go, rb, ro, quals, ids, n_pairs, 150, True, True, True)
        got = (tmp_path / "o.sam").read_bytes()
        cmd = " ".join([BIN, "--db", str(db), "--sam-file", "o.sam", "--output-file", "o.xml", "--num-reads-at-once", "150", str(r1), str(r2)]).encode()
        assert got == hdr.replace(b'CL:"CMD"', b'CL:"' + cmd + b'"') + sam and sam.count(b"\n") > 800 and b"\tXG:Z:" in sam
        for suffix, want in zip(("_PerRead", "", "_abbreviated"), outs):
            assert (tmp_path / ("o.xml" + suffix)).read_bytes() == want, suffix
        assert outs[1].count(b"<taxon>") >= 3
        # the same run with every batch split over three contexts (--devices; here all on GPU 0): identical files
        r = run("--db", db, "--sam-file", "m.sam", "--output-file", "m.xml", "--num-reads-at-once", 150, "--devices", "0,0,0", r1, r2)
        assert r.returncode == 0, r.stderr
        body = lambda t: t[t.index(b"@PG"):].split(b"\n", 1)[1]   # noqa: E731
        assert body((tmp_path / "m.sam").read_bytes()) == sam
        for suffix, want in zip(("_PerRead", "", "_abbreviated"), outs):
            assert (tmp_path / ("m.xml" + suffix)).read_bytes() == want, suffix
        # the genome k-mer index range-partitioned over two GPUs, read k-mers routed by the library's own NCCL all-to-alls
        import torch
        if torch.cuda.device_count() >= 2:
            r = run("--db", db, "--sam-file", "q.sam", "--output-file", "q.xml", "--num-reads-at-once", 150, "--devices", "0,1", "--partition-index", 1, r1, r2)
            assert r.returncode == 0, r.stderr
            assert body((tmp_path / "q.sam").read_bytes()) == sam
            for suffix, want in zip(("_PerRead", "", "_abbreviated"), outs):
                assert (tmp_path / ("q.xml" + suffix)).read_bytes() == want, suffix
            assert b"range-partitioned over 2 devices" in (tmp_path / "log.txt").read_bytes()
        # paired, taxonomy only (no SAM: the records are not re-sorted before the taxonomy step), XML to stdout, --num-reads cut
        r = run("--db=" + str(db), "--num-reads", 250, "--num-reads-at-once", 100, "--no-pseudo-assembly", "--score-fraction-threshold", 0.5, r1, r2)
        assert r.returncode == 0, r.stderr
        # (the --num-reads cut changes the batch sizes to 100, 100, 50: run the reference on exactly those pairs)
        keep = list(range(250)) + [n_pairs + i for i in range(250)]
        seqs = [rb[int(ro[i]):int(ro[i + 1])] for i in keep]
        k_rb, k_ro = np.concatenate(seqs), T.offsets_of(seqs)
        k_q = np.concatenate([quals[int(ro[i]):int(ro[i + 1])] for i in keep])
        k_ids = [ids[i] for i in keep]
        _, _, outs = reference_run(L, rt, gb, go, k_rb, k_ro, k_q, k_ids, 250, 100, True, False, True, pseudo=False, fraction=0.5)
        assert r.stdout == outs[1] and (tmp_path / "_PerRead").read_bytes() == outs[0]
        # single-end, --just-align
        r = run("--db", db, "--just-align", "--sam-file", "s.sam", "--num-reads-at-once", 300, "--num-alignments", 3, r1)
        assert r.returncode == 0, r.stderr
        s_seqs = [rb[int(ro[i]):int(ro[i + 1])] for i in range(n_pairs)]
        s_rb, s_ro = np.concatenate(s_seqs), T.offsets_of(s_seqs)
        hdr, sam, _ = reference_run(L, rt, gb, go, s_rb, s_ro, quals[:len(s_rb)], ids[:n_pairs], n_pairs // 2, 300, False, True, False, num_alignments=3)
        got = (tmp_path / "s.sam").read_bytes()
        assert got[got.index(b"@PG"):].split(b"\n", 1)[1] == sam and len(sam) > 10_000
        r = run("--db", db, "--just-align", "--sam-file", "s2.sam", "--num-reads-at-once", 300, "--num-alignments", 3, "--devices", "0,0", r1)
        assert r.returncode == 0, r.stderr
        assert body((tmp_path / "s2.sam").read_bytes()) == sam
    finally:
        L.kref_set_threads(os.cpu_count() or 1)
        L.kref_taxdb_close(rt)

