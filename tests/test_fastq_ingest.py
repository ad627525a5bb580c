"""CPU tier: the chunk-parallel FASTQ reader (kslam_fastq_*, host code of libkslam.so) against the reference's OWN
reader run here on the same files (oracle/_ref: getPairedSequencesFromFASTQFiles, FASTQsequence.h:110-165), and against
golden records that travel with the repo for boxes without /root/reference."""
import os

import numpy as np
import pytest

import _lib as T


def write(path, data: bytes):
    with open(path, "wb") as f:
        f.write(data)


def fastq_bytes(n, seed, eol=b"\n", read_len=(20, 160), final_newline=True, tricky=True):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.integers(*read_len))
        bases = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=L))
        qual = bytes(rng.integers(33, 74, size=L, dtype=np.uint8))
        if tricky and i % 7 == 3:
            qual = b"@" + qual[1:]                       # quality lines may start with '@'
        rid = [b"@r%d/1" % i, b"@r%d extra words/2" % i, b"@r%d/1 comment/x" % i, b"@", b"@ leading space", b"@x", b"r%dnoat" % i][i % 7] \
            if tricky else b"@r%d/1" % i
        plus = b"+" if i % 3 else b"+" + rid[1:]
        out.append(eol.join([rid, bases, plus, qual]))
    data = eol.join(out)
    return data + (eol if final_newline else b"")


def ours(pkg, r1, r2, max_reads, threads):
    got = []
    with pkg.FastqReader(r1, r2, threads=threads) as rd:
        while True:
            try:
                b = rd.next(max_reads)
            except pkg.KslamError as e:
                assert "mismatch in R1 and R2 size" in str(e)
                got.append("mismatch"); break
            if b is None:
                break
            got.append(b.records())
    return got


CASES = [dict(n1=57, n2=57, eol=b"\n", max_reads=1000), dict(n1=57, n2=57, eol=b"\r\n", max_reads=10),
         dict(n1=40, n2=40, eol=b"\r", max_reads=16), dict(n1=33, n2=33, eol=b"\n", max_reads=7, final_newline=False),
         dict(n1=5, n2=7, eol=b"\n", max_reads=100), dict(n1=9, n2=4, eol=b"\n", max_reads=100),
         dict(n1=12, n2=0, eol=b"\n", max_reads=5), dict(n1=0, n2=5, eol=b"\n", max_reads=5)]


def make_case(tmp_path, k, c):
    r1, r2 = str(tmp_path / f"c{k}_R1.fq"), str(tmp_path / f"c{k}_R2.fq")
    write(r1, fastq_bytes(c["n1"], 10 + k, c["eol"], final_newline=c.get("final_newline", True)))
    write(r2, fastq_bytes(c["n2"], 50 + k, c["eol"], final_newline=c.get("final_newline", True)))
    return r1, r2


@pytest.mark.skipif(not T.have_ref(), reason="needs oracle/_ref (built where /root/reference exists)")
@pytest.mark.parametrize("k", range(len(CASES)))
def test_reader_matches_the_reference_reader(pkg, tmp_path, k):
    c = CASES[k]
    r1, r2 = make_case(tmp_path, k, c)
    want = T.ref_read_fastq(r1, r2, c["max_reads"])
    for threads in (1, 5):
        assert ours(pkg, r1, r2, c["max_reads"], threads) == want
    want1 = T.ref_read_fastq(r1, None, c["max_reads"])                 # single-end
    assert ours(pkg, r1, None, c["max_reads"], 3) == want1


@pytest.mark.skipif(not T.have_ref(), reason="needs oracle/_ref")
def test_odd_line_structure(pkg, tmp_path):
    """Truncated last record, blank lines, mixed terminators, a trailing line without newline: lines are counted, never
    validated — exactly what the reference does."""
    blobs = [b"@a\nACGT\n+\nIIII\n@b\nAC\n+\n", b"@a\r\nACGT\n+\r\nIIII\r@b\nAAAA\n+\nIIII", b"\n\n@a\nACGT\n+\nIIII\n",
             b"@a\nACGT\n+\nIIII\n\n\n\n", b"", b"@only", b"@a\n\n+\n\n@b\nA\n+\nI\n", b"@a\r\r\nAC\n+\nII\n@b\nAC\n+\nII\n"]
    for k, blob in enumerate(blobs):
        p = str(tmp_path / f"odd{k}.fq")
        write(p, blob)
        for mr in (1, 2, 100):
            assert ours(pkg, p, None, mr, 4) == T.ref_read_fastq(p, None, mr), (k, mr)


@pytest.mark.skipif(not T.have_ref(), reason="needs oracle/_ref")
@pytest.mark.parametrize("final_newline", [True, False])
def test_long_reads_grow_the_scan_window(pkg, tmp_path, final_newline):
    """Records far longer than the first window estimate (512 bytes each): the single-scan path keeps what it has listed and
    scans on, over several threads, and the next batch sizes its window from the file's own line lengths."""
    r1, r2 = str(tmp_path / "long_R1.fq"), str(tmp_path / "long_R2.fq")
    write(r1, fastq_bytes(1500, 5, read_len=(600, 900), final_newline=final_newline))
    write(r2, fastq_bytes(1500, 6, read_len=(600, 900), final_newline=final_newline))
    for max_reads, threads in ((400, 4), (50, 3), (1499, 8), (5000, 2)):
        assert ours(pkg, r1, r2, max_reads, threads) == T.ref_read_fastq(r1, r2, max_reads), (max_reads, threads)


def test_reader_volume_and_batches(pkg, tmp_path):
    """20k pairs in batches of 3000: concatenated batches reproduce the files; R1 block precedes the R2 block."""
    n = 20_000
    r1, r2 = str(tmp_path / "big_R1.fq"), str(tmp_path / "big_R2.fq")
    write(r1, fastq_bytes(n, 1, tricky=False)); write(r2, fastq_bytes(n, 2, tricky=False))
    got = ours(pkg, r1, r2, 3000, 8)
    assert [len(b) for b in got] == [6000] * 6 + [4000]
    recs1 = [x for b in got for x in b[:len(b) // 2]]
    recs2 = [x for b in got for x in b[len(b) // 2:]]
    lines = open(r1, "rb").read().split(b"\n")
    assert [s for _, s, _ in recs1] == lines[1::4] and [q for _, _, q in recs1] == lines[3::4]
    assert [i for i, _, _ in recs1] == [ln[1:].split(b"/")[0] for ln in lines[0:-1:4]]
    assert len(recs2) == n and recs2[0][0] == b"r0"
    with pytest.raises(pkg.KslamError):
        pkg.FastqReader(str(tmp_path / "missing.fq"))


def test_golden_records(pkg, golden, tmp_path):
    """Fixture made from the reference reader (tests/golden/make_golden.py): travels to boxes without /root/reference."""
    g = golden("fastq_reader.npz")
    r1, r2 = str(tmp_path / "g_R1.fq"), str(tmp_path / "g_R2.fq")
    write(r1, g["r1"].tobytes()); write(r2, g["r2"].tobytes())
    got = ours(pkg, r1, r2, int(g["max_reads"]), 4)
    flat = [f for b in got for rec in b for f in rec]
    want = [bytes(x) for x in np.split(g["fields"], np.cumsum(g["field_lens"])[:-1])]
    assert flat == want and [len(b) for b in got] == g["batch_sizes"].tolist()


def test_ring_keeps_older_batches_valid(pkg, tmp_path):
    p = str(tmp_path / "ring.fq")
    write(p, fastq_bytes(90, 3, tricky=False))
    with pkg.FastqReader(p, None, threads=2, ring=3) as rd:
        views = [rd.next(10, copy=False) for _ in range(3)]          # three batches alive at once, no copies
        want = ours(pkg, p, None, 10, 2)
        assert [v.records() for v in views] == want[:3]
    with pkg.FastqReader(p, None) as rd:
        rd.next(10)
        assert rd.L.kslam_fastq_set_ring(rd.h, 2) != 0               # only before the first batch


def test_fifo_input_reads_like_a_file(pkg, tmp_path):
    """`SLAM ... <(zcat r1.fq.gz)`: a pipe has no size and cannot be mapped; the reader must drain it, not see an empty file."""
    import threading
    data1, data2 = fastq_bytes(300, 21, tricky=False), fastq_bytes(300, 22, tricky=False)
    f1, f2 = str(tmp_path / "p_R1.fq"), str(tmp_path / "p_R2.fq")
    write(f1, data1); write(f2, data2)
    want = ours(pkg, f1, f2, 128, 3)
    p1, p2 = str(tmp_path / "fifo1"), str(tmp_path / "fifo2")
    os.mkfifo(p1); os.mkfifo(p2)
    feeders = [threading.Thread(target=write, args=(p, d)) for p, d in ((p1, data1), (p2, data2))]
    [t.start() for t in feeders]
    got = ours(pkg, p1, p2, 128, 3)
    [t.join() for t in feeders]
    assert got == want and sum(len(b) for b in got) == 600
