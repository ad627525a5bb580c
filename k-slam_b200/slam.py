"""FASTQ + FASTA -> SAM through the library, batch by batch: the --just-align / --sam-file run of the reference
(`SLAM --db DB --just-align --sam-file out.sam R1 R2`, SLAM.h:159-268) with every stage of its batch loop taken from
libkslam.so: kslam_fastq_next (reader), kslam_align_pair_batch (alignToDatabase + screen + getPairedOverlaps on the GPU)
and kslam_sam_batch (insert-size / score screens, pseudo-assembly, SAM records). The database is given as FASTA files,
parsed the way --parse-fasta does (GenbankTools.h:224-260: locus tag = header up to the first space, bases upper-cased),
or as a --db directory (database.py). This is the Python form of the pipeline (tests, bench.py --workload sam); the C++
form with the reference's full command line, taxonomy and XML included, is k-slam_b200/SLAM (csrc/slam_main.cpp)."""
from __future__ import annotations

import time

import numpy as np


def parse_fasta(paths):
    """createIndexFromFASTA, GenbankTools.h:224-260. Returns (bases u8, offs u64, locus tags).
    A line is a header when it starts with '>' (locus tag = the text up to the first space, if there is one and it is not
    the first character); every other non-empty line is appended to the current entry; entries without bases are dropped;
    one GenbankEntry object is reused per file, so bases seen before the first header of a file form an entry too."""
    entries, tags = [], []

    def close(cur, tag):
        seq = b"".join(cur)
        if seq:
            entries.append(seq); tags.append(tag)

    for path in paths:
        with open(path, "rb") as f:
            data = f.read()
        if b"\r" in data:                       # safeGetline also ends lines at "\r\n" and lone "\r"
            data = data.replace(b"\r\n", b"\n").replace(b"\r", b"\n")
        cur, tag = [], b""
        pos, n = 0, len(data)
        while pos < n:
            if data[pos:pos + 1] == b">":
                close(cur, tag)
                cur, tag = [], b""
                end = data.find(b"\n", pos)
                end = n if end < 0 else end
                line = data[pos:end]
                sp = line.find(b" ")
                if sp not in (-1, 0):
                    tag = line[1:sp]
                pos = end + 1
            else:                                  # a block of sequence lines up to the next header line
                nxt = data.find(b"\n>", pos)
                end = n if nxt < 0 else nxt + 1
                cur.append(data[pos:end].translate(None, b"\n"))
                pos = end
        close(cur, tag)
    entries = [e.upper() for e in entries]            # inPlaceConvertToUpperCase
    offs = np.zeros(len(entries) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(e) for e in entries])
    bases = np.frombuffer(b"".join(entries), dtype=np.uint8).copy() if entries else np.zeros(0, np.uint8)
    return bases, offs, tags


def load_database(fasta_paths=None, db_dir=None):
    """The genomes either from FASTA files (parsed as --parse-fasta does) or from a SLAM database directory (`--db DIR`:
    DIR/database, database.py). -> (bases, offs, locus tags, taxonomy ids or None)"""
    if db_dir:
        import os
        from . import database
        gb, go, tags, tax = database.flatten(database.read_database(os.path.join(db_dir, "database")))
        return gb, go, tags, tax
    gb, go, tags = parse_fasta(fasta_paths)
    return gb, go, tags, None


def align_to_sam(pkg, fasta_paths, r1, r2, sam_path, reads_at_once=10_000_000, num_alignments=10, score_fraction_threshold=0.95,
                 pseudo_assembly=True, sam_xa=False, min_alignment_score=0, command_line="", device=0, log=None, db_dir=None):
    """The batch loop as a three-stage pipeline, one host thread per stage (every heavy call is a C call that releases the
    GIL): FASTQ ingest of batch i+2 | GPU matching of batch i+1 | host stages + SAM text + file write of batch i.
    The reader fills a ring of six buffer sets (one being filled, one in each queue, one in each of the two later stages,
    one spare), so batches are handed over without copies. Returns counts and per-stage busy times."""
    import queue
    import threading
    t = {"ingest": 0.0, "gpu": 0.0, "sam": 0.0, "write": 0.0}
    gb, go, tags, tax = load_database(fasta_paths, db_dir)
    stats = {"pairs": 0, "batches": 0, "sam_bytes": 0}
    errs = []
    q_reads, q_pairs = queue.Queue(maxsize=1), queue.Queue(maxsize=1)

    def stage_ingest(rd):
        try:
            while not errs:
                t0 = time.perf_counter()
                b = rd.next(reads_at_once, copy=False)     # ring of 6 buffer sets: at most 4 older batches are still in flight
                t["ingest"] += time.perf_counter() - t0
                q_reads.put(b)
                if b is None:
                    return
        except Exception as e:   # noqa: BLE001
            errs.append(e); q_reads.put(None)

    def stage_gpu(al):
        try:
            import torch
            torch.cuda.set_device(device)
        except Exception:   # noqa: BLE001
            pass
        try:
            while True:
                b = q_reads.get()
                if b is None or errs:
                    q_pairs.put(None)
                    return
                t0 = time.perf_counter()
                p = al.align_pair_batch(b.bases, b.offs, copy=True)
                t["gpu"] += time.perf_counter() - t0
                q_pairs.put((b, p))
        except Exception as e:   # noqa: BLE001
            errs.append(e); q_pairs.put(None)

    with pkg.Aligner(report_cigar=True, score_threshold=min_alignment_score, device=device) as al, \
            pkg.FastqReader(r1, r2, ring=6) as rd, open(sam_path, "wb") as out:
        al.set_debug_taps(False)
        al.load_genomes(gb, go)
        w = pkg.SamWriter(gb, go, tags, taxonomy_ids=tax, num_alignments=num_alignments, score_fraction_threshold=score_fraction_threshold,
                          pseudo_assembly=pseudo_assembly, report_cigar=True, sam_xa=sam_xa)
        out.write(w.header(command_line))
        th = [threading.Thread(target=stage_ingest, args=(rd,)), threading.Thread(target=stage_gpu, args=(al,))]
        [x.start() for x in th]
        while True:
            item = q_pairs.get()
            if item is None:
                break
            b, p = item
            t0 = time.perf_counter()
            n_text, _ = w.batch(b.bases, b.offs, b.quals, b.qual_offs, b.ids, b.id_offs, p.sorted_overlaps, p.cigar_pool, p.pairs, out_file=out)
            t["sam"] += time.perf_counter() - t0
            stats["pairs"] += b.n_r1; stats["batches"] += 1; stats["sam_bytes"] += n_text
            if log:
                log(f"batch {stats['batches']}: {b.n_r1} pairs done")
        [x.join() for x in th]
    if errs:
        raise errs[0]
    stats["seconds"] = t
    return stats
