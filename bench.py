#!/usr/bin/env python
"""bench.py — the matching path's headline metric on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config1|config3] [--pairs P]
  python bench.py --impl reference ...      # the reference's own CPU implementation (oracle/_ref), rank 0 only

A step = one pass of the hot path (alignToDatabase + score screen + getPairedOverlaps, SLAM.h:209-214) over
one batch of synthetic read pairs. `value` is measured with the batch already packed-resident in HBM
(kslam_align_resident + kslam_pair_batch, no result copy); `e2e` goes through the reference-facing C-ABI call
with HOST buffers (kslam_align_batch + kslam_pair_batch with results copied back), host<->device copies inside
the timed region. Ranks shard read pairs (weak scaling: every rank owns a full batch, the genome index is
replicated, no collective on the data path); the timed region is bracketed by barrier + synchronize and the
MAX over ranks is taken. One JSON line on stdout from rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

METRIC = "M read-pairs/min end-to-end at 1/2/4/8 B200; SW GCUPS; k-mer join GB/s"
UNIT = "M read-pairs/min"


# stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner ...) are sent to stderr
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples that arrived inside [t_begin, t_end] (the timed region)."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.05)]
        if not rows:   # region shorter than one sampling period: fall back to the samples taken under warm-up load
            rows = [r for _, r in self.rows[-5:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(pkg, name, pairs, seed_shift=0):
    synth = pkg.synth
    t0 = time.time()
    if name == "config1":
        gb, go = synth.random_genomes(50, 3_000_000, seed=1)
        desc = f"config1-shape: {pairs} x 150bp FR pairs from 50 x 3 Mbp random genomes"
    elif name == "config2":
        gb, go = synth.tree_genomes(500, 2_000_000, seed=1)
        desc = (f"config2-shape (metagenomic): {pairs} x 150bp FR pairs vs 500 x 2 Mbp genomes in a 5/25/100/500 "
                "phylogeny (multi-genome piles)")
    elif name == "config4":
        ng = int(os.environ.get("KSLAM_CONFIG4_GENOMES", "500"))
        gb, go = synth.random_genomes(ng, 4_000_000, seed=1)
        desc = (f"config4-shape (k-mer-range partitioned DB, scaled): {pairs} x 150bp FR pairs per GPU vs {ng} x 4 Mbp genomes "
                f"({ng * 250_000 / 1e6:.0f} M genome k-mer records range-partitioned across the GPUs, NCCL all-to-all both ways)")
    else:
        raise SystemExit(f"unknown workload {name}")
    rb, ro, _ = synth.paired_reads(gb, go, pairs, seed=2 + seed_shift)
    log(f"[bench] generated {desc} in {time.time() - t0:.1f}s")
    return gb, go, rb, ro, desc


def cpu_reference_sample(pkg, gb, go, n_pairs_sample, report_cigar, threshold):
    """The reference's own alignToDatabase + screen + getPairedOverlaps (oracle/_ref, all host threads) on a
    bounded sample of the same workload. Returns (pairs/min in millions, seconds, cores, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _lib as T
    rb, ro, _ = pkg.synth.paired_reads(gb, go, n_pairs_sample, seed=99)
    P = T.default_params(report_cigar=int(report_cigar), score_threshold=threshold)
    if T.have_ref():
        T.ref().kref_set_threads(os.cpu_count() or 1)     # torchrun exports OMP_NUM_THREADS=1: ask for every core
        R = T.Ref(gb, go, rb, ro, P)
        cores = T.ref().kref_num_threads()
        t0 = time.time()
        R.align_to_database()
        R.screen_and_pair()
        dt = time.time() - t0
        R.close()
        kind = "reference"
    else:
        cores = os.cpu_count() or 1
        t0 = time.time()
        T.ko_pipeline(gb, go, rb, ro, P, threads=cores)
        dt = time.time() - t0
        kind = "port"
    return n_pairs_sample / dt * 60 / 1e6, dt, cores, kind


def sw_pairs_chunked(pkg, n, read_len, window_len, seed):
    parts = [pkg.synth.sw_pairs(min(500_000, n - i), read_len, window_len, seed=seed + i // 500_000) for i in range(0, n, 500_000)]
    q = np.concatenate([p[0] for p in parts]); r = np.concatenate([p[2] for p in parts])
    return (q, np.arange(n + 1, dtype=np.uint64) * np.uint64(read_len), r, np.arange(n + 1, dtype=np.uint64) * np.uint64(window_len))


def run_config3(args, pkg):
    """Config 3, the Smith-Waterman microbenchmark: (read 150, window) pairs through the batched Aligner::Align entry
    point (kslam_ssw_batch), both shapes SURVEY.md §8d names (the live reference window of 150 and the 300-wide one),
    with and without CIGAR, next to ssw.c (oracle/_ref) on the host cores. GCUPS = read x window cells / all SW time."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the matching path has no CPU fallback")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _lib as T
    n = args.pairs or 2_000_000
    shapes = []
    for window in (150, 300):
        q, qo, r, ro = sw_pairs_chunked(pkg, n, 150, window, seed=300 + window)
        row = {"read_len": 150, "window_len": window, "pairs": n}
        for cigar in (False, True):
            with pkg.Aligner(report_cigar=cigar) as al:
                al.ssw_upload(q, qo, r, ro)
                for _ in range(args.warmup):
                    al.ssw_resident()
                ms = []
                for _ in range(args.steps):
                    al.ssw_resident(); ms.append(al.timings()["ms_total"])
                tm = al.timings()
                if not cigar:
                    int_peak = al.measure_int_peak()
            t = float(np.mean(ms)) / 1e3
            key = "cigar" if cigar else "score_only"
            row[key] = {"gcups": 150.0 * window * n / t / 1e9, "ms": t * 1e3, "M_pairs_per_min": n / t * 60 / 1e6,
                        "tiers": {k: tm[k] for k in ("n_sw_tier8", "n_sw_tier16", "n_sw_tier32", "n_sw_tier48", "n_sw_tier64", "n_sw_sweep32", "n_sw_fast", "n_sw_slow")},
                        "int_pipe_frac": 3.0 * tm["sw_cells_computed"] / ((tm["ms_sw_forward"] + tm["ms_sw_reverse"]) / 1e3) / int_peak}
        if not args.no_cpu_baseline and T.have_ref():
            cs = args.cpu_sample or 200_000
            P = T.default_params(report_cigar=1)
            t0 = time.time(); T.ref_ssw_batch(q[:cs * 150], qo[:cs + 1], r[:cs * window], ro[:cs + 1], P, cigar_cap=32, threads=os.cpu_count() or 1)
            dt = time.time() - t0
            row["ssw_c_reference"] = {"gcups": 150.0 * window * cs / dt / 1e9, "cores": os.cpu_count() or 1, "sample_pairs": cs, "with_cigar": True}
        shapes.append(row)
        log(f"[bench/config3] {row}")
    live = shapes[0]["score_only"]
    emit({"metric": METRIC, "value": live["M_pairs_per_min"], "unit": "M read/window pairs per min (SW microbench, 150 x 150, score + coordinates)",
          "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": live["ms"], "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
          "config": {"workload": f"config3 (SW microbench): {n} read/window pairs per shape, scoring 2/3/5/2; mix 70 % 1 % subs, 20 % one 1-5 bp indel, "
                                 "5 % unrelated, 5 % with N runs", "l2": "inputs larger than L2"},
          "sw_gcups": live["gcups"], "shapes": shapes, "int_peak_thread_ops_per_s": int_peak})


def run_fastq_to_sam(args, pkg):
    """--workload sam: the --just-align --sam-file run end to end, FASTQ text + FASTA in, SAM text out (config-1 data),
    through slam.align_to_sam (reader | GPU | host stages + SAM writer pipeline). Wall clock of the whole call, including
    FASTA parsing, index build and pipeline fill; per-stage busy times alongside."""
    import tempfile
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the matching path has no CPU fallback")
    from kslam_b200 import slam
    pairs = args.pairs or 4_000_000
    at_once = 1_000_000
    gb, go = pkg.synth.random_genomes(50, 3_000_000, seed=1)
    rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
    d = tempfile.mkdtemp(prefix="kslam_bench_")
    fa = os.path.join(d, "db.fa")
    with open(fa, "wb") as f:
        for i in range(len(go) - 1):
            f.write(b">g%02d synthetic\n" % i + gb[int(go[i]):int(go[i + 1])].tobytes() + b"\n")
    paths = []
    for k in range(2):
        rows = rb.reshape(-1, 150)[k * pairs:(k + 1) * pairs]
        p = os.path.join(d, f"R{k + 1}.fq"); paths.append(p)
        with open(p, "wb") as f:
            for lo in range(0, pairs, 100_000):
                blk = rows[lo:lo + 100_000]
                f.write(b"".join(b"@r%d/%d\n" % (lo + i, k + 1) + blk[i].tobytes() + b"\n+\n" + b"I" * 150 + b"\n" for i in range(len(blk))))
    in_bytes = sum(os.path.getsize(p) for p in paths)
    runs = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        st = slam.align_to_sam(pkg, [fa], paths[0], paths[1], os.path.join(d, "out.sam"), reads_at_once=at_once)
        dt = time.perf_counter() - t0
        log(f"[bench/sam] run {i}: {pairs} pairs in {dt:.2f}s, stage busy {st['seconds']}")
        if i >= args.warmup:
            runs.append((dt, st))
    dt = float(np.mean([r[0] for r in runs])); st = runs[-1][1]
    for p in paths + [fa, os.path.join(d, "out.sam")]:
        os.unlink(p)
    os.rmdir(d)
    emit({"metric": METRIC, "value": pairs / dt * 60 / 1e6, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16x2 (SW) / u64 (k-mers) / f64 (MAPQ)",
          "data": "synthetic",
          "config": {"workload": f"FASTQ -> SAM whole process: {pairs} x 150bp FR pairs (FASTQ text, {in_bytes / 1e9:.2f} GB) vs 50 x 3 Mbp genomes (FASTA), "
                                 f"--num-reads-at-once {at_once}, pseudo-assembly on, 10 alignments per read, CIGAR/MD/NM/MAPQ",
                     "includes": "FASTA parse, index build, FASTQ ingest, GPU matching path, host screens + pseudo-assembly + SAM text, file write"},
          "stage_busy_s": st["seconds"], "sam_bytes": st["sam_bytes"], "batches": st["batches"]})


def run_cli(args, pkg, meta):
    """--workload cli / cli-meta: the `SLAM` executable (k-slam_b200/csrc/slam_main.cpp, C++ host over the C ABI) as a user
    runs it, wall clock of the whole process: database load, index build, FASTQ ingest, GPU matching path, host stages, output.
    cli      = config-1 data, FASTA database (built with SLAM --parse-fasta), --just-align --sam-file   (configs 1 / 5)
    cli-meta = config-2 data scaled down (strains in a phylogeny as GenBank flat files with genes + names.dmp / nodes.dmp,
               built with SLAM --parse-genbank / --parse-taxonomy), SAM + LCA XML + _PerRead + _abbreviated   (config 2)."""
    import shutil
    import tempfile
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the matching path has no CPU fallback")
    exe = os.environ.get("KSLAM_BENCH_EXE") or os.path.join(ROOT, "k-slam_b200", "SLAM")   # (the override times another build of the CLI)
    pairs = args.pairs or (2_000_000 if meta else 4_000_000)
    at_once = 1_000_000
    d = tempfile.mkdtemp(prefix="kslam_cli_")
    db = os.path.join(d, "db"); os.mkdir(db)
    t_build = time.perf_counter()
    if meta:
        n_strains, length = 100, 1_000_000
        gb, go = pkg.synth.tree_genomes(n_strains, length, seed=1)
        nodes, strain_tax = pkg.synth.tree_taxonomy(n_strains)
        pkg.synth.write_taxonomy_dumps(nodes, os.path.join(d, "names.dmp"), os.path.join(d, "nodes.dmp"))
        gbff = os.path.join(d, "all.gbff")
        with open(gbff, "wb") as f:
            for i in range(n_strains):
                f.write(pkg.synth.genbank_text(gb[int(go[i]):int(go[i + 1])].tobytes(), f"NC_{i:06d}", 7_000_000 + i, int(strain_tax[i]),
                                               f"Synthetic strain {i}", i, i % (n_strains // 5), seed=i % (n_strains // 5)))
        subprocess.run([exe, "--parse-genbank", "--output-file", os.path.join(db, "database"), gbff], check=True, cwd=d)
        subprocess.run([exe, "--parse-taxonomy", "--output-file", os.path.join(db, "taxDB"), os.path.join(d, "names.dmp"), os.path.join(d, "nodes.dmp")],
                       check=True, cwd=d)
        what = f"{n_strains} strains x {length} bp in a phylogeny (GenBank flat files, ~1 gene / kb)"
    else:
        gb, go = pkg.synth.random_genomes(50, 3_000_000, seed=1)
        fa = os.path.join(d, "db.fa")
        with open(fa, "wb") as f:
            for i in range(len(go) - 1):
                f.write(b">g%02d synthetic\n" % i + gb[int(go[i]):int(go[i + 1])].tobytes() + b"\n")
        subprocess.run([exe, "--parse-fasta", "--output-file", os.path.join(db, "database"), fa], check=True, cwd=d)
        what = "50 x 3 Mbp genomes (FASTA)"
    t_build = time.perf_counter() - t_build
    rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
    paths = []
    for k in range(2):
        rows = rb.reshape(-1, 150)[k * pairs:(k + 1) * pairs]
        p = os.path.join(d, f"R{k + 1}.fq"); paths.append(p)
        with open(p, "wb") as f:
            for lo in range(0, pairs, 100_000):
                blk = rows[lo:lo + 100_000]
                f.write(b"".join(b"@r%d/%d\n" % (lo + i, k + 1) + blk[i].tobytes() + b"\n+\n" + b"I" * 150 + b"\n" for i in range(len(blk))))
    in_bytes = sum(os.path.getsize(p) for p in paths)
    cmd = [exe, "--db", db, "--sam-file", os.path.join(d, "out.sam"), "--num-reads-at-once", str(at_once)]
    cmd += ["--output-file", os.path.join(d, "out.xml")] if meta else ["--just-align"]
    cmd += paths
    runs = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = subprocess.run(cmd, cwd=d, capture_output=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            raise SystemExit(f"SLAM failed ({r.returncode}): {r.stderr.decode()[-500:]}")
        stages = open(os.path.join(d, "log.txt")).read().strip().splitlines()
        log(f"[bench/cli] run {i}: {pairs} pairs in {dt:.2f}s; last log line: {stages[-1] if stages else ''}")
        if i == args.warmup + args.steps - 1:
            log("[bench/cli] log.txt of the last run (first / last lines):\n  " + "\n  ".join(stages[:8] + ["..."] + stages[-6:]))
        if i >= args.warmup:
            runs.append(dt)
    dt = float(np.mean(runs))
    out_bytes = {n: os.path.getsize(os.path.join(d, n)) for n in os.listdir(d) if n.startswith("out.")}
    # the reference's OWN executable (oracle/_ref/SLAM_ref, its unmodified main.cpp) on the same files and database directory,
    # bounded by --num-reads; our executable on the same prefix must write the same files
    cpu_baseline = None
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "SLAM_ref")
    try:
        if not args.no_cpu_baseline and subprocess.run([ref_exe, "--version"], capture_output=True, timeout=60).stdout == b"1.0\n":
            sample = args.cpu_sample or (50_000 if meta else 200_000)
            cores = os.cpu_count() or 1
            both = {}
            for tag, exe in (("ref", ref_exe), ("ours", exe)):
                wd = os.path.join(d, "cmp_" + tag); os.mkdir(wd)
                cmd2 = [exe, "--db", db, "--sam-file", "o.sam", "--num-reads", str(sample), "--num-reads-at-once", str(at_once)]
                cmd2 += ["--output-file", "o.xml"] if meta else ["--just-align"]
                t0 = time.perf_counter()
                r = subprocess.run(cmd2 + paths, cwd=wd, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS=str(cores)))
                both[tag] = (time.perf_counter() - t0, r.returncode, wd)
            same = None
            if both["ref"][1] == 0 and both["ours"][1] == 0:
                def body(path):
                    t = open(path, "rb").read()
                    return t[t.index(b"@PG"):].split(b"\n", 1)[1] if path.endswith(".sam") else t
                names = ["o.sam"] + (["o.xml_PerRead"] if meta else [])   # (the XML's gene representatives depend on the reference's thread count)
                same = all(body(os.path.join(both["ref"][2], n)) == body(os.path.join(both["ours"][2], n)) for n in names)
            cpu_baseline = {"value": sample / both["ref"][0] * 60 / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
                            "sample": f"the reference's own executable (unmodified main.cpp) on the first {sample} pairs of the same files, "
                                      f"whole process, {both['ref'][0]:.1f}s", "same_output_as_ours": same,
                            "ours_on_the_same_sample_s": both["ours"][0]}
            log(f"[bench/cli] reference executable: {sample} pairs in {both['ref'][0]:.1f}s; ours {both['ours'][0]:.2f}s; same output: {same}")
    except Exception as e:   # noqa: BLE001
        log(f"[bench/cli] reference executable leg skipped: {e}")
    shutil.rmtree(d, ignore_errors=True)
    emit({"metric": METRIC, "value": pairs / dt * 60 / 1e6, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16x2 (SW) / u64 (k-mers) / f64 (MAPQ)",
          "data": "synthetic",
          "config": {"workload": f"SLAM executable, whole process: {pairs} x 150bp FR pairs (FASTQ text, {in_bytes / 1e9:.2f} GB) vs {what}, "
                                 f"--num-reads-at-once {at_once}, pseudo-assembly on, 10 alignments per read, "
                                 + ("SAM + LCA XML + _PerRead + _abbreviated" if meta else "--just-align SAM"),
                     "includes": "process start, CUDA context, DIR/database parse, index build, FASTQ ingest, GPU matching path, host stages, output files",
                     "command": " ".join(os.path.basename(c) if os.sep in c else c for c in cmd)},
          "output_bytes": out_bytes, "database_build_s": t_build, "cpu_baseline": cpu_baseline})


def run_reference_arm(args, pkg):
    """--impl reference: the reference's CPU path on this box's host cores, same config/metric/unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = args.pairs or DEFAULT_PAIRS[args.workload]
    sample = args.ref_sample or CPU_SAMPLE[args.workload]
    gb, go, _, _, desc = make_workload(pkg, args.workload, 1000)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, cores, kind = cpu_reference_sample(pkg, gb, go, sample, args.workload == "config1", 0)
        log(f"[bench/reference] step {i}: {sample} pairs in {dt:.2f}s -> {v:.3f} M pairs/min on {cores} threads")
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([x[0] for x in vals])); dt = float(np.mean([x[1] for x in vals]))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int16/int32/u64", "data": "synthetic",
           "config": {"workload": desc.replace("1000 x", f"{pairs} x"), "batch_pairs_per_gpu": pairs},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": f"{sample} pairs of the workload per step (same genomes), alignToDatabase+screen+getPairedOverlaps"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


DEFAULT_PAIRS = {"config1": 1_000_000, "config2": 10_000_000, "config4": 2_000_000}
# bounded CPU samples (pairs): sized for roughly 10-20 s of host work per step on a 16-core box
CPU_SAMPLE = {"config1": 400_000, "config2": 150_000, "config4": 100_000}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kslam", choices=["kslam", "reference"])
    ap.add_argument("--workload", default=os.environ.get("KSLAM_BENCH_WORKLOAD", "config2"), choices=["config1", "config2", "config3", "config4", "sam", "cli", "cli-meta"])
    ap.add_argument("--pairs", type=int, default=0, help="read pairs per batch per GPU (default: the config's)")
    ap.add_argument("--ref-sample", type=int, default=0, help="pairs per step for the CPU reference arm (0 = per workload)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs for the cpu_baseline leg (rank 0, N=1; 0 = per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    pkg = ge.load_pkg()
    if args.workload == "config3":
        return run_config3(args, pkg)
    if args.workload == "sam":
        return run_fastq_to_sam(args, pkg)
    if args.workload in ("cli", "cli-meta"):
        return run_cli(args, pkg, args.workload == "cli-meta")
    if args.impl == "reference":
        return run_reference_arm(args, pkg)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the matching path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pairs = args.pairs or DEFAULT_PAIRS[args.workload]
    gb, go, rb, ro, desc = make_workload(pkg, args.workload, pairs, seed_shift=rank)
    # pinned host staging for the e2e leg (the C ABI takes plain host pointers)
    rb_pin = torch.from_numpy(rb).pin_memory()
    rb_host = rb_pin.numpy()
    # config 1 is the --just-align / --sam-file run: reportCigar is on there (SLAM.h:169); configs 2 and 4 write XML only
    want_cigar = args.workload == "config1"
    al = pkg.Aligner(report_cigar=want_cigar, device=local)
    al.set_debug_taps(False)
    partitioned = args.workload == "config4"
    t0 = time.time()
    if partitioned:
        from kslam_b200 import dist as kd
        al.load_genomes_part(gb, go, rank, world)
        engine = kd.CudaEngine(al, local)
        exch = kd.TorchExchange(device=torch.device("cuda", local)) if world > 1 else kd.LoopbackGroup(1).exchange(0)
        n_reads = len(ro) - 1
        xstats = {}

        def step_resident():
            _, st = kd.align_partitioned(engine, exch, n_reads, fetch=False)
            xstats.update(st)
            al.pair_batch(fetch=False)

        def step_e2e():
            al.upload_reads(rb_host, ro)
            kd.align_partitioned(engine, exch, n_reads, fetch=False)
            return None, al.pair_batch(fetch=True, copy=False)
    else:
        al.load_genomes(gb, go)

        def step_resident():
            al.align_resident(fetch=False); al.pair_batch(fetch=False)

        def step_e2e():
            return None, al.align_pair_batch(rb_host, ro, copy=False)
    t_load = time.time() - t0
    log(f"[bench r{rank}] genome index built in {t_load:.2f}s")

    # ---- value: inputs resident in HBM -------------------------------------------------------
    al.upload_reads(rb_host, ro)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    launches0 = al.timings()["kernel_launches"]
    barrier()
    tw0 = time.perf_counter()
    tw_first = tw0
    stage = {}
    for _ in range(args.steps):
        step_resident()
        tm = al.timings()
        for k, v in tm.items():
            if k.startswith("ms_"):
                stage[k] = stage.get(k, 0.0) + v
    barrier()
    tw1 = time.perf_counter()
    t_res = tw1 - tw0                           # library calls are synchronous: wall == device time of the chain
    tm = al.timings()
    launches = (tm["kernel_launches"] - launches0) // max(1, args.steps)

    # ---- e2e: host buffers in, results back on the host ------------------------------------------
    # The reference-facing call with HOST buffers: kslam_align_pair_batch = the body of the batch loop (SLAM.h:209-214),
    # H2D of the reads and D2H of everything the loop keeps inside the timed region. Batches are streamed through TWO
    # contexts on the same GPU (include/kslam.h: "two per GPU to double-buffer"), each driven by its own host thread. A
    # batch is kslam_upload_reads (H2D) | kslam_align_resident + kslam_pair_batch (kernels) | kslam_fetch_pairs (D2H), and
    # a host mutex around the kernel phase hands the GPU from one context to the other, so the copies of one batch always
    # run under the kernels of the other. The partitioned workload runs one context (its NCCL collectives are issued in
    # program order).
    depth = 1 if partitioned else 2
    ctxs = [al]
    if depth == 2:
        al2 = pkg.Aligner(report_cigar=want_cigar, device=local, )
        al2.set_debug_taps(False)
        al2.load_genomes(gb, go)
        ctxs.append(al2)
    last = [None] * depth

    def run_e2e(n_batches):
        if depth == 1:
            for _ in range(n_batches):
                last[0] = step_e2e()[1]
            return
        errs = []
        gpu_turn = threading.Lock()
        trace = bool(os.environ.get("KSLAM_BENCH_TRACE"))

        def worker(k):
            try:
                torch.cuda.set_device(local)
                for _b in range(k, n_batches, depth):
                    t0 = time.perf_counter()
                    ctxs[k].upload_reads(rb_host, ro)
                    t1 = time.perf_counter()
                    with gpu_turn:
                        t2 = time.perf_counter()
                        ctxs[k].align_resident(fetch=False); ctxs[k].pair_batch(fetch=False)
                        t3 = time.perf_counter()
                    last[k] = ctxs[k].fetch_pairs(copy=False)
                    if trace:
                        log(f"[e2e ctx{k} batch{_b}] upload {t1 - t0:.3f}s wait {t2 - t1:.3f}s kernels {t3 - t2:.3f}s fetch {time.perf_counter() - t3:.3f}s")
            except Exception as e:   # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=worker, args=(k,)) for k in range(depth)]
        [t.start() for t in th]; [t.join() for t in th]
        if errs:
            raise errs[0]

    run_e2e(max(2, args.warmup - 1))
    barrier()
    tw0 = time.perf_counter()
    run_e2e(args.steps)
    barrier()
    tw3 = time.perf_counter()
    t_e2e = tw3 - tw0
    clocks = sampler.stop(tw_first, tw3) if rank == 0 else None
    pr = last[0]
    h2d = int(rb_host.nbytes + 3 * ro.nbytes)
    d2h = int(pr.sorted_overlaps.nbytes + pr.cigar_pool.nbytes + pr.pairs.nbytes)
    if depth == 2:
        al2.close()

    t_max = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    t_res, t_e2e = float(t_max[0]), float(t_max[1])
    total_pairs = pairs * world * args.steps
    value = total_pairs / t_res * 60 / 1e6
    e2e = total_pairs / t_e2e * 60 / 1e6

    int_peak = al.measure_int_peak() if rank == 0 else 0.0
    if rank == 0:
        peak, peak_src = measured_peaks()
        ms = {k: v / args.steps for k, v in stage.items()}
        n_rk, n_raw = tm["n_sorted_kmers"], tm["n_raw_seeds"]     # records actually sorted (after the prefilter)
        passes_kmer = max(1, al.kmer_sort_bits() // 8)     # 8-bit digits over the leading bits the join needs
        sort_s = ms["ms_sort"] / 1e3
        # dominant HBM-bound kernel: k_rs_onesweep of the read k-mer sort. algorithmic bytes per launch = 32 B x records
        # (one read + one write of every 16 B record per pass); duration = sort time / passes (histogram charged too)
        per_launch_s = sort_s / passes_kmer if sort_s > 0 else float("nan")
        achieved = 32.0 * n_rk / per_launch_s / 1e9 if sort_s > 0 else 0.0
        join_gbs = (32.0 * n_rk + 16.0 * n_rk + 16.0 * n_raw) / ((ms["ms_sort"] + ms["ms_join"]) / 1e3) / 1e9
        sw_s = (ms["ms_sw_forward"] + ms["ms_sw_reverse"] + ms["ms_sw_traceback"] + ms["ms_sw_slow"] + ms["ms_sw_prepare"]) / 1e3
        gcups = tm["sw_cells_forward"] / sw_s / 1e9 if sw_s > 0 else 0.0
        gcups_kernel = (tm["sw_cells_forward"] + tm["sw_cells_reverse"]) / ((ms["ms_sw_forward"] + ms["ms_sw_reverse"]) / 1e3) / 1e9 \
            if ms["ms_sw_forward"] + ms["ms_sw_reverse"] > 0 else 0.0
        hbm_roof = {"bound": "hbm", "kernel": "k_rs_onesweep (read k-mer LSD pass)", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak,
                    # ncu --set full of this kernel on 64 M records (profiles/r1_ncu_summaries.txt, prof_sort_r1b): dram read
                    # 1.037 GB + write 1.015 GB per launch = 32.07 B per record, i.e. no re-reads beyond the algorithmic bytes
                    "traffic": 32.07 * n_rk, "traffic_source": "ncu dram__bytes_read+write per record (64 M-record capture) x records per launch",
                    "peak_source": peak_src,
                    "share_of_step": ms["ms_sort"] / (t_res / args.steps * 1e3),
                    "note": f"algorithmic 32 B/record/pass; duration = (sort stage incl. histogram)/{passes_kmer} passes, CUDA events on the ctx stream"}
        # SW sweeps (k_sw_band / k_sw_fast): integer-pipe bound. 6 ALU thread-ops (1 PRMT + 5 s16x2 DPX) per two cells
        # (the seventh op of the recurrence, H - gapOpen, is an IMAD on the FMA pipe thanks to the biased cells);
        # numerator = cells the sweep kernels actually computed (band cells, not matrix cells), denominator = the
        # issue rate of VIADDMNMX.S16x2 measured on this GPU right now (kslam_measure_int_peak).
        sweep_s = (ms["ms_sw_forward"] + ms["ms_sw_reverse"]) / 1e3
        int_ops = 3.0 * tm["sw_cells_computed"]
        int_ach = int_ops / sweep_s / 1e12 if sweep_s > 0 else 0.0
        int_roof = {"bound": "int-pipe", "kernel": "k_sw_band / k_sw_fast (SW forward + reverse sweeps)", "achieved": int_ach,
                    "peak": int_peak / 1e12, "unit": "T int16x2 thread-ops/s", "frac": int_ach / (int_peak / 1e12) if int_peak else None,
                    "traffic": None, "peak_source": "measured live: dependency-free VIADDMNMX.S16x2 issue-rate microbenchmark (kslam_measure_int_peak)",
                    "share_of_step": sweep_s * 1e3 / (t_res / args.steps * 1e3),
                    "note": "3 ALU thread-ops per computed cell (1 PRMT + 5 DPX per s16x2 cell pair; H - gapOpen is an IMAD on the FMA pipe); cells computed = band cells "
                            "(32 or 64 per row) or the full matrix for fallback alignments; CUDA events on the ctx stream"}
        dominant = int_roof if sweep_s * 1e3 >= ms["ms_sort"] else hbm_roof
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "int16x2 (SW) / u64 (k-mers)", "data": "synthetic",
               "config": {"workload": desc, "batch_pairs_per_gpu": pairs, "sharding": (f"genome k-mer list range-partitioned over {world} ranks + read pairs per rank; all-to-all of k-mer records "
                                       f"and of raw matches (NCCL)" if partitioned else f"read pairs, {world} ranks, no collective"),
                          "l2": "inputs larger than L2 (3.8 GB+ of k-mer records per step)", "report_cigar": want_cigar},
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "call": "kslam_upload_reads + kslam_align_resident + kslam_pair_batch + kslam_fetch_pairs (host buffers in, pair-sorted overlaps + pairs back on the host)",
                       "contexts_per_gpu": depth},
               "gpu_launches": int(launches),
               "clocks": clocks,
               "roofline": dominant, "roofline_hbm": hbm_roof, "roofline_int": int_roof,
               "kmer_join_gbs": join_gbs, "sw_gcups": gcups, "sw_kernel_gcups_fwd_plus_rev": gcups_kernel,
               "stage_ms": ms,
               "counts": {k: tm[k] for k in ("n_read_kmers", "n_sorted_kmers", "n_genome_kmers", "n_raw_seeds", "n_seeds", "n_pairs",
                                             "n_sw_band", "n_sw_band64", "n_sw_fast", "n_sw_slow", "n_sw_band_rev",
                                             "n_sw_tier8", "n_sw_tier16", "n_sw_tier32", "n_sw_tier48", "n_sw_tier64", "n_sw_sweep32",
                                             "sw_cells_forward", "sw_cells_reverse", "sw_cells_computed", "n_sort_passes")},
               "genome_index_build_s": t_load}
        if partitioned:
            out["exchange"] = {"rank0_records_per_step": xstats, "record_bytes": 16,
                               "partition": {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in al.partition().items() if k != "splitters"}}
        if world == 1 and not args.no_cpu_baseline:
            cs = args.cpu_sample or CPU_SAMPLE[args.workload]
            v, dt, cores, kind = cpu_reference_sample(pkg, gb, go, cs, want_cigar, 0)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                   "sample": f"{cs} pairs of the same workload in {dt:.1f}s (alignToDatabase+screen+getPairedOverlaps, genome k-mers re-extracted and re-sorted per batch as the reference does)"}
        emit(out)
    al.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
