"""Summarise an .ncu-rep (full set) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second']


def summarize(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        out.append(f"## {row[hdr.index('Kernel Name')][:90]}  ({path})")
        for k in KEYS:
            if k in hdr:
                out.append(f"  {k:82s} {row[hdr.index(k)]} {units[hdr.index(k)]}")
        st = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
        vals = sorted(((float(row[hdr.index(h)].replace(',', '') or 0), h) for h in st), reverse=True)[:6]
        out.append("  top stall reasons (warps stalled per issue-active cycle):")
        for v, h in vals:
            out.append(f"    {v:7.2f}  {h.split('issue_stalled_')[1].split('_per_issue')[0]}")
    return "\n".join(out)


if __name__ == '__main__':
    for p in sys.argv[1:]:
        print(summarize(p))
