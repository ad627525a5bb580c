"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: time, share, launches."""
import collections, csv, sys
def main(path, last_step=False):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    rows = list(csv.DictReader(lines))
    if last_step:      # one batch = from the last read k-mer extraction up to (not including) the next genome / read packing
        starts = [i for i, r in enumerate(rows) if r['Kernel Name'].startswith('k_extract_reads_filtered')]
        rows = rows[starts[-1]:]
        stop = [i for i, r in enumerate(rows) if r['Kernel Name'].startswith(('k_pack', 'k_pairs_compact', 'k_insert_hist', 'k_int_peak'))]
        if stop:
            rows = rows[:stop[0]]
    for row in rows:
        k = row['Kernel Name'].split('(')[0].replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}{' (last batch)' if last_step else ''}: {sum(v[0] for v in agg.values())} launches, {tot/1e6:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}%  n={n:3d}  {k[:80]}")
if __name__ == '__main__':
    main(sys.argv[1], len(sys.argv) > 2 and sys.argv[2] == '--last-step')
