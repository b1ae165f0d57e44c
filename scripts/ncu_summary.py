"""Summarises ncu outputs for profiles/: launch list shares and per-kernel key metrics."""
import csv, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    cols = rows[hdr]
    ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
    agg = {}
    for r in rows[hdr + 1:]:
        if len(r) <= vi: continue
        n = r[ki].split('(')[0].replace('<unnamed>::', '')
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_ms,share_pct")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        csv.writer(sys.stdout).writerow([n, a[0], "%.3f" % a[1], "%.2f" % (100 * a[1] / tot)])

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second']

def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    names, units = rows[0], rows[1]
    kn = names.index('Kernel Name')
    idx = [(i, n) for i, n in enumerate(names) if n in WANT or any(n.endswith('.' + w) for w in WANT)]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel"] + ["%s [%s]" % (n, units[i]) for i, n in idx])
    for r in rows[2:]:
        w.writerow([r[kn].split('(')[0].replace('<unnamed>::', '')] + [r[i] for i, _ in idx])

if __name__ == '__main__':
    if sys.argv[1] == 'launches': launches(sys.argv[2])
    else: raw(sys.argv[2])
