"""Summarise an ncu launch list (+ optional full-set report) into markdown.  Usage:
   python profiles/summarize.py <launches.csv> [<report.ncu-rep>]"""
import collections
import csv
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'])
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {t / tot * 100:.1f}% |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n")


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']


def full(path):
    out = subprocess.run(f"ncu -i {path} --page raw --csv", shell=True, capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    rows = r[2:]
    print("| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(rows))) + " | unit |\n|---|" + "---|" * (len(rows) + 1))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | " + " | ".join(x[i] for x in rows) + f" | {units[i]} |")
    tens = [h for h in hdr if 'tensor' in h]
    print("\ntensor-related metrics present:", ", ".join(tens[:12]))


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        full(sys.argv[2])
