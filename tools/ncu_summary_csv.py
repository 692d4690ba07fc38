"""Summarises exported ncu pages (the .ncu-rep files are too large to travel back from the GPU box):
  <prefix>_raw.csv     = ncu -i x.ncu-rep --page raw --csv
  <prefix>_src.csv.gz  = ncu -i x.ncu-rep --page source --csv | gzip
Usage: python tools/ncu_summary_csv.py gpurun_out/r02_prof_mh_c4 [n_hot]"""
import csv, gzip, io, sys
pre = sys.argv[1]; nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rows = list(csv.reader(open(pre + "_raw.csv")))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg.per_second']
names = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[hdr.index('Kernel Name')]
    names.append(name)
    print("### %s" % name[:110])
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print("%-66s %s %s" % (k, r[i], units[i]))
    print()
src = gzip.open(pre + "_src.csv.gz", "rt").read()
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
num = lambda x: int(x) if x.strip().isdigit() else 0
for n, hi in enumerate(his):
    end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    hdr = rows[hi]
    ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    sc = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    tot = sum(num(r[isamp]) for r in body)
    agg = {}
    for r in body:
        for i in sc: agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
    print("--- launch %d (%s): stall samples: %s" % (n, names[n][:60] if n < len(names) else "?",
          ", ".join("%s %.1f%%" % (k, 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6])))
    for r in sorted(body, key=lambda r: -num(r[isamp]))[:nhot]:
        st = sorted([(num(r[i]), hdr[i]) for i in sc], reverse=True)[:2]
        print("%5.1f%% ex=%9s  %-56s %s" % (100 * num(r[isamp]) / max(tot, 1), r[iex], r[ia][:56], st))
