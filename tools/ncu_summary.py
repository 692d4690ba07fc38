"""Summarises an .ncu-rep (read on the CPU box with `ncu -i`): key metrics + top stall reasons
+ hottest SASS lines.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_hot]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_bytes.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:3]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print("%-70s %s %s" % (k, r[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
sc = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
num = lambda x: int(x) if x.strip().isdigit() else 0
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
tot = sum(num(r[isamp]) for r in body)
agg = {}
for r in body:
    for i in sc: agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
print("stall samples:", ", ".join("%s %.1f%%" % (k, 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
for r in sorted(body, key=lambda r: -num(r[isamp]))[:nhot]:
    st = sorted([(num(r[i]), hdr[i]) for i in sc], reverse=True)[:2]
    print("%5.1f%% ex=%9s  %-56s %s" % (100 * num(r[isamp]) / max(tot, 1), r[iex], r[ia][:56], st))
