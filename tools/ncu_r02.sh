#!/bin/bash
# r02 ncu evidence (one GPU): launch lists of the C2 bench pass and of the C5 north-star run, and --set full captures
# of the kernels that changed this round.  Each profiled command first runs plain with && (recipe in B200_PROFILING.md).
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu --no-north-star > $O/r02_plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 460 -c 1400 --csv --log-file $O/r02_launches_c2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-north-star > $O/r02_ncu_c2.log 2>&1
python tools/c5_run.py standard 23 > $O/r02_plain_c5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file $O/r02_launches_c5.csv \
    python tools/c5_run.py standard 23 > $O/r02_ncu_c5.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:"rw_coop_kernel|ess_search_coop_kernel|rs_fused_kernel|scan_lookback_kernel" \
    -s 40 -c 8 -o $O/r02_prof_c5_step python tools/c5_run.py standard 23 > $O/r02_ncu_c5_step.log 2>&1
python tools/mh_time.py c4 22 > $O/r02_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 6 -c 1 -o $O/r02_prof_mh_c4 \
    python tools/mh_time.py c4 22 > $O/r02_ncu_c4.log 2>&1
python tools/mh_time.py c3 20 > $O/r02_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 3 -c 1 -o $O/r02_prof_mh_c3 \
    python tools/mh_time.py c3 20 > $O/r02_ncu_c3.log 2>&1
python tools/mh_only.py c2 > $O/r02_plain_mh_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 3 -c 1 -o $O/r02_prof_mh_c2 \
    python tools/mh_only.py c2 > $O/r02_ncu_mh_c2.log 2>&1
# reports are too large to travel back (64 MiB limit): export the raw metrics and the per-instruction source page here
for r in c5_step mh_c4 mh_c3 mh_c2; do
  ncu -i $O/r02_prof_$r.ncu-rep --page raw --csv > $O/r02_prof_${r}_raw.csv 2>/dev/null
  ncu -i $O/r02_prof_$r.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_prof_${r}_src.csv.gz
  rm -f $O/r02_prof_$r.ncu-rep
done
ls -la $O/r02_*
