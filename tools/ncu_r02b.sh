#!/bin/bash
# r02 final-state ncu evidence (one GPU), after the SFU Box-Muller / early accept slots / dependent-launch changes:
# launch lists of one C2 bench pass and of the C5 north-star run, --set full captures of the three Metropolis
# kernels that changed (C2 linear, C5 wide identity, C4 MVNormal).  Each profiled command first runs plain with &&.
set -x
O=gpurun_out
python bench.py > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err
python bench.py --steps 1 --warmup 1 --no-cpu --no-north-star > $O/r02b_plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 460 -c 1400 --csv --log-file $O/r02b_launches_c2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-north-star > $O/r02b_ncu_c2.log 2>&1
python tools/c5_run.py standard 23 > $O/r02b_plain_c5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file $O/r02b_launches_c5.csv \
    python tools/c5_run.py standard 23 > $O/r02b_ncu_c5.log 2>&1
python tools/mh_only.py c2 > $O/r02b_plain_mh_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 3 -c 1 -o $O/r02b_prof_mh_c2 \
    python tools/mh_only.py c2 > $O/r02b_ncu_mh_c2.log 2>&1
python tools/mh_only.py c5 > $O/r02b_plain_mh_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 3 -c 1 -o $O/r02b_prof_mh_c5 \
    python tools/mh_only.py c5 > $O/r02b_ncu_mh_c5.log 2>&1
python tools/mh_time.py c4 22 > $O/r02b_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 6 -c 1 -o $O/r02b_prof_mh_c4 \
    python tools/mh_time.py c4 22 > $O/r02b_ncu_c4.log 2>&1
for r in mh_c2 mh_c5 mh_c4; do
  ncu -i $O/r02b_prof_$r.ncu-rep --page raw --csv > $O/r02b_prof_${r}_raw.csv 2>/dev/null
  ncu -i $O/r02b_prof_$r.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02b_prof_${r}_src.csv.gz
  rm -f $O/r02b_prof_$r.ncu-rep
done
python tools/run_config.py c3 > $O/r02b_c3_4M_1gpu.json 2> $O/r02b_c3.err
python tools/run_config.py c4 22 > $O/r02b_c4_4M_1gpu.json 2> $O/r02b_c4.err
ls -la $O/r02b_*
