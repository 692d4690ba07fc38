#!/bin/bash
# usage: tools/ncu_one.sh <tag> <kernel regex> <skip> <cmd...>   -> gpurun_out/<tag>_raw.csv, <tag>_src.csv.gz
tag=$1; shift; rx=$1; shift; skip=$1; shift
O=gpurun_out
"$@" > $O/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $O/$tag "$@" > $O/${tag}_ncu.log 2>&1
ncu -i $O/$tag.ncu-rep --page raw --csv > $O/${tag}_raw.csv 2>/dev/null
ncu -i $O/$tag.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${tag}_src.csv.gz
rm -f $O/$tag.ncu-rep
