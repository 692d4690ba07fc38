O=gpurun_out
python tools/cov_only.py > $O/g19_cov_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wmoments_dmma -s 3 -c 1 -o $O/g19_prof_cov python tools/cov_only.py > $O/g19_ncu_cov.log 2>&1
ncu -i $O/g19_prof_cov.ncu-rep --page raw --csv > $O/g19_prof_cov_raw.csv 2>/dev/null
ncu -i $O/g19_prof_cov.ncu-rep --page source --csv 2>/dev/null | gzip > $O/g19_prof_cov_src.csv.gz
rm -f $O/g19_prof_cov.ncu-rep
