O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -m gpu -x -q -k "spring or c3 or nvrtc" > $O/g14_pytest.log 2>&1
SMCB_SPRING_FORM=1 python tools/mh_time.py c3 22 >> $O/g14_time.log 2>&1
python tools/run_config.py c3 > $O/g14_c3.json 2> $O/g14_c3.err
python tools/mh_time.py c3 22 > $O/g14_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 3 -c 1 -o $O/g14_prof_mh_c3 \
    python tools/mh_time.py c3 22 > $O/g14_ncu_c3.log 2>&1
ncu -i $O/g14_prof_mh_c3.ncu-rep --page raw --csv > $O/g14_prof_mh_c3_raw.csv 2>/dev/null
ncu -i $O/g14_prof_mh_c3.ncu-rep --page source --csv 2>/dev/null | gzip > $O/g14_prof_mh_c3_src.csv.gz
rm -f $O/g14_prof_mh_c3.ncu-rep
