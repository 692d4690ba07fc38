O=gpurun_out
for n in 909312 1048576 1363968 1818624 2097152; do
  SMCB_N=$n python tools/mh_time.py c2 >> $O/g3_bal.log 2>&1
done
