O=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spring or c3" > $O/g13_pytest.log 2>&1
for f in 1 0; do
  SMCB_SPRING_FORM=$f python tools/mh_time.py c3 22 >> $O/g13_time.log 2>&1
  SMCB_SPRING_FORM=$f SMCB_SPRING_PPT=1 python tools/mh_time.py c3 22 >> $O/g13_time.log 2>&1
done
python tools/run_config.py c3 > $O/g13_c3.json 2> $O/g13_c3.err
SMCB_SPRING_FORM=0 python tools/run_config.py c3 > $O/g13_c3_stages.json 2> $O/g13_c3_stages.err
