O=gpurun_out
for i in 1 2; do
for sg in 0 200 400 700 1000 1500 2500; do
  echo "stagger $sg" >> $O/g5_stagger.log
  SMCB_STAGGER_NS=$sg python tools/mh_chain_time.py c2 >> $O/g5_stagger.log 2>&1
done
done
SMCB_STAGGER_NS=700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2 or dependent or full_size" > $O/g5_pytest.log 2>&1
