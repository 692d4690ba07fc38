set -x
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/g1_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/g1_smoke.log 2>&1
for i in 1 2 3; do
  python tools/mh_time.py c2 >> $O/g1_ab.log 2>&1
  SMCB_LIB=$PWD/tools/probe/libsmcb200_wt.so python tools/mh_time.py c2 >> $O/g1_ab.log 2>&1
done
SMCB_LIB=$PWD/tools/probe/libsmcb200_wt.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2 or variant or instantiation or linear or full_size" > $O/g1_wt_pytest.log 2>&1
( time python bench.py ) > $O/g1_bench.json 2> $O/g1_bench.err
