O=gpurun_out
for c in c2 c1 c5 c3; do
  for i in 1 2; do
    python tools/mh_chain_time.py $c >> $O/g4_pdl.log 2>&1
    SMCB_NO_PDL=1 python tools/mh_chain_time.py $c >> $O/g4_pdl.log 2>&1
  done
done
( time python -m pytest tests -m gpu -x -q ) > $O/g4_pytest.log 2>&1
python bench.py > $O/g4_bench.json 2> $O/g4_bench.err
