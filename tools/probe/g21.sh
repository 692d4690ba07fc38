O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/g21_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/g21_smoke.log 2>&1
( time python bench.py ) > $O/g21_n1.json 2> $O/g21_n1.err
