O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/g11_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/g11_smoke.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/g11_ref.json 2> $O/g11_ref.err
