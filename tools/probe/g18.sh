O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/g18_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/g18_smoke.log 2>&1
( time python bench.py ) > $O/g18_n1.json 2> $O/g18_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/g18_n2.json 2> $O/g18_n2.err
