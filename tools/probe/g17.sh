O=gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -x -q > $O/g17_pytest.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/g17_n2.json 2> $O/g17_n2.err
