set -x
O=gpurun_out
( time python -m pytest tests/test_multigpu.py -m gpu -x -q ) > $O/g7_pytest_2gpu.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/g7_n2.json 2> $O/g7_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/run_config.py c4 23 > $O/g7_c4_8M_2gpu.json 2> $O/g7_c4.err
python bench.py > $O/g7_n1.json 2> $O/g7_n1.err
