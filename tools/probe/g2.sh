set -x
O=gpurun_out
( time python -m pytest tests/test_multigpu.py -m gpu -x -q ) > $O/g2_pytest.log 2>&1
for t in a b; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/g2_n2$t.json 2> $O/g2_n2$t.err
done
python bench.py > $O/g2_n1.json 2> $O/g2_n1.err
./tools/probe/dfma_loop_probe > $O/g2_probe.log 2>&1
