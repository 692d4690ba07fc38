O=gpurun_out
( time python bench.py ) > $O/g12_n1.json 2> $O/g12_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/g12_n2.json 2> $O/g12_n2.err
