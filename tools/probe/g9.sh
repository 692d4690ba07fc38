O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 5 --warmup 3 > $O/g9_n4.json 2> $O/g9_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/run_config.py c4 24 > $O/g9_c4_16M_4gpu.json 2> $O/g9_c4.err
