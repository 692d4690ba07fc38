O=gpurun_out
for i in 1 2; do python bench.py --no-cpu > $O/g22_n1_$i.json 2> $O/g22_n1_$i.err; done
