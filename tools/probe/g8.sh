O=gpurun_out
for i in 1 2 3; do python bench.py --no-cpu --no-north-star > $O/g8_n1_$i.json 2> $O/g8_n1_$i.err; done
python bench.py > $O/g8_full.json 2> $O/g8_full.err
