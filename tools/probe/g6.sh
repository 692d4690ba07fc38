O=gpurun_out
python tools/cov_only.py > $O/g6_cov.log 2>&1
python tools/cov_only.py 20 >> $O/g6_cov.log 2>&1
python -m pytest tests -m gpu -x -q -k "cov or moments or c5 or wide or tma" > $O/g6_pytest.log 2>&1
python tools/c5_run.py standard 23 > $O/g6_c5.log 2>&1
