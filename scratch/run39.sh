python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python scratch/time_upd.py 2>&1 | tail -8
timeout 300 python scratch/time_tc.py 2>&1 | tail -12
