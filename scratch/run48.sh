python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python scratch/time_tc.py 2>&1 | tail -12
