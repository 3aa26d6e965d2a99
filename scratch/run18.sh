timeout 300 python scratch/time_tc.py 2>&1 | tail -20
