timeout 300 python scratch/time_tc.py 2>&1 | grep -v critic_grads | tail -8
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
