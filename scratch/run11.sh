python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scratch/time_upd.py 2>&1 | tail -14
