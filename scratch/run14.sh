python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scratch/env_roofline.py 2>&1 | tail -8
