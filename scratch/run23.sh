python -m pytest tests/test_env_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -6
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scratch/env_roofline.py 2>&1 | tail -8
