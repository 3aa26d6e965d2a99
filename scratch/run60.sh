timeout 300 python -m pytest tests/test_host_step_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python scratch/e2e_var.py
timeout 300 python scratch/e2e_gpu_side.py
