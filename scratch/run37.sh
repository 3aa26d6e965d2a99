timeout 300 python -m pytest tests/test_tensor_core_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -4
timeout 300 python scratch/time_tc.py 2>&1 | tail -12
