timeout 200 ./scratch/bin/umma_probe > gpurun_out/umma_probe.log 2>&1; echo "probe rc=$?"
grep -E "a_mn=2|PROBE" gpurun_out/umma_probe.log | tail -16
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
