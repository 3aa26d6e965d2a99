SECONDS=0
timeout 120 ./scratch/bin/umma_probe > gpurun_out/umma_probe.log 2>&1; echo "probe rc=$?"
grep -v "^[0-9,]*$" gpurun_out/umma_probe.log | tail -45
python -m pytest tests/test_host_step_gpu.py -x -q 2>&1 | tail -15
python bench.py --no-cpu-baseline --steps 500 --update-rounds 30 > gpurun_out/bench_out.log 2> gpurun_out/bench_err.log; echo "rc=$? elapsed=${SECONDS}s"
tail -3 gpurun_out/bench_err.log
tail -1 gpurun_out/bench_out.log > gpurun_out/bench_r1_hoststep.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_hoststep.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'])
print('e2e', json.dumps(d['e2e'], indent=1))
"
