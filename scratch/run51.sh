SECONDS=0
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_golden_gpu.py tests/test_distributed_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scratch/time_upd.py 2>&1 | head -4
timeout 600 python bench.py --no-tensor-section > gpurun_out/bench_fuse.log 2> gpurun_out/bench_fuse_err.log; echo "bench rc=$? elapsed=${SECONDS}s"
python -c "
import json; d=json.loads(open('gpurun_out/bench_fuse.log').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'upd', d['critic_updates']['value'], d['critic_updates']['ms_per_round'], 'grouped', d['critic_updates']['grouped']['value'])
"
