timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_golden_gpu.py tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -3
MDP_LIB_NAME=libmaddpg_b200_prof.so timeout 120 python scratch/prof_episode.py
timeout 600 python bench.py --no-tensor-section > gpurun_out/bench_ep.log 2> gpurun_out/bench_ep_err.log; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_ep.log').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'upd', d['critic_updates']['value'])
"
