timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-tensor-section > gpurun_out/bench_e2e.log 2> gpurun_out/bench_e2e_err.log; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_e2e.log').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'upd', d['critic_updates']['value'])
print(d['e2e']['api'])
"
