SECONDS=0
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_out.log 2> gpurun_out/bench_err.log; echo "bench rc=$? elapsed=${SECONDS}s"
tail -1 gpurun_out/bench_out.log > gpurun_out/bench_r1_v3.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_v3.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'upd', d['critic_updates']['value'], d['critic_updates']['ms_per_round'], 'grouped', d['critic_updates']['grouped']['value'])
print('roofline', d['roofline'])
t=d['tensor_core_td_target']; print('tensor', t['tcgen05'], t['simt_fp32'], t['frac'], t['grouped_update_round'])
"
