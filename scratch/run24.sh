SECONDS=0
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_out.log 2> gpurun_out/bench_err.log; echo "bench rc=$? elapsed=${SECONDS}s"
tail -3 gpurun_out/bench_err.log
tail -1 gpurun_out/bench_out.log > gpurun_out/bench_r1_final.json
python bench.py --impl reference --steps 300 --warmup 20 > gpurun_out/bench_ref_final.json 2>gpurun_out/bench_ref_err.log; echo "ref rc=$? elapsed=${SECONDS}s"
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_final.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('kernel','achieved','frac','avg_launch_us','fp32_fma_tflops')})
r=d['roofline_env_step_kernel']; print('env step', r['achieved'], r['frac'], r['avg_launch_us'], 'big', r['at_1M_envs'] and (r['at_1M_envs']['achieved'], r['at_1M_envs']['frac']))
print('upd', d['critic_updates']['value'], d['critic_updates']['grouped'])
print('e2e', d['e2e']['value'], d['e2e']['per_call_api']['value'], d['critic_updates']['e2e']['value'])
print('tensor', d.get('tensor_core_td_target'))
print('cpu', d['cpu_baseline'])
print('clocks', d['clocks'])
print('ref', open('gpurun_out/bench_ref_final.json').read()[:400])
"
