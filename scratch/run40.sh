SECONDS=0
python bench.py > gpurun_out/bench_out.log 2> gpurun_out/bench_err.log; echo "bench rc=$? elapsed=${SECONDS}s"
tail -3 gpurun_out/bench_err.log
tail -1 gpurun_out/bench_out.log > gpurun_out/bench_r1_final.json
python bench.py --impl reference --steps 300 --warmup 20 > gpurun_out/bench_ref_final.json 2>gpurun_out/bench_ref_err.log; echo "ref rc=$? elapsed=${SECONDS}s"
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_final.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
r=d['roofline_env_step_kernel']; print('env step', r['achieved'], r['frac'], r['avg_launch_us'], 'big', r['at_1M_envs'] and (r['at_1M_envs']['achieved'], r['at_1M_envs']['frac']))
print('upd', d['critic_updates']['value'], d['critic_updates']['ms_per_round'], d['critic_updates']['grouped'])
print('e2e', d['e2e']['value'], d['e2e']['per_call_api']['value'], d['critic_updates']['e2e']['value'])
print('tensor', json.dumps(d.get('tensor_core_td_target'), indent=1))
print('ref', json.load(open('gpurun_out/bench_ref_final.json'))['value'], json.load(open('gpurun_out/bench_ref_final.json'))['critic_updates'])
"
python tools/profile_tc.py simple_spread 24 2048 1024 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_critic_grads_tc|k_actor_grads_tc|k_td_target_tc" -c 3 -o gpurun_out/prof_tc5g_v4 python tools/profile_tc.py simple_spread 24 2048 1024 1 > gpurun_out/ncu_c.log 2>&1
echo "ncu rc=$?"
