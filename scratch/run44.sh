N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n${N}_err.log; echo "rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/bench_n${N}_err.log | tail -3
tail -1 gpurun_out/bench_n$N.log > gpurun_out/bench_r1_final_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_final_n$N.json'))
print('N', d['n_gpus'], 'value %.3fG' % (d['value']/1e9), 'e2e %.1fM' % (d['e2e']['value']/1e6), 'upd %.0f' % d['critic_updates']['value'], 'ms/round %.3f' % d['critic_updates']['ms_per_round'], d['critic_updates']['gradient_exchange'], 'grouped %.0f' % d['critic_updates']['grouped']['value'])
"
