N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/bench_n${N}_v3.log 2> gpurun_out/bench_n${N}_v3_err.log; echo "rc=$?"
tail -1 gpurun_out/bench_n${N}_v3.log | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'upd', d['critic_updates']['value'], d['critic_updates']['ms_per_round'], 'grouped', d['critic_updates']['grouped']['value'])
"
