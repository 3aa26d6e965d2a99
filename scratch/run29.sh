python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 100 --update-rounds 50 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2_err.log; echo "rc=$?"
tail -3 gpurun_out/bench_n2_err.log
tail -1 gpurun_out/bench_n2.log > gpurun_out/bench_r1_n2_peer.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n2_peer.json'))
print('value', d['value'], 'e2e', d['e2e']['value'])
print('upd', json.dumps(d['critic_updates'], indent=1))
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 500 --warmup 100 --update-rounds 50 --nccl-allreduce 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NCCL path upd', d['critic_updates']['value'], d['critic_updates']['ms_per_round'], d['critic_updates']['gradient_exchange'])"
