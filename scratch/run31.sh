python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 1000 --warmup 100 --update-rounds 50 --e2e-steps 50 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8_err.log; echo "rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/bench_n8_err.log | tail -5
tail -1 gpurun_out/bench_n8.log > gpurun_out/bench_r1_n8.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n8.json'))
print('value', d['value'], 'e2e', d['e2e']['value'])
print('upd', json.dumps(d['critic_updates'], indent=1))
"
