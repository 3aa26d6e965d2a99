set -x
nvidia-smi -L
python -m pytest tests/test_distributed_gpu.py -m gpu -q -x 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2000 --warmup 100 --update-rounds 100 2>&1 | tail -2 | tee gpurun_out/bench_r1_n2.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'upd', d['critic_updates']['value'], 'allreduce B/round', d['critic_updates']['allreduce_bytes_per_round'], 'e2e', d['e2e']['value'])
"
