timeout 300 python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 100 --update-rounds 50 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('peer path upd', d['critic_updates']['value'], d['critic_updates']['ms_per_round'], d['critic_updates']['gradient_exchange'], d['critic_updates']['grouped'])"
