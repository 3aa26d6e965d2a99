set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python bench.py --steps 2000 --warmup 100 --update-rounds 100 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1_v2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'upd', d['critic_updates']['value'], d['critic_updates']['ms_per_round'], 'envk', d['roofline']['avg_launch_us'], 'e2e', d['e2e']['value'], d['critic_updates']['e2e']['value'])"
