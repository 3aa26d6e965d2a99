set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python bench.py --steps 2000 --warmup 100 --update-rounds 100 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1_mega2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['critic_updates']['value'], d['roofline']['avg_launch_us'], d['e2e']['value'])"
