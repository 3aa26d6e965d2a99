set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 500 --warmup 50 --update-rounds 30 --no-graph 2>&1 | tail -5 | tee gpurun_out/bench_r1_nograph.json
python bench.py --impl reference --steps 300 --warmup 20 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
