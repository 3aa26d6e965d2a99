set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 2000 --warmup 100 --update-rounds 100 2>&1 | tail -3 | tee gpurun_out/bench_r1_graph.json
