for i in 1 2; do
timeout 600 python bench.py > gpurun_out/bench_full_$i.log 2> gpurun_out/bench_full_err.log; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_full_$i.log').read().strip().splitlines()[-1])
print('value', d['value'], 'ep_us', d['roofline']['avg_launch_us'], 'e2e', d['e2e']['value'], 'upd', d['critic_updates']['value'], 'grouped', d['critic_updates']['grouped']['value'])
"
done
nvidia-smi --query-gpu=name,temperature.gpu,power.draw,clocks.sm --format=csv
