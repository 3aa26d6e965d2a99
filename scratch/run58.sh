SECONDS=0
BARGS="--steps 100 --warmup 25 --update-rounds 5 --e2e-steps 25 --no-cpu-baseline --no-tensor-section"
python bench.py $BARGS > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1_v3.csv python bench.py $BARGS > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$? t=${SECONDS}"
python tools/profile_mega.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_rollout_episode -c 1 -o gpurun_out/prof_mega_v3 -f python tools/profile_mega.py > gpurun_out/ncu_a.log 2>&1
echo "mega rc=$? t=${SECONDS}"
