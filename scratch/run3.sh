set -x
python -m pytest tests/test_trainer_gpu.py -m gpu -q -x -k graph 2>&1 | tail -3
python tools/profile_workload.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_cfg2.csv python tools/profile_workload.py > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
python tools/profile_workload.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_env_step|k_critic_grads|k_actor_grads|k_td_target|k_actor_act' -c 10 -o gpurun_out/prof_cfg2 python tools/profile_workload.py > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
