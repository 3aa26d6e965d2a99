python tools/profile_workload.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_cfg2_v2.csv python tools/profile_workload.py > gpurun_out/ncu1.log 2>&1
python tools/profile_workload.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_critic_grads|k_td_target|k_actor_grads|k_clip_adam' -c 5 -o gpurun_out/prof_upd python tools/profile_workload.py > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
