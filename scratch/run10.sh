python tools/profile_workload.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_td_target_res|k_critic_grads_res' -c 2 -o gpurun_out/prof_upd2 python tools/profile_workload.py > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
