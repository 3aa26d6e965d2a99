python tools/profile_tc.py simple_spread 24 2048 1024 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_critic_grads_tc|k_actor_grads_tc|k_td_target_tc" -c 3 -o gpurun_out/prof_tc5g_v3 python tools/profile_tc.py simple_spread 24 2048 1024 1 > gpurun_out/ncu_c.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_c.log
