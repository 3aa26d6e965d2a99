python tools/profile_tc.py simple_spread 24 2048 1024 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_critic_grads_tc|k_actor_grads_tc|k_td_target_tc|k_dw1_tc" -c 5 -o gpurun_out/prof_tc5g_v5 python tools/profile_tc.py simple_spread 24 2048 1024 1 > gpurun_out/ncu_c.log 2>&1
echo "ncu rc=$?"
