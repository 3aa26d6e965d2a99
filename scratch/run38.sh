python tools/profile_tc_critic.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_critic_grads_tc -s 1 -c 1 -o gpurun_out/prof_tcc python tools/profile_tc_critic.py > gpurun_out/ncu_c.log 2>&1
echo rc=$?
