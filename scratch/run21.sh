python tools/profile_tc.py simple_spread 24 2048 1024 0 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_td_target_tc -c 1 -o gpurun_out/prof_tc5 python tools/profile_tc.py simple_spread 24 2048 1024 0 > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
