python tools/profile_tc.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_td_target_tc -c 1 -o gpurun_out/prof_tc2 python tools/profile_tc.py > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
