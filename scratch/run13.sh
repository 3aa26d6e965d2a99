python tools/profile_env.py simple_spread 262144 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_env_step -c 1 -o gpurun_out/prof_env python tools/profile_env.py simple_spread 262144 > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log
