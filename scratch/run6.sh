set -x
python -m pytest tests -m gpu -q -x -k "episode_kernel" 2>&1 | tail -3
python tools/profile_mega.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_rollout_episode -c 1 -o gpurun_out/prof_mega python tools/profile_mega.py > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
