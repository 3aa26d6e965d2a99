timeout 120 python scratch/ep_once.py || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_episode -s 2 -c 1 -o gpurun_out/ep_v3 -f python scratch/ep_once.py > gpurun_out/ncu_ep.log 2>&1; echo "ncu rc=$?"
