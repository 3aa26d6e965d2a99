for lib in libmaddpg_b200_pd2.so libmaddpg_b200.so libmaddpg_b200_pd4.so; do
  echo "== $lib"
  MDP_LIB_NAME=$lib timeout 200 python scratch/time_tc.py 2>&1 | grep "n=24" -B1 | grep -v "^--"
done
