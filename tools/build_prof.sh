# Builds libmaddpg_b200_prof.so: the library with the episode kernel's clock64 phase marks compiled in (-DMDP_EPISODE_PROF).
#   bash maddpg_b200/csrc/build.sh && bash tools/build_prof.sh
#   MDP_LIB_NAME=libmaddpg_b200_prof.so python tools/prof_episode.py        (on a GPU box)
set -e
cd "$(dirname "${BASH_SOURCE[0]}")/../maddpg_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static -DMDP_EPISODE_PROF -c mdp_rollout.cu -o ../_lib/mdp_rollout_prof.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static -DMDP_EPISODE_PROF ${PROF_DEFS:-} -c mdp_rollout_tc.cu -o ../_lib/mdp_rollout_tc_prof.o
objs=""; for f in mdp_api mdp_host mdp_env mdp_replay mdp_train mdp_train_tc mdp_optim; do objs="$objs ../_lib/$f.o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../_lib/libmaddpg_b200_prof${PROF_SUFFIX:-}.so $objs ../_lib/mdp_rollout_prof.o ../_lib/mdp_rollout_tc_prof.o
echo ok
