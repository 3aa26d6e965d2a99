#!/usr/bin/env bash
# Builds libmaddpg_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../_lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       -Xptxas -v --expt-relaxed-constexpr -cudart static)
objs=()
pids=()
srcs=(mdp_api mdp_host mdp_env mdp_replay mdp_prio mdp_train mdp_td3 mdp_train_tc mdp_optim mdp_rollout mdp_rollout_tc)
for f in "${srcs[@]}"; do  # one nvcc per translation unit, in parallel
  "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$out/$f.o" 2> "$out/$f.ptxas.log" &
  pids+=($!)
  objs+=("$out/$f.o")
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then cat "$out/${srcs[$i]}.ptxas.log"; fail=1; fi
done
[ "$fail" = 0 ] || exit 1
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$out/libmaddpg_b200.so" "${objs[@]}"
echo "built $out/libmaddpg_b200.so"
