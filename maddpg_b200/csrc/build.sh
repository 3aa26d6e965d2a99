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
for f in mdp_api mdp_env mdp_replay mdp_train mdp_optim mdp_rollout; do
  "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$out/$f.o" 2> "$out/$f.ptxas.log" || { cat "$out/$f.ptxas.log"; exit 1; }
  objs+=("$out/$f.o")
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$out/libmaddpg_b200.so" "${objs[@]}"
echo "built $out/libmaddpg_b200.so"
