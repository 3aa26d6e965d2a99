// Shared by the episode kernels (mdp_rollout.cu: fp32 SIMT actor tiles; mdp_rollout_tc.cu: tcgen05 actor tiles).
#pragma once
#include "mdp_env_dev.cuh"
#include "mdp_mlp.cuh"

namespace mdp {

constexpr int REB = 32;  // env instances per CTA == rows of the MLP tile

__host__ __device__ inline int actor_net_floats(int D, int U, int K) {
  return ((D * U + U + U * U + U + U * K + K) + 3) & ~3;
}

struct RolloutArgs {
  int E, steps, reset_after;
  int episodes;  // tcgen05 kernel: `episodes` x (steps, then reset_world when reset_after) in one launch
  void* state;  // SoA [state_comps][E], float32 or float64 (mdp_env_cfg.state_f64)
  float* obs;   // (E, obs_stride) joint current observations, in/out
  float* ring;
  long long capacity, cursor;
  unsigned long long seed, counter, env_seed, episode;
  float lm_lo, lm_hi;
  const unsigned long long* ctl;
  float* ep_return;  // optional (E, A): sum of rewards over the launch
};

CoreDev core_dev_for_rollout(const mdp_core* c);
// tcgen05 episode kernel (mdp_rollout_tc.cu): returns MDP_ENOTSUP when the configuration is not covered
int rollout_episode_tc(mdp_env* env, mdp_core* core, const mdp_ring_layout& lay, const RolloutArgs& R, cudaStream_t st);

}  // namespace mdp
