// Trainer-core structures shared by mdp_train.cu (forward/backward kernels) and mdp_optim.cu.
#pragma once
#include "mdp_common.cuh"

#include <vector>

namespace mdp {

struct MlpW {
  const float *W1, *b1, *W2, *b2, *W3, *b3;
  int in, out;
};
struct MlpG {
  float *W1, *b1, *W2, *b2, *W3, *b3;
};

struct AgentDev {
  MlpW net[4];  // MDP_NET_P, TARGET_P, Q, TARGET_Q
  MlpG grad[2]; // P, Q
  int obs_dim, act_dim, obs_off, act_off, n_heads, head_dim[2], local_q, q_in;
};

struct CoreDev {
  const AgentDev* agents;
  int n_agents, units;
  int obs_sum, act_sum, act_stride;
  double gamma, actor_reg;
  int* adam_t;
  double* stats;
  const unsigned long long* ctl;
};

// mdp_clip_adam_polyak[_all] with an optional programmatic dependent launch on the gradient kernel that precedes it (mdp_optim.cu)
int clip_adam_polyak_impl(mdp_core* c, int32_t agent, int32_t which, float grad_scale, int32_t do_polyak, void* stream, bool pdl);
int clip_adam_polyak_all_impl(mdp_core* c, int32_t which, float grad_scale, int32_t do_polyak, void* stream, bool pdl);

}  // namespace mdp

struct mdp_core {
  mdp_core_cfg cfg;
  mdp_core_layout lay;
  int obs_off[MDP_MAX_AGENTS], act_off[MDP_MAX_AGENTS];
  int obs_sum = 0, act_sum = 0, act_stride = 0;
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  int32_t* adam_t = nullptr;
  double* stats = nullptr;
  mdp::AgentDev* d_agents = nullptr;
  const unsigned long long* ctl = nullptr;
  std::vector<mdp::AgentDev> h_agents;
  // peer (NVLink) gradient exchange: every rank's gradient bucket and flag words, mapped into this process
  int peer_world = 0, peer_rank = 0;
  const float* const* d_peer_grads = nullptr;  // device array [world] of gradient-buffer bases
  unsigned* const* d_peer_flags = nullptr;     // device array [world] of flag-word bases
  unsigned* peer_epoch = nullptr;              // local epoch counters, one per (agent, variable) slot
  uint2* const* d_peer_recv = nullptr;         // device array [world] of low-latency receive buffers (or null: barrier mode)
  void* d_peer_tables = nullptr;
  int no_fuse = 0;                 // 1: keep TD target and critic step as two launches (mdp_core_set_fused_update)
  int prep_agent = -1, prep_count = 0;  // mdp_update_prepare zeroed these agents' statistics in the launch right before (mdp_train.cu)
  int tc_mode = 0;                 // tensor-core (tcgen05) kernels: 0 auto, 1 always where supported, -1 never
  float* tc_scratch = nullptr;     // sampled-action tiles of the tensor-core TD-target kernel when they exceed shared memory
  size_t tc_scratch_bytes = 0;
  float* norm2 = nullptr;             // squared gradient norms of wide W1 variables (mdp_optim.cu)
  float* peer_reduced = nullptr;      // rank-summed gradients of wide W1 variables (fused peer exchange, many-CTA path)
  float* tc_dz1 = nullptr;            // dz1 tiles handed from k_critic_grads_tc to k_dw1_tc (wide critics)
  size_t tc_dz1_bytes = 0;
  unsigned char* tc_arena = nullptr;  // pre-split UMMA weight images of every net (csrc/mdp_train_tc.cu)
  const void* tc_imgs = nullptr;      // device table of tc::AgentImg inside the arena
};

