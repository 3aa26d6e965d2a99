// Host-buffer entry points: the reference loop body (experiments/train.py:112-120) for E lockstep env
// instances in ONE call whose inputs and outputs are HOST arrays.  Everything between the two copies runs
// in the kernels of mdp_train.cu / mdp_env.cu / mdp_replay.cu; nothing here computes on the CPU.
#include "mdp_common.cuh"

using namespace mdp;

static inline int64_t up256(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int mdp_host_step_layout(const mdp_env* env, int32_t E, int64_t* offs4, int64_t* total_bytes) {
  MDP_REQUIRE(env && offs4 && total_bytes && E > 0, "mdp_host_step_layout: bad argument");
  mdp_env_dims d;
  int rc = mdp_env_get_dims(env, &d);
  if (rc) return rc;
  int64_t o = 0;
  offs4[0] = o; o = up256(o + 4ll * E * d.obs_stride);   // next observations (E, obs_stride) f32
  offs4[1] = o; o = up256(o + 4ll * E * d.n_agents);     // rewards (E, n_agents) f32
  offs4[2] = o; o = up256(o + 4ll * E * d.act_stride);   // sampled actions (E, act_stride) f32
  offs4[3] = o; o = up256(o + 1ll * E * d.n_agents);     // done (E, n_agents) u8
  *total_bytes = o;
  return MDP_OK;
}

extern "C" int mdp_host_step(mdp_env* env, mdp_core* core, int32_t E, void* state, const float* h_obs_in, float* d_obs_in,
                             void* d_out, void* h_out, float* ring, int64_t ring_capacity, int32_t ring_row_stride,
                             int64_t ring_cursor, uint64_t seed, uint64_t counter, void* stream) {
  MDP_REQUIRE(env && core && state && h_obs_in && d_obs_in && d_out && h_out && E > 0, "mdp_host_step: bad argument");
  mdp_env_dims d;
  int rc = mdp_env_get_dims(env, &d);
  if (rc) return rc;
  int64_t off[4], total;
  rc = mdp_host_step_layout(env, E, off, &total);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* dout = static_cast<char*>(d_out);
  float* d_obs = reinterpret_cast<float*>(dout + off[0]);
  float* d_rew = reinterpret_cast<float*>(dout + off[1]);
  float* d_act = reinterpret_cast<float*>(dout + off[2]);
  uint8_t* d_done = reinterpret_cast<uint8_t*>(dout + off[3]);
  // obs_n (host) -> device: the argument of agent.action(obs), train.py:112
  MDP_CUDA(cudaMemcpyAsync(d_obs_in, h_obs_in, 4ull * E * d.obs_stride, cudaMemcpyHostToDevice, st));
  // action_n = [agent.action(obs) ...]: grouped actor inference + Gumbel-softmax sampling
  rc = mdp_actor_act(core, 0, d.n_agents, 0, E, d_obs_in, d.obs_stride, d_act, d.act_stride, nullptr, seed, counter, nullptr, stream);
  if (rc) return rc;
  // new_obs_n, rew_n, done_n = env.step(action_n) (train.py:114) + agent.experience(...) for every agent (train.py:119-120)
  rc = mdp_env_step(env, E, state, d_act, d_obs, d_rew, d_done, ring ? d_obs_in : nullptr, ring, ring_capacity, ring_row_stride,
                    ring_cursor, stream);
  if (rc) return rc;
  // one packed device -> host copy of everything the loop reads back
  MDP_CUDA(cudaMemcpyAsync(h_out, d_out, (size_t)total, cudaMemcpyDeviceToHost, st));
  return MDP_OK;
}
