/*
 * maddpg_b200.h -- C ABI of the B200-native MADDPG hot path (libmaddpg_b200.so).
 *
 * The reference (adolfogonzalez3/maddpg) is 100% Python and has no FFI; the hot path sits
 * behind two duck-typed Python interfaces:
 *   - maddpg.AgentTrainer / MADDPGAgentTrainer      maddpg/__init__.py:1-15, maddpg/trainer/maddpg.py:112-196
 *   - multiagent.environment.MultiAgentEnv + Scenario   call sites experiments/train.py:49-60,104,114,128
 * Each entry point below names the reference function(s) it replaces.  The Python host code in
 * maddpg_b200/{env,replay,trainer}.py binds these with ctypes and mirrors the reference surface.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  All data pointers are DEVICE pointers unless the
 *     parameter name starts with h_.  The caller owns every buffer (the library never frees caller
 *     memory and keeps no pointer past the call, except the buffers bound with mdp_core_bind()).
 *   - Every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising.  Return value: 0 = MDP_OK, negative = error class; the message is available
 *     from mdp_last_error() (thread local).  Asynchronous launch failures surface at the next call.
 *   - Handles are not thread-safe.  One host thread per GPU.
 *
 * Joint layouts (row-major float32 unless noted; E = number of env instances in the call)
 *   obs   (E, obs_stride)   agent i's observation in columns [obs_off[i], obs_off[i]+obs_dim[i])
 *   act   (E, act_stride)   agent i's soft one-hot action in columns [act_off[i], act_off[i]+act_dim[i])
 *   rew   (E, n_agents)     done (E, n_agents) uint8
 *   state SoA [state_comps][E] of float32 (state_f64 = 0) or float64 (state_f64 = 1):
 *         comp 4*i+{0,1,2,3} = agent i pos.x, pos.y, vel.x, vel.y ; then comm_dim comm values of the
 *         non-silent agent(s) ; then 2*l+{0,1} = landmark l pos.x, pos.y
 *   replay ring (capacity, row_stride): one JOINT row per lockstep transition
 *         [ obs_0..obs_{n-1} | act_0..act_{n-1} | pad ]   columns [0, x_dim)           = critic input
 *         [ next_obs_0..next_obs_{n-1} | pad ]            columns [nx_off, nx_off+sum D)
 *         [ rew_0..rew_{n-1} | done_0..done_{n-1} | pad ] columns [rw_off, ..), [dn_off, ..)
 *         every sub-block starts on a 16-byte boundary; row_stride is a multiple of 4 floats.
 */
#ifndef MADDPG_B200_H_
#define MADDPG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDP_MAX_AGENTS 32
#define MDP_MAX_HEADS 2

enum { MDP_OK = 0, MDP_EINVAL = -1, MDP_ECUDA = -2, MDP_ENOTSUP = -3 };

/* scenario ids: multiagent/scenarios/<name>.py, loaded at experiments/train.py:53 */
enum { MDP_SIMPLE = 0, MDP_SIMPLE_SPREAD = 1, MDP_SIMPLE_TAG = 2, MDP_SIMPLE_WORLD_COMM = 3,
       /* the other scenarios `train.py --scenario` can name */
       MDP_SIMPLE_ADVERSARY = 4, MDP_SIMPLE_PUSH = 5, MDP_SIMPLE_SPEAKER_LISTENER = 6, MDP_SIMPLE_CRYPTO = 7,
       MDP_SIMPLE_REFERENCE = 8 };

typedef struct mdp_env_cfg {
  int32_t scenario;   /* MDP_SIMPLE ... */
  int32_t num_agents; /* simple_spread only: N agents = N landmarks (0 -> 3, max 32) */
  int32_t state_f64;  /* 0: float32 state (throughput); 1: float64 state like MPE's numpy (parity) */
} mdp_env_cfg;

typedef struct mdp_env_dims {
  int32_t n_agents, n_landmarks, comm_dim, collaborative;
  int32_t obs_dim[MDP_MAX_AGENTS], act_dim[MDP_MAX_AGENTS];
  int32_t obs_off[MDP_MAX_AGENTS], act_off[MDP_MAX_AGENTS];
  int32_t n_heads[MDP_MAX_AGENTS], head_dim[MDP_MAX_AGENTS][MDP_MAX_HEADS];
  int32_t obs_sum, act_sum, obs_stride, act_stride;
  int32_t state_comps;     /* 4*n_agents + comm_dim + 2*n_landmarks + n_goal */
  int32_t state_elem_size; /* 4 or 8 */
  int32_t env_bytes_per_step; /* algorithmic bytes of one env step, SURVEY 8(d) formula (fp32 state) */
  /* SoA state rows: [4i .. 4i+3] = agent i's (x, y, vx, vy); [4A + comm_off[i], + comm_len[i]) = agent i's state.c (speaking
   * agents only; comm_dim is the sum); then 2 rows per landmark; then n_goal rows holding the landmark indices reset_world
   * drew with np.random.choice (agent.goal_a / goal_b / key), stored as numbers of the state's type */
  int32_t n_goal;
  int32_t comm_off[MDP_MAX_AGENTS], comm_len[MDP_MAX_AGENTS];
  int32_t movable[MDP_MAX_AGENTS];
} mdp_env_dims;

typedef struct mdp_env mdp_env;

/* scenario.make_world() + MultiAgentEnv.__init__ (train.py:48-61): builds the entity table and the
 * observation column table.  No device work (usable without a GPU). */
int mdp_env_create(const mdp_env_cfg* cfg, mdp_env** out);
int mdp_env_get_dims(const mdp_env* env, mdp_env_dims* out);
void mdp_env_destroy(mdp_env* env);

/* scenario.reset_world() + env.reset() (train.py:104,128).  If init_state is non-null it is a device
 * SoA array of the env's state precision that is copied into `state` (injected reset, parity tests);
 * otherwise positions are drawn on device with Philox4x32-10 keyed by (seed, episode, env, comp).
 * Writes the reset observations to obs_out (joint layout). */
int mdp_env_reset(mdp_env* env, int32_t E, void* state, const void* init_state, uint64_t seed,
                  uint64_t episode, float* obs_out, void* stream);

/* MultiAgentEnv.step(action_n) -> World.step() -> per-agent observation()/reward() (train.py:114;
 * SURVEY Appendix A.2): one fused kernel.  state is updated in place.  ring (optional, may be null): when non-null the
 * SAME CALL also writes the joint replay rows of this transition (obs_t, act_t, next_obs, rew, done) to ring rows
 * (ring_cursor + e) % ring_capacity -- MADDPGAgentTrainer.experience / ReplayBuffer.add (maddpg.py:154-156,
 * replay_buffer.py:25-32) for all agents.  simple_spread with 7..32 agents (float32 state): the step kernel writes the rows
 * itself (one warp per env instance assembles obs_t | act_t | next_obs | rew | done while it produces the observations);
 * every other case: a second launch (the joint insert kernel) on the same stream.  The persistent episode kernels
 * (mdp_rollout_episode) always write the rows from inside the step.  obs_prev (joint obs_t, a buffer distinct from obs_out)
 * must then be non-null. */
int mdp_env_step(mdp_env* env, int32_t E, void* state, const float* act, float* obs_out, float* rew_out,
                 uint8_t* done_out, const float* obs_prev, float* ring, int64_t ring_capacity,
                 int32_t ring_row_stride, int64_t ring_cursor, void* stream);

/* mdp_env_step picks a register-resident one-thread-per-env kernel for simple_spread with <= 6 agents and a
 * one-warp-per-env kernel for 7..32 agents (float32 state); on = 1 forces the table-driven kernel that serves every
 * scenario (tests compare them). */
int mdp_env_force_generic(mdp_env* env, int32_t on);

/* Scenario.benchmark_data(agent, world) for every (env instance, agent): the info_n tape of `train.py --benchmark`
 * (experiments/train.py:139-148, MultiAgentEnv info_callback).  out: DEVICE (E, n_agents, 4) f32 --
 * simple_spread (reward, collisions, sum of the landmarks' closest-agent distances, occupied landmarks);
 * simple_tag / simple_world_comm (collisions with good agents, 0, 0, 0) for adversaries and zeros for good agents;
 * simple: zeros (the scenario defines no benchmark_data). */
int mdp_env_benchmark(mdp_env* env, int32_t E, const void* state, float* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* replay ring (maddpg/trainer/replay_buffer.py)                                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct mdp_ring_layout {
  int32_t n_agents;
  int32_t obs_dim[MDP_MAX_AGENTS], act_dim[MDP_MAX_AGENTS];
  int32_t obs_off[MDP_MAX_AGENTS], act_off[MDP_MAX_AGENTS]; /* within the obs / act blocks */
  int32_t obs_sum, act_sum;
  int32_t x_dim;      /* obs_sum + act_sum = centralized critic input width C */
  int32_t nx_off, rw_off, dn_off, row_stride; /* in floats */
} mdp_ring_layout;

/* Derives the joint-row layout from per-agent dims (host only). */
int mdp_ring_make_layout(int32_t n_agents, const int32_t* obs_dim, const int32_t* act_dim, mdp_ring_layout* out);

/* ReplayBuffer.add for E lockstep transitions (replay_buffer.py:25-32): row (cursor + e) % capacity
 * <- (obs[e], act[e], rew[e], next_obs[e], done[e]).  agent = -1 writes every agent's columns;
 * agent = i writes only agent i's columns (per-agent experience() calls, maddpg.py:154-156). */
int mdp_replay_insert(const mdp_ring_layout* lay, float* ring, int64_t capacity, int64_t cursor, int32_t E,
                      int32_t agent, const float* obs, int32_t obs_stride, const float* act, int32_t act_stride,
                      const float* rew, int32_t rew_stride, const float* next_obs, int32_t next_obs_stride,
                      const uint8_t* done, int32_t done_stride, void* stream);

/* ReplayBuffer.sample_index / _encode_sample (replay_buffer.py:34-44,55-56): out[b, :] = ring[idx[b], :]
 * for b < B (whole joint rows; per-agent fields are column views of `out`).  mode 0 = vectorised
 * warp-per-row copy, mode 1 = cp.async.bulk (TMA bulk copy engine) row copies through shared memory. */
int mdp_replay_gather(const float* ring, int64_t capacity, int32_t row_stride, const int64_t* idx, int32_t B,
                      float* out, int32_t mode, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* trainer core (maddpg/trainer/maddpg.py)                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct mdp_core_cfg {
  int32_t n_agents;
  int32_t num_units; /* args.num_units: 64 or 128 (train.py:24) */
  int32_t obs_dim[MDP_MAX_AGENTS], act_dim[MDP_MAX_AGENTS];
  int32_t n_heads[MDP_MAX_AGENTS], head_dim[MDP_MAX_AGENTS][MDP_MAX_HEADS];
  int32_t local_q[MDP_MAX_AGENTS]; /* local_q_func (ddpg mode), maddpg.py:51-52,86-87 */
  double lr, gamma, polyak, grad_clip, actor_reg, beta1, beta2, adam_eps; /* polyak = 0.99 (maddpg.py:21) */
} mdp_core_cfg;

/* network ids inside one agent's parameter block */
enum { MDP_NET_P = 0, MDP_NET_TARGET_P = 1, MDP_NET_Q = 2, MDP_NET_TARGET_Q = 3 };

typedef struct mdp_core_layout {
  int64_t total_params;               /* floats in the flat parameter buffer (all agents, 4 nets each) */
  int64_t total_train;                /* floats in the flat grad / adam_m / adam_v buffers (P and Q nets) */
  int64_t net_off[MDP_MAX_AGENTS][4]; /* offset of [W1|b1|W2|b2|W3|b3] of each net in `params` */
  int64_t train_off[MDP_MAX_AGENTS][2]; /* offset of the P (0) and Q (1) net in grads/adam buffers */
  int32_t net_in[MDP_MAX_AGENTS][4], net_out[MDP_MAX_AGENTS][4];
  int64_t net_size[MDP_MAX_AGENTS][4];
  int64_t update_flops_critic[MDP_MAX_AGENTS], update_flops_actor[MDP_MAX_AGENTS]; /* per batch row, SURVEY 8(d) */
} mdp_core_layout;

typedef struct mdp_core mdp_core;

/* MADDPGAgentTrainer.__init__ x n (maddpg.py:113-149): derives the flat parameter layout. Host only. */
int mdp_core_create(const mdp_core_cfg* cfg, mdp_core** out);
int mdp_core_get_layout(const mdp_core* core, mdp_core_layout* out);
void mdp_core_destroy(mdp_core* core);

/* Registers the caller-owned device buffers (lifetime: until mdp_core_destroy or the next bind).
 * params[total_params]; grads, adam_m, adam_v [total_train]; adam_t[2*n_agents] int32 step counters
 * (P, Q per agent); stats[8*n_agents] float64 accumulators. */
int mdp_core_bind(mdp_core* core, float* params, float* grads, float* adam_m, float* adam_v, int32_t* adam_t,
                  double* stats);

/* Selects the MLP kernels: 0 = automatic, 1 = the tcgen05 tensor-core kernels wherever a shape is supported
 * (3xTF32 split GEMMs with TMEM accumulators, csrc/mdp_train_tc.cu), -1 = the fp32 SIMT kernels only.  Both
 * compute the same functions (mlp_model, train.py:39-46) to ~fp32 accuracy. */
int mdp_core_set_tensor_cores(mdp_core* core, int32_t mode);

/* mdp_update_agent / mdp_update_all run the TD target (maddpg.py:181-189) and the critic step (q_train, :75-100)
 * of the same sampled rows in ONE launch when the row tile fits in shared memory (on = 1, the default); on = 0 keeps
 * the two launches.  Same arithmetic either way. */
int mdp_core_set_fused_update(mdp_core* core, int32_t on);

/* MADDPGAgentTrainer.action (maddpg.py:151-152) / p_debug['target_act'] (:70-71) for agents
 * [agent_begin, agent_begin+agent_count): act_i = gumbel_softmax(mlp(obs_i)) in one grouped launch.
 * obs/act are joint arrays.  u (optional, joint act layout): injected U[0,1) draws; when null the
 * kernel draws Philox4x32-10 uniforms keyed by (seed, counter, env, column).  use_target selects the
 * target actor.  logits_out (optional, joint act layout) receives the pre-noise logits (p_values). */
int mdp_actor_act(mdp_core* core, int32_t agent_begin, int32_t agent_count, int32_t use_target, int32_t E,
                  const float* obs, int32_t obs_stride, float* act, int32_t act_stride, const float* u,
                  uint64_t seed, uint64_t counter, float* logits_out, void* stream);

/* q_debug['q_values'] / ['target_q_values'] (maddpg.py:101,108): q[b] = Q_j(x[b, :]) where x is a
 * (B, x_stride) array whose first x_dim columns are the joint critic input. */
int mdp_critic_q(mdp_core* core, int32_t agent, int32_t use_target, int32_t B, const float* x, int32_t x_stride,
                 float* q_out, void* stream);

/* Row addressing of the three update kernels below: `batch` is a (rows, row_stride) array of joint ring rows.
 * idx == NULL: logical row b is batch row b (an already gathered batch).  idx != NULL (device int64[B]): logical
 * row b is batch row idx[b] -- pass the replay ring itself and the sampled indices and the gather
 * (ReplayBuffer.sample_index, replay_buffer.py:34-44,55-56) is fused into the kernels' tile loads.
 *
 * TD target of agent j (maddpg.py:181-187) on a gathered batch (B, row_stride) of joint ring rows:
 * a'_i = gumbel_softmax(target_p_i(next_obs_i)) for all i, q' = target_q_j(next_obs, a'),
 * y = float32(rew_j + gamma * (1 - done_j) * q') -- one fused kernel.  u_target (optional, (B, act_stride)
 * joint layout) injects the uniforms.  Accumulates sum(y), sum(y^2), sum(rew), sum(q') into stats. */
int mdp_td_target(mdp_core* core, int32_t agent, const mdp_ring_layout* lay, int32_t B, const float* batch,
                  const int64_t* idx, const float* u_target, int32_t u_stride, uint64_t seed, uint64_t counter,
                  float* y_out, float* target_act_out, void* stream);

/* mdp_td_target for every agent in ONE grouped launch (grid.y = agent; the first stage of mdp_update_all): idx is NULL or
 * int64 [n_agents][idx_agent_stride] index sets (stride 0 shares one set), y_out is float [n_agents][B]; in-kernel Philox
 * noise only. */
int mdp_td_target_all(mdp_core* core, const mdp_ring_layout* lay, int32_t B, const float* batch, const int64_t* idx,
                      int64_t idx_agent_stride, uint64_t seed, uint64_t counter, float* y_out, void* stream);

/* q_train forward/backward (maddpg.py:75-100): grads of mean((Q_j(x) - y)^2) wrt the critic's six
 * tensors are ACCUMULATED into the bound grad buffer (fused fwd + bwd kernel); sum((q-y)^2) -> stats. */
int mdp_critic_grads(mdp_core* core, int32_t agent, const mdp_ring_layout* lay, int32_t B, const float* batch,
                     const int64_t* idx, const float* y, float* q_out, void* stream);

/* mdp_critic_grads for every agent in ONE grouped launch (grid.y = agent): y is float [n_agents][B], idx NULL or int64
 * [n_agents][idx_agent_stride] (stride 0 shares one set). */
int mdp_critic_grads_all(mdp_core* core, const mdp_ring_layout* lay, int32_t B, const float* batch, const int64_t* idx,
                         int64_t idx_agent_stride, const float* y, void* stream);

/* p_train forward/backward (maddpg.py:28-61): grads of -mean(Q_j(o, a_-j, gumbel_softmax(p_j(o_j))))
 * + actor_reg * mean(logits^2) wrt the actor's tensors, through the RUNNING critic (fused kernel). */
int mdp_actor_grads(mdp_core* core, int32_t agent, const mdp_ring_layout* lay, int32_t B, const float* batch,
                    const int64_t* idx, const float* u_actor, int32_t u_stride, uint64_t seed, uint64_t counter,
                    void* stream);

/* U.minimize_and_clip + tf.train.AdamOptimizer + make_update_exp (tf_util.py:166-182, maddpg.py:20-26):
 * per-variable clip_by_norm, TF-Adam step (t = adam_t, incremented by the *_grads call), polyak update
 * of the matching target net, and re-zeroing of the grad segment -- one kernel.  which: 0 = actor (P),
 * 1 = critic (Q).  grad_scale multiplies the gradient first (1/world_size after an allreduce). */
int mdp_clip_adam_polyak(mdp_core* core, int32_t agent, int32_t which, float grad_scale, int32_t do_polyak,
                         void* stream);

/* Multi-GPU, fused exchange: binds every rank's gradient buffer and flag words (mapped into this process, e.g. by
 * torch.distributed._symmetric_memory; h_peer_grads[rank] must be the buffer given to mdp_core_bind) so that
 * mdp_clip_adam_polyak[_all] sums the gradient over ranks itself -- peer loads over NVLink inside the clip+Adam+polyak
 * kernel, flag barriers between the CTAs that own the same variable -- instead of expecting an all-reduced bucket.
 * With peers bound grad_scale is still applied (pass 1/world).  flags: uint32 [12 * n_agents][8] per rank (slot = 12 * agent + 6 * net + variable), zeroed;
 * epoch_local: uint32 [12 * n_agents] device words of this rank, zeroed.  h_peer_recv (optional): every rank's low-latency
 * receive buffer, uint32 pairs [2][world][total_train], zeroed -- when given, ranks PUSH (value, epoch) words into the peers'
 * buffers and poll their own (one NVLink traversal, no barrier); when NULL the kernel uses flag barriers and peer loads.
 * world <= 1 unbinds.  The gradient all-reduce is the only collective of the path. */
int mdp_core_bind_peers(mdp_core* core, int32_t world, int32_t rank, const void* const* h_peer_grads, void* const* h_peer_flags,
                        uint32_t* epoch_local, void* const* h_peer_recv);

/* The index draw of an update (ReplayBuffer.make_index, replay_buffer.py:46-47, on the device: the Philox stream of
 * mdp_replay_make_index, counter [+ ctl[0]], length <= 0 reads the control block's ring length) fused with the reset of the
 * statistics accumulators of agents [agent, agent + count): B_total = B draws per agent x count agents into idx_out.  When the
 * NEXT call on the core is mdp_update_agent(agent) (count = 1) or mdp_update_all (agent = 0, count = n_agents) on the same
 * stream, that call skips its own reset and its TD-target kernel starts as a programmatic dependent launch: its nets stream
 * into shared memory while the draw runs. */
int mdp_update_prepare(mdp_core* core, int32_t agent, int32_t count, int64_t* idx_out, int32_t B_total, int64_t length,
                       uint64_t seed, uint64_t counter, void* stream);

/* MADDPGAgentTrainer.update body for agent j on one stream (maddpg.py:181-194), single GPU:
 * td_target -> critic_grads -> clip_adam(Q) -> actor_grads -> clip_adam(P) + polyak(P) + polyak(Q). */
int mdp_update_agent(mdp_core* core, int32_t agent, const mdp_ring_layout* lay, int32_t B, const float* batch,
                     const int64_t* idx, const float* u_target, const float* u_actor, int32_t u_stride, uint64_t seed,
                     uint64_t counter, float* y_scratch, void* stream);

/* Grouped round ("Jacobi" order): the five kernels of mdp_update_agent launched ONCE for all agents
 * (grid.y = agent): every TD target uses the pre-round target actors, then all critics step, then all actors.
 * Differs from the reference's agent-by-agent order (train.py:160-161) only in that agent j does not see the
 * polyak step of agents i < j taken earlier in the same round -- a documented throughput mode, not the parity
 * mode.  idx: NULL, or int64 [n_agents][idx_agent_stride] index sets (idx_agent_stride = 0 shares one set);
 * y_scratch: float [n_agents][B].  grad_scale as in mdp_clip_adam_polyak. */
int mdp_update_all(mdp_core* core, const mdp_ring_layout* lay, int32_t B, const float* batch, const int64_t* idx,
                   int64_t idx_agent_stride, uint64_t seed, uint64_t counter, float* y_scratch, float grad_scale,
                   void* stream);
/* mdp_clip_adam_polyak for every agent's actor (which = 0) or critic (which = 1) in one launch. */
int mdp_clip_adam_polyak_all(mdp_core* core, int32_t which, float grad_scale, int32_t do_polyak, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* the fork's tanh-policy algorithms: MATD3 and the best/worst-policy "COMA" variant            */
/* (maddpg/modules/{policy,critic,laggingnetwork,matd3module,comamodule}.py) */
/* ------------------------------------------------------------------------------------------ */
/* A policy group or a critic group (policygroup.py:22-42, criticgroup.py:21-41, unshared) is one mdp_core: the group's
 * running / target policies live in its P nets, its running / target critics in its Q nets; a group that is only policies
 * (or only critics) leaves the other nets unused.  The critic step (critic.py:78-88: mse(values - target)) is
 * mdp_critic_grads, the Adam step mdp_clip_adam_polyak with grad_clip = 0 (grad_norm_clipping=None, tf_util.py:171-175) and
 * do_polyak = 0.  All cores of one algorithm share n_agents, the dims and num_units; local_q is not supported.
 *
 * Policy._build (policy.py:63-88) for every agent in one launch: act_i = t * scale_i + shift_i with t = tanh(mlp_i(obs_i))
 * (noise_std = 0: `predict` / `predict_target`) or t = clip(tanh(mlp_i(obs_i)) + clip(noise_std * z, -noise_clip, noise_clip),
 * -1, 1) (`noisy_target`, :72-75).  z: injected N(0,1) draws (B, noise_stride) in the joint action layout, or NULL for
 * in-kernel Philox draws keyed by (seed, counter [+ ctl[0] of an attached control block, mdp_core_set_ctl], row, column).
 * act_scale / act_shift: HOST float[n_agents], the Box rescale `interval` and `interval + low` (:76-84); NULL = 1 and 0.  obs (B, obs_stride) and act (B, act_stride) are joint arrays.
 * shared_agent >= 0: a PolicyGroup(shared=True) -- that agent's policy serves every name (policygroup.py:26-37, 54-70; equal
 * spaces required); -1: one policy per agent. */
int mdp_td3_policy_act(mdp_core* policies, int32_t use_target, int32_t B, const float* obs, int32_t obs_stride,
                       const float* noise, int32_t noise_stride, float noise_std, float noise_clip, uint64_t seed,
                       uint64_t counter, const float* act_scale, const float* act_shift, float* act, int32_t act_stride,
                       int32_t shared_agent, void* stream);

/* MaTD3Module.compute_qvalue (matd3module.py:113-123) / ComaModule.compute_{global,personal}_qvalue (comamodule.py:155-171)
 * for every agent in one launch: q_j = min over the given critic groups (critics_b may be NULL) of Q_j([x | act]) with the
 * target (use_target = 1) or running nets, x = the obs (obs_field = 0) or next_obs (1) columns of the (B, row_stride) joint
 * rows `batch`, act a joint (B, act_stride) action array; y_j = rew_j + gamma * (1 - done_j) * q_j in float32 (the graph's
 * arithmetic), rew_j from the rows, or rew_override[j] (float [n_agents][B]), minus rew_minus[j] when given (ComaModule's
 * personal reward `global value - worst value`, :104-107).  shared_agent >= 0: a CriticGroup(shared=True) -- that agent's critic
 * serves every name (criticgroup.py:24-34, 48-66); -1: one critic per agent.  q_out / y_out: float [n_agents][B], either may be
 * NULL. */
int mdp_td3_q_target(mdp_core* critics_a, mdp_core* critics_b, int32_t use_target, const mdp_ring_layout* lay, int32_t B,
                     const float* batch, int32_t obs_field, const float* act, int32_t act_stride, const float* rew_override,
                     const float* rew_minus, int32_t shared_agent, float gamma, float* q_out, float* y_out, void* stream);

/* The gradient behind Policy.create_optimizer (policy.py:90-100) for every agent in one launch: loss_j = -mean(sign *
 * Q_j(obs, a)) with a = every policy's current action (act_all, from mdp_td3_policy_act) and a_j recomputed from policy j, Q_j
 * the critic group's target (critic_use_target = 1: matd3module.py:96-97, comamodule.py:121,126) or running net;
 * differentiated wrt policy j's variables only (:95).  Accumulates into the policy core's gradient buffer, increments its
 * Adam step counters, adds sum(-sign * q) to stats[8 * j + 1].  sign = -1 is ComaModule's worst policy (:127).
 * critic_agent >= 0: every loss runs through that agent's critic (a shared critic group, or the first name's critic under a
 * shared policy group); shared_policy >= 0: one policy net, loss = -mean(sign * Q_critic_agent) differentiated through EVERY
 * name's action columns into that net's gradient segment (policygroup.py:129-135), step counter and loss counted once. */
int mdp_td3_policy_grads(mdp_core* policies, mdp_core* critics, int32_t critic_use_target, float sign,
                         const mdp_ring_layout* lay, int32_t B, const float* batch, const float* act_all, int32_t act_stride,
                         const float* act_scale, const float* act_shift, int32_t shared_policy, int32_t critic_agent,
                         void* stream);

/* LaggingNetwork.update_target (laggingnetwork.py:36-48): target <- polyak * target + (1 - polyak) * running for every agent's
 * policy (mask bit 0) and / or critic (mask bit 1).  The modules call it with polyak = 5e-3 (matd3module.py:104-107). */
int mdp_td3_polyak(mdp_core* core, int32_t mask, double polyak, void* stream);

/* Persistent episode kernel: `steps` lockstep iterations of experiments/train.py:112-133 (action ->
 * env.step -> experience, optional env.reset at the end) in ONE launch.  Each CTA keeps 32 env instances'
 * state, observation tile, sampled actions and (when they fit) all agents' actor weights in shared
 * memory for the whole episode; only the joint replay rows stream to `ring` at rows
 * (ring_cursor + s*E + e) % ring_capacity.  Step s draws its Gumbel noise from Philox counter
 * `counter + s + 1`, so the result equals `steps` x (mdp_actor_act(counter+s+1), mdp_env_step with ring)
 * [+ mdp_env_reset(env_seed, episode)].  obs: joint current observations (E, obs_stride), read at entry and
 * replaced by the final (post-reset) observations.  ep_return (optional, (E, n_agents)): += sum of rewards.
 * Two kernels serve the call.  simple_spread with 2-4 agents, num_units = 64, float32 OR float64 state: the actor MLP
 * (train.py:39-46) runs on the tcgen05 tensor cores -- weights resident in tensor memory, 3xTF32 split, fp32 accumulate;
 * actions agree with the fp32 SIMT kernels to ~5e-7 (csrc/mdp_rollout_tc.cu).  Everything else, and every core pinned
 * with mdp_core_set_tensor_cores(core, -1): fp32 SIMT actor tiles, bit-identical to the per-step kernels, float32 state only.
 * MDP_ENOTSUP when neither applies (callers then use the per-step kernels). */
int mdp_rollout_episode(mdp_env* env, mdp_core* core, int32_t E, void* state, float* obs, float* ring,
                        int64_t ring_capacity, int32_t ring_row_stride, int64_t ring_cursor, int32_t steps,
                        uint64_t seed, uint64_t counter, int32_t reset_after, uint64_t env_seed, uint64_t episode,
                        float* ep_return, void* stream);

/* `episodes` consecutive episodes of `steps` steps, each followed by env.reset() (train.py:110-133 run for
 * episodes * steps iterations with max_episode_len = steps): identical to `episodes` calls of mdp_rollout_episode with
 * reset_after = 1, counter + k*steps, ring_cursor + k*steps*E and episode + k -- the tcgen05 kernel loops over the episodes
 * inside ONE launch (weights stay in tensor memory, no per-episode prologue), other configurations launch per episode.
 * ring_capacity >= E * steps * episodes. */
int mdp_rollout_episodes(mdp_env* env, mdp_core* core, int32_t E, void* state, float* obs, float* ring,
                         int64_t ring_capacity, int32_t ring_row_stride, int64_t ring_cursor, int32_t steps, int32_t episodes,
                         uint64_t seed, uint64_t counter, uint64_t env_seed, uint64_t episode, float* ep_return, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* host-buffer loop body                                                                        */
/* ------------------------------------------------------------------------------------------ */
/* experiments/train.py:112-120 for E lockstep env instances with HOST input and output, one call:
 *     action_n = [agent.action(obs) for agent, obs in zip(trainers, obs_n)]          (:112)
 *     new_obs_n, rew_n, done_n, info_n = env.step(action_n)                           (:114)
 *     agent.experience(obs_n[i], action_n[i], rew_n[i], new_obs_n[i], done_n[i], ..)  (:119-120, all agents)
 * h_obs_in: HOST joint observations (E, obs_stride) -- copied to the caller's device staging d_obs_in;
 * then grouped actor inference + Gumbel sampling (Philox stream (seed, counter)), the fused env step and the
 * replay insert of the E joint rows at ring_cursor (ring may be NULL: no experience() calls); finally ONE
 * device->host copy of the packed result block d_out -> h_out, laid out as mdp_host_step_layout reports:
 *     offs4[0] next observations (E, obs_stride) f32 | offs4[1] rewards (E, n_agents) f32 |
 *     offs4[2] sampled actions (E, act_stride) f32   | offs4[3] done (E, n_agents) u8      (256-byte aligned blocks)
 * Both host buffers should be page-locked (the copies are then asynchronous DMA); the call only enqueues work
 * on `stream` -- h_out is valid once the stream has been synchronised. */
int mdp_host_step_layout(const mdp_env* env, int32_t E, int64_t* offs4, int64_t* total_bytes);
int mdp_host_step(mdp_env* env, mdp_core* core, int32_t E, void* state, const float* h_obs_in, float* d_obs_in,
                  void* d_out, void* h_out, float* ring, int64_t ring_capacity, int32_t ring_row_stride,
                  int64_t ring_cursor, uint64_t seed, uint64_t counter, void* stream);
/* Same call with the E env instances split into n_chunks equal ranges (E % n_chunks == 0, n_chunks <= 8), each on a
 * library-owned stream forked from / joined to `stream`: the host->device copy of range c+1 and the device->host copy
 * of range c-1 overlap the kernels of range c.  Same results bit for bit (the Philox streams are keyed by the env
 * index).  Capturable into a CUDA graph after one uncaptured call (which creates the streams and events). */
int mdp_host_step_pipelined(mdp_env* env, mdp_core* core, int32_t E, int32_t n_chunks, void* state, const float* h_obs_in,
                            float* d_obs_in, void* d_out, void* h_out, float* ring, int64_t ring_capacity,
                            int32_t ring_row_stride, int64_t ring_cursor, uint64_t seed, uint64_t counter, void* stream);

/* How mdp_host_step moves its two buffers: mode 0 = the copy engines (cudaMemcpyAsync), mode 1 = copy kernels (the SMs
 * read / write the page-locked host buffers through the unified address space; falls back to mode 0 when a buffer is
 * not device-accessible).  Same bytes either way; mode 1 avoids the copy engines' fixed per-transfer latency. */
int mdp_host_copy_mode(mdp_env* env, int32_t mode);

/* ------------------------------------------------------------------------------------------ */
/* device control block: lets a captured CUDA graph advance its own counters                    */
/* ------------------------------------------------------------------------------------------ */
/* ctl is a caller-owned DEVICE array of 4 uint64: {philox_counter, ring_cursor, episode, ring_length}.
 * Once attached, kernels ADD ctl[0] to the `counter` argument, ctl[2] to `episode`, and take the ring
 * cursor as (ring_cursor argument + ctl[1]) % capacity -- so the host can bake step-relative offsets
 * into a CUDA graph of a whole episode / update round and replay it.  Pass NULL to detach. */
int mdp_env_set_ctl(mdp_env* env, const uint64_t* ctl);
int mdp_core_set_ctl(mdp_core* core, const uint64_t* ctl);
/* ctl[0] += d_counter; ctl[1] = (ctl[1] + d_rows) % capacity; ctl[2] += d_episode;
 * ctl[3] = min(capacity, ctl[3] + d_rows)   (one tiny kernel, graph-capturable) */
int mdp_ctl_advance(uint64_t* ctl, uint64_t d_counter, int64_t d_rows, int64_t capacity, uint64_t d_episode,
                    void* stream);
/* ReplayBuffer.make_index on device (replay_buffer.py:46-47 draws B uniform indices in [0, len) with
 * replacement): Philox4x32-10 keyed by (seed, counter [+ctl[0]], b).  length <= 0 reads ctl[3]. */
int mdp_replay_make_index(int64_t* idx_out, int32_t B, int64_t length, uint64_t seed, uint64_t counter,
                          const uint64_t* ctl, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* prioritized replay: SumTree / PrioritizedReplayMemory on the device */
/* ------------------------------------------------------------------------------------------ */
/* Replaces /root/reference/maddpg/trainer/prioritized_replay_buffer.py: SumTree (:19-146) and the tree side of
 * PrioritizedReplayMemory.sample / batch_update (:171-201).  The tree is the reference's array -- 2^(k+1) - 1 float64 nodes,
 * k = ceil(log2(capacity)), the leaf of data slot d at index d + 2^k - 2 -- caller-owned, zero-initialised; `scratch` is a
 * caller-owned device array of `scratch_doubles` float64.  Rows are stored by mdp_replay_insert and read back by
 * mdp_replay_gather at the returned data indices.  Every float64 rounding follows the reference (sums of the children's deltas
 * in update_all, one delta per ancestor in batch order in update), so sampled indices are bit-exact given the same uniforms. */
int mdp_sumtree_layout(int64_t capacity, int64_t* tree_size, int32_t* k_out, int64_t* scratch_doubles);
/* SumTree.update_all (:58-100) for the adds still pending: data slots [start, start + count) (circular) take priority `value`
 * (SumTree.add (:45-56) defers its tree update to the next get_leaf; callers keep the pending range and pass it here or to
 * mdp_sumtree_sample). */
int mdp_sumtree_flush(double* tree, int64_t capacity, int64_t start, int64_t count, double value, double* scratch, void* stream);
/* PrioritizedReplayMemory.sample(n) (:171-194): reads total_p and the minimum over the last `capacity` tree entries, flushes the
 * pending adds (the first get_leaf does), then n stratified descents with v_i = a_i + (b_i - a_i) * uniforms[i] (numpy's
 * uniform(a, b)).  Outputs: tree index, data index, IS weight (prob / min_prob) ^ -beta per draw.  *flag |= 1 when a descent
 * ends on a data index >= capacity, where the reference raises IndexError (slot 0's node, see oracle/prioritized.py). */
int mdp_sumtree_sample(double* tree, int64_t capacity, int64_t dirty_start, int64_t dirty_count, double dirty_value, int32_t n,
                       const double* uniforms, double beta, int64_t* tree_idx_out, int64_t* data_idx_out, double* weights_out,
                       int32_t* flag, double* scratch, void* stream);
/* PrioritizedReplayMemory.batch_update(tree_idx, abs_errors) (:196-201): p = min(|err| + epsilon, abs_err_upper) ^ alpha (or the
 * caller's `priorities` when not NULL), then SumTree.update (:102-109) per element in batch order; B <= 4096.  *flag |= 2 when
 * the batch names a node that is not a true leaf (the one-thread reference loop ran instead of the parallel kernels). */
int mdp_sumtree_update(double* tree, int64_t capacity, const int64_t* tree_idx, int32_t B, const double* abs_errors,
                       const double* priorities, double epsilon, double abs_err_upper, double alpha, int32_t* flag,
                       double* scratch, void* stream);

/* cudaStreamSynchronize(stream): the host-side wait of the calls that fill HOST result buffers (mdp_host_step*). */
int mdp_stream_synchronize(void* stream);

const char* mdp_last_error(void);
const char* mdp_version(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
int64_t mdp_launch_count(void);

/* HOST helper for parity tests: the U[0,1) draws behind the kernels' Gumbel noise (SoftCategoricalPd.sample,
 * distributions.py:264-266, u = random_uniform): h_out[r * ncols + a] = Philox4x32-10 keyed by (seed, counter, agent tag,
 * row0 + r, a), exactly what mdp_actor_act / mdp_rollout_episode draw for population row row0 + r.  No device work. */
int mdp_philox_uniform(uint64_t seed, uint64_t counter, uint32_t tag, int64_t row0, int32_t nrows, int32_t ncols,
                       float* h_out);

#ifdef __cplusplus
}
#endif
#endif /* MADDPG_B200_H_ */
