// MADDPG trainer kernels: grouped actor inference + Gumbel-softmax, fused TD target, fused critic
// forward/backward, fused actor forward/backward through the running critic (sm_100a, fp32).
//
// Replaces, per SURVEY.md 8(a):
//   a2/a3  MADDPGAgentTrainer.action -> mlp_model -> SoftCategoricalPd.sample
//          (maddpg/trainer/maddpg.py:151-152,:62; experiments/train.py:39-46;
//           maddpg/common/distributions.py:264-266, 332-336)
//   a8     target computation                      maddpg/trainer/maddpg.py:181-187 (:70-71, :104,108)
//   a9     q_train                                 maddpg/trainer/maddpg.py:75-110
//   a10    p_train                                 maddpg/trainer/maddpg.py:28-73
//   a16    local_q_func (ddpg mode)                maddpg/trainer/maddpg.py:51-52, 86-87
//
// Design: every kernel owns a tile of TM batch rows per CTA and carries it through the WHOLE
// 3-layer MLP (and, for the update kernels, back again) without leaving shared memory: layer-1
// streams the (B, C) input and W1 through smem in K-chunks (C goes up to 3576), hidden activations
// live in smem, per-CTA weight-gradient partials are reduced into the flat gradient bucket with
// fp32 RED atomics.  The fp32 SIMT path is the parity path (1e-4 on Q/loss needs ~fp32 products);
// see DESIGN.md for the tensor-core plan.
#include "mdp_mlp.cuh"

#include <stdlib.h>

#include <algorithm>
#include <new>

namespace mdp {

// net buffer sizes (floats) of the resident-weights variants
__host__ __device__ inline int res_net_floats(int in, int U, int out) { return net_floats_padded(in, U, out); }

// ---------------------------------------------------------------------------------------------
// K1: grouped actor inference + Gumbel-softmax.  grid = (ceil(E/TM), agent_count)
// ---------------------------------------------------------------------------------------------
template <int U, int TM, bool RES>
__global__ void __launch_bounds__(NT) k_actor_act(CoreDev C, int agent_begin, int use_target, int E,
                                                  const float* __restrict__ obs, int obs_stride, float* __restrict__ act,
                                                  int act_stride, const float* __restrict__ u, uint64_t seed,
                                                  uint64_t counter, float* __restrict__ logits_out, int max_net,
                                                  long long rng_row_base) {
  if (C.ctl) counter += C.ctl[0];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sW = sm.take(RES ? max_net : KC * U);
  float* sX = sm.take(TM * XP);
  float* sH1 = sm.take(TM * (U + 4));
  float* sH2 = sm.take(TM * (U + 4));
  float* sL = sm.take(TM * KPAD);
  float* sA = sm.take(TM * KPAD);
  const int i = agent_begin + blockIdx.y;
  const AgentDev& ag = C.agents[i];
  MlpW w = ag.net[use_target ? MDP_NET_TARGET_P : MDP_NET_P];
  if (RES) w = load_net<U>(G, sW, w);  // visible after the first barrier inside layer1
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, E - row0);
  XSrc xs = make_xsrc(obs + ag.obs_off, obs_stride, ag.obs_dim);
  forward_hidden<U, TM, RES>(G, xs, w, row0, nrows, sX, sW, sH1, sH2);
  actor_head<U, TM>(G, sH2, w, sL);
  gumbel_softmax_tile<TM>(G, sL, sA, KPAD, nrows, ag.act_dim, ag.n_heads, ag.head_dim, u, act_stride, ag.act_off,
                          u ? row0 : row0 + rng_row_base, seed, counter, (uint32_t)i);
  for (int idx = threadIdx.x; idx < nrows * ag.act_dim; idx += NT) {
    const int r = idx / ag.act_dim, a = idx - r * ag.act_dim;
    act[(row0 + r) * act_stride + ag.act_off + a] = sA[r * KPAD + a];
    if (logits_out) logits_out[(row0 + r) * act_stride + ag.act_off + a] = sL[r * KPAD + a];
  }
}

// ---------------------------------------------------------------------------------------------
// q-values of one critic (debug surface q_debug[...]).  grid = ceil(B/TM)
// ---------------------------------------------------------------------------------------------
template <int U, int TM, bool RES>
__global__ void __launch_bounds__(NT) k_critic_q(CoreDev C, int agent, int use_target, int B, const float* __restrict__ x,
                                                 int x_stride, float* __restrict__ q_out, int max_net) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sW = sm.take(RES ? max_net : KC * U);
  float* sX = sm.take(TM * XP);
  float* sH1 = sm.take(TM * (U + 4));
  float* sH2 = sm.take(TM * (U + 4));
  float* sQ = sm.take(TM);
  const AgentDev& ag = C.agents[agent];
  MlpW w = ag.net[use_target ? MDP_NET_TARGET_Q : MDP_NET_Q];
  if (RES) w = load_net<U>(G, sW, w);
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  XSrc xs;
  if (ag.local_q) {
    xs = make_xsrc(x + ag.obs_off, x_stride, ag.obs_dim);
    xs.g1 = x + C.obs_sum + ag.act_off; xs.ld1 = x_stride; xs.n1 = ag.act_dim;
  } else {
    xs = make_xsrc(x, x_stride, C.obs_sum + C.act_sum);
  }
  forward_hidden<U, TM, RES>(G, xs, w, row0, nrows, sX, sW, sH1, sH2);
  critic_head<U, TM>(G, sH2, w, sQ);
  if (threadIdx.x < nrows) q_out[row0 + threadIdx.x] = sQ[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// K5: fused TD target of agent j.  grid = ceil(B/TM)
// ---------------------------------------------------------------------------------------------
template <int U, int TM, bool RES>
__global__ void __launch_bounds__(NT) k_td_target(CoreDev C, int j0, mdp_ring_layout L, int B, const float* __restrict__ batch,
                                                  const long long* __restrict__ ridx, const float* __restrict__ u_target,
                                                  int u_stride, uint64_t seed, uint64_t counter, float* __restrict__ y_out,
                                                  float* __restrict__ target_act_out, int max_net, long long idx_stride, long long y_stride) {
  const int j = j0 + blockIdx.y;  // grouped launch: one agent per grid.y slice
  if (ridx) ridx += blockIdx.y * idx_stride;
  y_out += blockIdx.y * y_stride;
  if (C.ctl) counter += C.ctl[0];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sW = sm.take(RES ? max_net : KC * U);
  float* sX = sm.take(TM * XP);
  float* sH1 = sm.take(TM * (U + 4));
  float* sH2 = sm.take(TM * (U + 4));
  float* sL = sm.take(TM * KPAD);
  float* sQ = sm.take(TM);
  const int ASP = C.act_stride | 1;
  float* sAct = sm.take(TM * ASP);
  const AgentDev& me = C.agents[j];
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  const int R = L.row_stride;

  // a'_i = gumbel_softmax(target_p_i(next_obs_i)) for every agent the critic sees
  const int i_begin = me.local_q ? j : 0, i_end = me.local_q ? j + 1 : C.n_agents;
  for (int i = i_begin; i < i_end; ++i) {
    const AgentDev& ag = C.agents[i];
    MlpW w = ag.net[MDP_NET_TARGET_P];
    if (RES) w = load_net<U>(G, sW, w);  // the previous user of sW finished before the last barrier
    XSrc xs = make_xsrc(batch + L.nx_off + ag.obs_off, R, ag.obs_dim);
    xs.idx = ridx;
    forward_hidden<U, TM, RES>(G, xs, w, row0, nrows, sX, sW, sH1, sH2);
    actor_head<U, TM>(G, sH2, w, sL);
    gumbel_softmax_tile<TM>(G, sL, sAct + ag.act_off, ASP, nrows, ag.act_dim, ag.n_heads, ag.head_dim, u_target, u_stride,
                            ag.act_off, row0, seed, counter, (uint32_t)(0x100 + i));
  }
  if (target_act_out) {
    for (int idx = threadIdx.x; idx < nrows * C.act_sum; idx += NT) {
      const int r = idx / C.act_sum, c = idx - r * C.act_sum;
      const bool mine = !me.local_q || (c >= me.act_off && c < me.act_off + me.act_dim);
      if (mine) target_act_out[(row0 + r) * u_stride + c] = sAct[r * ASP + c];
    }
  }
  // q' = target_q_j([next_obs | a'])
  XSrc xs;
  if (me.local_q) {
    xs = make_xsrc(batch + L.nx_off + me.obs_off, R, me.obs_dim);
    xs.s_over = sAct + me.act_off; xs.over_ld = ASP; xs.over_c0 = me.obs_dim; xs.over_n = me.act_dim;
  } else {
    xs = make_xsrc(batch + L.nx_off, R, C.obs_sum);
    xs.s_over = sAct; xs.over_ld = ASP; xs.over_c0 = C.obs_sum; xs.over_n = C.act_sum;
  }
  xs.idx = ridx;
  MlpW tq = me.net[MDP_NET_TARGET_Q];
  if (RES) tq = load_net<U>(G, sW, tq);
  forward_hidden<U, TM, RES>(G, xs, tq, row0, nrows, sX, sW, sH1, sH2);
  critic_head<U, TM>(G, sH2, tq, sQ);
  // y = float32(rew + gamma * (1 - done) * q')  -- float64 combine like numpy (maddpg.py:186)
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    double sy = 0, syy = 0, sr = 0, sq = 0;
    if (r < nrows) {
      const float* row = batch + (ridx ? ridx[row0 + r] : row0 + r) * R;
      const double rew = (double)row[L.rw_off + j], done = (double)row[L.dn_off + j];
      const float qn = sQ[r];
      const double y = rew + C.gamma * (1.0 - done) * (double)qn;
      y_out[row0 + r] = (float)y;
      sy = y; syy = y * y; sr = rew; sq = (double)qn;
    }
    for (int o = 16; o > 0; o >>= 1) {
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      syy += __shfl_xor_sync(0xffffffffu, syy, o);
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (r == 0) {
      double* st = C.stats + 8 * j;
      atomicAdd(st + 3, sy); atomicAdd(st + 4, syy); atomicAdd(st + 5, sr); atomicAdd(st + 6, sq);
      atomicAdd(st + 7, (double)nrows);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K6: fused critic forward + MSE + backward.  grid = ceil(B/TM)
// ---------------------------------------------------------------------------------------------
template <int U, int TM, bool RES>
__global__ void __launch_bounds__(NT) k_critic_grads(CoreDev C, int j0, mdp_ring_layout L, int B, const float* __restrict__ batch,
                                                     const long long* __restrict__ ridx, const float* __restrict__ y,
                                                     float* __restrict__ q_out, int max_net, long long idx_stride, long long y_stride) {
  const int j = j0 + blockIdx.y;  // grouped launch: one agent per grid.y slice
  if (ridx) ridx += blockIdx.y * idx_stride;
  y += blockIdx.y * y_stride;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int HP = U + 4;
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sW = sm.take(RES ? max_net : KC * U);
  float* sWT = sm.take(RES ? U * U : 4);
  float* sX = sm.take(TM * XP);
  float* sH1 = sm.take(TM * HP);
  float* sH2 = sm.take(TM * HP);
  float* sQ = sm.take(TM);
  float* sDq = sm.take(32);
  const AgentDev& me = C.agents[j];
  MlpW w = me.net[MDP_NET_Q];
  const MlpG& g = me.grad[1];
  if (RES) {
    load_wT_rows<U>(G, sWT, w.W2, 0, U);
    w = load_net<U>(G, sW, w);
  }
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  const int R = L.row_stride;
  if (blockIdx.x == 0 && threadIdx.x == 0) C.adam_t[2 * j + 1] += 1;  // one more Adam step for this net
  XSrc xs;
  if (me.local_q) {
    xs = make_xsrc(batch + me.obs_off, R, me.obs_dim);
    xs.g1 = batch + L.obs_sum + me.act_off; xs.ld1 = R; xs.n1 = me.act_dim;
  } else {
    xs = make_xsrc(batch, R, L.x_dim);
  }
  xs.idx = ridx;
  forward_hidden<U, TM, RES>(G, xs, w, row0, nrows, sX, sW, sH1, sH2);
  critic_head<U, TM>(G, sH2, w, sQ);
  // dL/dq = 2 (q - y) / B ; loss partial
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    float d = 0.f;
    double se = 0.0;
    if (r < nrows) {
      const float q = sQ[r];
      const float diff = q - y[row0 + r];
      d = 2.0f * diff / (float)B;
      se = (double)diff * (double)diff;
      if (q_out) q_out[row0 + r] = q;
    }
    sDq[r] = d;
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (r == 0) atomicAdd(C.stats + 8 * j + 0, se);
  }
  __syncthreads();
  // gW3[u] = sum_r h2[r][u] dq[r] ; gb3 = sum_r dq[r]
  if (threadIdx.x < U) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s = fmaf(sH2[r * HP + threadIdx.x], sDq[r], s);
    red_add(g.W3 + threadIdx.x, s);
  } else if (threadIdx.x == U) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s += sDq[r];
    red_add(g.b3, s);
  }
  __syncthreads();
  // dz2 = dq * W3^T * relu'(h2), in place over h2
  for (int idx = threadIdx.x; idx < TM * U; idx += NT) {
    const int r = idx / U, u = idx - r * U;
    const float h = sH2[r * HP + u];
    sH2[r * HP + u] = h > 0.f ? sDq[r] * w.W3[u] : 0.f;
  }
  __syncthreads();
  backward_hidden<U, TM, RES>(G, xs, w, sWT, &g, row0, nrows, sX, sW, sH1, sH2);
}

// ---------------------------------------------------------------------------------------------
// K7: fused actor forward -> Gumbel-softmax -> running critic forward -> backward to the action
// columns -> softmax Jacobian + logit regulariser -> actor backward.  grid = ceil(B/TM)
// ---------------------------------------------------------------------------------------------
template <int U, int TM, bool RES>
__global__ void __launch_bounds__(NT) k_actor_grads(CoreDev C, int j0, mdp_ring_layout L, int B, const float* __restrict__ batch,
                                                    const long long* __restrict__ ridx, const float* __restrict__ u_actor,
                                                    int u_stride, uint64_t seed, uint64_t counter, int max_net, long long idx_stride, long long y_stride) {
  const int j = j0 + blockIdx.y;  // grouped launch: one agent per grid.y slice
  if (ridx) ridx += blockIdx.y * idx_stride;
  if (C.ctl) counter += C.ctl[0];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int HP = U + 4;
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sW = sm.take(RES ? max_net : KC * U);   // RES: critic net ; else staging chunk
  float* sWp = sm.take(RES ? max_net : 4);       // RES: actor net
  float* sWT = sm.take(RES ? U * U : 4);         // RES: critic W2^T
  float* sWTp = sm.take(RES ? U * U : 4);        // RES: actor W2^T
  float* sX = sm.take(TM * XP);
  float* sH1 = sm.take(TM * HP);
  float* sH2 = sm.take(TM * HP);
  float* sP1 = sm.take(TM * HP);   // actor h1
  float* sP2 = sm.take(TM * HP);   // actor h2
  float* sL = sm.take(TM * KPAD);  // logits
  float* sA = sm.take(TM * KPAD);  // sampled action
  float* sDa = sm.take(TM * KPAD); // dQ/da then dL/dlogits
  float* sQ = sm.take(TM);
  const AgentDev& me = C.agents[j];
  MlpW pw = me.net[MDP_NET_P];
  MlpW qw = me.net[MDP_NET_Q];
  const MlpG& pg = me.grad[0];
  if (RES) {
    load_wT_rows<U>(G, sWT, qw.W2, 0, U);
    load_wT_rows<U>(G, sWTp, pw.W2, 0, U);
    qw = load_net<U>(G, sW, qw);
    pw = load_net<U>(G, sWp, pw);
  }
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  const int R = L.row_stride, K = me.act_dim;
  if (blockIdx.x == 0 && threadIdx.x == 0) C.adam_t[2 * j + 0] += 1;

  // actor forward on o_j, fresh Gumbel-softmax sample (maddpg.py:49)
  XSrc xp = make_xsrc(batch + me.obs_off, R, me.obs_dim);
  xp.idx = ridx;
  forward_hidden<U, TM, RES>(G, xp, pw, row0, nrows, sX, sW, sP1, sP2);
  actor_head<U, TM>(G, sP2, pw, sL);
  gumbel_softmax_tile<TM>(G, sL, sA, KPAD, nrows, K, me.n_heads, me.head_dim, u_actor, u_stride, me.act_off, row0, seed, counter,
                          (uint32_t)(0x200 + j));
  // running critic on [o, a_-j, a_hat_j]
  XSrc xq;
  int a_col0;
  if (me.local_q) {
    xq = make_xsrc(batch + me.obs_off, R, me.obs_dim);
    a_col0 = me.obs_dim;
  } else {
    xq = make_xsrc(batch, R, L.x_dim);
    a_col0 = L.obs_sum + me.act_off;
  }
  xq.s_over = sA; xq.over_ld = KPAD; xq.over_c0 = a_col0; xq.over_n = K;
  xq.idx = ridx;
  forward_hidden<U, TM, RES>(G, xq, qw, row0, nrows, sX, sW, sH1, sH2);
  critic_head<U, TM>(G, sH2, qw, sQ);
  // loss partials: sum(-q), sum(logits^2)
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    double sq = 0.0, sl = 0.0;
    if (r < nrows) {
      sq = -(double)sQ[r];
      for (int a = 0; a < K; ++a) sl += (double)sL[r * KPAD + a] * (double)sL[r * KPAD + a];
    }
    for (int o = 16; o > 0; o >>= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      sl += __shfl_xor_sync(0xffffffffu, sl, o);
    }
    if (r == 0) {
      atomicAdd(C.stats + 8 * j + 1, sq);
      atomicAdd(C.stats + 8 * j + 2, sl);
    }
  }
  // dz2 = (-1/B) * W3^T * relu'(h2) for valid rows
  const float dq = -1.0f / (float)B;
  for (int idx = threadIdx.x; idx < TM * U; idx += NT) {
    const int r = idx / U, u = idx - r * U;
    const float h = sH2[r * HP + u];
    sH2[r * HP + u] = (h > 0.f && r < nrows) ? dq * qw.W3[u] : 0.f;
  }
  __syncthreads();
  backward_hidden<U, TM, RES>(G, xq, qw, sWT, nullptr, row0, nrows, sX, sW, sH1, sH2);  // dz1 (critic) now in sH1
  // dQ/da[r][a] = dz1[r,:] . W1[a_col0 + a, :]
  for (int idx = threadIdx.x; idx < TM * K; idx += NT) {
    const int r = idx / K, a = idx - r * K;
    const float* w1row = qw.W1 + (size_t)(a_col0 + a) * U;
    float s = 0.f;
    for (int u = 0; u < U; ++u) s = fmaf(sH1[r * HP + u], w1row[u], s);
    sDa[r * KPAD + a] = s;
  }
  __syncthreads();
  // dL/dlogits = softmax Jacobian per head + 2 * reg * logits / (B * K)
  const float regc = (float)(2.0 * C.actor_reg / ((double)B * (double)K));
  for (int idx = threadIdx.x; idx < TM * me.n_heads; idx += NT) {
    const int r = idx / me.n_heads, h = idx - r * me.n_heads;
    const int o = h ? me.head_dim[0] : 0, n = me.head_dim[h];
    float dot = 0.f;
    for (int a = 0; a < n; ++a) dot = fmaf(sA[r * KPAD + o + a], sDa[r * KPAD + o + a], dot);
    for (int a = 0; a < n; ++a) {
      const float p = sA[r * KPAD + o + a];
      float dl = p * (sDa[r * KPAD + o + a] - dot) + regc * sL[r * KPAD + o + a];
      sDa[r * KPAD + o + a] = (r < nrows) ? dl : 0.f;
    }
  }
  __syncthreads();
  // actor head backward: gW3[u][a], gb3[a], dz2a = dl * W3^T * relu'(h2a)
  for (int idx = threadIdx.x; idx < U * K; idx += NT) {
    const int u = idx / K, a = idx - u * K;
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s = fmaf(sP2[r * HP + u], sDa[r * KPAD + a], s);
    red_add(pg.W3 + idx, s);
  }
  if (threadIdx.x < K) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s += sDa[r * KPAD + threadIdx.x];
    red_add(pg.b3 + threadIdx.x, s);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * U; idx += NT) {
    const int r = idx / U, u = idx - r * U;
    const float h = sP2[r * HP + u];
    float s = 0.f;
    if (h > 0.f)
      for (int a = 0; a < K; ++a) s = fmaf(sDa[r * KPAD + a], pw.W3[u * K + a], s);
    sP2[r * HP + u] = s;
  }
  __syncthreads();
  backward_hidden<U, TM, RES>(G, xp, pw, sWTp, &pg, row0, nrows, sX, sW, sP1, sP2);
}


// =============================================================================================
// Tile-resident variants (small configs, maddpg-mode critics): one prologue loads the X row tile (gathered
// through ridx straight from the replay ring) and every net the kernel touches into shared memory; the rest
// of the kernel never waits on global memory again.
// =============================================================================================
// q_train on one resident row tile (maddpg.py:75-100): forward, MSE gradient, backward; `ytile` = TD targets of the tile's rows
// (global or shared memory).  Runs on the 256 threads of group G.
template <int U, int TM>
__device__ __forceinline__ void critic_grads_body(const Grp& G, const CoreDev& C, int j, int B, const float* sXf, int XPf, const MlpW& w,
                                                  float* sWT, float* sH1, float* sH2, float* sQ, float* sDq,
                                                  const float* __restrict__ ytile, float* __restrict__ q_out_tile, int nrows,
                                                  int phase = 0) {
  // phase 0: the whole step; 1: forward pass only (q of the tile in sQ, activations in sH1 / sH2, W2^T in sWT); 2: loss and
  // backward pass from those -- the fused TD kernel runs phase 1 on an otherwise idle thread group next to the target critic
  constexpr int HP = U + 4;
  const MlpG& g = C.agents[j].grad[1];
  const int tid = G.tid;
  if (phase != 2) {
    build_wT_swz<U>(G, sWT, w.W2);  // published by the barriers inside forward_hidden_res
    forward_hidden_res<U, TM>(G, sXf, XPf, w, sH1, sH2);
    critic_head<U, TM>(G, sH2, w, sQ);
    if (phase == 1) return;
  }
  if (tid < 32) {
    const int r = tid;
    float d = 0.f;
    double se = 0.0;
    if (r < nrows && r < TM) {
      const float q = sQ[r];
      const float diff = q - ytile[r];
      d = 2.0f * diff / (float)B;
      se = (double)diff * (double)diff;
      if (q_out_tile) q_out_tile[r] = q;
    }
    sDq[r] = d;
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (r == 0) atomicAdd(C.stats + 8 * j + 0, se);
  }
  G.sync();
  if (tid < U) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s = fmaf(sH2[r * HP + tid], sDq[r], s);
    red_add(g.W3 + tid, s);
  } else if (tid == U) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s += sDq[r];
    red_add(g.b3, s);
  }
  G.sync();
  for (int idx = tid; idx < TM * U; idx += NT) {
    const int r = idx / U, u = idx - r * U;
    const float h = sH2[r * HP + u];
    sH2[r * HP + u] = h > 0.f ? sDq[r] * w.W3[u] : 0.f;
  }
  G.sync();
  backward_hidden_res<U, TM>(G, sXf, XPf, w, sWT, &g, sH1, sH2);
}

// FUSE: group 0 continues with the critic step (q_train) of the same rows -- y never leaves shared memory, the critic net and
// the [obs | act] columns of the rows ride in the same TMA prologue, and one launch + one prologue disappear from the serial
// chain of an agent update.
template <int U, int TM, bool FUSE>
__global__ void __launch_bounds__(3 * NT) k_td_target_res(CoreDev C, int j0, mdp_ring_layout L, int B, const float* __restrict__ batch,
                                                      const long long* __restrict__ ridx, const float* __restrict__ u_target,
                                                      int u_stride, uint64_t seed, uint64_t counter, float* __restrict__ y_out,
                                                      float* __restrict__ target_act_out, int XPf, long long idx_stride, long long y_stride) {
  pdl_launch_dependents();  // the optimizer launch that follows may become resident and prefetch its operands now
  const int j = j0 + blockIdx.y;  // grouped launch: one agent per grid.y slice
  if (ridx) ridx += blockIdx.y * idx_stride;
  y_out += blockIdx.y * y_stride;
  if (C.ctl) counter += C.ctl[0];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int HP = U + 4;
  // the target actors are independent of each other: NG = blockDim / 256 thread groups run one actor each, concurrently
  // (named barriers), then group 0 runs the target critic
  const int NG = blockDim.x / NT, grp = threadIdx.x / NT;
  const Grp G{(int)threadIdx.x % NT, NG > 1 ? grp + 1 : 0};
  SmemCarve sm(smem_raw);
  float* sXf = sm.take(TM * XPf);  // [next_obs_all | a'_all]
  float* sH1base = sm.take(NG * TM * HP);
  float* sH2base = sm.take(NG * TM * HP);
  float* sH1 = sH1base + grp * TM * HP;
  float* sH2 = sH2base + grp * TM * HP;
  float* sL = sm.take(NG * TM * KPAD) + grp * TM * KPAD;
  float* sQ = sm.take(TM);
  float* sRD = sm.take(2 * TM);    // rew_j, done_j
  const AgentDev& me = C.agents[j];
  const MlpW gq = me.net[MDP_NET_Q];
  const int x4 = (L.x_dim + 3) & ~3;
  float* sXc = FUSE ? sm.take(TM * XPf) : nullptr;  // [obs_all | act_all] of the same rows
  float* sY = FUSE ? sm.take(TM) : nullptr;
  float* sQc = FUSE ? sm.take(TM) : nullptr;  // q of the running critic (its forward pass runs next to the target critic's)
  float* sDq = FUSE ? sm.take(32) : nullptr;
  float* sWT = FUSE ? sm.take(U * U) : nullptr;
  float* sNetQ = FUSE ? sm.take(net_floats_padded(gq.in, U, 1)) : nullptr;
  float* sNets = sm.p;
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  const int R = L.row_stride, n = C.n_agents;
  if (FUSE && blockIdx.x == 0 && threadIdx.x == 0) C.adam_t[2 * j + 1] += 1;
  // prologue: ONE burst of TMA bulk copies (all target actors, the target critic, the gathered next_obs rows)
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  const int nx4 = (L.obs_sum + 3) & ~3;
  float* p = sNets;
  if (threadIdx.x == 0) {
    uint32_t total = (uint32_t)nrows * nx4 * 4u;
    if (FUSE) {
      total += (uint32_t)nrows * x4 * 4u + net_floats_padded(gq.in, U, 1) * 4u;
      bulk_g2s(sNetQ, gq.W1, net_floats_padded(gq.in, U, 1) * 4u, &bar);
    }
    for (int i = 0; i < n; ++i) total += net_floats_padded(C.agents[i].obs_dim, U, C.agents[i].act_dim) * 4u;
    total += net_floats_padded(me.net[MDP_NET_TARGET_Q].in, U, 1) * 4u;
    mbar_arrive_expect_tx(&bar, total);
    float* q = sNets;
    for (int i = 0; i < n; ++i) {
      const uint32_t nf = net_floats_padded(C.agents[i].obs_dim, U, C.agents[i].act_dim);
      bulk_g2s(q, C.agents[i].net[MDP_NET_TARGET_P].W1, nf * 4u, &bar);
      q += nf;
    }
    bulk_g2s(q, me.net[MDP_NET_TARGET_Q].W1, net_floats_padded(me.net[MDP_NET_TARGET_Q].in, U, 1) * 4u, &bar);
  }
  // Under a programmatic dependent launch the kernel before this one is mdp_update_prepare (index draw + statistics reset): the
  // nets above do not depend on it (every earlier kernel has completed), the sampled rows and the statistics do.
  pdl_wait();
  if (grp == 0) bulk_rows<TM>(G, sXf, XPf, batch, R, ridx, row0, nrows, L.nx_off, nx4, &bar);
  if (FUSE && grp == NG - 1) bulk_rows<TM>(G, sXc, XPf, batch, R, ridx, row0, nrows, 0, x4, &bar);
  for (int i = threadIdx.x; i < 2 * TM; i += blockDim.x) {
    const int r = i >> 1, which = i & 1;
    float v = 0.f;
    if (r < nrows) v = batch[(ridx ? ridx[row0 + r] : row0 + r) * R + (which ? L.dn_off : L.rw_off) + j];
    sRD[i] = v;
  }
  for (int i = 0; i < n; ++i) p += net_floats_padded(C.agents[i].obs_dim, U, C.agents[i].act_dim);
  const MlpW tq = net_at<U>(p, me.net[MDP_NET_TARGET_Q].in, 1);
  mbar_wait(&bar, 0);
  __syncthreads();
  // a'_i = gumbel_softmax(target_p_i(next_obs_i)) written straight into the critic-input columns of the tile
  p = sNets;
  for (int i = 0; i < n; ++i) {
    const AgentDev& ag = C.agents[i];
    const MlpW w = net_at<U>(p, ag.obs_dim, ag.act_dim);
    p += net_floats_padded(ag.obs_dim, U, ag.act_dim);
    if (i % NG != grp) continue;
    forward_hidden_res<U, TM>(G, sXf + ag.obs_off, XPf, w, sH1, sH2);
    actor_head<U, TM>(G, sH2, w, sL);
    gumbel_softmax_tile<TM>(G, sL, sXf + L.obs_sum + ag.act_off, XPf, nrows, ag.act_dim, ag.n_heads, ag.head_dim, u_target,
                            u_stride, ag.act_off, row0, seed, counter, (uint32_t)(0x100 + i));
  }
  if (NG > 1) __syncthreads();  // every group's a' is in the tile
  if (target_act_out) {
    for (int idx = threadIdx.x; idx < nrows * C.act_sum; idx += blockDim.x) {
      const int r = idx / C.act_sum, c = idx - r * C.act_sum;
      target_act_out[(row0 + r) * u_stride + c] = sXf[r * XPf + L.obs_sum + c];
    }
  }
  // FUSE with a second thread group: group 1 runs the running critic's forward pass on the replayed [obs | act] rows (it needs
  // neither a' nor y) while group 0 runs the target critic; a 512-thread barrier hands the activations over.
  const bool split = FUSE && NG > 1;
  if (split && grp == 1) {
    const MlpW wq = net_at<U>(sNetQ, gq.in, 1);
    critic_grads_body<U, TM>(G, C, j, B, sXc, XPf, wq, sWT, sH1, sH2, sQc, sDq, nullptr, nullptr, nrows, 1);
    asm volatile("bar.sync 8, 512;" ::: "memory");
    return;
  }
  if (grp != 0) return;
  forward_hidden_res<U, TM>(G, sXf, XPf, tq, sH1, sH2);
  critic_head<U, TM>(G, sH2, tq, sQ);
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    double sy = 0, syy = 0, sr = 0, sq = 0;
    if (r < nrows && r < TM) {
      const double rew = (double)sRD[2 * r], done = (double)sRD[2 * r + 1];
      const float qn = sQ[r];
      const double y = rew + C.gamma * (1.0 - done) * (double)qn;
      y_out[row0 + r] = (float)y;
      if (FUSE) sY[r] = (float)y;
      sy = y; syy = y * y; sr = rew; sq = (double)qn;
    }
    for (int o = 16; o > 0; o >>= 1) {
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      syy += __shfl_xor_sync(0xffffffffu, syy, o);
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (r == 0) {
      double* st = C.stats + 8 * j;
      atomicAdd(st + 3, sy); atomicAdd(st + 4, syy); atomicAdd(st + 5, sr); atomicAdd(st + 6, sq);
      atomicAdd(st + 7, (double)nrows);
    }
  }
  if (FUSE) {
    G.sync();  // y of the tile is in shared memory
    const MlpW wq = net_at<U>(sNetQ, gq.in, 1);
    if (split) {
      asm volatile("bar.sync 8, 512;" ::: "memory");  // group 1's forward pass is complete
      critic_grads_body<U, TM>(G, C, j, B, sXc, XPf, wq, sWT, sH1base + TM * HP, sH2base + TM * HP, sQc, sDq, sY, nullptr, nrows, 2);
    } else {
      critic_grads_body<U, TM>(G, C, j, B, sXc, XPf, wq, sWT, sH1, sH2, sQ, sDq, sY, nullptr, nrows);
    }
  }
}

template <int U, int TM>
__global__ void __launch_bounds__(NT) k_critic_grads_res(CoreDev C, int j0, mdp_ring_layout L, int B, const float* __restrict__ batch,
                                                         const long long* __restrict__ ridx, const float* __restrict__ y,
                                                         float* __restrict__ q_out, int XPf, long long idx_stride, long long y_stride) {
  pdl_launch_dependents();  // the optimizer launch that follows may become resident and prefetch its operands now
  const int j = j0 + blockIdx.y;  // grouped launch: one agent per grid.y slice
  if (ridx) ridx += blockIdx.y * idx_stride;
  y += blockIdx.y * y_stride;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int HP = U + 4;
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sXf = sm.take(TM * XPf);
  float* sH1 = sm.take(TM * HP);
  float* sH2 = sm.take(TM * HP);
  float* sQ = sm.take(TM);
  float* sDq = sm.take(32);
  float* sWT = sm.take(U * U);
  float* sNet = sm.p;
  const AgentDev& me = C.agents[j];
  const MlpG& g = me.grad[1];
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  if (blockIdx.x == 0 && threadIdx.x == 0) C.adam_t[2 * j + 1] += 1;
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  const int x4 = (L.x_dim + 3) & ~3;
  const MlpW gq = me.net[MDP_NET_Q];
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)nrows * x4 * 4u + net_floats_padded(gq.in, U, 1) * 4u);
    bulk_g2s(sNet, gq.W1, net_floats_padded(gq.in, U, 1) * 4u, &bar);
  }
  bulk_rows<TM>(G, sXf, XPf, batch, L.row_stride, ridx, row0, nrows, 0, x4, &bar);
  const MlpW w = net_at<U>(sNet, gq.in, 1);
  mbar_wait(&bar, 0);
  __syncthreads();
  critic_grads_body<U, TM>(G, C, j, B, sXf, XPf, w, sWT, sH1, sH2, sQ, sDq, y + row0, q_out ? q_out + row0 : nullptr, nrows);
}

template <int U, int TM>
__global__ void __launch_bounds__(NT) k_actor_grads_res(CoreDev C, int j0, mdp_ring_layout L, int B, const float* __restrict__ batch,
                                                        const long long* __restrict__ ridx, const float* __restrict__ u_actor,
                                                        int u_stride, uint64_t seed, uint64_t counter, int XPf, long long idx_stride, long long y_stride) {
  pdl_launch_dependents();  // the optimizer launch that follows may become resident and prefetch its operands now
  const int j = j0 + blockIdx.y;  // grouped launch: one agent per grid.y slice
  if (ridx) ridx += blockIdx.y * idx_stride;
  if (C.ctl) counter += C.ctl[0];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int HP = U + 4;
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sXf = sm.take(TM * XPf);
  float* sH1 = sm.take(TM * HP);
  float* sH2 = sm.take(TM * HP);
  float* sP1 = sm.take(TM * HP);
  float* sP2 = sm.take(TM * HP);
  float* sL = sm.take(TM * KPAD);
  float* sDa = sm.take(TM * KPAD);
  float* sQ = sm.take(TM);
  float* sWT = sm.take(U * U);
  float* sWTp = sm.take(U * U);
  float* sNets = sm.p;
  const AgentDev& me = C.agents[j];
  const MlpG& pg = me.grad[0];
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  const int K = me.act_dim;
  const int a_col0 = L.obs_sum + me.act_off;
  if (blockIdx.x == 0 && threadIdx.x == 0) C.adam_t[2 * j + 0] += 1;
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  const int x4 = (L.x_dim + 3) & ~3;
  const MlpW gq = me.net[MDP_NET_Q], gp = me.net[MDP_NET_P];
  const uint32_t nq = net_floats_padded(gq.in, U, 1), np_ = net_floats_padded(gp.in, U, gp.out);
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)nrows * x4 * 4u + (nq + np_) * 4u);
    bulk_g2s(sNets + nq, gp.W1, np_ * 4u, &bar);
  }
  bulk_rows<TM>(G, sXf, XPf, batch, L.row_stride, ridx, row0, nrows, 0, x4, &bar);
  // Under a programmatic dependent launch the kernel before this one is the critic's optimizer step: the actor net and the
  // sampled rows above do not depend on it, the critic's weights do.
  pdl_wait();
  if (threadIdx.x == 0) bulk_g2s(sNets, gq.W1, nq * 4u, &bar);
  const MlpW qw = net_at<U>(sNets, gq.in, 1);
  const MlpW pw = net_at<U>(sNets + nq, gp.in, gp.out);
  mbar_wait(&bar, 0);
  __syncthreads();
  build_wT_swz<U>(G, sWT, qw.W2);
  build_wT_swz<U>(G, sWTp, pw.W2);
  // actor forward on o_j; the fresh sample replaces the replayed action inside the tile (maddpg.py:49-50)
  forward_hidden_res<U, TM>(G, sXf + me.obs_off, XPf, pw, sP1, sP2);
  actor_head<U, TM>(G, sP2, pw, sL);
  gumbel_softmax_tile<TM>(G, sL, sXf + a_col0, XPf, nrows, K, me.n_heads, me.head_dim, u_actor, u_stride, me.act_off, row0, seed,
                          counter, (uint32_t)(0x200 + j));
  forward_hidden_res<U, TM>(G, sXf, XPf, qw, sH1, sH2);
  critic_head<U, TM>(G, sH2, qw, sQ);
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    double sq = 0.0, sl = 0.0;
    if (r < nrows && r < TM) {
      sq = -(double)sQ[r];
      for (int a = 0; a < K; ++a) sl += (double)sL[r * KPAD + a] * (double)sL[r * KPAD + a];
    }
    for (int o = 16; o > 0; o >>= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      sl += __shfl_xor_sync(0xffffffffu, sl, o);
    }
    if (r == 0) {
      atomicAdd(C.stats + 8 * j + 1, sq);
      atomicAdd(C.stats + 8 * j + 2, sl);
    }
  }
  const float dq = -1.0f / (float)B;
  for (int idx = threadIdx.x; idx < TM * U; idx += NT) {
    const int r = idx / U, u = idx - r * U;
    const float h = sH2[r * HP + u];
    sH2[r * HP + u] = (h > 0.f && r < nrows) ? dq * qw.W3[u] : 0.f;
  }
  __syncthreads();
  backward_hidden_res<U, TM>(G, sXf, XPf, qw, sWT, nullptr, sH1, sH2);  // dz1 (critic) now in sH1
  for (int idx = threadIdx.x; idx < TM * K; idx += NT) {
    const int r = idx / K, a = idx - r * K;
    const float* w1row = qw.W1 + (size_t)(a_col0 + a) * U;
    float s = 0.f;
    for (int u = 0; u < U; ++u) s = fmaf(sH1[r * HP + u], w1row[u], s);
    sDa[r * KPAD + a] = s;
  }
  __syncthreads();
  const float regc = (float)(2.0 * C.actor_reg / ((double)B * (double)K));
  for (int idx = threadIdx.x; idx < TM * me.n_heads; idx += NT) {
    const int r = idx / me.n_heads, h = idx - r * me.n_heads;
    const int o = h ? me.head_dim[0] : 0, nh = me.head_dim[h];
    const float* pa = sXf + r * XPf + a_col0;
    float dot = 0.f;
    for (int a = 0; a < nh; ++a) dot = fmaf(pa[o + a], sDa[r * KPAD + o + a], dot);
    for (int a = 0; a < nh; ++a) {
      const float dl = pa[o + a] * (sDa[r * KPAD + o + a] - dot) + regc * sL[r * KPAD + o + a];
      sDa[r * KPAD + o + a] = (r < nrows) ? dl : 0.f;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < U * K; idx += NT) {
    const int u = idx / K, a = idx - u * K;
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s = fmaf(sP2[r * HP + u], sDa[r * KPAD + a], s);
    red_add(pg.W3 + idx, s);
  }
  if (threadIdx.x < K) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s += sDa[r * KPAD + threadIdx.x];
    red_add(pg.b3 + threadIdx.x, s);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * U; idx += NT) {
    const int r = idx / U, u = idx - r * U;
    const float h = sP2[r * HP + u];
    float s = 0.f;
    if (h > 0.f)
      for (int a = 0; a < K; ++a) s = fmaf(sDa[r * KPAD + a], pw.W3[u * K + a], s);
    sP2[r * HP + u] = s;
  }
  __syncthreads();
  backward_hidden_res<U, TM>(G, sXf + me.obs_off, XPf, pw, sWTp, &pg, sP1, sP2);
}

// mdp_update_prepare: the index draw of an agent update (the same Philox stream as k_replay_make_index) and the reset of the
// agents' statistics accumulators in ONE launch, so that the TD-target kernel directly follows a kernel and can start as a
// programmatic dependent launch (a memset node in between would serialise the two).
__global__ void k_update_prepare(long long* __restrict__ idx_out, int B, long long length, uint64_t seed, uint64_t counter,
                                 const unsigned long long* __restrict__ ctl, double* __restrict__ stats, int n_stats) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_stats) stats[b] = 0.0;
  if (b < B) {
    if (ctl) counter += ctl[0];
    if (length <= 0) length = (long long)ctl[3];
    uint4 r = Philox::gen(seed, (uint32_t)b, 0x1D3Au, (uint32_t)counter, (uint32_t)(counter >> 32));
    long long i = (long long)(Philox::u01d(r.x, r.y) * (double)length);
    i = i < length ? i : length - 1;
    idx_out[b] = i > 0 ? i : 0;
  }
  // Launched as a programmatic dependent itself: nothing above reads what the kernel before it (the previous agent's optimizer
  // step) writes, so the draw runs under that kernel.  The TD-target kernel behind this one may only start staging its nets
  // once that optimizer step HAS completed: wait first, then release the dependents.
  pdl_wait();
  pdl_launch_dependents();
}

}  // namespace mdp

// =============================================================================================
// host side
// =============================================================================================
namespace mdp {

static inline int64_t net_floats(int in, int U, int out) { return (int64_t)in * U + U + (int64_t)U * U + U + (int64_t)U * out + out; }

static MlpW make_w(float* base, int in, int U, int out) {
  MlpW w;
  float* p = base;
  w.W1 = p; p += (size_t)in * U;
  w.b1 = p; p += U;
  w.W2 = p; p += (size_t)U * U;
  w.b2 = p; p += U;
  w.W3 = p; p += (size_t)U * out;
  w.b3 = p;
  w.in = in; w.out = out;
  return w;
}
static MlpG make_g(float* base, int in, int U, int out) {
  MlpW w = make_w(base, in, U, out);
  MlpG g;
  g.W1 = const_cast<float*>(w.W1); g.b1 = const_cast<float*>(w.b1); g.W2 = const_cast<float*>(w.W2);
  g.b2 = const_cast<float*>(w.b2); g.W3 = const_cast<float*>(w.W3); g.b3 = const_cast<float*>(w.b3);
  return g;
}

static CoreDev core_dev(const mdp_core* c) {
  CoreDev d;
  d.agents = c->d_agents;
  d.n_agents = c->cfg.n_agents;
  d.units = c->cfg.num_units;
  d.obs_sum = c->obs_sum; d.act_sum = c->act_sum; d.act_stride = c->act_stride;
  d.gamma = c->cfg.gamma; d.actor_reg = c->cfg.actor_reg;
  d.adam_t = c->adam_t; d.stats = c->stats; d.ctl = c->ctl;
  return d;
}

CoreDev core_dev_for_rollout(const mdp_core* c) { return core_dev(c); }

template <typename Kern>
static int set_smem(Kern kern, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      cudaFuncAttributes fa;
      memset(&fa, 0, sizeof(fa));
      cudaError_t e2 = cudaFuncGetAttributes(&fa, kern);
      cudaGetLastError();
      return fail(MDP_ECUDA, "cudaFuncSetAttribute(max dynamic shared memory = %zu): %s (kernel: %zu B static shared, %d registers, "
                  "attributes query: %s)", smem, cudaGetErrorString(e), fa.sharedSizeBytes, fa.numRegs, cudaGetErrorString(e2));
    }
  }
  return MDP_OK;
}

template <int V> struct IC { static constexpr int value = V; };
template <bool V> struct BC { static constexpr bool value = V; };

// Launch plan of the fused MLP kernels: tile rows (16 doubles the CTA count for the reference's batch of
// 1024 on 148 SMs), and whether every net a kernel touches fits in shared memory (resident variant).
struct Plan {
  int TM;
  bool res;
  int max_net;  // floats of the largest net (padded), for the resident buffers
};

static Plan make_plan(const mdp_core* c, int rows) {
  Plan p;
  p.TM = rows <= 2048 ? 16 : 32;
  int mx = 0;
  for (int i = 0; i < c->cfg.n_agents; ++i)
    for (int k = 0; k < 4; ++k) {
      const int n = net_floats_padded(c->lay.net_in[i][k], c->cfg.num_units, c->lay.net_out[i][k]);
      if (n > mx) mx = n;
    }
  p.max_net = mx;
  // resident budget: two nets + two transposed W2 + activations must stay under ~160 KB (k_actor_grads)
  p.res = c->cfg.num_units == 64 && (size_t)(2 * mx + 2 * 64 * 64) * 4 <= 120 * 1024;
  return p;
}

template <typename F>
static int dispatch(int U, const Plan& p, F&& f) {
  if (U == 64) {
    if (p.TM == 16) return p.res ? f(IC<64>{}, IC<16>{}, BC<true>{}) : f(IC<64>{}, IC<16>{}, BC<false>{});
    return p.res ? f(IC<64>{}, IC<32>{}, BC<true>{}) : f(IC<64>{}, IC<32>{}, BC<false>{});
  }
  if (p.TM == 16) return f(IC<128>{}, IC<16>{}, BC<false>{});
  return f(IC<128>{}, IC<32>{}, BC<false>{});
}

// floats of dynamic shared memory: weight buffers + x chunk + n_act activation tiles + extras
static size_t smem_for(int U, const Plan& p, int n_wbuf, int n_wT, int n_act, int extra_floats) {
  const int TMv = p.TM;
  size_t f = 0;
  f += (size_t)(p.res ? p.max_net : KC * U) + (size_t)(n_wbuf - 1) * (p.res ? p.max_net : 4);
  f += (size_t)n_wT * (p.res ? U * U : 4);
  f += (size_t)TMv * XP + (size_t)n_act * TMv * (U + 4) + extra_floats + 128;
  return f * sizeof(float);
}

static int check_lay(const mdp_core* c, const mdp_ring_layout* lay) {
  MDP_REQUIRE(lay, "null ring layout");
  MDP_REQUIRE(lay->n_agents == c->cfg.n_agents && lay->obs_sum == c->obs_sum && lay->act_sum == c->act_sum,
              "ring layout does not match the core (agents %d/%d, obs %d/%d, act %d/%d)", lay->n_agents, c->cfg.n_agents,
              lay->obs_sum, c->obs_sum, lay->act_sum, c->act_sum);
  return MDP_OK;
}

}  // namespace mdp

using namespace mdp;

extern "C" int mdp_core_create(const mdp_core_cfg* cfg, mdp_core** out) {
  MDP_REQUIRE(cfg && out, "mdp_core_create: null argument");
  MDP_REQUIRE(cfg->n_agents > 0 && cfg->n_agents <= MDP_MAX_AGENTS, "mdp_core_create: n_agents %d out of range", cfg->n_agents);
  MDP_REQUIRE(cfg->num_units == 64 || cfg->num_units == 128, "mdp_core_create: num_units must be 64 or 128 (got %d)", cfg->num_units);
  mdp_core* c = new (std::nothrow) mdp_core();
  MDP_REQUIRE(c, "mdp_core_create: out of memory");
  c->cfg = *cfg;
  memset(&c->lay, 0, sizeof(c->lay));
  const int n = cfg->n_agents, U = cfg->num_units;
  int od = 0, ad = 0;
  for (int i = 0; i < n; ++i) {
    int hsum = 0;
    if (!(cfg->n_heads[i] >= 1 && cfg->n_heads[i] <= MDP_MAX_HEADS)) { delete c; return fail(MDP_EINVAL, "agent %d: n_heads %d", i, cfg->n_heads[i]); }
    for (int h = 0; h < cfg->n_heads[i]; ++h) hsum += cfg->head_dim[i][h];
    if (hsum != cfg->act_dim[i] || cfg->act_dim[i] > MAXK || cfg->act_dim[i] < 1 || cfg->obs_dim[i] < 1) {
      delete c;
      return fail(MDP_EINVAL, "agent %d: act_dim %d (heads sum %d, max %d), obs_dim %d", i, cfg->act_dim[i], hsum, MAXK, cfg->obs_dim[i]);
    }
    c->obs_off[i] = od; c->act_off[i] = ad;
    od += cfg->obs_dim[i]; ad += cfg->act_dim[i];
  }
  c->obs_sum = od; c->act_sum = ad; c->act_stride = round_up(ad, 4);
  int64_t po = 0, to = 0;
  for (int i = 0; i < n; ++i) {
    const int q_in = cfg->local_q[i] ? cfg->obs_dim[i] + cfg->act_dim[i] : od + ad;
    const int ins[4] = {cfg->obs_dim[i], cfg->obs_dim[i], q_in, q_in};
    const int outs[4] = {cfg->act_dim[i], cfg->act_dim[i], 1, 1};
    for (int k = 0; k < 4; ++k) {
      c->lay.net_off[i][k] = po;
      c->lay.net_in[i][k] = ins[k];
      c->lay.net_out[i][k] = outs[k];
      c->lay.net_size[i][k] = net_floats(ins[k], U, outs[k]);
      po += round_up64(c->lay.net_size[i][k], 4);
    }
    c->lay.train_off[i][0] = to; to += round_up64(c->lay.net_size[i][MDP_NET_P], 4);
    c->lay.train_off[i][1] = to; to += round_up64(c->lay.net_size[i][MDP_NET_Q], 4);
  }
  c->lay.total_params = po;
  c->lay.total_train = to;
  // SURVEY 8(d) FLOP model per batch row: Fq = C*U + U^2 + U, Fpi_i = D_i*U + U^2 + U*K_i
  for (int i = 0; i < n; ++i) {
    const int64_t Cq = c->lay.net_in[i][MDP_NET_Q];
    const int64_t Fq = Cq * U + (int64_t)U * U + U;
    int64_t Fpi_all = 0;
    for (int k = 0; k < n; ++k)
      if (!cfg->local_q[i] || k == i) Fpi_all += (int64_t)cfg->obs_dim[k] * U + (int64_t)U * U + (int64_t)U * cfg->act_dim[k];
    const int64_t Fpj = (int64_t)cfg->obs_dim[i] * U + (int64_t)U * U + (int64_t)U * cfg->act_dim[i];
    c->lay.update_flops_critic[i] = 2 * (3 * Fq + ((int64_t)U * U + U) + Fpi_all);
    c->lay.update_flops_actor[i] = 2 * (3 * Fpj - (int64_t)cfg->obs_dim[i] * U + Fq + (U + (int64_t)U * U + (int64_t)cfg->act_dim[i] * U));
  }
  *out = c;
  return MDP_OK;
}

extern "C" int mdp_core_get_layout(const mdp_core* core, mdp_core_layout* out) {
  MDP_REQUIRE(core && out, "mdp_core_get_layout: null argument");
  *out = core->lay;
  return MDP_OK;
}

extern "C" void mdp_core_destroy(mdp_core* core) {
  if (!core) return;
  if (core->d_agents) cudaFree(core->d_agents);
  if (core->tc_scratch) cudaFree(core->tc_scratch);
  if (core->tc_arena) cudaFree(core->tc_arena);
  if (core->tc_dz1) cudaFree(core->tc_dz1);
  if (core->norm2) cudaFree(core->norm2);
  if (core->peer_reduced) cudaFree(core->peer_reduced);
  if (core->d_peer_tables) cudaFree(core->d_peer_tables);
  delete core;
}

extern "C" int mdp_core_set_ctl(mdp_core* c, const uint64_t* ctl) {
  MDP_REQUIRE(c, "mdp_core_set_ctl: null core");
  c->ctl = reinterpret_cast<const unsigned long long*>(ctl);
  return MDP_OK;
}

extern "C" int mdp_core_set_tensor_cores(mdp_core* c, int32_t mode) {
  MDP_REQUIRE(c && mode >= -1 && mode <= 1, "mdp_core_set_tensor_cores: mode must be -1, 0 or 1");
  c->tc_mode = mode;
  return MDP_OK;
}

extern "C" int mdp_core_set_fused_update(mdp_core* c, int32_t on) {
  MDP_REQUIRE(c, "mdp_core_set_fused_update: null core");
  c->no_fuse = on ? 0 : 1;
  return MDP_OK;
}

extern "C" int mdp_core_bind(mdp_core* c, float* params, float* grads, float* adam_m, float* adam_v, int32_t* adam_t,
                             double* stats) {
  MDP_REQUIRE(c && params && grads && adam_m && adam_v && adam_t && stats, "mdp_core_bind: null argument");
  MDP_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)adam_m | (uintptr_t)adam_v) & 15) == 0,
              "mdp_core_bind: buffers must be 16-byte aligned");
  c->params = params; c->grads = grads; c->adam_m = adam_m; c->adam_v = adam_v; c->adam_t = adam_t; c->stats = stats;
  const int n = c->cfg.n_agents, U = c->cfg.num_units;
  c->h_agents.assign(n, AgentDev());
  for (int i = 0; i < n; ++i) {
    AgentDev& a = c->h_agents[i];
    for (int k = 0; k < 4; ++k) a.net[k] = make_w(params + c->lay.net_off[i][k], c->lay.net_in[i][k], U, c->lay.net_out[i][k]);
    a.grad[0] = make_g(grads + c->lay.train_off[i][0], c->lay.net_in[i][MDP_NET_P], U, c->lay.net_out[i][MDP_NET_P]);
    a.grad[1] = make_g(grads + c->lay.train_off[i][1], c->lay.net_in[i][MDP_NET_Q], U, c->lay.net_out[i][MDP_NET_Q]);
    a.obs_dim = c->cfg.obs_dim[i]; a.act_dim = c->cfg.act_dim[i];
    a.obs_off = c->obs_off[i]; a.act_off = c->act_off[i];
    a.n_heads = c->cfg.n_heads[i]; a.head_dim[0] = c->cfg.head_dim[i][0]; a.head_dim[1] = c->cfg.head_dim[i][1];
    a.local_q = c->cfg.local_q[i]; a.q_in = c->lay.net_in[i][MDP_NET_Q];
  }
  if (c->d_agents) cudaFree(c->d_agents);
  c->d_agents = nullptr;
  MDP_CUDA(cudaMalloc(&c->d_agents, n * sizeof(AgentDev)));
  MDP_CUDA(cudaMemcpy(c->d_agents, c->h_agents.data(), n * sizeof(AgentDev), cudaMemcpyHostToDevice));
  return MDP_OK;
}


namespace mdp {
int launch_actor_act_tc(mdp_core* c, const CoreDev& d, int32_t agent_begin, int32_t agent_count, int32_t use_target, int32_t E,
                        const float* obs, int32_t obs_stride, float* act, int32_t act_stride, const float* u, uint64_t seed,
                        uint64_t counter, float* logits_out, long long row_base, cudaStream_t st);
}

extern "C" int mdp_actor_act(mdp_core* c, int32_t agent_begin, int32_t agent_count, int32_t use_target, int32_t E,
                             const float* obs, int32_t obs_stride, float* act, int32_t act_stride, const float* u,
                             uint64_t seed, uint64_t counter, float* logits_out, void* stream) {
  return mdp::actor_act_range(c, agent_begin, agent_count, use_target, E, obs, obs_stride, act, act_stride, u, seed, counter,
                              logits_out, 0, stream);
}

int mdp::actor_act_range(mdp_core* c, int32_t agent_begin, int32_t agent_count, int32_t use_target, int32_t E, const float* obs,
                         int32_t obs_stride, float* act, int32_t act_stride, const float* u, uint64_t seed, uint64_t counter,
                         float* logits_out, int64_t row_base, void* stream) {
  MDP_REQUIRE(c && c->d_agents, "mdp_actor_act: core not bound");
  MDP_REQUIRE(obs && act && E > 0 && agent_count > 0 && agent_begin >= 0 && agent_begin + agent_count <= c->cfg.n_agents,
              "mdp_actor_act: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CoreDev d = core_dev(c);
  // tensor-core kernel (128-row tiles) once a launch has at least one tile per SM (tc_mode 1: always; -1: never)
  if (c->cfg.num_units == 64 && c->tc_mode >= 0 && (c->tc_mode > 0 || (long long)cdiv(E, 128) * agent_count >= 148)) {
    int rc = launch_actor_act_tc(c, d, agent_begin, agent_count, use_target, E, obs, obs_stride, act, act_stride, u, seed, counter,
                                 logits_out, (long long)row_base, st);
    if (rc != MDP_ENOTSUP) return rc;
  }
  const Plan p = make_plan(c, E >= 4096 ? 4096 : E);
  return dispatch(c->cfg.num_units, p, [&](auto u_, auto tm_, auto res_) -> int {
    constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
    constexpr bool RES = decltype(res_)::value;
    auto kern = k_actor_act<U, TMv, RES>;
    const size_t smem = smem_for(U, p, 1, 0, 2, 2 * TMv * KPAD);
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    kern<<<dim3(cdiv(E, TMv), agent_count), NT, smem, st>>>(d, agent_begin, use_target, E, obs, obs_stride, act, act_stride, u,
                                                          seed, counter, logits_out, p.max_net, (long long)row_base);
    return check_launch("k_actor_act");
  });
}

extern "C" int mdp_critic_q(mdp_core* c, int32_t agent, int32_t use_target, int32_t B, const float* x, int32_t x_stride,
                            float* q_out, void* stream) {
  MDP_REQUIRE(c && c->d_agents, "mdp_critic_q: core not bound");
  MDP_REQUIRE(x && q_out && B > 0 && agent >= 0 && agent < c->cfg.n_agents, "mdp_critic_q: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CoreDev d = core_dev(c);
  const Plan p = make_plan(c, B);
  return dispatch(c->cfg.num_units, p, [&](auto u_, auto tm_, auto res_) -> int {
    constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
    constexpr bool RES = decltype(res_)::value;
    auto kern = k_critic_q<U, TMv, RES>;
    const size_t smem = smem_for(U, p, 1, 0, 2, TMv);
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    kern<<<cdiv(B, TMv), NT, smem, st>>>(d, agent, use_target, B, x, x_stride, q_out, p.max_net);
    return check_launch("k_critic_q");
  });
}

// tile-resident plan: every net a kernel touches + the X row tile fit one CTA's shared memory, and no agent
// uses a local critic (those keep the streaming kernels)
struct ResPlan {
  bool ok;
  int XPf;                      // pitch of the X row tile
  size_t td, critic, actor;     // dynamic shared memory (bytes) of the three kernels
  size_t td_fused;              // TD target with the critic step fused in
  bool fuse_ok;
  int td_groups;
};

static ResPlan make_res_plan(const mdp_core* c, const Plan& p, int agent) {
  ResPlan r;
  memset(&r, 0, sizeof(r));  // the early return below must leave fuse_ok false too (it was read uninitialised for local critics)
  r.ok = false;
  const int U = c->cfg.num_units, n = c->cfg.n_agents, TMv = p.TM, HP = U + 4;
  for (int i = 0; i < n; ++i)
    if (c->cfg.local_q[i]) return r;
  r.XPf = round_up(c->obs_sum + c->act_sum, 4) + 4;
  auto r4 = [](size_t x) { return (x + 3) & ~(size_t)3; };
  size_t actors = 0;
  for (int i = 0; i < n; ++i) actors += net_floats_padded(c->cfg.obs_dim[i], U, c->cfg.act_dim[i]);
  (void)agent;  // sized for the largest agent so that grouped launches (grid.y = agent) share one plan
  size_t crit = 0, act_j = 0;
  for (int i = 0; i < n; ++i) {
    crit = std::max(crit, (size_t)net_floats_padded(c->lay.net_in[i][MDP_NET_Q], U, 1));
    act_j = std::max(act_j, (size_t)net_floats_padded(c->cfg.obs_dim[i], U, c->cfg.act_dim[i]));
  }
  const size_t tile = r4((size_t)TMv * r.XPf);
  r.td_groups = std::min(n, 3);  // concurrent target-actor groups of 256 threads (k_td_target_res)
  r.td = (tile + r.td_groups * (2 * r4((size_t)TMv * HP) + r4(TMv * KPAD)) + r4(TMv) + r4(2 * TMv) + actors + crit + 64) * 4;
  r.critic = (tile + 2 * r4((size_t)TMv * HP) + r4(TMv) + 32 + (size_t)U * U + crit + 64) * 4;
  r.actor = (tile + 4 * r4((size_t)TMv * HP) + 2 * r4(TMv * KPAD) + r4(TMv) + 2 * (size_t)U * U + crit + act_j + 64) * 4;
  r.td_fused = r.td + (tile + 2 * r4(TMv) + 32 + (size_t)U * U + crit) * 4;  // k_td_target_res<FUSE>: + critic tile, y, q, dq, W2^T, q net
  const size_t limit = 200 * 1024;
  r.ok = r.td <= limit && r.critic <= limit && r.actor <= limit;
  r.fuse_ok = r.ok && r.td_fused <= limit;
  return r;
}

// Plan of a launch over `count` agents (grid.y): a grouped round of 16-row tiles is cdiv(rows, 16) * count CTAs, one resident CTA
// per SM -- past 148 they run as two waves, and 32-row tiles (half the CTAs, two rows per thread against every weight load) finish
// sooner: batch 1024 x 3 agents 0.088 -> 0.069 ms per round, simple_tag 0.107 -> 0.077, simple_world_comm with 128 units 0.71 ->
// 0.55 (tools/time_upd_scen.py).  Not when the taller tile would push the nets out of shared memory (simple_spread N=4: the
// resident plan fits at 16 rows only, 0.117 vs 0.151 ms), and never for a single agent's launch (64 CTAs: one wave either way).
static Plan plan_for(const mdp_core* c, int rows, int count) {
  Plan p = make_plan(c, rows);
  if (count > 1 && p.TM == 16 && (long long)cdiv(rows, 16) * count > 148) {
    Plan q = p;
    q.TM = 32;
    const ResPlan r16 = make_res_plan(c, p, 0), r32 = make_res_plan(c, q, 0);
    if (!r16.ok || r32.ok) p = q;  // also when the fused TD-target + critic tile no longer fits: two launches of 32-row tiles still win
  }
  static const int force_tm = []() { const char* e = getenv("MDP_PLAN_TM"); return e ? atoi(e) : 0; }();
  if (force_tm == 16 || force_tm == 32) p.TM = force_tm;  // diagnostic override (tools/time_upd_scen.py)
  return p;
}

namespace mdp {
int launch_td_target_tc(mdp_core* c, const CoreDev& d, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B,
                        const float* batch, const long long* ridx, long long idx_stride, const float* u_target, int32_t u_stride,
                        uint64_t seed, uint64_t counter, float* y_out, long long y_stride, float* target_act_out, cudaStream_t st);
}

namespace mdp {
int launch_critic_grads_tc(mdp_core* c, const CoreDev& d, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B,
                           const float* batch, const long long* ridx, long long idx_stride, const float* y, long long y_stride,
                           float* q_out, cudaStream_t st);
}

namespace mdp {
int launch_actor_grads_tc(mdp_core* c, const CoreDev& d, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B,
                          const float* batch, const long long* ridx, long long idx_stride, const float* u_actor, int32_t u_stride,
                          uint64_t seed, uint64_t counter, cudaStream_t st);
}

// tensor-core path policy: forced on (1), forced off (-1), or automatic (0)
static bool want_tc(const mdp_core* c, int B, int count, bool backward = false) {
  if (c->tc_mode < 0 || c->cfg.num_units != 64) return false;
  if (c->tc_mode > 0) return true;
  // automatic: the 128-row tensor-core tiles pay off once a launch fills most SMs with them, or when the critic input
  // is wide enough that streaming it through the UMMA pipeline beats the SIMT K-loop (measured on B200)
  (void)backward;
  int max_in = 0;
  for (int i = 0; i < c->cfg.n_agents; ++i) max_in = std::max(max_in, c->lay.net_in[i][MDP_NET_TARGET_Q]);
  return cdiv(B, 128) * count >= 96 || max_in >= 1024;
}

// mdp_update_prepare ran as the launch right before for exactly these agents: their statistics are already zero and the
// TD-target kernel may start as its programmatic dependent.  Consumes the mark.
static bool take_prepared(mdp_core* c, int32_t agent, int32_t count) {
  const bool ok = c->prep_agent == agent && c->prep_count == count;
  c->prep_agent = -1;
  c->prep_count = 0;
  return ok;
}

template <typename Kern, typename... Args>
static int launch_maybe_pdl(const char* what, bool pdl, Kern kern, dim3 grid, int block, size_t smem, cudaStream_t st, Args... args) {
  static const bool pdl_enabled = []() { const char* e = getenv("MDP_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t lc;
  cudaLaunchAttribute at[1];
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = grid; lc.blockDim = dim3(block); lc.dynamicSmemBytes = smem; lc.stream = st;
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = (pdl && pdl_enabled) ? 1 : 0;
  MDP_CUDA(cudaLaunchKernelEx(&lc, kern, args...));
  return check_launch(what);
}

static int launch_td_target(mdp_core* c, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B, const float* batch,
                            const int64_t* idx, long long idx_stride, const float* u_target, int32_t u_stride, uint64_t seed,
                            uint64_t counter, float* y_out, long long y_stride, float* target_act_out, void* stream) {
  MDP_REQUIRE(c && c->d_agents, "mdp_td_target: core not bound");
  int rc = check_lay(c, lay);
  if (rc) return rc;
  MDP_REQUIRE(batch && y_out && B > 0 && agent >= 0 && count > 0 && agent + count <= c->cfg.n_agents, "mdp_td_target: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const bool prepared = take_prepared(c, agent, count);
  if (!prepared) MDP_CUDA(cudaMemsetAsync(c->stats + 8 * agent, 0, 8 * sizeof(double) * count, st));
  CoreDev d = core_dev(c);
  const Plan p = plan_for(c, B, count);
  const ResPlan rp = make_res_plan(c, p, agent);
  const long long* ridx = (const long long*)idx;
  if (want_tc(c, B, count)) {
    rc = launch_td_target_tc(c, d, agent, count, lay, B, batch, ridx, idx_stride, u_target, u_stride, seed, counter, y_out, y_stride,
                             target_act_out, st);
    if (rc != MDP_ENOTSUP) return rc;
  }
  return dispatch(c->cfg.num_units, p, [&](auto u_, auto tm_, auto res_) -> int {
    constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
    constexpr bool RES = decltype(res_)::value;
    if (rp.ok) {
      auto kern = k_td_target_res<U, TMv, false>;
      int rc2 = set_smem(kern, rp.td);
      if (rc2) return rc2;
      return launch_maybe_pdl("k_td_target_res", prepared, kern, dim3(cdiv(B, TMv), count), rp.td_groups * NT, rp.td, st, d, (int)agent, *lay,
                              (int)B, batch, ridx, u_target, (int)u_stride, (unsigned long)seed, (unsigned long)counter, y_out,
                              target_act_out, (int)rp.XPf, (long long)idx_stride, (long long)y_stride);
    }
    auto kern = k_td_target<U, TMv, RES>;
    const size_t smem = smem_for(U, p, 1, 0, 2, TMv * KPAD + TMv + TMv * (c->act_stride | 1));
    int rc2 = set_smem(kern, smem);
    if (rc2) return rc2;
    kern<<<dim3(cdiv(B, TMv), count), NT, smem, st>>>(d, agent, *lay, B, batch, ridx, u_target, u_stride, seed, counter, y_out,
                                                     target_act_out, p.max_net, idx_stride, y_stride);
    return check_launch("k_td_target");
  });
}

static int launch_critic_grads(mdp_core* c, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B, const float* batch,
                               const int64_t* idx, long long idx_stride, const float* y, long long y_stride, float* q_out,
                               void* stream);

// TD target + critic gradients of the same rows in ONE launch (k_td_target_res<FUSE>) when the fused tile fits in shared memory
// and the SIMT path is the one selected; otherwise the two launches.
static int launch_td_critic(mdp_core* c, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B, const float* batch,
                            const int64_t* idx, long long idx_stride, const float* u_target, int32_t u_stride, uint64_t seed,
                            uint64_t counter, float* y_scratch, long long y_stride, void* stream) {
  MDP_REQUIRE(c && c->d_agents, "mdp_update: core not bound");
  int rc = check_lay(c, lay);
  if (rc) return rc;
  MDP_REQUIRE(batch && y_scratch && B > 0 && agent >= 0 && count > 0 && agent + count <= c->cfg.n_agents, "mdp_update: bad argument");
  const Plan p = plan_for(c, B, count);
  const ResPlan rp = make_res_plan(c, p, agent);
  if (rp.fuse_ok && !c->no_fuse && !want_tc(c, B, count) && !want_tc(c, B, count, true)) {
    cudaStream_t st = (cudaStream_t)stream;
    const bool prepared = take_prepared(c, agent, count);
    if (!prepared) MDP_CUDA(cudaMemsetAsync(c->stats + 8 * agent, 0, 8 * sizeof(double) * count, st));
    CoreDev d = core_dev(c);
    return dispatch(c->cfg.num_units, p, [&](auto u_, auto tm_, auto) -> int {
      constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
      auto kern = k_td_target_res<U, TMv, true>;
      int rc2 = set_smem(kern, rp.td_fused);
      if (rc2) return rc2;
      return launch_maybe_pdl("k_td_target_res<fused critic>", prepared, kern, dim3(cdiv(B, TMv), count), rp.td_groups * NT, rp.td_fused,
                              st, d, (int)agent, *lay, (int)B, batch, (const long long*)idx, u_target, (int)u_stride, (unsigned long)seed,
                              (unsigned long)counter, y_scratch, (float*)nullptr, (int)rp.XPf, (long long)idx_stride, (long long)y_stride);
    });
  }
  rc = launch_td_target(c, agent, count, lay, B, batch, idx, idx_stride, u_target, u_stride, seed, counter, y_scratch, y_stride, nullptr,
                        stream);
  if (rc) return rc;
  return launch_critic_grads(c, agent, count, lay, B, batch, idx, idx_stride, y_scratch, y_stride, nullptr, stream);
}

extern "C" int mdp_td_target(mdp_core* c, int32_t agent, const mdp_ring_layout* lay, int32_t B, const float* batch,
                             const int64_t* idx, const float* u_target, int32_t u_stride, uint64_t seed, uint64_t counter,
                             float* y_out, float* target_act_out, void* stream) {
  return launch_td_target(c, agent, 1, lay, B, batch, idx, 0, u_target, u_stride, seed, counter, y_out, 0, target_act_out, stream);
}

extern "C" int mdp_td_target_all(mdp_core* c, const mdp_ring_layout* lay, int32_t B, const float* batch, const int64_t* idx,
                                 int64_t idx_agent_stride, uint64_t seed, uint64_t counter, float* y_out, void* stream) {
  MDP_REQUIRE(c && y_out, "mdp_td_target_all: null argument");
  return launch_td_target(c, 0, c->cfg.n_agents, lay, B, batch, idx, idx_agent_stride, nullptr, 0, seed, counter, y_out, B, nullptr,
                          stream);
}

static int launch_critic_grads(mdp_core* c, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B, const float* batch,
                               const int64_t* idx, long long idx_stride, const float* y, long long y_stride, float* q_out,
                               void* stream) {
  MDP_REQUIRE(c && c->d_agents, "mdp_critic_grads: core not bound");
  int rc = check_lay(c, lay);
  if (rc) return rc;
  MDP_REQUIRE(batch && y && B > 0 && agent >= 0 && count > 0 && agent + count <= c->cfg.n_agents, "mdp_critic_grads: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CoreDev d = core_dev(c);
  const Plan p = plan_for(c, B, count);
  const ResPlan rp = make_res_plan(c, p, agent);
  const long long* ridx = (const long long*)idx;
  if (want_tc(c, B, count, true)) {
    rc = launch_critic_grads_tc(c, d, agent, count, lay, B, batch, ridx, idx_stride, y, y_stride, q_out, st);
    if (rc != MDP_ENOTSUP) return rc;
  }
  return dispatch(c->cfg.num_units, p, [&](auto u_, auto tm_, auto res_) -> int {
    constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
    constexpr bool RES = decltype(res_)::value;
    if (rp.ok) {
      auto kern = k_critic_grads_res<U, TMv>;
      int rc2 = set_smem(kern, rp.critic);
      if (rc2) return rc2;
      kern<<<dim3(cdiv(B, TMv), count), NT, rp.critic, st>>>(d, agent, *lay, B, batch, ridx, y, q_out, rp.XPf, idx_stride, y_stride);
      return check_launch("k_critic_grads_res");
    }
    auto kern = k_critic_grads<U, TMv, RES>;
    const size_t smem = smem_for(U, p, 1, 1, 2, TMv + 32);
    int rc2 = set_smem(kern, smem);
    if (rc2) return rc2;
    kern<<<dim3(cdiv(B, TMv), count), NT, smem, st>>>(d, agent, *lay, B, batch, ridx, y, q_out, p.max_net, idx_stride, y_stride);
    return check_launch("k_critic_grads");
  });
}

extern "C" int mdp_critic_grads(mdp_core* c, int32_t agent, const mdp_ring_layout* lay, int32_t B, const float* batch,
                                const int64_t* idx, const float* y, float* q_out, void* stream) {
  return launch_critic_grads(c, agent, 1, lay, B, batch, idx, 0, y, 0, q_out, stream);
}

static int launch_actor_grads(mdp_core* c, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B, const float* batch,
                              const int64_t* idx, long long idx_stride, const float* u_actor, int32_t u_stride, uint64_t seed,
                              uint64_t counter, void* stream, bool pdl = false) {
  MDP_REQUIRE(c && c->d_agents, "mdp_actor_grads: core not bound");
  int rc = check_lay(c, lay);
  if (rc) return rc;
  MDP_REQUIRE(batch && B > 0 && agent >= 0 && count > 0 && agent + count <= c->cfg.n_agents, "mdp_actor_grads: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CoreDev d = core_dev(c);
  const Plan p = plan_for(c, B, count);
  const ResPlan rp = make_res_plan(c, p, agent);
  const long long* ridx = (const long long*)idx;
  if (want_tc(c, B, count, true)) {
    rc = launch_actor_grads_tc(c, d, agent, count, lay, B, batch, ridx, idx_stride, u_actor, u_stride, seed, counter, st);
    if (rc != MDP_ENOTSUP) return rc;
  }
  return dispatch(c->cfg.num_units, p, [&](auto u_, auto tm_, auto res_) -> int {
    constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
    constexpr bool RES = decltype(res_)::value;
    if (rp.ok) {
      auto kern = k_actor_grads_res<U, TMv>;
      int rc2 = set_smem(kern, rp.actor);
      if (rc2) return rc2;
      // pdl (mdp_update_agent / mdp_update_all): programmatic dependent launch on the critic's optimizer kernel -- the sampled
      // rows (complete since the index draw) and the actor net stream into shared memory while that kernel still runs
      static const bool pdl_enabled = []() { const char* e = getenv("MDP_PDL"); return !(e && e[0] == '0'); }();
      cudaLaunchConfig_t lc;
      cudaLaunchAttribute at[1];
      memset(&lc, 0, sizeof(lc));
      lc.gridDim = dim3(cdiv(B, TMv), count); lc.blockDim = dim3(NT); lc.dynamicSmemBytes = rp.actor; lc.stream = st;
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      lc.attrs = at;
      lc.numAttrs = (pdl && pdl_enabled) ? 1 : 0;
      MDP_CUDA(cudaLaunchKernelEx(&lc, kern, d, (int)agent, *lay, (int)B, batch, ridx, u_actor, (int)u_stride, (unsigned long)seed,
                                  (unsigned long)counter, (int)rp.XPf, (long long)idx_stride, (long long)0));
      return check_launch("k_actor_grads_res");
    }
    auto kern = k_actor_grads<U, TMv, RES>;
    const size_t smem = smem_for(U, p, 2, 2, 4, 3 * TMv * KPAD + TMv);
    int rc2 = set_smem(kern, smem);
    if (rc2) return rc2;
    kern<<<dim3(cdiv(B, TMv), count), NT, smem, st>>>(d, agent, *lay, B, batch, ridx, u_actor, u_stride, seed, counter, p.max_net,
                                                     idx_stride, 0);
    return check_launch("k_actor_grads");
  });
}

extern "C" int mdp_critic_grads_all(mdp_core* c, const mdp_ring_layout* lay, int32_t B, const float* batch, const int64_t* idx,
                                    int64_t idx_agent_stride, const float* y, void* stream) {
  MDP_REQUIRE(c, "mdp_critic_grads_all: null core");
  return launch_critic_grads(c, 0, c->cfg.n_agents, lay, B, batch, idx, idx_agent_stride, y, B, nullptr, stream);
}

extern "C" int mdp_actor_grads(mdp_core* c, int32_t agent, const mdp_ring_layout* lay, int32_t B, const float* batch,
                               const int64_t* idx, const float* u_actor, int32_t u_stride, uint64_t seed, uint64_t counter,
                               void* stream) {
  return launch_actor_grads(c, agent, 1, lay, B, batch, idx, 0, u_actor, u_stride, seed, counter, stream);
}

// Grouped ("Jacobi") round: every agent's TD target is computed from the PRE-round target actors, then all
// critics step, then all actors -- five launches for the whole round instead of five per agent.  Deviates from
// the reference's sequential order (maddpg.py:181-194 run agent by agent, train.py:160-161) only in that agent j
// does not see the polyak step (1% blend of one Adam step) of agents i < j made earlier in the same round.
extern "C" int mdp_update_all(mdp_core* c, const mdp_ring_layout* lay, int32_t B, const float* batch, const int64_t* idx,
                              int64_t idx_agent_stride, uint64_t seed, uint64_t counter, float* y_scratch, float grad_scale,
                              void* stream) {
  MDP_REQUIRE(c && y_scratch, "mdp_update_all: null argument");
  const int n = c->cfg.n_agents;
  int rc = launch_td_critic(c, 0, n, lay, B, batch, idx, idx_agent_stride, nullptr, 0, seed, counter, y_scratch, B, stream);
  if (rc) return rc;
  rc = clip_adam_polyak_all_impl(c, 1, grad_scale, 1, stream, true);
  if (rc) return rc;
  rc = launch_actor_grads(c, 0, n, lay, B, batch, idx, idx_agent_stride, nullptr, 0, seed, counter, stream, true);
  if (rc) return rc;
  return clip_adam_polyak_all_impl(c, 0, grad_scale, 1, stream, true);
}

extern "C" int mdp_update_prepare(mdp_core* c, int32_t agent, int32_t count, int64_t* idx_out, int32_t B_total, int64_t length,
                                  uint64_t seed, uint64_t counter, void* stream) {
  MDP_REQUIRE(c && c->d_agents && idx_out && B_total > 0, "mdp_update_prepare: bad argument");
  MDP_REQUIRE(agent >= 0 && count > 0 && agent + count <= c->cfg.n_agents, "mdp_update_prepare: agents [%d, %d) out of range", agent,
              agent + count);
  MDP_REQUIRE(length > 0 || c->ctl, "mdp_update_prepare: length <= 0 needs a control block (mdp_core_set_ctl)");
  const int n = std::max(B_total, 8 * count);
  // The draw reads the control block before its dependency wait: fine inside a captured graph (the block only moves at the end
  // of a round, and graph launches are fully ordered) or without a block; otherwise keep plain stream order.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  MDP_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cap));
  const bool early = !c->ctl || cap == cudaStreamCaptureStatusActive;
  int rc = launch_maybe_pdl("k_update_prepare", early, k_update_prepare, dim3(cdiv(n, 256)), 256, 0, (cudaStream_t)stream, (long long*)idx_out,
                            (int)B_total, (long long)length, (unsigned long)seed, (unsigned long)counter,
                            (const unsigned long long*)c->ctl, c->stats + 8 * agent, (int)(8 * count));
  if (rc) return rc;
  c->prep_agent = agent;
  c->prep_count = count;
  return MDP_OK;
}

extern "C" int mdp_update_agent(mdp_core* c, int32_t agent, const mdp_ring_layout* lay, int32_t B, const float* batch,
                                const int64_t* idx, const float* u_target, const float* u_actor, int32_t u_stride,
                                uint64_t seed, uint64_t counter, float* y_scratch, void* stream) {
  int rc = launch_td_critic(c, agent, 1, lay, B, batch, idx, 0, u_target, u_stride, seed, counter, y_scratch, 0, stream);
  if (rc) return rc;
  const float scale = c->peer_world > 1 ? 1.0f / (float)c->peer_world : 1.0f;  // fused peer all-reduce (mdp_core_bind_peers)
  rc = clip_adam_polyak_impl(c, agent, 1, scale, 1, stream, true);
  if (rc) return rc;
  rc = launch_actor_grads(c, agent, 1, lay, B, batch, idx, 0, u_actor, u_stride, seed, counter, stream, true);
  if (rc) return rc;
  return clip_adam_polyak_impl(c, agent, 0, scale, 1, stream, true);
}
