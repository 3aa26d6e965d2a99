// The fork's tanh-policy algorithms on the MADDPG MLP tiles (sm_100a, fp32): MATD3 and the best/worst-policy "COMA" variant
// (SURVEY.md 8(f) rank 3).  Four building blocks; maddpg_b200/algorithms.py composes them exactly as the reference's graph does:
//
//   mdp_td3_policy_act     Policy._build                       maddpg/modules/policy.py:63-88   (tanh, clipped target noise, Box rescale)
//   mdp_td3_q_target       MaTD3Module.compute_qvalue          maddpg/modules/matd3module.py:113-123 (min over twin target critics, TD combine)
//                          ComaModule.compute_{global,personal}_qvalue   maddpg/modules/comamodule.py:155-171
//   mdp_td3_policy_grads   Policy.create_optimizer's gradient  maddpg/modules/policy.py:90-100 through the critic's TARGET network
//                          (matd3module.py:96-99, comamodule.py:118-129)
//   mdp_td3_polyak         LaggingNetwork.update_target        maddpg/modules/laggingnetwork.py:36-48
//
// The critic step itself (Critic.create_optimizer, critic.py:78-88: mse(values - target)) is mdp_critic_grads, and the Adam step
// is mdp_clip_adam_polyak with grad_clip = 0 (grad_norm_clipping=None, tf_util.py:171-175) -- unchanged kernels.
//
// Every kernel owns a tile of TM batch rows per CTA and carries it through the whole 3-layer MLP(s) in shared memory (the tiles
// of mdp_mlp.cuh, FFMA2 inner loops); grid.y = agent.  The fork's 64-unit nets stay resident in shared memory for the whole
// kernel (RES: one load per CTA, a critic + a policy + both transposed W2 for the gradient kernel); wider nets stream their
// weights through a staging chunk.
#include "mdp_mlp.cuh"

#include <algorithm>
#include <type_traits>

namespace mdp {
CoreDev core_dev_for_rollout(const mdp_core* c);

struct Td3Head {  // Box rescale of the tanh output, per agent: interval = (high - low) / 2, adjust = interval + low (policy.py:76-84)
  float scale[MDP_MAX_AGENTS], shift[MDP_MAX_AGENTS];
};

// N(0, 1) draw of element (row, col): Box-Muller on two Philox uniforms
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t counter, uint32_t tag, long long row, int col) {
  const uint4 r = Philox::gen(seed, (uint32_t)row, (uint32_t)(row >> 32) ^ (tag << 8) ^ (uint32_t)col, (uint32_t)counter,
                              (uint32_t)(counter >> 32));
  const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = Philox::u01(r.y);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__device__ __forceinline__ float rescale(float t, float scale, float shift) { return __fadd_rn(__fmul_rn(t, scale), shift); }

// ---------------------------------------------------------------------------------------------
// tanh-policy actions of every agent.  grid = (ceil(B/TM), n_agents)
// ---------------------------------------------------------------------------------------------
template <int U, int TM, bool RES>
__global__ void __launch_bounds__(NT) k_td3_policy_act(CoreDev C, Td3Head H, int max_net, int use_target, int B, const float* __restrict__ obs,
                                                       int obs_stride, const float* __restrict__ noise, int noise_stride,
                                                       float noise_std, float noise_clip, uint64_t seed, uint64_t counter,
                                                       float* __restrict__ act, int act_stride, int shared_agent) {
  if (C.ctl) counter += C.ctl[0];  // control block (mdp_core_set_ctl): a replayed CUDA graph draws fresh noise every time
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sW = sm.take(RES ? max_net : KC * U);   // RES: the whole net stays in shared memory; else a staging chunk
  float* sX = sm.take(TM * XP);
  float* sH1 = sm.take(TM * (U + 4));
  float* sH2 = sm.take(TM * (U + 4));
  float* sL = sm.take(TM * KPAD);
  const int i = blockIdx.y;
  const AgentDev& ag = C.agents[i];
  // PolicyGroup(shared=True): one policy serves every name (policygroup.py:26-37, 54-70)
  MlpW w = C.agents[shared_agent >= 0 ? shared_agent : i].net[use_target ? MDP_NET_TARGET_P : MDP_NET_P];
  if (RES) w = load_net<U>(G, sW, w);  // visible after the first barrier inside layer1
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  XSrc xs = make_xsrc(obs + ag.obs_off, obs_stride, ag.obs_dim);
  forward_hidden<U, TM, RES>(G, xs, w, row0, nrows, sX, sW, sH1, sH2);
  actor_head<U, TM>(G, sH2, w, sL);
  const int K = ag.act_dim;
  for (int idx = threadIdx.x; idx < nrows * K; idx += NT) {
    const int r = idx / K, a = idx - r * K;
    float t = tanhf(sL[r * KPAD + a]);
    if (noise_std > 0.f) {  // clip(tanh + clip(N(0, std), -c, c), -1, 1)   policy.py:72-75
      const float z = noise ? noise[(row0 + r) * noise_stride + ag.act_off + a] : philox_normal(seed, counter, 0x300u + i, row0 + r, a);
      const float n = fminf(fmaxf(z * noise_std, -noise_clip), noise_clip);
      t = fminf(fmaxf(t + n, -1.0f), 1.0f);
    }
    act[(row0 + r) * act_stride + ag.act_off + a] = rescale(t, H.scale[i], H.shift[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// q_j = min_c Q_{c,j}([x | act]) and y_j = rew_j + gamma (1 - done_j) q_j.  grid = (ceil(B/TM), n_agents)
// ---------------------------------------------------------------------------------------------
template <int U, int TM, bool RES>
__global__ void __launch_bounds__(NT) k_td3_q_target(CoreDev Ca, CoreDev Cb, int max_net, int n_critics, int use_target, mdp_ring_layout L, int B,
                                                     const float* __restrict__ batch, int obs_col0, const float* __restrict__ act,
                                                     int act_stride, const float* __restrict__ rew_override,
                                                     const float* __restrict__ rew_minus, int shared_agent, float gamma,
                                                     float* __restrict__ q_out, float* __restrict__ y_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sW = sm.take(RES ? max_net : KC * U);
  float* sX = sm.take(TM * XP);
  float* sH1 = sm.take(TM * (U + 4));
  float* sH2 = sm.take(TM * (U + 4));
  float* sQ = sm.take(2 * TM);
  const int j = blockIdx.y;
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  const int R = L.row_stride;
  XSrc xs = make_xsrc(batch + obs_col0, R, L.obs_sum);
  xs.g1 = act; xs.ld1 = act_stride; xs.n1 = L.act_sum;
  for (int c = 0; c < n_critics; ++c) {
    const AgentDev& me = (c ? Cb : Ca).agents[shared_agent >= 0 ? shared_agent : j];  // shared group: one critic serves every name
    MlpW w = me.net[use_target ? MDP_NET_TARGET_Q : MDP_NET_Q];
    if (RES) w = load_net<U>(G, sW, w);  // the previous user of sW finished before the barrier that ends critic_head
    forward_hidden<U, TM, RES>(G, xs, w, row0, nrows, sX, sW, sH1, sH2);
    critic_head<U, TM>(G, sH2, w, sQ + c * TM);
  }
  if (threadIdx.x < nrows) {
    const int r = threadIdx.x;
    float q = sQ[r];
    if (n_critics > 1) q = fminf(q, sQ[TM + r]);
    if (q_out) q_out[(long long)j * B + row0 + r] = q;
    if (y_out) {
      const float* row = batch + (row0 + r) * R;
      float rew = rew_override ? rew_override[(long long)j * B + row0 + r] : row[L.rw_off + j];
      if (rew_minus) rew = __fsub_rn(rew, rew_minus[(long long)j * B + row0 + r]);
      const float done = row[L.dn_off + j];
      // float32 graph arithmetic, left to right: R + ((gamma * (1 - D)) * Q)
      y_out[(long long)j * B + row0 + r] = __fadd_rn(rew, __fmul_rn(__fmul_rn(gamma, __fsub_rn(1.0f, done)), q));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// gradient of -mean(sign * Q_j(o, [a_-j, a_j])) wrt policy j, a_j = tanh(policy_j(o_j)) * scale + shift, a_-j from `act_all`
// (the other policies' current actions: the loss of policy j is differentiated wrt policy j's variables only, policy.py:95-99).
// grid = (ceil(B/TM), n_agents)
// ---------------------------------------------------------------------------------------------
template <int U, int TM, bool RES>
__global__ void __launch_bounds__(NT) k_td3_policy_grads(CoreDev Cp, CoreDev Cq, Td3Head H, int max_net, int critic_use_target, float sign,
                                                         mdp_ring_layout L, int B, const float* __restrict__ batch,
                                                         const float* __restrict__ act_all, int act_stride, int shared_policy,
                                                         int critic_agent) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int HP = U + 4;
  const Grp G{(int)threadIdx.x, 0};
  SmemCarve sm(smem_raw);
  float* sW = sm.take(RES ? max_net : KC * U);   // RES: critic net ; else staging chunk
  float* sWp = sm.take(RES ? max_net : 4);       // RES: policy net
  float* sWT = sm.take(RES ? U * U : 4);         // RES: critic W2^T
  float* sWTp = sm.take(RES ? U * U : 4);        // RES: policy W2^T
  float* sX = sm.take(TM * XP);
  float* sH1 = sm.take(TM * HP);
  float* sH2 = sm.take(TM * HP);
  float* sP1 = sm.take(TM * HP);   // policy h1
  float* sP2 = sm.take(TM * HP);   // policy h2
  float* sT = sm.take(TM * KPAD);  // pre-activation, then tanh
  float* sA = sm.take(TM * KPAD);  // rescaled action
  float* sDa = sm.take(TM * KPAD); // dQ/da, then dL/d(pre-activation)
  float* sQ = sm.take(TM);
  const int j = blockIdx.y;
  const AgentDev& me = Cp.agents[j];
  // shared policy group: name j's observation and action columns, the one shared net and its gradient bucket; its loss is
  // -mean(value of the group's first name) and reaches the shared variables through EVERY name's action (policygroup.py:129-135)
  const int pj = shared_policy >= 0 ? shared_policy : j, cj = critic_agent >= 0 ? critic_agent : j;
  const bool lead = shared_policy < 0 || j == 0;   // step counter and loss are the group's, counted once
  MlpW pw = Cp.agents[pj].net[MDP_NET_P];
  MlpW qw = Cq.agents[cj].net[critic_use_target ? MDP_NET_TARGET_Q : MDP_NET_Q];
  const MlpG& pg = Cp.agents[pj].grad[0];
  if (RES) {
    qw = load_net<U>(G, sW, qw);   // coalesced 16-byte loads; the transposed W2 copies are then made from shared memory
    pw = load_net<U>(G, sWp, pw);
    __syncthreads();
    for (int idx = threadIdx.x; idx < U * U; idx += NT) {
      const int k = idx / U, ul = idx - k * U;
      sWT[ul * U + k] = qw.W2[idx];
      sWTp[ul * U + k] = pw.W2[idx];
    }  // published by the first barrier inside layer1
  }
  const long long row0 = (long long)blockIdx.x * TM;
  const int nrows = (int)min((long long)TM, B - row0);
  const int R = L.row_stride, K = me.act_dim;
  if (blockIdx.x == 0 && threadIdx.x == 0 && lead) Cp.adam_t[2 * pj + 0] += 1;

  XSrc xp = make_xsrc(batch + me.obs_off, R, me.obs_dim);
  forward_hidden<U, TM, RES>(G, xp, pw, row0, nrows, sX, sW, sP1, sP2);
  actor_head<U, TM>(G, sP2, pw, sT);
  const float scale = H.scale[j], shift = H.shift[j];
  for (int idx = threadIdx.x; idx < TM * K; idx += NT) {
    const int r = idx / K, a = idx - r * K;
    const float t = tanhf(sT[r * KPAD + a]);
    sT[r * KPAD + a] = t;
    sA[r * KPAD + a] = rescale(t, scale, shift);
  }
  __syncthreads();
  // critic on [o | a_all with a_j replaced]
  const int a_col0 = L.obs_sum + me.act_off;
  XSrc xq = make_xsrc(batch, R, L.obs_sum);
  xq.g1 = act_all; xq.ld1 = act_stride; xq.n1 = L.act_sum;
  xq.s_over = sA; xq.over_ld = KPAD; xq.over_c0 = a_col0; xq.over_n = K;
  forward_hidden<U, TM, RES>(G, xq, qw, row0, nrows, sX, sW, sH1, sH2);
  critic_head<U, TM>(G, sH2, qw, sQ);
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    double sq = (r < nrows && r < TM) ? -(double)sign * (double)sQ[r] : 0.0;
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (r == 0 && lead) atomicAdd(Cp.stats + 8 * pj + 1, sq);
  }
  const float dq = -sign / (float)B;
  for (int idx = threadIdx.x; idx < TM * U; idx += NT) {
    const int r = idx / U, u = idx - r * U;
    const float h = sH2[r * HP + u];
    sH2[r * HP + u] = (h > 0.f && r < nrows) ? dq * qw.W3[u] : 0.f;
  }
  __syncthreads();
  backward_hidden<U, TM, RES>(G, xq, qw, sWT, nullptr, row0, nrows, sX, sW, sH1, sH2);  // dz1 of the critic -> sH1
  for (int idx = threadIdx.x; idx < TM * K; idx += NT) {
    const int r = idx / K, a = idx - r * K;
    const float* w1row = qw.W1 + (size_t)(a_col0 + a) * U;
    float s = 0.f;
    for (int u = 0; u < U; ++u) s = fmaf(sH1[r * HP + u], w1row[u], s);
    const float t = sT[r * KPAD + a];
    sDa[r * KPAD + a] = (r < nrows) ? s * scale * (1.0f - t * t) : 0.f;  // through the rescale and the tanh
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
  backward_hidden<U, TM, RES>(G, xp, pw, sWTp, &pg, row0, nrows, sX, sW, sP1, sP2);
}

// ---------------------------------------------------------------------------------------------
// target <- polyak * target + (1 - polyak) * running for the policy (mask bit 0) and / or critic (bit 1) of every agent.
// grid = (chunks, 2 * n_agents)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_td3_polyak(const AgentDev* __restrict__ agents, int units, int mask, float pol, float opol) {
  const int j = blockIdx.y >> 1, which = blockIdx.y & 1;
  if (!((mask >> which) & 1)) return;
  const MlpW w = agents[j].net[which ? MDP_NET_Q : MDP_NET_P], wt = agents[j].net[which ? MDP_NET_TARGET_Q : MDP_NET_TARGET_P];
  const int n = w.in * units + units + units * units + units + units * w.out + w.out;  // [W1|b1|W2|b2|W3|b3], contiguous
  const float* __restrict__ p = w.W1;
  float* __restrict__ t = const_cast<float*>(wt.W1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    t[i] = __fadd_rn(__fmul_rn(pol, t[i]), __fmul_rn(opol, p[i]));
}

template <typename Kern>
static int td3_smem(Kern kern, size_t smem) {
  if (smem > 48 * 1024) MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return MDP_OK;
}

// Launch plan, like the MADDPG kernels' (mdp_train.cu make_plan): 64-unit nets small enough that a critic, a policy and their two
// transposed W2 fit one CTA's shared memory stay resident there (RES); wider nets stream their weights in KC-row chunks.
struct Td3Plan {
  bool res;
  int max_net;
};

static Td3Plan td3_plan(const mdp_core* c) {
  Td3Plan p;
  int mx = 0;
  for (int i = 0; i < c->cfg.n_agents; ++i)
    for (int k = 0; k < 4; ++k) mx = std::max(mx, net_floats_padded(c->lay.net_in[i][k], c->cfg.num_units, c->lay.net_out[i][k]));
  p.max_net = mx;
  p.res = c->cfg.num_units == 64 && (size_t)(2 * mx + 2 * 64 * 64) * 4 <= 120 * 1024;
  return p;
}

// bytes of dynamic shared memory: weight buffers + transposed W2 copies + x chunk + n_act activation tiles + extras
static size_t td3_floats(int U, int TM, const Td3Plan& p, int n_wbuf, int n_wT, int n_act, int extra) {
  size_t f = (size_t)(p.res ? p.max_net : KC * U) + (size_t)(n_wbuf - 1) * (p.res ? p.max_net : 4);
  f += (size_t)n_wT * (p.res ? U * U : 4);
  f += (size_t)TM * XP + (size_t)n_act * TM * (U + 4) + extra + 128;
  return f * sizeof(float);
}

template <int V> using IC = std::integral_constant<int, V>;
template <bool V> using BC = std::integral_constant<bool, V>;

template <typename F>
static int td3_dispatch(int U, int rows, int n_agents, const Td3Plan& p, F&& f) {
  // 16-row tiles while all of a launch's CTAs (rows / 16 per agent, one resident CTA per SM) fit the 148 SMs in one wave; 32-row
  // tiles beyond that: batch 1024 x 3 agents is 192 CTAs of 16 rows (two waves) or 96 of 32 -- measured 8 % faster per step
  const bool small = (long long)cdiv(rows, 16) * n_agents <= 148;
  if (U == 64) {
    if (small) return p.res ? f(IC<64>{}, IC<16>{}, BC<true>{}) : f(IC<64>{}, IC<16>{}, BC<false>{});
    return p.res ? f(IC<64>{}, IC<32>{}, BC<true>{}) : f(IC<64>{}, IC<32>{}, BC<false>{});
  }
  if (U == 128) return small ? f(IC<128>{}, IC<16>{}, BC<false>{}) : f(IC<128>{}, IC<32>{}, BC<false>{});
  return fail(MDP_ENOTSUP, "num_units %d: the MLP tiles are built for 64 and 128", U);
}

static int td3_check_core(const mdp_core* c, const char* what) {
  MDP_REQUIRE(c && c->d_agents, "%s: core not bound", what);
  for (int i = 0; i < c->cfg.n_agents; ++i) {
    MDP_REQUIRE(!c->cfg.local_q[i], "%s: the fork's critics are centralized (modules/critic.py:71-73), local_q is not", what);
    MDP_REQUIRE(c->cfg.act_dim[i] <= MAXK, "%s: action width %d exceeds %d", what, c->cfg.act_dim[i], MAXK);
  }
  return MDP_OK;
}

static int td3_same_shape(const mdp_core* a, const mdp_core* b, const char* what) {
  MDP_REQUIRE(a->cfg.n_agents == b->cfg.n_agents && a->obs_sum == b->obs_sum && a->act_sum == b->act_sum &&
                  a->cfg.num_units == b->cfg.num_units, "%s: the cores describe different agents", what);
  return MDP_OK;
}

static int td3_check_lay(const mdp_core* c, const mdp_ring_layout* lay, const char* what) {
  MDP_REQUIRE(lay && lay->n_agents == c->cfg.n_agents && lay->obs_sum == c->obs_sum && lay->act_sum == c->act_sum,
              "%s: the row layout does not match the core", what);
  return MDP_OK;
}

// a shared policy serves names with equal spaces only (policygroup.py:32-34)
static int td3_check_shared(const mdp_core* c, int shared, const char* what) {
  MDP_REQUIRE(shared < c->cfg.n_agents, "%s: shared agent %d of %d", what, shared, c->cfg.n_agents);
  if (shared >= 0)
    for (int i = 0; i < c->cfg.n_agents; ++i)
      MDP_REQUIRE(c->cfg.obs_dim[i] == c->cfg.obs_dim[shared] && c->cfg.act_dim[i] == c->cfg.act_dim[shared],
                  "%s: a shared policy needs equal observation and action spaces (policygroup.py:32-34)", what);
  return MDP_OK;
}

static Td3Head make_head(const mdp_core* c, const float* scale, const float* shift) {
  Td3Head h;
  for (int i = 0; i < MDP_MAX_AGENTS; ++i) {
    h.scale[i] = (scale && i < c->cfg.n_agents) ? scale[i] : 1.0f;
    h.shift[i] = (shift && i < c->cfg.n_agents) ? shift[i] : 0.0f;
  }
  return h;
}

}  // namespace mdp

using namespace mdp;

extern "C" int mdp_td3_policy_act(mdp_core* c, int32_t use_target, int32_t B, const float* obs, int32_t obs_stride,
                                  const float* noise, int32_t noise_stride, float noise_std, float noise_clip, uint64_t seed,
                                  uint64_t counter, const float* act_scale, const float* act_shift, float* act,
                                  int32_t act_stride, int32_t shared_agent, void* stream) {
  int rc = td3_check_core(c, "mdp_td3_policy_act");
  if (rc) return rc;
  MDP_REQUIRE(obs && act && B > 0 && obs_stride >= c->obs_sum && act_stride >= c->act_sum, "mdp_td3_policy_act: bad argument");
  MDP_REQUIRE(!noise || noise_stride >= c->act_sum, "mdp_td3_policy_act: bad noise stride");
  if ((rc = td3_check_shared(c, shared_agent, "mdp_td3_policy_act"))) return rc;
  const CoreDev d = core_dev_for_rollout(c);
  const Td3Head h = make_head(c, act_scale, act_shift);
  const Td3Plan p = td3_plan(c);
  return td3_dispatch(c->cfg.num_units, B, c->cfg.n_agents, p, [&](auto u_, auto tm_, auto res_) -> int {
    constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
    constexpr bool RES = decltype(res_)::value;
    auto kern = k_td3_policy_act<U, TMv, RES>;
    const size_t smem = td3_floats(U, TMv, p, 1, 0, 2, TMv * KPAD);
    int rc2 = td3_smem(kern, smem);
    if (rc2) return rc2;
    kern<<<dim3(cdiv(B, TMv), c->cfg.n_agents), NT, smem, (cudaStream_t)stream>>>(d, h, p.max_net, use_target, B, obs, obs_stride, noise,
                                                                                noise_stride, noise_std, noise_clip, seed, counter,
                                                                                act, act_stride, shared_agent);
    return check_launch("k_td3_policy_act");
  });
}

extern "C" int mdp_td3_q_target(mdp_core* ca, mdp_core* cb, int32_t use_target, const mdp_ring_layout* lay, int32_t B,
                                const float* batch, int32_t obs_field, const float* act, int32_t act_stride,
                                const float* rew_override, const float* rew_minus, int32_t shared_agent, float gamma,
                                float* q_out, float* y_out, void* stream) {
  int rc = td3_check_core(ca, "mdp_td3_q_target");
  if (rc) return rc;
  if (cb) {
    if ((rc = td3_check_core(cb, "mdp_td3_q_target"))) return rc;
    if ((rc = td3_same_shape(ca, cb, "mdp_td3_q_target"))) return rc;
  }
  if ((rc = td3_check_lay(ca, lay, "mdp_td3_q_target"))) return rc;
  MDP_REQUIRE(batch && act && B > 0 && act_stride >= ca->act_sum && (q_out || y_out) && (obs_field == 0 || obs_field == 1) &&
                  shared_agent < ca->cfg.n_agents, "mdp_td3_q_target: bad argument");
  const CoreDev da = core_dev_for_rollout(ca), db = core_dev_for_rollout(cb ? cb : ca);
  const Td3Plan p = td3_plan(ca);
  return td3_dispatch(ca->cfg.num_units, B, ca->cfg.n_agents, p, [&](auto u_, auto tm_, auto res_) -> int {
    constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
    constexpr bool RES = decltype(res_)::value;
    auto kern = k_td3_q_target<U, TMv, RES>;
    const size_t smem = td3_floats(U, TMv, p, 1, 0, 2, 2 * TMv);
    int rc2 = td3_smem(kern, smem);
    if (rc2) return rc2;
    kern<<<dim3(cdiv(B, TMv), ca->cfg.n_agents), NT, smem, (cudaStream_t)stream>>>(da, db, p.max_net, cb ? 2 : 1, use_target, *lay, B, batch,
                                                                                 obs_field ? lay->nx_off : 0, act, act_stride,
                                                                                 rew_override, rew_minus, shared_agent, gamma, q_out,
                                                                                 y_out);
    return check_launch("k_td3_q_target");
  });
}

extern "C" int mdp_td3_policy_grads(mdp_core* policy, mdp_core* critic, int32_t critic_use_target, float sign,
                                    const mdp_ring_layout* lay, int32_t B, const float* batch, const float* act_all,
                                    int32_t act_stride, const float* act_scale, const float* act_shift, int32_t shared_policy,
                                    int32_t critic_agent, void* stream) {
  int rc = td3_check_core(policy, "mdp_td3_policy_grads");
  if (rc) return rc;
  if ((rc = td3_check_core(critic, "mdp_td3_policy_grads"))) return rc;
  if ((rc = td3_same_shape(policy, critic, "mdp_td3_policy_grads"))) return rc;
  if ((rc = td3_check_lay(policy, lay, "mdp_td3_policy_grads"))) return rc;
  MDP_REQUIRE(batch && act_all && B > 0 && act_stride >= policy->act_sum && critic_agent < critic->cfg.n_agents,
              "mdp_td3_policy_grads: bad argument");
  if ((rc = td3_check_shared(policy, shared_policy, "mdp_td3_policy_grads"))) return rc;
  const CoreDev dp = core_dev_for_rollout(policy), dq = core_dev_for_rollout(critic);
  const Td3Head h = make_head(policy, act_scale, act_shift);
  const Td3Plan p = td3_plan(policy);
  return td3_dispatch(policy->cfg.num_units, B, policy->cfg.n_agents, p, [&](auto u_, auto tm_, auto res_) -> int {
    constexpr int U = decltype(u_)::value, TMv = decltype(tm_)::value;
    constexpr bool RES = decltype(res_)::value;
    auto kern = k_td3_policy_grads<U, TMv, RES>;
    const size_t smem = td3_floats(U, TMv, p, 2, 2, 4, 3 * TMv * KPAD + TMv);
    int rc2 = td3_smem(kern, smem);
    if (rc2) return rc2;
    kern<<<dim3(cdiv(B, TMv), policy->cfg.n_agents), NT, smem, (cudaStream_t)stream>>>(dp, dq, h, p.max_net, critic_use_target, sign, *lay, B,
                                                                                     batch, act_all, act_stride, shared_policy,
                                                                                     critic_agent);
    return check_launch("k_td3_policy_grads");
  });
}

extern "C" int mdp_td3_polyak(mdp_core* c, int32_t mask, double polyak, void* stream) {
  MDP_REQUIRE(c && c->d_agents, "mdp_td3_polyak: core not bound");
  MDP_REQUIRE(mask > 0 && mask < 4, "mdp_td3_polyak: mask selects the policy (1), the critic (2) or both (3)");
  k_td3_polyak<<<dim3(8, 2 * c->cfg.n_agents), 256, 0, (cudaStream_t)stream>>>(c->d_agents, c->cfg.num_units, mask, (float)polyak,
                                                                              (float)(1.0 - polyak));
  return check_launch("k_td3_polyak");
}
