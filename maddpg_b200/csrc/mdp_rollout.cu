// Persistent episode kernel: the whole lockstep rollout of experiments/train.py:110-133 --
//   action_n = [agent.action(obs)]          maddpg/trainer/maddpg.py:151-152 (mlp_model + SoftCategoricalPd.sample)
//   new_obs, rew, done = env.step(action_n) multiagent.environment.MultiAgentEnv.step
//   agent.experience(...)                   maddpg/trainer/maddpg.py:154-156 -> ReplayBuffer.add
//   env.reset() every max_episode_len steps train.py:127-129
// -- for `steps` steps in ONE launch.  A CTA owns 32 env instances for the whole episode: their SoA
// state, the joint observation tile, the sampled actions and (when they fit) ALL agents' actor
// weights stay resident in shared memory; per step the only global traffic is the joint replay row
// (obs_t, act_t, next_obs, rew, done) streaming out to the ring.  Results are identical to the
// per-step path (mdp_actor_act + mdp_env_step with ring + mdp_env_reset) on the same Philox counters.
#include "mdp_rollout.cuh"

namespace mdp {

#ifndef MDP_EP_UNROLL
#define MDP_EP_UNROLL 2
#endif
constexpr int EP_UNROLL = MDP_EP_UNROLL;  // k-loop unroll (in float4 A loads) of the episode kernel's hidden layer
constexpr int TM = REB;

// ---- actor tile of the episode kernel: 32 env rows x U units per group of GTH threads ----------------------------------
// Thread (ty, tx) = (tid >> 4, tid & 15) owns rows {RM*ty .. +RM-1}, RM = 512 / GTH, and columns {64g + 4tx .. +3}.  With
// resident weights a group is 128 threads (RM = 4, 16 outputs per thread): per 4 k-steps a thread issues 8 LDS.128 for 32
// FFMA2 -- half the shared-memory wavefronts per FMA of the 256-thread 2x4 tile, which ran at 38 % of the FMA pipe
// (125 cycles per k-step of layer 2 against 48; clock64 phase profile in DESIGN.md).  Every output still accumulates
// k = 0, 1, ... in order with fused multiply-adds, so the results are bit-identical to the per-step kernels.
// Measured on B200 (clock64 phase profile): the hidden layer now runs at ~84 cycles per k-step for 3 agents x
// 32 rows x 64 units = 73 FMA/clk/SM, insensitive to k-loop unroll depth and to software-pipelined operand loads; a pure
// register-operand FMA stream sustains 113 FMA/clk/SM (tools/fp32_probe.cu), FFMA2 halving the instruction count only.
template <int U, int GTH>
struct EpTile {
  static constexpr int RM = 512 / GTH, HP = U + 4;
  typedef float2 Acc[RM][U / 32];

  static __device__ __forceinline__ void zero(Acc& acc) {
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
      for (int c = 0; c < U / 32; ++c) acc[r][c] = make_float2(0.f, 0.f);
  }
  // acc += sA[rows][0..kc) * sW[0..kc)[cols]; scalar A loads (any alignment, any kc)
  static __device__ __forceinline__ void mma_sa(int tid, Acc& acc, const float* __restrict__ sA, int lda, const float* __restrict__ sW,
                                                int kc) {
    const int ty = tid >> 4, tx = tid & 15;
    const float* ap = sA + (RM * ty) * lda;
#pragma unroll 2
    for (int k = 0; k < kc; ++k) {
      float av[RM];
#pragma unroll
      for (int rr = 0; rr < RM; ++rr) av[rr] = ap[rr * lda + k];
#pragma unroll
      for (int g = 0; g < U / 64; ++g) {
        const float4 w = *reinterpret_cast<const float4*>(sW + k * U + g * 64 + 4 * tx);
#pragma unroll
        for (int rr = 0; rr < RM; ++rr) {
          const float2 a2 = make_float2(av[rr], av[rr]);
          acc[rr][2 * g + 0] = __ffma2_rn(a2, make_float2(w.x, w.y), acc[rr][2 * g + 0]);
          acc[rr][2 * g + 1] = __ffma2_rn(a2, make_float2(w.z, w.w), acc[rr][2 * g + 1]);
        }
      }
    }
  }
  // A given TRANSPOSED, sAT[k][32 rows]: the thread's RM = 4 rows are one float4 per k (128-thread groups only)
  static __device__ __forceinline__ void mma_t(int tid, Acc& acc, const float* __restrict__ sAT, const float* __restrict__ sW, int kc) {
    static_assert(RM == 4 || GTH != 128, "mma_t: 4 rows per thread");
    const int ty = tid >> 4, tx = tid & 15;
    const float* ap = sAT + 4 * ty;
#pragma unroll 2
    for (int k = 0; k < kc; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(ap + k * 32);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int g = 0; g < U / 64; ++g) {
        const float4 w = *reinterpret_cast<const float4*>(sW + k * U + g * 64 + 4 * tx);
#pragma unroll
        for (int rr = 0; rr < (RM < 4 ? RM : 4); ++rr) {
          const float2 a2 = make_float2(av[rr], av[rr]);
          acc[rr][2 * g + 0] = __ffma2_rn(a2, make_float2(w.x, w.y), acc[rr][2 * g + 0]);
          acc[rr][2 * g + 1] = __ffma2_rn(a2, make_float2(w.z, w.w), acc[rr][2 * g + 1]);
        }
      }
    }
  }
  // same with float4 A loads (lda % 4 == 0, kc % 4 == 0, 16-byte aligned rows)
  static __device__ __forceinline__ void mma(int tid, Acc& acc, const float* __restrict__ sA, int lda, const float* __restrict__ sW,
                                             int kc) {
    const int ty = tid >> 4, tx = tid & 15;
    const float* ap = sA + (RM * ty) * lda;
#pragma unroll EP_UNROLL
    for (int k = 0; k < kc; k += 4) {
      float av[RM][4];
#pragma unroll
      for (int rr = 0; rr < RM; ++rr) {
        const float4 a = *reinterpret_cast<const float4*>(ap + rr * lda + k);
        av[rr][0] = a.x; av[rr][1] = a.y; av[rr][2] = a.z; av[rr][3] = a.w;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int g = 0; g < U / 64; ++g) {
          const float4 w = *reinterpret_cast<const float4*>(sW + (k + kk) * U + g * 64 + 4 * tx);
#pragma unroll
          for (int rr = 0; rr < RM; ++rr) {
            const float2 a2 = make_float2(av[rr][kk], av[rr][kk]);
            acc[rr][2 * g + 0] = __ffma2_rn(a2, make_float2(w.x, w.y), acc[rr][2 * g + 0]);
            acc[rr][2 * g + 1] = __ffma2_rn(a2, make_float2(w.z, w.w), acc[rr][2 * g + 1]);
          }
        }
      }
    }
  }
  // sH[r][c] = relu(acc + bias[c])   (caller synchronises)
  static __device__ __forceinline__ void store_bias_relu(int tid, const Acc& acc, const float* __restrict__ bias, float* sH) {
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int g = 0; g < U / 64; ++g) {
      const int c = g * 64 + 4 * tx;
      const float4 b = *reinterpret_cast<const float4*>(bias + c);
#pragma unroll
      for (int rr = 0; rr < RM; ++rr) {
        float4 v;
        v.x = fmaxf(acc[rr][2 * g + 0].x + b.x, 0.f);
        v.y = fmaxf(acc[rr][2 * g + 0].y + b.y, 0.f);
        v.z = fmaxf(acc[rr][2 * g + 1].x + b.z, 0.f);
        v.w = fmaxf(acc[rr][2 * g + 1].y + b.w, 0.f);
        *reinterpret_cast<float4*>(sH + (RM * ty + rr) * HP + c) = v;
      }
    }
  }
  // output head, same arithmetic as actor_head_k (mdp_mlp.cuh): 8 threads per row, strided units, xor-shuffle tree
  template <int KK>
  static __device__ __forceinline__ void head_k(const Grp& G, const float* __restrict__ sH2, const MlpW& w, float* sL) {
    for (int r = G.tid >> 3; r < TM; r += GTH / 8) {
      const int part = G.tid & 7;
      float s[KK];
#pragma unroll
      for (int a = 0; a < KK; ++a) s[a] = 0.f;
      for (int u = part; u < U; u += 8) {
        const float h = sH2[r * HP + u];
        const float* w3 = w.W3 + u * KK;
#pragma unroll
        for (int a = 0; a < KK; ++a) s[a] = fmaf(h, w3[a], s[a]);
      }
#pragma unroll
      for (int a = 0; a < KK; ++a) {
        float v = s[a];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (part == 0) sL[r * KPAD + a] = v + w.b3[a];
      }
    }
    G.sync();
  }
  static __device__ __forceinline__ void head(const Grp& G, const float* __restrict__ sH2, const MlpW& w, float* sL) {
    if (w.out == 5) return head_k<5>(G, sH2, w, sL);
    if (w.out == 9) return head_k<9>(G, sH2, w, sL);
    const int K = w.out;
    for (int idx = G.tid; idx < TM * K; idx += GTH) {
      const int r = idx / K, a = idx - r * K;
      float s = 0.f;
      for (int u = 0; u < U; ++u) s = fmaf(sH2[r * HP + u], w.W3[u * K + a], s);
      sL[r * KPAD + a] = s + w.b3[a];
    }
    G.sync();
  }
  // Single-head actors with KK <= 8 outputs: head + Gumbel-softmax fused, 8 lanes per row, no shared-memory round trip and no
  // group barrier in between.  noise[p] = -log(-log u) of element (row head_row(tid, p), column tid & 7), drawn by the
  // caller BEFORE layer 1 so the Philox / log chains overlap the GEMM loads.  Same arithmetic, in the same order, as
  // head_k + gumbel_softmax below: strided partial sums, xor-shuffle tree, sequential softmax sum.
  static constexpr int NP = TM * 8 / GTH;
  // head row p of thread tid: the rows a warp takes here are the rows its own layer-2 tile produced (a warp of the GEMM tile
  // owns ty = 2w, 2w+1, i.e. rows [2*RM*w, 2*RM*(w+1)) ), so layer 1 -> layer 2 -> head hand-offs are warp-local
  static __device__ __forceinline__ int head_row(int tid, int p) { return 2 * RM * (tid >> 5) + ((tid & 31) >> 3) + 4 * p; }
  template <int KK>
  static __device__ __forceinline__ void head_gumbel(const Grp& G, const float* __restrict__ sH2, const MlpW& w, float* __restrict__ sOut,
                                                     int out_ld, int nrows, const float (&noise)[NP]) {
    const int part = G.tid & 7, lane_base = (G.tid & 31) & ~7;
    float s[NP][KK];
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int a = 0; a < KK; ++a) s[p][a] = 0.f;
#pragma unroll
    for (int j = 0; j < U / 8; ++j) {
      const int u = part + 8 * j;
      float h[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) h[p] = sH2[head_row(G.tid, p) * HP + u];
      const float* w3 = w.W3 + u * KK;
#pragma unroll
      for (int a = 0; a < KK; ++a) {
        const float wv = w3[a];
#pragma unroll
        for (int p = 0; p < NP; ++p) s[p][a] = fmaf(h[p], wv, s[p][a]);
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int r = head_row(G.tid, p);
      float mine = -INFINITY;  // perturbed logit of column `part`
#pragma unroll
      for (int a = 0; a < KK; ++a) {
        float v = s[p][a];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (part == a) mine = (v + w.b3[a]) + noise[p];
      }
      float m = -INFINITY;
#pragma unroll
      for (int a = 0; a < KK; ++a) m = fmaxf(m, __shfl_sync(0xffffffffu, mine, lane_base + a));
      const float z = part < KK ? expf(mine - m) : 0.f;
      float sum = 0.f;
#pragma unroll
      for (int a = 0; a < KK; ++a) sum += __shfl_sync(0xffffffffu, z, lane_base + a);
      if (part < KK && r < nrows) sOut[r * out_ld + part] = z / sum;
    }
  }
  // gumbel_softmax_tile (mdp_mlp.cuh) for a GTH-thread group, in-kernel Philox draws only
  static __device__ __forceinline__ void gumbel_softmax(const Grp& G, const float* __restrict__ sL, float* __restrict__ sOut, int out_ld,
                                                        int nrows, int K, int n_heads, const int* head_dim, long long row0,
                                                        uint64_t seed, uint64_t counter, uint32_t tag) {
    for (int idx = G.tid; idx < TM * K; idx += GTH) {
      const int r = idx / K, a = idx - r * K;
      if (r >= nrows) continue;
      sOut[r * out_ld + a] = sL[r * KPAD + a] + gumbel_from_u(philox_u(seed, counter, tag, row0 + r, a));
    }
    G.sync();
    for (int idx = G.tid; idx < TM * n_heads; idx += GTH) {
      const int r = idx / n_heads, h = idx - r * n_heads;
      if (r >= nrows) continue;
      const int o = h ? head_dim[0] : 0, n = head_dim[h];
      float m = -INFINITY;
      for (int a = 0; a < n; ++a) m = fmaxf(m, sOut[r * out_ld + o + a]);
      float z[MAXK];
      float s = 0.f;
#pragma unroll
      for (int a = 0; a < MAXK; ++a) {
        if (a < n) {
          z[a] = expf(sOut[r * out_ld + o + a] - m);
          s += z[a];
        }
      }
#pragma unroll
      for (int a = 0; a < MAXK; ++a)
        if (a < n) sOut[r * out_ld + o + a] = z[a] / s;
    }
    G.sync();
  }
};


// NG groups of GTH threads (128 with resident weights, else 256); group g runs the actor MLPs of agents g, g+NG, ...
// concurrently with the other groups (named barriers 1..NG); the env phases use all NG*GTH threads.
//
// Replay rows are ASSEMBLED IN PLACE: the CTA owns two [32][row_stride] row buffers (ping-pong).  In the buffer
// of step s the actors read obs_t from columns [0, sum D), the Gumbel-softmax writes act_t into [sum D, C), the
// env phase writes next_obs / rew / done into their columns (and next_obs also into the OTHER buffer's obs_t
// columns); then ONE thread streams the 32 finished rows to the ring with a TMA bulk store
// (cp.async.bulk.global.shared::cta) that overlaps with the next step's compute.
//
// SA > 0: simple_spread with SA agents -- the env phase of a step runs in REGISTERS on warp 0 (thread = env instance,
// spread_step / spread_obs of mdp_env_dev.cuh, the code of k_env_step_spread) instead of the table-driven CTA-collective
// phases: the state never leaves the 32 threads' registers between the prologue and the epilogue.
template <int U, bool RESIDENT, int SA>
__global__ void __launch_bounds__(!RESIDENT ? 1024 : SA == 2 ? 256 : SA == 3 ? 384 : 512) k_rollout_episode(CoreDev C, EnvParams P, const ObsCol* __restrict__ cols,
                                                          mdp_ring_layout L, RolloutArgs R, int NG) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int HP = U + 4, EBP = REB + 1;
  const int tid = threadIdx.x, NTB = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = NTB >> 5;
  constexpr int GTH = RESIDENT ? 128 : NT;
  typedef EpTile<U, GTH> Tile;
  const int grp = tid / GTH;
  const Grp G{tid - grp * GTH, grp + 1, GTH};
  const int OS = P.obs_stride, A = P.A, RS = L.row_stride;
  const int e0 = blockIdx.x * REB;
  const int nE = min(REB, R.E - e0);

  // ---- shared memory carve-up ------------------------------------------------------------------
  SmemCarve sm(smem_raw);
  float* sRow = sm.take(2 * TM * RS);  // two row buffers, 16-byte aligned rows
  float* sH1 = sm.take(NG * TM * HP) + grp * TM * HP;
  float* sH2 = sm.take(NG * TM * HP) + grp * TM * HP;
  float* sL = sm.take(NG * TM * KPAD) + grp * TM * KPAD;
  float* sRet = sm.take(A * EBP);
  float* sXT = sm.take(SA > 0 ? 6 * SA * SA * 32 : 0);  // SA > 0: transposed observation tile [column][env]
  float* sNoise = sm.take(8 * 256);     // SA > 0: Gumbel noise of the next step, [agent][row][8]
  float* sPart = sm.take(2 * 8 * 32);  // SA > 0: per-agent reward partials (landmark minima, collision counts)
  int* sOff = reinterpret_cast<int*>(sm.take(MDP_MAX_AGENTS + 1));
  ObsCol* sCols = reinterpret_cast<ObsCol*>(sm.take(2 * OS));
  float* sEnv = sm.take((int)(EnvTile<float, REB>::bytes(P.scomp, A, P.act_stride, false) / 4));
  // RESIDENT: every agent's actor net ; else one [KC][U] staging chunk per group
  float* sWts = RESIDENT ? sm.p : sm.p + grp * KC * U;
  EnvTile<float, REB> T;
  T.carve(sEnv, P, sRow + L.obs_sum);
  T.ASP = RS;

  unsigned long long counter = R.counter, episode = R.episode;
  long long cursor = R.cursor;
  if (R.ctl) {
    counter += R.ctl[0];
    cursor = (cursor + (long long)R.ctl[1]) % R.capacity;
    episode += R.ctl[2];
  }

  // ---- prologue: state tile, observation tile, column table, actor weights -------------------------
  env_load_state<float, REB>(P, T, (const float*)R.state, R.E, e0, nE);
  for (int i = tid; i < 2 * TM * RS; i += NTB) sRow[i] = 0.f;  // padding / done columns stay zero for good
  for (int c = tid; c < OS; c += NTB) sCols[c] = cols[c];
  for (int idx = tid; idx < A * EBP; idx += NTB) sRet[idx] = 0.f;
  if (tid == 0) {
    int o = 0;
    for (int i = 0; i < A; ++i) {
      sOff[i] = o;
      o += actor_net_floats(C.agents[i].obs_dim, U, C.agents[i].act_dim);
    }
    sOff[A] = o;
  }
  __syncthreads();
  for (int ee = warp; ee < nE; ee += nwarps)
    for (int c = lane; c < L.obs_sum; c += 32) sRow[ee * RS + c] = R.obs[(size_t)(e0 + ee) * OS + c];
  if (SA > 0) {
    for (int idx = tid; idx < L.obs_sum * 32; idx += NTB) {
      const int c = idx >> 5, ee = idx & 31;
      sXT[idx] = ee < nE ? R.obs[(size_t)(e0 + ee) * OS + c] : 0.f;
    }
  }
  if (RESIDENT) {
    for (int i = 0; i < A; ++i) {
      const float4* src = reinterpret_cast<const float4*>(C.agents[i].net[MDP_NET_P].W1);
      float4* dst = reinterpret_cast<float4*>(sWts + sOff[i]);
      const int n4 = (sOff[i + 1] - sOff[i]) >> 2;
      for (int q = tid; q < n4; q += NTB) dst[q] = src[q];
    }
  }
  __syncthreads();

  // SA > 0: thread (warp i < SA, lane) owns agent i of env instance `lane`: its state and the landmarks stay in registers;
  // positions are exchanged through the state tile (T.sS rows 4j, 4j+1), reward partials through sPart
  constexpr int SAc = SA > 0 ? SA : 1;
  float pxi = 0.f, pyi = 0.f, vxi = 0.f, vyi = 0.f, lxi = 0.f, lyi = 0.f, ret_reg = 0.f;
  float lx[SAc], ly[SAc];
  if (SA > 0 && warp < SA) {
    pxi = T.sS[(4 * warp + 0) * EBP + lane]; pyi = T.sS[(4 * warp + 1) * EBP + lane];
    vxi = T.sS[(4 * warp + 2) * EBP + lane]; vyi = T.sS[(4 * warp + 3) * EBP + lane];
    lxi = T.sS[(4 * SA + 2 * warp + 0) * EBP + lane]; lyi = T.sS[(4 * SA + 2 * warp + 1) * EBP + lane];
#pragma unroll
    for (int l = 0; l < SAc; ++l) { lx[l] = T.sS[(4 * SA + 2 * l + 0) * EBP + lane]; ly[l] = T.sS[(4 * SA + 2 * l + 1) * EBP + lane]; }
  }
  // The rewards run on a second set of warps (RW0 + i, the first warps of the next actor group) next to the observation
  // writes of the physics warps: barrier 8 = physics warps only, barrier 9 = both sets ("new positions are in the state
  // tile"), barrier 10 = reward warps only.
  constexpr int RW0 = 4;
  const bool phys_warp = SA > 0 && warp < SA, rew_warp = SA > 0 && warp >= RW0 && warp < RW0 + SA;
  if (rew_warp) {
    lxi = T.sS[(4 * SA + 2 * (warp - RW0) + 0) * EBP + lane];
    lyi = T.sS[(4 * SA + 2 * (warp - RW0) + 1) * EBP + lane];
  }
  const SpreadConsts<SAc> Cn = spread_consts<SAc>(P);
  const SpreadAgentConsts Ai = spread_agent_consts(P, phys_warp ? warp : rew_warp ? warp - RW0 : 0);
  auto env_bar = [] { asm volatile("bar.sync 8, %0;" ::"r"(32 * SAc) : "memory"); };
  auto pos_bar = [] { asm volatile("bar.sync 9, %0;" ::"r"(64 * SAc) : "memory"); };
  auto rew_bar = [] { asm volatile("bar.sync 10, %0;" ::"r"(32 * SAc) : "memory"); };

  // SA > 0 (every agent is Discrete(5)): the Gumbel noise of step s is drawn ahead of time into sNoise[agent][row][8] by
  // `nthr` threads -- all of them in the prologue, afterwards the warps that have no part in the env phase
  auto draw_noise = [&](int s, int t, int nthr) {
    for (int idx = t; idx < SAc * 160; idx += nthr) {
      const int i = idx / 160, rem = idx - i * 160, r = rem / 5, a = rem - r * 5;
      if (r < nE)
        sNoise[i * 256 + r * 8 + a] =
            gumbel_from_u(philox_u(R.seed, counter + (unsigned long long)s + 1ull, (uint32_t)i, (long long)e0 + r, a));
    }
  };
  if (SA > 0) {
    for (int idx = tid; idx < SAc * 256; idx += NTB) sNoise[idx] = 0.f;
    __syncthreads();
    draw_noise(0, tid, NTB);
    __syncthreads();
  }

  // descriptor of the group's first agent (its only one when A <= NG), loaded once
  const int ag_first = grp < A ? grp : 0;
  const int D0 = C.agents[ag_first].obs_dim, K0 = C.agents[ag_first].act_dim, nh0 = C.agents[ag_first].n_heads;
  const int obs_off0 = C.agents[ag_first].obs_off, act_off0 = C.agents[ag_first].act_off;
  MlpW w0 = C.agents[ag_first].net[MDP_NET_P];
  if (RESIDENT) w0 = net_at<U>(sWts + sOff[ag_first], D0, K0);

  // ---- the episode -----------------------------------------------------------------------------------
  long long ring_row = (cursor + e0) % R.capacity;  // ring row of this CTA's first env at step s (thread 0 keeps it current)
#ifdef MDP_EPISODE_PROF
  long long prof_t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, prof_c = clock64();
#define PROF_MARK(k) { const long long t_ = clock64(); prof_t[k] += t_ - prof_c; prof_c = t_; }
#else
#define PROF_MARK(k)
#endif
  for (int s = 0; s < R.steps; ++s) {
    float* buf = sRow + (s & 1) * TM * RS;        // rows of this step
    float* nxt = sRow + ((s & 1) ^ 1) * TM * RS;  // rows of the next step (receive obs_{t+1} as their obs_t)
    T.sA = buf + L.obs_sum;
    // (1) actions: a_i = gumbel_softmax(mlp_i(obs_i)) for every agent, written into the row's act columns
    for (int i = grp; i < A; i += NG) {
      const AgentDev& ag = C.agents[i];
      const bool first = i == grp;  // the group's first agent: descriptor cached in registers before the step loop
      const int D = first ? D0 : ag.obs_dim, K = first ? K0 : ag.act_dim;
      const int obs_off_i = first ? obs_off0 : ag.obs_off, act_off_i = first ? act_off0 : ag.act_off;
      MlpW w = w0;
      if (!first) {
        w = ag.net[MDP_NET_P];
        if (RESIDENT) w = net_at<U>(sWts + sOff[i], D, K);
      }
      const bool fused_head = (first ? nh0 : ag.n_heads) == 1 && K == 5;  // Discrete(5): every MPE movement head
      float noise[Tile::NP];
      if (fused_head && SA == 0) {
#pragma unroll
        for (int p = 0; p < Tile::NP; ++p) {
          const int r = Tile::head_row(G.tid, p), a = G.tid & 7;
          noise[p] = (a < K && r < nE)
                         ? gumbel_from_u(philox_u(R.seed, counter + (unsigned long long)s + 1ull, (uint32_t)i, (long long)e0 + r, a))
                         : 0.f;
        }
      }
      typename Tile::Acc acc;
      Tile::zero(acc);
      PROF_MARK(2)
      if (SA > 0) {
        Tile::mma_t(G.tid, acc, sXT + obs_off_i * 32, w.W1, D);
        PROF_MARK(7)
      } else if (RESIDENT) {
        Tile::mma_sa(G.tid, acc, buf + obs_off_i, RS, w.W1, D);
      } else {
        for (int k0 = 0; k0 < D; k0 += KC) {
          load_w_rows<U>(G, sWts, w.W1, k0, D);
          G.sync();
          Tile::mma_sa(G.tid, acc, buf + obs_off_i + k0, RS, sWts, min(KC, D - k0));
          G.sync();
        }
      }
      Tile::store_bias_relu(G.tid, acc, w.b1, sH1);
      PROF_MARK(9)
      // a warp's layer-2 rows are the rows its own lanes just wrote (thread (ty, tx): rows RM*ty.., ty = tid >> 4): with resident
      // weights nothing else is shared between the warps of a group, so the hand-offs need a warp barrier only
      if (RESIDENT) __syncwarp(); else G.sync();
      PROF_MARK(5)
      Tile::zero(acc);
      if (RESIDENT) {
        Tile::mma(G.tid, acc, sH1, HP, w.W2, U);
      } else {
        for (int k0 = 0; k0 < U; k0 += KC) {
          load_w_rows<U>(G, sWts, w.W2, k0, U);
          G.sync();
          Tile::mma(G.tid, acc, sH1 + k0, HP, sWts, KC);
          G.sync();
        }
      }
      Tile::store_bias_relu(G.tid, acc, w.b2, sH2);
      if (RESIDENT && fused_head) __syncwarp(); else G.sync();
      PROF_MARK(6)
      if (fused_head) {
        if (SA > 0) {  // drawn by the warps that idle during the previous step's env phase
#pragma unroll
          for (int p = 0; p < Tile::NP; ++p) noise[p] = sNoise[i * 256 + Tile::head_row(G.tid, p) * 8 + (G.tid & 7)];
        }
        Tile::template head_gumbel<5>(G, sH2, w, buf + L.obs_sum + act_off_i, RS, nE, noise);
        if (!RESIDENT) G.sync();  // streaming groups reuse their weight staging chunk for the next agent
      } else {
        Tile::head(G, sH2, w, sL);
        PROF_MARK(7)
        Tile::gumbel_softmax(G, sL, buf + L.obs_sum + act_off_i, RS, nE, K, ag.n_heads, ag.head_dim, (long long)e0, R.seed,
                             counter + (unsigned long long)s + 1ull, (uint32_t)i);
      }
      PROF_MARK(8)
    }
    // the bulk store of step s-1 must have finished READING `nxt` before phase (4) overwrites its obs columns
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();  // all groups' actions are in the row buffer
    PROF_MARK(0)
    if (SA > 0) {
      // (3)+(4) in registers: World.step, shared reward and observation of agent `warp` of env instance `lane`
      if (phys_warp) {
        const int i = warp, D = 6 * SAc;
        float px[SAc], py[SAc];
#pragma unroll
        for (int j = 0; j < SAc; ++j) { px[j] = T.sS[(4 * j + 0) * EBP + lane]; py[j] = T.sS[(4 * j + 1) * EBP + lane]; }
        const float* arow = buf + lane * RS + L.obs_sum + 5 * i;
        const float a[5] = {arow[0], arow[1], arow[2], arow[3], arow[4]};
        spread_agent_step<SAc>(Cn, Ai, i, px, py, pxi, pyi, vxi, vyi, a);
        env_bar();  // every agent has read the old positions
        T.sS[(4 * i + 0) * EBP + lane] = pxi;
        T.sS[(4 * i + 1) * EBP + lane] = pyi;
        pos_bar();  // new positions of all agents are in the state tile
#pragma unroll
        for (int j = 0; j < SAc; ++j) { px[j] = T.sS[(4 * j + 0) * EBP + lane]; py[j] = T.sS[(4 * j + 1) * EBP + lane]; }
        float* nx = buf + lane * RS + L.nx_off + i * D;
        float* ob = nxt + lane * RS + i * D;
        float* obT = sXT + (i * D) * 32 + lane;  // transposed copy [column][env]: layer 1 reads its 4 rows as one float4
        spread_obs_agent<SAc>(i, px, py, pxi, pyi, vxi, vyi, lx, ly, [&](int c, float v) { nx[c] = v; ob[c] = v; obT[c * 32] = v; });
      } else {
        if (s + 1 < R.steps) draw_noise(s + 1, tid - 32 * SAc, NTB - 32 * SAc);
        if (rew_warp) {
          // Scenario.reward of agent i (shared: every agent receives the sum) from the new positions
          const int i = warp - RW0;
          pos_bar();
          float px[SAc], py[SAc];
#pragma unroll
          for (int j = 0; j < SAc; ++j) { px[j] = T.sS[(4 * j + 0) * EBP + lane]; py[j] = T.sS[(4 * j + 1) * EBP + lane]; }
          const float pxn = T.sS[(4 * i + 0) * EBP + lane], pyn = T.sS[(4 * i + 1) * EBP + lane];
          sPart[i * 32 + lane] = spread_landmark_min<SAc>(px, py, lxi, lyi);
          sPart[(SAc + i) * 32 + lane] = __int_as_float(spread_collisions<SAc>(Cn, Ai, px, py, pxn, pyn));
          rew_bar();
          float m[SAc];
          int cnt[SAc];
#pragma unroll
          for (int j = 0; j < SAc; ++j) { m[j] = sPart[j * 32 + lane]; cnt[j] = __float_as_int(sPart[(SAc + j) * 32 + lane]); }
          const float msum = spread_reward_sum<SAc>(m, cnt);
          buf[lane * RS + L.rw_off + i] = msum;
          ret_reg += msum;
        }
      }
      PROF_MARK(1)
    } else {
    // (3) World.step and rewards (both CTA-collective, synchronised on return)
    env_physics<float, REB>(P, T, nE);
    PROF_MARK(1)
    env_flags_rewards<float, REB, true>(P, T, nE);
    PROF_MARK(2)
    // (4) next observations: into this row's next_obs columns and into the next row buffer's obs columns
    for (int c = lane; c < L.obs_sum; c += 32) {
      const ObsCol d = sCols[c];
      for (int ee = warp; ee < nE; ee += nwarps) {
        const float v = env_obs_value<float, REB>(T, d, ee);
        buf[ee * RS + L.nx_off + c] = v;
        nxt[ee * RS + c] = v;
      }
    }
    for (int idx = tid; idx < nE * A; idx += NTB) {
      const int ee = idx / A, ii = idx - ee * A;
      const float r = env_reward_out<float, REB>(P, T, ee, ii);
      buf[ee * RS + L.rw_off + ii] = r;
      sRet[ii * EBP + ee] += r;
    }
    }
    // (5) hand the finished rows to the TMA engine: generic-proxy writes -> async proxy, then one bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    PROF_MARK(3)
    if (tid == 0) {
      const long long r0 = ring_row;
      ring_row += R.E;
      if (ring_row >= R.capacity) ring_row -= R.capacity;  // capacity >= E * steps (checked on the host)
      const long long first = min((long long)nE, R.capacity - r0);  // rows before the ring wraps
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(R.ring + r0 * RS), "r"(smem_u32(buf)),
                   "r"((uint32_t)(first * RS * 4))
                   : "memory");
      if (first < nE)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(R.ring), "r"(smem_u32(buf + first * RS)),
                     "r"((uint32_t)((nE - first) * RS * 4))
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    PROF_MARK(4)
  }
#ifdef MDP_EPISODE_PROF
  if (R.ep_return && blockIdx.x == 0 && tid == 0)
    for (int k = 0; k < 10; ++k) R.ep_return[(size_t)R.E * A + k] = (float)prof_t[k];  // caller over-allocates ep_return by 16 floats
#endif

  // ---- epilogue: optional reset_world, then hand state and observations back ----------------------------
  float* fin = sRow + (R.steps & 1) * TM * RS;  // obs_{T} lives in the obs columns of the next buffer
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if (phys_warp) {  // registers -> state tile (positions are current there already)
    T.sS[(4 * warp + 2) * EBP + lane] = vxi;
    T.sS[(4 * warp + 3) * EBP + lane] = vyi;
  }
  if (rew_warp) sRet[(warp - RW0) * EBP + lane] = ret_reg;
  __syncthreads();
  if (R.reset_after) {
    for (int idx = tid; idx < P.scomp * REB; idx += NTB) {
      const int comp = idx / REB, e = idx % REB;
      T.sS[comp * EBP + e] = env_reset_value<float>(P, comp, e0 + e, R.env_seed, episode, R.lm_lo, R.lm_hi);
    }
    __syncthreads();
    env_flags_rewards<float, REB, false>(P, T, nE);
    for (int c = lane; c < L.obs_sum; c += 32) {
      const ObsCol d = sCols[c];
      for (int ee = warp; ee < nE; ee += nwarps) fin[ee * RS + c] = env_obs_value<float, REB>(T, d, ee);
    }
    __syncthreads();
  }
  env_store_state<float, REB>(P, T, (float*)R.state, R.E, e0, nE, R.reset_after != 0);
  for (int ee = warp; ee < nE; ee += nwarps)
    for (int c = lane; c < OS; c += 32) R.obs[(size_t)(e0 + ee) * OS + c] = (c < L.obs_sum) ? fin[ee * RS + c] : 0.f;
  if (R.ep_return)
    for (int idx = tid; idx < nE * A; idx += NTB) {
      const int ee = idx / A, ii = idx - ee * A;
      R.ep_return[(size_t)(e0 + ee) * A + ii] += sRet[ii * EBP + ee];
    }
}

}  // namespace mdp

using namespace mdp;


static int rollout_launch(mdp_env* env, mdp_core* core, int32_t E, void* state, float* obs, float* ring, int64_t ring_capacity,
                          int32_t ring_row_stride, int64_t ring_cursor, int32_t steps, int32_t episodes, uint64_t seed,
                          uint64_t counter, int32_t reset_after, uint64_t env_seed, uint64_t episode, float* ep_return,
                          void* stream);

extern "C" int mdp_rollout_episode(mdp_env* env, mdp_core* core, int32_t E, void* state, float* obs, float* ring,
                                   int64_t ring_capacity, int32_t ring_row_stride, int64_t ring_cursor, int32_t steps,
                                   uint64_t seed, uint64_t counter, int32_t reset_after, uint64_t env_seed,
                                   uint64_t episode, float* ep_return, void* stream) {
  return rollout_launch(env, core, E, state, obs, ring, ring_capacity, ring_row_stride, ring_cursor, steps, 1, seed, counter,
                        reset_after, env_seed, episode, ep_return, stream);
}

extern "C" int mdp_rollout_episodes(mdp_env* env, mdp_core* core, int32_t E, void* state, float* obs, float* ring,
                                    int64_t ring_capacity, int32_t ring_row_stride, int64_t ring_cursor, int32_t steps,
                                    int32_t episodes, uint64_t seed, uint64_t counter, uint64_t env_seed, uint64_t episode,
                                    float* ep_return, void* stream) {
  MDP_REQUIRE(episodes > 0, "mdp_rollout_episodes: episodes %d", episodes);
  int rc = rollout_launch(env, core, E, state, obs, ring, ring_capacity, ring_row_stride, ring_cursor, steps, episodes, seed, counter,
                          1, env_seed, episode, ep_return, stream);
  if (rc != MDP_ENOTSUP || episodes == 1) return rc;
  // kernels without the in-kernel episode loop: one launch per episode (the ring cursor / counters advance like the kernel's)
  for (int ep = 0; ep < episodes; ++ep) {
    rc = rollout_launch(env, core, E, state, obs, ring, ring_capacity, ring_row_stride,
                        (ring_cursor + (int64_t)ep * steps * E) % ring_capacity, steps, 1, seed, counter + (uint64_t)ep * steps, 1,
                        env_seed, episode + ep, ep_return, stream);
    if (rc) return rc;
  }
  return MDP_OK;
}

static int rollout_launch(mdp_env* env, mdp_core* core, int32_t E, void* state, float* obs, float* ring, int64_t ring_capacity,
                          int32_t ring_row_stride, int64_t ring_cursor, int32_t steps, int32_t episodes, uint64_t seed,
                          uint64_t counter, int32_t reset_after, uint64_t env_seed, uint64_t episode, float* ep_return,
                          void* stream) {
  MDP_REQUIRE(env && core && core->d_agents, "mdp_rollout_episode: env/core not ready");
  MDP_REQUIRE(state && obs && ring && E > 0 && steps > 0 && ring_capacity >= (int64_t)E * steps * episodes,
              "mdp_rollout_episode: bad argument (E %d, steps %d x %d, capacity %lld)", E, steps, episodes, (long long)ring_capacity);
  const EnvParams& P = env->P;
  MDP_REQUIRE(P.A == core->cfg.n_agents, "mdp_rollout_episode: env has %d agents, core %d", P.A, core->cfg.n_agents);
  for (int i = 0; i < P.A; ++i)
    MDP_REQUIRE(env->dims.obs_dim[i] == core->cfg.obs_dim[i] && env->dims.act_dim[i] == core->cfg.act_dim[i],
                "mdp_rollout_episode: agent %d dims differ between env and core", i);
  mdp_ring_layout lay;
  int rc = mdp_ring_make_layout(P.A, env->dims.obs_dim, env->dims.act_dim, &lay);
  if (rc) return rc;
  MDP_REQUIRE(lay.row_stride == ring_row_stride, "mdp_rollout_episode: ring_row_stride %d != layout %d", ring_row_stride, lay.row_stride);
  rc = env_ensure_cols(env);
  if (rc) return rc;
  RolloutArgs R;
  R.E = E; R.steps = steps; R.reset_after = reset_after; R.episodes = episodes;
  R.state = state; R.obs = obs; R.ring = ring;
  R.capacity = ring_capacity; R.cursor = ring_cursor;
  R.seed = seed; R.counter = counter; R.env_seed = env_seed; R.episode = episode;
  R.lm_lo = env->reset_lo_lm; R.lm_hi = env->reset_hi_lm;
  R.ctl = env->ctl ? env->ctl : core->ctl;
  R.ep_return = ep_return;
  cudaStream_t st = (cudaStream_t)stream;
  // tcgen05 actor tiles (mdp_rollout_tc.cu) unless the core is pinned to the SIMT kernels (mdp_core_set_tensor_cores(core, -1))
  if (core->tc_mode >= 0) {
    rc = rollout_episode_tc(env, core, lay, R, st);
    if (rc != MDP_ENOTSUP) return rc;
  }
  if (episodes != 1) return MDP_ENOTSUP;  // the caller loops (mdp_rollout_episodes)
  if (env->cfg.state_f64) return fail(MDP_ENOTSUP, "mdp_rollout_episode: float64 state needs the tensor-core episode kernel (simple_spread, 2-4 agents, 64 units)");
  const int U = core->cfg.num_units, HP = U + 4;
  size_t wts = 0;
  for (int i = 0; i < P.A; ++i) wts += actor_net_floats(core->cfg.obs_dim[i], U, core->cfg.act_dim[i]);
  auto r4 = [](size_t x) { return (x + 3) & ~(size_t)3; };
  const size_t limit = 200 * 1024;
  // agent groups of 256 threads running concurrently: as many as fit (<= 1024 threads, <= ~200 KB smem)
  int NG = 0;
  bool resident = false;
  size_t smem = 0;
  for (int ng = P.A < 4 ? P.A : 4; ng >= 1 && NG == 0; --ng) {
    const size_t base = 2 * r4((size_t)ng * TM * HP) + r4((size_t)ng * TM * KPAD) + r4(2 * (size_t)TM * lay.row_stride) +
                        r4((size_t)P.A * (REB + 1)) + r4(6 * 4 * 4 * 32) + r4(8 * 256) + r4(2 * 8 * 32) + r4(MDP_MAX_AGENTS + 1) +
                        r4(2 * (size_t)P.obs_stride) + r4(EnvTile<float, REB>::bytes(P.scomp, P.A, P.act_stride, false) / 4);
    const size_t smem_res = (base + wts + 16) * 4, smem_str = (base + (size_t)ng * KC * U + 16) * 4;
    // the physics phase needs one thread per (env, agent); resident groups are 128 threads, streaming groups 256
    if (smem_res <= limit && ng * 128 >= REB * P.A) { NG = ng; resident = true; smem = smem_res; }
    else if (smem_str <= limit && ng * NT >= REB * P.A) { NG = ng; resident = false; smem = smem_str; }
  }
  if (NG == 0)
    return fail(MDP_ENOTSUP, "mdp_rollout_episode: %d agents x %d observation floats do not fit one CTA's shared memory",
                P.A, P.obs_stride);
  CoreDev d = core_dev_for_rollout(core);
  auto go = [&](auto kern) -> int {
    if (smem > 48 * 1024) MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(E, REB), NG * (resident ? 128 : NT), smem, st>>>(d, P, env->d_cols, lay, R, NG);
    return check_launch("k_rollout_episode");
  };
  if (U == 64 && resident && P.scenario == MDP_SIMPLE_SPREAD && !env->force_generic && !env->cfg.state_f64) {
    switch (P.A) {  // register-resident env phase (as mdp_env_step picks k_env_step_spread)
      case 2: return go(k_rollout_episode<64, true, 2>);
      case 3: return go(k_rollout_episode<64, true, 3>);
      case 4: return go(k_rollout_episode<64, true, 4>);
      default: break;  // A >= 5: 6*A*A observation floats per thread no longer fit the register file
    }
  }
  if (U == 64) return resident ? go(k_rollout_episode<64, true, 0>) : go(k_rollout_episode<64, false, 0>);
  return resident ? go(k_rollout_episode<128, true, 0>) : go(k_rollout_episode<128, false, 0>);
}
