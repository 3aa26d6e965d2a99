// Persistent episode kernel with the actor MLP on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract as k_rollout_episode (mdp_rollout.cu): the lockstep rollout of experiments/train.py:110-133
//   action_n = [agent.action(obs)]          maddpg/trainer/maddpg.py:151-152 (mlp_model train.py:39-46 + SoftCategoricalPd.sample
//                                           distributions.py:264-266)
//   new_obs, rew, done = env.step(action_n) multiagent.environment.MultiAgentEnv.step
//   agent.experience(...)                   maddpg/trainer/maddpg.py:154-156 -> ReplayBuffer.add (replay_buffer.py:25-32)
//   env.reset() every max_episode_len steps train.py:127-129
// for `steps` steps in ONE launch, a CTA owning 32 env instances of simple_spread (A = 2..4 agents, num_units = 64).
//
// GEMM orientation.  A CTA has only 32 rows (env instances) per agent, far below the UMMA M of 128, so the hidden layers are
// computed TRANSPOSED: D[unit][env] = W^T[unit][k] * X^T[k][env].
//   * A operand = the weights, resident in TENSOR MEMORY for the whole episode (lane = output unit, one 32-bit column per k):
//     two agents share one 128-lane image (agent 2p in lanes 0-63, agent 2p+1 in lanes 64-127), hi | lo halves of the 3xTF32
//     split side by side.  An MMA for agent i therefore also produces 64 garbage lanes (the pair partner's weights applied to
//     agent i's activations) -- the UMMA cost floor is max(M, 128) * N / 256 cycles anyway.
//   * B operand = the activations of the CTA's 32 env instances, a K-major SWIZZLE_128B image [32 env][32 k] per 32-wide k panel
//     in shared memory, written by the epilogue threads (hi = the fp32 value itself -- kind::tf32 reads its upper 19 bits --
//     and lo = x - trunc(x), exact).  Every product is issued as lo*hi + hi*lo + hi*hi (fp32 accumulate in TMEM): measured
//     fp32-class, ~2e-6 relative (tools/umma_probe.cu).
//   * D (32 fp32 columns per agent in TMEM) comes back with tcgen05.ld 32x32b: thread = one hidden unit, 32 env instances in
//     registers, bias + ReLU + split, stores to the next layer's image.  N = 32: 16 cycles per MMA.
// The 64 x 5 output head and the Gumbel-softmax run on two threads per (env, agent) row straight from the fp32 h2 image, and
// the SAME thread continues into World.step for its (env, agent) -- an agent's chain never waits for the other agents'
// actions, only for their positions.
//
// Warp roles (A agents): warps [0, 2A) actor warps (agent w/2; TMEM lane quadrant w%4 == the quadrant its units land in),
// warps [2A, 3A) reward warps (Scenario.reward, next step's Gumbel noise, the replay-row TMA bulk store), warp 3A MMA issuer.
#include "mdp_rollout.cuh"
#include "mdp_umma.cuh"

namespace mdp {
namespace eptc {

constexpr int EB = REB;        // env instances per CTA == N of every MMA
constexpr int U = 64;          // hidden units
constexpr int IMG = EB * 128;  // bytes of one [32 env][32 float] SWIZZLE_128B panel
constexpr int W3P = 36;        // pitch of one 32-unit half of a transposed W3 row (the two halves land in disjoint banks)
constexpr int W3A = 2 * W3P;   // floats per action row of the transposed head weights

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// a lost arrival must fail loudly (~2 s), never hang the GPU
__device__ __forceinline__ void mbar_wait_b(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 1;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if ((spins & 1023u) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
}
// (no memory clobber: volatile asm statements keep their order among themselves -- barriers included --, and these never alias the
// plain C++ accesses of the same barrier interval)
__device__ __forceinline__ void st_s32(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v)); }
__device__ __forceinline__ float4 ld_s128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---- simple_spread pieces for a state precision `real` -----------------------------------------------------------------------
// float: the pinned-rounding pieces of mdp_env_dev.cuh (bit-identical to k_env_step_spread); double: the arithmetic of
// env_physics / env_flags_rewards / env_obs_value <double> (the float64 parity mode of the per-step kernel).
template <int A, typename real> struct Spread;

template <int A>
struct Spread<A, float> {
  SpreadConsts<A> Cn;
  SpreadAgentConsts Ai;
  __device__ __forceinline__ void init(const EnvParams& P, int i) { Cn = spread_consts<A>(P); Ai = spread_agent_consts(P, i); }
  __device__ __forceinline__ void step(int i, const float (&px)[A], const float (&py)[A], float& pxi, float& pyi, float& vxi, float& vyi,
                                       const float (&a)[5]) const {
    spread_agent_step<A>(Cn, Ai, i, px, py, pxi, pyi, vxi, vyi, a);
  }
  __device__ __forceinline__ float landmark_min(const float (&px)[A], const float (&py)[A], float lx, float ly) const {
    return spread_landmark_min<A>(px, py, lx, ly);
  }
  __device__ __forceinline__ int collisions(const float (&px)[A], const float (&py)[A], float pxi, float pyi) const {
    return spread_collisions<A>(Cn, Ai, px, py, pxi, pyi);
  }
  __device__ __forceinline__ float reward_sum(const float (&m)[A], const int (&cnt)[A]) const { return spread_reward_sum<A>(m, cnt); }
};

template <int A>
struct Spread<A, double> {
  double k, cf, damp, dt, size[A], sens, ms, si;
  __device__ __forceinline__ void init(const EnvParams& P, int i) {
    k = P.contact_margin; cf = P.contact_force; damp = 1.0 - P.damping; dt = P.dt;
#pragma unroll
    for (int j = 0; j < A; ++j) size[j] = P.size[j];
    sens = P.sens[i]; ms = P.max_speed[i]; si = P.size[i];
  }
  // env_physics<double>: _set_action in float32, soft contact in partner order, damping, integration
  __device__ __forceinline__ void step(int i, const double (&px)[A], const double (&py)[A], double& pxi, double& pyi, double& vxi,
                                       double& vyi, const float (&a)[5]) const {
    double fx = (double)(a[1] - a[2]), fy = (double)(a[3] - a[4]);
    fx *= sens; fy *= sens;
#pragma unroll
    for (int j = 0; j < A; ++j) {
      if (j == i) continue;
      const double dx = pxi - px[j], dy = pyi - py[j];
      const double dist = sqrt(dx * dx + dy * dy);
      const double dmin = si + size[j];
      const double z = -(dist - dmin) / k;
      if (z < -746.0) continue;
      const double pen = logaddexp0<double>(z) * k;
      fx = cf * dx / dist * pen + fx;
      fy = cf * dy / dist * pen + fy;
    }
    double wx = vxi * damp, wy = vyi * damp;
    wx += fx * dt; wy += fy * dt;
    if (ms > 0.0) {
      const double speed = sqrt(wx * wx + wy * wy);
      if (speed > ms) { wx = wx / speed * ms; wy = wy / speed * ms; }
    }
    vxi = wx; vyi = wy;
    pxi += wx * dt; pyi += wy * dt;
  }
  __device__ __forceinline__ double landmark_min(const double (&px)[A], const double (&py)[A], double lx, double ly) const {
    double best = 0.0;
#pragma unroll
    for (int q = 0; q < A; ++q) {
      const double dx = px[q] - lx, dy = py[q] - ly;
      const double d = sqrt(dx * dx + dy * dy);
      best = (q == 0 || d < best) ? d : best;
    }
    return best;
  }
  __device__ __forceinline__ int collisions(const double (&px)[A], const double (&py)[A], double pxi, double pyi) const {
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < A; ++q) {
      const double dx = px[q] - pxi, dy = py[q] - pyi;
      cnt += (sqrt(dx * dx + dy * dy) < size[q] + si) ? 1 : 0;
    }
    return cnt;
  }
  // env_flags_rewards + env_reward_out <double>: r_i = -sum_l m[l] - cnt[i]; every agent receives sum_i r_i
  __device__ __forceinline__ float reward_sum(const double (&m)[A], const int (&cnt)[A]) const {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < A; ++i) {
      double r = 0.0;
#pragma unroll
      for (int l = 0; l < A; ++l) r -= m[l];
      r -= (double)cnt[i];
      tot += r;
    }
    return (float)tot;
  }
};

// Scenario.observation of agent i (simple_spread): [vel, pos, landmarks - pos, others - pos, silent comm zeros]
// out(c, v): c known at compile time; out_rt(c, v): c depends on the (runtime) agent index
template <int A, typename real, typename Out, typename OutRt>
__device__ __forceinline__ void obs_agent(int i, const real (&px)[A], const real (&py)[A], real pxi, real pyi, real vxi, real vyi,
                                          const real (&lx)[A], const real (&ly)[A], Out&& out, OutRt&& out_rt) {
  constexpr int L = A, D = 6 * A;
  out(0, (float)vxi); out(1, (float)vyi); out(2, (float)pxi); out(3, (float)pyi);
#pragma unroll
  for (int l = 0; l < L; ++l) { out(4 + 2 * l, (float)(lx[l] - pxi)); out(5 + 2 * l, (float)(ly[l] - pyi)); }
#pragma unroll
  for (int q = 0; q < A; ++q) {
    if (q == i) continue;
    const int c = 4 + 2 * L + 2 * (q < i ? q : q - 1);
    out_rt(c, (float)(px[q] - pxi));
    out_rt(c + 1, (float)(py[q] - pyi));
  }
#pragma unroll
  for (int c = 4 + 2 * L + 2 * (A - 1); c < D; ++c) out(c, 0.f);
}

template <int SA>
struct Cfg {
  static constexpr int A = SA, D = 6 * SA, NPAIR = (SA + 1) / 2;
  static constexpr int K1 = (D + 7) / 8 * 8;  // layer-1 K padded to whole kind::tf32 MMAs (8 per instruction)
  static constexpr int NW = 3 * SA + 1, NTB = 32 * NW;
  // tensor-memory columns
  static constexpr uint32_t T_W2 = 0;                        // pair p: hi at 128 p, lo at 128 p + 64
  static constexpr uint32_t T_W1 = 128 * NPAIR;              // pair p: hi at T_W1 + 2 K1 p, lo K1 further
  static constexpr uint32_t T_D = T_W1 + 2 * K1 * NPAIR;     // agent i: 32 accumulator columns at T_D + 32 i
  static constexpr uint32_t T_USED = T_D + 32 * SA;
  static constexpr uint32_t T_COLS = T_USED <= 128 ? 128 : T_USED <= 256 ? 256 : 512;
  // shared memory (bytes from the 1024-aligned base)
  static constexpr uint32_t OFF_OBS = 0;                     // agent i: hi | lo panels of its observation tile
  static constexpr uint32_t OFF_H = OFF_OBS + SA * 2 * IMG;  // agent i: hi panel 0, hi panel 1, lo panel 0, lo panel 1
  static constexpr uint32_t OFF_REST = OFF_H + SA * 4 * IMG;
};

__host__ __device__ inline size_t rest_floats(int A, int RS, int OS, size_t env_tile_bytes) {
  auto r4 = [](size_t x) { return (x + 3) & ~(size_t)3; };
  return r4(2 * (size_t)EB * RS) + r4((size_t)A * (5 * W3A + 8)) + r4(2 * (size_t)A * EB * 8) + r4(2 * (size_t)A * EB) +
         r4((size_t)A * (EB + 1)) + r4(2 * (size_t)OS) + r4(env_tile_bytes / 4) + r4(2 * 4 * (size_t)A) + 16;
}

template <int SA, typename real>
__global__ void __launch_bounds__(Cfg<SA>::NTB, 1) k_rollout_episode_tc(CoreDev C, EnvParams P, const ObsCol* __restrict__ cols,
                                                                         mdp_ring_layout L, RolloutArgs R) {
  using CF = Cfg<SA>;
  constexpr int A = SA, D = CF::D, K1 = CF::K1, NTB = CF::NTB, EBP = EB + 1;
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint32_t tmem_slot;
  unsigned char* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int OS = P.obs_stride, RS = L.row_stride;
  const int e0 = blockIdx.x * EB;
  const int nE = min(EB, R.E - e0);

  // ---- shared memory carve-up ---------------------------------------------------------------------------------------------
  SmemCarve sm(smem + CF::OFF_REST);
  float* sRow = sm.take(2 * EB * RS);       // two replay-row buffers (ping-pong), assembled in place
  float* sW3 = sm.take(A * (5 * W3A + 8));  // per agent: W3 transposed [5][2 x 36] + b3
  float* sNoise = sm.take(2 * A * EB * 8);  // Gumbel noise, [step parity][agent][env][8]
  float* sPart = sm.take(2 * A * EB);       // reward partials (landmark minima, collision counts)
  float* sRet = sm.take(A * EBP);
  ObsCol* sCols = reinterpret_cast<ObsCol*>(sm.take(2 * OS));
  real* sEnvBase = reinterpret_cast<real*>(sm.take((int)(EnvTile<real, EB>::bytes(P.scomp, A, P.act_stride, false) / 4)));
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm.take(2 * 4 * A));
  unsigned long long* bar_obs = bars;          // [A] count 2: agent i's observation images of the next step are written
  unsigned long long* bar_l1 = bars + A;       // [A] tcgen05.commit: layer-1 accumulator of agent i complete
  unsigned long long* bar_h1 = bars + 2 * A;   // [A] count 2: agent i's h1 images are written
  unsigned long long* bar_l2 = bars + 3 * A;   // [A] tcgen05.commit: layer-2 accumulator complete
  EnvTile<real, EB> T;
  T.carve(sEnvBase, P, sRow + L.obs_sum);
  T.ASP = RS;

  unsigned long long counter = R.counter, episode = R.episode;
  long long cursor = R.cursor;
  if (R.ctl) {
    counter += R.ctl[0];
    cursor = (cursor + (long long)R.ctl[1]) % R.capacity;
    episode += R.ctl[2];
  }

  // ---- prologue ---------------------------------------------------------------------------------------------------------------
  if (warp == CF::NW - 1) umma::tmem_alloc(&tmem_slot, CF::T_COLS);
  if (tid == 0) {
    for (int i = 0; i < A; ++i) {
      mbar_init(&bar_obs[i], 2); mbar_init(&bar_l1[i], 1); mbar_init(&bar_h1[i], 2); mbar_init(&bar_l2[i], 1);
    }
  }
  env_load_state<real, EB>(P, T, (const real*)R.state, R.E, e0, nE);
  for (int i = tid; i < 2 * EB * RS; i += NTB) sRow[i] = 0.f;  // padding / done columns stay zero for good
  for (int i = tid; i < (int)(CF::OFF_H / 16); i += NTB) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = tid; c < OS; c += NTB) sCols[c] = cols[c];
  for (int idx = tid; idx < A * EBP; idx += NTB) sRet[idx] = 0.f;
  for (int idx = tid; idx < 2 * A * EB * 8; idx += NTB) sNoise[idx] = 0.f;
  for (int idx = tid; idx < A * (5 * W3A + 8); idx += NTB) sW3[idx] = 0.f;
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  // the actors' weights: W1^T / W2^T into tensor memory (hi | lo), W3 transposed into shared memory
  if (warp < 4) {
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    const int half = warp >> 1, u = 32 * (warp & 1) + lane;
#pragma unroll 1
    for (int p = 0; p < CF::NPAIR; ++p) {
      const int ag = 2 * p + half;
      const MlpW w = C.agents[ag < A ? ag : 0].net[MDP_NET_P];
      const bool live = ag < A;
#pragma unroll 1
      for (int k0 = 0; k0 < U; k0 += 16) {
        float hi[16], lo[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const float x = live ? w.W2[(k0 + kk) * U + u] : 0.f;
          umma::split_tf32(x, hi[kk], lo[kk]);
        }
        umma::tmem_st16(tbase + lane_base + CF::T_W2 + 128 * p + k0, hi);
        umma::tmem_st16(tbase + lane_base + CF::T_W2 + 128 * p + 64 + k0, lo);
      }
#pragma unroll 1
      for (int k0 = 0; k0 < K1; k0 += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const float x = (live && k0 + kk < D) ? w.W1[(k0 + kk) * U + u] : 0.f;
          umma::split_tf32(x, hi[kk], lo[kk]);
        }
        umma::tmem_st8(tbase + lane_base + CF::T_W1 + 2 * K1 * p + k0, hi);
        umma::tmem_st8(tbase + lane_base + CF::T_W1 + 2 * K1 * p + K1 + k0, lo);
      }
    }
    umma::tmem_st_wait();
  }
  for (int idx = tid; idx < A * U * 5; idx += NTB) {
    const int i = idx / (U * 5), rem = idx - i * (U * 5), u = rem / 5, a = rem - u * 5;
    sW3[i * (5 * W3A + 8) + a * W3A + W3P * (u >> 5) + (u & 31)] = C.agents[i].net[MDP_NET_P].W3[rem];
  }
  for (int idx = tid; idx < A * 5; idx += NTB) {
    const int i = idx / 5, a = idx - i * 5;
    sW3[i * (5 * W3A + 8) + 5 * W3A + a] = C.agents[i].net[MDP_NET_P].b3[a];
  }
  // current observations: obs_t columns of the first row buffer + the layer-1 operand images
  for (int idx = tid; idx < EB * L.obs_sum; idx += NTB) {
    const int ee = idx / L.obs_sum, c = idx - ee * L.obs_sum;
    const float v = ee < nE ? R.obs[(size_t)(e0 + ee) * OS + c] : 0.f;
    sRow[ee * RS + c] = v;
    const int i = c / D, cc = c - i * D;
    unsigned char* img = smem + CF::OFF_OBS + i * 2 * IMG;
    *reinterpret_cast<float*>(img + umma::sw128_off(ee, cc)) = v;
    *reinterpret_cast<float*>(img + IMG + umma::sw128_off(ee, cc)) = tf32_lo(v);
  }
  // Gumbel noise of step 0 (afterwards the reward warps draw one step ahead)
  for (int idx = tid; idx < A * EB * 5; idx += NTB) {
    const int i = idx / (EB * 5), rem = idx - i * (EB * 5), r = rem / 5, a = rem - r * 5;
    if (r < nE) sNoise[(i * EB + r) * 8 + a] = gumbel_from_u(philox_u(R.seed, counter + 1ull, (uint32_t)i, (long long)e0 + r, a));
  }
  umma::fence_async_smem();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();

  auto phys_bar = [] { asm volatile("bar.sync 8, %0;" ::"r"(64 * A) : "memory"); };   // actor warps: old positions are read
  auto pos_bar = [] { asm volatile("bar.sync 9, %0;" ::"r"(96 * A) : "memory"); };    // actor + reward warps: new positions are in the tile
  auto rew_bar = [] { asm volatile("bar.sync 10, %0;" ::"r"(32 * A) : "memory"); };   // reward warps only
  auto row_arrive = [] { asm volatile("bar.arrive 11, %0;" ::"r"(96 * A) : "memory"); };
  auto row_sync = [] { asm volatile("bar.sync 11, %0;" ::"r"(96 * A) : "memory"); };  // the step's replay rows are complete

  if (warp < 2 * A) {
    // =================================== actor warps =========================================================================
    const int i = warp >> 1, h = warp & 1;
    const int e = 16 * h + (lane >> 1), par = lane & 1, e7 = e & 7;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const uint32_t tD = tbase + lane_base + CF::T_D + 32 * i;
    const MlpW wg = C.agents[i].net[MDP_NET_P];
    const float b1u = wg.b1[32 * h + lane], b2u = wg.b2[32 * h + lane];
    // unit role: element (env n, unit 32 h + lane) of the hidden images
    const uint32_t hU = sbase + CF::OFF_H + i * 4 * IMG + h * IMG;
    uint32_t xo[8];
#pragma unroll
    for (int n7 = 0; n7 < 8; ++n7) xo[n7] = (uint32_t)((((lane >> 2) ^ n7) << 4) | ((lane & 3) << 2));
    // row role: (env e, half `par` of the units)
    const uint32_t hR = sbase + CF::OFF_H + i * 4 * IMG + par * IMG + (uint32_t)((e >> 3) << 10) + (uint32_t)(e7 << 7);
    const float* w3 = sW3 + i * (5 * W3A + 8) + W3P * par;
    float b3[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) b3[a] = sW3[i * (5 * W3A + 8) + 5 * W3A + a];
    // observation writes: parity 0 -> the replay row (next_obs of this step, obs_t of the next), parity 1 -> hi | lo images
    const uint32_t rowoff = (uint32_t)((e >> 3) << 10) + (uint32_t)(e7 << 7);
    uint32_t cofs[(D + 3) / 4];
#pragma unroll
    for (int j = 0; j < (D + 3) / 4; ++j) cofs[j] = par ? (uint32_t)((j ^ e7) << 4) : (uint32_t)(j << 4);
    const int x7 = par ? e7 : 0;
    const uint32_t oImg = sbase + CF::OFF_OBS + i * 2 * IMG + rowoff;
    // state of (env e, agent i) in registers for the whole episode (both lanes of a row keep identical copies)
    Spread<A, real> S;
    S.init(P, i);
    real pxi = T.sS[(4 * i + 0) * EBP + e], pyi = T.sS[(4 * i + 1) * EBP + e];
    real vxi = T.sS[(4 * i + 2) * EBP + e], vyi = T.sS[(4 * i + 3) * EBP + e];
    real lx[A], ly[A];
#pragma unroll
    for (int l = 0; l < A; ++l) { lx[l] = T.sS[(4 * A + 2 * l + 0) * EBP + e]; ly[l] = T.sS[(4 * A + 2 * l + 1) * EBP + e]; }

    for (int s = 0; s < R.steps; ++s) {
      float* buf = sRow + (s & 1) * EB * RS;
      float* nxt = sRow + ((s & 1) ^ 1) * EB * RS;
      const uint32_t ph = (uint32_t)(s & 1);
      // ---- epilogue 1: h1 = relu(acc + b1) -> hi | lo images (B operand of layer 2)
      mbar_wait_b(&bar_l1[i], ph);
      umma::fence_after();
      {
        float v[32];
        umma::tmem_ld32(tD, v);
#pragma unroll
        for (int n = 0; n < 32; ++n) {
          const float x = fmaxf(v[n] + b1u, 0.f);
          const uint32_t a = hU + (uint32_t)(((n >> 3) << 10) + ((n & 7) << 7)) + xo[n & 7];
          st_s32(a, x);
          st_s32(a + 2 * IMG, tf32_lo(x));
        }
      }
      umma::fence_before();
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_h1[i]);
      // ---- epilogue 2: h2 = relu(acc + b2) -> fp32 image (read by the head below)
      mbar_wait_b(&bar_l2[i], ph);
      umma::fence_after();
      {
        float v[32];
        umma::tmem_ld32(tD, v);
#pragma unroll
        for (int n = 0; n < 32; ++n) {
          const float x = fmaxf(v[n] + b2u, 0.f);
          st_s32(hU + (uint32_t)(((n >> 3) << 10) + ((n & 7) << 7)) + xo[n & 7], x);
        }
      }
      umma::fence_before();
      asm volatile("bar.sync %0, 64;" ::"r"(1 + i) : "memory");  // both warps of the agent: h2 rows are complete
      // ---- output head on two threads per row (32 units each), then Gumbel-softmax (distributions.py:264-266)
      float act[5];
      {
        float sa[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int jj = j ^ (par << 2);  // the two lanes of a row walk the 16-byte chunks in different orders: no bank conflicts
          const float4 hv = ld_s128(hR + (uint32_t)((jj ^ e7) << 4));
#pragma unroll
          for (int a = 0; a < 5; ++a) {
            const float4 wv = *reinterpret_cast<const float4*>(w3 + a * W3A + 4 * jj);
            sa[a] = fmaf(hv.x, wv.x, sa[a]);
            sa[a] = fmaf(hv.y, wv.y, sa[a]);
            sa[a] = fmaf(hv.z, wv.z, sa[a]);
            sa[a] = fmaf(hv.w, wv.w, sa[a]);
          }
        }
        const float* nz = sNoise + ((s & 1) * A * EB + i * EB + e) * 8;
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const float o = __shfl_xor_sync(0xffffffffu, sa[a], 1);
          act[a] = ((sa[a] + o) + b3[a]) + nz[a];  // units [0, 32) + units [32, 64): the same value on both lanes
          m = fmaxf(m, act[a]);
        }
        float sum = 0.f;
#pragma unroll
        for (int a = 0; a < 5; ++a) { act[a] = expf(act[a] - m); sum += act[a]; }
#pragma unroll
        for (int a = 0; a < 5; ++a) act[a] = act[a] / sum;
        if (par == 0 && e < nE) {
          float* arow = buf + e * RS + L.obs_sum + 5 * i;
#pragma unroll
          for (int a = 0; a < 5; ++a) arow[a] = act[a];
        }
      }
      // ---- World.step for (env e, agent i): needs the other agents' OLD positions only
      {
        real px[A], py[A];
#pragma unroll
        for (int j = 0; j < A; ++j) { px[j] = T.sS[(4 * j + 0) * EBP + e]; py[j] = T.sS[(4 * j + 1) * EBP + e]; }
        S.step(i, px, py, pxi, pyi, vxi, vyi, act);
        phys_bar();
        if (par == 0) { T.sS[(4 * i + 0) * EBP + e] = pxi; T.sS[(4 * i + 1) * EBP + e] = pyi; }
        pos_bar();
#pragma unroll
        for (int j = 0; j < A; ++j) { px[j] = T.sS[(4 * j + 0) * EBP + e]; py[j] = T.sS[(4 * j + 1) * EBP + e]; }
        // ---- Scenario.observation: parity 0 writes the replay-row copies, parity 1 the layer-1 operand images of step s+1
        const uint32_t p1 = par ? oImg : smem_u32(buf + e * RS + L.nx_off + i * D);
        const uint32_t p2 = par ? oImg + IMG : smem_u32(nxt + e * RS + i * D);
        obs_agent<A, real>(
            i, px, py, pxi, pyi, vxi, vyi, lx, ly,
            [&](int c, float v) {
              const uint32_t o = cofs[c >> 2] + (uint32_t)((c & 3) << 2);
              st_s32(p1 + o, v);
              st_s32(p2 + o, par ? tf32_lo(v) : v);
            },
            [&](int c, float v) {
              const uint32_t o = (uint32_t)((((c >> 2) ^ x7) << 4) | ((c & 3) << 2));
              st_s32(p1 + o, v);
              st_s32(p2 + o, par ? tf32_lo(v) : v);
            });
      }
      umma::fence_async_smem();  // images -> UMMA (async proxy), replay rows -> TMA bulk store
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_obs[i]);
      row_arrive();
    }
    // registers -> state tile (positions are current there already)
    if (par == 0) { T.sS[(4 * i + 2) * EBP + e] = vxi; T.sS[(4 * i + 3) * EBP + e] = vyi; }
  } else if (warp < 3 * A) {
    // =================================== reward warps ========================================================================
    const int i = warp - 2 * A, e = lane;
    Spread<A, real> S;
    S.init(P, i);
    const real lxi = T.sS[(4 * A + 2 * i + 0) * EBP + e], lyi = T.sS[(4 * A + 2 * i + 1) * EBP + e];
    float ret_reg = 0.f;
    long long ring_row = (cursor + e0) % R.capacity;  // ring row of this CTA's first env at step s
    const bool storer = warp == 2 * A && lane == 0;
    for (int s = 0; s < R.steps; ++s) {
      float* buf = sRow + (s & 1) * EB * RS;
      if (s + 1 < R.steps && e < nE) {  // Gumbel noise of the next step, agent i
        float* nz = sNoise + (((s + 1) & 1) * A * EB + i * EB + e) * 8;
#pragma unroll
        for (int a = 0; a < 5; ++a)
          nz[a] = gumbel_from_u(philox_u(R.seed, counter + (unsigned long long)s + 2ull, (uint32_t)i, (long long)e0 + e, a));
      }
      // the bulk store of step s-1 must have finished READING the other buffer before the actor warps overwrite its obs columns
      if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      pos_bar();
      real px[A], py[A];
#pragma unroll
      for (int j = 0; j < A; ++j) { px[j] = T.sS[(4 * j + 0) * EBP + e]; py[j] = T.sS[(4 * j + 1) * EBP + e]; }
      const real mine = S.landmark_min(px, py, lxi, lyi);
      const int cn = S.collisions(px, py, T.sS[(4 * i + 0) * EBP + e], T.sS[(4 * i + 1) * EBP + e]);
      // partials through shared memory as float pairs (double: split into two words)
      real* sp = reinterpret_cast<real*>(sPart);
      if (sizeof(real) == 4) {
        sPart[i * EB + e] = (float)mine;
        sPart[(A + i) * EB + e] = __int_as_float(cn);
      } else {
        sp[i * EB + e] = mine;  // sPart holds 2 A EB floats = A EB doubles; counts go to sRet's tail below
        reinterpret_cast<int*>(sRet)[i * EBP + e] = cn;
      }
      rew_bar();
      real m[A];
      int cnt[A];
#pragma unroll
      for (int j = 0; j < A; ++j) {
        if (sizeof(real) == 4) { m[j] = (real)sPart[j * EB + e]; cnt[j] = __float_as_int(sPart[(A + j) * EB + e]); }
        else { m[j] = sp[j * EB + e]; cnt[j] = reinterpret_cast<int*>(sRet)[j * EBP + e]; }
      }
      const float rsum = S.reward_sum(m, cnt);
      if (e < nE) buf[e * RS + L.rw_off + i] = rsum;
      ret_reg += rsum;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      row_sync();
      if (storer) {
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
    }
    if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    rew_bar();
    sRet[i * EBP + e] = ret_reg;
  } else {
    // =================================== MMA issuer ==========================================================================
    constexpr uint32_t idesc = umma::idesc_tf32(128, EB, 0, 0);
    for (int s = 0; s < R.steps; ++s) {
#pragma unroll 1
      for (int i = 0; i < A; ++i) {  // layer 1: D_i = W1_i^T x obs_i^T
        if (s > 0) mbar_wait_b(&bar_obs[i], (uint32_t)((s - 1) & 1));
        umma::fence_after();
        const uint32_t a_hi = tbase + CF::T_W1 + 2 * K1 * (i >> 1), a_lo = a_hi + K1;
        const uint32_t tacc = tbase + CF::T_D + 32 * i;
        const uint64_t b_hi = umma::desc_k(sbase + CF::OFF_OBS + i * 2 * IMG, IMG, 0);
        const uint64_t b_lo = umma::desc_k(sbase + CF::OFF_OBS + i * 2 * IMG + IMG, IMG, 0);
        if (umma::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < K1 / 8; ++ks) {
            umma::mma_tf32_ta(tacc, a_lo + 8 * ks, b_hi + 2 * ks, idesc, ks > 0 ? 1u : 0u);
            umma::mma_tf32_ta(tacc, a_hi + 8 * ks, b_lo + 2 * ks, idesc, 1u);
            umma::mma_tf32_ta(tacc, a_hi + 8 * ks, b_hi + 2 * ks, idesc, 1u);
          }
          umma::commit(&bar_l1[i]);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int i = 0; i < A; ++i) {  // layer 2: D_i = W2_i^T x h1_i^T
        mbar_wait_b(&bar_h1[i], (uint32_t)(s & 1));
        umma::fence_after();
        const uint32_t a_hi = tbase + CF::T_W2 + 128 * (i >> 1), a_lo = a_hi + 64;
        const uint32_t tacc = tbase + CF::T_D + 32 * i;
        const uint64_t b_hi = umma::desc_k(sbase + CF::OFF_H + i * 4 * IMG, IMG, 0);
        const uint64_t b_lo = umma::desc_k(sbase + CF::OFF_H + i * 4 * IMG + 2 * IMG, IMG, 0);
        if (umma::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < U / 8; ++ks) {
            const uint32_t bo = (uint32_t)(ks >> 2) * (IMG >> 4) + (uint32_t)(ks & 3) * 2u;
            umma::mma_tf32_ta(tacc, a_lo + 8 * ks, b_hi + bo, idesc, ks > 0 ? 1u : 0u);
            umma::mma_tf32_ta(tacc, a_hi + 8 * ks, b_lo + bo, idesc, 1u);
            umma::mma_tf32_ta(tacc, a_hi + 8 * ks, b_hi + bo, idesc, 1u);
          }
          umma::commit(&bar_l2[i]);
        }
        __syncwarp();
      }
    }
  }

  // ---- epilogue: optional reset_world, then hand state and observations back --------------------------------------------------
  umma::fence_before();
  __syncthreads();
  if (warp == CF::NW - 1) {
    umma::fence_after();
    umma::tmem_free(tbase, CF::T_COLS);
  }
  const int nwarps = NTB >> 5;
  float* fin = sRow + (R.steps & 1) * EB * RS;  // obs_T lives in the obs columns of the next buffer
  if (R.reset_after) {
    for (int idx = tid; idx < P.scomp * EB; idx += NTB) {
      const int comp = idx / EB, e = idx % EB;
      T.sS[comp * EBP + e] = env_reset_value<real>(P, comp, e0 + e, R.env_seed, episode, R.lm_lo, R.lm_hi);
    }
    __syncthreads();
    env_flags_rewards<real, EB, false>(P, T, nE);
    for (int c = lane; c < L.obs_sum; c += 32) {
      const ObsCol d = sCols[c];
      for (int ee = warp; ee < nE; ee += nwarps) fin[ee * RS + c] = env_obs_value<real, EB>(T, d, ee);
    }
    __syncthreads();
  }
  env_store_state<real, EB>(P, T, (real*)R.state, R.E, e0, nE, R.reset_after != 0);
  for (int ee = warp; ee < nE; ee += nwarps)
    for (int c = lane; c < OS; c += 32) R.obs[(size_t)(e0 + ee) * OS + c] = (c < L.obs_sum) ? fin[ee * RS + c] : 0.f;
  if (R.ep_return)
    for (int idx = tid; idx < nE * A; idx += NTB) {
      const int ee = idx / A, ii = idx - ee * A;
      R.ep_return[(size_t)(e0 + ee) * A + ii] += sRet[ii * EBP + ee];
    }
}

template <int SA, typename real>
int launch(const CoreDev& d, mdp_env* env, const mdp_ring_layout& lay, const RolloutArgs& R, cudaStream_t st) {
  using CF = Cfg<SA>;
  const EnvParams& P = env->P;
  const size_t smem = 1024 + CF::OFF_REST +
                      4 * rest_floats(SA, lay.row_stride, P.obs_stride, EnvTile<real, EB>::bytes(P.scomp, P.A, P.act_stride, false));
  if (smem > 227 * 1024) return fail(MDP_ENOTSUP, "mdp_rollout_episode: shared memory");
  auto kern = k_rollout_episode_tc<SA, real>;
  MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<cdiv(R.E, EB), CF::NTB, smem, st>>>(d, P, env->d_cols, lay, R);
  return check_launch("k_rollout_episode_tc");
}

}  // namespace eptc

// simple_spread with 2..4 agents, num_units 64, every agent Discrete(5); float32 or float64 state
int rollout_episode_tc(mdp_env* env, mdp_core* core, const mdp_ring_layout& lay, const RolloutArgs& R, cudaStream_t st) {
  const EnvParams& P = env->P;
  if (core->cfg.num_units != 64 || P.scenario != MDP_SIMPLE_SPREAD || P.A < 2 || P.A > 4 || env->force_generic)
    return MDP_ENOTSUP;
  for (int i = 0; i < P.A; ++i)
    if (core->cfg.n_heads[i] != 1 || core->cfg.act_dim[i] != 5) return MDP_ENOTSUP;
  const CoreDev d = core_dev_for_rollout(core);
  const bool f64 = env->cfg.state_f64 != 0;
  switch (P.A) {
    case 2: return f64 ? eptc::launch<2, double>(d, env, lay, R, st) : eptc::launch<2, float>(d, env, lay, R, st);
    case 3: return f64 ? eptc::launch<3, double>(d, env, lay, R, st) : eptc::launch<3, float>(d, env, lay, R, st);
    default: return f64 ? eptc::launch<4, double>(d, env, lay, R, st) : eptc::launch<4, float>(d, env, lay, R, st);
  }
}

}  // namespace mdp
