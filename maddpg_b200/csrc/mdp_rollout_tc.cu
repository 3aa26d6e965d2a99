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
//     in shared memory, written by the epilogue threads as hi = rna_tf32(x) and lo = x - hi (exact).  Every product is issued
//     as lo*hi + hi*lo + hi*hi (fp32 accumulate in TMEM): fp32-class, actions within ~2e-7 of the fp32 SIMT kernels.
//   * D (32 fp32 columns per agent in TMEM) comes back with tcgen05.ld 32x32b: thread = one hidden unit, 32 env instances in
//     registers, bias + ReLU + split, stores to the next layer's image.  N = 32: 16 cycles per MMA.
// The 64 x 5 output head and the Gumbel-softmax run on two threads per (env, agent) row straight from the fp32 h2 image, and
// the SAME thread continues into World.step for its (env, agent) -- an agent's chain never waits for the other agents'
// actions, only for their positions.  The two threads of a row then write the x / y halves of the new observation straight
// into the next step's layer-1 operand images; the replay rows are assembled from those images by the reward warps, off the
// critical path, and leave through one TMA bulk store per row and step.
//
// Warp roles (A agents): warps [0, 2A) actor warps (agent w/2; TMEM lane quadrant w%4 == the quadrant its units land in; the
// even warp of an agent also issues that agent's MMAs -- no issuer warp, no cross-agent serialisation of the issue loops),
// warps [2A, 3A) reward warps (Scenario.reward, next step's Gumbel noise, the replay-row TMA bulk store).
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
// 3xTF32 split with ROUND-TO-NEAREST on the high part: hi = rna_tf32(x) (low 13 mantissa bits zero), lo = x - hi (exact,
// signed, |lo| <= 2^-11 |x|; x == hi + lo exactly).  kind::tf32 reads the upper 19 bits of lo, i.e. truncates it toward zero:
// a sign-symmetric error <= 2^-21 |x|.  With the truncating split (hi = x & 0xFFFFE000) lo is one-sided and twice as large,
// and both the dropped lo*lo term and the truncation of lo become BIASED errors that add up along K: measured against the
// float64 oracle the free-running 25-step rollout drifted 5x more (actions 5e-7 instead of 1-2e-7).
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_rn(float x, float& hi, float& lo) {
  hi = tf32_rn(x);
  lo = x - hi;
}

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


template <int SA>
struct Cfg {
  static constexpr int A = SA, D = 6 * SA, NPAIR = (SA + 1) / 2;
  static constexpr int K1 = (D + 7) / 8 * 8;  // layer-1 K padded to whole kind::tf32 MMAs (8 per instruction)
  static constexpr int NW = 3 * SA, NTB = 32 * NW;
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

// pitch of the shared-memory replay rows: a multiple of 4 floats (TMA bulk source alignment) that is 4 mod 8, so that the
// 16 rows a warp touches per access fall on 8 different 4-bank groups (2-way instead of 4-way conflicts at pitch = 136)
__host__ __device__ inline int row_pitch(int RS) { return (RS & 7) == 0 ? RS + 4 : RS; }

__host__ __device__ inline size_t rest_floats(int A, int RS, int OS, size_t env_tile_bytes) {
  auto r4 = [](size_t x) { return (x + 3) & ~(size_t)3; };
  return r4(2 * (size_t)EB * row_pitch(RS)) + r4((size_t)A * (5 * W3A + 8)) + r4(2 * (size_t)A * EB * 8) + r4(2 * (size_t)A * EB) + r4((size_t)A * EB) +
         r4((size_t)A * (EB + 1)) + r4(2 * (size_t)OS) + r4(env_tile_bytes / 4) + r4(2 * 4 * (size_t)A) + 16;
}

#ifdef MDP_EPISODE_PROF
#ifndef MDP_PROF_AGENT
#define MDP_PROF_AGENT 0
#endif
#define TPROF_DECL long long prof_t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, prof_c = clock64();
#define TPROF(k) { const long long t_ = clock64(); prof_t[k] += t_ - prof_c; prof_c = t_; }
#else
#define TPROF_DECL
#define TPROF(k)
#endif

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
  const int OS = P.obs_stride, RS = L.row_stride, RSP = row_pitch(RS);
  const int e0 = blockIdx.x * EB;
  const int nE = min(EB, R.E - e0);
  const int nwarps = NTB >> 5;

  // ---- shared memory carve-up ---------------------------------------------------------------------------------------------
  SmemCarve sm(smem + CF::OFF_REST);
  float* sRow = sm.take(2 * EB * RSP);      // two replay-row buffers (ping-pong by step parity), assembled in place
  float* sW3 = sm.take(A * (5 * W3A + 8));  // per agent: W3 transposed [5][2 x 36] + b3
  float* sNoise = sm.take(2 * A * EB * 8);  // Gumbel noise, [step parity][agent][env][8]
  float* sPart = sm.take(2 * A * EB);       // reward partials: landmark minima (room for float64)
  int* sCnt = reinterpret_cast<int*>(sm.take(A * EB));  // ... and collision counts
  float* sRet = sm.take(A * EBP);
  ObsCol* sCols = reinterpret_cast<ObsCol*>(sm.take(2 * OS));
  real* sEnvBase = reinterpret_cast<real*>(sm.take((int)(EnvTile<real, EB>::bytes(P.scomp, A, P.act_stride, false) / 4)));
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm.take(2 * 4 * A));
  unsigned long long* bar_l1 = bars;           // [A] tcgen05.commit: layer-1 accumulator of agent i complete
  unsigned long long* bar_l2 = bars + 2 * A;   // [A] tcgen05.commit: layer-2 accumulator complete
  EnvTile<real, EB> T;
  T.carve(sEnvBase, P, sRow + L.obs_sum);
  T.ASP = RSP;

  unsigned long long counter = R.counter, episode = R.episode;
  long long cursor = R.cursor;
  if (R.ctl) {
    counter += R.ctl[0];
    cursor = (cursor + (long long)R.ctl[1]) % R.capacity;
    episode += R.ctl[2];
  }
  const int total_steps = R.steps * R.episodes;

#ifdef MDP_EPISODE_PROF
  const long long k_t0 = clock64();
  long long k_loop = 0;
#endif
  // ---- prologue ---------------------------------------------------------------------------------------------------------------
  if (warp == 0) umma::tmem_alloc(&tmem_slot, CF::T_COLS);
  if (tid == 0) {
    for (int i = 0; i < 2 * A; ++i) { mbar_init(&bar_l1[i], 1); mbar_init(&bar_l2[i], 1); }
  }
  env_load_state<real, EB>(P, T, (const real*)R.state, R.E, e0, nE);
  for (int i = tid; i < 2 * EB * RSP; i += NTB) sRow[i] = 0.f;  // padding / done / silent-comm columns stay zero for good
  for (int i = tid; i < (int)(CF::OFF_H / 16); i += NTB) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = tid; c < OS; c += NTB) sCols[c] = cols[c];
  for (int idx = tid; idx < A * EBP; idx += NTB) sRet[idx] = 0.f;
  for (int idx = tid; idx < 2 * A * EB * 8; idx += NTB) sNoise[idx] = 0.f;
  for (int idx = tid; idx < A * (5 * W3A + 8); idx += NTB) sW3[idx] = 0.f;
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  // the actors' weights: W1^T / W2^T into tensor memory (hi | lo), W3 transposed into shared memory.  Warps [4p, 4p + 4) load
  // agent pair p (lane quadrant = warp % 4); every thread first issues ALL its global loads (one unit column of W2 and W1:
  // 64 + D independent loads, coalesced across the warp), so the whole image costs one memory round trip.
  if (warp < 4 * CF::NPAIR) {
    const int p = warp >> 2, q = warp & 3;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    const int half = q >> 1, u = 32 * (q & 1) + lane;
    const int ag = 2 * p + half;
    const MlpW w = C.agents[ag < A ? ag : 0].net[MDP_NET_P];
    const bool live = ag < A;
    float x2[U], x1[K1];
#pragma unroll
    for (int k = 0; k < U; ++k) x2[k] = live ? w.W2[k * U + u] : 0.f;
#pragma unroll
    for (int k = 0; k < K1; ++k) x1[k] = (live && k < D) ? w.W1[k * U + u] : 0.f;
#pragma unroll
    for (int k0 = 0; k0 < U; k0 += 16) {
      float hi[16], lo[16];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) split_rn(x2[k0 + kk], hi[kk], lo[kk]);
      umma::tmem_st16(tbase + lane_base + CF::T_W2 + 128 * p + k0, hi);
      umma::tmem_st16(tbase + lane_base + CF::T_W2 + 128 * p + 64 + k0, lo);
    }
#pragma unroll
    for (int k0 = 0; k0 < K1; k0 += 8) {
      float hi[8], lo[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) split_rn(x1[k0 + kk], hi[kk], lo[kk]);
      umma::tmem_st8(tbase + lane_base + CF::T_W1 + 2 * K1 * p + k0, hi);
      umma::tmem_st8(tbase + lane_base + CF::T_W1 + 2 * K1 * p + K1 + k0, lo);
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
    sRow[ee * RSP + c] = v;
    const int i = c / D, cc = c - i * D;
    unsigned char* img = smem + CF::OFF_OBS + i * 2 * IMG;
    float vh, vl;
    split_rn(v, vh, vl);
    *reinterpret_cast<float*>(img + umma::sw128_off(ee, cc)) = vh;
    *reinterpret_cast<float*>(img + IMG + umma::sw128_off(ee, cc)) = vl;
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
  auto row_sync = [] { asm volatile("bar.sync 11, %0;" ::"r"(96 * A) : "memory"); };  // the step's observation images are written

  for (int ep = 0; ep < R.episodes; ++ep) {
    const int g0 = ep * R.steps;  // global step index of the episode's first step: row-buffer / noise / mbarrier parities follow it
#ifdef MDP_EPISODE_PROF
    const long long k_t1 = clock64();
#endif
    if (warp < 2 * A) {
      // =================================== actor warps =======================================================================
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
      const uint32_t rowoff = (uint32_t)((e >> 3) << 10) + (uint32_t)(e7 << 7);
      const uint32_t hR = sbase + CF::OFF_H + i * 4 * IMG + par * IMG + rowoff;
      const float* w3 = sW3 + i * (5 * W3A + 8) + W3P * par;
      float b3[5];
#pragma unroll
      for (int a = 0; a < 5; ++a) b3[a] = sW3[i * (5 * W3A + 8) + 5 * W3A + a];
      // observation writes: lane parity 0 owns the x components, parity 1 the y components of every (x, y) column pair
      constexpr int NPR = 2 + A;  // compile-time pairs: velocity, position, landmarks; then A - 1 pairs of the other agents
      uint32_t cofs[(NPR + A) / 2 + 1];
#pragma unroll
      for (int j = 0; j < (NPR + A) / 2 + 1; ++j) cofs[j] = (uint32_t)(((j ^ e7) << 4) | (par << 2));
      const uint32_t oImg = sbase + CF::OFF_OBS + i * 2 * IMG + rowoff;
      // state of (env e, agent i) in registers for the whole episode (both lanes of a row keep identical copies)
      Spread<A, real> S;
      S.init(P, i);
      real pxi = T.sS[(4 * i + 0) * EBP + e], pyi = T.sS[(4 * i + 1) * EBP + e];
      real vxi = T.sS[(4 * i + 2) * EBP + e], vyi = T.sS[(4 * i + 3) * EBP + e];
      real ml[A];  // this lane's component of the landmark positions
#pragma unroll
      for (int l = 0; l < A; ++l) ml[l] = T.sS[(4 * A + 2 * l + par) * EBP + e];

      // The agent's even warp issues its MMAs (one elected lane): D_i = W^T x act^T as lo*hi + hi*lo + hi*hi, N = 32.
      // (Splitting every GEMM into two N = 16 halves to overlap one half's epilogue with the other half's MMAs was measured
      // and bought nothing: an N = 16 MMA is not faster than an N = 32 one here, the step got 2 % slower.)
      constexpr uint32_t idesc = umma::idesc_tf32(128, EB, 0, 0);
      const uint32_t tacc = tbase + CF::T_D + 32 * i;
      auto pair_bar = [&] { asm volatile("bar.sync %0, 64;" ::"r"(1 + i) : "memory"); };
      auto issue_l1 = [&] {  // layer 1: B = the observation images
        umma::fence_after();
        const uint32_t a_hi = tbase + CF::T_W1 + 2 * K1 * (i >> 1), a_lo = a_hi + K1;
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
      };
      auto issue_l2 = [&] {  // layer 2: B = the h1 images
        umma::fence_after();
        const uint32_t a_hi = tbase + CF::T_W2 + 128 * (i >> 1), a_lo = a_hi + 64;
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
      };
      // epilogue: relu(acc + bias) of this thread's unit for the 32 env instances -> the hidden images
      uint32_t hx[8];
#pragma unroll
      for (int n7 = 0; n7 < 8; ++n7) hx[n7] = hU + xo[n7];
      auto epilogue = [&](float bias, bool with_lo) {
        float v[32];
        umma::tmem_ld32(tD, v);
#pragma unroll
        for (int n = 0; n < 32; ++n) {
          const float x = fmaxf(v[n] + bias, 0.f);
          const uint32_t a = hx[n & 7] + (uint32_t)(((n >> 3) << 10) + ((n & 7) << 7));
          if (with_lo) {  // layer-2 operand: hi | lo images
            float xh, xl;
            split_rn(x, xh, xl);
            st_s32(a, xh);
            st_s32(a + 2 * IMG, xl);
          } else {  // h2: plain fp32 for the head
            st_s32(a, x);
          }
        }
      };
      if (h == 0) issue_l1();
      TPROF_DECL
      for (int s = 0; s < R.steps; ++s) {
        const int g = g0 + s;
        float* buf = sRow + (g & 1) * EB * RSP;
        const uint32_t ph = (uint32_t)(g & 1);
        // ---- epilogue 1: h1 = relu(acc + b1) -> hi | lo images (B operand of layer 2)
        mbar_wait_b(&bar_l1[i], ph);
        TPROF(0)
        umma::fence_after();
        epilogue(b1u, true);
        umma::fence_before();
        umma::fence_async_smem();
        TPROF(1)
        pair_bar();
        if (h == 0) issue_l2();
        TPROF(8)
        // ---- epilogue 2: h2 = relu(acc + b2) -> fp32 image (read by the head below)
        mbar_wait_b(&bar_l2[i], ph);
        TPROF(2)
        umma::fence_after();
        epilogue(b2u, false);
        umma::fence_before();
        pair_bar();  // both warps of the agent: h2 rows are complete
        TPROF(3)
        // ---- output head on two threads per row (32 units each), then Gumbel-softmax (distributions.py:264-266)
        float act[5];
        {
          float2 sa[5];
#pragma unroll
          for (int a = 0; a < 5; ++a) sa[a] = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int jj = j ^ (par << 2);  // the two lanes of a row walk the 16-byte chunks in different orders: no bank conflicts
            const float4 hv = ld_s128(hR + (uint32_t)((jj ^ e7) << 4));
#pragma unroll
            for (int a = 0; a < 5; ++a) {
              const float4 wv = *reinterpret_cast<const float4*>(w3 + a * W3A + 4 * jj);
              sa[a] = __ffma2_rn(make_float2(hv.x, hv.y), make_float2(wv.x, wv.y), sa[a]);
              sa[a] = __ffma2_rn(make_float2(hv.z, hv.w), make_float2(wv.z, wv.w), sa[a]);
            }
          }
          TPROF(11)
          const float* nz = sNoise + ((g & 1) * A * EB + i * EB + e) * 8;
          const float4 n4 = *reinterpret_cast<const float4*>(nz);
          const float nzv[5] = {n4.x, n4.y, n4.z, n4.w, nz[4]};
          float m = -INFINITY;
#pragma unroll
          for (int a = 0; a < 5; ++a) {
            const float mine = sa[a].x + sa[a].y;
            const float o = __shfl_xor_sync(0xffffffffu, mine, 1);
            act[a] = ((par ? o + mine : mine + o) + b3[a]) + nzv[a];  // units [0, 32) + units [32, 64): the same value on both lanes
            m = fmaxf(m, act[a]);
          }
          float sum = 0.f;
#pragma unroll
          for (int a = 0; a < 5; ++a) { act[a] = __expf(act[a] - m); sum += act[a]; }  // arguments <= 0: ex2.approx, ~2 ulp near 0
          const float inv = 1.0f / sum;
#pragma unroll
          for (int a = 0; a < 5; ++a) act[a] = act[a] * inv;
          if (par == 0 && e < nE) {
            float* arow = buf + e * RSP + L.obs_sum + 5 * i;
#pragma unroll
            for (int a = 0; a < 5; ++a) arow[a] = act[a];
          }
        }
        TPROF(4)
        // ---- World.step for (env e, agent i): needs the other agents' OLD positions only
        {
          real px[A], py[A];
#pragma unroll
          for (int j = 0; j < A; ++j) { px[j] = T.sS[(4 * j + 0) * EBP + e]; py[j] = T.sS[(4 * j + 1) * EBP + e]; }
          S.step(i, px, py, pxi, pyi, vxi, vyi, act);
          TPROF(5)
          phys_bar();
          if (par == 0) { T.sS[(4 * i + 0) * EBP + e] = pxi; T.sS[(4 * i + 1) * EBP + e] = pyi; }
          pos_bar();
          TPROF(6)
        }
        // ---- Scenario.observation of agent i, this lane's component of every pair: [vel, pos, landmarks - pos, others - pos]
        // (the silent agents' comm columns stay zero), straight into the layer-1 operand images of the next step
        {
          const real mp = par ? pyi : pxi, mv = par ? vyi : vxi;
          auto put = [&](uint32_t off, float v) {
            float vh, vl;
            split_rn(v, vh, vl);
            st_s32(oImg + off, vh);
            st_s32(oImg + IMG + off, vl);
          };
          put(cofs[0], (float)mv);
          put(cofs[0] + 8u, (float)mp);
#pragma unroll
          for (int l = 0; l < A; ++l) put(cofs[(2 + l) >> 1] + (uint32_t)(((2 + l) & 1) << 3), (float)(ml[l] - mp));
#pragma unroll
          for (int q = 0; q < A; ++q) {
            if (q == i) continue;
            const int k = NPR + (q < i ? q : q - 1);  // runtime pair index
            const real oq = T.sS[(4 * q + par) * EBP + e];
            put((uint32_t)((((k >> 1) ^ e7) << 4) | ((k & 1) << 3) | (par << 2)), (float)(oq - mp));
          }
        }
        TPROF(7)
        umma::fence_async_smem();  // images -> UMMA (async proxy)
        TPROF(9)
        pair_bar();
        if (h == 0 && s + 1 < R.steps) issue_l1();
        row_arrive();
        TPROF(10)
      }
      // registers -> state tile (positions are current there already)
      if (par == 0) { T.sS[(4 * i + 2) * EBP + e] = vxi; T.sS[(4 * i + 3) * EBP + e] = vyi; }
#ifdef MDP_EPISODE_PROF
      if (R.ep_return && blockIdx.x == 0 && warp == 2 * MDP_PROF_AGENT && lane == 0)
        for (int k = 0; k < 12; ++k) R.ep_return[(size_t)R.E * A + k] += (float)prof_t[k];  // caller over-allocates ep_return by 16 floats
#endif
    } else {
      // =================================== reward warps ======================================================================
      const int i = warp - 2 * A, e = lane, e7 = e & 7;
      Spread<A, real> S;
      S.init(P, i);
      const real lxi = T.sS[(4 * A + 2 * i + 0) * EBP + e], lyi = T.sS[(4 * A + 2 * i + 1) * EBP + e];
      float ret_reg = 0.f;
      const bool storer = warp == 2 * A;  // lane e of this warp streams row e of every step to the ring
      long long ring_row = (cursor + (long long)g0 * R.E + e0 + e) % R.capacity;
      const uint32_t oImg = sbase + CF::OFF_OBS + i * 2 * IMG + (uint32_t)((e >> 3) << 10) + (uint32_t)(e7 << 7);
      for (int s = 0; s < R.steps; ++s) {
        const int g = g0 + s;
        float* buf = sRow + (g & 1) * EB * RSP;
        float* nxt = sRow + ((g & 1) ^ 1) * EB * RSP;
        if (g + 1 < total_steps && e < nE) {  // Gumbel noise of the next step, agent i
          float* nz = sNoise + (((g + 1) & 1) * A * EB + i * EB + e) * 8;
#pragma unroll
          for (int a = 0; a < 5; ++a)
            nz[a] = gumbel_from_u(philox_u(R.seed, counter + (unsigned long long)g + 2ull, (uint32_t)i, (long long)e0 + e, a));
        }
        // the bulk stores of step g-1 must have finished READING the other buffer before its obs columns are overwritten below
        if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        pos_bar();
        real px[A], py[A];
#pragma unroll
        for (int j = 0; j < A; ++j) { px[j] = T.sS[(4 * j + 0) * EBP + e]; py[j] = T.sS[(4 * j + 1) * EBP + e]; }
        const real mine = S.landmark_min(px, py, lxi, lyi);
        const int cn = S.collisions(px, py, T.sS[(4 * i + 0) * EBP + e], T.sS[(4 * i + 1) * EBP + e]);
        real* sp = reinterpret_cast<real*>(sPart);
        sp[i * EB + e] = mine;
        sCnt[i * EB + e] = cn;
        rew_bar();
        real m[A];
        int cnt[A];
#pragma unroll
        for (int j = 0; j < A; ++j) { m[j] = sp[j * EB + e]; cnt[j] = sCnt[j * EB + e]; }
        const float rsum = S.reward_sum(m, cnt);
        buf[e * RSP + L.rw_off + i] = rsum;
        ret_reg += rsum;
        row_sync();  // every actor warp has written its observation images of step g + 1
        // replay rows: next_obs of this step and obs_t of the next one are copies of agent i's observation image row
        {
          float* d1 = buf + e * RSP + L.nx_off + i * D;
          float* d2 = nxt + e * RSP + i * D;
#pragma unroll
          for (int j = 0; j < (D + 3) / 4; ++j) {
            const float4 vh = ld_s128(oImg + (uint32_t)((j ^ e7) << 4)), vl = ld_s128(oImg + IMG + (uint32_t)((j ^ e7) << 4));
            const float4 v = make_float4(vh.x + vl.x, vh.y + vl.y, vh.z + vl.z, vh.w + vl.w);  // hi + lo == the observation, exactly
            if (4 * j < D) {
              *reinterpret_cast<float2*>(d1 + 4 * j) = make_float2(v.x, v.y);
              *reinterpret_cast<float2*>(d2 + 4 * j) = make_float2(v.x, v.y);
            }
            if (4 * j + 2 < D) {
              *reinterpret_cast<float2*>(d1 + 4 * j + 2) = make_float2(v.z, v.w);
              *reinterpret_cast<float2*>(d2 + 4 * j + 2) = make_float2(v.z, v.w);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // replay rows -> TMA bulk store (async proxy)
        rew_bar();
        if (storer) {
          if (e < nE)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(R.ring + ring_row * RS),
                         "r"(smem_u32(buf + e * RSP)), "r"((uint32_t)(RS * 4))
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          ring_row += R.E;
          if (ring_row >= R.capacity) ring_row -= R.capacity;  // capacity >= E * steps * episodes (checked on the host)
        }
      }
      if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      sRet[i * EBP + e] += ret_reg;
    }

    // ---- end of the episode: optional reset_world; after the last one hand state and observations back ------------------------
    umma::fence_before();
    __syncthreads();
#ifdef MDP_EPISODE_PROF
    k_loop += clock64() - k_t1;
#endif
    const int gn = g0 + R.steps;
    float* fin = sRow + (gn & 1) * EB * RSP;  // obs of the next step lives in the obs_t columns of the next buffer
    const bool last = ep + 1 == R.episodes;
    if (R.reset_after) {
      for (int idx = tid; idx < P.scomp * EB; idx += NTB) {
        const int comp = idx / EB, e = idx % EB;
        T.sS[comp * EBP + e] = env_reset_value<real>(P, comp, e0 + e, R.env_seed, episode + (unsigned long long)ep, R.lm_lo, R.lm_hi);
      }
      __syncthreads();
      env_flags_rewards<real, EB, false>(P, T, nE);
      for (int c = lane; c < L.obs_sum; c += 32) {
        const ObsCol d = sCols[c];
        const int i = c / D, cc = c - i * D;
        unsigned char* img = smem + CF::OFF_OBS + i * 2 * IMG;
        for (int ee = warp; ee < EB; ee += nwarps) {
          const float v = ee < nE ? env_obs_value<real, EB>(T, d, ee) : 0.f;
          fin[ee * RSP + c] = v;
          if (!last) {
            float vh, vl;
            split_rn(v, vh, vl);
            *reinterpret_cast<float*>(img + umma::sw128_off(ee, cc)) = vh;
            *reinterpret_cast<float*>(img + IMG + umma::sw128_off(ee, cc)) = vl;
          }
        }
      }
      umma::fence_async_smem();
      __syncthreads();
    }
    if (last) {
      env_store_state<real, EB>(P, T, (real*)R.state, R.E, e0, nE, R.reset_after != 0);
      for (int ee = warp; ee < nE; ee += nwarps)
        for (int c = lane; c < OS; c += 32) R.obs[(size_t)(e0 + ee) * OS + c] = (c < L.obs_sum) ? fin[ee * RSP + c] : 0.f;
      if (R.ep_return)
        for (int idx = tid; idx < nE * A; idx += NTB) {
          const int ee = idx / A, ii = idx - ee * A;
          R.ep_return[(size_t)(e0 + ee) * A + ii] += sRet[ii * EBP + ee];
        }
    }
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) {
    umma::fence_after();
    umma::tmem_free(tbase, CF::T_COLS);
  }
#ifdef MDP_EPISODE_PROF
  if (R.ep_return && blockIdx.x == 0 && tid == 0) {
    R.ep_return[(size_t)R.E * A + 12] = (float)(clock64() - k_t0 - k_loop);
    R.ep_return[(size_t)R.E * A + 13] = (float)k_loop;
  }
#endif
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
