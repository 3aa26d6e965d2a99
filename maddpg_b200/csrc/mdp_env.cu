// Fused Multi-Agent Particle Environment step for thousands of lockstep env instances (sm_100a).
//
// Replaces, per SURVEY.md 8(a) row a4 / Appendix A (upstream openai/multiagent-particle-envs,
// imported by the reference at experiments/train.py:49-60 and driven at :104,:114,:128):
//   MultiAgentEnv.step/_set_action/reset, World.step/apply_action_force/apply_environment_force/
//   get_collision_force/integrate_state/update_agent_state, Scenario.{reset_world,reward,observation}
//
// Design (B200): one CTA owns EB consecutive env instances; the SoA state of those envs is staged
// in shared memory as [component][env] (coalesced global reads, conflict-free smem access);
// phase 1 runs one thread per (env, agent) -- warps are agent-uniform so role branches never
// diverge -- and integrates the physics in registers; phase 2 computes scenario rewards from the
// post-integration state; phase 3 streams out state, the joint observation rows (driven by a
// per-scenario column table so that consecutive lanes write consecutive floats) and rewards.
// The kernel is HBM/latency bound: algorithmic bytes per env step are SURVEY 8(d)'s bytes_env.
#include "mdp_env_dev.cuh"

#include <vector>
#include <new>

namespace mdp {

// One CTA = EB env instances.  Threads [0, EB*A) run the physics (one per (env, agent)); ALL threads
// (at least 256) take part in the load / store phases.
template <typename real, int EB, bool DO_STEP>
__global__ void __launch_bounds__(1024) k_env_step(EnvParams P, int E, int ES, real* __restrict__ state,
                                                   const float* __restrict__ act, const ObsCol* __restrict__ cols,
                                                   float* __restrict__ obs_out, float* __restrict__ rew_out,
                                                   uint8_t* __restrict__ done_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EnvTile<real, EB> T;
  T.carve(smem_raw, P, nullptr);
  const int tid = threadIdx.x, NT = blockDim.x;
  const int e0 = blockIdx.x * EB;
  const int nE = min(EB, E - e0);

  // E = env instances of this launch, ES = stride of the SoA state arrays (the whole population; a launch may cover a range)
  env_load_state<real, EB>(P, T, state, ES, e0, nE);
  if (DO_STEP) env_load_actions<real, EB>(P, T, act, e0, nE);
  __syncthreads();
  if (DO_STEP) env_physics<real, EB>(P, T, nE);
  env_flags_rewards<real, EB, DO_STEP>(P, T, nE);

  if (DO_STEP) {
    env_store_state<real, EB>(P, T, state, ES, e0, nE, false);
    for (int idx = tid; idx < nE * P.A; idx += NT) {
      const int ee = idx / P.A, ii = idx - ee * P.A;
      rew_out[(size_t)(e0 + ee) * P.A + ii] = env_reward_out<real, EB>(P, T, ee, ii);
      done_out[(size_t)(e0 + ee) * P.A + ii] = 0;  // MPE has no done callback: always False
    }
  }
  // observations: flat (row, column-quad) sweep, one float4 store per thread iteration
  env_write_obs<real, EB>(P, T, cols, obs_out + (size_t)e0 * P.obs_stride, P.obs_stride, nE);
}

// --------------------------------------------------------------------------------------------
// simple_spread fast path (float32 state, A <= 6): ONE THREAD PER ENV INSTANCE, everything in registers.
// The generic kernel above is table-driven (runtime entity loops, per-column observation decode) and issue-bound
// at ~170 warp instructions per env instance; with A known at compile time the whole step unrolls to ~15, which
// leaves the kernel waiting on HBM only.  I/O: the SoA state is read/written coalesced straight from registers;
// the AoS action row is the thread's own 4*AS4 floats (float4 loads); the joint observation rows of the CTA's
// 128 env instances are assembled in shared memory and streamed out as one contiguous block of float4 stores.
// Arithmetic and operation order follow env_physics / env_flags_rewards / env_obs_value exactly.
// --------------------------------------------------------------------------------------------
constexpr int SPREAD_EB = 128;

template <int A>
__global__ void __launch_bounds__(SPREAD_EB) k_env_step_spread(EnvParams P, int E, int ES, float* __restrict__ state,
                                                               const float* __restrict__ act, float* __restrict__ obs_out,
                                                               float* __restrict__ rew_out, uint8_t* __restrict__ done_out) {
  constexpr int L = A, D = 6 * A, OS = (A * D + 3) / 4 * 4, AS = (5 * A + 3) / 4 * 4, OS4 = OS / 4;
  constexpr int PITCH4 = OS4 | 1;  // odd pitch in float4 units
  extern __shared__ __align__(16) float sObs[];  // [SPREAD_EB][PITCH4 float4]
  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * SPREAD_EB;
  const int nE = min(SPREAD_EB, E - e0);
  const int e = e0 + tid;
  if (tid < nE) {
    SpreadRegs<A> S;
    float a[AS];
#pragma unroll
    for (int i = 0; i < A; ++i) {
      S.px[i] = state[(size_t)(4 * i + 0) * ES + e];
      S.py[i] = state[(size_t)(4 * i + 1) * ES + e];
      S.vx[i] = state[(size_t)(4 * i + 2) * ES + e];
      S.vy[i] = state[(size_t)(4 * i + 3) * ES + e];
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
      S.lx[l] = state[(size_t)(4 * A + 2 * l + 0) * ES + e];
      S.ly[l] = state[(size_t)(4 * A + 2 * l + 1) * ES + e];
    }
    const float4* arow = reinterpret_cast<const float4*>(act + (size_t)e * AS);
#pragma unroll
    for (int q = 0; q < AS / 4; ++q) {
      const float4 v = arow[q];
      a[4 * q + 0] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
    }
    const float msum = spread_step<A>(P, S, a);
#pragma unroll
    for (int i = 0; i < A; ++i) {
      state[(size_t)(4 * i + 0) * ES + e] = S.px[i];
      state[(size_t)(4 * i + 1) * ES + e] = S.py[i];
      state[(size_t)(4 * i + 2) * ES + e] = S.vx[i];
      state[(size_t)(4 * i + 3) * ES + e] = S.vy[i];
    }
#pragma unroll
    for (int i = 0; i < A; ++i) rew_out[(size_t)e * A + i] = msum;
    float o[OS];
    spread_obs<A>(S, o);
#pragma unroll
    for (int c = A * D; c < OS; ++c) o[c] = 0.f;
    // row pitch OS4 + 1 float4 (odd for every A here, or made odd): conflict-free 16-byte stores and reads
    float4* orow = reinterpret_cast<float4*>(sObs) + tid * PITCH4;
#pragma unroll
    for (int q = 0; q < OS4; ++q) orow[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  }
  __syncthreads();
  // the CTA's observation rows are contiguous in global memory: one flat float4 sweep
  float4* dst = reinterpret_cast<float4*>(obs_out + (size_t)e0 * OS);
  const float4* src = reinterpret_cast<const float4*>(sObs);
  for (int i = tid; i < nE * OS4; i += SPREAD_EB) {
    const int ee = i / OS4, q = i - ee * OS4;
    dst[i] = src[ee * PITCH4 + q];
  }
  // done is identically False in MPE (no done callback)
  uint8_t* dn = done_out + (size_t)e0 * A;
  for (int i = tid; i < nE * A; i += SPREAD_EB) dn[i] = 0;
}

// --------------------------------------------------------------------------------------------
// simple_spread with 7..32 agents (float32 state): ONE WARP PER ENV INSTANCE, lane = agent = landmark.  BASELINE.json
// configs[4] (N = 24: 48 entities, 276 collidable pairs, 15 384 B per env step) is the one configuration whose env step is
// genuinely HBM-sized (504 MB per step at 32 768 env instances).  A lane keeps its agent's state and its landmark in
// registers; every pairwise term (soft-contact forces, landmark minima, collision counts) runs as a shuffle loop over the
// entities -- in entity order, with the pinned roundings of spread_agent_step, so the results equal the register kernel's
// -- and the observation rows are produced column-major across the lanes (lane = column: the partner a column needs is one
// shuffle away), so every store is a full 128-byte line.  RING: the same warp also writes the joint replay row of the
// transition (obs_t copied from obs_prev, act_t, next_obs, rew, done) -- ReplayBuffer.add for all agents fused into the
// step: one pass over the observations instead of the separate insert kernel's re-read (1.85 GB instead of 2.9 GB of traffic
// per step at the configs[4] size).
// --------------------------------------------------------------------------------------------
template <int AT, bool RING>
__global__ void __launch_bounds__(256) k_env_step_spread_warp(EnvParams P, int E, int ES, float* __restrict__ state,
                                                              const float* __restrict__ act, float* __restrict__ obs_out,
                                                              float* __restrict__ rew_out, uint8_t* __restrict__ done_out,
                                                              const float* __restrict__ obs_prev, float* __restrict__ ring,
                                                              long long capacity, long long cursor, mdp_ring_layout L,
                                                              const unsigned long long* __restrict__ ctl) {
  // positions of the warp's env instance: [0, A) agents, [32, 32 + A) landmarks, as (x, y) pairs; velocities at [64, 64 + A)
  __shared__ float2 sPos[8][96];
  const int e = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (e >= E) return;  // warp-uniform
  float2* sp = sPos[threadIdx.x >> 5];
  const float* spf = reinterpret_cast<const float*>(sp);
  const unsigned FULL = 0xffffffffu;
  // AT > 0: the agent count is a template constant -- every loop over the entities unrolls, the per-agent values a lane needs
  // for its observation columns live in registers and the row offsets become store immediates (BASELINE configs[4]: 24)
  const int A = AT ? AT : P.A, D = 6 * A, OS = P.obs_stride, AS = P.act_stride;
  const bool live = lane < A;
  const int li = live ? lane : 0;
  float px = state[(size_t)(4 * li + 0) * ES + e], py = state[(size_t)(4 * li + 1) * ES + e];
  float vx = state[(size_t)(4 * li + 2) * ES + e], vy = state[(size_t)(4 * li + 3) * ES + e];
  const float lx = state[(size_t)(4 * A + 2 * li + 0) * ES + e], ly = state[(size_t)(4 * A + 2 * li + 1) * ES + e];
  const float* arow = act + (size_t)e * AS + 5 * li;
  const float a1 = arow[1], a2 = arow[2], a3 = arow[3], a4 = arow[4];
  const float k = (float)P.contact_margin, cf = (float)P.contact_force, damp = 1.0f - (float)P.damping, dt = (float)P.dt;
  const float si = P.sizef[li], sens = (float)P.sens[li], ms = (float)P.max_speed[li];
  sp[lane] = make_float2(px, py);
  sp[32 + lane] = make_float2(lx, ly);
  __syncwarp();
  // World.step: action force, then the soft contact with every other agent in agent order (landmarks do not collide).
  // A pair further apart than dmin + 0.11 has z = -(dist - dmin) / k < -104: exp(z) underflows and the penetration is exactly 0
  // -- decided on the squared distance, so that far pairs (nearly all of them) cost neither the sqrt nor the division.
  float fx = __fmul_rn(__fsub_rn(a1, a2), sens), fy = __fmul_rn(__fsub_rn(a3, a4), sens);
  unsigned near = 0;  // pass 1 (unrolled, branch-free): which agents are close enough for a non-zero force
#pragma unroll
  for (int j = 0; j < A; ++j) {
    const float2 pj = sp[j];
    const float d2 = sq_norm2(__fsub_rn(px, pj.x), __fsub_rn(py, pj.y));
    const float far = __fadd_rn(si, P.sizef[j]) + 0.11f;
    near |= (d2 <= far * far ? 1u : 0u) << j;
  }
  near &= ~(1u << lane);
  while (near) {  // pass 2: the few contact pairs, in agent order (one copy of the IEEE sqrt / div / exp / log1p code)
    const int j = __ffs(near) - 1;
    near &= near - 1;
    const float2 pj = sp[j];
    const float dx = __fsub_rn(px, pj.x), dy = __fsub_rn(py, pj.y);
    const float dmin = __fadd_rn(si, P.sizef[j]);
    const float dist = __fsqrt_rn(sq_norm2(dx, dy));
    const float z = __fdiv_rn(-__fsub_rn(dist, dmin), k);
    if (z < -104.0f) continue;
    // logaddexp(0, z) = log1p(exp(z)); below z = -17, exp(z) < 2^-24 and log1p(x) rounds to x itself in float32
    const float pen = __fmul_rn(z < -17.0f ? expf(z) : logaddexp0<float>(z), k);
    fx = __fmaf_rn(__fdiv_rn(__fmul_rn(cf, dx), dist), pen, fx);
    fy = __fmaf_rn(__fdiv_rn(__fmul_rn(cf, dy), dist), pen, fy);
  }
  float wx = __fmaf_rn(fx, dt, __fmul_rn(vx, damp)), wy = __fmaf_rn(fy, dt, __fmul_rn(vy, damp));
  if (ms > 0.f) {
    const float speed = __fsqrt_rn(sq_norm2(wx, wy));
    if (speed > ms) { wx = __fmul_rn(__fdiv_rn(wx, speed), ms); wy = __fmul_rn(__fdiv_rn(wy, speed), ms); }
  }
  vx = wx; vy = wy;
  px = __fmaf_rn(wx, dt, px);
  py = __fmaf_rn(wy, dt, py);
  __syncwarp();  // every lane has read the old positions
  sp[lane] = make_float2(px, py);
  sp[64 + lane] = make_float2(vx, vy);
  if (live) {
    state[(size_t)(4 * lane + 0) * ES + e] = px; state[(size_t)(4 * lane + 1) * ES + e] = py;
    state[(size_t)(4 * lane + 2) * ES + e] = vx; state[(size_t)(4 * lane + 3) * ES + e] = vy;
  }
  __syncwarp();
  // Scenario.reward from the new positions: lane l = landmark l's closest agent (sqrt is monotone: the minimum of the distances
  // is the sqrt of the minimum squared distance, bit for bit), lane i = agent i's collision count (sqrt only near the threshold)
  float best2 = 0.f;
  int cnt = 0;
  unsigned amb = 0;  // pairs within 0.1 % of the collision threshold: decided on the rounded distance below (one sqrt copy)
#pragma unroll
  for (int q = 0; q < A; ++q) {
    const float2 pq = sp[q];
    const float d2l = sq_norm2(__fsub_rn(pq.x, lx), __fsub_rn(pq.y, ly));
    best2 = (q == 0 || d2l < best2) ? d2l : best2;
    const float d2a = sq_norm2(__fsub_rn(pq.x, px), __fsub_rn(pq.y, py));
    const float s = __fadd_rn(P.sizef[q], si), s2 = s * s;
    cnt += d2a < 0.999f * s2 ? 1 : 0;
    amb |= ((d2a >= 0.999f * s2 && d2a <= 1.001f * s2) ? 1u : 0u) << q;
  }
  while (amb) {
    const int q = __ffs(amb) - 1;
    amb &= amb - 1;
    const float2 pq = sp[q];
    const float d2a = sq_norm2(__fsub_rn(pq.x, px), __fsub_rn(pq.y, py));
    cnt += __fsqrt_rn(d2a) < __fadd_rn(P.sizef[q], si) ? 1 : 0;
  }
  const float best = __fsqrt_rn(best2);
  float ri = 0.f;
#pragma unroll
  for (int l = 0; l < A; ++l) ri -= __shfl_sync(FULL, best, l);
  ri -= (float)cnt;
  float msum = 0.f;  // shared reward: every agent receives the sum over agents, accumulated in agent order
#pragma unroll
  for (int i = 0; i < A; ++i) msum += __shfl_sync(FULL, ri, i);
  if (live) {
    rew_out[(size_t)e * A + lane] = msum;
    done_out[(size_t)e * A + lane] = 0;
  }
  // replay row of this transition (RING): obs_t | act_t from the caller's arrays, rew / done here, next_obs below
  float* row = nullptr;
  if (RING) {
    if (ctl) cursor = (cursor + (long long)ctl[1]) % capacity;
    row = ring + ((cursor + e) % capacity) * (long long)L.row_stride;
    const float* op = obs_prev + (size_t)e * OS;
    if ((L.obs_sum & 3) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(op);
      float4* d4 = reinterpret_cast<float4*>(row);
      const int n4 = L.obs_sum >> 2;
      int c = lane;
      for (; c + 96 < n4; c += 128) {  // four independent 16-byte loads in flight per lane
        const float4 v0 = __ldcs(s4 + c), v1 = __ldcs(s4 + c + 32), v2 = __ldcs(s4 + c + 64), v3 = __ldcs(s4 + c + 96);
        __stcs(d4 + c, v0); __stcs(d4 + c + 32, v1); __stcs(d4 + c + 64, v2); __stcs(d4 + c + 96, v3);
      }
      for (; c < n4; c += 32) __stcs(d4 + c, __ldcs(s4 + c));
    } else {
      for (int c = lane; c < L.obs_sum; c += 32) row[c] = op[c];
    }
    const float* ap = act + (size_t)e * AS;
    for (int c = lane; c < L.act_sum; c += 32) row[L.obs_sum + c] = ap[c];
    if (live) { row[L.rw_off + lane] = msum; row[L.dn_off + lane] = 0.f; }
  }
  // Scenario.observation; lane = column (mod 32), so a lane's component (x / y) is its parity in every 32-column chunk:
  //   [vel(2) pos(2) | landmarks - pos (2A) | other agents - pos (2(A-1)) | silent comm zeros (2(A-1))]
  // A column of row i is  base - m * p_i  with (base, m) fixed per chunk and lane: landmark columns base = the landmark's
  // component; "other agent" columns base = agent t's component for rows i > t and agent t+1's for rows i <= t; comm columns
  // base = m = 0.  So a row costs one select, one FFMA (exactly base - p_i: one rounding) and the store.
  float* orow = obs_out + (size_t)e * OS;
  float* nrow = RING ? row + L.nx_off : nullptr;
  const int c_lm = 4, c_ot = 4 + 2 * A, c_cm = c_ot + 2 * (A - 1);
  const int comp = lane & 1;
  float pc[AT ? AT : 1];  // this lane's component of every agent's position
  if (AT) {
#pragma unroll
    for (int i = 0; i < (AT ? AT : 1); ++i) pc[i] = spf[2 * i + comp];
  }
#pragma unroll
  for (int c0 = 0; c0 < D; c0 += 32) {
    const int c = c0 + lane;
    const int kind = c < c_lm ? 0 : c < c_ot ? 1 : c < c_cm ? 2 : 3;
    const int t = kind == 1 ? (c - c_lm) >> 1 : kind == 2 ? (c - c_ot) >> 1 : 0;
    const float b_hi = kind == 1 ? spf[64 + 2 * t + comp] : kind == 2 ? spf[2 * (t + 1) + comp] : 0.f;  // rows i <= t
    const float b_lo = kind == 2 ? spf[2 * t + comp] : b_hi;                                            // rows i > t
    const float m = (kind == 1 || kind == 2) ? -1.0f : 0.0f;
    const int thr = kind == 2 ? t : 1 << 30;
    const bool on = c < D;
    const int k0 = (c < 2 ? 128 : 0) + comp;  // kind 0: velocity (columns 0, 1) or position (2, 3) of the row's agent
    float* o = orow + c;
    float* n = RING ? nrow + c : nullptr;
#pragma unroll
    for (int i = 0; i < A; ++i) {
      const float pci = AT ? pc[AT ? i : 0] : spf[2 * i + comp];
      float v = __fmaf_rn(pci, m, i > thr ? b_lo : b_hi);
      if (c0 == 0 && kind == 0) v = spf[k0 + 2 * i];
      if (on) {
        o[i * D] = v;
        if (RING) n[i * D] = v;
      }
    }
  }
  for (int c = A * D + lane; c < OS; c += 32) orow[c] = 0.f;
}

// Scenario.benchmark_data(agent, world) (the info_n tape of `train.py --benchmark`, train.py:139-148) for every (env, agent):
// four floats.  simple_spread: (reward, collisions, sum over landmarks of the closest agent's distance, occupied landmarks);
// simple_tag / simple_world_comm: (collisions with good agents, 0, 0, 0) for adversaries, zeros for good agents; simple: zeros.
template <typename real>
__global__ void k_env_benchmark(EnvParams P, int E, const real* __restrict__ state, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E * P.A) return;
  const int e = idx / P.A, i = idx - e * P.A;
  auto px = [&](int ent) { return state[(size_t)(ent_comp(P, ent) + 0) * E + e]; };
  auto py = [&](int ent) { return state[(size_t)(ent_comp(P, ent) + 1) * E + e]; };
  auto collides = [&](int a, int b) {
    const real dx = px(a) - px(b), dy = py(a) - py(b);
    return r_sqrt<real>(dx * dx + dy * dy) < ent_size<real>(P, a) + ent_size<real>(P, b);
  };
  float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
  if (P.scenario == MDP_SIMPLE_SPREAD) {
    real rew = 0, min_dists = 0;
    int occupied = 0, collisions = 0;
    for (int l = 0; l < P.L; ++l) {
      real best = 0;
      for (int a = 0; a < P.A; ++a) {
        const real dx = px(a) - px(P.A + l), dy = py(a) - py(P.A + l);
        const real d = r_sqrt<real>(dx * dx + dy * dy);
        best = (a == 0 || d < best) ? d : best;
      }
      min_dists += best;
      rew -= best;
      if (best < (real)0.1) ++occupied;
    }
    if ((P.collide_mask >> i) & 1ull)
      for (int a = 0; a < P.A; ++a)
        if (collides(a, i)) { rew -= 1; ++collisions; }
    o0 = (float)rew; o1 = (float)collisions; o2 = (float)min_dists; o3 = (float)occupied;
  } else if (P.scenario == MDP_SIMPLE_TAG || P.scenario == MDP_SIMPLE_WORLD_COMM) {
    if (i < P.n_adv) {
      int collisions = 0;
      for (int a = P.n_adv; a < P.A; ++a)
        if (collides(a, i)) ++collisions;
      o0 = (float)collisions;
    }
  }
  reinterpret_cast<float4*>(out)[idx] = make_float4(o0, o1, o2, o3);
}

// scenario.reset_world: agents U(-1,1), velocities / comm 0, landmarks U(lo,hi)
template <typename real>
__global__ void k_env_reset(EnvParams P, int E, real* __restrict__ state, uint64_t seed, uint64_t episode,
                            float lm_lo, float lm_hi, const unsigned long long* __restrict__ ctl) {
  if (ctl) episode += ctl[2];
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)P.scomp * E;
  if (idx >= total) return;
  int comp = (int)(idx / E);
  int e = (int)(idx - (size_t)comp * E);
  state[idx] = env_reset_value<real>(P, comp, e, seed, episode, lm_lo, lm_hi);
}

// --------------------------------------------------------------------------------------------
// host: scenario tables (make_world) -- SURVEY Appendix A.3
// --------------------------------------------------------------------------------------------
static void push_col(std::vector<ObsCol>& v, uint8_t kind, int i, int o, int k, int a, int b) {
  ObsCol c;
  c.kind = kind; c.i = (uint8_t)i; c.o = (uint8_t)o; c.k = (uint8_t)k; c.a = (uint8_t)a; c.b = (uint8_t)b;
  c.pad0 = c.pad1 = 0;
  v.push_back(c);
}

static int build_env(mdp_env* env) {
  EnvParams& P = env->P;
  mdp_env_dims& D = env->dims;
  memset(&P, 0, sizeof(P));
  memset(&D, 0, sizeof(D));
  const int sc = env->cfg.scenario;
  P.scenario = sc;
  P.dt = 0.1; P.damping = 0.25; P.contact_force = 1e+2; P.contact_margin = 1e-3;
  int A = 0, L = 0;
  uint32_t movable = 0xffffffffu;
  env->reset_lo_lm = -1.f; env->reset_hi_lm = 1.f;
  if (sc == MDP_SIMPLE) {
    A = 1; L = 1; P.cdim = 0; P.collide_mask = 0;
    P.size[0] = 0.05; P.size[1] = 0.05; P.sens[0] = 5.0; P.max_speed[0] = -1;
  } else if (sc == MDP_SIMPLE_SPREAD) {
    A = env->cfg.num_agents > 0 ? env->cfg.num_agents : 3;
    if (A > MDP_MAX_AGENTS) return fail(MDP_EINVAL, "simple_spread: num_agents %d > %d", A, MDP_MAX_AGENTS);
    L = A; P.cdim = 0;  // every agent is silent: comm state is identically zero, not stored
    P.collaborative = 1;
    for (int i = 0; i < A; ++i) { P.size[i] = 0.15; P.sens[i] = 5.0; P.max_speed[i] = -1; P.collide_mask |= 1ull << i; }
    for (int l = 0; l < L; ++l) P.size[A + l] = 0.05;
  } else if (sc == MDP_SIMPLE_TAG) {
    A = 4; L = 2; P.cdim = 0; P.n_adv = 3;
    for (int i = 0; i < A; ++i) {
      bool adv = i < P.n_adv;
      P.size[i] = adv ? 0.075 : 0.05; P.sens[i] = adv ? 3.0 : 4.0; P.max_speed[i] = adv ? 1.0 : 1.3;
      P.collide_mask |= 1ull << i;
    }
    for (int l = 0; l < L; ++l) { P.size[A + l] = 0.2; P.collide_mask |= 1ull << (A + l); }
    env->reset_lo_lm = -0.9f; env->reset_hi_lm = 0.9f;
  } else if (sc == MDP_SIMPLE_WORLD_COMM) {
    A = 6; L = 5; P.cdim = 4; P.n_adv = 4; P.food0 = 1; P.n_food = 2; P.forest0 = 3; P.n_forest = 2;
    for (int i = 0; i < A; ++i) {
      bool adv = i < P.n_adv;
      P.size[i] = adv ? 0.075 : 0.045; P.sens[i] = adv ? 3.0 : 4.0; P.max_speed[i] = adv ? 1.0 : 1.3;
      P.collide_mask |= 1ull << i;
    }
    // agent 0 (leader) speaks: c_dim[0] below
    P.size[A + 0] = 0.2; P.collide_mask |= 1ull << (A + 0);
    P.size[A + 1] = P.size[A + 2] = 0.03;
    P.size[A + 3] = P.size[A + 4] = 0.3;
    env->reset_lo_lm = -0.9f; env->reset_hi_lm = 0.9f;
    P.c_dim[0] = 4;
  } else if (sc == MDP_SIMPLE_ADVERSARY) {
    // physical deception: agent 0 is the adversary, N = 2 good agents, N landmarks, one of them the goal; nothing collides
    A = 3; L = 2; P.n_adv = 1; P.n_goal = 1; P.collide_mask = 0;
    for (int i = 0; i < A; ++i) { P.size[i] = 0.15; P.sens[i] = 5.0; P.max_speed[i] = -1; }
    for (int l = 0; l < L; ++l) P.size[A + l] = 0.08;
  } else if (sc == MDP_SIMPLE_PUSH) {
    // keep-away: agent 0 adversary, agent 1 good; the two agents collide, the landmarks do not
    A = 2; L = 2; P.n_adv = 1; P.n_goal = 1;
    for (int i = 0; i < A; ++i) { P.size[i] = 0.05; P.sens[i] = 5.0; P.max_speed[i] = -1; P.collide_mask |= 1ull << i; }
    for (int l = 0; l < L; ++l) P.size[A + l] = 0.05;
  } else if (sc == MDP_SIMPLE_SPEAKER_LISTENER) {
    // cooperative communication: agent 0 = speaker (not movable, dim_c = 3), agent 1 = listener (silent); shared reward
    A = 2; L = 3; P.n_goal = 1; P.collide_mask = 0; P.collaborative = 1;
    for (int i = 0; i < A; ++i) { P.size[i] = 0.075; P.sens[i] = 5.0; P.max_speed[i] = -1; }
    for (int l = 0; l < L; ++l) P.size[A + l] = 0.04;
    movable = 0x2u; P.c_dim[0] = 3;
  } else if (sc == MDP_SIMPLE_REFERENCE) {
    // two agents, each told (through the other's 10-channel message) which of 3 landmarks the OTHER should reach; both move
    // and speak: MultiDiscrete([[0, 4], [0, 9]]) action blocks; goal slot i = agent i's goal_b; shared reward
    A = 2; L = 3; P.n_goal = 2; P.collide_mask = 0; P.collaborative = 1;
    for (int i = 0; i < A; ++i) { P.size[i] = 0.05; P.sens[i] = 5.0; P.max_speed[i] = -1; P.c_dim[i] = 10; }
    for (int l = 0; l < L; ++l) P.size[A + l] = 0.05;
  } else if (sc == MDP_SIMPLE_CRYPTO) {
    // covert communication: agent 0 = adversary (Eve), agent 1 = listener (Bob), agent 2 = speaker (Alice); nobody moves, everybody
    // speaks (dim_c = 4); goal slot 0 = the goal landmark, slot 1 = the key
    A = 3; L = 2; P.n_adv = 1; P.n_goal = 2; P.collide_mask = 0;
    for (int i = 0; i < A; ++i) { P.size[i] = 0.05; P.sens[i] = 5.0; P.max_speed[i] = -1; P.c_dim[i] = 4; }
    for (int l = 0; l < L; ++l) P.size[A + l] = 0.05;
    movable = 0;
  } else {
    return fail(MDP_ENOTSUP, "unknown scenario id %d", sc);
  }
  P.A = A; P.L = L; P.NE = A + L;
  P.movable_mask = movable;
  P.cdim = 0;
  P.silent_mask = 0xffffffffu;
  for (int i = 0; i < A; ++i) {
    P.c_off[i] = P.cdim;
    P.cdim += P.c_dim[i];
    if (P.c_dim[i]) P.silent_mask &= ~(1u << i);
  }
  for (int j = 0; j < MAX_ENT; ++j) P.sizef[j] = (float)P.size[j];
  P.gcomp0 = 4 * A + P.cdim + 2 * L;
  P.scomp = P.gcomp0 + P.n_goal;
  auto pc = [&](int ent, int c) { return (ent < A ? 4 * ent : 4 * A + P.cdim + 2 * (ent - A)) + c; };
  auto vc = [&](int ag, int c) { return 4 * ag + 2 + c; };

  // observation column table, in the reference's concatenation order
  std::vector<ObsCol>& cols = env->cols;
  cols.clear();
  int aoff = 0;
  for (int i = 0; i < A; ++i) {
    D.obs_off[i] = (int)cols.size();
    D.act_off[i] = aoff; P.act_off[i] = aoff;
    D.n_heads[i] = 1; D.head_dim[i][0] = 5; D.head_dim[i][1] = 0;
    if (sc == MDP_SIMPLE_WORLD_COMM && i == 0) { D.n_heads[i] = 2; D.head_dim[i][1] = 4; }
    D.act_dim[i] = D.head_dim[i][0] + D.head_dim[i][1];
    aoff += D.act_dim[i];
    if (sc >= MDP_SIMPLE_ADVERSARY) {
      // head layout of MultiAgentEnv.__init__: Discrete(5) if movable, Discrete(dim_c) if not silent, MultiDiscrete of both
      const bool mv = (movable >> i) & 1u;
      D.head_dim[i][0] = mv ? 5 : P.c_dim[i];
      if (mv && P.c_dim[i]) { D.n_heads[i] = 2; D.head_dim[i][1] = P.c_dim[i]; }  // MultiDiscrete([movement, message])
      D.act_dim[i] = D.head_dim[i][0] + D.head_dim[i][1];
      aoff += D.act_dim[i] - 5;
      const int lm0 = pc(A, 0), g0 = P.gcomp0;
      auto vel = [&]() { for (int c = 0; c < 2; ++c) push_col(cols, OK_DIRECT, i, 0, 0, vc(i, c), 0); };
      auto entity_pos = [&]() {
        for (int l = 0; l < L; ++l)
          for (int c = 0; c < 2; ++c) push_col(cols, OK_REL, i, 0, 0, pc(A + l, c), pc(i, c));
      };
      auto other_pos = [&]() {
        for (int o = 0; o < A; ++o) if (o != i)
          for (int c = 0; c < 2; ++c) push_col(cols, OK_REL, i, o, 0, pc(o, c), pc(i, c));
      };
      auto goal_rel = [&]() { for (int c = 0; c < 2; ++c) push_col(cols, OK_REL_GOAL, 0, lm0, g0, c, pc(i, c)); };
      // colour channel that depends on a goal slot: LUT index per goal value 0..3, one nibble each
      auto goal_lut = [&](int slot, int v0, int v1, int v2, int v3) {
        const unsigned pack = (unsigned)v0 | ((unsigned)v1 << 4) | ((unsigned)v2 << 8) | ((unsigned)v3 << 12);
        push_col(cols, OK_GOAL_LUT, (int)(pack & 0xff), (int)(pack >> 8), g0 + slot, 0, 0);
      };
      auto comm_of = [&](int o) { for (int c = 0; c < P.c_dim[o]; ++c) push_col(cols, OK_DIRECT, i, o, 0, 4 * A + P.c_off[o] + c, 0); };
      if (sc == MDP_SIMPLE_ADVERSARY) {
        if (i >= P.n_adv) goal_rel();
        entity_pos(); other_pos();
      } else if (sc == MDP_SIMPLE_PUSH) {
        vel();
        if (i >= P.n_adv) {
          goal_rel();
          push_col(cols, OK_CONST, 0, 0, 0, LUT_025, 0);      // agent.color = [0.25, 0.25, 0.25]; [goal.index + 1] += 0.5
          goal_lut(0, LUT_075, LUT_025, LUT_025, LUT_025);
          goal_lut(0, LUT_025, LUT_075, LUT_025, LUT_025);
          entity_pos();
          for (int l = 0; l < L; ++l)                         // landmark.color = [0.1, 0.1, 0.1]; [l + 1] += 0.8
            for (int ch = 0; ch < 3; ++ch) push_col(cols, OK_CONST, 0, 0, 0, ch == l + 1 ? LUT_090 : LUT_010, 0);
          other_pos();
        } else {
          entity_pos(); other_pos();
        }
      } else if (sc == MDP_SIMPLE_SPEAKER_LISTENER) {
        if (i == 0) {  // speaker: the goal landmark's colour (0.65 on its own channel, 0.15 elsewhere)
          for (int ch = 0; ch < 3; ++ch)
            goal_lut(0, ch == 0 ? LUT_065 : LUT_015, ch == 1 ? LUT_065 : LUT_015, ch == 2 ? LUT_065 : LUT_015, LUT_015);
        } else {       // listener
          vel(); entity_pos(); comm_of(0);
        }
      } else if (sc == MDP_SIMPLE_REFERENCE) {
        vel(); entity_pos();
        for (int ch = 0; ch < 3; ++ch)  // goal_b's colour: 0.75 on its own channel, 0.25 elsewhere
          goal_lut(i, ch == 0 ? LUT_075 : LUT_025, ch == 1 ? LUT_075 : LUT_025, ch == 2 ? LUT_075 : LUT_025, LUT_025);
        comm_of(1 - i);
      } else {  // MDP_SIMPLE_CRYPTO: one-hot "colours" in dim_c = 4 channels
        auto onehot = [&](int slot) {
          for (int ch = 0; ch < 4; ++ch)
            goal_lut(slot, ch == 0 ? LUT_1 : LUT_0, ch == 1 ? LUT_1 : LUT_0, ch == 2 ? LUT_1 : LUT_0, ch == 3 ? LUT_1 : LUT_0);
        };
        if (i == A - 1) { onehot(0); onehot(1); }     // speaker: [goal colour, key]
        else if (i >= P.n_adv) { onehot(1); comm_of(A - 1); }  // listener: [key, speaker's message]
        else comm_of(A - 1);                            // adversary: the message only
      }
      D.obs_dim[i] = (int)cols.size() - D.obs_off[i];
      continue;
    }
    for (int c = 0; c < 2; ++c) push_col(cols, OK_DIRECT, i, 0, 0, vc(i, c), 0);                  // p_vel
    if (sc != MDP_SIMPLE)
      for (int c = 0; c < 2; ++c) push_col(cols, OK_DIRECT, i, 0, 0, pc(i, c), 0);                // p_pos
    for (int l = 0; l < L; ++l)
      for (int c = 0; c < 2; ++c) push_col(cols, OK_REL, i, 0, 0, pc(A + l, c), pc(i, c));        // entity_pos
    if (sc == MDP_SIMPLE_SPREAD) {
      for (int o = 0; o < A; ++o) if (o != i)
        for (int c = 0; c < 2; ++c) push_col(cols, OK_REL, i, o, 0, pc(o, c), pc(i, c));          // other_pos
      for (int o = 0; o < A; ++o) if (o != i)
        for (int c = 0; c < 2; ++c) push_col(cols, OK_ZERO, i, o, 0, 0, 0);                       // comm (silent)
    } else if (sc == MDP_SIMPLE_TAG) {
      for (int o = 0; o < A; ++o) if (o != i)
        for (int c = 0; c < 2; ++c) push_col(cols, OK_REL, i, o, 0, pc(o, c), pc(i, c));
      for (int o = 0; o < A; ++o) if (o != i && o >= P.n_adv)
        for (int c = 0; c < 2; ++c) push_col(cols, OK_DIRECT, i, o, 0, vc(o, c), 0);              // other_vel (good only)
    } else if (sc == MDP_SIMPLE_WORLD_COMM) {
      for (int o = 0; o < A; ++o) if (o != i)
        for (int c = 0; c < 2; ++c) push_col(cols, OK_REL_MASK, i, o, 0, pc(o, c), pc(i, c));
      auto other_vel = [&]() {
        for (int o = 0; o < A; ++o) if (o != i && o >= P.n_adv)
          for (int c = 0; c < 2; ++c) push_col(cols, OK_DIR_MASK, i, o, 0, vc(o, c), 0);
      };
      auto in_forest = [&]() { for (int q = 0; q < 2; ++q) push_col(cols, OK_FOREST, i, 0, q, 0, 0); };
      if (i < P.n_adv) {
        other_vel(); in_forest();
        for (int k = 0; k < 4; ++k) push_col(cols, OK_DIRECT, i, 0, 0, 4 * A + k, 0);             // leader comm
      } else {
        in_forest(); other_vel();
      }
    }
    D.obs_dim[i] = (int)cols.size() - D.obs_off[i];
  }
  D.obs_sum = (int)cols.size();
  D.act_sum = aoff;
  D.obs_stride = round_up(D.obs_sum, 4);
  D.act_stride = round_up(D.act_sum, 4);
  while ((int)cols.size() < D.obs_stride) push_col(cols, OK_PAD, 0, 0, 0, 0, 0);
  for (auto& c : cols) {  // plain columns are evaluated as S[a] - S[b]; row `scomp` of the smem tile is all zeros
    if (c.kind == OK_DIRECT || c.kind == OK_DIR_MASK) c.b = (uint8_t)P.scomp;
    if (c.kind == OK_ZERO || c.kind == OK_PAD) c.a = c.b = (uint8_t)P.scomp;
    if (c.kind > OK_ZERO) P.has_mask = 1;
  }
  if (P.scomp + 1 > 255) return fail(MDP_ENOTSUP, "scenario has %d state components (max 254)", P.scomp);
  P.obs_stride = D.obs_stride; P.act_stride = D.act_stride; P.obs_sum = D.obs_sum; P.act_sum = D.act_sum;
  P.os4_magic = (uint32_t)((0x100000000ull + (uint64_t)(D.obs_stride / 4) - 1) / (uint64_t)(D.obs_stride / 4));
  P.as4_magic = (uint32_t)((0x100000000ull + (uint64_t)(D.act_stride / 4) - 1) / (uint64_t)(D.act_stride / 4));
  D.n_agents = A; D.n_landmarks = L; D.comm_dim = P.cdim; D.collaborative = P.collaborative;
  D.state_comps = P.scomp;
  D.n_goal = P.n_goal;
  for (int i = 0; i < A; ++i) { D.comm_off[i] = P.c_off[i]; D.comm_len[i] = P.c_dim[i]; D.movable[i] = (int)((movable >> i) & 1u); }
  D.state_elem_size = env->cfg.state_f64 ? 8 : 4;
  // bytes_env = 4(4A + c + 2L + sum K) + 4(4A + c + sum D + A) + A
  D.env_bytes_per_step = 4 * (4 * A + P.cdim + 2 * L + D.act_sum) + 4 * (4 * A + P.cdim + D.obs_sum + A) + A;
  return MDP_OK;
}

int env_ensure_cols(mdp_env* env) {
  int dev = -1;
  MDP_CUDA(cudaGetDevice(&dev));
  if (env->d_cols && env->d_cols_device == dev) return MDP_OK;
  if (env->d_cols) cudaFree(env->d_cols);
  env->d_cols = nullptr;
  MDP_CUDA(cudaMalloc(&env->d_cols, env->cols.size() * sizeof(ObsCol)));
  MDP_CUDA(cudaMemcpy(env->d_cols, env->cols.data(), env->cols.size() * sizeof(ObsCol), cudaMemcpyHostToDevice));
  env->d_cols_device = dev;
  return MDP_OK;
}

template <typename real, int EB, bool DO_STEP>
static int launch_step_eb(mdp_env* env, int E, int ES, void* state, const float* act, float* obs, float* rew, uint8_t* done,
                          cudaStream_t st) {
  const EnvParams& P = env->P;
  int NT = round_up(EB * P.A, 32);
  if (NT < 256) NT = 256;
  const size_t smem = EnvTile<real, EB>::bytes(P.scomp, P.A, P.act_stride, true);
  auto kern = k_env_step<real, EB, DO_STEP>;
  if (smem > 48 * 1024) MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<cdiv(E, EB), NT, smem, st>>>(P, E, ES, (real*)state, act, env->d_cols, obs, rew, done);
  return check_launch("k_env_step");
}

// threads per CTA >= EB * A with agent-uniform warps (EB % 32 == 0)
template <int A>
static int launch_spread(mdp_env* env, int E, int ES, float* state, const float* act, float* obs, float* rew, uint8_t* done,
                         cudaStream_t st) {
  const size_t smem = (size_t)SPREAD_EB * ((env->P.obs_stride / 4) | 1) * 16;
  auto kern = k_env_step_spread<A>;
  if (smem > 48 * 1024) MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<cdiv(E, SPREAD_EB), SPREAD_EB, smem, st>>>(env->P, E, ES, state, act, obs, rew, done);
  return check_launch("k_env_step_spread");
}

// simple_spread, 7..32 agents, float32 state: warp-per-env kernel; lay != null fuses the joint replay insert into the step
static bool spread_warp_ok(const mdp_env* env) {
  return !env->cfg.state_f64 && env->P.scenario == MDP_SIMPLE_SPREAD && !env->force_generic && env->P.A > 6 && env->P.A <= 32;
}
static int launch_spread_warp(mdp_env* env, int E, int ES, float* state, const float* act, float* obs, float* rew, uint8_t* done,
                              const float* obs_prev, float* ring, long long capacity, long long cursor, const mdp_ring_layout* lay,
                              cudaStream_t st) {
  const int grid = cdiv(E, 8);
  mdp_ring_layout none;
  memset(&none, 0, sizeof(none));
#define MDP_WARP_LAUNCH(AT)                                                                                                 \
  do {                                                                                                                      \
    if (lay)                                                                                                                \
      k_env_step_spread_warp<AT, true><<<grid, 256, 0, st>>>(env->P, E, ES, state, act, obs, rew, done, obs_prev, ring, capacity, \
                                                             cursor, *lay, env->ctl);                                      \
    else                                                                                                                    \
      k_env_step_spread_warp<AT, false><<<grid, 256, 0, st>>>(env->P, E, ES, state, act, obs, rew, done, nullptr, nullptr, 1, 0,  \
                                                              none, nullptr);                                              \
  } while (0)
  if (env->P.A == 24) MDP_WARP_LAUNCH(24);  // BASELINE configs[4]: every entity loop unrolled
  else MDP_WARP_LAUNCH(0);
#undef MDP_WARP_LAUNCH
  return check_launch("k_env_step_spread_warp");
}

template <typename real, bool DO_STEP>
static int launch_step(mdp_env* env, int E, int ES, void* state, const float* act, float* obs, float* rew, uint8_t* done,
                       cudaStream_t st) {
  if (DO_STEP && sizeof(real) == 4 && spread_warp_ok(env))
    return launch_spread_warp(env, E, ES, (float*)state, act, obs, rew, done, nullptr, nullptr, 1, 0, nullptr, st);
  if (DO_STEP && sizeof(real) == 4 && env->P.scenario == MDP_SIMPLE_SPREAD && !env->force_generic) {
    switch (env->P.A) {
      case 2: return launch_spread<2>(env, E, ES, (float*)state, act, obs, rew, done, st);
      case 3: return launch_spread<3>(env, E, ES, (float*)state, act, obs, rew, done, st);
      case 4: return launch_spread<4>(env, E, ES, (float*)state, act, obs, rew, done, st);
      case 5: return launch_spread<5>(env, E, ES, (float*)state, act, obs, rew, done, st);
      case 6: return launch_spread<6>(env, E, ES, (float*)state, act, obs, rew, done, st);
      default: break;
    }
  }
  if (env->P.A == 1) return launch_step_eb<real, 128, DO_STEP>(env, E, ES, state, act, obs, rew, done, st);
  if (env->P.A == 2) return launch_step_eb<real, 64, DO_STEP>(env, E, ES, state, act, obs, rew, done, st);
  return launch_step_eb<real, 32, DO_STEP>(env, E, ES, state, act, obs, rew, done, st);
}

}  // namespace mdp

using namespace mdp;

extern "C" int mdp_env_create(const mdp_env_cfg* cfg, mdp_env** out) {
  MDP_REQUIRE(cfg && out, "mdp_env_create: null argument");
  mdp_env* env = new (std::nothrow) mdp_env();
  MDP_REQUIRE(env, "mdp_env_create: out of memory");
  env->cfg = *cfg;
  int rc = build_env(env);
  if (rc != MDP_OK) { delete env; return rc; }
  *out = env;
  return MDP_OK;
}

extern "C" int mdp_env_get_dims(const mdp_env* env, mdp_env_dims* out) {
  MDP_REQUIRE(env && out, "mdp_env_get_dims: null argument");
  *out = env->dims;
  return MDP_OK;
}

extern "C" void mdp_env_destroy(mdp_env* env) {
  if (!env) return;
  if (env->d_cols) cudaFree(env->d_cols);
  if (env->pipeline_ready) {
    for (int i = 0; i < mdp_env::kMaxChunks; ++i) {
      cudaStreamDestroy(env->chunk_stream[i]);
      cudaEventDestroy(env->chunk_done[i]);
    }
    cudaEventDestroy(env->fork_ev);
  }
  delete env;
}

extern "C" int mdp_env_force_generic(mdp_env* env, int32_t on) {
  MDP_REQUIRE(env, "mdp_env_force_generic: null env");
  env->force_generic = on ? 1 : 0;
  return MDP_OK;
}

extern "C" int mdp_env_benchmark(mdp_env* env, int32_t E, const void* state, float* out, void* stream) {
  MDP_REQUIRE(env && state && out && E > 0, "mdp_env_benchmark: bad argument");
  const int n = E * env->P.A;
  cudaStream_t st = (cudaStream_t)stream;
  if (env->cfg.state_f64)
    k_env_benchmark<double><<<cdiv(n, 256), 256, 0, st>>>(env->P, E, (const double*)state, out);
  else
    k_env_benchmark<float><<<cdiv(n, 256), 256, 0, st>>>(env->P, E, (const float*)state, out);
  return check_launch("k_env_benchmark");
}

extern "C" int mdp_env_set_ctl(mdp_env* env, const uint64_t* ctl) {
  MDP_REQUIRE(env, "mdp_env_set_ctl: null env");
  env->ctl = reinterpret_cast<const unsigned long long*>(ctl);
  return MDP_OK;
}

extern "C" int mdp_env_reset(mdp_env* env, int32_t E, void* state, const void* init_state, uint64_t seed,
                             uint64_t episode, float* obs_out, void* stream) {
  MDP_REQUIRE(env && state && obs_out && E > 0, "mdp_env_reset: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = env_ensure_cols(env);
  if (rc) return rc;
  const EnvParams& P = env->P;
  const size_t esz = env->cfg.state_f64 ? 8 : 4;
  if (init_state) {
    MDP_CUDA(cudaMemcpyAsync(state, init_state, (size_t)P.scomp * E * esz, cudaMemcpyDeviceToDevice, st));
  } else {
    size_t total = (size_t)P.scomp * E;
    int nb = (int)((total + 255) / 256);
    if (env->cfg.state_f64)
      k_env_reset<double><<<nb, 256, 0, st>>>(P, E, (double*)state, seed, episode, env->reset_lo_lm, env->reset_hi_lm, env->ctl);
    else
      k_env_reset<float><<<nb, 256, 0, st>>>(P, E, (float*)state, seed, episode, env->reset_lo_lm, env->reset_hi_lm, env->ctl);
    rc = check_launch("k_env_reset");
    if (rc) return rc;
  }
  if (env->cfg.state_f64) return launch_step<double, false>(env, E, E, state, nullptr, obs_out, nullptr, nullptr, st);
  return launch_step<float, false>(env, E, E, state, nullptr, obs_out, nullptr, nullptr, st);
}

// env instances [e_base, e_base + n) of a population of E (SoA state stride E): the pointers are those of the whole
// population.  Chunks of one step may run on different streams (mdp_host_step pipelines its copies against them).
int mdp::env_step_range(mdp_env* env, int32_t E, int32_t e_base, int32_t n, void* state, const float* act, float* obs_out,
                        float* rew_out, uint8_t* done_out, const float* obs_prev, float* ring, int64_t ring_capacity,
                        int32_t ring_row_stride, int64_t ring_cursor, void* stream) {
  MDP_REQUIRE(env && state && act && obs_out && rew_out && done_out && E > 0 && e_base >= 0 && n > 0 && e_base + n <= E,
              "mdp_env_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = env_ensure_cols(env);
  if (rc) return rc;
  const mdp_env_dims& d = env->dims;
  act += (size_t)e_base * d.act_stride;
  obs_out += (size_t)e_base * d.obs_stride;
  rew_out += (size_t)e_base * d.n_agents;
  done_out += (size_t)e_base * d.n_agents;
  if (ring && spread_warp_ok(env)) {  // the step kernel writes the replay rows itself
    MDP_REQUIRE(obs_prev && obs_prev != obs_out, "mdp_env_step: ring insert needs a distinct obs_prev buffer");
    mdp_ring_layout lay;
    rc = mdp_ring_make_layout(d.n_agents, d.obs_dim, d.act_dim, &lay);
    if (rc) return rc;
    MDP_REQUIRE(lay.row_stride == ring_row_stride, "mdp_env_step: ring_row_stride %d != layout %d", ring_row_stride,
                lay.row_stride);
    MDP_REQUIRE(ring_capacity > 0 && n <= ring_capacity && ring_cursor >= 0 && ring_cursor < ring_capacity,
                "mdp_env_step: bad ring sizes (capacity %lld, cursor %lld, E %d)", (long long)ring_capacity, (long long)ring_cursor, n);
    return launch_spread_warp(env, n, E, (float*)state + e_base, act, obs_out, rew_out, done_out,
                              obs_prev + (size_t)e_base * d.obs_stride, ring, ring_capacity, (ring_cursor + e_base) % ring_capacity,
                              &lay, st);
  }
  if (env->cfg.state_f64)
    rc = launch_step<double, true>(env, n, E, (double*)state + e_base, act, obs_out, rew_out, done_out, st);
  else
    rc = launch_step<float, true>(env, n, E, (float*)state + e_base, act, obs_out, rew_out, done_out, st);
  if (rc) return rc;
  if (ring) {
    MDP_REQUIRE(obs_prev && obs_prev != obs_out, "mdp_env_step: ring insert needs a distinct obs_prev buffer");
    mdp_ring_layout lay;
    rc = mdp_ring_make_layout(d.n_agents, d.obs_dim, d.act_dim, &lay);
    if (rc) return rc;
    MDP_REQUIRE(lay.row_stride == ring_row_stride, "mdp_env_step: ring_row_stride %d != layout %d", ring_row_stride,
                lay.row_stride);
    return mdp::replay_insert_ctl(&lay, ring, ring_capacity, (ring_cursor + e_base) % ring_capacity, n, -1,
                                  obs_prev + (size_t)e_base * d.obs_stride, d.obs_stride, act, d.act_stride, rew_out, d.n_agents,
                                  obs_out, d.obs_stride, done_out, d.n_agents, env->ctl, stream);
  }
  return MDP_OK;
}

extern "C" int mdp_env_step(mdp_env* env, int32_t E, void* state, const float* act, float* obs_out, float* rew_out,
                            uint8_t* done_out, const float* obs_prev, float* ring, int64_t ring_capacity,
                            int32_t ring_row_stride, int64_t ring_cursor, void* stream) {
  return mdp::env_step_range(env, E, 0, E, state, act, obs_out, rew_out, done_out, obs_prev, ring, ring_capacity, ring_row_stride,
                             ring_cursor, stream);
}
