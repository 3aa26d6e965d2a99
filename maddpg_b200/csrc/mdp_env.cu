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
#include "mdp_common.cuh"

#include <vector>
#include <new>

namespace mdp {

enum ObsKind : uint8_t {
  OK_PAD = 0,
  OK_DIRECT = 1,    // S[a]
  OK_REL = 2,       // S[a] - S[b]
  OK_ZERO = 3,      // literal zero (silent agents' comm in simple_spread)
  OK_REL_MASK = 4,  // visible(i,o) ? S[a] - S[b] : 0        (simple_world_comm)
  OK_DIR_MASK = 5,  // visible(i,o) ? S[a] : 0
  OK_FOREST = 6,    // in_forest(i, k) ? +1 : -1
};

struct ObsCol {
  uint8_t kind, i, o, k;
  uint8_t a, b, pad0, pad1;
};

constexpr int MAX_ENT = 2 * MDP_MAX_AGENTS;

struct EnvParams {
  int scenario, A, L, NE, cdim, scomp, obs_stride, act_stride, obs_sum, act_sum;
  int n_adv;       // adversaries are agents [0, n_adv)
  int food0, n_food, forest0, n_forest;  // landmark indices (simple_world_comm)
  int collaborative;
  int act_off[MDP_MAX_AGENTS];
  double size[MAX_ENT];
  double sens[MDP_MAX_AGENTS];       // accel if set else 5.0
  double max_speed[MDP_MAX_AGENTS];  // <= 0: no clamp
  uint64_t collide_mask;             // bit per entity
  uint32_t silent_mask;              // bit per agent
  double dt, damping, contact_force, contact_margin;
};

}  // namespace mdp

struct mdp_env {
  mdp_env_cfg cfg;
  mdp_env_dims dims;
  mdp::EnvParams P;
  std::vector<mdp::ObsCol> cols;
  mdp::ObsCol* d_cols = nullptr;
  int d_cols_device = -1;
  float reset_lo_lm, reset_hi_lm;
};

namespace mdp {

// --------------------------------------------------------------------------------------------
// device helpers
// --------------------------------------------------------------------------------------------
template <typename real> __device__ __forceinline__ real r_sqrt(real x);
template <> __device__ __forceinline__ float r_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double r_sqrt<double>(double x) { return sqrt(x); }
template <typename real> __device__ __forceinline__ real r_exp(real x);
template <> __device__ __forceinline__ float r_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double r_exp<double>(double x) { return exp(x); }
template <typename real> __device__ __forceinline__ real r_log1p(real x);
template <> __device__ __forceinline__ float r_log1p<float>(float x) { return log1pf(x); }
template <> __device__ __forceinline__ double r_log1p<double>(double x) { return log1p(x); }

// numpy.logaddexp(0, z): the soft-contact penetration of World.get_collision_force
template <typename real>
__device__ __forceinline__ real logaddexp0(real z) {
  if (z == (real)0) return (real)0.693147180559945309417232121458;
  if (z < (real)0) return r_log1p<real>(r_exp<real>(z));
  return z + r_log1p<real>(r_exp<real>(-z));
}

// Scenario.bound(x) of simple_tag / simple_world_comm
template <typename real>
__device__ __forceinline__ real bound_pen(real x) {
  if (x < (real)0.9) return (real)0;
  if (x < (real)1.0) return (x - (real)0.9) * (real)10;
  real e = r_exp<real>((real)2 * x - (real)2);
  return e < (real)10 ? e : (real)10;
}

__device__ __forceinline__ int ent_comp(const EnvParams& P, int ent) {
  return ent < P.A ? 4 * ent : 4 * P.A + P.cdim + 2 * (ent - P.A);
}

// Shared memory carve-up: sS[scomp][EBP] real | sR[A][EBP] real | sT[A][EBP] real | sF[A][EBP] int | sA[EB][ASP] float
template <typename real, bool DO_STEP>
__global__ void __launch_bounds__(1024) k_env_step(EnvParams P, int E, int EB, real* __restrict__ state,
                                                   const float* __restrict__ act, const ObsCol* __restrict__ cols,
                                                   float* __restrict__ obs_out, float* __restrict__ rew_out,
                                                   uint8_t* __restrict__ done_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int EBP = EB + 1;
  const int ASP = P.act_stride | 1;
  real* sS = reinterpret_cast<real*>(smem_raw);
  real* sR = sS + (size_t)P.scomp * EBP;
  real* sT = sR + (size_t)P.A * EBP;
  int* sF = reinterpret_cast<int*>(sT + (size_t)P.A * EBP);
  float* sA = reinterpret_cast<float*>(sF + (size_t)P.A * EBP);

  const int tid = threadIdx.x, NT = blockDim.x;
  const int e0 = blockIdx.x * EB;
  const int nE = min(EB, E - e0);

  // ---- phase 0: stage state (and actions) --------------------------------------------------
  for (int idx = tid; idx < P.scomp * EB; idx += NT) {
    int comp = idx / EB, e = idx - comp * EB;
    sS[comp * EBP + e] = (e < nE) ? state[(size_t)comp * E + e0 + e] : (real)0;
  }
  if (DO_STEP) {
    const float* arow = act + (size_t)e0 * P.act_stride;
    for (int idx = tid; idx < nE * P.act_stride; idx += NT) {
      int e = idx / P.act_stride, c = idx - e * P.act_stride;
      sA[e * ASP + c] = arow[idx];
    }
  }
  __syncthreads();

  const int e = tid % EB, i = tid / EB;  // agent-uniform warps
  const bool live = (i < P.A) && (e < nE);

  // ---- phase 1: action decode, forces, integration (World.step) ------------------------------
  if (DO_STEP) {
    real px = 0, py = 0, vx = 0, vy = 0;
    if (live) {
      px = sS[(4 * i + 0) * EBP + e];
      py = sS[(4 * i + 1) * EBP + e];
      vx = sS[(4 * i + 2) * EBP + e];
      vy = sS[(4 * i + 3) * EBP + e];
      const float* a = sA + e * ASP + P.act_off[i];
      // _set_action: float32 differences, then float64 scaling by accel / 5.0
      real fx = (real)(a[1] - a[2]);
      real fy = (real)(a[3] - a[4]);
      fx *= (real)P.sens[i];
      fy *= (real)P.sens[i];
      // apply_environment_force: soft contact with every other collidable entity, partner order
      if ((P.collide_mask >> i) & 1ull) {
        const real k = (real)P.contact_margin, cf = (real)P.contact_force;
        const real si = (real)P.size[i];
        // exact shortcut: exp(z) underflows to +0 below this, so the penetration is exactly 0
        const real zmin = sizeof(real) == 4 ? (real)-104.0 : (real)-746.0;
        for (int j = 0; j < P.NE; ++j) {
          if (j == i || !((P.collide_mask >> j) & 1ull)) continue;
          int cj = ent_comp(P, j);
          real dx = px - sS[cj * EBP + e];
          real dy = py - sS[(cj + 1) * EBP + e];
          real dist = r_sqrt<real>(dx * dx + dy * dy);
          real dmin = si + (real)P.size[j];
          real z = -(dist - dmin) / k;
          if (z < zmin) continue;
          real pen = logaddexp0<real>(z) * k;
          fx = cf * dx / dist * pen + fx;
          fy = cf * dy / dist * pen + fy;
        }
      }
      // integrate_state (mass = 1)
      const real damp = (real)1 - (real)P.damping, dt = (real)P.dt;
      vx = vx * damp;
      vy = vy * damp;
      vx += fx * dt;
      vy += fy * dt;
      real ms = (real)P.max_speed[i];
      if (ms > (real)0) {
        real speed = r_sqrt<real>(vx * vx + vy * vy);
        if (speed > ms) {
          vx = vx / speed * ms;
          vy = vy / speed * ms;
        }
      }
      px += vx * dt;
      py += vy * dt;
    }
    __syncthreads();  // everyone has read the old positions
    if (live) {
      sS[(4 * i + 0) * EBP + e] = px;
      sS[(4 * i + 1) * EBP + e] = py;
      sS[(4 * i + 2) * EBP + e] = vx;
      sS[(4 * i + 3) * EBP + e] = vy;
      // update_agent_state: non-silent agents publish their comm head (simple_world_comm leader)
      if (P.cdim > 0 && !((P.silent_mask >> i) & 1u)) {
        const float* a = sA + e * ASP + P.act_off[i] + 5;
        for (int c = 0; c < P.cdim; ++c) sS[(4 * P.A + c) * EBP + e] = (real)a[c];
      }
    }
    __syncthreads();
  }

  // ---- phase 2: scenario rewards / visibility flags from the post-integration state -----------
  auto PX = [&](int ent) -> real { return sS[ent_comp(P, ent) * EBP + e]; };
  auto PY = [&](int ent) -> real { return sS[(ent_comp(P, ent) + 1) * EBP + e]; };
  auto dist_ee = [&](int a, int b) -> real {
    real dx = PX(a) - PX(b), dy = PY(a) - PY(b);
    return r_sqrt<real>(dx * dx + dy * dy);
  };
  auto collide_ee = [&](int a, int b) -> bool { return dist_ee(a, b) < (real)(P.size[a] + P.size[b]); };

  if (P.scenario == MDP_SIMPLE_WORLD_COMM) {
    if (live) {
      int f = 0;
      for (int q = 0; q < P.n_forest; ++q)
        if (collide_ee(i, P.A + P.forest0 + q)) f |= (1 << q);
      sF[i * EBP + e] = f;
    }
  }
  if (DO_STEP) {
    if (live) {
      real r = 0;
      if (P.scenario == MDP_SIMPLE) {
        real dx = PX(0) - PX(P.A), dy = PY(0) - PY(P.A);
        r = -(dx * dx + dy * dy);
      } else if (P.scenario == MDP_SIMPLE_SPREAD) {
        // thread i owns landmark i: min over agents of the distance (the "occupied landmark" term)
        real m = dist_ee(0, P.A + i);
        for (int a = 1; a < P.A; ++a) {
          real d = dist_ee(a, P.A + i);
          m = d < m ? d : m;
        }
        sT[i * EBP + e] = m;
        int cnt = 0;  // includes a == i (distance 0 < 2*size): the reference's constant -1
        for (int a = 0; a < P.A; ++a) cnt += collide_ee(a, i) ? 1 : 0;
        r = (real)cnt;  // finished after the sync below
      } else if (P.scenario == MDP_SIMPLE_TAG) {
        if (i < P.n_adv) {
          for (int g = P.n_adv; g < P.A; ++g)
            for (int a = 0; a < P.n_adv; ++a)
              if (collide_ee(g, a)) r += (real)10;
        } else {
          for (int a = 0; a < P.n_adv; ++a)
            if (collide_ee(a, i)) r -= (real)10;
          real ax = PX(i), ay = PY(i);
          r -= bound_pen<real>(ax < 0 ? -ax : ax);
          r -= bound_pen<real>(ay < 0 ? -ay : ay);
        }
      } else {  // MDP_SIMPLE_WORLD_COMM
        if (i < P.n_adv) {
          real m = dist_ee(P.n_adv, i);
          for (int g = P.n_adv + 1; g < P.A; ++g) {
            real d = dist_ee(g, i);
            m = d < m ? d : m;
          }
          r -= (real)0.1 * m;
          for (int g = P.n_adv; g < P.A; ++g)
            for (int a = 0; a < P.n_adv; ++a)
              if (collide_ee(g, a)) r += (real)5;
        } else {
          for (int a = 0; a < P.n_adv; ++a)
            if (collide_ee(a, i)) r -= (real)5;
          real ax = PX(i), ay = PY(i);
          r -= (real)2 * bound_pen<real>(ax < 0 ? -ax : ax);
          r -= (real)2 * bound_pen<real>(ay < 0 ? -ay : ay);
          real m = 0;
          for (int q = 0; q < P.n_food; ++q) {
            int fe = P.A + P.food0 + q;
            if (collide_ee(i, fe)) r += (real)2;
            real d = dist_ee(fe, i);
            m = (q == 0 || d < m) ? d : m;
          }
          r += (real)0.05 * m;
        }
      }
      sR[i * EBP + e] = r;
    }
    if (P.scenario == MDP_SIMPLE_SPREAD) {
      __syncthreads();
      real r = 0;
      if (live) {
        for (int l = 0; l < P.L; ++l) r -= sT[l * EBP + e];
        r -= sR[i * EBP + e];  // collision count
      }
      __syncthreads();
      if (live) sR[i * EBP + e] = r;
    }
  }
  __syncthreads();

  // ---- phase 3: stream out ---------------------------------------------------------------------
  if (DO_STEP) {
    // movable state (agents' pos/vel + comm); landmarks never move
    const int wcomp = 4 * P.A + P.cdim;
    for (int idx = tid; idx < wcomp * EB; idx += NT) {
      int comp = idx / EB, ee = idx - comp * EB;
      if (ee < nE) state[(size_t)comp * E + e0 + ee] = sS[comp * EBP + ee];
    }
    // rewards (shared reward: every agent receives the sum, environment.py step())
    for (int idx = tid; idx < nE * P.A; idx += NT) {
      int ee = idx / P.A, ii = idx - ee * P.A;
      real r;
      if (P.collaborative) {
        r = 0;
        for (int a = 0; a < P.A; ++a) r += sR[a * EBP + ee];
      } else {
        r = sR[ii * EBP + ee];
      }
      rew_out[(size_t)(e0 + ee) * P.A + ii] = (float)r;
      done_out[(size_t)(e0 + ee) * P.A + ii] = 0;  // MPE has no done callback: always False
    }
  }
  // observations: one warp per env row, lanes sweep the joint columns (coalesced 128-byte stores)
  const int warp = tid >> 5, lane = tid & 31, nwarps = NT >> 5;
  for (int ee = warp; ee < nE; ee += nwarps) {
    float* orow = obs_out + (size_t)(e0 + ee) * P.obs_stride;
    for (int c = lane; c < P.obs_stride; c += 32) {
      ObsCol d = cols[c];
      real v = 0;
      switch (d.kind) {
        case OK_DIRECT: v = sS[d.a * EBP + ee]; break;
        case OK_REL: v = sS[d.a * EBP + ee] - sS[d.b * EBP + ee]; break;
        case OK_REL_MASK:
        case OK_DIR_MASK: {
          int fi = sF[d.i * EBP + ee], fo = sF[d.o * EBP + ee];
          bool inf1 = fi & 1, inf2 = fi & 2, of1 = fo & 1, of2 = fo & 2;
          bool vis = (inf1 && of1) || (inf2 && of2) || (!inf1 && !of1 && !inf2 && !of2) || (d.i == 0);
          if (vis) v = (d.kind == OK_REL_MASK) ? sS[d.a * EBP + ee] - sS[d.b * EBP + ee] : sS[d.a * EBP + ee];
          break;
        }
        case OK_FOREST: v = ((sF[d.i * EBP + ee] >> d.k) & 1) ? (real)1 : (real)-1; break;
        default: break;
      }
      orow[c] = (float)v;
    }
  }
}

// scenario.reset_world: agents U(-1,1), velocities / comm 0, landmarks U(lo,hi)
template <typename real>
__global__ void k_env_reset(EnvParams P, int E, real* __restrict__ state, uint64_t seed, uint64_t episode,
                            float lm_lo, float lm_hi) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)P.scomp * E;
  if (idx >= total) return;
  int comp = (int)(idx / E);
  int e = (int)(idx - (size_t)comp * E);
  real v = 0;
  bool agent_pos = comp < 4 * P.A && (comp & 3) < 2;
  bool lm_pos = comp >= 4 * P.A + P.cdim;
  if (agent_pos || lm_pos) {
    uint4 r = Philox::gen(seed, (uint32_t)e, (uint32_t)comp, (uint32_t)episode, (uint32_t)(episode >> 32) ^ 0x5EEDu);
    real u = sizeof(real) == 4 ? (real)Philox::u01(r.x) : (real)Philox::u01d(r.x, r.y);
    real lo = agent_pos ? (real)-1 : (real)lm_lo, hi = agent_pos ? (real)1 : (real)lm_hi;
    v = lo + (hi - lo) * u;
  }
  state[idx] = v;
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
  env->reset_lo_lm = -1.f; env->reset_hi_lm = 1.f;
  P.silent_mask = 0xffffffffu;
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
    P.silent_mask &= ~1u;  // agent 0 (leader) speaks
    P.size[A + 0] = 0.2; P.collide_mask |= 1ull << (A + 0);
    P.size[A + 1] = P.size[A + 2] = 0.03;
    P.size[A + 3] = P.size[A + 4] = 0.3;
    env->reset_lo_lm = -0.9f; env->reset_hi_lm = 0.9f;
  } else {
    return fail(MDP_ENOTSUP, "unknown scenario id %d", sc);
  }
  P.A = A; P.L = L; P.NE = A + L;
  P.scomp = 4 * A + P.cdim + 2 * L;
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
  P.obs_stride = D.obs_stride; P.act_stride = D.act_stride; P.obs_sum = D.obs_sum; P.act_sum = D.act_sum;
  D.n_agents = A; D.n_landmarks = L; D.comm_dim = P.cdim; D.collaborative = P.collaborative;
  D.state_comps = P.scomp;
  D.state_elem_size = env->cfg.state_f64 ? 8 : 4;
  // bytes_env = 4(4A + c + 2L + sum K) + 4(4A + c + sum D + A) + A
  D.env_bytes_per_step = 4 * (4 * A + P.cdim + 2 * L + D.act_sum) + 4 * (4 * A + P.cdim + D.obs_sum + A) + A;
  return MDP_OK;
}

static int ensure_cols(mdp_env* env) {
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

static int pick_eb(const EnvParams& P) {
  // threads per CTA = EB * A, agent-uniform warps need EB % 32 == 0
  if (P.A == 1) return 128;
  if (P.A <= 2) return 64;
  return 32;
}

template <typename real, bool DO_STEP>
static int launch_step(mdp_env* env, int E, void* state, const float* act, float* obs, float* rew, uint8_t* done,
                       cudaStream_t st) {
  const EnvParams& P = env->P;
  const int EB = pick_eb(P), EBP = EB + 1, ASP = P.act_stride | 1;
  const int NT = EB * P.A;
  size_t smem = ((size_t)P.scomp + 2 * P.A) * EBP * sizeof(real) + (size_t)P.A * EBP * sizeof(int) +
                (size_t)EB * ASP * sizeof(float) + 16;
  auto kern = k_env_step<real, DO_STEP>;
  if (smem > 48 * 1024) MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<cdiv(E, EB), NT, smem, st>>>(P, E, EB, (real*)state, act, env->d_cols, obs, rew, done);
  return check_launch("k_env_step");
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
  delete env;
}

extern "C" int mdp_env_reset(mdp_env* env, int32_t E, void* state, const void* init_state, uint64_t seed,
                             uint64_t episode, float* obs_out, void* stream) {
  MDP_REQUIRE(env && state && obs_out && E > 0, "mdp_env_reset: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_cols(env);
  if (rc) return rc;
  const EnvParams& P = env->P;
  const size_t esz = env->cfg.state_f64 ? 8 : 4;
  if (init_state) {
    MDP_CUDA(cudaMemcpyAsync(state, init_state, (size_t)P.scomp * E * esz, cudaMemcpyDeviceToDevice, st));
  } else {
    size_t total = (size_t)P.scomp * E;
    int nb = (int)((total + 255) / 256);
    if (env->cfg.state_f64)
      k_env_reset<double><<<nb, 256, 0, st>>>(P, E, (double*)state, seed, episode, env->reset_lo_lm, env->reset_hi_lm);
    else
      k_env_reset<float><<<nb, 256, 0, st>>>(P, E, (float*)state, seed, episode, env->reset_lo_lm, env->reset_hi_lm);
    rc = check_launch("k_env_reset");
    if (rc) return rc;
  }
  if (env->cfg.state_f64) return launch_step<double, false>(env, E, state, nullptr, obs_out, nullptr, nullptr, st);
  return launch_step<float, false>(env, E, state, nullptr, obs_out, nullptr, nullptr, st);
}

extern "C" int mdp_env_step(mdp_env* env, int32_t E, void* state, const float* act, float* obs_out, float* rew_out,
                            uint8_t* done_out, const float* obs_prev, float* ring, int64_t ring_capacity,
                            int32_t ring_row_stride, int64_t ring_cursor, void* stream) {
  MDP_REQUIRE(env && state && act && obs_out && rew_out && done_out && E > 0, "mdp_env_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_cols(env);
  if (rc) return rc;
  if (env->cfg.state_f64)
    rc = launch_step<double, true>(env, E, state, act, obs_out, rew_out, done_out, st);
  else
    rc = launch_step<float, true>(env, E, state, act, obs_out, rew_out, done_out, st);
  if (rc) return rc;
  if (ring) {
    MDP_REQUIRE(obs_prev && obs_prev != obs_out, "mdp_env_step: ring insert needs a distinct obs_prev buffer");
    mdp_ring_layout lay;
    rc = mdp_ring_make_layout(env->dims.n_agents, env->dims.obs_dim, env->dims.act_dim, &lay);
    if (rc) return rc;
    MDP_REQUIRE(lay.row_stride == ring_row_stride, "mdp_env_step: ring_row_stride %d != layout %d", ring_row_stride,
                lay.row_stride);
    return mdp_replay_insert(&lay, ring, ring_capacity, ring_cursor, E, -1, obs_prev, env->dims.obs_stride, act,
                             env->dims.act_stride, rew_out, env->dims.n_agents, obs_out, env->dims.obs_stride,
                             done_out, env->dims.n_agents, stream);
  }
  return MDP_OK;
}
