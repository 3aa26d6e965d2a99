// Device-side MPE physics shared by the per-step kernel (mdp_env.cu) and the persistent episode
// kernel (mdp_rollout.cu).  A CTA owns a tile of EB consecutive env instances whose SoA state lives
// in shared memory as [component][env]; all functions here are CTA-collective (every thread of the
// block must call them; they synchronise internally where noted).
//
// Semantics follow SURVEY.md Appendix A (upstream openai/multiagent-particle-envs):
//   env_physics        MultiAgentEnv._set_action + World.step (apply_action_force,
//                      apply_environment_force/get_collision_force, integrate_state, update_agent_state)
//   env_flags_rewards  Scenario.reward (+ shared-reward sum of MultiAgentEnv.step) and the forest
//                      visibility flags of simple_world_comm's observation()
//   env_obs_value      one column of Scenario.observation(), driven by the per-scenario column table
#pragma once
#include "mdp_common.cuh"

#include <vector>

namespace mdp {

enum ObsKind : uint8_t {
  OK_PAD = 0,
  OK_DIRECT = 1,    // S[a]
  OK_REL = 2,       // S[a] - S[b]
  OK_ZERO = 3,      // literal zero (silent agents' comm in simple_spread)
  OK_REL_MASK = 4,  // visible(i,o) ? S[a] - S[b] : 0        (simple_world_comm)
  OK_DIR_MASK = 5,  // visible(i,o) ? S[a] : 0
  OK_FOREST = 6,    // in_forest(i, k) ? +1 : -1
  OK_REL_GOAL = 7,  // S[o + 2 * goal + a] - S[b]: goal = (int)S[k], o = first landmark component   (agent.goal_a.state.p_pos - p_pos)
  OK_GOAL_LUT = 8,  // kColorLut[nibble `goal` of (i | o << 8)], goal = (int)S[k]                     (goal-dependent colour channel)
  OK_CONST = 9,     // kColorLut[a]                                                                 (fixed colour channel)
};

// every colour value the scenarios' observations contain (entity.color channels, one-hot "colours" of simple_crypto)
__device__ __constant__ float kColorLut[8] = {0.0f, 0.1f, 0.15f, 0.25f, 0.65f, 0.75f, 0.9f, 1.0f};
enum { LUT_0 = 0, LUT_010 = 1, LUT_015 = 2, LUT_025 = 3, LUT_065 = 4, LUT_075 = 5, LUT_090 = 6, LUT_1 = 7 };

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
  float sizef[MAX_ENT];
  double sens[MDP_MAX_AGENTS];       // accel if set else 5.0
  double max_speed[MDP_MAX_AGENTS];  // <= 0: no clamp
  uint64_t collide_mask;             // bit per entity
  uint32_t silent_mask;              // bit per agent
  int has_mask;                      // scenario has visibility-masked / flag columns (simple_world_comm)
  uint32_t movable_mask;             // bit per agent: World.integrate_state skips the others (speakers, simple_crypto)
  int n_goal, gcomp0;                // goal landmark indices drawn by reset_world, stored as state components [gcomp0, gcomp0 + n_goal)
  int c_off[MDP_MAX_AGENTS];         // agent.state.c of a speaking agent: components [4A + c_off, 4A + c_off + c_dim)
  int c_dim[MDP_MAX_AGENTS];         // 0 for silent agents; cdim = sum
  uint32_t as4_magic;                // ceil(2^32 / (act_stride/4)): row = (i * magic) >> 32 for i < 2^15
  uint32_t os4_magic;                // same for obs_stride/4
  double dt, damping, contact_force, contact_margin;
};

template <typename real> __device__ __forceinline__ real r_sqrt(real x);
template <> __device__ __forceinline__ float r_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double r_sqrt<double>(double x) { return sqrt(x); }
template <typename real> __device__ __forceinline__ real r_exp(real x);
template <> __device__ __forceinline__ float r_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double r_exp<double>(double x) { return exp(x); }
template <typename real> __device__ __forceinline__ real r_log1p(real x);
template <> __device__ __forceinline__ float r_log1p<float>(float x) { return log1pf(x); }
template <> __device__ __forceinline__ double r_log1p<double>(double x) { return log1p(x); }
template <typename real> __device__ __forceinline__ real ent_size(const EnvParams& P, int j);
template <> __device__ __forceinline__ float ent_size<float>(const EnvParams& P, int j) { return P.sizef[j]; }
template <> __device__ __forceinline__ double ent_size<double>(const EnvParams& P, int j) { return P.size[j]; }

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

// Shared-memory tile of EB env instances.
template <typename real, int EB>
struct EnvTile {
  static constexpr int EBP = EB + 1;
  real* sS;   // [scomp][EBP] state
  real* sR;   // [A][EBP] per-agent reward
  real* sT;   // [A][EBP] scratch (simple_spread landmark minima)
  int* sF;    // [A][EBP] forest flags (simple_world_comm)
  float* sA;  // [EB][ASP] joint actions of the tile
  int ASP;

  __host__ __device__ static size_t bytes(int scomp, int A, int act_stride, bool with_actions) {
    size_t b = ((size_t)scomp + 1 + 2 * A) * EBP * sizeof(real) + (size_t)A * EBP * sizeof(int);
    if (with_actions) b += (size_t)EB * (act_stride | 1) * sizeof(float);
    return (b + 15) & ~(size_t)15;
  }
  __device__ void carve(void* base, const EnvParams& P, float* actions_or_null) {
    sS = reinterpret_cast<real*>(base);
    sR = sS + (size_t)(P.scomp + 1) * EBP;  // row P.scomp of sS is all zeros (the 'b' operand of plain columns)
    sT = sR + (size_t)P.A * EBP;
    sF = reinterpret_cast<int*>(sT + (size_t)P.A * EBP);
    ASP = P.act_stride | 1;
    sA = actions_or_null ? actions_or_null : reinterpret_cast<float*>(sF + (size_t)P.A * EBP);
  }
};

template <typename real, int EB>
__device__ __forceinline__ void env_load_state(const EnvParams& P, EnvTile<real, EB>& T, const real* __restrict__ state,
                                               int E, int e0, int nE) {
  constexpr int EBP = EB + 1;
  for (int idx = threadIdx.x; idx < (P.scomp + 1) * EB; idx += blockDim.x) {
    const int comp = idx / EB, e = idx % EB;  // EB is a power of two: shifts
    T.sS[comp * EBP + e] = (e < nE && comp < P.scomp) ? state[(size_t)comp * E + e0 + e] : (real)0;
  }
}

// movable state only (agents' pos/vel + comm); landmarks never move
template <typename real, int EB>
__device__ __forceinline__ void env_store_state(const EnvParams& P, const EnvTile<real, EB>& T, real* __restrict__ state,
                                                int E, int e0, int nE, bool all_comps) {
  constexpr int EBP = EB + 1;
  const int wcomp = all_comps ? P.scomp : 4 * P.A + P.cdim;
  for (int idx = threadIdx.x; idx < wcomp * EB; idx += blockDim.x) {
    const int comp = idx / EB, e = idx % EB;
    if (e < nE) state[(size_t)comp * E + e0 + e] = T.sS[comp * EBP + e];
  }
}

template <typename real, int EB>
__device__ __forceinline__ void env_load_actions(const EnvParams& P, EnvTile<real, EB>& T, const float* __restrict__ act,
                                                 int e0, int nE) {
  // the tile's action rows are contiguous in global memory: one flat float4 sweep, row = i / (act_stride/4)
  const float4* src = reinterpret_cast<const float4*>(act + (size_t)e0 * P.act_stride);
  const int as4 = P.act_stride >> 2;
  for (int i = threadIdx.x; i < nE * as4; i += blockDim.x) {
    const int e = P.as4_magic ? (int)__umulhi((unsigned)i, P.as4_magic) : i;
    const int q = i - e * as4;
    const float4 v = src[i];
    float* d = T.sA + e * T.ASP + 4 * q;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

// World.step for the tile: thread tid < EB*A owns (env e = tid % EB, agent i = tid / EB) so warps are
// agent-uniform.  Contains two __syncthreads; the new state is visible to the whole CTA on return.
template <typename real, int EB>
__device__ __forceinline__ void env_physics(const EnvParams& P, EnvTile<real, EB>& T, int nE) {
  constexpr int EBP = EB + 1;
  const int tid = threadIdx.x;
  const int e = tid % EB, i = tid / EB;
  const bool live = (i < P.A) && (e < nE);
  real px = 0, py = 0, vx = 0, vy = 0;
  const bool movable = (P.movable_mask >> i) & 1u;
  if (live) {
    real* sS = T.sS;
    px = sS[(4 * i + 0) * EBP + e];
    py = sS[(4 * i + 1) * EBP + e];
    vx = sS[(4 * i + 2) * EBP + e];
    vy = sS[(4 * i + 3) * EBP + e];
  }
  if (live && movable) {
    real* sS = T.sS;
    const float* a = T.sA + e * T.ASP + P.act_off[i];
    // _set_action: float32 differences, then scaling by accel / 5.0 in the state precision
    real fx = (real)(a[1] - a[2]);
    real fy = (real)(a[3] - a[4]);
    const real sens = (real)P.sens[i];
    fx *= sens;
    fy *= sens;
    // apply_environment_force: soft contact with every other collidable entity, in partner order
    if ((P.collide_mask >> i) & 1ull) {
      const real k = (real)P.contact_margin, cf = (real)P.contact_force;
      const real si = ent_size<real>(P, i);
      // exact shortcut: exp(z) underflows to +0 below this, so the penetration is exactly 0
      const real zmin = sizeof(real) == 4 ? (real)-104.0 : (real)-746.0;
      for (int j = 0; j < P.NE; ++j) {
        if (j == i || !((P.collide_mask >> j) & 1ull)) continue;
        const int cj = ent_comp(P, j);
        const real dx = px - sS[cj * EBP + e];
        const real dy = py - sS[(cj + 1) * EBP + e];
        const real dist = r_sqrt<real>(dx * dx + dy * dy);
        const real dmin = si + ent_size<real>(P, j);
        const real z = -(dist - dmin) / k;
        if (z < zmin) continue;
        const real pen = logaddexp0<real>(z) * k;
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
    const real ms = (real)P.max_speed[i];
    if (ms > (real)0) {
      const real speed = r_sqrt<real>(vx * vx + vy * vy);
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
    real* sS = T.sS;
    sS[(4 * i + 0) * EBP + e] = px;
    sS[(4 * i + 1) * EBP + e] = py;
    sS[(4 * i + 2) * EBP + e] = vx;
    sS[(4 * i + 3) * EBP + e] = vy;
    // update_agent_state: non-silent agents publish their comm head (simple_world_comm leader)
    if (P.c_dim[i] > 0) {  // the communication head follows the movement head (if the agent has one) in its action block
      const float* a = T.sA + e * T.ASP + P.act_off[i] + (movable ? 5 : 0);
      for (int c = 0; c < P.c_dim[i]; ++c) sS[(4 * P.A + P.c_off[i] + c) * EBP + e] = (real)a[c];
    }
  }
  __syncthreads();
}

// Scenario rewards (DO_REWARD) and simple_world_comm forest flags from the current tile state.
// On return sR[i][e] holds agent i's own reward (before the shared-reward sum).  Ends synchronised.
template <typename real, int EB, bool DO_REWARD>
__device__ __forceinline__ void env_flags_rewards(const EnvParams& P, EnvTile<real, EB>& T, int nE) {
  constexpr int EBP = EB + 1;
  const int tid = threadIdx.x;
  const int e = tid % EB, i = tid / EB;
  const bool live = (i < P.A) && (e < nE);
  real* sS = T.sS;
  auto PX = [&](int ent) -> real { return sS[ent_comp(P, ent) * EBP + e]; };
  auto PY = [&](int ent) -> real { return sS[(ent_comp(P, ent) + 1) * EBP + e]; };
  auto dist_ee = [&](int a, int b) -> real {
    const real dx = PX(a) - PX(b), dy = PY(a) - PY(b);
    return r_sqrt<real>(dx * dx + dy * dy);
  };
  auto collide_ee = [&](int a, int b) -> bool { return dist_ee(a, b) < ent_size<real>(P, a) + ent_size<real>(P, b); };

  if (P.scenario == MDP_SIMPLE_WORLD_COMM && live) {
    int f = 0;
    for (int q = 0; q < P.n_forest; ++q)
      if (collide_ee(i, P.A + P.forest0 + q)) f |= (1 << q);
    T.sF[i * EBP + e] = f;
  }
  if (DO_REWARD) {
    if (live) {
      real r = 0;
      if (P.scenario == MDP_SIMPLE) {
        const real dx = PX(0) - PX(P.A), dy = PY(0) - PY(P.A);
        r = -(dx * dx + dy * dy);
      } else if (P.scenario == MDP_SIMPLE_SPREAD) {
        // thread i owns landmark i: min over agents of the distance (the "occupied landmark" term)
        real m = dist_ee(0, P.A + i);
        for (int a = 1; a < P.A; ++a) {
          const real d = dist_ee(a, P.A + i);
          m = d < m ? d : m;
        }
        T.sT[i * EBP + e] = m;
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
          const real ax = PX(i), ay = PY(i);
          r -= bound_pen<real>(ax < 0 ? -ax : ax);
          r -= bound_pen<real>(ay < 0 ? -ay : ay);
        }
      } else if (P.scenario == MDP_SIMPLE_ADVERSARY) {
        const int ge = P.A + (int)sS[P.gcomp0 * EBP + e];  // agent.goal_a
        if (i < P.n_adv) {  // adversary_reward (shaped): -sum(square(p_pos - goal))
          const real dx = PX(i) - PX(ge), dy = PY(i) - PY(ge);
          r = -(dx * dx + dy * dy);
        } else {            // agent_reward (shaped): -min over good agents of dist + sum over adversaries of dist
          real adv = 0;
          for (int a = 0; a < P.n_adv; ++a) adv += dist_ee(a, ge);
          real m = dist_ee(P.n_adv, ge);
          for (int g = P.n_adv + 1; g < P.A; ++g) {
            const real d = dist_ee(g, ge);
            m = d < m ? d : m;
          }
          r = -m + adv;
        }
      } else if (P.scenario == MDP_SIMPLE_PUSH) {
        const int ge = P.A + (int)sS[P.gcomp0 * EBP + e];
        if (i < P.n_adv) {  // keep the nearest good agent away from the goal, stay close to it
          real m = dist_ee(P.n_adv, ge);
          for (int g = P.n_adv + 1; g < P.A; ++g) {
            const real d = dist_ee(g, ge);
            m = d < m ? d : m;
          }
          r = m - dist_ee(ge, i);
        } else {
          r = -dist_ee(i, ge);
        }
      } else if (P.scenario == MDP_SIMPLE_SPEAKER_LISTENER) {
        const int ge = P.A + (int)sS[P.gcomp0 * EBP + e];  // the speaker's goal_b; goal_a is the listener (agent 1)
        const real dx = PX(1) - PX(ge), dy = PY(1) - PY(ge);
        r = -(dx * dx + dy * dy);
      } else if (P.scenario == MDP_SIMPLE_REFERENCE) {
        // agent i wants the OTHER agent (goal_a) on landmark goal_b_i: -sum(square(goal_a.p_pos - goal_b.p_pos))
        const int ge = P.A + (int)sS[(P.gcomp0 + i) * EBP + e];
        const real dx = PX(1 - i) - PX(ge), dy = PY(1 - i) - PY(ge);
        r = -(dx * dx + dy * dy);
      } else if (P.scenario == MDP_SIMPLE_CRYPTO) {
        const int g = (int)sS[P.gcomp0 * EBP + e];  // goal landmark: its "colour" is the one-hot of its index in dim_c channels
        auto all_zero = [&](int a) -> bool {
          bool z = true;
          for (int c = 0; c < P.c_dim[a]; ++c) z = z && (sS[(4 * P.A + P.c_off[a] + c) * EBP + e] == (real)0);
          return z;
        };
        auto sq_err = [&](int a) -> real {
          real s = 0;
          for (int c = 0; c < P.c_dim[a]; ++c) {
            const real d = sS[(4 * P.A + P.c_off[a] + c) * EBP + e] - (c == g ? (real)1 : (real)0);
            s += d * d;
          }
          return s;
        };
        if (i < P.n_adv) {
          if (!all_zero(i)) r -= sq_err(i);
        } else {  // good listeners: not adversary, not speaker (the speaker is the last agent)
          real good = 0, adv = 0;
          for (int a = P.n_adv; a < P.A - 1; ++a)
            if (!all_zero(a)) good -= sq_err(a);
          for (int a = 0; a < P.n_adv; ++a)
            if (!all_zero(a)) adv += sq_err(a);
          r = adv + good;
        }
      } else {  // MDP_SIMPLE_WORLD_COMM
        if (i < P.n_adv) {
          real m = dist_ee(P.n_adv, i);
          for (int g = P.n_adv + 1; g < P.A; ++g) {
            const real d = dist_ee(g, i);
            m = d < m ? d : m;
          }
          r -= (real)0.1 * m;
          for (int g = P.n_adv; g < P.A; ++g)
            for (int a = 0; a < P.n_adv; ++a)
              if (collide_ee(g, a)) r += (real)5;
        } else {
          for (int a = 0; a < P.n_adv; ++a)
            if (collide_ee(a, i)) r -= (real)5;
          const real ax = PX(i), ay = PY(i);
          r -= (real)2 * bound_pen<real>(ax < 0 ? -ax : ax);
          r -= (real)2 * bound_pen<real>(ay < 0 ? -ay : ay);
          real m = 0;
          for (int q = 0; q < P.n_food; ++q) {
            const int fe = P.A + P.food0 + q;
            if (collide_ee(i, fe)) r += (real)2;
            const real d = dist_ee(fe, i);
            m = (q == 0 || d < m) ? d : m;
          }
          r += (real)0.05 * m;
        }
      }
      T.sR[i * EBP + e] = r;
    }
    if (P.scenario == MDP_SIMPLE_SPREAD) {
      __syncthreads();
      real r = 0;
      if (live) {
        for (int l = 0; l < P.L; ++l) r -= T.sT[l * EBP + e];
        r -= T.sR[i * EBP + e];  // collision count
      }
      __syncthreads();
      if (live) T.sR[i * EBP + e] = r;
    }
  }
  __syncthreads();
  if (P.scenario == MDP_SIMPLE_WORLD_COMM) {
    // visibility of every other agent, once per (env, agent) instead of once per observation column: o is visible to i when both
    // stand in the same forest, or neither stands in any, or i is the leader (simple_world_comm.py observation())
    if (live) {
      const int fi = T.sF[i * EBP + e] & 3;
      int vis = 0;
      for (int o = 0; o < P.A; ++o) {
        const int fo = T.sF[o * EBP + e] & 3;
        if ((fi & fo) || (fi == 0 && fo == 0) || i == 0) vis |= 1 << o;
      }
      T.sF[i * EBP + e] = fi | (vis << 8);  // the other threads only read bits 0-1, which do not change
    }
    __syncthreads();
  }
}

// reward agent ii of env ee receives (shared reward: the sum over agents, environment.py step())
template <typename real, int EB>
__device__ __forceinline__ float env_reward_out(const EnvParams& P, const EnvTile<real, EB>& T, int ee, int ii) {
  constexpr int EBP = EB + 1;
  if (P.collaborative) {
    real r = 0;
    for (int a = 0; a < P.A; ++a) r += T.sR[a * EBP + ee];
    return (float)r;
  }
  return (float)T.sR[ii * EBP + ee];
}

// one observation column of env ee (joint layout), from the tile state
template <typename real, int EB>
__device__ __forceinline__ float env_obs_value(const EnvTile<real, EB>& T, const ObsCol d, int ee) {
  constexpr int EBP = EB + 1;
  const real* sS = T.sS;
  if (d.kind <= OK_ZERO) return (float)(sS[d.a * EBP + ee] - sS[d.b * EBP + ee]);  // plain columns: b may be the zero row
  real v = 0;
  switch (d.kind) {
    case OK_DIRECT: v = sS[d.a * EBP + ee]; break;
    case OK_REL: v = sS[d.a * EBP + ee] - sS[d.b * EBP + ee]; break;
    case OK_REL_MASK:
    case OK_DIR_MASK: {
      // bit 8 + o of agent i's flag word: "i sees o" (env_flags_rewards); OK_DIR_MASK columns carry b = the zero row
      if ((T.sF[d.i * EBP + ee] >> (8 + d.o)) & 1) v = sS[d.a * EBP + ee] - sS[d.b * EBP + ee];
      break;
    }
    case OK_FOREST: v = ((T.sF[d.i * EBP + ee] >> d.k) & 1) ? (real)1 : (real)-1; break;
    case OK_REL_GOAL: {
      const int g = (int)sS[d.k * EBP + ee];
      v = sS[(d.o + 2 * g + d.a) * EBP + ee] - sS[d.b * EBP + ee];
      break;
    }
    case OK_GOAL_LUT: {
      const int g = (int)sS[d.k * EBP + ee];
      return kColorLut[(((unsigned)d.i | ((unsigned)d.o << 8)) >> (4 * g)) & 7u];
    }
    case OK_CONST: return kColorLut[d.a & 7];
    default: break;
  }
  return (float)v;
}

// Joint observation rows of the tile -> dst (row pitch ld floats, 16-byte aligned rows): every thread
// produces four consecutive columns of one row and stores them as one float4 (flat (row, quad) sweep).
template <typename real, int EB>
__device__ __forceinline__ void env_write_obs(const EnvParams& P, const EnvTile<real, EB>& T, const ObsCol* __restrict__ cols,
                                              float* __restrict__ dst, long long ld, int nE) {
  const int os4 = P.obs_stride >> 2;
  for (int i = threadIdx.x; i < nE * os4; i += blockDim.x) {
    const int ee = P.os4_magic ? (int)__umulhi((unsigned)i, P.os4_magic) : i;  // magic 0 <=> divisor 1
    const int q = i - ee * os4;
    const uint4 t0 = *reinterpret_cast<const uint4*>(cols + 4 * q);      // columns 4q, 4q+1
    const uint4 t1 = *reinterpret_cast<const uint4*>(cols + 4 * q + 2);  // columns 4q+2, 4q+3
    ObsCol d[4];
    *reinterpret_cast<uint2*>(&d[0]) = make_uint2(t0.x, t0.y);
    *reinterpret_cast<uint2*>(&d[1]) = make_uint2(t0.z, t0.w);
    *reinterpret_cast<uint2*>(&d[2]) = make_uint2(t1.x, t1.y);
    *reinterpret_cast<uint2*>(&d[3]) = make_uint2(t1.z, t1.w);
    float4 v;
    v.x = env_obs_value<real, EB>(T, d[0], ee);
    v.y = env_obs_value<real, EB>(T, d[1], ee);
    v.z = env_obs_value<real, EB>(T, d[2], ee);
    v.w = env_obs_value<real, EB>(T, d[3], ee);
    *reinterpret_cast<float4*>(dst + (long long)ee * ld + 4 * q) = v;
  }
}

// reset_world draw for state component `comp` of env `e_global` (agents U(-1,1), landmarks U(lo,hi))
template <typename real>
__device__ __forceinline__ real env_reset_value(const EnvParams& P, int comp, long long e_global, uint64_t seed,
                                                uint64_t episode, float lm_lo, float lm_hi) {
  const bool agent_pos = comp < 4 * P.A && (comp & 3) < 2;
  const bool lm_pos = comp >= 4 * P.A + P.cdim && comp < 4 * P.A + P.cdim + 2 * P.L;
  const bool goal = P.n_goal > 0 && comp >= P.gcomp0;
  if (!(agent_pos || lm_pos || goal)) return (real)0;
  uint4 r = Philox::gen(seed, (uint32_t)e_global, (uint32_t)comp, (uint32_t)episode, (uint32_t)(episode >> 32) ^ 0x5EEDu);
  const real u = sizeof(real) == 4 ? (real)Philox::u01(r.x) : (real)Philox::u01d(r.x, r.y);
  if (goal) {  // np.random.choice(world.landmarks): a uniform landmark index
    const int g = (int)(u * (real)P.L);
    return (real)(g < P.L ? g : P.L - 1);
  }
  const real lo = agent_pos ? (real)-1 : (real)lm_lo, hi = agent_pos ? (real)1 : (real)lm_hi;
  return lo + (hi - lo) * u;
}

// --------------------------------------------------------------------------------------------
// simple_spread in registers (float32 state, A agents == A landmarks known at compile time): one thread owns one env
// instance.  Used by k_env_step_spread (mdp_env.cu) and by the persistent episode kernel (mdp_rollout.cu).
// Arithmetic and operation order follow env_physics / env_flags_rewards / env_obs_value exactly.
// --------------------------------------------------------------------------------------------
template <int A>
struct SpreadRegs {
  float px[A], py[A], vx[A], vy[A], lx[A], ly[A];
};

// dx*dx + dy*dy with the contraction pinned (one multiply, one fused multiply-add): the pieces below are inlined into
// different kernels and must round identically in all of them
__device__ __forceinline__ float sq_norm2(float dx, float dy) { return __fmaf_rn(dx, dx, __fmul_rn(dy, dy)); }

// scenario constants as floats, converted once per kernel (EnvParams keeps the reference's float64 values)
template <int A>
struct SpreadConsts {
  float k, cf, damp, dt, size[A];
};
struct SpreadAgentConsts {
  float sens, ms, size;
};
template <int A>
__device__ __forceinline__ SpreadConsts<A> spread_consts(const EnvParams& P) {
  SpreadConsts<A> c;
  c.k = (float)P.contact_margin; c.cf = (float)P.contact_force;
  c.damp = 1.0f - (float)P.damping; c.dt = (float)P.dt;
#pragma unroll
  for (int j = 0; j < A; ++j) c.size[j] = P.sizef[j];
  return c;
}
__device__ __forceinline__ SpreadAgentConsts spread_agent_consts(const EnvParams& P, int i) {
  SpreadAgentConsts a;
  a.sens = (float)P.sens[i]; a.ms = (float)P.max_speed[i]; a.size = P.sizef[i];
  return a;
}

// The step is written as per-agent pieces so that it can run either entirely in one thread (k_env_step_spread: i is a
// compile-time constant after unrolling) or spread over one thread per (env instance, agent) (episode kernel: i = warp index).

// World.step for agent i: action force, soft contact with the other agents (landmarks do not collide in simple_spread),
// damping, integration.  px/py = positions of ALL agents BEFORE the step; (pxi, pyi, vxi, vyi) = agent i, updated in place;
// a = agent i's 5 action floats.
template <int A, typename ActT>
__device__ __forceinline__ void spread_agent_step(const SpreadConsts<A>& Cn, const SpreadAgentConsts& Ai, int i, const float (&px)[A],
                                                  const float (&py)[A], float& pxi, float& pyi, float& vxi, float& vyi,
                                                  const ActT& a) {
  const float k = Cn.k, cf = Cn.cf, zmin = -104.0f;
  const float damp = Cn.damp, dt = Cn.dt;
  const float sens = Ai.sens;
  float fx = __fmul_rn(__fsub_rn(a[1], a[2]), sens);
  float fy = __fmul_rn(__fsub_rn(a[3], a[4]), sens);
  const float si = Ai.size;
#pragma unroll
  for (int j = 0; j < A; ++j) {
    if (j == i) continue;
    const float dx = __fsub_rn(pxi, px[j]), dy = __fsub_rn(pyi, py[j]);
    const float dist = __fsqrt_rn(sq_norm2(dx, dy));
    const float dmin = __fadd_rn(si, Cn.size[j]);
    const float z = __fdiv_rn(-__fsub_rn(dist, dmin), k);
    if (z < zmin) continue;
    const float pen = __fmul_rn(logaddexp0<float>(z), k);
    fx = __fmaf_rn(__fdiv_rn(__fmul_rn(cf, dx), dist), pen, fx);
    fy = __fmaf_rn(__fdiv_rn(__fmul_rn(cf, dy), dist), pen, fy);
  }
  float wx = __fmaf_rn(fx, dt, __fmul_rn(vxi, damp));
  float wy = __fmaf_rn(fy, dt, __fmul_rn(vyi, damp));
  const float ms = Ai.ms;
  if (ms > 0.f) {
    const float speed = __fsqrt_rn(sq_norm2(wx, wy));
    if (speed > ms) {
      wx = __fmul_rn(__fdiv_rn(wx, speed), ms);
      wy = __fmul_rn(__fdiv_rn(wy, speed), ms);
    }
  }
  vxi = wx; vyi = wy;
  pxi = __fmaf_rn(wx, dt, pxi);
  pyi = __fmaf_rn(wy, dt, pyi);
}

// Scenario.reward pieces (positions AFTER the step): distance of the closest agent to one landmark ...
template <int A>
__device__ __forceinline__ float spread_landmark_min(const float (&px)[A], const float (&py)[A], float lx, float ly) {
  float best = 0.f;
#pragma unroll
  for (int q = 0; q < A; ++q) {
    const float dx = __fsub_rn(px[q], lx), dy = __fsub_rn(py[q], ly);
    const float d = __fsqrt_rn(sq_norm2(dx, dy));
    best = (q == 0 || d < best) ? d : best;
  }
  return best;
}
// ... the collision count of agent i (self included) ...
template <int A>
__device__ __forceinline__ int spread_collisions(const SpreadConsts<A>& Cn, const SpreadAgentConsts& Ai, const float (&px)[A],
                                                 const float (&py)[A], float pxi, float pyi) {
  int cnt = 0;
#pragma unroll
  for (int q = 0; q < A; ++q) {
    const float dx = __fsub_rn(px[q], pxi), dy = __fsub_rn(py[q], pyi);
    cnt += (__fsqrt_rn(sq_norm2(dx, dy)) < __fadd_rn(Cn.size[q], Ai.size)) ? 1 : 0;
  }
  return cnt;
}
// ... and the shared reward: every agent receives the sum over agents of (-sum_l m[l] - cnt[i])
template <int A>
__device__ __forceinline__ float spread_reward_sum(const float (&m)[A], const int (&cnt)[A]) {
  float msum = 0.f;
#pragma unroll
  for (int i = 0; i < A; ++i) {
    float ri = 0.f;
#pragma unroll
    for (int l = 0; l < A; ++l) ri -= m[l];
    ri -= (float)cnt[i];
    msum += ri;
  }
  return msum;
}

// Scenario.observation of agent i: [vel, pos, landmarks - pos, others - pos, silent comm zeros], 6A floats through out(c, v)
template <int A, typename Out>
__device__ __forceinline__ void spread_obs_agent(int i, const float (&px)[A], const float (&py)[A], float pxi, float pyi, float vxi,
                                                 float vyi, const float (&lx)[A], const float (&ly)[A], Out&& out) {
  constexpr int L = A, D = 6 * A;
  out(0, vxi); out(1, vyi); out(2, pxi); out(3, pyi);
#pragma unroll
  for (int l = 0; l < L; ++l) { out(4 + 2 * l, lx[l] - pxi); out(5 + 2 * l, ly[l] - pyi); }
#pragma unroll
  for (int q = 0; q < A; ++q) {
    if (q == i) continue;
    const int c = 4 + 2 * L + 2 * (q < i ? q : q - 1);
    out(c, px[q] - pxi);
    out(c + 1, py[q] - pyi);
  }
#pragma unroll
  for (int c = 4 + 2 * L + 2 * (A - 1); c < D; ++c) out(c, 0.f);
}

// World.step + Scenario.reward in one thread: advances S in place, returns the shared reward.
// a = the joint action row, 5 floats per agent.
template <int A, typename ActT>
__device__ __forceinline__ float spread_step(const EnvParams& P, SpreadRegs<A>& S, const ActT& a) {
  const SpreadConsts<A> Cn = spread_consts<A>(P);
  float npx[A], npy[A], nvx[A], nvy[A];
#pragma unroll
  for (int i = 0; i < A; ++i) {
    npx[i] = S.px[i]; npy[i] = S.py[i]; nvx[i] = S.vx[i]; nvy[i] = S.vy[i];
    const float ai[5] = {a[5 * i], a[5 * i + 1], a[5 * i + 2], a[5 * i + 3], a[5 * i + 4]};
    spread_agent_step<A>(Cn, spread_agent_consts(P, i), i, S.px, S.py, npx[i], npy[i], nvx[i], nvy[i], ai);
  }
#pragma unroll
  for (int i = 0; i < A; ++i) { S.px[i] = npx[i]; S.py[i] = npy[i]; S.vx[i] = nvx[i]; S.vy[i] = nvy[i]; }
  float m[A];
  int cnt[A];
#pragma unroll
  for (int l = 0; l < A; ++l) m[l] = spread_landmark_min<A>(S.px, S.py, S.lx[l], S.ly[l]);
#pragma unroll
  for (int i = 0; i < A; ++i) cnt[i] = spread_collisions<A>(Cn, spread_agent_consts(P, i), S.px, S.py, S.px[i], S.py[i]);
  return spread_reward_sum<A>(m, cnt);
}

// every agent's observation; o holds >= A * 6A floats
template <int A, typename ObsT>
__device__ __forceinline__ void spread_obs(const SpreadRegs<A>& S, ObsT& o) {
#pragma unroll
  for (int i = 0; i < A; ++i)
    spread_obs_agent<A>(i, S.px, S.py, S.px[i], S.py[i], S.vx[i], S.vy[i], S.lx, S.ly, [&](int c, float v) { o[i * 6 * A + c] = v; });
}

}  // namespace mdp

struct mdp_env {
  mdp_env_cfg cfg;
  mdp_env_dims dims;
  mdp::EnvParams P;
  std::vector<mdp::ObsCol> cols;
  mdp::ObsCol* d_cols = nullptr;
  int d_cols_device = -1;
  float reset_lo_lm, reset_hi_lm;
  const unsigned long long* ctl = nullptr;
  int force_generic = 0;  // 1: always the table-driven kernel (tests compare the two)
  // mdp_host_step pipeline: one stream per chunk of env instances, fork / join events (created on first use)
  static constexpr int kMaxChunks = 8;
  cudaStream_t chunk_stream[kMaxChunks] = {};
  cudaEvent_t chunk_done[kMaxChunks] = {};
  cudaEvent_t fork_ev = nullptr;
  int pipeline_ready = 0;
  int host_copy_mode = 0;  // mdp_host_step: 0 = copy engines (cudaMemcpyAsync), 1 = copy kernels over the unified address space
};


namespace mdp {
// uploads the observation column table to the current device (lazy; not graph-capturable)
int env_ensure_cols(mdp_env* env);
}  // namespace mdp
