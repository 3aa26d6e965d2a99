// Fused per-variable clip_by_norm + TF-Adam + polyak target sync + gradient re-zeroing (sm_100a).
//
// Replaces, per SURVEY.md 8(a) rows a11/a12:
//   U.minimize_and_clip              maddpg/common/tf_util.py:166-182  (clip_by_norm PER VARIABLE, 0.5)
//   tf.train.AdamOptimizer(lr)       maddpg/trainer/maddpg.py:129,141  (TF-1.8 formulation, SURVEY B.4:
//                                    eps is added to the uncorrected sqrt(v); lr_t carries the bias terms)
//   make_update_exp (polyak 0.99)    maddpg/trainer/maddpg.py:20-26, invoked :193-194
// One CTA per variable (six per network): pass 1 reduces the squared norm of the (optionally
// 1/world_size-scaled, i.e. all-reduced) gradient, pass 2 streams grad/m/v/param/target once.
// HBM bound: 7 floats read + 5 written per parameter.
#include "mdp_core.cuh"

namespace mdp {

struct OptSeg {
  long long off[6];  // offsets of W1,b1,W2,b2,W3,b3 inside the net block
  long long len[6];
};

// One CTA = one variable.  Latency matters more than bandwidth here (a cfg-2 net has 14 k parameters and the kernel sits
// on the serial path of every agent update), so the second pass's operands (m, v, param, target) are loaded BEFORE the norm
// reduction, and the bias-corrected step size (two double-precision pow) is computed by the last warp while the others reduce.
constexpr int OPT_PF = 4;  // elements per thread held in registers across the reduction (covers len <= 4096 at 1024 threads)

__device__ __forceinline__ void clip_adam_polyak_var(float* __restrict__ g, float* __restrict__ p, float* __restrict__ tg,
                                                     float* __restrict__ mm, float* __restrict__ vv, long long len, int t,
                                                     float grad_scale, float clip, double lr, double beta1, double beta2,
                                                     float eps, float polyak, int do_polyak) {
  __shared__ float red[32];
  __shared__ float s_factor, s_lr_t;
  const int nt = blockDim.x, tid = threadIdx.x;
  if (tid == nt - 32)  // last warp, lane 0: overlaps with the loads / reduction of the other warps
    s_lr_t = (float)(lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t)));
  float gr[OPT_PF], mr[OPT_PF], vr[OPT_PF], pr[OPT_PF], tr[OPT_PF];
#pragma unroll
  for (int k = 0; k < OPT_PF; ++k) {
    const long long i = tid + (long long)k * nt;
    const bool ok = i < len;
    gr[k] = ok ? g[i] : 0.f;
    mr[k] = ok ? mm[i] : 0.f;
    vr[k] = ok ? vv[i] : 0.f;
    pr[k] = ok ? p[i] : 0.f;
    tr[k] = (ok && do_polyak) ? tg[i] : 0.f;
  }
  // pass 1: ||scale * g||_2
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < OPT_PF; ++k) {
    const float x = gr[k] * grad_scale;
    ss = fmaf(x, x, ss);
  }
  for (long long i = tid + (long long)OPT_PF * nt; i < len; i += nt) {
    const float x = g[i] * grad_scale;
    ss = fmaf(x, x, ss);
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  if (tid < 32) {
    float s = tid < (nt >> 5) ? red[tid] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tid == 0) {
      const float norm = sqrtf(s);
      s_factor = clip > 0.f ? clip / fmaxf(norm, clip) : 1.0f;  // tf.clip_by_norm: g * clip / max(||g||, clip)
    }
  }
  __syncthreads();
  const float factor = s_factor * grad_scale;
  const float lr_t = s_lr_t;
  const float b1 = (float)beta1, b2 = (float)beta2, ob1 = (float)(1.0 - beta1), ob2 = (float)(1.0 - beta2);
  const float opol = 1.0f - polyak;
#pragma unroll
  for (int k = 0; k < OPT_PF; ++k) {
    const long long i = tid + (long long)k * nt;
    if (i < len) {
      const float gi = gr[k] * factor;
      const float mi = b1 * mr[k] + ob1 * gi;
      const float vi = b2 * vr[k] + ob2 * gi * gi;
      const float pi = pr[k] - lr_t * mi / (sqrtf(vi) + eps);
      mm[i] = mi;
      vv[i] = vi;
      p[i] = pi;
      if (do_polyak) tg[i] = polyak * tr[k] + opol * pi;
      g[i] = 0.f;  // the *_grads kernels accumulate with atomics: leave the bucket clean for the next round
    }
  }
  for (long long i = tid + (long long)OPT_PF * nt; i < len; i += nt) {
    const float gi = g[i] * factor;
    const float mi = b1 * mm[i] + ob1 * gi;
    const float vi = b2 * vv[i] + ob2 * gi * gi;
    const float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    mm[i] = mi;
    vv[i] = vi;
    p[i] = pi;
    if (do_polyak) tg[i] = polyak * tg[i] + opol * pi;
    g[i] = 0.f;
  }
}

__global__ void __launch_bounds__(1024) k_clip_adam_polyak(float* __restrict__ param, float* __restrict__ target,
                                                           float* __restrict__ grad, float* __restrict__ m,
                                                           float* __restrict__ v, OptSeg seg, const int* __restrict__ t_ptr,
                                                           float grad_scale, float clip, double lr, double beta1,
                                                           double beta2, float eps, float polyak, int do_polyak) {
  const long long off = seg.off[blockIdx.x], len = seg.len[blockIdx.x];
  clip_adam_polyak_var(grad + off, param + off, target + off, m + off, v + off, len, *t_ptr, grad_scale, clip, lr, beta1, beta2,
                       eps, polyak, do_polyak);
}

// all agents in one launch: grid = (6 variables, n_agents); pointers come from the device agent table
__global__ void __launch_bounds__(1024) k_clip_adam_polyak_all(const AgentDev* __restrict__ agents, int which, int units,
                                                               float* __restrict__ grads_base, float* __restrict__ m_base,
                                                               float* __restrict__ v_base, const int* __restrict__ adam_t,
                                                               float grad_scale, float clip, double lr, double beta1,
                                                               double beta2, float eps, float polyak, int do_polyak) {
  const int j = blockIdx.y, var = blockIdx.x;
  const AgentDev& ag = agents[j];
  const MlpW& w = ag.net[which == 0 ? MDP_NET_P : MDP_NET_Q];
  const MlpW& wt = ag.net[which == 0 ? MDP_NET_TARGET_P : MDP_NET_TARGET_Q];
  const long long U = units, in = w.in, out = w.out;
  const long long lens[6] = {in * U, U, U * U, U, U * out, out};
  long long off = 0;
  for (int k = 0; k < var; ++k) off += lens[k];
  const long long len = lens[var];
  float* g = ag.grad[which].W1 + off;
  const long long goff = g - grads_base;
  clip_adam_polyak_var(g, const_cast<float*>(w.W1) + off, const_cast<float*>(wt.W1) + off, m_base + goff, v_base + goff, len,
                       adam_t[2 * j + which], grad_scale, clip, lr, beta1, beta2, eps, polyak, do_polyak);
}

}  // namespace mdp

using namespace mdp;

extern "C" int mdp_clip_adam_polyak_all(mdp_core* c, int32_t which, float grad_scale, int32_t do_polyak, void* stream) {
  MDP_REQUIRE(c && c->d_agents, "mdp_clip_adam_polyak_all: core not bound");
  MDP_REQUIRE(which == 0 || which == 1, "mdp_clip_adam_polyak_all: bad argument");
  k_clip_adam_polyak_all<<<dim3(6, c->cfg.n_agents), 1024, 0, (cudaStream_t)stream>>>(
      c->d_agents, which, c->cfg.num_units, c->grads, c->adam_m, c->adam_v, c->adam_t, grad_scale, (float)c->cfg.grad_clip,
      c->cfg.lr, c->cfg.beta1, c->cfg.beta2, (float)c->cfg.adam_eps, (float)c->cfg.polyak, do_polyak);
  return check_launch("k_clip_adam_polyak_all");
}

extern "C" int mdp_clip_adam_polyak(mdp_core* c, int32_t agent, int32_t which, float grad_scale, int32_t do_polyak,
                                    void* stream) {
  MDP_REQUIRE(c && c->d_agents, "mdp_clip_adam_polyak: core not bound");
  MDP_REQUIRE(agent >= 0 && agent < c->cfg.n_agents && (which == 0 || which == 1), "mdp_clip_adam_polyak: bad argument");
  const int U = c->cfg.num_units;
  const int net = which == 0 ? MDP_NET_P : MDP_NET_Q;
  const int tnet = which == 0 ? MDP_NET_TARGET_P : MDP_NET_TARGET_Q;
  const long long in = c->lay.net_in[agent][net], out = c->lay.net_out[agent][net];
  OptSeg seg;
  const long long lens[6] = {in * U, U, (long long)U * U, U, (long long)U * out, out};
  long long o = 0;
  for (int k = 0; k < 6; ++k) {
    seg.off[k] = o;
    seg.len[k] = lens[k];
    o += lens[k];
  }
  float* param = c->params + c->lay.net_off[agent][net];
  float* target = c->params + c->lay.net_off[agent][tnet];
  float* grad = c->grads + c->lay.train_off[agent][which];
  float* m = c->adam_m + c->lay.train_off[agent][which];
  float* v = c->adam_v + c->lay.train_off[agent][which];
  k_clip_adam_polyak<<<6, 1024, 0, (cudaStream_t)stream>>>(param, target, grad, m, v, seg, c->adam_t + 2 * agent + which,
                                                           grad_scale, (float)c->cfg.grad_clip, c->cfg.lr, c->cfg.beta1,
                                                           c->cfg.beta2, (float)c->cfg.adam_eps, (float)c->cfg.polyak,
                                                           do_polyak);
  return check_launch("k_clip_adam_polyak");
}
