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

#include <algorithm>
#include <stdlib.h>

namespace mdp {

struct OptSeg {
  long long off[6];  // offsets of W1,b1,W2,b2,W3,b3 inside the net block
  long long len[6];
};

// One CTA = one variable.  Latency matters more than bandwidth here (a cfg-2 net has 14 k parameters and the kernel sits
// on the serial path of every agent update), so the second pass's operands (m, v, param, target) are loaded BEFORE the norm
// reduction, and the bias-corrected step size (two double-precision pow) is computed by the last warp while the others reduce.
constexpr int OPT_PF = 4;  // elements per thread held in registers across the reduction (covers len <= 4096 at 1024 threads)
constexpr int PEER_MAXW = 8;  // flag words per slot: one per rank of an 8-GPU box

// Fused all-reduce: when `world` > 1 the gradient of element i is the sum over ranks r = 0..world-1 (fixed order, so every
// replica computes bit-identical parameters) of peer_grads[r][goff + i], read straight from the peers' HBM over NVLink
// (buffers mapped by torch's symmetric memory).  Two flag barriers per CTA, scoped to the CTAs that own the same variable
// on every rank: (1) before the reads -- a rank's flag implies its gradient kernels are complete (stream order);
// (2) after the reads -- only then may a rank zero its own bucket.  Flags carry a monotonically growing epoch, so the
// kernel is replayable from a CUDA graph without host involvement.
//
// Low-latency mode (recv != null; the default for the small MADDPG buckets): instead of barrier + pull, every rank PUSHES
// its gradient values into the peers' receive buffers as 8-byte (value, epoch) words -- one posted NVLink store per element
// and peer -- and each rank polls its own receive buffer until the epoch tag matches: data and flag travel together (the
// protocol NCCL calls LL), so the exchange costs one NVLink traversal, no barrier, no fence, and a rank may zero its own
// bucket immediately because nobody reads it remotely.  Receive buffers are double-buffered by epoch parity: a peer can run
// at most one launch ahead of the slowest rank for a given variable.
struct PeerCtx {
  int world, rank;
  const float* const* grads;  // [world] bases of the ranks' flat gradient buffers
  unsigned* const* flags;     // [world] bases of the ranks' flag words ([slot][PEER_MAXW])
  unsigned* epoch;            // local per-slot epoch counters
  uint2* const* recv;         // [world] bases of the receive buffers [parity][source rank][total] of (value bits, epoch)
  long long total;            // floats in the flat gradient buffer
};

__device__ __forceinline__ void ll_push(const PeerCtx& P, long long goff, long long i, float x, unsigned epoch) {
  const long long at = ((long long)((epoch & 1u) * P.world + P.rank)) * P.total + goff + i;
  for (int r = 0; r < P.world; ++r) {
    if (r == P.rank) continue;
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(P.recv[r] + at), "r"(__float_as_uint(x)), "r"(epoch) : "memory");
  }
}
// sum over ranks in rank order: own value from registers, the peers' from this rank's receive buffer (spins on the epoch tag)
__device__ __forceinline__ float ll_sum(const PeerCtx& P, long long goff, long long i, float own, unsigned epoch) {
  float s = 0.f;
  for (int r = 0; r < P.world; ++r) {
    if (r == P.rank) { s += own; continue; }
    const uint2* src = P.recv[P.rank] + ((long long)((epoch & 1u) * P.world + r)) * P.total + goff + i;
    unsigned v, e;
    long long t0 = 0;
    for (unsigned spins = 1;; ++spins) {
      asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(e) : "l"(src) : "memory");
      if (e == epoch) break;
      if ((spins & 255u) == 0) {  // a missing peer must fail loudly (~4 s), never hang the box
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 8000000000ll) __trap();
      }
    }
    s += __uint_as_float(v);
  }
  return s;
}

// ll_sum for the OPT_PF register-prefetched elements of a thread (tid + k nt), with the polls of two source ranks x OPT_PF
// elements in flight at once.  One poll is an L2 round trip of a system-scope load (~0.7 us); polling element after element and
// rank after rank (the first version) serialised 7 x 4 of them per thread: ~19 us of a 27 us exchange on 8 GPUs.
// The sum still runs over the ranks in rank order (own value at its rank's position): replicas stay bit-identical.
template <int PF>
__device__ __forceinline__ void ll_sum_pf(const PeerCtx& P, long long goff, int tid, int nt, long long len, float (&gr)[PF], unsigned epoch) {
  float own[PF];
#pragma unroll
  for (int k = 0; k < PF; ++k) { own[k] = gr[k]; gr[k] = 0.f; }
  for (int r0 = 0; r0 < P.world; r0 += 2) {
    unsigned v[2][PF], e[2][PF];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const int r = r0 + d;
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        const long long i = tid + (long long)k * nt;
        e[d][k] = epoch; v[d][k] = 0u;
        if (r < P.world && r != P.rank && i < len) {
          const uint2* src = P.recv[P.rank] + ((long long)((epoch & 1u) * P.world + r)) * P.total + goff + i;
          asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[d][k]), "=r"(e[d][k]) : "l"(src) : "memory");
        }
      }
    }
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const int r = r0 + d;
      if (r >= P.world) break;
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        const long long i = tid + (long long)k * nt;
        if (i >= len) continue;
        if (r == P.rank) { gr[k] += own[k]; continue; }
        if (e[d][k] != epoch) {  // not there yet: spin on this word
          const uint2* src = P.recv[P.rank] + ((long long)((epoch & 1u) * P.world + r)) * P.total + goff + i;
          long long t0 = 0;
          for (unsigned spins = 1;; ++spins) {
            asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[d][k]), "=r"(e[d][k]) : "l"(src) : "memory");
            if (e[d][k] == epoch) break;
            if ((spins & 255u) == 0) {  // a missing peer must fail loudly (~4 s), never hang the box
              if (t0 == 0) t0 = clock64();
              else if (clock64() - t0 > 8000000000ll) __trap();
            }
          }
        }
        gr[k] += __uint_as_float(v[d][k]);
      }
    }
  }
}

// flag = epoch on every peer (posted NVLink stores); relaxed: what the flag publishes is already complete when it is written
// (gradient kernels finished in stream order / peer loads consumed by the norm reduction)
__device__ __forceinline__ void peer_signal(const PeerCtx& P, int slot, unsigned epoch) {
  if ((int)threadIdx.x < P.world) {
    unsigned* dst = P.flags[threadIdx.x] + slot * PEER_MAXW + P.rank;
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
  }
}
// every peer's flag in this rank's flag words has reached epoch (ends with a CTA barrier)
__device__ __forceinline__ void peer_wait(const PeerCtx& P, int slot, unsigned epoch) {
  if ((int)threadIdx.x < P.world) {
    const unsigned* src = P.flags[P.rank] + slot * PEER_MAXW + threadIdx.x;
    long long t0 = 0;
    for (unsigned spins = 1;; ++spins) {
      unsigned v;
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if ((spins & 255u) == 0) {  // a missing peer must fail loudly (~4 s), never hang the box
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 8000000000ll) __trap();
      }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  __syncthreads();
}

__device__ __forceinline__ float peer_grad(const PeerCtx& P, const float* g_local, long long goff, long long i) {
  if (P.world <= 1) return g_local[i];
  float s = 0.f;
  for (int r = 0; r < P.world; ++r) s += __ldcv(P.grads[r] + goff + i);
  return s;
}

__device__ __forceinline__ void clip_adam_polyak_var(float* __restrict__ g, float* __restrict__ p, float* __restrict__ tg,
                                                     float* __restrict__ mm, float* __restrict__ vv, long long len,
                                                     const int* __restrict__ t_ptr, float grad_scale, float clip, double lr, double beta1, double beta2,
                                                     float eps, float polyak, int do_polyak, const PeerCtx& P, long long goff,
                                                     int slot) {
  __shared__ float red[32];
  __shared__ float s_factor, s_lr_t;
  const int nt = blockDim.x, tid = threadIdx.x;
  unsigned epoch = 0;
  const bool small = len <= (long long)OPT_PF * nt;  // every peer read of this variable fits the register prefetch
  const bool ll = P.world > 1 && P.recv != nullptr;
  // operands the gradient kernel before this launch does not write (Adam slots, running and target parameters): loaded BEFORE
  // the dependency wait, so that under a programmatic dependent launch they are in flight while that kernel still runs
  float gr[OPT_PF], mr[OPT_PF], vr[OPT_PF], pr[OPT_PF], tr[OPT_PF];
#pragma unroll
  for (int k = 0; k < OPT_PF; ++k) {
    const long long i = tid + (long long)k * nt;
    const bool ok = i < len;
    mr[k] = ok ? mm[i] : 0.f;
    vr[k] = ok ? vv[i] : 0.f;
    pr[k] = ok ? p[i] : 0.f;
    tr[k] = (ok && do_polyak) ? tg[i] : 0.f;
  }
  pdl_wait();  // gradients (and the step counter) of the preceding launch are complete from here on
  pdl_launch_dependents();  // a dependent launch (the actor step after the critic's optimizer) may stage its own inputs now: every
                            // kernel before this one has completed
  const int t = *t_ptr;
  if (P.world > 1) {
    // one slot (flag words + epoch counter) per (agent, net, variable): the actor and critic steps of an agent never share
    // epochs, so a critic-only or reordered step sequence cannot alias another net's exchange (all ranks must still issue the
    // SAME sequence of optimizer launches).  Barrier mode uses epoch ("my gradients are complete") and epoch + 1 ("I have read yours")
    epoch = P.epoch[slot] + 1u;
    if (ll) {
      // push this rank's values first (posted stores), then everything below overlaps their flight
      for (long long i = tid; i < len; i += nt) ll_push(P, goff, i, g[i], epoch);
    } else {
      peer_signal(P, slot, epoch);
      peer_wait(P, slot, epoch);
    }
  }
  if (tid == nt - 32)  // last warp, lane 0: overlaps with the loads / reduction of the other warps
    s_lr_t = (float)(lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t)));
#pragma unroll
  for (int k = 0; k < OPT_PF; ++k) {
    const long long i = tid + (long long)k * nt;
    gr[k] = !(i < len) ? 0.f : ll ? g[i] : peer_grad(P, g, goff, i);
  }
  if (ll) ll_sum_pf(P, goff, tid, nt, len, gr, epoch);
  // pass 1: ||scale * g||_2
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < OPT_PF; ++k) {
    const float x = gr[k] * grad_scale;
    ss = fmaf(x, x, ss);
  }
  for (long long i = tid + (long long)OPT_PF * nt; i < len; i += nt) {
    const float x = (ll ? ll_sum(P, goff, i, g[i], epoch) : peer_grad(P, g, goff, i)) * grad_scale;
    ss = fmaf(x, x, ss);
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();  // every thread's peer loads have returned (their values fed the reduction)
  if (P.world > 1 && !ll && small) peer_signal(P, slot, epoch + 1u);  // early: the peers' wait overlaps this rank's Adam math
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
      if (P.world <= 1 || ll) g[i] = 0.f;  // the *_grads kernels accumulate with atomics: leave the bucket clean for the next round
    }
  }
  for (long long i = tid + (long long)OPT_PF * nt; i < len; i += nt) {
    const float gi = (ll ? ll_sum(P, goff, i, g[i], epoch) : peer_grad(P, g, goff, i)) * factor;
    const float mi = b1 * mm[i] + ob1 * gi;
    const float vi = b2 * vv[i] + ob2 * gi * gi;
    const float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    mm[i] = mi;
    vv[i] = vi;
    p[i] = pi;
    if (do_polyak) tg[i] = polyak * tg[i] + opol * pi;
    if (P.world <= 1 || ll) g[i] = 0.f;
  }
  if (ll) {
    if (tid == 0) P.epoch[slot] = epoch;
  } else if (P.world > 1) {
    if (!small) {
      __syncthreads();
      peer_signal(P, slot, epoch + 1u);
    }
    peer_wait(P, slot, epoch + 1u);  // every rank has read this bucket
    for (long long i = tid; i < len; i += nt) g[i] = 0.f;
    if (tid == 0) P.epoch[slot] = epoch + 1u;
  }
}

__global__ void __launch_bounds__(1024) k_clip_adam_polyak(float* __restrict__ param, float* __restrict__ target,
                                                           float* __restrict__ grad, float* __restrict__ m,
                                                           float* __restrict__ v, OptSeg seg, const int* __restrict__ t_ptr,
                                                           float grad_scale, float clip, double lr, double beta1,
                                                           double beta2, float eps, float polyak, int do_polyak, PeerCtx P,
                                                           long long goff_net, int slot0, int var0) {
  const int var = blockIdx.x + var0;
  const long long off = seg.off[var], len = seg.len[var];
  clip_adam_polyak_var(grad + off, param + off, target + off, m + off, v + off, len, t_ptr, grad_scale, clip, lr, beta1, beta2,
                       eps, polyak, do_polyak, P, goff_net + off, slot0 + var);
}


constexpr int W1_MAXB = 128;  // CTAs per variable of the wide-W1 path (norm partials per agent)

// Wide layer-1 weights (simple_spread N=24: 3576 x 64 per critic): one CTA per variable would stream 229 k parameters through
// 1024 threads (455 us for 24 critics, the longest kernel of the round), so W1 gets a two-kernel path on a single GPU:
// squared norm by many CTAs (one float atomic per CTA), then the clip + Adam + polyak sweep by many CTAs.
// With peers bound (barrier protocol, large buckets): every CTA first joins the flag barrier "all ranks' gradients are complete",
// then sums its chunk over the ranks with peer loads (rank order: bit-identical replicas) and leaves the SUM in the local
// scratch `reduced` for k_w1_adam -- the single-CTA-per-variable kernel pulled the 229 k-float W1 of a simple_spread N=24 critic
// through one CTA (~1.1 ms per agent on 8 GPUs; the whole sequential round ran at 34 % of the exchange-free rate).
__global__ void __launch_bounds__(256) k_w1_sqnorm(const AgentDev* __restrict__ agents, int which, int units, int a0, float grad_scale,
                                                   float* __restrict__ norm2, PeerCtx P, const float* __restrict__ grads_base,
                                                   float* __restrict__ reduced) {
  __shared__ float red[8];
  const int j = a0 + blockIdx.y;
  const AgentDev& ag = agents[j];
  const long long len = (long long)ag.net[which == 0 ? MDP_NET_P : MDP_NET_Q].in * units;
  const float4* g4 = reinterpret_cast<const float4*>(ag.grad[which].W1);
  const long long goff = ag.grad[which].W1 - grads_base;
  if (P.world > 1) {
    const int slot = 12 * j + 6 * which;
    const unsigned epoch = P.epoch[slot] + 1u;
    peer_signal(P, slot, epoch);
    peer_wait(P, slot, epoch);
  }
  float ss = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len / 4; i += (long long)gridDim.x * blockDim.x) {
    float4 x;
    if (P.world > 1) {
      x = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < P.world; ++r) {
        const float4 y = __ldcv(reinterpret_cast<const float4*>(P.grads[r] + goff) + i);
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
      }
      reinterpret_cast<float4*>(reduced + goff)[i] = x;
    } else {
      x = g4[i];
    }
    x.x *= grad_scale; x.y *= grad_scale; x.z *= grad_scale; x.w *= grad_scale;
    ss = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, ss))));
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) norm2[j * W1_MAXB + blockIdx.x] = s;  // per-CTA partial: k_w1_adam adds them in a fixed order, so the
  }                                                              // norm (and with it every replica's parameters) is reproducible
}

__global__ void __launch_bounds__(256) k_w1_adam(const AgentDev* __restrict__ agents, int which, int units, int a0,
                                                 float* __restrict__ grads_base, float* __restrict__ m_base, float* __restrict__ v_base,
                                                 const int* __restrict__ adam_t, const float* __restrict__ norm2, float grad_scale,
                                                 float clip, double lr, double beta1, double beta2, float eps, float polyak,
                                                 int do_polyak, PeerCtx P, const float* __restrict__ reduced) {
  __shared__ float s_lr_t, s_norm2;
  const int j = a0 + blockIdx.y;
  const int slot = 12 * j + 6 * which;
  unsigned epoch = 0;
  if (P.world > 1) {  // k_w1_sqnorm has completed (stream order): this rank has read every peer's bucket
    epoch = P.epoch[slot] + 1u;
    peer_signal(P, slot, epoch + 1u);
  }
  const AgentDev& ag = agents[j];
  const MlpW& w = ag.net[which == 0 ? MDP_NET_P : MDP_NET_Q];
  const MlpW& wt = ag.net[which == 0 ? MDP_NET_TARGET_P : MDP_NET_TARGET_Q];
  const long long len = (long long)w.in * units;
  if (threadIdx.x == 0) {
    const int t = adam_t[2 * j + which];
    s_lr_t = (float)(lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t)));
  }
  if (threadIdx.x == 32) {
    float n2 = 0.f;
    for (int b = 0; b < (int)gridDim.x; ++b) n2 += norm2[j * W1_MAXB + b];
    s_norm2 = n2;
  }
  __syncthreads();
  const float norm = sqrtf(s_norm2);
  const float factor = (clip > 0.f ? clip / fmaxf(norm, clip) : 1.0f) * grad_scale, lr_t = s_lr_t;
  const float b1 = (float)beta1, b2 = (float)beta2, ob1 = (float)(1.0 - beta1), ob2 = (float)(1.0 - beta2), opol = 1.0f - polyak;
  float4* g4 = reinterpret_cast<float4*>(ag.grad[which].W1);
  const long long goff4 = (ag.grad[which].W1 - grads_base) / 4;
  const float4* gs4 = P.world > 1 ? reinterpret_cast<const float4*>(reduced) + goff4 : g4;  // the rank-summed gradient
  float4* m4 = reinterpret_cast<float4*>(m_base) + goff4;
  float4* v4 = reinterpret_cast<float4*>(v_base) + goff4;
  float4* p4 = reinterpret_cast<float4*>(const_cast<float*>(w.W1));
  float4* t4 = reinterpret_cast<float4*>(const_cast<float*>(wt.W1));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len / 4; i += (long long)gridDim.x * blockDim.x) {
    float4 g = gs4[i], m = m4[i], v = v4[i], p = p4[i];
    float4 t = do_polyak ? t4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    auto step = [&](float gi, float& mi, float& vi, float& pi, float& ti) {
      gi *= factor;
      mi = b1 * mi + ob1 * gi;
      vi = b2 * vi + ob2 * gi * gi;
      pi = pi - lr_t * mi / (sqrtf(vi) + eps);
      ti = polyak * ti + opol * pi;
    };
    step(g.x, m.x, v.x, p.x, t.x); step(g.y, m.y, v.y, p.y, t.y); step(g.z, m.z, v.z, p.z, t.z); step(g.w, m.w, v.w, p.w, t.w);
    m4[i] = m; v4[i] = v; p4[i] = p;
    if (do_polyak) t4[i] = t;
    if (P.world <= 1) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (P.world > 1) {  // every rank has read this rank's bucket: only now may it be re-zeroed for the next round
    peer_wait(P, slot, epoch + 1u);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len / 4; i += (long long)gridDim.x * blockDim.x)
      g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// the epoch counters of the W1 slots advance by 2 once every CTA of k_w1_adam is done (stream order)
__global__ void k_w1_epoch_bump(unsigned* __restrict__ epoch, int which, int a0, int count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) epoch[12 * (a0 + k) + 6 * which] += 2u;
}

// all agents in one launch: grid = (6 variables, n_agents); pointers come from the device agent table
__global__ void __launch_bounds__(1024) k_clip_adam_polyak_all(const AgentDev* __restrict__ agents, int which, int units,
                                                               float* __restrict__ grads_base, float* __restrict__ m_base,
                                                               float* __restrict__ v_base, const int* __restrict__ adam_t,
                                                               float grad_scale, float clip, double lr, double beta1,
                                                               double beta2, float eps, float polyak, int do_polyak, PeerCtx P,
                                                               int a0, int var0) {
  const int j = a0 + blockIdx.y, var = blockIdx.x + var0;
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
                       adam_t + 2 * j + which, grad_scale, clip, lr, beta1, beta2, eps, polyak, do_polyak, P, goff, 12 * j + 6 * which + var);
}

}  // namespace mdp

using namespace mdp;

// Launch configuration of the optimizer kernels.  pdl: programmatic dependent launch on the preceding kernel of the stream (the
// gradient kernel): this launch's CTAs become resident and prefetch their Adam slots / parameters while it still runs, and
// block in pdl_wait() until its gradients are complete.  MDP_PDL=0 in the environment keeps plain stream order.
static void pdl_config(cudaLaunchConfig_t* lc, cudaLaunchAttribute* at, dim3 grid, int block, cudaStream_t st, bool pdl) {
  static const bool enabled = []() { const char* e = getenv("MDP_PDL"); return !(e && e[0] == '0'); }();
  memset(lc, 0, sizeof(*lc));
  lc->gridDim = grid; lc->blockDim = dim3(block); lc->dynamicSmemBytes = 0; lc->stream = st;
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc->attrs = at;
  lc->numAttrs = (pdl && enabled) ? 1 : 0;
}

static PeerCtx peer_ctx(const mdp_core* c) {
  PeerCtx P;
  P.world = c->peer_world; P.rank = c->peer_rank;
  P.grads = c->d_peer_grads; P.flags = c->d_peer_flags; P.epoch = c->peer_epoch;
  P.recv = c->d_peer_recv; P.total = c->lay.total_train;
  return P;
}

extern "C" int mdp_core_bind_peers(mdp_core* c, int32_t world, int32_t rank, const void* const* h_peer_grads,
                                   void* const* h_peer_flags, uint32_t* epoch_local, void* const* h_peer_recv) {
  MDP_REQUIRE(c, "mdp_core_bind_peers: null core");
  if (c->d_peer_tables) cudaFree(c->d_peer_tables);
  c->d_peer_tables = nullptr;
  c->peer_world = 0;
  c->d_peer_recv = nullptr;
  if (world <= 1) return MDP_OK;
  MDP_REQUIRE(world <= PEER_MAXW && rank >= 0 && rank < world && h_peer_grads && h_peer_flags && epoch_local,
              "mdp_core_bind_peers: bad argument (world %d, rank %d)", world, rank);
  MDP_REQUIRE(h_peer_grads[rank] == (const void*)c->grads,
              "mdp_core_bind_peers: this rank's peer entry must be the gradient buffer bound with mdp_core_bind");
  void* tab = nullptr;
  MDP_CUDA(cudaMalloc(&tab, 3 * world * sizeof(void*)));
  if (h_peer_recv) {
    MDP_CUDA(cudaMemcpy((char*)tab + 2 * world * sizeof(void*), h_peer_recv, world * sizeof(void*), cudaMemcpyHostToDevice));
    c->d_peer_recv = reinterpret_cast<uint2* const*>((char*)tab + 2 * world * sizeof(void*));
  }
  MDP_CUDA(cudaMemcpy(tab, h_peer_grads, world * sizeof(void*), cudaMemcpyHostToDevice));
  MDP_CUDA(cudaMemcpy((char*)tab + world * sizeof(void*), h_peer_flags, world * sizeof(void*), cudaMemcpyHostToDevice));
  c->d_peer_tables = tab;
  c->d_peer_grads = reinterpret_cast<const float* const*>(tab);
  c->d_peer_flags = reinterpret_cast<unsigned* const*>((char*)tab + world * sizeof(void*));
  c->peer_epoch = epoch_local;
  c->peer_world = world;
  c->peer_rank = rank;
  return MDP_OK;
}

// W1 of agents [a0, a0 + count) through the many-CTA path when it is wide.  With peers bound this applies to the barrier protocol
// only (the low-latency protocol keeps small buckets in the per-variable kernel); the rank-summed gradient goes through the
// local scratch c->peer_reduced.  Returns 1 if it handled W1 (the caller then starts at variable 1), 0 if not.
static int wide_w1_step(mdp_core* c, int which, int a0, int count, float grad_scale, int do_polyak, cudaStream_t st, int* handled) {
  *handled = 0;
  const bool peers = c->peer_world > 1;
  if (peers && c->d_peer_recv) return MDP_OK;
  const int U = c->cfg.num_units, net = which == 0 ? MDP_NET_P : MDP_NET_Q;
  long long min_len = 1ll << 60, max_len = 0;
  for (int j = a0; j < a0 + count; ++j) {
    const long long len = (long long)c->lay.net_in[j][net] * U;
    min_len = std::min(min_len, len); max_len = std::max(max_len, len);
  }
  if (min_len < 32768) return MDP_OK;
  if (!c->norm2) MDP_CUDA(cudaMalloc(&c->norm2, MDP_MAX_AGENTS * W1_MAXB * sizeof(float)));
  if (peers && !c->peer_reduced) MDP_CUDA(cudaMalloc(&c->peer_reduced, (size_t)c->lay.total_train * sizeof(float)));
  // with peers every CTA spins on the flag barrier: all CTAs of a launch must be co-resident (<= 148 x 8 of 256 threads)
  const int nb = (int)std::min<long long>(peers ? std::min(W1_MAXB, std::max(1, 592 / count)) : W1_MAXB, (max_len / 4 + 1023) / 1024);
  const PeerCtx P = peer_ctx(c);
  k_w1_sqnorm<<<dim3(nb, count), 256, 0, st>>>(c->d_agents, which, U, a0, grad_scale, c->norm2, P, c->grads, c->peer_reduced);
  int rc = check_launch("k_w1_sqnorm");
  if (rc) return rc;
  k_w1_adam<<<dim3(nb, count), 256, 0, st>>>(c->d_agents, which, U, a0, c->grads, c->adam_m, c->adam_v, c->adam_t, c->norm2, grad_scale,
                                             (float)c->cfg.grad_clip, c->cfg.lr, c->cfg.beta1, c->cfg.beta2, (float)c->cfg.adam_eps,
                                             (float)c->cfg.polyak, do_polyak, P, c->peer_reduced);
  rc = check_launch("k_w1_adam");
  if (rc) return rc;
  if (peers) {
    k_w1_epoch_bump<<<cdiv(count, 32), 32, 0, st>>>(c->peer_epoch, which, a0, count);
    rc = check_launch("k_w1_epoch_bump");
    if (rc) return rc;
  }
  *handled = 1;
  return MDP_OK;
}

// pdl: the caller guarantees that the preceding launch on `stream` is the gradient kernel of the same net(s) (mdp_update_agent /
// mdp_update_all), which writes none of the Adam slots / parameters this kernel prefetches before its dependency wait
int mdp::clip_adam_polyak_all_impl(mdp_core* c, int32_t which, float grad_scale, int32_t do_polyak, void* stream, bool pdl) {
  MDP_REQUIRE(c && c->d_agents, "mdp_clip_adam_polyak_all: core not bound");
  MDP_REQUIRE(which == 0 || which == 1, "mdp_clip_adam_polyak_all: bad argument");
  int wide = 0;
  int rc = wide_w1_step(c, which, 0, c->cfg.n_agents, grad_scale, do_polyak, (cudaStream_t)stream, &wide);
  if (rc) return rc;
  cudaLaunchConfig_t lc;
  cudaLaunchAttribute at[1];
  pdl_config(&lc, at, dim3(6 - wide, c->cfg.n_agents), 1024, (cudaStream_t)stream, pdl && wide == 0);
  MDP_CUDA(cudaLaunchKernelEx(&lc, k_clip_adam_polyak_all, (const AgentDev*)c->d_agents, (int)which, (int)c->cfg.num_units, c->grads,
                              c->adam_m, c->adam_v, (const int*)c->adam_t, grad_scale, (float)c->cfg.grad_clip, c->cfg.lr, c->cfg.beta1,
                              c->cfg.beta2, (float)c->cfg.adam_eps, (float)c->cfg.polyak, (int)do_polyak, peer_ctx(c), 0, wide));
  return check_launch("k_clip_adam_polyak_all");
}
extern "C" int mdp_clip_adam_polyak_all(mdp_core* c, int32_t which, float grad_scale, int32_t do_polyak, void* stream) {
  return mdp::clip_adam_polyak_all_impl(c, which, grad_scale, do_polyak, stream, false);
}

int mdp::clip_adam_polyak_impl(mdp_core* c, int32_t agent, int32_t which, float grad_scale, int32_t do_polyak, void* stream,
                               bool pdl) {
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
  int wide = 0;
  int rc = wide_w1_step(c, which, agent, 1, grad_scale, do_polyak, (cudaStream_t)stream, &wide);
  if (rc) return rc;
  cudaLaunchConfig_t lc;
  cudaLaunchAttribute at[1];
  pdl_config(&lc, at, dim3(6 - wide), 1024, (cudaStream_t)stream, pdl && wide == 0);
  MDP_CUDA(cudaLaunchKernelEx(&lc, k_clip_adam_polyak, param, target, grad, m, v, seg, (const int*)(c->adam_t + 2 * agent + which),
                              grad_scale, (float)c->cfg.grad_clip, c->cfg.lr, c->cfg.beta1, c->cfg.beta2, (float)c->cfg.adam_eps,
                              (float)c->cfg.polyak, (int)do_polyak, peer_ctx(c), (long long)c->lay.train_off[agent][which],
                              12 * agent + 6 * which, wide));
  return check_launch("k_clip_adam_polyak");
}
extern "C" int mdp_clip_adam_polyak(mdp_core* c, int32_t agent, int32_t which, float grad_scale, int32_t do_polyak,
                                    void* stream) {
  return mdp::clip_adam_polyak_impl(c, agent, which, grad_scale, do_polyak, stream, false);
}
