// Shared-memory tile primitives of the fused MLP kernels (fp32 SIMT).  A group of NT = 256 threads owns a
// tile of TM (16 or 32) rows and carries it through whole layers; thread (ty, tx) = (tid >> 4, tid & 15)
// owns rows {RM*ty .. RM*ty+RM-1} (RM = TM/16) and columns {64g + 4tx .. +3 : g < U/64}.
// RES = the network's weights are resident in shared memory (small configs); otherwise weights stream
// from global/L2 through a [KC][U] staging chunk.
// Used by mdp_train.cu (update kernels) and mdp_rollout.cu (persistent episode kernel).
#pragma once
#include "mdp_core.cuh"

#ifndef MDP_KUNROLL
#define MDP_KUNROLL 2  // k-loop unroll of the SIMT tile GEMMs (shared-memory loads in flight per thread)
#endif

namespace mdp {

// A "group" is the NT = 256 threads that cooperate on one MLP tile.  The update kernels run one group per
// CTA (barrier 0 == __syncthreads); the episode kernel runs several groups per CTA (one agent each), each
// synchronising on its own named barrier.
struct Grp {
  int tid;     // thread index inside the group
  int bar_id;  // 0: whole CTA, else named barrier id
  int nthr = 256;  // threads synchronising on the named barrier
  __device__ __forceinline__ void sync() const {
    if (bar_id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthr) : "memory");
  }
};

constexpr int NT = 256;   // threads per group: 16 (row groups) x 16 (column quads)
constexpr int KUNROLL = MDP_KUNROLL;
constexpr int KC = 32;    // K-chunk streamed through shared memory
constexpr int XP = KC + 4;
constexpr int KPAD = 16;  // pitch of per-row action/logit scratch
constexpr int MAXK = 15;  // widest action block: MultiDiscrete([5, 10]) of simple_reference (the tcgen05 kernels carry 9: TC_MAXK)

// Layer-1 input: up to two global column segments plus an optional shared-memory override range
// (the freshly sampled action that replaces the replayed one).
struct XSrc {
  const float* g0; int ld0, n0;
  const float* g1; int ld1, n1;
  const float* s_over; int over_ld, over_c0, over_n;
  const long long* idx;  // optional row indirection: logical row b lives at physical row idx[b] (fused replay gather)
  __device__ __forceinline__ float get(int r_local, long long r_global, int c) const {
    if (c >= over_c0 && c < over_c0 + over_n) return s_over[r_local * over_ld + (c - over_c0)];
    if (idx) r_global = idx[r_global];
    if (c < n0) return g0[r_global * ld0 + c];
    c -= n0;
    if (c < n1) return g1[r_global * ld1 + c];
    return 0.f;
  }
};

__device__ __forceinline__ XSrc make_xsrc(const float* g0, int ld0, int n0) {
  XSrc x; x.g0 = g0; x.ld0 = ld0; x.n0 = n0; x.g1 = nullptr; x.ld1 = 0; x.n1 = 0;
  x.s_over = nullptr; x.over_ld = 0; x.over_c0 = 0; x.over_n = 0; x.idx = nullptr; return x;
}

template <int U, int TM>
__device__ __forceinline__ void zero_acc(float2 (&acc)[TM / 16][U / 32]) {
#pragma unroll
  for (int r = 0; r < TM / 16; ++r)
#pragma unroll
    for (int c = 0; c < U / 32; ++c) acc[r][c] = make_float2(0.f, 0.f);
}

// acc += sA[rows][0..kc) * sW[0..kc)[cols]; A read as float4 (lda % 4 == 0, kc % 4 == 0, 16-byte aligned)
template <int U, int TM>
__device__ __forceinline__ void mma_tile(const Grp& G, float2 (&acc)[TM / 16][U / 32], const float* __restrict__ sA, int lda,
                                         const float* __restrict__ sW, int kc) {
  constexpr int RM = TM / 16;
  const int ty = G.tid >> 4, tx = G.tid & 15;
  const float* ap = sA + (RM * ty) * lda;
#pragma unroll KUNROLL
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
          const float2 a2 = make_float2(av[rr][kk], av[rr][kk]);  // packed FFMA2: two fp32 FMAs per instruction
          acc[rr][2 * g + 0] = __ffma2_rn(a2, make_float2(w.x, w.y), acc[rr][2 * g + 0]);
          acc[rr][2 * g + 1] = __ffma2_rn(a2, make_float2(w.z, w.w), acc[rr][2 * g + 1]);
        }
      }
    }
  }
}

// same with scalar A loads (no alignment or multiple-of-4 demands)
template <int U, int TM>
__device__ __forceinline__ void mma_tile_sa(const Grp& G, float2 (&acc)[TM / 16][U / 32], const float* __restrict__ sA, int lda,
                                            const float* __restrict__ sW, int kc) {
  constexpr int RM = TM / 16;
  const int ty = G.tid >> 4, tx = G.tid & 15;
  const float* ap = sA + (RM * ty) * lda;
#pragma unroll KUNROLL
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

// rows [k0, k0+KC) of a row-major (K, U) weight -> sW[KC][U], zero past K
template <int U>
__device__ __forceinline__ void load_w_rows(const Grp& G, float* __restrict__ sW, const float* __restrict__ W, int k0, int K) {
  for (int idx = G.tid * 4; idx < KC * U; idx += NT * 4) {
    const int k = idx / U;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + k < K) v = *reinterpret_cast<const float4*>(W + (size_t)(k0 + k) * U + (idx - k * U));
    *reinterpret_cast<float4*>(sW + idx) = v;
  }
}

// transposed rows of a (U, U) weight: sW[ul][k] = W[k][u0 + ul], ul < nu  (for dX = dY * W^T)
template <int U>
__device__ __forceinline__ void load_wT_rows(const Grp& G, float* __restrict__ sW, const float* __restrict__ W, int u0, int nu) {
  for (int idx = G.tid; idx < nu * U; idx += NT) {
    const int ul = idx / U, k = idx - ul * U;
    sW[idx] = W[(size_t)k * U + u0 + ul];
  }
}

// contiguous block of n4*4 floats global -> shared (16-byte aligned both sides)
__device__ __forceinline__ void load_block(const Grp& G, float* __restrict__ dst, const float* __restrict__ src, int nfloats) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int q = G.tid; q < (nfloats + 3) / 4; q += NT) d4[q] = s4[q];
}

__host__ __device__ inline int net_floats_padded(int in, int U, int out) {
  return ((in * U + U + U * U + U + U * out + out) + 3) & ~3;
}

// resident copy of one net: returns an MlpW whose pointers are in shared memory (call G.sync() afterwards)
template <int U>
__device__ __forceinline__ MlpW load_net(const Grp& G, float* __restrict__ dst, const MlpW& w) {
  load_block(G, dst, w.W1, net_floats_padded(w.in, U, w.out));
  MlpW s = w;
  const float* b = dst;
  s.W1 = b; b += w.in * U;
  s.b1 = b; b += U;
  s.W2 = b; b += U * U;
  s.b2 = b; b += U;
  s.W3 = b; b += U * w.out;
  s.b3 = b;
  return s;
}

template <int TM>
__device__ __forceinline__ void load_x_chunk(const Grp& G, float* __restrict__ sX, const XSrc& xs, long long row0, int nrows, int k0) {
  for (int idx = G.tid; idx < TM * KC; idx += NT) {
    const int r = idx >> 5, c = idx & 31;
    sX[r * XP + c] = (r < nrows) ? xs.get(r, row0 + r, k0 + c) : 0.f;
  }
}

// acc = X[rows] * W1, X streamed in K-chunks through sX; W1 from smem (RES) or staged through sW (ends synchronised)
template <int U, int TM, bool RES>
__device__ __forceinline__ void layer1(const Grp& G, float2 (&acc)[TM / 16][U / 32], const XSrc& xs, int K, const float* __restrict__ W1,
                                       long long row0, int nrows, float* sX, float* sW) {
  zero_acc<U, TM>(acc);
  for (int k0 = 0; k0 < K; k0 += KC) {
    load_x_chunk<TM>(G, sX, xs, row0, nrows, k0);
    if (!RES) load_w_rows<U>(G, sW, W1, k0, K);
    G.sync();
    if (RES) mma_tile_sa<U, TM>(G, acc, sX, XP, W1 + (size_t)k0 * U, min(KC, K - k0));
    else mma_tile<U, TM>(G, acc, sX, XP, sW, KC);
    G.sync();
  }
}

// acc = sA[TM][U] * W (U,U); W from smem (RES, no barrier) or streamed in KC-row chunks (ends synchronised)
template <int U, int TM, bool RES>
__device__ __forceinline__ void layer_h(const Grp& G, float2 (&acc)[TM / 16][U / 32], const float* __restrict__ sA, const float* __restrict__ W,
                                        float* sW) {
  constexpr int HP = U + 4;
  zero_acc<U, TM>(acc);
  if (RES) {
    mma_tile<U, TM>(G, acc, sA, HP, W, U);
  } else {
    for (int k0 = 0; k0 < U; k0 += KC) {
      load_w_rows<U>(G, sW, W, k0, U);
      G.sync();
      mma_tile<U, TM>(G, acc, sA + k0, HP, sW, KC);
      G.sync();
    }
  }
}

// acc = sA[TM][U] * W^T; WT = resident transposed copy (RES) else W streamed transposed (ends synchronised)
template <int U, int TM, bool RES>
__device__ __forceinline__ void layer_hT(const Grp& G, float2 (&acc)[TM / 16][U / 32], const float* __restrict__ sA, const float* __restrict__ W,
                                         const float* __restrict__ WT, float* sW) {
  constexpr int HP = U + 4;
  zero_acc<U, TM>(acc);
  if (RES) {
    mma_tile<U, TM>(G, acc, sA, HP, WT, U);
    G.sync();
  } else {
    for (int k0 = 0; k0 < U; k0 += KC) {
      load_wT_rows<U>(G, sW, W, k0, KC);
      G.sync();
      mma_tile<U, TM>(G, acc, sA + k0, HP, sW, KC);
      G.sync();
    }
  }
}

// sH[r][c] = relu(acc + bias[c])   (caller synchronises)
template <int U, int TM>
__device__ __forceinline__ void store_bias_relu(const Grp& G, const float2 (&acc)[TM / 16][U / 32], const float* __restrict__ bias, float* sH) {
  constexpr int HP = U + 4, RM = TM / 16;
  const int ty = G.tid >> 4, tx = G.tid & 15;
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

// sH[r][c] = (sH[r][c] > 0) ? acc : 0     in place: dz = dh * relu'(h)   (caller synchronises)
template <int U, int TM>
__device__ __forceinline__ void store_masked(const Grp& G, const float2 (&acc)[TM / 16][U / 32], float* sH) {
  constexpr int HP = U + 4, RM = TM / 16;
  const int ty = G.tid >> 4, tx = G.tid & 15;
#pragma unroll
  for (int g = 0; g < U / 64; ++g) {
    const int c = g * 64 + 4 * tx;
#pragma unroll
    for (int rr = 0; rr < RM; ++rr) {
      float4 h = *reinterpret_cast<const float4*>(sH + (RM * ty + rr) * HP + c);
      h.x = h.x > 0.f ? acc[rr][2 * g + 0].x : 0.f;
      h.y = h.y > 0.f ? acc[rr][2 * g + 0].y : 0.f;
      h.z = h.z > 0.f ? acc[rr][2 * g + 1].x : 0.f;
      h.w = h.w > 0.f ? acc[rr][2 * g + 1].y : 0.f;
      *reinterpret_cast<float4*>(sH + (RM * ty + rr) * HP + c) = h;
    }
  }
}

// h1 -> sH1, h2 -> sH2 for the tile (ends synchronised)
template <int U, int TM, bool RES>
__device__ __forceinline__ void forward_hidden(const Grp& G, const XSrc& xs, const MlpW& w, long long row0, int nrows, float* sX,
                                               float* sW, float* sH1, float* sH2) {
  float2 acc[TM / 16][U / 32];
  layer1<U, TM, RES>(G, acc, xs, w.in, w.W1, row0, nrows, sX, sW);
  store_bias_relu<U, TM>(G, acc, w.b1, sH1);
  G.sync();
  layer_h<U, TM, RES>(G, acc, sH1, w.W2, sW);
  store_bias_relu<U, TM>(G, acc, w.b2, sH2);
  G.sync();
}

// out_dim == 1 head: sQ[r] = h2[r,:] . W3 + b3   (256/TM threads per row; ends synchronised)
template <int U, int TM>
__device__ __forceinline__ void critic_head(const Grp& G, const float* __restrict__ sH2, const MlpW& w, float* sQ) {
  constexpr int HP = U + 4, PARTS = NT / TM;
  const int row = G.tid / PARTS, part = G.tid % PARTS;
  float s = 0.f;
  for (int u = part; u < U; u += PARTS) s = fmaf(sH2[row * HP + u], w.W3[u], s);
#pragma unroll
  for (int o = PARTS / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (part == 0) sQ[row] = s + w.b3[0];
  G.sync();
}

// general head: sL[r][a] = h2[r,:] . W3[:,a] + b3[a], a < out.  8 threads per row, each sums U/8 hidden
// units for all outputs, then a 3-step shuffle reduction (ends synchronised).  KK = compile-time out width.
template <int U, int TM, int KK>
__device__ __forceinline__ void actor_head_k(const Grp& G, const float* __restrict__ sH2, const MlpW& w, float* sL) {
  constexpr int HP = U + 4;
  for (int r = G.tid >> 3; r < TM; r += NT / 8) {  // one pass for TM <= 32
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

template <int U, int TM>
__device__ __forceinline__ void actor_head(const Grp& G, const float* __restrict__ sH2, const MlpW& w, float* sL) {
  if (w.out == 5) return actor_head_k<U, TM, 5>(G, sH2, w, sL);   // Discrete(5): every MPE movement head
  if (w.out == 9) return actor_head_k<U, TM, 9>(G, sH2, w, sL);   // MultiDiscrete([5, 4]): simple_world_comm leader
  constexpr int HP = U + 4;
  const int K = w.out;
  for (int idx = G.tid; idx < TM * K; idx += NT) {
    const int r = idx / K, a = idx - r * K;
    float s = 0.f;
    for (int u = 0; u < U; ++u) s = fmaf(sH2[r * HP + u], w.W3[u * K + a], s);
    sL[r * KPAD + a] = s + w.b3[a];
  }
  G.sync();
}

// U[0,1) draw of element (row, col) for Philox stream (seed, counter, tag)
__device__ __forceinline__ float philox_u(uint64_t seed, uint64_t counter, uint32_t tag, long long row, int col) {
  const uint4 r = Philox::gen(seed, (uint32_t)row, (uint32_t)(row >> 32) ^ (tag << 8) ^ (uint32_t)col, (uint32_t)counter,
                              (uint32_t)(counter >> 32));
  return Philox::u01(r.x);
}

// Gumbel-softmax of a TM x K logits tile: softmax(logits - log(-log u)) per head.  Pass 1 (one thread
// per element) writes the perturbed logits to sOut, pass 2 (one thread per (row, head)) normalises them
// in place.  sL (the clean logits) is preserved.  Ends synchronised.
template <int TM>
__device__ __forceinline__ void gumbel_softmax_tile(const Grp& G, const float* __restrict__ sL, float* __restrict__ sOut, int out_ld,
                                                    int nrows, int K, int n_heads, const int* head_dim,
                                                    const float* __restrict__ u_glob, int u_ld, int u_col0,
                                                    long long row0, uint64_t seed, uint64_t counter, uint32_t tag) {
  for (int idx = G.tid; idx < TM * K; idx += NT) {
    const int r = idx / K, a = idx - r * K;
    if (r >= nrows) continue;
    const float u = u_glob ? u_glob[(row0 + r) * u_ld + u_col0 + a] : philox_u(seed, counter, tag, row0 + r, a);
    sOut[r * out_ld + a] = sL[r * KPAD + a] + gumbel_from_u(u);
  }
  G.sync();
  for (int idx = G.tid; idx < TM * n_heads; idx += NT) {
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

// gW (U,U) += sA^T (h, TM x U) * sD (dz, TM x U); thread owns rows ty*CN.. and the usual columns
template <int U, int TM>
__device__ __forceinline__ void grad_w_hidden(const Grp& G, const float* __restrict__ sA, const float* __restrict__ sD, float* __restrict__ gW) {
  constexpr int HP = U + 4, RN = U / 16;
  const int ty = G.tid >> 4, tx = G.tid & 15;
  float acc[RN][RN];
#pragma unroll
  for (int a = 0; a < RN; ++a)
#pragma unroll
    for (int b = 0; b < RN; ++b) acc[a][b] = 0.f;
  for (int r = 0; r < TM; ++r) {
    float av[RN], dv[RN];
#pragma unroll
    for (int q = 0; q < RN / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(sA + r * HP + ty * RN + 4 * q);
      av[4 * q + 0] = t.x; av[4 * q + 1] = t.y; av[4 * q + 2] = t.z; av[4 * q + 3] = t.w;
      const float4 d = *reinterpret_cast<const float4*>(sD + r * HP + q * 64 + 4 * tx);
      dv[4 * q + 0] = d.x; dv[4 * q + 1] = d.y; dv[4 * q + 2] = d.z; dv[4 * q + 3] = d.w;
    }
#pragma unroll
    for (int a = 0; a < RN; ++a)
#pragma unroll
      for (int b = 0; b < RN; ++b) acc[a][b] = fmaf(av[a], dv[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < RN; ++a)
#pragma unroll
    for (int b = 0; b < RN; b += 4)
      red_add4(gW + (size_t)(ty * RN + a) * U + (b / 4) * 64 + 4 * tx, acc[a][b], acc[a][b + 1], acc[a][b + 2], acc[a][b + 3]);
}

// gW1 rows [k0, k0+KC) += sX^T (TM x KC chunk) * sD (TM x U)
template <int U, int TM>
__device__ __forceinline__ void grad_w_chunk(const Grp& G, const float* __restrict__ sX, int ldx, const float* __restrict__ sD,
                                             float* __restrict__ gW1, int k0, int K) {
  constexpr int HP = U + 4, RN = U / 16;
  const int ty = G.tid >> 4, tx = G.tid & 15;
  float acc[2][RN];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < RN; ++b) acc[a][b] = 0.f;
  for (int r = 0; r < TM; ++r) {
    float2 a;  // scalar loads: the tile base may sit at an odd column offset
    a.x = sX[r * ldx + 2 * ty];
    a.y = sX[r * ldx + 2 * ty + 1];
#pragma unroll
    for (int q = 0; q < RN / 4; ++q) {
      const float4 d = *reinterpret_cast<const float4*>(sD + r * HP + q * 64 + 4 * tx);
      acc[0][4 * q + 0] = fmaf(a.x, d.x, acc[0][4 * q + 0]);
      acc[0][4 * q + 1] = fmaf(a.x, d.y, acc[0][4 * q + 1]);
      acc[0][4 * q + 2] = fmaf(a.x, d.z, acc[0][4 * q + 2]);
      acc[0][4 * q + 3] = fmaf(a.x, d.w, acc[0][4 * q + 3]);
      acc[1][4 * q + 0] = fmaf(a.y, d.x, acc[1][4 * q + 0]);
      acc[1][4 * q + 1] = fmaf(a.y, d.y, acc[1][4 * q + 1]);
      acc[1][4 * q + 2] = fmaf(a.y, d.z, acc[1][4 * q + 2]);
      acc[1][4 * q + 3] = fmaf(a.y, d.w, acc[1][4 * q + 3]);
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int k = k0 + 2 * ty + rr;
    if (k < K) {
#pragma unroll
      for (int b = 0; b < RN; b += 4)
        red_add4(gW1 + (size_t)k * U + (b / 4) * 64 + 4 * tx, acc[rr][b], acc[rr][b + 1], acc[rr][b + 2], acc[rr][b + 3]);
    }
  }
}

template <int U, int TM>
__device__ __forceinline__ void grad_bias(const Grp& G, const float* __restrict__ sD, float* __restrict__ gb) {
  constexpr int HP = U + 4;
  if (G.tid < U) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s += sD[r * HP + G.tid];
    red_add(gb + G.tid, s);
  }
}

// full backward below the second hidden layer: given dz2 in sH2 (already masked) and h1 in sH1,
// accumulates gW2, gb2, then dz1 -> sH1 (in place), then optionally gW1/gb1 (re-streaming X).
// w2T: resident transposed W2 (RES) or unused.
template <int U, int TM, bool RES>
__device__ __forceinline__ void backward_hidden(const Grp& G, const XSrc& xs, const MlpW& w, const float* w2T, const MlpG* g,
                                                long long row0, int nrows, float* sX, float* sW, float* sH1, float* sH2) {
  if (g) {
    grad_w_hidden<U, TM>(G, sH1, sH2, g->W2);
    grad_bias<U, TM>(G, sH2, g->b2);
  }
  float2 acc[TM / 16][U / 32];
  G.sync();
  layer_hT<U, TM, RES>(G, acc, sH2, w.W2, w2T, sW);  // dh1 = dz2 * W2^T
  store_masked<U, TM>(G, acc, sH1);                   // dz1 = dh1 * relu'(h1)
  G.sync();
  if (g) {
    grad_bias<U, TM>(G, sH1, g->b1);
    for (int k0 = 0; k0 < w.in; k0 += KC) {
      load_x_chunk<TM>(G, sX, xs, row0, nrows, k0);
      G.sync();
      grad_w_chunk<U, TM>(G, sX, XP, sH1, g->W1, k0, w.in);
      G.sync();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tile-resident helpers (small configs): whole X row tile, nets and a swizzled W2^T live in smem
// ---------------------------------------------------------------------------------------------
// sXf[r][c] = base[(idx ? idx[row0+r] : row0+r) * ld + col0 + c], c < ncols (zero rows past nrows)
template <int TM>
__device__ __forceinline__ void load_rows(const Grp& G, float* __restrict__ sXf, int ldx, const float* __restrict__ base, long long ld,
                                          const long long* __restrict__ idx, long long row0, int nrows, int col0, int ncols) {
  for (int i = G.tid; i < TM * ncols; i += NT) {
    const int r = i / ncols, c = i - r * ncols;
    float v = 0.f;
    if (r < nrows) {
      const long long rg = idx ? idx[row0 + r] : row0 + r;
      v = base[rg * ld + col0 + c];
    }
    sXf[r * ldx + c] = v;
  }
}

// MlpW view of a net stored contiguously at `base` (shared memory copy made by a bulk load)
template <int U>
__device__ __forceinline__ MlpW net_at(const float* base, int in, int out) {
  MlpW s;
  const float* b = base;
  s.W1 = b; b += in * U;
  s.b1 = b; b += U;
  s.W2 = b; b += U * U;
  s.b2 = b; b += U;
  s.W3 = b; b += U * out;
  s.b3 = b;
  s.in = in; s.out = out;
  return s;
}

// Gathers the tile's rows with one TMA bulk copy per row: sXf[r][0..ncols4) = base[(idx ? idx[row0+r] : row0+r) * ld
// + col0 ..], issued by threads 0..TM-1 (completion on `bar`); rows past nrows are zero-filled with plain stores.
// ncols4 and col0 are multiples of 4 floats.  Returns nothing; the caller accounts nrows * ncols4 * 4 bytes.
template <int TM>
__device__ __forceinline__ void bulk_rows(const Grp& G, float* __restrict__ sXf, int ldx, const float* __restrict__ base, long long ld,
                                          const long long* __restrict__ idx, long long row0, int nrows, int col0, int ncols4,
                                          unsigned long long* bar) {
  if (G.tid < nrows) {
    const long long rg = idx ? idx[row0 + G.tid] : row0 + G.tid;
    bulk_g2s(sXf + G.tid * ldx, base + rg * ld + col0, (uint32_t)ncols4 * 4u, bar);
  }
  for (int i = G.tid; i < (TM - nrows) * ncols4; i += NT) sXf[(nrows + i / ncols4) * ldx + i % ncols4] = 0.f;
}

// sWT[u][k] = W2[k][u] with the 4-float column groups XOR-swizzled by the row so that both the transposing
// writes (consecutive threads = consecutive u) and the float4 reads of mma_tile_swz are conflict-free
template <int U>
__device__ __forceinline__ void build_wT_swz(const Grp& G, float* __restrict__ sWT, const float* __restrict__ sW2) {
  constexpr int GM = U / 4 - 1;
  for (int i = G.tid; i < U * U; i += NT) {
    const int k = i / U, u = i - k * U;
    sWT[u * U + ((((k >> 2) ^ (u & GM)) << 2) | (k & 3))] = sW2[i];
  }
}

// acc += sA[rows][0..U) * W^T where sWT is the swizzled transposed copy built by build_wT_swz
template <int U, int TM>
__device__ __forceinline__ void mma_tile_swz(const Grp& G, float2 (&acc)[TM / 16][U / 32], const float* __restrict__ sA, int lda,
                                             const float* __restrict__ sWT) {
  constexpr int RM = TM / 16, GM = U / 4 - 1;
  const int ty = G.tid >> 4, tx = G.tid & 15;
  const float* ap = sA + (RM * ty) * lda;
#pragma unroll KUNROLL
  for (int k = 0; k < U; k += 4) {
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
        const float4 w = *reinterpret_cast<const float4*>(sWT + (k + kk) * U + ((((g * 16 + tx) ^ ((k + kk) & GM))) << 2));
#pragma unroll
        for (int rr = 0; rr < RM; ++rr) {
          const float2 a2 = make_float2(av[rr][kk], av[rr][kk]);  // packed FFMA2: two fp32 FMAs per instruction
          acc[rr][2 * g + 0] = __ffma2_rn(a2, make_float2(w.x, w.y), acc[rr][2 * g + 0]);
          acc[rr][2 * g + 1] = __ffma2_rn(a2, make_float2(w.z, w.w), acc[rr][2 * g + 1]);
        }
      }
    }
  }
}

// forward through the two hidden layers with everything resident: X tile sXf (pitch ldx, K columns), net w
// in smem.  h1 -> sH1, h2 -> sH2 (ends synchronised)
template <int U, int TM>
__device__ __forceinline__ void forward_hidden_res(const Grp& G, const float* __restrict__ sXf, int ldx, const MlpW& w,
                                                   float* sH1, float* sH2) {
  float2 acc[TM / 16][U / 32];
  zero_acc<U, TM>(acc);
  mma_tile_sa<U, TM>(G, acc, sXf, ldx, w.W1, w.in);
  store_bias_relu<U, TM>(G, acc, w.b1, sH1);
  G.sync();
  zero_acc<U, TM>(acc);
  mma_tile<U, TM>(G, acc, sH1, U + 4, w.W2, U);
  store_bias_relu<U, TM>(G, acc, w.b2, sH2);
  G.sync();
}

// backward below the second hidden layer, everything resident: dz2 in sH2, h1 in sH1, X tile in sXf.
template <int U, int TM>
__device__ __forceinline__ void backward_hidden_res(const Grp& G, const float* __restrict__ sXf, int ldx, const MlpW& w,
                                                    const float* __restrict__ sWT, const MlpG* g, float* sH1, float* sH2) {
  if (g) {
    grad_w_hidden<U, TM>(G, sH1, sH2, g->W2);
    grad_bias<U, TM>(G, sH2, g->b2);
  }
  float2 acc[TM / 16][U / 32];
  zero_acc<U, TM>(acc);
  mma_tile_swz<U, TM>(G, acc, sH2, U + 4, sWT);  // dh1 = dz2 * W2^T
  G.sync();                                       // every reader of h1 (grad_w_hidden) is done
  store_masked<U, TM>(G, acc, sH1);               // dz1 = dh1 * relu'(h1)
  G.sync();
  if (g) {
    grad_bias<U, TM>(G, sH1, g->b1);
    for (int k0 = 0; k0 < w.in; k0 += KC) grad_w_chunk<U, TM>(G, sXf + k0, ldx, sH1, g->W1, k0, w.in);
  }
}

struct SmemCarve {
  float* p;
  __device__ explicit SmemCarve(void* base) : p(reinterpret_cast<float*>(base)) {}
  __device__ float* take(int nfloats) {
    float* r = p;
    p += (nfloats + 3) & ~3;
    return r;
  }
};

}  // namespace mdp
