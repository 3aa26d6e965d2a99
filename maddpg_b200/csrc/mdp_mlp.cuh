// Shared-memory tile primitives of the fused MLP kernels (fp32 SIMT): a CTA of NT = 256 threads owns a
// tile of TM = 32 rows and carries it through whole layers; thread (ty, tx) = (tid >> 4, tid & 15) owns
// rows {2ty, 2ty+1} and columns {64g + 4tx .. +3 : g < U/64}.  Used by mdp_train.cu (update kernels) and
// mdp_rollout.cu (persistent episode kernel).
#pragma once
#include "mdp_core.cuh"

namespace mdp {

// A "group" is the NT = 256 threads that cooperate on one MLP tile.  The update kernels run one group per
// CTA (barrier 0 == __syncthreads); the episode kernel runs several groups per CTA (one agent each), each
// synchronising on its own named barrier.
struct Grp {
  int tid;     // thread index inside the group
  int bar_id;  // 0: whole CTA, else named barrier id
  __device__ __forceinline__ void sync() const {
    if (bar_id == 0) __syncthreads();
    else asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
  }
};

constexpr int TM = 32;    // batch rows per CTA
constexpr int NT = 256;   // threads per CTA: 16 (row pairs) x 16 (column quads)
constexpr int KC = 32;    // K-chunk streamed through shared memory
constexpr int XP = KC + 4;
constexpr int KPAD = 12;  // pitch of per-row action/logit scratch (max act_dim 9)
constexpr int MAXK = 9;

// Layer-1 input: up to two global column segments plus an optional shared-memory override range
// (the freshly sampled action that replaces the replayed one).
struct XSrc {
  const float* g0; int ld0, n0;
  const float* g1; int ld1, n1;
  const float* s_over; int over_ld, over_c0, over_n;
  __device__ __forceinline__ float get(int r_local, long long r_global, int c) const {
    if (c >= over_c0 && c < over_c0 + over_n) return s_over[r_local * over_ld + (c - over_c0)];
    if (c < n0) return g0[r_global * ld0 + c];
    c -= n0;
    if (c < n1) return g1[r_global * ld1 + c];
    return 0.f;
  }
};

__device__ __forceinline__ XSrc make_xsrc(const float* g0, int ld0, int n0) {
  XSrc x; x.g0 = g0; x.ld0 = ld0; x.n0 = n0; x.g1 = nullptr; x.ld1 = 0; x.n1 = 0;
  x.s_over = nullptr; x.over_ld = 0; x.over_c0 = 0; x.over_n = 0; return x;
}

// ---------------------------------------------------------------------------------------------
// tile primitives.  Thread (ty, tx) = (tid >> 4, tid & 15) owns rows {2ty, 2ty+1} and columns
// {64g + 4tx .. +3 : g < U/64} of a TM x U tile.
// ---------------------------------------------------------------------------------------------
template <int U>
__device__ __forceinline__ void zero_acc(float (&acc)[2][U / 16]) {
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < U / 16; ++c) acc[r][c] = 0.f;
}

template <int U>
__device__ __forceinline__ void mma_tile(const Grp& G, float (&acc)[2][U / 16], const float* __restrict__ sA, int lda,
                                         const float* __restrict__ sW, int kc) {
  const int ty = G.tid >> 4, tx = G.tid & 15;
  const float* a0p = sA + (2 * ty) * lda;
  const float* a1p = a0p + lda;
#pragma unroll 2
  for (int k = 0; k < kc; k += 4) {
    const float4 a0 = *reinterpret_cast<const float4*>(a0p + k);
    const float4 a1 = *reinterpret_cast<const float4*>(a1p + k);
    const float a0v[4] = {a0.x, a0.y, a0.z, a0.w};
    const float a1v[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int g = 0; g < U / 64; ++g) {
        const float4 w = *reinterpret_cast<const float4*>(sW + (k + kk) * U + g * 64 + 4 * tx);
        acc[0][4 * g + 0] = fmaf(a0v[kk], w.x, acc[0][4 * g + 0]);
        acc[0][4 * g + 1] = fmaf(a0v[kk], w.y, acc[0][4 * g + 1]);
        acc[0][4 * g + 2] = fmaf(a0v[kk], w.z, acc[0][4 * g + 2]);
        acc[0][4 * g + 3] = fmaf(a0v[kk], w.w, acc[0][4 * g + 3]);
        acc[1][4 * g + 0] = fmaf(a1v[kk], w.x, acc[1][4 * g + 0]);
        acc[1][4 * g + 1] = fmaf(a1v[kk], w.y, acc[1][4 * g + 1]);
        acc[1][4 * g + 2] = fmaf(a1v[kk], w.z, acc[1][4 * g + 2]);
        acc[1][4 * g + 3] = fmaf(a1v[kk], w.w, acc[1][4 * g + 3]);
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

// transposed chunk of a (U, U) weight: sW[ul][k] = W[k][u0 + ul]  (for dX = dY * W^T)
template <int U>
__device__ __forceinline__ void load_wT_rows(const Grp& G, float* __restrict__ sW, const float* __restrict__ W, int u0) {
  for (int idx = G.tid; idx < KC * U; idx += NT) {
    const int ul = idx / U, k = idx - ul * U;
    sW[idx] = W[(size_t)k * U + u0 + ul];
  }
}

__device__ __forceinline__ void load_x_chunk(const Grp& G, float* __restrict__ sX, const XSrc& xs, long long row0, int nrows, int k0) {
  for (int idx = G.tid; idx < TM * KC; idx += NT) {
    const int r = idx >> 5, c = idx & 31;
    sX[r * XP + c] = (r < nrows) ? xs.get(r, row0 + r, k0 + c) : 0.f;
  }
}

// acc = X[rows] * W1 streamed in K-chunks (ends synchronised)
template <int U>
__device__ __forceinline__ void layer1(const Grp& G, float (&acc)[2][U / 16], const XSrc& xs, int K, const float* __restrict__ W1,
                                       long long row0, int nrows, float* sX, float* sW) {
  zero_acc<U>(acc);
  for (int k0 = 0; k0 < K; k0 += KC) {
    load_x_chunk(G, sX, xs, row0, nrows, k0);
    load_w_rows<U>(G, sW, W1, k0, K);
    G.sync();
    mma_tile<U>(G, acc, sX, XP, sW, KC);
    G.sync();
  }
}

// acc = sA[TM][U] * W (U,U)   (TRANSPOSED: * W^T), W streamed in KC-row chunks (ends synchronised)
template <int U, bool TRANSPOSED>
__device__ __forceinline__ void layer_h(const Grp& G, float (&acc)[2][U / 16], const float* __restrict__ sA, const float* __restrict__ W,
                                        float* sW) {
  constexpr int HP = U + 4;
  zero_acc<U>(acc);
  for (int k0 = 0; k0 < U; k0 += KC) {
    if (TRANSPOSED) load_wT_rows<U>(G, sW, W, k0); else load_w_rows<U>(G, sW, W, k0, U);
    G.sync();
    mma_tile<U>(G, acc, sA + k0, HP, sW, KC);
    G.sync();
  }
}

// sH[r][c] = relu(acc + bias[c])   (caller synchronises)
template <int U>
__device__ __forceinline__ void store_bias_relu(const Grp& G, const float (&acc)[2][U / 16], const float* __restrict__ bias, float* sH) {
  constexpr int HP = U + 4;
  const int ty = G.tid >> 4, tx = G.tid & 15;
#pragma unroll
  for (int g = 0; g < U / 64; ++g) {
    const int c = g * 64 + 4 * tx;
    const float4 b = *reinterpret_cast<const float4*>(bias + c);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float4 v;
      v.x = fmaxf(acc[rr][4 * g + 0] + b.x, 0.f);
      v.y = fmaxf(acc[rr][4 * g + 1] + b.y, 0.f);
      v.z = fmaxf(acc[rr][4 * g + 2] + b.z, 0.f);
      v.w = fmaxf(acc[rr][4 * g + 3] + b.w, 0.f);
      *reinterpret_cast<float4*>(sH + (2 * ty + rr) * HP + c) = v;
    }
  }
}

// sH[r][c] = (sH[r][c] > 0) ? acc : 0     in place: dz = dh * relu'(h)   (caller synchronises)
template <int U>
__device__ __forceinline__ void store_masked(const Grp& G, const float (&acc)[2][U / 16], float* sH) {
  constexpr int HP = U + 4;
  const int ty = G.tid >> 4, tx = G.tid & 15;
#pragma unroll
  for (int g = 0; g < U / 64; ++g) {
    const int c = g * 64 + 4 * tx;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float4 h = *reinterpret_cast<const float4*>(sH + (2 * ty + rr) * HP + c);
      h.x = h.x > 0.f ? acc[rr][4 * g + 0] : 0.f;
      h.y = h.y > 0.f ? acc[rr][4 * g + 1] : 0.f;
      h.z = h.z > 0.f ? acc[rr][4 * g + 2] : 0.f;
      h.w = h.w > 0.f ? acc[rr][4 * g + 3] : 0.f;
      *reinterpret_cast<float4*>(sH + (2 * ty + rr) * HP + c) = h;
    }
  }
}

// h1 -> sH1, h2 -> sH2 for the tile (ends synchronised)
template <int U>
__device__ __forceinline__ void forward_hidden(const Grp& G, const XSrc& xs, const MlpW& w, long long row0, int nrows, float* sX,
                                               float* sW, float* sH1, float* sH2) {
  float acc[2][U / 16];
  layer1<U>(G, acc, xs, w.in, w.W1, row0, nrows, sX, sW);
  store_bias_relu<U>(G, acc, w.b1, sH1);
  G.sync();
  layer_h<U, false>(G, acc, sH1, w.W2, sW);
  store_bias_relu<U>(G, acc, w.b2, sH2);
  G.sync();
}

// out_dim == 1 head: sQ[r] = h2[r,:] . W3 + b3   (8 threads per row; ends synchronised)
template <int U>
__device__ __forceinline__ void critic_head(const Grp& G, const float* __restrict__ sH2, const MlpW& w, float* sQ) {
  constexpr int HP = U + 4;
  const int row = G.tid >> 3, part = G.tid & 7;
  float s = 0.f;
  for (int u = part; u < U; u += 8) s = fmaf(sH2[row * HP + u], w.W3[u], s);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (part == 0) sQ[row] = s + w.b3[0];
  G.sync();
}

// general head: sL[r][a] = h2[r,:] . W3[:,a] + b3[a], a < out   (ends synchronised)
template <int U>
__device__ __forceinline__ void actor_head(const Grp& G, const float* __restrict__ sH2, const MlpW& w, float* sL) {
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

// gW (U,U) += sA^T (h, TM x U) * sD (dz, TM x U); thread owns rows ty*RN.. and the usual columns
template <int U>
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
    for (int b = 0; b < RN; ++b)
      atomicAdd(gW + (size_t)(ty * RN + a) * U + (b / 4) * 64 + 4 * tx + (b & 3), acc[a][b]);
}

// gW1 rows [k0, k0+KC) += sX^T (TM x KC chunk) * sD (TM x U)
template <int U>
__device__ __forceinline__ void grad_w_chunk(const Grp& G, const float* __restrict__ sX, const float* __restrict__ sD,
                                             float* __restrict__ gW1, int k0, int K) {
  constexpr int HP = U + 4, RN = U / 16;
  const int ty = G.tid >> 4, tx = G.tid & 15;
  float acc[2][RN];
  zero_acc<U>(acc);
  for (int r = 0; r < TM; ++r) {
    const float2 a = *reinterpret_cast<const float2*>(sX + r * XP + 2 * ty);
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
      for (int b = 0; b < RN; ++b) atomicAdd(gW1 + (size_t)k * U + (b / 4) * 64 + 4 * tx + (b & 3), acc[rr][b]);
    }
  }
}

template <int U>
__device__ __forceinline__ void grad_bias(const Grp& G, const float* __restrict__ sD, float* __restrict__ gb) {
  constexpr int HP = U + 4;
  if (G.tid < U) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s += sD[r * HP + G.tid];
    atomicAdd(gb + G.tid, s);
  }
}

// full backward below the second hidden layer: given dz2 in sH2 (already masked) and h1 in sH1,
// accumulates gW2, gb2, then dz1 -> sH1 (in place), then optionally gW1/gb1 (re-streaming X).
template <int U>
__device__ __forceinline__ void backward_hidden(const Grp& G, const XSrc& xs, const MlpW& w, const MlpG* g, long long row0, int nrows,
                                                float* sX, float* sW, float* sH1, float* sH2) {
  if (g) {
    grad_w_hidden<U>(G, sH1, sH2, g->W2);
    grad_bias<U>(G, sH2, g->b2);
  }
  float acc[2][U / 16];
  G.sync();
  layer_h<U, true>(G, acc, sH2, w.W2, sW);  // dh1 = dz2 * W2^T
  store_masked<U>(G, acc, sH1);              // dz1 = dh1 * relu'(h1)
  G.sync();
  if (g) {
    grad_bias<U>(G, sH1, g->b1);
    for (int k0 = 0; k0 < w.in; k0 += KC) {
      load_x_chunk(G, sX, xs, row0, nrows, k0);
      G.sync();
      grad_w_chunk<U>(G, sX, sH1, g->W1, k0, w.in);
      G.sync();
    }
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

template <int U>
constexpr int smem_floats_base() { return KC * U + TM * XP + 2 * TM * (U + 4); }


}  // namespace mdp
