// Tensor-core (tcgen05 / TMEM) forward path of the MADDPG update: the fused TD-target kernel
//   a'_i = gumbel_softmax(target_p_i(o'_i)) for all i ; q' = target_q_j(o', a') ; y = r_j + gamma (1 - d_j) q'
// (maddpg/trainer/maddpg.py:181-187, :70-71, :104,108) with every MLP layer as a UMMA GEMM.
//
// One CTA owns 128 batch rows (UMMA_M = 128, cta_group::1, N = num_units).  Operands are staged in shared memory
// as SWIZZLE_128B K-major images (mdp_umma.cuh) by the CTA's own threads, because they are produced on the fly:
// the layer-1 input is gathered row by row from the replay ring through the sampled indices (fused
// ReplayBuffer.sample_index), the hidden activations come out of the previous epilogue, and every fp32 operand
// is split into a TF32 "hi" image and an exact remainder "lo" image.  Each GEMM is then issued as THREE
// kind::tf32 MMAs (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM), which restores ~fp32 products -- the
// 1e-4 parity bar on Q values does not survive plain TF32.  Layer 1 streams K in 32-column chunks
// through a two-stage ring (the MMAs of chunk c overlap the gather + split of chunk c+1; tcgen05.commit frees a
// stage); accumulators live in TMEM and are read back with tcgen05.ld (one thread = one batch row x 32 units),
// so bias + ReLU, the output head, the Gumbel-softmax and the TD combine are row-local epilogues in registers.
#include "mdp_mlp.cuh"
#include "mdp_umma.cuh"

#include <algorithm>
#include <vector>

namespace mdp {
namespace tc {

constexpr int TMR = 128;  // batch rows per CTA (UMMA M)
constexpr int NTC = 256;  // compute threads: 8 warps; warp w reads TMEM lanes [32 (w & 3), +32), unit half w >> 2
constexpr int NTT = 320;  // + a TMA producer warp (pre-split weight images) + an MMA issuer warp

template <int U>
struct Lay {  // byte offsets from the 1024-byte aligned base of dynamic shared memory
  static constexpr uint32_t X_IMG = TMR * 128;           // [128 rows][32 cols]   16 KB
  static constexpr uint32_t W_IMG = U * 128;             // [U rows][32 cols]
  static constexpr uint32_t STAGE = 2 * X_IMG + 2 * W_IMG;
  static constexpr uint32_t H_IMG = (U / 32) * X_IMG;    // [128][U]
  static constexpr uint32_t W2_IMG = (U / 32) * W_IMG;   // [U][U]
  static constexpr uint32_t OFF_H = 2 * STAGE;
  static constexpr uint32_t OFF_W2 = OFF_H + 2 * H_IMG;
  static constexpr uint32_t OFF_MISC = OFF_W2 + 2 * W2_IMG;
  static constexpr int MISC_FLOATS = U + U + U * MAXK + 16 + TMR * KPAD + TMR + 2 * TMR + 2 * TMR;  // b1 b2 W3 b3 part q rd rowoff
};

__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, uint32_t parity) {
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
    if ((spins & 1023u) == 0) {  // a lost arrival must fail loudly (~2 s), never hang the GPU
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void named_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Pre-split weight images in global memory (built by k_build_images from the flat parameter buffer right before
// every tensor-core launch): per net, W1^T as nchunks x [hi | lo] chunk images of [U][32] and W2^T as a [hi | lo]
// pair of [U][U] images, already in the SWIZZLE_128B byte order, so one cp.async.bulk drops them into place.
struct NetImg {
  const unsigned char* w1;
  const unsigned char* w2;
};
struct AgentImg {
  NetImg net[4];
};

template <int U>
__global__ void __launch_bounds__(256) k_build_images(CoreDev C, const AgentImg* __restrict__ imgs, int n_actor_jobs, int a_begin,
                                                      int q_begin, int actor_net, int critic_net) {
  using L = Lay<U>;
  const bool is_actor = (int)blockIdx.y < n_actor_jobs;
  const int agent = is_actor ? a_begin + blockIdx.y : q_begin + (blockIdx.y - n_actor_jobs);
  const int net = is_actor ? actor_net : critic_net;
  const MlpW w = C.agents[agent].net[net];
  unsigned char* w1 = const_cast<unsigned char*>(imgs[agent].net[net].w1);
  unsigned char* w2 = const_cast<unsigned char*>(imgs[agent].net[net].w2);
  const int nchunks = (w.in + 31) / 32;
  const long long n1 = (long long)nchunks * 32 * U, n2 = (long long)U * U;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n1 + n2; t += (long long)gridDim.x * blockDim.x) {
    float hi, lo;
    if (t < n1) {
      const int u = (int)(t % U), col = (int)(t / U);
      umma::split_tf32(col < w.in ? w.W1[(size_t)col * U + u] : 0.f, hi, lo);
      unsigned char* dst = w1 + (size_t)(col >> 5) * (2 * L::W_IMG) + umma::sw128_off(u, col & 31);
      *reinterpret_cast<float*>(dst) = hi;
      *reinterpret_cast<float*>(dst + L::W_IMG) = lo;
    } else {
      const int e = (int)(t - n1), u2 = e % U, k1 = e / U;
      umma::split_tf32(w.W2[e], hi, lo);
      unsigned char* dst = w2 + (size_t)(k1 >> 5) * L::W_IMG + umma::sw128_off(u2, k1 & 31);
      *reinterpret_cast<float*>(dst) = hi;
      *reinterpret_cast<float*>(dst + L::W2_IMG) = lo;
    }
  }
}

// layer-1 input of one net: global columns [0, n0) of the gathered row, with an override range served from the
// sampled-action tile (shared memory, or an L2-resident scratch when the joint action is too wide for smem)
struct XT {
  const float* g0;
  int n0;
  const float* over;
  int over_ld, over_c0, over_n;
  int vec_ok;  // g0 + row offset is 16-byte aligned: whole float4 units inside [0, n0) use one LDG.128
};

struct XRegs {
  float4 v[4];  // thread's four float4 units of a [128][32] chunk: unit e = j * 256 + tid -> row e >> 3, columns 4 (e & 7)..+3
};

__device__ __forceinline__ float x_get(const XT& xs, const long long* sRow, int r, int col) {
  if (col >= xs.over_c0 && col < xs.over_c0 + xs.over_n) return xs.over[(size_t)r * xs.over_ld + (col - xs.over_c0)];
  if (col < xs.n0) return __ldg(xs.g0 + sRow[r] + col);
  return 0.f;
}

__device__ __forceinline__ void load_x(XRegs& x, const XT& xs, const long long* __restrict__ sRow, int k0, int tid) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = j * NTC + tid, r = e >> 3, col = k0 + 4 * (e & 7);
    const bool in_over = col + 4 > xs.over_c0 && col < xs.over_c0 + xs.over_n;
    if (xs.vec_ok && col + 4 <= xs.n0 && !in_over) {
      x.v[j] = __ldg(reinterpret_cast<const float4*>(xs.g0 + sRow[r] + col));
    } else {
      x.v[j].x = x_get(xs, sRow, r, col);
      x.v[j].y = x_get(xs, sRow, r, col + 1);
      x.v[j].z = x_get(xs, sRow, r, col + 2);
      x.v[j].w = x_get(xs, sRow, r, col + 3);
    }
  }
}

template <int U>
__device__ __forceinline__ void store_x(unsigned char* stage, const XRegs& x, int tid) {
  unsigned char* xhi = stage;
  unsigned char* xlo = stage + Lay<U>::X_IMG;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = j * NTC + tid, r = e >> 3;
    float4 hi, lo;
    umma::split_tf32(x.v[j].x, hi.x, lo.x);
    umma::split_tf32(x.v[j].y, hi.y, lo.y);
    umma::split_tf32(x.v[j].z, hi.z, lo.z);
    umma::split_tf32(x.v[j].w, hi.w, lo.w);
    const uint32_t off = umma::sw128_off(r, 4 * (e & 7));
    *reinterpret_cast<float4*>(xhi + off) = hi;
    *reinterpret_cast<float4*>(xlo + off) = lo;
  }
}

// D[128 x U] (+)= A[128 x 8*nsteps] * B[U x 8*nsteps]^T as lo*hi + hi*lo + hi*hi (one elected thread)
template <int U>
__device__ __forceinline__ void issue_3x(uint32_t tacc, uint32_t a_hi, uint32_t a_lo, uint32_t a_panel, uint32_t b_hi, uint32_t b_lo,
                                         uint32_t b_panel, int nsteps, bool accumulate) {
  constexpr uint32_t idesc = umma::idesc_tf32(TMR, U, 0, 0);
  for (int s = 0; s < nsteps; ++s) {
    umma::mma_tf32(tacc, umma::desc_k(a_lo, a_panel, s), umma::desc_k(b_hi, b_panel, s), idesc, (accumulate || s > 0) ? 1u : 0u);
    umma::mma_tf32(tacc, umma::desc_k(a_hi, a_panel, s), umma::desc_k(b_lo, b_panel, s), idesc, 1u);
    umma::mma_tf32(tacc, umma::desc_k(a_hi, a_panel, s), umma::desc_k(b_hi, b_panel, s), idesc, 1u);
  }
}

struct Pipe {       // uniform across the CTA (the producer warp keeps its own copy in step)
  uint32_t chunks;  // layer-1 chunks so far (stage = chunks & 1)
  uint32_t accs;    // completed waits on the accumulator barrier
  uint32_t nets;    // nets finished
};

struct Bars {
  unsigned long long stage_w[2];     // TMA: W1^T chunk images landed in stage s
  unsigned long long stage_x[2];     // compute warps (8 arrivals): X chunk images stored in stage s
  unsigned long long h1_full;        // compute warps (8 arrivals): h1 images stored
  unsigned long long stage_free[2];  // tcgen05.commit: the MMAs reading stage s are done
  unsigned long long w2_full;        // TMA: W2^T images landed
  unsigned long long w2_free;        // tcgen05.commit: layer-2 MMAs done, W2^T may be overwritten
  unsigned long long acc;            // tcgen05.commit: accumulator complete
};

// producer warp (one lane): streams one net's weight images; runs ahead of the compute warps, throttled by the
// stage_free / w2_free barriers
template <int U>
__device__ __forceinline__ void produce_net(unsigned char* smem, Bars* bars, Pipe& pipe, const NetImg& img, int in_dim) {
  using L = Lay<U>;
  if (pipe.nets >= 1) mbar_wait_bounded(&bars->w2_free, (pipe.nets - 1u) & 1u);
  mbar_arrive_expect_tx(&bars->w2_full, 2 * L::W2_IMG);
  bulk_g2s(smem + L::OFF_W2, img.w2, 2 * L::W2_IMG, &bars->w2_full);
  const int nchunks = (in_dim + 31) / 32;
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t s = pipe.chunks & 1u;
    if (pipe.chunks >= 2) mbar_wait_bounded(&bars->stage_free[s], ((pipe.chunks >> 1) - 1u) & 1u);
    mbar_arrive_expect_tx(&bars->stage_w[s], 2 * L::W_IMG);
    bulk_g2s(smem + s * L::STAGE + 2 * L::X_IMG, img.w1 + (size_t)c * (2 * L::W_IMG), 2 * L::W_IMG, &bars->stage_w[s]);
    pipe.chunks++;
  }
  pipe.nets++;
}

__device__ __forceinline__ void warp_arrive(unsigned long long* bar, int lane) {
  umma::fence_async_smem();  // this lane's generic-proxy stores -> async proxy
  __syncwarp();
  if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MMA issuer warp (one lane): both GEMMs of one net.  Waits for the operands (TMA weight images + the compute warps'
// X / h1 images), issues the 3xTF32 MMAs and commits the barriers that free the buffers / publish the accumulators.
template <int U>
__device__ __forceinline__ void mma_net(unsigned char* smem, Bars* bars, Pipe& pipe, uint32_t tbase, int in_dim) {
  using L = Lay<U>;
  const int nchunks = (in_dim + 31) / 32;
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t s = pipe.chunks & 1u, ph = (pipe.chunks >> 1) & 1u;
    mbar_wait_bounded(&bars->stage_x[s], ph);
    mbar_wait_bounded(&bars->stage_w[s], ph);
    umma::fence_after();
    const uint32_t sa = smem_u32(smem + s * L::STAGE);
    issue_3x<U>(tbase, sa, sa + L::X_IMG, L::X_IMG, sa + 2 * L::X_IMG, sa + 2 * L::X_IMG + L::W_IMG, L::W_IMG, 4, c > 0);
    umma::commit(&bars->stage_free[s]);
    if (c == nchunks - 1) umma::commit(&bars->acc);
    pipe.chunks++;
  }
  mbar_wait_bounded(&bars->h1_full, pipe.nets & 1u);
  mbar_wait_bounded(&bars->w2_full, pipe.nets & 1u);
  umma::fence_after();
  const uint32_t ha = smem_u32(smem + L::OFF_H), wa = smem_u32(smem + L::OFF_W2);
  issue_3x<U>(tbase + U, ha, ha + L::H_IMG, L::X_IMG, wa, wa + L::W2_IMG, L::W_IMG, U / 8, false);
  umma::commit(&bars->acc);
  umma::commit(&bars->w2_free);
  pipe.nets++;
}

// h2 = relu(relu(X W1 + b1) W2 + b2) for the CTA's 128 rows (compute warps); thread (warp w, lane l) ends up with units
// [32 (w >> 2) + 64 g, +32) of row 32 (w & 3) + l in h2[32 g ..].  Warps run decoupled: each one gathers, splits and
// stores its 16 rows of a chunk and arrives on the stage barrier; nothing but the accumulator barrier joins them.
template <int U>
__device__ __forceinline__ void forward_hidden_tc(unsigned char* smem, Bars* bars, Pipe& pipe, uint32_t tbase, const XT& xs, const MlpW& w,
                                                  const long long* sRow, float* sB1, float* sB2, float* sW3, float* sB3,
                                                  float (&h2)[U / 2]) {
  using L = Lay<U>;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half = warp >> 2;
  unsigned char* h1hi = smem + L::OFF_H;
  const int nchunks = (w.in + 31) / 32;
  XRegs xr[2];  // register prefetch, two chunks deep
  load_x(xr[0], xs, sRow, 0, tid);
  if (nchunks > 1) load_x(xr[1], xs, sRow, 32, tid);
  for (int i = tid; i < U; i += NTC) { sB1[i] = w.b1[i]; sB2[i] = w.b2[i]; }
  for (int i = tid; i < U * w.out; i += NTC) sW3[i] = w.W3[i];
  if (tid < w.out) sB3[tid] = w.b3[tid];
  named_sync();  // small tensors visible to every compute warp
  // ---- layer 1: K chunks through the two-stage ring
  for (int c = 0; c < nchunks; c += 2) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      if (c + b < nchunks) {
        const uint32_t s = pipe.chunks & 1u;
        if (pipe.chunks >= 2) mbar_wait_bounded(&bars->stage_free[s], ((pipe.chunks >> 1) - 1u) & 1u);
        store_x<U>(smem + s * L::STAGE, xr[b], tid);
        warp_arrive(&bars->stage_x[s], lane);
        if (c + b + 2 < nchunks) load_x(xr[b], xs, sRow, 32 * (c + b + 2), tid);
        pipe.chunks++;
      }
    }
  }
  mbar_wait_bounded(&bars->acc, pipe.accs & 1u);
  pipe.accs++;
  umma::fence_after();
  // ---- epilogue 1: h1 = relu(acc + b1) -> split -> K-major image pair (A operand of layer 2)
#pragma unroll
  for (int g = 0; g < U / 64; ++g) {
    const int c0 = 32 * half + 64 * g;  // this thread's 32 units of the pass
    float v[32];
    umma::tmem_ld32(tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
    unsigned char* phi = h1hi + (uint32_t)(c0 >> 5) * L::X_IMG;
    unsigned char* plo = phi + L::H_IMG;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 hi, lo;
      umma::split_tf32(fmaxf(v[4 * q + 0] + sB1[c0 + 4 * q + 0], 0.f), hi.x, lo.x);
      umma::split_tf32(fmaxf(v[4 * q + 1] + sB1[c0 + 4 * q + 1], 0.f), hi.y, lo.y);
      umma::split_tf32(fmaxf(v[4 * q + 2] + sB1[c0 + 4 * q + 2], 0.f), hi.z, lo.z);
      umma::split_tf32(fmaxf(v[4 * q + 3] + sB1[c0 + 4 * q + 3], 0.f), hi.w, lo.w);
      const uint32_t off = umma::sw128_off(row, 4 * q);
      *reinterpret_cast<float4*>(phi + off) = hi;
      *reinterpret_cast<float4*>(plo + off) = lo;
    }
  }
  umma::fence_before();
  warp_arrive(&bars->h1_full, lane);
  pipe.nets++;
  // ---- layer 2 (issued by the MMA warp): acc2 = h1 W2
  mbar_wait_bounded(&bars->acc, pipe.accs & 1u);
  pipe.accs++;
  umma::fence_after();
  // ---- epilogue 2: h2 = relu(acc2 + b2), kept in registers
#pragma unroll
  for (int g = 0; g < U / 64; ++g) {
    const int c0 = 32 * half + 64 * g;
    float v[32];
    umma::tmem_ld32(tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(U + c0), v);
#pragma unroll
    for (int i = 0; i < 32; ++i) h2[32 * g + i] = fmaxf(v[i] + sB2[c0 + i], 0.f);
  }
  umma::fence_before();
}

// out[a] = sum over this thread's units of h2 * W3[:, a]   (a < KK); the two unit halves of a row are combined by the caller
template <int U, int KK>
__device__ __forceinline__ void head_partial(const float (&h2)[U / 2], const float* __restrict__ sW3, int half, float (&out)[MAXK]) {
#pragma unroll
  for (int a = 0; a < KK; ++a) out[a] = 0.f;
#pragma unroll
  for (int g = 0; g < U / 64; ++g)
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float* w3 = sW3 + (32 * half + 64 * g + i) * KK;
#pragma unroll
      for (int a = 0; a < KK; ++a) out[a] = fmaf(h2[32 * g + i], w3[a], out[a]);
    }
}

template <int U>
__global__ void __launch_bounds__(NTT, 1) k_td_target_tc(CoreDev C, const AgentImg* __restrict__ imgs, int j0, mdp_ring_layout L, int B,
                                                         const float* __restrict__ batch, const long long* __restrict__ ridx,
                                                         const float* __restrict__ u_target, int u_stride, uint64_t seed,
                                                         uint64_t counter, float* __restrict__ y_out, float* __restrict__ target_act_out,
                                                         long long idx_stride, long long y_stride, float* __restrict__ act_scratch,
                                                         int act_in_smem) {
  using LY = Lay<U>;
  const int j = j0 + blockIdx.y;  // grouped launch: one agent per grid.y slice
  if (ridx) ridx += blockIdx.y * idx_stride;
  y_out += blockIdx.y * y_stride;
  if (C.ctl) counter += C.ctl[0];
  extern __shared__ unsigned char smem_raw[];
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  unsigned char* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  float* misc = reinterpret_cast<float*>(smem + LY::OFF_MISC);
  float* sB1 = misc;
  float* sB2 = sB1 + U;
  float* sW3 = sB2 + U;
  float* sB3 = sW3 + U * MAXK;
  float* sPart = sB3 + 16;
  float* sQ = sPart + TMR * KPAD;
  float* sRD = sQ + TMR;
  long long* sRow = reinterpret_cast<long long*>(sRD + 2 * TMR);
  const int ASP = C.act_stride | 1;
  float* sAct = reinterpret_cast<float*>(sRow + TMR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half = warp >> 2;
  const AgentDev& me = C.agents[j];
  const long long row0 = (long long)blockIdx.x * TMR;
  const int nrows = (int)min((long long)TMR, B - row0);
  const int R = L.row_stride;
  // the sampled-action tile: shared memory when it fits, otherwise an L2-resident scratch (rows of this CTA only)
  float* actT = act_in_smem ? sAct : act_scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * TMR * ASP;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 2 * U);
  if (tid == 0) {
    mbar_init(&bars.stage_w[0], 1);
    mbar_init(&bars.stage_w[1], 1);
    mbar_init(&bars.stage_x[0], NTC / 32);
    mbar_init(&bars.stage_x[1], NTC / 32);
    mbar_init(&bars.h1_full, NTC / 32);
    mbar_init(&bars.stage_free[0], 1);
    mbar_init(&bars.stage_free[1], 1);
    mbar_init(&bars.w2_full, 1);
    mbar_init(&bars.w2_free, 1);
    mbar_init(&bars.acc, 1);
  }
  if (tid < TMR) {
    const long long rl = row0 + min(tid, nrows - 1);  // tail rows replay the last valid row (their results are dropped)
    const long long rg = ridx ? ridx[rl] : rl;
    sRow[tid] = rg * R;
    sRD[2 * tid] = batch[rg * R + L.rw_off + j];
    sRD[2 * tid + 1] = batch[rg * R + L.dn_off + j];
  }
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  Pipe pipe{0u, 0u, 0u};
  const int i_begin = me.local_q ? j : 0, i_end = me.local_q ? j + 1 : C.n_agents;

  if (warp == NTC / 32) {
    // ===== producer warp: weight images of every net of this CTA's job list, in order =====
    if (lane == 0) {
      for (int i = i_begin; i < i_end; ++i) produce_net<U>(smem, &bars, pipe, imgs[i].net[MDP_NET_TARGET_P], C.agents[i].obs_dim);
      produce_net<U>(smem, &bars, pipe, imgs[j].net[MDP_NET_TARGET_Q], me.net[MDP_NET_TARGET_Q].in);
    }
  } else if (warp == NTC / 32 + 1) {
    // ===== MMA issuer warp =====
    if (lane == 0) {
      for (int i = i_begin; i < i_end; ++i) mma_net<U>(smem, &bars, pipe, tbase, C.agents[i].obs_dim);
      mma_net<U>(smem, &bars, pipe, tbase, me.net[MDP_NET_TARGET_Q].in);
    }
  } else {
    // ===== compute warps =====
    float h2[U / 2];
    float part[MAXK];
    // a'_i = gumbel_softmax(target_p_i(next_obs_i)) for every agent the critic sees
    for (int i = i_begin; i < i_end; ++i) {
      const AgentDev& ag = C.agents[i];
      const MlpW w = ag.net[MDP_NET_TARGET_P];
      XT xs{batch + L.nx_off + ag.obs_off, ag.obs_dim, nullptr, 0, 0, 0, ((L.nx_off + ag.obs_off) & 3) == 0};
      forward_hidden_tc<U>(smem, &bars, pipe, tbase, xs, w, sRow, sB1, sB2, sW3, sB3, h2);
      const int K = ag.act_dim;
      if (K == 5) head_partial<U, 5>(h2, sW3, half, part);
      else if (K == 9) head_partial<U, 9>(h2, sW3, half, part);
      else {
        for (int a = 0; a < MAXK; ++a) part[a] = 0.f;
        for (int a = 0; a < K; ++a)
          for (int g = 0; g < U / 64; ++g)
            for (int t = 0; t < 32; ++t) part[a] = fmaf(h2[32 * g + t], sW3[(32 * half + 64 * g + t) * K + a], part[a]);
      }
      if (half == 1)
        for (int a = 0; a < K; ++a) sPart[row * KPAD + a] = part[a];
      named_sync();
      if (half == 0) {  // one thread per batch row: logits -> Gumbel-softmax per head (distributions.py:264-266, 332-336)
        float z[MAXK];
#pragma unroll
        for (int a = 0; a < MAXK; ++a) {
          if (a < K) {
            const float logit = part[a] + sPart[row * KPAD + a] + sB3[a];
            const float u = u_target ? u_target[(row0 + min(row, nrows - 1)) * u_stride + ag.act_off + a]
                                     : philox_u(seed, counter, (uint32_t)(0x100 + i), row0 + row, a);
            z[a] = logit + gumbel_from_u(u);
          }
        }
        for (int h = 0; h < ag.n_heads; ++h) {
          const int o = h ? ag.head_dim[0] : 0, n = ag.head_dim[h];
          float m = -INFINITY;
#pragma unroll
          for (int a = 0; a < MAXK; ++a)
            if (a >= o && a < o + n) m = fmaxf(m, z[a]);
          float ssum = 0.f;
#pragma unroll
          for (int a = 0; a < MAXK; ++a)
            if (a >= o && a < o + n) { z[a] = expf(z[a] - m); ssum += z[a]; }
#pragma unroll
          for (int a = 0; a < MAXK; ++a)
            if (a >= o && a < o + n) z[a] = z[a] / ssum;
        }
#pragma unroll
        for (int a = 0; a < MAXK; ++a)
          if (a < K) {
            actT[(size_t)row * ASP + ag.act_off + a] = z[a];
            if (target_act_out && row < nrows) target_act_out[(row0 + row) * u_stride + ag.act_off + a] = z[a];
          }
      }
      named_sync();  // a' visible to the critic's gather; sPart / small tensors free for the next net
    }

    // q' = target_q_j([next_obs | a'])
    XT xq;
    if (me.local_q) xq = XT{batch + L.nx_off + me.obs_off, me.obs_dim, actT + me.act_off, ASP, me.obs_dim, me.act_dim,
                            ((L.nx_off + me.obs_off) & 3) == 0};
    else xq = XT{batch + L.nx_off, C.obs_sum, actT, ASP, C.obs_sum, C.act_sum, (L.nx_off & 3) == 0};
    const MlpW tq = me.net[MDP_NET_TARGET_Q];
    forward_hidden_tc<U>(smem, &bars, pipe, tbase, xq, tq, sRow, sB1, sB2, sW3, sB3, h2);
    head_partial<U, 1>(h2, sW3, half, part);
    if (half == 1) sPart[row * KPAD] = part[0];
    named_sync();
    if (half == 0) {
      // y = float32(rew + gamma * (1 - done) * q')  -- float64 combine like numpy (maddpg.py:186)
      double sy = 0, syy = 0, sr = 0, sq = 0;
      if (row < nrows) {
        const float qn = part[0] + sPart[row * KPAD] + sB3[0];
        const double rew = (double)sRD[2 * row], done = (double)sRD[2 * row + 1];
        const double y = rew + C.gamma * (1.0 - done) * (double)qn;
        y_out[row0 + row] = (float)y;
        sy = y; syy = y * y; sr = rew; sq = (double)qn;
      }
      for (int o = 16; o > 0; o >>= 1) {
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        syy += __shfl_xor_sync(0xffffffffu, syy, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      if (lane == 0) {
        double* st = C.stats + 8 * j;
        atomicAdd(st + 3, sy); atomicAdd(st + 4, syy); atomicAdd(st + 5, sr); atomicAdd(st + 6, sq);
        atomicAdd(st + 7, (double)max(0, min(32, nrows - 32 * warp)));
      }
    }
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tbase, 2 * U);
}

}  // namespace tc

// host side -------------------------------------------------------------------------------------------------
// weight-image arena: [agent][net] -> {W1^T chunk images, W2^T images}, plus the device pointer table
static int ensure_images(mdp_core* c) {
  if (c->tc_imgs) return MDP_OK;
  constexpr int U = 64;
  using LY = tc::Lay<U>;
  const int n = c->cfg.n_agents;
  size_t total = 0;
  std::vector<tc::AgentImg> h(n);
  std::vector<size_t> off1((size_t)n * 4), off2((size_t)n * 4);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) {
      const size_t nchunks = (size_t)(c->lay.net_in[i][k] + 31) / 32;
      off1[i * 4 + k] = total; total += nchunks * 2 * LY::W_IMG;
      off2[i * 4 + k] = total; total += 2 * LY::W2_IMG;
    }
  unsigned char* arena = nullptr;
  MDP_CUDA(cudaMalloc(&arena, total + n * sizeof(tc::AgentImg)));
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) {
      h[i].net[k].w1 = arena + off1[i * 4 + k];
      h[i].net[k].w2 = arena + off2[i * 4 + k];
    }
  MDP_CUDA(cudaMemcpy(arena + total, h.data(), n * sizeof(tc::AgentImg), cudaMemcpyHostToDevice));
  c->tc_arena = arena;
  c->tc_imgs = arena + total;
  return MDP_OK;
}

// Returns MDP_ENOTSUP when the shape is outside the tensor-core path (the caller then uses the SIMT kernels).
int launch_td_target_tc(mdp_core* c, const CoreDev& d, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B,
                        const float* batch, const long long* ridx, long long idx_stride, const float* u_target, int32_t u_stride,
                        uint64_t seed, uint64_t counter, float* y_out, long long y_stride, float* target_act_out, cudaStream_t st) {
  if (c->cfg.num_units != 64) return fail(MDP_ENOTSUP, "tensor-core path: num_units %d (64 only)", c->cfg.num_units);
  constexpr int U = 64;
  using LY = tc::Lay<U>;
  int rc = ensure_images(c);
  if (rc) return rc;
  const tc::AgentImg* imgs = reinterpret_cast<const tc::AgentImg*>(c->tc_imgs);
  const int n = c->cfg.n_agents, ASP = c->act_stride | 1;
  const size_t fixed = LY::OFF_MISC + (size_t)LY::MISC_FLOATS * 4 + 1024 + 64;
  const size_t limit = 227 * 1024 - 256;  // static __shared__ (barriers, TMEM slot) comes on top
  const size_t act_bytes = (size_t)tc::TMR * ASP * 4;
  const int act_in_smem = fixed + act_bytes <= limit;
  const size_t smem = fixed + (act_in_smem ? act_bytes : 0);
  const int tiles = cdiv(B, tc::TMR);
  if (!act_in_smem) {
    const size_t need = (size_t)tiles * count * tc::TMR * ASP * sizeof(float);
    if (need > c->tc_scratch_bytes) {
      if (c->tc_scratch) cudaFree(c->tc_scratch);
      c->tc_scratch = nullptr;
      c->tc_scratch_bytes = 0;
      MDP_CUDA(cudaMalloc(&c->tc_scratch, need));
      c->tc_scratch_bytes = need;
    }
  }
  // weight images of the nets this launch reads: the target actors the critics see + the target critics of the slice
  bool any_global = false;
  for (int k = agent; k < agent + count; ++k) any_global |= !c->cfg.local_q[k];
  const int a_begin = any_global ? 0 : agent, a_count = any_global ? n : count;
  int max_in = 0;
  for (int k = 0; k < n; ++k) max_in = std::max(max_in, c->lay.net_in[k][MDP_NET_TARGET_Q]);
  const int bx = std::min(64, cdiv(((max_in + 31) / 32 * 32 + U) * U, 256 * 4));
  tc::k_build_images<U><<<dim3(bx, a_count + count), 256, 0, st>>>(d, imgs, a_count, a_begin, agent, MDP_NET_TARGET_P, MDP_NET_TARGET_Q);
  rc = check_launch("k_build_images");
  if (rc) return rc;
  auto kern = tc::k_td_target_tc<U>;
  MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(tiles, count), tc::NTT, smem, st>>>(d, imgs, agent, *lay, B, batch, ridx, u_target, u_stride, seed, counter, y_out,
                                                   target_act_out, idx_stride, y_stride, c->tc_scratch, act_in_smem);
  return check_launch("k_td_target_tc");
}

}  // namespace mdp
