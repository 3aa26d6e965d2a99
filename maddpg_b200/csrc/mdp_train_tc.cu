// Tensor-core (tcgen05 / TMEM) forward path of the MADDPG update: the fused TD-target kernel
//   a'_i = gumbel_softmax(target_p_i(o'_i)) for all i ; q' = target_q_j(o', a') ; y = r_j + gamma (1 - d_j) q'
// (maddpg/trainer/maddpg.py:181-187, :70-71, :104,108) with every MLP layer as a UMMA GEMM.
//
// One CTA owns 128 batch rows (UMMA_M = 128, cta_group::1, N = num_units).  Every fp32 operand is split into a TF32
// "hi" part and an exact remainder "lo" and each GEMM is issued as THREE kind::tf32 MMAs (lo*hi + hi*lo + hi*hi, fp32
// accumulation in TMEM), which restores ~fp32 products -- the 1e-4 parity bar on Q values does not survive plain TF32
// (SURVEY H4).  Operand placement:
//   A (activations: gathered replay rows, h1)  -> TENSOR MEMORY.  One thread owns one batch row (TMEM lane): it gathers its
//       row's 16 columns of a 32-column chunk from the replay ring through the sampled index (fused
//       ReplayBuffer.sample_index), splits them in registers and writes them with tcgen05.st; the epilogue of layer 1 puts
//       relu(acc + b1) straight back as layer 2's A operand the same way.  No shared-memory images, no swizzle, no proxy fence.
//   B (weights) -> shared memory, SWIZZLE_128B K-major images streamed by a TMA producer warp (cp.async.bulk) from
//       pre-split global images that k_build_images refreshes from the flat parameter buffer before every launch.
//   D (accumulators) -> TMEM, read back with tcgen05.ld (thread = batch row x 32 units), so bias + ReLU, the output head, the
//       Gumbel-softmax and the float64 TD combine are row-local register code.
// Warp roles: 8 compute warps (gather/split/tcgen05.st, epilogues), 1 TMA producer warp, 1 MMA issuer warp; mbarriers carry
// "stage filled" (compute -> MMA), "weights landed" (TMA -> MMA), "stage free" (tcgen05.commit -> compute, TMA) and
// "accumulator ready" (tcgen05.commit -> compute).  Layer 1 streams K through a 4-stage ring of TMEM A slots + smem B slots.
#include "mdp_mlp.cuh"
#include "mdp_umma.cuh"

#include <algorithm>
#include <vector>

// The tensor-core epilogues keep per-row action vectors in REGISTER arrays: they carry at most 9 action columns per agent
// (Discrete(5), MultiDiscrete([5, 4])); wider heads (simple_reference: [5, 10]) run the SIMT kernels (MDP_ENOTSUP here).
constexpr int TC_KPAD = 12;
constexpr int TC_MAXK = 9;
static inline bool tc_heads_ok(const mdp_core* c) {
  for (int i = 0; i < c->cfg.n_agents; ++i)
    if (c->cfg.act_dim[i] > TC_MAXK) return false;
  return true;
}

namespace mdp {
namespace tc {

constexpr int TMR = 128;  // batch rows per CTA (UMMA M)
constexpr int NTC = 256;  // compute threads: 8 warps; warp w reads TMEM lanes [32 (w & 3), +32), unit half w >> 2
constexpr int NTT = 320;  // + a TMA producer warp (pre-split weight images) + an MMA issuer warp

constexpr int NS = 4;      // layer-1 ring depth (TMEM A slots and shared-memory B slots)

template <int U>
struct Lay {
  // shared memory (bytes from the 1024-byte aligned base): weight images only
  static constexpr uint32_t W_IMG = U * 128;             // [U rows][32 cols] image
  static constexpr uint32_t W2_IMG = (U / 32) * W_IMG;   // [U][U] image
  static constexpr uint32_t OFF_W2 = NS * 2 * W_IMG;
  static constexpr uint32_t OFF_MISC = OFF_W2 + 2 * W2_IMG;
  static constexpr int MISC_FLOATS = U + U + U * TC_MAXK + 16 + 2 * TMR * TC_KPAD + TMR + 2 * TMR + 2 * TMR;  // b1 b2 W3 b3 part noise q rd rowoff
  // tensor memory columns (512 allocated): accumulators, h1 (hi | lo), NS x-chunk slots (hi | lo, 32 columns each)
  static constexpr uint32_t T_ACC1 = 0, T_ACC2 = U, T_H1 = 2 * U, T_X = 4 * U;
  static constexpr uint32_t T_COLS = 512;
  static_assert(4 * U + NS * 64 <= 512, "TMEM budget");
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
  const unsigned char* w1;   // W1^T chunk images: nchunks x [hi | lo] of [U][32]
  const unsigned char* w2;   // W2^T image pair [u2][k1]   (B operand of h1 W2)
  const unsigned char* w2n;  // W2 image pair [k1][u2]     (B operand of dz2 W2^T, the backward pass)
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
  unsigned char* w2n = const_cast<unsigned char*>(imgs[agent].net[net].w2n);
  const int nchunks = (w.in + 31) / 32;
  // one thread = four consecutive K columns of one image row: 16-byte image stores (hi and lo), coalesced weight reads
  const long long n1 = (long long)nchunks * 8 * U, n2 = (long long)(U / 4) * U;
  auto put = [](unsigned char* dst, uint32_t lo_off, float a, float b, float c, float d) {
    float4 hi, lo;
    umma::split_tf32(a, hi.x, lo.x); umma::split_tf32(b, hi.y, lo.y); umma::split_tf32(c, hi.z, lo.z); umma::split_tf32(d, hi.w, lo.w);
    *reinterpret_cast<float4*>(dst) = hi;
    *reinterpret_cast<float4*>(dst + lo_off) = lo;
  };
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n1 + 2 * n2; t += (long long)gridDim.x * blockDim.x) {
    if (t < n1) {  // W1^T chunk images: row u, columns col0..col0+3
      const int u = (int)(t % U), col0 = 4 * (int)(t / U);
      float x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = col0 + q < w.in ? w.W1[(size_t)(col0 + q) * U + u] : 0.f;
      put(w1 + (size_t)(col0 >> 5) * (2 * L::W_IMG) + umma::sw128_off(u, col0 & 31), L::W_IMG, x[0], x[1], x[2], x[3]);
    } else if (t < n1 + n2) {  // W2^T images: row u2, columns k1..k1+3
      const int e = (int)(t - n1), u2 = e % U, k1 = 4 * (e / U);
      put(w2 + (size_t)(k1 >> 5) * L::W_IMG + umma::sw128_off(u2, k1 & 31), L::W2_IMG, w.W2[(size_t)k1 * U + u2],
          w.W2[(size_t)(k1 + 1) * U + u2], w.W2[(size_t)(k1 + 2) * U + u2], w.W2[(size_t)(k1 + 3) * U + u2]);
    } else {  // W2 images: row k1, columns u2..u2+3
      const int e = (int)(t - n1 - n2), u2 = 4 * (e % (U / 4)), k1 = e / (U / 4);
      const float4 x = *reinterpret_cast<const float4*>(w.W2 + (size_t)k1 * U + u2);
      put(w2n + (size_t)(u2 >> 5) * L::W_IMG + umma::sw128_off(k1, u2 & 31), L::W2_IMG, x.x, x.y, x.z, x.w);
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
  float v[16];  // this thread's 16 columns (half of a 32-column chunk) of its batch row
};

__device__ __forceinline__ float x_get(const XT& xs, long long rowoff, int r, int col) {
  if (col >= xs.over_c0 && col < xs.over_c0 + xs.over_n) return xs.over[(size_t)r * xs.over_ld + (col - xs.over_c0)];
  if (col < xs.n0) return __ldg(xs.g0 + rowoff + col);
  return 0.f;
}

// columns [k0, k0 + 16) of batch row r (row offset rowoff in the ring)
__device__ __forceinline__ void load_x(XRegs& x, const XT& xs, long long rowoff, int r, int k0) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = k0 + 4 * q;
    const bool in_over = col + 4 > xs.over_c0 && col < xs.over_c0 + xs.over_n;
    if (xs.vec_ok && col + 4 <= xs.n0 && !in_over) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(xs.g0 + rowoff + col));
      x.v[4 * q + 0] = t.x; x.v[4 * q + 1] = t.y; x.v[4 * q + 2] = t.z; x.v[4 * q + 3] = t.w;
    } else {
      x.v[4 * q + 0] = x_get(xs, rowoff, r, col);
      x.v[4 * q + 1] = x_get(xs, rowoff, r, col + 1);
      x.v[4 * q + 2] = x_get(xs, rowoff, r, col + 2);
      x.v[4 * q + 3] = x_get(xs, rowoff, r, col + 3);
    }
  }
}

// split and write the thread's 16 columns into TMEM slot `slot` (hi at +0, lo at +32), columns [c16, c16 + 16)
__device__ __forceinline__ void store_x(uint32_t slot, uint32_t lane_base, int c16, const XRegs& x) {
  float hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) umma::split_tf32(x.v[i], hi[i], lo[i]);
  umma::tmem_st16(slot + lane_base + c16, hi);
  umma::tmem_st16(slot + lane_base + 32 + c16, lo);
}

// D[128 x U] (+)= A[128 x 8*NSTEPS] (TMEM, hi | lo `a_lo_off` columns apart) * B[U x 8*NSTEPS]^T (smem image pair, descriptors
// of k-step 0 given) as lo*hi + hi*lo + hi*hi.  The issuing thread is the serial bottleneck of the whole pipeline, so the
// loop is fully unrolled and a k-step only costs adds: +8 TMEM columns on A, +32 bytes (2 descriptor units) on B inside a
// panel, +panel bytes across panels.
template <int U, int NSTEPS>
__device__ __forceinline__ void issue_3x(uint32_t tacc, uint32_t a_hi, uint32_t a_lo_off, uint64_t b_hi, uint64_t b_lo, uint32_t accumulate) {
  constexpr uint32_t idesc = umma::idesc_tf32(TMR, U, 0, 0);
  constexpr uint32_t PANEL16 = (U * 128) >> 4;  // panel stride of a [U][32] image in 16-byte descriptor units
#pragma unroll
  for (int s = 0; s < NSTEPS; ++s) {
    const uint32_t bo = (uint32_t)(s >> 2) * PANEL16 + (uint32_t)(s & 3) * 2u;
    umma::mma_tf32_ta(tacc, a_hi + a_lo_off + 8 * s, b_hi + bo, idesc, s == 0 ? accumulate : 1u);
    umma::mma_tf32_ta(tacc, a_hi + 8 * s, b_lo + bo, idesc, 1u);
    umma::mma_tf32_ta(tacc, a_hi + 8 * s, b_hi + bo, idesc, 1u);
  }
}

struct Pipe {       // every role keeps its own copy, advanced in lock step
  uint32_t chunks;  // layer-1 chunks so far (slot = chunks % NS)
  uint32_t accs;    // completed waits on the accumulator barrier
  uint32_t nets;    // nets finished
};

struct Bars {
  unsigned long long stage_w[NS];     // TMA: W1^T chunk images landed in smem slot s
  unsigned long long stage_x[NS];     // compute warps (8 arrivals): X chunk written to TMEM slot s
  unsigned long long stage_free[NS];  // tcgen05.commit: the MMAs reading slot s (TMEM A + smem B) are done
  unsigned long long h1_full;         // compute warps (8 arrivals): h1 written to TMEM
  unsigned long long w2_full;         // TMA: W2^T images landed
  unsigned long long w2_free;         // tcgen05.commit: layer-2 MMAs done, W2^T may be overwritten
  unsigned long long acc;             // tcgen05.commit: accumulator complete
};

__device__ __forceinline__ void wait_slot_free(Bars* bars, const Pipe& pipe) {
  if (pipe.chunks >= NS) mbar_wait_bounded(&bars->stage_free[pipe.chunks % NS], ((pipe.chunks / NS) - 1u) & 1u);
}

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
    const uint32_t s = pipe.chunks % NS;
    wait_slot_free(bars, pipe);
    mbar_arrive_expect_tx(&bars->stage_w[s], 2 * L::W_IMG);
    bulk_g2s(smem + s * (2 * L::W_IMG), img.w1 + (size_t)c * (2 * L::W_IMG), 2 * L::W_IMG, &bars->stage_w[s]);
    pipe.chunks++;
  }
  pipe.nets++;
}

// this warp's tcgen05.st writes are complete and ordered before the MMA warp's reads
__device__ __forceinline__ void warp_arrive_tmem(unsigned long long* bar, int lane) {
  umma::tmem_st_wait();
  umma::fence_before();
  __syncwarp();
  if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MMA issuer warp (one lane): both GEMMs of one net.  Waits for the operands (TMA weight images + the compute warps'
// X / h1 writes), issues the 3xTF32 MMAs and commits the barriers that free the slots / publish the accumulators.
template <int U>
__device__ __forceinline__ void mma_net(unsigned char* smem, Bars* bars, Pipe& pipe, uint32_t tbase, int in_dim) {
  using L = Lay<U>;
  const int nchunks = (in_dim + 31) / 32;
  const uint64_t w_hi0 = umma::desc_k(smem_u32(smem), L::W_IMG, 0), w_lo0 = umma::desc_k(smem_u32(smem) + L::W_IMG, L::W_IMG, 0);
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t s = pipe.chunks % NS, ph = (pipe.chunks / NS) & 1u;
    mbar_wait_bounded(&bars->stage_x[s], ph);
    mbar_wait_bounded(&bars->stage_w[s], ph);
    umma::fence_after();
    const uint32_t so = s * ((2 * L::W_IMG) >> 4);  // smem slot offset in descriptor units
    if (umma::elect_one()) {
      issue_3x<U, 4>(tbase + L::T_ACC1, tbase + L::T_X + s * 64, 32, w_hi0 + so, w_lo0 + so, c > 0 ? 1u : 0u);
      umma::commit(&bars->stage_free[s]);
      if (c == nchunks - 1) umma::commit(&bars->acc);
    }
    __syncwarp();
    pipe.chunks++;
  }
  mbar_wait_bounded(&bars->h1_full, pipe.nets & 1u);
  mbar_wait_bounded(&bars->w2_full, pipe.nets & 1u);
  umma::fence_after();
  const uint32_t wa = smem_u32(smem + L::OFF_W2);
  if (umma::elect_one()) {
    issue_3x<U, U / 8>(tbase + L::T_ACC2, tbase + L::T_H1, U, umma::desc_k(wa, L::W_IMG, 0), umma::desc_k(wa + L::W2_IMG, L::W_IMG, 0), 0u);
    umma::commit(&bars->acc);
    umma::commit(&bars->w2_free);
  }
  __syncwarp();
  pipe.nets++;
}


// Layer-1 staging loop shared by the three kernels: the gathered X chunk goes to TMEM slot (chunk % NS) as a hi | lo pair.
// Warps 0-3 stage the even chunks of a net, warps 4-7 the odd ones (a chunk needs all 128 TMEM lanes = four warps, each
// thread writing its row's 32 columns), so a warp's wait -> split -> tcgen05.st -> arrive chain is only on the critical path
// of every second chunk.  stage_x counts four warp arrivals.  `chunk0` = chunks staged before this net (ring position).
template <typename BarsT>
__device__ __forceinline__ void l1_stage_loop(BarsT& bars, uint32_t t_x, const XT& xs, long long rowoff, int n, uint32_t chunk0,
                                              int warp, int lane) {
  const int par = warp >> 2, row = 32 * (warp & 3) + lane;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  XRegs xa[2][2];  // two of this warp's chunks in flight, 2 x 16 columns each
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const int k = par + 2 * b;
    if (k < n) {
      load_x(xa[b][0], xs, rowoff, row, 32 * k);
      load_x(xa[b][1], xs, rowoff, row, 32 * k + 16);
    }
  }
  for (int k0 = par; k0 < n; k0 += 4) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int k = k0 + 2 * b;
      if (k < n) {
        const uint32_t cc = chunk0 + (uint32_t)k, s = cc % NS;
        if (cc >= NS) mbar_wait_bounded(&bars.stage_free[s], ((cc / NS) - 1u) & 1u);
        umma::fence_after();
        store_x(t_x + s * 64, lane_base, 0, xa[b][0]);
        store_x(t_x + s * 64, lane_base, 16, xa[b][1]);
        warp_arrive_tmem(&bars.stage_x[s], lane);
        if (k + 4 < n) {
          load_x(xa[b][0], xs, rowoff, row, 32 * (k + 4));
          load_x(xa[b][1], xs, rowoff, row, 32 * (k + 4) + 16);
        }
      }
    }
  }
}

// h2 = relu(relu(X W1 + b1) W2 + b2) for the CTA's 128 rows (compute warps); thread (warp w, lane l) ends up with units
// [32 (w >> 2) + 64 g, +32) of row 32 (w & 3) + l in h2[32 g ..].  Warps run decoupled: each one gathers, splits and
// writes its rows of a chunk and arrives on the slot barrier; nothing but the accumulator barrier joins them.

template <int U, typename Overlap>
__device__ __forceinline__ void forward_hidden_tc(Bars* bars, Pipe& pipe, uint32_t tbase, const XT& xs, const MlpW& w,
                                                  const long long* sRow, float* sB1, float* sB2, float* sW3, float* sB3,
                                                  float (&h2)[U / 2], Overlap&& overlap) {
  using L = Lay<U>;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half = warp >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  const long long rowoff = sRow[row];
  const int nchunks = (w.in + 31) / 32;
  for (int i = tid; i < U; i += NTC) { sB1[i] = w.b1[i]; sB2[i] = w.b2[i]; }
  for (int i = tid; i < U * w.out; i += NTC) sW3[i] = w.W3[i];
  if (tid < w.out) sB3[tid] = w.b3[tid];
  // ---- layer 1: K chunks through the NS-slot ring
  l1_stage_loop(*bars, tbase + L::T_X, xs, rowoff, nchunks, pipe.chunks, warp, lane);
  pipe.chunks += (uint32_t)nchunks;
  named_sync();  // small tensors visible to every compute warp
  overlap();  // work that does not depend on the net's output (the Gumbel noise) hides behind the layer-1 MMAs
  mbar_wait_bounded(&bars->acc, pipe.accs & 1u);
  pipe.accs++;
  umma::fence_after();
  // ---- epilogue 1: h1 = relu(acc + b1) -> split -> TMEM (A operand of layer 2)
#pragma unroll
  for (int g = 0; g < U / 64; ++g) {
    const int c0 = 32 * half + 64 * g;  // this thread's 32 units of the pass
    float v[32];
    umma::tmem_ld32(tbase + L::T_ACC1 + lane_base + (uint32_t)c0, v);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) umma::split_tf32(fmaxf(v[16 * q + i] + sB1[c0 + 16 * q + i], 0.f), hi[i], lo[i]);
      umma::tmem_st16(tbase + L::T_H1 + lane_base + (uint32_t)(c0 + 16 * q), hi);
      umma::tmem_st16(tbase + L::T_H1 + U + lane_base + (uint32_t)(c0 + 16 * q), lo);
    }
  }
  warp_arrive_tmem(&bars->h1_full, lane);
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
    umma::tmem_ld32(tbase + L::T_ACC2 + lane_base + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) h2[32 * g + i] = fmaxf(v[i] + sB2[c0 + i], 0.f);
  }
  umma::fence_before();
}

// out[a] = sum over this thread's units of h2 * W3[:, a]   (a < KK); the two unit halves of a row are combined by the caller
template <int U, int KK>
__device__ __forceinline__ void head_partial(const float (&h2)[U / 2], const float* __restrict__ sW3, int half, float (&out)[TC_MAXK]) {
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
                                                         int act_in_smem, int phase) {
  // phase 0: all target actors, then the target critic, in one CTA (a' stays in shared memory when it fits).
  // phase 1 / 2 (split launch, many agents): grid.z actor groups each compute their share of a' into the L2-resident scratch
  // (phase 1), then one CTA per tile runs the critic on it (phase 2) -- the actor passes of a tile are independent, so the
  // split turns 25 serial nets per CTA into <= 5 and fills the SMs even for a single-agent launch.
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
  float* sB3 = sW3 + U * TC_MAXK;
  float* sPart = sB3 + 16;
  float* sG = sPart + TMR * TC_KPAD;  // Gumbel noise -log(-log u) of the current actor, [row][TC_KPAD]
  float* sQ = sG + TMR * TC_KPAD;
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

  if (warp == 0) umma::tmem_alloc(&tmem_slot, LY::T_COLS);
  if (tid == 0) {
    for (int k = 0; k < NS; ++k) {
      mbar_init(&bars.stage_w[k], 1);
      mbar_init(&bars.stage_x[k], NTC / 64);
      mbar_init(&bars.stage_free[k], 1);
    }
    mbar_init(&bars.h1_full, NTC / 32);
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
  int i_begin = me.local_q ? j : 0, i_end = me.local_q ? j + 1 : C.n_agents;
  if (phase == 1) {  // this CTA's contiguous share of the actors
    const int per = (i_end - i_begin + (int)gridDim.z - 1) / (int)gridDim.z;
    i_begin += (int)blockIdx.z * per;
    i_end = min(i_end, i_begin + per);
  } else if (phase == 2) {
    i_end = i_begin;  // no actor passes
  }
  const bool do_critic = phase != 1;

  if (warp == NTC / 32) {
    // ===== producer warp: weight images of every net of this CTA's job list, in order =====
    if (lane == 0) {
      for (int i = i_begin; i < i_end; ++i) produce_net<U>(smem, &bars, pipe, imgs[i].net[MDP_NET_TARGET_P], C.agents[i].obs_dim);
      if (do_critic) produce_net<U>(smem, &bars, pipe, imgs[j].net[MDP_NET_TARGET_Q], me.net[MDP_NET_TARGET_Q].in);
    }
  } else if (warp == NTC / 32 + 1) {
    // ===== MMA issuer warp (all lanes run the loops; one elected lane issues) =====
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    for (int i = i_begin; i < i_end; ++i) mma_net<U>(smem, &bars, pipe, tb, C.agents[i].obs_dim);
    if (do_critic) mma_net<U>(smem, &bars, pipe, tb, me.net[MDP_NET_TARGET_Q].in);
  } else {
    // ===== compute warps =====
    float h2[U / 2];
    float part[TC_MAXK];
    // a'_i = gumbel_softmax(target_p_i(next_obs_i)) for every agent the critic sees
    for (int i = i_begin; i < i_end; ++i) {
      const AgentDev& ag = C.agents[i];
      const MlpW w = ag.net[MDP_NET_TARGET_P];
      XT xs{batch + L.nx_off + ag.obs_off, ag.obs_dim, nullptr, 0, 0, 0, ((L.nx_off + ag.obs_off) & 3) == 0};
      const int K = ag.act_dim;
      forward_hidden_tc<U>(&bars, pipe, tbase, xs, w, sRow, sB1, sB2, sW3, sB3, h2, [&]() {
        // both unit-halves of a row share its K draws: half 0 takes the even actions, half 1 the odd ones
        for (int a = half; a < K; a += 2) {
          const float u = u_target ? u_target[(row0 + min(row, nrows - 1)) * u_stride + ag.act_off + a]
                                   : philox_u(seed, counter, (uint32_t)(0x100 + i), row0 + row, a);
          sG[row * TC_KPAD + a] = gumbel_from_u(u);
        }
      });
      if (K == 5) head_partial<U, 5>(h2, sW3, half, part);
      else if (K == 9) head_partial<U, 9>(h2, sW3, half, part);
      else {
        for (int a = 0; a < TC_MAXK; ++a) part[a] = 0.f;
        for (int a = 0; a < K; ++a)
          for (int g = 0; g < U / 64; ++g)
            for (int t = 0; t < 32; ++t) part[a] = fmaf(h2[32 * g + t], sW3[(32 * half + 64 * g + t) * K + a], part[a]);
      }
      if (half == 1)
        for (int a = 0; a < K; ++a) sPart[row * TC_KPAD + a] = part[a];
      named_sync();
      if (half == 0) {  // one thread per batch row: logits -> Gumbel-softmax per head (distributions.py:264-266, 332-336)
        float z[TC_MAXK];
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a) {
          if (a < K) {
            const float logit = part[a] + sPart[row * TC_KPAD + a] + sB3[a];
            z[a] = logit + sG[row * TC_KPAD + a];
          }
        }
        for (int h = 0; h < ag.n_heads; ++h) {
          const int o = h ? ag.head_dim[0] : 0, n = ag.head_dim[h];
          float m = -INFINITY;
#pragma unroll
          for (int a = 0; a < TC_MAXK; ++a)
            if (a >= o && a < o + n) m = fmaxf(m, z[a]);
          float ssum = 0.f;
#pragma unroll
          for (int a = 0; a < TC_MAXK; ++a)
            if (a >= o && a < o + n) { z[a] = expf(z[a] - m); ssum += z[a]; }
#pragma unroll
          for (int a = 0; a < TC_MAXK; ++a)
            if (a >= o && a < o + n) z[a] = z[a] / ssum;
        }
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a < K) {
            actT[(size_t)row * ASP + ag.act_off + a] = z[a];
            if (target_act_out && row < nrows) target_act_out[(row0 + row) * u_stride + ag.act_off + a] = z[a];
          }
      }
      named_sync();  // a' visible to the critic's gather; sPart / small tensors free for the next net
    }

    if (do_critic) {
    // q' = target_q_j([next_obs | a'])
    XT xq;
    if (me.local_q) xq = XT{batch + L.nx_off + me.obs_off, me.obs_dim, actT + me.act_off, ASP, me.obs_dim, me.act_dim,
                            ((L.nx_off + me.obs_off) & 3) == 0};
    else xq = XT{batch + L.nx_off, C.obs_sum, actT, ASP, C.obs_sum, C.act_sum, (L.nx_off & 3) == 0};
    const MlpW tq = me.net[MDP_NET_TARGET_Q];
    forward_hidden_tc<U>(&bars, pipe, tbase, xq, tq, sRow, sB1, sB2, sW3, sB3, h2, []() {});
    head_partial<U, 1>(h2, sW3, half, part);
    if (half == 1) sPart[row * TC_KPAD] = part[0];
    named_sync();
    if (half == 0) {
      // y = float32(rew + gamma * (1 - done) * q')  -- float64 combine like numpy (maddpg.py:186)
      double sy = 0, syy = 0, sr = 0, sq = 0;
      if (row < nrows) {
        const float qn = part[0] + sPart[row * TC_KPAD] + sB3[0];
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
    }  // do_critic
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tbase, LY::T_COLS);
}


// =============================================================================================================
// Grouped actor inference + Gumbel-softmax on tensor cores: MADDPGAgentTrainer.action for E lockstep env instances
// (maddpg/trainer/maddpg.py:151-152 -> act :62 -> mlp_model train.py:39-46 + SoftCategoricalPd.sample distributions.py:264-266).
// The per-step rollout of the larger configs spends most of its time here (measured on B200, fp32 SIMT kernel: 55 of 70 us
// per lockstep step for simple_tag with 16384 env instances, 1.2 of 2.0 ms for simple_spread N=24 with 32768).
// grid = (ceil(E / 128), agent_count): one CTA = 128 env instances of one agent, same building blocks as the TD-target
// kernel above (observation rows -> TMEM A operand, W^T images by TMA, 3xTF32, head + Gumbel-softmax as row-local register
// code); same Philox keys as k_actor_act, so the two kernels draw the same noise.
template <int U>
__global__ void __launch_bounds__(NTT, 1) k_actor_act_tc(CoreDev C, const AgentImg* __restrict__ imgs, int agent_begin, int net,
                                                         int E, const float* __restrict__ obs, int obs_stride,
                                                         float* __restrict__ act, int act_stride, const float* __restrict__ u,
                                                         uint64_t seed, uint64_t counter, float* __restrict__ logits_out,
                                                         long long rng_row_base) {
  using LY = Lay<U>;
  const int i = agent_begin + blockIdx.y;
  if (C.ctl) counter += C.ctl[0];
  extern __shared__ unsigned char smem_raw[];
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  unsigned char* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  float* misc = reinterpret_cast<float*>(smem + LY::OFF_MISC);
  float* sB1 = misc;
  float* sB2 = sB1 + U;
  float* sW3 = sB2 + U;
  float* sB3 = sW3 + U * TC_MAXK;
  float* sPart = sB3 + 16;
  float* sG = sPart + TMR * TC_KPAD;
  float* sQ = sG + TMR * TC_KPAD;
  float* sRD = sQ + TMR;
  long long* sRow = reinterpret_cast<long long*>(sRD + 2 * TMR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half = warp >> 2;
  const AgentDev& ag = C.agents[i];
  const long long row0 = (long long)blockIdx.x * TMR;
  const int nrows = (int)min((long long)TMR, E - row0);

  if (warp == 0) umma::tmem_alloc(&tmem_slot, LY::T_COLS);
  if (tid == 0) {
    for (int k = 0; k < NS; ++k) {
      mbar_init(&bars.stage_w[k], 1);
      mbar_init(&bars.stage_x[k], NTC / 64);
      mbar_init(&bars.stage_free[k], 1);
    }
    mbar_init(&bars.h1_full, NTC / 32);
    mbar_init(&bars.w2_full, 1);
    mbar_init(&bars.w2_free, 1);
    mbar_init(&bars.acc, 1);
  }
  if (tid < TMR) sRow[tid] = (row0 + min(tid, nrows - 1)) * (long long)obs_stride;  // tail rows replay the last valid row
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  Pipe pipe{0u, 0u, 0u};
  if (warp == NTC / 32) {
    if (lane == 0) produce_net<U>(smem, &bars, pipe, imgs[i].net[net], ag.obs_dim);
  } else if (warp == NTC / 32 + 1) {
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    mma_net<U>(smem, &bars, pipe, tb, ag.obs_dim);
  } else {
    float h2[U / 2];
    float part[TC_MAXK];
    const MlpW w = ag.net[net];
    XT xs{obs + ag.obs_off, ag.obs_dim, nullptr, 0, 0, 0, ((ag.obs_off & 3) == 0 && (obs_stride & 3) == 0) ? 1 : 0};
    const int K = ag.act_dim;
    forward_hidden_tc<U>(&bars, pipe, tbase, xs, w, sRow, sB1, sB2, sW3, sB3, h2, [&]() {
      // both unit-halves of a row share its K draws: half 0 takes the even actions, half 1 the odd ones
      for (int a = half; a < K; a += 2) {
        const long long r = row0 + min(row, nrows - 1);
        const float uu = u ? u[r * act_stride + ag.act_off + a] : philox_u(seed, counter, (uint32_t)i, r + rng_row_base, a);
        sG[row * TC_KPAD + a] = gumbel_from_u(uu);
      }
    });
    if (K == 5) head_partial<U, 5>(h2, sW3, half, part);
    else if (K == 9) head_partial<U, 9>(h2, sW3, half, part);
    else {
      for (int a = 0; a < TC_MAXK; ++a) part[a] = 0.f;
      for (int a = 0; a < K; ++a)
        for (int g = 0; g < U / 64; ++g)
          for (int t = 0; t < 32; ++t) part[a] = fmaf(h2[32 * g + t], sW3[(32 * half + 64 * g + t) * K + a], part[a]);
    }
    if (half == 1)
      for (int a = 0; a < K; ++a) sPart[row * TC_KPAD + a] = part[a];
    named_sync();
    if (half == 0 && row < nrows) {  // one thread per env instance: logits -> Gumbel-softmax per head
      float z[TC_MAXK], lg[TC_MAXK];
#pragma unroll
      for (int a = 0; a < TC_MAXK; ++a) {
        if (a < K) {
          lg[a] = part[a] + sPart[row * TC_KPAD + a] + sB3[a];
          z[a] = lg[a] + sG[row * TC_KPAD + a];
        }
      }
      for (int h = 0; h < ag.n_heads; ++h) {
        const int o = h ? ag.head_dim[0] : 0, n = ag.head_dim[h];
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a >= o && a < o + n) m = fmaxf(m, z[a]);
        float ssum = 0.f;
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a >= o && a < o + n) { z[a] = expf(z[a] - m); ssum += z[a]; }
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a >= o && a < o + n) z[a] = z[a] / ssum;
      }
      float* arow = act + (row0 + row) * (long long)act_stride + ag.act_off;
#pragma unroll
      for (int a = 0; a < TC_MAXK; ++a)
        if (a < K) {
          arow[a] = z[a];
          if (logits_out) logits_out[(row0 + row) * (long long)act_stride + ag.act_off + a] = lg[a];
        }
    }
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tbase, LY::T_COLS);
}


// =============================================================================================================
// Fused critic forward + MSE + backward on tensor cores (q_train, maddpg.py:75-100), maddpg-mode critics, U = 64.
//   forward   z1 = X W1            A = X (TMEM, gathered rows), B = W1^T chunk images (TMA)        -> acc1
//             z2 = h1 W2           A = h1 (TMEM),               B = W2^T images (TMA)              -> acc2
//   backward  dW2 = h1^T dz2       A = h1, B = dz2: MN-major SWIZZLE_128B_BASE32B smem images       -> acc1 (M = 64)
//             dh1 = dz2 W2^T       A = dz2 (TMEM),              B = W2 images (TMA)                 -> acc2
//             dW1^T = dz1^T X      A = dz1 (MN-major smem), B = X chunk (MN-major smem), per chunk  -> ring of 32-col slots
// Every GEMM is 3xTF32.  Bias / W3 gradients and the loss are row-local register code plus a 31-shuffle column sum.
// Weight-gradient tiles leave TMEM as fp32 RED atomics into the flat gradient bucket (8x fewer than the 16-row SIMT tiles).
// Shared memory regions are recycled by phase: R1 = W1^T slots -> dz2 images -> X chunk slots; R2 = W2^T -> W2;
// R3 = h1 images -> dz1 images.
// =============================================================================================================
template <int U>
struct LayB {
  using L = Lay<U>;
  static constexpr uint32_t ACT_IMG = TMR * 128 * (U / 32);  // [128 rows][U] MN-major image            32 KB
  static constexpr uint32_t XS_IMG = TMR * 128;               // [128 rows][32] MN-major X chunk image   16 KB
  static constexpr uint32_t R1 = 0, R2 = L::OFF_W2, R3 = L::OFF_W2 + 2 * L::W2_IMG;
  static constexpr uint32_t OFF_MISC = R3 + 2 * ACT_IMG;
  static constexpr int MISC_FLOATS = U + U + U + 16 + 3 * TMR + 2 * TMR;  // b1 b2 W3 b3 | part dq y | rowoff
  static_assert(2 * ACT_IMG <= NS * 2 * L::W_IMG, "dz2 images must fit the W1^T slot region");
  static constexpr uint32_t T_DW = L::T_X;  // dW1 chunk accumulators: 4 slots x 32 columns
};

struct BarsB {
  unsigned long long stage_w[NS], stage_x[NS], stage_free[NS];
  unsigned long long w2_full, h1_full, acc, l2_done, w2n_full, dz2_full, dz1_full;
  unsigned long long xs_full[2], xs_free[2], dw_full[4], dw_free[4];
};

// lane l returns the sum over the warp's 32 lanes of v[l] (butterfly transpose-reduce, 31 shuffles); v is clobbered
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// D (+)= A * B^T, both operands MN-major smem image pairs (hi at +0, lo at +lo_off), k-step = 8 image rows = 1024 bytes
template <int M, int N, int NSTEPS>
__device__ __forceinline__ void issue_3x_mn(uint32_t tacc, uint32_t a_img, uint32_t a_lo_off, uint32_t a_panel, uint32_t b_img,
                                            uint32_t b_lo_off, uint32_t b_panel) {
  constexpr uint32_t idesc = umma::idesc_tf32(M, N, 1, 1);
  const uint64_t ah = umma::desc_mn(a_img, a_panel, 0), al = umma::desc_mn(a_img + a_lo_off, a_panel, 0);
  const uint64_t bh = umma::desc_mn(b_img, b_panel, 0), bl = umma::desc_mn(b_img + b_lo_off, b_panel, 0);
#pragma unroll
  for (int s = 0; s < NSTEPS; ++s) {
    const uint32_t o = 64u * s;  // 1024 bytes in descriptor units
    umma::mma_tf32(tacc, al + o, bh + o, idesc, s == 0 ? 0u : 1u);
    umma::mma_tf32(tacc, ah + o, bl + o, idesc, 1u);
    umma::mma_tf32(tacc, ah + o, bh + o, idesc, 1u);
  }
}

// this thread's 32 units of batch row `row` -> MN-major image pair (panel = 32-unit group c0 / 32)
__device__ __forceinline__ void store_act_mn(unsigned char* img_hi, uint32_t lo_off, int row, int c0, const float (&v)[32]) {
  unsigned char* phi = img_hi + (uint32_t)(c0 >> 5) * (TMR * 128);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 hi, lo;
    umma::split_tf32(v[4 * q + 0], hi.x, lo.x);
    umma::split_tf32(v[4 * q + 1], hi.y, lo.y);
    umma::split_tf32(v[4 * q + 2], hi.z, lo.z);
    umma::split_tf32(v[4 * q + 3], hi.w, lo.w);
    const uint32_t off = umma::sw128b32_off(row, 4 * q);
    *reinterpret_cast<float4*>(phi + off) = hi;
    *reinterpret_cast<float4*>(phi + lo_off + off) = lo;
  }
}
// the same values as the A operand of a TMEM-sourced GEMM (hi at region + c0, lo at region + U + c0)
template <int U>
__device__ __forceinline__ void store_act_tmem(uint32_t region, uint32_t lane_base, int c0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) umma::split_tf32(v[16 * q + i], hi[i], lo[i]);
    umma::tmem_st16(region + lane_base + (uint32_t)(c0 + 16 * q), hi);
    umma::tmem_st16(region + U + lane_base + (uint32_t)(c0 + 16 * q), lo);
  }
}
__device__ __forceinline__ void warp_arrive_both(unsigned long long* bar, int lane) {
  umma::fence_async_smem();
  umma::tmem_st_wait();
  umma::fence_before();
  __syncwarp();
  if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// dW1^T = dz1^T X, the chunk loop of the weight-gradient phase, with the eight compute warps split into two roles so that
// neither sits on the other's latency chain: warps 0-3 ("stagers", thread = batch row) gather all 32 columns of a chunk,
// split them and store the MN-major image pair; warps 4-7 ("readers") pull the finished 64 x 32 accumulator of the previous
// chunks out of TMEM and reduce it into the gradient bucket.  xs_full / dw_free therefore count four warp arrivals.
template <int U, typename BarsT>
__device__ __forceinline__ void dw1_chunk_loop(BarsT& bars, unsigned char* x_slots, uint32_t xs_img_bytes, uint32_t tdw, const XT& xs,
                                               long long rowoff, int nchunks, int in_dim, float* __restrict__ gW1, int warp, int lane,
                                               int chunk0 = 0) {
  const int row = 32 * (warp & 3) + lane;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  if (warp < 4) {
    XRegs xa[2][2];  // two chunks in flight, 2 x 16 columns each
#pragma unroll
    for (int b = 0; b < 2; ++b)
      if (b < nchunks) {
        load_x(xa[b][0], xs, rowoff, row, 32 * (chunk0 + b));
        load_x(xa[b][1], xs, rowoff, row, 32 * (chunk0 + b) + 16);
      }
    for (int c = 0; c < nchunks; c += 2) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (c + b < nchunks) {
          const int cc = c + b, s = cc & 1;
          if (cc >= 2) mbar_wait_bounded(&bars.xs_free[s], ((cc >> 1) - 1) & 1);
          unsigned char* img = x_slots + s * (2 * xs_img_bytes);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 hi, lo;
              umma::split_tf32(xa[b][h].v[4 * q + 0], hi.x, lo.x);
              umma::split_tf32(xa[b][h].v[4 * q + 1], hi.y, lo.y);
              umma::split_tf32(xa[b][h].v[4 * q + 2], hi.z, lo.z);
              umma::split_tf32(xa[b][h].v[4 * q + 3], hi.w, lo.w);
              const uint32_t off = umma::sw128b32_off(row, 16 * h + 4 * q);
              *reinterpret_cast<float4*>(img + off) = hi;
              *reinterpret_cast<float4*>(img + xs_img_bytes + off) = lo;
            }
          umma::fence_async_smem();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars.xs_full[s])) : "memory");
          if (cc + 2 < nchunks) {
            load_x(xa[b][0], xs, rowoff, row, 32 * (chunk0 + cc + 2));
            load_x(xa[b][1], xs, rowoff, row, 32 * (chunk0 + cc + 2) + 16);
          }
        }
      }
    }
  } else {
    for (int cc = 0; cc < nchunks; ++cc) {
      const int t = cc & 3;
      mbar_wait_bounded(&bars.dw_full[t], (cc >> 2) & 1);
      umma::fence_after();
      float d[32];
      umma::tmem_ld32(tdw + t * 32 + lane_base, d);  // M = 64: unit u = 16 (w & 3) + lane lives on lanes < 16 of each quarter
      umma::fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars.dw_free[t])) : "memory");
      if (lane < 16) {
        const int u = 16 * (warp & 3) + lane, f0 = 32 * (chunk0 + cc);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (f0 + i < in_dim) red_add(gW1 + (size_t)(f0 + i) * U + u, d[i]);
      }
    }
  }
}

template <int U>
__global__ void __launch_bounds__(NTT, 1) k_critic_grads_tc(CoreDev C, const AgentImg* __restrict__ imgs, int j0, mdp_ring_layout L, int B,
                                                            const float* __restrict__ batch, const long long* __restrict__ ridx,
                                                            const float* __restrict__ y, float* __restrict__ q_out, long long idx_stride,
                                                            long long y_stride, float* __restrict__ dz1_scratch) {
  // dz1_scratch != null (wide critics): the dW1 phase runs in k_dw1_tc with the K range spread over many CTAs; this kernel
  // then ends by writing the tile's dz1 (fp32) to the scratch.
  using LY = Lay<U>;
  using LB = LayB<U>;
  const int j = j0 + blockIdx.y;
  if (ridx) ridx += blockIdx.y * idx_stride;
  y += blockIdx.y * y_stride;
  extern __shared__ unsigned char smem_raw[];
  __shared__ BarsB bars;
  __shared__ uint32_t tmem_slot;
  unsigned char* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  float* misc = reinterpret_cast<float*>(smem + LB::OFF_MISC);
  float* sB1 = misc;
  float* sB2 = sB1 + U;
  float* sW3 = sB2 + U;
  float* sB3 = sW3 + U;
  float* sPart = sB3 + 16;
  float* sDq = sPart + TMR;
  float* sY = sDq + TMR;
  long long* sRow = reinterpret_cast<long long*>(sY + TMR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const AgentDev& me = C.agents[j];
  const MlpW w = me.net[MDP_NET_Q];
  const MlpG& g = me.grad[1];
  const long long row0 = (long long)blockIdx.x * TMR;
  const int nrows = (int)min((long long)TMR, B - row0);
  const int R = L.row_stride;
  const int nchunks = (w.in + 31) / 32;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, LY::T_COLS);
  if (tid == 0) {
    for (int k = 0; k < NS; ++k) {
      mbar_init(&bars.stage_w[k], 1);
      mbar_init(&bars.stage_x[k], NTC / 64);
      mbar_init(&bars.stage_free[k], 1);
    }
    mbar_init(&bars.w2_full, 1); mbar_init(&bars.h1_full, NTC / 32); mbar_init(&bars.acc, 1);
    mbar_init(&bars.l2_done, 1); mbar_init(&bars.w2n_full, 1);
    mbar_init(&bars.dz2_full, NTC / 32); mbar_init(&bars.dz1_full, NTC / 32);
    for (int k = 0; k < 2; ++k) { mbar_init(&bars.xs_full[k], NTC / 64); mbar_init(&bars.xs_free[k], 1); }
    for (int k = 0; k < 4; ++k) { mbar_init(&bars.dw_full[k], 1); mbar_init(&bars.dw_free[k], NTC / 64); }
    if (blockIdx.x == 0) C.adam_t[2 * j + 1] += 1;  // one more Adam step for this net
  }
  if (tid < TMR) {
    const long long rl = row0 + min(tid, nrows - 1);
    const long long rg = ridx ? ridx[rl] : rl;
    sRow[tid] = rg * R;
    sY[tid] = y[rl];
  }
  for (int i = tid; i < U; i += NTT) { sB1[i] = w.b1[i]; sB2[i] = w.b2[i]; sW3[i] = w.W3[i]; }
  if (tid == 0) sB3[0] = w.b3[0];
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == NTC / 32) {
    // ===== TMA producer =====
    if (lane == 0) {
      const NetImg img = imgs[j].net[MDP_NET_Q];
      mbar_arrive_expect_tx(&bars.w2_full, 2 * LY::W2_IMG);
      bulk_g2s(smem + LB::R2, img.w2, 2 * LY::W2_IMG, &bars.w2_full);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % NS;
        if (c >= NS) mbar_wait_bounded(&bars.stage_free[s], ((c / NS) - 1) & 1);
        mbar_arrive_expect_tx(&bars.stage_w[s], 2 * LY::W_IMG);
        bulk_g2s(smem + LB::R1 + s * (2 * LY::W_IMG), img.w1 + (size_t)c * (2 * LY::W_IMG), 2 * LY::W_IMG, &bars.stage_w[s]);
      }
      mbar_wait_bounded(&bars.l2_done, 0);  // layer-2 MMAs have consumed W2^T: the region now takes W2 for the backward pass
      mbar_arrive_expect_tx(&bars.w2n_full, 2 * LY::W2_IMG);
      bulk_g2s(smem + LB::R2, img.w2n, 2 * LY::W2_IMG, &bars.w2n_full);
    }
  } else if (warp == NTC / 32 + 1) {
    // ===== MMA issuer (all lanes run the loops; one elected lane issues) =====
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    const uint64_t w_hi0 = umma::desc_k(sbase + LB::R1, LY::W_IMG, 0), w_lo0 = umma::desc_k(sbase + LB::R1 + LY::W_IMG, LY::W_IMG, 0);
    const uint64_t r2_hi = umma::desc_k(sbase + LB::R2, LY::W_IMG, 0), r2_lo = umma::desc_k(sbase + LB::R2 + LY::W2_IMG, LY::W_IMG, 0);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % NS, ph = (c / NS) & 1;
      mbar_wait_bounded(&bars.stage_x[s], ph);
      mbar_wait_bounded(&bars.stage_w[s], ph);
      umma::fence_after();
      const uint32_t so = s * ((2 * LY::W_IMG) >> 4);
      if (umma::elect_one()) {
        issue_3x<U, 4>(tb + LY::T_ACC1, tb + LY::T_X + s * 64, 32, w_hi0 + so, w_lo0 + so, c > 0 ? 1u : 0u);
        umma::commit(&bars.stage_free[s]);
        if (c == nchunks - 1) umma::commit(&bars.acc);
      }
      __syncwarp();
    }
    mbar_wait_bounded(&bars.h1_full, 0);
    mbar_wait_bounded(&bars.w2_full, 0);
    umma::fence_after();
    if (umma::elect_one()) {
      issue_3x<U, U / 8>(tb + LY::T_ACC2, tb + LY::T_H1, U, r2_hi, r2_lo, 0u);
      umma::commit(&bars.acc);
      umma::commit(&bars.l2_done);
    }
    __syncwarp();
    // backward: dW2 = h1^T dz2 (acc1, M = 64) and dh1 = dz2 W2^T (acc2)
    mbar_wait_bounded(&bars.dz2_full, 0);
    mbar_wait_bounded(&bars.w2n_full, 0);
    umma::fence_after();
    if (umma::elect_one()) {
      issue_3x_mn<U, U, TMR / 8>(tb + LY::T_ACC1, sbase + LB::R3, LB::ACT_IMG, TMR * 128, sbase + LB::R1, LB::ACT_IMG, TMR * 128);
      issue_3x<U, U / 8>(tb + LY::T_ACC2, tb + LY::T_H1, U, r2_hi, r2_lo, 0u);
      umma::commit(&bars.acc);
    }
    __syncwarp();
    // dW1^T chunks: D[u][f] = sum_r dz1[r][u] X[r][f]
    if (!dz1_scratch) mbar_wait_bounded(&bars.dz1_full, 0);
    for (int c = 0; c < (dz1_scratch ? 0 : nchunks); ++c) {
      const int s = c & 1, t = c & 3;
      mbar_wait_bounded(&bars.xs_full[s], (c >> 1) & 1);
      if (c >= 4) mbar_wait_bounded(&bars.dw_free[t], ((c >> 2) - 1) & 1);
      umma::fence_after();
      if (umma::elect_one()) {
        issue_3x_mn<U, 32, TMR / 8>(tb + LB::T_DW + t * 32, sbase + LB::R3, LB::ACT_IMG, TMR * 128,
                                   sbase + LB::R1 + s * (2 * LB::XS_IMG), LB::XS_IMG, TMR * 128);
        umma::commit(&bars.xs_free[s]);
        umma::commit(&bars.dw_full[t]);
      }
      __syncwarp();
    }
  } else {
    // ===== compute warps: thread = (batch row, unit half) =====
    const int row = 32 * (warp & 3) + lane, half = warp >> 2, c0 = 32 * half;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const long long rowoff = sRow[row];
    const XT xs{batch, L.x_dim, nullptr, 0, 0, 0, 1};
    l1_stage_loop(bars, tbase + LY::T_X, xs, rowoff, nchunks, 0u, warp, lane);
    // ---- epilogue 1: h1 -> TMEM (A of layer 2) and MN-major smem images (A of dW2); relu mask kept in a register
    mbar_wait_bounded(&bars.acc, 0);
    umma::fence_after();
    uint32_t mask1 = 0, mask2 = 0;
    float v[32];
    umma::tmem_ld32(tbase + LY::T_ACC1 + lane_base + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = fmaxf(v[i] + sB1[c0 + i], 0.f);
      mask1 |= (v[i] > 0.f ? 1u : 0u) << i;
    }
    store_act_tmem<U>(tbase + LY::T_H1, lane_base, c0, v);
    store_act_mn(smem + LB::R3, LB::ACT_IMG, row, c0, v);
    warp_arrive_both(&bars.h1_full, lane);
    // ---- epilogue 2: h2, q, loss, dq, dW3/db3, dz2
    mbar_wait_bounded(&bars.acc, 1);
    umma::fence_after();
    umma::tmem_ld32(tbase + LY::T_ACC2 + lane_base + (uint32_t)c0, v);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = fmaxf(v[i] + sB2[c0 + i], 0.f);
      mask2 |= (v[i] > 0.f ? 1u : 0u) << i;
      part = fmaf(v[i], sW3[c0 + i], part);
    }
    if (half == 1) sPart[row] = part;
    named_sync();
    if (half == 0) {
      float d = 0.f;
      double se = 0.0;
      if (row < nrows) {
        const float q = part + sPart[row] + sB3[0];
        const float diff = q - sY[row];
        d = 2.0f * diff / (float)B;  // dL/dq of mean((q - y)^2)
        se = (double)diff * (double)diff;
        if (q_out) q_out[row0 + row] = q;
      }
      sDq[row] = d;
      float ds = d;
      for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
      }
      if (lane == 0) {
        atomicAdd(C.stats + 8 * j + 0, se);
        red_add(g.b3, ds);
      }
    }
    named_sync();
    const float dq = sDq[row];
    {
      float t[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = v[i] * dq;  // gW3[u] = sum_r h2[r][u] dq[r]
      const float sum = warp_colsum32(t, lane);
      red_add(g.W3 + c0 + lane, sum);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = ((mask2 >> i) & 1u) ? dq * sW3[c0 + i] : 0.f;  // dz2 = dq W3^T relu'(h2)
    store_act_tmem<U>(tbase + LY::T_H1, lane_base, c0, v);   // layer-2 MMAs are complete: the h1 operand region is free
    store_act_mn(smem + LB::R1, LB::ACT_IMG, row, c0, v);    // layer-1 MMAs are complete: the W1^T slots are free
    warp_arrive_both(&bars.dz2_full, lane);
    {
      const float sum = warp_colsum32(v, lane);  // gb2 = sum_r dz2
      red_add(g.b2 + c0 + lane, sum);
    }
    // ---- epilogue 3: dW2 tile -> gradient bucket; dz1 = dh1 relu'(h1) -> MN-major images (A of dW1^T)
    mbar_wait_bounded(&bars.acc, 0);
    umma::fence_after();
    umma::tmem_ld32(tbase + LY::T_ACC1 + lane_base + (uint32_t)c0, v);  // M = 64: row m lives on lane 32 (m / 16) + m % 16
    if (lane < 16) {
      float* dst = g.W2 + (size_t)(16 * (warp & 3) + lane) * U + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 4) red_add4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    umma::tmem_ld32(tbase + LY::T_ACC2 + lane_base + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = ((mask1 >> i) & 1u) ? v[i] : 0.f;
    umma::fence_before();
    if (dz1_scratch) {
      float4* dst = reinterpret_cast<float4*>(dz1_scratch + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * TMR + row) * U + c0);
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
      store_act_mn(smem + LB::R3, LB::ACT_IMG, row, c0, v);  // dW2's MMAs are complete: the h1 images are dead
      warp_arrive_both(&bars.dz1_full, lane);
    }
    {
      const float sum = warp_colsum32(v, lane);  // gb1 = sum_r dz1
      red_add(g.b1 + c0 + lane, sum);
    }
    // ---- dW1^T chunks: stager warps / reader warps
    if (!dz1_scratch)
      dw1_chunk_loop<U>(bars, smem + LB::R1, LB::XS_IMG, tbase + LB::T_DW, xs, rowoff, nchunks, w.in, g.W1, warp, lane);
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tbase, LY::T_COLS);
}


// dW1^T = dz1^T X for wide critics, spread over grid.z K-ranges: every CTA rebuilds the tile's dz1 images from the fp32
// scratch that k_critic_grads_tc left behind and then runs the same stager / MMA / reader pipeline on its share of the chunks.
struct BarsD {
  unsigned long long xs_full[2], xs_free[2], dw_full[4], dw_free[4];
};
constexpr int NTD = NTC + 32;  // 8 compute warps + the MMA issuer warp

template <int U>
__global__ void __launch_bounds__(NTD, 1) k_dw1_tc(CoreDev C, int j0, mdp_ring_layout L, int B, const float* __restrict__ batch,
                                                   const long long* __restrict__ ridx, long long idx_stride,
                                                   const float* __restrict__ dz1_scratch, int chunks_per_cta) {
  constexpr uint32_t ACT_IMG = TMR * 128 * (U / 32), XS_IMG = TMR * 128;
  constexpr uint32_t R_DZ = 0, R_X = 2 * ACT_IMG, OFF_MISC = R_X + 4 * XS_IMG;
  const int j = j0 + blockIdx.y;
  if (ridx) ridx += blockIdx.y * idx_stride;
  extern __shared__ unsigned char smem_raw[];
  __shared__ BarsD bars;
  __shared__ uint32_t tmem_slot;
  unsigned char* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  long long* sRow = reinterpret_cast<long long*>(smem + OFF_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const AgentDev& me = C.agents[j];
  const MlpW w = me.net[MDP_NET_Q];
  const long long row0 = (long long)blockIdx.x * TMR;
  const int nrows = (int)min((long long)TMR, B - row0);
  const int nchunks_all = (w.in + 31) / 32;
  const int chunk0 = (int)blockIdx.z * chunks_per_cta;
  const int nchunks = max(0, min(chunks_per_cta, nchunks_all - chunk0));
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
  if (tid == 0) {
    for (int k = 0; k < 2; ++k) { mbar_init(&bars.xs_full[k], NTC / 64); mbar_init(&bars.xs_free[k], 1); }
    for (int k = 0; k < 4; ++k) { mbar_init(&bars.dw_full[k], 1); mbar_init(&bars.dw_free[k], NTC / 64); }
  }
  if (tid < TMR) {
    const long long rl = row0 + min(tid, nrows - 1);
    sRow[tid] = (ridx ? ridx[rl] : rl) * L.row_stride;
  }
  if (warp < NTC / 32) {  // dz1 (fp32 scratch) -> MN-major image pair
    const int row = 32 * (warp & 3) + lane, c0 = 32 * (warp >> 2);
    const float4* src = reinterpret_cast<const float4*>(dz1_scratch + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * TMR + row) * U + c0);
    float v[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = src[q];
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    store_act_mn(smem + R_DZ, ACT_IMG, row, c0, v);
    umma::fence_async_smem();
  }
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  if (warp == NTC / 32) {
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1, t = c & 3;
      mbar_wait_bounded(&bars.xs_full[s], (c >> 1) & 1);
      if (c >= 4) mbar_wait_bounded(&bars.dw_free[t], ((c >> 2) - 1) & 1);
      umma::fence_after();
      if (umma::elect_one()) {
        issue_3x_mn<U, 32, TMR / 8>(tb + t * 32, sbase + R_DZ, ACT_IMG, TMR * 128, sbase + R_X + s * (2 * XS_IMG), XS_IMG, TMR * 128);
        umma::commit(&bars.xs_free[s]);
        umma::commit(&bars.dw_full[t]);
      }
      __syncwarp();
    }
  } else {
    const int row = 32 * (warp & 3) + lane;
    const XT xs{batch, L.x_dim, nullptr, 0, 0, 0, 1};
    dw1_chunk_loop<U>(bars, smem + R_X, XS_IMG, tbase, xs, sRow[row], nchunks, w.in, me.grad[1].W1, warp, lane, chunk0);
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tbase, 128);
}

// =============================================================================================================
// Fused actor update on tensor cores (p_train, maddpg.py:28-61), maddpg-mode critics, U = 64:
//   actor forward (2 GEMMs) -> Gumbel-softmax sample -> running critic forward on [o, a_-j, a_hat_j] (2 GEMMs) ->
//   critic backward to the action columns (dh1q = dz2q W2q^T on tensor cores, dQ/da in registers) -> softmax Jacobian +
//   logit regulariser -> actor backward (dW2p = h1p^T dz2p, dh1p = dz2p W2p^T, dW1p^T = dz1p^T X_p per chunk).
// Same operand placement as k_critic_grads_tc; the weight-image region R2 is reloaded four times by the TMA producer
// (W2p^T, W2q^T, W2q, W2p), gated by commit barriers.
// =============================================================================================================
template <int U>
struct LayA {
  using L = Lay<U>;
  static constexpr uint32_t ACT_IMG = TMR * 128 * (U / 32);
  static constexpr uint32_t XS_IMG = TMR * 128;
  static constexpr uint32_t R1 = 0, R2 = L::OFF_W2, S1 = L::OFF_W2 + 2 * L::W2_IMG;
  static constexpr uint32_t OFF_MISC = S1 + 2 * ACT_IMG;
  // b1p b2p W3p b3p | b1q b2q W3q b3q | W1q action rows | part dl act noise | rowoff
  static constexpr int MISC_FLOATS = (2 * U + U * TC_MAXK + 16) + (3 * U + 16) + TC_MAXK * U + 4 * TMR * TC_KPAD + 2 * TMR;
  static constexpr uint32_t T_DW = L::T_X;
};

struct BarsA {
  unsigned long long stage_w[NS], stage_x[NS], stage_free[NS];
  unsigned long long a_full, r2_full, r2_free, acc, dz1_full;
  unsigned long long xs_full[2], xs_free[2], dw_full[4], dw_free[4];
};

template <int U>
__global__ void __launch_bounds__(NTT, 1) k_actor_grads_tc(CoreDev C, const AgentImg* __restrict__ imgs, int j0, mdp_ring_layout L, int B,
                                                           const float* __restrict__ batch, const long long* __restrict__ ridx,
                                                           const float* __restrict__ u_actor, int u_stride, uint64_t seed,
                                                           uint64_t counter, long long idx_stride) {
  using LY = Lay<U>;
  using LA = LayA<U>;
  const int j = j0 + blockIdx.y;
  if (ridx) ridx += blockIdx.y * idx_stride;
  if (C.ctl) counter += C.ctl[0];
  extern __shared__ unsigned char smem_raw[];
  __shared__ BarsA bars;
  __shared__ uint32_t tmem_slot;
  unsigned char* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  float* misc = reinterpret_cast<float*>(smem + LA::OFF_MISC);
  float* sB1p = misc;
  float* sB2p = sB1p + U;
  float* sW3p = sB2p + U;
  float* sB3p = sW3p + U * TC_MAXK;
  float* sB1q = sB3p + 16;
  float* sB2q = sB1q + U;
  float* sW3q = sB2q + U;
  float* sB3q = sW3q + U;
  float* sW1a = sB3q + 16;             // [K][U]: critic W1 rows of agent j's action columns
  float* sPart = sW1a + TC_MAXK * U;      // [TMR][TC_KPAD] half-1 partial sums
  float* sDl = sPart + TMR * TC_KPAD;     // [TMR][TC_KPAD] dL/dlogits
  float* sAct = sDl + TMR * TC_KPAD;      // [TMR][TC_KPAD] fresh action sample
  float* sG = sAct + TMR * TC_KPAD;       // [TMR][TC_KPAD] Gumbel noise
  long long* sRow = reinterpret_cast<long long*>(sG + TMR * TC_KPAD);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const AgentDev& me = C.agents[j];
  const MlpW pw = me.net[MDP_NET_P], qw = me.net[MDP_NET_Q];
  const MlpG& pg = me.grad[0];
  const long long row0 = (long long)blockIdx.x * TMR;
  const int nrows = (int)min((long long)TMR, B - row0);
  const int R = L.row_stride, K = me.act_dim;
  const int a_col0 = L.obs_sum + me.act_off;
  const int np = (pw.in + 31) / 32, nq = (qw.in + 31) / 32;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, LY::T_COLS);
  if (tid == 0) {
    for (int k = 0; k < NS; ++k) {
      mbar_init(&bars.stage_w[k], 1);
      mbar_init(&bars.stage_x[k], NTC / 64);
      mbar_init(&bars.stage_free[k], 1);
    }
    mbar_init(&bars.a_full, NTC / 32); mbar_init(&bars.r2_full, 1); mbar_init(&bars.r2_free, 1);
    mbar_init(&bars.acc, 1); mbar_init(&bars.dz1_full, NTC / 32);
    for (int k = 0; k < 2; ++k) { mbar_init(&bars.xs_full[k], NTC / 64); mbar_init(&bars.xs_free[k], 1); }
    for (int k = 0; k < 4; ++k) { mbar_init(&bars.dw_full[k], 1); mbar_init(&bars.dw_free[k], NTC / 64); }
    if (blockIdx.x == 0) C.adam_t[2 * j + 0] += 1;
  }
  if (tid < TMR) {
    const long long rl = row0 + min(tid, nrows - 1);
    sRow[tid] = (ridx ? ridx[rl] : rl) * R;
  }
  for (int i = tid; i < U; i += NTT) {
    sB1p[i] = pw.b1[i]; sB2p[i] = pw.b2[i];
    sB1q[i] = qw.b1[i]; sB2q[i] = qw.b2[i]; sW3q[i] = qw.W3[i];
  }
  for (int i = tid; i < U * K; i += NTT) sW3p[i] = pw.W3[i];
  for (int i = tid; i < K * U; i += NTT) sW1a[i] = qw.W1[(size_t)(a_col0 + i / U) * U + (i % U)];
  if (tid < K) sB3p[tid] = pw.b3[tid];
  if (tid == 0) sB3q[0] = qw.b3[0];
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == NTC / 32) {
    // ===== TMA producer =====
    if (lane == 0) {
      const NetImg ip = imgs[j].net[MDP_NET_P], iq = imgs[j].net[MDP_NET_Q];
      auto load_r2 = [&](const unsigned char* src) {
        mbar_arrive_expect_tx(&bars.r2_full, 2 * LY::W2_IMG);
        bulk_g2s(smem + LA::R2, src, 2 * LY::W2_IMG, &bars.r2_full);
      };
      auto chunks = [&](const unsigned char* w1, int n, int cbase) {
        for (int k = 0; k < n; ++k) {
          const int c = cbase + k, s = c % NS;
          if (c >= NS) mbar_wait_bounded(&bars.stage_free[s], ((c / NS) - 1) & 1);
          mbar_arrive_expect_tx(&bars.stage_w[s], 2 * LY::W_IMG);
          bulk_g2s(smem + LA::R1 + s * (2 * LY::W_IMG), w1 + (size_t)k * (2 * LY::W_IMG), 2 * LY::W_IMG, &bars.stage_w[s]);
        }
      };
      load_r2(ip.w2);
      chunks(ip.w1, np, 0);
      mbar_wait_bounded(&bars.r2_free, 0);
      load_r2(iq.w2);
      chunks(iq.w1, nq, np);
      mbar_wait_bounded(&bars.r2_free, 1);
      load_r2(iq.w2n);
      mbar_wait_bounded(&bars.r2_free, 0);
      load_r2(ip.w2n);
    }
  } else if (warp == NTC / 32 + 1) {
    // ===== MMA issuer (all lanes run the loops; one elected lane issues) =====
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    const uint64_t w_hi0 = umma::desc_k(sbase + LA::R1, LY::W_IMG, 0), w_lo0 = umma::desc_k(sbase + LA::R1 + LY::W_IMG, LY::W_IMG, 0);
    const uint64_t r2_hi = umma::desc_k(sbase + LA::R2, LY::W_IMG, 0), r2_lo = umma::desc_k(sbase + LA::R2 + LY::W2_IMG, LY::W_IMG, 0);
    auto layer1 = [&](int n, int cbase) {
      for (int k = 0; k < n; ++k) {
        const int c = cbase + k, s = c % NS, ph = (c / NS) & 1;
        mbar_wait_bounded(&bars.stage_x[s], ph);
        mbar_wait_bounded(&bars.stage_w[s], ph);
        umma::fence_after();
        const uint32_t so = s * ((2 * LY::W_IMG) >> 4);
        if (umma::elect_one()) {
          issue_3x<U, 4>(tb + LY::T_ACC1, tb + LY::T_X + s * 64, 32, w_hi0 + so, w_lo0 + so, k > 0 ? 1u : 0u);
          umma::commit(&bars.stage_free[s]);
          if (k == n - 1) umma::commit(&bars.acc);
        }
        __syncwarp();
      }
    };
    auto from_h1 = [&](int a_phase, int r2_phase, bool frees_r2) {  // acc2 = (TMEM H1 region) x (R2 images)^T
      mbar_wait_bounded(&bars.a_full, a_phase);
      mbar_wait_bounded(&bars.r2_full, r2_phase);
      umma::fence_after();
      if (umma::elect_one()) {
        issue_3x<U, U / 8>(tb + LY::T_ACC2, tb + LY::T_H1, U, r2_hi, r2_lo, 0u);
        umma::commit(&bars.acc);
        if (frees_r2) umma::commit(&bars.r2_free);
      }
      __syncwarp();
    };
    layer1(np, 0);            // acc phase 0: z1p
    from_h1(0, 0, true);      // acc phase 1: z2p = h1p W2p
    layer1(nq, np);           // acc phase 2: z1q
    from_h1(1, 1, true);      // acc phase 3: z2q = h1q W2q
    from_h1(0, 0, true);      // acc phase 4: dh1q = dz2q W2q^T
    mbar_wait_bounded(&bars.a_full, 1);  // dz2p: TMEM operand + MN-major images
    mbar_wait_bounded(&bars.r2_full, 1);
    umma::fence_after();
    if (umma::elect_one()) {
      issue_3x_mn<U, U, TMR / 8>(tb + LY::T_ACC1, sbase + LA::S1, LA::ACT_IMG, TMR * 128, sbase + LA::R1, LA::ACT_IMG, TMR * 128);
      issue_3x<U, U / 8>(tb + LY::T_ACC2, tb + LY::T_H1, U, r2_hi, r2_lo, 0u);  // dh1p = dz2p W2p^T
      umma::commit(&bars.acc);  // acc phase 5
    }
    __syncwarp();
    mbar_wait_bounded(&bars.dz1_full, 0);
    for (int c = 0; c < np; ++c) {  // dW1p^T chunks
      const int s = c & 1, t = c & 3;
      mbar_wait_bounded(&bars.xs_full[s], (c >> 1) & 1);
      if (c >= 4) mbar_wait_bounded(&bars.dw_free[t], ((c >> 2) - 1) & 1);
      umma::fence_after();
      if (umma::elect_one()) {
        issue_3x_mn<U, 32, TMR / 8>(tb + LA::T_DW + t * 32, sbase + LA::S1, LA::ACT_IMG, TMR * 128,
                                   sbase + LA::R1 + s * (2 * LA::XS_IMG), LA::XS_IMG, TMR * 128);
        umma::commit(&bars.xs_free[s]);
        umma::commit(&bars.dw_full[t]);
      }
      __syncwarp();
    }
  } else {
    // ===== compute warps: thread = (batch row, unit half) =====
    const int row = 32 * (warp & 3) + lane, half = warp >> 2, c0 = 32 * half;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const long long rowoff = sRow[row];
    const bool valid = row < nrows;
    uint32_t chunk = 0;  // running layer-1 chunk counter (actor pass, then critic pass)
    auto layer1 = [&](const XT& xs, int n) {
      l1_stage_loop(bars, tbase + LY::T_X, xs, rowoff, n, chunk, warp, lane);
      chunk += (uint32_t)n;
    };
    float v[32];
    uint32_t mask1p = 0, mask2p = 0, mask1q = 0, mask2q = 0;
    // ---- actor forward on o_j
    const XT xp{batch + me.obs_off, me.obs_dim, nullptr, 0, 0, 0, (me.obs_off & 3) == 0};
    layer1(xp, np);
    for (int a = half; a < K; a += 2) {  // Gumbel noise of this row, hidden behind the layer-1 MMAs
      const float u = u_actor ? u_actor[(row0 + min(row, nrows - 1)) * u_stride + me.act_off + a]
                              : philox_u(seed, counter, (uint32_t)(0x200 + j), row0 + row, a);
      sG[row * TC_KPAD + a] = gumbel_from_u(u);
    }
    mbar_wait_bounded(&bars.acc, 0);
    umma::fence_after();
    umma::tmem_ld32(tbase + LY::T_ACC1 + lane_base + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = fmaxf(v[i] + sB1p[c0 + i], 0.f);
      mask1p |= (v[i] > 0.f ? 1u : 0u) << i;
    }
    store_act_tmem<U>(tbase + LY::T_H1, lane_base, c0, v);
    store_act_mn(smem + LA::S1, LA::ACT_IMG, row, c0, v);  // h1p: A operand of dW2p
    warp_arrive_both(&bars.a_full, lane);                  // a_full phase 0
    mbar_wait_bounded(&bars.acc, 1);
    umma::fence_after();
    float h2p[32];
    umma::tmem_ld32(tbase + LY::T_ACC2 + lane_base + (uint32_t)c0, h2p);
    float part[TC_MAXK];
#pragma unroll
    for (int a = 0; a < TC_MAXK; ++a) part[a] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      h2p[i] = fmaxf(h2p[i] + sB2p[c0 + i], 0.f);
      mask2p |= (h2p[i] > 0.f ? 1u : 0u) << i;
#pragma unroll
      for (int a = 0; a < TC_MAXK; ++a)
        if (a < K) part[a] = fmaf(h2p[i], sW3p[(c0 + i) * K + a], part[a]);
    }
    umma::fence_before();
    if (half == 1)
      for (int a = 0; a < K; ++a) sPart[row * TC_KPAD + a] = part[a];
    named_sync();
    float logit[TC_MAXK], act[TC_MAXK];
    if (half == 0) {  // logits -> fresh Gumbel-softmax sample (maddpg.py:49), per head
      double sl = 0.0;
#pragma unroll
      for (int a = 0; a < TC_MAXK; ++a) {
        logit[a] = 0.f; act[a] = 0.f;
        if (a < K) {
          logit[a] = part[a] + sPart[row * TC_KPAD + a] + sB3p[a];
          act[a] = logit[a] + sG[row * TC_KPAD + a];
          if (valid) sl += (double)logit[a] * (double)logit[a];
        }
      }
      for (int h = 0; h < me.n_heads; ++h) {
        const int o = h ? me.head_dim[0] : 0, n = me.head_dim[h];
        float m = -INFINITY, ssum = 0.f;
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a >= o && a < o + n) m = fmaxf(m, act[a]);
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a >= o && a < o + n) { act[a] = expf(act[a] - m); ssum += act[a]; }
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a >= o && a < o + n) act[a] = act[a] / ssum;
      }
#pragma unroll
      for (int a = 0; a < TC_MAXK; ++a)
        if (a < K) sAct[row * TC_KPAD + a] = act[a];
      for (int o = 16; o > 0; o >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, o);
      if (lane == 0) atomicAdd(C.stats + 8 * j + 2, sl);
    }
    named_sync();  // the sample is visible to the critic's gather
    // ---- running critic on [o, a_-j, a_hat_j]
    const XT xq{batch, L.x_dim, sAct, TC_KPAD, a_col0, K, 1};
    layer1(xq, nq);
    mbar_wait_bounded(&bars.acc, 0);
    umma::fence_after();
    umma::tmem_ld32(tbase + LY::T_ACC1 + lane_base + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = fmaxf(v[i] + sB1q[c0 + i], 0.f);
      mask1q |= (v[i] > 0.f ? 1u : 0u) << i;
    }
    store_act_tmem<U>(tbase + LY::T_H1, lane_base, c0, v);
    warp_arrive_tmem(&bars.a_full, lane);  // a_full phase 1
    mbar_wait_bounded(&bars.acc, 1);
    umma::fence_after();
    umma::tmem_ld32(tbase + LY::T_ACC2 + lane_base + (uint32_t)c0, v);
    float qpart = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float h = fmaxf(v[i] + sB2q[c0 + i], 0.f);
      mask2q |= (h > 0.f ? 1u : 0u) << i;
      qpart = fmaf(h, sW3q[c0 + i], qpart);
    }
    if (half == 1) sPart[row * TC_KPAD] = qpart;
    // dz2q = (-1/B) W3q^T relu'(h2q) for valid rows  (the actor loss is -mean(q))
    const float dq = valid ? -1.0f / (float)B : 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = ((mask2q >> i) & 1u) ? dq * sW3q[c0 + i] : 0.f;
    store_act_tmem<U>(tbase + LY::T_H1, lane_base, c0, v);
    warp_arrive_tmem(&bars.a_full, lane);  // a_full phase 2
    named_sync();
    if (half == 0) {
      double sq = valid ? -(double)(qpart + sPart[row * TC_KPAD] + sB3q[0]) : 0.0;
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) atomicAdd(C.stats + 8 * j + 1, sq);
    }
    // ---- critic backward to agent j's action columns
    mbar_wait_bounded(&bars.acc, 0);
    umma::fence_after();
    umma::tmem_ld32(tbase + LY::T_ACC2 + lane_base + (uint32_t)c0, v);
    umma::fence_before();
#pragma unroll
    for (int a = 0; a < TC_MAXK; ++a) part[a] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float dz = ((mask1q >> i) & 1u) ? v[i] : 0.f;
#pragma unroll
      for (int a = 0; a < TC_MAXK; ++a)
        if (a < K) part[a] = fmaf(dz, sW1a[a * U + c0 + i], part[a]);  // dQ/da = dz1q . W1q[a_col0 + a, :]
    }
    named_sync();  // the q partials in sPart have been consumed
    if (half == 1)
      for (int a = 0; a < K; ++a) sPart[row * TC_KPAD + a] = part[a];
    named_sync();
    if (half == 0) {  // softmax Jacobian per head + logit regulariser (maddpg.py:55-58)
      const float regc = (float)(2.0 * C.actor_reg / ((double)B * (double)K));
      float dqa[TC_MAXK];
#pragma unroll
      for (int a = 0; a < TC_MAXK; ++a) dqa[a] = a < K ? part[a] + sPart[row * TC_KPAD + a] : 0.f;
      for (int h = 0; h < me.n_heads; ++h) {
        const int o = h ? me.head_dim[0] : 0, n = me.head_dim[h];
        float dot = 0.f;
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a >= o && a < o + n) dot = fmaf(act[a], dqa[a], dot);
#pragma unroll
        for (int a = 0; a < TC_MAXK; ++a)
          if (a >= o && a < o + n) {
            const float dl = act[a] * (dqa[a] - dot) + regc * logit[a];
            sDl[row * TC_KPAD + a] = valid ? dl : 0.f;
          }
      }
    }
    named_sync();
    float dl[TC_MAXK];
#pragma unroll
    for (int a = 0; a < TC_MAXK; ++a) dl[a] = a < K ? sDl[row * TC_KPAD + a] : 0.f;
    // ---- actor head backward: gW3p, gb3p, dz2p
    for (int a = 0; a < K; ++a) {
      float t[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = h2p[i] * dl[a];
      const float sum = warp_colsum32(t, lane);
      red_add(pg.W3 + (size_t)(c0 + lane) * K + a, sum);
    }
    if (half == 0)
      for (int a = 0; a < K; ++a) {
        float sb = dl[a];
        for (int o = 16; o > 0; o >>= 1) sb += __shfl_xor_sync(0xffffffffu, sb, o);
        if (lane == 0) red_add(pg.b3 + a, sb);
      }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float sacc = 0.f;
#pragma unroll
      for (int a = 0; a < TC_MAXK; ++a)
        if (a < K) sacc = fmaf(dl[a], sW3p[(c0 + i) * K + a], sacc);
      v[i] = ((mask2p >> i) & 1u) ? sacc : 0.f;  // dz2p
    }
    store_act_tmem<U>(tbase + LY::T_H1, lane_base, c0, v);   // A operand of dh1p
    store_act_mn(smem + LA::R1, LA::ACT_IMG, row, c0, v);    // B operand of dW2p (the W1^T slots are idle)
    warp_arrive_both(&bars.a_full, lane);                    // a_full phase 3
    {
      const float sum = warp_colsum32(v, lane);
      red_add(pg.b2 + c0 + lane, sum);
    }
    // ---- dW2p tile out, dz1p images in
    mbar_wait_bounded(&bars.acc, 1);
    umma::fence_after();
    umma::tmem_ld32(tbase + LY::T_ACC1 + lane_base + (uint32_t)c0, v);
    if (lane < 16) {
      float* dst = pg.W2 + (size_t)(16 * (warp & 3) + lane) * U + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 4) red_add4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    umma::tmem_ld32(tbase + LY::T_ACC2 + lane_base + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = ((mask1p >> i) & 1u) ? v[i] : 0.f;
    umma::fence_before();
    store_act_mn(smem + LA::S1, LA::ACT_IMG, row, c0, v);  // dz1p over the dead h1p images
    warp_arrive_both(&bars.dz1_full, lane);
    {
      const float sum = warp_colsum32(v, lane);
      red_add(pg.b1 + c0 + lane, sum);
    }
    // ---- dW1p^T chunks: stager warps / reader warps
    dw1_chunk_loop<U>(bars, smem + LA::R1, LA::XS_IMG, tbase + LA::T_DW, xp, rowoff, np, pw.in, pg.W1, warp, lane);
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tbase, LY::T_COLS);
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
      off2[i * 4 + k] = total; total += 4 * LY::W2_IMG;  // W2^T pair + W2 pair
    }
  unsigned char* arena = nullptr;
  MDP_CUDA(cudaMalloc(&arena, total + n * sizeof(tc::AgentImg)));
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) {
      h[i].net[k].w1 = arena + off1[i * 4 + k];
      h[i].net[k].w2 = arena + off2[i * 4 + k];
      h[i].net[k].w2n = arena + off2[i * 4 + k] + 2 * LY::W2_IMG;
    }
  MDP_CUDA(cudaMemcpy(arena + total, h.data(), n * sizeof(tc::AgentImg), cudaMemcpyHostToDevice));
  c->tc_arena = arena;
  c->tc_imgs = arena + total;
  return MDP_OK;
}

// mdp_actor_act on tensor cores (num_units 64).  MDP_ENOTSUP -> the caller uses k_actor_act.
int launch_actor_act_tc(mdp_core* c, const CoreDev& d, int32_t agent_begin, int32_t agent_count, int32_t use_target, int32_t E,
                        const float* obs, int32_t obs_stride, float* act, int32_t act_stride, const float* u, uint64_t seed,
                        uint64_t counter, float* logits_out, long long row_base, cudaStream_t st) {
  if (c->cfg.num_units != 64 || !tc_heads_ok(c)) return MDP_ENOTSUP;
  constexpr int U = 64;
  using LY = tc::Lay<U>;
  int rc = ensure_images(c);
  if (rc) return rc;
  const tc::AgentImg* imgs = reinterpret_cast<const tc::AgentImg*>(c->tc_imgs);
  const int net = use_target ? MDP_NET_TARGET_P : MDP_NET_P;
  int max_in = 0;
  for (int k = agent_begin; k < agent_begin + agent_count; ++k) max_in = std::max(max_in, c->cfg.obs_dim[k]);
  const int bx = std::min(64, cdiv(((max_in + 31) / 32 * 32 + U) * U, 256 * 4));
  tc::k_build_images<U><<<dim3(bx, agent_count), 256, 0, st>>>(d, imgs, agent_count, agent_begin, 0, net, MDP_NET_Q);
  rc = check_launch("k_build_images");
  if (rc) return rc;
  // a CTA owns all 512 TMEM columns of its SM: ask for more than half of the shared memory so that two never share one
  const size_t smem = std::max(LY::OFF_MISC + (size_t)LY::MISC_FLOATS * 4 + 1024 + 64, (size_t)116 * 1024);
  auto kern = tc::k_actor_act_tc<U>;
  MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(cdiv(E, tc::TMR), agent_count), tc::NTT, smem, st>>>(d, imgs, agent_begin, net, E, obs, obs_stride, act, act_stride, u,
                                                                  seed, counter, logits_out, row_base);
  return check_launch("k_actor_act_tc");
}


// Returns MDP_ENOTSUP when the shape is outside the tensor-core path (the caller then uses the SIMT kernels).
int launch_td_target_tc(mdp_core* c, const CoreDev& d, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B,
                        const float* batch, const long long* ridx, long long idx_stride, const float* u_target, int32_t u_stride,
                        uint64_t seed, uint64_t counter, float* y_out, long long y_stride, float* target_act_out, cudaStream_t st) {
  if (c->cfg.num_units != 64) return fail(MDP_ENOTSUP, "tensor-core path: num_units %d (64 only)", c->cfg.num_units);
  if (!tc_heads_ok(c)) return fail(MDP_ENOTSUP, "tensor-core path: an action block wider than %d columns", TC_MAXK);
  constexpr int U = 64;
  using LY = tc::Lay<U>;
  int rc = ensure_images(c);
  if (rc) return rc;
  const tc::AgentImg* imgs = reinterpret_cast<const tc::AgentImg*>(c->tc_imgs);
  const int n = c->cfg.n_agents, ASP = c->act_stride | 1;
  const size_t fixed = LY::OFF_MISC + (size_t)LY::MISC_FLOATS * 4 + 1024 + 64;
  const size_t limit = 227 * 1024 - 512;  // static __shared__ (barriers, TMEM slot) comes on top
  const size_t act_bytes = (size_t)tc::TMR * ASP * 4;
  // split launch (actor groups, then critics) once there are enough actor passes per tile to be worth a second launch
  bool any_local = false;
  for (int k = agent; k < agent + count; ++k) any_local |= c->cfg.local_q[k] != 0;
  const int groups = (!any_local && n >= 8) ? std::min(8, (n + 2) / 3) : 1;
  const int act_in_smem = groups == 1 && fixed + act_bytes <= limit;
  // a CTA owns all 512 TMEM columns of its SM: ask for more than half of the shared memory so that two never share one
  const size_t smem = std::max(fixed + (act_in_smem ? act_bytes : 0), (size_t)116 * 1024);
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
  if (groups > 1) {
    kern<<<dim3(tiles, count, groups), tc::NTT, smem, st>>>(d, imgs, agent, *lay, B, batch, ridx, u_target, u_stride, seed, counter,
                                                             y_out, target_act_out, idx_stride, y_stride, c->tc_scratch, 0, 1);
    rc = check_launch("k_td_target_tc (actors)");
    if (rc) return rc;
    kern<<<dim3(tiles, count), tc::NTT, smem, st>>>(d, imgs, agent, *lay, B, batch, ridx, u_target, u_stride, seed, counter, y_out,
                                                     target_act_out, idx_stride, y_stride, c->tc_scratch, 0, 2);
    return check_launch("k_td_target_tc (critic)");
  }
  kern<<<dim3(tiles, count), tc::NTT, smem, st>>>(d, imgs, agent, *lay, B, batch, ridx, u_target, u_stride, seed, counter, y_out,
                                                   target_act_out, idx_stride, y_stride, c->tc_scratch, act_in_smem, 0);
  return check_launch("k_td_target_tc");
}


// Returns MDP_ENOTSUP when the shape is outside the tensor-core path (local critics, num_units != 64).
int launch_critic_grads_tc(mdp_core* c, const CoreDev& d, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B,
                           const float* batch, const long long* ridx, long long idx_stride, const float* y, long long y_stride,
                           float* q_out, cudaStream_t st) {
  if (c->cfg.num_units != 64) return fail(MDP_ENOTSUP, "tensor-core path: num_units %d (64 only)", c->cfg.num_units);
  if (!tc_heads_ok(c)) return fail(MDP_ENOTSUP, "tensor-core path: an action block wider than %d columns", TC_MAXK);
  for (int k = agent; k < agent + count; ++k)
    if (c->cfg.local_q[k]) return fail(MDP_ENOTSUP, "tensor-core critic path: local critics use the SIMT kernels");
  if ((lay->row_stride & 3) != 0) return fail(MDP_ENOTSUP, "tensor-core path: row stride");
  constexpr int U = 64;
  using LB = tc::LayB<U>;
  int rc = ensure_images(c);
  if (rc) return rc;
  const tc::AgentImg* imgs = reinterpret_cast<const tc::AgentImg*>(c->tc_imgs);
  int max_in = 0;
  for (int k = agent; k < agent + count; ++k) max_in = std::max(max_in, c->lay.net_in[k][MDP_NET_Q]);
  const int bx = std::min(64, cdiv(((max_in + 31) / 32 * 32 + U) * U, 256 * 4));
  tc::k_build_images<U><<<dim3(bx, count), 256, 0, st>>>(d, imgs, 0, 0, agent, MDP_NET_P, MDP_NET_Q);
  rc = check_launch("k_build_images");
  if (rc) return rc;
  const size_t smem = LB::OFF_MISC + (size_t)LB::MISC_FLOATS * 4 + 1024 + 64;
  auto kern = tc::k_critic_grads_tc<U>;
  MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = cdiv(B, tc::TMR), nchunks = (max_in + 31) / 32;
  // wide critics: the dW1 phase (as long as the forward pass, but embarrassingly parallel over K) moves to its own launch
  const int ksplit = nchunks >= 16 ? std::min(8, nchunks / 8) : 1;
  float* dz1 = nullptr;
  if (ksplit > 1) {
    const size_t need = (size_t)tiles * count * tc::TMR * U * sizeof(float);
    if (need > c->tc_dz1_bytes) {
      if (c->tc_dz1) cudaFree(c->tc_dz1);
      c->tc_dz1 = nullptr;
      c->tc_dz1_bytes = 0;
      MDP_CUDA(cudaMalloc(&c->tc_dz1, need));
      c->tc_dz1_bytes = need;
    }
    dz1 = c->tc_dz1;
  }
  kern<<<dim3(tiles, count), tc::NTT, smem, st>>>(d, imgs, agent, *lay, B, batch, ridx, y, q_out, idx_stride, y_stride, dz1);
  rc = check_launch("k_critic_grads_tc");
  if (rc || ksplit == 1) return rc;
  auto kd = tc::k_dw1_tc<U>;
  const size_t smem_d = 2 * (size_t)tc::TMR * 128 * (U / 32) + 4 * (size_t)tc::TMR * 128 + tc::TMR * 8 + 1024 + 64;
  MDP_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
  kd<<<dim3(tiles, count, ksplit), tc::NTD, smem_d, st>>>(d, agent, *lay, B, batch, ridx, idx_stride, dz1, cdiv(nchunks, ksplit));
  return check_launch("k_dw1_tc");
}

int launch_actor_grads_tc(mdp_core* c, const CoreDev& d, int32_t agent, int32_t count, const mdp_ring_layout* lay, int32_t B,
                          const float* batch, const long long* ridx, long long idx_stride, const float* u_actor, int32_t u_stride,
                          uint64_t seed, uint64_t counter, cudaStream_t st) {
  if (c->cfg.num_units != 64) return fail(MDP_ENOTSUP, "tensor-core path: num_units %d (64 only)", c->cfg.num_units);
  if (!tc_heads_ok(c)) return fail(MDP_ENOTSUP, "tensor-core path: an action block wider than %d columns", TC_MAXK);
  for (int k = agent; k < agent + count; ++k)
    if (c->cfg.local_q[k]) return fail(MDP_ENOTSUP, "tensor-core actor path: local critics use the SIMT kernels");
  constexpr int U = 64;
  using LA = tc::LayA<U>;
  int rc = ensure_images(c);
  if (rc) return rc;
  const tc::AgentImg* imgs = reinterpret_cast<const tc::AgentImg*>(c->tc_imgs);
  int max_in = 0;
  for (int k = agent; k < agent + count; ++k) max_in = std::max(max_in, c->lay.net_in[k][MDP_NET_Q]);
  const int bx = std::min(64, cdiv(((max_in + 31) / 32 * 32 + U) * U, 256 * 4));
  // images of the running actor and the running critic of every agent in the slice
  tc::k_build_images<U><<<dim3(bx, 2 * count), 256, 0, st>>>(d, imgs, count, agent, agent, MDP_NET_P, MDP_NET_Q);
  rc = check_launch("k_build_images");
  if (rc) return rc;
  const size_t smem = LA::OFF_MISC + (size_t)LA::MISC_FLOATS * 4 + 1024 + 64;
  auto kern = tc::k_actor_grads_tc<U>;
  MDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(cdiv(B, tc::TMR), count), tc::NTT, smem, st>>>(d, imgs, agent, *lay, B, batch, ridx, u_actor, u_stride, seed, counter,
                                                             idx_stride);
  return check_launch("k_actor_grads_tc");
}

}  // namespace mdp
