// Shared host/device helpers for libmaddpg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/maddpg_b200.h"

namespace mdp {

// ---------------------------------------------------------------------------------------------
// error plumbing: int status across the ABI, message in a thread-local buffer (header contract)
// ---------------------------------------------------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MDP_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return MDP_OK;
}

#define MDP_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) return mdp::fail(MDP_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

#define MDP_REQUIRE(cond, ...)                                    \
  do {                                                            \
    if (!(cond)) return mdp::fail(MDP_EINVAL, __VA_ARGS__);       \
  } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011).  key = (seed lo, seed hi); counter =
// (c0, c1, c2, c3).  Used for on-device reset draws and Gumbel noise (the reference's
// TF / numpy / python RNG streams cannot be reproduced; parity tests inject the draws instead).
// ---------------------------------------------------------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  __host__ __device__ static inline uint4 gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(M0, c0, hi0, lo0);
      mulhilo(M1, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  // U[0,1) with 24 random bits, like tf.random_uniform(float32) / numpy float32 draws
  __host__ __device__ static inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
  // U[0,1) with 53 random bits
  __host__ __device__ static inline double u01d(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
  }
};

// mdp_replay_insert with an optional device control block (see mdp_env_set_ctl)
int env_step_range(mdp_env* env, int32_t E, int32_t e_base, int32_t n, void* state, const float* act, float* obs_out,
                   float* rew_out, uint8_t* done_out, const float* obs_prev, float* ring, int64_t ring_capacity,
                   int32_t ring_row_stride, int64_t ring_cursor, void* stream);
// mdp_actor_act for rows [row_base, row_base + E) of a larger population (obs / act point at row row_base): the Philox
// streams are keyed by the population row, so a chunked launch draws the same numbers as one launch over all rows
int actor_act_range(mdp_core* c, int32_t agent_begin, int32_t agent_count, int32_t use_target, int32_t E, const float* obs,
                    int32_t obs_stride, float* act, int32_t act_stride, const float* u, uint64_t seed, uint64_t counter,
                    float* logits_out, int64_t row_base, void* stream);
int replay_insert_ctl(const mdp_ring_layout* lay, float* ring, int64_t capacity, int64_t cursor, int32_t E, int32_t agent,
                      const float* obs, int32_t obs_stride, const float* act, int32_t act_stride, const float* rew,
                      int32_t rew_stride, const float* next_obs, int32_t next_obs_stride, const uint8_t* done,
                      int32_t done_stride, const unsigned long long* ctl, void* stream);

// ---------------------------------------------------------------------------------------------
// TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP) + mbarrier helpers.  Sizes and both addresses must be
// multiples of 16 bytes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// Fire-and-forget fp32 reductions into the gradient bucket.  `atomicAdd(float*)` on a pointer the compiler cannot prove to be
// global (the gradient pointers come out of a device-side table) lowers to a RETURNING generic atomic plus shared/local CAS-spin
// fallbacks (ATOM.E.ADD + ATOM.CAST.SPIN); the explicit .global RED forms are one instruction, non-blocking, and vectorisable.
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {  // p 16-byte aligned
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Programmatic dependent launch (griddepcontrol): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream still runs; everything it reads that the predecessor writes must come after
// pdl_wait() (which returns once the predecessor grid has completed and flushed).  pdl_launch_dependents() in the predecessor
// lets the dependent grid become resident early.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Gumbel-softmax noise term -log(-log(u)) in float32 (distributions.py:264-266)
__device__ __forceinline__ float gumbel_from_u(float u) { return -logf(-logf(u)); }

}  // namespace mdp
