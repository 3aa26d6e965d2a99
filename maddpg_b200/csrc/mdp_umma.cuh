// tcgen05 / TMEM primitives for the tensor-core MLP kernels (sm_100a inline PTX; no CUTLASS dependency).
//
// Operand images.  Every UMMA operand lives in shared memory as a stack of PANELS of [rows][32 floats]
// (128-byte rows), panels rows*128 bytes apart.
//   K-major  view (MN index = image row, K index = image column, panel = K / 32): SWIZZLE_128B, eight rows
//       form one 1024-byte atom, 16-byte chunk index XOR (row & 7).
//   MN-major view (MN index = image column, panel = MN / 32, K index = image row): 32-bit operands only accept
//       SWIZZLE_128B_BASE32B, four rows form one 512-byte atom, 32-byte chunk index XOR (row & 3).
// (The two swizzles differ, so one fp32 image cannot serve both views; tools/umma_probe.cu checks both on a B200.)
// Bit layouts follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor" tables.
#pragma once
#include "mdp_common.cuh"

namespace mdp {
namespace umma {

constexpr int PANEL_COLS = 32;  // floats per 128-byte image row

// byte offset of element (r, c), c < 32, inside one panel
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2)));
}

// byte offset of element (r, c), c < 32, inside one MN-major (SWIZZLE_128B_BASE32B) panel
__device__ __forceinline__ uint32_t sw128b32_off(int r, int c) {
  return (uint32_t)(((r >> 2) << 9) + ((r & 3) << 7) + ((((c >> 3) ^ (r & 3)) << 5) | ((c & 7) << 2)));
}

// split for 3xTF32: hi = x rounded to NEAREST tf32 (13 low mantissa bits zero: exactly what kind::tf32 reads), lo = x - hi
// (exact, signed, |lo| <= 2^-11 |x|; the MMA truncates lo toward zero: a sign-symmetric 2^-21 |x| error).  Round 1 cleared the
// low bits instead (truncation): lo was one-sided and twice as large, so the dropped lo*lo term and the truncation of lo were
// biased errors that add up along K -- measured 4x the error of this split on the rollout's actions against the oracle.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  hi = __uint_as_float(r);
  lo = x - hi;
}

// shared-memory matrix descriptor, SWIZZLE_128B, descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2u) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // version
  d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// K-major view, k-step s (8 tf32 = 32 bytes) of an image whose panels are panel_bytes apart
__device__ __forceinline__ uint64_t desc_k(uint32_t img, uint32_t panel_bytes, int s) {
  return smem_desc(img + (uint32_t)(s >> 2) * panel_bytes + (uint32_t)(s & 3) * 32u, 16u, 1024u);
}
// MN-major view, k-step s (8 image rows = two 512-byte atoms); MN groups of 32 are panel_bytes apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t img, uint32_t panel_bytes, int s) {
  return smem_desc(img + (uint32_t)s * 1024u, panel_bytes, 512u, 1u);
}

// instruction descriptor for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in TMEM (lane = row, one 32-bit column per K element; 8 columns per kind::tf32 MMA)
__device__ __forceinline__ void mma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (always the same one).  The MMA issuer warps run their loops with ALL lanes and put only
// the tcgen05.mma / tcgen05.commit under this predicate: the descriptors are then warp-uniform values that ptxas keeps in
// uniform registers (0-1 SASS instructions between MMAs), whereas an `if (lane == 0)` region makes them per-thread values that
// need an ELECT / R2UR sequence per operand (5-15 instructions per MMA -- the issuing thread was the bottleneck of the pipeline).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// arrives on `bar` once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one full warp allocates / frees `cols` (power of two >= 32) TMEM columns; the base address lands in *slot (smem)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// TMEM -> registers: lane (32 * (warp % 4) + laneid), 32 / 16 consecutive fp32 columns starting at taddr's column
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: lane (32 * (warp % 4) + laneid), 16 consecutive 32-bit columns starting at taddr's column
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace umma
}  // namespace mdp
