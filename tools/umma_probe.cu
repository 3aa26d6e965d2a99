// Stand-alone check of the tcgen05 building blocks in maddpg_b200/csrc/mdp_umma.cuh on a real B200:
// SWIZZLE_128B operand images in both views (K-major / MN-major), kind::tf32 MMAs with M = 64 / 128,
// the 3xTF32 split, the TMEM accumulator layout and tcgen05.ld.  D[M][N] = A[M][K] * B[N][K]^T.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <array>
#include <vector>

#include "../maddpg_b200/csrc/mdp_umma.cuh"

namespace mdp {
thread_local char g_err[512];
std::atomic<long long> g_launches{0};
}  // namespace mdp
using namespace mdp;

template <int SPLIT>
__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                               int M, int N, int K, int a_mn, int b_mn, int tmem_cols) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_bytes = (uint32_t)((M + 31) / 32 * 32) * K * 4, b_bytes = (uint32_t)((N + 31) / 32 * 32) * K * 4;
  unsigned char *Ahi = gbase, *Alo = Ahi + a_bytes, *Bhi = Alo + a_bytes, *Blo = Bhi + b_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, tmem_cols);
  if (tid == 0) mbar_init(&bar, 1);
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    float hi, lo;
    umma::split_tf32(A[i], hi, lo);
    const uint32_t off = a_mn ? (uint32_t)(m >> 5) * (K * 128) + umma::sw128b32_off(k, m & 31)
                              : (uint32_t)(k >> 5) * (M * 128) + umma::sw128_off(m, k & 31);
    *reinterpret_cast<float*>(Ahi + off) = hi;
    *reinterpret_cast<float*>(Alo + off) = lo;
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    float hi, lo;
    umma::split_tf32(B[i], hi, lo);
    const uint32_t off = b_mn ? (uint32_t)(n >> 5) * (K * 128) + umma::sw128b32_off(k, n & 31)
                              : (uint32_t)(k >> 5) * (N * 128) + umma::sw128_off(n, k & 31);
    *reinterpret_cast<float*>(Bhi + off) = hi;
    *reinterpret_cast<float*>(Blo + off) = lo;
  }
  umma::fence_async_smem();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma::idesc_tf32(M, N, a_mn, b_mn);
    const uint32_t sAhi = smem_u32(Ahi), sAlo = smem_u32(Alo), sBhi = smem_u32(Bhi), sBlo = smem_u32(Blo);
    uint32_t acc = 0;
    for (int s = 0; s < K / 8; ++s) {
      auto da = [&](uint32_t img) { return a_mn ? umma::desc_mn(img, K * 128, s) : umma::desc_k(img, M * 128, s); };
      auto db = [&](uint32_t img) { return b_mn ? umma::desc_mn(img, K * 128, s) : umma::desc_k(img, N * 128, s); };
      if (SPLIT == 3) {
        umma::mma_tf32(tbase, da(sAlo), db(sBhi), idesc, acc); acc = 1;
        umma::mma_tf32(tbase, da(sAhi), db(sBlo), idesc, acc);
      }
      umma::mma_tf32(tbase, da(sAhi), db(sBhi), idesc, acc); acc = 1;
    }
    umma::commit(&bar);
  }
  {  // bounded wait: a wrong descriptor must not hang the box
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      if (!done && clock64() - t0 > 2000000000ll) __trap();
    }
  }
  umma::fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * N + c0 + i] = v[i];
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tbase, tmem_cols);
}

// A operand from TMEM: thread (warp w, lane l) writes row 32 w + l of A (hi and lo) with tcgen05.st, 16 columns at a time
template <int SPLIT>
__global__ void __launch_bounds__(128) k_probe_ta(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                  int N, int K, int b_mn, int tmem_cols) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t b_bytes = (uint32_t)((N + 31) / 32 * 32) * K * 4;
  unsigned char *Bhi = gbase, *Blo = Bhi + b_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, tmem_cols);
  if (tid == 0) mbar_init(&bar, 1);
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    float hi, lo;
    umma::split_tf32(B[i], hi, lo);
    const uint32_t off = b_mn ? (uint32_t)(n >> 5) * (K * 128) + umma::sw128b32_off(k, n & 31)
                              : (uint32_t)(k >> 5) * (N * 128) + umma::sw128_off(n, k & 31);
    *reinterpret_cast<float*>(Bhi + off) = hi;
    *reinterpret_cast<float*>(Blo + off) = lo;
  }
  umma::fence_async_smem();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t a_hi = tbase + N, a_lo = a_hi + K;  // A regions after the accumulator columns
  const int row = 32 * warp + lane;
  for (int k0 = 0; k0 < K; k0 += 16) {
    float hi[16], lo[16];
    for (int i = 0; i < 16; ++i) umma::split_tf32(A[row * K + k0 + i], hi[i], lo[i]);
    umma::tmem_st16(a_hi + ((uint32_t)(32 * warp) << 16) + k0, hi);
    umma::tmem_st16(a_lo + ((uint32_t)(32 * warp) << 16) + k0, lo);
  }
  umma::tmem_st_wait();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma::idesc_tf32(128, N, 0, b_mn);
    const uint32_t sBhi = smem_u32(Bhi), sBlo = smem_u32(Blo);
    uint32_t acc = 0;
    for (int s = 0; s < K / 8; ++s) {
      auto db = [&](uint32_t img) { return b_mn ? umma::desc_mn(img, K * 128, s) : umma::desc_k(img, N * 128, s); };
      if (SPLIT == 3) {
        umma::mma_tf32_ta(tbase, a_lo + 8 * s, db(sBhi), idesc, acc); acc = 1;
        umma::mma_tf32_ta(tbase, a_hi + 8 * s, db(sBlo), idesc, acc);
      }
      umma::mma_tf32_ta(tbase, a_hi + 8 * s, db(sBhi), idesc, acc); acc = 1;
    }
    umma::commit(&bar);
  }
  {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      if (!done && clock64() - t0 > 2000000000ll) __trap();
    }
  }
  umma::fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * N + c0 + i] = v[i];
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tbase, tmem_cols);
}

static float trunc_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

static int run(int M, int N, int K, int a_mn, int b_mn, int split) {
  std::vector<float> A((size_t)M * K), B((size_t)N * K), D(128 * (size_t)N, -777.f);
  srand(1234 + M + 3 * N + 7 * K + 11 * a_mn + 13 * b_mn);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * ((size_t)((M + 31) / 32 * 32) * K + (size_t)((N + 31) / 32 * 32) * K) * 4 + 1024;
  int cols = 32;
  while (cols < (a_mn == 2 ? N + 2 * K : N)) cols *= 2;
  if (a_mn == 2) {  // A from TMEM (M = 128 only)
    auto kern = split == 3 ? k_probe_ta<3> : k_probe_ta<1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<1, 128, smem>>>(dA, dB, dD, N, K, b_mn, cols);
  } else {
    auto kern = split == 3 ? k_probe<3> : k_probe<1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<1, 128, smem>>>(dA, dB, dD, M, N, K, a_mn, b_mn, cols);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("M=%d N=%d K=%d a_mn=%d b_mn=%d split=%d: CUDA error %s\n", M, N, K, a_mn, b_mn, split, cudaGetErrorString(e));
    return 1;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  std::vector<double> R((size_t)M * N);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) {
        const float a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
        s += split == 3 ? (double)a * (double)b : (double)trunc_tf32(a) * (double)trunc_tf32(b);
      }
      R[(size_t)m * N + n] = s;
    }
  // which TMEM lane holds row m?  (M = 128: lane m.  M = 64: reported below)
  std::vector<int> lane_of(M, -1);
  double worst = 0;
  for (int m = 0; m < M; ++m) {
    double best = 1e30;
    for (int l = 0; l < 128; ++l) {
      double err = 0;
      for (int n = 0; n < N; ++n) err = fmax(err, fabs((double)D[(size_t)l * N + n] - R[(size_t)m * N + n]));
      if (err < best) { best = err; lane_of[m] = l; }
    }
    worst = fmax(worst, best);
  }
  bool ident = true, m64 = true;
  for (int m = 0; m < M; ++m) {
    ident &= lane_of[m] == m;
    m64 &= lane_of[m] == (m / 16) * 32 + m % 16;
  }
  const double tol = split == 3 ? 2e-7 * K * 2 : 1e-4;  // ~fp32 accumulation error of K products
  printf("M=%3d N=%3d K=%3d a_mn=%d b_mn=%d split=%d: max|err|=%.3e (%s)  lane map: %s\n", M, N, K, a_mn, b_mn, split, worst,
         worst < tol ? "PASS" : "FAIL", ident ? "row m -> lane m" : m64 ? "row m -> lane 32*(m/16)+m%16" : "OTHER");
  if (!ident && !m64) {
    for (int m = 0; m < M; ++m) printf("%d%s", lane_of[m], m + 1 < M ? "," : "\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return worst < tol ? 0 : 1;
}

// one case per process (a faulting case must not poison the others): umma_probe M N K a_mn b_mn split
int main(int argc, char** argv) {
  if (argc == 7) return run(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
  int bad = 0;
  const int shapes[5][3] = {{128, 64, 64}, {128, 128, 64}, {64, 64, 128}, {64, 32, 128}, {128, 16, 64}};
  for (int split : {1, 3})
    for (int a_mn : {0, 1})
      for (int b_mn : {0, 1})
        for (auto& sh : shapes) {
          char cmd[256];
          snprintf(cmd, sizeof(cmd), "%s %d %d %d %d %d %d", argv[0], sh[0], sh[1], sh[2], a_mn, b_mn, split);
          bad += system(cmd) != 0;
        }
  for (int split : {1, 3})  // A operand in TMEM (a_mn = 2)
    for (int b_mn : {0, 1})
      for (auto& sh : {std::array<int, 3>{128, 64, 64}, std::array<int, 3>{128, 128, 64}, std::array<int, 3>{128, 64, 128}}) {
        char cmd[256];
        snprintf(cmd, sizeof(cmd), "%s %d %d %d 2 %d %d", argv[0], sh[0], sh[1], sh[2], b_mn, split);
        bad += system(cmd) != 0;
      }
  printf("%s (%d failing cases)\n", bad ? "PROBE FAILED" : "PROBE OK", bad);
  return bad ? 1 : 0;
}
