// FP32 pipe issue-rate probe (sm_100a): cycles per warp instruction and FMA/clk/SM for register-operand FFMA and FFMA2
// with W warps per SM sub-partition, 16 independent accumulators per thread (no dependent-issue stalls).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_probe tools/fp32_probe.cu && ./fp32_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0: FFMA (scalar), 1: FFMA2 (packed pair), 2: FFMA2 with a broadcast scalar multiplicand
__global__ void k_probe(float* out, long long* cyc, int iters, float a0, float w0) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 a = make_float2(a0, a0 * 0.5f), w = make_float2(w0, w0 + 1e-3f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {
        acc[i].x = fmaf(a.x, w.x, acc[i].x);
        acc[i].y = fmaf(a.y, w.y, acc[i].y);
      } else if (MODE == 1) {
        acc[i] = __ffma2_rn(a, w, acc[i]);
      } else {
        acc[i] = __ffma2_rn(make_float2(a.x, a.x), w, acc[i]);
      }
    }
    a.x += 1e-7f;  // keep the operands live registers
  }
  const long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int warps_per_smsp) {
  const int threads = 128 * warps_per_smsp, iters = 4096, blocks = 148;
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  k_probe<MODE><<<blocks, threads>>>(out, cyc, iters, 1.0001f, 0.9999f);
  k_probe<MODE><<<blocks, threads>>>(out, cyc, iters, 1.0001f, 0.9999f);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; ++i) c += (double)h[i]; c /= blocks;
  const double inst_per_warp = (double)iters * 16 * (MODE == 0 ? 2 : 1);
  const double fma_per_sm = (double)iters * 16 * 2 * threads;
  printf("%-28s %d warps/SMSP: %.2f cycles per warp instruction per SMSP, %.1f FMA/clk/SM\n", name, warps_per_smsp,
         c / (inst_per_warp * warps_per_smsp), fma_per_sm / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 3, 4, 8}) {
    run<0>("FFMA  (3 register operands)", w);
    run<1>("FFMA2 (packed pairs)", w);
    run<2>("FFMA2 (broadcast scalar a)", w);
  }
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
