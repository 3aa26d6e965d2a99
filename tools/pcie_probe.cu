// Host <-> device transfer probe for the host-step sizes (sm_100a box, PCIe Gen5): copy engines vs copy kernels, alone and
// with both directions at once.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_probe tools/pcie_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_copy16(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
}
static void kcopy(void* d, const void* s, size_t bytes, cudaStream_t st) {
  const size_t n16 = bytes / 16;
  int grid = (int)((n16 + 255) / 256);
  if (grid > 148 * 4) grid = 148 * 4;
  k_copy16<<<grid, 256, 0, st>>>((uint4*)d, (const uint4*)s, n16);
}

template <typename F>
static float time_us(F f, cudaStream_t st, int reps = 200) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 10; ++i) f();
  cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b, st);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms * 1e3f / reps;
}

int main() {
  const size_t in_b = 917504, out_b = 1241088;
  void *h_in, *h_out, *d_in, *d_out;
  cudaHostAlloc(&h_in, in_b, cudaHostAllocDefault); cudaHostAlloc(&h_out, out_b, cudaHostAllocDefault);
  cudaMalloc(&d_in, in_b); cudaMalloc(&d_out, out_b);
  cudaStream_t s0, s1;
  cudaStreamCreate(&s0); cudaStreamCreate(&s1);
  cudaEvent_t fork, join;
  cudaEventCreateWithFlags(&fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
  printf("back-to-back on one stream, us per transfer (GB/s):\n");
  float t;
  t = time_us([&] { cudaMemcpyAsync(d_in, h_in, in_b, cudaMemcpyHostToDevice, s0); }, s0);
  printf("  copy engine  H2D %7zu B: %6.1f us (%.1f GB/s)\n", in_b, t, in_b / t / 1e3);
  t = time_us([&] { cudaMemcpyAsync(h_out, d_out, out_b, cudaMemcpyDeviceToHost, s0); }, s0);
  printf("  copy engine  D2H %7zu B: %6.1f us (%.1f GB/s)\n", out_b, t, out_b / t / 1e3);
  t = time_us([&] { kcopy(d_in, h_in, in_b, s0); }, s0);
  printf("  copy kernel  H2D %7zu B: %6.1f us (%.1f GB/s)\n", in_b, t, in_b / t / 1e3);
  t = time_us([&] { kcopy(h_out, d_out, out_b, s0); }, s0);
  printf("  copy kernel  D2H %7zu B: %6.1f us (%.1f GB/s)\n", out_b, t, out_b / t / 1e3);
  for (int parts : {4, 8}) {
    t = time_us([&] { for (int c = 0; c < parts; ++c) kcopy((char*)d_in + c * (in_b / parts), (char*)h_in + c * (in_b / parts), in_b / parts, s0); }, s0);
    printf("  copy kernel  H2D in %d launches: %6.1f us\n", parts, t);
    t = time_us([&] { for (int c = 0; c < parts; ++c) cudaMemcpyAsync((char*)d_in + c * (in_b / parts), (char*)h_in + c * (in_b / parts), in_b / parts, cudaMemcpyHostToDevice, s0); }, s0);
    printf("  copy engine  H2D in %d transfers: %6.1f us\n", parts, t);
  }
  printf("both directions at once (fork/join on two streams), us per pair:\n");
  t = time_us([&] {
    cudaEventRecord(fork, s0); cudaStreamWaitEvent(s1, fork, 0);
    cudaMemcpyAsync(d_in, h_in, in_b, cudaMemcpyHostToDevice, s0);
    cudaMemcpyAsync(h_out, d_out, out_b, cudaMemcpyDeviceToHost, s1);
    cudaEventRecord(join, s1); cudaStreamWaitEvent(s0, join, 0); }, s0);
  printf("  copy engines: %6.1f us\n", t);
  t = time_us([&] {
    cudaEventRecord(fork, s0); cudaStreamWaitEvent(s1, fork, 0);
    kcopy(d_in, h_in, in_b, s0);
    kcopy(h_out, d_out, out_b, s1);
    cudaEventRecord(join, s1); cudaStreamWaitEvent(s0, join, 0); }, s0);
  printf("  copy kernels: %6.1f us\n", t);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
