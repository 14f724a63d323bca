// Micro-benchmark: DRAM bytes fetched per random 8-byte / 24-byte gather on B200
// for different load flavours and L2 fetch granularities.  Run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <random>

template <int MODE> __device__ inline unsigned long long ld(const unsigned long long* p) {
  unsigned long long v;
  if (MODE == 0) v = *p;
  else if (MODE == 1) v = __ldg(p);
  else if (MODE == 2) v = __ldcs(p);
  else if (MODE == 3) v = __ldcg(p);
  else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  else asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
template <int MODE, int W>
__global__ void gather(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst, const int* __restrict__ order, size_t n) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const unsigned long long* e = src + (size_t)order[k] * W;
#pragma unroll
  for (int w = 0; w < W; ++w) dst[k * W + w] = ld<MODE>(e + w);
}
template <int MODE, int W> void run(const char* name, const unsigned long long* src, unsigned long long* dst, const int* order, size_t n) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  gather<MODE, W><<<(unsigned)((n + 255) / 256), 256>>>(src, dst, order, n);
  cudaEventRecord(a);
  gather<MODE, W><<<(unsigned)((n + 255) / 256), 256>>>(src, dst, order, n);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-28s W=%d  %.3f ms  (%.1f GB/s payload r+w)\n", name, W, ms, 2.0 * n * W * 8 / ms / 1e6);
}
int main() {
  const size_t n = 32000000;
  std::vector<int> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (int)i;
  std::mt19937_64 g(1); std::shuffle(h.begin(), h.end(), g);
  int* order; unsigned long long *src, *dst;
  cudaMalloc(&order, n * 4); cudaMalloc(&src, n * 24); cudaMalloc(&dst, n * 24);
  cudaMemcpy(order, h.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemset(src, 1, n * 24);
  for (int gran : {64, 32, 128}) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("--- L2 fetch granularity request %d -> %zu (%s)\n", gran, got, cudaGetErrorString(e));
    run<0, 1>("plain", src, dst, order, n);
    run<1, 1>("ldg(nc)", src, dst, order, n);
    run<2, 1>("ldcs", src, dst, order, n);
    run<3, 1>("ldcg", src, dst, order, n);
    run<4, 1>("nc.L1::no_allocate", src, dst, order, n);
    run<5, 1>("L1::no_allocate", src, dst, order, n);
    run<1, 3>("ldg(nc)", src, dst, order, n);
    run<3, 3>("ldcg", src, dst, order, n);
  }
  return 0;
}
