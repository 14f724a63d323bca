// Probe: throughput of fp64 reductions (red.global.add.f64) with the address pattern a
// symmetric (Newton's-third-law) sparse product would produce: every target bucket of a
// 147^3 grid (10 particles per bucket, sorted by bucket) adds ~215 contributions to the
// particles of its 13 forward neighbour buckets + itself.  Decides whether a half-stencil
// kernel with y[j] += F * b[i] scatter is worth building.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_probe tools/red_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int MODE> // 0: RED fp64, 1: plain store (no atomicity; traffic reference), 2: RED but one per candidate (pre-reduced)
__global__ void __launch_bounds__(128) k_red(double *y, int S, int per, uint32_t *counter, int pairs_per_bucket) {
  const int lane = threadIdx.x & 31;
  const uint32_t ncell = (uint32_t)S * S * S;
  while (true) {
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(counter, 8u);
    c = __shfl_sync(0xFFFFFFFFu, c, 0);
    if (c >= ncell) break;
    for (uint32_t cell = c; cell < min(c + 8u, ncell); ++cell) {
      const int z = cell % S, yy = (cell / S) % S, x = cell / (S * S);
      const int rounds = (pairs_per_bucket + 31) / 32;
      for (int r = 0; r < rounds; ++r) {
        const uint32_t h = hash32(cell * 131u + r * 32u + lane);
        // forward half stencil: 14 buckets = self + 13
        const int k = h % 14;
        int ox, oy, oz;
        if (k == 0) { ox = 0; oy = 0; oz = 0; }
        else if (k == 1) { ox = 0; oy = 0; oz = 1; }
        else if (k < 5) { ox = 0; oy = 1; oz = k - 3; }
        else { ox = 1; oy = (k - 5) / 3 - 1; oz = (k - 5) % 3 - 1; }
        const int nx = (x + ox) % S, ny = (yy + oy + S) % S, nz = (z + oz + S) % S;
        const uint32_t nb = ((uint32_t)nx * S + ny) * S + nz;
        const uint32_t j = nb * per + (h >> 8) % per;
        const double v = 1e-9 * lane;
        if (MODE == 1) y[j] = v;
        else atomicAdd(&y[j], v);
      }
    }
  }
}

int main() {
  const int S = 147, per = 10;
  const size_t n = (size_t)S * S * S * per;
  double *y; uint32_t *counter;
  cudaMalloc(&y, n * sizeof(double));
  cudaMalloc(&counter, 4);
  cudaMemset(y, 0, n * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int mode = 0; mode < 2; ++mode)
    for (int pairs : {215, 140, 435}) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(counter, 0, 4);
        cudaEventRecord(e0);
        if (mode == 0) k_red<0><<<sms * 8, 128>>>(y, S, per, counter, pairs);
        else k_red<1><<<sms * 8, 128>>>(y, S, per, counter, pairs);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      const double ops = (double)S * S * S * ((pairs + 31) / 32) * 32;
      printf("mode=%s pairs/bucket=%d: %.3f ms for %.0fM lane-ops = %.1f G/s (%s)\n", mode == 0 ? "RED.ADD.F64" : "plain ST.64", pairs, best, ops / 1e6,
             ops / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
