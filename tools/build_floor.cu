// Floor of the cell-list build on this GPU: the byte movement of each build stage executed as
// PLAIN streaming reads and writes (16-byte vector accesses, grid-stride, no sorting logic).
// What a build with this data flow could reach at best; the number quoted next to the >= 60 %
// HBM target of BASELINE.md §2.  Stages = those of the two-level radix build
// (aboria_b200/csrc/abr_build.cu) for the bench workload: N particles, position (24 B) + id (8 B)
// + alive (1 B) reordered, C buckets.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build_floor tools/build_floor.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_stream(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t nread, size_t nwrite) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = t; i < nread; i += stride) {
    const uint4 v = src[i];
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  for (size_t i = t; i < nwrite; i += stride) dst[i] = acc;
}

int main(int argc, char **argv) {
  const size_t N = argc > 1 ? strtoull(argv[1], nullptr, 10) : 32000000ull;
  const size_t C = argc > 2 ? strtoull(argv[2], nullptr, 10) : 3176523ull;
  struct Stage { const char *name; double rd, wr; };
  const double n = (double)N, c = (double)C;
  const Stage stages[] = {
      {"key: read position + alive, write key", 25 * n, 4 * n},
      {"level 1: read key + record, write record + key + original index", 41 * n, 45 * n},
      {"level 2, pass 1: read (key, index), write (key, index)", 8 * n, 8 * n},
      {"level 2, pass 2: read (key, index), write (key, index)", 8 * n, 8 * n},
      {"bounds: read sorted keys, write bucket_begin/end", 4 * n, 8 * c},
      {"reorder: read permutation + binned record + original index, write record + order", 45 * n, 41 * n},
      {"ALGORITHMIC (SURVEY 8d): positions, id, alive read once and written once + order + bucket ranges", 33 * n, 37 * n + 8 * c},
  };
  size_t maxb = 0;
  for (const Stage &s : stages) { if (s.rd > maxb) maxb = (size_t)s.rd; if (s.wr > maxb) maxb = (size_t)s.wr; }
  uint4 *src, *dst;
  cudaMalloc(&src, maxb + 64); cudaMalloc(&dst, maxb + 64);
  cudaMemset(src, 1, maxb + 64);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double total_ms = 0, total_bytes = 0;
  printf("build floor, N = %zu particles, C = %zu buckets (plain streaming copies of each stage's bytes)\n", N, C);
  for (size_t k = 0; k < sizeof(stages) / sizeof(stages[0]); ++k) {
    const Stage &s = stages[k];
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      k_stream<<<sms * 16, 256>>>(src, dst, (size_t)(s.rd / 16), (size_t)(s.wr / 16));
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const bool alg = k + 1 == sizeof(stages) / sizeof(stages[0]);
    if (!alg) { total_ms += best; total_bytes += s.rd + s.wr; }
    printf("  %-100s %7.3f GB  %7.3f ms  %7.1f GB/s\n", s.name, (s.rd + s.wr) / 1e9, best, (s.rd + s.wr) / best / 1e6);
    if (k + 2 == sizeof(stages) / sizeof(stages[0]))
      printf("  %-100s %7.3f GB  %7.3f ms   <- floor of the two-level data flow\n", "SUM of the six stages", total_bytes / 1e9, total_ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
