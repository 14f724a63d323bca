// Counting-sort build of the ordered cell list for sm_100a: an alternative large-N strategy of
// abr_update_positions (CellListOrdered::update_positions_impl + Particles::reorder,
// /root/reference/src/CellListOrdered.h:190-259, src/Particles.h:694-724), selected with
// abr_set_option("counting_min_n", n).  OFF by default: bit-identical to the radix builds of
// abr_build.cu (tests/test_gpu_parity.py runs all three against the oracle) but measured slower
// on B200 — 3.4 ms against 2.7 ms at 32 M particles (profiles/r2f_counting_build.txt) — for the
// reason the measurements below give.  Kept because it moves the fewest bytes of the three
// (6.0 GB against 7.2 GB) and because its stages are the ones a hardware with cheaper
// scattered stores would want.
//
// The reference sorts (bucket, index) pairs and then gathers every column through the
// permutation.  Moving a particle record from a random place to its sorted place needs two
// passes over the records whatever the sort is; here everything else is arranged so that
// nothing BUT those two record passes and one key pass touches HBM:
//
//   K1 k_cs_enforce_key   wrap / kill / bucket key (enforce_one) + the bin histogram of every
//                         block's slice of the input (top B1 key bits = bin, 2^11..2^14 bins)
//   K1b/c                 totals per (block class, bin), bin starts, cursors, alive counts
//   K2 k_cs_scatter       one thread per particle: a slot of its bin from an atomicAdd on the
//                         cursor of (class, bin), then the whole RECORD (all columns, padded to
//                         32-byte groups) with 256-bit stores, and (key | original index << 32)
//                         into an 8-byte side array.  Neither stable nor deterministic — K3
//                         restores the order.
//   K3 k_cs_binsort       one CTA per bin (a bin is a contiguous range of buckets and of the
//                         output): bucket histogram of the bin in shared memory -> exclusive
//                         scan = bucket_begin / bucket_end directly (no boundary search, no
//                         fill) -> placement with a shared-memory cursor per bucket -> every
//                         bucket's few entries ordered by ORIGINAL INDEX in registers (this is
//                         what makes the result the stable sort the reference's
//                         thrust::sort_by_key gives, independent of the atomics' timing) ->
//                         gather of the records from the bin (prefetched into L2 in address
//                         order) into the output columns, sorted keys and m_alive_indices.
//
// Measured variants of K2 at 32 M particles (gpurun_out/r2f_*, ncu):
//   tiles of 2048 particles staged by cp.async.bulk, ranks by shared-memory atomics, one
//     claimed range per (tile, bin), word-parallel 8-byte stores:              1.32 ms
//   one global cursor per bin (2^11), thread per particle, 256-bit stores:      2.11 ms  (atomics serialise per address)
//   private range per (block, bin), no atomics:                                 2.36 ms  (4.8 M open write streams: 2x HBM traffic)
//   one cursor per (class of 37 blocks, bin) — this file:                       1.40 ms  (HBM traffic 1.05x algorithmic)
// against 0.77 ms for the radix partition of abr_build.cu, whose 49 bins keep every warp
// store a contiguous run.  K3: 1.39 ms (3.2 GB at 2.3 TB/s, latency bound at 2 CTAs/SM).
// Compiled with -fmad=false like the rest of the build.
#include <algorithm>

#include <cstring>

#include "abr_internal.h"
#include "abr_enforce.cuh"

namespace abr {

constexpr int CS_THREADS = 512;
constexpr int CS_MAXW = 20;        // 8-byte words per record, key/orig word included
constexpr int CS_SMALL = 32;       // buckets up to this size are ordered by one thread

struct CsCols {
  int ncols;                       // columns that move (the alive column does not: survivors are alive)
  const uint8_t *src[GP_MAXC];
  uint8_t *dst[GP_MAXC];
  uint32_t words[GP_MAXC];  // 8-byte words per element
  uint32_t woff[GP_MAXC];   // first word of the column inside a record
  uint32_t W;                      // words per record; word W-1 = key | original index << 32
  uint8_t *alive_dst;              // may be null
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------
// K1: keys + bin histogram of this block's slice of the input.  Block g owns the particles
// [g * chunk, (g + 1) * chunk); its histogram goes to block_hist[g][*].  K2 runs with the same
// slices, so the exclusive prefix over g gives every block a private range in every bin: the
// scatter needs no global atomics (32 M atomicAdds on 2^11 cursors cost 2.1 ms: measured).
// ---------------------------------------------------------------------------
template <int D, bool WINDOWED>
__global__ void __launch_bounds__(CS_THREADS) k_cs_enforce_key(double *__restrict__ pos, uint8_t *__restrict__ alive, uint32_t n, uint32_t chunk, Grid g,
                                                               uint32_t *__restrict__ keys, DevScalars *sc, uint32_t *__restrict__ block_hist,
                                                               int shift, uint32_t NB) {
  extern __shared__ uint32_t s_hist[];
  for (uint32_t b = threadIdx.x; b < NB; b += CS_THREADS) s_hist[b] = 0;
  __syncthreads();
  uint32_t dead = 0;
  const uint32_t p0 = blockIdx.x * chunk, p1 = min(n, p0 + chunk);
  for (uint32_t pb = p0 + threadIdx.x; pb < p1; pb += 4 * CS_THREADS) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t p = pb + u * CS_THREADS;
      if (p < p1) {
        const uint32_t key = enforce_one<D, WINDOWED>(pos, alive, p, g, sc);
        keys[p] = key;
        atomicAdd(&s_hist[key >> shift], 1u);
        dead += key == g.key_bound;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dead += __shfl_xor_sync(0xFFFFFFFFu, dead, o);
  if ((threadIdx.x & 31) == 0 && dead) atomicAdd(&sc->pad[0], dead);
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < NB; b += CS_THREADS) block_hist[(size_t)blockIdx.x * NB + b] = s_hist[b];
}

// K1b: totals per (class, bin): class q = the blocks g with g % CS_Q == q.  The scatter keeps one
// global cursor per (class, bin): few enough streams (CS_Q * NB lines) for L2 to assemble full
// lines before they go to HBM, many enough addresses for the atomics not to serialise
// (one cursor per bin: 32 M atomicAdds on 2^11 addresses took 2.1 ms; one private range per
// (block, bin): no atomics but 4.8 M open write streams, 2x the HBM traffic and 2.4 ms — measured).
constexpr uint32_t CS_Q = 16;
__global__ void __launch_bounds__(256) k_cs_class_totals(const uint32_t *__restrict__ block_hist, uint32_t G, uint32_t NB, uint32_t *__restrict__ qtot) {
  const uint32_t t = blockIdx.x * 256u + threadIdx.x; // (q, bin), bin fastest
  if (t >= CS_Q * NB) return;
  const uint32_t q = t / NB, b = t - q * NB;
  uint32_t sum = 0;
  for (uint32_t g0 = q; g0 < G; g0 += 4 * CS_Q) {
    uint32_t c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c[u] = g0 + u * CS_Q < G ? block_hist[(size_t)(g0 + u * CS_Q) * NB + b] : 0u;
    sum += (c[0] + c[1]) + (c[2] + c[3]);
  }
  qtot[t] = sum;
}

// K1c: bin starts (exclusive scan of the bin totals), the cursors of every (class, bin), alive / in-cell counts
__global__ void __launch_bounds__(1024) k_cs_bin_scan(const uint32_t *__restrict__ qtot, uint32_t NB, uint32_t n, uint32_t *__restrict__ bin_start,
                                                      uint32_t *__restrict__ cursor, DevScalars *sc) {
  __shared__ uint32_t s_w[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t per = (NB + 1023u) / 1024u; // consecutive bins per thread
  const uint32_t b0 = threadIdx.x * per;
  uint32_t local = 0;
  for (uint32_t k = 0; k < per; ++k)
    if (b0 + k < NB)
      for (uint32_t q = 0; q < CS_Q; ++q) local += qtot[q * NB + b0 + k];
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_w[w];
  uint32_t run = woff + incl - local;
  for (uint32_t k = 0; k < per; ++k) {
    if (b0 + k < NB) {
      bin_start[b0 + k] = run;
      for (uint32_t q = 0; q < CS_Q; ++q) {
        cursor[q * NB + b0 + k] = run;
        run += qtot[q * NB + b0 + k];
      }
    }
  }
  if (threadIdx.x == 0) {
    bin_start[NB] = n;
    const uint32_t n_alive = n - sc->pad[0];
    sc->n_alive = n_alive;
    sc->n_incell = n_alive - sc->n_aliased;
  }
}

// ---------------------------------------------------------------------------
// K2: scatter of whole records into the bins.  One slot per particle from an atomicAdd on the
// cursor of (class of this block, bin); the record — every moving column, padded to a multiple
// of 32 bytes — leaves with 256-bit stores (one full, aligned sector each: nothing for L2 to
// read-modify-write), and (key | original index << 32) goes to a separate 8-byte array that K3
// streams twice.  The order inside a bin is whatever the atomics give; K3 restores the
// original order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void st256(void *p, const unsigned long long *v) {
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]) : "memory");
}
__device__ __forceinline__ void ld256(const void *p, unsigned long long *v) {
  asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}

template <int NG> // 32-byte groups per record
__global__ void __launch_bounds__(CS_THREADS) k_cs_scatter(const uint32_t *__restrict__ keys, uint32_t n, uint32_t chunk, int shift, uint32_t NB,
                                                           const CsCols cols, uint32_t *__restrict__ cursor, unsigned long long *__restrict__ rec,
                                                           unsigned long long *__restrict__ ko) {
  uint32_t *cur = cursor + (size_t)(blockIdx.x % CS_Q) * NB;
  const uint32_t p0 = blockIdx.x * chunk, p1 = min(n, p0 + chunk);
  constexpr int U = NG == 1 ? 4 : 2; // particles in flight per thread
  for (uint32_t pb = p0 + threadIdx.x; pb < p1; pb += U * CS_THREADS) {
    uint32_t key[U];
    unsigned long long v[U][NG * 4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = pb + u * CS_THREADS;
      key[u] = p < p1 ? keys[p] : 0u;
#pragma unroll
      for (int w = 0; w < NG * 4; ++w) v[u][w] = 0;
      if (p < p1) {
        int w = 0;
        for (int c = 0; c < cols.ncols; ++c) {
          const unsigned long long *src = reinterpret_cast<const unsigned long long *>(cols.src[c]) + (size_t)p * cols.words[c];
          for (uint32_t k = 0; k < cols.words[c]; ++k) {
            const unsigned long long x = __ldg(src + k);
#pragma unroll
            for (int q = 0; q < NG * 4; ++q)
              if (q == w) v[u][q] = x; // static indexing: the record stays in registers
            ++w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = pb + u * CS_THREADS;
      if (p < p1) {
        const uint32_t slot = atomicAdd(&cur[key[u] >> shift], 1u);
        unsigned long long *dst = rec + (size_t)slot * (NG * 4);
#pragma unroll
        for (int g = 0; g < NG; ++g) st256(dst + 4 * g, v[u] + 4 * g);
        ko[slot] = (unsigned long long)key[u] | ((unsigned long long)p << 32);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K3: per-bin bucket sort, bucket ranges, final reorder
// ---------------------------------------------------------------------------
// ascending order of up to 16 distinct 64-bit values held in registers (odd-even transposition,
// fully unrolled: no dynamic indexing, unused places hold ~0)
__device__ __forceinline__ void sort16(unsigned long long (&v)[16]) {
#pragma unroll
  for (int round = 0; round < 16; ++round) {
#pragma unroll
    for (int k = round & 1; k + 1 < 16; k += 2) {
      const unsigned long long a = v[k], b = v[k + 1];
      const bool sw = a > b;
      v[k] = sw ? b : a;
      v[k + 1] = sw ? a : b;
    }
  }
}

template <int NG>
__global__ void __launch_bounds__(CS_THREADS, 2)
k_cs_binsort(const unsigned long long *__restrict__ ko, const unsigned long long *__restrict__ rec, unsigned long long *__restrict__ scratch,
             unsigned long long *__restrict__ scratch2, const uint32_t *__restrict__ bin_start, int shift, uint32_t ncells, uint32_t dead_key,
             const CsCols cols, uint32_t *__restrict__ bb, uint32_t *__restrict__ be, uint32_t *__restrict__ sorted_keys,
             int32_t *__restrict__ order_out, uint32_t *__restrict__ max_bucket) {
  extern __shared__ __align__(128) unsigned char cs_raw[];
  const uint32_t S = 1u << shift;
  uint32_t *off = reinterpret_cast<uint32_t *>(cs_raw); // [S + 1] exclusive offsets
  uint32_t *cur = off + S + 1;                          // [S] cursors; afterwards the list of big buckets
  __shared__ uint32_t s_w[CS_THREADS / 32];
  __shared__ uint32_t s_nbig;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bin = blockIdx.x;
  const uint32_t bs = bin_start[bin], nb = bin_start[bin + 1] - bs;
  const uint32_t mask = S - 1u;
  const unsigned long long *kob = ko + bs;
  // the bin's records are read in random order by phase E: pull the (contiguous) range into L2 now,
  // full lines in address order, so that HBM sees one sequential read of the bin
  {
    const char *rbase = reinterpret_cast<const char *>(rec + (size_t)bs * (NG * 4));
    const size_t rbytes = (size_t)nb * (NG * 32);
    for (size_t o = (size_t)tid * 128; o < rbytes; o += (size_t)CS_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rbase + o));
  }

  // A. bucket histogram of the bin
  for (uint32_t s = tid; s < S; s += CS_THREADS) off[s] = 0;
  if (tid == 0) s_nbig = 0;
  __syncthreads();
  for (uint32_t e0 = tid; e0 < nb; e0 += 4 * CS_THREADS) {
    unsigned long long v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = e0 + u * CS_THREADS < nb ? kob[e0 + u * CS_THREADS] : 0ull;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (e0 + u * CS_THREADS < nb) atomicAdd(&off[(uint32_t)v[u] & mask], 1u);
  }
  __syncthreads();
  // B. exclusive scan over the S buckets (each thread a consecutive slice) -> bucket ranges
  {
    const uint32_t per = (S + CS_THREADS - 1) / CS_THREADS;
    const uint32_t s0 = tid * per;
    uint32_t local = 0;
    for (uint32_t k = 0; k < per; ++k)
      if (s0 + k < S) local += off[s0 + k];
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_w[w];
    uint32_t run = woff + incl - local;
    for (uint32_t k = 0; k < per; ++k) {
      if (s0 + k < S) {
        const uint32_t c = off[s0 + k];
        off[s0 + k] = run;
        cur[s0 + k] = run;
        const uint32_t bucket = (bin << shift) + s0 + k;
        if (bucket < ncells) {
          // == lower_bound / upper_bound of the bucket id in the sorted keys (src/CellListOrdered.h:229-239)
          bb[bucket] = bs + run;
          be[bucket] = bs + run + c;
          if (c > 64) atomicMax(max_bucket, c);
        }
        run += c;
      }
    }
    if (tid == CS_THREADS - 1) off[S] = nb;
  }
  __syncthreads();
  // C. placement (arrival order inside a bucket)
  for (uint32_t e0 = tid; e0 < nb; e0 += 4 * CS_THREADS) {
    unsigned long long v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = e0 + u * CS_THREADS < nb ? kob[e0 + u * CS_THREADS] : 0ull;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t e = e0 + u * CS_THREADS;
      if (e < nb) {
        const uint32_t slot = atomicAdd(&cur[(uint32_t)v[u] & mask], 1u);
        scratch[bs + slot] = (v[u] & 0xFFFFFFFF00000000ull) | e; // original index << 32 | position in the bin
      }
    }
  }
  __syncthreads();
  // D. order every bucket by original index: the stable order of the reference's sort_by_key
  for (uint32_t s = tid; s < S; s += CS_THREADS) {
    const uint32_t a = off[s], c = off[s + 1] - a;
    unsigned long long *base = scratch + bs + a;
    if (c > 1 && c <= 16) {
      unsigned long long v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = (uint32_t)k < c ? base[k] : ~0ull;
      sort16(v);
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if ((uint32_t)k < c) base[k] = v[k];
    } else if (c > 16 && c <= CS_SMALL) {
      unsigned long long v[CS_SMALL];
      for (uint32_t k = 0; k < c; ++k) v[k] = base[k];
      for (uint32_t k = 1; k < c; ++k) { // insertion sort
        const unsigned long long x = v[k];
        uint32_t m = k;
        while (m > 0 && v[m - 1] > x) {
          v[m] = v[m - 1];
          --m;
        }
        v[m] = x;
      }
      for (uint32_t k = 0; k < c; ++k) base[k] = v[k];
    } else if (c > CS_SMALL) {
      const uint32_t k = atomicAdd(&s_nbig, 1u);
      cur[k] = s; // cursors are no longer needed: reuse as the list of big buckets
    }
  }
  __syncthreads();
  // big buckets (clustered clouds): rank by counting, one warp per bucket
  const uint32_t nbig = s_nbig;
  for (uint32_t kb = warp; kb < nbig; kb += CS_THREADS / 32) {
    const uint32_t s = cur[kb];
    const uint32_t a = off[s], c = off[s + 1] - a;
    const unsigned long long *base = scratch + bs + a;
    for (uint32_t k = lane; k < c; k += 32) {
      const unsigned long long x = base[k];
      uint32_t r = 0;
      for (uint32_t m = 0; m < c; ++m) r += base[m] < x;
      scratch2[bs + a + r] = x;
    }
    __syncwarp();
    for (uint32_t k = lane; k < c; k += 32) scratch[bs + a + k] = scratch2[bs + a + k];
  }
  __syncthreads();
  // E. one thread per output particle: its record (256-bit loads from the bin, L2 resident) -> the output columns
  for (uint32_t k0 = tid; k0 < nb; k0 += 2 * CS_THREADS) {
    unsigned long long se[2], kv[2], v[2][NG * 4];
#pragma unroll
    for (int u = 0; u < 2; ++u) se[u] = k0 + u * CS_THREADS < nb ? scratch[bs + k0 + u * CS_THREADS] : 0ull;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t e = (uint32_t)se[u];
      kv[u] = kob[e];
#pragma unroll
      for (int g = 0; g < NG; ++g) ld256(rec + (size_t)(bs + e) * (NG * 4) + 4 * g, v[u] + 4 * g);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t k = k0 + u * CS_THREADS;
      if (k >= nb) continue;
      const uint32_t key = (uint32_t)kv[u];
      sorted_keys[bs + k] = key;
      order_out[bs + k] = (int32_t)(se[u] >> 32);
      if (cols.alive_dst) cols.alive_dst[bs + k] = key != dead_key ? 1 : 0;
      int w = 0;
      for (int c = 0; c < cols.ncols; ++c) {
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(cols.dst[c]) + (size_t)(bs + k) * cols.words[c];
        for (uint32_t q = 0; q < cols.words[c]; ++q) {
          unsigned long long x = 0;
#pragma unroll
          for (int r = 0; r < NG * 4; ++r)
            if (r == w) x = v[u][r];
          dst[q] = x;
          ++w;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
// Can this update use the counting-sort build?  Every moving column made of 8-byte words
// (position, id, doubles, the 104-byte generator state ...), the alive column recognised by
// its pointer, a record of at most CS_MAXW words.
bool counting_build_applicable(const Handle *h, size_t n, int bits, const ReorderSpec *reorder, const uint8_t *alive) {
  if (!reorder || n < h->counting_min_n || n >= 0x7FFFFFF0ull) return false;
  if (bits < 8 || bits > 27) return false;
  uint32_t W = 1;
  int moving = 0;
  for (int c = 0; c < reorder->ncols; ++c) {
    if (reorder->src[c] == alive) {
      if (reorder->elem_bytes[c] != 1) return false;
      continue;
    }
    const size_t eb = reorder->elem_bytes[c];
    if (eb == 0 || (eb & 7u) || ((uintptr_t)reorder->src[c] & 7u) || ((uintptr_t)reorder->dst[c] & 7u)) return false;
    W += (uint32_t)(eb / 8);
    ++moving;
  }
  if (!(moving >= 1 && moving <= GP_MAXC && W <= CS_MAXW)) return false;
  return true;
}

template <int D, bool WIN>
static void launch_cs_key(Handle *h, double *pos, uint8_t *alive, uint32_t n, uint32_t chunk, uint32_t G, const Grid &g, uint32_t *keys,
                          uint32_t *block_hist, int shift, uint32_t NB) {
  cudaFuncSetAttribute(k_cs_enforce_key<D, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NB * sizeof(uint32_t)));
  k_cs_enforce_key<D, WIN><<<G, CS_THREADS, NB * sizeof(uint32_t), h->stream>>>(pos, alive, n, chunk, g, keys, h->d_scalars, block_hist, shift, NB);
}

int build_counting(Handle *h, double *pos, uint8_t *alive, uint32_t n, const Grid &g, int bits, const ReorderSpec *reorder, int32_t *order_out) {
  CsCols cols;
  memset(&cols, 0, sizeof(cols));
  uint32_t W = 0;
  for (int c = 0; c < reorder->ncols; ++c) {
    if (reorder->src[c] == alive) {
      cols.alive_dst = static_cast<uint8_t *>(reorder->dst[c]);
      continue;
    }
    const int k = cols.ncols++;
    cols.src[k] = static_cast<const uint8_t *>(reorder->src[c]);
    cols.dst[k] = static_cast<uint8_t *>(reorder->dst[c]);
    cols.words[k] = (uint32_t)(reorder->elem_bytes[c] / 8);
    cols.woff[k] = W;
    W += cols.words[k];
  }
  cols.W = W + 1;
  const int NG = (int)((W + 3) / 4); // 32-byte groups per record

  // Digits: the top B1 bits of the key pick the bin, the remaining `shift` bits the bucket inside
  // the bin.  Bins are sized so that the records of all bins in flight (2 CTAs per SM) stay in L2
  // (phase E of K3 reads them in random order): ~32 MB in flight.
  const double in_flight = 2.0 * h->sm_count * (double)NG * 32.0;
  int B1 = 8;
  while (B1 < 14 && ((double)n / (double)(1u << B1)) * in_flight > 32.0e6) ++B1;
  B1 = std::max(B1, bits - 13); // at most 2^13 buckets per bin (their counters live in shared memory)
  B1 = std::min(B1, bits);
  if (bits - B1 > 13) return set_error(h, ABR_ERR_UNSUPPORTED, "counting build: too many buckets per bin");
  const int shift = bits - B1;
  const uint32_t NB = 1u << B1;
  const uint32_t S = 1u << shift;
  // slices of the input: one per block, the same in K1 and K2
  uint32_t G = (uint32_t)h->sm_count * 4u;
  uint32_t chunk = (n + G - 1) / G;
  chunk = (chunk + CS_THREADS - 1) / CS_THREADS * CS_THREADS;
  G = (n + chunk - 1) / chunk;

  ABR_CUDA(h, h->keys[0].reserve((size_t)n * sizeof(uint32_t)));
  ABR_CUDA(h, h->keys[1].reserve((size_t)n * sizeof(uint32_t)));
  ABR_CUDA(h, h->tmp_cols.reserve((size_t)n * NG * 32));
  ABR_CUDA(h, h->cs_ko.reserve((size_t)n * 8));
  ABR_CUDA(h, h->cs_scratch.reserve((size_t)n * 8));
  ABR_CUDA(h, h->cs_scratch2.reserve((size_t)n * 8));
  ABR_CUDA(h, h->cs_bins.reserve(((size_t)(G + 2 * CS_Q + 1) * NB + 8) * sizeof(uint32_t)));
  uint32_t *keys = h->keys[0].as<uint32_t>();
  uint32_t *sorted_keys = h->keys[1].as<uint32_t>();
  uint32_t *block_hist = h->cs_bins.as<uint32_t>(); // G x NB
  uint32_t *qtot = block_hist + (size_t)G * NB;     // CS_Q x NB
  uint32_t *cursor = qtot + (size_t)CS_Q * NB;      // CS_Q x NB
  uint32_t *bin_start = cursor + (size_t)CS_Q * NB; // NB + 1
  unsigned long long *rec = h->tmp_cols.as<unsigned long long>();
  unsigned long long *ko = h->cs_ko.as<unsigned long long>();

  const int D = h->D;
  if (h->windowed) {
    if (D == 2) launch_cs_key<2, true>(h, pos, alive, n, chunk, G, g, keys, block_hist, shift, NB);
    else launch_cs_key<3, true>(h, pos, alive, n, chunk, G, g, keys, block_hist, shift, NB);
  } else {
    if (D == 1) launch_cs_key<1, false>(h, pos, alive, n, chunk, G, g, keys, block_hist, shift, NB);
    else if (D == 2) launch_cs_key<2, false>(h, pos, alive, n, chunk, G, g, keys, block_hist, shift, NB);
    else launch_cs_key<3, false>(h, pos, alive, n, chunk, G, g, keys, block_hist, shift, NB);
  }
  k_cs_class_totals<<<(CS_Q * NB + 255) / 256, 256, 0, h->stream>>>(block_hist, G, NB, qtot);
  k_cs_bin_scan<<<1, 1024, 0, h->stream>>>(qtot, NB, n, bin_start, cursor, h->d_scalars);

  const size_t sort_smem = ((size_t)2 * S + 2) * sizeof(uint32_t);
  unsigned long long *scr = h->cs_scratch.as<unsigned long long>(), *scr2 = h->cs_scratch2.as<unsigned long long>();
  uint32_t *bbp = h->bucket_begin.as<uint32_t>(), *bep = h->bucket_end.as<uint32_t>();
#define ABR_CS_LAUNCH(NGV)                                                                                                             \
  {                                                                                                                                    \
    k_cs_scatter<NGV><<<G, CS_THREADS, 0, h->stream>>>(keys, n, chunk, shift, NB, cols, cursor, rec, ko);                               \
    ABR_CUDA(h, cudaFuncSetAttribute(k_cs_binsort<NGV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));                  \
    k_cs_binsort<NGV><<<NB, CS_THREADS, sort_smem, h->stream>>>(ko, rec, scr, scr2, bin_start, shift, g.ncells, g.key_bound, cols, bbp, bep, \
                                                                 sorted_keys, order_out, &h->d_scalars->max_bucket);                   \
  }
  switch (NG) {
  case 1: ABR_CS_LAUNCH(1) break;
  case 2: ABR_CS_LAUNCH(2) break;
  case 3: ABR_CS_LAUNCH(3) break;
  case 4: ABR_CS_LAUNCH(4) break;
  default: ABR_CS_LAUNCH(5) break;
  }
#undef ABR_CS_LAUNCH
  h->launches += 5;
  ABR_CUDA(h, cudaGetLastError());
  h->sorted_keys = sorted_keys;
  return ABR_OK;
}

} // namespace abr
