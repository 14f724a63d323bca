// get_neighbouring_buckets(query) -> bucket_pair_iterator
// (/root/reference/src/Search.h:498-764, :857-860): the "fast cell-list search"
// of tests/neighbours.h:281-300 — every pair of touching buckets once, plus the
// pairs across a periodic boundary with their position offset.
//
// The reference walks a serial iterator.  Here the same SEQUENCE is produced in
// parallel: the iterator is a concatenation of phases (domain-domain first, then
// one per periodic quadrant after the zero quadrant, lattice order), each phase is
// "for i in a box of buckets (row-major), for j in a box around i (row-major)",
// so the position of every (i, j) in the sequence follows from an exclusive scan
// of the per-i pair counts.  abr_bucket_pairs writes the sequence (bit-exact
// against the oracle's literal restatement of the iterator);
// abr_fast_bucket_search_counts runs the user-side loops the reference documents
// for it (tests/neighbours.h:892-951) on the device.
#include <cstring>

#include "abr_internal.h"

namespace abr {

int scan_exclusive_u32(Handle *h, uint32_t *data, uint64_t m); // abr_build.cu

struct PairPhase {
  int q[MAXD];      // periodic quadrant (all zero: domain-domain)
  int lo[MAXD];     // box of the i buckets: lo .. lo + n - 1 per dimension
  int n[MAXD];
  uint32_t n_i;     // number of i buckets visited (domain-domain: all but the last)
  int domain;       // 1: j runs over the part of the box AFTER i (m_j = *m_i; ++m_j)
};

// src/Search.h:675-707 get_neighbouring_buckets(query, bucket): box [lo, lo+n) or empty
template <int D> __device__ inline bool neighbour_box(const Grid &g, const int *b, int *lo, int *n) {
  bool none = false;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    int s = b[d] - 1, e = b[d] + 1;
    if (s < 0) {
      s = 0;
    } else if (s > g.end[d]) {
      none = true;
      s = g.end[d];
    }
    if (e < 0) {
      none = true;
      e = 0;
    } else if (e > g.end[d]) {
      e = g.end[d];
    }
    lo[d] = s;
    n[d] = e - s + 1;
  }
  return !none;
}

template <int D> __device__ inline void decode_i(const PairPhase &ph, uint32_t t, int *iv) {
  uint32_t rem = t;
#pragma unroll
  for (int d = D - 1; d >= 0; --d) {
    iv[d] = ph.lo[d] + (int)(rem % (uint32_t)ph.n[d]);
    rem /= (uint32_t)ph.n[d];
  }
}

// pairs contributed by the t-th i bucket of the phase
template <int D> __global__ void __launch_bounds__(256) k_pair_counts(const Grid g, const PairPhase ph, uint32_t *__restrict__ counts) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ph.n_i) return;
  int iv[D], b[D], lo[D], n[D];
  decode_i<D>(ph, t, iv);
#pragma unroll
  for (int d = 0; d < D; ++d) b[d] = iv[d] + ph.q[d] * (g.end[d] + 1); // :709-720
  uint32_t c = 0;
  if (neighbour_box<D>(g, b, lo, n)) {
    uint32_t box = 1, rank = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      box *= (uint32_t)n[d];
      rank = rank * (uint32_t)n[d] + (uint32_t)(iv[d] - lo[d]);
    }
    c = ph.domain ? box - rank - 1u : box;
  }
  counts[t] = c;
}

template <int D>
__global__ void __launch_bounds__(256) k_pair_fill(const Grid g, const PairPhase ph, const uint32_t *__restrict__ offsets, uint64_t base,
                                                   uint64_t capacity, uint32_t *__restrict__ bucket_i, uint32_t *__restrict__ bucket_j,
                                                   int8_t *__restrict__ quadrant) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ph.n_i) return;
  int iv[D], b[D], lo[D], n[D];
  decode_i<D>(ph, t, iv);
#pragma unroll
  for (int d = 0; d < D; ++d) b[d] = iv[d] + ph.q[d] * (g.end[d] + 1);
  if (!neighbour_box<D>(g, b, lo, n)) return;
  uint32_t box = 1, rank = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    box *= (uint32_t)n[d];
    rank = rank * (uint32_t)n[d] + (uint32_t)(iv[d] - lo[d]);
  }
  const uint32_t i_lin = (uint32_t)collapse_index<D>(g, iv);
  uint64_t out = base + offsets[t];
  for (uint32_t e = ph.domain ? rank + 1u : 0u; e < box; ++e, ++out) {
    if (out >= capacity) return;
    int jv[D];
    uint32_t rem = e;
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
      jv[d] = lo[d] + (int)(rem % (uint32_t)n[d]);
      rem /= (uint32_t)n[d];
    }
    bucket_i[out] = i_lin;
    bucket_j[out] = (uint32_t)collapse_index<D>(g, jv);
#pragma unroll
    for (int d = 0; d < D; ++d) quadrant[out * D + d] = (int8_t)ph.q[d];
  }
}

// the phases of the iterator in order (host): zero quadrant, then the quadrants after it
// in the lattice over {-1,0,1} (periodic dimensions) / {0}, last dimension fastest
static int make_phases(const Handle *h, PairPhase *out) {
  const int D = h->D;
  int np = 0;
  int q[MAXD] = {0, 0, 0};
  bool first = true;
  while (true) {
    PairPhase ph;
    memset(&ph, 0, sizeof(ph));
    uint64_t ni = 1;
    for (int d = 0; d < D; ++d) {
      ph.q[d] = q[d];
      ph.lo[d] = 0;
      ph.n[d] = (int)h->size[d];
      if (h->periodic[d]) { // :722-751 get_regular_buckets
        if (q[d] > 0) {
          ph.lo[d] = 0;
          ph.n[d] = 1;
        } else if (q[d] < 0) {
          ph.lo[d] = (int)h->size[d] - 1;
          ph.n[d] = 1;
        }
      }
      ni *= (uint64_t)ph.n[d];
    }
    for (int d = D; d < MAXD; ++d) ph.n[d] = 1;
    ph.domain = first ? 1 : 0;
    ph.n_i = (uint32_t)(first ? ni - 1 : ni); // domain-domain stops when m_i + 1 is the end
    if (ph.n_i > 0) out[np++] = ph;
    first = false;
    // ++m_periodic (lattice_iterator::increment over [-1,2) / [0,1))
    int d = D - 1;
    for (; d >= 0; --d) {
      const int mx = h->periodic[d] ? 2 : 1, mn = h->periodic[d] ? -1 : 0;
      ++q[d];
      if (q[d] < mx) break;
      if (d != 0) q[d] = mn;
    }
    if (d < 0 || q[0] >= (h->periodic[0] ? 2 : 1)) break;
  }
  return np;
}

template <int D>
static int bucket_pairs_impl(Handle *h, uint32_t *bucket_i, uint32_t *bucket_j, int8_t *quadrant, uint64_t capacity, uint64_t *n_host) {
  PairPhase phases[32];
  const int np = make_phases(h, phases);
  const Grid g = h->grid();
  uint64_t total = 0;
  for (int p = 0; p < np; ++p) {
    const PairPhase &ph = phases[p];
    ABR_CUDA(h, h->scan_tmp2.reserve(((size_t)ph.n_i + 1) * sizeof(uint32_t)));
    uint32_t *counts = h->scan_tmp2.as<uint32_t>();
    fill_u32(h, counts + ph.n_i, 0u, 1);
    const unsigned grid = (ph.n_i + 255) / 256;
    k_pair_counts<D><<<grid, 256, 0, h->stream>>>(g, ph, counts);
    h->launches += 1;
    int rc = scan_exclusive_u32(h, counts, (uint64_t)ph.n_i + 1);
    if (rc) return rc;
    uint32_t phase_total = 0;
    ABR_CUDA(h, cudaMemcpyAsync(&phase_total, counts + ph.n_i, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    ABR_CUDA(h, cudaStreamSynchronize(h->stream));
    if (bucket_i) {
      k_pair_fill<D><<<grid, 256, 0, h->stream>>>(g, ph, counts, total, capacity, bucket_i, bucket_j, quadrant);
      h->launches += 1;
      ABR_CUDA(h, cudaStreamSynchronize(h->stream)); // counts is reused by the next phase
    }
    total += phase_total;
  }
  ABR_CUDA(h, cudaGetLastError());
  if (n_host) *n_host = total;
  return ABR_OK;
}

int run_bucket_pairs(Handle *h, uint32_t *bucket_i, uint32_t *bucket_j, int8_t *quadrant, uint64_t capacity, uint64_t *n_host) {
  if (!h->built) return set_error(h, ABR_ERR_STATE, "bucket pairs: cell list has not been built");
  if (h->windowed) return set_error(h, ABR_ERR_UNSUPPORTED, "bucket pairs: not available on a slab window");
  if (bucket_i && (!bucket_j || !quadrant)) return set_error(h, ABR_ERR_INVALID, "bucket pairs: null output");
  uint64_t cells = 1;
  for (int d = 0; d < h->D; ++d) cells *= h->size[d];
  if (cells * 14 >= 0xFFFFFFFFull) return set_error(h, ABR_ERR_UNSUPPORTED, "bucket pairs: too many buckets");
  switch (h->D) {
  case 1: return bucket_pairs_impl<1>(h, bucket_i, bucket_j, quadrant, capacity, n_host);
  case 2: return bucket_pairs_impl<2>(h, bucket_i, bucket_j, quadrant, capacity, n_host);
  default: return bucket_pairs_impl<3>(h, bucket_i, bucket_j, quadrant, capacity, n_host);
  }
}

// ---------------------------------------------------------------------------
// tests/neighbours.h:892-951: for every bucket pair, every particle pair with
// |p_i + offset - p_j|^2 < r^2 (strict) counts for both particles; inside a bucket every
// unordered pair once, and every particle counts itself.
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_fast_pairs(const Grid g, const double *__restrict__ pos, const uint32_t *__restrict__ bb,
                                                   const uint32_t *__restrict__ be, const uint32_t *__restrict__ bucket_i,
                                                   const uint32_t *__restrict__ bucket_j, const int8_t *__restrict__ quadrant, uint64_t n_pairs,
                                                   double r2, uint32_t *__restrict__ count) {
  const uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; // one warp per bucket pair
  const int lane = threadIdx.x & 31;
  if (k >= n_pairs) return;
  const uint32_t bi = bucket_i[k], bj = bucket_j[k];
  const uint32_t a0 = bb[bi], na = be[bi] - a0, b0 = bb[bj], nb = be[bj] - b0;
  double off[D];
#pragma unroll
  for (int d = 0; d < D; ++d) off[d] = (double)quadrant[k * D + d] * (g.bmax[d] - g.bmin[d]);
  for (uint32_t e = lane; e < na * nb; e += 32) {
    const uint32_t a = a0 + e / nb, b = b0 + e % nb;
    double n2 = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double t = (pos[(size_t)a * D + d] + off[d]) - pos[(size_t)b * D + d];
      n2 += t * t;
    }
    if (n2 < r2) {
      atomicAdd(&count[a], 1u);
      atomicAdd(&count[b], 1u);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256) k_fast_self(const double *__restrict__ pos, const uint32_t *__restrict__ bb, const uint32_t *__restrict__ be,
                                                  uint32_t n_buckets, double r2, uint32_t *__restrict__ count) {
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; // one warp per bucket
  const int lane = threadIdx.x & 31;
  if (c >= n_buckets) return;
  const uint32_t a0 = bb[c], na = be[c] - a0;
  for (uint32_t e = lane; e < na * na; e += 32) {
    const uint32_t x = e / na, y = e % na;
    if (x == y) atomicAdd(&count[a0 + x], 1u); // self is a neighbour
    if (y <= x) continue;
    double n2 = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double t = pos[(size_t)(a0 + x) * D + d] - pos[(size_t)(a0 + y) * D + d];
      n2 += t * t;
    }
    if (n2 < r2) {
      atomicAdd(&count[a0 + x], 1u);
      atomicAdd(&count[a0 + y], 1u);
    }
  }
}

template <int D> static int fast_counts_impl(Handle *h, double radius, uint32_t *count, uint64_t n_pairs) {
  const Grid g = h->grid();
  const double r2 = radius * radius;
  const uint32_t *bi = h->pair_i.as<uint32_t>(), *bj = h->pair_j.as<uint32_t>();
  const int8_t *qd = h->pair_q.as<int8_t>();
  const uint32_t *bb = h->bucket_begin.as<uint32_t>(), *be = h->bucket_end.as<uint32_t>();
  if (n_pairs) {
    const uint64_t blocks = (n_pairs * 32 + 255) / 256;
    k_fast_pairs<D><<<(unsigned)blocks, 256, 0, h->stream>>>(g, h->pos_sorted, bb, be, bi, bj, qd, n_pairs, r2, count);
  }
  const uint64_t blocks2 = ((uint64_t)h->ncells * 32 + 255) / 256;
  k_fast_self<D><<<(unsigned)blocks2, 256, 0, h->stream>>>(h->pos_sorted, bb, be, (uint32_t)h->ncells, r2, count);
  h->launches += 2;
  ABR_CUDA(h, cudaGetLastError());
  return ABR_OK;
}

int run_fast_bucket_search_counts(Handle *h, double radius, uint32_t *count) {
  if (!h->built) return set_error(h, ABR_ERR_STATE, "fast bucket search: cell list has not been built");
  if (!h->pos_sorted && h->n_sorted > 0) return set_error(h, ABR_ERR_STATE, "fast bucket search: abr_query_set_particles not called");
  if (!count) return set_error(h, ABR_ERR_INVALID, "fast bucket search: null output");
  if (h->n_sorted == 0) return ABR_OK;
  uint64_t n_pairs = 0;
  int rc = run_bucket_pairs(h, nullptr, nullptr, nullptr, 0, &n_pairs);
  if (rc) return rc;
  if (n_pairs * 32 / 256 >= 0x7FFFFFFFull) return set_error(h, ABR_ERR_UNSUPPORTED, "fast bucket search: too many bucket pairs");
  ABR_CUDA(h, h->pair_i.reserve((n_pairs + 1) * sizeof(uint32_t)));
  ABR_CUDA(h, h->pair_j.reserve((n_pairs + 1) * sizeof(uint32_t)));
  ABR_CUDA(h, h->pair_q.reserve((n_pairs + 1) * (size_t)h->D));
  rc = run_bucket_pairs(h, h->pair_i.as<uint32_t>(), h->pair_j.as<uint32_t>(), h->pair_q.as<int8_t>(), n_pairs, &n_pairs);
  if (rc) return rc;
  fill_u32(h, count, 0u, h->n_sorted);
  switch (h->D) {
  case 1: return fast_counts_impl<1>(h, radius, count, n_pairs);
  case 2: return fast_counts_impl<2>(h, radius, count, n_pairs);
  default: return fast_counts_impl<3>(h, radius, count, n_pairs);
  }
}

} // namespace abr
