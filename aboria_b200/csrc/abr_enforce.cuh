// enforce_domain_lambda (/root/reference/src/NeighbourSearchBase.h:208-237) followed by
// point_to_bucket_index (src/detail/SpatialUtil.h:118-131) for one particle: shared by the
// key kernels of abr_build.cu (radix builds) and abr_build2.cu (counting-sort build).
// Compiled with -fmad=false (see grid.cuh).
#ifndef ABR_ENFORCE_CUH_
#define ABR_ENFORCE_CUH_
#include "abr_internal.h"

namespace abr {

// the per-particle part: returns the bucket key (key_bound for a dead particle)
template <int D, bool WINDOWED>
__device__ __forceinline__ uint32_t enforce_one(double *__restrict__ pos, uint8_t *__restrict__ alive, uint32_t p, const Grid &g,
                                                DevScalars *sc) {
  double r[D], r0[D];
#pragma unroll
  for (int d = 0; d < D; ++d) r0[d] = r[d] = pos[(size_t)p * D + d];
  const uint8_t a0 = alive[p];
  uint8_t a = a0;
  // the common case first: a particle already inside [bmin, bmax) in every dimension (NaN and infinities
  // fail the comparisons) is neither wrapped nor killed nor written back — 2 D comparisons instead of the
  // general path's finite test, wrap loops and write-back checks
  bool inside = true;
#pragma unroll
  for (int d = 0; d < D; ++d) inside = inside && (r[d] >= g.bmin[d]) && (r[d] < g.bmax[d]);
  if (!inside) {
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (!isfinite(r[d])) {
      a = 0;
    } else if (g.periodic[d]) {
      // The reference loops without bound (and never terminates once |r|/L
      // exceeds 2^53).  A GPU kernel must not hang: after 2^20 steps the
      // particle is killed like a non-finite one (documented deviation).
      int guard = 0;
      while (r[d] < g.bmin[d] && ++guard < (1 << 20)) r[d] += (g.bmax[d] - g.bmin[d]);
      while (r[d] >= g.bmax[d] && ++guard < (1 << 20)) r[d] -= (g.bmax[d] - g.bmin[d]);
      if (guard >= (1 << 20)) {
        a = 0;
        r[d] = r0[d];
      }
    } else {
      if ((r[d] < g.bmin[d]) || (r[d] >= g.bmax[d])) a = 0;
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
    if (__double_as_longlong(r[d]) != __double_as_longlong(r0[d])) pos[(size_t)p * D + d] = r[d];
  if (a != a0) alive[p] = a;
  }
  uint32_t key = g.key_bound;
  if (a) {
    int v[D];
    bool overflow = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      v[d] = (int)floor((r[d] - g.bmin[d]) * g.inv_side[d]);
      overflow |= (v[d] >= g.size[d]) | (v[d] < 0);
    }
    if (WINDOWED) {
      const int cl = overflow ? -1 : local_collapse<D>(g, v);
      if (cl < 0) {
        atomicAdd(&sc->n_outside, 1u); // not this rank's particle: reported as an error by the host
        key = g.key_bound;
      } else {
        key = (uint32_t)cl;
      }
    } else {
      key = (uint32_t)collapse_index<D>(g, v);
      if (overflow) atomicAdd(&sc->n_aliased, 1u);
      if (key >= g.key_bound) key = g.key_bound - 1; // cannot happen (v[d] <= size[d]); keeps the sort in range
    }
  }
  return key;
}


} // namespace abr
#endif
