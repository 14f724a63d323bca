// Device-side view of the ordered cell list (the CellListOrderedQuery POD,
// /root/reference/src/CellListOrdered.h:285-599) and the exact per-row search
// walk (search_iterator + lattice_iterator_within_distance).
//
// Everything here must be compiled WITHOUT floating-point contraction
// (nvcc -fmad=false): the acceptance predicate ||dx||^2 <= r^2 and the bucket
// index arithmetic have to round exactly like the reference's un-fused host
// code.
#ifndef ABORIA_B200_DETAIL_GRID_CUH_
#define ABORIA_B200_DETAIL_GRID_CUH_

#include <stdint.h>

namespace abr {

constexpr int MAXD = 3;

// by-value kernel argument; mirrors the fields of CellListOrderedQuery
// (src/CellListOrdered.h:305-365) that the search path reads
struct Grid {
  int D;
  int size[MAXD];      // m_size
  int end[MAXD];       // m_end_bucket = m_size - 1
  int periodic[MAXD];  // m_periodic
  double bmin[MAXD], bmax[MAXD];
  double side[MAXD];      // m_bucket_side_length
  double inv_side[MAXD];  // 1.0 / side (src/detail/SpatialUtil.h:116)
  double L[MAXD];         // bmax - bmin (one rounding, as in Search.h:188-190)
  uint32_t ncells;        // number of buckets stored locally (= m_size.prod() without a window)
  uint32_t key_bound;     // keys of alive particles are < key_bound; dead = key_bound
  // Slab window along dimension 0 (multi-GPU).  The arithmetic above
  // always uses the GLOBAL grid; a rank stores only the bucket layers
  // win_lo .. win_lo+win_n-1 (unwrapped, may run past either end of a periodic
  // dimension) and computes rows only for its own layers own_lo .. own_lo+own_n-1
  // (local numbering).  Single GPU: win_lo = 0, win_n = size[0], own everything.
  int win_lo, win_n, own_lo, own_n;
};

// global layer index (dimension 0) -> local layer, or -1 when not stored here
__host__ __device__ inline int local_layer(const Grid &g, int v0) {
  int l = v0 - g.win_lo;
  if (l < 0) l += g.size[0];
  if (l >= g.size[0]) l -= g.size[0];
  return (l >= 0 && l < g.win_n) ? l : -1;
}

// collapse a GLOBAL bucket index vector to the LOCAL bucket number (-1: not stored)
template <int D> __host__ __device__ inline int local_collapse(const Grid &g, const int *v) {
  const int l = local_layer(g, v[0]);
  if (l < 0) return -1;
  int index = l;
#pragma unroll
  for (int i = 1; i < D; ++i) index = index * g.size[i] + v[i];
  return index;
}

struct Query {
  Grid g;
  const double *pos;            // sorted positions, n x D AoS
  const uint32_t *bucket_begin; // m_bucket_begin
  const uint32_t *bucket_end;   // m_bucket_end
  uint32_t n;
};

// src/detail/SpatialUtil.h:49-59 collapse_index_vector (last dim fastest,
// unsigned multiplier)
template <int D> __host__ __device__ inline int collapse_index(const Grid &g, const int *v) {
  int index = 0;
  unsigned int multiplier = 1;
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    if (i != D - 1) multiplier *= (unsigned)g.size[i + 1];
    index += multiplier * v[i];
  }
  return index;
}

// src/detail/SpatialUtil.h:118-131: floor((r - bmin) * inv_side) per dim
template <int D> __device__ inline int point_to_bucket(const Grid &g, const double *r) {
  int v[D];
#pragma unroll
  for (int d = 0; d < D; ++d) v[d] = (int)floor((r[d] - g.bmin[d]) * g.inv_side[d]);
  return collapse_index<D>(g, v);
}

// ---------------------------------------------------------------------------
// distance_helper<LNormNumber> (src/detail/Distance.h:46-139) for the norms that
// round identically on host and device: -1 Chebyshev (max |x|), 1 Manhattan
// (sum |x|), 2 Euclidean (sum x*x).  p >= 3 goes through std::pow in the
// reference and is not offered on the device.
// ---------------------------------------------------------------------------
template <int LN> struct DistHelper {
  __host__ __device__ static inline double value(double x) { return LN == 2 ? x * x : fabs(x); }
  __host__ __device__ static inline double accumulate(double acc, double v) {
    if (LN == -1) return v > acc ? v : acc;
    return acc + v;
  }
};

// ---------------------------------------------------------------------------
// lattice_iterator_within_distance<Query,LNormNumber,IdentityTransform>
// (src/NeighbourSearchBase.h:1720-2009), restated for the device.
// ---------------------------------------------------------------------------
// The Transform argument of the search iterators (src/Transform.h), by-value kernel argument.
//   TK = 0  IdentityTransform (no code at all)
//   TK = 1  ScaleTransform (:140-160): v -> v * s; a box's extent -> (bmax - bmin) * s
//   TK = 2  LinearTransform (:61-137) over a linear functor given as its D x D matrix (row
//           major, zero entries skipped, a 1.0 entry is a plain copy) with the "eigen vertex"
//           the constructor finds (:82-99, computed on the host)
struct Xform {
  double s[MAXD];
  double m[MAXD * MAXD];
  int eig[MAXD];
};
template <int D, int TK> __host__ __device__ inline void xform_point(const Xform &x, const double *v, double *out) {
  if (TK == 1) {
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] = v[i] * x.s[i];
  } else if (TK == 2) {
    double tmp[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      bool first = true;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const double c = x.m[i * D + j];
        if (c == 0.0) continue;
        const double t = c == 1.0 ? v[j] : c * v[j];
        acc = first ? t : acc + t;
        first = false;
      }
      tmp[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] = tmp[i];
  }
}
// extent of the axis-aligned box bounding the transformed box [bmin, bmax]
template <int D, int TK> __host__ __device__ inline void xform_box(const Xform &x, const double *bmin, const double *bmax, double *out) {
  if (TK == 1) {
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] = (bmax[i] - bmin[i]) * x.s[i];
  } else if (TK == 2) {
    double mx[D], mn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double centre = 0.5 * (bmax[i] + bmin[i]);
      mx[i] = (x.eig[i] ? bmax[i] : bmin[i]) - centre;
      mn[i] = (x.eig[i] ? bmin[i] : bmax[i]) - centre;
    }
    xform_point<D, TK>(x, mx, mx);
    xform_point<D, TK>(x, mn, mn);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double hi = fmax(mx[i], mn[i]), lo = fmin(mx[i], mn[i]);
      out[i] = hi - lo;
    }
  }
}

template <int D, int LN = 2, int TK = 0> struct BucketWalk {
  const Grid &g;
  const Xform *xf;
  double qp[D];
  double half[D];
  double r2;
  int quadrant;
  bool valid;
  int mn[D];
  int index[D];

  __device__ inline bool qbit(int i) const { return 1 == ((quadrant >> i) & 1); }

  // :1884-1892 with find_bucket_centre = (v + 0.5) * side + bmin
  __device__ inline double min_dist2(const int *b) const {
    double acc = 0;
    double dx[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double centre = ((double)b[i] + 0.5) * g.side[i] + g.bmin[i];
      dx[i] = centre - qp[i];
    }
    if (TK) xform_point<D, TK>(*xf, dx, dx); // m_transform(centre - m_query_point)
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double t = fmax(fabs(dx[i]) - half[i], 0.0);
      acc = DistHelper<LN>::accumulate(acc, DistHelper<LN>::value(t));
    }
    return acc;
  }
  // :1950-1958
  __device__ inline bool outside_domain() const {
    double acc = 0;
    double dx[D], ext[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      dx[i] = 0.5 * (g.bmin[i] + g.bmax[i]) - qp[i];
      ext[i] = g.bmax[i] - g.bmin[i];
    }
    if (TK) {
      xform_point<D, TK>(*xf, dx, dx);
      xform_box<D, TK>(*xf, g.bmin, g.bmax, ext);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double t = fmax(fabs(dx[i]) - 0.5 * ext[i], 0.0);
      acc = DistHelper<LN>::accumulate(acc, DistHelper<LN>::value(t));
    }
    return acc > r2;
  }
  // :1895-1947
  __device__ inline void reset_min_and_index() {
    bool no_buckets = true;
    while (valid && no_buckets) {
#pragma unroll
      for (int i = 0; i < D; ++i)
        mn[i] = (int)floor((qp[i] + (qbit(i) ? 0.5 : -0.5) * g.side[i] - g.bmin[i]) * g.inv_side[i]);
      no_buckets = min_dist2(mn) > r2;
      if (!no_buckets) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          if (qbit(i)) {
            if (mn[i] < 0) {
              mn[i] = 0;
            } else if (mn[i] > g.end[i]) {
              no_buckets = true;
              mn[i] = g.end[i];
            }
          } else {
            if (mn[i] < 0) {
              no_buckets = true;
              mn[i] = 0;
            } else if (mn[i] > g.end[i]) {
              mn[i] = g.end[i];
            }
          }
        }
      }
      if (no_buckets) {
        ++quadrant;
        if (quadrant >= (1 << D)) valid = false;
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) index[i] = mn[i];
      }
    }
  }
  // :1779-1804
  __device__ inline BucketWalk(const Grid &grid, const double *point, double R2, const Xform *xform = nullptr)
      : g(grid), xf(xform), r2(R2), quadrant(0), valid(true) {
#pragma unroll
    for (int i = 0; i < D; ++i) qp[i] = point[i];
    if (outside_domain()) {
      valid = false;
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) half[i] = 0.5 * g.side[i];
      if (TK) { // :1790-1799: 0.5 * m_transform(bbox(-0.5 side, 0.5 side))
        double lo[D], hi[D], ext[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          lo[i] = -0.5 * g.side[i];
          hi[i] = 0.5 * g.side[i];
        }
        xform_box<D, TK>(*xf, lo, hi, ext);
#pragma unroll
        for (int i = 0; i < D; ++i) half[i] = 0.5 * ext[i];
      }
      reset_min_and_index();
    }
  }
  // :1960-2004
  __device__ inline void increment() {
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
      bool potential = true;
      if (qbit(i)) {
        ++index[i];
        potential = index[i] <= g.end[i];
      } else {
        --index[i];
        potential = index[i] >= 0;
      }
      if (potential) potential = min_dist2(index) <= r2;
      if (potential) break;
      index[i] = mn[i];
      if (i == 0) {
        ++quadrant;
        if (quadrant < (1 << D)) {
          reset_min_and_index();
        } else {
          valid = false;
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// search_iterator<Query,LNormNumber> (src/Search.h:66-496) as a visitor: image
// lattice (last dim fastest, :152-159) x buckets near cur = r + image*L
// (:188-190) x particles of the bucket, accept iff !(norm(dx) > R2) (:438-446),
// R2 = get_value_to_accumulate(R).  LN = 2: euclidean_search, -1:
// chebyshev_search, 1: manhatten_search (src/Search.h:794-845).
// visit(j, dx, accumulated norm, image_linear_index)
// ---------------------------------------------------------------------------
template <int D, int LN = 2, int TK = 0, typename Visit>
__device__ inline void search_walk(const Query &q, const double *r, double R, Visit &&visit, const Xform *xform = nullptr) {
  const Grid &g = q.g;
  const double R2 = DistHelper<LN>::value(R);
  int img[D];
#pragma unroll
  for (int i = 0; i < D; ++i) img[i] = g.periodic[i] ? -1 : 0;
  int image_counter = 0;
  while (true) {
    double cur[D];
#pragma unroll
    for (int i = 0; i < D; ++i) cur[i] = r[i] + (double)img[i] * g.L[i];
    for (BucketWalk<D, LN, TK> b(g, cur, R2, xform); b.valid; b.increment()) {
      const int cl = local_collapse<D>(g, b.index);
      if (cl < 0) continue; // bucket layer held by another rank (never within reach of an owned row)
      const unsigned c = (unsigned)cl;
      const unsigned jb = q.bucket_begin[c], je = q.bucket_end[c];
      for (unsigned j = jb; j < je; ++j) {
        double dx[D];
        double acc = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) dx[i] = q.pos[(size_t)j * D + i] - cur[i];
        if (TK) xform_point<D, TK>(*xform, dx, dx); // m_dx = m_transform(p - m_current_point), src/Search.h:443
#pragma unroll
        for (int i = 0; i < D; ++i) acc = DistHelper<LN>::accumulate(acc, DistHelper<LN>::value(dx[i]));
        if (!(acc > R2)) visit(j, dx, acc, image_counter);
      }
    }
    ++image_counter;
    int i = D - 1;
    for (; i >= 0; --i) {
      const int hi = g.periodic[i] ? 2 : 1;
      if (++img[i] < hi) break;
      img[i] = g.periodic[i] ? -1 : 0;
    }
    if (i < 0) break;
  }
}

__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

} // namespace abr
#endif
