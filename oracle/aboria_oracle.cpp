// ============================================================================
// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement (C++17 + OpenMP, no Boost / Eigen) of the Aboria reference
// functions on the hot path "ordered cell-list build -> sparse kernel matvec".
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product (libabr.so) never
// does.
//
// The real reference cannot be compiled here (needs Boost >=1.50 and Eigen 3.3,
// neither present) so this file follows the reference line by
// line, in the same expression order, and is compiled with
//     g++ -O2 -fopenmp -ffp-contract=off
// (reference flags are plain -std=c++14, i.e. no FMA contraction).
//
// Parity pinning: the known-answer vectors of the reference's own tests
// (tests/operators.h:810-967, tests/utils.h:52-93, tests/iterators.h:49-211,
// tests/neighbours.h:520-686 with cases :1252-1260, brute force :739-764) are
// checked against this oracle in tests/test_oracle_golden.py.  The within-cell
// particle order produced by the reference's CPU std::sort is NOT pinned by any
// reference test ("parity unpinned" for that one property) —
// comparisons are therefore made by particle id / per-cell id sets.
//
// Every function cites the reference file:line (under /root/reference) it
// restates.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int MAXD = 4;

// ---------------------------------------------------------------------------
// src/detail/SpatialUtil.h:36-59  bucket_index<D>::collapse_index_vector
// (row-major, last dimension fastest; multiplier is unsigned)
// ---------------------------------------------------------------------------
inline int collapse_index_vector(int D, const unsigned *size, const int *v) {
  int index = 0;
  unsigned int multiplier = 1;
  for (int i = D - 1; i >= 0; --i) {
    if (i != D - 1) multiplier *= size[i + 1];
    index += multiplier * v[i];
  }
  return index;
}

struct Oracle {
  int D = 3;
  // neighbour_search_base state (src/NeighbourSearchBase.h:252-268)
  double bmin[MAXD], bmax[MAXD];
  bool periodic[MAXD];
  double n_leaf = 10.0;
  bool domain_has_been_set = false;
  // CellListOrdered state (src/CellListOrdered.h:272-283)
  unsigned size[MAXD];
  double side[MAXD], inv_side[MAXD];
  int end_bucket[MAXD];
  size_t size_calculated_with_n = std::numeric_limits<size_t>::max();
  std::vector<unsigned> bucket_begin, bucket_end, bucket_indices;
  std::vector<int> alive_sum, alive_indices;
  // query view of the (reordered) column particle set
  const double *pos = nullptr; // n x D AoS
  size_t n = 0;

  // src/detail/SpatialUtil.h:118-125 find_bucket_index_vector
  // floor((r - bmin) * inv_side) cast to int, per dimension
  inline void find_bucket_index_vector(const double *r, int *v) const {
    for (int d = 0; d < D; ++d)
      v[d] = static_cast<int>(std::floor((r[d] - bmin[d]) * inv_side[d]));
  }
  // src/detail/SpatialUtil.h:129-131 find_bucket_index
  inline int find_bucket_index(const double *r) const {
    int v[MAXD];
    find_bucket_index_vector(r, v);
    return collapse_index_vector(D, size, v);
  }
  // src/detail/SpatialUtil.h:150-156 get_min_index_by_quadrant
  inline int get_min_index_by_quadrant(double r, int i, bool up) const {
    return static_cast<int>(
        std::floor((r + (up ? 0.5 : -0.5) * side[i] - bmin[i]) * inv_side[i]));
  }

  // src/CellListOrdered.h:132-186 set_domain_impl
  bool set_domain_impl() {
    const size_t nn = alive_indices.size();
    if (nn < 0.5 * size_calculated_with_n || nn > 2 * size_calculated_with_n) {
      size_calculated_with_n = nn;
      if (n_leaf > nn) {
        for (int d = 0; d < D; ++d) size[d] = 1;
      } else {
        double total_volume = 1.0; // Vector::prod(): ret=1; ret*=mem[i]
        for (int d = 0; d < D; ++d) total_volume *= (bmax[d] - bmin[d]);
        const double box_volume = n_leaf / double(nn) * total_volume;
        const double box_side_length = std::pow(box_volume, 1.0 / D);
        for (int d = 0; d < D; ++d) {
          size[d] = static_cast<unsigned int>(
              std::floor((bmax[d] - bmin[d]) / box_side_length));
          if (size[d] == 0) size[d] = 1;
        }
      }
      size_t prod = 1;
      for (int d = 0; d < D; ++d) {
        side[d] = (bmax[d] - bmin[d]) / size[d];
        inv_side[d] = 1.0 / side[d]; // SpatialUtil.h:116
        end_bucket[d] = int(size[d]) - 1;
        prod *= size[d];
      }
      bucket_begin.resize(prod);
      bucket_end.resize(prod);
      return true;
    }
    return false;
  }
};

// ---------------------------------------------------------------------------
// src/detail/Distance.h:46-139 distance_helper<LNormNumber>: -1 = Chebyshev
// (max), 1 = Manhattan, 2 = Euclidean (pow(x,2) == x*x), p >= 3 via
// std::pow as the reference does.
// ---------------------------------------------------------------------------
template <int LN> struct dist_helper {
  static inline double value(double x) {
    if (LN == -1 || LN == 1) return std::abs(x);
    if (LN == 0) return x != 0;
    if (LN == 2) return x * x;
    if (LN == 4) return std::pow(x, LN);
    return std::abs(std::pow(x, LN));
  }
  static inline double accumulate(double accum, double v) {
    if (LN == -1) return v > accum ? v : accum;
    return accum + v;
  }
};

// ---------------------------------------------------------------------------
// src/NeighbourSearchBase.h:1720-2009 lattice_iterator_within_distance<Query,
// LNormNumber,IdentityTransform>; enumerates buckets near a point exactly as the
// reference does (quadrant by quadrant, row-wise with early exit).
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// src/Transform.h: the Transform argument of the search iterators.
//   kind 1  ScaleTransform (:140-160): v -> v * scale; box -> (bmax - bmin) * scale
//   kind 2  LinearTransform (:61-137) over a LINEAR user functor, given here as its matrix
//           (row major; zero entries are skipped so that e.g. the reference tests'
//           SkewTransform `v[0] + 0.3 * v[1]`, tests/neighbours.h:1262-1267, is reproduced
//           operation for operation); the box transform uses the "eigen vertex" found by the
//           constructor (:82-99).
// ---------------------------------------------------------------------------
struct Xform {
  int kind = 0;
  int D = 0;
  double s[MAXD];
  double m[MAXD * MAXD];
  bool eig[MAXD];
  inline void point(const double *v, double *out) const {
    if (kind == 1) {
      for (int i = 0; i < D; ++i) out[i] = v[i] * s[i];
    } else {
      double tmp[MAXD];
      for (int i = 0; i < D; ++i) {
        bool first = true;
        double acc = 0.0;
        for (int j = 0; j < D; ++j) {
          const double c = m[i * D + j];
          if (c == 0.0) continue;
          const double t = c == 1.0 ? v[j] : c * v[j];
          acc = first ? t : acc + t;
          first = false;
        }
        tmp[i] = acc;
      }
      for (int i = 0; i < D; ++i) out[i] = tmp[i];
    }
  }
  // side lengths of the axis-aligned box bounding the transformed box
  inline void box(const double *bmin, const double *bmax, double *out) const {
    if (kind == 1) {
      for (int i = 0; i < D; ++i) out[i] = (bmax[i] - bmin[i]) * s[i];
      return;
    }
    double mx[MAXD], mn[MAXD]; // :112-136
    for (int i = 0; i < D; ++i) {
      const double centre = 0.5 * (bmax[i] + bmin[i]);
      if (eig[i]) {
        mx[i] = bmax[i] - centre;
        mn[i] = bmin[i] - centre;
      } else {
        mx[i] = bmin[i] - centre;
        mn[i] = bmax[i] - centre;
      }
    }
    point(mx, mx);
    point(mn, mn);
    for (int i = 0; i < D; ++i) {
      const double tmp = std::max(mx[i], mn[i]);
      mn[i] = std::min(mx[i], mn[i]);
      mx[i] = tmp;
      out[i] = mx[i] - mn[i];
    }
  }
  // LinearTransform constructor (:82-99): the vertex of [-1,1]^D whose image is longest
  void find_eigen_vertices() {
    double best = 0;
    for (int i = 0; i < D; ++i) eig[i] = false;
    for (int code = 0; code < (1 << D); ++code) { // lattice_iterator over {0,1}^D, last dimension fastest
      double pt[MAXD], q[MAXD];
      bool bits[MAXD];
      for (int j = 0; j < D; ++j) {
        bits[j] = (code >> (D - 1 - j)) & 1;
        pt[j] = bits[j] ? 1.0 : -1.0;
      }
      point(pt, q);
      double n2 = 0;
      for (int j = 0; j < D; ++j) n2 += q[j] * q[j];
      if (n2 > best) {
        for (int j = 0; j < D; ++j) eig[j] = bits[j];
        best = n2;
      }
    }
  }
};

template <int LN = 2> struct BucketIter {
  const Oracle *q;
  int D;
  double query_point[MAXD];
  double half_bucket_length[MAXD];
  double max_distance2;
  int quadrant = 0;
  bool valid = true;
  int mn[MAXD];
  int index[MAXD];
  const Xform *xf = nullptr; // null = IdentityTransform

  inline bool ith_quadrant_bit(int i) const { return 1 == ((quadrant >> i) & 1); }

  // :1884-1892 get_min_distance_to_bucket; find_bucket_centre is
  // (vindex + 0.5) * side + bmin  (src/detail/SpatialUtil.h:144-147)
  inline double get_min_distance_to_bucket(const int *bucket) const {
    double dx[MAXD];
    for (int i = 0; i < D; ++i) {
      const double centre = (bucket[i] + 0.5) * q->side[i] + q->bmin[i];
      dx[i] = centre - query_point[i];
    }
    if (xf) xf->point(dx, dx); // m_transform(centre - m_query_point)
    for (int i = 0; i < D; ++i)
      dx[i] = std::max(std::abs(dx[i]) - half_bucket_length[i], 0.0);
    double accum = 0; // src/detail/Distance.h:131-138
    for (int i = 0; i < D; ++i) accum = dist_helper<LN>::accumulate(accum, dist_helper<LN>::value(dx[i]));
    return accum;
  }

  // :1950-1958 outside_domain
  inline bool outside_domain(const double *position) const {
    double dx[MAXD], ext[MAXD];
    for (int i = 0; i < D; ++i) {
      dx[i] = 0.5 * (q->bmin[i] + q->bmax[i]) - position[i];
      ext[i] = q->bmax[i] - q->bmin[i];
    }
    if (xf) {
      xf->point(dx, dx);
      xf->box(q->bmin, q->bmax, ext);
    }
    for (int i = 0; i < D; ++i) {
      const double half_domain_side_length = 0.5 * ext[i];
      dx[i] = std::max(std::abs(dx[i]) - half_domain_side_length, 0.0);
    }
    double accum = 0;
    for (int i = 0; i < D; ++i) accum = dist_helper<LN>::accumulate(accum, dist_helper<LN>::value(dx[i]));
    return accum > max_distance2;
  }

  // :1895-1947 reset_min_and_index
  void reset_min_and_index() {
    bool no_buckets = true;
    while (valid && no_buckets) {
      for (int i = 0; i < D; ++i)
        mn[i] = q->get_min_index_by_quadrant(query_point[i], i, ith_quadrant_bit(i));
      const double accum = get_min_distance_to_bucket(mn);
      no_buckets = accum > max_distance2;
      if (!no_buckets) {
        for (int i = 0; i < D; i++) {
          if (ith_quadrant_bit(i)) {
            if (mn[i] < 0) {
              mn[i] = 0;
            } else if (mn[i] > q->end_bucket[i]) {
              no_buckets = true;
              mn[i] = q->end_bucket[i];
            }
          } else {
            if (mn[i] < 0) {
              no_buckets = true;
              mn[i] = 0;
            } else if (mn[i] > q->end_bucket[i]) {
              mn[i] = q->end_bucket[i];
            }
          }
        }
      }
      if (no_buckets) {
        ++quadrant;
        if (quadrant >= (1 << D)) valid = false;
      } else {
        for (int i = 0; i < D; ++i) index[i] = mn[i];
      }
    }
  }

  // :1779-1804 constructor
  BucketIter(const Oracle *query, const double *point, double max_distance, const Xform *xf_ = nullptr)
      : q(query), D(query->D), xf(xf_) {
    for (int i = 0; i < D; ++i) query_point[i] = point[i];
    max_distance2 = dist_helper<LN>::value(max_distance); // pow(x,2) -> x*x (§0.5)
    if (outside_domain(point)) {
      valid = false;
    } else {
      // :1790-1799: identity: 0.5 * side; otherwise 0.5 * transform(bbox(-0.5 side, 0.5 side))
      if (xf) {
        double lo[MAXD], hi[MAXD], ext[MAXD];
        for (int i = 0; i < D; ++i) {
          lo[i] = -0.5 * q->side[i];
          hi[i] = 0.5 * q->side[i];
        }
        xf->box(lo, hi, ext);
        for (int i = 0; i < D; ++i) half_bucket_length[i] = 0.5 * ext[i];
      } else {
        for (int i = 0; i < D; ++i) half_bucket_length[i] = 0.5 * q->side[i];
      }
      reset_min_and_index();
    }
  }

  // :1960-2004 increment
  void increment() {
    for (int i = D - 1; i >= 0; --i) {
      bool potential_bucket = true;
      if (ith_quadrant_bit(i)) {
        ++index[i];
        potential_bucket = index[i] <= q->end_bucket[i];
      } else {
        --index[i];
        potential_bucket = index[i] >= 0;
      }
      if (potential_bucket) {
        const double accum = get_min_distance_to_bucket(index);
        potential_bucket = accum <= max_distance2;
      }
      if (potential_bucket) break;
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
// src/Search.h:66-496 search_iterator<Query,2> as a visitor: for each periodic
// image (lattice_iterator, last dim fastest; src/LatticeIterator.h:269-280 and
// src/Search.h:152-159), for each bucket near the shifted point, for each
// particle in the bucket range: dx = p_j - cur; accept iff !(sum dx^2 > r^2)
// (src/Search.h:438-446).  Calls visit(j, dx, image_linear_index).
// ---------------------------------------------------------------------------
template <int LN, typename Visit>
inline void distance_search(const Oracle &q, const double *r, double max_distance,
                            Visit &&visit, const Xform *xf = nullptr) {
  const int D = q.D;
  const double max_distance2 = dist_helper<LN>::value(max_distance);
  int start[MAXD], end[MAXD], img[MAXD];
  for (int i = 0; i < D; ++i) {
    start[i] = q.periodic[i] ? -1 : 0;
    end[i] = q.periodic[i] ? 2 : 1;
    img[i] = start[i];
  }
  bool images_left = true;
  int image_counter = 0;
  while (images_left) {
    double cur[MAXD];
    for (int i = 0; i < D; ++i)
      cur[i] = r[i] + img[i] * (q.bmax[i] - q.bmin[i]); // Search.h:188-190
    for (BucketIter<LN> b(&q, cur, max_distance, xf); b.valid; b.increment()) {
      const unsigned c = (unsigned)collapse_index_vector(D, q.size, b.index);
      const unsigned jb = q.bucket_begin[c], je = q.bucket_end[c];
      for (unsigned j = jb; j < je; ++j) {
        double dx[MAXD];
        double accum = 0;
        for (int i = 0; i < D; ++i) dx[i] = q.pos[(size_t)j * D + i] - cur[i];
        if (xf) xf->point(dx, dx); // m_dx = m_transform(p - m_current_point), src/Search.h:443
        for (int i = 0; i < D; ++i) accum = dist_helper<LN>::accumulate(accum, dist_helper<LN>::value(dx[i]));
        if (!(accum > max_distance2)) visit(j, dx, image_counter);
      }
    }
    // lattice_iterator increment, last dimension fastest
    ++image_counter;
    int i = D - 1;
    for (; i >= 0; --i) {
      if (++img[i] < end[i]) break;
      img[i] = start[i];
    }
    if (i < 0) images_left = false;
  }
}

// euclidean_search (src/Search.h:839-845) = distance_search<2>
template <typename Visit>
inline void euclidean_search(const Oracle &q, const double *r, double max_distance, Visit &&visit) {
  distance_search<2>(q, r, max_distance, visit);
}

// ---------------------------------------------------------------------------
// Kernel functions.  These are the *user lambdas* of the reference's tests and
// examples, restated with the std:: calls the reference uses.  ids are mirrored
// (by value) in include/abr.h; there is no shared code with the product.
// ---------------------------------------------------------------------------
enum {
  K_CONST_SUM = 0,      // tests/operators.h:842-847  s1(a)+s2(b)
  K_CONST_SUM_DIFF = 1, // tests/operators.h:905-911  (s1(a)+s2(b), s1(a)-s2(b)) 2x1
  K_INV_DIST = 2,       // 1/(norm(dx)+eps)
  K_INV_DIST_AA = 3,    // tests/operators.h:251-256  a_i*a_j/(norm(dx)+eps)
  K_WENDLAND_C2 = 4,    // tests/rbf_interpolation.h:310-313
  K_LJ_FORCE = 5,       // SURVEY §8d c3 (md.h pattern tests/md.h:166-174) Dx1
  K_SPH_DENSITY = 6,    // tests/sph.h:154-165 W_fun, times mass
  K_SPH_PRESSURE = 7,   // tests/sph.h:140-152 F_fun; m(Pa/ra^2+Pb/rb^2) F dx, Dx1
  K_LINEAR_SPRING = 8,  // tests/md.h:166-174  -k (diameter/r - 1) dx for r != 0, Dx1
};

struct KernelCtx {
  int id;
  int D;
  const double *params;
  const double *const *row_vars;
  const double *const *col_vars;
};

inline double norm_of(const double *dx, int D) {
  double ret = 0; // src/Vector.h:313-324
  for (int i = 0; i < D; ++i) ret += dx[i] * dx[i];
  return std::sqrt(ret);
}

// writes a BRxBC block (row-major) for the pair (i,j)
inline void eval_kernel(const KernelCtx &k, const double *dx, size_t i, size_t j,
                        double *blk) {
  const int D = k.D;
  switch (k.id) {
  case K_CONST_SUM:
    blk[0] = k.row_vars[0][i] + k.col_vars[0][j];
    break;
  case K_CONST_SUM_DIFF:
    blk[0] = k.row_vars[0][i] + k.col_vars[0][j];
    blk[1] = k.row_vars[0][i] - k.col_vars[0][j];
    break;
  case K_INV_DIST:
    blk[0] = 1.0 / (norm_of(dx, D) + k.params[0]);
    break;
  case K_INV_DIST_AA:
    blk[0] = (k.row_vars[0][i] * k.col_vars[0][j]) / (norm_of(dx, D) + k.params[0]);
    break;
  case K_WENDLAND_C2: {
    const double h = k.params[0];
    blk[0] = std::pow(2.0 - norm_of(dx, D) / h, 4) * (1.0 + 2.0 * norm_of(dx, D) / h);
    break;
  }
  case K_LJ_FORCE: {
    // params: sigma, epsilon.  f = 24 eps (2 (s/r)^12 - (s/r)^6) / r^2 * dx, 0 at r==0
    const double sigma = k.params[0], eps = k.params[1];
    const double r = norm_of(dx, D);
    if (r == 0) {
      for (int d = 0; d < D; ++d) blk[d] = 0.0;
    } else {
      const double sr = sigma / r;
      const double sr6 = std::pow(sr, 6);
      const double f = 24.0 * eps * (2.0 * sr6 * sr6 - sr6) / (r * r);
      for (int d = 0; d < D; ++d) blk[d] = f * dx[d];
    }
    break;
  }
  case K_LINEAR_SPRING: {
    // params: k, diameter.  tests/md.h:166-174: if (r != 0) sum += -k * (diameter / r - 1.0) * dx
    const double kk = k.params[0], diameter = k.params[1];
    const double r = norm_of(dx, D);
    for (int d = 0; d < D; ++d) blk[d] = (r != 0) ? -kk * (diameter / r - 1.0) * dx[d] : 0.0;
    break;
  }
  case K_SPH_DENSITY: {
    // params: h, mass, wcon.  tests/sph.h:154-165
    const double h = k.params[0], mass = k.params[1], wcon = k.params[2];
    const double r = norm_of(dx, D);
    const double q = r / h;
    double W = 0.0;
    if (q <= 2.0)
      W = (1 / std::pow(h, D)) * wcon * std::pow(2.0 - q, 4) * (1.0 + 2.0 * q);
    blk[0] = mass * W;
    break;
  }
  case K_SPH_PRESSURE: {
    // params: h, mass, wcon; row_vars[0]=pdr2 (P/rho^2) of rows, col_vars[0] of cols
    const double h = k.params[0], mass = k.params[1], wcon = k.params[2];
    const double r = norm_of(dx, D);
    double F = 0.0;
    if (r != 0) {
      const double q = r / h;
      if (q <= 2.0)
        F = (1 / std::pow(h, D + 2)) * wcon *
            (-4 * std::pow(2 - q, 3) * (1 + 2 * q) + 2 * std::pow(2 - q, 4)) / q;
    }
    const double c = mass * (k.row_vars[0][i] + k.col_vars[0][j]) * F;
    for (int d = 0; d < D; ++d) blk[d] = c * dx[d];
    break;
  }
  default:
    blk[0] = 0.0;
  }
}

inline uint64_t mix64(uint64_t x) { // splitmix64 finaliser, for pair-set hashes
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

} // namespace

extern "C" {

void *orc_create(int D) {
  if (D < 1 || D > MAXD) return nullptr;
  Oracle *o = new Oracle();
  o->D = D;
  return o;
}
void orc_destroy(void *h) { delete static_cast<Oracle *>(h); }

// src/detail/SpatialUtil.h:49-59 (KAT: tests/utils.h:52-74)
int orc_collapse_index_vector(int D, const unsigned *size, const int *v) {
  return collapse_index_vector(D, size, v);
}

// src/NeighbourSearchBase.h:252-268 set_domain (+ CellListOrdered::set_domain_impl)
void orc_set_domain(void *h, const double *bmin, const double *bmax,
                    const uint8_t *periodic, double n_leaf) {
  Oracle *o = static_cast<Oracle *>(h);
  o->domain_has_been_set = true;
  for (int d = 0; d < o->D; ++d) {
    o->bmin[d] = bmin[d];
    o->bmax[d] = bmax[d];
    o->periodic[d] = periodic[d] != 0;
  }
  o->n_leaf = n_leaf;
  o->set_domain_impl();
}

// direct grid override used by KATs that construct point_to_bucket_index by
// hand (tests/utils.h:76-93) or need a known grid (tests/iterators.h uses
// CellList with 10 buckets per dim)
void orc_force_grid(void *h, const double *bmin, const double *bmax,
                    const uint8_t *periodic, const unsigned *size) {
  Oracle *o = static_cast<Oracle *>(h);
  o->domain_has_been_set = true;
  size_t prod = 1;
  for (int d = 0; d < o->D; ++d) {
    o->bmin[d] = bmin[d];
    o->bmax[d] = bmax[d];
    o->periodic[d] = periodic[d] != 0;
    o->size[d] = size[d];
    o->side[d] = (bmax[d] - bmin[d]) / size[d];
    o->inv_side[d] = 1.0 / o->side[d];
    o->end_bucket[d] = int(size[d]) - 1;
    prod *= size[d];
  }
  o->bucket_begin.assign(prod, 0);
  o->bucket_end.assign(prod, 0);
}

void orc_get_grid(void *h, unsigned *size, double *side) {
  Oracle *o = static_cast<Oracle *>(h);
  for (int d = 0; d < o->D; ++d) {
    size[d] = o->size[d];
    side[d] = o->side[d];
  }
}

int orc_point_to_bucket_index(void *h, const double *r, int *vindex) {
  Oracle *o = static_cast<Oracle *>(h);
  o->find_bucket_index_vector(r, vindex);
  return collapse_index_vector(o->D, o->size, vindex);
}

// ---------------------------------------------------------------------------
// src/NeighbourSearchBase.h:350-495 update_positions, restricted to the ordered
// case (update range == whole set), followed by
// src/CellListOrdered.h:190-259 update_positions_impl.
//   pos   : n x D, wrapped in place (enforce_domain_lambda, :185-238)
//   alive : n, cleared in place for killed particles
//   sort_mode 0: std::sort with a key-only comparator (detail/Algorithms.h:173-182)
//   sort_mode 1: stable (what thrust::sort_by_key / the CUDA build produce)
// returns number of alive particles; alive_indices/bucket arrays via getters.
// ---------------------------------------------------------------------------
long orc_update_positions(void *h, double *pos, uint8_t *alive, size_t n,
                          int sort_mode) {
  Oracle *o = static_cast<Oracle *>(h);
  const int D = o->D;
  if (n == 0) return 0; // :375-376

  // enforce domain (:380-385, lambda :208-237)
  if (o->domain_has_been_set) {
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; ++p) {
      double r[MAXD];
      for (int d = 0; d < D; ++d) r[d] = pos[p * D + d];
      for (int d = 0; d < D; ++d) {
        if (!std::isfinite(r[d])) {
          alive[p] = uint8_t(false);
        } else if (o->periodic[d]) {
          while (r[d] < o->bmin[d]) r[d] += (o->bmax[d] - o->bmin[d]);
          while (r[d] >= o->bmax[d]) r[d] -= (o->bmax[d] - o->bmin[d]);
        } else {
          if ((r[d] < o->bmin[d]) || (r[d] >= o->bmax[d])) alive[p] = uint8_t(false);
        }
      }
      for (int d = 0; d < D; ++d) pos[p * D + d] = r[d];
    }
  }

  // exclusive scan of alive (:393), num_dead (:394-397)
  o->alive_sum.resize(n);
  int running = 0;
  for (size_t p = 0; p < n; ++p) {
    o->alive_sum[p] = running;
    running += static_cast<int>(alive[p]);
  }
  const int num_alive = o->alive_sum.back() + static_cast<int>(alive[n - 1]);
  const int num_dead = int(n) - num_alive;

  // scatter_if (:417-434)
  o->alive_indices.resize(n - num_dead);
  for (size_t p = 0; p < n; ++p)
    if (alive[p]) o->alive_indices[o->alive_sum[p]] = int(p);

  // CellListOrdered::update_positions_impl (src/CellListOrdered.h:190-259)
  o->set_domain_impl();
  const size_t na = o->alive_indices.size();
  o->bucket_indices.resize(na);
  if (na > 0) {
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < na; ++k)
      o->bucket_indices[k] =
          (unsigned)o->find_bucket_index(pos + (size_t)o->alive_indices[k] * D);

    // sort_by_key(bucket_indices, alive_indices)
    std::vector<std::pair<unsigned, int>> zip(na);
    for (size_t k = 0; k < na; ++k) zip[k] = {o->bucket_indices[k], o->alive_indices[k]};
    auto key_less = [](const std::pair<unsigned, int> &a,
                       const std::pair<unsigned, int> &b) { return a.first < b.first; };
    if (sort_mode == 0)
      std::sort(zip.begin(), zip.end(), key_less);
    else
      std::stable_sort(zip.begin(), zip.end(), key_less);
    for (size_t k = 0; k < na; ++k) {
      o->bucket_indices[k] = zip[k].first;
      o->alive_indices[k] = zip[k].second;
    }
  }
  // lower_bound / upper_bound of 0..C-1 in the sorted keys (:229-239)
  const size_t C = o->bucket_begin.size();
#pragma omp parallel for schedule(static)
  for (size_t c = 0; c < C; ++c) {
    o->bucket_begin[c] = (unsigned)(std::lower_bound(o->bucket_indices.begin(),
                                                     o->bucket_indices.end(), (unsigned)c) -
                                    o->bucket_indices.begin());
    o->bucket_end[c] = (unsigned)(std::upper_bound(o->bucket_indices.begin(),
                                                   o->bucket_indices.end(), (unsigned)c) -
                                  o->bucket_indices.begin());
  }
  return (long)na;
}

size_t orc_num_buckets(void *h) { return static_cast<Oracle *>(h)->bucket_begin.size(); }
void orc_get_build(void *h, int *alive_indices, unsigned *bucket_indices,
                   unsigned *bucket_begin, unsigned *bucket_end) {
  Oracle *o = static_cast<Oracle *>(h);
  if (alive_indices)
    std::memcpy(alive_indices, o->alive_indices.data(), o->alive_indices.size() * sizeof(int));
  if (bucket_indices)
    std::memcpy(bucket_indices, o->bucket_indices.data(),
                o->bucket_indices.size() * sizeof(unsigned));
  if (bucket_begin)
    std::memcpy(bucket_begin, o->bucket_begin.data(), o->bucket_begin.size() * sizeof(unsigned));
  if (bucket_end)
    std::memcpy(bucket_end, o->bucket_end.data(), o->bucket_end.size() * sizeof(unsigned));
}

// src/Particles.h:694-724 reorder -> detail::gather (detail/Algorithms.h:718-745)
// for one column: dst[k] = src[order[k]]
void orc_gather(const int *order, size_t n_out, const void *src, void *dst,
                size_t elem_bytes) {
  const char *s = static_cast<const char *>(src);
  char *d = static_cast<char *>(dst);
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < n_out; ++k)
    std::memcpy(d + k * elem_bytes, s + (size_t)order[k] * elem_bytes, elem_bytes);
}

// point the query at the reordered positions (src/NeighbourSearchBase.h:504-512)
void orc_update_iterators(void *h, const double *pos_sorted, size_t n) {
  Oracle *o = static_cast<Oracle *>(h);
  o->pos = pos_sorted;
  o->n = n;
}

// get_buckets_near_point<2> (src/CellListOrdered.h:527-536): writes up to max
// D-tuples of bucket indices, returns the number the iterator yields.
long orc_buckets_near_point(void *h, const double *point, double max_distance,
                            int *out, long max_out) {
  Oracle *o = static_cast<Oracle *>(h);
  long count = 0;
  for (BucketIter<2> b(o, point, max_distance); b.valid; b.increment()) {
    if (out && count < max_out)
      for (int d = 0; d < o->D; ++d) out[count * o->D + d] = b.index[d];
    ++count;
  }
  return count;
}

// euclidean_search from one point (src/Search.h:839-845): returns count,
// writes up to max_out (j, image) and dx
long orc_search_point(void *h, const double *r, double max_distance, int *out_j,
                      int *out_image, double *out_dx, long max_out) {
  Oracle *o = static_cast<Oracle *>(h);
  long count = 0;
  euclidean_search(*o, r, max_distance, [&](unsigned j, const double *dx, int image) {
    if (count < max_out) {
      if (out_j) out_j[count] = (int)j;
      if (out_image) out_image[count] = image;
      if (out_dx)
        for (int d = 0; d < o->D; ++d) out_dx[count * o->D + d] = dx[d];
    }
    ++count;
  });
  return count;
}

// per-row neighbour count and order-independent pair-set hash
// hash_i = sum_j mix64(j * 27 * 3 + image) (mod 2^64)
void orc_pair_stats(void *h, const double *row_pos, size_t n_rows, double radius,
                    const double *radius_per_row, uint32_t *count, uint64_t *hash) {
  Oracle *o = static_cast<Oracle *>(h);
  const int D = o->D;
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < n_rows; ++i) {
    uint32_t c = 0;
    uint64_t hs = 0;
    const double R = radius_per_row ? radius_per_row[i] : radius;
    euclidean_search(*o, row_pos + i * D, R, [&](unsigned j, const double *, int image) {
      ++c;
      hs += mix64((uint64_t)j * 81u + (uint64_t)image);
    });
    if (count) count[i] = c;
    if (hash) hash[i] = hs;
  }
}

// distance_search<LNormNumber> / chebyshev_search / manhatten_search
// (src/Search.h:794-831): per-row neighbour count and pair-set hash for the
// norms -1 (Chebyshev), 1 (Manhattan), 2 (Euclidean), 3, 4
int orc_pair_stats_norm_scaled(void *h, const double *row_pos, size_t n_rows, double radius, int lnorm, const double *scale,
                               uint32_t *count, uint64_t *hash);
int orc_pair_stats_norm(void *h, const double *row_pos, size_t n_rows, double radius, int lnorm, uint32_t *count,
                        uint64_t *hash) {
  return orc_pair_stats_norm_scaled(h, row_pos, n_rows, radius, lnorm, nullptr, count, hash);
}
// the same with a ScaleTransform (create_scale_transform, src/Transform.h:140-172; used as
// euclidean_search(query, centre, 1.0, create_scale_transform(1/radius)) in tests/neighbours.h:553-561)
int orc_pair_stats_norm_xform(void *h, const double *row_pos, size_t n_rows, double radius, int lnorm, const Xform *xf, uint32_t *count,
                              uint64_t *hash);
int orc_pair_stats_norm_scaled(void *h, const double *row_pos, size_t n_rows, double radius, int lnorm, const double *scale,
                               uint32_t *count, uint64_t *hash) {
  Oracle *o = static_cast<Oracle *>(h);
  Xform xf;
  xf.kind = 1;
  xf.D = o->D;
  if (scale)
    for (int d = 0; d < o->D; ++d) xf.s[d] = scale[d];
  return orc_pair_stats_norm_xform(h, row_pos, n_rows, radius, lnorm, scale ? &xf : nullptr, count, hash);
}
// create_linear_transform<D>(functor) for a linear functor given as its D x D matrix (row major)
int orc_pair_stats_norm_linear(void *h, const double *row_pos, size_t n_rows, double radius, int lnorm, const double *matrix,
                               uint32_t *count, uint64_t *hash) {
  Oracle *o = static_cast<Oracle *>(h);
  Xform xf;
  xf.kind = 2;
  xf.D = o->D;
  for (int e = 0; e < o->D * o->D; ++e) xf.m[e] = matrix[e];
  xf.find_eigen_vertices();
  return orc_pair_stats_norm_xform(h, row_pos, n_rows, radius, lnorm, &xf, count, hash);
}
int orc_pair_stats_norm_xform(void *h, const double *row_pos, size_t n_rows, double radius, int lnorm, const Xform *scale, uint32_t *count,
                              uint64_t *hash) {
  Oracle *o = static_cast<Oracle *>(h);
  const int D = o->D;
  if (lnorm != -1 && lnorm != 1 && lnorm != 2 && lnorm != 3 && lnorm != 4) return 1;
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < n_rows; ++i) {
    uint32_t c = 0;
    uint64_t hs = 0;
    auto visit = [&](unsigned j, const double *, int image) {
      ++c;
      hs += mix64((uint64_t)j * 81u + (uint64_t)image);
    };
    switch (lnorm) {
    case -1: distance_search<-1>(*o, row_pos + i * D, radius, visit, scale); break;
    case 1: distance_search<1>(*o, row_pos + i * D, radius, visit, scale); break;
    case 3: distance_search<3>(*o, row_pos + i * D, radius, visit, scale); break;
    case 4: distance_search<4>(*o, row_pos + i * D, radius, visit, scale); break;
    default: distance_search<2>(*o, row_pos + i * D, radius, visit, scale); break;
    }
    if (count) count[i] = c;
    if (hash) hash[i] = hs;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// src/Kernels.h:720-751 KernelSparse::evaluate (Eigen overload): OpenMP
// parallel for over rows; lhs.segment<BR>(i*BR) += F(dx,a_i,b_j) *
// rhs.segment<BC>(j*BC).  lhs is ACCUMULATED into (Eigen zeroes it first for
// y = K*b; src/detail/Operators.h:219-232).
// returns the total number of accepted pairs.
// ---------------------------------------------------------------------------
uint64_t orc_sparse_matvec(void *h, const double *row_pos, size_t n_rows,
                           int kernel_id, const double *params,
                           const double *const *row_vars,
                           const double *const *col_vars, double radius,
                           const double *radius_per_row, int BR, int BC,
                           const double *rhs, double *lhs, int nthreads) {
  Oracle *o = static_cast<Oracle *>(h);
  const int D = o->D;
  KernelCtx k{kernel_id, D, params, row_vars, col_vars};
  uint64_t total = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : total)
  for (size_t i = 0; i < n_rows; ++i) {
    const double R = radius_per_row ? radius_per_row[i] : radius;
    euclidean_search(*o, row_pos + i * D, R, [&](unsigned j, const double *dx, int) {
      double blk[MAXD * MAXD];
      eval_kernel(k, dx, i, j, blk);
      for (int p = 0; p < BR; ++p) {
        double s = 0; // fixed-size Eigen product: sum over q of blk(p,q)*rhs(q)
        for (int q = 0; q < BC; ++q) s += blk[p * BC + q] * rhs[(size_t)j * BC + q];
        lhs[i * BR + p] += s;
      }
      ++total;
    });
  }
  return total;
}

// src/Kernels.h:653-685 KernelSparse::assemble(std::vector<Triplet>&): for every
// row i (serial loop), for every search hit j: one BRxBC block.  Emitted as CSR
// (row_ptr over block entries, col = particle index j, values row-major blocks),
// entries of a row in the iterator's own order.  Pass col_idx == NULL to count.
uint64_t orc_sparse_assemble(void *h, const double *row_pos, size_t n_rows, int kernel_id, const double *params,
                             const double *const *row_vars, const double *const *col_vars, double radius,
                             const double *radius_per_row, int BR, int BC, uint32_t *row_ptr, int32_t *col_idx,
                             double *values) {
  Oracle *o = static_cast<Oracle *>(h);
  const int D = o->D;
  KernelCtx k{kernel_id, D, params, row_vars, col_vars};
  uint64_t nnz = 0;
  for (size_t i = 0; i < n_rows; ++i) {
    row_ptr[i] = (uint32_t)nnz;
    const double R = radius_per_row ? radius_per_row[i] : radius;
    euclidean_search(*o, row_pos + i * D, R, [&](unsigned j, const double *dx, int) {
      if (col_idx) {
        col_idx[nnz] = (int32_t)j;
        if (values) eval_kernel(k, dx, i, j, values + nnz * (size_t)(BR * BC));
      }
      ++nnz;
    });
  }
  row_ptr[n_rows] = (uint32_t)nnz;
  return nnz;
}

// ---------------------------------------------------------------------------
// src/NeighbourSearchBase.h:440-486 update of the id map for the ordered case
// (update range == whole set): m_id_map_value = sequence(0..n), m_id_map_key[k] =
// id of the particle at (post-reorder) position k, then sort_by_key(key, value)
// (detail/Algorithms.h:173-182: std::sort on the zipped pair with a key-only
// comparator).  ids are unique in the reference (src/Particles.h next_id), so the
// result does not depend on the stability of the sort.
// ---------------------------------------------------------------------------
void orc_id_map_build(const uint64_t *ids, size_t n, uint64_t *key, uint64_t *value) {
  std::vector<std::pair<uint64_t, uint64_t>> kv(n);
  for (size_t k = 0; k < n; ++k) kv[k] = {ids[k], (uint64_t)k};
  std::sort(kv.begin(), kv.end(), [](const std::pair<uint64_t, uint64_t> &a, const std::pair<uint64_t, uint64_t> &b) { return a.first < b.first; });
  for (size_t k = 0; k < n; ++k) {
    key[k] = kv[k].first;
    value[k] = kv[k].second;
  }
}

// src/CellListOrdered.h:379-388 find(id): lower_bound over m_id_map_key; returns the
// particle index m_id_map_value[...] or n (the reference's "end" pointer) when absent
void orc_id_find(const uint64_t *key, const uint64_t *value, size_t n, const uint64_t *query, size_t m, uint64_t *index_out) {
  for (size_t q = 0; q < m; ++q) {
    const uint64_t id = query[q];
    const uint64_t *first = std::lower_bound(key, key + n, id);
    index_out[q] = (first != key + n && !(id < *first)) ? value[first - key] : (uint64_t)n;
  }
}

// ---------------------------------------------------------------------------
// KernelBase::coeff(i, j) (src/Kernels.h:102-112) over detail::sparse_kernel
// (src/detail/Kernels.h:336-367): pi = i / BR, pj = j / BC;
// dx = correct_dx_for_periodicity(pos_col[pj] - pos_row[pi]) (src/Particles.h:480-494:
// per periodic dimension `while (dx > w/2) dx -= w; while (dx <= -w/2) dx += w`);
// block = dx.squaredNorm() < pow(radius(a), 2) ? F(dx, a, b) : 0   -- STRICT '<', unlike
// the search predicate of the product; returns block(i % BR, j % BC).
// ---------------------------------------------------------------------------
void orc_sparse_coeff(void *h, const double *row_pos, const double *col_pos, const uint64_t *ii, const uint64_t *jj, size_t m,
                      int kernel_id, const double *params, const double *const *row_vars, const double *const *col_vars,
                      double radius, const double *radius_per_row, int BR, int BC, double *out) {
  Oracle *o = static_cast<Oracle *>(h);
  const int D = o->D;
  KernelCtx k{kernel_id, D, params, row_vars, col_vars};
  for (size_t q = 0; q < m; ++q) {
    const size_t pi = ii[q] / BR, ioff = ii[q] - pi * BR;
    const size_t pj = jj[q] / BC, joff = jj[q] - pj * BC;
    double dx[MAXD];
    for (int d = 0; d < D; ++d) {
      dx[d] = col_pos[pj * D + d] - row_pos[pi * D + d];
      if (o->periodic[d]) {
        const double w = o->bmax[d] - o->bmin[d];
        while (dx[d] > w / 2) dx[d] -= w;
        while (dx[d] <= -w / 2) dx[d] += w;
      }
    }
    double n2 = 0; // src/Vector.h squaredNorm: sequential sum from 0
    for (int d = 0; d < D; ++d) n2 += dx[d] * dx[d];
    const double R = radius_per_row ? radius_per_row[pi] : radius;
    double blk[MAXD * MAXD];
    for (int e = 0; e < BR * BC; ++e) blk[e] = 0.0;
    if (n2 < R * R) eval_kernel(k, dx, pi, pj, blk);
    out[q] = blk[ioff * BC + joff];
  }
}

// ---------------------------------------------------------------------------
// get_neighbouring_buckets(query) -> bucket_pair_iterator (src/Search.h:498-764,
// :857-860), the "fast cell-list search" of tests/neighbours.h:281-300.
// Restated literally: the iterator state (m_periodic, m_i, m_j, m_domain_domain) and
// its increment(), on top of lattice_iterator (src/LatticeIterator.h:48-291, last
// dimension fastest).
// ---------------------------------------------------------------------------
namespace {
struct LatticeIt {
  int D = 0;
  int mn[MAXD], mx[MAXD], idx[MAXD];
  bool valid = false;
  LatticeIt() {}
  LatticeIt(int D_, const int *a, const int *b) : D(D_), valid(true) { // [a, b)
    for (int d = 0; d < D; ++d) {
      mn[d] = a[d];
      mx[d] = b[d];
      idx[d] = a[d];
    }
  }
  void set(const int *v) { // operator=(const int_d&), LatticeIterator.h:118-123
    valid = true;
    for (int d = 0; d < D; ++d) {
      idx[d] = v[d];
      valid = valid && v[d] >= mn[d] && v[d] < mx[d];
    }
  }
  void inc() { // LatticeIterator.h:269-280
    for (int i = D - 1; i >= 0; --i) {
      ++idx[i];
      if (idx[i] < mx[i]) break;
      if (i != 0) {
        idx[i] = mn[i];
      } else {
        valid = false;
      }
    }
  }
  bool next_is_end() const { // (it + 1) == false
    LatticeIt t = *this;
    t.inc();
    return !t.valid;
  }
};

// src/Search.h:675-707 get_neighbouring_buckets(query, bucket)
LatticeIt neighbouring_buckets(const Oracle &q, const int *bucket) {
  int start[MAXD], end[MAXD];
  bool no_buckets = false;
  for (int i = 0; i < q.D; ++i) {
    start[i] = bucket[i] - 1;
    end[i] = bucket[i] + 1;
    if (start[i] < 0) {
      start[i] = 0;
    } else if (start[i] > q.end_bucket[i]) {
      no_buckets = true;
      start[i] = q.end_bucket[i];
    }
    if (end[i] < 0) {
      no_buckets = true;
      end[i] = 0;
    } else if (end[i] > q.end_bucket[i]) {
      end[i] = q.end_bucket[i];
    }
  }
  if (no_buckets) return LatticeIt();
  int endp1[MAXD];
  for (int i = 0; i < q.D; ++i) endp1[i] = end[i] + 1;
  return LatticeIt(q.D, start, endp1);
}
// :709-720 get_neighbouring_buckets(query, bucket, quadrant): bucket + quadrant * (end_bucket + 1)
LatticeIt neighbouring_buckets_q(const Oracle &q, const int *bucket, const int *quadrant) {
  int b[MAXD];
  for (int i = 0; i < q.D; ++i) b[i] = bucket[i] + quadrant[i] * (q.end_bucket[i] + 1);
  return neighbouring_buckets(q, b);
}
// :722-751 get_regular_buckets(query, quadrant)
LatticeIt regular_buckets(const Oracle &q, const int *quadrant) {
  int start[MAXD], endp1[MAXD];
  for (int i = 0; i < q.D; ++i) {
    start[i] = 0;
    int end = q.end_bucket[i];
    if (q.periodic[i]) {
      if (quadrant[i] > 0) {
        start[i] = 0;
        end = 0;
      } else if (quadrant[i] < 0) {
        start[i] = q.end_bucket[i];
        end = q.end_bucket[i];
      }
    }
    endp1[i] = end + 1;
  }
  return LatticeIt(q.D, start, endp1);
}
} // namespace

// Walks the whole iterator; per step writes the collapsed bucket numbers of i and j and
// the periodic quadrant (D ints in -1..1; m_position_offset = quadrant * (bmax - bmin)).
// Pass null outputs to count.
uint64_t orc_bucket_pairs(void *h, uint32_t *bucket_i, uint32_t *bucket_j, int8_t *quadrant, uint64_t capacity) {
  const Oracle &q = *static_cast<Oracle *>(h);
  const int D = q.D;
  // constructor, :541-575
  int pstart[MAXD], pend[MAXD], zero[MAXD];
  for (int i = 0; i < D; ++i) {
    pstart[i] = q.periodic[i] ? -1 : 0;
    pend[i] = q.periodic[i] ? 2 : 1;
    zero[i] = 0;
  }
  LatticeIt m_periodic(D, pstart, pend);
  m_periodic.set(zero);
  bool m_valid = true, m_domain_domain = true;
  LatticeIt m_i = regular_buckets(q, m_periodic.idx), m_j;
  if (m_i.next_is_end()) {
    m_domain_domain = false;
    m_periodic.inc();
    if (!m_periodic.valid) {
      m_valid = false;
    } else {
      m_i = regular_buckets(q, m_periodic.idx);
      m_j = neighbouring_buckets_q(q, m_i.idx, m_periodic.idx);
    }
  } else {
    m_j = neighbouring_buckets_q(q, m_i.idx, m_periodic.idx);
    m_j.set(m_i.idx);
    m_j.inc();
  }
  uint64_t count = 0;
  while (m_valid) {
    if (bucket_i && count < capacity) {
      bucket_i[count] = (uint32_t)collapse_index_vector(D, q.size, m_i.idx);
      bucket_j[count] = (uint32_t)collapse_index_vector(D, q.size, m_j.idx);
      for (int d = 0; d < D; ++d) quadrant[count * D + d] = (int8_t)m_periodic.idx[d];
    }
    ++count;
    // increment(), :627-667
    m_j.inc();
    if (!m_j.valid) {
      m_i.inc();
      if (m_domain_domain ? m_i.next_is_end() : !m_i.valid) {
        m_domain_domain = false;
        m_periodic.inc();
        if (!m_periodic.valid) {
          m_valid = false;
        } else {
          m_i = regular_buckets(q, m_periodic.idx);
          m_j = neighbouring_buckets_q(q, m_i.idx, m_periodic.idx);
        }
      } else {
        m_j = neighbouring_buckets_q(q, m_i.idx, m_periodic.idx);
        if (m_domain_domain) {
          m_j.set(m_i.idx);
          m_j.inc();
        }
      }
    }
  }
  return count;
}

// The fast cell-list search of tests/neighbours.h:892-951 (the user-side loops the
// reference documents for this iterator): per particle the number of neighbours with
// |p_i + offset - p_j|^2 < r^2 (STRICT), each unordered pair found once and counted for
// both particles, plus the self count.
void orc_fast_bucket_search_counts(void *h, double radius, uint32_t *count) {
  const Oracle &q = *static_cast<Oracle *>(h);
  const int D = q.D;
  const double r2 = radius * radius;
  for (size_t i = 0; i < q.n; ++i) count[i] = 0;
  const uint64_t np = orc_bucket_pairs(h, nullptr, nullptr, nullptr, 0);
  std::vector<uint32_t> bi(np), bj(np);
  std::vector<int8_t> qd(np * D);
  orc_bucket_pairs(h, bi.data(), bj.data(), qd.data(), np);
  for (uint64_t k = 0; k < np; ++k) {
    double off[MAXD];
    for (int d = 0; d < D; ++d) off[d] = qd[k * D + d] * (q.bmax[d] - q.bmin[d]);
    for (unsigned a = q.bucket_begin[bi[k]]; a < q.bucket_end[bi[k]]; ++a) {
      double pa[MAXD];
      for (int d = 0; d < D; ++d) pa[d] = q.pos[(size_t)a * D + d] + off[d];
      for (unsigned b = q.bucket_begin[bj[k]]; b < q.bucket_end[bj[k]]; ++b) {
        double n2 = 0;
        for (int d = 0; d < D; ++d) {
          const double t = pa[d] - q.pos[(size_t)b * D + d];
          n2 += t * t;
        }
        if (n2 < r2) {
          count[a]++;
          count[b]++;
        }
      }
    }
  }
  for (size_t c = 0; c < q.bucket_begin.size(); ++c) {
    for (unsigned a = q.bucket_begin[c]; a < q.bucket_end[c]; ++a) {
      count[a]++; // self is a neighbour
      for (unsigned b = a + 1; b < q.bucket_end[c]; ++b) {
        double n2 = 0;
        for (int d = 0; d < D; ++d) {
          const double t = q.pos[(size_t)a * D + d] - q.pos[(size_t)b * D + d];
          n2 += t * t;
        }
        if (n2 < r2) {
          count[a]++;
          count[b]++;
        }
      }
    }
  }
}

// brute force of tests/neighbours.h:739-764: counts j with
// squaredNorm(pj - pi - image*(max-min)) <= r2 over 3^D images (periodic) or
// the single image (non periodic).  NOTE the operation order differs from the
// search iterator (this is the reference's *test* oracle, used as a property).
void orc_brute_force_counts(int D, const double *pos, size_t n, const double *bmin,
                            const double *bmax, int is_periodic, double r2,
                            uint32_t *count) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    uint32_t c = 0;
    for (size_t j = 0; j < n; ++j) {
      if (is_periodic) {
        int img[MAXD];
        for (int d = 0; d < D; ++d) img[d] = -1;
        while (true) {
          double s = 0;
          for (int d = 0; d < D; ++d) {
            const double dx = pos[j * D + d] - pos[i * D + d] - img[d] * (bmax[d] - bmin[d]);
            s += dx * dx;
          }
          if (s <= r2) c++;
          int d = D - 1;
          for (; d >= 0; --d) {
            if (++img[d] < 2) break;
            img[d] = -1;
          }
          if (d < 0) break;
        }
      } else {
        double s = 0;
        for (int d = 0; d < D; ++d) {
          const double dx = pos[j * D + d] - pos[i * D + d];
          s += dx * dx;
        }
        if (s <= r2) c++;
      }
    }
    count[i] = c;
  }
}

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Sets the OpenMP team size for every later parallel region of the oracle.  bench.py calls
// it with the number of host cores it may use: launchers such as torchrun export
// OMP_NUM_THREADS=1, which would silently turn the timed CPU arm into a single-thread run.
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

} // extern "C"
