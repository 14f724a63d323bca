// Sparse kernel operator product y += K b for sm_100a
// (KernelSparse::evaluate, /root/reference/src/Kernels.h:720-751).
//
// Two kernels, both templated on the dimension and on a device functor F
// (the user's lambda f(dx, a_i, b_j), src/detail/Kernels.h:336-357):
//
//  * tiled_kernel  — the hot path when the row set is the column set
//    (create_sparse_operator(p, p, r, f)).  One warp per target bucket; the
//    candidates are the contiguous particle runs of the neighbouring buckets
//    (last dimension contiguous in the sorted array), 32 candidates per step
//    (one per lane, positions and b_j in registers), the rows of the target
//    bucket are broadcast from shared memory.  The acceptance test is the
//    reference's exact un-fused fp64 predicate.  Partial sums live in a
//    per-warp shared-memory table part[row][lane] and are reduced once per
//    bucket in a fixed order (deterministic results).
//    Rows whose result could depend on rounding in the reference's bucket
//    iterator (a coordinate within `tolf` of a bucket face / centre, or an
//    accepted pair within `r2lo..r2` of the cut-off) are NOT finished here:
//    they are appended to a list and recomputed by walk_kernel, which restates
//    the reference iterator exactly.  So the pair set is the reference's, always.
//
//  * walk_kernel — one thread per row, the reference's search_iterator walk
//    (grid.cuh).  Used for arbitrary row sets, per-row radii, the listed rows
//    above, and whenever the tiled preconditions do not hold.
//
// No tensor cores: this is a gather + pointwise fp64 path (north_star).
#ifndef ABORIA_B200_DETAIL_MATVEC_KERNELS_CUH_
#define ABORIA_B200_DETAIL_MATVEC_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "aboria_b200/detail/grid.cuh"

// opaque in abr.h
struct abr_matvec_plan {
  abr::Query q;
  // rows
  const double *row_pos;
  uint32_t n_rows;
  int rows_are_cols;
  double radius;
  const double *radius_per_row;
  const double *b;
  double *y;
  // tiled path
  const double *posb;      // packed (x, y, z, b) per column particle, 32-byte records (b only when BC == 1)
  int use_tiled;
  int w[abr::MAXD];        // stencil half width per dimension
  uint32_t grab;           // consecutive buckets a warp claims per scheduler step (1..8)
  int trim;                // some w >= 2: trim the stencil by distance (nothing to trim when all w == 1)
  double r2, r2lo;         // cut-off^2 and the "rounding sensitive" lower edge
  float pre_r2;            // fp32 pre-filter threshold: r2 * (1 + tol), never rejects a pair the exact test accepts
  double tolf[abr::MAXD];  // fractional bucket coordinate tolerance
  uint32_t *work_counter;
  uint32_t *danger_count;
  uint32_t *danger_list;
  uint32_t danger_capacity;
  // stats (pair_stats): when non-null the kernels count/hash instead of evaluating F
  uint32_t *stat_count;
  uint64_t *stat_hash;
  // launch
  cudaStream_t stream;
  int sm_count;
  int walk_only_list; // walk kernel processes danger_list instead of all rows
  // tiled path with a row set that is NOT the column set (create_sparse_operator(test, knots, ...),
  // tests/rbf_interpolation.h:326): the row points bucketed into the column grid by an internal build
  const double *xrow_pos;     // row positions sorted by column-grid bucket (null: rows are the columns)
  const uint32_t *xrow_bb, *xrow_be; // per bucket: range of sorted rows
  const int32_t *xrow_index;  // original index of sorted row k (y, row columns and the exact-walk list use it)
  const uint8_t *xrow_alive;  // 0: the row point was dropped by the row build (outside the domain): exact walk
  // heavy buckets (clustered clouds): (bucket, first batch number) items appended by the first launch, consumed by the second
  uint2 *heavy_list;
  uint32_t heavy_capacity;
  unsigned long long *heavy_state; // items << 40 | row batches
  uint32_t *heavy_work;
  int heavy_phase;    // 1: the second launch (consumes the heavy list)
  int symmetric;      // tiled path: half stencil + y[j] scatter for functors that declare SYMMETRY (see tiled_kernel)
  double *ytmp;       //   zeroed scratch the symmetric kernel accumulates into (n_rows * BR)
  uint32_t *row_bits; //   one bit per row: result comes from the exact walk, not from ytmp
  int variant;        // tiled path: 0 = tiled_kernel (gathers from L2), 1 = staged_kernel (bulk-copy staging in shared memory)
};

namespace abr {

constexpr int TILED_WARPS = 4;
constexpr int TILED_THREADS = TILED_WARPS * 32;

// functor used by pair_stats; never evaluated
struct StatsFunctor {
  static constexpr int BR = 1, BC = 1;
  __device__ void operator()(const double *, double, uint32_t, uint32_t, double *blk) const { blk[0] = 0; }
};

template <int D> __device__ inline int image_linear_index(const Grid &g, const int *img) {
  // position of `img` in the reference's periodic lattice_iterator
  // (src/Search.h:152-159; last dimension fastest)
  int idx = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) idx = idx * (g.periodic[d] ? 3 : 1) + (g.periodic[d] ? img[d] + 1 : 0);
  return idx;
}

// ---------------------------------------------------------------------------
// walk_kernel
// ---------------------------------------------------------------------------
template <int D, class F, bool STATS>
__global__ void __launch_bounds__(128) walk_kernel(const abr_matvec_plan p, const F f) {
  constexpr int BR = F::BR, BC = F::BC;
  const uint32_t total = p.walk_only_list ? min(*p.danger_count, p.danger_capacity) : p.n_rows;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t i = p.walk_only_list ? p.danger_list[t] : t;
    double r[D];
#pragma unroll
    for (int d = 0; d < D; ++d) r[d] = p.row_pos[(size_t)i * D + d];
    const double R = p.radius_per_row ? p.radius_per_row[i] : p.radius;
    if (STATS) {
      uint32_t cnt = 0;
      uint64_t hs = 0;
      search_walk<D, 2>(p.q, r, R, [&](unsigned j, const double *, double, int image) {
        ++cnt;
        hs += mix64((uint64_t)j * 81u + (uint64_t)image);
      });
      if (p.stat_count) p.stat_count[i] = cnt;
      if (p.stat_hash) p.stat_hash[i] = hs;
    } else {
      double acc[BR];
#pragma unroll
      for (int a = 0; a < BR; ++a) acc[a] = p.y[(size_t)i * BR + a];
      search_walk<D, 2>(p.q, r, R, [&](unsigned j, const double *dx, double d2, int) {
        double blk[BR * BC];
        f(dx, d2, i, j, blk);
#pragma unroll
        for (int a = 0; a < BR; ++a) {
          double s = 0;
#pragma unroll
          for (int c = 0; c < BC; ++c) s += blk[a * BC + c] * p.b[(size_t)j * BC + c];
          acc[a] += s;
        }
      });
#pragma unroll
      for (int a = 0; a < BR; ++a) p.y[(size_t)i * BR + a] = acc[a];
    }
  }
}

// ---------------------------------------------------------------------------
// norm_stats_kernel: distance_search<LN> / chebyshev_search / manhatten_search
// (src/Search.h:794-831) — per-row neighbour count and pair-set hash
// ---------------------------------------------------------------------------
template <int D, int LN, int TK = 0>
__global__ void __launch_bounds__(128) norm_stats_kernel(const abr_matvec_plan p, const Xform xform = Xform()) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n_rows; i += gridDim.x * blockDim.x) {
    double r[D];
#pragma unroll
    for (int d = 0; d < D; ++d) r[d] = p.row_pos[(size_t)i * D + d];
    const double R = p.radius_per_row ? p.radius_per_row[i] : p.radius;
    uint32_t cnt = 0;
    uint64_t hs = 0;
    search_walk<D, LN, TK>(p.q, r, R, [&](unsigned j, const double *, double, int image) {
      ++cnt;
      hs += mix64((uint64_t)j * 81u + (uint64_t)image);
    }, &xform);
    if (p.stat_count) p.stat_count[i] = cnt;
    if (p.stat_hash) p.stat_hash[i] = hs;
  }
}

// ---------------------------------------------------------------------------
// assemble_kernel: KernelSparse::assemble to triplets (src/Kernels.h:653-685) as
// CSR.  One thread per row walks the reference iterator, so the entries of a row
// appear in exactly the order the reference pushes its triplets.
// ---------------------------------------------------------------------------
template <int D, class F>
__global__ void __launch_bounds__(128) assemble_kernel(const abr_matvec_plan p, const F f, const uint32_t *__restrict__ row_ptr,
                                                      int32_t *__restrict__ col_idx, double *__restrict__ values) {
  constexpr int BR = F::BR, BC = F::BC;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n_rows; i += gridDim.x * blockDim.x) {
    double r[D];
#pragma unroll
    for (int d = 0; d < D; ++d) r[d] = p.row_pos[(size_t)i * D + d];
    const double R = p.radius_per_row ? p.radius_per_row[i] : p.radius;
    uint32_t k = row_ptr[i];
    const uint32_t kend = row_ptr[i + 1];
    search_walk<D, 2>(p.q, r, R, [&](unsigned j, const double *dx, double d2, int) {
      if (k < kend) {
        col_idx[k] = (int32_t)j;
        if (values) {
          double blk[BR * BC];
          f(dx, d2, i, j, blk);
#pragma unroll
          for (int e = 0; e < BR * BC; ++e) values[(size_t)k * (BR * BC) + e] = blk[e];
        }
      }
      ++k;
    });
  }
}

template <int D, class F> inline int launch_assemble(const abr_matvec_plan &p, const F &f, const uint32_t *row_ptr, int32_t *col_idx,
                                                     double *values) {
  const unsigned grid = (unsigned)((p.n_rows + 127) / 128);
  if (grid > 0) assemble_kernel<D, F><<<grid, 128, 0, p.stream>>>(p, f, row_ptr, col_idx, values);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// coeff_kernel: KernelBase::coeff(i, j) over detail::sparse_kernel
// (src/Kernels.h:102-112, src/detail/Kernels.h:336-367) for m (i, j) entries:
// dx = correct_dx_for_periodicity(p_col - p_row) (src/Particles.h:480-494), the
// block is F(dx, a, b) when dx.squaredNorm() < r^2 — STRICT, unlike the search
// predicate — and zero otherwise.
// ---------------------------------------------------------------------------
template <int D, class F>
__global__ void __launch_bounds__(128) coeff_kernel(const abr_matvec_plan p, const F f, const uint64_t *__restrict__ ii,
                                                   const uint64_t *__restrict__ jj, uint64_t m, double *__restrict__ out) {
  constexpr int BR = F::BR, BC = F::BC;
  const Grid &g = p.q.g;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t pi = ii[q] / BR, pj = jj[q] / BC;
    const int ioff = (int)(ii[q] - pi * BR), joff = (int)(jj[q] - pj * BC);
    if (pi >= p.n_rows || pj >= p.q.n) { // the reference ASSERTs (src/Kernels.h:103-104); through the C ABI: a zero entry, no wild read
      out[q] = 0.0;
      continue;
    }
    double dx[D];
    double n2 = 0;
    bool finite = true;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double v = p.q.pos[pj * D + d] - p.row_pos[pi * D + d];
      if (!isfinite(v)) { // a non-finite row point would never leave the wrap loops: a GPU kernel must not hang
        finite = false;
        v = 0.0;
      }
      if (g.periodic[d]) {
        const double w = g.bmax[d] - g.bmin[d];
        int guard = 0;
        while (v > w / 2 && ++guard < (1 << 20)) v -= w;
        while (v <= -w / 2 && ++guard < (1 << 20)) v += w;
        if (guard >= (1 << 20)) finite = false;
      }
      dx[d] = v;
      n2 += v * v;
    }
    if (!finite) {
      out[q] = 0.0;
      continue;
    }
    const double R = p.radius_per_row ? p.radius_per_row[pi] : p.radius;
    double val = 0.0;
    if (n2 < R * R) {
      double blk[BR * BC];
      f(dx, n2, (uint32_t)pi, (uint32_t)pj, blk);
      val = blk[ioff * BC + joff];
    }
    out[q] = val;
  }
}

template <int D, class F> inline int launch_coeff(const abr_matvec_plan &p, const F &f, const uint64_t *ii, const uint64_t *jj, uint64_t m,
                                                  double *out) {
  const unsigned grid = (unsigned)((m + 127) / 128 < 65535u * 16u ? (m + 127) / 128 : 65535u * 16u);
  if (grid > 0) coeff_kernel<D, F><<<grid, 128, 0, p.stream>>>(p, f, ii, jj, m, out);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// tiled_kernel
// ---------------------------------------------------------------------------
// does the functor read dx?  (default yes; functors that only need |dx|^2 set
// `static constexpr bool NEEDS_DX = false` and save three subtractions per pair)
template <class F, class = void> struct needs_dx { static constexpr bool value = true; };
template <class F> struct needs_dx<F, decltype((void)F::NEEDS_DX)> { static constexpr bool value = F::NEEDS_DX; };

// Symmetry of the kernel function under exchange of the two particles, declared by the functor:
//   static constexpr int SYMMETRY = +1   block(-dx, b, a) == +block(dx, a, b)   (1/(r+eps), Wendland, SPH density ...)
//   static constexpr int SYMMETRY = -1   block(-dx, b, a) == -block(dx, a, b)   (central forces: LJ, linear spring)
// default 0: no assumption.  With a symmetric functor and rows == columns the product can evaluate every
// unordered pair ONCE (half stencil, "fast cell-list search" of tests/neighbours.h:281-300 /
// src/Search.h:498-764) and add the result to both rows.
template <class F, class = void> struct symmetry { static constexpr int value = 0; };
template <class F> struct symmetry<F, decltype((void)F::SYMMETRY)> { static constexpr int value = F::SYMMETRY; };

#ifndef ABR_QDRAIN
#define ABR_QDRAIN 14
#endif
#ifndef ABR_TILED_CTAS
#define ABR_TILED_CTAS 7
#endif
constexpr int QDRAIN = ABR_QDRAIN;    // drain when any lane holds this many (a test step adds <= 4)
constexpr int QCAP = QDRAIN - 1 + 4;  // most accepted pairs a lane can hold
constexpr int ROW_BITS = 4;
#ifndef ABR_SKEW_LANES
#define ABR_SKEW_LANES 1
#endif
#ifndef ABR_WQ
#define ABR_WQ 256
#endif
constexpr int WQ = ABR_WQ;            // pairs compacted per drain pass
constexpr uint32_t HEAVY_ROWS = 64;   // a bucket with more rows than this is split into row-batch work items (second launch)
// partial-sum table update in the drain (profiles/r2w_acc_modes.txt):
//   0  16 columns, the two half-warps update one after the other (default)
//   1  16 columns, one pass: lanes l and l+16 share a column; when they also hold the same row the lower
//      lane adds both terms (two shuffles), otherwise the two addresses differ
//   2  32 columns, one pass (2 KB more shared memory per warp: one or two resident CTAs fewer per SM)
#ifndef ABR_ACC_MODE
#define ABR_ACC_MODE 0
#endif
constexpr int PCOL = ABR_ACC_MODE == 2 ? 32 : 16; // columns of the partial-sum table

// resident CTAs per SM the tiled kernel is compiled for: 7 (72 registers) by default; a
// functor whose math is light enough for 64 registers declares
// `static constexpr int TILED_CTAS = 8` (measured: InvDist +2 %, SphDensity -10 % at 8)
template <class F, class = void> struct tiled_ctas { static constexpr int value = ABR_TILED_CTAS; };
#ifdef ABR_FORCE_CTAS
template <class F> struct tiled_ctas<F, decltype((void)F::TILED_CTAS)> { static constexpr int value = ABR_FORCE_CTAS; };
#else
template <class F> struct tiled_ctas<F, decltype((void)F::TILED_CTAS)> { static constexpr int value = F::TILED_CTAS; };
#endif

template <int D, class F, bool STATS> struct TiledCfg {
  static constexpr int NACC = STATS ? 2 : F::BR;
  static constexpr int CTAS = NACC == 1 ? tiled_ctas<F>::value : 5;
  // rows per batch: 16 (buckets hold ~n_particles_in_leaf = 10 rows, so one sweep over the
  // candidates serves the whole bucket); block kernels keep BR partial-sum tables of
  // RB x 16 doubles each (ABR_BLOCK_RB=8 restores the smaller tables)
#ifndef ABR_BLOCK_RB
#define ABR_BLOCK_RB 16
#endif
  static constexpr int RB = NACC == 1 ? (1 << ROW_BITS) : ABR_BLOCK_RB;
};

// what the symmetric kernel's drain needs besides DrainCtx; kept in the warp's shared memory (a by-value
// argument of this size would travel through local memory at every call of the non-inlined drain)
struct SymCtx {
  double *ytmp;
  uint32_t *row_bits, *danger_count, *danger_list;
  uint32_t danger_capacity;
  uint32_t own_p0, own_p1; // rows this rank owns: [own_p0, own_p1)
  uint32_t pad_;
};

// per-warp shared memory (a warp works on one target bucket at a time and never
// synchronises with the other warps of its CTA)
template <int D, class F, bool STATS> struct WarpSmem {
  static constexpr int NACC = TiledCfg<D, F, STATS>::NACC;
  static constexpr int RB = TiledCfg<D, F, STATS>::RB;
  // rows of the batch, one array per dimension: a drain round reads rows[d][i] for 32
  // arbitrary i — 16 doubles span the 32 banks exactly once, so any pattern is conflict free
  double rows0[MAXD][RB];
  double rowsS[MAXD][RB];                 // rows shifted by a periodic image
  // rows relative to the bucket-stencil origin, fp32, for the pre-filter: row PAIRS packed
  // per dimension (x_2p, x_2p+1), (y..), (z..), pad — the operands of f32x2 instructions;
  // padded with far-away dummy rows
  unsigned long long rowsf2[RB / 2 + 1][4];
  unsigned long long part[NACC][RB][PCOL]; // partial sums [row][lane & 15]
  uint32_t lq[QCAP][32];                  // lane-private accepted-pair queues, slot major: (j << ROW_BITS) | row
  uint32_t wq[WQ];                        // the same pairs compacted for the drain, WQ at a time
  uint32_t run_pref[32];                  // candidate-run directory: inclusive prefix of run lengths
  uint32_t run_delta[32];                 //   j = k + run_delta[run]
  uint32_t danger;
  uint32_t danger_pre;                    // rows flagged before the candidates are seen (symmetric kernel: their partners are flagged too)
  uint32_t pad_[2];
  uint32_t rowid[RB];                     // index of the batch's rows in the caller's numbering (y, row columns)
  double rowsb[F::BC][RB];                // b of the rows (symmetric kernel only: y[j] += K(j,i) b[i])
  SymCtx sym;
};

#ifndef ABR_TILED_GRAB
#define ABR_TILED_GRAB 8
#endif
constexpr uint32_t TILED_GRAB = ABR_TILED_GRAB; // most consecutive buckets a warp claims per scheduler step (plan.grab)

// Stencil trimming.  A neighbour bucket at offset o (in buckets) from the target
// bucket is at least max(|o|-1, 0) * side away from every point of the target
// in that dimension.  Given the squared gap already spent in the slow dimensions,
// returns how many buckets the last dimension can still reach (-1: none), with a
// 1e-6 relative safety margin — far above rounding, far below anything that
// changes which buckets are needed.  (r/side = 1.09, w = 2: 81 of 125 buckets.)
__device__ __forceinline__ int reach_last_dim(double r2, double gap2, double side_last, int w_last) {
  const double rem2 = r2 * (1.0 + 1e-6) - gap2;
  if (rem2 < 0.0) return -1;
  const int m = (int)(sqrt(rem2) / side_last * (1.0 + 1e-6) + 1e-6) + 1;
  return m < w_last ? m : w_last;
}

// one 32-byte record (x, y, z, b) per lane: a single 256-bit load (LDG.E.256, sm_100)
__device__ __forceinline__ void ld_rec(const double *base, uint32_t j, double &x, double &y, double &z, double &w) {
  unsigned long long a, b, c, d, addr;
  asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(addr) : "r"(j), "l"(base));
  asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(addr));
  x = __longlong_as_double((long long)a);
  y = __longlong_as_double((long long)b);
  z = __longlong_as_double((long long)c);
  w = __longlong_as_double((long long)d);
}

// 32-bit shared-window addresses for the queue: a store and a bump per accepted pair
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

struct DrainCtx {
  const double *pos; // packed records (plan.posb)
  const double *b;
  double r2lo;
  double r2;
};
// symmetric kernel: bit 31 of the drain's image_id argument marks the primary image, where a pair
// (i, j > i) counts for both rows; periodic-image runs are one-sided (every pair, row i only)
constexpr uint32_t TWO_SIDED = 0x80000000u;
// symmetric kernel: row j's result must come from the exact walk (first flagging appends it to the list)
__device__ __forceinline__ void flag_row(const SymCtx &p, uint32_t j) {
  const uint32_t bit = 1u << (j & 31u);
  const uint32_t old = atomicOr(&p.row_bits[j >> 5], bit);
  if (!(old & bit)) {
    const uint32_t slot = atomicAdd(p.danger_count, 1u);
    if (slot < p.danger_capacity) p.danger_list[slot] = j;
  }
}

// Drain the lane-private queues.  The queued (j, row) pairs of all lanes are
// first compacted into one warp-wide list (an exclusive scan of the queue
// lengths gives every lane its offset), then dealt out 32 per round, so the
// expensive part of the product (sqrt, divide, the user's math) runs at full
// lane utilisation instead of on the ~15 % of lanes that pass the cut-off test.
// dx and |dx|^2 are recomputed here in fp64 with the reference's operations in
// the reference's order.
template <int D, class F, bool STATS, int SYM, bool GEN, class SM>
__device__ __noinline__ void drain_queues(SM &sm, const DrainCtx p, const F f, int lane, uint32_t cnt, uint32_t r0,
                                          const double (*rowp)[SM::RB], uint32_t image_id) {
  constexpr int BR = F::BR, BC = F::BC;
  uint32_t pin = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pin, o);
    if (lane >= o) pin += t;
  }
  const uint32_t total = __shfl_sync(0xFFFFFFFFu, pin, 31);
  const uint32_t excl = pin - cnt; // this lane's pairs are numbers excl .. pin-1 of the warp's list
  const uint32_t lq0 = (uint32_t)__cvta_generic_to_shared(&sm.lq[0][lane]);
  const uint32_t wq0 = (uint32_t)__cvta_generic_to_shared(&sm.wq[0]);
  const int col = lane & (PCOL - 1);
  // the compacted list is produced WQ pairs at a time (almost always a single pass)
  for (uint32_t wbase = 0; wbase < total; wbase += WQ) {
    {
      const uint32_t lo = max(excl, wbase), hi = min(pin, wbase + WQ);
      for (uint32_t k = lo; k < hi; ++k) sts32(wq0 + (k - wbase) * 4u, lds32(lq0 + (k - excl) * 128u));
    }
    __syncwarp();
    const uint32_t wtotal = min(total - wbase, (uint32_t)WQ);
    for (uint32_t base = 0; base < wtotal; base += 32) {
      const uint32_t k = base + lane;
      const bool live = k < wtotal;
      const uint32_t ent = sm.wq[live ? k : 0u];
      const uint32_t j = ent >> ROW_BITS;
      const uint32_t i = ent & ((1u << ROW_BITS) - 1u);
      // one 256-bit load brings the candidate's position and b
      double pj[D], bj[BC];
      {
        double rec[4];
        ld_rec(p.pos, j, rec[0], rec[1], rec[2], rec[3]);
  #pragma unroll
        for (int d = 0; d < D; ++d) pj[d] = rec[d];
        if (!STATS) {
          if (BC == 1) {
            bj[0] = rec[3];
          } else {
  #pragma unroll
            for (int c = 0; c < BC; ++c) bj[c] = p.b[(size_t)j * BC + c];
          }
        }
      }
      double dx[D];
      double d2 = 0;
  #pragma unroll
      for (int d = 0; d < D; ++d) {
        dx[d] = pj[d] - rowp[d][i];
        d2 = d2 + dx[d] * dx[d];
      }
      // the queue holds the survivors of the conservative fp32 pre-filter; this is
      // the reference's exact predicate (src/Search.h:438-446)
      bool ok = live && !(d2 > p.r2);
      bool both = false; // symmetric kernel: this pair also counts for row j
      if (SYM != 0) {
        // primary image: every unordered pair of the half stencil once (j > i; the self pair j == i one-sided);
        // j < i only occurs inside the target bucket itself and is the pair (j, i) seen from row j
        const uint32_t gi = r0 + i;
        if (image_id & TWO_SIDED) {
          const uint32_t o0 = sm.sym.own_p0, o1 = sm.sym.own_p1;
          ok = ok && j >= gi;
          both = ok && j > gi && j >= o0 && j < o1;
          ok = ok && gi >= o0 && gi < o1; // a ghost row (slab ranks) only serves its owned partners
        }
        if (both && (d2 > p.r2lo || ((sm.danger_pre >> i) & 1u))) flag_row(sm.sym, j);
      }
      if (ok && d2 > p.r2lo) atomicOr(&sm.danger, 1u << i);
      // lanes l and l+16 share a column of the partial-sum table: the two halves of
      // the warp update it one after the other
      if (STATS) {
        const unsigned long long hv = mix64((uint64_t)j * 81u + (uint64_t)image_id);
  #pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (ok && (lane >> 4) == half) {
            sm.part[0][i][col] += 1ull;
            sm.part[1][i][col] += hv;
          }
          __syncwarp();
        }
      } else {
        // F is evaluated unconditionally (it is pure; j, i are valid indices even for
        // a pair that fails the test); only the accumulation is predicated
        double blk[BR * BC];
        f(dx, d2, GEN ? sm.rowid[i] : r0 + i, j, blk);
        double s[BR];
  #pragma unroll
        for (int a2 = 0; a2 < BR; ++a2) {
          s[a2] = blk[a2 * BC] * bj[0];
  #pragma unroll
          for (int c = 1; c < BC; ++c) s[a2] += blk[a2 * BC + c] * bj[c];
        }
        if (SYM != 0) {
          // y[j] += K(j, i) b[i] with K(j, i) = SYM * K(i, j): one fp64 reduction per block row (RED.E.ADD.F64)
          if (both) {
  #pragma unroll
            for (int a2 = 0; a2 < BR; ++a2) {
              double t = blk[a2 * BC] * sm.rowsb[0][i];
  #pragma unroll
              for (int c = 1; c < BC; ++c) t += blk[a2 * BC + c] * sm.rowsb[c][i];
              atomicAdd(&sm.sym.ytmp[(size_t)j * BR + a2], SYM > 0 ? t : -t);
            }
          }
        }
        if (ABR_ACC_MODE == 2) {
          if (ok) {
  #pragma unroll
            for (int a2 = 0; a2 < BR; ++a2) {
              double *slot = reinterpret_cast<double *>(&sm.part[a2][i][col]);
              *slot += s[a2];
            }
          }
          __syncwarp();
        } else if (ABR_ACC_MODE == 1) {
          // the partner lane (l ^ 16) uses the same column: same row too -> the lower lane carries both terms
          const uint32_t mine = ok ? i : 0xFFFFu;
          const uint32_t theirs = __shfl_xor_sync(0xFFFFFFFFu, mine, 16);
          const bool merge = ok && mine == theirs;
  #pragma unroll
          for (int a2 = 0; a2 < BR; ++a2) {
            const double t = __shfl_xor_sync(0xFFFFFFFFu, s[a2], 16);
            if (merge) s[a2] += t;
          }
          if (ok && !(merge && lane >= 16)) {
  #pragma unroll
            for (int a2 = 0; a2 < BR; ++a2) {
              double *slot = reinterpret_cast<double *>(&sm.part[a2][i][col]);
              *slot += s[a2];
            }
          }
          __syncwarp();
        } else {
  #pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (ok && (lane >> 4) == half) {
  #pragma unroll
              for (int a2 = 0; a2 < BR; ++a2) {
                double *slot = reinterpret_cast<double *>(&sm.part[a2][i][col]);
                *slot += s[a2];
              }
            }
            __syncwarp();
          }
        }
      }
    }
    __syncwarp();
  }
}

// packed fp32 pairs (Blackwell FADD2 / FFMA2): one instruction, two rows
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// One step of the hot loop: this lane's TWO candidates (jA, jB) against the nr
// rows of the batch, two rows at a time.  This is a conservative fp32
// PRE-FILTER on coordinates relative to the stencil origin: a pair survives if
// |dx|^2 <= r^2 (1 + tol) in fp32, tol chosen on the host so that no pair the
// exact fp64 test accepts is ever dropped.  The two rows of a step are the two
// halves of packed f32x2 operands (FADD2 / FFMA2): 4 tests cost 4 D packed
// instructions.  Survivors cost one predicated 4-byte store into the lane's own
// queue (slot-major layout: the bank is the lane, never a conflict) and a
// pointer bump; the exact un-fused fp64 predicate is applied to them in
// drain_queues.  (6.45 candidates are tested per accepted pair, so this loop is
// kept off the fp64 pipe altogether.)
template <int D, class F, bool STATS, int SYM, bool GEN, class SM>
__device__ __forceinline__ void test_rows(SM &sm, const DrainCtx &dc, float pre_r2, const F &f, int lane,
                                          const float *pA, const float *pB, uint32_t jA, uint32_t jB, bool vA, bool vB,
                                          int nr, const double (*rowp)[SM::RB], uint32_t image_id, uint32_t r0,
                                          uint32_t &qa) {
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sm.lq[0][lane]);
  // invalid candidates are parked far away instead of being predicated out
  unsigned long long a[D], b[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float av = vA ? pA[d] : -3.0e18f, bv = vB ? pB[d] : -3.0e18f;
    a[d] = pack2(av, av);
    b[d] = pack2(bv, bv);
  }
  uint32_t eA = jA << ROW_BITS, eB = jB << ROW_BITS;
  const int npairs = (nr + 1) >> 1; // an odd tail pairs with a dummy row
  for (int pr = 0; pr < npairs; ++pr) {
    unsigned long long accA, accB;
    {
      const unsigned long long r = sm.rowsf2[pr][0];
      const unsigned long long ta = sub2(a[0], r), tb = sub2(b[0], r);
      accA = mul2(ta, ta);
      accB = mul2(tb, tb);
    }
#pragma unroll
    for (int d = 1; d < D; ++d) {
      const unsigned long long r = sm.rowsf2[pr][d];
      const unsigned long long ta = sub2(a[d], r), tb = sub2(b[d], r);
      accA = fma2(ta, ta, accA);
      accB = fma2(tb, tb, accB);
    }
    float a0, a1, b0, b1;
    unpack2(accA, a0, a1);
    unpack2(accB, b0, b1);
    if (a0 <= pre_r2) { sts32(qa, eA); qa += 128u; }
    if (a1 <= pre_r2) { sts32(qa, eA + 1u); qa += 128u; }
    if (b0 <= pre_r2) { sts32(qa, eB); qa += 128u; }
    if (b1 <= pre_r2) { sts32(qa, eB + 1u); qa += 128u; }
    eA += 2u;
    eB += 2u;
    if (__any_sync(0xFFFFFFFFu, qa >= q0 + QDRAIN * 128u)) {
      drain_queues<D, F, STATS, SYM, GEN>(sm, dc, f, lane, (qa - q0) >> 7, r0, rowp, image_id);
      qa = q0;
    }
  }
}

// occupancy target: TiledCfg::CTAS CTAs/SM (6.4 KB of shared memory per warp): 7 or 8 for
// scalar kernels, 5 for D x 1 block kernels
//
// SYM != 0 (functor-declared symmetry, rows == columns): the symmetric form.  A target bucket is tested
// against the FORWARD half of its stencil only (itself and the neighbours after it in bucket order); an
// accepted pair (i, j > i) is evaluated once and added to both rows — row i through the warp's partial-sum
// table, row j with one fp64 reduction (RED.E.ADD.F64) per block row.  All sums go to the zeroed scratch
// `ytmp` (k_sym_combine adds it to y afterwards) because a row flagged for the exact walk — by its own
// bucket or by a partner — must not keep partial sums.  Periodic-image runs stay one-sided (cur = r +
// image * L is not antisymmetric under exchange in floating point).  On slab ranks the lower ghost layers
// are targets too (their forward neighbours are owned rows); only owned rows receive sums.
// Results differ from the ordered form by summation order only (fp64 atomics are not ordered).
#ifndef ABR_SYM_CTAS_LESS
#define ABR_SYM_CTAS_LESS 1 // the symmetric form keeps a little more state: one resident CTA less per SM
#endif
// GEN: the general form — rows from another particle set bucketed into this grid (plan.xrow_*) and heavy
// buckets split into row-batch work items (plan.heavy_*).  GEN = false compiles both away: the common
// rows == columns product on a cloud without heavy buckets pays nothing for them (measured: 1.8 %).
template <int D, class F, bool STATS, int SYM = 0, bool GEN = false>
__global__ void __launch_bounds__(TILED_THREADS, (TiledCfg<D, F, STATS>::CTAS - (SYM != 0 && TiledCfg<D, F, STATS>::CTAS > 5 ? ABR_SYM_CTAS_LESS : 0)))
tiled_kernel(const abr_matvec_plan p, const F f) {
  constexpr int BR = F::BR;
  constexpr int NACC = TiledCfg<D, F, STATS>::NACC;
  constexpr int RB = TiledCfg<D, F, STATS>::RB;
  using WS = WarpSmem<D, F, STATS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WS &sm = reinterpret_cast<WS *>(smem_raw)[warp];
  const Grid &g = p.q.g;
  const double *__restrict__ pos = p.q.pos;
  const uint32_t *__restrict__ bbeg = p.q.bucket_begin;
  const uint32_t *__restrict__ bend = p.q.bucket_end;
  constexpr int L = D - 1; // last (fastest, memory-contiguous) dimension
  constexpr int DS = D > 1 ? D - 1 : 1;
  // number of offset tuples in the D-1 slow dimensions
  int nslow = 1;
#pragma unroll
  for (int d = 0; d < D - 1; ++d) nslow *= 2 * p.w[d] + 1;
  int img0[D];
#pragma unroll
  for (int d = 0; d < D; ++d) img0[d] = 0;
  const uint32_t image_id0 = STATS ? (uint32_t)image_linear_index<D>(g, img0) : (SYM != 0 ? TWO_SIDED : 0u);
  // buckets of the bucket layers this rank owns (all of them on a single GPU)
  uint32_t per_layer = 1;
#pragma unroll
  for (int d = 1; d < D; ++d) per_layer *= (uint32_t)g.size[d];
  const uint32_t first_cell = (D > 1 ? (uint32_t)g.own_lo * per_layer : 0u);
  const uint32_t own_cells = (D > 1 ? (uint32_t)g.own_n * per_layer : g.ncells);
  // target buckets: the owned ones; the symmetric kernel also visits the lower ghost layers of a slab rank
  const uint32_t tfirst = SYM != 0 ? 0u : first_cell;
  const uint32_t tcount = first_cell + own_cells - tfirst;
  uint32_t own_p0 = 0, own_p1 = 0xFFFFFFFFu;
  if (SYM != 0 && own_cells > 0) {
    own_p0 = bbeg[first_cell];
    own_p1 = bend[first_cell + own_cells - 1];
  }
  const DrainCtx dc{p.posb, p.b, p.r2lo, p.r2};
  if (SYM != 0) {
    if (lane == 0) sm.sym = SymCtx{p.ytmp, p.row_bits, p.danger_count, p.danger_list, p.danger_capacity, own_p0, own_p1, 0u};
    __syncwarp();
  }
  const double *__restrict__ posb = p.posb;
  const float pre_r2 = p.pre_r2;
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sm.lq[0][lane]);
  const int S = g.size[L];

  // offsets of run `rid` of the stencil in the slow dimensions, the squared gap they
  // already spend, and how far the last dimension still reaches (trimmed stencil)
  auto decode_run = [&](int rid, int *od, int &wz) {
    int rem = rid;
    double gap2 = 0.0;
#pragma unroll
    for (int d = D - 2; d >= 0; --d) {
      const int span = 2 * p.w[d] + 1;
      od[d] = (rem % span) - p.w[d];
      rem /= span;
      const double gap = (double)max(abs(od[d]) - 1, 0) * g.side[d];
      gap2 += gap * gap;
    }
    wz = p.trim ? reach_last_dim(p.r2, gap2, g.side[L], p.w[L]) : p.w[L];
  };
  // the first 32 runs are the same for every bucket: decode them once
  int od_first[DS], wz_first;
  od_first[0] = 0;
  decode_run(lane, od_first, wz_first);

  // everything a target bucket needs; `batch` (heavy launch only): which row batch of the bucket
  int tc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) tc[d] = 0;
  const bool HEAVY = GEN && p.heavy_phase != 0;
  auto do_cell = [&](const uint32_t cell, const uint32_t batch) {
      const bool xrows = GEN && SYM == 0 && p.xrow_pos != nullptr; // rows from another particle set, bucketed into this grid
      const uint32_t rb = xrows ? p.xrow_bb[cell] : bbeg[cell], re = xrows ? p.xrow_be[cell] : bend[cell];
      if (rb == re) return;
      if (GEN && !HEAVY && p.heavy_list && re - rb > HEAVY_ROWS) {
        // a heavy bucket (clustered clouds): its row batches become independent work items of the second launch
        if (lane == 0) {
          const uint32_t nbatch = (re - rb + RB - 1) / RB;
          const unsigned long long old = atomicAdd(p.heavy_state, (1ull << 40) + nbatch);
          const uint32_t slot = (uint32_t)(old >> 40);
          if (slot < p.heavy_capacity) p.heavy_list[slot] = make_uint2(cell, (uint32_t)(old & ((1ull << 40) - 1ull)));
        }
        return;
      }
      const uint32_t r_lo = HEAVY ? rb + batch * RB : rb, r_hi = HEAVY ? min(re, r_lo + RB) : re;
      const double *__restrict__ rpos = xrows ? p.xrow_pos : pos;
      const bool owned_target = SYM == 0 || first_cell == 0u || cell >= first_cell; // else: a lower ghost layer (symmetric kernel on a slab rank)
      const int zlo = tc[L] - p.w[L], zhi = tc[L] + p.w[L];
      // does any neighbour of this bucket lie across a periodic boundary / outside?
      bool boundary = (zlo < 0) | (zhi >= S);
#pragma unroll
      for (int d = 0; d < D - 1; ++d) boundary |= (tc[d] - p.w[d] < 0) | (tc[d] + p.w[d] >= g.size[d]);

      // origin of the fp32 pre-filter coordinates: lower corner of the stencil
      double origin[D];
#pragma unroll
      for (int d = 0; d < D; ++d) origin[d] = g.bmin[d] + (double)(tc[d] - p.w[d]) * g.side[d];

      for (uint32_t r0 = r_lo; r0 < r_hi; r0 += RB) {
        const int nr = (int)min((uint32_t)RB, re - r0);
        // ---- load the rows of this batch, flag rounding-sensitive ones ----
        bool my_danger = false;
        __syncwarp();
        if (lane < nr) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const double r = rpos[(size_t)(r0 + lane) * D + d];
            sm.rows0[d][lane] = r;
            const double fl = (r - g.bmin[d]) * g.inv_side[d];
            const double fr = fl - floor(fl);
            my_danger |= ((int)floor(fl) != tc[d]) | (fr < p.tolf[d]) | (fr > 1.0 - p.tolf[d]) |
                         (fabs(fr - 0.5) < p.tolf[d]);
          }
        }
        if (lane == 0) sm.danger = 0;
        if (GEN && lane < nr) sm.rowid[lane] = xrows ? (uint32_t)p.xrow_index[r0 + lane] : r0 + lane;
        if (SYM != 0) {
          if (lane < nr) {
#pragma unroll
            for (int c = 0; c < F::BC; ++c) sm.rowsb[c][lane] = p.b[(size_t)(r0 + lane) * F::BC + c];
          }
          const uint32_t pre = __ballot_sync(0xFFFFFFFFu, my_danger);
          if (lane == 0) sm.danger_pre = pre;
        }
        __syncwarp();
        if (lane < RB + 2) {
          // fp32 copy relative to the stencil origin (lower corner of the first neighbour
          // bucket), packed in row pairs; rows >= nr are dummies far away from everything
          float *rf = reinterpret_cast<float *>(&sm.rowsf2[0][0]);
#pragma unroll
          for (int d = 0; d < D; ++d)
            rf[(((lane >> 1) * 4) + d) * 2 + (lane & 1)] = lane < nr ? (float)(sm.rows0[d][lane] - origin[d]) : 3.0e18f;
        }
#pragma unroll
        for (int a = 0; a < NACC; ++a)
          for (int e = lane; e < nr * PCOL; e += 32) (&sm.part[a][0][0])[e] = 0ull;
        __syncwarp();
        uint32_t qa = q0; // next free slot of this lane's queue (shared-window address)

        // ---- phase 1: neighbour runs in the primary image, concatenated so that
        //      every step tests 64 candidates (one run = the 2w+1 buckets along the
        //      last dimension, contiguous in the sorted arrays) ----
        for (int rbase = 0; rbase < nslow; rbase += 32) {
          uint32_t len = 0, jb = 0;
          const int rid = rbase + lane;
          if (rid < nslow) {
            int od[DS], wz;
            if (rbase == 0) {
#pragma unroll
              for (int d = 0; d < DS; ++d) od[d] = od_first[d];
              wz = wz_first;
            } else {
              decode_run(rid, od, wz);
            }
            int nc[D];
            bool ok = true;
#pragma unroll
            for (int d = 0; d < D - 1; ++d) {
              const int u = tc[d] + od[d];
              ok &= (u >= 0) & (u < g.size[d]);
              nc[d] = u;
            }
            int zfirst = tc[L] - wz;
            if (SYM != 0) {
              // forward half of the stencil: runs after the target's own run in bucket order, and of the
              // own run the part from the target bucket on
              int sgn = 0;
#pragma unroll
              for (int d = 0; d < D - 1; ++d)
                if (sgn == 0 && od[d] != 0) sgn = od[d] > 0 ? 1 : -1;
              ok &= sgn >= 0;
              if (sgn == 0) zfirst = tc[L];
            }
            const int a = max(zfirst, 0), bnd = min(tc[L] + wz, S - 1);
            if (ok && wz >= 0 && a <= bnd) {
              nc[L] = a;
              const int c_lo = local_collapse<D>(g, nc);
              if (c_lo >= 0) {
                jb = bbeg[c_lo];
                len = bend[c_lo + (bnd - a)] - jb;
              }
            }
          }
          uint32_t pin = len;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pin, o);
            if (lane >= o) pin += t;
          }
          const uint32_t total = __shfl_sync(0xFFFFFFFFu, pin, 31);
          __syncwarp();
          sm.run_pref[lane] = pin;
          sm.run_delta[lane] = jb - (pin - len);
          __syncwarp();
          for (uint32_t kb = 0; kb < total; kb += 64) {
            // two candidates per lane: k and k + 32
            uint32_t jj[2];
            bool vv[2];
            float pj[2][D];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#if ABR_SKEW_LANES
              // the second candidate of a lane sits half a run away from the first one: a lane whose
              // first candidate is in the middle of a run (accepted by most rows) gets a peripheral
              // second one, which evens out the queue lengths (fewer, fuller drains)
              const uint32_t k = kb + 32 * h + (h ? ((lane + 16) & 31) : lane);
#else
              const uint32_t k = kb + 32 * h + lane;
#endif
              vv[h] = k < total;
              const uint32_t ks = vv[h] ? k : total - 1;
              uint32_t rho = 0;
#pragma unroll
              for (int step = 16; step > 0; step >>= 1)
                if (sm.run_pref[rho + step - 1] <= ks) rho += step;
              jj[h] = ks + sm.run_delta[rho];
              double rec[4];
              ld_rec(posb, jj[h], rec[0], rec[1], rec[2], rec[3]);
#pragma unroll
              for (int d = 0; d < D; ++d) pj[h][d] = (float)(rec[d] - origin[d]);
            }
            test_rows<D, F, STATS, SYM, GEN>(sm, dc, pre_r2, f, lane, pj[0], pj[1], jj[0], jj[1], vv[0], vv[1], nr, sm.rows0,
                                        image_id0, r0, qa);
          }
        }
        // pairs queued so far belong to the primary image
        if (__any_sync(0xFFFFFFFFu, qa != q0)) {
          drain_queues<D, F, STATS, SYM, GEN>(sm, dc, f, lane, (qa - q0) >> 7, r0, sm.rows0, image_id0);
          qa = q0;
        }
        if (boundary && owned_target) {
          // ---- phase 2 (buckets at a periodic boundary only): runs reached through
          //      a periodic image; cur = r + image * L exactly as src/Search.h:188-190 ----
          int o[DS];
#pragma unroll
          for (int d = 0; d < D - 1; ++d) o[d] = -p.w[d];
          bool more = true;
          while (more) {
            int nc[D], img[D];
            bool ok_slow = true, slow_shifted = false;
            double gap2 = 0.0;
#pragma unroll
            for (int d = 0; d < D - 1; ++d) {
              const double gap = (double)max(abs(o[d]) - 1, 0) * g.side[d];
              gap2 += gap * gap;
              int u = tc[d] + o[d];
              img[d] = 0;
              if (u < 0) {
                u += g.size[d];
                img[d] = 1;
              } else if (u >= g.size[d]) {
                u -= g.size[d];
                img[d] = -1;
              }
              ok_slow &= (u >= 0) & (u < g.size[d]) & (img[d] == 0 || g.periodic[d]);
              slow_shifted |= (img[d] != 0);
              nc[d] = u;
            }
            const int wz = p.trim ? reach_last_dim(p.r2, gap2, g.side[L], p.w[L]) : p.w[L]; // trimmed stencil
            if (ok_slow && wz >= 0) {
              for (int m = (g.periodic[L] ? -1 : 0); m <= (g.periodic[L] ? 1 : 0); ++m) {
                if (m == 0 && !slow_shifted) continue; // primary image: done in phase 1
                // unwrapped indices [m*S, m*S + S - 1] map to buckets [0,S-1] with image -m
                const int a = max(tc[L] - wz, m * S), bnd = min(tc[L] + wz, m * S + S - 1);
                if (a > bnd) continue;
                img[L] = -m;
                nc[L] = a - m * S;
                const int c_lo = local_collapse<D>(g, nc);
                if (c_lo < 0) continue;
                const uint32_t jb = bbeg[c_lo], je = bend[c_lo + (bnd - a)];
                if (jb >= je) continue;
                __syncwarp();
                if (lane < nr) {
#pragma unroll
                  for (int d = 0; d < D; ++d) sm.rowsS[d][lane] = sm.rows0[d][lane] + (double)img[d] * g.L[d];
                }
                __syncwarp();
                const uint32_t image_id = STATS ? (uint32_t)image_linear_index<D>(g, img) : 0u;
                for (uint32_t cb = jb; cb < je; cb += 64) {
                  uint32_t jj[2];
                  bool vv[2];
                  float pj[2][D];
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    jj[h] = min(cb + 32 * h + lane, je - 1);
                    vv[h] = cb + 32 * h + lane < je;
                    // pre-filter only: move the candidate by -image*L instead of the row by +image*L
                    double rec[4];
                    ld_rec(posb, jj[h], rec[0], rec[1], rec[2], rec[3]);
#pragma unroll
                    for (int d = 0; d < D; ++d) pj[h][d] = (float)((rec[d] - (double)img[d] * g.L[d]) - origin[d]);
                  }
                  test_rows<D, F, STATS, SYM, GEN>(sm, dc, pre_r2, f, lane, pj[0], pj[1], jj[0], jj[1], vv[0], vv[1], nr, sm.rowsS,
                                              image_id, r0, qa);
                }
                // leave no pair of this image in the queues (rowsS is reused)
                if (__any_sync(0xFFFFFFFFu, qa != q0)) {
                  drain_queues<D, F, STATS, SYM, GEN>(sm, dc, f, lane, (qa - q0) >> 7, r0, sm.rowsS, image_id);
                  qa = q0;
                }
              }
            }
            // next offset tuple in the slow dimensions (odometer)
            more = false;
#pragma unroll
            for (int d = D - 2; d >= 0; --d) {
              if (!more) {
                if (++o[d] <= p.w[d]) {
                  more = true;
                } else {
                  o[d] = -p.w[d];
                }
              }
            }
          }
        }
        __syncwarp();

        // ---- reduce part[row][*] in a fixed (skewed, conflict-free) order ----
        const uint32_t dmask = sm.danger | __ballot_sync(0xFFFFFFFFu, my_danger);
        if (lane < nr && owned_target) {
          const bool dangerous = (dmask >> lane) & 1u;
          if (dangerous) {
            if (SYM != 0) {
              flag_row(sm.sym, r0 + lane); // a partner's drain may have flagged it already
            } else {
              const uint32_t slot = atomicAdd(p.danger_count, 1u);
              if (slot < p.danger_capacity) p.danger_list[slot] = GEN ? sm.rowid[lane] : r0 + lane;
            }
          } else if (STATS) {
            unsigned long long c = 0, hsum = 0;
#pragma unroll
            for (int k = 0; k < PCOL; ++k) {
              c += sm.part[0][lane][(k + lane) & (PCOL - 1)];
              hsum += sm.part[1][lane][(k + lane) & (PCOL - 1)];
            }
            const uint32_t orow = GEN ? sm.rowid[lane] : r0 + lane;
            if (p.stat_count) p.stat_count[orow] = (uint32_t)c;
            if (p.stat_hash) p.stat_hash[orow] = hsum;
          } else {
#pragma unroll
            for (int a2 = 0; a2 < NACC; ++a2) {
              double s = 0;
#pragma unroll
              for (int k = 0; k < PCOL; ++k) s += *reinterpret_cast<double *>(&sm.part[a2][lane][(k + lane) & (PCOL - 1)]);
              if (SYM != 0)
                atomicAdd(&p.ytmp[(size_t)(r0 + lane) * BR + a2], s); // partners add to the same entry concurrently
              else
                p.y[(size_t)(GEN ? sm.rowid[lane] : r0 + lane) * BR + a2] += s;
            }
          }
        }
        __syncwarp();
      }
  };

  if (HEAVY) {
    // second launch: (heavy bucket, row batch) items handed out one at a time
    const unsigned long long st = *p.heavy_state;
    const uint32_t n_items = min((uint32_t)(st >> 40), p.heavy_capacity);
    const uint32_t n_batches = (uint32_t)(st & ((1ull << 40) - 1ull));
    while (true) {
      uint32_t t = 0;
      if (lane == 0) t = atomicAdd(p.heavy_work, 1u);
      t = __shfl_sync(0xFFFFFFFFu, t, 0);
      if (t >= n_batches) break;
      // the item whose batches contain t (the list is ordered by its base)
      uint32_t lo = 0, hi = n_items;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p.heavy_list[mid].y <= t) lo = mid; else hi = mid;
      }
      const uint2 item = p.heavy_list[lo];
      const uint32_t cell = item.x;
      uint32_t rem = cell;
#pragma unroll
      for (int d = D - 1; d >= 0; --d) {
        tc[d] = (int)(rem % (uint32_t)g.size[d]);
        rem /= (uint32_t)g.size[d];
      }
      if (D > 1) {
        int gl = g.win_lo + (int)(cell / per_layer);
        if (gl < 0) gl += g.size[0];
        if (gl >= g.size[0]) gl -= g.size[0];
        tc[0] = gl;
      }
      do_cell(cell, t - item.y);
    }
    return;
  }

  while (true) {
    // warp-level dynamic scheduler: no block barrier anywhere in this kernel
    uint32_t grab = 0;
    if (lane == 0) grab = atomicAdd(p.work_counter, p.grab);
    grab = __shfl_sync(0xFFFFFFFFu, grab, 0);
    if (grab >= tcount) break;
    const uint32_t grab_end = min(grab + p.grab, tcount);

    // bucket coordinates of the first bucket of the grab (inverse of collapse_index);
    // the following ones are reached by counting up, last dimension fastest
    {
      const uint32_t cell = tfirst + grab;
      uint32_t rem = cell;
#pragma unroll
      for (int d = D - 1; d >= 0; --d) {
        tc[d] = (int)(rem % (uint32_t)g.size[d]);
        rem /= (uint32_t)g.size[d];
      }
      if (D > 1) { // local layer -> global layer (slab window; identity on a single GPU)
        int gl = g.win_lo + (int)(cell / per_layer);
        if (gl < 0) gl += g.size[0];
        if (gl >= g.size[0]) gl -= g.size[0];
        tc[0] = gl;
      }
    }
    --tc[L]; // the loop advances before it works

    for (uint32_t cell = tfirst + grab; cell < tfirst + grab_end; ++cell) {
      // advance the bucket coordinates to `cell`
      if (++tc[L] == S) {
        if (D > 1) {
          tc[L] = 0;
          if (D > 2) {
            if (++tc[D > 2 ? 1 : 0] == g.size[D > 2 ? 1 : 0]) {
              tc[D > 2 ? 1 : 0] = 0;
              if (++tc[0] >= g.size[0]) tc[0] -= g.size[0]; // next (global) layer of the window
            }
          } else {
            if (++tc[0] >= g.size[0]) tc[0] -= g.size[0];
          }
        }
      }
      do_cell(cell, 0u);
    }
  }
}
// ---------------------------------------------------------------------------
// staged_kernel — the cell-tiled product with the candidate records staged in
// shared memory by the bulk-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP /
// SYNCS): "stages the neighbouring cells' sorted positions and b values into shared
// memory with TMA" (north_star).
//
// Same decomposition as tiled_kernel (one warp per target bucket, candidates = the
// contiguous particle runs of the neighbouring buckets, conservative fp32 pre-filter,
// exact fp64 predicate + F in a compacted drain), but the unit of work is a CHUNK of 64
// consecutive candidates of the concatenated runs:
//   * the lanes that own a run (lane = run) each issue ONE bulk copy for the piece of
//     their run that falls into the chunk — 32-byte (x, y, z, b) records, contiguous in
//     the sorted array — into a per-warp double buffer; lane 0 arms the buffer's mbarrier
//     with the byte count.  Chunk c+1 is in flight while chunk c is tested and drained;
//   * a lane's two candidates are slots `lane` and `32 + (lane+16)%32` of the buffer: no
//     run lookup (the binary search of tiled_kernel is gone), no global gather;
//   * queue entries are (slot << 4 | row): the drain runs at the end of every chunk (and
//     whenever a lane's queue fills) and reads the candidate records from shared memory
//     instead of re-gathering them from L2;
//   * the column index j is reconstructed (a few shuffles per chunk) only for functors
//     that read it (uses_j), for block columns (BC > 1) and for the pair-set hashes.
// Periodic-image runs (boundary buckets only) are staged by plain loads into the same
// buffers, so test and drain code is shared.
// ---------------------------------------------------------------------------
template <class F, class = void> struct uses_j { static constexpr bool value = true; };
template <class F> struct uses_j<F, decltype((void)F::USES_J)> { static constexpr bool value = F::USES_J; };

constexpr int CHUNK = 64;
#ifndef ABR_STAGED_PCOL
#define ABR_STAGED_PCOL 16
#endif
#ifndef ABR_STAGED_CTAS
#define ABR_STAGED_CTAS 6
#endif
constexpr int SPCOL = ABR_STAGED_PCOL; // columns of the partial-sum table: 16 (two half-warp passes) or 32 (one pass)

template <int D, class F, bool STATS> struct StagedCfg {
  static constexpr int NACC = STATS ? 2 : F::BR;
  static constexpr int RB = TiledCfg<D, F, STATS>::RB;
  static constexpr bool NEEDJ = STATS || uses_j<F>::value || F::BC != 1;
  static constexpr int CTAS = NACC == 1 ? ABR_STAGED_CTAS : (NACC == 2 ? 5 : 4);
};

template <int D, class F, bool STATS> struct alignas(128) StagedSmem {
  static constexpr int NACC = StagedCfg<D, F, STATS>::NACC;
  static constexpr int RB = StagedCfg<D, F, STATS>::RB;
  double stage[2][CHUNK][4];               // candidate records (x, y, z, b) of two chunks
  unsigned long long mbar[2];              // one mbarrier per buffer
  double rows0[MAXD][RB];
  double rowsS[MAXD][RB];
  unsigned long long rowsf2[RB / 2 + 1][4];
  unsigned long long part[NACC][RB][SPCOL];
  uint16_t lq[QCAP][32];                   // lane-private queues, slot major: (slot << ROW_BITS) | row
  uint16_t wq[WQ];                         // compacted for the drain
  uint32_t sj[CHUNK];                      // column index of each slot (NEEDJ only)
  uint32_t danger;
  uint32_t pad_[3];
};

__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!done && ++spins > (1 << 22)) __trap(); // a lost copy must fail loudly, never hang the GPU
  }
}
// global -> shared bulk copy, completion reported to an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

struct SDrainCtx {
  const double *b;   // b column (BC > 1 only)
  double r2lo;
  double r2;
  uint32_t stage;    // shared-window address of the chunk's records
};

template <int D, class F, bool STATS, class SM>
__device__ __noinline__ void sdrain(SM &sm, const SDrainCtx p, const F f, int lane, uint32_t cnt, uint32_t r0, const double (*rowp)[SM::RB],
                                    uint32_t image_id) {
  constexpr int BR = F::BR, BC = F::BC;
  constexpr bool NEEDJ = StagedCfg<D, F, STATS>::NEEDJ;
  uint32_t pin = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pin, o);
    if (lane >= o) pin += t;
  }
  const uint32_t total = __shfl_sync(0xFFFFFFFFu, pin, 31);
  const uint32_t excl = pin - cnt;
  const uint32_t lq0 = (uint32_t)__cvta_generic_to_shared(&sm.lq[0][lane]);
  const uint32_t wq0 = (uint32_t)__cvta_generic_to_shared(&sm.wq[0]);
  for (uint32_t wbase = 0; wbase < total; wbase += WQ) {
    {
      const uint32_t lo = max(excl, wbase), hi = min(pin, wbase + WQ);
      for (uint32_t k = lo; k < hi; ++k) sts16(wq0 + (k - wbase) * 2u, lds16(lq0 + (k - excl) * 64u));
    }
    __syncwarp();
    const uint32_t wtotal = min(total - wbase, (uint32_t)WQ);
    for (uint32_t base = 0; base < wtotal; base += 32) {
      const uint32_t k = base + lane;
      const bool live = k < wtotal;
      const uint32_t ent = sm.wq[live ? k : 0u];
      const uint32_t slot = ent >> ROW_BITS;
      const uint32_t i = ent & ((1u << ROW_BITS) - 1u);
      double pj[3], bj[BC];
      uint32_t j = 0;
      {
        double rx, ry, rz, rb;
        const uint32_t a = p.stage + slot * 32u;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rx), "=d"(ry) : "r"(a));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rz), "=d"(rb) : "r"(a + 16u));
        pj[0] = rx;
        pj[1] = ry;
        pj[2] = rz;
        if (NEEDJ) j = sm.sj[slot];
        if (!STATS) {
          if (BC == 1) {
            bj[0] = rb;
          } else {
#pragma unroll
            for (int c = 0; c < BC; ++c) bj[c] = p.b[(size_t)j * BC + c];
          }
        }
      }
      double dx[D];
      double d2 = 0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        dx[d] = pj[d] - rowp[d][i];
        d2 = d2 + dx[d] * dx[d];
      }
      // the reference's exact predicate (src/Search.h:438-446) on the survivors of the pre-filter
      const bool ok = live && !(d2 > p.r2);
      if (ok && d2 > p.r2lo) atomicOr(&sm.danger, 1u << i);
      if (STATS) {
        const unsigned long long hv = mix64((uint64_t)j * 81u + (uint64_t)image_id);
        if (SPCOL == 32) {
          if (ok) {
            sm.part[0][i][lane & (SPCOL - 1)] += 1ull;
            sm.part[1][i][lane & (SPCOL - 1)] += hv;
          }
        } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (ok && (lane >> 4) == half) {
              sm.part[0][i][lane & (SPCOL - 1)] += 1ull;
              sm.part[1][i][lane & (SPCOL - 1)] += hv;
            }
            __syncwarp();
          }
        }
      } else {
        double blk[BR * BC];
        f(dx, d2, r0 + i, j, blk);
        double s[BR];
#pragma unroll
        for (int a2 = 0; a2 < BR; ++a2) {
          s[a2] = blk[a2 * BC] * bj[0];
#pragma unroll
          for (int c = 1; c < BC; ++c) s[a2] += blk[a2 * BC + c] * bj[c];
        }
        if (SPCOL == 32) {
          if (ok) {
#pragma unroll
            for (int a2 = 0; a2 < BR; ++a2) {
              double *cell = reinterpret_cast<double *>(&sm.part[a2][i][lane & (SPCOL - 1)]);
              *cell += s[a2];
            }
          }
        } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (ok && (lane >> 4) == half) {
#pragma unroll
              for (int a2 = 0; a2 < BR; ++a2) {
                double *cell = reinterpret_cast<double *>(&sm.part[a2][i][lane & (SPCOL - 1)]);
                *cell += s[a2];
              }
            }
            __syncwarp();
          }
        }
      }
    }
    __syncwarp();
  }
}

// pre-filter of one chunk: this lane's two candidates (slots sA, sB of the staged buffer)
// against the nr rows of the batch, two rows per packed instruction — see test_rows
template <int D, class F, bool STATS, class SM>
__device__ __forceinline__ void stest_rows(SM &sm, const SDrainCtx &dc, float pre_r2, const F &f, int lane, const float *pA, const float *pB,
                                           uint32_t sA, uint32_t sB, bool vA, bool vB, int nr, const double (*rowp)[SM::RB], uint32_t image_id,
                                           uint32_t r0, uint32_t &qa) {
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sm.lq[0][lane]);
  unsigned long long a[D], b[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float av = vA ? pA[d] : -3.0e18f, bv = vB ? pB[d] : -3.0e18f;
    a[d] = pack2(av, av);
    b[d] = pack2(bv, bv);
  }
  uint32_t eA = sA << ROW_BITS, eB = sB << ROW_BITS;
  const int npairs = (nr + 1) >> 1;
  for (int pr = 0; pr < npairs; ++pr) {
    unsigned long long accA, accB;
    {
      const unsigned long long r = sm.rowsf2[pr][0];
      const unsigned long long ta = sub2(a[0], r), tb = sub2(b[0], r);
      accA = mul2(ta, ta);
      accB = mul2(tb, tb);
    }
#pragma unroll
    for (int d = 1; d < D; ++d) {
      const unsigned long long r = sm.rowsf2[pr][d];
      const unsigned long long ta = sub2(a[d], r), tb = sub2(b[d], r);
      accA = fma2(ta, ta, accA);
      accB = fma2(tb, tb, accB);
    }
    float a0, a1, b0, b1;
    unpack2(accA, a0, a1);
    unpack2(accB, b0, b1);
    if (a0 <= pre_r2) { sts16(qa, eA); qa += 64u; }
    if (a1 <= pre_r2) { sts16(qa, eA + 1u); qa += 64u; }
    if (b0 <= pre_r2) { sts16(qa, eB); qa += 64u; }
    if (b1 <= pre_r2) { sts16(qa, eB + 1u); qa += 64u; }
    eA += 2u;
    eB += 2u;
    if (__any_sync(0xFFFFFFFFu, qa >= q0 + QDRAIN * 64u)) {
      sdrain<D, F, STATS>(sm, dc, f, lane, (qa - q0) >> 6, r0, rowp, image_id);
      qa = q0;
    }
  }
}

template <int D, class F, bool STATS>
__global__ void __launch_bounds__(TILED_THREADS, (StagedCfg<D, F, STATS>::CTAS))
staged_kernel(const abr_matvec_plan p, const F f) {
  constexpr int BR = F::BR;
  constexpr int NACC = StagedCfg<D, F, STATS>::NACC;
  constexpr int RB = StagedCfg<D, F, STATS>::RB;
  constexpr bool NEEDJ = StagedCfg<D, F, STATS>::NEEDJ;
  using WS = StagedSmem<D, F, STATS>;
  extern __shared__ __align__(16) unsigned char smem_raw[]; // no static shared memory in this kernel: the window starts 1024-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WS &sm = reinterpret_cast<WS *>(smem_raw)[warp];
  const Grid &g = p.q.g;
  const double *__restrict__ pos = p.q.pos;
  const uint32_t *__restrict__ bbeg = p.q.bucket_begin;
  const uint32_t *__restrict__ bend = p.q.bucket_end;
  constexpr int L = D - 1;
  constexpr int DS = D > 1 ? D - 1 : 1;
  int nslow = 1;
#pragma unroll
  for (int d = 0; d < D - 1; ++d) nslow *= 2 * p.w[d] + 1;
  int img0[D];
#pragma unroll
  for (int d = 0; d < D; ++d) img0[d] = 0;
  const uint32_t image_id0 = STATS ? (uint32_t)image_linear_index<D>(g, img0) : 0u;
  uint32_t per_layer = 1;
#pragma unroll
  for (int d = 1; d < D; ++d) per_layer *= (uint32_t)g.size[d];
  const uint32_t first_cell = (D > 1 ? (uint32_t)g.own_lo * per_layer : 0u);
  const uint32_t own_cells = (D > 1 ? (uint32_t)g.own_n * per_layer : g.ncells);
  const char *__restrict__ posb = reinterpret_cast<const char *>(p.posb);
  const float pre_r2 = p.pre_r2;
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sm.lq[0][lane]);
  const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(&sm.stage[0][0][0]);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&sm.mbar[0]);
  const int S = g.size[L];
  // this lane's two slots of a chunk (the second one half a run away from the first: evens out the queue lengths)
  const uint32_t sA = (uint32_t)lane, sB = 32u + (uint32_t)((lane + 16) & 31);

  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t n_issued = 0, n_waited = 0; // chunk k uses buffer k & 1, completes phase (k >> 1) & 1 of its barrier
  bool generic_dirty = false;          // a buffer was written by plain stores since the last bulk copy

  auto decode_run = [&](int rid, int *od, int &wz) {
    int rem = rid;
    double gap2 = 0.0;
#pragma unroll
    for (int d = D - 2; d >= 0; --d) {
      const int span = 2 * p.w[d] + 1;
      od[d] = (rem % span) - p.w[d];
      rem /= span;
      const double gap = (double)max(abs(od[d]) - 1, 0) * g.side[d];
      gap2 += gap * gap;
    }
    wz = p.trim ? reach_last_dim(p.r2, gap2, g.side[L], p.w[L]) : p.w[L];
  };
  int od_first[DS], wz_first;
  od_first[0] = 0;
  decode_run(lane, od_first, wz_first);

  while (true) {
    uint32_t grab = 0;
    if (lane == 0) grab = atomicAdd(p.work_counter, p.grab);
    grab = __shfl_sync(0xFFFFFFFFu, grab, 0);
    if (grab >= own_cells) break;
    const uint32_t grab_end = min(grab + p.grab, own_cells);
    int tc[D];
    {
      const uint32_t cell = first_cell + grab;
      uint32_t rem = cell;
#pragma unroll
      for (int d = D - 1; d >= 0; --d) {
        tc[d] = (int)(rem % (uint32_t)g.size[d]);
        rem /= (uint32_t)g.size[d];
      }
      if (D > 1) {
        int gl = g.win_lo + (int)(cell / per_layer);
        if (gl < 0) gl += g.size[0];
        if (gl >= g.size[0]) gl -= g.size[0];
        tc[0] = gl;
      }
    }
    --tc[L];

    for (uint32_t cell = first_cell + grab; cell < first_cell + grab_end; ++cell) {
      if (++tc[L] == S) {
        if (D > 1) {
          tc[L] = 0;
          if (D > 2) {
            if (++tc[D > 2 ? 1 : 0] == g.size[D > 2 ? 1 : 0]) {
              tc[D > 2 ? 1 : 0] = 0;
              if (++tc[0] >= g.size[0]) tc[0] -= g.size[0];
            }
          } else {
            if (++tc[0] >= g.size[0]) tc[0] -= g.size[0];
          }
        }
      }
      const uint32_t rb = bbeg[cell], re = bend[cell];
      if (rb == re) continue;
      const int zlo = tc[L] - p.w[L], zhi = tc[L] + p.w[L];
      bool boundary = (zlo < 0) | (zhi >= S);
#pragma unroll
      for (int d = 0; d < D - 1; ++d) boundary |= (tc[d] - p.w[d] < 0) | (tc[d] + p.w[d] >= g.size[d]);
      double origin[D];
#pragma unroll
      for (int d = 0; d < D; ++d) origin[d] = g.bmin[d] + (double)(tc[d] - p.w[d]) * g.side[d];

      for (uint32_t r0 = rb; r0 < re; r0 += RB) {
        const int nr = (int)min((uint32_t)RB, re - r0);
        // ---- phase 1: candidate runs of the primary image, 32 runs per directory batch ----
        bool rows_ready = false;
        bool my_danger = false;
        uint32_t qa = q0;
        for (int rbase = 0; rbase < nslow; rbase += 32) {
          uint32_t len = 0, jb = 0;
          const int rid = rbase + lane;
          if (rid < nslow) {
            int od[DS], wz;
            if (rbase == 0) {
#pragma unroll
              for (int d = 0; d < DS; ++d) od[d] = od_first[d];
              wz = wz_first;
            } else {
              decode_run(rid, od, wz);
            }
            int nc[D];
            bool ok = true;
#pragma unroll
            for (int d = 0; d < D - 1; ++d) {
              const int u = tc[d] + od[d];
              ok &= (u >= 0) & (u < g.size[d]);
              nc[d] = u;
            }
            const int a = max(tc[L] - wz, 0), bnd = min(tc[L] + wz, S - 1);
            if (ok && wz >= 0 && a <= bnd) {
              nc[L] = a;
              const int c_lo = local_collapse<D>(g, nc);
              if (c_lo >= 0) {
                jb = bbeg[c_lo];
                len = bend[c_lo + (bnd - a)] - jb;
              }
            }
          }
          uint32_t pin = len;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pin, o);
            if (lane >= o) pin += t;
          }
          const uint32_t total = __shfl_sync(0xFFFFFFFFu, pin, 31);
          const uint32_t excl = pin - len;              // this lane's run = candidates excl .. pin-1 of the batch
          const uint32_t jdelta = jb - excl;            // j = k + jdelta inside the run
          const uint32_t nchunks = (total + CHUNK - 1) / CHUNK;

          // chunk c of this batch -> buffer n_issued & 1
          auto issue = [&](uint32_t c) {
            __syncwarp(); // every lane is done with the buffer's previous contents
            const uint32_t buf = n_issued & 1u;
            const uint32_t k0 = c * CHUNK;
            const uint32_t lo = max(excl, k0), hi = min(pin, k0 + CHUNK);
            if (generic_dirty) {
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              generic_dirty = false;
            }
            if (lane == 0) mbar_arrive_expect_tx(bar0 + buf * 8u, min((uint32_t)CHUNK, total - k0) * 32u);
            if (lo < hi) bulk_g2s(stage0 + buf * (CHUNK * 32u) + (lo - k0) * 32u, posb + (size_t)(lo + jdelta) * 32u, (hi - lo) * 32u, bar0 + buf * 8u);
            ++n_issued;
          };
          if (nchunks > 0) issue(0);

          if (!rows_ready) {
            // ---- rows of this batch (while the first chunk is in flight) ----
            rows_ready = true;
            if (lane < nr) {
#pragma unroll
              for (int d = 0; d < D; ++d) {
                const double r = pos[(size_t)(r0 + lane) * D + d];
                sm.rows0[d][lane] = r;
                const double fl = (r - g.bmin[d]) * g.inv_side[d];
                const double fr = fl - floor(fl);
                my_danger |= ((int)floor(fl) != tc[d]) | (fr < p.tolf[d]) | (fr > 1.0 - p.tolf[d]) | (fabs(fr - 0.5) < p.tolf[d]);
              }
            }
            if (lane == 0) sm.danger = 0;
            __syncwarp();
            if (lane < RB + 2) {
              float *rf = reinterpret_cast<float *>(&sm.rowsf2[0][0]);
#pragma unroll
              for (int d = 0; d < D; ++d)
                rf[(((lane >> 1) * 4) + d) * 2 + (lane & 1)] = lane < nr ? (float)(sm.rows0[d][lane] - origin[d]) : 3.0e18f;
            }
#pragma unroll
            for (int a = 0; a < NACC; ++a)
              for (int e = lane; e < nr * SPCOL; e += 32) (&sm.part[a][0][0])[e] = 0ull;
            __syncwarp();
          }

          for (uint32_t c = 0; c < nchunks; ++c) {
            if (c + 1 < nchunks) issue(c + 1);
            const uint32_t buf = n_waited & 1u;
            mbar_wait(bar0 + buf * 8u, (n_waited >> 1) & 1u);
            ++n_waited;
            const uint32_t k0 = c * CHUNK;
            const uint32_t nvalid = min((uint32_t)CHUNK, total - k0);
            const uint32_t stg = stage0 + buf * (CHUNK * 32u);
            if (NEEDJ) {
              // column index of each slot: pieces of the runs that intersect this chunk
              const uint32_t lo = max(excl, k0), hi = min(pin, k0 + CHUNK);
              uint32_t pm = __ballot_sync(0xFFFFFFFFu, lo < hi);
              uint32_t dA = 0, dB = 0;
              while (pm) {
                const int r = __ffs(pm) - 1;
                pm &= pm - 1;
                const uint32_t plo = __shfl_sync(0xFFFFFFFFu, lo, r) - k0;
                const uint32_t pd = __shfl_sync(0xFFFFFFFFu, jdelta, r) + k0;
                if (sA >= plo) dA = pd;
                if (sB >= plo) dB = pd;
              }
              sm.sj[sA] = sA + dA;
              sm.sj[sB] = sB + dB;
              __syncwarp();
            }
            float pj[2][D];
            const bool vA = sA < nvalid, vB = sB < nvalid;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t a = stg + (h ? sB : sA) * 32u;
              double rx, ry, rz = 0.0;
              asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rx), "=d"(ry) : "r"(a));
              if (D > 2) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rz) : "r"(a + 16u));
              const double rec[3] = {rx, ry, rz};
#pragma unroll
              for (int d = 0; d < D; ++d) pj[h][d] = (float)(rec[d] - origin[d]);
            }
            const SDrainCtx dc{p.b, p.r2lo, p.r2, stg};
            stest_rows<D, F, STATS>(sm, dc, pre_r2, f, lane, pj[0], pj[1], sA, sB, vA, vB, nr, sm.rows0, image_id0, r0, qa);
            // the records of this chunk leave the buffer with the next bulk copy: drain now
            if (__any_sync(0xFFFFFFFFu, qa != q0)) {
              sdrain<D, F, STATS>(sm, dc, f, lane, (qa - q0) >> 6, r0, sm.rows0, image_id0);
              qa = q0;
            }
          }
        }
        if (boundary) {
          // ---- phase 2 (buckets at a periodic boundary only): runs reached through a periodic
          //      image, cur = r + image * L exactly as src/Search.h:188-190; staged by plain loads ----
          int o[DS];
#pragma unroll
          for (int d = 0; d < D - 1; ++d) o[d] = -p.w[d];
          bool more = true;
          while (more) {
            int nc[D], img[D];
            bool ok_slow = true, slow_shifted = false;
            double gap2 = 0.0;
#pragma unroll
            for (int d = 0; d < D - 1; ++d) {
              const double gap = (double)max(abs(o[d]) - 1, 0) * g.side[d];
              gap2 += gap * gap;
              int u = tc[d] + o[d];
              img[d] = 0;
              if (u < 0) {
                u += g.size[d];
                img[d] = 1;
              } else if (u >= g.size[d]) {
                u -= g.size[d];
                img[d] = -1;
              }
              ok_slow &= (u >= 0) & (u < g.size[d]) & (img[d] == 0 || g.periodic[d]);
              slow_shifted |= (img[d] != 0);
              nc[d] = u;
            }
            const int wz = p.trim ? reach_last_dim(p.r2, gap2, g.side[L], p.w[L]) : p.w[L];
            if (ok_slow && wz >= 0) {
              for (int m = (g.periodic[L] ? -1 : 0); m <= (g.periodic[L] ? 1 : 0); ++m) {
                if (m == 0 && !slow_shifted) continue;
                const int a = max(tc[L] - wz, m * S), bnd = min(tc[L] + wz, m * S + S - 1);
                if (a > bnd) continue;
                img[L] = -m;
                nc[L] = a - m * S;
                const int c_lo = local_collapse<D>(g, nc);
                if (c_lo < 0) continue;
                const uint32_t jb = bbeg[c_lo], je = bend[c_lo + (bnd - a)];
                if (jb >= je) continue;
                __syncwarp();
                if (lane < nr) {
#pragma unroll
                  for (int d = 0; d < D; ++d) sm.rowsS[d][lane] = sm.rows0[d][lane] + (double)img[d] * g.L[d];
                }
                __syncwarp();
                const uint32_t image_id = STATS ? (uint32_t)image_linear_index<D>(g, img) : 0u;
                const SDrainCtx dc{p.b, p.r2lo, p.r2, stage0};
                generic_dirty = true;
                for (uint32_t cb = jb; cb < je; cb += CHUNK) {
                  float pj[2][D];
                  const uint32_t nvalid = min((uint32_t)CHUNK, je - cb);
                  const bool vA = sA < nvalid, vB = sB < nvalid;
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const uint32_t sl = h ? sB : sA;
                    const uint32_t jx = min(cb + sl, je - 1);
                    double rec[4];
                    ld_rec(p.posb, jx, rec[0], rec[1], rec[2], rec[3]);
                    const uint32_t a = stage0 + sl * 32u;
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(rec[0]), "d"(rec[1]) : "memory");
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a + 16u), "d"(rec[2]), "d"(rec[3]) : "memory");
                    if (NEEDJ) sm.sj[sl] = jx;
                    // pre-filter only: move the candidate by -image*L instead of the row by +image*L
#pragma unroll
                    for (int d = 0; d < D; ++d) pj[h][d] = (float)((rec[d] - (double)img[d] * g.L[d]) - origin[d]);
                  }
                  __syncwarp();
                  stest_rows<D, F, STATS>(sm, dc, pre_r2, f, lane, pj[0], pj[1], sA, sB, vA, vB, nr, sm.rowsS, image_id, r0, qa);
                  if (__any_sync(0xFFFFFFFFu, qa != q0)) {
                    sdrain<D, F, STATS>(sm, dc, f, lane, (qa - q0) >> 6, r0, sm.rowsS, image_id);
                    qa = q0;
                  }
                  __syncwarp();
                }
              }
            }
            more = false;
#pragma unroll
            for (int d = D - 2; d >= 0; --d) {
              if (!more) {
                if (++o[d] <= p.w[d]) {
                  more = true;
                } else {
                  o[d] = -p.w[d];
                }
              }
            }
          }
        }
        __syncwarp();

        // ---- reduce part[row][*] in a fixed (skewed, conflict-free) order ----
        const uint32_t dmask = sm.danger | __ballot_sync(0xFFFFFFFFu, my_danger);
        if (lane < nr) {
          const bool dangerous = (dmask >> lane) & 1u;
          if (dangerous) {
            const uint32_t slot = atomicAdd(p.danger_count, 1u);
            if (slot < p.danger_capacity) p.danger_list[slot] = r0 + lane;
          } else if (STATS) {
            unsigned long long c = 0, hsum = 0;
#pragma unroll
            for (int k = 0; k < SPCOL; ++k) {
              c += sm.part[0][lane][(k + lane) & (SPCOL - 1)];
              hsum += sm.part[1][lane][(k + lane) & (SPCOL - 1)];
            }
            if (p.stat_count) p.stat_count[r0 + lane] = (uint32_t)c;
            if (p.stat_hash) p.stat_hash[r0 + lane] = hsum;
          } else {
#pragma unroll
            for (int a2 = 0; a2 < NACC; ++a2) {
              double s = 0;
#pragma unroll
              for (int k = 0; k < SPCOL; ++k) s += *reinterpret_cast<double *>(&sm.part[a2][lane][(k + lane) & (SPCOL - 1)]);
              p.y[(size_t)(r0 + lane) * BR + a2] += s;
            }
          }
        }
        __syncwarp();
      }
    }
  }
}

// ---------------------------------------------------------------------------
// launcher, instantiated per (D, Functor) — in libabr.so for the built-in
// functors, in the user's nvcc-compiled TU for custom ones.
// ---------------------------------------------------------------------------
// symmetric kernel, second step: y += ytmp for the owned rows that keep their tiled result
// (flagged rows are recomputed from y by the exact walk that follows)
static __global__ void __launch_bounds__(256) k_sym_combine(const abr_matvec_plan p, int BR) {
  const Grid &g = p.q.g;
  uint32_t per_layer = 1;
  for (int d = 1; d < g.D; ++d) per_layer *= (uint32_t)g.size[d];
  const uint32_t first_cell = g.D > 1 ? (uint32_t)g.own_lo * per_layer : 0u;
  const uint32_t own_cells = g.D > 1 ? (uint32_t)g.own_n * per_layer : g.ncells;
  if (own_cells == 0) return;
  const uint32_t p0 = p.q.bucket_begin[first_cell], p1 = p.q.bucket_end[first_cell + own_cells - 1];
  const uint64_t total = (uint64_t)(p1 - p0) * BR;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t i = p0 + (uint32_t)(e / BR);
    if (!((p.row_bits[i >> 5] >> (i & 31u)) & 1u)) {
      const uint64_t idx = (uint64_t)p0 * BR + e;
      p.y[idx] += p.ytmp[idx];
    }
  }
}

template <int D, class F, bool STATS> inline int launch_plan(const abr_matvec_plan &p, const F &f) {
  cudaError_t e;
  if (p.use_tiled) {
    constexpr int SYMV = STATS ? 0 : symmetry<F>::value;
    const bool sym = SYMV != 0 && p.symmetric && p.ytmp && p.row_bits;
    const bool staged = p.variant == 1 && !sym && !p.xrow_pos;
    const size_t smem = (staged ? sizeof(StagedSmem<D, F, STATS>) : sizeof(WarpSmem<D, F, STATS>)) * TILED_WARPS;
    const bool gen = !sym && !staged && (p.xrow_pos || p.heavy_list);
    void (*kern)(const abr_matvec_plan, const F) = staged ? staged_kernel<D, F, STATS>
                                                    : (sym ? tiled_kernel<D, F, STATS, SYMV, false>
                                                           : (gen ? tiled_kernel<D, F, STATS, 0, true> : tiled_kernel<D, F, STATS, 0, false>));
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TILED_THREADS, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    {
      // ask for no more shared memory than the resident CTAs need: what is left of the
      // 256 KB array serves as L1
      int pct = (int)((100.0 * per_sm * (smem + 1024)) / (228.0 * 1024.0)) + 1;
      if (pct > 100) pct = 100;
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    const unsigned max_chunks = (p.q.g.ncells + p.grab * TILED_WARPS - 1) / (p.grab * TILED_WARPS);
    unsigned grid = (unsigned)(p.sm_count * per_sm);
    if (grid > max_chunks) grid = max_chunks;
    if (grid < 1) grid = 1;
    kern<<<grid, TILED_THREADS, smem, p.stream>>>(p, f);
    if (gen && p.heavy_list) {
      // heavy buckets of the first launch, one (bucket, row batch) per warp at a time; returns at once when there are none
      abr_matvec_plan ph = p;
      ph.heavy_phase = 1;
      kern<<<grid, TILED_THREADS, smem, p.stream>>>(ph, f);
    }
    if (sym) k_sym_combine<<<p.sm_count * 8, 256, 0, p.stream>>>(p, F::BR);
    // rows handed over by the tiled kernel: exact per-row walk
    abr_matvec_plan p2 = p;
    p2.walk_only_list = 1;
    // grid sized for the worst case (every row handed over: lattice inputs sit exactly on bucket faces);
    // blocks beyond the list length leave at once
    {
      unsigned wgrid = (unsigned)((p.n_rows + 127) / 128);
      const unsigned wmax = (unsigned)p.sm_count * 16u;
      if (wgrid > wmax) wgrid = wmax;
      if (wgrid < 1) wgrid = 1;
      walk_kernel<D, F, STATS><<<wgrid, 128, 0, p.stream>>>(p2, f);
    }
  } else {
    const unsigned grid = (unsigned)((p.n_rows + 127) / 128);
    if (grid > 0) walk_kernel<D, F, STATS><<<grid, 128, 0, p.stream>>>(p, f);
  }
  e = cudaGetLastError();
  return (int)e;
}

template <int D, class F> struct sparse_launcher {
  static int launch(const abr_matvec_plan *plan, const void *functor_host) {
    return launch_plan<D, F, false>(*plan, *static_cast<const F *>(functor_host));
  }
};

} // namespace abr
#endif
