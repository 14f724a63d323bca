// Host planner + built-in functor dispatch of the sparse kernel product.
// Kernels live in include/aboria_b200/detail/matvec_kernels.cuh.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

#include "abr_internal.h"
#include "aboria_b200/device_kernel.cuh"

#ifndef ABR_FAST_INVDIST
#define ABR_FAST_INVDIST 1
#endif

// This file is compiled once per dimension (Makefile: -DABR_MATVEC_D=1|2|3; the kernels are templates
// on D and on the functor, ~70 instantiations per dimension): the D = 3 object also carries the planner
// and the host entry points, the others only their `matvec_entry_d<k>` (the single-TU build without the
// macro still works).
#ifndef ABR_MATVEC_D
#define ABR_MATVEC_D 0 // everything in one translation unit
#endif
#define ABR_D_HERE(d) (ABR_MATVEC_D == 0 || ABR_MATVEC_D == (d))
#define ABR_MATVEC_MAIN (ABR_MATVEC_D == 0 || ABR_MATVEC_D == 3)

namespace abr {

// one entry point per dimension (defined at the end of this file, each in its own object)
struct MvArgs {
  const uint32_t *row_ptr;
  int32_t *col_idx;
  double *values;
  const uint64_t *ii, *jj;
  uint64_t m;
  double *out;
  int lnorm, transform_kind;
  const double *t_host;
};
enum { MV_BUILTIN = 0, MV_ASSEMBLE, MV_COEFF, MV_STATS, MV_NORM };
int matvec_entry_d1(Handle *h, int op, const abr_matvec_plan &p, const abr_kernel_desc *k, const MvArgs &a);
int matvec_entry_d2(Handle *h, int op, const abr_matvec_plan &p, const abr_kernel_desc *k, const MvArgs &a);
int matvec_entry_d3(Handle *h, int op, const abr_matvec_plan &p, const abr_kernel_desc *k, const MvArgs &a);
__attribute__((unused)) static inline int matvec_entry(Handle *h, int op, const abr_matvec_plan &p, const abr_kernel_desc *k, const MvArgs &a) {
  switch (h->D) {
  case 1: return matvec_entry_d1(h, op, p, k, a);
  case 2: return matvec_entry_d2(h, op, p, k, a);
  default: return matvec_entry_d3(h, op, p, k, a);
  }
}

#if ABR_MATVEC_MAIN

// (x, y, z, b) records for the tiled kernel: the drain gathers position and b of a
// column particle with ONE 256-bit load instead of four scattered 8-byte loads
// (the LSU data pipe was the binding limit, profiles/r1m_*).  Costs one streaming
// pass (64 B/particle) per product.
template <int D> __global__ void __launch_bounds__(256) k_pack_posb(const double *__restrict__ pos, const double *__restrict__ b, double *__restrict__ posb, uint32_t n, DevScalars *scal) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) { // reset the tile scheduler and the exact-walk list of the product that follows (no memset node: see abr_build.cu)
    scal->work_counter = 0;
    scal->danger_count = 0;
    scal->heavy_state = 0ull;
    scal->heavy_work = 0;
  }
  if (j >= n) return;
  double r[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < D; ++d) r[d] = pos[(size_t)j * D + d];
  if (b) r[3] = b[j];
  double4 v = make_double4(r[0], r[1], r[2], r[3]);
  reinterpret_cast<double4 *>(posb)[j] = v;
}

// rows whose point the internal row build dropped (outside the domain, non-finite): to the exact walk
__global__ void __launch_bounds__(256) k_rows_dropped(const uint8_t *__restrict__ alive, uint32_t n_rows, uint32_t *__restrict__ danger_count,
                                                      uint32_t *__restrict__ danger_list, uint32_t capacity) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows && !alive[i]) {
    const uint32_t slot = atomicAdd(danger_count, 1u);
    if (slot < capacity) danger_list[slot] = i;
  }
}

// Buckets the row points of a rows != columns product into the COLUMN grid: an ordinary build (keys, sort,
// bucket ranges, reorder of the position column) on an internal handle that shares the column set's grid but
// treats every dimension as non-periodic, so that row points outside the domain are dropped (they go to the
// exact walk, which searches from outside through the images like the reference) instead of being wrapped.
static int bucket_rows(Handle *h, const MatvecCall &c, abr_matvec_plan *p) {
  if (!h->rows_h) {
    h->rows_h = new Handle();
    h->rows_h->device = h->device;
    h->rows_h->sm_count = h->sm_count;
    cudaError_t e = cudaMalloc(&h->rows_h->d_scalars, sizeof(DevScalars));
    if (e == cudaSuccess) e = cudaMallocHost(&h->rows_h->h_scalars, 3 * sizeof(DevScalars));
    if (e != cudaSuccess) return check_cuda(h, e, "row build: scalars");
  }
  Handle *r = h->rows_h;
  r->stream = h->stream;
  r->domain_set = true;
  r->grid_forced = true;
  r->D = h->D;
  for (int d = 0; d < MAXD; ++d) {
    r->bmin[d] = h->bmin[d];
    r->bmax[d] = h->bmax[d];
    r->periodic[d] = false;
    r->size[d] = h->size[d];
    r->side[d] = h->side[d];
    r->inv_side[d] = h->inv_side[d];
  }
  r->n_leaf = h->n_leaf;
  const size_t n = c.n_rows;
  const size_t pos_bytes = n * (size_t)h->D * sizeof(double);
  ABR_CUDA(h, r->posb.reserve(2 * pos_bytes + 64));            // [input copy | sorted]
  ABR_CUDA(h, r->danger_list.reserve(n * sizeof(int32_t) + n)); // [order | alive]
  double *pin = r->posb.as<double>();
  double *psorted = reinterpret_cast<double *>(r->posb.as<char>() + ((pos_bytes + 63) / 64) * 64);
  int32_t *order = r->danger_list.as<int32_t>();
  uint8_t *alive = reinterpret_cast<uint8_t *>(order + n);
  ABR_CUDA(h, cudaMemcpyAsync(pin, c.row_pos, pos_bytes, cudaMemcpyDeviceToDevice, h->stream));
  fill_u32(h, reinterpret_cast<uint32_t *>(alive), 0x01010101u, (n + 3) / 4);
  const void *src[1] = {pin};
  void *dst[1] = {psorted};
  const size_t eb[1] = {(size_t)h->D * sizeof(double)};
  ReorderSpec spec{1, src, dst, eb};
  // asynchronous form: the alive count is never read on the host — the bucket ranges bound every access
  int rc = build_celllist(r, pin, alive, n, order, nullptr, &spec);
  if (rc) return set_error(h, rc, "row build: " + r->err);
  h->launches += r->launches;
  r->launches = 0;
  ABR_CUDA(h, h->danger_list.reserve(n * sizeof(uint32_t)));
  p->danger_list = h->danger_list.as<uint32_t>();
  p->danger_capacity = (uint32_t)n;
  p->xrow_pos = psorted;
  p->xrow_bb = r->bucket_begin.as<uint32_t>();
  p->xrow_be = r->bucket_end.as<uint32_t>();
  p->xrow_index = order;
  p->xrow_alive = alive;
  return ABR_OK;
}

// Fills the plan: picks the tiled kernel when its preconditions hold, and the
// rounding-safety margins that decide which rows go to the exact walk.
static int make_plan(Handle *h, const MatvecCall &c, int BR, abr_matvec_plan *p, int BC = 1) {
  if (!h->built) return set_error(h, ABR_ERR_STATE, "matvec: cell list has not been built");
  if (!h->pos_sorted && h->n_sorted > 0) return set_error(h, ABR_ERR_STATE, "matvec: abr_query_set_particles not called");
  if (c.n_rows >= 0xFFFFFFFFull) return set_error(h, ABR_ERR_UNSUPPORTED, "matvec: too many rows");
  memset(p, 0, sizeof(*p));
  p->q.g = h->grid();
  p->q.pos = h->pos_sorted;
  p->q.bucket_begin = h->bucket_begin.as<uint32_t>();
  p->q.bucket_end = h->bucket_end.as<uint32_t>();
  p->q.n = (uint32_t)h->n_sorted;
  p->row_pos = c.row_pos;
  p->n_rows = (uint32_t)c.n_rows;
  p->rows_are_cols = c.rows_are_cols;
  p->radius = c.radius;
  p->radius_per_row = c.radius_per_row;
  p->b = c.b;
  p->y = c.y;
  p->stat_count = c.count;
  p->stat_hash = c.hash;
  p->stream = h->stream;
  p->sm_count = h->sm_count;
  p->work_counter = &h->d_scalars->work_counter;
  p->danger_count = &h->d_scalars->danger_count;

  const int D = h->D;
  bool tiled = c.rows_are_cols && !c.radius_per_row && h->n_aliased == 0 &&
               c.row_pos == h->pos_sorted && c.n_rows == h->n_sorted && c.radius > 0 &&
               std::isfinite(c.radius) && h->n_sorted < (1ull << 28); // queue entries pack (j << 4) | row
  if (c.force_path == 1) tiled = false;
  if (c.force_path == 0 && !tiled)
    return set_error(h, ABR_ERR_INVALID, "tiled path needs rows_are_cols, a constant radius and no aliased keys");
  if (tiled) {
    // absolute slack of any coordinate computed on the path (a few ulps of the
    // largest magnitude involved), see DESIGN.md "exactness of the tiled kernel"
    double xmax = 0;
    for (int d = 0; d < D; ++d)
      xmax = std::fmax(xmax, std::fmax(std::fabs(h->bmin[d]), std::fabs(h->bmax[d])) + (h->bmax[d] - h->bmin[d]));
    xmax += c.radius;
    const double delta = 256.0 * 2.220446049250313e-16 * xmax;
    const double R = c.radius;
    p->r2 = R * R; // pow(r,2) == r*r
    p->r2lo = (R > 4 * delta) ? (R - 4 * delta) * (R - 4 * delta) : -1.0;
    double stencil = 1;
    for (int d = 0; d < D; ++d) {
      // half width of the bucket stencil.  The 1e-9 keeps r == k * side (a
      // common choice) at k instead of k + 1 when the product rounds up; rows
      // closer than tolf (>= 4e-9) to a bucket face are recomputed exactly anyway.
      const double wr = std::ceil(R * h->inv_side[d] - 1e-9);
      if (!(wr < 1e6)) {
        tiled = false;
        break;
      }
      p->w[d] = (int)wr;
      if (p->w[d] < 1) p->w[d] = 1;
      p->tolf[d] = std::fmax(4e-9, 8.0 * delta * h->inv_side[d]);
      if (p->tolf[d] > 0.125) tiled = false;
      stencil *= (2.0 * p->w[d] + 1.0);
      if (p->w[d] >= 2) p->trim = 1;
    }
    if (stencil > 8192.0) tiled = false; // radius >> bucket side: the walk is the better kernel
    if (tiled) {
      // fp32 pre-filter: coordinates relative to the stencil origin have magnitude
      // M <= (2w+2) side (+ slack); |dx| is known to ~2 ulp32(M) per component, so the
      // relative error of the fp32 |dx|^2 near the cut-off is <= 2 D * 2^-22 * M / R plus a
      // few ulp32 of arithmetic.  An 8x safety factor on top; tests compare pair-set
      // hashes with the oracle.
      double M = 0;
      for (int d = 0; d < D; ++d) M = std::fmax(M, (2.0 * p->w[d] + 2.0) * h->side[d]);
      const double tol = 8.0 * (2.0 * D * std::ldexp(1.0, -22) * M / R + std::ldexp(1.0, -20));
      double pr = p->r2 * (1.0 + tol);
      float prf = (float)pr;
      if ((double)prf < pr) prf = std::nextafterf(prf, INFINITY);
      if (!(prf < 1.0e30f) || !(p->r2 > 1e-30)) tiled = false; // out of comfortable fp32 range
      p->pre_r2 = prf;
    }
    if (!tiled && c.force_path == 0)
      return set_error(h, ABR_ERR_INVALID, "tiled path not applicable for this radius / grid");
  }
  // Rows that are NOT the column set (RBF evaluation at test points, tests/rbf_interpolation.h:326) with a
  // constant radius: bucket the row points into the column grid with an internal build and run the cell-tiled
  // kernel on them instead of one thread per row.  Row points outside the domain (the reference searches from
  // there too, through the images) and non-finite ones go to the exact walk.
  bool xrows = false;
  if (!tiled && c.force_path != 1 && !c.rows_are_cols && !c.radius_per_row && h->n_aliased == 0 && !h->windowed && c.radius > 0 &&
      std::isfinite(c.radius) && h->n_sorted > 0 && h->n_sorted < (1ull << 28) && c.n_rows >= h->xrows_min_n) {
    tiled = true;
    xrows = true;
  }
  if (xrows) {
    double xmax = 0;
    for (int d = 0; d < D; ++d)
      xmax = std::fmax(xmax, std::fmax(std::fabs(h->bmin[d]), std::fabs(h->bmax[d])) + (h->bmax[d] - h->bmin[d]));
    xmax += c.radius;
    const double delta = 256.0 * 2.220446049250313e-16 * xmax;
    const double R = c.radius;
    p->r2 = R * R;
    p->r2lo = (R > 4 * delta) ? (R - 4 * delta) * (R - 4 * delta) : -1.0;
    double stencil = 1;
    for (int d = 0; d < D && tiled; ++d) {
      const double wr = std::ceil(R * h->inv_side[d] - 1e-9);
      if (!(wr < 1e6)) {
        tiled = false;
        break;
      }
      p->w[d] = std::max(1, (int)wr);
      p->tolf[d] = std::fmax(4e-9, 8.0 * delta * h->inv_side[d]);
      if (p->tolf[d] > 0.125) tiled = false;
      stencil *= (2.0 * p->w[d] + 1.0);
      if (p->w[d] >= 2) p->trim = 1;
    }
    if (stencil > 8192.0) tiled = false;
    if (tiled) {
      double M = 0;
      for (int d = 0; d < D; ++d) M = std::fmax(M, (2.0 * p->w[d] + 2.0) * h->side[d]);
      const double tol = 8.0 * (2.0 * D * std::ldexp(1.0, -22) * M / R + std::ldexp(1.0, -20));
      double pr = p->r2 * (1.0 + tol);
      float prf = (float)pr;
      if ((double)prf < pr) prf = std::nextafterf(prf, INFINITY);
      if (!(prf < 1.0e30f) || !(p->r2 > 1e-30)) tiled = false;
      p->pre_r2 = prf;
    }
    if (tiled) {
      const int rc = bucket_rows(h, c, p);
      if (rc) return rc;
    } else {
      xrows = false;
    }
  }
  p->use_tiled = tiled ? 1 : 0;
  p->variant = h->matvec_variant;
  if (tiled && h->symmetric && c.b && c.y && !c.count && !c.hash) {
    // symmetric product (functors that declare SYMMETRY): zeroed accumulation scratch + one flag bit per row
    const size_t ny = (size_t)c.n_rows * (size_t)(BR > 0 ? BR : 1);
    ABR_CUDA(h, h->ytmp.reserve(ny * sizeof(double)));
    ABR_CUDA(h, h->row_bits.reserve(((size_t)c.n_rows / 32 + 2) * sizeof(uint32_t)));
    fill_u32(h, h->ytmp.as<uint32_t>(), 0u, ny * 2);
    fill_u32(h, h->row_bits.as<uint32_t>(), 0u, (uint64_t)c.n_rows / 32 + 2);
    p->symmetric = 1;
    p->ytmp = h->ytmp.as<double>();
    p->row_bits = h->row_bits.as<uint32_t>();
  }
  {
    // small grids: hand out single buckets so that every resident warp gets work
    const uint64_t warps = (uint64_t)h->sm_count * 32;
    uint64_t gsz = p->q.g.ncells / (4 * warps);
    p->grab = (uint32_t)(gsz < 1 ? 1 : (gsz > TILED_GRAB ? TILED_GRAB : gsz));
  }
  if (tiled) {
    if (!xrows) ABR_CUDA(h, h->danger_list.reserve((size_t)c.n_rows * sizeof(uint32_t)));
    p->danger_list = h->danger_list.as<uint32_t>();
    p->danger_capacity = (uint32_t)c.n_rows;
    // packed column records; b rides along when the block has one column
    ABR_CUDA(h, h->posb.reserve((size_t)h->n_sorted * 4 * sizeof(double) + 32));
    double *posb = h->posb.as<double>();
    const uint32_t n32 = (uint32_t)h->n_sorted;
    const double *bpack = (BC == 1 && c.b) ? c.b : nullptr;
    const unsigned gb = (n32 + 255) / 256 > 0 ? (n32 + 255) / 256 : 1;
    switch (D) {
    case 1: k_pack_posb<1><<<gb, 256, 0, h->stream>>>(h->pos_sorted, bpack, posb, n32, h->d_scalars); break;
    case 2: k_pack_posb<2><<<gb, 256, 0, h->stream>>>(h->pos_sorted, bpack, posb, n32, h->d_scalars); break;
    default: k_pack_posb<3><<<gb, 256, 0, h->stream>>>(h->pos_sorted, bpack, posb, n32, h->d_scalars); break;
    }
    h->launches += 1;
    p->posb = posb;
    // buckets with many rows (clustered clouds) are split into row-batch work items; `max_bucket` is what the
    // last VERIFIED build saw (an asynchronous build publishes it with abr_check_async) — a performance hint only
    if (h->max_bucket > HEAVY_ROWS || xrows) {
      const size_t cap = c.n_rows / HEAVY_ROWS + 2;
      ABR_CUDA(h, h->heavy_list.reserve(cap * sizeof(uint2)));
      p->heavy_list = h->heavy_list.as<uint2>();
      p->heavy_capacity = (uint32_t)cap;
      p->heavy_state = &h->d_scalars->heavy_state;
      p->heavy_work = &h->d_scalars->heavy_work;
    }
    if (xrows) { // after k_pack_posb has reset the exact-walk list
      k_rows_dropped<<<(unsigned)((c.n_rows + 255) / 256), 256, 0, h->stream>>>(p->xrow_alive, (uint32_t)c.n_rows, p->danger_count, p->danger_list,
                                                                                p->danger_capacity);
      h->launches += 1;
    }
  }
  (void)BR;
  return ABR_OK;
}

#endif // ABR_MATVEC_MAIN (planner)

template <int D, class F, bool STATS>
static int launch_checked(Handle *h, const abr_matvec_plan &p, const F &f) {
  const int e = launch_plan<D, F, STATS>(p, f);
  if (e != 0) return check_cuda(h, (cudaError_t)e, "matvec launch");
  const bool sym = !STATS && symmetry<F>::value != 0 && p.symmetric && p.ytmp && p.row_bits;
  const bool staged = p.variant == 1 && !sym && !p.xrow_pos;
  h->counters[2] = p.use_tiled ? (sym ? 3 : 2) + ((!sym && !staged && p.heavy_list) ? 1 : 0) : 1;
  h->launches += h->counters[2];
  return ABR_OK;
}

template <int D> static int dispatch_builtin(Handle *h, const abr_matvec_plan &p, const abr_kernel_desc *k) {
  using namespace functors;
  switch (k->kernel_id) {
  case ABR_K_CONST_SUM:
    if (k->block_rows != 1 || k->block_cols != 1) break;
    return launch_checked<D, ConstSum, false>(h, p, ConstSum{k->row_vars[0], k->col_vars[0]});
  case ABR_K_CONST_SUM_DIFF:
    if (k->block_rows != 2 || k->block_cols != 1) break;
    return launch_checked<D, ConstSumDiff, false>(h, p, ConstSumDiff{k->row_vars[0], k->col_vars[0]});
  case ABR_K_INV_DIST:
    if (k->block_rows != 1 || k->block_cols != 1) break;
    if (ABR_FAST_INVDIST && k->params[0] >= 1e-100 && k->params[0] <= 1e100 && p.use_tiled) return launch_checked<D, InvDistFast, false>(h, p, InvDistFast{k->params[0]});
    return launch_checked<D, InvDist, false>(h, p, InvDist{k->params[0]});
  case ABR_K_INV_DIST_AA:
    if (k->block_rows != 1 || k->block_cols != 1) break;
    return launch_checked<D, InvDistAA, false>(h, p, InvDistAA{k->params[0], k->row_vars[0], k->col_vars[0]});
  case ABR_K_WENDLAND_C2:
    if (k->block_rows != 1 || k->block_cols != 1) break;
    return launch_checked<D, WendlandC2, false>(h, p, WendlandC2::make(k->params[0]));
  case ABR_K_LJ_FORCE:
    if (k->block_rows != D || k->block_cols != 1) break;
    return launch_checked<D, LJForce<D>, false>(h, p, LJForce<D>::make(k->params[0], k->params[1]));
  case ABR_K_LINEAR_SPRING:
    if (k->block_rows != D || k->block_cols != 1) break;
    return launch_checked<D, LinearSpring<D>, false>(h, p, LinearSpring<D>::make(k->params[0], k->params[1]));
  case ABR_K_SPH_DENSITY:
    if (k->block_rows != 1 || k->block_cols != 1) break;
    return launch_checked<D, SphDensity<D>, false>(h, p, SphDensity<D>::make(k->params[0], k->params[1], k->params[2]));
  case ABR_K_SPH_PRESSURE:
    if (k->block_rows != D || k->block_cols != 1) break;
    return launch_checked<D, SphPressure<D>, false>(
        h, p, SphPressure<D>::make(k->params[0], k->params[1], k->params[2], k->row_vars[0], k->col_vars[0]));
  default:
    return set_error(h, ABR_ERR_INVALID, "matvec: unknown kernel_id");
  }
  return set_error(h, ABR_ERR_INVALID, "matvec: block_rows/block_cols do not match the kernel");
}

// ---- assemble (CSR) --------------------------------------------------------
template <int D> static int dispatch_assemble(Handle *h, const abr_matvec_plan &p, const abr_kernel_desc *k, const uint32_t *row_ptr,
                                             int32_t *col_idx, double *values) {
  using namespace functors;
  int e = -1;
  switch (k->kernel_id) {
  case ABR_K_CONST_SUM: e = launch_assemble<D>(p, ConstSum{k->row_vars[0], k->col_vars[0]}, row_ptr, col_idx, values); break;
  case ABR_K_CONST_SUM_DIFF: e = launch_assemble<D>(p, ConstSumDiff{k->row_vars[0], k->col_vars[0]}, row_ptr, col_idx, values); break;
  case ABR_K_INV_DIST: e = launch_assemble<D>(p, InvDist{k->params[0]}, row_ptr, col_idx, values); break;
  case ABR_K_INV_DIST_AA: e = launch_assemble<D>(p, InvDistAA{k->params[0], k->row_vars[0], k->col_vars[0]}, row_ptr, col_idx, values); break;
  case ABR_K_WENDLAND_C2: e = launch_assemble<D>(p, WendlandC2::make(k->params[0]), row_ptr, col_idx, values); break;
  case ABR_K_LJ_FORCE: e = launch_assemble<D>(p, LJForce<D>::make(k->params[0], k->params[1]), row_ptr, col_idx, values); break;
  case ABR_K_LINEAR_SPRING: e = launch_assemble<D>(p, LinearSpring<D>::make(k->params[0], k->params[1]), row_ptr, col_idx, values); break;
  case ABR_K_SPH_DENSITY: e = launch_assemble<D>(p, SphDensity<D>::make(k->params[0], k->params[1], k->params[2]), row_ptr, col_idx, values); break;
  case ABR_K_SPH_PRESSURE:
    e = launch_assemble<D>(p, SphPressure<D>::make(k->params[0], k->params[1], k->params[2], k->row_vars[0], k->col_vars[0]), row_ptr, col_idx, values);
    break;
  default: return set_error(h, ABR_ERR_INVALID, "assemble: unknown kernel_id");
  }
  if (e != 0) return check_cuda(h, (cudaError_t)e, "assemble launch");
  h->launches += 1;
  return ABR_OK;
}

// ---- coeff --------------------------------------------------------------------
template <int D> static int dispatch_coeff(Handle *h, const abr_matvec_plan &p, const abr_kernel_desc *k, const uint64_t *ii, const uint64_t *jj,
                                          uint64_t m, double *out) {
  using namespace functors;
  int e = -1;
  switch (k->kernel_id) {
  case ABR_K_CONST_SUM: e = launch_coeff<D>(p, ConstSum{k->row_vars[0], k->col_vars[0]}, ii, jj, m, out); break;
  case ABR_K_CONST_SUM_DIFF: e = launch_coeff<D>(p, ConstSumDiff{k->row_vars[0], k->col_vars[0]}, ii, jj, m, out); break;
  case ABR_K_INV_DIST: e = launch_coeff<D>(p, InvDist{k->params[0]}, ii, jj, m, out); break;
  case ABR_K_INV_DIST_AA: e = launch_coeff<D>(p, InvDistAA{k->params[0], k->row_vars[0], k->col_vars[0]}, ii, jj, m, out); break;
  case ABR_K_WENDLAND_C2: e = launch_coeff<D>(p, WendlandC2::make(k->params[0]), ii, jj, m, out); break;
  case ABR_K_LJ_FORCE: e = launch_coeff<D>(p, LJForce<D>::make(k->params[0], k->params[1]), ii, jj, m, out); break;
  case ABR_K_LINEAR_SPRING: e = launch_coeff<D>(p, LinearSpring<D>::make(k->params[0], k->params[1]), ii, jj, m, out); break;
  case ABR_K_SPH_DENSITY: e = launch_coeff<D>(p, SphDensity<D>::make(k->params[0], k->params[1], k->params[2]), ii, jj, m, out); break;
  case ABR_K_SPH_PRESSURE:
    e = launch_coeff<D>(p, SphPressure<D>::make(k->params[0], k->params[1], k->params[2], k->row_vars[0], k->col_vars[0]), ii, jj, m, out);
    break;
  default: return set_error(h, ABR_ERR_INVALID, "coeff: unknown kernel_id");
  }
  if (e != 0) return check_cuda(h, (cudaError_t)e, "coeff launch");
  h->launches += 1;
  return ABR_OK;
}

#if ABR_MATVEC_MAIN
int run_coeff(Handle *h, const MatvecCall &c, const abr_kernel_desc *k, const uint64_t *ii, const uint64_t *jj, size_t m, double *out) {
  if (!k) return set_error(h, ABR_ERR_INVALID, "coeff: null kernel descriptor");
  if (m == 0) return ABR_OK;
  if (!c.row_pos || !ii || !jj || !out) return set_error(h, ABR_ERR_INVALID, "coeff: null pointer");
  if (!h->domain_set) return set_error(h, ABR_ERR_STATE, "coeff: domain has not been set");
  if (!h->pos_sorted) return set_error(h, ABR_ERR_STATE, "coeff: abr_query_set_particles not called");
  abr_matvec_plan p;
  memset(&p, 0, sizeof(p));
  p.q.g = h->grid();
  p.q.pos = h->pos_sorted;
  p.q.n = (uint32_t)h->n_sorted;
  p.row_pos = c.row_pos;
  p.n_rows = (uint32_t)c.n_rows;
  p.radius = c.radius;
  p.radius_per_row = c.radius_per_row;
  p.stream = h->stream;
  MvArgs a{};
  a.ii = ii;
  a.jj = jj;
  a.m = m;
  a.out = out;
  return matvec_entry(h, MV_COEFF, p, k, a);
}

int scan_exclusive_u32(Handle *h, uint32_t *data, uint64_t m); // abr_build.cu

__global__ void k_zero_u64(unsigned long long *p) { *p = 0ull; }
__global__ void __launch_bounds__(256) k_sum_u32_64(const uint32_t *__restrict__ v, uint64_t n, unsigned long long *__restrict__ out) {
  unsigned long long acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

int run_assemble(Handle *h, const MatvecCall &c, const abr_kernel_desc *k, uint32_t *row_ptr, int32_t *col_idx, double *values,
                 size_t capacity, uint64_t *nnz_host) {
  if (!k || !row_ptr) return set_error(h, ABR_ERR_INVALID, "assemble: null pointer");
  if (c.n_rows >= 0xFFFFFFF0ull) return set_error(h, ABR_ERR_UNSUPPORTED, "assemble: too many rows");
  ABR_CUDA(h, cudaMemsetAsync(row_ptr, 0, (c.n_rows + 1) * sizeof(uint32_t), h->stream));
  uint64_t nnz = 0;
  if (c.n_rows > 0) {
    // 1. entries per row (pair_stats), 2. exclusive scan -> row_ptr, 3. exact walk fills the rows
    MatvecCall s = c;
    s.count = row_ptr;
    s.hash = nullptr;
    s.b = nullptr;
    s.y = nullptr;
    int rc = run_pair_stats(h, s);
    if (rc) return rc;
    // the row pointers are 32 bit: a total of 2^32 entries or more would wrap silently in the scan, so the
    // counts are summed in 64 bits first
    {
      unsigned long long *tot = reinterpret_cast<unsigned long long *>(&h->d_scalars->pair_count);
      k_zero_u64<<<1, 1, 0, h->stream>>>(tot);
      k_sum_u32_64<<<(unsigned)std::min<uint64_t>((c.n_rows + 255) / 256, (uint64_t)h->sm_count * 8), 256, 0, h->stream>>>(row_ptr, c.n_rows, tot);
      unsigned long long t64 = 0;
      ABR_CUDA(h, cudaMemcpyAsync(&t64, tot, sizeof(t64), cudaMemcpyDeviceToHost, h->stream));
      ABR_CUDA(h, cudaStreamSynchronize(h->stream));
      h->launches += 2;
      if (t64 >= 0xFFFFFFFFull) return set_error(h, ABR_ERR_UNSUPPORTED, "assemble: " + std::to_string(t64) + " entries do not fit 32-bit row pointers");
    }
    rc = scan_exclusive_u32(h, row_ptr, c.n_rows + 1);
    if (rc) return rc;
    uint32_t last = 0;
    ABR_CUDA(h, cudaMemcpyAsync(&last, row_ptr + c.n_rows, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    ABR_CUDA(h, cudaStreamSynchronize(h->stream));
    nnz = last;
  }
  if (nnz_host) *nnz_host = nnz;
  if (!col_idx || c.n_rows == 0) return ABR_OK; // count only
  if (capacity < nnz) return set_error(h, ABR_ERR_INVALID, "assemble: capacity smaller than the number of entries");
  abr_matvec_plan p;
  MatvecCall w = c;
  w.force_path = 1;
  w.count = nullptr;
  int rc = make_plan(h, w, k->block_rows, &p);
  if (rc) return rc;
  MvArgs a{};
  a.row_ptr = row_ptr;
  a.col_idx = col_idx;
  a.values = values;
  return matvec_entry(h, MV_ASSEMBLE, p, k, a);
}

int run_builtin_matvec(Handle *h, const MatvecCall &c, const abr_kernel_desc *k) {
  if (!k) return set_error(h, ABR_ERR_INVALID, "matvec: null kernel descriptor");
  if (c.n_rows == 0) return ABR_OK;
  if (!c.row_pos || !c.b || !c.y) return set_error(h, ABR_ERR_INVALID, "matvec: null pointer");
  abr_matvec_plan p;
  int rc = make_plan(h, c, k->block_rows, &p, k->block_cols);
  if (rc) return rc;
  return matvec_entry(h, MV_BUILTIN, p, k, MvArgs{});
}

int run_pair_stats(Handle *h, const MatvecCall &c) {
  if (c.n_rows == 0) return ABR_OK;
  if (!c.row_pos) return set_error(h, ABR_ERR_INVALID, "pair_stats: null pointer");
  abr_matvec_plan p;
  int rc = make_plan(h, c, 1, &p);
  if (rc) return rc;
  return matvec_entry(h, MV_STATS, p, nullptr, MvArgs{});
}
#endif // ABR_MATVEC_MAIN (coeff / assemble / product / stats entry points)

template <int D, int TK> static int launch_norm_stats_t(Handle *h, const abr_matvec_plan &p, int lnorm, const Xform &xf) {
  const unsigned grid = (unsigned)((p.n_rows + 127) / 128);
  switch (lnorm) {
  case -1: norm_stats_kernel<D, -1, TK><<<grid, 128, 0, p.stream>>>(p, xf); break;
  case 1: norm_stats_kernel<D, 1, TK><<<grid, 128, 0, p.stream>>>(p, xf); break;
  case 2: norm_stats_kernel<D, 2, TK><<<grid, 128, 0, p.stream>>>(p, xf); break;
  default: return set_error(h, ABR_ERR_UNSUPPORTED, "distance_search: norm must be -1 (Chebyshev), 1 (Manhattan) or 2 (Euclidean)");
  }
  h->launches += 1;
  ABR_CUDA(h, cudaGetLastError());
  return ABR_OK;
}

// transform_kind: 0 identity, 1 ScaleTransform (D factors), 2 LinearTransform (D x D matrix, row major)
template <int D> static int launch_norm_stats(Handle *h, const abr_matvec_plan &p, int lnorm, int transform_kind, const double *t_host) {
  Xform xf;
  memset(&xf, 0, sizeof(xf));
  if (transform_kind == 1) {
    for (int d = 0; d < D; ++d) xf.s[d] = t_host[d];
    return launch_norm_stats_t<D, 1>(h, p, lnorm, xf);
  }
  if (transform_kind == 2) {
    for (int e = 0; e < D * D; ++e) xf.m[e] = t_host[e];
    // LinearTransform constructor (src/Transform.h:82-99): the vertex of [-1,1]^D whose image is
    // longest, lattice order (last dimension fastest), strict '>'
    double best = 0;
    for (int code = 0; code < (1 << D); ++code) {
      double pt[D], q[D];
      int bits[D];
      for (int j = 0; j < D; ++j) {
        bits[j] = (code >> (D - 1 - j)) & 1;
        pt[j] = bits[j] ? 1.0 : -1.0;
      }
      xform_point<D, 2>(xf, pt, q);
      double n2 = 0;
      for (int j = 0; j < D; ++j) n2 += q[j] * q[j];
      if (n2 > best) {
        for (int j = 0; j < D; ++j) xf.eig[j] = bits[j];
        best = n2;
      }
    }
    return launch_norm_stats_t<D, 2>(h, p, lnorm, xf);
  }
  return launch_norm_stats_t<D, 0>(h, p, lnorm, xf);
}

#if ABR_MATVEC_MAIN
int run_norm_stats(Handle *h, const MatvecCall &c, int lnorm, int transform_kind, const double *t_host) {
  if (transform_kind != 0 && !t_host) return set_error(h, ABR_ERR_INVALID, "distance_search: null transform");
  if (c.n_rows == 0) return ABR_OK;
  if (!c.row_pos) return set_error(h, ABR_ERR_INVALID, "distance_search: null pointer");
  abr_matvec_plan p;
  MatvecCall w = c;
  w.force_path = 1;
  int rc = make_plan(h, w, 1, &p);
  if (rc) return rc;
  MvArgs a{};
  a.lnorm = lnorm;
  a.transform_kind = transform_kind;
  a.t_host = t_host;
  return matvec_entry(h, MV_NORM, p, nullptr, a);
}

int run_custom_matvec(Handle *h, const MatvecCall &c, abr_launch_fn launch, const void *functor,
                      int BR, int BC) {
  if (!launch || !functor) return set_error(h, ABR_ERR_INVALID, "custom matvec: null launcher/functor");
  if (c.n_rows == 0) return ABR_OK;
  if (!c.row_pos || !c.b || !c.y) return set_error(h, ABR_ERR_INVALID, "matvec: null pointer");
  abr_matvec_plan p;
  int rc = make_plan(h, c, BR, &p, BC);
  if (rc) return rc;
  const int e = launch(&p, functor);
  if (e != 0) return check_cuda(h, (cudaError_t)e, "custom matvec launch");
  h->counters[2] = p.use_tiled ? 2 + ((p.heavy_list && !(p.variant == 1 && !p.xrow_pos)) ? 1 : 0) : 1;
  h->launches += h->counters[2];
  return ABR_OK;
}

#endif // ABR_MATVEC_MAIN (norm stats / custom entry points)

template <int D> static int matvec_entry_t(Handle *h, int op, const abr_matvec_plan &p, const abr_kernel_desc *k, const MvArgs &a) {
  switch (op) {
  case MV_BUILTIN: return dispatch_builtin<D>(h, p, k);
  case MV_ASSEMBLE: return dispatch_assemble<D>(h, p, k, a.row_ptr, a.col_idx, a.values);
  case MV_COEFF: return dispatch_coeff<D>(h, p, k, a.ii, a.jj, a.m, a.out);
  case MV_STATS: return launch_checked<D, StatsFunctor, true>(h, p, StatsFunctor());
  default: return launch_norm_stats<D>(h, p, a.lnorm, a.transform_kind, a.t_host);
  }
}
#if ABR_D_HERE(1)
int matvec_entry_d1(Handle *h, int op, const abr_matvec_plan &p, const abr_kernel_desc *k, const MvArgs &a) { return matvec_entry_t<1>(h, op, p, k, a); }
#endif
#if ABR_D_HERE(2)
int matvec_entry_d2(Handle *h, int op, const abr_matvec_plan &p, const abr_kernel_desc *k, const MvArgs &a) { return matvec_entry_t<2>(h, op, p, k, a); }
#endif
#if ABR_D_HERE(3)
int matvec_entry_d3(Handle *h, int op, const abr_matvec_plan &p, const abr_kernel_desc *k, const MvArgs &a) { return matvec_entry_t<3>(h, op, p, k, a); }
#endif

} // namespace abr
