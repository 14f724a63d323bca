// Test TU for the user-functor path of the sparse operator product: "the user kernel
// lambda, compiled as a device functor" (north_star; the reference takes a host lambda
// f(dx, a, b), /root/reference/src/Operators.h:478-516, evaluated in
// src/Kernels.h:737-749).  This file plays the USER's role: it defines a functor that
// reads a per-particle column of the row and of the column particle — the pattern
// `get<w>(a) * get<w>(b) / (dx.norm() + eps)` of tests/operators.h:251-256 — compiles it
// with nvcc against include/aboria_b200/device_kernel.cuh, and hands
// abr::sparse_launcher<D, F>::launch to abr_sparse_matvec_custom.  The planning (tiled
// kernel vs exact walk, danger rows, ...) stays inside libabr.so.
// Driven from tests/test_custom_functor.py, which checks the result against the oracle.
#include "aboria_b200/device_kernel.cuh"

namespace {

struct WeightedInvDist {
  static constexpr int BR = 1, BC = 1;
  const double *w_row; // get<w>(a) -> w_row[i]
  const double *w_col; // get<w>(b) -> w_col[j]
  double eps;
  __device__ void operator()(const double *dx, double d2, uint32_t i, uint32_t j, double *blk) const {
    (void)d2;
    double n2 = 0;
    for (int d = 0; d < 3; ++d) n2 += dx[d] * dx[d]; // dx itself is used, not only |dx|^2
    blk[0] = (w_row[i] * w_col[j]) / (sqrt(n2) + eps);
  }
};

// a 2 x 1 block functor through the same path (BR != 1: the block kernels' row batches)
struct SumDiffBlock {
  static constexpr int BR = 2, BC = 1;
  const double *s_row, *s_col;
  __device__ void operator()(const double *, double, uint32_t i, uint32_t j, double *blk) const {
    blk[0] = s_row[i] + s_col[j];
    blk[1] = s_row[i] - s_col[j];
  }
};

} // namespace

extern "C" int custom_weighted_inv_dist(abr_handle h, const double *row_pos, size_t n_rows, int rows_are_cols, const double *w_row,
                                        const double *w_col, double eps, double radius, const double *b, double *y,
                                        uint64_t *n_pairs_host) {
  WeightedInvDist f{w_row, w_col, eps};
  return abr_sparse_matvec_custom(h, row_pos, n_rows, rows_are_cols, &abr::sparse_launcher<3, WeightedInvDist>::launch, &f, 1, 1, radius,
                                  nullptr, b, y, n_pairs_host);
}

extern "C" int custom_sum_diff(abr_handle h, const double *row_pos, size_t n_rows, int rows_are_cols, const double *s_row,
                               const double *s_col, double radius, const double *b, double *y) {
  SumDiffBlock f{s_row, s_col};
  return abr_sparse_matvec_custom(h, row_pos, n_rows, rows_are_cols, &abr::sparse_launcher<3, SumDiffBlock>::launch, &f, 2, 1, radius,
                                  nullptr, b, y, nullptr);
}
