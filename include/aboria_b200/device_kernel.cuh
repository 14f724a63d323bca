// Device functors for the sparse kernel operator: the built-in set selected
// by abr_kernel_desc::kernel_id, and the pattern a user follows for a custom
// one (the reference takes a host lambda f(dx, a, b),
// /root/reference/src/Operators.h:478-516; a GPU needs it as a device functor
// compiled by nvcc).
//
// A functor provides
//     static constexpr int BR, BC;                       // block size
//     __device__ void operator()(const double *dx,       // r_b - r_a (periodic image applied)
//                                double d2,              // sum dx^2, same order as the cut-off test
//                                uint32_t i, uint32_t j, // row / column particle index
//                                double *blk) const;     // BR x BC, row major
// and captures whatever per-particle columns it needs as raw device pointers
// (get<variable>(a) -> col[i]).
//
// Custom functor in a user .cu file:
//     struct MyKernel { static constexpr int BR = 1, BC = 1; const double *w;
//       static constexpr bool NEEDS_DX = false;   // optional: dx itself is not read
//       __device__ void operator()(const double *dx, double d2, uint32_t i, uint32_t j, double *blk) const
//       { blk[0] = w[i] * w[j] * exp(-d2); } };
//     MyKernel k{w_dev};
//     abr_sparse_matvec_custom(h, pos, n, 1, &abr::sparse_launcher<3, MyKernel>::launch, &k,
//                              1, 1, radius, nullptr, b, y, nullptr);
#ifndef ABORIA_B200_DEVICE_KERNEL_CUH_
#define ABORIA_B200_DEVICE_KERNEL_CUH_

#include "abr.h"
#include "aboria_b200/detail/matvec_kernels.cuh"

namespace abr {

// sqrt for a squared distance.  d2 == 0 (the self pair, present in every row of
// create_sparse_operator(p, p, ...)) would send the whole warp through the
// library's special-case path; route it around instead.  rsqrt * x is within
// 2 ulp of sqrt(x) — far inside the 1e-12 relative L2 budget of the product.
__device__ __forceinline__ double sqrt_d2(double d2) {
  const bool pos = d2 > 0.0;
  const double x = pos ? d2 : 1.0;
  const double s = x * rsqrt(x);
  return pos ? s : 0.0;
}

// Variants without the library's special-case paths, for arguments known to be
// normal numbers (the host checks the kernel's parameters before choosing them):
// hardware seed (MUFU.RSQ64H / RCP64H, ~2^-22) + one third-order correction, error
// ~1 ulp — inside the 1e-12 budget — and no divergent slow-path call in the drain.
__device__ __forceinline__ double sqrt_d2_fast(double d2) {
  const bool pos = d2 > 1e-280; // below: a coincident pair
  const double x = pos ? d2 : 1.0;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  const double c = fma(e, 0.375, 0.5);
  y = fma(c, y * e, y);
  const double s = x * y;
  return pos ? s : 0.0;
}
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
}

namespace functors {

// tests/operators.h:842-847
struct ConstSum {
  static constexpr int TILED_CTAS = 8;
  static constexpr bool NEEDS_DX = false;
  static constexpr int BR = 1, BC = 1;
  const double *s1, *s2;
  __device__ void operator()(const double *, double, uint32_t i, uint32_t j, double *blk) const {
    blk[0] = s1[i] + s2[j];
  }
};
// tests/operators.h:905-911
struct ConstSumDiff {
  static constexpr bool NEEDS_DX = false;
  static constexpr int BR = 2, BC = 1;
  const double *s1, *s2;
  __device__ void operator()(const double *, double, uint32_t i, uint32_t j, double *blk) const {
    blk[0] = s1[i] + s2[j];
    blk[1] = s1[i] - s2[j];
  }
};
// 1/(|dx| + eps)
struct InvDist {
  static constexpr int SYMMETRY = +1; // block(-dx, b, a) == +block(dx, a, b): eligible for the symmetric product
  static constexpr bool USES_J = false; // the column index is not read: the staged kernel skips its reconstruction
  static constexpr int TILED_CTAS = 8;
  static constexpr bool NEEDS_DX = false;
  static constexpr int BR = 1, BC = 1;
  double eps;
  __device__ void operator()(const double *, double d2, uint32_t, uint32_t, double *blk) const {
    blk[0] = __drcp_rn(sqrt_d2(d2) + eps);
  }
};
// the same for eps in [1e-100, 1e100] (|dx| + eps is then a normal number far from
// overflow): branch-free sqrt and reciprocal
struct InvDistFast {
  static constexpr int SYMMETRY = +1; // block(-dx, b, a) == +block(dx, a, b): eligible for the symmetric product
  static constexpr bool USES_J = false; // the column index is not read: the staged kernel skips its reconstruction
  static constexpr int TILED_CTAS = 8;
  static constexpr bool NEEDS_DX = false;
  static constexpr int BR = 1, BC = 1;
  double eps;
  __device__ void operator()(const double *, double d2, uint32_t, uint32_t, double *blk) const {
    blk[0] = rcp_fast(sqrt_d2_fast(d2) + eps);
  }
};
// tests/operators.h:251-256
struct InvDistAA {
  static constexpr bool NEEDS_DX = false;
  static constexpr int BR = 1, BC = 1;
  double eps;
  const double *ai, *aj;
  __device__ void operator()(const double *, double d2, uint32_t i, uint32_t j, double *blk) const {
    blk[0] = (ai[i] * aj[j]) * __drcp_rn(sqrt_d2(d2) + eps);
  }
};
// The functors below keep loop-invariant scalars (1/h, 1/h^D * wcon * mass, sigma^2 ...) as
// members computed once on the host by make(): inside the drain a division costs ~20
// instructions with a divergent slow path, a multiplication one.  Their values differ from
// the reference's expression order by a few ulp (inside the 1e-12 budget); cut-offs inside a
// kernel function (q <= 2) sit where the function is continuous and zero.
//
// tests/rbf_interpolation.h:310-313: pow(2 - r/h, 4) * (1 + 2 r/h)
struct WendlandC2 {
  static constexpr int SYMMETRY = +1; // block(-dx, b, a) == +block(dx, a, b): eligible for the symmetric product
  static constexpr bool USES_J = false; // the column index is not read: the staged kernel skips its reconstruction
  static constexpr int TILED_CTAS = 8;
  static constexpr bool NEEDS_DX = false;
  static constexpr int BR = 1, BC = 1;
  double inv_h;
  static WendlandC2 make(double h) { return WendlandC2{1.0 / h}; }
  __device__ void operator()(const double *, double d2, uint32_t, uint32_t, double *blk) const {
    const double q = sqrt_d2_fast(d2) * inv_h;
    const double t = 2.0 - q;
    const double t2 = t * t;
    blk[0] = (t2 * t2) * (1.0 + 2.0 * q);
  }
};
// SURVEY §8d c3 (tests/md.h:166-174 pattern): Lennard-Jones force, D x 1 block:
// 24 eps (2 (s/r)^12 - (s/r)^6) / r^2 * dx needs 1/r^2 only — no square root
template <int D> struct LJForce {
  static constexpr int SYMMETRY = -1; // block(-dx, b, a) == -block(dx, a, b): eligible for the symmetric product
  static constexpr bool USES_J = false; // the column index is not read: the staged kernel skips its reconstruction
  static constexpr int BR = D, BC = 1;
  double sigma2, eps24;
  static LJForce make(double sigma, double eps) { return LJForce{sigma * sigma, 24.0 * eps}; }
  __device__ void operator()(const double *dx, double d2, uint32_t, uint32_t, double *blk) const {
    const bool regular = d2 > 1e-280; // d2 == 0: the self pair, force 0
    double fmag = 0.0;
    if (regular) {
      const double inv = rcp_fast(d2);
      const double sr2 = sigma2 * inv;
      const double sr6 = sr2 * sr2 * sr2;
      fmag = eps24 * (2.0 * sr6 * sr6 - sr6) * inv;
    } else if (d2 != 0) { // closer than 1e-140: the plain expression (overflows like the reference's)
      const double sr2 = sigma2 / d2;
      const double sr6 = sr2 * sr2 * sr2;
      fmag = eps24 * (2.0 * sr6 * sr6 - sr6) / d2;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) blk[d] = fmag * dx[d];
  }
};
// tests/md.h:166-174: linear spring between overlapping discs/spheres, D x 1 block:
// -k (diameter / r - 1) dx for r != 0
template <int D> struct LinearSpring {
  static constexpr int SYMMETRY = -1; // block(-dx, b, a) == -block(dx, a, b): eligible for the symmetric product
  static constexpr bool USES_J = false; // the column index is not read: the staged kernel skips its reconstruction
  static constexpr int BR = D, BC = 1;
  double k, diameter;
  static LinearSpring make(double k, double diameter) { return LinearSpring{k, diameter}; }
  __device__ void operator()(const double *dx, double d2, uint32_t, uint32_t, double *blk) const {
    const bool regular = d2 > 1e-280;
    const double r = sqrt_d2_fast(d2);
    const double f = regular ? -k * (diameter * rcp_fast(regular ? r : 1.0) - 1.0) : 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) blk[d] = f * dx[d];
  }
};
// tests/sph.h:154-165 W_fun (Wendland), times the particle mass
template <int D> struct SphDensity {
  static constexpr int SYMMETRY = +1; // block(-dx, b, a) == +block(dx, a, b): eligible for the symmetric product
  static constexpr bool USES_J = false; // the column index is not read: the staged kernel skips its reconstruction
  static constexpr bool NEEDS_DX = false;
  static constexpr int BR = 1, BC = 1;
  double inv_h, pref; // pref = (1 / h^D) * wcon * mass
  static SphDensity make(double h, double mass, double wcon) {
    double hD = h;
    for (int d = 1; d < D; ++d) hD *= h;
    return SphDensity{1.0 / h, mass * ((1 / hD) * wcon)};
  }
  __device__ void operator()(const double *, double d2, uint32_t, uint32_t, double *blk) const {
    const double q = sqrt_d2_fast(d2) * inv_h;
    const double t = 2.0 - q;
    const double t2 = t * t;
    const double W = pref * (t2 * t2) * (1.0 + 2.0 * q);
    blk[0] = q <= 2.0 ? W : 0.0;
  }
};
// tests/sph.h:140-152 F_fun and the pressure term of :333-339, D x 1 block
template <int D> struct SphPressure {
  static constexpr int BR = D, BC = 1;
  double inv_h, pref; // pref = mass * (1 / h^(D+2)) * wcon
  const double *pdr2_i, *pdr2_j;
  static SphPressure make(double h, double mass, double wcon, const double *pi, const double *pj) {
    double hD2 = h * h;
    for (int d = 0; d < D; ++d) hD2 *= h;
    return SphPressure{1.0 / h, mass * ((1 / hD2) * wcon), pi, pj};
  }
  __device__ void operator()(const double *dx, double d2, uint32_t i, uint32_t j, double *blk) const {
    const bool regular = d2 > 1e-280; // r == 0: F_fun returns 0
    const double q = sqrt_d2_fast(d2) * inv_h;
    const double t = 2.0 - q;
    const double t3 = t * t * t;
    const double Fv = pref * (-4 * t3 * (1 + 2 * q) + 2 * (t3 * t)) * rcp_fast(regular ? q : 1.0);
    const double c = (regular && q <= 2.0) ? (pdr2_i[i] + pdr2_j[j]) * Fv : 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) blk[d] = c * dx[d];
  }
};

} // namespace functors
} // namespace abr
#endif
