// The Eigen product glue (include/aboria_b200/EigenGlue.h) type-checked and run against the
// stand-in under tests/cpp/mock_eigen (Eigen itself is not in this image): `y = K * b` through
// Eigen::Product -> generic_product_impl::scaleAndAddTo -> SparseOperator::evaluate, the
// reference's own call chain (/root/reference/src/detail/Operators.h:206-232), on the golden
// operator of tests/operators.h:810-933 and on a conjugate-gradient loop written against the
// Eigen-facing interface (the shape of tests/rbf_interpolation.h:340-379).
#include <cmath>
#include <cstdio>

#include <Eigen/Core>
#include <Eigen/Sparse>

#include "aboria_b200/EigenGlue.h"

using namespace Aboria;

int main() {
  ABORIA_VARIABLE(scalar1, double, "scalar1")
  ABORIA_VARIABLE(scalar2, double, "scalar2")
  typedef Particles<std::tuple<scalar1, scalar2>> ParticlesType;
  typedef position_d<3> position;
  ParticlesType particles;
  const double diameter = 0.1;
  ParticlesType::value_type p;
  for (int i = 0; i < 3; ++i) {
    get<position>(p) = vdouble3(diameter * 0.9 * i, 0, 0);
    get<scalar1>(p) = 1.0;
    get<scalar2>(p) = 2.0;
    particles.push_back(p);
  }
  particles.init_neighbour_search(vdouble3::Constant(-1), vdouble3::Constant(1), vbool3::Constant(false));
  auto A = make_eigen_operator(create_sparse_operator(particles, particles, diameter, kernels::const_sum<scalar1, scalar2>()));
  int failures = 0;
  Eigen::VectorXd v(3);
  v[0] = 1;
  v[1] = 2;
  v[2] = 3;
  Eigen::VectorXd ans = A * v; // Product -> generic_product_impl -> evaluate
  const double expect[3] = {9.0, 18.0, 15.0};
  for (int i = 0; i < 3; ++i)
    if (ans[i] != expect[i]) {
      std::printf("FAIL golden product: ans[%d] = %g\n", i, ans[i]);
      ++failures;
    }
  if (A.rows() != 3 || A.cols() != 3 || A.coeff(0, 1) != 3.0 || A.coeff(0, 2) != 0.0) {
    std::printf("FAIL rows/cols/coeff\n");
    ++failures;
  }
  // a CG loop on (I + 0.05 K): symmetric positive definite for this K
  auto apply = [&](const Eigen::VectorXd &x) {
    Eigen::VectorXd y = A * x;
    for (size_t i = 0; i < y.size(); ++i) y[i] = x[i] + 0.05 * y[i];
    return y;
  };
  Eigen::VectorXd x(3), r(3), d(3);
  x.setZero();
  for (int i = 0; i < 3; ++i) r[i] = d[i] = v[i];
  double rr = 0;
  for (int i = 0; i < 3; ++i) rr += r[i] * r[i];
  for (int it = 0; it < 10 && rr > 1e-28; ++it) {
    Eigen::VectorXd Ad = apply(d);
    double dAd = 0;
    for (int i = 0; i < 3; ++i) dAd += d[i] * Ad[i];
    const double alpha = rr / dAd;
    double rr_new = 0;
    for (int i = 0; i < 3; ++i) {
      x[i] += alpha * d[i];
      r[i] -= alpha * Ad[i];
      rr_new += r[i] * r[i];
    }
    for (int i = 0; i < 3; ++i) d[i] = r[i] + rr_new / rr * d[i];
    rr = rr_new;
  }
  Eigen::VectorXd check = apply(x);
  for (int i = 0; i < 3; ++i)
    if (std::fabs(check[i] - v[i]) > 1e-12) {
      std::printf("FAIL cg residual %d: %g\n", i, check[i] - v[i]);
      ++failures;
    }
  if (failures == 0) std::printf("eigen glue tests passed\n");
  return failures;
}
