// Eigen product glue for the B200 sparse operator: makes `Eigen::VectorXd y = K * b` and the
// matrix-free iterative solvers (Eigen::ConjugateGradient / BiCGSTAB / GMRES with
// IdentityPreconditioner, /root/reference/tests/rbf_interpolation.h:340-379) work on top of
// Aboria::SparseOperator / BlockOperator (include/aboria_b200/Aboria.h), exactly the way the
// reference wires its MatrixReplacement into Eigen:
//   Eigen::internal::traits<MatrixReplacement>            src/detail/Operators.h:50-56
//   MatrixReplacement (EigenBase, rows/cols, operator*)   src/Operators.h:75-160
//   generic_product_impl<...>::scaleAndAddTo              src/detail/Operators.h:206-232
//
// Include <Eigen/Core>, <Eigen/Sparse> (and <Eigen/IterativeLinearSolvers>) BEFORE this header.
// Eigen is not part of this image; tests/cpp/mock_eigen holds a minimal stand-in of the
// templates touched here so that the overload set is type-checked and exercised
// (tests/cpp/test_eigen_glue.cpp).  Against the real Eigen 3.3 nothing else is needed.
#ifndef ABORIA_B200_EIGEN_GLUE_H_
#define ABORIA_B200_EIGEN_GLUE_H_

#include <cassert>

#include "aboria_b200/Aboria.h"

namespace Aboria {

// The operator behind an Eigen-visible matrix type.  Holds the operator by value (operators hold
// references to their particle sets, like the reference's kernels, src/Kernels.h:133-134).
template <typename Operator> class MatrixReplacement : public Eigen::EigenBase<MatrixReplacement<Operator>> {
public:
  // compile-time information for Eigen (src/Operators.h:83-96)
  typedef double Scalar;
  typedef double RealScalar;
  typedef size_t Index;
  typedef int StorageIndex;
  enum {
    ColsAtCompileTime = Eigen::Dynamic,
    RowsAtCompileTime = Eigen::Dynamic,
    MaxColsAtCompileTime = Eigen::Dynamic,
    MaxRowsAtCompileTime = Eigen::Dynamic,
    IsRowMajor = false
  };
  explicit MatrixReplacement(const Operator &op) : op_(op) {}
  Index rows() const { return op_.rows(); }
  Index cols() const { return op_.cols(); }
  Index innerSize() const { return rows(); }
  Index outerSize() const { return cols(); }
  void resize(Index a_rows, Index a_cols) {
    assert((a_rows == 0 && a_cols == 0) || (a_rows == rows() && a_cols == cols()));
    (void)a_rows;
    (void)a_cols;
  }
  Scalar coeff(const Index i, const Index j) const { return op_.coeff(i, j); }
  // src/Operators.h:153-158
  template <typename Rhs> Eigen::Product<MatrixReplacement, Rhs, Eigen::AliasFreeProduct> operator*(const Eigen::MatrixBase<Rhs> &x) const {
    return Eigen::Product<MatrixReplacement, Rhs, Eigen::AliasFreeProduct>(*this, x.derived());
  }
  const Operator &get_operator() const { return op_; }

private:
  Operator op_;
};

// `auto A = make_eigen_operator(create_sparse_operator(...)); Eigen::VectorXd y = A * b;`
template <typename Operator> MatrixReplacement<Operator> make_eigen_operator(const Operator &op) { return MatrixReplacement<Operator>(op); }

} // namespace Aboria

namespace Eigen {
namespace internal {
// MatrixReplacement looks like a SparseMatrix, so it inherits its traits (src/detail/Operators.h:50-56)
template <typename Operator> struct traits<Aboria::MatrixReplacement<Operator>> : public Eigen::internal::traits<Eigen::SparseMatrix<double>> {};

// MatrixReplacement * dense vector (src/detail/Operators.h:206-232)
template <typename Rhs, typename Operator>
struct generic_product_impl<Aboria::MatrixReplacement<Operator>, Rhs, SparseShape, DenseShape, GemvProduct>
    : generic_product_impl_base<Aboria::MatrixReplacement<Operator>, Rhs, generic_product_impl<Aboria::MatrixReplacement<Operator>, Rhs>> {
  typedef typename Product<Aboria::MatrixReplacement<Operator>, Rhs>::Scalar Scalar;
  template <typename Dest> static void scaleAndAddTo(Dest &y, const Aboria::MatrixReplacement<Operator> &lhs, const Rhs &rhs, const Scalar &alpha) {
    // "y += alpha * lhs * rhs"; the iterative solvers always pass alpha == 1 (the reference asserts the same)
    assert(alpha == Scalar(1) && "scaling is not implemented");
    (void)alpha;
    // plain, contiguous copies: rhs may be an expression and y a block of a larger vector
    Eigen::Matrix<double, Eigen::Dynamic, 1> b = rhs;
    Eigen::Matrix<double, Eigen::Dynamic, 1> acc = y;
    lhs.get_operator().evaluate(acc, b); // acc += K b on the GPU (KernelSparse::evaluate, src/Kernels.h:720-751)
    y = acc;
  }
};
} // namespace internal
} // namespace Eigen

#endif
