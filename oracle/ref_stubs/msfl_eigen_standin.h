// msfl_eigen_standin.h -- TEST INFRASTRUCTURE.  A minimal stand-in for the part of Eigen 3.3 that the reference's
// scan-matching sources use, so that those files can be compiled UNMODIFIED from /root/reference in an image that has
// no Eigen (oracle/Makefile, target `ref`):
//   src/slam/local/scan_matching/{lidar_factor,odometry_scan_matcher,mapping_scan_matcher,scan_matcher}.cc,
//   src/slam/imu_fusion/{pose_local_parameterization,scan_undistortion}.cc and the headers they pull in
//   (utility.h, rigid_transform.h, common.h, timestamped_pointcloud.h, integration_base.h, estimator.h ...).
//
// What this pins and what it does not.  Every expression of the reference's own code -- which operands, which
// products, which signs, the Jacobian block layout, the parameter-block order, the association loops, gates and
// thresholds, what is handed to the solver -- comes from the reference's source text.  The elementary operations
// underneath are restated here from Eigen's published definitions (Geometry/Quaternion.h: _transformVector,
// toRotationMatrix, quaternion product, slerp, conjugate; Geometry/OrthoMethods.h: cross; dense coefficient-wise
// arithmetic and small fixed-size products in index order) and evaluate eagerly into fixed-size temporaries instead of
// expression templates.  The two decompositions the matcher calls (SelfAdjointEigenSolver<Matrix3d>,
// colPivHouseholderQr on 5x3) delegate to the oracle's restatements (msflo_sym_eig3 / msflo_lstsq_5x3, themselves
// checked against numpy eigh / lstsq in tests/test_oracle.py): third-party numerics stay "restated, not pinned".
// Nothing outside oracle/ includes this file.
#ifndef MSFL_EIGEN_STANDIN_H
#define MSFL_EIGEN_STANDIN_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

extern "C" {  // oracle/msfl_oracle.h (the restated third-party decompositions)
void msflo_sym_eig3(const double A[9], double evals[3], double evecs[9]);
void msflo_lstsq_5x3(const double A[15], const double b[5], double x[3]);
}

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 __attribute__((aligned(16)))

namespace Eigen {

const int Dynamic = -1;
enum StorageOptions { ColMajor = 0, RowMajor = 1 };

template <typename S, int R, int C, int O = ColMajor>
class Matrix;
template <typename T>
class Map;
template <typename X, int BR, int BC>
class Block;
template <typename S>
class Quaternion;

// compile-time shape of every dense type (usable while the type itself is still incomplete, i.e. inside MatrixBase<D>)
template <typename T>
struct traits;
template <typename S, int R, int C, int O>
struct traits<Matrix<S, R, C, O>> {
  typedef S Scalar;
  static constexpr int Rows = R, Cols = C;
};
template <typename S, int R, int C, int O>
struct traits<Map<Matrix<S, R, C, O>>> : traits<Matrix<S, R, C, O>> {};
template <typename S, int R, int C, int O>
struct traits<Map<const Matrix<S, R, C, O>>> : traits<Matrix<S, R, C, O>> {};
template <typename X, int BR, int BC>
struct traits<Block<X, BR, BC>> {
  typedef typename traits<X>::Scalar Scalar;
  static constexpr int Rows = BR, Cols = BC;
};

template <typename D>
class CommaInitializer {
 public:
  typedef typename traits<D>::Scalar S;
  CommaInitializer(D &m, S v) : m_(m), k_(0) { put(v); }
  CommaInitializer &operator,(S v) {
    put(v);
    return *this;
  }

 private:
  void put(S v) {
    m_.coeffRef(k_ / D::Cols, k_ % D::Cols) = v;  // row by row, like Eigen's comma initialiser
    ++k_;
  }
  D &m_;
  int k_;
};

template <typename D>
struct RowwiseOp;
template <typename D>
struct ColwiseOp;
template <typename M>
struct ColPivQRStandin;
template <typename S, int N>
class Array;

template <typename D>
class MatrixBase {
 public:
  typedef typename traits<D>::Scalar Scalar;
  static constexpr int Rows = traits<D>::Rows, Cols = traits<D>::Cols;
  const D &derived() const { return *static_cast<const D *>(this); }
  D &derived() { return *static_cast<D *>(this); }

  auto operator()(int i, int j) const { return derived().coeff(i, j); }
  auto &operator()(int i, int j) { return derived().coeffRef(i, j); }
  auto operator()(int i) const { return Cols == 1 ? derived().coeff(i, 0) : derived().coeff(0, i); }
  auto &operator()(int i) { return Cols == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
  auto operator[](int i) const { return (*this)(i); }
  auto &operator[](int i) { return (*this)(i); }
  auto x() const { return (*this)(0); }
  auto y() const { return (*this)(1); }
  auto z() const { return (*this)(2); }

  auto eval() const {
    Matrix<Scalar, Rows, Cols> m;
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) m.coeffRef(i, j) = derived().coeff(i, j);
    return m;
  }
  template <typename T>
  auto cast() const {
    Matrix<T, Rows, Cols> m;
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) m.coeffRef(i, j) = (T)derived().coeff(i, j);
    return m;
  }
  auto squaredNorm() const {
    Scalar s = 0;
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) s += derived().coeff(i, j) * derived().coeff(i, j);
    return s;
  }
  auto norm() const { return std::sqrt(squaredNorm()); }
  auto normalized() const {  // Eigen: n = squaredNorm(); n > 0 ? *this / sqrt(n) : *this
    auto m = eval();
    const auto n = squaredNorm();
    if (n > 0) {
      const auto d = std::sqrt(n);
      for (int j = 0; j < Cols; ++j)
        for (int i = 0; i < Rows; ++i) m.coeffRef(i, j) = m.coeff(i, j) / d;
    }
    return m;
  }
  void normalize() { derived() = normalized(); }
  template <typename O>
  auto dot(const MatrixBase<O> &o) const {
    static_assert(Rows * Cols == O::Rows * O::Cols, "dot: size mismatch");
    Scalar s = 0;
    for (int i = 0; i < Rows * Cols; ++i) s += (*this)(i) * o(i);
    return s;
  }
  template <typename O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O> &o) const;
  auto transpose() const {
    Matrix<Scalar, Cols, Rows> m;
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) m.coeffRef(j, i) = derived().coeff(i, j);
    return m;
  }
  D &setConstant(Scalar v) {
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) = v;
    return derived();
  }
  D &setZero() { return setConstant(0); }
  D &setIdentity() {
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) = i == j ? 1 : 0;
    return derived();
  }
  // sub-views: writable on a non-const object, a copy on a const one
  template <int BR, int BC>
  Block<D, BR, BC> block(int r, int c) {
    return Block<D, BR, BC>(derived(), r, c);
  }
  template <int BR, int BC>
  auto block(int r, int c) const {
    Matrix<Scalar, BR, BC> m;
    for (int j = 0; j < BC; ++j)
      for (int i = 0; i < BR; ++i) m.coeffRef(i, j) = derived().coeff(r + i, c + j);
    return m;
  }
  template <int N>
  Block<D, N, 1> head() {
    static_assert(Cols == 1, "head: column vectors only");
    return Block<D, N, 1>(derived(), 0, 0);
  }
  template <int N>
  auto head() const {
    return this->template block<N, 1>(0, 0);
  }
  Block<D, Rows, 1> col(int j) { return Block<D, Rows, 1>(derived(), 0, j); }
  auto col(int j) const { return this->template block<Rows, 1>(0, j); }
  Block<D, 1, Cols> row(int i) { return Block<D, 1, Cols>(derived(), i, 0); }
  auto row(int i) const { return this->template block<1, Cols>(i, 0); }
  auto array() const {
    static_assert(Cols == 1, "array(): column vectors only in this stand-in");
    Array<Scalar, Rows> a;
    for (int i = 0; i < Rows; ++i) a(i) = derived().coeff(i, 0);
    return a;
  }
  template <typename O>
  D &operator+=(const MatrixBase<O> &o) {
    static_assert(Rows == O::Rows && Cols == O::Cols, "+=: size mismatch");
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) += o.derived().coeff(i, j);
    return derived();
  }
  RowwiseOp<D> rowwise() const { return RowwiseOp<D>{derived()}; }
  ColwiseOp<D> colwise() const { return ColwiseOp<D>{derived()}; }
  ColPivQRStandin<D> colPivHouseholderQr() const { return ColPivQRStandin<D>{derived()}; }
  CommaInitializer<D> operator<<(Scalar v) { return CommaInitializer<D>(derived(), v); }

 protected:
  // dense assignment; a column vector may be assigned to a row vector and vice versa (Eigen's implicit transposition
  // of vectors, used by lidar_factor.cc:38 "block<1, 3>(0, 0) = last_plane_N_")
  template <typename O>
  void assign_from(const MatrixBase<O> &o) {
    constexpr bool same = Rows == O::Rows && Cols == O::Cols;
    constexpr bool tvec = Rows == O::Cols && Cols == O::Rows && (Rows == 1 || Cols == 1);
    static_assert(same || tvec, "assignment: size mismatch");
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) = same ? o.derived().coeff(i, j) : o.derived().coeff(j, i);
  }
};

template <typename S, int R, int C, int O>
class Matrix : public MatrixBase<Matrix<S, R, C, O>> {
 public:
  typedef S Scalar;
  static constexpr int Rows = R, Cols = C, Options = O;
  Matrix() {
    for (int i = 0; i < R * C; ++i) d_[i] = 0;
  }
  Matrix(S x, S y, S z) {
    static_assert(R * C == 3, "3-vector constructor");
    d_[0] = x, d_[1] = y, d_[2] = z;
  }
  Matrix(const Matrix &o) = default;
  template <typename X>
  Matrix(const MatrixBase<X> &o) {
    this->assign_from(o);
  }
  Matrix &operator=(const Matrix &o) = default;
  template <typename X>
  Matrix &operator=(const MatrixBase<X> &o) {
    this->assign_from(o);
    return *this;
  }
  S coeff(int i, int j) const { return O == RowMajor ? d_[i * C + j] : d_[j * R + i]; }
  S &coeffRef(int i, int j) { return O == RowMajor ? d_[i * C + j] : d_[j * R + i]; }
  S *data() { return d_; }
  const S *data() const { return d_; }
  static Matrix Identity() {
    Matrix m;
    m.setIdentity();
    return m;
  }
  static Matrix Zero() { return Matrix(); }
  static Matrix Ones() {
    Matrix m;
    m.setConstant(1);
    return m;
  }

 private:
  S d_[R * C];
};

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, Dynamic, 1> VectorXd;

// writable view of caller memory
template <typename S, int R, int C, int O>
class Map<Matrix<S, R, C, O>> : public MatrixBase<Map<Matrix<S, R, C, O>>> {
 public:
  typedef S Scalar;
  static constexpr int Rows = R, Cols = C, Options = O;
  explicit Map(S *p) : p_(p) {}
  Map(const Map &o) = default;
  Map &operator=(const Map &o) {
    this->assign_from(o);
    return *this;
  }
  template <typename X>
  Map &operator=(const MatrixBase<X> &o) {
    this->assign_from(o);
    return *this;
  }
  S coeff(int i, int j) const { return O == RowMajor ? p_[i * C + j] : p_[j * R + i]; }
  S &coeffRef(int i, int j) { return O == RowMajor ? p_[i * C + j] : p_[j * R + i]; }

 private:
  S *p_;
};

// read-only view
template <typename S, int R, int C, int O>
class Map<const Matrix<S, R, C, O>> : public MatrixBase<Map<const Matrix<S, R, C, O>>> {
 public:
  typedef S Scalar;
  static constexpr int Rows = R, Cols = C, Options = O;
  explicit Map(const S *p) : p_(p) {}
  S coeff(int i, int j) const { return O == RowMajor ? p_[i * C + j] : p_[j * R + i]; }

 private:
  const S *p_;
};

// the one dynamic-size use (scan_matcher.cc:45, inside the never-called RefineByRejectOutliersWithFrac)
template <>
class Map<Matrix<double, Dynamic, 1, ColMajor>> {
 public:
  Map(double *p, int n) : p_(p), n_(n) {}
  double norm() const {
    double s = 0;
    for (int i = 0; i < n_; ++i) s += p_[i] * p_[i];
    return std::sqrt(s);
  }

 private:
  double *p_;
  int n_;
};

template <typename X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC>> {
 public:
  typedef typename X::Scalar Scalar;
  static constexpr int Rows = BR, Cols = BC;
  Block(X &x, int r, int c) : x_(x), r_(r), c_(c) {}
  Block(const Block &o) = default;
  Block &operator=(const Block &o) {
    this->assign_from(o);
    return *this;
  }
  template <typename O>
  Block &operator=(const MatrixBase<O> &o) {
    this->assign_from(o);
    return *this;
  }
  Scalar coeff(int i, int j) const { return x_.coeff(r_ + i, c_ + j); }
  Scalar &coeffRef(int i, int j) { return x_.coeffRef(r_ + i, c_ + j); }

 private:
  X &x_;
  int r_, c_;
};

// ---- arithmetic: evaluated eagerly, element order = index order ----
#define MSFL_STANDIN_CWISE(OP)                                                                                          \
  template <typename A, typename B>                                                                                     \
  auto operator OP(const MatrixBase<A> &a, const MatrixBase<B> &b) {                                                    \
    static_assert(A::Rows == B::Rows && A::Cols == B::Cols, "size mismatch");                                           \
    Matrix<typename A::Scalar, A::Rows, A::Cols> m;                                                                     \
    for (int j = 0; j < A::Cols; ++j)                                                                                   \
      for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = a.derived().coeff(i, j) OP b.derived().coeff(i, j);          \
    return m;                                                                                                           \
  }
MSFL_STANDIN_CWISE(+)
MSFL_STANDIN_CWISE(-)
#undef MSFL_STANDIN_CWISE
template <typename A>
auto operator-(const MatrixBase<A> &a) {
  Matrix<typename A::Scalar, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = -a.derived().coeff(i, j);
  return m;
}
template <typename A, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
auto operator*(const MatrixBase<A> &a, T s) {
  Matrix<typename A::Scalar, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = a.derived().coeff(i, j) * (typename A::Scalar)s;
  return m;
}
template <typename A, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
auto operator*(T s, const MatrixBase<A> &a) {
  Matrix<typename A::Scalar, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = (typename A::Scalar)s * a.derived().coeff(i, j);
  return m;
}
template <typename A, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
auto operator/(const MatrixBase<A> &a, T s) {
  Matrix<typename A::Scalar, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = a.derived().coeff(i, j) / (typename A::Scalar)s;
  return m;
}
template <typename A, typename B>
auto operator*(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  static_assert(A::Cols == B::Rows, "*: inner size mismatch");
  Matrix<typename A::Scalar, A::Rows, B::Cols> m;
  for (int j = 0; j < B::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) {
      typename A::Scalar s = a.derived().coeff(i, 0) * b.derived().coeff(0, j);
      for (int k = 1; k < A::Cols; ++k) s += a.derived().coeff(i, k) * b.derived().coeff(k, j);
      m.coeffRef(i, j) = s;
    }
  return m;
}
template <typename D>
template <typename O>
Matrix<typename traits<D>::Scalar, 3, 1> MatrixBase<D>::cross(const MatrixBase<O> &o) const {
  static_assert(Rows * Cols == 3 && O::Rows * O::Cols == 3, "cross: 3-vectors only");
  const MatrixBase<D> &a = *this;
  return Matrix<Scalar, 3, 1>(a(1) * o(2) - a(2) * o(1), a(2) * o(0) - a(0) * o(2), a(0) * o(1) - a(1) * o(0));
}
template <typename D>
std::ostream &operator<<(std::ostream &os, const MatrixBase<D> &m) {
  for (int i = 0; i < D::Rows; ++i)
    for (int j = 0; j < D::Cols; ++j) os << m.derived().coeff(i, j) << (j + 1 < D::Cols ? " " : (i + 1 < D::Rows ? "\n" : ""));
  return os;
}

// ---- partial reductions (mapping_scan_matcher.cc:136-137, :210) ----
template <typename D>
struct RowwiseOp {
  const D &m;
  auto mean() const {  // mean of every row: sum / Cols
    Matrix<typename D::Scalar, D::Rows, 1> r;
    for (int i = 0; i < D::Rows; ++i) {
      typename D::Scalar s = m.coeff(i, 0);
      for (int j = 1; j < D::Cols; ++j) s += m.coeff(i, j);
      r.coeffRef(i, 0) = s / (typename D::Scalar)D::Cols;
    }
    return r;
  }
};
template <typename D>
struct ColwiseOp {
  const D &m;
  auto mean() const {  // mean of every column: sum / Rows
    Matrix<typename D::Scalar, 1, D::Cols> r;
    for (int j = 0; j < D::Cols; ++j) {
      typename D::Scalar s = m.coeff(0, j);
      for (int i = 1; i < D::Rows; ++i) s += m.coeff(i, j);
      r.coeffRef(0, j) = s / (typename D::Scalar)D::Rows;
    }
    return r;
  }
  template <typename V>
  auto operator-(const MatrixBase<V> &v) const {  // subtract a column vector from every column
    static_assert(V::Rows == D::Rows && V::Cols == 1, "colwise() - v: v must be a column of matching height");
    Matrix<typename D::Scalar, D::Rows, D::Cols> r;
    for (int j = 0; j < D::Cols; ++j)
      for (int i = 0; i < D::Rows; ++i) r.coeffRef(i, j) = m.coeff(i, j) - v.derived().coeff(i, 0);
    return r;
  }
};

// ---- the two decompositions of the map matcher: delegated to the oracle's restatements ----
template <typename M>
class SelfAdjointEigenSolver {
 public:
  template <typename D>
  explicit SelfAdjointEigenSolver(const MatrixBase<D> &a) {
    static_assert(D::Rows == 3 && D::Cols == 3, "stand-in: 3x3 only (mapping_scan_matcher.cc:141)");
    double A[9], ev[3], V[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A[i * 3 + j] = a.derived().coeff(i, j);
    msflo_sym_eig3(A, ev, V);  // ascending eigenvalues, columns = eigenvectors (row-major storage)
    for (int i = 0; i < 3; ++i) {
      vals_.coeffRef(i, 0) = ev[i];
      for (int j = 0; j < 3; ++j) vecs_.coeffRef(i, j) = V[i * 3 + j];
    }
  }
  const Matrix<double, 3, 1> &eigenvalues() const { return vals_; }
  const Matrix<double, 3, 3> &eigenvectors() const { return vecs_; }

 private:
  Matrix<double, 3, 1> vals_;
  Matrix<double, 3, 3> vecs_;
};
template <typename M>
struct ColPivQRStandin {
  const M &a;
  template <typename B>
  Matrix<double, 3, 1> solve(const MatrixBase<B> &b) const {
    static_assert(M::Rows == 5 && M::Cols == 3 && B::Rows == 5 && B::Cols == 1, "stand-in: 5x3 least squares only (mapping_scan_matcher.cc:209)");
    double A[15], bb[5], x[3];
    for (int i = 0; i < 5; ++i) {
      bb[i] = b.derived().coeff(i, 0);
      for (int j = 0; j < 3; ++j) A[i * 3 + j] = a.coeff(i, j);
    }
    msflo_lstsq_5x3(A, bb, x);
    return Matrix<double, 3, 1>(x[0], x[1], x[2]);
  }
};

// ---- coefficient-wise arrays (slam/map/hybrid_grid.cc: cell indices) ----
template <typename S, int N>
class Array;
template <int N>
struct BoolArray {
  bool v[N];
  bool all() const {
    for (int i = 0; i < N; ++i)
      if (!v[i]) return false;
    return true;
  }
  bool any() const {
    for (int i = 0; i < N; ++i)
      if (v[i]) return true;
    return false;
  }
};
template <typename S, int N>
class Array {
 public:
  typedef S Scalar;
  Array() {
    for (int i = 0; i < N; ++i) d_[i] = 0;
  }
  Array(S x, S y, S z) : d_{x, y, z} { static_assert(N == 3, "3-array constructor"); }
  template <typename D>
  Array(const MatrixBase<D> &m) {  // Eigen converts between the matrix and array worlds on assignment / construction
    static_assert(D::Rows * D::Cols == N, "Array(matrix): size mismatch");
    for (int i = 0; i < N; ++i) d_[i] = m(i);
  }
  S x() const { return d_[0]; }
  S y() const { return d_[1]; }
  S z() const { return d_[2]; }
  S operator()(int i) const { return d_[i]; }
  S &operator()(int i) { return d_[i]; }
  Matrix<S, N, 1> matrix() const {
    Matrix<S, N, 1> m;
    for (int i = 0; i < N; ++i) m.coeffRef(i, 0) = d_[i];
    return m;
  }
  operator Matrix<S, N, 1>() const { return matrix(); }
  template <typename T>
  Array<T, N> cast() const {
    Array<T, N> r;
    for (int i = 0; i < N; ++i) r(i) = (T)d_[i];
    return r;
  }

 private:
  S d_[N];
};
#define MSFL_STANDIN_ARRAY_OP(OP)                                                  \
  template <typename S, int N>                                                     \
  Array<S, N> operator OP(const Array<S, N> &a, const Array<S, N> &b) {            \
    Array<S, N> r;                                                                 \
    for (int i = 0; i < N; ++i) r(i) = a(i) OP b(i);                               \
    return r;                                                                      \
  }                                                                                \
  template <typename S, int N, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type> \
  Array<S, N> operator OP(const Array<S, N> &a, T s) {                             \
    Array<S, N> r;                                                                 \
    for (int i = 0; i < N; ++i) r(i) = a(i) OP(S) s;                               \
    return r;                                                                      \
  }
MSFL_STANDIN_ARRAY_OP(+)
MSFL_STANDIN_ARRAY_OP(-)
MSFL_STANDIN_ARRAY_OP(*)
MSFL_STANDIN_ARRAY_OP(/)
#undef MSFL_STANDIN_ARRAY_OP
#define MSFL_STANDIN_ARRAY_CMP(OP)                                                 \
  template <typename S, int N, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type> \
  BoolArray<N> operator OP(const Array<S, N> &a, T s) {                            \
    BoolArray<N> r;                                                                \
    for (int i = 0; i < N; ++i) r.v[i] = a(i) OP(S) s;                             \
    return r;                                                                      \
  }
MSFL_STANDIN_ARRAY_CMP(>=)
MSFL_STANDIN_ARRAY_CMP(<)
#undef MSFL_STANDIN_ARRAY_CMP
typedef Array<int, 3> Array3i;
typedef Array<float, 3> Array3f;
// read-only array view of three floats (pcl point getArray3fMap()); converts to a vector where one is expected
template <>
class Map<const Array<float, 3>> {
 public:
  explicit Map(const float *p) : p_(p) {}
  operator Matrix<float, 3, 1>() const { return Matrix<float, 3, 1>(p_[0], p_[1], p_[2]); }

 private:
  const float *p_;
};

// ---- quaternions: coefficients stored x, y, z, w (Eigen's layout; Map<Quaterniond>(x + 3) relies on it) ----
template <typename D>
struct quat_scalar;
template <typename S>
struct quat_scalar<Quaternion<S>> {
  typedef S type;
};
template <typename S>
struct quat_scalar<Map<Quaternion<S>>> {
  typedef S type;
};
template <typename S>
struct quat_scalar<Map<const Quaternion<S>>> {
  typedef S type;
};

template <typename D>
class QuaternionBase {
 public:
  typedef typename quat_scalar<D>::type S;
  const D &derived() const { return *static_cast<const D *>(this); }
  D &derived() { return *static_cast<D *>(this); }
  S x() const { return derived().data()[0]; }
  S y() const { return derived().data()[1]; }
  S z() const { return derived().data()[2]; }
  S w() const { return derived().data()[3]; }
  Matrix<S, 3, 1> vec() const { return Matrix<S, 3, 1>(x(), y(), z()); }
  Matrix<S, 4, 1> coeffs() const {
    Matrix<S, 4, 1> c;
    c(0) = x(), c(1) = y(), c(2) = z(), c(3) = w();
    return c;
  }
  S squaredNorm() const { return x() * x() + y() * y() + z() * z() + w() * w(); }
  S norm() const { return std::sqrt(squaredNorm()); }
  Quaternion<S> normalized() const {  // coeffs / norm
    const S n = norm();
    return Quaternion<S>(w() / n, x() / n, y() / n, z() / n);
  }
  Quaternion<S> conjugate() const { return Quaternion<S>(w(), -x(), -y(), -z()); }
  template <typename O>
  S dot(const QuaternionBase<O> &o) const {
    return x() * o.x() + y() * o.y() + z() * o.z() + w() * o.w();
  }
  template <typename O>
  Quaternion<S> operator*(const QuaternionBase<O> &b) const {  // Eigen's quat_product
    const QuaternionBase<D> &a = *this;
    return Quaternion<S>(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                         a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                         a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                         a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  // rotate a 3-vector: Eigen's QuaternionBase::_transformVector
  template <typename V>
  Matrix<S, 3, 1> operator*(const MatrixBase<V> &v) const {
    const Matrix<S, 3, 1> q = vec();
    const Matrix<S, 3, 1> uv = S(2) * q.cross(v);
    return v + w() * uv + q.cross(uv);
  }
  // Eigen's QuaternionBase::slerp
  template <typename O>
  Quaternion<S> slerp(const S &t, const QuaternionBase<O> &other) const {
    const S one = S(1) - std::numeric_limits<S>::epsilon();
    const S d = this->dot(other);
    const S absD = std::abs(d);
    S scale0, scale1;
    if (absD >= one) {
      scale0 = S(1) - t;
      scale1 = t;
    } else {
      const S theta = std::acos(absD);
      const S sinTheta = std::sin(theta);
      scale0 = std::sin((S(1) - t) * theta) / sinTheta;
      scale1 = std::sin((t * theta)) / sinTheta;
    }
    if (d < S(0)) scale1 = -scale1;
    return Quaternion<S>(scale0 * w() + scale1 * other.w(), scale0 * x() + scale1 * other.x(), scale0 * y() + scale1 * other.y(),
                         scale0 * z() + scale1 * other.z());
  }
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> r;
    const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
    const S twx = tx * w(), twy = ty * w(), twz = tz * w();
    const S txx = tx * x(), txy = ty * x(), txz = tz * x();
    const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    r.coeffRef(0, 0) = S(1) - (tyy + tzz);
    r.coeffRef(0, 1) = txy - twz;
    r.coeffRef(0, 2) = txz + twy;
    r.coeffRef(1, 0) = txy + twz;
    r.coeffRef(1, 1) = S(1) - (txx + tzz);
    r.coeffRef(1, 2) = tyz - twx;
    r.coeffRef(2, 0) = txz - twy;
    r.coeffRef(2, 1) = tyz + twx;
    r.coeffRef(2, 2) = S(1) - (txx + tyy);
    return r;
  }
  template <typename T>
  Quaternion<T> cast() const {
    return Quaternion<T>((T)w(), (T)x(), (T)y(), (T)z());
  }
};

template <typename S>
class Quaternion : public QuaternionBase<Quaternion<S>> {
 public:
  typedef S Scalar;
  Quaternion() : d_{0, 0, 0, 1} {}
  Quaternion(S w, S x, S y, S z) : d_{x, y, z, w} {}
  Quaternion(const Quaternion &o) = default;
  template <typename O>
  Quaternion(const QuaternionBase<O> &o) : d_{(S)o.x(), (S)o.y(), (S)o.z(), (S)o.w()} {}
  template <typename V>
  explicit Quaternion(const MatrixBase<V> &c) : d_{c(0), c(1), c(2), c(3)} {  // from a 4-vector of coefficients x y z w
    static_assert(V::Rows * V::Cols == 4, "Quaternion(coeffs): 4-vector");
  }
  Quaternion &operator=(const Quaternion &o) = default;
  static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
  const S *data() const { return d_; }
  S *data() { return d_; }

 private:
  S d_[4];
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <typename S>
class Map<Quaternion<S>> : public QuaternionBase<Map<Quaternion<S>>> {
 public:
  typedef S Scalar;
  explicit Map(S *p) : p_(p) {}
  template <typename O>
  Map &operator=(const QuaternionBase<O> &o) {
    const S x = o.x(), y = o.y(), z = o.z(), w = o.w();
    p_[0] = x, p_[1] = y, p_[2] = z, p_[3] = w;
    return *this;
  }
  const S *data() const { return p_; }
  S *data() { return p_; }

 private:
  S *p_;
};
template <typename S>
class Map<const Quaternion<S>> : public QuaternionBase<Map<const Quaternion<S>>> {
 public:
  typedef S Scalar;
  explicit Map(const S *p) : p_(p) {}
  const S *data() const { return p_; }

 private:
  const S *p_;
};

}  // namespace Eigen
#endif
