// msfl_eigen_standin.h -- TEST INFRASTRUCTURE.  A minimal stand-in for the part of Eigen 3.3 that the reference's
// factor sources use (src/slam/local/scan_matching/lidar_factor.{h,cc}, src/slam/imu_fusion/utility.h,
// src/slam/imu_fusion/pose_local_parameterization.cc), so that those files can be compiled UNMODIFIED from
// /root/reference in an image that has no Eigen (oracle/Makefile, target _ref/libmsfl_ref_factors.so).
//
// What this pins and what it does not: every expression of the reference's Evaluate() / Plus() / deltaQ() --
// which operands, which products, which signs, the Jacobian block layout, the parameter-block ordering -- comes from
// the reference's own source text.  The elementary operations underneath (3-vector cross / dot, 3x3 product, unit
// quaternion * vector, quaternion -> rotation matrix, quaternion product, normalisation) are restated here from
// Eigen's published definitions (Eigen/src/Geometry/Quaternion.h: _transformVector, toRotationMatrix, operator*;
// Eigen/src/Geometry/OrthoMethods.h: cross) and evaluate eagerly into fixed-size temporaries instead of expression
// templates -- the same arithmetic in the same order for these sizes.  Nothing outside oracle/ includes this file.
#ifndef MSFL_EIGEN_STANDIN_H
#define MSFL_EIGEN_STANDIN_H

#include <cmath>
#include <cstddef>

namespace Eigen {

enum StorageOptions { ColMajor = 0, RowMajor = 1 };

template <typename S, int R, int C, int O = ColMajor>
class Matrix;
template <typename T>
class Map;
template <typename X, int BR, int BC>
class Block;
template <typename S>
class Quaternion;

template <typename D>
class CommaInitializer {
 public:
  CommaInitializer(D &m, double v) : m_(m), k_(0) { put(v); }
  CommaInitializer &operator,(double v) {
    put(v);
    return *this;
  }

 private:
  void put(double v) {
    m_.coeffRef(k_ / D::Cols, k_ % D::Cols) = v;  // row by row, like Eigen's comma initialiser
    ++k_;
  }
  D &m_;
  int k_;
};

template <typename D>
class MatrixBase {
 public:
  const D &derived() const { return *static_cast<const D *>(this); }
  D &derived() { return *static_cast<D *>(this); }

  double operator()(int i, int j) const { return derived().coeff(i, j); }
  double &operator()(int i, int j) { return derived().coeffRef(i, j); }
  double operator()(int i) const { return D::Cols == 1 ? derived().coeff(i, 0) : derived().coeff(0, i); }
  double &operator()(int i) { return D::Cols == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
  double x() const { return (*this)(0); }
  double y() const { return (*this)(1); }
  double z() const { return (*this)(2); }

  auto eval() const {
    Matrix<double, D::Rows, D::Cols> m;
    for (int j = 0; j < D::Cols; ++j)
      for (int i = 0; i < D::Rows; ++i) m.coeffRef(i, j) = derived().coeff(i, j);
    return m;
  }
  double squaredNorm() const {
    double s = 0;
    for (int j = 0; j < D::Cols; ++j)
      for (int i = 0; i < D::Rows; ++i) s += derived().coeff(i, j) * derived().coeff(i, j);
    return s;
  }
  double norm() const { return std::sqrt(squaredNorm()); }
  auto normalized() const {  // Eigen: n = squaredNorm(); n > 0 ? *this / sqrt(n) : *this
    auto m = eval();
    const double n = squaredNorm();
    if (n > 0) {
      const double d = std::sqrt(n);
      for (int j = 0; j < D::Cols; ++j)
        for (int i = 0; i < D::Rows; ++i) m.coeffRef(i, j) = m.coeff(i, j) / d;
    }
    return m;
  }
  template <typename O>
  double dot(const MatrixBase<O> &o) const {
    static_assert(D::Rows * D::Cols == O::Rows * O::Cols, "dot: size mismatch");
    double s = 0;
    for (int i = 0; i < D::Rows * D::Cols; ++i) s += (*this)(i) * o(i);
    return s;
  }
  template <typename O>
  Matrix<double, 3, 1> cross(const MatrixBase<O> &o) const;
  auto transpose() const {
    Matrix<double, D::Cols, D::Rows> m;
    for (int j = 0; j < D::Cols; ++j)
      for (int i = 0; i < D::Rows; ++i) m.coeffRef(j, i) = derived().coeff(i, j);
    return m;
  }
  D &setConstant(double v) {
    for (int j = 0; j < D::Cols; ++j)
      for (int i = 0; i < D::Rows; ++i) derived().coeffRef(i, j) = v;
    return derived();
  }
  D &setZero() { return setConstant(0.0); }
  D &setIdentity() {
    for (int j = 0; j < D::Cols; ++j)
      for (int i = 0; i < D::Rows; ++i) derived().coeffRef(i, j) = i == j ? 1.0 : 0.0;
    return derived();
  }
  template <int BR, int BC>
  Block<D, BR, BC> block(int r, int c) {
    return Block<D, BR, BC>(derived(), r, c);
  }
  CommaInitializer<D> operator<<(double v) { return CommaInitializer<D>(derived(), v); }

 protected:
  // dense assignment; a column vector may be assigned to a row vector and vice versa (Eigen's implicit transposition
  // of vectors, used by lidar_factor.cc:38 "block<1, 3>(0, 0) = last_plane_N_")
  template <typename O>
  void assign_from(const MatrixBase<O> &o) {
    constexpr bool same = D::Rows == O::Rows && D::Cols == O::Cols;
    constexpr bool tvec = D::Rows == O::Cols && D::Cols == O::Rows && (D::Rows == 1 || D::Cols == 1);
    static_assert(same || tvec, "assignment: size mismatch");
    for (int j = 0; j < D::Cols; ++j)
      for (int i = 0; i < D::Rows; ++i) derived().coeffRef(i, j) = same ? o.derived().coeff(i, j) : o.derived().coeff(j, i);
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
  Matrix(double x, double y, double z) {
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
  double coeff(int i, int j) const { return O == RowMajor ? d_[i * C + j] : d_[j * R + i]; }
  double &coeffRef(int i, int j) { return O == RowMajor ? d_[i * C + j] : d_[j * R + i]; }
  static Matrix Identity() {
    Matrix m;
    m.setIdentity();
    return m;
  }
  static Matrix Zero() { return Matrix(); }

 private:
  S d_[R * C];
};

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 3, 3> Matrix3d;

// writable view of caller memory
template <typename S, int R, int C, int O>
class Map<Matrix<S, R, C, O>> : public MatrixBase<Map<Matrix<S, R, C, O>>> {
 public:
  typedef S Scalar;
  static constexpr int Rows = R, Cols = C, Options = O;
  explicit Map(S *p) : p_(p) {}
  Map &operator=(const Map &o) {
    this->assign_from(o);
    return *this;
  }
  template <typename X>
  Map &operator=(const MatrixBase<X> &o) {
    this->assign_from(o);
    return *this;
  }
  double coeff(int i, int j) const { return O == RowMajor ? p_[i * C + j] : p_[j * R + i]; }
  double &coeffRef(int i, int j) { return O == RowMajor ? p_[i * C + j] : p_[j * R + i]; }

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
  double coeff(int i, int j) const { return O == RowMajor ? p_[i * C + j] : p_[j * R + i]; }

 private:
  const S *p_;
};

template <typename X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC>> {
 public:
  typedef typename X::Scalar Scalar;
  static constexpr int Rows = BR, Cols = BC;
  Block(X &x, int r, int c) : x_(x), r_(r), c_(c) {}
  Block &operator=(const Block &o) {
    this->assign_from(o);
    return *this;
  }
  template <typename O>
  Block &operator=(const MatrixBase<O> &o) {
    this->assign_from(o);
    return *this;
  }
  double coeff(int i, int j) const { return x_.coeff(r_ + i, c_ + j); }
  double &coeffRef(int i, int j) { return x_.coeffRef(r_ + i, c_ + j); }

 private:
  X &x_;
  int r_, c_;
};

// ---- arithmetic: evaluated eagerly, element order = Eigen's for these fixed sizes ----
template <typename A, typename B>
auto operator+(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  static_assert(A::Rows == B::Rows && A::Cols == B::Cols, "+: size mismatch");
  Matrix<double, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = a.derived().coeff(i, j) + b.derived().coeff(i, j);
  return m;
}
template <typename A, typename B>
auto operator-(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  static_assert(A::Rows == B::Rows && A::Cols == B::Cols, "-: size mismatch");
  Matrix<double, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = a.derived().coeff(i, j) - b.derived().coeff(i, j);
  return m;
}
template <typename A>
auto operator-(const MatrixBase<A> &a) {
  Matrix<double, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = -a.derived().coeff(i, j);
  return m;
}
template <typename A>
auto operator*(const MatrixBase<A> &a, double s) {
  Matrix<double, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = a.derived().coeff(i, j) * s;
  return m;
}
template <typename A>
auto operator*(double s, const MatrixBase<A> &a) {
  Matrix<double, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = s * a.derived().coeff(i, j);
  return m;
}
template <typename A>
auto operator/(const MatrixBase<A> &a, double s) {
  Matrix<double, A::Rows, A::Cols> m;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) m.coeffRef(i, j) = a.derived().coeff(i, j) / s;
  return m;
}
template <typename A, typename B>
auto operator*(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  static_assert(A::Cols == B::Rows, "*: inner size mismatch");
  Matrix<double, A::Rows, B::Cols> m;
  for (int j = 0; j < B::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) {
      double s = a.derived().coeff(i, 0) * b.derived().coeff(0, j);
      for (int k = 1; k < A::Cols; ++k) s += a.derived().coeff(i, k) * b.derived().coeff(k, j);
      m.coeffRef(i, j) = s;
    }
  return m;
}
template <typename D>
template <typename O>
Matrix<double, 3, 1> MatrixBase<D>::cross(const MatrixBase<O> &o) const {
  static_assert(D::Rows * D::Cols == 3 && O::Rows * O::Cols == 3, "cross: 3-vectors only");
  const MatrixBase<D> &a = *this;
  return Matrix<double, 3, 1>(a(1) * o(2) - a(2) * o(1), a(2) * o(0) - a(0) * o(2), a(0) * o(1) - a(1) * o(0));
}

// ---- quaternions: coefficients stored x, y, z, w (Eigen's layout; Map<Quaterniond>(x + 3) relies on it) ----
template <typename D>
class QuaternionBase {
 public:
  const D &derived() const { return *static_cast<const D *>(this); }
  D &derived() { return *static_cast<D *>(this); }
  double x() const { return derived().data()[0]; }
  double y() const { return derived().data()[1]; }
  double z() const { return derived().data()[2]; }
  double w() const { return derived().data()[3]; }
  Matrix<double, 3, 1> vec() const { return Matrix<double, 3, 1>(x(), y(), z()); }
  double squaredNorm() const { return x() * x() + y() * y() + z() * z() + w() * w(); }
  double norm() const { return std::sqrt(squaredNorm()); }
  Quaternion<double> normalized() const;
  template <typename O>
  Quaternion<double> operator*(const QuaternionBase<O> &b) const;
  // rotate a 3-vector: Eigen's QuaternionBase::_transformVector
  template <typename V>
  Matrix<double, 3, 1> operator*(const MatrixBase<V> &v) const {
    const Matrix<double, 3, 1> q = vec();
    const Matrix<double, 3, 1> uv = 2.0 * q.cross(v);
    return v + w() * uv + q.cross(uv);
  }
  Matrix<double, 3, 3> toRotationMatrix() const {
    Matrix<double, 3, 3> r;
    const double tx = 2.0 * x(), ty = 2.0 * y(), tz = 2.0 * z();
    const double twx = tx * w(), twy = ty * w(), twz = tz * w();
    const double txx = tx * x(), txy = ty * x(), txz = tz * x();
    const double tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    r.coeffRef(0, 0) = 1.0 - (tyy + tzz);
    r.coeffRef(0, 1) = txy - twz;
    r.coeffRef(0, 2) = txz + twy;
    r.coeffRef(1, 0) = txy + twz;
    r.coeffRef(1, 1) = 1.0 - (txx + tzz);
    r.coeffRef(1, 2) = tyz - twx;
    r.coeffRef(2, 0) = txz - twy;
    r.coeffRef(2, 1) = tyz + twx;
    r.coeffRef(2, 2) = 1.0 - (txx + tyy);
    return r;
  }
  template <typename T>
  Quaternion<T> cast() const;
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
  Quaternion &operator=(const Quaternion &o) = default;
  const S *data() const { return d_; }
  S *data() { return d_; }

 private:
  S d_[4];
};
typedef Quaternion<double> Quaterniond;

template <>
class Map<Quaternion<double>> : public QuaternionBase<Map<Quaternion<double>>> {
 public:
  typedef double Scalar;
  explicit Map(double *p) : p_(p) {}
  template <typename O>
  Map &operator=(const QuaternionBase<O> &o) {
    const double x = o.x(), y = o.y(), z = o.z(), w = o.w();
    p_[0] = x, p_[1] = y, p_[2] = z, p_[3] = w;
    return *this;
  }
  const double *data() const { return p_; }
  double *data() { return p_; }

 private:
  double *p_;
};
template <>
class Map<const Quaternion<double>> : public QuaternionBase<Map<const Quaternion<double>>> {
 public:
  typedef double Scalar;
  explicit Map(const double *p) : p_(p) {}
  const double *data() const { return p_; }

 private:
  const double *p_;
};

template <typename D>
Quaternion<double> QuaternionBase<D>::normalized() const {  // coeffs / norm
  const double n = norm();
  return Quaternion<double>(w() / n, x() / n, y() / n, z() / n);
}
template <typename D>
template <typename O>
Quaternion<double> QuaternionBase<D>::operator*(const QuaternionBase<O> &b) const {  // Eigen's quat_product
  const QuaternionBase<D> &a = *this;
  return Quaternion<double>(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                            a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                            a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                            a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
}
template <typename D>
template <typename T>
Quaternion<T> QuaternionBase<D>::cast() const {
  return Quaternion<T>((T)w(), (T)x(), (T)y(), (T)z());
}

}  // namespace Eigen
#endif
