// Minimal stand-ins for the Eigen types that appear in EdgeFEM's public API for the
// frequency-domain hot path (Eigen 3.4 is not available in this build environment).
// Same spellings as the reference uses after `using` (VecC, SpMatC, Eigen::MatrixXcd ->
// MatrixXcd, Eigen::Vector3d -> Vector3d), and the subset of members its callers touch.
// SpMatC stores true row-major CSR (the reference's Eigen matrices are column-major; for the
// complex-symmetric systems of this path the arrays coincide).
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace edgefem {

using cplx = std::complex<double>;

struct Vector3d {
  double v[3]{0.0, 0.0, 0.0};
  Vector3d() = default;
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  static Vector3d Zero() { return Vector3d(); }
  static Vector3d Ones() { return Vector3d(1.0, 1.0, 1.0); }
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double &x() { return v[0]; }
  double &y() { return v[1]; }
  double &z() { return v[2]; }
  Vector3d operator+(const Vector3d &o) const { return {v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]}; }
  Vector3d operator-(const Vector3d &o) const { return {v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}; }
  Vector3d operator-() const { return {-v[0], -v[1], -v[2]}; }
  Vector3d operator*(double s) const { return {v[0] * s, v[1] * s, v[2] * s}; }
  Vector3d operator/(double s) const { return {v[0] / s, v[1] / s, v[2] / s}; }
  double dot(const Vector3d &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
  Vector3d cross(const Vector3d &o) const {
    return {v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]};
  }
  double norm() const { return std::sqrt(dot(*this)); }
  Vector3d normalized() const {
    double n = norm();
    return n > 0 ? (*this) / n : *this;
  }
};
inline Vector3d operator*(double s, const Vector3d &a) { return a * s; }

struct Vector2d {
  double v[2]{0.0, 0.0};
  Vector2d() = default;
  Vector2d(double x, double y) : v{x, y} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
};

template <typename T>
class DenseVector {
public:
  DenseVector() = default;
  explicit DenseVector(std::size_t n) : d_(n) {}
  DenseVector(std::size_t n, const T &val) : d_(n, val) {}
  static DenseVector Zero(std::size_t n) { return DenseVector(n, T(0)); }
  std::size_t size() const { return d_.size(); }
  void resize(std::size_t n) { d_.assign(n, T(0)); }
  void setZero() { std::fill(d_.begin(), d_.end(), T(0)); }
  T &operator()(std::size_t i) { return d_[i]; }
  const T &operator()(std::size_t i) const { return d_[i]; }
  T &operator[](std::size_t i) { return d_[i]; }
  const T &operator[](std::size_t i) const { return d_[i]; }
  T *data() { return d_.data(); }
  const T *data() const { return d_.data(); }
  double norm() const {
    double s = 0.0;
    for (const auto &x : d_) s += std::norm(x);
    return std::sqrt(s);
  }
  DenseVector &operator*=(double a) {
    for (auto &x : d_) x *= a;
    return *this;
  }
  std::vector<T> &vec() { return d_; }
  const std::vector<T> &vec() const { return d_; }

private:
  std::vector<T> d_;
};
using VectorXcd = DenseVector<cplx>;
using VectorXd = DenseVector<double>;

template <typename T>
class DenseMatrix {  // column-major like Eigen's default
public:
  DenseMatrix() = default;
  DenseMatrix(int r, int c) : r_(r), c_(c), d_((std::size_t)r * c, T(0)) {}
  static DenseMatrix Zero(int r, int c) { return DenseMatrix(r, c); }
  int rows() const { return r_; }
  int cols() const { return c_; }
  void resize(int r, int c) {
    r_ = r;
    c_ = c;
    d_.assign((std::size_t)r * c, T(0));
  }
  T &operator()(int i, int j) { return d_[(std::size_t)j * r_ + i]; }
  const T &operator()(int i, int j) const { return d_[(std::size_t)j * r_ + i]; }
  T *data() { return d_.data(); }
  const T *data() const { return d_.data(); }

private:
  int r_ = 0, c_ = 0;
  std::vector<T> d_;
};
using MatrixXcd = DenseMatrix<cplx>;
using MatrixXd = DenseMatrix<double>;

// Compressed sparse ROW matrix with sorted column indices; explicit zeros are kept.
template <typename T>
class SparseMatrix {
public:
  SparseMatrix() : rowptr_(1, 0) {}
  SparseMatrix(int r, int c) : r_(r), c_(c), rowptr_((std::size_t)r + 1, 0) {}
  void resize(int r, int c) {
    r_ = r;
    c_ = c;
    rowptr_.assign((std::size_t)r + 1, 0);
    col_.clear();
    val_.clear();
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  std::int64_t nonZeros() const { return (std::int64_t)col_.size(); }
  T coeff(int i, int j) const {
    if (i < 0 || i >= r_ || j < 0 || j >= c_) return T(0);
    int lo = rowptr_[i], hi = rowptr_[i + 1];
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (col_[mid] == j) return val_[mid];
      if (col_[mid] < j) lo = mid + 1; else hi = mid;
    }
    return T(0);
  }
  // raw CSR access
  std::vector<int> &rowptr() { return rowptr_; }
  std::vector<int> &colidx() { return col_; }
  std::vector<T> &values() { return val_; }
  const std::vector<int> &rowptr() const { return rowptr_; }
  const std::vector<int> &colidx() const { return col_; }
  const std::vector<T> &values() const { return val_; }
  // y = A x
  template <typename V>
  DenseVector<V> operator*(const DenseVector<V> &x) const {
    DenseVector<V> y(r_, V(0));
    for (int i = 0; i < r_; ++i) {
      V acc(0);
      for (int k = rowptr_[i]; k < rowptr_[i + 1]; ++k) acc += val_[k] * x[col_[k]];
      y[i] = acc;
    }
    return y;
  }

private:
  int r_ = 0, c_ = 0;
  std::vector<int> rowptr_, col_;
  std::vector<T> val_;
};

} // namespace edgefem
