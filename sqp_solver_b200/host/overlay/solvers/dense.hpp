// Dense column-major containers for the host-side mirror of the reference interface.
//
// The reference's interface types are Eigen matrices (include/solvers/qp.hpp:21-27). Eigen is
// an external dependency that is not vendored; when <Eigen/Dense> is on the include path these
// aliases ARE the Eigen types, so user code written against the reference compiles unchanged.
// Otherwise a minimal stand-in with the same storage order (column-major, contiguous) and the
// handful of members the solvers use (rows, cols, data, operator(), resize, setZero, ...) is used.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <limits>
#include <vector>

#if defined(__has_include)
#if __has_include(<Eigen/Dense>) && !defined(SQPB200_NO_EIGEN)
#define SQPB200_HAVE_EIGEN 1
#endif
#endif

#ifdef SQPB200_HAVE_EIGEN
#include <Eigen/Dense>
namespace sqpb200_dense {
template <typename S> using Vector = Eigen::Matrix<S, Eigen::Dynamic, 1>;
template <typename S> using Matrix = Eigen::Matrix<S, Eigen::Dynamic, Eigen::Dynamic>;
using VectorXi = Eigen::VectorXi;
}  // namespace sqpb200_dense
#else
namespace sqpb200_dense {

template <typename S>
class Vector {
   public:
    using Scalar = S;
    Vector() = default;
    explicit Vector(std::ptrdiff_t n) : v_(static_cast<size_t>(n)) {}
    Vector(std::initializer_list<S> il) : v_(il) {}
    static Vector Zero(std::ptrdiff_t n) { Vector r(n); r.setZero(); return r; }
    static Vector Constant(std::ptrdiff_t n, S c) { Vector r(n); r.setConstant(c); return r; }
    std::ptrdiff_t rows() const { return static_cast<std::ptrdiff_t>(v_.size()); }
    std::ptrdiff_t cols() const { return 1; }
    std::ptrdiff_t size() const { return rows(); }
    void resize(std::ptrdiff_t n) { v_.resize(static_cast<size_t>(n)); }
    S *data() { return v_.data(); }
    const S *data() const { return v_.data(); }
    S &operator()(std::ptrdiff_t i) { assert(i >= 0 && i < rows()); return v_[static_cast<size_t>(i)]; }
    const S &operator()(std::ptrdiff_t i) const { assert(i >= 0 && i < rows()); return v_[static_cast<size_t>(i)]; }
    S &operator[](std::ptrdiff_t i) { return (*this)(i); }
    const S &operator[](std::ptrdiff_t i) const { return (*this)(i); }
    Vector &setZero() { for (auto &e : v_) e = S(0); return *this; }
    Vector &setZero(std::ptrdiff_t n) { resize(n); return setZero(); }
    Vector &setConstant(S c) { for (auto &e : v_) e = c; return *this; }
    S sum() const { S s = 0; for (auto e : v_) s += e; return s; }
    S dot(const Vector &o) const { assert(o.rows() == rows()); S s = 0; for (size_t i = 0; i < v_.size(); ++i) s += v_[i] * o.v_[i]; return s; }
    S squaredNorm() const { return dot(*this); }
    S norm() const { return std::sqrt(squaredNorm()); }
    // Eigen's isApprox: ||a-b||^2 <= prec^2 * min(||a||^2, ||b||^2)
    bool isApprox(const Vector &o, S prec = std::numeric_limits<S>::epsilon() * 100) const {
        S d = 0;
        for (size_t i = 0; i < v_.size(); ++i) d += (v_[i] - o.v_[i]) * (v_[i] - o.v_[i]);
        S a = squaredNorm(), b = o.squaredNorm();
        return d <= prec * prec * (a < b ? a : b);
    }

   private:
    std::vector<S> v_;
};

template <typename S>
class Matrix {
   public:
    using Scalar = S;
    Matrix() = default;
    Matrix(std::ptrdiff_t r, std::ptrdiff_t c) : r_(r), c_(c), v_(static_cast<size_t>(r * c)) {}
    static Matrix Identity(std::ptrdiff_t r, std::ptrdiff_t c) { Matrix m(r, c); m.setIdentity(); return m; }
    static Matrix Zero(std::ptrdiff_t r, std::ptrdiff_t c) { Matrix m(r, c); m.setZero(); return m; }
    std::ptrdiff_t rows() const { return r_; }
    std::ptrdiff_t cols() const { return c_; }
    void resize(std::ptrdiff_t r, std::ptrdiff_t c) { r_ = r; c_ = c; v_.resize(static_cast<size_t>(r * c)); }
    S *data() { return v_.data(); }
    const S *data() const { return v_.data(); }
    S &operator()(std::ptrdiff_t i, std::ptrdiff_t j) { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return v_[static_cast<size_t>(i + r_ * j)]; }
    const S &operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return v_[static_cast<size_t>(i + r_ * j)]; }
    Matrix &setZero() { for (auto &e : v_) e = S(0); return *this; }
    Matrix &setIdentity() { setZero(); for (std::ptrdiff_t i = 0; i < (r_ < c_ ? r_ : c_); ++i) (*this)(i, i) = S(1); return *this; }

   private:
    std::ptrdiff_t r_ = 0, c_ = 0;
    std::vector<S> v_;
};

using VectorXi = Vector<int>;

}  // namespace sqpb200_dense
#endif
