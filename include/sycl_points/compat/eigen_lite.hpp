// Minimal fixed-size linear algebra with Eigen's names and memory layout, used ONLY when the real
// Eigen is not installed (this image has none: SURVEY.md §8(c)).  With Eigen present the facade
// includes <Eigen/Dense> / <Eigen/Geometry> instead and this file is empty.
//
// Layout contract (what the C-ABI relies on): column-major, densely packed, Matrix4f / Vector4f
// 16-byte aligned — identical to Eigen::Matrix<float, R, C> — so PointType / Covariance /
// TransformMatrix arrays can be handed to libspx as float[n][4] / float[n][16] unchanged.
#pragma once

#if __has_include(<Eigen/Dense>) && !defined(SPX_FORCE_EIGEN_LITE)
#include <Eigen/Dense>
#include <Eigen/Geometry>
#else

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <initializer_list>
#include <limits>
#include <memory>
#include <new>
#include <ostream>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define SPX_EIGEN_LITE 1

namespace Eigen {

template <typename T, int R, int C>
struct Matrix;

/// writable view of a BR x BC block of a matrix (what Eigen's non-const block<>() returns): assignable,
/// accumulable, and convertible to the block's value
template <typename T, int R, int C, int BR, int BC>
struct BlockRef {
    Matrix<T, R, C>& m;
    int i0, j0;
    Matrix<T, BR, BC> eval() const {
        Matrix<T, BR, BC> r;
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) r(i, j) = m(i0 + i, j0 + j);
        return r;
    }
    operator Matrix<T, BR, BC>() const { return eval(); }
    BlockRef& operator=(const Matrix<T, BR, BC>& b) {
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) m(i0 + i, j0 + j) = b(i, j);
        return *this;
    }
    BlockRef& operator=(const BlockRef& o) { return *this = o.eval(); }
    BlockRef& operator+=(const Matrix<T, BR, BC>& b) { return *this = eval() + b; }
    BlockRef& operator-=(const Matrix<T, BR, BC>& b) { return *this = eval() - b; }
    BlockRef& operator*=(T s) { return *this = eval() * s; }
    BlockRef& operator/=(T s) { return *this = eval() / s; }
    T& operator()(int i, int j) { return m(i0 + i, j0 + j); }
    T operator()(int i, int j) const { return m(i0 + i, j0 + j); }
    T& operator()(int i) { return BC == 1 ? m(i0 + i, j0) : m(i0, j0 + i); }
    T& x() { return (*this)(0); }
    T& y() { return (*this)(1); }
    T& z() { return (*this)(2); }
    T norm() const { return eval().norm(); }
    T squaredNorm() const { return eval().squaredNorm(); }
    Matrix<T, BC, BR> transpose() const { return eval().transpose(); }
    void setZero() { *this = Matrix<T, BR, BC>::Zero(); }
    void setIdentity() { *this = Matrix<T, BR, BC>::Identity(); }
    Matrix<T, BR, BC> operator+(const Matrix<T, BR, BC>& o) const { return eval() + o; }
    Matrix<T, BR, BC> operator-(const Matrix<T, BR, BC>& o) const { return eval() - o; }
    Matrix<T, BR, BC> operator-() const { return -eval(); }
    Matrix<T, BR, BC> operator*(T s) const { return eval() * s; }
    template <int K>
    Matrix<T, BR, K> operator*(const Matrix<T, BC, K>& o) const { return eval() * o; }
    T dot(const Matrix<T, BR, BC>& o) const { return eval().dot(o); }
    Matrix<T, BR, BC> normalized() const { return eval().normalized(); }
};

/// `m << a, b, c, ...;` — coefficients in row-major order, as Eigen's comma initialiser
template <typename T, int R, int C>
struct CommaInit {
    Matrix<T, R, C>& m;
    int n;
    CommaInit& operator,(T v) {
        if (n < R * C) m(n / C, n % C) = v;
        ++n;
        return *this;
    }
};
/// writable view of the main diagonal
template <typename T, int R, int C>
struct DiagRef {
    Matrix<T, R, C>& m;
    static constexpr int N = R < C ? R : C;
    DiagRef& operator=(const Matrix<T, N, 1>& v) {
        for (int i = 0; i < N; ++i) m(i, i) = v(i);
        return *this;
    }
    operator Matrix<T, N, 1>() const {
        Matrix<T, N, 1> r;
        for (int i = 0; i < N; ++i) r(i) = m(i, i);
        return r;
    }
};

template <typename T, int R, int C>
struct Matrix {
    static constexpr int Rows = R, Cols = C, Size = R * C;
    alignas((sizeof(T) * R * C) % 16 == 0 ? 16 : alignof(T)) T m[R * C];

    Matrix() = default;
    Matrix(std::initializer_list<T> v) {
        int i = 0;
        for (T x : v)
            if (i < Size) m[i++] = x;
        for (; i < Size; ++i) m[i] = T(0);
    }
    template <int S = Size, typename = std::enable_if_t<S == 3>>
    Matrix(T x, T y, T z) : m{x, y, z} {}
    template <int S = Size, typename = std::enable_if_t<S == 4>>
    Matrix(T x, T y, T z, T w) : m{x, y, z, w} {}

    static Matrix Zero() {
        Matrix r;
        for (int i = 0; i < Size; ++i) r.m[i] = T(0);
        return r;
    }
    static Matrix Ones() {
        Matrix r;
        for (int i = 0; i < Size; ++i) r.m[i] = T(1);
        return r;
    }
    static Matrix Constant(T v) {
        Matrix r;
        for (int i = 0; i < Size; ++i) r.m[i] = v;
        return r;
    }
    static Matrix Identity() {
        Matrix r = Zero();
        for (int i = 0; i < (R < C ? R : C); ++i) r(i, i) = T(1);
        return r;
    }
    Matrix& setZero() { return *this = Zero(); }
    Matrix& setIdentity() { return *this = Identity(); }
    Matrix& setConstant(T v) { return *this = Constant(v); }

    T& operator()(int i, int j) { return m[j * R + i]; }
    const T& operator()(int i, int j) const { return m[j * R + i]; }
    T& operator()(int i) { return m[i]; }
    const T& operator()(int i) const { return m[i]; }
    T& operator[](int i) { return m[i]; }
    const T& operator[](int i) const { return m[i]; }
    T* data() { return m; }
    const T* data() const { return m; }
    static constexpr int rows() { return R; }
    static constexpr int cols() { return C; }
    static constexpr int size() { return Size; }
    T& x() { return m[0]; }
    T& y() { return m[1]; }
    T& z() { return m[2]; }
    T& w() { return m[3]; }
    const T& x() const { return m[0]; }
    const T& y() const { return m[1]; }
    const T& z() const { return m[2]; }
    const T& w() const { return m[3]; }

    Matrix operator+(const Matrix& o) const {
        Matrix r;
        for (int i = 0; i < Size; ++i) r.m[i] = m[i] + o.m[i];
        return r;
    }
    Matrix operator-(const Matrix& o) const {
        Matrix r;
        for (int i = 0; i < Size; ++i) r.m[i] = m[i] - o.m[i];
        return r;
    }
    Matrix operator-() const {
        Matrix r;
        for (int i = 0; i < Size; ++i) r.m[i] = -m[i];
        return r;
    }
    Matrix operator*(T s) const {
        Matrix r;
        for (int i = 0; i < Size; ++i) r.m[i] = m[i] * s;
        return r;
    }
    Matrix operator/(T s) const {
        Matrix r;
        for (int i = 0; i < Size; ++i) r.m[i] = m[i] / s;
        return r;
    }
    Matrix& operator+=(const Matrix& o) { return *this = *this + o; }
    Matrix& operator-=(const Matrix& o) { return *this = *this - o; }
    Matrix& operator*=(T s) { return *this = *this * s; }
    Matrix& operator/=(T s) { return *this = *this / s; }
    template <int K>
    Matrix<T, R, K> operator*(const Matrix<T, C, K>& o) const {
        Matrix<T, R, K> r = Matrix<T, R, K>::Zero();
        for (int j = 0; j < K; ++j)
            for (int k = 0; k < C; ++k)
                for (int i = 0; i < R; ++i) r(i, j) += (*this)(i, k) * o(k, j);
        return r;
    }
    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> r;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) r(j, i) = (*this)(i, j);
        return r;
    }
    T dot(const Matrix& o) const {
        T s = T(0);
        for (int i = 0; i < Size; ++i) s += m[i] * o.m[i];
        return s;
    }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    T sum() const {
        T s = T(0);
        for (int i = 0; i < Size; ++i) s += m[i];
        return s;
    }
    T trace() const {
        T s = T(0);
        for (int i = 0; i < (R < C ? R : C); ++i) s += (*this)(i, i);
        return s;
    }
    Matrix normalized() const { return *this / norm(); }
    bool allFinite() const {
        for (int i = 0; i < Size; ++i)
            if (!std::isfinite(m[i])) return false;
        return true;
    }
    bool isApprox(const Matrix& o, T prec = T(1e-5)) const {
        const T a = (*this - o).squaredNorm(), b = squaredNorm(), c = o.squaredNorm();
        return a <= prec * prec * (b < c ? b : c);
    }
    /// copy of the BR x BC block at (i0, j0) (read-only, unlike Eigen's view)
    template <int BR, int BC>
    Matrix<T, BR, BC> block(int i0, int j0) const {
        Matrix<T, BR, BC> r;
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) r(i, j) = (*this)(i0 + i, j0 + j);
        return r;
    }
    template <int BR, int BC>
    BlockRef<T, R, C, BR, BC> block(int i0, int j0) {
        return BlockRef<T, R, C, BR, BC>{*this, i0, j0};
    }
    template <int BR, int BC>
    Matrix<T, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
    template <int BR, int BC>
    BlockRef<T, R, C, BR, BC> topLeftCorner() { return block<BR, BC>(0, 0); }
    CommaInit<T, R, C> operator<<(T v) {
        (*this)(0, 0) = v;
        return CommaInit<T, R, C>{*this, 1};
    }
    DiagRef<T, R, C> diagonal() { return DiagRef<T, R, C>{*this}; }
    Matrix<T, (R < C ? R : C), 1> diagonal() const {
        Matrix<T, (R < C ? R : C), 1> r;
        for (int i = 0; i < (R < C ? R : C); ++i) r(i) = (*this)(i, i);
        return r;
    }
    Matrix eval() const { return *this; }
    template <int N>
    BlockRef<T, R, C, N, 1> head() {
        static_assert(C == 1, "head<N>() of a column vector");
        return BlockRef<T, R, C, N, 1>{*this, 0, 0};
    }
    Matrix<T, 1, C> row(int i) const { return block<1, C>(i, 0); }
    Matrix<T, R, 1> col(int j) const { return block<R, 1>(0, j); }
    BlockRef<T, R, C, 1, C> row(int i) { return BlockRef<T, R, C, 1, C>{*this, i, 0}; }
    BlockRef<T, R, C, R, 1> col(int j) { return BlockRef<T, R, C, R, 1>{*this, 0, j}; }
    static Matrix Unit(int k) {
        Matrix r = Zero();
        r.m[k] = T(1);
        return r;
    }
    static Matrix UnitX() { return Unit(0); }
    static Matrix UnitY() { return Unit(1); }
    static Matrix UnitZ() { return Unit(2); }
    Matrix cross(const Matrix& o) const {
        static_assert(Size == 3, "cross product of 3-vectors");
        return Matrix(m[1] * o.m[2] - m[2] * o.m[1], m[2] * o.m[0] - m[0] * o.m[2], m[0] * o.m[1] - m[1] * o.m[0]);
    }
    template <int BR, int BC>
    void set_block(int i0, int j0, const Matrix<T, BR, BC>& b) {
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) (*this)(i0 + i, j0 + j) = b(i, j);
    }
    template <int N>
    Matrix<T, N, 1> head() const {
        Matrix<T, N, 1> r;
        for (int i = 0; i < N; ++i) r.m[i] = m[i];
        return r;
    }
    template <int N>
    Matrix<T, N, 1> tail() const {
        Matrix<T, N, 1> r;
        for (int i = 0; i < N; ++i) r.m[i] = m[Size - N + i];
        return r;
    }
    /// Eigen's isometries expose .matrix(); a plain matrix is its own matrix
    Matrix& matrix() { return *this; }
    const Matrix& matrix() const { return *this; }
};

template <typename T, int R, int C>
inline Matrix<T, R, C> operator*(T s, const Matrix<T, R, C>& a) {
    return a * s;
}
template <typename T, int R, int C>
inline std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& a) {
    for (int i = 0; i < R; ++i) {
        for (int j = 0; j < C; ++j) os << (j ? " " : "") << a(i, j);
        if (i + 1 < R) os << "\n";
    }
    return os;
}

template <typename T, int N>
using Vector = Matrix<T, N, 1>;
using Vector2f = Matrix<float, 2, 1>;
using Vector3f = Matrix<float, 3, 1>;
using Vector4f = Matrix<float, 4, 1>;
using Matrix3f = Matrix<float, 3, 3>;
using Matrix4f = Matrix<float, 4, 4>;
using Vector3d = Matrix<double, 3, 1>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;

/// Eigen::Isometry3f: a 4x4 homogeneous rigid transform
struct Isometry3f {
    Matrix4f M = Matrix4f::Identity();
    Isometry3f() = default;
    explicit Isometry3f(const Matrix4f& mat) : M(mat) {}
    static Isometry3f Identity() { return Isometry3f(); }
    Matrix4f& matrix() { return M; }
    const Matrix4f& matrix() const { return M; }
    Matrix3f linear() const { return M.block<3, 3>(0, 0); }
    BlockRef<float, 4, 4, 3, 3> linear() { return M.block<3, 3>(0, 0); }
    Matrix3f rotation() const { return linear(); }
    Vector3f translation() const { return Vector3f(M(0, 3), M(1, 3), M(2, 3)); }
    BlockRef<float, 4, 4, 3, 1> translation() { return M.block<3, 1>(0, 3); }
    void set_translation(const Vector3f& t) {
        for (int i = 0; i < 3; ++i) M(i, 3) = t(i);
    }
    void set_linear(const Matrix3f& R) { M.set_block<3, 3>(0, 0, R); }
    Isometry3f operator*(const Isometry3f& o) const { return Isometry3f(M * o.M); }
    Matrix4f operator*(const Matrix4f& o) const { return M * o; }
    Vector3f operator*(const Vector3f& p) const { return linear() * p + translation(); }
    Isometry3f inverse() const {
        const Matrix3f Rt = linear().transpose();
        Isometry3f r;
        r.set_linear(Rt);
        r.set_translation(-(Rt * translation()));
        return r;
    }
    const float* data() const { return M.data(); }
    float* data() { return M.data(); }
};

/// Eigen::Transform<float, 3, Eigen::Isometry>: the only transform the path uses
template <typename T, int Dim, int Mode, int Options = 0>
using Transform = Isometry3f;
constexpr int Isometry = 1;

/// Eigen::AngleAxisf: rotation by `angle` about the unit vector `axis` (Rodrigues)
struct AngleAxisf {
    float a;
    Vector3f ax;
    AngleAxisf() : a(0.0f), ax(1.0f, 0.0f, 0.0f) {}
    AngleAxisf(float angle, const Vector3f& axis) : a(angle), ax(axis) {}
    /// from a rotation matrix: angle in [0, pi]
    explicit AngleAxisf(const Matrix3f& R) {
        const float c = std::fmin(1.0f, std::fmax(-1.0f, 0.5f * (R(0, 0) + R(1, 1) + R(2, 2) - 1.0f)));
        Vector3f v(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
        const float s = 0.5f * v.norm();
        a = std::atan2(s, c);
        if (s > 1e-6f) {
            ax = v / (2.0f * s);
        } else if (c > 0.0f) {
            a = 0.0f;
            ax = Vector3f(1.0f, 0.0f, 0.0f);
        } else {  // angle ~ pi: the axis is the dominant column of (R + I) / 2
            int k = 0;
            if (R(1, 1) > R(k, k)) k = 1;
            if (R(2, 2) > R(k, k)) k = 2;
            Vector3f col(R(0, k) + (k == 0 ? 1.0f : 0.0f), R(1, k) + (k == 1 ? 1.0f : 0.0f), R(2, k) + (k == 2 ? 1.0f : 0.0f));
            ax = col.normalized();
        }
    }
    static AngleAxisf Identity() { return AngleAxisf(); }
    float angle() const { return a; }
    const Vector3f& axis() const { return ax; }
    Matrix3f toRotationMatrix() const {
        const float c = std::cos(a), s = std::sin(a), t = 1.0f - c;
        const float x = ax.x(), y = ax.y(), z = ax.z();
        Matrix3f R;
        R(0, 0) = t * x * x + c;     R(0, 1) = t * x * y - s * z; R(0, 2) = t * x * z + s * y;
        R(1, 0) = t * x * y + s * z; R(1, 1) = t * y * y + c;     R(1, 2) = t * y * z - s * x;
        R(2, 0) = t * x * z - s * y; R(2, 1) = t * y * z + s * x; R(2, 2) = t * z * z + c;
        return R;
    }
    Matrix3f matrix() const { return toRotationMatrix(); }
};

/// Eigen::Translation3f: only as the left factor of an isometry product
struct Translation3f {
    Vector3f t;
    Translation3f(float x, float y, float z) : t(x, y, z) {}
    explicit Translation3f(const Vector3f& v) : t(v) {}
    Isometry3f operator*(const Isometry3f& o) const {
        Isometry3f r = o;
        r.set_translation(o.translation() + t);
        return r;
    }
};

template <typename T>
struct aligned_allocator {
    using value_type = T;
    aligned_allocator() = default;
    template <typename U>
    aligned_allocator(const aligned_allocator<U>&) {}
    T* allocate(std::size_t n) {
        constexpr std::size_t al = alignof(T) < 16 ? 16 : alignof(T);
        void* p = ::operator new(n * sizeof(T), std::align_val_t(al));
        return static_cast<T*>(p);
    }
    void deallocate(T* p, std::size_t) {
        constexpr std::size_t al = alignof(T) < 16 ? 16 : alignof(T);
        ::operator delete(p, std::align_val_t(al));
    }
    template <typename U>
    bool operator==(const aligned_allocator<U>&) const { return true; }
    template <typename U>
    bool operator!=(const aligned_allocator<U>&) const { return false; }
};

}  // namespace Eigen
#endif
