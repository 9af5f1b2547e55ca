// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the small-matrix math the reference's hot path runs inside its
// SYCL kernels.  Nothing under oracle/ may be imported, linked or executed by the
// product library (sycl_points_b200/); only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, and only as the checker.
//
// Every function cites the reference lines it restates.  Abbreviation:
//   I/ = /root/reference/cpp/include/sycl_points/
//
// Arithmetic contract: fp32, std::fma exactly where the reference writes sycl::fma,
// plain * and + elsewhere; the translation unit is compiled with -ffp-contract=off so
// the compiler adds no contractions of its own.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>

namespace orc {

// Transcendentals: the reference's sycl::cos / acos / cbrt / sin / log / pow are implementation-
// defined; the oracle fixes them as the CORRECTLY ROUNDED fp32 value (fp64
// evaluation, one rounding) — the same contract as the CUDA side (spx_math.cuh cr_*).
inline float cr_cos(float x) { return (float)std::cos((double)x); }
inline float cr_sin(float x) { return (float)std::sin((double)x); }
inline float cr_acos(float x) { return (float)std::acos((double)x); }
inline float cr_atan2(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float cr_cbrt(float x) { return (float)std::cbrt((double)x); }
inline float cr_log(float x) { return (float)std::log((double)x); }
inline float cr_exp(float x) { return (float)std::exp((double)x); }
inline float cr_cube(float x) { const double d = (double)x; return (float)(d * d * d); }

template <int M, int N>
struct Mat {
    float v[M][N];
    float& operator()(int i, int j) { return v[i][j]; }
    float operator()(int i, int j) const { return v[i][j]; }
    static Mat zero() {
        Mat r;
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) r.v[i][j] = 0.0f;
        return r;
    }
    static Mat identity() {
        Mat r = zero();
        for (int i = 0; i < (M < N ? M : N); ++i) r.v[i][i] = 1.0f;
        return r;
    }
};

template <int N>
struct Vec {
    float v[N];
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    static Vec zero() {
        Vec r;
        for (int i = 0; i < N; ++i) r.v[i] = 0.0f;
        return r;
    }
};

using M3 = Mat<3, 3>;
using M4 = Mat<4, 4>;
using M6 = Mat<6, 6>;
using V3 = Vec<3>;
using V4 = Vec<4>;
using V6 = Vec<6>;

// I/utils/eigen_utils.hpp:88-105  matrix product, fma accumulation over k ascending from 0
template <int M, int K, int N>
inline Mat<M, N> mul(const Mat<M, K>& A, const Mat<K, N>& B) {
    Mat<M, N> r = Mat<M, N>::zero();
    for (int j = 0; j < N; ++j)
        for (int k = 0; k < K; ++k) {
            const float b = B(k, j);
            for (int i = 0; i < M; ++i) r(i, j) = std::fma(A(i, k), b, r(i, j));
        }
    return r;
}

// I/utils/eigen_utils.hpp:113-127  matrix * vector, fma accumulation over j ascending from 0
template <int M, int N>
inline Vec<M> mul(const Mat<M, N>& A, const Vec<N>& x) {
    Vec<M> r;
    for (int i = 0; i < M; ++i) {
        float s = 0.0f;
        for (int j = 0; j < N; ++j) s = std::fma(A(i, j), x(j), s);
        r(i) = s;
    }
    return r;
}

// I/utils/eigen_utils.hpp:135-161  scalar products (plain multiply)
template <int M, int N>
inline Mat<M, N> scale(const Mat<M, N>& A, float s) {
    Mat<M, N> r;
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < M; ++i) r(i, j) = A(i, j) * s;
    return r;
}
template <int N>
inline Vec<N> scale(const Vec<N>& a, float s) {
    Vec<N> r;
    for (int i = 0; i < N; ++i) r(i) = a(i) * s;
    return r;
}

template <int M, int N>
inline Mat<M, N> add(const Mat<M, N>& A, const Mat<M, N>& B) {  // :32-43
    Mat<M, N> r;
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < M; ++i) r(i, j) = A(i, j) + B(i, j);
    return r;
}
template <int M, int N>
inline Mat<M, N> sub(const Mat<M, N>& A, const Mat<M, N>& B) {  // :67-79
    Mat<M, N> r;
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < M; ++i) r(i, j) = A(i, j) - B(i, j);
    return r;
}

// I/utils/eigen_utils.hpp:208-219
template <int M>
inline Mat<M, M> ensure_symmetric(const Mat<M, M>& A) {
    Mat<M, M> r;
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i) r(i, j) = (i == j) ? A(i, j) : (A(i, j) + A(j, i)) * 0.5f;
    return r;
}

template <int M, int N>
inline Mat<N, M> transpose(const Mat<M, N>& A) {  // :226-237
    Mat<N, M> r;
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < M; ++i) r(j, i) = A(i, j);
    return r;
}

// I/utils/eigen_utils.hpp:245-253
template <int N>
inline float dot(const Vec<N>& u, const Vec<N>& w) {
    float r = 0.0f;
    for (int i = 0; i < N; ++i) r = std::fma(u(i), w(i), r);
    return r;
}

// I/utils/eigen_utils.hpp:272-284
template <int N>
inline Mat<N, N> outer(const Vec<N>& u, const Vec<N>& w) {
    Mat<N, N> r;
    for (int j = 0; j < N; ++j) {
        const float wj = w(j);
        for (int i = 0; i < N; ++i) r(i, j) = u(i) * wj;
    }
    return r;
}

inline float trace3(const M3& A) {  // :291-298
    float r = 0.0f;
    for (int i = 0; i < 3; ++i) r += A(i, i);
    return r;
}

// I/utils/eigen_utils.hpp:303-307
inline float determinant(const M3& A) {
    return std::fma(A(0, 0), std::fma(A(1, 1), A(2, 2), -A(1, 2) * A(2, 1)),
                    std::fma(-A(0, 1), std::fma(A(1, 0), A(2, 2), -A(1, 2) * A(2, 0)),
                             A(0, 2) * std::fma(A(1, 0), A(2, 1), -A(1, 1) * A(2, 0))));
}

// I/utils/eigen_utils.hpp:403-423  adjugate inverse, zero matrix when |det| < 1e-6
inline M3 inverse(const M3& s) {
    const float det = determinant(s);
    if (std::fabs(det) < 1e-6f) return M3::zero();
    const float id = 1.0f / det;
    M3 r;
    r(0, 0) = std::fma(s(1, 1), s(2, 2), -s(1, 2) * s(2, 1)) * id;
    r(1, 0) = std::fma(s(1, 2), s(2, 0), -s(1, 0) * s(2, 2)) * id;
    r(2, 0) = std::fma(s(1, 0), s(2, 1), -s(1, 1) * s(2, 0)) * id;
    r(0, 1) = std::fma(s(0, 2), s(2, 1), -s(0, 1) * s(2, 2)) * id;
    r(1, 1) = std::fma(s(0, 0), s(2, 2), -s(0, 2) * s(2, 0)) * id;
    r(2, 1) = std::fma(s(0, 1), s(2, 0), -s(0, 0) * s(2, 1)) * id;
    r(0, 2) = std::fma(s(0, 1), s(1, 2), -s(0, 2) * s(1, 1)) * id;
    r(1, 2) = std::fma(s(0, 2), s(1, 0), -s(0, 0) * s(1, 2)) * id;
    r(2, 2) = std::fma(s(0, 0), s(1, 1), -s(0, 1) * s(1, 0)) * id;
    return r;
}

// I/utils/eigen_utils.hpp:443-562  scaled trigonometric-Cardano eigen-decomposition,
// eigenvalues ascending, eigenvector k = largest-norm column of adj(A - lambda_k I).
inline void eigen3(const M3& A, V3& evals, M3& evecs) {
    constexpr float EPS = std::numeric_limits<float>::epsilon();
    constexpr float PI = 3.14159265358979323846f;
    float mx = 0.0f;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) mx = std::fmax(mx, std::fabs(A(i, j)));
    if (mx < std::numeric_limits<float>::min()) {
        evals = V3::zero();
        evecs = M3::identity();
        return;
    }
    const float sinv = 1.0f / mx;
    const M3 S = scale(A, sinv);

    const float c2 = -trace3(S);
    const float c1 = std::fma(S(0, 0), S(1, 1), std::fma(S(0, 0), S(2, 2), S(1, 1) * S(2, 2))) -
                     std::fma(S(0, 1), S(1, 0), std::fma(S(0, 2), S(2, 0), S(1, 2) * S(2, 1)));
    const float c0 = -determinant(S);

    const float p = c1 - c2 * c2 / 3.0f;
    const float q = 2.0f * c2 * c2 * c2 / 27.0f - c2 * c1 / 3.0f + c0;
    const float disc = 4.0f * p * p * p + 27.0f * q * q;

    if (std::fabs(disc) <= EPS) {
        const float u = q >= 0 ? -cr_cbrt(q / 2.0f) : cr_cbrt(-q / 2.0f);
        evals(0) = 2.0f * u - c2 / 3.0f;
        evals(1) = evals(2) = -u - c2 / 3.0f;
    } else {
        const float sp = std::sqrt(-p / 3.0f);
        const float cs = std::max(-1.0f, std::min(1.0f, -q / (2.0f * sp * sp * sp)));
        float phi = std::fabs(p) < EPS ? 0.0f : cr_acos(cs);
        if (phi < 0.0f) phi += PI;
        evals(0) = std::fma(2.0f * sp, cr_cos(phi / 3.0f), -c2 / 3.0f);
        evals(2) = std::fma(2.0f * sp, cr_cos((phi + 4.0f * PI) / 3.0f), -c2 / 3.0f);
        evals(1) = std::fma(2.0f * sp, cr_cos((phi + 2.0f * PI) / 3.0f), -c2 / 3.0f);
    }
    if (evals(0) > evals(1)) std::swap(evals(0), evals(1));
    if (evals(1) > evals(2)) std::swap(evals(1), evals(2));
    if (evals(0) > evals(1)) std::swap(evals(1), evals(0));

    evecs = M3::zero();
    for (int k = 0; k < 3; ++k) {
        M3 Mk = S;  // S - lambda_k * I  (identity * lambda then subtract: off-diagonals subtract 0)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Mk(i, j) = S(i, j) - ((i == j ? 1.0f : 0.0f) * evals(k));

        const float m00 = std::fma(Mk(1, 1), Mk(2, 2), -Mk(1, 2) * Mk(2, 1));
        const float m01 = std::fma(Mk(1, 2), Mk(2, 0), -Mk(1, 0) * Mk(2, 2));
        const float m02 = std::fma(Mk(1, 0), Mk(2, 1), -Mk(1, 1) * Mk(2, 0));
        const float m10 = std::fma(Mk(0, 2), Mk(2, 1), -Mk(0, 1) * Mk(2, 2));
        const float m11 = std::fma(Mk(0, 0), Mk(2, 2), -Mk(0, 2) * Mk(2, 0));
        const float m12 = std::fma(Mk(0, 1), Mk(2, 0), -Mk(0, 0) * Mk(2, 1));
        const float m20 = std::fma(Mk(0, 1), Mk(1, 2), -Mk(0, 2) * Mk(1, 1));
        const float m21 = std::fma(Mk(0, 2), Mk(1, 0), -Mk(0, 0) * Mk(1, 2));
        const float m22 = std::fma(Mk(0, 0), Mk(1, 1), -Mk(0, 1) * Mk(1, 0));

        const float s0 = std::fma(m00, m00, std::fma(m10, m10, m20 * m20));
        const float s1 = std::fma(m01, m01, std::fma(m11, m11, m21 * m21));
        const float s2 = std::fma(m02, m02, std::fma(m12, m12, m22 * m22));

        float vx, vy, vz;
        if (s0 >= s1 && s0 >= s2) {
            vx = m00; vy = m10; vz = m20;
        } else if (s1 >= s0 && s1 >= s2) {
            vx = m01; vy = m11; vz = m21;
        } else {
            vx = m02; vy = m12; vz = m22;
        }
        // sycl::dot(float3): contraction left to the SYCL implementation; restated as the
        // plain left-to-right sum of products.
        float n2 = vx * vx + vy * vy + vz * vz;
        if (n2 < std::numeric_limits<float>::min()) {
            vx = 1.0f; vy = 0.0f; vz = 0.0f;
            n2 = 1.0f;
        }
        const float il = 1.0f / std::sqrt(n2);
        evecs(0, k) = vx * il;
        evecs(1, k) = vy * il;
        evecs(2, k) = vz * il;
    }
    evals = scale(evals, mx);
}

// ---------------------------------------------------------------- Lie group helpers
// I/utils/eigen_utils.hpp:808-836
inline M3 quat_to_rot(const V4& qt) {
    const float x = qt(0), y = qt(1), z = qt(2), w = qt(3);
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    const float xy = x * y, xz = x * z, yz = y * z;
    const float wx = w * x, wy = w * y, wz = w * z;
    M3 R;
    R(0, 0) = 1.0f - 2.0f * (y2 + z2);
    R(0, 1) = 2.0f * (xy - wz);
    R(0, 2) = 2.0f * (xz + wy);
    R(1, 0) = 2.0f * (xy + wz);
    R(1, 1) = 1.0f - 2.0f * (x2 + z2);
    R(1, 2) = 2.0f * (yz - wx);
    R(2, 0) = 2.0f * (xz - wy);
    R(2, 1) = 2.0f * (yz + wx);
    R(2, 2) = 1.0f - 2.0f * (x2 + y2);
    return R;
}

// I/utils/eigen_utils.hpp:860-866
inline M3 skew(float x, float y, float z) {
    M3 r;
    r(0, 0) = 0.0f; r(0, 1) = -z;   r(0, 2) = y;
    r(1, 0) = z;    r(1, 1) = 0.0f; r(1, 2) = -x;
    r(2, 0) = -y;   r(2, 1) = x;    r(2, 2) = 0.0f;
    return r;
}

// I/utils/eigen_utils.hpp:886-902
inline V4 so3_exp(const V3& om) {
    const float th2 = dot<3>(om, om);
    float imag, real;
    if (th2 < 1e-6f) {
        const float th4 = th2 * th2;
        imag = 0.5f - 1.0f / 48.0f * th2 + 1.0f / 3840.0f * th4;
        real = 1.0f - 1.0f / 8.0f * th2 + 1.0f / 384.0f * th4;
    } else {
        const float th = std::sqrt(th2);
        const float h = 0.5f * th;
        imag = cr_sin(h) / th;
        real = cr_cos(h);
    }
    V4 q;
    q(0) = imag * om(0); q(1) = imag * om(1); q(2) = imag * om(2); q(3) = real;
    return q;
}

// I/utils/eigen_utils.hpp:909-943  twist = [rx ry rz tx ty tz], rotation first
inline M4 se3_exp(const V6& a) {
    V3 om; om(0) = a(0); om(1) = a(1); om(2) = a(2);
    V3 tv; tv(0) = a(3); tv(1) = a(4); tv(2) = a(5);
    const float th2 = dot<3>(om, om);
    const float th = std::sqrt(th2);
    const M3 R = quat_to_rot(so3_exp(om));
    M4 T = M4::identity();
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) T(r, c) = R(r, c);
    V3 t;
    if (th < 1e-6f) {
        t = mul<3, 3>(R, tv);
    } else {
        const M3 Om = skew(om(0), om(1), om(2));
        const M3 Om2 = mul<3, 3, 3>(Om, Om);
        const float A = (1.0f - cr_cos(th)) / th2;
        const float B = (th - cr_sin(th)) / (th2 * th);
        const M3 V = add(M3::identity(), add(scale(Om, A), scale(Om2, B)));
        t = mul<3, 3>(V, tv);
    }
    T(0, 3) = t(0); T(1, 3) = t(1); T(2, 3) = t(2);
    return T;
}

// I/utils/eigen_utils.hpp:774-803
inline V4 rot_to_quat(const M3& R) {
    V4 q;
    const float tr = R(0, 0) + R(1, 1) + R(2, 2);
    if (tr > 0.0f) {
        const float S = std::sqrt(tr + 1.0f) * 2.0f;
        q(0) = (R(2, 1) - R(1, 2)) / S;
        q(1) = (R(0, 2) - R(2, 0)) / S;
        q(2) = (R(1, 0) - R(0, 1)) / S;
        q(3) = 0.25f * S;
    } else if ((R(0, 0) > R(1, 1)) && (R(0, 0) > R(2, 2))) {
        const float S = std::sqrt(1.0f + R(0, 0) - R(1, 1) - R(2, 2)) * 2.0f;
        q(0) = 0.25f * S;
        q(1) = (R(0, 1) + R(1, 0)) / S;
        q(2) = (R(0, 2) + R(2, 0)) / S;
        q(3) = (R(2, 1) - R(1, 2)) / S;
    } else if (R(1, 1) > R(2, 2)) {
        const float S = std::sqrt(1.0f + R(1, 1) - R(0, 0) - R(2, 2)) * 2.0f;
        q(0) = (R(0, 1) + R(1, 0)) / S;
        q(1) = 0.25f * S;
        q(2) = (R(1, 2) + R(2, 1)) / S;
        q(3) = (R(0, 2) - R(2, 0)) / S;
    } else {
        const float S = std::sqrt(1.0f + R(2, 2) - R(0, 0) - R(1, 1)) * 2.0f;
        q(2) = 0.25f * S;
        q(3) = (R(1, 0) - R(0, 1)) / S;
        q(0) = (R(0, 2) + R(2, 0)) / S;
        q(1) = (R(1, 2) + R(2, 1)) / S;
    }
    return q;
}

// I/utils/eigen_utils.hpp:948-986
inline V3 so3_log(const V4& quat) {
    V4 q = quat;
    const float nrm = std::sqrt(dot<4>(q, q));
    if (nrm < 1e-6f) {
        q = V4::zero();
    } else {
        q = scale(q, 1.0f / nrm);
    }
    if (q(3) < 0.0f) {
        for (int i = 0; i < 4; ++i) q(i) *= -1.0f;
    }
    const float w = q(3);
    V3 xyz; xyz(0) = q(0); xyz(1) = q(1); xyz(2) = q(2);
    const float n = std::sqrt(dot<3>(xyz, xyz));
    if (n < 1e-6f) {
        const float s = 2.0f / w * (1.0f + n * n / (6.0f * w * w));
        return scale(xyz, s);
    }
    if (std::fabs(w) < 1e-6f) {
        const float th = 3.14159265358979323846f;
        return scale(xyz, th / n);
    }
    const float th = 2.0f * cr_atan2(n, std::fabs(w));
    return scale(xyz, th / n);
}

// I/utils/eigen_utils.hpp:991-1034 (Eigen expression arithmetic restated as plain fp32)
inline V6 se3_log(const M4& T) {
    M3 R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R(i, j) = T(i, j);
    const float t[3] = {T(0, 3), T(1, 3), T(2, 3)};
    const V3 om = so3_log(rot_to_quat(R));
    const float th = std::sqrt(dot<3>(om, om));
    V6 r;
    r(0) = om(0); r(1) = om(1); r(2) = om(2);
    const M3 Om = skew(om(0), om(1), om(2));
    M3 Vinv = M3::identity();
    if (th < 1e-6f) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Vinv(i, j) = Vinv(i, j) - 0.5f * Om(i, j);
    } else {
        const float h = 0.5f * th;
        const float sh = cr_sin(h), ch = cr_cos(h);
        const float coeff = (1.0f - th * ch / (2.0f * sh)) / (th * th);
        M3 Om2 = M3::zero();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                float s = 0.0f;
                for (int k = 0; k < 3; ++k) s += Om(i, k) * Om(k, j);
                Om2(i, j) = s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Vinv(i, j) = Vinv(i, j) - 0.5f * Om(i, j) + coeff * Om2(i, j);
    }
    for (int i = 0; i < 3; ++i) {
        float s = 0.0f;
        for (int k = 0; k < 3; ++k) s += Vinv(i, k) * t[k];
        r(3 + i) = s;
    }
    return r;
}

// Isometry3f product as Eigen evaluates it for Transform<float,3,Isometry>:
// linear = L1*L2, translation = L1*t2 + t1, last row fixed (registration.hpp:814 call site).
inline M4 isometry_mul(const M4& A, const M4& B) {
    M4 r = M4::identity();
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            float s = 0.0f;
            for (int k = 0; k < 3; ++k) s += A(i, k) * B(k, j);
            r(i, j) = s;
        }
        float s = 0.0f;
        for (int k = 0; k < 3; ++k) s += A(i, k) * B(k, 3);
        r(i, 3) = s + A(i, 3);
    }
    return r;
}

}  // namespace orc
