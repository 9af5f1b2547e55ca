// Small fixed-size math used inside the kernels and by the host-side optimiser pieces.
// Rounding contract: where the reference spells an fma chain (I/utils/eigen_utils.hpp), the same
// chain is spelled here with __fmaf_rn/__fmul_rn (device) so nvcc neither fuses nor splits
// anything; symmetric 3x3 matrices are carried as their 6 upper-triangle entries.
#pragma once

#include <cmath>

#include "spx_common.cuh"

#ifdef __CUDA_ARCH__
#define SPX_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define SPX_MUL(a, b) __fmul_rn((a), (b))
#define SPX_ADD(a, b) __fadd_rn((a), (b))
#define SPX_SUB(a, b) __fsub_rn((a), (b))
#define SPX_DIV(a, b) __fdiv_rn((a), (b))
#else
#define SPX_FMA(a, b, c) fmaf((a), (b), (c))
#define SPX_MUL(a, b) ((a) * (b))
#define SPX_ADD(a, b) ((a) + (b))
#define SPX_SUB(a, b) ((a) - (b))
#define SPX_DIV(a, b) ((a) / (b))
#endif

#define SPX_HD __host__ __device__ __forceinline__

namespace spx {

// Transcendentals.  The reference calls sycl::cos / acos / cbrt / sin / log / pow, whose rounding is
// implementation-defined (unpinned).  The contract here and in the oracle is the
// CORRECTLY ROUNDED fp32 value, obtained by evaluating in fp64 and rounding once: CUDA's and glibc's
// fp64 routines are both < 2 ulp(fp64), so the two sides round to the same float except when the
// exact value lies within ~1e-16 relative of a rounding boundary (probability ~1e-8 per call).
// None of these sit on a per-iteration hot loop (plane regularisation and normals run once per
// point per align; se3_exp once per iteration on one thread).
SPX_HD float cr_cosf(float x) { return (float)cos((double)x); }
SPX_HD float cr_sinf(float x) { return (float)sin((double)x); }
SPX_HD float cr_acosf(float x) { return (float)acos((double)x); }
SPX_HD float cr_cbrtf(float x) { return (float)cbrt((double)x); }
SPX_HD float cr_logf(float x) { return (float)log((double)x); }
SPX_HD float cr_cubef(float x) {  // pow(x, 3.0f)
    const double d = (double)x;
    return (float)(d * d * d);
}

// symmetric 3x3: xx xy xz yy yz zz
struct Sym3 {
    float xx, xy, xz, yy, yz, zz;
};

struct Mat3 {
    float m[3][3];
};

// a*b - c*d the way eigen_utils.hpp writes its 2x2 minors: fma(a, b, -(c*d))
SPX_HD float minor2(float a, float b, float c, float d) { return SPX_FMA(a, b, -SPX_MUL(c, d)); }

// determinant(A) — eigen_utils.hpp:303-307, A symmetric
SPX_HD float sym_det(const Sym3& a) {
    return SPX_FMA(a.xx, minor2(a.yy, a.zz, a.yz, a.yz),
                   SPX_FMA(-a.xy, minor2(a.xy, a.zz, a.yz, a.xz), SPX_MUL(a.xz, minor2(a.xy, a.yz, a.yy, a.xz))));
}

// inverse(A) — eigen_utils.hpp:403-423: adjugate / det, the ZERO matrix when |det| < 1e-6
SPX_HD Sym3 sym_inverse(const Sym3& a) {
    const float det = sym_det(a);
    Sym3 r;
    if (fabsf(det) < 1e-6f) {
        r.xx = r.xy = r.xz = r.yy = r.yz = r.zz = 0.0f;
        return r;
    }
    const float id = SPX_DIV(1.0f, det);
    r.xx = SPX_MUL(minor2(a.yy, a.zz, a.yz, a.yz), id);  // (0,0)
    r.xy = SPX_MUL(minor2(a.xz, a.yz, a.xy, a.zz), id);  // (0,1) = fma(s02, s21, -s01*s22)
    r.xz = SPX_MUL(minor2(a.xy, a.yz, a.xz, a.yy), id);  // (0,2) = fma(s01, s12, -s02*s11)
    r.yy = SPX_MUL(minor2(a.xx, a.zz, a.xz, a.xz), id);  // (1,1)
    r.yz = SPX_MUL(minor2(a.xz, a.xy, a.xx, a.yz), id);  // (1,2) = fma(s02, s10, -s00*s12)
    r.zz = SPX_MUL(minor2(a.xx, a.yy, a.xy, a.xy), id);  // (2,2)
    return r;
}

#ifdef __CUDACC__
// symmetric_eigen_decomposition_3x3 — eigen_utils.hpp:443-562 (scaled trigonometric Cardano,
// eigenvalues ascending, eigenvector k = largest-norm column of adj(A - l_k I)).  Only the
// eigenvectors V (columns) are returned scaled-matrix-exact; eigenvalues are rescaled at the end.
__device__ inline void sym_eigen3(const Sym3& A, float ev[3], float V[3][3]) {
    constexpr float EPS = 1.1920929e-07f;
    constexpr float FMIN = 1.17549435e-38f;
    constexpr float PI = 3.14159265358979323846f;
    float mx = fmaxf(fmaxf(fmaxf(fabsf(A.xx), fabsf(A.xy)), fmaxf(fabsf(A.xz), fabsf(A.yy))),
                     fmaxf(fabsf(A.yz), fabsf(A.zz)));
    if (mx < FMIN) {
        ev[0] = ev[1] = ev[2] = 0.0f;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0f : 0.0f;
        return;
    }
    const float si = __fdiv_rn(1.0f, mx);
    Sym3 S;
    S.xx = __fmul_rn(A.xx, si); S.xy = __fmul_rn(A.xy, si); S.xz = __fmul_rn(A.xz, si);
    S.yy = __fmul_rn(A.yy, si); S.yz = __fmul_rn(A.yz, si); S.zz = __fmul_rn(A.zz, si);

    const float c2 = -__fadd_rn(__fadd_rn(__fadd_rn(0.0f, S.xx), S.yy), S.zz);
    const float c1 = __fsub_rn(__fmaf_rn(S.xx, S.yy, __fmaf_rn(S.xx, S.zz, __fmul_rn(S.yy, S.zz))),
                               __fmaf_rn(S.xy, S.xy, __fmaf_rn(S.xz, S.xz, __fmul_rn(S.yz, S.yz))));
    const float c0 = -sym_det(S);

    const float p = __fsub_rn(c1, __fdiv_rn(__fmul_rn(c2, c2), 3.0f));
    const float q = __fadd_rn(__fsub_rn(__fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(2.0f, c2), c2), c2), 27.0f),
                                        __fdiv_rn(__fmul_rn(c2, c1), 3.0f)),
                              c0);
    const float disc = __fadd_rn(__fmul_rn(__fmul_rn(__fmul_rn(4.0f, p), p), p), __fmul_rn(__fmul_rn(27.0f, q), q));
    const float c2_3 = __fdiv_rn(c2, 3.0f);
    if (fabsf(disc) <= EPS) {
        const float u = q >= 0 ? -cr_cbrtf(__fdiv_rn(q, 2.0f)) : cr_cbrtf(__fdiv_rn(-q, 2.0f));
        ev[0] = __fsub_rn(__fmul_rn(2.0f, u), c2_3);
        ev[1] = ev[2] = __fsub_rn(-u, c2_3);
    } else {
        const float sp = __fsqrt_rn(__fdiv_rn(-p, 3.0f));
        const float den = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, sp), sp), sp);
        const float cs = fmaxf(-1.0f, fminf(1.0f, __fdiv_rn(-q, den)));
        const float phi = fabsf(p) < EPS ? 0.0f : cr_acosf(cs);
        const float two_sp = __fmul_rn(2.0f, sp);
        ev[0] = __fmaf_rn(two_sp, cr_cosf(__fdiv_rn(phi, 3.0f)), -c2_3);
        ev[2] = __fmaf_rn(two_sp, cr_cosf(__fdiv_rn(__fadd_rn(phi, __fmul_rn(4.0f, PI)), 3.0f)), -c2_3);
        ev[1] = __fmaf_rn(two_sp, cr_cosf(__fdiv_rn(__fadd_rn(phi, __fmul_rn(2.0f, PI)), 3.0f)), -c2_3);
    }
    float t;
    if (ev[0] > ev[1]) { t = ev[0]; ev[0] = ev[1]; ev[1] = t; }
    if (ev[1] > ev[2]) { t = ev[1]; ev[1] = ev[2]; ev[2] = t; }
    if (ev[0] > ev[1]) { t = ev[0]; ev[0] = ev[1]; ev[1] = t; }

#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Sym3 M = S;
        M.xx = __fsub_rn(S.xx, ev[k]);
        M.yy = __fsub_rn(S.yy, ev[k]);
        M.zz = __fsub_rn(S.zz, ev[k]);
        // cofactors (symmetric): m00 m01 m02 m11 m12 m22
        const float m00 = minor2(M.yy, M.zz, M.yz, M.yz);
        const float m01 = minor2(M.yz, M.xz, M.xy, M.zz);
        const float m02 = minor2(M.xy, M.yz, M.yy, M.xz);
        const float m11 = minor2(M.xx, M.zz, M.xz, M.xz);
        const float m12 = minor2(M.xy, M.xz, M.xx, M.yz);
        const float m22 = minor2(M.xx, M.yy, M.xy, M.xy);
        const float s0 = __fmaf_rn(m00, m00, __fmaf_rn(m01, m01, __fmul_rn(m02, m02)));
        const float s1 = __fmaf_rn(m01, m01, __fmaf_rn(m11, m11, __fmul_rn(m12, m12)));
        const float s2 = __fmaf_rn(m02, m02, __fmaf_rn(m12, m12, __fmul_rn(m22, m22)));
        float vx, vy, vz;
        if (s0 >= s1 && s0 >= s2) {
            vx = m00; vy = m01; vz = m02;
        } else if (s1 >= s0 && s1 >= s2) {
            vx = m01; vy = m11; vz = m12;
        } else {
            vx = m02; vy = m12; vz = m22;
        }
        float n2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
        if (n2 < FMIN) {
            vx = 1.0f; vy = 0.0f; vz = 0.0f;
            n2 = 1.0f;
        }
        const float il = __fdiv_rn(1.0f, __fsqrt_rn(n2));
        V[0][k] = __fmul_rn(vx, il);
        V[1][k] = __fmul_rn(vy, il);
        V[2][k] = __fmul_rn(vz, il);
    }
    ev[0] = __fmul_rn(ev[0], mx);
    ev[1] = __fmul_rn(ev[1], mx);
    ev[2] = __fmul_rn(ev[2], mx);
}

// update_covariance_plane — I/algorithms/feature/covariance.hpp:67-74: C <- V diag(1e-3,1,1) V^T
// (eigenvalues discarded).  Upper triangle of the reference's fma chains.
__device__ inline Sym3 plane_regularize(const Sym3& C) {
    float ev[3], V[3][3];
    sym_eigen3(C, ev, V);
    float VD[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        VD[i][0] = __fmul_rn(V[i][0], 1e-3f);
        VD[i][1] = V[i][1];
        VD[i][2] = V[i][2];
    }
    auto e = [&](int i, int j) {
        return __fmaf_rn(VD[i][2], V[j][2], __fmaf_rn(VD[i][1], V[j][1], __fmul_rn(VD[i][0], V[j][0])));
    };
    Sym3 r;
    r.xx = e(0, 0); r.xy = e(0, 1); r.xz = e(0, 2);
    r.yy = e(1, 1); r.yz = e(1, 2); r.zz = e(2, 2);
    return r;
}

// covariance stored by the reference: float[16] column-major 4x4; upper 3x3 is symmetric
__device__ __forceinline__ Sym3 load_cov16(const float* __restrict__ c) {
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(c));
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(c) + 1);
    const float4 c2 = __ldg(reinterpret_cast<const float4*>(c) + 2);
    Sym3 s;
    s.xx = c0.x; s.xy = c1.x; s.xz = c2.x;
    s.yy = c1.y; s.yz = c2.y; s.zz = c2.z;
    return s;
}
#endif  // __CUDACC__

// ---------------------------------------------------------------- SE(3) pieces (host + device)
// lie::se3_exp — eigen_utils.hpp:886-943 (so3_exp -> quaternion -> rotation; V matrix).  Row-major out.
SPX_HD void se3_exp_rm(const float a[6], float T[4][4]) {
    const float ox = a[0], oy = a[1], oz = a[2];
    const float th2 = SPX_FMA(oz, oz, SPX_FMA(oy, oy, SPX_MUL(ox, ox)));
    float imag, real;
    if (th2 < 1e-6f) {
        const float th4 = SPX_MUL(th2, th2);
        imag = SPX_ADD(SPX_SUB(0.5f, SPX_MUL(SPX_DIV(1.0f, 48.0f), th2)), SPX_MUL(SPX_DIV(1.0f, 3840.0f), th4));
        real = SPX_ADD(SPX_SUB(1.0f, SPX_MUL(SPX_DIV(1.0f, 8.0f), th2)), SPX_MUL(SPX_DIV(1.0f, 384.0f), th4));
    } else {
        const float th = sqrtf(th2);
        const float h = SPX_MUL(0.5f, th);
        imag = SPX_DIV(cr_sinf(h), th);
        real = cr_cosf(h);
    }
    const float x = SPX_MUL(imag, ox), y = SPX_MUL(imag, oy), z = SPX_MUL(imag, oz), w = real;
    const float x2 = SPX_MUL(x, x), y2 = SPX_MUL(y, y), z2 = SPX_MUL(z, z);
    const float xy = SPX_MUL(x, y), xz = SPX_MUL(x, z), yz = SPX_MUL(y, z);
    const float wx = SPX_MUL(w, x), wy = SPX_MUL(w, y), wz = SPX_MUL(w, z);
    float R[3][3];
    R[0][0] = SPX_SUB(1.0f, SPX_MUL(2.0f, SPX_ADD(y2, z2)));
    R[0][1] = SPX_MUL(2.0f, SPX_SUB(xy, wz));
    R[0][2] = SPX_MUL(2.0f, SPX_ADD(xz, wy));
    R[1][0] = SPX_MUL(2.0f, SPX_ADD(xy, wz));
    R[1][1] = SPX_SUB(1.0f, SPX_MUL(2.0f, SPX_ADD(x2, z2)));
    R[1][2] = SPX_MUL(2.0f, SPX_SUB(yz, wx));
    R[2][0] = SPX_MUL(2.0f, SPX_SUB(xz, wy));
    R[2][1] = SPX_MUL(2.0f, SPX_ADD(yz, wx));
    R[2][2] = SPX_SUB(1.0f, SPX_MUL(2.0f, SPX_ADD(x2, y2)));
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T[i][j] = (i == j) ? 1.0f : 0.0f;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i][j] = R[i][j];
    const float tv[3] = {a[3], a[4], a[5]};
    const float th = sqrtf(th2);
    float Vm[3][3];
    if (th < 1e-6f) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Vm[i][j] = R[i][j];
    } else {
        const float Om[3][3] = {{0.0f, -oz, oy}, {oz, 0.0f, -ox}, {-oy, ox, 0.0f}};
        float Om2[3][3];
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) {
                float s = 0.0f;
                for (int k = 0; k < 3; ++k) s = SPX_FMA(Om[i][k], Om[k][j], s);
                Om2[i][j] = s;
            }
        const float A = SPX_DIV(SPX_SUB(1.0f, cr_cosf(th)), th2);
        const float B = SPX_DIV(SPX_SUB(th, cr_sinf(th)), SPX_MUL(th2, th));
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                Vm[i][j] = SPX_ADD((i == j) ? 1.0f : 0.0f, SPX_ADD(SPX_MUL(Om[i][j], A), SPX_MUL(Om2[i][j], B)));
    }
    for (int i = 0; i < 3; ++i) {
        float s = 0.0f;
        for (int j = 0; j < 3; ++j) s = SPX_FMA(Vm[i][j], tv[j], s);
        T[i][3] = s;
    }
}

// result.T * Isometry3f(se3_exp(delta)) — registration.hpp:814 (Eigen isometry product:
// linear = L1*L2, translation = L1*t2 + t1).  Row-major 4x4.
SPX_HD void isometry_mul_rm(const float A[4][4], const float B[4][4], float out[4][4]) {
    float r[4][4];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            float s = 0.0f;
            for (int k = 0; k < 3; ++k) s = SPX_ADD(s, SPX_MUL(A[i][k], B[k][j]));
            r[i][j] = s;
        }
        float s = 0.0f;
        for (int k = 0; k < 3; ++k) s = SPX_ADD(s, SPX_MUL(A[i][k], B[k][3]));
        r[i][3] = SPX_ADD(s, A[i][3]);
    }
    r[3][0] = r[3][1] = r[3][2] = 0.0f;
    r[3][3] = 1.0f;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i][j] = r[i][j];
}

// ---------------------------------------------------------------- 6x6 LDL^T (fp64, diagonal pivoting)
// Stands in for Eigen::LDLT<Matrix<float,6,6>> at registration.hpp:791-801 / dogleg_step.hpp:43-50
// (Eigen is an unpinned third-party dependency of the reference; the factorisation is evaluated
// in fp64 and the solution cast to fp32, which agrees with any backward-stable fp32 LDLT to far
// below the 1e-5 pose tolerance).
struct Ldlt6 {
    double L[6][6];
    double D[6];
    int perm[6];
    bool ok;
};

SPX_HD void ldlt6_compute(const double Ain[6][6], Ldlt6& f) {
    double A[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) A[i][j] = Ain[i][j];
    for (int i = 0; i < 6; ++i) f.perm[i] = i;
    f.ok = true;
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        double best = fabs(A[k][k]);
        for (int i = k + 1; i < 6; ++i)
            if (fabs(A[i][i]) > best) {
                best = fabs(A[i][i]);
                piv = i;
            }
        if (piv != k) {
            for (int j = 0; j < 6; ++j) {
                const double t = A[k][j];
                A[k][j] = A[piv][j];
                A[piv][j] = t;
            }
            for (int j = 0; j < 6; ++j) {
                const double t = A[j][k];
                A[j][k] = A[j][piv];
                A[j][piv] = t;
            }
            const int t = f.perm[k];
            f.perm[k] = f.perm[piv];
            f.perm[piv] = t;
        }
        const double d = A[k][k];
        if (d == 0.0) {
            for (int i = k + 1; i < 6; ++i) {
                if (A[i][k] != 0.0) f.ok = false;
                A[i][k] = 0.0;
            }
            continue;
        }
        for (int i = k + 1; i < 6; ++i) A[i][k] /= d;
        for (int i = k + 1; i < 6; ++i)
            for (int j = k + 1; j <= i; ++j) {
                A[i][j] -= A[i][k] * d * A[j][k];
                A[j][i] = A[i][j];
            }
    }
    for (int i = 0; i < 6; ++i) {
        f.D[i] = A[i][i];
        for (int j = 0; j < 6; ++j) f.L[i][j] = (j < i) ? A[i][j] : (i == j ? 1.0 : 0.0);
    }
}

SPX_HD void ldlt6_solve(const Ldlt6& f, const double rhs[6], double x[6]) {
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = rhs[f.perm[i]];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < i; ++j) y[i] -= f.L[i][j] * y[j];
    for (int i = 0; i < 6; ++i) y[i] = (fabs(f.D[i]) > 2.2250738585072014e-308) ? y[i] / f.D[i] : 0.0;
    for (int i = 5; i >= 0; --i)
        for (int j = i + 1; j < 6; ++j) y[i] -= f.L[j][i] * y[j];
    for (int i = 0; i < 6; ++i) x[f.perm[i]] = y[i];
}

// solve (H + lambda I) delta = -b; H row-major 36 floats.  registration.hpp:791-801,806-808.
SPX_HD bool solve_damped6(const float* H, const float* b, float lambda, float* delta) {
    double A[6][6], rhs[6], x[6];
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) A[i][j] = (double)H[i * 6 + j];
        A[i][i] = (double)SPX_ADD(H[i * 6 + i], lambda);
        rhs[i] = -(double)b[i];
    }
    Ldlt6 f;
    ldlt6_compute(A, f);
    if (!f.ok) {
        for (int i = 0; i < 6; ++i) delta[i] = 0.0f;
        return false;
    }
    ldlt6_solve(f, rhs, x);
    for (int i = 0; i < 6; ++i) delta[i] = (float)x[i];
    return true;
}

#ifdef __CUDACC__
// Device fast path of solve_damped6 for the in-kernel Gauss-Newton step: H + lambda I is symmetric
// positive definite in every healthy iteration, so the factorisation needs no pivoting and, fully
// unrolled, lives in registers (the pivoted version indexes its arrays dynamically, i.e. local
// memory: ~9 us on one thread, which every block of the cooperative kernel waited for).  Any
// non-positive or tiny pivot falls back to the pivoted routine, so degenerate systems behave as
// before.  Same fp64 arithmetic, different elimination order: solutions agree to ~1e-15 relative.
__device__ __forceinline__ bool solve_damped6_device(const float* H, const float* b, float lambda, float* delta) {
    double A[6][6];
    double maxd = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) A[i][j] = (double)H[i * 6 + j];
        A[i][i] = (double)SPX_ADD(H[i * 6 + i], lambda);
        maxd = fmax(maxd, A[i][i]);
    }
    bool ok = maxd > 0.0;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = -(double)b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double d = A[k][k];
        ok = ok && d > 1e-10 * maxd && isfinite(d);
        const double inv = 1.0 / d;
        double l[6];
#pragma unroll
        for (int i = k + 1; i < 6; ++i) l[i] = A[i][k] * inv;
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
#pragma unroll
            for (int j = k + 1; j <= i; ++j) A[i][j] -= l[i] * A[j][k];
        }
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            A[i][k] = l[i];
            y[i] -= l[i] * y[k];  // forward substitution rides along
        }
    }
    if (!ok) return solve_damped6(H, b, lambda, delta);
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] /= A[i][i];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
#pragma unroll
        for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) delta[i] = (float)y[i];
    return true;
}
#endif

SPX_HD float norm3f(const float* v) {
    return sqrtf(SPX_ADD(SPX_ADD(SPX_MUL(v[0], v[0]), SPX_MUL(v[1], v[1])), SPX_MUL(v[2], v[2])));
}

}  // namespace spx
