// Per-point covariance and normal estimation from k neighbours
// (I/algorithms/feature/covariance.hpp:16-74,260-311,417-495).
#include <algorithm>

#include "spx_math.cuh"

using namespace spx;

namespace {

constexpr int FEAT_THREADS = 128;

// kernel::estimate — covariance.hpp:16-47: plain (un-fused) sums of p and p p^T over the
// neighbour ranks in order, identity when fewer than 4 valid neighbours, otherwise
// sym(sum_outer / n - mean mean^T).  The un-centred fp32 formulation is kept on purpose: it is
// what the reference computes.
__device__ __forceinline__ bool estimate_cov(const float4* __restrict__ pts, const int32_t* __restrict__ row, int k,
                                             Sym3& C) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    float oxx = 0.f, oxy = 0.f, oxz = 0.f, oyy = 0.f, oyz = 0.f, ozz = 0.f;
    int cnt = 0;
    for (int j = 0; j < k; ++j) {
        const int id = __ldg(row + j);
        if (id < 0) continue;
        const float4 p = __ldg(pts + id);
        sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z);
        // outer(u, v)(i, j) = u(i) * v(j): (i,j) and (j,i) are the same product bit for bit
        oxx = __fadd_rn(oxx, __fmul_rn(p.x, p.x));
        oxy = __fadd_rn(oxy, __fmul_rn(p.x, p.y));
        oxz = __fadd_rn(oxz, __fmul_rn(p.x, p.z));
        oyy = __fadd_rn(oyy, __fmul_rn(p.y, p.y));
        oyz = __fadd_rn(oyz, __fmul_rn(p.y, p.z));
        ozz = __fadd_rn(ozz, __fmul_rn(p.z, p.z));
        ++cnt;
    }
    if (cnt < 4) {
        C.xx = C.yy = C.zz = 1.0f;
        C.xy = C.xz = C.yz = 0.0f;
        return false;
    }
    const float inv = __fdiv_rn(1.0f, (float)cnt);
    const float mx = __fmul_rn(sx, inv), my = __fmul_rn(sy, inv), mz = __fmul_rn(sz, inv);
    // ensure_symmetric of an exactly symmetric matrix: (a + a) * 0.5 == a
    C.xx = __fsub_rn(__fmul_rn(oxx, inv), __fmul_rn(mx, mx));
    C.xy = __fsub_rn(__fmul_rn(oxy, inv), __fmul_rn(mx, my));
    C.xz = __fsub_rn(__fmul_rn(oxz, inv), __fmul_rn(mx, mz));
    C.yy = __fsub_rn(__fmul_rn(oyy, inv), __fmul_rn(my, my));
    C.yz = __fsub_rn(__fmul_rn(oyz, inv), __fmul_rn(my, mz));
    C.zz = __fsub_rn(__fmul_rn(ozz, inv), __fmul_rn(mz, mz));
    return true;
}


// extract_normal — covariance.hpp:49-65: eigenvector of the smallest eigenvalue, kept when
// dot(n, p) <= 1.0 and negated otherwise (the reference's literal test), w = 0.
__device__ __forceinline__ float4 normal_of(const float4 p, const Mat3& C) {
    float ev[3], V[3][3];
    mat3_eigen(C, ev, V);
    const float nx = V[0][0], ny = V[1][0], nz = V[2][0];
    const float d = __fmaf_rn(nz, p.z, __fmaf_rn(ny, p.y, __fmul_rn(nx, p.x)));
    return d <= 1.0f ? make_float4(nx, ny, nz, 0.f) : make_float4(-nx, -ny, -nz, 0.f);
}

__global__ void __launch_bounds__(FEAT_THREADS) covariance_kernel(const float4* __restrict__ pts, uint32_t n,
                                                                  const int32_t* __restrict__ idx, int k,
                                                                  float* __restrict__ covs) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    Sym3 C;
    estimate_cov(pts, idx + (size_t)i * k, k, C);
    store_cov16(covs + (size_t)i * 16, mat3_from_sym(C));
}

// kernel::estimate_weighted — covariance.hpp:97-134: weighted sums over the valid neighbours in rank order, identity
// when fewer than 4 of them or no weight; returns success
__device__ __forceinline__ bool estimate_cov_weighted(const float4* __restrict__ pts, const int32_t* __restrict__ row, int k,
                                                      const float* w, Mat3& C, float mean[3]) {
    float sp[3] = {0.f, 0.f, 0.f};
    float so[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    int cnt = 0;
    float tw = 0.0f;
    for (int j = 0; j < k; ++j) {
        const int id = __ldg(row + j);
        if (id < 0) continue;
        const float4 p4 = __ldg(pts + id);
        const float p[3] = {p4.x, p4.y, p4.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            sp[a] = __fadd_rn(sp[a], __fmul_rn(p[a], w[j]));
#pragma unroll
            for (int b = 0; b < 3; ++b) so[a][b] = __fadd_rn(so[a][b], __fmul_rn(__fmul_rn(p[a], p[b]), w[j]));
        }
        ++cnt;
        tw = __fadd_rn(tw, w[j]);
    }
    if (cnt < 4 || tw < 1.1920929e-07f) {
        C = mat3_identity();
        return false;
    }
    const float inv = __fdiv_rn(1.0f, tw);
#pragma unroll
    for (int a = 0; a < 3; ++a) mean[a] = __fmul_rn(sp[a], inv);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) C.m[a][b] = __fsub_rn(__fmul_rn(so[a][b], inv), __fmul_rn(mean[a], mean[b]));
    return true;  // exactly symmetric: ensure_symmetric is the identity on it
}

// kernel::estimate_robust — covariance.hpp:182-222: M-estimated covariance: weights from the SQUARED Mahalanobis
// distances (passed to the robust weight as the residual, as the reference does) at scale mad_scale x median
// (floored at min_robust_scale; the median runs over all k entries, unfilled ones count as 0, :159-173)
constexpr int ROBUST_MAX_K = 64;
__global__ void __launch_bounds__(FEAT_THREADS) covariance_robust_kernel(const float4* __restrict__ pts, uint32_t n,
                                                                         const int32_t* __restrict__ idx, int k, int loss,
                                                                         float mad_scale, float min_scale, int max_iter,
                                                                         float* __restrict__ covs) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    const int32_t* row = idx + (size_t)i * k;
    float w[ROBUST_MAX_K], d2[ROBUST_MAX_K];
    for (int j = 0; j < ROBUST_MAX_K; ++j) {
        w[j] = 1.0f;
        d2[j] = 0.0f;
    }
    Mat3 C;
    float mean[3] = {0.f, 0.f, 0.f};
    bool ok = estimate_cov_weighted(pts, row, k, w, C, mean);
    for (int it = 0; ok && it < max_iter; ++it) {
        const Mat3 Ci = mat3_inverse(C);
        for (int j = 0; j < k; ++j) {
            const int id = __ldg(row + j);
            if (id < 0) continue;
            const float4 p = __ldg(pts + id);
            const float d0 = __fsub_rn(p.x, mean[0]), d1 = __fsub_rn(p.y, mean[1]), dz = __fsub_rn(p.z, mean[2]);
            const float m0 = __fmaf_rn(Ci.m[0][2], dz, __fmaf_rn(Ci.m[0][1], d1, __fmul_rn(Ci.m[0][0], d0)));
            const float m1 = __fmaf_rn(Ci.m[1][2], dz, __fmaf_rn(Ci.m[1][1], d1, __fmul_rn(Ci.m[1][0], d0)));
            const float m2 = __fmaf_rn(Ci.m[2][2], dz, __fmaf_rn(Ci.m[2][1], d1, __fmul_rn(Ci.m[2][0], d0)));
            d2[j] = __fmaf_rn(dz, m2, __fmaf_rn(d1, m1, __fmul_rn(d0, m0)));
        }
        // median by insertion sort in the weights buffer (compute_median, :143-173)
        for (int j = 0; j < k; ++j) w[j] = d2[j];
        for (int a = 1; a < k; ++a) {
            const float key = w[a];
            int b = a;
            while (b > 0 && w[b - 1] > key) {
                w[b] = w[b - 1];
                --b;
            }
            w[b] = key;
        }
        const int mid = k / 2;
        const float median = (k % 2 == 0) ? __fmul_rn(__fadd_rn(w[mid - 1], w[mid]), 0.5f) : w[mid];
        float scale = __fmul_rn(mad_scale, median);
        if (scale < min_scale) scale = min_scale;
        for (int j = 0; j < k; ++j) w[j] = robust_weight(loss, d2[j], scale);
        ok = estimate_cov_weighted(pts, row, k, w, C, mean);
    }
    store_cov16(covs + (size_t)i * 16, C);
}

__global__ void __launch_bounds__(FEAT_THREADS) normals_kernel(const float4* __restrict__ pts, uint32_t n,
                                                               const int32_t* __restrict__ idx, int k,
                                                               float4* __restrict__ normals) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    Sym3 C;
    estimate_cov(pts, idx + (size_t)i * k, k, C);
    normals[i] = normal_of(__ldg(pts + i), mat3_from_sym(C));
}

__global__ void __launch_bounds__(FEAT_THREADS) normals_from_covs_kernel(const float4* __restrict__ pts,
                                                                         const float* __restrict__ covs, uint32_t n,
                                                                         float4* __restrict__ normals) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    normals[i] = normal_of(__ldg(pts + i), load_cov16(covs + (size_t)i * 16));
}

// eigen_utils::symmetric_eigen_decomposition_3x3 on stored covariances (eigen_utils.hpp:443-562):
// evals ascending, evecs 3x3 ROW-major (column k = eigenvector k)
__global__ void __launch_bounds__(FEAT_THREADS) eigen3_kernel(const float* __restrict__ covs, uint32_t n,
                                                              float* __restrict__ evals, float* __restrict__ evecs) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    float ev[3], V[3][3];
    mat3_eigen(load_cov16(covs + (size_t)i * 16), ev, V);
    for (int a = 0; a < 3; ++a) {
        evals[(size_t)i * 3 + a] = ev[a];
        for (int b = 0; b < 3; ++b) evecs[(size_t)i * 9 + a * 3 + b] = V[a][b];
    }
}

// kernel::update_covariance_plane — covariance.hpp:67-74, in place on the 4x4 layout
__global__ void __launch_bounds__(FEAT_THREADS) update_plane_kernel(float* __restrict__ covs, uint32_t n) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    store_cov16(covs + (size_t)i * 16, plane_regularize(load_cov16(covs + (size_t)i * 16)));
}

// packed xyz (12 B per point, what a LiDAR driver delivers) -> PointType xyz1 (types.hpp:11: Vector4f, w = 1)
__global__ void __launch_bounds__(256) expand_xyz_kernel(const float* __restrict__ xyz, uint32_t n, float4* __restrict__ out) {
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        out[i] = make_float4(__ldg(xyz + 3 * (size_t)i), __ldg(xyz + 3 * (size_t)i + 1), __ldg(xyz + 3 * (size_t)i + 2), 1.0f);
}

}  // namespace

// ------------------------------------------------------------------ cloud transform
// transform::transform_async — I/algorithms/common/transform.hpp:45-94, in place.  Same fma chains
// as eigen_utils::multiply (eigen_utils.hpp:88-127): covariance = T (C T^T) with the 4x4 products
// written out (terms with an exact zero factor — the 4th row / column of C — are no-ops and are
// skipped), normal = T n (NOT re-normalised: the reference discards normalize<4>()'s result,
// transform.hpp:24-30), point = T p.  One kernel for all three attributes.
namespace {
__global__ void __launch_bounds__(FEAT_THREADS) transform_cloud_kernel(float4* __restrict__ pts, float* __restrict__ covs,
                                                                       float4* __restrict__ normals, uint32_t n, Xform T) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    const float4 rows[4] = {T.r0, T.r1, T.r2, T.r3};
    if (covs) {
        float4* c = reinterpret_cast<float4*>(covs + (size_t)i * 16);  // column-major: c[j] = column j
        const float4 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
        const float Cm[4][4] = {{c0.x, c1.x, c2.x, c3.x}, {c0.y, c1.y, c2.y, c3.y}, {c0.z, c1.z, c2.z, c3.z},
                                {c0.w, c1.w, c2.w, c3.w}};  // Cm[row][col]
        const float Tm[4][4] = {{rows[0].x, rows[0].y, rows[0].z, rows[0].w}, {rows[1].x, rows[1].y, rows[1].z, rows[1].w},
                                {rows[2].x, rows[2].y, rows[2].z, rows[2].w}, {rows[3].x, rows[3].y, rows[3].z, rows[3].w}};
        float X[4][4], R[4][4];
        // X = C * T^T : X(i,j) = fma chain over k of C(i,k) * T(j,k), k ascending, starting from 0
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = __fmaf_rn(Cm[a][k], Tm[b][k], acc);
                X[a][b] = acc;
            }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = __fmaf_rn(Tm[a][k], X[k][b], acc);
                R[a][b] = acc;
            }
#pragma unroll
        for (int b = 0; b < 4; ++b) c[b] = make_float4(R[0][b], R[1][b], R[2][b], R[3][b]);
    }
    if (normals) {
        const float4 v = normals[i];
        float o[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            o[a] = __fmaf_rn(rows[a].w, v.w, __fmaf_rn(rows[a].z, v.z, __fmaf_rn(rows[a].y, v.y, __fmaf_rn(rows[a].x, v.x, 0.0f))));
        normals[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
    const float4 p = pts[i];
    float o[4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
        o[a] = __fmaf_rn(rows[a].w, p.w, __fmaf_rn(rows[a].z, p.z, __fmaf_rn(rows[a].y, p.y, __fmaf_rn(rows[a].x, p.x, 0.0f))));
    pts[i] = make_float4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------ constant-velocity deskew
// deskew::deskew_point_cloud_constant_velocity — I/algorithms/deskew/relative_pose_deskew.hpp:121-174.  One
// thread per point: tau = clamp(t_ms * 1e-3 / duration, 0, 1), M = se3_exp(tau * twist), point = M p; normal =
// R n and covariance = R (C R^T) with R the rotation of M (quaternion_to_rotation_matrix(so3_exp(tau * omega)),
// :155-157, is that same matrix); a point whose timestamp is not finite is copied (:127-137).  In place when the
// output pointers equal the inputs (every thread reads its own element before it writes it).
struct Twist6 {
    float a[6];
};
__global__ void __launch_bounds__(FEAT_THREADS)
    deskew_kernel(const float4* __restrict__ pin, const float4* __restrict__ nin, const float* __restrict__ cin,
                  const float* __restrict__ ts, uint32_t n, Twist6 tw, float duration, float4* pout, float4* nout, float* cout) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    const float t_s = __fmul_rn(ts[i], 1e-3f);
    const float4 p = pin[i];
    float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 c[4];
    if (nin) nv = nin[i];
    if (cin)
        for (int b = 0; b < 4; ++b) c[b] = reinterpret_cast<const float4*>(cin + (size_t)i * 16)[b];
    if (!isfinite(t_s)) {
        pout[i] = p;
        if (nin) nout[i] = nv;
        if (cin)
            for (int b = 0; b < 4; ++b) reinterpret_cast<float4*>(cout + (size_t)i * 16)[b] = c[b];
        return;
    }
    const float tau = fminf(fmaxf(__fdiv_rn(t_s, duration), 0.0f), 1.0f);
    float a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = __fmul_rn(tw.a[k], tau);
    float M[4][4];
    se3_exp_rm(a, M);
    const float pv[4] = {p.x, p.y, p.z, p.w};
    float o[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = __fmaf_rn(M[r][k], pv[k], acc);
        o[r] = acc;
    }
    pout[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (nin) {
        const float v[3] = {nv.x, nv.y, nv.z};
        float r3[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc = __fmaf_rn(M[r][k], v[k], acc);
            r3[r] = acc;
        }
        nout[i] = make_float4(r3[0], r3[1], r3[2], 0.0f);
    }
    if (cin) {
        const float Cm[3][3] = {{c[0].x, c[1].x, c[2].x}, {c[0].y, c[1].y, c[2].y}, {c[0].z, c[1].z, c[2].z}};
        float X[3][3], R[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int b = 0; b < 3; ++b) {  // X = C R^T
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) acc = __fmaf_rn(Cm[r][k], M[b][k], acc);
                X[r][b] = acc;
            }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) acc = __fmaf_rn(M[r][k], X[k][b], acc);
                R[r][b] = acc;
            }
        float4* co = reinterpret_cast<float4*>(cout + (size_t)i * 16);
#pragma unroll
        for (int b = 0; b < 3; ++b) co[b] = make_float4(R[0][b], R[1][b], R[2][b], 0.0f);
        co[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
}  // namespace

extern "C" {

int spx_covariance(spx_queue_t q, const float* points, size_t n, const int32_t* knn_idx, int k, float* covs) {
    return guard([&] {
        SPX_REQUIRE(q, "[covariance::estimate_async] null queue");
        SPX_REQUIRE(n < (1ull << 31), "[covariance::estimate_async] too many points");
        if (n == 0) return;
        SPX_REQUIRE(points && knn_idx && covs && k >= 1, "[covariance::estimate_async] null pointer or k < 1");
        DeviceGuard g(q->device);
        covariance_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(
            reinterpret_cast<const float4*>(points), (uint32_t)n, knn_idx, k, covs);
        SPX_LAUNCH_CHECK();
    });
}

int spx_covariance_robust(spx_queue_t q, const float* points, size_t n, const int32_t* knn_idx, int k, int robust_loss,
                          float mad_scale, float min_robust_scale, int robust_max_iterations, float* covs) {
    return guard([&] {
        SPX_REQUIRE(q, "[covariance::estimate_robust_async] null queue");
        SPX_REQUIRE(n < (1ull << 31), "[covariance::estimate_robust_async] too many points");
        if (k > ROBUST_MAX_K)
            throw Error(SPX_ERR_INVALID_ARGUMENT, "[covariance::estimate_robust_async] neighbor K is too large. MAX_K is 64");
        SPX_REQUIRE(robust_loss >= SPX_LOSS_NONE && robust_loss <= SPX_LOSS_GEMAN_MCCLURE,
                    "[covariance::estimate_robust_async] unknown robust loss");
        if (n == 0) return;
        SPX_REQUIRE(points && knn_idx && covs && k >= 1, "[covariance::estimate_robust_async] null pointer or k < 1");
        DeviceGuard g(q->device);
        if (robust_loss == SPX_LOSS_NONE) {  // covariance.hpp:241-244: the plain estimate
            covariance_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(
                reinterpret_cast<const float4*>(points), (uint32_t)n, knn_idx, k, covs);
        } else {
            covariance_robust_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(
                reinterpret_cast<const float4*>(points), (uint32_t)n, knn_idx, k, robust_loss, mad_scale, min_robust_scale,
                std::max(robust_max_iterations, 0), covs);
        }
        SPX_LAUNCH_CHECK();
    });
}

int spx_normals(spx_queue_t q, const float* points, size_t n, const int32_t* knn_idx, int k, float* normals) {
    return guard([&] {
        SPX_REQUIRE(q, "[covariance::estimate_normals_async] null queue");
        SPX_REQUIRE(n < (1ull << 31), "[covariance::estimate_normals_async] too many points");
        if (n == 0) return;
        SPX_REQUIRE(points && knn_idx && normals && k >= 1,
                    "[covariance::estimate_normals_async] null pointer or k < 1");
        DeviceGuard g(q->device);
        normals_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(
            reinterpret_cast<const float4*>(points), (uint32_t)n, knn_idx, k, reinterpret_cast<float4*>(normals));
        SPX_LAUNCH_CHECK();
    });
}

int spx_normals_from_covs(spx_queue_t q, const float* points, const float* covs, size_t n, float* normals) {
    return guard([&] {
        SPX_REQUIRE(q, "[covariance::extract_normals_async] null queue");
        SPX_REQUIRE(n < (1ull << 31), "[covariance::extract_normals_async] too many points");
        if (n == 0) return;
        SPX_REQUIRE(covs, "[covariance::extract_normals_async] covariances not computed");
        SPX_REQUIRE(points && normals, "[covariance::extract_normals_async] null pointer");
        DeviceGuard g(q->device);
        normals_from_covs_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(
            reinterpret_cast<const float4*>(points), covs, (uint32_t)n, reinterpret_cast<float4*>(normals));
        SPX_LAUNCH_CHECK();
    });
}

int spx_points_from_xyz(spx_queue_t q, const float* xyz, size_t n, float* points) {
    return guard([&] {
        SPX_REQUIRE(q, "[PointCloudShared] null queue");
        SPX_REQUIRE(n < (1ull << 31), "[PointCloudShared] too many points");
        if (n == 0) return;
        SPX_REQUIRE(xyz && points, "[PointCloudShared] null pointer");
        DeviceGuard g(q->device);
        expand_xyz_kernel<<<std::min(div_up(n, 256), q->sm_count * 16), 256, 0, q->stream>>>(xyz, (uint32_t)n,
                                                                                       reinterpret_cast<float4*>(points));
        SPX_LAUNCH_CHECK();
    });
}

int spx_eigen3(spx_queue_t q, const float* covs, size_t n, float* evals, float* evecs) {
    return guard([&] {
        SPX_REQUIRE(q, "[eigen_utils::symmetric_eigen_decomposition_3x3] null queue");
        SPX_REQUIRE(n < (1ull << 31), "[eigen_utils::symmetric_eigen_decomposition_3x3] too many matrices");
        if (n == 0) return;
        SPX_REQUIRE(covs && evals && evecs, "[eigen_utils::symmetric_eigen_decomposition_3x3] null pointer");
        DeviceGuard g(q->device);
        eigen3_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(covs, (uint32_t)n, evals, evecs);
        SPX_LAUNCH_CHECK();
    });
}

int spx_covariance_update_plane(spx_queue_t q, float* covs, size_t n) {
    return guard([&] {
        SPX_REQUIRE(q, "[covariance::update_covariance_plane] null queue");
        SPX_REQUIRE(n < (1ull << 31), "[covariance::update_covariance_plane] too many matrices");
        if (n == 0) return;
        SPX_REQUIRE(covs, "[covariance::update_covariance_plane] null pointer");
        DeviceGuard g(q->device);
        update_plane_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(covs, (uint32_t)n);
        SPX_LAUNCH_CHECK();
    });
}

int spx_transform(spx_queue_t q, float* points, float* covs, float* normals, size_t n, const float* T_host) {
    return guard([&] {
        SPX_REQUIRE(q && T_host, "[transform::transform] null argument");
        SPX_REQUIRE(n < (1ull << 31), "[transform::transform] too many points");
        if (n == 0) return;
        SPX_REQUIRE(points, "[transform::transform] null points");
        DeviceGuard g(q->device);
        transform_cloud_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(
            reinterpret_cast<float4*>(points), covs, reinterpret_cast<float4*>(normals), (uint32_t)n,
            xform_from_colmajor(T_host));
        SPX_LAUNCH_CHECK();
    });
}

int spx_deskew_constant_velocity(spx_queue_t q, const float* points, const float* normals, const float* covs,
                                 const float* timestamp_offsets_ms, size_t n, const float* twist6_host,
                                 float scan_duration_s, float* points_out, float* normals_out, float* covs_out) {
    return guard([&] {
        SPX_REQUIRE(q && twist6_host, "[deskew_point_cloud_constant_velocity] null argument");
        SPX_REQUIRE(n < (1ull << 31), "[deskew_point_cloud_constant_velocity] too many points");
        SPX_REQUIRE(scan_duration_s > 0.0f, "[deskew_point_cloud_constant_velocity] scan duration must be positive");
        if (n == 0) return;
        SPX_REQUIRE(points && points_out && timestamp_offsets_ms, "[deskew_point_cloud_constant_velocity] null points or timestamps");
        SPX_REQUIRE((normals == nullptr) == (normals_out == nullptr) && (covs == nullptr) == (covs_out == nullptr),
                    "[deskew_point_cloud_constant_velocity] an attribute needs both its input and its output");
        DeviceGuard g(q->device);
        Twist6 tw;
        for (int k = 0; k < 6; ++k) tw.a[k] = twist6_host[k];
        deskew_kernel<<<div_up(n, FEAT_THREADS), FEAT_THREADS, 0, q->stream>>>(
            reinterpret_cast<const float4*>(points), reinterpret_cast<const float4*>(normals), covs, timestamp_offsets_ms,
            (uint32_t)n, tw, scan_duration_s, reinterpret_cast<float4*>(points_out), reinterpret_cast<float4*>(normals_out),
            covs_out);
        SPX_LAUNCH_CHECK();
    });
}

int spx_se3_log(const float* T_host, float* twist6_host) {
    return guard([&] {
        SPX_REQUIRE(T_host && twist6_host, "[lie::se3_log] null argument");
        float T[4][4];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) T[r][c] = T_host[c * 4 + r];
        se3_log_rm(T, twist6_host);
    });
}

}  // extern "C"
