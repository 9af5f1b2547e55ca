// Per-point covariance and normal estimation from k neighbours
// (I/algorithms/feature/covariance.hpp:16-74,260-311,417-495).
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

__device__ __forceinline__ void store_cov16(float* __restrict__ out, const Sym3& C) {
    float4* o = reinterpret_cast<float4*>(out);
    o[0] = make_float4(C.xx, C.xy, C.xz, 0.f);
    o[1] = make_float4(C.xy, C.yy, C.yz, 0.f);
    o[2] = make_float4(C.xz, C.yz, C.zz, 0.f);
    o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// extract_normal — covariance.hpp:49-65: eigenvector of the smallest eigenvalue, kept when
// dot(n, p) <= 1.0 and negated otherwise (the reference's literal test), w = 0.
__device__ __forceinline__ float4 normal_of(const float4 p, const Sym3& C) {
    float ev[3], V[3][3];
    sym_eigen3(C, ev, V);
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
    store_cov16(covs + (size_t)i * 16, C);
}

__global__ void __launch_bounds__(FEAT_THREADS) normals_kernel(const float4* __restrict__ pts, uint32_t n,
                                                               const int32_t* __restrict__ idx, int k,
                                                               float4* __restrict__ normals) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    Sym3 C;
    estimate_cov(pts, idx + (size_t)i * k, k, C);
    normals[i] = normal_of(__ldg(pts + i), C);
}

__global__ void __launch_bounds__(FEAT_THREADS) normals_from_covs_kernel(const float4* __restrict__ pts,
                                                                         const float* __restrict__ covs, uint32_t n,
                                                                         float4* __restrict__ normals) {
    const uint32_t i = blockIdx.x * FEAT_THREADS + threadIdx.x;
    if (i >= n) return;
    normals[i] = normal_of(__ldg(pts + i), load_cov16(covs + (size_t)i * 16));
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

}  // extern "C"
