// mapping::VoxelHashMap — the submap on the far side of the registration path.
// Replaces I/algorithms/mapping/voxel_hash_map.hpp:22-1066.
//
// Same table as the reference, so the semantics carry over entry for entry:
//   * open addressing, double hashing  slot_j = (key + j * hash2(key)) % capacity, hash2 = (cap-2) - key % (cap-2)
//     (:607-612), at most 100 probes (:498) — a point whose probe sequence is exhausted is dropped (:583-604);
//   * capacities from the same list of primes (:481-482), rehash into the next one when
//     voxel_num / capacity exceeds rehash_threshold BEFORE an insertion round (:120-126);
//   * a slot holds the key, sum x/y/z + count, the sum of the LOG-Euclidean images of the rotated covariances
//     (6 upper-triangle entries, :420-476), sum rgba, sum intensity and the staleness stamp of its last update;
//   * every remove_old_data_cycle-th call removes the voxels not touched for more than max_staleness calls (:794-845);
//   * the export keeps a voxel iff count >= min_num_point and its centroid lies in the axis-aligned box
//     (:395-418), in SLOT order (the reference's NVIDIA branch: flags -> prefix sum -> write, :950-1028), and
//     writes centroid, exp of the mean log-covariance, mean rgba, mean intensity (:348-393).
// What differs is how it runs: the reference sorts and merges inside a work-group of 32/64 points before its
// atomics (:753-756) and blocks the host three times per call; here one thread per point claims the slot with a
// 64-bit CAS and adds with RED.F32 (the accumulation ORDER is unspecified on both sides — fp32 atomics — so sums
// agree to rounding, not bit for bit; every test states that tolerance), and one call = one launch + one 4-byte
// read-back of the voxel count (the façade needs it for the next rehash decision and to size the export).
// Arithmetic of the covariance images is the library's exact restatement of eigen_utils.hpp:443-562,646-677.
#include <memory>

#include "spx_common.cuh"
#include "spx_math.cuh"
#include "spx_scan.cuh"

using namespace spx;

namespace {

constexpr unsigned long long VM_INVALID = ~0ull;  // VoxelConstants::invalid_coord (voxel_constants.hpp:12)
constexpr int VM_MAX_PROBE = 100;                 // voxel_hash_map.hpp:498
constexpr int VM_THREADS = 256;
const size_t VM_CAPACITIES[11] = {30029,   60013,   120011,  240007,   480013,  960017,
                                  1920001, 3840007, 7680017, 15360013, 30720007};  // :481-482

struct VmCore {  // VoxelCoreData :256-262
    float sx, sy, sz;
    uint32_t count;
};
struct VmCov {  // VoxelCovarianceData :274-282
    float xx, xy, xz, yy, yz, zz;
};

struct VmTable {
    unsigned long long* key;
    VmCore* core;
    VmCov* cov;
    float4* color;
    float* intensity;
    uint32_t* last;
    unsigned long long cap;
};

// compute_voxel_bit — voxel_constants.hpp:36-62
__device__ __forceinline__ unsigned long long vm_key(const float4 p, float inv) {
    if (!isfinite(p.x) || !isfinite(p.y) || !isfinite(p.z)) return VM_INVALID;
    const float fx = floorf(__fmul_rn(p.x, inv)), fy = floorf(__fmul_rn(p.y, inv)), fz = floorf(__fmul_rn(p.z, inv));
    const float lim = 1048576.0f;
    if (!(fx >= -lim && fx < lim && fy >= -lim && fy < lim && fz >= -lim && fz < lim)) return VM_INVALID;
    const unsigned long long c0 = (unsigned long long)((long long)fx + (1 << 20));
    const unsigned long long c1 = (unsigned long long)((long long)fy + (1 << 20));
    const unsigned long long c2 = (unsigned long long)((long long)fz + (1 << 20));
    return c0 | (c1 << 21) | (c2 << 42);
}

// f(A) = V diag(f(l)) V^T, symmetrised — log_spd_3x3 / exp_spd_3x3 (eigen_utils.hpp:646-677):
// multiply<3,3,3>(multiply<3,3,3>(V, D), V^T) with fma accumulation over k ascending, then ensure_symmetric.
template <bool LOG>
__device__ inline Mat3 spd_function(const Mat3& A, float min_eigenvalue = 1e-6f) {
    float ev[3], V[3][3];
    mat3_eigen(A, ev, V);
    float f[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) f[i] = LOG ? cr_logf(fmaxf(ev[i], min_eigenvalue)) : (float)exp((double)ev[i]);
    float VD[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) VD[i][j] = __fmul_rn(V[i][j], f[j]);
    Mat3 P;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            P.m[i][j] = __fmaf_rn(VD[i][2], V[j][2], __fmaf_rn(VD[i][1], V[j][1], __fmul_rn(VD[i][0], V[j][0])));
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.m[i][j] = (i == j) ? P.m[i][j] : __fmul_rn(__fadd_rn(P.m[i][j], P.m[j][i]), 0.5f);
    return r;
}

// rotate_covariance_upper_triangle (:420-456) then encode_covariance_for_aggregation (:458-476)
__device__ inline VmCov encode_covariance(const float* __restrict__ cov16, const Xform& T) {
    const Mat3 C = load_cov16(cov16);
    const float cxx = C.m[0][0], cxy = C.m[0][1], cxz = C.m[0][2], cyy = C.m[1][1], cyz = C.m[1][2], czz = C.m[2][2];
    const float r00 = T.r0.x, r01 = T.r0.y, r02 = T.r0.z, r10 = T.r1.x, r11 = T.r1.y, r12 = T.r1.z, r20 = T.r2.x,
                r21 = T.r2.y, r22 = T.r2.z;
    const float a00 = __fmaf_rn(r02, cxz, __fmaf_rn(r01, cxy, __fmul_rn(r00, cxx)));
    const float a01 = __fmaf_rn(r02, cyz, __fmaf_rn(r01, cyy, __fmul_rn(r00, cxy)));
    const float a02 = __fmaf_rn(r02, czz, __fmaf_rn(r01, cyz, __fmul_rn(r00, cxz)));
    const float a10 = __fmaf_rn(r12, cxz, __fmaf_rn(r11, cxy, __fmul_rn(r10, cxx)));
    const float a11 = __fmaf_rn(r12, cyz, __fmaf_rn(r11, cyy, __fmul_rn(r10, cxy)));
    const float a12 = __fmaf_rn(r12, czz, __fmaf_rn(r11, cyz, __fmul_rn(r10, cxz)));
    const float a20 = __fmaf_rn(r22, cxz, __fmaf_rn(r21, cxy, __fmul_rn(r20, cxx)));
    const float a21 = __fmaf_rn(r22, cyz, __fmaf_rn(r21, cyy, __fmul_rn(r20, cxy)));
    const float a22 = __fmaf_rn(r22, czz, __fmaf_rn(r21, cyz, __fmul_rn(r20, cxz)));
    Sym3 s;
    s.xx = __fmaf_rn(a02, r02, __fmaf_rn(a01, r01, __fmul_rn(a00, r00)));
    s.xy = __fmaf_rn(a02, r12, __fmaf_rn(a01, r11, __fmul_rn(a00, r10)));
    s.xz = __fmaf_rn(a02, r22, __fmaf_rn(a01, r21, __fmul_rn(a00, r20)));
    s.yy = __fmaf_rn(a12, r12, __fmaf_rn(a11, r11, __fmul_rn(a10, r10)));
    s.yz = __fmaf_rn(a12, r22, __fmaf_rn(a11, r21, __fmul_rn(a10, r20)));
    s.zz = __fmaf_rn(a22, r22, __fmaf_rn(a21, r21, __fmul_rn(a20, r20)));
    const Mat3 L = spd_function<true>(mat3_from_sym(s));
    return VmCov{L.m[0][0], L.m[0][1], L.m[0][2], L.m[1][1], L.m[1][2], L.m[2][2]};
}

struct VmEntry {
    VmCore core;
    VmCov cov;
    float4 color;
    float intensity;
};

// global_reduction — :574-605.  The probe sequence is walked incrementally: slot_{j+1} = (slot_j + step) % cap
// equals (key + (j+1) * step) % cap because neither expression wraps 64 bits (key < 2^63, 100 * step < 2^32).
__device__ __forceinline__ void vm_insert(const VmTable t, unsigned long long key, const VmEntry& e, uint32_t stamp,
                                          bool has_cov, bool has_rgb, bool has_intensity, uint32_t* voxel_counter) {
    if (key == VM_INVALID) return;
    unsigned long long slot = key % t.cap;
    const unsigned long long step = (t.cap - 2) - key % (t.cap - 2);
    for (int j = 0; j < VM_MAX_PROBE; ++j) {
        unsigned long long seen = t.key[slot];
        if (seen == VM_INVALID) {
            seen = atomicCAS(t.key + slot, VM_INVALID, key);
            if (seen == VM_INVALID) {
                atomicAdd(voxel_counter, 1u);
                seen = key;
            }
        }
        if (seen == key) {
            VmCore* c = t.core + slot;
            atomicAdd(&c->sx, e.core.sx);
            atomicAdd(&c->sy, e.core.sy);
            atomicAdd(&c->sz, e.core.sz);
            atomicAdd(&c->count, e.core.count);
            if (has_cov) {
                float* v = reinterpret_cast<float*>(t.cov + slot);
                atomicAdd(v + 0, e.cov.xx);
                atomicAdd(v + 1, e.cov.xy);
                atomicAdd(v + 2, e.cov.xz);
                atomicAdd(v + 3, e.cov.yy);
                atomicAdd(v + 4, e.cov.yz);
                atomicAdd(v + 5, e.cov.zz);
            }
            if (has_rgb) atomicAdd(t.color + slot, e.color);  // RED.E.ADD.F32x4 (sm_90+): one 16-byte atomic
            if (has_intensity) atomicAdd(t.intensity + slot, e.intensity);
            t.last[slot] = stamp;  // atomic_store_timestamp :343-346 (every writer of a round stores the same value)
            return;
        }
        slot += step;
        if (slot >= t.cap) slot -= t.cap;
    }
}

// add_point_cloud_impl — :614-792 (load_entry :661-704)
__global__ void __launch_bounds__(VM_THREADS) vm_add_kernel(VmTable t, const float4* __restrict__ pts,
                                                            const float* __restrict__ covs, const float4* __restrict__ rgb,
                                                            const float* __restrict__ intensity, uint32_t n, Xform T,
                                                            float inv, uint32_t stamp, uint32_t* voxel_counter) {
    const uint32_t i = blockIdx.x * VM_THREADS + threadIdx.x;
    if (i >= n) return;
    const float4 w = transform_point(T, __ldg(pts + i));
    const unsigned long long key = vm_key(w, inv);
    if (key == VM_INVALID) return;
    VmEntry e;
    e.core = VmCore{w.x, w.y, w.z, 1u};
    if (covs) e.cov = encode_covariance(covs + (size_t)i * 16, T);
    if (rgb) e.color = __ldg(rgb + i);
    if (intensity) e.intensity = __ldg(intensity + i);
    vm_insert(t, key, e, stamp, covs != nullptr, rgb != nullptr, intensity != nullptr, voxel_counter);
}

// rehash — :847-934: every live slot of the old table is re-inserted with its sums and its own stamp
__global__ void __launch_bounds__(VM_THREADS) vm_rehash_kernel(VmTable old_t, VmTable new_t, bool has_cov, bool has_rgb,
                                                               bool has_intensity, uint32_t* voxel_counter) {
    const unsigned long long i = (unsigned long long)blockIdx.x * VM_THREADS + threadIdx.x;
    if (i >= old_t.cap) return;
    const unsigned long long key = old_t.key[i];
    if (key == VM_INVALID) return;
    VmEntry e;
    e.core = old_t.core[i];
    if (has_cov) e.cov = old_t.cov[i];
    if (has_rgb) e.color = old_t.color[i];
    if (has_intensity) e.intensity = old_t.intensity[i];
    vm_insert(new_t, key, e, old_t.last[i], has_cov, has_rgb, has_intensity, voxel_counter);
}

// remove_old_data_impl — :794-845
__global__ void __launch_bounds__(VM_THREADS) vm_remove_old_kernel(VmTable t, uint32_t remove_staleness,
                                                                   uint32_t* voxel_counter) {
    const unsigned long long i = (unsigned long long)blockIdx.x * VM_THREADS + threadIdx.x;
    uint32_t kept = 0;
    if (i < t.cap && t.key[i] != VM_INVALID) {
        if (t.last[i] >= remove_staleness) {
            kept = 1;
        } else {
            t.key[i] = VM_INVALID;
            t.core[i] = VmCore{0.f, 0.f, 0.f, 0u};
            t.cov[i] = VmCov{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            t.color[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            t.intensity[i] = 0.f;
            t.last[i] = 0;
        }
    }
    const uint32_t warp_kept = __popc(__ballot_sync(0xffffffffu, kept));
    if ((threadIdx.x & 31) == 0 && warp_kept) atomicAdd(voxel_counter, warp_kept);
}

// should_include_voxel / centroid_inside_bbox — :395-418
__global__ void __launch_bounds__(VM_THREADS) vm_flag_kernel(VmTable t, uint32_t min_num_point, float3 lo, float3 hi,
                                                             uint32_t* __restrict__ flags) {
    const unsigned long long i = (unsigned long long)blockIdx.x * VM_THREADS + threadIdx.x;
    if (i >= t.cap) return;
    uint32_t keep = 0;
    const VmCore c = t.core[i];
    if (t.key[i] != VM_INVALID && c.count >= min_num_point && c.count != 0u) {
        const float ic = __fdiv_rn(1.0f, (float)c.count);
        const float x = __fmul_rn(c.sx, ic), y = __fmul_rn(c.sy, ic), z = __fmul_rn(c.sz, ic);
        keep = (x >= lo.x && x <= hi.x) && (y >= lo.y && y <= hi.y) && (z >= lo.z && z <= hi.z);
    }
    flags[i] = keep;
}

// compute_averaged_attributes — :348-393 (the branch count < min_num_point never runs for a flagged voxel)
__global__ void __launch_bounds__(VM_THREADS) vm_export_kernel(VmTable t, const uint32_t* __restrict__ flags,
                                                               const uint32_t* __restrict__ pos, float4* __restrict__ out_pts,
                                                               float* __restrict__ out_covs, float4* __restrict__ out_rgb,
                                                               float* __restrict__ out_intensity,
                                                               unsigned long long* __restrict__ out_keys) {
    const unsigned long long i = (unsigned long long)blockIdx.x * VM_THREADS + threadIdx.x;
    if (i >= t.cap || !flags[i]) return;
    const uint32_t o = pos[i];
    const VmCore c = t.core[i];
    const float ic = __fdiv_rn(1.0f, (float)c.count);
    out_pts[o] = make_float4(__fmul_rn(c.sx, ic), __fmul_rn(c.sy, ic), __fmul_rn(c.sz, ic), 1.0f);
    if (out_covs) {
        const VmCov v = t.cov[i];
        Sym3 s{__fmul_rn(v.xx, ic), __fmul_rn(v.xy, ic), __fmul_rn(v.xz, ic),
               __fmul_rn(v.yy, ic), __fmul_rn(v.yz, ic), __fmul_rn(v.zz, ic)};
        store_cov16(out_covs + (size_t)o * 16, spd_function<false>(mat3_from_sym(s)));
    }
    if (out_rgb) {
        const float4 k = t.color[i];
        out_rgb[o] = make_float4(__fmul_rn(k.x, ic), __fmul_rn(k.y, ic), __fmul_rn(k.z, ic), __fmul_rn(k.w, ic));
    }
    if (out_intensity) out_intensity[o] = __fmul_rn(t.intensity[i], ic);
    if (out_keys) out_keys[o] = t.key[i];
}

// eigen_utils::log_spd_3x3 / exp_spd_3x3 on column-major 3x3 matrices (eigen_utils.hpp:646-677)
__global__ void __launch_bounds__(VM_THREADS) spd_function_kernel(const float* __restrict__ in, uint32_t n, int is_log,
                                                                  float min_eigenvalue, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * VM_THREADS + threadIdx.x;
    if (i >= n) return;
    Mat3 A;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) A.m[r][c] = in[(size_t)i * 9 + c * 3 + r];
    const Mat3 R = is_log ? spd_function<true>(A, min_eigenvalue) : spd_function<false>(A);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) out[(size_t)i * 9 + c * 3 + r] = R.m[r][c];
}

// compute_overlap_ratio — :194-246
__global__ void __launch_bounds__(VM_THREADS) vm_overlap_kernel(VmTable t, const float4* __restrict__ pts, uint32_t n,
                                                                Xform T, float inv, uint32_t min_num_point,
                                                                uint32_t* counter) {
    const uint32_t i = blockIdx.x * VM_THREADS + threadIdx.x;
    uint32_t hit = 0;
    if (i < n) {
        const unsigned long long key = vm_key(transform_point(T, __ldg(pts + i)), inv);
        if (key != VM_INVALID) {
            unsigned long long slot = key % t.cap;
            const unsigned long long step = (t.cap - 2) - key % (t.cap - 2);
            for (int j = 0; j < VM_MAX_PROBE; ++j) {
                const unsigned long long seen = t.key[slot];
                if (seen == key) {
                    hit = t.core[slot].count >= min_num_point;
                    break;
                }
                if (seen == VM_INVALID) break;
                slot += step;
                if (slot >= t.cap) slot -= t.cap;
            }
        }
    }
    const uint32_t warp_hits = __popc(__ballot_sync(0xffffffffu, hit));
    if ((threadIdx.x & 31) == 0 && warp_hits) atomicAdd(counter, warp_hits);
}

}  // namespace

struct spx_voxelmap_s {
    spx_queue_t q = nullptr;
    float voxel = 0.f, inv = 0.f;
    VmTable t{};
    uint32_t* counter = nullptr;  // device: [0] voxel count of the running round, [1] overlap hits
    uint32_t staleness_counter = 0, max_staleness = 100, remove_old_data_cycle = 10, min_num_point = 1;
    float rehash_threshold = 0.7f;
    size_t voxel_num = 0;
    bool has_cov = false, has_rgb = false, has_intensity = false;
};

namespace {

void vm_free_table(VmTable& t) {
    cudaFree(t.key);
    cudaFree(t.core);
    cudaFree(t.cov);
    cudaFree(t.color);
    cudaFree(t.intensity);
    cudaFree(t.last);
    t = VmTable{};
}

// allocate_storage (:519-533) + the fills of clear() (:98-111): keys all ones, everything else zero
VmTable vm_alloc_table(size_t cap, cudaStream_t st) {
    VmTable t{};
    t.cap = cap;
    try {
        SPX_CUDA(cudaMalloc(&t.key, cap * sizeof(unsigned long long)));
        SPX_CUDA(cudaMalloc(&t.core, cap * sizeof(VmCore)));
        SPX_CUDA(cudaMalloc(&t.cov, cap * sizeof(VmCov)));
        SPX_CUDA(cudaMalloc(&t.color, cap * sizeof(float4)));
        SPX_CUDA(cudaMalloc(&t.intensity, cap * sizeof(float)));
        SPX_CUDA(cudaMalloc(&t.last, cap * sizeof(uint32_t)));
        SPX_CUDA(cudaMemsetAsync(t.key, 0xff, cap * sizeof(unsigned long long), st));
        SPX_CUDA(cudaMemsetAsync(t.core, 0, cap * sizeof(VmCore), st));
        SPX_CUDA(cudaMemsetAsync(t.cov, 0, cap * sizeof(VmCov), st));
        SPX_CUDA(cudaMemsetAsync(t.color, 0, cap * sizeof(float4), st));
        SPX_CUDA(cudaMemsetAsync(t.intensity, 0, cap * sizeof(float), st));
        SPX_CUDA(cudaMemsetAsync(t.last, 0, cap * sizeof(uint32_t), st));
    } catch (...) {
        vm_free_table(t);
        throw;
    }
    return t;
}

uint32_t vm_read_counter(spx_voxelmap_t m, int which) {
    uint32_t* h = static_cast<uint32_t*>(m->q->pinned_get(64));
    SPX_CUDA(cudaMemcpyAsync(h, m->counter + which, 4, cudaMemcpyDeviceToHost, m->q->stream));
    m->q->sync();
    return *h;
}

void vm_set_voxel_num(spx_voxelmap_t m, size_t v) {  // update_voxel_num_and_flags :510-517
    m->voxel_num = v;
    if (v == 0) m->has_cov = m->has_rgb = m->has_intensity = false;
}

void vm_remove_old(spx_voxelmap_t m) {  // :794-845
    if (m->staleness_counter <= m->max_staleness) return;
    cudaStream_t st = m->q->stream;
    SPX_CUDA(cudaMemsetAsync(m->counter, 0, 4, st));
    vm_remove_old_kernel<<<div_up(m->t.cap, VM_THREADS), VM_THREADS, 0, st>>>(
        m->t, m->staleness_counter - m->max_staleness, m->counter);
    SPX_LAUNCH_CHECK();
    vm_set_voxel_num(m, vm_read_counter(m, 0));
}

void vm_rehash(spx_voxelmap_t m, size_t new_cap) {  // :847-934
    if (m->t.cap >= new_cap) return;
    cudaStream_t st = m->q->stream;
    VmTable nt = vm_alloc_table(new_cap, st);
    SPX_CUDA(cudaMemsetAsync(m->counter, 0, 4, st));
    vm_rehash_kernel<<<div_up(m->t.cap, VM_THREADS), VM_THREADS, 0, st>>>(m->t, nt, m->has_cov, m->has_rgb,
                                                                         m->has_intensity, m->counter);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        vm_free_table(nt);
        SPX_CUDA(e);
    }
    const uint32_t v = vm_read_counter(m, 0);  // also: the old table is no longer read
    vm_free_table(m->t);
    m->t = nt;
    vm_set_voxel_num(m, v);
}

}  // namespace

extern "C" {

int spx_voxelmap_create(spx_queue_t q, float voxel_size, spx_voxelmap_t* out) {
    if (out) *out = nullptr;
    return guard([&] {
        SPX_REQUIRE(q && out, "[VoxelHashMap] null argument");
        SPX_REQUIRE(voxel_size > 0.0f, "voxel_size must be positive.");
        DeviceGuard dg(q->device);
        std::unique_ptr<spx_voxelmap_s> m(new spx_voxelmap_s);
        m->q = q;
        m->voxel = voxel_size;
        m->inv = 1.0f / voxel_size;
        SPX_CUDA(cudaMalloc(&m->counter, 64));
        try {
            SPX_CUDA(cudaMemsetAsync(m->counter, 0, 64, q->stream));
            m->t = vm_alloc_table(VM_CAPACITIES[0], q->stream);
        } catch (...) {
            cudaFree(m->counter);
            throw;
        }
        *out = m.release();
    });
}

int spx_voxelmap_destroy(spx_voxelmap_t m) {
    return guard([&] {
        if (!m) return;
        if (queue_is_live(m->q)) {
            DeviceGuard dg(m->q->device);
            cudaStreamSynchronize(m->q->stream);
        } else {
            cudaDeviceSynchronize();
        }
        vm_free_table(m->t);
        cudaFree(m->counter);
        delete m;
    });
}

int spx_voxelmap_set_params(spx_voxelmap_t m, float voxel_size, uint32_t max_staleness, uint32_t remove_old_data_cycle,
                            float rehash_threshold, uint32_t min_num_point) {
    return guard([&] {
        SPX_REQUIRE(m, "[VoxelHashMap] null map");
        SPX_REQUIRE(voxel_size > 0.0f, "voxel_size must be positive.");
        m->voxel = voxel_size;
        m->inv = 1.0f / voxel_size;
        m->max_staleness = max_staleness;
        m->remove_old_data_cycle = remove_old_data_cycle;
        m->rehash_threshold = rehash_threshold;
        m->min_num_point = min_num_point;
    });
}

int spx_voxelmap_clear(spx_voxelmap_t m) {  // :83-112
    return guard([&] {
        SPX_REQUIRE(m, "[VoxelHashMap] null map");
        DeviceGuard dg(m->q->device);
        cudaStream_t st = m->q->stream;
        if (m->t.cap != VM_CAPACITIES[0]) {
            SPX_CUDA(cudaStreamSynchronize(st));
            VmTable nt = vm_alloc_table(VM_CAPACITIES[0], st);
            vm_free_table(m->t);
            m->t = nt;
        } else {
            const size_t cap = m->t.cap;
            SPX_CUDA(cudaMemsetAsync(m->t.key, 0xff, cap * sizeof(unsigned long long), st));
            SPX_CUDA(cudaMemsetAsync(m->t.core, 0, cap * sizeof(VmCore), st));
            SPX_CUDA(cudaMemsetAsync(m->t.cov, 0, cap * sizeof(VmCov), st));
            SPX_CUDA(cudaMemsetAsync(m->t.color, 0, cap * sizeof(float4), st));
            SPX_CUDA(cudaMemsetAsync(m->t.intensity, 0, cap * sizeof(float), st));
            SPX_CUDA(cudaMemsetAsync(m->t.last, 0, cap * sizeof(uint32_t), st));
        }
        m->voxel_num = 0;
        m->staleness_counter = 0;
        m->has_cov = m->has_rgb = m->has_intensity = false;
    });
}

int spx_voxelmap_add(spx_voxelmap_t m, const float* points, const float* covs, const float* rgb, const float* intensities,
                     size_t n, const float* sensor_pose16) {  // add_point_cloud :117-140
    return guard([&] {
        SPX_REQUIRE(m, "[VoxelHashMap::add_point_cloud] null map");
        SPX_REQUIRE(n < (1ull << 31), "[VoxelHashMap::add_point_cloud] too many points");
        SPX_REQUIRE(n == 0 || points, "[VoxelHashMap::add_point_cloud] null points");
        DeviceGuard dg(m->q->device);
        cudaStream_t st = m->q->stream;
        if (m->rehash_threshold < (float)m->voxel_num / (float)m->t.cap) {
            size_t next = m->t.cap;
            for (size_t c : VM_CAPACITIES)
                if (c > m->t.cap) {
                    next = c;
                    break;
                }
            if (next > m->t.cap) vm_rehash(m, next);
        }
        if (n > 0) {
            m->has_cov |= covs != nullptr;
            m->has_rgb |= rgb != nullptr;
            m->has_intensity |= intensities != nullptr;
            const uint32_t start = (uint32_t)m->voxel_num;
            SPX_CUDA(cudaMemcpyAsync(m->counter, &start, 4, cudaMemcpyHostToDevice, st));
            const Xform T = sensor_pose16 ? xform_from_colmajor(sensor_pose16) : xform_identity();
            vm_add_kernel<<<div_up(n, VM_THREADS), VM_THREADS, 0, st>>>(
                m->t, reinterpret_cast<const float4*>(points), covs, reinterpret_cast<const float4*>(rgb), intensities,
                (uint32_t)n, T, m->inv, m->staleness_counter, m->counter);
            SPX_LAUNCH_CHECK();
            m->voxel_num = vm_read_counter(m, 0);
        }
        if (m->remove_old_data_cycle > 0 && (m->staleness_counter % m->remove_old_data_cycle) == 0) vm_remove_old(m);
        ++m->staleness_counter;
    });
}

int spx_voxelmap_remove_old(spx_voxelmap_t m) {
    return guard([&] {
        SPX_REQUIRE(m, "[VoxelHashMap::remove_old_data] null map");
        DeviceGuard dg(m->q->device);
        vm_remove_old(m);
    });
}

int spx_voxelmap_info(spx_voxelmap_t m, uint64_t* capacity, uint64_t* voxel_num, uint32_t* staleness_counter,
                      int* has_cov, int* has_rgb, int* has_intensity) {
    return guard([&] {
        SPX_REQUIRE(m, "[VoxelHashMap] null map");
        if (capacity) *capacity = m->t.cap;
        if (voxel_num) *voxel_num = m->voxel_num;
        if (staleness_counter) *staleness_counter = m->staleness_counter;
        if (has_cov) *has_cov = m->has_cov;
        if (has_rgb) *has_rgb = m->has_rgb;
        if (has_intensity) *has_intensity = m->has_intensity;
    });
}

int spx_voxelmap_downsample(spx_voxelmap_t m, const float* center3, float distance, float* out_points, float* out_covs,
                            float* out_rgb, float* out_intensities, uint64_t* out_keys, size_t out_capacity,
                            size_t* m_host) {  // downsampling :146-188, downsampling_impl :936-1065
    return guard([&] {
        SPX_REQUIRE(m && m_host && center3, "[VoxelHashMap::downsampling] null argument");
        *m_host = 0;
        if (m->voxel_num == 0) return;
        SPX_REQUIRE(out_points, "[VoxelHashMap::downsampling] null output");
        SPX_REQUIRE(out_capacity >= m->voxel_num, "[VoxelHashMap::downsampling] outputs must hold voxel_num entries");
        spx_queue_t q = m->q;
        DeviceGuard dg(q->device);
        cudaStream_t st = q->stream;
        const size_t cap = m->t.cap;
        q->arena_reset();
        q->arena_reserve(cap * 8 + scan_scratch_elems(cap) * 4 + 4096);
        uint32_t* flags = q->take<uint32_t>(cap);
        uint32_t* pos = q->take<uint32_t>(cap);
        uint32_t* scan_tmp = q->take<uint32_t>(scan_scratch_elems(cap));
        uint32_t* total_dev = q->take<uint32_t>(16);
        const float3 lo = make_float3(center3[0] - distance, center3[1] - distance, center3[2] - distance);
        const float3 hi = make_float3(center3[0] + distance, center3[1] + distance, center3[2] + distance);
        vm_flag_kernel<<<div_up(cap, VM_THREADS), VM_THREADS, 0, st>>>(m->t, m->min_num_point, lo, hi, flags);
        SPX_LAUNCH_CHECK();
        exclusive_scan_u32(st, flags, pos, cap, scan_tmp, total_dev);
        vm_export_kernel<<<div_up(cap, VM_THREADS), VM_THREADS, 0, st>>>(
            m->t, flags, pos, reinterpret_cast<float4*>(out_points), m->has_cov ? out_covs : nullptr,
            m->has_rgb ? reinterpret_cast<float4*>(out_rgb) : nullptr, m->has_intensity ? out_intensities : nullptr,
            reinterpret_cast<unsigned long long*>(out_keys));
        SPX_LAUNCH_CHECK();
        uint32_t* htotal = static_cast<uint32_t*>(q->pinned_get(64));
        SPX_CUDA(cudaMemcpyAsync(htotal, total_dev, 4, cudaMemcpyDeviceToHost, st));
        q->sync();
        *m_host = *htotal;
    });
}

int spx_spd_function(spx_queue_t q, const float* mats, size_t n, int is_log, float min_eigenvalue, float* out) {
    return guard([&] {
        SPX_REQUIRE(q, "[eigen_utils::log_spd_3x3] null queue");
        SPX_REQUIRE(n < (1ull << 31), "[eigen_utils::log_spd_3x3] too many matrices");
        if (n == 0) return;
        SPX_REQUIRE(mats && out, "[eigen_utils::log_spd_3x3] null pointer");
        DeviceGuard dg(q->device);
        spd_function_kernel<<<div_up(n, VM_THREADS), VM_THREADS, 0, q->stream>>>(mats, (uint32_t)n, is_log, min_eigenvalue,
                                                                                 out);
        SPX_LAUNCH_CHECK();
    });
}

int spx_voxelmap_overlap_ratio(spx_voxelmap_t m, const float* points, size_t n, const float* sensor_pose16,
                               float* ratio) {  // :194-246
    return guard([&] {
        SPX_REQUIRE(m && ratio, "[VoxelHashMap::compute_overlap_ratio] null argument");
        *ratio = 0.0f;
        if (n == 0 || m->voxel_num == 0) return;
        SPX_REQUIRE(points, "[VoxelHashMap::compute_overlap_ratio] null points");
        SPX_REQUIRE(n < (1ull << 31), "[VoxelHashMap::compute_overlap_ratio] too many points");
        DeviceGuard dg(m->q->device);
        cudaStream_t st = m->q->stream;
        SPX_CUDA(cudaMemsetAsync(m->counter + 1, 0, 4, st));
        const Xform T = sensor_pose16 ? xform_from_colmajor(sensor_pose16) : xform_identity();
        vm_overlap_kernel<<<div_up(n, VM_THREADS), VM_THREADS, 0, st>>>(m->t, reinterpret_cast<const float4*>(points),
                                                                       (uint32_t)n, T, m->inv, m->min_num_point,
                                                                       m->counter + 1);
        SPX_LAUNCH_CHECK();
        *ratio = (float)vm_read_counter(m, 1) / (float)n;
    });
}

}  // extern "C"
