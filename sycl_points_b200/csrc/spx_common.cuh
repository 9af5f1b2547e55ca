// Shared internals of libspx: error plumbing, the queue (one in-order CUDA stream + a scratch
// arena), launch accounting and the exact-arithmetic device helpers every kernel uses.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/spx.h"

namespace spx {

// ------------------------------------------------------------------ errors
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

#define SPX_CUDA(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            throw ::spx::Error(SPX_ERR_CUDA, std::string("[CUDA] ") + cudaGetErrorString(_e) + " at " +     \
                                                 __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")"); \
    } while (0)

#define SPX_REQUIRE(cond, msg)                                              \
    do {                                                                    \
        if (!(cond)) throw ::spx::Error(SPX_ERR_INVALID_ARGUMENT, (msg));   \
    } while (0)

// Every extern "C" entry point is `return guard([&]{ ... });`
template <typename F>
inline int guard(F&& f) {
    try {
        f();
        return SPX_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        if (e.code == SPX_ERR_CUDA) (void)cudaGetLastError();  // reported once: do not leave it pending for the next launch check
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return SPX_ERR_INTERNAL;
    } catch (...) {
        set_last_error("unknown exception");
        return SPX_ERR_INTERNAL;
    }
}

extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Checks the launch itself (configuration errors); execution errors surface at the next sync.
#define SPX_LAUNCH_CHECK()            \
    do {                              \
        ::spx::count_launch();        \
        SPX_CUDA(cudaGetLastError()); \
    } while (0)

}  // namespace spx

// ------------------------------------------------------------------ queue
struct spx_queue_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    int sm_count = 148;
    // scratch arena: grows monotonically, bump-allocated per API call (no cudaMalloc on the hot path)
    char* arena = nullptr;
    size_t arena_cap = 0;
    size_t arena_off = 0;
    std::vector<void*> retired;  // old arenas kept alive until the stream is known idle
    // pinned staging for small host<->device scalars
    char* pinned = nullptr;
    size_t pinned_cap = 0;
    // optional blocking wait (cudaEventBlockingSync) instead of the default spin
    cudaEvent_t block_ev = nullptr;
    // voxel-grid key geometry of the previous down-sampling on this queue (spx_voxel.cu: lets the next
    // call of the same voxel size start its sort without waiting for its own bounding box)
    struct {
        bool valid = false;
        float voxel = 0.0f;
        int mn[3] = {0, 0, 0};
        int mx[3] = {0, 0, 0};
    } voxel_geom;
    // voxel box of the LAST down-sampled cloud on this queue (its own, not the joined guess): the metric bounding
    // box of its output follows from it without touching the points (spx_voxel_last_box -> hinted index build)
    struct {
        bool valid = false;
        float voxel = 0.0f;
        int mn[3] = {0, 0, 0};
        int mx[3] = {0, 0, 0};
    } voxel_last;

    // the host thread whose API call owns the arena right now (0: none yet).  A queue — its arena, its pinned staging
    // block, its voxel-grid state — is single-threaded by contract; a second thread that enters the library on the
    // same queue while a call is still carving its scratch is caught at the next take instead of silently handing
    // two calls the same memory (what an illegal address in a kernel much later would otherwise be the only sign of).
    unsigned long long arena_owner = 0;
    static unsigned long long this_thread_tag() {
        static std::atomic<unsigned long long> next{1};
        static thread_local unsigned long long tag = next.fetch_add(1, std::memory_order_relaxed);
        return tag;
    }
    void arena_reset() {
        arena_off = 0;
        arena_owner = this_thread_tag();
    }
    // Reserve the total a call needs BEFORE taking pointers: growing invalidates nothing in flight
    // (the old block is retired, not freed) but earlier pointers of this call would go stale.
    void arena_reserve(size_t bytes);
    void* arena_take(size_t bytes);
    template <typename T>
    T* take(size_t count) {
        return static_cast<T*>(arena_take(count * sizeof(T)));
    }
    void* pinned_get(size_t bytes);
    void sync();
};

struct spx_event_s {
    cudaEvent_t ev = nullptr;
};

// Multi-GPU exchange of the per-iteration sums row over NVLink peer memory.
// Every rank owns one mailbox (cudaMalloc, exported by CUDA IPC or mapped by peer access):
//   rows  [2 parities][SPX_MAX_RANKS][32] doubles — rank r's partial sums, written BY rank r
//   flags [2][SPX_MAX_RANKS] u64                  — sequence number of the row that has landed
//   gsum  [2][32] doubles, ready [2] u64          — the folded sums, published to the local grid
//   error u32                                     — set when a peer did not answer in time
constexpr int SPX_MAX_RANKS = 8;
constexpr size_t SPX_MBOX_ROWS = 0;
constexpr size_t SPX_MBOX_FLAGS = SPX_MBOX_ROWS + 2 * SPX_MAX_RANKS * 32 * sizeof(double);
constexpr size_t SPX_MBOX_GSUM = SPX_MBOX_FLAGS + 2 * SPX_MAX_RANKS * sizeof(unsigned long long);
constexpr size_t SPX_MBOX_READY = SPX_MBOX_GSUM + 2 * 32 * sizeof(double);
constexpr size_t SPX_MBOX_ERROR = SPX_MBOX_READY + 2 * sizeof(unsigned long long);
constexpr size_t SPX_MBOX_BYTES = 8192;

struct spx_comm_s {
    spx_queue_t q = nullptr;
    int rank = 0, world = 1;
    char* local = nullptr;                  // this rank's mailbox
    char* peer[SPX_MAX_RANKS] = {};         // every rank's mailbox as mapped here (peer[rank] == local)
    bool ipc_opened[SPX_MAX_RANKS] = {};    // mapped through cudaIpcOpenMemHandle (to be closed)
    bool connected = false;
    unsigned long long seq = 1;             // next row sequence number; advances identically on every rank
};

namespace spx {

// Handles (device arrays, indices, registrations, voxel maps) outlive their queue in garbage-collected callers
// (Python finalises the members of a reference cycle in no particular order): every path that releases memory
// checks that its queue still exists and falls back to a synchronous cudaFree when it does not.
bool queue_is_live(spx_queue_t q);

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        SPX_CUDA(cudaGetDevice(&prev));
        if (prev != dev) SPX_CUDA(cudaSetDevice(dev));
        cur = dev;
    }
    ~DeviceGuard() {
        if (prev != cur) cudaSetDevice(prev);
    }
    int cur = 0;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int div_up(size_t a, size_t b) { return (int)((a + b - 1) / b); }

// 4x4 transform passed to kernels by value, row-major rows as float4 (the reference hands its
// kernels std::array<sycl::float4,4> rows: eigen_utils.hpp:701-708)
struct Xform {
    float4 r0, r1, r2, r3;
};
inline Xform xform_from_colmajor(const float* T) {
    Xform x;
    x.r0 = make_float4(T[0], T[4], T[8], T[12]);
    x.r1 = make_float4(T[1], T[5], T[9], T[13]);
    x.r2 = make_float4(T[2], T[6], T[10], T[14]);
    x.r3 = make_float4(T[3], T[7], T[11], T[15]);
    return x;
}
inline Xform xform_identity() {
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    return xform_from_colmajor(I);
}

#ifdef __CUDACC__
// transform_point: row i = fma(T(i,3),p3, fma(T(i,2),p2, fma(T(i,1),p1, fma(T(i,0),p0, 0))))
// (I/algorithms/common/transform.hpp:32-37 via eigen_utils.hpp:113-127).  __fmaf_rn/__fmul_rn are
// never re-associated or contracted by nvcc, so the result is bit-identical to the fma chain.
__device__ __forceinline__ float row_dot(const float4 r, const float4 p) {
    return __fmaf_rn(r.w, p.w, __fmaf_rn(r.z, p.z, __fmaf_rn(r.y, p.y, __fmul_rn(r.x, p.x))));
}
__device__ __forceinline__ float4 transform_point(const Xform& T, const float4 p) {
    return make_float4(row_dot(T.r0, p), row_dot(T.r1, p), row_dot(T.r2, p), row_dot(T.r3, p));
}
// squared distance exactly as kdtree.hpp:509-511 evaluates dot<4>(q-p, q-p) with w-difference 0:
// fma(dz,dz,fma(dy,dy,dx*dx)).  (The 4th term fma(0,0,s) == s.)
__device__ __forceinline__ float dist_sq(const float qx, const float qy, const float qz, const float px,
                                         const float py, const float pz) {
    const float dx = __fsub_rn(qx, px);
    const float dy = __fsub_rn(qy, py);
    const float dz = __fsub_rn(qz, pz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
// (dist, idx) lexicographic "a before b"; idx < 0 marks an empty slot (always last)
__device__ __forceinline__ bool lex_less(float da, int ia, float db, int ib) {
    return da < db || (da == db && (ib < 0 || ia < ib));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace spx
