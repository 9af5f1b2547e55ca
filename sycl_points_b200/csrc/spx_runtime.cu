// Runtime shim of libspx: replaces the reference's SYCL queue / USM / event layer
// (I/utils/sycl_utils.hpp:234-280,491-635) with one in-order CUDA stream per queue, explicit
// device memory and CUDA events.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <queue>
#include <random>
#include <unordered_map>
#include <unordered_set>

#include <cuda_profiler_api.h>

#include "spx_common.cuh"

namespace spx {

static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }
std::atomic<uint64_t> g_launches{0};

}  // namespace spx

using namespace spx;

void spx_queue_s::arena_reserve(size_t bytes) {
    bytes = align_up(bytes + 4096, 1 << 20);
    if (bytes <= arena_cap) return;
    size_t cap = arena_cap ? arena_cap : (size_t)(8 << 20);
    while (cap < bytes) cap *= 2;
    void* p = nullptr;
    SPX_CUDA(cudaMalloc(&p, cap));
    if (arena) retired.push_back(arena);
    arena = static_cast<char*>(p);
    arena_cap = cap;
    arena_off = 0;
}

void* spx_queue_s::arena_take(size_t bytes) {
    if (arena_owner != 0 && arena_owner != this_thread_tag())
        throw Error(SPX_ERR_INVALID_ARGUMENT,
                    "[spx] this queue is being used by two host threads at once (a queue and everything created on it "
                    "is single-threaded: drive each queue from one thread at a time)");
    const size_t off = align_up(arena_off, 256);
    if (off + bytes > arena_cap)
        throw Error(SPX_ERR_INTERNAL, "[spx] scratch arena overflow (arena_reserve under-estimated)");
    arena_off = off + bytes;
    return arena + off;
}

void* spx_queue_s::pinned_get(size_t bytes) {
    if (bytes < (64u << 10)) bytes = 64u << 10;  // one allocation serves every small staging use
    if (bytes > pinned_cap) {
        if (pinned) {
            cudaStreamSynchronize(stream);  // nothing may still be copying through the old block
            cudaFreeHost(pinned);
        }
        pinned = nullptr;
        pinned_cap = 0;
        void* p = nullptr;
        SPX_CUDA(cudaMallocHost(&p, align_up(bytes, 4096)));
        pinned = static_cast<char*>(p);
        pinned_cap = align_up(bytes, 4096);
    }
    return pinned;
}

void spx_queue_s::sync() {
    if (block_ev) {
        SPX_CUDA(cudaEventRecord(block_ev, stream));
        SPX_CUDA(cudaEventSynchronize(block_ev));
    } else {
        SPX_CUDA(cudaStreamSynchronize(stream));
    }
    for (void* p : retired) cudaFree(p);
    retired.clear();
}

namespace spx {
namespace {
std::mutex g_queues_mu;
std::unordered_set<spx_queue_t> g_queues;
}  // namespace
bool queue_is_live(spx_queue_t q) {
    if (!q) return false;
    std::lock_guard<std::mutex> lk(g_queues_mu);
    return g_queues.count(q) != 0;
}
static void queue_register(spx_queue_t q, bool live) {
    std::lock_guard<std::mutex> lk(g_queues_mu);
    if (live) g_queues.insert(q);
    else g_queues.erase(q);
}
}  // namespace spx

extern "C" {

const char* spx_last_error(void) { return t_last_error.c_str(); }
int spx_abi_version(void) { return SPX_ABI_VERSION; }
uint64_t spx_kernel_launch_count(void) { return g_launches.load(); }

int spx_profiler_range(int start) {
    return guard([&] {
        if (start) SPX_CUDA(cudaProfilerStart());
        else SPX_CUDA(cudaProfilerStop());
    });
}

int spx_device_count(int* count) {
    return guard([&] {
        SPX_REQUIRE(count, "[spx_device_count] null output");
        SPX_CUDA(cudaGetDeviceCount(count));
    });
}

int spx_device_info(int device, char* name, int* sm, int* sm_count, size_t* global_mem_bytes, int* l2_bytes) {
    return guard([&] {
        cudaDeviceProp p;
        SPX_CUDA(cudaGetDeviceProperties(&p, device));
        if (name) {
            std::strncpy(name, p.name, 255);
            name[255] = 0;
        }
        if (sm) *sm = p.major * 10 + p.minor;
        if (sm_count) *sm_count = p.multiProcessorCount;
        if (global_mem_bytes) *global_mem_bytes = p.totalGlobalMem;
        if (l2_bytes) *l2_bytes = p.l2CacheSize;
    });
}

static int queue_create(int device, void* stream, bool own, spx_queue_t* out, int priority = 0) {
    return guard([&] {
        SPX_REQUIRE(out, "[DeviceQueue::DeviceQueue] null output");
        int n = 0;
        SPX_CUDA(cudaGetDeviceCount(&n));
        SPX_REQUIRE(device >= 0 && device < n,
                    "[DeviceQueue::DeviceQueue] device ordinal " + std::to_string(device) + " is not supported.");
        DeviceGuard g(device);
        auto* q = new spx_queue_s();
        q->device = device;
        q->owns_stream = own;
        if (own) {
            int lo = 0, hi = 0;  // numerically: hi <= lo, lower = more urgent
            SPX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            const int pr = priority > 0 ? hi : (priority < 0 ? lo : 0);
            SPX_CUDA(cudaStreamCreateWithPriority(&q->stream, cudaStreamNonBlocking, pr));
        } else {
            q->stream = static_cast<cudaStream_t>(stream);
        }
        SPX_CUDA(cudaDeviceGetAttribute(&q->sm_count, cudaDevAttrMultiProcessorCount, device));
        // spx_malloc / spx_free are stream-ordered pool allocations: keep freed blocks cached so a
        // per-scan alloc/free cycle never reaches the driver
        cudaMemPool_t pool;
        SPX_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;
        SPX_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        // ... and start the pool with a cushion (once per device): a pool that has to grow in the middle of a
        // stream of scans — whenever a block freed on one queue is not yet reusable on another — stalls the
        // allocating host thread for 3-30 ms per growth (measured: profiles/r2_pool_growth_stalls.txt).
        // SPX_POOL_RESERVE_MB overrides the 2 GiB default; 0 leaves the pool to grow on demand.
        {
            static std::mutex mu;
            static bool grown[64] = {};
            std::lock_guard<std::mutex> lk(mu);
            if (device < 64 && !grown[device]) {
                grown[device] = true;
                const char* e = std::getenv("SPX_POOL_RESERVE_MB");
                const size_t mb = e ? (size_t)std::strtoull(e, nullptr, 10) : 2048;
                size_t free_b = 0, total_b = 0;
                if (mb && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (mb << 20) < free_b / 2) {
                    void* cushion = nullptr;
                    if (cudaMallocAsync(&cushion, mb << 20, q->stream) == cudaSuccess) {
                        cudaFreeAsync(cushion, q->stream);
                        cudaStreamSynchronize(q->stream);
                    } else {
                        cudaGetLastError();
                    }
                }
            }
        }
        q->arena_reserve(8 << 20);
        queue_register(q, true);
        *out = q;
    });
}

int spx_queue_create(int device, spx_queue_t* out) { return queue_create(device, nullptr, true, out); }
int spx_queue_create_with_priority(int device, int priority, spx_queue_t* out) {
    return queue_create(device, nullptr, true, out, priority);
}
int spx_queue_create_on_stream(int device, void* cuda_stream, spx_queue_t* out) {
    return queue_create(device, cuda_stream, false, out);
}

int spx_queue_destroy(spx_queue_t q) {
    return guard([&] {
        if (!q || !queue_is_live(q)) return;  // (a second destroy of the same handle is ignored)
        queue_register(q, false);
        DeviceGuard g(q->device);
        cudaStreamSynchronize(q->stream);
        for (void* p : q->retired) cudaFree(p);
        if (q->arena) cudaFree(q->arena);
        if (q->pinned) cudaFreeHost(q->pinned);
        if (q->block_ev) cudaEventDestroy(q->block_ev);
        if (q->owns_stream && q->stream) cudaStreamDestroy(q->stream);
        delete q;
    });
}

int spx_queue_sync(spx_queue_t q) {
    return guard([&] {
        SPX_REQUIRE(q, "[spx_queue_sync] null queue");
        DeviceGuard g(q->device);
        q->sync();
    });
}

int spx_queue_set_blocking_sync(spx_queue_t q, int blocking) {
    return guard([&] {
        SPX_REQUIRE(q, "[spx_queue_set_blocking_sync] null queue");
        DeviceGuard g(q->device);
        if (blocking && !q->block_ev) SPX_CUDA(cudaEventCreateWithFlags(&q->block_ev, cudaEventBlockingSync | cudaEventDisableTiming));
        if (!blocking && q->block_ev) {
            cudaEventDestroy(q->block_ev);
            q->block_ev = nullptr;
        }
    });
}

int spx_queue_device(spx_queue_t q, int* device) {
    return guard([&] {
        SPX_REQUIRE(q && device, "[spx_queue_device] null argument");
        *device = q->device;
    });
}

int spx_malloc(spx_queue_t q, size_t bytes, void** out) {
    return guard([&] {
        SPX_REQUIRE(q && out, "[spx_malloc] null argument");
        DeviceGuard g(q->device);
        *out = nullptr;
        if (bytes == 0) return;
        SPX_CUDA(cudaMallocAsync(out, bytes, q->stream));
    });
}

int spx_free(spx_queue_t q, void* ptr) {
    return guard([&] {
        if (!ptr) return;
        if (!queue_is_live(q)) {  // the queue is gone (or NULL): nothing to order against, free synchronously
            SPX_CUDA(cudaFree(ptr));
            return;
        }
        DeviceGuard g(q->device);
        SPX_CUDA(cudaFreeAsync(ptr, q->stream));  // ordered after everything already enqueued
    });
}

int spx_malloc_host(size_t bytes, void** out) {
    return guard([&] {
        SPX_REQUIRE(out, "[spx_malloc_host] null output");
        *out = nullptr;
        if (bytes == 0) return;
        SPX_CUDA(cudaMallocHost(out, bytes));
    });
}

int spx_free_host(void* ptr) {
    return guard([&] {
        if (ptr) SPX_CUDA(cudaFreeHost(ptr));
    });
}

int spx_memcpy_h2d(spx_queue_t q, void* dst, const void* src_host, size_t bytes) {
    return guard([&] {
        SPX_REQUIRE(q, "[spx_memcpy_h2d] null queue");
        if (!bytes) return;
        DeviceGuard g(q->device);
        SPX_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, q->stream));
    });
}

int spx_memcpy_d2h(spx_queue_t q, void* dst_host, const void* src, size_t bytes) {
    return guard([&] {
        SPX_REQUIRE(q, "[spx_memcpy_d2h] null queue");
        if (!bytes) return;
        DeviceGuard g(q->device);
        SPX_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, q->stream));
    });
}

int spx_memcpy_d2d(spx_queue_t q, void* dst, const void* src, size_t bytes) {
    return guard([&] {
        SPX_REQUIRE(q, "[spx_memcpy_d2d] null queue");
        if (!bytes) return;
        DeviceGuard g(q->device);
        SPX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, q->stream));
    });
}

int spx_memset(spx_queue_t q, void* dst, int value, size_t bytes) {
    return guard([&] {
        SPX_REQUIRE(q, "[spx_memset] null queue");
        if (!bytes) return;
        DeviceGuard g(q->device);
        SPX_CUDA(cudaMemsetAsync(dst, value, bytes, q->stream));
    });
}

// Managed blocks are recycled through size-class free lists: cudaMallocManaged / cudaFree cost
// 100+ us each and cudaFree synchronises the whole device, while the C++ facade's shared_vector
// allocates and releases result arrays on every call (as the reference does with USM).
namespace {
std::mutex g_managed_mu;
std::unordered_map<size_t, std::vector<void*>> g_managed_free;  // size class -> idle blocks
std::unordered_map<void*, size_t> g_managed_size;               // live or idle block -> size class
size_t managed_class(size_t bytes) {
    size_t c = 4096;
    while (c < bytes) c <<= 1;
    return c;
}
}  // namespace

int spx_malloc_managed(size_t bytes, void** out) {
    return guard([&] {
        SPX_REQUIRE(out, "[spx_malloc_managed] null output");
        *out = nullptr;
        if (bytes == 0) return;
        const size_t cls = managed_class(bytes);
        {
            std::lock_guard<std::mutex> lk(g_managed_mu);
            auto it = g_managed_free.find(cls);
            if (it != g_managed_free.end() && !it->second.empty()) {
                *out = it->second.back();
                it->second.pop_back();
                return;
            }
        }
        void* p = nullptr;
        SPX_CUDA(cudaMallocManaged(&p, cls, cudaMemAttachGlobal));
        std::lock_guard<std::mutex> lk(g_managed_mu);
        g_managed_size[p] = cls;
        *out = p;
    });
}

int spx_free_managed(void* ptr) {
    return guard([&] {
        if (!ptr) return;
        std::lock_guard<std::mutex> lk(g_managed_mu);
        auto it = g_managed_size.find(ptr);
        if (it == g_managed_size.end()) {  // not ours: release it the plain way
            SPX_CUDA(cudaFree(ptr));
            return;
        }
        g_managed_free[it->second].push_back(ptr);
    });
}

int spx_managed_trim(void) {
    return guard([&] {
        std::lock_guard<std::mutex> lk(g_managed_mu);
        for (auto& kv : g_managed_free) {
            for (void* p : kv.second) {
                g_managed_size.erase(p);
                cudaFree(p);
            }
            kv.second.clear();
        }
    });
}

int spx_prefetch(spx_queue_t q, const void* ptr, size_t bytes, int to_device) {
    return guard([&] {
        SPX_REQUIRE(q, "[spx_prefetch] null queue");
        if (!ptr || !bytes) return;
        static const bool disabled = std::getenv("SPX_NO_PREFETCH") != nullptr;  // tuning aid
        // cudaMemPrefetchAsync costs 50-100 us per call on this platform whatever the size (measured
        // through the facade's example: 5.1 ms per loop with it, 1.8 ms without, on 6 k-point clouds);
        // below a few MB on-demand migration is cheaper, so only bulk arrays are prefetched
        if (disabled || bytes < ((size_t)4 << 20)) return;
        DeviceGuard g(q->device);
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess || at.type != cudaMemoryTypeManaged) {
            (void)cudaGetLastError();
            return;  // plain device / host memory: nothing to migrate
        }
        SPX_CUDA(cudaMemPrefetchAsync(ptr, bytes, to_device ? q->device : cudaCpuDeviceId, q->stream));
    });
}

int spx_event_create(spx_event_t* out) {
    return guard([&] {
        SPX_REQUIRE(out, "[spx_event_create] null output");
        auto* e = new spx_event_s();
        SPX_CUDA(cudaEventCreate(&e->ev));
        *out = e;
    });
}

int spx_event_destroy(spx_event_t e) {
    return guard([&] {
        if (!e) return;
        cudaEventDestroy(e->ev);
        delete e;
    });
}

int spx_event_record(spx_queue_t q, spx_event_t e) {
    return guard([&] {
        SPX_REQUIRE(q && e, "[spx_event_record] null argument");
        DeviceGuard g(q->device);
        SPX_CUDA(cudaEventRecord(e->ev, q->stream));
    });
}

int spx_queue_wait_event(spx_queue_t q, spx_event_t e) {
    return guard([&] {
        SPX_REQUIRE(q && e, "[spx_queue_wait_event] null argument");
        DeviceGuard g(q->device);
        SPX_CUDA(cudaStreamWaitEvent(q->stream, e->ev, 0));
    });
}

int spx_event_elapsed_ms(spx_event_t start, spx_event_t stop, float* ms) {
    return guard([&] {
        SPX_REQUIRE(start && stop && ms, "[spx_event_elapsed_ms] null argument");
        SPX_CUDA(cudaEventSynchronize(stop->ev));
        SPX_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
    });
}

// ------------------------------------------------------------------ random sampling (host RNG, like the reference)
struct spx_rng_s {
    std::mt19937 mt;
};

int spx_rng_create(uint32_t seed, spx_rng_t* out) {
    return guard([&] {
        SPX_REQUIRE(out, "[spx_rng_create] null output");
        auto* r = new spx_rng_s();
        r->mt.seed(seed);
        *out = r;
    });
}
int spx_rng_seed(spx_rng_t rng, uint32_t seed) {
    return guard([&] {
        SPX_REQUIRE(rng, "[spx_rng_seed] null rng");
        rng->mt.seed(seed);
    });
}
int spx_rng_destroy(spx_rng_t rng) {
    return guard([&] { delete rng; });
}

// random_sampling_operator.hpp:24-52: the draw order (and therefore the RNG stream carried over to
// the next call) is the reference's; the compaction keeps source order, so ascending indices.
int spx_random_sampling(spx_queue_t q, spx_rng_t rng, size_t n, size_t sampling_num, int32_t* idx_out, size_t* m_host) {
    return guard([&] {
        SPX_REQUIRE(q && rng && m_host, "[PreprocessFilter::random_sampling] null argument");
        SPX_REQUIRE(n < (1ull << 31), "[PreprocessFilter::random_sampling] too many points");
        std::vector<int32_t> sel;
        if (n <= sampling_num) {  // keep everything, no draw
            sel.resize(n);
            std::iota(sel.begin(), sel.end(), 0);
        } else {
            std::vector<size_t> indices(n);
            std::iota(indices.begin(), indices.end(), (size_t)0);
            for (size_t i = 0; i < sampling_num; ++i) {
                std::uniform_int_distribution<size_t> dist(i, n - 1);
                const size_t j = dist(rng->mt);
                std::swap(indices[i], indices[j]);
            }
            sel.resize(sampling_num);
            for (size_t i = 0; i < sampling_num; ++i) sel[i] = (int32_t)indices[i];
            std::sort(sel.begin(), sel.end());
        }
        *m_host = sel.size();
        if (sel.empty()) return;
        SPX_REQUIRE(idx_out, "[PreprocessFilter::random_sampling] null output");
        DeviceGuard g(q->device);
        int32_t* pin = static_cast<int32_t*>(q->pinned_get(sel.size() * sizeof(int32_t)));
        std::memcpy(pin, sel.data(), sel.size() * sizeof(int32_t));
        SPX_CUDA(cudaMemcpyAsync(idx_out, pin, sel.size() * sizeof(int32_t), cudaMemcpyDefault, q->stream));
        q->sync();  // the pinned staging block is reused by the next call
    });
}

// mixed_random_sampling_operator.hpp:29-107: the first floor(sampling_num * weighted_ratio) points by weighted
// reservoir keys log(u) / w (u ~ uniform_real_distribution<float>(FLT_MIN, 1), one draw per positive weight, the
// smallest key evicted first: std::priority_queue with std::greater on (key, index)), the rest by a partial
// Fisher-Yates over the not yet selected indices; the persistent mt19937 advances exactly as in the reference.
int spx_mixed_random_sampling(spx_queue_t q, spx_rng_t rng, const float* weights, size_t n, size_t sampling_num,
                              float weighted_ratio, int32_t* idx_out, size_t* m_host) {
    return guard([&] {
        SPX_REQUIRE(q && rng && m_host, "[PreprocessFilter::mixed_random_sampling] null argument");
        SPX_REQUIRE(n < (1ull << 31), "[PreprocessFilter::mixed_random_sampling] too many points");
        DeviceGuard g(q->device);
        std::vector<int32_t> sel;
        if (n <= sampling_num) {
            sel.resize(n);
            std::iota(sel.begin(), sel.end(), 0);
        } else {
            SPX_REQUIRE(weights, "[PreprocessFilter::mixed_random_sampling] weights size must match points");
            SPX_REQUIRE(std::isfinite(weighted_ratio) && weighted_ratio >= 0.0f && weighted_ratio <= 1.0f,
                        "[PreprocessFilter::mixed_random_sampling] weighted_ratio must be within [0.0, 1.0]");
            std::vector<float> w(n);
            SPX_CUDA(cudaMemcpyAsync(w.data(), weights, n * sizeof(float), cudaMemcpyDefault, q->stream));
            q->sync();
            const size_t weighted_target = static_cast<size_t>(std::floor(static_cast<double>(sampling_num) * weighted_ratio));
            std::vector<uint8_t> flags(n, 0);
            using KeyIndexPair = std::pair<float, size_t>;
            std::priority_queue<KeyIndexPair, std::vector<KeyIndexPair>, std::greater<KeyIndexPair>> selected;
            std::uniform_real_distribution<float> weighted_dist(std::numeric_limits<float>::min(), 1.0f);
            for (size_t i = 0; i < n; ++i) {
                const float weight = w[i];
                SPX_REQUIRE(std::isfinite(weight) && weight >= 0.0f,
                            "[PreprocessFilter::mixed_random_sampling] weights must be finite and non-negative");
                if (weight <= 0.0f || weighted_target == 0) continue;
                const float key = std::log(weighted_dist(rng->mt)) / weight;
                if (selected.size() < weighted_target) {
                    selected.emplace(key, i);
                    continue;
                }
                if (!selected.empty() && selected.top().first < key) {
                    selected.pop();
                    selected.emplace(key, i);
                }
            }
            while (!selected.empty()) {
                flags[selected.top().second] = 1;
                selected.pop();
            }
            std::vector<size_t> remaining;
            remaining.reserve(n);
            size_t selected_count = 0;
            for (size_t i = 0; i < n; ++i) {
                if (flags[i]) ++selected_count;
                else remaining.push_back(i);
            }
            const size_t uniform_target = std::min(sampling_num - selected_count, remaining.size());
            for (size_t i = 0; i < uniform_target; ++i) {
                std::uniform_int_distribution<size_t> uniform_dist(i, remaining.size() - 1);
                const size_t j = uniform_dist(rng->mt);
                std::swap(remaining[i], remaining[j]);
                flags[remaining[i]] = 1;
            }
            for (size_t i = 0; i < n; ++i)
                if (flags[i]) sel.push_back((int32_t)i);
        }
        *m_host = sel.size();
        if (sel.empty()) return;
        SPX_REQUIRE(idx_out, "[PreprocessFilter::mixed_random_sampling] null output");
        int32_t* pin = static_cast<int32_t*>(q->pinned_get(sel.size() * sizeof(int32_t)));
        std::memcpy(pin, sel.data(), sel.size() * sizeof(int32_t));
        SPX_CUDA(cudaMemcpyAsync(idx_out, pin, sel.size() * sizeof(int32_t), cudaMemcpyDefault, q->stream));
        q->sync();
    });
}

// weighted_sampling_operator.hpp:29-96: sampling_num points by weighted reservoir keys log(u) / w over the points of
// positive weight (std::priority_queue with std::greater on (key, index), the smallest key evicted first).
int spx_weighted_random_sampling(spx_queue_t q, spx_rng_t rng, const float* weights, size_t n, size_t sampling_num,
                                 int32_t* idx_out, size_t* m_host) {
    return guard([&] {
        SPX_REQUIRE(q && rng && m_host, "[PreprocessFilter::weighted_random_sampling] null argument");
        SPX_REQUIRE(n < (1ull << 31), "[PreprocessFilter::weighted_random_sampling] too many points");
        DeviceGuard g(q->device);
        std::vector<int32_t> sel;
        if (n <= sampling_num) {
            sel.resize(n);
            std::iota(sel.begin(), sel.end(), 0);
        } else {
            SPX_REQUIRE(weights, "[PreprocessFilter::weighted_random_sampling] weights size must match points");
            std::vector<float> w(n);
            SPX_CUDA(cudaMemcpyAsync(w.data(), weights, n * sizeof(float), cudaMemcpyDefault, q->stream));
            q->sync();
            size_t positive = 0;
            for (size_t i = 0; i < n; ++i) {
                SPX_REQUIRE(std::isfinite(w[i]) && w[i] >= 0.0f,
                            "[PreprocessFilter::weighted_random_sampling] weights must be finite and non-negative");
                if (w[i] > 0.0f) ++positive;
            }
            SPX_REQUIRE(positive > 0, "[PreprocessFilter::weighted_random_sampling] at least one weight must be positive");
            SPX_REQUIRE(sampling_num <= positive,
                        "[PreprocessFilter::weighted_random_sampling] sampling_num exceeds positive-weight points");
            using KeyIndexPair = std::pair<float, size_t>;
            std::priority_queue<KeyIndexPair, std::vector<KeyIndexPair>, std::greater<KeyIndexPair>> selected;
            std::uniform_real_distribution<float> dist(std::numeric_limits<float>::min(), 1.0f);
            for (size_t i = 0; i < n; ++i) {
                if (w[i] <= 0.0f) continue;
                const float key = std::log(dist(rng->mt)) / w[i];
                if (selected.size() < sampling_num) {
                    selected.emplace(key, i);
                    continue;
                }
                if (selected.top().first < key) {
                    selected.pop();
                    selected.emplace(key, i);
                }
            }
            while (!selected.empty()) {
                sel.push_back((int32_t)selected.top().second);
                selected.pop();
            }
            std::sort(sel.begin(), sel.end());
        }
        *m_host = sel.size();
        if (sel.empty()) return;
        SPX_REQUIRE(idx_out, "[PreprocessFilter::weighted_random_sampling] null output");
        int32_t* pin = static_cast<int32_t*>(q->pinned_get(sel.size() * sizeof(int32_t)));
        std::memcpy(pin, sel.data(), sel.size() * sizeof(int32_t));
        SPX_CUDA(cudaMemcpyAsync(idx_out, pin, sel.size() * sizeof(int32_t), cudaMemcpyDefault, q->stream));
        q->sync();
    });
}

// the first point of farthest_point_sampling: uniform_int_distribution<size_t>(0, N - 1) on the operator's mt19937
// (farthest_point_sampling_operator.hpp:51-53)
int spx_rng_uniform_index(spx_rng_t rng, size_t n, size_t* out) {
    return guard([&] {
        SPX_REQUIRE(rng && out && n > 0, "[spx_rng_uniform_index] bad argument");
        std::uniform_int_distribution<size_t> dist(0, n - 1);
        *out = dist(rng->mt);
    });
}

// ------------------------------------------------------------------ multi-GPU mailboxes
int spx_comm_create(spx_queue_t q, int rank, int world, spx_comm_t* out) {
    return guard([&] {
        SPX_REQUIRE(q && out, "[spx_comm_create] null argument");
        SPX_REQUIRE(world >= 1 && world <= SPX_MAX_RANKS && rank >= 0 && rank < world,
                    "[spx_comm_create] need 0 <= rank < world <= 8");
        DeviceGuard g(q->device);
        auto* c = new spx_comm_s();
        c->q = q;
        c->rank = rank;
        c->world = world;
        void* p = nullptr;
        // plain cudaMalloc: pool (cudaMallocAsync) memory cannot be exported through CUDA IPC
        cudaError_t e = cudaMalloc(&p, SPX_MBOX_BYTES);
        if (e != cudaSuccess) {
            delete c;
            throw Error(SPX_ERR_CUDA, std::string("[spx_comm_create] cudaMalloc: ") + cudaGetErrorString(e));
        }
        c->local = static_cast<char*>(p);
        c->peer[rank] = c->local;
        SPX_CUDA(cudaMemsetAsync(p, 0, SPX_MBOX_BYTES, q->stream));
        SPX_CUDA(cudaStreamSynchronize(q->stream));
        c->connected = (world == 1);
        *out = c;
    });
}

int spx_comm_destroy(spx_comm_t c) {
    return guard([&] {
        if (!c) return;
        if (queue_is_live(c->q)) {
            DeviceGuard g(c->q->device);
            cudaStreamSynchronize(c->q->stream);
        } else {
            cudaDeviceSynchronize();
        }
        for (int r = 0; r < c->world; ++r)
            if (c->ipc_opened[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
        if (c->local) cudaFree(c->local);
        delete c;
    });
}

int spx_comm_ipc_handle(spx_comm_t c, uint8_t* handle64_host) {
    return guard([&] {
        SPX_REQUIRE(c && handle64_host, "[spx_comm_ipc_handle] null argument");
        static_assert(sizeof(cudaIpcMemHandle_t) == SPX_IPC_HANDLE_BYTES, "CUDA IPC handle size");
        DeviceGuard g(c->q->device);
        cudaIpcMemHandle_t h;
        SPX_CUDA(cudaIpcGetMemHandle(&h, c->local));
        std::memcpy(handle64_host, &h, sizeof(h));
    });
}

int spx_comm_connect_ipc(spx_comm_t c, const uint8_t* handles_host) {
    return guard([&] {
        SPX_REQUIRE(c && handles_host, "[spx_comm_connect_ipc] null argument");
        DeviceGuard g(c->q->device);
        for (int r = 0; r < c->world; ++r) {
            if (r == c->rank || c->peer[r]) continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles_host + (size_t)r * SPX_IPC_HANDLE_BYTES, sizeof(h));
            void* p = nullptr;
            SPX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            c->peer[r] = static_cast<char*>(p);
            c->ipc_opened[r] = true;
        }
        c->connected = true;
    });
}

int spx_comm_connect_local(spx_comm_t* comms, int world) {
    return guard([&] {
        SPX_REQUIRE(comms && world >= 1 && world <= SPX_MAX_RANKS, "[spx_comm_connect_local] bad arguments");
        for (int a = 0; a < world; ++a) {
            SPX_REQUIRE(comms[a] && comms[a]->world == world && comms[a]->rank == a,
                        "[spx_comm_connect_local] communicator a must have rank a of `world`");
        }
        for (int a = 0; a < world; ++a) {
            DeviceGuard g(comms[a]->q->device);
            for (int b = 0; b < world; ++b) {
                if (a == b) continue;
                const int da = comms[a]->q->device, db = comms[b]->q->device;
                if (da != db) {
                    int can = 0;
                    SPX_CUDA(cudaDeviceCanAccessPeer(&can, da, db));
                    SPX_REQUIRE(can, "[spx_comm_connect_local] devices " + std::to_string(da) + " and " +
                                         std::to_string(db) + " have no peer access");
                    const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SPX_CUDA(e);
                    (void)cudaGetLastError();
                }
                comms[a]->peer[b] = comms[b]->local;
            }
            comms[a]->connected = true;
        }
    });
}

}  // extern "C"
