// Batched odometry front end (BASELINE config 5, SURVEY.md §8(b) `spx_align_batch`): P independent scan pairs
// from RAW clouds to registration results in one call.
//
//   per cloud (2P of them, identical (pointer, size) inputs processed once):
//       VoxelGrid::downsampling -> KDTree::build -> knn_search(k) -> covariance::estimate
//       (voxel_downsampling.hpp:50-79, kdtree.hpp:165-224, covariance.hpp:260-311)
//     — the clouds are independent, so they are spread over `lanes` internal queues (one CUDA stream + scratch
//       arena each), every lane driven by its own host thread: the kernels of this stage are latency-bound at
//       60 k points and overlap on the device, and a lane's host round trips (voxel count, grid plan) only
//       stall that lane;
//   all pairs together:
//       Registration::align (registration.hpp:201-276) as ONE set-up launch + ONE persistent cooperative launch
//       (align_batch_kernel, spx_registration.cu), a single device-to-host copy of the P results.
//
// Results are bit for bit those of the single-pair entry points on the same inputs.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

#include "spx_common.cuh"

using namespace spx;

struct spx_batch_s {
    spx_queue_t q = nullptr;
    spx_registration_t reg = nullptr;
    std::vector<spx_queue_t> lanes;
    float voxel = 0.25f;
    int k = 10;
};

namespace {

struct CloudJob {
    const float* raw = nullptr;
    size_t n_in = 0;
    // produced
    float* pts = nullptr;
    float* covs = nullptr;
    size_t m = 0;
    spx_index_t index = nullptr;
    spx_queue_t lane = nullptr;
    int rc = SPX_OK;
    std::string err;
};

void run_cloud(spx_batch_t b, spx_queue_t lane, CloudJob& c) {
    auto fail = [&](int rc) {
        c.rc = rc;
        c.err = spx_last_error();
    };
    c.lane = lane;
    if (c.n_in == 0) return;
    int rc;
    void* p = nullptr;
    if ((rc = spx_malloc(lane, c.n_in * 16, &p)) != SPX_OK) return fail(rc);
    c.pts = static_cast<float*>(p);
    if ((rc = spx_voxel_downsample(lane, c.raw, c.n_in, b->voxel, 1, c.pts, &c.m)) != SPX_OK) return fail(rc);
    // the voxel grid already knows a box around its output: the index build skips its own bounding-box / occupancy pass
    // and host round trip (cells of 1.85 / 2.4 voxels: ~3 / ~5 points per occupied cell on LiDAR surfaces)
    float lo[3], hi[3];
    static const bool hints = !(std::getenv("SPX_INDEX_HINT") && std::getenv("SPX_INDEX_HINT")[0] == '0');
    if (hints && c.m > 0 && spx_voxel_last_box(lane, lo, hi, nullptr) == SPX_OK) {
        if ((rc = spx_index_build_hinted(lane, c.pts, c.m, lo, hi, 1.85f * b->voxel, 2.4f * b->voxel, &c.index)) != SPX_OK) return fail(rc);
    } else if ((rc = spx_index_build(lane, c.pts, c.m, 0.0f, &c.index)) != SPX_OK) {
        return fail(rc);
    }
    if (c.m == 0) return;
    void *idx = nullptr, *dist = nullptr, *cov = nullptr;
    if ((rc = spx_malloc(lane, c.m * (size_t)b->k * 4, &idx)) != SPX_OK) return fail(rc);
    if ((rc = spx_malloc(lane, c.m * (size_t)b->k * 4, &dist)) != SPX_OK) return fail(rc);
    if ((rc = spx_malloc(lane, c.m * 64, &cov)) != SPX_OK) return fail(rc);
    c.covs = static_cast<float*>(cov);
    if ((rc = spx_index_knn(c.index, c.pts, c.m, b->k, nullptr, static_cast<int32_t*>(idx), static_cast<float*>(dist))) != SPX_OK)
        return fail(rc);
    if ((rc = spx_covariance(lane, c.pts, c.m, static_cast<const int32_t*>(idx), b->k, c.covs)) != SPX_OK) return fail(rc);
    spx_free(lane, idx);  // stream-ordered: after the covariance kernel
    spx_free(lane, dist);
}

}  // namespace

extern "C" {

int spx_batch_create(spx_queue_t q, const spx_registration_params* params, float voxel_size, int k_correspondences, int lanes,
                     spx_batch_t* out) {
    return guard([&] {
        SPX_REQUIRE(q && out, "[spx_batch_create] null argument");
        if (!(voxel_size > 0.0f)) throw Error(SPX_ERR_INVALID_ARGUMENT, "voxel_size must be positive");
        SPX_REQUIRE(k_correspondences >= 1 && k_correspondences <= 128, "[spx_batch_create] k must be in [1, 128]");
        if (lanes <= 0) {
            lanes = 8;
            if (const char* e = std::getenv("SPX_BATCH_LANES")) lanes = std::max(1, std::atoi(e));
        }
        lanes = std::min(lanes, 32);
        std::unique_ptr<spx_batch_s> b(new spx_batch_s());
        b->q = q;
        b->voxel = voxel_size;
        b->k = k_correspondences;
        try {
            if (spx_registration_create(q, params, &b->reg) != SPX_OK) throw Error(SPX_ERR_INTERNAL, spx_last_error());
            const unsigned cores = std::max(1u, std::thread::hardware_concurrency());
            for (int l = 0; l < lanes; ++l) {
                spx_queue_t lq = nullptr;
                if (spx_queue_create(q->device, &lq) != SPX_OK) throw Error(SPX_ERR_INTERNAL, spx_last_error());
                b->lanes.push_back(lq);
                // more lanes than spare cores: wait on an OS primitive instead of spinning
                if ((unsigned)lanes * 2 > cores || std::getenv("SPX_BATCH_BLOCKING")) spx_queue_set_blocking_sync(lq, 1);
            }
        } catch (...) {
            for (spx_queue_t lq : b->lanes) spx_queue_destroy(lq);
            if (b->reg) spx_registration_destroy(b->reg);
            throw;
        }
        *out = b.release();
    });
}

int spx_batch_destroy(spx_batch_t b) {
    return guard([&] {
        if (!b) return;
        for (spx_queue_t lq : b->lanes) spx_queue_destroy(lq);
        if (b->reg) spx_registration_destroy(b->reg);
        delete b;
    });
}

int spx_batch_set_params(spx_batch_t b, const spx_registration_params* params) {
    return guard([&] {
        SPX_REQUIRE(b && params, "[spx_batch_set_params] null argument");
        if (spx_registration_set_params(b->reg, params) != SPX_OK) throw Error(SPX_ERR_INTERNAL, spx_last_error());
    });
}

int spx_align_batch(spx_batch_t b, size_t n_pairs, const spx_scan_pair* pairs_host, spx_registration_result* results_host,
                    uint32_t* n_src_out, uint32_t* n_tgt_out) {
    return guard([&] {
        SPX_REQUIRE(b && (n_pairs == 0 || (pairs_host && results_host)), "[spx_align_batch] null argument");
        if (n_pairs == 0) return;
        DeviceGuard g(b->q->device);
        // distinct clouds (a target shared by several pairs — scans against one submap — is processed once)
        std::vector<CloudJob> jobs;
        jobs.reserve(2 * n_pairs);
        std::unordered_map<const float*, size_t> seen;
        std::vector<size_t> src_job(n_pairs), tgt_job(n_pairs);
        auto job_of = [&](const float* raw, size_t n) {
            auto it = seen.find(raw);
            if (it != seen.end() && jobs[it->second].n_in == n) return it->second;
            CloudJob c;
            c.raw = raw;
            c.n_in = n;
            jobs.push_back(c);
            seen[raw] = jobs.size() - 1;
            return jobs.size() - 1;
        };
        for (size_t p = 0; p < n_pairs; ++p) {
            SPX_REQUIRE((pairs_host[p].src_raw || pairs_host[p].n_src == 0) && (pairs_host[p].tgt_raw || pairs_host[p].n_tgt == 0),
                        "[spx_align_batch] null cloud");
            tgt_job[p] = job_of(pairs_host[p].tgt_raw, pairs_host[p].n_tgt);
            src_job[p] = job_of(pairs_host[p].src_raw, pairs_host[p].n_src);
        }
        // inputs may still be in flight on the caller's queue (uploads): the lanes start after it
        SPX_CUDA(cudaStreamSynchronize(b->q->stream));

        std::atomic<size_t> next{0};
        const size_t W = std::min(b->lanes.size(), jobs.size());
        auto worker = [&](size_t w) {
            cudaSetDevice(b->q->device);
            for (;;) {
                const size_t j = next.fetch_add(1);
                if (j >= jobs.size()) break;
                run_cloud(b, b->lanes[w], jobs[j]);
            }
            spx_queue_sync(b->lanes[w]);
        };
        std::vector<std::thread> threads;
        for (size_t w = 1; w < W; ++w) threads.emplace_back(worker, w);
        worker(0);
        for (auto& t : threads) t.join();

        auto cleanup = [&] {
            for (CloudJob& c : jobs) {
                if (c.index) spx_index_destroy(c.index);
                if (c.pts) spx_free(c.lane, c.pts);
                if (c.covs) spx_free(c.lane, c.covs);
            }
        };
        try {
            for (CloudJob& c : jobs)
                if (c.rc != SPX_OK) throw Error(c.rc, c.err);
            std::vector<spx_align_pair> ap(n_pairs);
            for (size_t p = 0; p < n_pairs; ++p) {
                const CloudJob &s = jobs[src_job[p]], &t = jobs[tgt_job[p]];
                spx_align_pair& a = ap[p];
                std::memset(&a, 0, sizeof(a));
                a.src_points = s.pts;
                a.src_covs = s.covs;
                a.ns = s.m;
                a.tgt_points = t.pts;
                a.tgt_covs = t.covs;
                a.nt = t.m;
                a.target_index = t.index;
                a.T_init_host = pairs_host[p].T_init_host;
                a.robust_scale = -1.0f;
                if (n_src_out) n_src_out[p] = (uint32_t)s.m;
                if (n_tgt_out) n_tgt_out[p] = (uint32_t)t.m;
            }
            if (spx_registration_align_batch(b->reg, n_pairs, ap.data(), results_host) != SPX_OK)
                throw Error(SPX_ERR_INTERNAL, spx_last_error());
        } catch (...) {
            cudaStreamSynchronize(b->q->stream);
            cleanup();
            throw;
        }
        cleanup();  // stream-ordered frees on the lanes; the align has synchronised
    });
}

int spx_batch_last_timing(spx_batch_t b, float* align_ms, int32_t* iterations) {
    return guard([&] {
        SPX_REQUIRE(b, "[spx_batch_last_timing] null handle");
        int32_t launches = 0;
        if (spx_registration_last_timing(b->reg, align_ms, &launches, iterations) != SPX_OK)
            throw Error(SPX_ERR_INTERNAL, spx_last_error());
    });
}

}  // extern "C"
