// Facade tests: the reference's own known-answer tests for the hot path, restated against
// include/sycl_points (no GTest in the image: a 20-line harness).  Run on a GPU box:
//   g++ -std=c++20 -O1 -Iinclude tests/cpp/test_facade.cpp -Lsycl_points_b200 -lspx -Wl,-rpath,... -o t && ./t
// Sources of the cases: T/test_kdtree.cpp:301-317,358-389,472-476; T/test_downsampling_filters.cpp:27-88;
// T/test_registration_pipeline.cpp:16-61,128-156,360-508.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <random>
#include <set>
#include <vector>

#include "sycl_points/algorithms/common/transform.hpp"
#include "sycl_points/algorithms/feature/covariance.hpp"
#include "sycl_points/algorithms/filter/preprocess_filter.hpp"
#include "sycl_points/algorithms/filter/voxel_downsampling.hpp"
#include "sycl_points/algorithms/knn/bruteforce.hpp"
#include "sycl_points/algorithms/knn/kdtree.hpp"
#include "sycl_points/algorithms/registration/registration_pipeline.hpp"

using namespace sycl_points;
using namespace sycl_points::algorithms;
using namespace sycl_points::algorithms::registration;

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                               \
    do {                                                                          \
        ++g_checks;                                                               \
        if (!(cond)) {                                                            \
            ++g_failed;                                                           \
            std::printf("  FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);       \
        }                                                                         \
    } while (0)
#define CHECK_NEAR(a, b, tol) CHECK(std::fabs((double)(a) - (double)(b)) <= (tol))

static PointCloudShared random_cloud(const sycl_utils::DeviceQueue& q, size_t n, std::mt19937& gen) {
    std::uniform_real_distribution<float> dist(-10.0f, 10.0f);
    PointCloudShared c(q);
    c.points->resize(n);
    for (size_t i = 0; i < n; ++i) {
        const float x = dist(gen), y = dist(gen), z = dist(gen);
        (*c.points)[i] = PointType(x, y, z, 1.0f);
    }
    return c;
}

// host KNN that counts its calls (T/test_registration_pipeline.cpp:28-61)
class CountingNearestKNN : public knn::KNNBase {
public:
    explicit CountingNearestKNN(const PointCloudShared& target) : target_(&target) {}
    sycl_utils::events knn_search_async(const PointCloudShared& queries, const size_t k, knn::KNNResult& result,
                                        const std::vector<sycl::event>& = {},
                                        const TransformMatrix& transT = TransformMatrix::Identity()) const override {
        ++calls;
        if (result.indices == nullptr) result.allocate(queries.queue, queries.size(), k);
        else result.resize(queries.size(), k);
        queries.queue.ptr->wait();
        for (size_t i = 0; i < queries.size(); ++i) {
            const PointType p = transT * (*queries.points)[i];
            float best = std::numeric_limits<float>::max();
            int32_t arg = -1;
            for (size_t j = 0; j < target_->size(); ++j) {
                const PointType d = p - (*target_->points)[j];
                const float dist = d.x() * d.x() + d.y() * d.y() + d.z() * d.z();
                if (dist < best) best = dist, arg = (int32_t)j;
            }
            (*result.indices)[i * k] = arg;
            (*result.distances)[i * k] = best;
        }
        return sycl_utils::events();
    }
    mutable int calls = 0;

private:
    const PointCloudShared* target_;
};

class DummyKNN : public knn::KNNBase {
public:
    sycl_utils::events knn_search_async(const PointCloudShared&, const size_t, knn::KNNResult&,
                                        const std::vector<sycl::event>& = {},
                                        const TransformMatrix& = TransformMatrix::Identity()) const override {
        return sycl_utils::events();
    }
};

static PointCloudShared make_cloud(const sycl_utils::DeviceQueue& queue, size_t size) {
    PointCloudShared cloud(queue);
    cloud.points->resize(size);
    cloud.intensities->resize(size);
    for (size_t i = 0; i < size; ++i) {
        (*cloud.points)[i] = PointType((float)i, (float)(i + 1), (float)(i + 2), 1.0f);
        (*cloud.intensities)[i] = (float)i;
    }
    return cloud;
}

static void test_knn(const sycl_utils::DeviceQueue& queue) {
    std::printf("[knn] known answer, self search, index vs brute force (seed 1234)\n");
    {  // T/test_kdtree.cpp:358-389
        PointCloudShared target(queue), query(queue);
        target.points->push_back(PointType(0.f, 0.f, 0.f, 1.f));
        query.points->push_back(PointType(1.f, 1.f, 1.f, 1.f));
        const auto tree = knn::KDTree::build(queue, target);
        const auto r = tree->knn_search(query, 1);
        CHECK((*r.indices)[0] == 0);
        CHECK_NEAR((*r.distances)[0], 3.0f, 1e-6);
    }
    std::mt19937 gen(1234);
    const auto target = random_cloud(queue, 1000, gen);
    const auto query = random_cloud(queue, 100, gen);
    const auto tree = knn::KDTree::build(queue, target);
    for (size_t k : {1, 3, 5, 10, 20}) {  // T/test_kdtree.cpp:301-317
        const auto a = tree->knn_search(query, k);
        const auto b = knn::knn_search_bruteforce(queue, query, target, k);
        bool same = a.indices->size() == b.indices->size();
        for (size_t i = 0; same && i < a.indices->size(); ++i)
            same = (*a.indices)[i] == (*b.indices)[i] && (*a.distances)[i] == (*b.distances)[i];
        CHECK(same);
        // and against a plain host scan, ties grouped within 1e-4 like the reference's comparison
        for (size_t qi = 0; qi < query.size(); ++qi) {
            std::vector<std::pair<float, int>> all;
            for (size_t j = 0; j < target.size(); ++j) {
                const PointType d = (*query.points)[qi] - (*target.points)[j];
                all.push_back({d.x() * d.x() + d.y() * d.y() + d.z() * d.z(), (int)j});
            }
            std::sort(all.begin(), all.end());
            for (size_t j = 0; j < k; ++j) CHECK_NEAR((*a.distances)[qi * k + j], all[j].first, 1e-4);
        }
    }
    {  // T/test_kdtree.cpp:472-476
        const auto r = tree->knn_search(target, 1);
        bool ok = true;
        for (size_t i = 0; i < target.size(); ++i) ok = ok && (*r.indices)[i] == (int32_t)i && (*r.distances)[i] == 0.0f;
        CHECK(ok);
    }
    {  // transform inside the search (knn.hpp:22-24)
        TransformMatrix T = TransformMatrix::Identity();
        T(0, 3) = 0.25f;
        const auto r = tree->knn_search(query, 1, {}, T);
        PointCloudShared moved(queue);
        for (size_t i = 0; i < query.size(); ++i) moved.points->push_back(T * (*query.points)[i]);
        const auto ref = knn::knn_search_bruteforce(queue, moved, target, 1);
        bool ok = true;
        for (size_t i = 0; i < query.size(); ++i) ok = ok && (*r.indices)[i] == (*ref.indices)[i];
        CHECK(ok);
    }
}

static void test_voxel(const sycl_utils::DeviceQueue& queue) {
    std::printf("[voxel] mean / median aggregation known answer\n");
    PointCloudCPU cpu;  // T/test_downsampling_filters.cpp:27-88
    const float xs[5] = {0.10f, 0.40f, 1.10f, 1.40f, 0.20f};
    const float rgbs[5][3] = {{10, 20, 30}, {20, 40, 60}, {30, 60, 90}, {50, 70, 90}, {70, 80, 90}};
    const float inten[5] = {1.f, 3.f, 5.f, 7.f, 100.f};
    for (int i = 0; i < 5; ++i) {
        cpu.points->push_back(PointType(xs[i], 0.f, 0.f, 1.f));
        cpu.rgb->push_back(RGBType(rgbs[i][0], rgbs[i][1], rgbs[i][2], 1.f));
        cpu.intensities->push_back(inten[i]);
        cpu.timestamp_offsets->push_back(2.0f * i);
    }
    PointCloudShared cloud(queue, cpu), result(queue);
    filter::VoxelGrid vg(queue, 1.0f);
    vg.set_min_voxel_count(2);
    vg.downsampling(cloud, result);
    CHECK(result.size() == 2);
    CHECK(result.has_rgb() && result.has_intensity() && result.has_timestamps());
    int first = -1;
    for (size_t i = 0; i < result.size(); ++i)
        if (std::fabs((*result.points)[i].x() - 0.233333f) < 1e-5f) first = (int)i;
    CHECK(first != -1);
    if (first >= 0) {
        CHECK_NEAR((*result.intensities)[first], 3.0f, 1e-5);
        CHECK_NEAR((*result.timestamp_offsets)[first], 3.333333f, 1e-5);
        CHECK_NEAR((*result.rgb)[first].x(), 33.333333f, 1e-5);
        CHECK_NEAR((*result.rgb)[first].y(), 46.666667f, 1e-5);
        CHECK_NEAR((*result.rgb)[first].z(), 60.0f, 1e-5);
    }
    bool threw = false;
    try {
        filter::VoxelGrid bad(queue, 0.0f);
    } catch (const std::invalid_argument&) {
        threw = true;
    }
    CHECK(threw);
}

static void test_pipeline(const sycl_utils::DeviceQueue& queue) {
    std::printf("[pipeline] random sampling, robust-scale schedule, robust weights through injected KNN\n");
    {  // T/test_registration_pipeline.cpp:128-156
        RegistrationPipelineParams params;
        params.random_sampling.enable = true;
        params.random_sampling.num = 3;
        size_t aligned = 0;
        bool has_intensity = false;
        auto aligner = [&](const PointCloudShared& source, const PointCloudShared&, const knn::KNNBase&,
                           const TransformMatrix&, const Registration::ExecutionOptions&) {
            aligned = source.size();
            has_intensity = source.has_intensity();
            RegistrationResult r;
            r.inlier = (uint32_t)source.size();
            return r;
        };
        RegistrationPipeline pipeline(aligner, params);
        DummyKNN knn;
        const auto res = pipeline.align(make_cloud(queue, 6), make_cloud(queue, 4), knn);
        CHECK(aligned == 3 && has_intensity);
        CHECK(pipeline.get_registration_input_point_cloud()->size() == 3);
        CHECK_NEAR(pipeline.get_inlier_ratio(res), 1.0f, 1e-6);
        // order-preserving compaction: x strictly increasing
        const auto* in = pipeline.get_registration_input_point_cloud();
        CHECK((*in->points)[0].x() < (*in->points)[1].x() && (*in->points)[1].x() < (*in->points)[2].x());
    }
    {  // :360-409
        RegistrationPipelineParams params;
        params.registration.robust.type = robust::RobustLossType::HUBER;
        params.registration.robust.default_scale = 8.0f;
        params.random_sampling.enable = false;
        std::vector<float> scales, rot;
        auto aligner = [&](const PointCloudShared&, const PointCloudShared&, const knn::KNNBase&, const TransformMatrix&,
                           const Registration::ExecutionOptions& o) {
            scales.push_back(o.robust_scale);
            rot.push_back(o.rotation_robust_scale);
            return RegistrationResult{};
        };
        DummyKNN knn;
        RegistrationPipeline fixed(aligner, params);
        fixed.align(make_cloud(queue, 3), make_cloud(queue, 3), knn);
        CHECK(scales.size() == 1 && scales[0] == -1.0f);
        params.robust.auto_scale = true;
        params.robust.init_scale = 6.0f;
        params.robust.min_scale = 2.0f;
        params.robust.rotation_init_scale = 9.0f;
        params.robust.rotation_min_scale = 3.0f;
        params.robust.auto_scaling_iter = 3;
        scales.clear();
        rot.clear();
        RegistrationPipeline annealed(aligner, params);
        annealed.align(make_cloud(queue, 3), make_cloud(queue, 3), knn);
        CHECK(scales.size() == 3);
        if (scales.size() == 3) {
            CHECK(scales[0] == 6.0f);
            CHECK_NEAR(scales[1], std::sqrt(12.0f), 1e-5);
            CHECK_NEAR(scales[2], 2.0f, 1e-5);
            CHECK(rot[0] == 9.0f);
            CHECK_NEAR(rot[1], std::sqrt(27.0f), 1e-5);
            CHECK_NEAR(rot[2], 3.0f, 1e-5);
        }
    }
    {  // :411-508: source {0,1,5} on x, target {0,1}; P2P; NONE -> {1,1,0} at max_corr 1.5; HUBER on residual 3
        PointCloudShared source(queue), target(queue);
        for (float x : {0.f, 1.f, 5.f}) source.points->push_back(PointType(x, 0.f, 0.f, 1.f));
        for (float x : {0.f, 1.f}) target.points->push_back(PointType(x, 0.f, 0.f, 1.f));
        RegistrationParams params;
        params.reg_type = RegType::POINT_TO_POINT;
        params.max_correspondence_distance = 1.5f;
        Registration reg(queue, params);
        CountingNearestKNN knn(target);
        shared_vector<float> w;
        reg.compute_icp_robust_weights(source, target, knn, TransformMatrix::Identity(), 1.0f, w);
        queue.ptr->wait();
        CHECK(knn.calls == 1);
        CHECK(w.size() == 3 && w[0] == 1.0f && w[1] == 1.0f && w[2] == 0.0f);

        params.robust.type = robust::RobustLossType::HUBER;
        params.max_correspondence_distance = 10.0f;
        Registration huber(queue, params);
        PointCloudShared far(queue);
        far.points->push_back(PointType(4.f, 0.f, 0.f, 1.f));  // residual 3 to target x = 1
        for (float s : {1.0f, 2.0f}) {
            huber.compute_icp_robust_weights(far, target, knn, TransformMatrix::Identity(), s, w);
            CHECK_NEAR(w[0], s / 3.0f, 1e-5);
        }
    }
}

static void test_align(const sycl_utils::DeviceQueue& queue) {
    std::printf("[align] synthetic surface, GICP / P2Plane / P2P, index path vs injected-KNN path\n");
    std::mt19937 gen(7);
    std::uniform_real_distribution<float> u(-6.f, 6.f);
    std::normal_distribution<float> noise(0.f, 0.005f);
    PointCloudShared target(queue), source(queue);
    TransformMatrix T = TransformMatrix::Identity();  // source = T_gt^-1 * target, T_gt = small yaw + shift
    const float a = 0.02f;
    T(0, 0) = std::cos(a), T(0, 1) = -std::sin(a), T(1, 0) = std::sin(a), T(1, 1) = std::cos(a);
    T(0, 3) = 0.15f, T(1, 3) = -0.08f, T(2, 3) = 0.03f;
    const Eigen::Isometry3f Tinv = Eigen::Isometry3f(T).inverse();
    for (int i = 0; i < 6000; ++i) {
        const float x = u(gen), y = u(gen);
        PointType p;
        if (i % 3 == 0) p = PointType(x, y, 0.1f * x - 1.0f + noise(gen), 1.f);
        else if (i % 3 == 1) p = PointType(x, 5.0f + noise(gen), 0.4f * y + 1.0f, 1.f);
        else p = PointType(-5.5f + noise(gen), x, 0.4f * y + 1.0f, 1.f);
        target.points->push_back(p);
        source.points->push_back(Tinv.matrix() * p);
    }
    const auto tree_t = knn::KDTree::build(queue, target);
    const auto tree_s = knn::KDTree::build(queue, source);
    covariance::estimate_async(*tree_t, target, 10).wait_and_throw();
    covariance::estimate_async(*tree_s, source, 10).wait_and_throw();
    covariance::estimate_normals_async(*tree_t, target, 10).wait_and_throw();
    CHECK(target.has_cov() && source.has_cov() && target.has_normal());

    {  // transform::transform_copy (device) vs transform_cpu_copy (host loop): T * source == target
        auto moved = transform::transform_copy(source, T);
        auto moved_cpu = transform::transform_cpu_copy(source, T);
        queue.ptr->wait();
        float worst = 0.f, worst_cov = 0.f;
        for (size_t i = 0; i < moved.size(); ++i) {
            worst = std::max(worst, ((*moved.points)[i] - (*target.points)[i]).norm());
            worst = std::max(worst, ((*moved.points)[i] - (*moved_cpu.points)[i]).norm());
            worst_cov = std::max(worst_cov, ((*moved.covs)[i] - (*moved_cpu.covs)[i]).norm());
        }
        CHECK(moved.size() == source.size() && moved.has_cov());
        CHECK(worst < 1e-4f);
        CHECK(worst_cov < 1e-5f);
    }

    struct Wrap : knn::KNNBase {  // hides the KDTree type: forces the generic (injected) path
        const knn::KDTree& t;
        explicit Wrap(const knn::KDTree& tree) : t(tree) {}
        sycl_utils::events knn_search_async(const PointCloudShared& q, const size_t k, knn::KNNResult& r,
                                            const std::vector<sycl::event>& d = {},
                                            const TransformMatrix& transT = TransformMatrix::Identity()) const override {
            return t.knn_search_async(q, k, r, d, transT);
        }
    };
    for (RegType rt : {RegType::GICP, RegType::POINT_TO_PLANE, RegType::POINT_TO_POINT})
        for (OptimizationMethod om : {OptimizationMethod::GAUSS_NEWTON, OptimizationMethod::LEVENBERG_MARQUARDT,
                                      OptimizationMethod::POWELL_DOGLEG}) {
            RegistrationParams params;
            params.reg_type = rt;
            params.optimization_method = om;
            params.robust.type = robust::RobustLossType::HUBER;
            params.robust.default_scale = 1.0f;
            params.max_iterations = 30;
            Registration reg(queue, params), reg2(queue, params);
            const auto r1 = reg.align(source, target, *tree_t);
            const auto r2 = reg2.align(source, target, Wrap(*tree_t));
            float d_gt = 0.f, d_12 = 0.f;
            for (int i = 0; i < 3; ++i) {
                d_gt += std::pow(r1.T.matrix()(i, 3) - T(i, 3), 2.f);
                d_12 += std::pow(r1.T.matrix()(i, 3) - r2.T.matrix()(i, 3), 2.f);
            }
            CHECK(r1.converged);
            CHECK(std::sqrt(d_gt) < 0.01f);   // recovers the ground-truth shift to < 1 cm
            CHECK(std::sqrt(d_12) < 1e-4f);   // both call paths agree
            CHECK(r1.inlier > 5000);
        }
    bool threw = false;  // registration.hpp:166-171
    try {
        PointCloudShared bare(queue);
        bare.points->push_back(PointType(0.f, 0.f, 0.f, 1.f));
        Registration reg(queue);
        reg.align(bare, target, *tree_t);
    } catch (const std::runtime_error&) {
        threw = true;
    }
    CHECK(threw);
}

int main() {
    sycl::device dev(sycl_utils::device_selector::default_selector_v);
    sycl_utils::DeviceQueue queue(dev);
    queue.print_device_info();
    test_knn(queue);
    test_voxel(queue);
    test_pipeline(queue);
    test_align(queue);
    std::printf("%d checks, %d failed\n", g_checks, g_failed);
    return g_failed == 0 ? 0 : 1;
}
