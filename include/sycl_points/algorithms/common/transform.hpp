// transform::transform_async / transform / transform_copy / transform_cpu —
// I/algorithms/common/transform.hpp:45-188.  The device path is one libspx kernel (spx_transform).
#pragma once

#include <array>

#include "sycl_points/points/point_cloud.hpp"
#include "sycl_points/utils/eigen_utils.hpp"

namespace sycl_points {
namespace algorithms {
namespace transform {

namespace kernel {
/// trans * cov * trans^T on the host (the reference's device helper of the same name, transform.hpp:14-22; the
/// device side of it lives in csrc/spx_features.cu)
inline void transform_covs(const Covariance& cov, Covariance& result, const std::array<sycl::float4, 4>& trans) {
    const Eigen::Matrix4f T = eigen_utils::from_sycl_vec(trans);
    const Eigen::Matrix4f r = T * (cov * T.transpose());
    result = r;
}
}  // namespace kernel

/// in place: points T p, covariances T C T^T, normals T n (not re-normalised, like the reference's kernel)
inline sycl_utils::events transform_async(PointCloudShared& cloud, const TransformMatrix& trans) {
    sycl_utils::events events;
    const size_t N = cloud.size();
    if (N == 0) return events;
    const auto& q = cloud.queue;
    cloud.index_hint = PointCloudShared::IndexHint{};  // the points move: the voxel box no longer describes them
    q.set_accessed_by_device(cloud.points_ptr(), N);
    if (cloud.has_cov()) q.set_accessed_by_device(cloud.covs_ptr(), N);
    if (cloud.has_normal()) q.set_accessed_by_device(cloud.normals_ptr(), N);
    detail::spx_check(spx_transform(q.handle(), reinterpret_cast<float*>(cloud.points_ptr()),
                                    cloud.has_cov() ? reinterpret_cast<float*>(cloud.covs_ptr()) : nullptr,
                                    cloud.has_normal() ? reinterpret_cast<float*>(cloud.normals_ptr()) : nullptr, N,
                                    trans.data()));
    events += q.checkpoint();
    return events;
}

inline void transform(PointCloudShared& cloud, const TransformMatrix& trans) {
    transform_async(cloud, trans).wait_and_throw();
}

inline PointCloudShared transform_copy(const PointCloudShared& cloud, const TransformMatrix& trans) {
    PointCloudShared ret(cloud);  // device-side copy of every attribute
    transform(ret, trans);
    return ret;
}

/// host loop over the shared containers (transform.hpp:148-170); normals ARE re-normalised here
inline void transform_cpu(PointCloudShared& cloud, const TransformMatrix& trans) {
    const size_t N = cloud.size();
    if (N == 0) return;
    cloud.queue.ptr->wait();
    for (size_t i = 0; i < N; ++i) (*cloud.points)[i] = trans * (*cloud.points)[i];
    if (cloud.has_cov()) {
        const TransformMatrix trans_T = trans.transpose();
        for (size_t i = 0; i < N; ++i) (*cloud.covs)[i] = trans * (*cloud.covs)[i] * trans_T;
    }
    if (cloud.has_normal()) {
        const Eigen::Matrix3f R = trans.template block<3, 3>(0, 0);
        for (size_t i = 0; i < N; ++i) {
            const Normal& v = (*cloud.normals)[i];
            const Eigen::Vector3f r = (R * Eigen::Vector3f(v.x(), v.y(), v.z())).normalized();
            (*cloud.normals)[i] = Normal(r.x(), r.y(), r.z(), v.w());
        }
    }
}

inline PointCloudShared transform_cpu_copy(PointCloudShared& cloud, const TransformMatrix& trans) {
    PointCloudShared ret(cloud);
    transform_cpu(ret, trans);
    return ret;
}

}  // namespace transform
}  // namespace algorithms
}  // namespace sycl_points
