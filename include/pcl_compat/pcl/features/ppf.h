// pcl::PPFEstimation<PointInT, PointNT, PointOutT> over libb200ppf (K1).
// Replaces [PCL] features/include/pcl/features/ppf.h + impl/ppf.hpp.
#pragma once

#include <cstddef>
#include <vector>

#include "../b200_context.h"
#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {

template <typename PointInT, typename PointNT, typename PointOutT>
class PPFEstimation {
public:
    using Ptr = shared_ptr<PPFEstimation<PointInT, PointNT, PointOutT>>;
    using ConstPtr = shared_ptr<const PPFEstimation<PointInT, PointNT, PointOutT>>;
    using PointCloudIn = PointCloud<PointInT>;
    using PointCloudN = PointCloud<PointNT>;
    using PointCloudOut = PointCloud<PointOutT>;

    PPFEstimation() = default;

    void setInputCloud(const typename PointCloudIn::ConstPtr &cloud) { input_ = cloud; }
    void setInputNormals(const typename PointCloudN::ConstPtr &normals) { normals_ = normals; }
    typename PointCloudIn::ConstPtr getInputCloud() const { return input_; }
    typename PointCloudN::ConstPtr getInputNormals() const { return normals_; }
    // PCL lets a subset of reference points be selected; the table build needs all of them
    void setIndices(const shared_ptr<const std::vector<int>> &indices) { indices_ = indices; }

    // Feature::compute -> PPFEstimation::computeFeature: output[i*N + j] for every ordered pair,
    // NaN signature (and is_dense = false) for i == j and failed pairs.
    void compute(PointCloudOut &output) {
        static_assert(sizeof(PointOutT) == sizeof(b200ppf_signature), "PointOutT must be pcl::PPFSignature");
        output.clear();
        if (!input_ || input_->empty()) {
            PCL_ERROR("[pcl::PPFEstimation::compute] input cloud is not set or empty\n");
            return;
        }
        if (!normals_ || normals_->size() != input_->size()) {
            PCL_ERROR("[pcl::PPFEstimation::compute] normals are not set or differ in size from the input cloud\n");
            return;
        }
        if (indices_ && indices_->size() != input_->size()) {
            PCL_ERROR("[pcl::PPFEstimation::compute] index subsets are not supported by the B200 engine\n");
            return;
        }
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx) return;
        const std::size_t n = input_->size();
        std::vector<float> packed(n * 6);
        for (std::size_t i = 0; i < n; ++i) {
            const PointInT &p = (*input_)[i];
            const PointNT &q = (*normals_)[i];
            packed[6 * i + 0] = p.x; packed[6 * i + 1] = p.y; packed[6 * i + 2] = p.z;
            packed[6 * i + 3] = q.normal_x; packed[6 * i + 4] = q.normal_y; packed[6 * i + 5] = q.normal_z;
        }
        b200::CloudHandle cloud;
        b200::FeaturesHandle feats;
        if (b200ppf_cloud_upload(ctx, packed.data(), n, 6, 3, &cloud.h) != B200PPF_OK ||
            b200ppf_cloud_size(cloud.h) != n ||  // NaN points would shift the pair indexing
            b200ppf_features_compute(ctx, cloud.h, &feats.h) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFEstimation::compute] %s\n", b200ppf_cloud_size(cloud.h) != n
                                                                 ? "input contains NaN points"
                                                                 : b200ppf_last_error(ctx));
            return;
        }
        output.points.resize(n * n);
        output.width = static_cast<std::uint32_t>(n * n);
        output.height = 1;
        if (b200ppf_features_download(ctx, feats.h, 0, n * n,
                                      reinterpret_cast<b200ppf_signature *>(output.points.data())) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFEstimation::compute] %s\n", b200ppf_last_error(ctx));
            output.clear();
            return;
        }
        output.is_dense = false;  // the diagonal is NaN by construction (PCL sets the same)
    }

private:
    typename PointCloudIn::ConstPtr input_;
    typename PointCloudN::ConstPtr normals_;
    shared_ptr<const std::vector<int>> indices_;
};

}  // namespace pcl
