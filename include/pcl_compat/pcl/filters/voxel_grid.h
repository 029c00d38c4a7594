// pcl::VoxelGrid<PointT> over libb200ppf (P1, prep.cu).  Replaces [PCL] filters/include/pcl/filters/voxel_grid.h +
// impl/voxel_grid.hpp for the call the reference makes: CloudProcessor::Subsampling,
// include/CloudProcessing.h:359-377 (setInputCloud, setLeafSize(Eigen::Vector4f), filter).
#pragma once

#include <cstddef>
#include <vector>

#include "../b200_context.h"
#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {

template <typename PointT>
class VoxelGrid {
public:
    using PointCloudT = PointCloud<PointT>;
    using Ptr = shared_ptr<VoxelGrid<PointT>>;

    VoxelGrid() = default;
    void setInputCloud(const typename PointCloudT::ConstPtr &cloud) { input_ = cloud; }
    typename PointCloudT::ConstPtr getInputCloud() const { return input_; }
    void setLeafSize(const Eigen::Vector4f &leaf_size) { setLeafSize(leaf_size[0], leaf_size[1], leaf_size[2]); }
    void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
    Eigen::Vector3f getLeafSize() const { return Eigen::Vector3f(leaf_[0], leaf_[1], leaf_[2]); }

    // one centroid per occupied leaf, ascending leaf index; PCL's error behaviour: message + empty output
    void filter(PointCloudT &output) {
        output.clear();
        if (!input_) {
            PCL_ERROR("[pcl::VoxelGrid::filter] No input dataset given!\n");
            return;
        }
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx) return;
        b200::CloudHandle in, out;
        if (b200ppf_cloud_upload_xyz(ctx, reinterpret_cast<const float *>(input_->points.data()), input_->size(),
                                     sizeof(PointT) / sizeof(float), &in.h) != B200PPF_OK ||
            b200ppf_voxel_grid(ctx, in.h, leaf_, &out.h) != B200PPF_OK) {
            PCL_ERROR("[pcl::VoxelGrid::filter] %s\n", b200ppf_last_error(ctx));
            return;
        }
        if (b200ppf_last_error(ctx)[0]) PCL_WARN("[pcl::VoxelGrid::applyFilter] %s\n", b200ppf_last_error(ctx));
        b200::downloadXYZ(ctx, out.h, output);
    }

private:
    typename PointCloudT::ConstPtr input_;
    float leaf_[3] = {0.f, 0.f, 0.f};
};

}  // namespace pcl
