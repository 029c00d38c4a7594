// pcl::NormalEstimationOMP<PointInT, PointOutT> (and NormalEstimation) over libb200ppf (P2 + P4, prep.cu).
// Replaces [PCL] features/include/pcl/features/normal_3d_omp.h + impl/normal_3d_omp.hpp and the helpers it
// calls (normal_3d.h computePointNormal / flipNormalTowardsViewpoint, common/impl/centroid.hpp,
// common/impl/eigen.hpp) for the call the reference makes: CloudProcessor::NormalEstimation,
// include/CloudProcessing.h:378-401 (setInputCloud, setNumberOfThreads(12), setSearchMethod(tree), setKSearch(30),
// compute).  Radius search (setRadiusSearch) and a separate search surface are not offered.
#pragma once

#include <cstddef>
#include <limits>
#include <vector>

#include "../b200_context.h"
#include "../point_cloud.h"
#include "../point_types.h"
#include "../search/kdtree.h"

namespace pcl {

template <typename PointInT, typename PointOutT>
class NormalEstimationOMP {
public:
    using PointCloudIn = PointCloud<PointInT>;
    using PointCloudOut = PointCloud<PointOutT>;
    using KdTreePtr = typename search::KdTree<PointInT>::Ptr;

    explicit NormalEstimationOMP(unsigned int nr_threads = 0) : threads_(nr_threads) {}
    void setInputCloud(const typename PointCloudIn::ConstPtr &cloud) { input_ = cloud; }
    void setNumberOfThreads(unsigned int nr_threads = 0) { threads_ = nr_threads; }  // the device has its own
    void setSearchMethod(const KdTreePtr &tree) { tree_ = tree; }
    void setKSearch(int k) { k_ = k; }
    int getKSearch() const { return k_; }
    void setViewPoint(float vpx, float vpy, float vpz) { vp_[0] = vpx; vp_[1] = vpy; vp_[2] = vpz; }
    void getViewPoint(float &vpx, float &vpy, float &vpz) const { vpx = vp_[0]; vpy = vp_[1]; vpz = vp_[2]; }
    // extension: which computeMeanAndCovarianceMatrix to follow (B200PPF_COVARIANCE_SHIFTED = PCL >= 1.12, default)
    void setCovarianceMode(int mode) { cov_mode_ = mode; }

    void compute(PointCloudOut &output) {
        output.clear();
        if (!input_ || input_->empty()) {
            PCL_ERROR("[pcl::NormalEstimationOMP::compute] input cloud is not set or empty\n");
            return;
        }
        if (k_ <= 0) {
            PCL_ERROR("[pcl::NormalEstimationOMP::compute] Neither radius nor K defined! Set one of them to zero first and then re-run compute ().\n");
            return;
        }
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx) return;
        b200::CloudHandle cloud;
        const std::size_t n = input_->size();
        if (b200ppf_cloud_upload_xyz(ctx, reinterpret_cast<const float *>(input_->points.data()), n,
                                     sizeof(PointInT) / sizeof(float), &cloud.h) != B200PPF_OK ||
            b200ppf_cloud_size(cloud.h) != n ||
            b200ppf_normal_estimation(ctx, cloud.h, k_, vp_, cov_mode_) != B200PPF_OK) {
            PCL_ERROR("[pcl::NormalEstimationOMP::compute] %s\n", cloud.h && b200ppf_cloud_size(cloud.h) != n
                                                                      ? "input contains non-finite points"
                                                                      : b200ppf_last_error(ctx));
            return;
        }
        std::vector<float> rows(n * 8);
        if (b200ppf_cloud_download(ctx, cloud.h, rows.data(), 8, 3, 6) != B200PPF_OK) {
            PCL_ERROR("[pcl::NormalEstimationOMP::compute] %s\n", b200ppf_last_error(ctx));
            return;
        }
        output.points.resize(n);
        output.width = input_->width;
        output.height = input_->height;
        output.is_dense = true;
        for (std::size_t i = 0; i < n; ++i) {
            PointOutT &o = output.points[i];
            o.normal_x = rows[8 * i + 3]; o.normal_y = rows[8 * i + 4]; o.normal_z = rows[8 * i + 5];
            o.curvature = rows[8 * i + 6];
            if (o.normal_x != o.normal_x) output.is_dense = false;  // NaN normal: fewer than three neighbours
        }
    }

private:
    typename PointCloudIn::ConstPtr input_;
    KdTreePtr tree_;
    unsigned int threads_;
    int k_ = 0;
    float vp_[3] = {0.f, 0.f, 0.f};
    int cov_mode_ = B200PPF_COVARIANCE_SHIFTED;
};

template <typename PointInT, typename PointOutT>
using NormalEstimation = NormalEstimationOMP<PointInT, PointOutT>;

}  // namespace pcl
