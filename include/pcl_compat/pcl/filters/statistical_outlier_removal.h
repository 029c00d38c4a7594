// pcl::StatisticalOutlierRemoval<PointT> over libb200ppf (P2 + P3, prep.cu).  Replaces [PCL]
// filters/include/pcl/filters/statistical_outlier_removal.h + impl for the call the reference makes:
// CloudProcessor::OutlierProcessing, include/CloudProcessing.h:340-358 (setInputCloud, setMeanK,
// setStddevMulThresh, filter).  setNegative / setKeepOrganized are not offered.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "../b200_context.h"
#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {

template <typename PointT>
class StatisticalOutlierRemoval {
public:
    using PointCloudT = PointCloud<PointT>;
    using Ptr = shared_ptr<StatisticalOutlierRemoval<PointT>>;

    explicit StatisticalOutlierRemoval(bool extract_removed_indices = false) : extract_removed_(extract_removed_indices) {}
    void setInputCloud(const typename PointCloudT::ConstPtr &cloud) { input_ = cloud; }
    void setMeanK(int nr_k) { mean_k_ = nr_k; }
    int getMeanK() const { return mean_k_; }
    void setStddevMulThresh(double stddev_mult) { std_mul_ = stddev_mult; }
    double getStddevMulThresh() const { return std_mul_; }
    // indices of the points filter() kept / removed (removed only when constructed with extract_removed_indices)
    const std::vector<int> &getKeptIndices() const { return kept_; }
    const std::vector<int> &getRemovedIndices() const { return removed_; }

    void filter(PointCloudT &output) {
        output.clear();
        kept_.clear();
        removed_.clear();
        if (!input_) {
            PCL_ERROR("[pcl::StatisticalOutlierRemoval::filter] No input dataset given!\n");
            return;
        }
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx) return;
        b200::CloudHandle in, out;
        std::vector<std::uint32_t> kept(input_->size());
        if (b200ppf_cloud_upload_xyz(ctx, reinterpret_cast<const float *>(input_->points.data()), input_->size(),
                                     sizeof(PointT) / sizeof(float), &in.h) != B200PPF_OK ||
            b200ppf_cloud_size(in.h) != input_->size() ||  // NaN points would shift the indices
            b200ppf_statistical_outlier_removal(ctx, in.h, mean_k_, std_mul_, &out.h, kept.data(), nullptr, nullptr) !=
                B200PPF_OK) {
            PCL_ERROR("[pcl::StatisticalOutlierRemoval::applyFilter] %s\n",
                      in.h && b200ppf_cloud_size(in.h) != input_->size() ? "input contains non-finite points"
                                                                         : b200ppf_last_error(ctx));
            return;
        }
        const std::size_t m = b200ppf_cloud_size(out.h);
        output.points.reserve(m);
        kept_.reserve(m);
        std::size_t next = 0;
        for (std::size_t k = 0; k < m; ++k) {
            const std::size_t i = kept[k];
            if (extract_removed_)
                for (; next < i; ++next) removed_.push_back(static_cast<int>(next));
            next = i + 1;
            kept_.push_back(static_cast<int>(i));
            output.points.push_back((*input_)[i]);  // the input's own points, every field, as PCL copies them
        }
        if (extract_removed_)
            for (; next < input_->size(); ++next) removed_.push_back(static_cast<int>(next));
        output.width = static_cast<std::uint32_t>(m);
        output.height = 1;
        output.is_dense = true;
    }

private:
    typename PointCloudT::ConstPtr input_;
    int mean_k_ = 1;       // PCL defaults
    double std_mul_ = 0.0;
    bool extract_removed_;
    std::vector<int> kept_, removed_;
};

}  // namespace pcl
