// pcl::PPFHashMapSearch and pcl::PPFRegistration<PointSource, PointTarget> over libb200ppf
// (K2..K5).  Replaces [PCL] registration/include/pcl/registration/ppf_registration.h,
// impl/ppf_registration.hpp and registration/src/ppf_registration.cpp.
#pragma once

#include <cmath>
#include <cstddef>
#include <utility>
#include <string>
#include <vector>

#include "../b200_context.h"
#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {

class PPFHashMapSearch {
public:
    using Ptr = shared_ptr<PPFHashMapSearch>;
    using ConstPtr = shared_ptr<const PPFHashMapSearch>;

    // PCL defaults: 12 degrees, 0.01
    PPFHashMapSearch(float angle_discretization_step = 12.0f / 180.0f * static_cast<float>(M_PI),
                     float distance_discretization_step = 0.01f)
        : angle_discretization_step_(angle_discretization_step),
          distance_discretization_step_(distance_discretization_step) {}

    // builds the (device) table from the N*N signatures of PPFEstimation::compute
    void setInputFeatureCloud(PointCloud<PPFSignature>::ConstPtr feature_cloud) {
        internals_initialized_ = false;
        table_.reset();
        alpha_m_.clear();
        max_dist_ = -1.0f;
        if (!feature_cloud) {
            PCL_ERROR("[pcl::PPFHashMapSearch::setInputFeatureCloud] null feature cloud\n");
            return;
        }
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx) return;
        b200::FeaturesHandle feats;
        if (b200ppf_features_upload(ctx, reinterpret_cast<const b200ppf_signature *>(feature_cloud->points.data()),
                                    feature_cloud->size(), &feats.h) != B200PPF_OK ||
            b200ppf_table_build(ctx, feats.h, angle_discretization_step_, distance_discretization_step_, &table_.h) !=
                B200PPF_OK) {
            PCL_ERROR("[pcl::PPFHashMapSearch::setInputFeatureCloud] %s\n", b200ppf_last_error(ctx));
            table_.reset();
            return;
        }
        b200ppf_table_info info;
        b200ppf_table_get_info(table_.h, &info);
        max_dist_ = info.max_dist;
        // the public alpha_m_[i][j] member, as PCL fills it
        const std::size_t n = static_cast<std::size_t>(info.n_model);
        alpha_m_.assign(n, std::vector<float>(n));
        for (std::size_t i = 0; i < n; ++i)
            for (std::size_t j = 0; j < n; ++j) alpha_m_[i][j] = (*feature_cloud)[i * n + j].alpha_m;
        internals_initialized_ = true;
    }

    // extension: K1+K2 fused on the device, no N*N host feature cloud (alpha_m_ stays empty)
    template <typename PointT>
    bool setInputModel(const PointCloud<PointT> &model) {
        internals_initialized_ = false;
        table_.reset();
        alpha_m_.clear();
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx) return false;
        b200::CloudHandle cloud;
        if (b200ppf_cloud_upload(ctx, reinterpret_cast<const float *>(model.points.data()), model.size(),
                                 sizeof(PointT) / sizeof(float), 4, &cloud.h) != B200PPF_OK ||
            b200ppf_table_build_from_cloud(ctx, cloud.h, angle_discretization_step_, distance_discretization_step_,
                                           &table_.h) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFHashMapSearch::setInputModel] %s\n", b200ppf_last_error(ctx));
            table_.reset();
            return false;
        }
        b200ppf_table_info info;
        b200ppf_table_get_info(table_.h, &info);
        max_dist_ = info.max_dist;
        internals_initialized_ = true;
        return true;
    }

    // extension: trained-model persistence — what the reference does with detector.write(FileStorage) /
    // detector.read() (include/CloudProcessing.h:242-258, :106-121).  load() replaces the table without
    // re-running PPFEstimation; the discretisation steps become the file's; alpha_m_ stays empty.
    bool saveTrained(const std::string &path) const {
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx || !internals_initialized_) {
            PCL_ERROR("[pcl::PPFHashMapSearch::saveTrained] the search object has no table\n");
            return false;
        }
        if (b200ppf_table_save(ctx, table_.h, path.c_str()) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFHashMapSearch::saveTrained] %s\n", b200ppf_last_error(ctx));
            return false;
        }
        return true;
    }
    bool loadTrained(const std::string &path) {
        internals_initialized_ = false;
        table_.reset();
        alpha_m_.clear();
        max_dist_ = -1.0f;
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx) return false;
        if (b200ppf_table_load(ctx, path.c_str(), &table_.h) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFHashMapSearch::loadTrained] %s\n", b200ppf_last_error(ctx));
            table_.reset();
            return false;
        }
        b200ppf_table_info info;
        b200ppf_table_get_info(table_.h, &info);
        angle_discretization_step_ = info.angle_step;
        distance_discretization_step_ = info.dist_step;
        max_dist_ = info.max_dist;
        internals_initialized_ = true;
        return true;
    }

    void nearestNeighborSearch(float &f1, float &f2, float &f3, float &f4,
                               std::vector<std::pair<std::size_t, std::size_t>> &indices) {
        indices.clear();
        if (!internals_initialized_) {
            PCL_ERROR("[pcl::PPFHashMapSearch::nearestNeighborSearch] the search object has no input feature cloud\n");
            return;
        }
        b200ppf_ctx *ctx = b200::defaultContext();
        std::vector<uint64_t> buf(2 * 256);
        std::size_t found = 0;
        for (;;) {
            if (b200ppf_table_query(ctx, table_.h, f1, f2, f3, f4, buf.data(), buf.size() / 2, &found) != B200PPF_OK) {
                PCL_ERROR("[pcl::PPFHashMapSearch::nearestNeighborSearch] %s\n", b200ppf_last_error(ctx));
                return;
            }
            if (found <= buf.size() / 2) break;
            buf.resize(2 * found);
        }
        indices.reserve(found);
        for (std::size_t e = 0; e < found; ++e) indices.emplace_back(buf[2 * e], buf[2 * e + 1]);
    }

    // PCL's makeShared(): a copy of the search object — here a device-to-device copy of the table (b200ppf_table_clone)
    Ptr makeShared() const {
        Ptr copy(new PPFHashMapSearch(angle_discretization_step_, distance_discretization_step_));
        if (!internals_initialized_) return copy;
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx || b200ppf_table_clone(ctx, table_.h, &copy->table_.h) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFHashMapSearch::makeShared] %s\n", b200ppf_last_error(ctx));
            copy->table_.reset();
            return copy;
        }
        copy->alpha_m_ = alpha_m_;
        copy->max_dist_ = max_dist_;
        copy->internals_initialized_ = true;
        return copy;
    }

    float getAngleDiscretizationStep() const { return angle_discretization_step_; }
    float getDistanceDiscretizationStep() const { return distance_discretization_step_; }
    float getModelDiameter() const { return max_dist_; }

    std::vector<std::vector<float>> alpha_m_;  // public in PCL

    // device table (used by PPFRegistration)
    const b200ppf_table *deviceTable() const { return internals_initialized_ ? table_.h : nullptr; }

private:
    PPFHashMapSearch(const PPFHashMapSearch &) = delete;  // owns a device table: copies go through makeShared()
    float angle_discretization_step_, distance_discretization_step_;
    float max_dist_ = -1.0f;
    bool internals_initialized_ = false;
    b200::TableHandle table_;
};

template <typename PointSource, typename PointTarget>
class PPFRegistration {
public:
    struct PoseWithVotes {
        PoseWithVotes(const Eigen::Affine3f &a_pose, unsigned int a_votes) : pose(a_pose), votes(a_votes) {}
        Eigen::Affine3f pose;
        unsigned int votes;
    };
    using PoseWithVotesList = std::vector<PoseWithVotes>;
    using PointCloudSource = PointCloud<PointSource>;
    using PointCloudSourcePtr = typename PointCloudSource::Ptr;
    using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
    using PointCloudTarget = PointCloud<PointTarget>;
    using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
    using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
    using Matrix4 = Eigen::Matrix4f;

    // PCL defaults: sampling rate 5, 0.01 m, 20 degrees
    PPFRegistration()
        : scene_reference_point_sampling_rate_(5),
          clustering_position_diff_threshold_(0.01f),
          clustering_rotation_diff_threshold_(20.0f / 180.0f * static_cast<float>(M_PI)) {}

    void setPositionClusteringThreshold(float t) { clustering_position_diff_threshold_ = t; }
    float getPositionClusteringThreshold() { return clustering_position_diff_threshold_; }
    void setRotationClusteringThreshold(float t) { clustering_rotation_diff_threshold_ = t; }
    float getRotationClusteringThreshold() { return clustering_rotation_diff_threshold_; }
    void setSceneReferencePointSamplingRate(unsigned int r) { scene_reference_point_sampling_rate_ = r; }
    unsigned int getSceneReferencePointSamplingRate() { return scene_reference_point_sampling_rate_; }
    void setSearchMethod(PPFHashMapSearch::Ptr search_method) { search_method_ = search_method; }
    PPFHashMapSearch::Ptr getSearchMethod() { return search_method_; }

    // model
    void setInputSource(const PointCloudSourceConstPtr &cloud) {
        input_ = cloud;
        source_dev_.reset();
        multi_source_ = nullptr;
    }
    void setInputCloud(const PointCloudSourceConstPtr &cloud) { setInputSource(cloud); }  // PCL < 1.7 name
    // scene (PCL also builds its kd-tree here; the device grid is built inside align)
    void setInputTarget(const PointCloudTargetConstPtr &cloud) {
        target_ = cloud;
        target_dev_.reset();
        multi_target_ = nullptr;
    }
    PointCloudSourceConstPtr getInputSource() const { return input_; }
    PointCloudTargetConstPtr getInputTarget() const { return target_; }

    void align(PointCloudSource &output) { align(output, Matrix4::Identity()); }

    // Registration::align + PPFRegistration::computeTransformation
    void align(PointCloudSource &output, const Matrix4 &guess) {
        converged_ = false;
        final_transformation_ = transformation_ = previous_transformation_ = Matrix4::Identity();
        results_.clear();
        if (!input_ || input_->empty() || !target_ || target_->empty()) {
            PCL_ERROR("[pcl::PPFRegistration::align] source or target cloud is not set or empty\n");
            return;
        }
        // Registration::align resizes the output to the model and copies it
        if (&output != input_.get()) output = *input_;
        if (!search_method_ || !search_method_->deviceTable()) {
            PCL_ERROR("[pcl::PPFRegistration::computeTransformation] Search method not set - skipping computeTransformation!\n");
            return;
        }
        if (guess != Matrix4::Identity())
            PCL_ERROR("[pcl::PPFRegistration::computeTransformation] setting initial transform (guess) not implemented!\n");
        b200ppf_ctx *ctx = b200::defaultContext();
        if (!ctx) return;
        if (!source_dev_.h &&
            b200ppf_cloud_upload(ctx, reinterpret_cast<const float *>(input_->points.data()), input_->size(),
                                 sizeof(PointSource) / sizeof(float), 4, &source_dev_.h) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFRegistration::align] %s\n", b200ppf_last_error(ctx));
            return;
        }
        if (!target_dev_.h &&
            b200ppf_cloud_upload(ctx, reinterpret_cast<const float *>(target_->points.data()), target_->size(),
                                 sizeof(PointTarget) / sizeof(float), 4, &target_dev_.h) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFRegistration::align] %s\n", b200ppf_last_error(ctx));
            return;
        }
        float final16[16], poses[3 * 16];
        uint32_t votes[3];
        std::size_t n_out = 0;
        const unsigned int rate = scene_reference_point_sampling_rate_ ? scene_reference_point_sampling_rate_ : 1;
        if (b200ppf_multi *multi = b200::defaultMulti()) {
            // B200PPF_DEVICES lists several GPUs: the scene reference points are interleaved over them
            // (model + table replicated once per (source, table) pair, the scene once per target)
            if (multi_table_ != search_method_->deviceTable() || multi_source_ != input_.get()) {
                if (b200ppf_multi_adopt(multi, reinterpret_cast<const float *>(input_->points.data()), input_->size(),
                                        sizeof(PointSource) / sizeof(float), 4, search_method_->deviceTable()) != B200PPF_OK) {
                    PCL_ERROR("[pcl::PPFRegistration::align] %s\n", b200ppf_multi_last_error(multi));
                    return;
                }
                multi_table_ = search_method_->deviceTable();
                multi_source_ = input_.get();
                multi_target_ = nullptr;
            }
            if (multi_target_ != target_.get()) {
                if (b200ppf_multi_scene(multi, reinterpret_cast<const float *>(target_->points.data()), target_->size(),
                                        sizeof(PointTarget) / sizeof(float), 4) != B200PPF_OK) {
                    PCL_ERROR("[pcl::PPFRegistration::align] %s\n", b200ppf_multi_last_error(multi));
                    return;
                }
                multi_target_ = target_.get();
            }
            if (b200ppf_multi_register(multi, rate, clustering_position_diff_threshold_, clustering_rotation_diff_threshold_,
                                       final16, poses, votes, &n_out) != B200PPF_OK ||
                n_out == 0) {
                PCL_ERROR("[pcl::PPFRegistration::computeTransformation] %s\n", b200ppf_multi_last_error(multi));
                return;
            }
        } else if (b200ppf_register(ctx, source_dev_.h, search_method_->deviceTable(), target_dev_.h, rate,
                                    clustering_position_diff_threshold_, clustering_rotation_diff_threshold_, final16, poses,
                                    votes, &n_out) != B200PPF_OK ||
                   n_out == 0) {
            PCL_ERROR("[pcl::PPFRegistration::computeTransformation] %s\n", b200ppf_last_error(ctx));
            return;
        }
        for (std::size_t k = 0; k < n_out; ++k) results_.emplace_back(Eigen::Affine3f(rowMajor(poses + 16 * k)), votes[k]);
        // transformPointCloud(*input_, output, results.front().pose): xyz only
        std::vector<float> xyz(3 * input_->size());
        if (b200ppf_transform(ctx, source_dev_.h, final16, xyz.data(), 3) != B200PPF_OK) {
            PCL_ERROR("[pcl::PPFRegistration::computeTransformation] %s\n", b200ppf_last_error(ctx));
            return;
        }
        if (b200ppf_cloud_size(source_dev_.h) == output.size())
            for (std::size_t i = 0; i < output.size(); ++i) {
                output[i].x = xyz[3 * i];
                output[i].y = xyz[3 * i + 1];
                output[i].z = xyz[3 * i + 2];
            }
        transformation_ = final_transformation_ = rowMajor(final16);
        converged_ = true;
    }

    Matrix4 getFinalTransformation() { return final_transformation_; }
    Matrix4 getLastIncrementalTransformation() { return transformation_; }
    bool hasConverged() const { return converged_; }
    // extension: the (up to three) averaged cluster poses computeTransformation produced
    const PoseWithVotesList &getBestPoseCandidates() const { return results_; }

private:
    static Matrix4 rowMajor(const float *r) {
        Matrix4 M;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) M(i, j) = r[i * 4 + j];
        return M;
    }
    PPFHashMapSearch::Ptr search_method_;
    unsigned int scene_reference_point_sampling_rate_;
    float clustering_position_diff_threshold_, clustering_rotation_diff_threshold_;
    PointCloudSourceConstPtr input_;
    PointCloudTargetConstPtr target_;
    b200::CloudHandle source_dev_, target_dev_;
    // what the multi-GPU handle currently holds (B200PPF_DEVICES)
    const b200ppf_table *multi_table_ = nullptr;
    const void *multi_source_ = nullptr, *multi_target_ = nullptr;
    Matrix4 final_transformation_ = Matrix4::Identity(), transformation_ = Matrix4::Identity(),
            previous_transformation_ = Matrix4::Identity();
    bool converged_ = false;
    PoseWithVotesList results_;
};

}  // namespace pcl
