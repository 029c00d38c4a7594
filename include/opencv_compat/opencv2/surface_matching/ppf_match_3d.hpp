// cv::ppf_match_3d::PPF3DDetector over libb200ppf (csrc/k7_cvppf.cu): the class the reference trains and matches with
//     ppf_match_3d::PPF3DDetector detector(relativeSamplingStep, relativeDistanceStep);   include/CloudProcessing.h:205,217,234
//     detector.trainModel(pc);                                                           :236
//     detector.match(scene, results, relativeSceneSampleStep, relativeSceneDistance);    :442
//     detector.match_S2B(scene, edge, results, relativeSceneSampleStep, relativeSceneDistance);   :495 (private fork)
// (opencv_contrib surface_matching/include/opencv2/surface_matching/ppf_match_3d.hpp).  Clouds are the reference's
// N x 6 CV_32F matrices [x y z nx ny nz]; results are Pose3DPtr, best cluster first, as upstream returns them.
// Error behaviour: a message on stderr, results left empty (there is no CPU fallback).  Copying a detector — the
// reference copies one per call, CloudProcessing.h:432,485 — shares the trained device table.
// match_S2B: the fork's source is not available; see b200ppf.h for what this implementation infers.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../../b200ppf.h"
#include "../core_min.hpp"
#include "pose_3d.hpp"

namespace cv {
namespace ppf_match_3d {

class PPF3DDetector {
public:
    PPF3DDetector() : PPF3DDetector(0.05, 0.05, 30) {}
    PPF3DDetector(const double relativeSamplingStep, const double relativeDistanceStep = 0.05, const double numAngles = 30)
        : sampling_step_(relativeSamplingStep), distance_step_(relativeDistanceStep), num_angles_(numAngles) {}
    virtual ~PPF3DDetector() {}

    // a negative threshold keeps the default (position: relativeSamplingStep, rotation: 2 pi / numAngles);
    // useWeightedClustering is accepted for source compatibility (upstream's weighted average is not built: the
    // reference never enables it)
    void setSearchParams(const double positionThreshold = -1, const double rotationThreshold = -1, const bool useWeightedClustering = false) {
        position_threshold_ = positionThreshold;
        rotation_threshold_ = rotationThreshold;
        (void)useWeightedClustering;
        if (det_) b200cv_detector_set_search_params(det_.get(), position_threshold_, rotation_threshold_);
    }

    void trainModel(const Mat &Model) {
        det_.reset();
        if (Model.empty() || Model.cols < 6) {
            std::fprintf(stderr, "[cv::ppf_match_3d::PPF3DDetector::trainModel] the model must be an N x 6 CV_32F matrix with normals\n");
            return;
        }
        b200ppf_ctx *ctx = context();
        if (!ctx) return;
        b200cv_detector *d = nullptr;
        if (b200cv_detector_create(ctx, sampling_step_, distance_step_, num_angles_, &d) != B200PPF_OK ||
            b200cv_detector_set_search_params(d, position_threshold_, rotation_threshold_) != B200PPF_OK ||
            b200cv_detector_train(d, Model.ptr<float>(0), (size_t)Model.rows, (size_t)Model.cols) != B200PPF_OK) {
            std::fprintf(stderr, "[cv::ppf_match_3d::PPF3DDetector::trainModel] %s\n", b200ppf_last_error(ctx));
            if (d) b200cv_detector_free(d);
            return;
        }
        det_ = std::shared_ptr<b200cv_detector>(d, b200cv_detector_free);
    }

    void match(const Mat &scene, std::vector<Pose3DPtr> &results, const double relativeSceneSampleStep = 1.0 / 5.0,
               const double relativeSceneDistance = 0.03) {
        run(scene, nullptr, results, relativeSceneSampleStep, relativeSceneDistance);
    }

    void match_S2B(const Mat &scene, const Mat &edge, std::vector<Pose3DPtr> &results, const double relativeSceneSampleStep = 1.0 / 5.0,
                   const double relativeSceneDistance = 0.03) {
        run(scene, &edge, results, relativeSceneSampleStep, relativeSceneDistance);
    }

    bool trained() const { return (bool)det_; }
    // extension: the device handle (parity tests, timings)
    b200cv_detector *deviceDetector() const { return det_.get(); }

private:
    void run(const Mat &scene, const Mat *edge, std::vector<Pose3DPtr> &results, double sample_step, double distance) {
        results.clear();
        if (!det_) {
            std::fprintf(stderr, "[cv::ppf_match_3d::PPF3DDetector::match] the model is not trained\n");
            return;
        }
        if (scene.empty() || scene.cols < 6 || (edge && (edge->empty() || edge->cols < 6))) {
            std::fprintf(stderr, "[cv::ppf_match_3d::PPF3DDetector::match] clouds must be N x 6 CV_32F matrices with normals\n");
            return;
        }
        b200ppf_ctx *ctx = context();
        std::size_t n = 0;
        std::vector<b200cv_pose> poses(64);
        for (int pass = 0; pass < 2; ++pass) {
            const int rc = edge ? b200cv_detector_match_s2b(det_.get(), scene.ptr<float>(0), (size_t)scene.rows, (size_t)scene.cols,
                                                            edge->ptr<float>(0), (size_t)edge->rows, (size_t)edge->cols, sample_step,
                                                            distance, poses.data(), poses.size(), &n)
                                : b200cv_detector_match(det_.get(), scene.ptr<float>(0), (size_t)scene.rows, (size_t)scene.cols,
                                                        sample_step, distance, poses.data(), poses.size(), &n);
            if (rc != B200PPF_OK) {
                std::fprintf(stderr, "[cv::ppf_match_3d::PPF3DDetector::match] %s\n", b200ppf_last_error(ctx));
                return;
            }
            if (n <= poses.size()) break;
            poses.resize(n);  // upstream returns every cluster
        }
        for (std::size_t k = 0; k < n && k < poses.size(); ++k) {
            Pose3DPtr p(new Pose3D(poses[k].alpha, poses[k].model_index, poses[k].num_votes));
            p->updatePose(Matx44d(poses[k].pose));
            results.push_back(p);
        }
    }

    // one device context per process (B200PPF_DEVICE / LOCAL_RANK select the GPU)
    static b200ppf_ctx *context() {
        static b200ppf_ctx *ctx = nullptr;
        static bool tried = false;
        if (!tried) {
            tried = true;
            int dev = 0;
            if (const char *e = std::getenv("B200PPF_DEVICE")) dev = std::atoi(e);
            else if (const char *l = std::getenv("LOCAL_RANK")) dev = std::atoi(l);
            if (b200ppf_create(dev, &ctx) != B200PPF_OK) {
                std::fprintf(stderr, "[cv::ppf_match_3d::PPF3DDetector] cannot create a device context: %s\n", b200ppf_last_error(nullptr));
                ctx = nullptr;
            }
        }
        return ctx;
    }

    double sampling_step_, distance_step_, num_angles_;
    double position_threshold_ = -1, rotation_threshold_ = -1;
    std::shared_ptr<b200cv_detector> det_;
};

}  // namespace ppf_match_3d
}  // namespace cv
