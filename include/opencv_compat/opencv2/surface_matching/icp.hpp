// cv::ppf_match_3d::ICP over libb200ppf (K6): the call the reference makes right after matching,
//     ICP icp(100, 0.005f, 2.5f, 8);  icp.registerModelToScene(models[id], pc_scene, resultsSub);
// (include/CloudProcessing.h:465-470, :518-523; opencv_contrib surface_matching/include/.../icp.hpp, src/icp.cpp).
// Clouds are the reference's N x 6 CV_32F matrices [x y z nx ny nz].  All poses of one call are refined
// concurrently in one kernel launch.  Error behaviour: a message on stderr and return value -1, poses untouched
// (no CPU fallback).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../../b200ppf.h"
#include "../core_min.hpp"
#include "pose_3d.hpp"

namespace cv {
namespace ppf_match_3d {

class ICP {
public:
    enum { ICP_SAMPLING_TYPE_UNIFORM = 0, ICP_SAMPLING_TYPE_GELFAND = 1 };

    ICP() : m_tolerance(0.005f), m_rejectionScale(2.5f), m_maxIterations(250), m_numLevels(6), m_sampleType(0), m_numNeighborsCorr(1) {}
    explicit ICP(const int iterations, const float tolerence = 0.05f, const float rejectionScale = 2.5f, const int numLevels = 6,
                 const int sampleType = ICP_SAMPLING_TYPE_UNIFORM, const int numMaxCorr = 1)
        : m_tolerance(tolerence), m_rejectionScale(rejectionScale), m_maxIterations(iterations), m_numLevels(numLevels),
          m_sampleType(sampleType), m_numNeighborsCorr(numMaxCorr) {}

    // one pose: srcPC is registered to dstPC from the identity; residual and the 4x4 pose are returned
    int registerModelToScene(const Mat &srcPC, const Mat &dstPC, double &residual, Matx44d &pose) {
        std::vector<Pose3DPtr> one(1, Pose3DPtr(new Pose3D()));
        const int rc = registerModelToScene(srcPC, dstPC, one);
        if (rc == 0) {
            residual = one[0]->residual;
            pose = one[0]->pose;
        }
        return rc;
    }

    // every pose in `poses` is refined in place (Pose3D::appendPose) and receives its residual
    int registerModelToScene(const Mat &srcPC, const Mat &dstPC, std::vector<Pose3DPtr> &poses) {
        if (poses.empty()) return 0;
        if (srcPC.cols < 6 || dstPC.cols < 6 || srcPC.empty() || dstPC.empty()) {
            std::fprintf(stderr, "[cv::ppf_match_3d::ICP] model and scene must be N x 6 CV_32F matrices with normals\n");
            return -1;
        }
        b200ppf_ctx *ctx = context();
        if (!ctx) return -1;
        b200ppf_cloud *model = nullptr, *scene = nullptr;
        int rc = b200ppf_cloud_upload(ctx, srcPC.ptr<float>(0), (size_t)srcPC.rows, (size_t)srcPC.cols, 3, &model);
        if (rc == B200PPF_OK) rc = b200ppf_cloud_upload(ctx, dstPC.ptr<float>(0), (size_t)dstPC.rows, (size_t)dstPC.cols, 3, &scene);
        std::vector<double> p16(poses.size() * 16), residuals(poses.size(), 0.0);
        for (size_t k = 0; k < poses.size(); ++k)
            for (int e = 0; e < 16; ++e) p16[16 * k + e] = poses[k]->pose.val[e];
        if (rc == B200PPF_OK) {
            const b200ppf_icp_params prm = {m_maxIterations, m_tolerance, m_rejectionScale, m_numLevels};
            rc = b200ppf_icp_refine(ctx, model, scene, &prm, p16.data(), poses.size(), residuals.data(), nullptr);
        }
        if (rc != B200PPF_OK) std::fprintf(stderr, "[cv::ppf_match_3d::ICP] %s\n", b200ppf_last_error(ctx));
        if (model) b200ppf_cloud_free(model);
        if (scene) b200ppf_cloud_free(scene);
        if (rc != B200PPF_OK) return -1;
        for (size_t k = 0; k < poses.size(); ++k) {
            poses[k]->updatePose(Matx44d(&p16[16 * k]));
            poses[k]->residual = residuals[k];
        }
        return 0;
    }

private:
    // one device context per process (B200PPF_DEVICE / LOCAL_RANK select the GPU), shared with nothing else
    static b200ppf_ctx *context() {
        static b200ppf_ctx *ctx = nullptr;
        static bool tried = false;
        if (!tried) {
            tried = true;
            int dev = 0;
            if (const char *e = std::getenv("B200PPF_DEVICE")) dev = std::atoi(e);
            else if (const char *l = std::getenv("LOCAL_RANK")) dev = std::atoi(l);
            if (b200ppf_create(dev, &ctx) != B200PPF_OK) {
                std::fprintf(stderr, "[cv::ppf_match_3d::ICP] cannot create a device context: %s\n", b200ppf_last_error(nullptr));
                ctx = nullptr;
            }
        }
        return ctx;
    }

    float m_tolerance, m_rejectionScale;
    int m_maxIterations, m_numLevels, m_sampleType, m_numNeighborsCorr;
};

}  // namespace ppf_match_3d
}  // namespace cv
