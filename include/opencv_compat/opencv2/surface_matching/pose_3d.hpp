// cv::ppf_match_3d::Pose3D — the pose record the reference passes between its PPF match and its ICP
// (include/CloudProcessing.h:436-477, :497-530; opencv_contrib surface_matching/include/.../pose_3d.hpp).
// Fields and methods the reference touches: pose, residual, numVotes, alpha, modelIndex, angle, t, updatePose,
// appendPose, printPose, clone.  The quaternion member of the original is not carried.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <vector>

#include "../core_min.hpp"

namespace cv {
namespace ppf_match_3d {

class Pose3D;
typedef Ptr<Pose3D> Pose3DPtr;

class Pose3D {
public:
    Pose3D() : alpha(0), residual(0), modelIndex(0), numVotes(0), pose(Matx44d::eye()), angle(0), t(0, 0, 0) {}
    Pose3D(double Alpha, std::size_t ModelIndex = 0, std::size_t NumVotes = 0)
        : alpha(Alpha), residual(0), modelIndex(ModelIndex), numVotes(NumVotes), pose(Matx44d::eye()), angle(0), t(0, 0, 0) {}

    // new 4x4 pose: rotation angle from the trace as the original forms it (0 / pi at the ends), translation column
    void updatePose(const Matx44d &NewPose) {
        pose = NewPose;
        const double trace = pose(0, 0) + pose(1, 1) + pose(2, 2);
        const double eps = 1.192092896e-07;
        if (std::fabs(trace - 3) <= eps) angle = 0;
        else if (std::fabs(trace + 1) <= eps) angle = 3.14159265358979323846;
        else angle = std::acos((trace - 1) / 2);
        t = Vec3d(pose(0, 3), pose(1, 3), pose(2, 3));
    }
    // pose = IncrementalPose * pose (what ICP::registerModelToScene does with its result)
    void appendPose(const Matx44d &IncrementalPose) { updatePose(IncrementalPose * pose); }
    void printPose() const {
        std::printf("\n-- Pose to Model Index %d: NumVotes = %d, Residual = %f\n", (int)modelIndex, (int)numVotes, residual);
        for (int r = 0; r < 4; ++r) std::printf("[%.9g, %.9g, %.9g, %.9g]\n", pose(r, 0), pose(r, 1), pose(r, 2), pose(r, 3));
    }
    Pose3DPtr clone() const { return Pose3DPtr(new Pose3D(*this)); }

    double alpha, residual;
    std::size_t modelIndex, numVotes;
    Matx44d pose;
    double angle;
    Vec3d t;
};

}  // namespace ppf_match_3d
}  // namespace cv
