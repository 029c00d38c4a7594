// The reference's refinement step (include/CloudProcessing.h:508-530) written against include/opencv_compat exactly as
// it is written against opencv_contrib: top poses -> ICP icp(100, 0.005f, 2.5f, 8) -> registerModelToScene -> first pose.
//
// usage: cv_icp_shim_example <dir>   with model.f32 / scene.f32 (N x 6 float32) and poses.f64 (P x 16 doubles) in <dir>;
// writes refined.f64 (P x 17 doubles: pose, residual) and prints the first pose.  Exit code 3 when the ICP failed.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <opencv2/surface_matching/icp.hpp>
#include <opencv2/surface_matching/pose_3d.hpp>

using namespace cv;
using namespace ppf_match_3d;

static Mat load6(const std::string &path) {
    std::vector<float> v;
    if (FILE *f = std::fopen(path.c_str(), "rb")) {
        float row[6];
        while (std::fread(row, sizeof(float), 6, f) == 6) v.insert(v.end(), row, row + 6);
        std::fclose(f);
    }
    Mat m((int)(v.size() / 6), 6, CV_32F);
    for (std::size_t i = 0; i < v.size(); ++i) m.ptr<float>(0)[i] = v[i];
    return m;
}

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    Mat model = load6(dir + "/model.f32"), pc_scene = load6(dir + "/scene.f32");
    std::vector<Pose3DPtr> results;
    if (FILE *f = std::fopen((dir + "/poses.f64").c_str(), "rb")) {
        double p[16];
        while (std::fread(p, sizeof(double), 16, f) == 16) {
            Pose3DPtr pose(new Pose3D(0.0, 0, 100 - results.size()));
            pose->updatePose(Matx44d(p));
            results.push_back(pose);
        }
        std::fclose(f);
    }
    if (model.empty() || pc_scene.empty() || results.empty()) {  // nothing given: a small pair keeps the binary runnable
        model = Mat(64, 6, CV_32F);
        for (int i = 0; i < 64; ++i) {
            float *r = model.ptr<float>(i);
            r[0] = 0.01f * (i % 8); r[1] = 0.01f * (i / 8); r[2] = 0.5f + 0.002f * ((i * 7) % 11);
            r[3] = 0.f; r[4] = 0.f; r[5] = -1.f;
        }
        pc_scene = model.clone();
        results.assign(1, Pose3DPtr(new Pose3D()));
    }
    // top N poses, as the reference selects them
    size_t N = 5;
    if (results.size() < N) N = results.size();
    std::vector<Pose3DPtr> resultsSub(results.begin(), results.begin() + N);

    ICP icp(100, 0.005f, 2.5f, 8);
    if (icp.registerModelToScene(model, pc_scene, resultsSub) != 0) return 3;

    if (FILE *f = std::fopen((dir + "/refined.f64").c_str(), "wb")) {
        for (size_t k = 0; k < resultsSub.size(); ++k) {
            std::fwrite(resultsSub[k]->pose.val, sizeof(double), 16, f);
            std::fwrite(&resultsSub[k]->residual, sizeof(double), 1, f);
        }
        std::fclose(f);
    }
    resultsSub[0]->printPose();
    return 0;
}
