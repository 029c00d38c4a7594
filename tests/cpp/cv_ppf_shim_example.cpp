// The reference's own call sequence on its PPF engine (include/CloudProcessing.h:205-236, :428-533), compiled against
// include/opencv_compat:  PPF3DDetector detector(0.04, 0.05); detector.trainModel(pc); detector.match(scene, results, ...);
// detector.match_S2B(scene, edge, results, ...); top-N poses -> ICP(100, 0.005f, 2.5f, 8).registerModelToScene.
// usage: cv_ppf_shim_example DIR   (DIR holds model.f32, scene.f32, edge.f32: N x 6 float32 rows)
#include <cstdio>
#include <string>
#include <vector>

#include "opencv2/surface_matching/icp.hpp"
#include "opencv2/surface_matching/ppf_match_3d.hpp"

using namespace cv;
using namespace cv::ppf_match_3d;

static Mat load(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return Mat();
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    Mat m((int)(bytes / 24), 6, CV_32F);
    if (std::fread(m.ptr<float>(0), 1, (size_t)bytes, f) != (size_t)bytes) m = Mat();
    std::fclose(f);
    return m;
}

static void print(const char *tag, const std::vector<Pose3DPtr> &r, size_t k) {
    std::printf("%s %zu", tag, r.size());
    for (size_t i = 0; i < k && i < r.size(); ++i) {
        std::printf(" | %zu %zu", r[i]->numVotes, r[i]->modelIndex);
        for (int e = 0; e < 16; ++e) std::printf(" %.17g", r[i]->pose.val[e]);
    }
    std::printf("\n");
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    Mat model = load(dir + "/model.f32"), scene = load(dir + "/scene.f32"), edge = load(dir + "/edge.f32");
    if (model.empty() || scene.empty() || edge.empty()) return 3;
    PPF3DDetector detector(0.04, 0.05);
    detector.trainModel(model);
    if (!detector.trained()) return 4;
    PPF3DDetector copy = detector;  // the reference copies the detector by value on every call (CloudProcessing.h:432)
    std::vector<Pose3DPtr> results;
    copy.match(scene, results, 1.0 / 5.0, 0.04);
    print("match", results, 2);
    std::vector<Pose3DPtr> s2b;
    copy.match_S2B(scene, edge, s2b, 1.0 / 5.0, 0.04);
    print("match_S2B", s2b, 2);
    // Matching(): the N best poses go through ICP and the first one is returned
    const size_t N = results.size() < 2 ? results.size() : 2;
    std::vector<Pose3DPtr> sub(results.begin(), results.begin() + N);
    ICP icp(100, 0.005f, 2.5f, 8);
    if (icp.registerModelToScene(model, scene, sub) != 0) return 5;
    std::printf("icp %.9g", sub[0]->residual);
    for (int e = 0; e < 16; ++e) std::printf(" %.17g", sub[0]->pose.val[e]);
    std::printf("\n");
    return 0;
}
