// The canonical PCL PPF pipeline (PCL's apps/ppf_object_recognition.cpp), written
// against include/pcl_compat exactly as it would be written against PCL:
//   PPFEstimation::compute -> PPFHashMapSearch::setInputFeatureCloud -> PPFRegistration::align.
// This is also what a maintainer would put into the reference's CloudProcessor in place of
// detector.match(...) (include/CloudProcessing.h:442/495), see INTEGRATION.md.
//
// usage: pcl_shim_example <dir with bottle_1cm.f32 / scene_crop_1cm.f32>   (N x 6 float32 dumps)
// prints the final 4x4 transformation (row-major) on stdout; exit code 3 when align did not converge.
#include <cstdio>
#include <string>
#include <vector>

#include <pcl/features/ppf.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/ppf_registration.h>

static pcl::PointCloud<pcl::PointNormal>::Ptr load(const std::string &path) {
    pcl::PointCloud<pcl::PointNormal>::Ptr cloud(new pcl::PointCloud<pcl::PointNormal>());
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return cloud;
    float row[6];
    while (std::fread(row, sizeof(float), 6, f) == 6) {
        pcl::PointNormal p;
        p.x = row[0]; p.y = row[1]; p.z = row[2];
        p.normal_x = row[3]; p.normal_y = row[4]; p.normal_z = row[5];
        cloud->push_back(p);
    }
    std::fclose(f);
    return cloud;
}

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    pcl::PointCloud<pcl::PointNormal>::Ptr model = load(dir + "/bottle_1cm.f32");
    pcl::PointCloud<pcl::PointNormal>::Ptr scene = load(dir + "/scene_crop_1cm.f32");
    if (model->empty() || scene->empty()) {  // no dumps: a tiny synthetic pair keeps the binary runnable
        for (int i = 0; i < 64; ++i) {
            pcl::PointNormal p;
            p.x = 0.01f * (i % 8); p.y = 0.01f * (i / 8); p.z = 0.5f + 0.0005f * ((i * 7) % 11);
            p.normal_x = 0.f; p.normal_y = 0.f; p.normal_z = -1.f;
            model->push_back(p);
            scene->push_back(p);
        }
    }

    // train: pair features of the model, hashed
    pcl::PointCloud<pcl::PPFSignature>::Ptr cloud_model_ppf(new pcl::PointCloud<pcl::PPFSignature>());
    pcl::PPFEstimation<pcl::PointNormal, pcl::PointNormal, pcl::PPFSignature> ppf_estimator;
    ppf_estimator.setInputCloud(model);
    ppf_estimator.setInputNormals(model);
    ppf_estimator.compute(*cloud_model_ppf);

    pcl::PPFHashMapSearch::Ptr hashmap_search(new pcl::PPFHashMapSearch(12.0f / 180.0f * float(M_PI), 0.01f));
    hashmap_search->setInputFeatureCloud(cloud_model_ppf);

    // match
    pcl::PPFRegistration<pcl::PointNormal, pcl::PointNormal> ppf_registration;
    ppf_registration.setSceneReferencePointSamplingRate(5);
    ppf_registration.setPositionClusteringThreshold(0.01f);
    ppf_registration.setRotationClusteringThreshold(20.0f / 180.0f * float(M_PI));
    ppf_registration.setSearchMethod(hashmap_search);
    ppf_registration.setInputSource(model);
    ppf_registration.setInputTarget(scene);

    pcl::PointCloud<pcl::PointNormal> cloud_output;
    ppf_registration.align(cloud_output);
    if (!ppf_registration.hasConverged()) {
        std::fprintf(stderr, "align did not converge\n");
        return 3;
    }
    Eigen::Matrix4f mat = ppf_registration.getFinalTransformation();
    for (int r = 0; r < 4; ++r) std::printf("%.9g %.9g %.9g %.9g\n", mat(r, 0), mat(r, 1), mat(r, 2), mat(r, 3));
    std::printf("model_diameter %.9g features %zu output %zu candidates %zu\n", hashmap_search->getModelDiameter(),
                cloud_model_ppf->size(), cloud_output.size(), ppf_registration.getBestPoseCandidates().size());
    float f1 = (*cloud_model_ppf)[1].f1, f2 = (*cloud_model_ppf)[1].f2, f3 = (*cloud_model_ppf)[1].f3,
          f4 = (*cloud_model_ppf)[1].f4;
    std::vector<std::pair<std::size_t, std::size_t>> nn;
    hashmap_search->nearestNeighborSearch(f1, f2, f3, f4, nn);
    std::printf("bucket_of_pair_0_1 %zu first %zu %zu\n", nn.size(), nn.empty() ? 0 : nn[0].first,
                nn.empty() ? 0 : nn[0].second);
    std::printf("output0 %.9g %.9g %.9g\n", cloud_output[0].x, cloud_output[0].y, cloud_output[0].z);

    // the reference's run-time order (src/YOLO_cropping_ppf_test.cpp:117-122): LoadTrainedDetector, then match
    const std::string file = dir + "/trained.b200ppf";
    pcl::PPFHashMapSearch::Ptr loaded(new pcl::PPFHashMapSearch());
    if (!hashmap_search->saveTrained(file) || !loaded->loadTrained(file)) return 4;
    pcl::PPFRegistration<pcl::PointNormal, pcl::PointNormal> again;
    again.setSceneReferencePointSamplingRate(5);
    again.setSearchMethod(loaded);
    again.setInputSource(model);
    again.setInputTarget(scene);
    pcl::PointCloud<pcl::PointNormal> out2;
    again.align(out2);
    const bool same = again.hasConverged() && again.getFinalTransformation() == mat &&
                      loaded->getModelDiameter() == hashmap_search->getModelDiameter() &&
                      loaded->getDistanceDiscretizationStep() == 0.01f;
    std::printf("reloaded_table_same_pose %d\n", same ? 1 : 0);

    // PPFHashMapSearch::makeShared(): the copy owns its own device table and answers like the original
    pcl::PPFHashMapSearch::Ptr copy = hashmap_search->makeShared();
    hashmap_search.reset();
    std::vector<std::pair<std::size_t, std::size_t>> nn2;
    copy->nearestNeighborSearch(f1, f2, f3, f4, nn2);
    pcl::PPFRegistration<pcl::PointNormal, pcl::PointNormal> third;
    third.setSceneReferencePointSamplingRate(5);
    third.setSearchMethod(copy);
    third.setInputSource(model);
    third.setInputTarget(scene);
    pcl::PointCloud<pcl::PointNormal> out3;
    third.align(out3);
    std::printf("copied_table_same_answers %d\n", (nn2 == nn && third.hasConverged() && third.getFinalTransformation() == mat) ? 1 : 0);
    return 0;
}
