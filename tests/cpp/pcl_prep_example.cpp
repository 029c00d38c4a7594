// The reference's per-object pre-processing (src/YOLO_cropping_ppf_test.cpp:96-103,120-121:
// Subsampling -> OutlierProcessing -> NormalEstimation -> EdgeExtraction -> PointCloudXYZNormalToMat), written
// against include/pcl_compat exactly as include/CloudProcessing.h:340-427 writes it against PCL — every operator
// below runs on the B200 through libb200ppf.  See INTEGRATION.md.
//
// usage: pcl_prep_example <scene_crop_raw.f32 (N x 3 float32)> [leaf = 0.01] [outlier threshold = 1.0]
// prints the stage sizes, the first rows of the N x 6 matrix handed to the PPF engine and a checksum;
// exit code 3 when a stage produced nothing (no GPU: PCL-style error messages on stderr, no fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <pcl/common/io.h>
#include <pcl/features/normal_3d_omp.h>
#include <pcl/filters/statistical_outlier_removal.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/search/kdtree.h>

int main(int argc, char **argv) {
    const std::string path = argc > 1 ? argv[1] : "scene_crop_raw.f32";
    const double leafsize = argc > 2 ? std::atof(argv[2]) : 0.01;
    const double thresh = argc > 3 ? std::atof(argv[3]) : 1.0;

    pcl::PointCloud<pcl::PointXYZ>::Ptr object(new pcl::PointCloud<pcl::PointXYZ>());
    if (FILE *f = std::fopen(path.c_str(), "rb")) {
        float row[3];
        while (std::fread(row, sizeof(float), 3, f) == 3) object->push_back(pcl::PointXYZ(row[0], row[1], row[2]));
        std::fclose(f);
    }
    if (object->empty())  // no dump given: a small wavy sheet keeps the binary runnable
        for (int i = 0; i < 4096; ++i)
            object->push_back(pcl::PointXYZ(0.002f * (i % 64), 0.002f * (i / 64), 0.6f + 0.004f * std::sin(0.3f * (i % 64))));

    // Subsampling (CloudProcessing.h:359-377)
    pcl::PointCloud<pcl::PointXYZ>::Ptr sampled(new pcl::PointCloud<pcl::PointXYZ>());
    pcl::VoxelGrid<pcl::PointXYZ> sub;
    const float lf = static_cast<float>(leafsize);
    Eigen::Vector4f leaf(lf, lf, lf, lf);
    sub.setInputCloud(object);
    sub.setLeafSize(leaf);
    sub.filter(*sampled);

    // OutlierProcessing (CloudProcessing.h:340-358)
    pcl::PointCloud<pcl::PointXYZ>::Ptr filtered(new pcl::PointCloud<pcl::PointXYZ>());
    pcl::StatisticalOutlierRemoval<pcl::PointXYZ> sor;
    sor.setInputCloud(sampled);
    sor.setMeanK(50);
    sor.setStddevMulThresh(thresh);
    sor.filter(*filtered);

    // NormalEstimation (CloudProcessing.h:378-401)
    pcl::NormalEstimationOMP<pcl::PointXYZ, pcl::Normal> ne;
    pcl::search::KdTree<pcl::PointXYZ>::Ptr tree(new pcl::search::KdTree<pcl::PointXYZ>());
    pcl::PointCloud<pcl::Normal>::Ptr normals(new pcl::PointCloud<pcl::Normal>());
    pcl::PointCloud<pcl::PointNormal>::Ptr with_normals(new pcl::PointCloud<pcl::PointNormal>());
    ne.setInputCloud(filtered);
    ne.setNumberOfThreads(12);
    ne.setSearchMethod(tree);
    ne.setKSearch(30);
    ne.compute(*normals);
    pcl::concatenateFields(*filtered, *normals, *with_normals);

    // EdgeExtraction (CloudProcessing.h:402-427)
    pcl::PointCloud<pcl::PointNormal> edges;
    for (const pcl::PointNormal &p : *with_normals)
        if (p.curvature > 0.03f) edges.push_back(p);

    std::printf("sizes %zu %zu %zu %zu %zu\n", object->size(), sampled->size(), filtered->size(), normals->size(), edges.size());
    if (sampled->empty() || filtered->empty() || normals->size() != filtered->size()) return 3;

    // PointCloudXYZNormalToMat (CloudProcessing.h:163-190): N x 6 float rows, normals re-normalised
    std::vector<float> mat(with_normals->size() * 6);
    double checksum = 0.0;
    for (std::size_t i = 0; i < with_normals->size(); ++i) {
        const pcl::PointNormal &p = (*with_normals)[i];
        float *d = &mat[6 * i];
        d[0] = p.x; d[1] = p.y; d[2] = p.z;
        d[3] = p.normal_x; d[4] = p.normal_y; d[5] = p.normal_z;
        const double A = std::sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
        if (A > 0.00001) { d[3] /= float(A); d[4] /= float(A); d[5] /= float(A); }
        for (int k = 0; k < 6; ++k) checksum += d[k];
    }
    for (std::size_t i = 0; i < 3 && i < with_normals->size(); ++i)
        std::printf("row%zu %.9g %.9g %.9g %.9g %.9g %.9g\n", i, mat[6 * i], mat[6 * i + 1], mat[6 * i + 2], mat[6 * i + 3],
                    mat[6 * i + 4], mat[6 * i + 5]);
    std::printf("checksum %.17g\n", checksum);
    return 0;
}
