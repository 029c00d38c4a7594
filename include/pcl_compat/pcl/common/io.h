// pcl::concatenateFields for the pair the reference joins: PointXYZ + Normal -> PointNormal
// ([PCL] common/include/pcl/common/impl/io.hpp; reference include/CloudProcessing.h:396).
#pragma once

#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {

inline void concatenateFields(const PointCloud<PointXYZ> &cloud1_in, const PointCloud<Normal> &cloud2_in,
                              PointCloud<PointNormal> &cloud_out) {
    if (cloud1_in.size() != cloud2_in.size()) {
        PCL_ERROR("[pcl::concatenateFields] The number of points in the two input datasets differs!\n");
        return;
    }
    cloud_out.points.resize(cloud1_in.size());
    cloud_out.width = cloud1_in.width;
    cloud_out.height = cloud1_in.height;
    cloud_out.is_dense = cloud1_in.is_dense && cloud2_in.is_dense;
    for (std::size_t i = 0; i < cloud1_in.size(); ++i) {
        PointNormal &o = cloud_out.points[i];
        o.x = cloud1_in[i].x; o.y = cloud1_in[i].y; o.z = cloud1_in[i].z;
        o.normal_x = cloud2_in[i].normal_x; o.normal_y = cloud2_in[i].normal_y; o.normal_z = cloud2_in[i].normal_z;
        o.curvature = cloud2_in[i].curvature;
    }
}

}  // namespace pcl
