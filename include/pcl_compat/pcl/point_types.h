// pcl::PointNormal / pcl::PPFSignature with PCL's memory layout
// ([PCL] common/include/pcl/impl/point_types.hpp).
#pragma once

#include "pcl_macros.h"

namespace pcl {

struct alignas(16) PointXYZ {
    union {
        float data[4];
        struct { float x, y, z; };
    };
    PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
    PointXYZ(float x_, float y_, float z_) : data{x_, y_, z_, 1.f} {}
};

// pcl::Normal: 32 bytes, data_n[4] = {nx,ny,nz,0}, curvature, pad
struct alignas(16) Normal {
    union {
        float data_n[4];
        float normal[3];
        struct { float normal_x, normal_y, normal_z; };
    };
    union {
        struct { float curvature; };
        float data_c[4];
    };
    Normal() : data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
static_assert(sizeof(Normal) == 32, "pcl::Normal is 32 bytes");

// 48 bytes, 16-byte aligned: data[4] = {x,y,z,1}, data_n[4] = {nx,ny,nz,0}, curvature, pad
struct alignas(16) PointNormal {
    union {
        float data[4];
        struct { float x, y, z; };
    };
    union {
        float data_n[4];
        float normal[3];
        struct { float normal_x, normal_y, normal_z; };
    };
    union {
        struct { float curvature; };
        float data_c[4];
    };
    PointNormal() : data{0.f, 0.f, 0.f, 1.f}, data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
static_assert(sizeof(PointNormal) == 48, "pcl::PointNormal is 48 bytes");

// 20 bytes, unaligned
struct PPFSignature {
    float f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f, alpha_m = 0.f;
};
static_assert(sizeof(PPFSignature) == 20, "pcl::PPFSignature is 20 bytes");

}  // namespace pcl
