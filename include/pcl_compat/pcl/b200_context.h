// Process-wide default device context shared by the PCL-shaped classes (one GPU per process).
// Select the GPU with the environment variable B200PPF_DEVICE (default 0; under torchrun use
// LOCAL_RANK).  There is no CPU fallback: without a usable B200 every operator reports a
// PCL_ERROR and leaves its outputs untouched, exactly as PCL does on bad input.
#pragma once

#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../b200ppf.h"
#include "pcl_macros.h"

namespace pcl {
namespace b200 {

inline b200ppf_ctx *defaultContext() {
    static b200ppf_ctx *ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (const char *e = std::getenv("B200PPF_DEVICE")) dev = std::atoi(e);
        else if (const char *l = std::getenv("LOCAL_RANK")) dev = std::atoi(l);
        if (b200ppf_create(dev, &ctx) != B200PPF_OK) {
            PCL_ERROR("[pcl::b200] cannot create a device context: %s\n", b200ppf_last_error(nullptr));
            ctx = nullptr;
        }
    });
    return ctx;
}

// Several GPUs: B200PPF_DEVICES=0,1,2,3 makes PPFRegistration::align shard the scene reference points over those
// devices of this process (b200ppf_multi_*; the first one should be the default context's device).  nullptr = one GPU.
inline b200ppf_multi *defaultMulti() {
    static b200ppf_multi *multi = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *e = std::getenv("B200PPF_DEVICES");
        if (!e) return;
        std::vector<int> dev;
        for (const char *p = e; *p;) {
            char *end = nullptr;
            long v = std::strtol(p, &end, 10);
            if (end == p) break;
            dev.push_back(static_cast<int>(v));
            p = (*end == ',') ? end + 1 : end;
        }
        if (dev.size() < 2) return;
        if (b200ppf_multi_create(dev.data(), static_cast<int>(dev.size()), &multi) != B200PPF_OK) {
            PCL_ERROR("[pcl::b200] B200PPF_DEVICES: %s\n", b200ppf_last_error(nullptr));
            multi = nullptr;
        }
    });
    return multi;
}

// RAII owners of the opaque handles
struct CloudHandle {
    b200ppf_cloud *h = nullptr;
    CloudHandle() = default;
    CloudHandle(const CloudHandle &) = delete;
    CloudHandle &operator=(const CloudHandle &) = delete;
    ~CloudHandle() { reset(); }
    void reset(b200ppf_cloud *n = nullptr) {
        if (h) b200ppf_cloud_free(h);
        h = n;
    }
};
struct FeaturesHandle {
    b200ppf_features *h = nullptr;
    FeaturesHandle() = default;
    FeaturesHandle(const FeaturesHandle &) = delete;
    FeaturesHandle &operator=(const FeaturesHandle &) = delete;
    ~FeaturesHandle() { reset(); }
    void reset(b200ppf_features *n = nullptr) {
        if (h) b200ppf_features_free(h);
        h = n;
    }
};
struct TableHandle {
    b200ppf_table *h = nullptr;
    TableHandle() = default;
    TableHandle(const TableHandle &) = delete;
    TableHandle &operator=(const TableHandle &) = delete;
    ~TableHandle() { reset(); }
    void reset(b200ppf_table *n = nullptr) {
        if (h) b200ppf_table_free(h);
        h = n;
    }
};

// device cloud -> PointCloud<PointT> (x y z of every point; the other fields keep their defaults)
template <typename CloudT>
inline bool downloadXYZ(b200ppf_ctx *ctx, const b200ppf_cloud *cloud, CloudT &output) {
    const std::size_t n = b200ppf_cloud_size(cloud);
    std::vector<float> rows(n * 3);
    if (b200ppf_cloud_download(ctx, cloud, rows.data(), 3, 0, 0) != B200PPF_OK) {
        PCL_ERROR("[pcl::b200] %s\n", b200ppf_last_error(ctx));
        return false;
    }
    output.points.resize(n);
    for (std::size_t i = 0; i < n; ++i) {
        output.points[i].x = rows[3 * i];
        output.points[i].y = rows[3 * i + 1];
        output.points[i].z = rows[3 * i + 2];
    }
    output.width = static_cast<std::uint32_t>(n);
    output.height = 1;
    output.is_dense = true;
    return true;
}

}  // namespace b200
}  // namespace pcl
