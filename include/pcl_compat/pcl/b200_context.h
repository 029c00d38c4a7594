// Process-wide default device context shared by the PCL-shaped classes (one GPU per process).
// Select the GPU with the environment variable B200PPF_DEVICE (default 0; under torchrun use
// LOCAL_RANK).  There is no CPU fallback: without a usable B200 every operator reports a
// PCL_ERROR and leaves its outputs untouched, exactly as PCL does on bad input.
#pragma once

#include <cstdlib>
#include <mutex>

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

}  // namespace b200
}  // namespace pcl
