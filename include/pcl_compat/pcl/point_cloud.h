// pcl::PointCloud<T>: the subset the PPF path uses ([PCL] common/include/pcl/point_cloud.h)
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "pcl_macros.h"

namespace pcl {

template <typename PointT>
class PointCloud {
public:
    using Ptr = shared_ptr<PointCloud<PointT>>;
    using ConstPtr = shared_ptr<const PointCloud<PointT>>;
    using iterator = typename std::vector<PointT>::iterator;
    using const_iterator = typename std::vector<PointT>::const_iterator;

    std::vector<PointT> points;
    std::uint32_t width = 0;
    std::uint32_t height = 0;
    bool is_dense = true;

    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void resize(std::size_t n) {
        points.resize(n);
        if (width * height != n) {
            width = static_cast<std::uint32_t>(n);
            height = 1;
        }
    }
    void clear() {
        points.clear();
        width = height = 0;
    }
    void push_back(const PointT &p) {
        points.push_back(p);
        width = static_cast<std::uint32_t>(points.size());
        height = 1;
    }
    PointT &operator[](std::size_t i) { return points[i]; }
    const PointT &operator[](std::size_t i) const { return points[i]; }
    PointT &at(std::size_t i) { return points.at(i); }
    const PointT &at(std::size_t i) const { return points.at(i); }
    iterator begin() { return points.begin(); }
    iterator end() { return points.end(); }
    const_iterator begin() const { return points.begin(); }
    const_iterator end() const { return points.end(); }
    Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};

}  // namespace pcl
