// pcl::search::KdTree<PointT>: accepted by setSearchMethod for source compatibility
// (reference include/CloudProcessing.h:381,392).  The device operators search a uniform cell-sorted grid
// (prep.cu) and never consult it: exact k nearest neighbours are the same set whatever structure finds them.
#pragma once

#include "../pcl_macros.h"

namespace pcl {
namespace search {

template <typename PointT>
class KdTree {
public:
    using Ptr = shared_ptr<KdTree<PointT>>;
    using ConstPtr = shared_ptr<const KdTree<PointT>>;
    explicit KdTree(bool sorted = true) : sorted_(sorted) {}
    bool getSortedResults() const { return sorted_; }

private:
    bool sorted_;
};

}  // namespace search
}  // namespace pcl
