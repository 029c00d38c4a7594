// pcl_compat — minimal PCL-shaped surface over libb200ppf (include/b200ppf.h).
// Only what the PPF path touches is provided; names, defaults and error behaviour follow PCL
// (SURVEY.md Appendix A.7).  With real Eigen on the include path its types are used, otherwise
// the few Eigen types the PPF API exposes are provided by Eigen_min.h.
#pragma once

#include <cstdarg>
#include <cstdio>
#include <memory>

#ifndef PCL_ERROR
#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#endif
#ifndef PCL_WARN
#define PCL_WARN(...) std::fprintf(stderr, __VA_ARGS__)
#endif
#ifndef PCL_INFO
#define PCL_INFO(...) ((void)0)
#endif

#if defined(__has_include)
#if __has_include(<Eigen/Geometry>)
#include <Eigen/Geometry>
#define PCL_COMPAT_HAVE_EIGEN 1
#endif
#endif
#ifndef PCL_COMPAT_HAVE_EIGEN
#include "Eigen_min.h"
#endif

namespace pcl {
template <typename T>
using shared_ptr = std::shared_ptr<T>;  // PCL >= 1.11
}
