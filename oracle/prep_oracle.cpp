/*
 * prep_oracle.cpp — CPU restatement of the scene pre-processing the reference runs between its YOLO crop
 * and the PPF engine ("next" row):
 *
 *   reference include/CloudProcessing.h:359-377  Subsampling       pcl::VoxelGrid<PointXYZ>
 *   reference include/CloudProcessing.h:340-358  OutlierProcessing pcl::StatisticalOutlierRemoval<PointXYZ>(50, thr)
 *   reference include/CloudProcessing.h:378-401  NormalEstimation  pcl::NormalEstimationOMP<PointXYZ,Normal> k = 30
 *   reference include/CloudProcessing.h:402-427  EdgeExtraction    curvature > threshold
 *   reference include/CloudProcessing.h:163-190  PointCloudXYZNormalToMat (normals re-normalised)
 *
 * TEST INFRASTRUCTURE ONLY (see ppf_oracle.h).  PARITY UNPINNED: PCL is neither vendored in the reference
 * nor installed here; the arithmetic below is recalled from the upstream files named at each function
 * ([PCL] ...), float where PCL computes in float, un-fused (-ffp-contract=off).
 *
 * Choices where PCL leaves the result unspecified (the device follows the same ones):
 *   - VoxelGrid sorts (voxel, point) records with std::sort, which is not stable: the order in which a
 *     voxel's points are summed is unspecified.  Here: original point order.
 *   - FLANN returns neighbours sorted by distance; among equal distances the order is unspecified.
 *     Here: (squared distance, point index) ascending.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "ppf_oracle.h"

namespace {

/* FLANN L2_Simple<float>::operator(): result += diff*diff over x, y, z, starting from 0
 * ([FLANN] algorithms/dist.h; the metric behind pcl::KdTreeFLANN<PointXYZ>) */
inline float l2_simple(const float *a, const float *b) {
    float result = 0.0f;
    for (int k = 0; k < 3; ++k) {
        const float diff = a[k] - b[k];
        result += diff * diff;
    }
    return result;
}

struct Neighbour {
    float d2;
    uint32_t idx;
    bool operator<(const Neighbour &o) const { return d2 < o.d2 || (d2 == o.d2 && idx < o.idx); }
};

/* exact k nearest neighbours of point q (itself included), sorted by (d2, idx) */
void knn_one(const float *xyz, size_t n, size_t stride, size_t q, size_t k, std::vector<Neighbour> &heap) {
    heap.clear();
    const float *pq = xyz + q * stride;
    for (size_t j = 0; j < n; ++j) {
        Neighbour c{l2_simple(pq, xyz + j * stride), (uint32_t)j};
        if (heap.size() < k) {
            heap.push_back(c);
            std::push_heap(heap.begin(), heap.end());
        } else if (c < heap.front()) {
            std::pop_heap(heap.begin(), heap.end());
            heap.back() = c;
            std::push_heap(heap.begin(), heap.end());
        }
    }
    std::sort_heap(heap.begin(), heap.end());
}

/* [PCL] common/include/pcl/common/impl/eigen.hpp computeRoots2 */
inline void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = b * b - 4.0f * c;
    if (d < 0.0f) d = 0.0f;  // no real roots: should not happen for a symmetric matrix
    const float sd = std::sqrt(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

/* [PCL] common/include/pcl/common/impl/eigen.hpp computeRoots: eigenvalues of a symmetric 3x3, ascending */
inline void compute_roots(const float m[9], float roots[3]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    const float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
    const float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    const float c2 = m00 + m11 + m22;
    if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {  // one root is 0 -> quadratic equation
        compute_roots2(c2, c1, roots);
        return;
    }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    const float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    const float rho = std::sqrt(-a_over_3);
    const float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
    const float cos_theta = std::cos(theta);
    const float sin_theta = std::sin(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
        std::swap(roots[1], roots[2]);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f)  // eigenvalues of a positive semi-definite matrix cannot be negative: set the smallest to 0
        compute_roots2(c2, c1, roots);
}

inline void cross3(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* [PCL] common/include/pcl/common/impl/eigen.hpp eigen33(mat, eigenvalue, eigenvector): smallest eigenpair */
inline void eigen33_smallest(const float cov[9], float *eigenvalue, float evec[3]) {
    float scale = 0.0f;
    for (int k = 0; k < 9; ++k) scale = std::max(scale, std::fabs(cov[k]));
    if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
    float s[9];
    for (int k = 0; k < 9; ++k) s[k] = cov[k] / scale;
    float roots[3];
    compute_roots(s, roots);
    *eigenvalue = roots[0] * scale;
    s[0] -= roots[0];
    s[4] -= roots[0];
    s[8] -= roots[0];
    float v[3][3];
    cross3(s + 0, s + 3, v[0]);
    cross3(s + 0, s + 6, v[1]);
    cross3(s + 3, s + 6, v[2]);
    float len[3];
    for (int k = 0; k < 3; ++k) len[k] = std::sqrt(v[k][0] * v[k][0] + v[k][1] * v[k][1] + v[k][2] * v[k][2]);
    int best = 0;  // Eigen maxCoeff: the first maximum wins
    if (len[1] > len[best]) best = 1;
    if (len[2] > len[best]) best = 2;
    for (int k = 0; k < 3; ++k) evec[k] = v[best][k] / len[best];
}

}  // namespace

extern "C" {

/* [PCL] filters/include/pcl/filters/impl/voxel_grid.hpp VoxelGrid<PointT>::applyFilter (PointXYZ, no filter
 * field, min_points_per_voxel 0).  Returns the number of output points; *status = 1 when the leaf is so small
 * that the voxel index would overflow an int (PCL warns and returns the input cloud unchanged). */
size_t oracle_voxel_grid(const float *xyz, size_t n, size_t stride, const float *leaf3, float *out_xyz, int *status) {
    if (status) *status = 0;
    if (n == 0) return 0;
    float inv[3];
    for (int k = 0; k < 3; ++k) inv[k] = 1.0f / leaf3[k];  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
    float min_p[3], max_p[3];  // getMinMax3D
    for (int k = 0; k < 3; ++k) min_p[k] = std::numeric_limits<float>::max(), max_p[k] = -std::numeric_limits<float>::max();
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            min_p[k] = std::min(min_p[k], xyz[i * stride + k]);
            max_p[k] = std::max(max_p[k], xyz[i * stride + k]);
        }
    int64_t d[3];
    for (int k = 0; k < 3; ++k) d[k] = (int64_t)((max_p[k] - min_p[k]) * inv[k]) + 1;
    if (d[0] * d[1] * d[2] > (int64_t)std::numeric_limits<int32_t>::max()) {
        if (status) *status = 1;
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) out_xyz[i * 3 + k] = xyz[i * stride + k];
        return n;
    }
    int min_b[3], max_b[3], div_b[3], mul[3];
    for (int k = 0; k < 3; ++k) {
        min_b[k] = (int)std::floor(min_p[k] * inv[k]);
        max_b[k] = (int)std::floor(max_p[k] * inv[k]);
        div_b[k] = max_b[k] - min_b[k] + 1;
    }
    mul[0] = 1, mul[1] = div_b[0], mul[2] = div_b[0] * div_b[1];
    std::vector<std::pair<uint32_t, uint32_t>> rec(n);  // (voxel, point)
    for (size_t i = 0; i < n; ++i) {
        const float *p = xyz + i * stride;
        const int ijk0 = (int)(std::floor(p[0] * inv[0]) - (float)min_b[0]);
        const int ijk1 = (int)(std::floor(p[1] * inv[1]) - (float)min_b[1]);
        const int ijk2 = (int)(std::floor(p[2] * inv[2]) - (float)min_b[2]);
        rec[i] = {(uint32_t)(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]), (uint32_t)i};
    }
    std::stable_sort(rec.begin(), rec.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    size_t m = 0;
    for (size_t first = 0; first < n;) {
        size_t last = first;
        float c[3] = {0.0f, 0.0f, 0.0f};  // CentroidPoint / AccumulatorXYZ: Vector3f sum, then / n
        while (last < n && rec[last].first == rec[first].first) {
            const float *p = xyz + (size_t)rec[last].second * stride;
            c[0] += p[0], c[1] += p[1], c[2] += p[2];
            ++last;
        }
        const float cnt = (float)(last - first);
        out_xyz[m * 3 + 0] = c[0] / cnt;
        out_xyz[m * 3 + 1] = c[1] / cnt;
        out_xyz[m * 3 + 2] = c[2] / cnt;
        ++m;
        first = last;
    }
    return m;
}

/* exact k nearest neighbours of every point (itself included): idx / d2 are n*k, row-major, each row sorted
 * by (squared distance, index).  k is clamped to n by the caller. */
void oracle_knn(const float *xyz, size_t n, size_t stride, int k, uint32_t *idx, float *d2, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        std::vector<Neighbour> heap;
        heap.reserve((size_t)k + 1);
#pragma omp for schedule(dynamic, 64)
        for (int64_t q = 0; q < (int64_t)n; ++q) {
            knn_one(xyz, n, stride, (size_t)q, (size_t)k, heap);
            for (size_t j = 0; j < heap.size(); ++j) {
                if (idx) idx[(size_t)q * k + j] = heap[j].idx;
                if (d2) d2[(size_t)q * k + j] = heap[j].d2;
            }
        }
    }
}

/* [PCL] filters/include/pcl/filters/impl/statistical_outlier_removal.hpp applyFilterIndices (negative_ = false).
 * distances[n] (optional) = mean distance to the mean_k nearest other points; keep[n] = 1 for the points that
 * stay; *threshold = mean + std_mul * stddev.  Requires n > mean_k (PCL reads past the neighbour list otherwise).
 * Returns the number of points kept. */
size_t oracle_sor(const float *xyz, size_t n, size_t stride, int mean_k, double std_mul, float *distances, uint8_t *keep,
                  double *threshold, int n_threads) {
    if (n == 0 || (size_t)mean_k + 1 > n) return 0;
    std::vector<float> dist(n);
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        std::vector<Neighbour> heap;
        heap.reserve((size_t)mean_k + 2);
#pragma omp for schedule(dynamic, 64)
        for (int64_t q = 0; q < (int64_t)n; ++q) {
            knn_one(xyz, n, stride, (size_t)q, (size_t)mean_k + 1, heap);
            double dist_sum = 0.0;
            for (int k = 1; k < mean_k + 1; ++k) dist_sum += std::sqrt(heap[k].d2);  // k = 0 is the query point
            dist[q] = (float)(dist_sum / mean_k);
        }
    }
    double sum = 0.0, sq_sum = 0.0;
    for (size_t i = 0; i < n; ++i) {
        sum += dist[i];
        sq_sum += dist[i] * dist[i];  // float product, as PCL writes it
    }
    const double mean = sum / (double)n;
    const double variance = (sq_sum - sum * sum / (double)n) / ((double)n - 1.0);
    const double stddev = std::sqrt(variance);
    const double thr = mean + std_mul * stddev;
    if (threshold) *threshold = thr;
    size_t kept = 0;
    for (size_t i = 0; i < n; ++i) {
        const bool out = dist[i] > thr;
        if (keep) keep[i] = out ? 0 : 1;
        kept += out ? 0 : 1;
        if (distances) distances[i] = dist[i];
    }
    return kept;
}

/* [PCL] features/include/pcl/features/impl/normal_3d_omp.hpp computeFeature ->
 * features/include/pcl/features/normal_3d.h computePointNormal / flipNormalTowardsViewpoint ->
 * common/include/pcl/common/impl/centroid.hpp computeMeanAndCovarianceMatrix (float) ->
 * features/include/pcl/features/impl/feature.hpp solvePlaneParameters -> common/impl/eigen.hpp eigen33.
 * cov_mode 0 = PCL >= 1.12 (sums shifted by the first neighbour), 1 = PCL 1.8-1.11 (raw sums).
 * out4[n*4] = {nx, ny, nz, curvature}; NaN when fewer than 3 neighbours exist. */
void oracle_normals(const float *xyz, size_t n, size_t stride, int k, const float *viewpoint3, int cov_mode, float *out4,
                    int n_threads) {
    const size_t kk = std::min<size_t>((size_t)std::max(k, 0), n);
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        std::vector<Neighbour> heap;
        heap.reserve(kk + 1);
#pragma omp for schedule(dynamic, 64)
        for (int64_t q = 0; q < (int64_t)n; ++q) {
            float *o = out4 + (size_t)q * 4;
            if (kk < 3) {
                o[0] = o[1] = o[2] = o[3] = std::numeric_limits<float>::quiet_NaN();
                continue;
            }
            knn_one(xyz, n, stride, (size_t)q, kk, heap);
            float K[3] = {0.0f, 0.0f, 0.0f};
            if (cov_mode == 0) {
                const float *p0 = xyz + (size_t)heap[0].idx * stride;
                K[0] = p0[0], K[1] = p0[1], K[2] = p0[2];
            }
            float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (size_t j = 0; j < kk; ++j) {
                const float *p = xyz + (size_t)heap[j].idx * stride;
                const float x = p[0] - K[0], y = p[1] - K[1], z = p[2] - K[2];
                accu[0] += x * x;
                accu[1] += x * y;
                accu[2] += x * z;
                accu[3] += y * y;
                accu[4] += y * z;
                accu[5] += z * z;
                accu[6] += x;
                accu[7] += y;
                accu[8] += z;
            }
            const float cnt = (float)kk;
            for (int a = 0; a < 9; ++a) accu[a] /= cnt;
            float cov[9];
            cov[0] = accu[0] - accu[6] * accu[6];
            cov[1] = accu[1] - accu[6] * accu[7];
            cov[2] = accu[2] - accu[6] * accu[8];
            cov[4] = accu[3] - accu[7] * accu[7];
            cov[5] = accu[4] - accu[7] * accu[8];
            cov[8] = accu[5] - accu[8] * accu[8];
            cov[3] = cov[1], cov[6] = cov[2], cov[7] = cov[5];
            float ev, nrm[3];
            eigen33_smallest(cov, &ev, nrm);
            const float eig_sum = cov[0] + cov[4] + cov[8];
            float curvature = 0.0f;
            if (eig_sum != 0.0f) curvature = std::fabs(ev / eig_sum);
            // flipNormalTowardsViewpoint
            const float *p = xyz + (size_t)q * stride;
            const float vx = viewpoint3[0] - p[0], vy = viewpoint3[1] - p[1], vz = viewpoint3[2] - p[2];
            const float cos_theta = vx * nrm[0] + vy * nrm[1] + vz * nrm[2];
            if (cos_theta < 0.0f) nrm[0] *= -1.0f, nrm[1] *= -1.0f, nrm[2] *= -1.0f;
            o[0] = nrm[0], o[1] = nrm[1], o[2] = nrm[2], o[3] = curvature;
        }
    }
}

/* reference include/CloudProcessing.h:270-300 (SceneCropping, per box) + include/Camera.h:50-61
 * (back_projection_bbox): the box grown by 30 pixels and clamped, the mean of the depth image at its four
 * corners, the four corner rays at that depth, pushed 0.15 m further.  corners12 = left_top, left_bot,
 * right_top, right_bot (the order the reference pushes them into its hull cloud). */
void oracle_frustum_corners(const float *depth, int rows, int cols, int bx, int by, int bw, int bh, double fx, double fy,
                            double ppx, double ppy, float *corners12) {
    double left = bx - 30;
    if (left < 0) left = 0;
    double top = by - 30;
    if (top < 0) top = 0;
    double right = bx + bw + 30;
    if (right >= cols) right = cols - 1;
    double bottom = by + bh + 30;
    if (bottom >= rows) bottom = rows - 1;
    const float depth_1 = depth[(size_t)(int)top * cols + (int)left];
    const float depth_2 = depth[(size_t)(int)top * cols + (int)right];
    const float depth_3 = depth[(size_t)(int)bottom * cols + (int)left];
    const float depth_4 = depth[(size_t)(int)bottom * cols + (int)right];
    const float depth_avg = (depth_1 + depth_2 + depth_3 + depth_4) / 4;
    const int us[4] = {(int)left, (int)left, (int)right, (int)right};
    const int vs[4] = {(int)top, (int)bottom, (int)top, (int)bottom};
    for (int c = 0; c < 4; ++c) {
        const float z = depth_avg;
        float x = (float)((float)(us[c] - ppx) * z / fx);
        float y = (float)((float)(vs[c] - ppy) * z / fy);
        float zz = z;
        zz += 0.15;  // float += double literal, as the reference writes it
        corners12[3 * c + 0] = x, corners12[3 * c + 1] = y, corners12[3 * c + 2] = zz;
    }
}

/* reference include/CloudProcessing.h:312-332: ConvexHull of {four corners, origin} + CropHull (dim 3) = the points
 * inside that pyramid.  PCL decides by casting rays at the hull's triangles; here the same set is formed another
 * way (the device uses five half-spaces): p is inside iff the ray from the apex through p meets the base plane at
 * or beyond p, inside the base quadrilateral.  All in double; a point within rounding of a face may fall either
 * way in any of the three formulations.  keep[n] = 1 inside; returns the count. */
size_t oracle_crop_pyramid(const float *xyz, size_t n, size_t stride, const float *corners12, uint8_t *keep) {
    double c[4][3];
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 3; ++k) c[i][k] = corners12[3 * i + k];
    const int cyc[4] = {0, 1, 3, 2};  // left_top -> left_bot -> right_bot -> right_top
    double e1[3], e2[3], nrm[3];
    for (int k = 0; k < 3; ++k) e1[k] = c[1][k] - c[0][k], e2[k] = c[2][k] - c[0][k];
    nrm[0] = e1[1] * e2[2] - e1[2] * e2[1];
    nrm[1] = e1[2] * e2[0] - e1[0] * e2[2];
    nrm[2] = e1[0] * e2[1] - e1[1] * e2[0];
    const double d = nrm[0] * c[0][0] + nrm[1] * c[0][1] + nrm[2] * c[0][2];
    size_t kept = 0;
    for (size_t i = 0; i < n; ++i) {
        const double p[3] = {xyz[i * stride], xyz[i * stride + 1], xyz[i * stride + 2]};
        const double np_ = nrm[0] * p[0] + nrm[1] * p[1] + nrm[2] * p[2];
        bool in = false;
        if (p[0] == 0.0 && p[1] == 0.0 && p[2] == 0.0) {
            in = true;  // the apex
        } else if (np_ != 0.0 && d / np_ >= 1.0) {  // t * p lies in the base plane at t = d / (n.p) >= 1
            const double t = d / np_;
            const double h[3] = {t * p[0], t * p[1], t * p[2]};
            int pos = 0, neg = 0;
            for (int s = 0; s < 4; ++s) {  // h inside the convex quadrilateral: the same side of its four edges
                const double *a = c[cyc[s]], *b = c[cyc[(s + 1) & 3]];
                const double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ah[3] = {h[0] - a[0], h[1] - a[1], h[2] - a[2]};
                const double cr[3] = {ab[1] * ah[2] - ab[2] * ah[1], ab[2] * ah[0] - ab[0] * ah[2], ab[0] * ah[1] - ab[1] * ah[0]};
                const double sgn = cr[0] * nrm[0] + cr[1] * nrm[1] + cr[2] * nrm[2];
                if (sgn > 0) ++pos;
                if (sgn < 0) ++neg;
            }
            in = (pos == 0 || neg == 0);
        }
        keep[i] = in ? 1 : 0;
        kept += in ? 1 : 0;
    }
    return kept;
}

/* reference include/CloudProcessing.h:181-186: double A = sqrt(nx*nx + ny*ny + nz*nz) — a float sum, and under the
 * file's `using namespace std` the call resolves to the float overload, widened to double afterwards; if
 * A > 0.00001 the three components are divided by (float)A.  In place on n rows of `stride` floats. */
void oracle_renormalize_normals(float *nrm, size_t n, size_t stride) {
    for (size_t i = 0; i < n; ++i) {
        float *d = nrm + i * stride;
        const double A = (double)std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (A > 0.00001) {
            d[0] /= (float)A;
            d[1] /= (float)A;
            d[2] /= (float)A;
        }
    }
}

}  // extern "C"
