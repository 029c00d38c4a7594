/*
 * icp_oracle.cpp — CPU restatement of the ICP refinement the reference runs right after PPF matching
 * (reference include/CloudProcessing.h:465-470 and :518-523:
 *      ICP icp(100, 0.005f, 2.5f, 8);  icp.registerModelToScene(models[id], pc_scene, resultsSub);)
 *
 * TEST INFRASTRUCTURE ONLY (see ppf_oracle.h).  PARITY UNPINNED: the algorithm lives in opencv_contrib
 * modules/surface_matching/src/icp.cpp (+ ppf_helpers.cpp, c_utils.hpp), which is neither vendored in
 * the reference nor installed here; it is restated from the published source as recalled in SURVEY.md
 * Appendix B ("picky" point-to-plane ICP: Birdal & Ilic):
 *   - both clouds are centred on the mean of their two centroids and scaled by n / mean distance to origin;
 *   - coarse-to-fine over num_levels: level L uses every round(n / round(n / 2^L))-th point of both clouds,
 *     tolerance * (L+1)^2 and max_iterations / (L+1) iterations;
 *   - one iteration: nearest scene sample of every (moved) model sample (squared distances, as FLANN
 *     reports them); robust rejection at median + scale * 1.48257968 * MAD; of several model samples that
 *     chose the same scene sample the closest survives; point-to-plane least squares (minimum-norm solution, as
 *     cv::solve(DECOMP_SVD) gives it), linearised about the level's un-moved samples, gives (roll, pitch, yaw, t)
 *     -> PoseX = [Rx*Ry*Rz | t];
 *     error = || [Src_Match - Dst_Match] ||_F over all six columns / samples; stop when its ratio to the
 *     previous error is within 1 +- tolerance;
 *   - pose <- PoseX * pose per level; the normalisation is undone at the end.
 * Clouds are float32 N x 6, transforms are evaluated in double and stored back as float32, exactly
 * where the cv::Mat types of the original force it.
 */
#include "ppf_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Pose {
    double m[16];  // row-major 4x4
};

Pose identity() {
    Pose p;
    std::memset(p.m, 0, sizeof(p.m));
    p.m[0] = p.m[5] = p.m[10] = p.m[15] = 1.0;
    return p;
}

Pose mul(const Pose &a, const Pose &b) {
    Pose c;
    for (int r = 0; r < 4; ++r)
        for (int k = 0; k < 4; ++k) {
            double s = 0.0;
            for (int q = 0; q < 4; ++q) s += a.m[r * 4 + q] * b.m[q * 4 + k];
            c.m[r * 4 + k] = s;
        }
    return c;
}

/* transformPCPose: xyz -> R p + t, normals -> normalised R n; float32 in, double maths, float32 out */
std::vector<float> transform_cloud(const std::vector<float> &pc, const Pose &P) {
    const size_t n = pc.size() / 6;
    std::vector<float> out(pc.size());
    for (size_t i = 0; i < n; ++i) {
        const float *s = &pc[6 * i];
        float *d = &out[6 * i];
        const double x = s[0], y = s[1], z = s[2];
        double w = P.m[12] * x + P.m[13] * y + P.m[14] * z + P.m[15];
        if (w == 0.0) w = 1.0;
        d[0] = (float)((P.m[0] * x + P.m[1] * y + P.m[2] * z + P.m[3]) / w);
        d[1] = (float)((P.m[4] * x + P.m[5] * y + P.m[6] * z + P.m[7]) / w);
        d[2] = (float)((P.m[8] * x + P.m[9] * y + P.m[10] * z + P.m[11]) / w);
        const double nx = s[3], ny = s[4], nz = s[5];
        double rx = P.m[0] * nx + P.m[1] * ny + P.m[2] * nz;
        double ry = P.m[4] * nx + P.m[5] * ny + P.m[6] * nz;
        double rz = P.m[8] * nx + P.m[9] * ny + P.m[10] * nz;
        const double len = std::sqrt(rx * rx + ry * ry + rz * rz);
        if (len > 1e-12) {
            rx /= len;
            ry /= len;
            rz /= len;
        }
        d[3] = (float)rx;
        d[4] = (float)ry;
        d[5] = (float)rz;
    }
    return out;
}

std::vector<float> sample_uniform(const std::vector<float> &pc, int step) {
    const size_t n = pc.size() / 6;
    std::vector<float> out;
    if (step < 1) step = 1;
    for (size_t i = 0; i < n; i += (size_t)step) out.insert(out.end(), pc.begin() + 6 * i, pc.begin() + 6 * i + 6);
    return out;
}

/* exact nearest neighbour through a uniform grid with growing shells (FLANN's single kd-tree is exact too) */
struct Grid {
    const std::vector<float> *pc = nullptr;
    double lo[3] = {0, 0, 0}, cell = 1.0;
    int dim[3] = {1, 1, 1};
    std::vector<uint32_t> start, items;

    void build(const std::vector<float> &cloud) {
        pc = &cloud;
        const size_t n = cloud.size() / 6;
        double hi[3] = {-1e300, -1e300, -1e300};
        lo[0] = lo[1] = lo[2] = 1e300;
        for (size_t i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) {
                lo[c] = std::min(lo[c], (double)cloud[6 * i + c]);
                hi[c] = std::max(hi[c], (double)cloud[6 * i + c]);
            }
        if (n == 0) lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.0;
        const double ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-9});
        const double target = std::cbrt((double)std::max<size_t>(n, 1) / 2.0);  // ~2 points per cell along a surface-ish cloud
        cell = ext / std::max(1.0, std::min(target, 128.0));
        for (int c = 0; c < 3; ++c) dim[c] = std::max(1, (int)std::floor((hi[c] - lo[c]) / cell) + 1);
        const size_t cells = (size_t)dim[0] * dim[1] * dim[2];
        start.assign(cells + 1, 0);
        std::vector<uint32_t> cid(n);
        for (size_t i = 0; i < n; ++i) {
            cid[i] = index(coord(cloud[6 * i], 0), coord(cloud[6 * i + 1], 1), coord(cloud[6 * i + 2], 2));
            ++start[cid[i] + 1];
        }
        for (size_t c = 0; c < cells; ++c) start[c + 1] += start[c];
        items.resize(n);
        std::vector<uint32_t> fill(start.begin(), start.end() - 1);
        for (size_t i = 0; i < n; ++i) items[fill[cid[i]]++] = (uint32_t)i;
    }
    int coord(double v, int c) const {
        int k = (int)std::floor((v - lo[c]) / cell);
        return k < 0 ? 0 : (k >= dim[c] ? dim[c] - 1 : k);
    }
    uint32_t index(int x, int y, int z) const { return (uint32_t)(((size_t)z * dim[1] + y) * dim[0] + x); }

    /* nearest point (ties: lowest index), squared distance in float as FLANN's L2 functor accumulates it */
    void nearest(const float *q, int &best, float &best_d2) const {
        best = -1;
        best_d2 = std::numeric_limits<float>::max();
        const int cx = coord(q[0], 0), cy = coord(q[1], 1), cz = coord(q[2], 2);
        const int max_r = std::max({dim[0], dim[1], dim[2]});
        for (int r = 0; r <= max_r; ++r) {
            if (best >= 0) {
                // every unvisited cell lies at least (r - 1) * cell away from q along one axis (q may sit
                // outside the grid: then the clamped cell only makes the bound more conservative)
                const double reach = (double)(r - 1) * cell;
                if (reach > 0 && reach * reach > (double)best_d2) break;
            }
            for (int z = cz - r; z <= cz + r; ++z) {
                if (z < 0 || z >= dim[2]) continue;
                for (int y = cy - r; y <= cy + r; ++y) {
                    if (y < 0 || y >= dim[1]) continue;
                    const bool shell_zy = (z == cz - r || z == cz + r || y == cy - r || y == cy + r);
                    for (int x = cx - r; x <= cx + r; ++x) {
                        if (x < 0 || x >= dim[0]) continue;
                        if (!shell_zy && x != cx - r && x != cx + r) continue;  // interior: visited at a smaller r
                        const uint32_t c = index(x, y, z);
                        for (uint32_t s = start[c]; s < start[c + 1]; ++s) {
                            const uint32_t j = items[s];
                            const float *p = &(*pc)[6 * (size_t)j];
                            const float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
                            const float d2 = dx * dx + dy * dy + dz * dz;
                            if (d2 < best_d2 || (d2 == best_d2 && (int)j < best)) {
                                best_d2 = d2;
                                best = (int)j;
                            }
                        }
                    }
                }
            }
        }
    }
};

/* lower median by selection, as the quick-select of the original returns it */
float median_of(std::vector<float> v) {
    if (v.empty()) return 0.0f;
    const size_t k = (v.size() - 1) / 2;
    std::nth_element(v.begin(), v.begin() + k, v.end());
    return v[k];
}

float rejection_threshold(const std::vector<float> &r, float scale) {
    const float med = median_of(r);
    std::vector<float> t(r.size());
    for (size_t i = 0; i < r.size(); ++i) t[i] = (float)std::fabs((double)r[i] - (double)med);
    const float s = 1.48257968f * median_of(t);
    return scale * s + med;
}

/* The original solves the n x 6 system A x = b with cv::solve(DECOMP_SVD): the least-squares solution of MINIMUM NORM, which
 * is what keeps the coarsest pyramid levels of a small cloud sane — with fewer than six surviving correspondences (the
 * reference's own 681-point object has 4-5 samples at level 7) the system is rank-deficient, and an elimination of the
 * normal equations divides by rounding noise and sends the pose metres away.  The same solution from the 6 x 6 Gram matrix
 * N = A^T A, g = A^T b:  N = V diag(l) V^T by cyclic Jacobi rotations,  x = sum over l_i > cut of v_i (v_i . g) / l_i.
 * cut = 1e-12 * l_max: singular values below 1e-6 of the largest count as zero (the SVD of A itself would resolve them
 * down to ~1e-15; a Gram matrix cannot, and a direction that weak is unobservable in float32 point data anyway). */
/* the 15 index pairs of a sweep as 5 rounds of 3 disjoint pairs: the three rotations of a round touch different rows and
 * columns, so the device evaluates their (long, dependent) divide / square-root chains side by side.  Columns of all three
 * first, then rows of all three — in that order here too, so that both sides round alike. */
static const int JACOBI_ROUNDS[5][3][2] = {{{0, 5}, {1, 4}, {2, 3}}, {{1, 5}, {0, 2}, {3, 4}}, {{2, 5}, {1, 3}, {0, 4}},
                                           {{3, 5}, {2, 4}, {0, 1}}, {{4, 5}, {0, 3}, {1, 2}}};

bool solve6_min_norm(double A[36], double b[6], double x[6]) {
    double V[36];
    for (int k = 0; k < 36; ++k) V[k] = 0.0;
    for (int k = 0; k < 6; ++k) V[k * 6 + k] = 1.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < 6; ++p) {
            diag += std::fabs(A[p * 6 + p]);
            for (int q = p + 1; q < 6; ++q) off += std::fabs(A[p * 6 + q]);
        }
        if (!(off > 1e-300) || off <= 1e-18 * diag) break;
        for (int round = 0; round < 5; ++round) {
            double c[3], sn[3];
            for (int u = 0; u < 3; ++u) {  // the three rotations, from the matrix as the round finds it
                const int p = JACOBI_ROUNDS[round][u][0], q = JACOBI_ROUNDS[round][u][1];
                const double apq = A[p * 6 + q];
                if (apq == 0.0) {
                    c[u] = 1.0;
                    sn[u] = 0.0;
                    continue;
                }
                const double theta = (A[q * 6 + q] - A[p * 6 + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                c[u] = 1.0 / std::sqrt(t * t + 1.0);
                sn[u] = t * c[u];
            }
            for (int u = 0; u < 3; ++u) {  // columns p, q of A and V
                const int p = JACOBI_ROUNDS[round][u][0], q = JACOBI_ROUNDS[round][u][1];
                for (int k = 0; k < 6; ++k) {
                    const double akp = A[k * 6 + p], akq = A[k * 6 + q];
                    A[k * 6 + p] = c[u] * akp - sn[u] * akq;
                    A[k * 6 + q] = sn[u] * akp + c[u] * akq;
                    const double vkp = V[k * 6 + p], vkq = V[k * 6 + q];
                    V[k * 6 + p] = c[u] * vkp - sn[u] * vkq;
                    V[k * 6 + q] = sn[u] * vkp + c[u] * vkq;
                }
            }
            for (int u = 0; u < 3; ++u) {  // rows p, q of A
                const int p = JACOBI_ROUNDS[round][u][0], q = JACOBI_ROUNDS[round][u][1];
                for (int k = 0; k < 6; ++k) {
                    const double apk = A[p * 6 + k], aqk = A[q * 6 + k];
                    A[p * 6 + k] = c[u] * apk - sn[u] * aqk;
                    A[q * 6 + k] = sn[u] * apk + c[u] * aqk;
                }
            }
        }
    }
    double lmax = 0.0;
    for (int k = 0; k < 6; ++k) lmax = std::fmax(lmax, A[k * 6 + k]);
    if (!(lmax > 0.0)) return false;
    const double cut = 1e-12 * lmax;
    for (int r = 0; r < 6; ++r) x[r] = 0.0;
    for (int k = 0; k < 6; ++k) {
        const double l = A[k * 6 + k];
        if (!(l > cut)) continue;
        double proj = 0.0;
        for (int r = 0; r < 6; ++r) proj += V[r * 6 + k] * b[r];
        proj /= l;
        for (int r = 0; r < 6; ++r) x[r] += V[r * 6 + k] * proj;
    }
    for (int c = 0; c < 6; ++c)
        if (!std::isfinite(x[c])) return false;
    return true;
}

/* Fast path: the elimination of round 1 (partial pivoting) while every pivot is at least 1e-9 of the largest diagonal
 * entry — then the system has full rank and the least-squares solution is the minimum-norm one; otherwise (fewer than
 * six independent correspondences) the Jacobi form above.  Both implementations take the same branch on the same numbers. */
bool solve6(double A[36], double b[6], double x[6]) {
    double E[36], g[6], dmax = 0.0;
    for (int k = 0; k < 36; ++k) E[k] = A[k];
    for (int k = 0; k < 6; ++k) {
        g[k] = b[k];
        dmax = std::fmax(dmax, std::fabs(A[k * 6 + k]));
    }
    const double floor_pivot = 1e-9 * dmax;
    int perm[6] = {0, 1, 2, 3, 4, 5};
    bool full_rank = dmax > 0.0;
    for (int c = 0; c < 6 && full_rank; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r)
            if (std::fabs(E[perm[r] * 6 + c]) > std::fabs(E[perm[piv] * 6 + c])) piv = r;
        const int t = perm[c];
        perm[c] = perm[piv];
        perm[piv] = t;
        const double d = E[perm[c] * 6 + c];
        if (!(std::fabs(d) > floor_pivot)) {
            full_rank = false;
            break;
        }
        for (int r = c + 1; r < 6; ++r) {
            const double f = E[perm[r] * 6 + c] / d;
            for (int k = c; k < 6; ++k) E[perm[r] * 6 + k] -= f * E[perm[c] * 6 + k];
            g[perm[r]] -= f * g[perm[c]];
        }
    }
    if (!full_rank) return solve6_min_norm(A, b, x);
    for (int c = 5; c >= 0; --c) {
        double s = g[perm[c]];
        for (int k = c + 1; k < 6; ++k) s -= E[perm[c] * 6 + k] * x[k];
        x[c] = s / E[perm[c] * 6 + c];
    }
    for (int c = 0; c < 6; ++c)
        if (!std::isfinite(x[c])) return false;
    return true;
}

Pose pose_from_euler(const double *rpy, const double *t) {
    const double cx = std::cos(rpy[0]), sx = std::sin(rpy[0]);
    const double cy = std::cos(rpy[1]), sy = std::sin(rpy[1]);
    const double cz = std::cos(rpy[2]), sz = std::sin(rpy[2]);
    const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
    const double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    const double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    double T[9], R[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[r * 3 + c] = Ry[r * 3] * Rz[c] + Ry[r * 3 + 1] * Rz[3 + c] + Ry[r * 3 + 2] * Rz[6 + c];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r * 3 + c] = Rx[r * 3] * T[c] + Rx[r * 3 + 1] * T[3 + c] + Rx[r * 3 + 2] * T[6 + c];
    Pose P = identity();
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) P.m[r * 4 + c] = R[r * 3 + c];
        P.m[r * 4 + 3] = t[r];
    }
    return P;
}

double icp_single(const std::vector<float> &src_in, const std::vector<float> &dst_in, int max_iterations, float tolerance,
                  float rejection_scale, int num_levels, Pose &pose, uint64_t *iterations_run) {
    const int n = (int)(src_in.size() / 6);
    const size_t nd = dst_in.size() / 6;
    pose = identity();
    if (n == 0 || nd == 0) return 0.0;
    std::vector<float> src = src_in, dst = dst_in;
    // centre on the mean of the two centroids, scale by n / mean distance to the origin
    double ms[3] = {0, 0, 0}, md[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) ms[c] += src[6 * (size_t)i + c];
    for (size_t i = 0; i < nd; ++i)
        for (int c = 0; c < 3; ++c) md[c] += dst[6 * i + c];
    double mean[3];
    for (int c = 0; c < 3; ++c) mean[c] = 0.5 * (ms[c] / n + md[c] / (double)nd);
    auto centre = [&](std::vector<float> &pc) {
        double dist = 0.0;
        for (size_t i = 0; i < pc.size() / 6; ++i) {
            for (int c = 0; c < 3; ++c) pc[6 * i + c] = (float)((double)pc[6 * i + c] - mean[c]);
            dist += std::sqrt((double)pc[6 * i] * pc[6 * i] + (double)pc[6 * i + 1] * pc[6 * i + 1] +
                              (double)pc[6 * i + 2] * pc[6 * i + 2]);
        }
        return dist;
    };
    const double dist_src = centre(src), dist_dst = centre(dst);
    const double scale = (double)n / ((dist_src + dist_dst) * 0.5);
    for (size_t i = 0; i < src.size() / 6; ++i)
        for (int c = 0; c < 3; ++c) src[6 * i + c] = (float)((double)src[6 * i + c] * scale);
    for (size_t i = 0; i < dst.size() / 6; ++i)
        for (int c = 0; c < 3; ++c) dst[6 * i + c] = (float)((double)dst[6 * i + c] * scale);

    double residual = 0.0;
    for (int level = num_levels - 1; level >= 0; --level) {
        const double div = std::pow(2.0, (double)level);
        const int num_samples = (int)std::nearbyint((double)n / div);
        const double tol_p = (double)tolerance * (double)(level + 1) * (double)(level + 1);
        const int max_it = (int)std::nearbyint((double)max_iterations / (double)(level + 1));
        if (num_samples < 1) continue;
        const int step = std::max(1, (int)std::nearbyint((double)n / (double)num_samples));
        const std::vector<float> src_t = sample_uniform(transform_cloud(src, pose), step);
        const std::vector<float> dst_s = sample_uniform(dst, step);
        const size_t m = src_t.size() / 6, md_s = dst_s.size() / 6;
        if (m == 0 || md_s == 0) continue;
        Grid grid;
        grid.build(dst_s);
        double fval_old = 9999999999.0, fval_perc = 0.0, fval_min = 9999999999.0;
        std::vector<float> moved = src_t;
        Pose pose_x = identity();
        std::vector<int> nn(m);
        std::vector<float> d2(m);
        int it = 0;
        while (!(fval_perc < (1.0 + tol_p) && fval_perc > (1.0 - tol_p)) && it < max_it) {
            for (size_t i = 0; i < m; ++i) grid.nearest(&moved[6 * i], nn[i], d2[i]);
            std::vector<char> keep(m, 1);
            if (rejection_scale > 0.0f) {
                const float thr = rejection_threshold(d2, rejection_scale);
                for (size_t i = 0; i < m; ++i) keep[i] = d2[i] < thr;
            }
            // several model samples on one scene sample: the closest one survives (first on ties)
            std::vector<int> winner(md_s, -1);
            for (size_t i = 0; i < m; ++i) {
                if (!keep[i]) continue;
                int &w = winner[(size_t)nn[i]];
                if (w < 0 || d2[i] < d2[(size_t)w]) w = (int)i;
            }
            double A[36], bb[6];
            std::memset(A, 0, sizeof(A));
            std::memset(bb, 0, sizeof(bb));
            double err2 = 0.0;
            size_t matches = 0;
            for (size_t j = 0; j < md_s; ++j) {
                if (winner[j] < 0) continue;
                const float *s = &src_t[6 * (size_t)winner[j]];
                const float *d = &dst_s[6 * j];
                const double sp[3] = {s[0], s[1], s[2]}, dp[3] = {d[0], d[1], d[2]}, nr[3] = {d[3], d[4], d[5]};
                const double row[6] = {sp[1] * nr[2] - sp[2] * nr[1], sp[2] * nr[0] - sp[0] * nr[2],
                                       sp[0] * nr[1] - sp[1] * nr[0], nr[0], nr[1], nr[2]};
                const double rhs = (dp[0] - sp[0]) * nr[0] + (dp[1] - sp[1]) * nr[1] + (dp[2] - sp[2]) * nr[2];
                for (int r = 0; r < 6; ++r) {
                    for (int c = 0; c < 6; ++c) A[r * 6 + c] += row[r] * row[c];
                    bb[r] += row[r] * rhs;
                }
                for (int c = 0; c < 6; ++c) {
                    const double e = (double)s[c] - (double)d[c];
                    err2 += e * e;
                }
                ++matches;
            }
            if (matches == 0) break;
            double x[6];
            if (!solve6(A, bb, x)) break;
            pose_x = pose_from_euler(x, x + 3);
            moved = transform_cloud(src_t, pose_x);
            const double fval = std::sqrt(err2) / (double)m;
            fval_perc = fval / fval_old;
            fval_old = fval;
            if (fval < fval_min) fval_min = fval;
            ++it;
            if (iterations_run) ++*iterations_run;
        }
        pose = mul(pose_x, pose);
        residual = fval_min;
    }
    // undo the normalisation: t <- t / scale + mean - R * mean
    double t[3];
    for (int r = 0; r < 3; ++r) {
        const double rm = pose.m[r * 4] * mean[0] + pose.m[r * 4 + 1] * mean[1] + pose.m[r * 4 + 2] * mean[2];
        t[r] = pose.m[r * 4 + 3] / scale + mean[r] - rm;
    }
    for (int r = 0; r < 3; ++r) pose.m[r * 4 + 3] = t[r];
    return residual;
}

}  // namespace

extern "C" int oracle_icp_refine(const float *model, size_t n_model, const float *scene, size_t n_scene, int max_iterations,
                                 float tolerance, float rejection_scale, int num_levels, double *poses16, size_t n_poses,
                                 double *residuals, uint64_t *iterations_run) {
    if (!model || !scene || !poses16) return -1;
    const std::vector<float> m(model, model + 6 * n_model), s(scene, scene + 6 * n_scene);
    if (iterations_run) *iterations_run = 0;
    for (size_t p = 0; p < n_poses; ++p) {
        Pose start;
        std::memcpy(start.m, poses16 + 16 * p, sizeof(start.m));
        const std::vector<float> moved = transform_cloud(m, start);
        Pose delta;
        const double res = icp_single(moved, s, max_iterations, tolerance, rejection_scale, num_levels, delta, iterations_run);
        const Pose out = mul(delta, start);  // Pose3D::appendPose
        std::memcpy(poses16 + 16 * p, out.m, sizeof(out.m));
        if (residuals) residuals[p] = res;
    }
    return 0;
}

/* the minimum-norm solve alone (tests): N 6x6 row-major (destroyed), g[6] -> x[6]; 0 on success */
extern "C" int oracle_icp_solve6(double *N, double *g, double *x) { return solve6(N, g, x) ? 0 : -1; }
