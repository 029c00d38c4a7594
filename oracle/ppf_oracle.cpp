/*
 * ppf_oracle.cpp — CPU restatement of PCL's Point-Pair-Feature path.
 *
 * TEST INFRASTRUCTURE ONLY (see ppf_oracle.h).  PARITY UNPINNED: PCL itself is not in the
 * container and the reference repository (EmilyJrxx/YOLO_PPF_Pose_Estimation) holds no tests
 * or golden vectors for this path; this file restates the published PCL algorithm as recorded
 * in SURVEY.md Appendix A.  Reference call sites that enter this path:
 *   include/CloudProcessing.h:222-261 (train), :106-121 (load), :428-533 (match),
 *   :163-190 (N x 6 float layout), src/YOLO_cropping_ppf_test.cpp:113-127.
 *
 * Arithmetic conventions (fixed here; the device code has to match THESE):
 *   - everything IEEE binary32 unless the PCL source computes in double (noted inline);
 *   - no FMA contraction (built with -ffp-contract=off), dot products summed left to right
 *     (Eigen's SSE path sums (x0y0+x2y2)+(x1y1+x3y3): an ulp-level difference that falls
 *     under the 1e-5 bin-edge rule);
 *   - transcendental functions are the host libm's float versions (std::acos(float) etc.,
 *     exactly what PCL calls);
 *   - Affine inverse of a rigid frame is taken as the transpose (<=1e-7);
 *   - Affine3f::rotation() is taken as linear();
 *   - determinism rules of SURVEY.md A.8 (stable sorts, lowest flat index wins ties).
 */
#include "ppf_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 {
    float x, y, z;
};

inline V3 ld3(const float *p) { return V3{p[0], p[1], p[2]}; }
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross3(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float norm3(V3 a) { return std::sqrt(dot3(a, a)); }

/* ---- A.1  pcl::computePairFeatures (features/src/pfh.cpp) ---------------------------- */
bool pair_features_pfh(V3 p1, V3 n1, V3 p2, V3 n2, float *f) {
    V3 d{p2.x - p1.x, p2.y - p1.y, p2.z - p1.z};
    float f4 = norm3(d);
    if (f4 == 0.0f) {
        f[0] = f[1] = f[2] = f[3] = 0.0f;
        return false;
    }
    float angle1 = dot3(n1, d) / f4;
    float angle2 = dot3(n2, d) / f4;
    V3 u, nt;
    float f3;
    if (std::acos(std::fabs(angle1)) > std::acos(std::fabs(angle2))) {
        u = n2;
        nt = n1;
        d = V3{d.x * -1.0f, d.y * -1.0f, d.z * -1.0f};
        f3 = -angle2;
    } else {
        u = n1;
        nt = n2;
        f3 = angle1;
    }
    V3 v = cross3(d, u);
    float vn = norm3(v);
    if (vn == 0.0f) {
        f[0] = f[1] = f[2] = f[3] = 0.0f;
        return false;
    }
    v = V3{v.x / vn, v.y / vn, v.z / vn};
    V3 w = cross3(u, v);
    f[1] = dot3(v, nt);
    f[0] = std::atan2(dot3(w, nt), dot3(u, nt));
    f[2] = f3;
    f[3] = f4;
    return true;
}

/* ---- A.1' pcl::computePPFPairFeature (features/src/ppf.cpp), cosines ------------------ */
bool pair_features_drost_cos(V3 p1, V3 n1, V3 p2, V3 n2, float *f) {
    V3 d{p2.x - p1.x, p2.y - p1.y, p2.z - p1.z};
    float f4 = norm3(d);
    if (f4 == 0.0f) { /* PCL would produce NaN here; our builds treat it as a failed pair */
        f[0] = f[1] = f[2] = f[3] = 0.0f;
        return false;
    }
    d = V3{d.x / f4, d.y / f4, d.z / f4};
    f[0] = dot3(n1, d);
    f[1] = dot3(n2, d);
    f[2] = dot3(n1, n2);
    f[3] = f4;
    return true;
}

inline float clamp_unit(float c) { return c > 1.0f ? 1.0f : (c < -1.0f ? -1.0f : c); }

/* Drost tuple as angles (north_star wording): acos of the cosines above */
bool pair_features_drost_angle(V3 p1, V3 n1, V3 p2, V3 n2, float *f) {
    if (!pair_features_drost_cos(p1, n1, p2, n2, f)) return false;
    f[0] = std::acos(clamp_unit(f[0]));
    f[1] = std::acos(clamp_unit(f[1]));
    f[2] = std::acos(clamp_unit(f[2]));
    return true;
}

bool pair_features(int mode, V3 p1, V3 n1, V3 p2, V3 n2, float *f) {
    switch (mode) {
        case ORACLE_FEATURE_DROST_COS: return pair_features_drost_cos(p1, n1, p2, n2, f);
        case ORACLE_FEATURE_DROST_ANGLE: return pair_features_drost_angle(p1, n1, p2, n2, f);
        default: return pair_features_pfh(p1, n1, p2, n2, f);
    }
}

/* ---- A.6  Eigen::AngleAxisf::toRotationMatrix ---------------------------------------- */
void angle_axis_matrix(float angle, V3 a, float *R) {
    float s = std::sin(angle), c = std::cos(angle);
    V3 sa{s * a.x, s * a.y, s * a.z};
    float omc = 1.0f - c;
    V3 ca{omc * a.x, omc * a.y, omc * a.z};
    float tmp;
    tmp = ca.x * a.y;
    R[1] = tmp - sa.z;
    R[3] = tmp + sa.z;
    tmp = ca.x * a.z;
    R[2] = tmp + sa.y;
    R[6] = tmp - sa.y;
    tmp = ca.y * a.z;
    R[5] = tmp - sa.x;
    R[7] = tmp + sa.x;
    R[0] = ca.x * a.x + c;
    R[4] = ca.y * a.y + c;
    R[8] = ca.z * a.z + c;
}

inline V3 matvec(const float *R, V3 v) {
    return V3{(R[0] * v.x + R[1] * v.y) + R[2] * v.z, (R[3] * v.x + R[4] * v.y) + R[5] * v.z,
              (R[6] * v.x + R[7] * v.y) + R[8] * v.z};
}

/* ---- A.2  frame that moves (p_r, n_r) to the origin with n_r on +x -------------------- */
struct Frame {
    float R[9];
    float t[3];
};

Frame ref_frame(V3 p, V3 n) {
    Frame F;
    float angle = std::acos(n.x); /* n . UnitX */
    bool parallel = (n.y == 0.0f && n.z == 0.0f);
    V3 axis;
    if (parallel) {
        axis = V3{0.0f, 1.0f, 0.0f};
    } else {
        V3 c{0.0f, n.z, -n.y}; /* n x UnitX */
        float z = (c.x * c.x + c.y * c.y) + c.z * c.z;
        if (z > 0.0f) {
            float s = std::sqrt(z);
            c = V3{c.x / s, c.y / s, c.z / s};
        }
        axis = c;
    }
    angle_axis_matrix(angle, axis, F.R);
    V3 mp{-1.0f * p.x, -1.0f * p.y, -1.0f * p.z};
    V3 t = matvec(F.R, mp);
    F.t[0] = t.x;
    F.t[1] = t.y;
    F.t[2] = t.z;
    return F;
}

inline V3 apply(const Frame &F, V3 m) {
    V3 r = matvec(F.R, m);
    return V3{r.x + F.t[0], r.y + F.t[1], r.z + F.t[2]};
}

float planar_alpha(const Frame &F, V3 m) {
    V3 mt = apply(F, m);
    float a = std::atan2(-mt.z, mt.y);
    if (std::sin(a) * mt.z < 0.0f) a *= -1.0f;
    return -a;
}

/* ---- A.3 hash key ---------------------------------------------------------------------- */
struct Key {
    int32_t d[4];
    bool operator==(const Key &o) const {
        return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2] && d[3] == o.d[3];
    }
};
struct KeyHash {
    /* PCL >= 1.9: h1 ^ (h2<<1) ^ (h3<<2) ^ (h4<<3).  That function takes only a few hundred distinct
     * values, so every insert and lookup walks a chain of thousands of unrelated keys and the table
     * build is quadratic (minutes at 2 000 model points, hours at 10 000).  Above
     * ORACLE_PCL_HASH_MAX_POINTS model points the oracle therefore mixes the four integers properly.
     * Only speed and iteration order change: bucket contents (what every consumer reads, sorted or
     * order-independent) are those of PCL's container. */
    bool mixed = false;
    size_t operator()(const Key &k) const {
        std::hash<int> h;
        if (!mixed) return h(k.d[0]) ^ (h(k.d[1]) << 1) ^ (h(k.d[2]) << 2) ^ (h(k.d[3]) << 3);
        uint64_t x = (uint64_t)(uint32_t)k.d[0] * 0x9E3779B97F4A7C15ull;
        x = (x ^ (uint32_t)k.d[1]) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (uint32_t)k.d[2]) * 0x94D049BB133111EBull;
        x = (x ^ (uint32_t)k.d[3]) * 0x9E3779B97F4A7C15ull;
        return (size_t)(x ^ (x >> 31));
    }
};
constexpr size_t ORACLE_PCL_HASH_MAX_POINTS = 1024;

}  // namespace

struct oracle_hashmap {
    float angle_step;
    float dist_step;
    float max_dist;
    int nalpha_rule = ORACLE_NALPHA_CEIL;
    size_t n;
    /* PCL's one unordered_multimap; built by several threads it is split into shards by key (every key lives in
     * exactly one shard, so equal_range on its shard returns what the single container would) */
    typedef std::unordered_multimap<Key, std::pair<size_t, size_t>, KeyHash> Map;
    std::vector<Map> maps;
    std::vector<std::vector<float>> alpha_m;
    size_t n_keys;
    size_t shard_of(const Key &k) const {
        if (maps.size() <= 1) return 0;
        KeyHash h;
        h.mixed = true;
        return (h(k) >> 7) % maps.size();
    }
    std::pair<Map::const_iterator, Map::const_iterator> equal_range(const Key &k) const {
        return maps[shard_of(k)].equal_range(k);
    }
    size_t size() const {
        size_t t = 0;
        for (const Map &m : maps) t += m.size();
        return t;
    }
};

namespace {

inline Key quantise(const oracle_hashmap *hm, const float *f) {
    Key k;
    k.d[0] = static_cast<int>(std::floor(f[0] / hm->angle_step));
    k.d[1] = static_cast<int>(std::floor(f[1] / hm->angle_step));
    k.d[2] = static_cast<int>(std::floor(f[2] / hm->angle_step));
    k.d[3] = static_cast<int>(std::floor(f[3] / hm->dist_step));
    return k;
}

/* Columns of the accumulator (ppf_registration.hpp, `aux_size`).  PCL 1.8 .. 1.11 take the floor of
 * 2*pi / step, current PCL the ceiling; for PCL's float 12 degrees the quotient is 30 - 1.3e-6, i.e. 29 or 30
 * columns while the binning formula below produces bins 0 .. 29.  ppf_oracle.h ORACLE_NALPHA_*. */
inline uint32_t num_alpha_bins(float angle_step, int rule) {
    const double t = 2 * M_PI / angle_step; /* double */
    return static_cast<uint32_t>(rule == ORACLE_NALPHA_CEIL ? std::ceil(t) : std::floor(t));
}

constexpr uint32_t BIN_DROPPED = UINT32_MAX - 1;

/* what becomes of a bin past the last column */
inline uint32_t place_bin(int rule, uint32_t n_alpha, uint32_t bin) {
    if (bin < n_alpha) return bin;
    /* FLOOR_DROP: PCL <= 1.11 executes accumulator_array[i][29]++ on a 29-element std::vector<unsigned int>: the word
     * behind the row's 116 bytes, inside the 120 usable bytes of its glibc chunk — never read again, the vote is lost.
     * CEIL reaches this line only when 2*pi/step is an integer to the last bit (then PCL writes out of bounds too);
     * the oracle clamps there, as it does under FLOOR_CLAMP, the rule of this repository's round 1. */
    return rule == ORACLE_NALPHA_FLOOR_DROP ? BIN_DROPPED : n_alpha - 1;
}

inline uint32_t alpha_bin(int mode, float angle_step, uint32_t n_alpha, int rule, float alpha_m,
                          float alpha_s) {
    float alpha = alpha_m - alpha_s;
    if (std::isnan(alpha)) return UINT32_MAX;
    uint32_t bin;
    if (mode == ORACLE_ALPHA_MODE_B) {
        double b = std::floor(alpha) + std::floor(M_PI / angle_step);
        bin = b < 0 ? 0u : static_cast<uint32_t>(b);
    } else {
        if (alpha < -M_PI) {
            alpha += (2 * M_PI); /* double add, stored back to float */
        } else if (alpha > M_PI) {
            alpha -= (2 * M_PI);
        }
        double b = std::floor((alpha + M_PI) / angle_step);
        bin = b < 0 ? 0u : static_cast<uint32_t>(b);
    }
    return place_bin(rule, n_alpha, bin);
}

/* pose = T_sg^-1 * Rx(theta) * T_mg  as 3x4 row-major */
void compose_pose(const Frame &sg, float theta, const Frame &mg, float *P) {
    float Rx[9];
    angle_axis_matrix(theta, V3{1.0f, 0.0f, 0.0f}, Rx);
    /* inverse of the rigid scene frame: R^T, -R^T t */
    float Ri[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Ri[r * 3 + c] = sg.R[c * 3 + r];
    V3 ti = matvec(Ri, V3{sg.t[0], sg.t[1], sg.t[2]});
    ti = V3{-ti.x, -ti.y, -ti.z};
    /* A = Ri * Rx (translation ti) */
    float A[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            A[r * 3 + c] = (Ri[r * 3 + 0] * Rx[0 * 3 + c] + Ri[r * 3 + 1] * Rx[1 * 3 + c]) +
                           Ri[r * 3 + 2] * Rx[2 * 3 + c];
    /* P = A * T_mg */
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            P[r * 4 + c] = (A[r * 3 + 0] * mg.R[0 * 3 + c] + A[r * 3 + 1] * mg.R[1 * 3 + c]) +
                           A[r * 3 + 2] * mg.R[2 * 3 + c];
    }
    V3 t = matvec(A, V3{mg.t[0], mg.t[1], mg.t[2]});
    P[3] = t.x + ti.x;
    P[7] = t.y + ti.y;
    P[11] = t.z + ti.z;
}

float peak_theta(int alpha_mode, float angle_step, uint32_t bin) {
    if (alpha_mode == ORACLE_ALPHA_MODE_B) {
        float k = static_cast<float>(static_cast<double>(bin) - std::floor(M_PI / angle_step));
        return k * angle_step;
    }
    /* static_cast<float>(max_j + 0.5) * step - M_PI  (float product, double subtract) */
    float k = static_cast<float>(bin + 0.5);
    return static_cast<float>(k * angle_step - M_PI);
}

/* ---- A.6 quaternion helpers (coeffs order x,y,z,w) ------------------------------------- */
void quat_from_matrix(const float *R /*row-major 3x3, stride 3*/, float *q) {
    float t = R[0] + R[4] + R[8];
    if (t > 0.0f) {
        t = std::sqrt(t + 1.0f);
        q[3] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0f);
        q[i] = 0.5f * t;
        t = 0.5f / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
}

void quat_to_matrix(const float *q, float *R) {
    float tx = 2.0f * q[0], ty = 2.0f * q[1], tz = 2.0f * q[2];
    float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1.0f - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = 1.0f - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = 1.0f - (txx + tyy);
}

/* |AngleAxisf(R).angle()| : matrix -> quaternion -> 2*atan2(|vec|, |w|) */
float rotation_angle(const float *R) {
    float q[4];
    quat_from_matrix(R, q);
    float n = std::sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    if (n < FLT_EPSILON) { /* Eigen: stableNorm() */
        float m = std::max(std::fabs(q[0]), std::max(std::fabs(q[1]), std::fabs(q[2])));
        if (m > 0.0f) {
            float a = q[0] / m, b = q[1] / m, c = q[2] / m;
            n = m * std::sqrt((a * a + b * b) + c * c);
        } else {
            n = 0.0f;
        }
    }
    if (n != 0.0f) return std::fabs(2.0f * std::atan2(n, std::fabs(q[3])));
    return 0.0f;
}

bool poses_within(const float *a, const float *b, float pos_thr, float rot_thr) {
    V3 dt{a[3] - b[3], a[7] - b[7], a[11] - b[11]};
    float position_diff = norm3(dt);
    /* PCL evaluates both differences and ANDs them; a pose too far away fails whatever its rotation, so the
     * rotation is only worked out for the others (same result, and the greedy loop over 10^5 poses stays in seconds) */
    if (!(position_diff < pos_thr)) return false;
    /* R_a^-1 * R_b, inverse as transpose */
    float M[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            M[r * 3 + c] = (a[0 * 4 + r] * b[0 * 4 + c] + a[1 * 4 + r] * b[1 * 4 + c]) +
                           a[2 * 4 + r] * b[2 * 4 + c];
    float rotation_diff_angle = rotation_angle(M);
    return position_diff < pos_thr && rotation_diff_angle < rot_thr;
}

/* uniform grid over the scene for the d/2 radius search (stands in for FLANN; favours the CPU) */
struct Grid {
    float cell;
    float origin[3];
    std::unordered_map<uint64_t, std::vector<uint32_t>> cells;
    static uint64_t pack(int64_t x, int64_t y, int64_t z) {
        return (static_cast<uint64_t>(x & 0x1FFFFF) << 42) | (static_cast<uint64_t>(y & 0x1FFFFF) << 21) |
               static_cast<uint64_t>(z & 0x1FFFFF);
    }
    void build(const float *scene, size_t n, float cell_size) {
        cell = cell_size;
        origin[0] = origin[1] = origin[2] = 0.0f;
        for (size_t s = 0; s < n; ++s) {
            const float *p = scene + 6 * s;
            if (std::isnan(p[0]) || std::isnan(p[1]) || std::isnan(p[2])) continue;
            cells[pack(cx(p[0]), cx(p[1]), cx(p[2]))].push_back(static_cast<uint32_t>(s));
        }
    }
    int64_t cx(float v) const { return static_cast<int64_t>(std::floor(static_cast<double>(v) / cell)); }
    template <class Fn>
    void for_each_candidate(const float *p, Fn fn) const {
        int64_t x = cx(p[0]), y = cx(p[1]), z = cx(p[2]);
        for (int64_t dx = -1; dx <= 1; ++dx)
            for (int64_t dy = -1; dy <= 1; ++dy)
                for (int64_t dz = -1; dz <= 1; ++dz) {
                    auto it = cells.find(pack(x + dx, y + dy, z + dz));
                    if (it == cells.end()) continue;
                    for (uint32_t s : it->second) fn(s);
                }
    }
};

/* one scene point paired with the reference point: the body of PCL's inner loop.  counters = examined, in radius,
 * non-empty lookups, votes */
inline void vote_one_pair(const oracle_hashmap *hm, int feature_mode, int alpha_mode, const float *scene, size_t s_r,
                          size_t s_i, const V3 &p_r, const V3 &n_r, const Frame &sg, float radius, uint32_t n_alpha,
                          uint32_t *acc, std::vector<std::pair<size_t, size_t>> &bucket, uint64_t *counters) {
    if (s_i == s_r) return;
    ++counters[0];
    const float *pi = scene + 6 * s_i;
    V3 p_i = ld3(pi), n_i = ld3(pi + 3);
    /* radius predicate on f4 itself (SURVEY.md A.8 rule 7) */
    V3 d{p_i.x - p_r.x, p_i.y - p_r.y, p_i.z - p_r.z};
    float dist = norm3(d);
    if (!(dist < radius)) return;
    float f[4];
    if (!pair_features(feature_mode, p_r, n_r, p_i, n_i, f)) return;
    ++counters[1];
    Key k = quantise(hm, f);
    auto range = hm->equal_range(k);
    bucket.clear();
    for (auto it = range.first; it != range.second; ++it) bucket.push_back(it->second);
    if (bucket.empty()) return;
    ++counters[2];
    float alpha_s = planar_alpha(sg, p_i);
    for (const auto &ij : bucket) {
        uint32_t bin = alpha_bin(alpha_mode, hm->angle_step, n_alpha, hm->nalpha_rule, hm->alpha_m[ij.first][ij.second], alpha_s);
        if (bin == UINT32_MAX) continue;
        if (bin != BIN_DROPPED) acc[ij.first * n_alpha + bin]++;
        ++counters[3]; /* a dropped vote is still an increment PCL executes */
    }
}

/* first maximum, i-major / bin-minor, strict '>' ; reset ; pose of the peak */
inline void peak_and_reset(const oracle_hashmap *hm, int alpha_mode, const float *model, size_t n_m, const Frame &sg,
                           size_t s_r, uint32_t n_alpha, uint32_t *acc, oracle_hypothesis *hyp) {
    uint32_t max_v = 0;
    size_t max_i = 0, max_j = 0;
    for (size_t i = 0; i < n_m; ++i)
        for (size_t j = 0; j < n_alpha; ++j) {
            uint32_t v = acc[i * n_alpha + j];
            if (v > max_v) {
                max_v = v;
                max_i = i;
                max_j = j;
            }
            acc[i * n_alpha + j] = 0;
        }
    const float *pm = model + 6 * max_i;
    Frame mg = ref_frame(ld3(pm), ld3(pm + 3));
    compose_pose(sg, peak_theta(alpha_mode, hm->angle_step, static_cast<uint32_t>(max_j)), mg, hyp->pose);
    hyp->votes = max_v;
    hyp->model_index = static_cast<uint32_t>(max_i);
    hyp->alpha_bin = static_cast<uint32_t>(max_j);
    hyp->scene_index = static_cast<uint32_t>(s_r);
}

/* one reference point of the A.4 voting loop; acc must be zero on entry and is zero on exit */
void vote_one_reference(const oracle_hashmap *hm, int feature_mode, int alpha_mode,
                        const float *model, size_t n_m, const float *scene, size_t n_s,
                        const Grid *grid, size_t s_r, uint32_t n_alpha, uint32_t *acc,
                        oracle_hypothesis *hyp, uint64_t *stats) {
    const float *pr = scene + 6 * s_r;
    V3 p_r = ld3(pr), n_r = ld3(pr + 3);
    Frame sg = ref_frame(p_r, n_r);
    const float radius = hm->max_dist * 0.5f;
    std::vector<std::pair<size_t, size_t>> bucket;
    uint64_t counters[4] = {0, 0, 0, 0};
    auto visit = [&](size_t s_i) {
        vote_one_pair(hm, feature_mode, alpha_mode, scene, s_r, s_i, p_r, n_r, sg, radius, n_alpha, acc, bucket, counters);
    };
    if (grid) {
        grid->for_each_candidate(pr, visit);
    } else {
        for (size_t s_i = 0; s_i < n_s; ++s_i) visit(s_i);
    }
    peak_and_reset(hm, alpha_mode, model, n_m, sg, s_r, n_alpha, acc, hyp);
    if (stats)
        for (int k = 0; k < 4; ++k) stats[k] += counters[k];
}

#ifdef _OPENMP
/* the same reference point with its scene pairs spread over the threads (thread-private accumulators summed before
 * the peak scan): what keeps every core busy when a pass holds fewer reference points than threads, or a few very
 * expensive ones.  Counters are integers and accumulator adds commute: the hypothesis is the serial one. */
void vote_one_reference_parallel(const oracle_hashmap *hm, int feature_mode, int alpha_mode, const float *model, size_t n_m,
                                 const float *scene, size_t n_s, const Grid *grid, size_t s_r, uint32_t n_alpha,
                                 std::vector<std::vector<uint32_t>> &accs, oracle_hypothesis *hyp, uint64_t *stats) {
    const float *pr = scene + 6 * s_r;
    V3 p_r = ld3(pr), n_r = ld3(pr + 3);
    Frame sg = ref_frame(p_r, n_r);
    const float radius = hm->max_dist * 0.5f;
    std::vector<size_t> cand;
    if (grid) {
        grid->for_each_candidate(pr, [&](size_t s_i) { cand.push_back(s_i); });
    } else {
        cand.resize(n_s);
        for (size_t s_i = 0; s_i < n_s; ++s_i) cand[s_i] = s_i;
    }
    const int n_threads = static_cast<int>(accs.size());
    uint64_t tot[4] = {0, 0, 0, 0};
#pragma omp parallel num_threads(n_threads)
    {
        uint32_t *acc = accs[omp_get_thread_num()].data();
        std::vector<std::pair<size_t, size_t>> bucket;
        uint64_t counters[4] = {0, 0, 0, 0};
#pragma omp for schedule(dynamic, 16)
        for (long long c = 0; c < static_cast<long long>(cand.size()); ++c)
            vote_one_pair(hm, feature_mode, alpha_mode, scene, s_r, cand[c], p_r, n_r, sg, radius, n_alpha, acc, bucket, counters);
#pragma omp critical
        for (int k = 0; k < 4; ++k) tot[k] += counters[k];
#pragma omp barrier
        /* sum the private accumulators into the first one, rows split over the threads; clear the others */
#pragma omp for schedule(static)
        for (long long i = 0; i < static_cast<long long>(n_m); ++i)
            for (int t = 1; t < n_threads; ++t) {
                uint32_t *src = accs[t].data() + static_cast<size_t>(i) * n_alpha, *dst = accs[0].data() + static_cast<size_t>(i) * n_alpha;
                for (uint32_t j = 0; j < n_alpha; ++j) {
                    dst[j] += src[j];
                    src[j] = 0;
                }
            }
    }
    peak_and_reset(hm, alpha_mode, model, n_m, sg, s_r, n_alpha, accs[0].data(), hyp);
    if (stats)
        for (int k = 0; k < 4; ++k) stats[k] += tot[k];
}
#endif

}  // namespace

extern "C" {

int oracle_pair_feature(int feature_mode, const float *p1, const float *n1, const float *p2,
                        const float *n2, float *f) {
    return pair_features(feature_mode, ld3(p1), ld3(n1), ld3(p2), ld3(n2), f) ? 1 : 0;
}

float oracle_alpha(const float *p_r, const float *n_r, const float *p) {
    Frame F = ref_frame(ld3(p_r), ld3(n_r));
    return planar_alpha(F, ld3(p));
}

void oracle_ref_frame(const float *p_r, const float *n_r, float *R, float *t) {
    Frame F = ref_frame(ld3(p_r), ld3(n_r));
    std::memcpy(R, F.R, sizeof(F.R));
    std::memcpy(t, F.t, sizeof(F.t));
}

size_t oracle_ppf_estimation(int feature_mode, const float *cloud, size_t n, float *out) {
    return oracle_ppf_estimation_mt(feature_mode, cloud, n, out, 1);
}

/* rows of the N x N feature cloud are independent: n_threads > 1 computes them in parallel (same values) */
size_t oracle_ppf_estimation_mt(int feature_mode, const float *cloud, size_t n, float *out, int n_threads) {
    const float nan = std::numeric_limits<float>::quiet_NaN();
    size_t valid = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(n_threads > 1 ? n_threads : 1) schedule(static) reduction(+ : valid)
#else
    (void)n_threads;
#endif
    for (long long ii = 0; ii < static_cast<long long>(n); ++ii) {
        const size_t i = static_cast<size_t>(ii);
        V3 p_i = ld3(cloud + 6 * i), n_i = ld3(cloud + 6 * i + 3);
        for (size_t j = 0; j < n; ++j) {
            float *o = out + (i * n + j) * 5;
            float f[4];
            if (i != j &&
                pair_features(feature_mode, p_i, n_i, ld3(cloud + 6 * j), ld3(cloud + 6 * j + 3), f)) {
                /* PCL recomputes the frame of i for every j; hoisting it changes nothing */
                Frame mg = ref_frame(p_i, n_i);
                o[0] = f[0];
                o[1] = f[1];
                o[2] = f[2];
                o[3] = f[3];
                o[4] = planar_alpha(mg, ld3(cloud + 6 * j));
                ++valid;
            } else {
                o[0] = o[1] = o[2] = o[3] = o[4] = nan;
            }
        }
    }
    return valid;
}

oracle_hashmap *oracle_hashmap_create(float angle_step, float dist_step) {
    oracle_hashmap *hm = new oracle_hashmap();
    hm->angle_step = angle_step;
    hm->dist_step = dist_step;
    hm->max_dist = -1.0f;
    hm->n = 0;
    hm->n_keys = 0;
    return hm;
}

void oracle_hashmap_destroy(oracle_hashmap *hm) { delete hm; }

void oracle_hashmap_set_features(oracle_hashmap *hm, const float *feats, size_t count) {
    oracle_hashmap_set_features_mt(hm, feats, count, 1);
}

void oracle_hashmap_set_features_mt(oracle_hashmap *hm, const float *feats, size_t count, int n_threads) {
    unsigned int n = static_cast<unsigned int>(std::sqrt(static_cast<float>(count)));
    KeyHash kh;
    kh.mixed = n > ORACLE_PCL_HASH_MAX_POINTS && !std::getenv("ORACLE_FORCE_PCL_HASH");
    const size_t shards = n_threads > 1 ? static_cast<size_t>(n_threads) : 1;
    hm->maps.clear();
    for (size_t t = 0; t < shards; ++t) hm->maps.emplace_back(0, kh);
    hm->n = n;
    hm->max_dist = -1.0f;
    hm->alpha_m.assign(n, std::vector<float>());
    for (size_t i = 0; i < n; ++i) {
        std::vector<float> row(n);
        for (size_t j = 0; j < n; ++j) {
            const float *s = feats + (i * n + j) * 5;
            row[j] = s[4];
            if (std::isnan(s[0]) || std::isnan(s[1]) || std::isnan(s[2]) || std::isnan(s[3])) continue;
            if (hm->max_dist < s[3]) hm->max_dist = s[3];
        }
        hm->alpha_m[i] = std::move(row);
    }
    /* the inserts, in PCL's order (i ascending, j ascending) inside every shard */
    auto fill = [&](size_t shard) {
        oracle_hashmap::Map &map = hm->maps[shard];
        if (kh.mixed) map.reserve(count / shards + count / (8 * shards) + 16);
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < n; ++j) {
                const float *s = feats + (i * n + j) * 5;
                /* PCL inserts the NaN diagonal too (key = (int)NaN, UB); no finite query can reach
                 * those nodes, so skipping them is observationally identical. */
                if (std::isnan(s[0]) || std::isnan(s[1]) || std::isnan(s[2]) || std::isnan(s[3])) continue;
                Key k = quantise(hm, s);
                if (shards > 1 && hm->shard_of(k) != shard) continue;
                map.insert(std::make_pair(k, std::make_pair(i, j)));
            }
    };
#ifdef _OPENMP
    if (shards > 1) {
#pragma omp parallel for num_threads(static_cast<int>(shards)) schedule(static, 1)
        for (long long t = 0; t < static_cast<long long>(shards); ++t) fill(static_cast<size_t>(t));
    } else {
        fill(0);
    }
#else
    for (size_t t = 0; t < shards; ++t) fill(t);
#endif
    size_t keys = 0;
    for (const oracle_hashmap::Map &map : hm->maps)
        for (auto it = map.begin(); it != map.end();) {
            ++keys;
            it = map.equal_range(it->first).second;
        }
    hm->n_keys = keys;
}

float oracle_hashmap_model_diameter(const oracle_hashmap *hm) { return hm->max_dist; }
size_t oracle_hashmap_num_entries(const oracle_hashmap *hm) { return hm->size(); }
size_t oracle_hashmap_num_keys(const oracle_hashmap *hm) { return hm->n_keys; }

void oracle_hashmap_quantise(const oracle_hashmap *hm, const float *f, int32_t *d) {
    Key k = quantise(hm, f);
    std::memcpy(d, k.d, sizeof(k.d));
}

size_t oracle_hashmap_query_key(const oracle_hashmap *hm, const int32_t *d, uint64_t *pairs,
                                size_t cap) {
    Key k;
    std::memcpy(k.d, d, sizeof(k.d));
    auto range = hm->equal_range(k);
    std::vector<std::pair<size_t, size_t>> v;
    for (auto it = range.first; it != range.second; ++it) v.push_back(it->second);
    std::sort(v.begin(), v.end());
    for (size_t e = 0; e < v.size() && e < cap; ++e) {
        pairs[2 * e] = v[e].first;
        pairs[2 * e + 1] = v[e].second;
    }
    return v.size();
}

size_t oracle_hashmap_query(const oracle_hashmap *hm, float f1, float f2, float f3, float f4,
                            uint64_t *pairs, size_t cap) {
    float f[4] = {f1, f2, f3, f4};
    Key k = quantise(hm, f);
    return oracle_hashmap_query_key(hm, k.d, pairs, cap);
}

void oracle_hashmap_dump_keys(const oracle_hashmap *hm, int32_t *keys, uint32_t *lengths) {
    size_t idx = 0;
    for (const oracle_hashmap::Map &map : hm->maps)
        for (auto it = map.begin(); it != map.end();) {
            auto range = map.equal_range(it->first);
            std::memcpy(keys + 4 * idx, it->first.d, 4 * sizeof(int32_t));
            lengths[idx] = static_cast<uint32_t>(std::distance(range.first, range.second));
            ++idx;
            it = range.second;
        }
}

void oracle_hashmap_set_nalpha_rule(oracle_hashmap *hm, int rule) { hm->nalpha_rule = rule; }

uint32_t oracle_num_alpha_bins(float angle_step, int nalpha_rule) { return num_alpha_bins(angle_step, nalpha_rule); }

uint32_t oracle_alpha_bin(int alpha_mode, int nalpha_rule, float angle_step, float alpha_m, float alpha_s) {
    return alpha_bin(alpha_mode, angle_step, num_alpha_bins(angle_step, nalpha_rule), nalpha_rule, alpha_m, alpha_s);
}

size_t oracle_scene_pairs(const oracle_hashmap *hm, int feature_mode, const float *scene,
                          size_t n_s, size_t s_r, uint8_t *in_radius, int32_t *d, float *alpha_s) {
    const float *pr = scene + 6 * s_r;
    V3 p_r = ld3(pr), n_r = ld3(pr + 3);
    Frame sg = ref_frame(p_r, n_r);
    const float radius = hm->max_dist * 0.5f;
    size_t count = 0;
    for (size_t s = 0; s < n_s; ++s) {
        in_radius[s] = 0;
        d[4 * s] = d[4 * s + 1] = d[4 * s + 2] = d[4 * s + 3] = 0;
        alpha_s[s] = 0.0f;
        if (s == s_r) continue;
        V3 p_i = ld3(scene + 6 * s), n_i = ld3(scene + 6 * s + 3);
        V3 dd{p_i.x - p_r.x, p_i.y - p_r.y, p_i.z - p_r.z};
        if (!(norm3(dd) < radius)) continue;
        float f[4];
        if (!pair_features(feature_mode, p_r, n_r, p_i, n_i, f)) continue;
        Key k = quantise(hm, f);
        in_radius[s] = 1;
        std::memcpy(d + 4 * s, k.d, sizeof(k.d));
        alpha_s[s] = planar_alpha(sg, p_i);
        ++count;
    }
    return count;
}

uint64_t oracle_vote_accumulate_from_pairs(const oracle_hashmap *hm, int alpha_mode, size_t n_m,
                                           size_t n_pairs, const int32_t *d, const float *alpha_s,
                                           uint32_t *acc) {
    return oracle_vote_accumulate_from_pairs_mt(hm, alpha_mode, n_m, n_pairs, d, alpha_s, acc, 1);
}

/* n_threads > 1: the pairs are spread over threads with private accumulators that are summed at the end (integer
 * adds commute: the result is the serial one) */
uint64_t oracle_vote_accumulate_from_pairs_mt(const oracle_hashmap *hm, int alpha_mode, size_t n_m, size_t n_pairs,
                                              const int32_t *d, const float *alpha_s, uint32_t *acc, int n_threads) {
    const uint32_t n_alpha = num_alpha_bins(hm->angle_step, hm->nalpha_rule);
    const size_t len = n_m * n_alpha;
    std::memset(acc, 0, len * sizeof(uint32_t));
    uint64_t votes = 0;
    auto walk = [&](size_t p, uint32_t *a, uint64_t &v) {
        Key k;
        std::memcpy(k.d, d + 4 * p, sizeof(k.d));
        auto range = hm->equal_range(k);
        for (auto it = range.first; it != range.second; ++it) {
            uint32_t bin = alpha_bin(alpha_mode, hm->angle_step, n_alpha, hm->nalpha_rule,
                                     hm->alpha_m[it->second.first][it->second.second], alpha_s[p]);
            if (bin == UINT32_MAX) continue;
            if (bin != BIN_DROPPED) a[it->second.first * n_alpha + bin]++;
            ++v;
        }
    };
#ifdef _OPENMP
    if (n_threads > 1) {
#pragma omp parallel num_threads(n_threads)
        {
            std::vector<uint32_t> priv(len, 0u);
            uint64_t v = 0;
#pragma omp for schedule(dynamic, 8)
            for (long long p = 0; p < static_cast<long long>(n_pairs); ++p) walk(static_cast<size_t>(p), priv.data(), v);
#pragma omp critical
            {
                for (size_t e = 0; e < len; ++e) acc[e] += priv[e];
                votes += v;
            }
        }
        return votes;
    }
#else
    (void)n_threads;
#endif
    for (size_t p = 0; p < n_pairs; ++p) walk(p, acc, votes);
    return votes;
}

uint64_t oracle_vote_accumulate(const oracle_hashmap *hm, int feature_mode, int alpha_mode,
                                size_t n_m, const float *scene, size_t n_s, size_t s_r,
                                uint32_t *acc) {
    std::vector<uint8_t> in_radius(n_s);
    std::vector<int32_t> d(4 * n_s);
    std::vector<float> alpha_s(n_s);
    oracle_scene_pairs(hm, feature_mode, scene, n_s, s_r, in_radius.data(), d.data(), alpha_s.data());
    std::vector<int32_t> dc;
    std::vector<float> ac;
    for (size_t s = 0; s < n_s; ++s)
        if (in_radius[s]) {
            dc.insert(dc.end(), d.begin() + 4 * s, d.begin() + 4 * s + 4);
            ac.push_back(alpha_s[s]);
        }
    return oracle_vote_accumulate_from_pairs(hm, alpha_mode, n_m, ac.size(), dc.data(), ac.data(), acc);
}

int oracle_vote(const oracle_hashmap *hm, int feature_mode, int alpha_mode, const float *model,
                size_t n_m, const float *scene, size_t n_s, size_t ref_first, size_t ref_step,
                size_t ref_count, int n_threads, oracle_hypothesis *hyps, uint64_t *stats) {
    if (ref_step == 0) return -1;
    std::vector<size_t> refs(ref_count);
    for (size_t r = 0; r < ref_count; ++r) refs[r] = ref_first + r * ref_step;
    return oracle_vote_refs(hm, feature_mode, alpha_mode, model, n_m, scene, n_s, refs.data(), ref_count, n_threads, hyps, stats);
}

int oracle_vote_refs(const oracle_hashmap *hm, int feature_mode, int alpha_mode, const float *model, size_t n_m,
                     const float *scene, size_t n_s, const size_t *refs, size_t ref_count, int n_threads,
                     oracle_hypothesis *hyps, uint64_t *stats) {
    if (!hm || hm->n != n_m || (ref_count && !refs)) return -1;
    const uint32_t n_alpha = num_alpha_bins(hm->angle_step, hm->nalpha_rule);
    Grid grid;
    const float radius = hm->max_dist * 0.5f;
    const bool use_grid = radius > 0.0f && n_s > 2048;
    if (use_grid) grid.build(scene, n_s, radius);
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
#ifdef _OPENMP
    if (n_threads > 1 && (ref_count < 4 * static_cast<size_t>(n_threads) || n_m >= 4096)) {
        /* few reference points, or a large model: threads share each point's scene pairs (a 10 000-point model makes
         * single reference points cost from a second to minutes, so one thread per point would leave most cores
         * idle behind the expensive ones) */
        std::vector<std::vector<uint32_t>> accs(static_cast<size_t>(n_threads), std::vector<uint32_t>(n_m * n_alpha, 0u));
        for (size_t r = 0; r < ref_count; ++r) {
            size_t s_r = refs[r];
            if (s_r >= n_s) continue;
            vote_one_reference_parallel(hm, feature_mode, alpha_mode, model, n_m, scene, n_s, use_grid ? &grid : nullptr, s_r,
                                        n_alpha, accs, &hyps[r], stats);
        }
        return 0;
    }
    if (n_threads > 1) {
        uint64_t tot[4] = {0, 0, 0, 0};
#pragma omp parallel num_threads(n_threads)
        {
            std::vector<uint32_t> acc(n_m * n_alpha, 0u);
            uint64_t local[4] = {0, 0, 0, 0};
#pragma omp for schedule(dynamic, 4)
            for (long long r = 0; r < static_cast<long long>(ref_count); ++r) {
                size_t s_r = refs[static_cast<size_t>(r)];
                if (s_r >= n_s) continue;
                vote_one_reference(hm, feature_mode, alpha_mode, model, n_m, scene, n_s,
                                   use_grid ? &grid : nullptr, s_r, n_alpha, acc.data(), &hyps[r], local);
            }
#pragma omp critical
            for (int k = 0; k < 4; ++k) tot[k] += local[k];
        }
        if (stats)
            for (int k = 0; k < 4; ++k) stats[k] = tot[k];
        return 0;
    }
#else
    (void)n_threads;
#endif
    std::vector<uint32_t> acc(n_m * n_alpha, 0u);
    for (size_t r = 0; r < ref_count; ++r) {
        size_t s_r = refs[r];
        if (s_r >= n_s) continue;
        vote_one_reference(hm, feature_mode, alpha_mode, model, n_m, scene, n_s,
                           use_grid ? &grid : nullptr, s_r, n_alpha, acc.data(), &hyps[r], stats);
    }
    return 0;
}

void oracle_peak_pose(int alpha_mode, float angle_step, const float *model, size_t model_index,
                      uint32_t bin, const float *scene, size_t s_r, float *pose12) {
    Frame sg = ref_frame(ld3(scene + 6 * s_r), ld3(scene + 6 * s_r + 3));
    Frame mg = ref_frame(ld3(model + 6 * model_index), ld3(model + 6 * model_index + 3));
    compose_pose(sg, peak_theta(alpha_mode, angle_step, bin), mg, pose12);
}

int oracle_poses_within(const float *a12, const float *b12, float pos_thr, float rot_thr) {
    return poses_within(a12, b12, pos_thr, rot_thr) ? 1 : 0;
}

size_t oracle_cluster(const oracle_hypothesis *hyps, size_t n, float pos_thr, float rot_thr,
                      float *out_poses, uint32_t *out_votes, uint32_t *assignment,
                      size_t *n_clusters_out) {
    /* sort by votes descending; ties keep input order (A.8 rule 1) */
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = static_cast<uint32_t>(i);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return hyps[a].votes > hyps[b].votes; });

    std::vector<std::vector<uint32_t>> clusters;
    std::vector<std::pair<size_t, unsigned int>> cluster_votes;
    for (size_t k = 0; k < n; ++k) {
        const oracle_hypothesis &h = hyps[order[k]];
        bool found = false;
        for (size_t c = 0; c < clusters.size(); ++c) {
            if (poses_within(h.pose, hyps[clusters[c].front()].pose, pos_thr, rot_thr)) {
                found = true;
                clusters[c].push_back(order[k]);
                cluster_votes[c].second += h.votes;
                if (assignment) assignment[order[k]] = static_cast<uint32_t>(c);
                break;
            }
        }
        if (!found) {
            clusters.push_back(std::vector<uint32_t>(1, order[k]));
            cluster_votes.push_back(std::make_pair(clusters.size() - 1, h.votes));
            if (assignment) assignment[order[k]] = static_cast<uint32_t>(clusters.size() - 1);
        }
    }
    if (n_clusters_out) *n_clusters_out = clusters.size();
    /* clusters by summed votes descending; ties keep creation order (A.8 rule 2) */
    std::stable_sort(cluster_votes.begin(), cluster_votes.end(),
                     [](const std::pair<size_t, unsigned int> &a,
                        const std::pair<size_t, unsigned int> &b) { return a.second > b.second; });
    size_t n_out = clusters.size() < 3 ? clusters.size() : 3;
    for (size_t c = 0; c < n_out; ++c) {
        const std::vector<uint32_t> &members = clusters[cluster_votes[c].first];
        float t[3] = {0.0f, 0.0f, 0.0f};
        float q[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (uint32_t m : members) {
            const float *P = hyps[m].pose;
            t[0] += P[3];
            t[1] += P[7];
            t[2] += P[11];
            float R[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
            float qm[4];
            quat_from_matrix(R, qm);
            for (int k = 0; k < 4; ++k) q[k] += qm[k];
        }
        float cnt = static_cast<float>(members.size());
        for (int k = 0; k < 3; ++k) t[k] /= cnt;
        for (int k = 0; k < 4; ++k) q[k] /= cnt;
        float qn = std::sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
        if (qn > 0.0f)
            for (int k = 0; k < 4; ++k) q[k] /= qn;
        float R[9];
        quat_to_matrix(q, R);
        float *M = out_poses + 16 * c;
        M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = t[0];
        M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = t[1];
        M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = t[2];
        M[12] = 0.0f; M[13] = 0.0f; M[14] = 0.0f; M[15] = 1.0f;
        out_votes[c] = cluster_votes[c].second;
    }
    return n_out;
}

void oracle_transform(const float *cloud, size_t n, const float *M, float *out_xyz) {
    for (size_t i = 0; i < n; ++i) {
        const float *p = cloud + 6 * i;
        out_xyz[3 * i + 0] = ((M[0] * p[0] + M[1] * p[1]) + M[2] * p[2]) + M[3];
        out_xyz[3 * i + 1] = ((M[4] * p[0] + M[5] * p[1]) + M[6] * p[2]) + M[7];
        out_xyz[3 * i + 2] = ((M[8] * p[0] + M[9] * p[1]) + M[10] * p[2]) + M[11];
    }
}

size_t oracle_register(const oracle_hashmap *hm, int feature_mode, int alpha_mode,
                       const float *model, size_t n_m, const float *scene, size_t n_s,
                       size_t ref_rate, float pos_thr, float rot_thr, int n_threads,
                       float *final16, float *out_poses, uint32_t *out_votes, uint64_t *stats) {
    if (ref_rate == 0) ref_rate = 1;
    size_t ref_count = (n_s + ref_rate - 1) / ref_rate;
    std::vector<oracle_hypothesis> hyps(ref_count);
    if (oracle_vote(hm, feature_mode, alpha_mode, model, n_m, scene, n_s, 0, ref_rate, ref_count,
                    n_threads, hyps.data(), stats) != 0)
        return 0;
    size_t n_out = oracle_cluster(hyps.data(), ref_count, pos_thr, rot_thr, out_poses, out_votes,
                                  nullptr, nullptr);
    if (n_out > 0 && final16) std::memcpy(final16, out_poses, 16 * sizeof(float));
    return n_out;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
