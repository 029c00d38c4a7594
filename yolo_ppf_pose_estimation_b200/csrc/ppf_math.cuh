// ppf_math.cuh — device arithmetic of the PPF path (fp32, un-fused: build with -fmad=false).
//
// Restates, for the device, the arithmetic PCL executes on this path:
//   [PCL] features/src/pfh.cpp                         computePairFeatures      
//   [PCL] features/src/ppf.cpp                         computePPFPairFeature    (A.1')
//   [PCL] features/include/pcl/features/impl/ppf.hpp   alpha_m                  (A.2)
//   [PCL] registration/impl/ppf_registration.hpp       alpha bin, pose assembly (A.4, A.5)
//   Eigen Geometry (AngleAxis / Quaternion)                                     (A.6)
// Operation order is left-to-right, transcendental functions are CUDA's libm-accurate
// float versions (never the __fast intrinsics), sqrt and division are IEEE (nvcc defaults).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace b200ppf {

struct V3 {
    float x, y, z;
};

__host__ __device__ __forceinline__ V3 make_v3(float x, float y, float z) { return V3{x, y, z}; }
__host__ __device__ __forceinline__ V3 v3_of(const float4 &p) { return V3{p.x, p.y, p.z}; }
__host__ __device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__host__ __device__ __forceinline__ float norm3(V3 a) { return sqrtf(dot3(a, a)); }

enum { FEATURE_PCL_PFH = 0, FEATURE_DROST_COS = 1, FEATURE_DROST_ANGLE = 2 };
enum { ALPHA_MODE_A = 0, ALPHA_MODE_B = 1 };
// How many alpha columns the accumulator has and what becomes of a vote whose bin is past the last one
// (include/b200ppf.h B200PPF_NALPHA_*): CEIL = ceil(2*pi/step) columns (no such vote exists for PCL's 12
// degrees), FLOOR_DROP = floor(2*pi/step) columns, the vote is lost (PCL <= 1.11 increments one word past the
// row's heap block), FLOOR_CLAMP = floor columns, the vote joins the last column (this library's round 1).
enum { NALPHA_CEIL = 0, NALPHA_FLOOR_DROP = 1, NALPHA_FLOOR_CLAMP = 2 };

// computePairFeatures ([PCL] features/src/pfh.cpp).  f = {f1,f2,f3,f4}.
__host__ __device__ __forceinline__ bool pair_features_pfh(V3 p1, V3 n1, V3 p2, V3 n2, float *f) {
    V3 d = make_v3(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z);
    float f4 = norm3(d);
    if (f4 == 0.0f) return false;
    float angle1 = dot3(n1, d) / f4;
    float angle2 = dot3(n2, d) / f4;
    V3 u, nt;
    float f3;
    if (acosf(fabsf(angle1)) > acosf(fabsf(angle2))) {
        u = n2;
        nt = n1;
        d = make_v3(d.x * -1.0f, d.y * -1.0f, d.z * -1.0f);
        f3 = -angle2;
    } else {
        u = n1;
        nt = n2;
        f3 = angle1;
    }
    V3 v = cross3(d, u);
    float vn = norm3(v);
    if (vn == 0.0f) return false;
    v = make_v3(v.x / vn, v.y / vn, v.z / vn);
    V3 w = cross3(u, v);
    f[1] = dot3(v, nt);
    f[0] = atan2f(dot3(w, nt), dot3(u, nt));
    f[2] = f3;
    f[3] = f4;
    return true;
}

// computePPFPairFeature ([PCL] features/src/ppf.cpp): Drost tuple as cosines
__host__ __device__ __forceinline__ bool pair_features_drost_cos(V3 p1, V3 n1, V3 p2, V3 n2, float *f) {
    V3 d = make_v3(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z);
    float f4 = norm3(d);
    if (f4 == 0.0f) return false;
    d = make_v3(d.x / f4, d.y / f4, d.z / f4);
    f[0] = dot3(n1, d);
    f[1] = dot3(n2, d);
    f[2] = dot3(n1, n2);
    f[3] = f4;
    return true;
}

__host__ __device__ __forceinline__ float clamp_unit(float c) {
    return c > 1.0f ? 1.0f : (c < -1.0f ? -1.0f : c);
}

__host__ __device__ __forceinline__ bool pair_features(int mode, V3 p1, V3 n1, V3 p2, V3 n2, float *f) {
    if (mode == FEATURE_PCL_PFH) return pair_features_pfh(p1, n1, p2, n2, f);
    if (!pair_features_drost_cos(p1, n1, p2, n2, f)) return false;
    if (mode == FEATURE_DROST_ANGLE) {
        f[0] = acosf(clamp_unit(f[0]));
        f[1] = acosf(clamp_unit(f[1]));
        f[2] = acosf(clamp_unit(f[2]));
    }
    return true;
}

// Eigen::AngleAxisf::toRotationMatrix (Eigen Geometry/AngleAxis.h), R row-major
__host__ __device__ __forceinline__ void angle_axis_matrix(float angle, V3 a, float *R) {
    float s = sinf(angle), c = cosf(angle);
    V3 sa = make_v3(s * a.x, s * a.y, s * a.z);
    float omc = 1.0f - c;
    V3 ca = make_v3(omc * a.x, omc * a.y, omc * a.z);
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

__host__ __device__ __forceinline__ V3 matvec(const float *R, V3 v) {
    return make_v3((R[0] * v.x + R[1] * v.y) + R[2] * v.z, (R[3] * v.x + R[4] * v.y) + R[5] * v.z,
                   (R[6] * v.x + R[7] * v.y) + R[8] * v.z);
}

// rigid frame moving (p, n) to the origin with n on +x ([PCL] impl/ppf.hpp, transform_mg / _sg)
struct Frame {
    float R[9];
    float t[3];
};

__host__ __device__ __forceinline__ void ref_frame(V3 p, V3 n, Frame &F) {
    float angle = acosf(n.x);
    bool parallel = (n.y == 0.0f && n.z == 0.0f);
    V3 axis;
    if (parallel) {
        axis = make_v3(0.0f, 1.0f, 0.0f);
    } else {
        V3 c = make_v3(0.0f, n.z, -n.y);
        float z = (c.x * c.x + c.y * c.y) + c.z * c.z;
        if (z > 0.0f) {
            float s = sqrtf(z);
            c = make_v3(c.x / s, c.y / s, c.z / s);
        }
        axis = c;
    }
    angle_axis_matrix(angle, axis, F.R);
    V3 mp = make_v3(-1.0f * p.x, -1.0f * p.y, -1.0f * p.z);
    V3 t = matvec(F.R, mp);
    F.t[0] = t.x;
    F.t[1] = t.y;
    F.t[2] = t.z;
}

// only rows 1 and 2 of the frame are needed for alpha (y and z of the transformed point)
__host__ __device__ __forceinline__ float planar_alpha(const Frame &F, V3 m) {
    float y = ((F.R[3] * m.x + F.R[4] * m.y) + F.R[5] * m.z) + F.t[1];
    float z = ((F.R[6] * m.x + F.R[7] * m.y) + F.R[8] * m.z) + F.t[2];
    float a = atan2f(-z, y);
    if (sinf(a) * z < 0.0f) a *= -1.0f;
    return -a;
}

// ---- alpha binning ([PCL] impl/ppf_registration.hpp, voting loop) ---------------------------

// literal form: float subtract, double wrap / divide / floor.  Returns 0xFFFFFFFF for NaN.  A bin past the last
// column becomes overflow_bin: n_alpha - 1 (clamp) or n_alpha (the spare cell every accumulator row carries).
__host__ __device__ __forceinline__ uint32_t alpha_bin_exact(int mode, float angle_step, uint32_t n_alpha,
                                                             uint32_t overflow_bin, float alpha_m, float alpha_s) {
    const double PI_D = 3.14159265358979323846;
    float alpha = alpha_m - alpha_s;
    if (alpha != alpha) return 0xFFFFFFFFu;
    uint32_t bin;
    if (mode == ALPHA_MODE_B) {
        double b = (double)floorf(alpha) + floor(PI_D / (double)angle_step);
        bin = b < 0.0 ? 0u : (uint32_t)b;
    } else {
        if ((double)alpha < -PI_D) {
            alpha = (float)((double)alpha + 2.0 * PI_D);
        } else if ((double)alpha > PI_D) {
            alpha = (float)((double)alpha - 2.0 * PI_D);
        }
        double b = floor(((double)alpha + PI_D) / (double)angle_step);
        bin = b < 0.0 ? 0u : (uint32_t)b;
    }
    if (bin >= n_alpha) bin = overflow_bin;
    return bin;
}

// Hot-loop form for mode A: an fp32 estimate of (alpha+pi)/step whose error is < 1e-5 bins
// (three float roundings of magnitudes <= 2*pi, scaled by 1/step <= ~60); whenever the
// estimate lies within bp.guard of an integer — which also covers every input near the
// +-pi wrap decisions, since those map to the ends of the bin range — the literal
// double-precision form decides.  Results are therefore identical to alpha_bin_exact.
struct BinParams {
    float angle_step;
    float inv_step;
    uint32_t n_alpha;
    uint32_t overflow_bin;  // where a bin >= n_alpha goes: n_alpha - 1 (CEIL, FLOOR_CLAMP) or the spare cell n_alpha (FLOOR_DROP)
    uint32_t fold;          // FLOOR_CLAMP: the spare cell is added to column n_alpha - 1 before the peak scan
    int nalpha_rule;
    int mode;
    int mode_b_offset;  // floor(pi / angle_step)
    float guard;        // max(2e-4, 1e-5 / angle_step): > 20x the fp32 estimate's error bound
    // fixed-point form of the hot loop (alpha_bin_fixed below)
    uint32_t acc_cols;     // accumulator columns held in shared memory: n_alpha, plus the spare column n_alpha under the FLOOR rules
    uint32_t fix_mul;      // round(T * 2^fix_shift), T = 2*pi/angle_step bins per turn
    uint32_t fix_shift;    // fractional bits of fix_mul: high word of X*fix_mul = bin . (fix_shift-bit position inside the bin)
    uint32_t frac_mul;     // 2^(32 - fix_shift): a second multiply splits that word into (bin, position << (32-fix_shift))
    uint32_t fix_guard;    // guard band around bin edges, in 2^-32 bins (>= 10x the worst disagreement with PCL's floats)
    uint32_t seam_guard;   // guard band around the +-pi seam in 2^-32 turns; 0 when the bin guard already covers it
    // phase-sorted buckets (alpha mode A, T within the guard of an integer N_T): see "constant-shift voting" below
    uint32_t bulk;         // 1 when the table carries phase cells and hot words
    uint32_t n_turn;       // N_T = round(T): alpha positions per turn
    uint32_t cells_log2;   // log2 of the phase cells per bucket (0: buckets are not subdivided)
    uint32_t phase_guard;  // guard band around the scene phase, in 2^-fix_shift bins
};

// ---- constant-shift voting ---------------------------------------------------------------------
// With u_e = (alpha_m + pi) / step = B_e + f_e (integer bin, phase in [0,1)) and the scene pair's
// c = (alpha_s / step) mod T = q + phi, PCL's bin is floor((u_e - c) mod T).  When T is an integer
// N_T (up to the guard band) this is (B_e - q - [f_e < phi]) mod N_T: inside a bucket whose entries
// are ordered by phase cell, every entry of the cells below phi's cell shifts by q + 1, every entry
// of the cells above it by q, and phi's own cell decides per entry by comparing the phases.
//
// The accumulator slice is held BIN-MAJOR: word(bin, row) = bin * pitch + row, pitch = the slice's rows
// rounded up to a multiple of 32.  A hot word is the byte offset of the entry's own cell,
//   word = 4 * (B_e * pitch + row),
// and a vote with shift q' is
//   t = word - 4 * pitch * q' ;  offset = umin(t, t + 4 * pitch * N_T)
// (t is "negative", i.e. huge as an unsigned number, exactly when B_e < q', and adding one turn then wraps
// it back into range).  Because pitch is a multiple of 32, neither the shift nor the wrap changes the
// shared-memory bank of the vote: it is (row mod 32) for every scene pair, which is what lets the table
// build order each phase cell so that 32 consecutive entries hit 32 different banks (k2_table.cu).
__host__ __device__ __forceinline__ uint32_t phase_of_fix(const BinParams &bp, uint32_t a_fix) {
    return (uint32_t)(((unsigned long long)a_fix * bp.fix_mul) >> 32);  // B . f with fix_shift fractional bits
}
__host__ __device__ __forceinline__ uint32_t phase_cell(const BinParams &bp, uint32_t u) {
    return (u & ((1u << bp.fix_shift) - 1u)) >> (bp.fix_shift - bp.cells_log2);
}
// integer position B_e of a phase word (T a hair above N_T: the sliver is the start of the turn)
__host__ __device__ __forceinline__ uint32_t phase_bin(const BinParams &bp, uint32_t u) {
    uint32_t B = u >> bp.fix_shift;
    if (B >= bp.n_turn) B -= bp.n_turn;
    return B;
}
__host__ __device__ __forceinline__ uint32_t hot_word(const BinParams &bp, uint32_t pitch, uint32_t row, uint32_t u) {
    return 4u * (phase_bin(bp, u) * pitch + row);
}
// The unmerged entry arrays carry, above the 24-bit hot word, the top 8 bits of the entry's position INSIDE its phase
// cell ("sub-phase").  For a scene pair whose phase lies in the same cell, an entry with a smaller sub-phase is below
// the scene phase and one with a larger sub-phase above it — one load and one byte compare per entry; only the entries
// whose sub-phase is the scene's own (up to the guard band: sub_lo..sub_hi) need the full-precision comparison.
constexpr uint32_t HOT_MASK = 0xFFFFFFu;
__host__ __device__ __forceinline__ uint32_t sub_phase_shift(const BinParams &bp) { return bp.fix_shift - bp.cells_log2 - 8u; }
// byte offset of a vote: the hot word shifted by q' alpha positions (c = 4 * pitch * q'), wrapped
__host__ __device__ __forceinline__ uint32_t shifted_offset(uint32_t w, uint32_t c, uint32_t wrap_bytes) {
    const uint32_t t = w - c, t2 = t + wrap_bytes;
    return t < t2 ? t : t2;
}
// scene side: returns false when the whole bucket must take the per-entry path
__host__ __device__ __forceinline__ bool phase_split(const BinParams &bp, uint32_t c_s, uint32_t &q, uint32_t &cell) {
    const uint32_t c = phase_of_fix(bp, c_s);
    q = c >> bp.fix_shift;
    const uint32_t phi = c & ((1u << bp.fix_shift) - 1u);
    const uint32_t cellw = 1u << (bp.fix_shift - bp.cells_log2);
    cell = phi >> (bp.fix_shift - bp.cells_log2);
    const uint32_t in = phi & (cellw - 1u);
    return q < bp.n_turn && in >= bp.phase_guard && in < cellw - bp.phase_guard;
}

// ---- fixed-point alpha arithmetic -------------------------------------------------------------
// An angle a in [-pi, pi] is held as A = round((a + pi) / (2*pi) * 2^32) (mod 2^32), one "turn" per
// 2^32, so differences wrap for free.  For a model entry A_m and a scene pair constant
// C_s = A_s - 2^31:   X = A_m - C_s  ==  ((alpha_m - alpha_s + pi) mod 2*pi) / (2*pi) * 2^32,
// which is exactly the quantity PCL bins: floor((wrap(alpha_m - alpha_s) + pi) / step) = floor(X/2^32 * T).
// The high word of X * fix_mul is the bin followed by fix_shift bits of position inside the bin.  This evaluates the EXACT real-number bin; PCL rounds the
// difference to float (<= 3.6e-7 rad off), so whenever the position lies within fix_guard of an
// edge (or X within seam_guard of the +-pi seam) the literal double form decides instead.
__host__ __device__ __forceinline__ uint32_t alpha_to_fix(float a) {
    const double K = 4294967296.0 / 6.283185307179586476925286766559;
    const double v = ((double)a + 3.14159265358979323846) * K;
#ifdef __CUDA_ARCH__
    return (uint32_t)(unsigned long long)__double2ll_rn(v);
#else
    return (uint32_t)(unsigned long long)llrint(v);
#endif
}

// sub-phase of phase word u relative to phase cell `cell` (the table files an entry under the cell its fp32 phase
// estimate names — filed_phase_cell, within 1e-6 rad of the fixed-point phase — so an entry next to a cell edge may lie
// just outside its cell: it gets the cell's first or last sub-phase, which orders it against every scene phase the
// constant-shift path admits, those being at least phase_guard away from the cell edges)
__host__ __device__ __forceinline__ uint32_t sub_phase(const BinParams &bp, uint32_t u, uint32_t cell) {
    const uint32_t wbits = bp.fix_shift - bp.cells_log2;
    const int rel = (int)(u & ((1u << bp.fix_shift) - 1u)) - (int)(cell << wbits);
    if (rel < 0) return 0u;
    if (rel >= (int)(1u << wbits)) return 0xFFu;
    return (uint32_t)rel >> sub_phase_shift(bp);
}
// Phase cell an entry is filed under, from alpha_m in fp32: frac((alpha + pi) / step) * cells.  The estimate is within
// 1e-6 rad of the fixed-point phase the voting kernel reasons with; an entry that lands on the other side of a cell
// edge because of it is harmless, since the kernel compares every entry of the bucket in full whenever the scene phase
// is within phase_guard (>= 4e-6 rad) of a cell edge — otherwise such an entry compares with the scene phase exactly
// like the edge itself does.  The one edge that must be exact is the circular one (phase 0 == phase 1: an entry moved
// across it would sit below every scene phase instead of above): close to it the fixed-point phase decides.
__host__ __device__ __forceinline__ uint32_t filed_phase_cell(const BinParams &bp, float alpha) {
    if (!(alpha == alpha)) return 0u;
    const float u = (alpha + 3.14159274f) * bp.inv_step;
    const float fr = u - floorf(u);
    const float thr = fmaxf(1e-4f, 2e-6f * (float)bp.n_turn);  // >> the fp32 error of u (a few ulps of T)
    if (fr < thr || fr > 1.0f - thr) return phase_cell(bp, phase_of_fix(bp, alpha_to_fix(alpha)));
    const uint32_t c = (uint32_t)(fr * (float)(1u << bp.cells_log2)), last = (1u << bp.cells_log2) - 1u;
    return c < last ? c : last;
}
// word of the unmerged entry arrays of a phase-sorted table
__host__ __device__ __forceinline__ uint32_t entry_word(const BinParams &bp, uint32_t pitch, uint32_t row, float alpha) {
    const uint32_t u = phase_of_fix(bp, alpha_to_fix(alpha));
    return (sub_phase(bp, u, filed_phase_cell(bp, alpha)) << 24) | hot_word(bp, pitch, row, u);
}

// returns the bin in [0, n_alpha] (n_alpha = PCL's out-of-range bin) or 0xFFFFFFFF when the guard
// bands say "decide with the literal form"
__host__ __device__ __forceinline__ uint32_t alpha_bin_fixed(const BinParams &bp, uint32_t a_m, uint32_t c_s) {
    const uint32_t x = a_m - c_s;
    const uint32_t hi = (uint32_t)(((unsigned long long)x * bp.fix_mul) >> 32);
    // two multiplies instead of shifts: they run on the FMA pipe, the ALU pipe is the busy one
    const unsigned long long p2 = (unsigned long long)hi * bp.frac_mul;
    const uint32_t bin = (uint32_t)(p2 >> 32), frac = (uint32_t)p2;
    bool ok = (frac - bp.fix_guard) < (0u - 2u * bp.fix_guard);
    if (bp.seam_guard) ok = ok && ((x + bp.seam_guard) >= 2u * bp.seam_guard);
    return ok ? bin : 0xFFFFFFFFu;
}

// fp32 form (used by alpha mode B and kept as a cross-check of mode A)
__host__ __device__ __forceinline__ uint32_t alpha_bin_fast(const BinParams &bp, float alpha_m, float alpha_s) {
    float d = alpha_m - alpha_s;
    if (bp.mode == ALPHA_MODE_B) {
        if (d != d) return 0xFFFFFFFFu;
        int b = (int)floorf(d) + bp.mode_b_offset;
        uint32_t bin = b < 0 ? 0u : (uint32_t)b;
        return bin >= bp.n_alpha ? bp.overflow_bin : bin;
    }
    float w = d;
    // (double)d < -pi  <=>  d <= -3.14159274f (= float(pi), just beyond pi); same on the + side
    if (d <= -3.14159274f) w = d + 6.28318548f;
    else if (d >= 3.14159274f) w = d - 6.28318548f;
    float q = (w + 3.14159274f) * bp.inv_step;
    float fl = floorf(q);
    float fr = q - fl;
    if (!(fr > bp.guard && fr < 1.0f - bp.guard))  // also catches NaN
        return alpha_bin_exact(ALPHA_MODE_A, bp.angle_step, bp.n_alpha, bp.overflow_bin, alpha_m, alpha_s);
    if (fl < 0.0f) return 0u;  // only reachable for |alpha_m - alpha_s| > 3*pi; same as the literal form
    uint32_t bin = (uint32_t)(int)fl;
    return bin >= bp.n_alpha ? bp.overflow_bin : bin;
}

// what the voting kernel computes for one vote (mode A: fixed-point estimate, literal form inside
// the guard bands; mode B: the legacy formula, already cheap).  Domain: alpha_m, alpha_s in
// [-pi, pi] as produced by atan2f — the table build rejects anything else.
__host__ __device__ __forceinline__ uint32_t alpha_bin_hot(const BinParams &bp, float alpha_m, float alpha_s) {
    if (bp.mode == ALPHA_MODE_B) return alpha_bin_fast(bp, alpha_m, alpha_s);
    if (alpha_m != alpha_m || alpha_s != alpha_s) return 0xFFFFFFFFu;
    // the fixed-point form holds one turn: both angles must be atan2f results (|a| <= float(pi)), which
    // is what the table and the scene frames produce; anything else takes the literal form
    if (!(fabsf(alpha_m) <= 3.14159274f && fabsf(alpha_s) <= 3.14159274f))
        return alpha_bin_exact(ALPHA_MODE_A, bp.angle_step, bp.n_alpha, bp.overflow_bin, alpha_m, alpha_s);
    const uint32_t b = alpha_bin_fixed(bp, alpha_to_fix(alpha_m), alpha_to_fix(alpha_s) - 0x80000000u);
    if (b == 0xFFFFFFFFu) return alpha_bin_exact(ALPHA_MODE_A, bp.angle_step, bp.n_alpha, bp.overflow_bin, alpha_m, alpha_s);
    return b >= bp.n_alpha ? bp.overflow_bin : b;
}

// what the voting kernel computes for one (entry, scene pair) of a phase-sorted table: the constant-shift form
// outside the scene phase's cell, the phase comparison inside it (the literal form within the guard band)
__host__ __device__ __forceinline__ uint32_t alpha_bin_phase(const BinParams &bp, float alpha_m, float alpha_s) {
    if (!bp.bulk || bp.mode == ALPHA_MODE_B) return alpha_bin_hot(bp, alpha_m, alpha_s);
    if (alpha_m != alpha_m || alpha_s != alpha_s) return 0xFFFFFFFFu;
    if (!(fabsf(alpha_m) <= 3.14159274f && fabsf(alpha_s) <= 3.14159274f)) return alpha_bin_hot(bp, alpha_m, alpha_s);
    const uint32_t u = phase_of_fix(bp, alpha_to_fix(alpha_m));
    const uint32_t c_s = alpha_to_fix(alpha_s) - 0x80000000u;
    uint32_t q, cell;
    const bool split = phase_split(bp, c_s, q, cell);
    if (q >= bp.n_turn)  // the scene pair sits in the sliver between N_T and T: the literal form for every entry
        return alpha_bin_exact(ALPHA_MODE_A, bp.angle_step, bp.n_alpha, bp.overflow_bin, alpha_m, alpha_s);
    const uint32_t ce = filed_phase_cell(bp, alpha_m);  // the cell the table files the entry under
    uint32_t qp;
    if (split && ce != cell) {
        qp = q + (ce < cell ? 1u : 0u);
    } else {
        // the scene phase's own cell — or the whole bucket when the scene phase is close to a cell edge
        const uint32_t fmask = (1u << bp.fix_shift) - 1u;
        const uint32_t phi = phase_of_fix(bp, c_s) & fmask;
        const int d = (int)(u & fmask) - (int)phi;
        bool settled = false;
        if (split) {  // the kernel's first look: the entry word's sub-phase byte against the scene's own sub-phases
            const uint32_t in = phi & ((1u << (bp.fix_shift - bp.cells_log2)) - 1u), sh = sub_phase_shift(bp);
            const uint32_t sub_lo = (in - bp.phase_guard) >> sh, span = ((in + bp.phase_guard - 1u) >> sh) - sub_lo;
            const uint32_t rel = sub_phase(bp, u, ce) - sub_lo;
            if (rel > span) {
                qp = q + ((int)rel > (int)span ? 0u : 1u);
                settled = true;
            }
        }
        if (!settled) {
            if (((uint32_t)(d + (int)bp.phase_guard) & fmask) < 2u * bp.phase_guard)  // circular: across a bin edge too
                return alpha_bin_exact(ALPHA_MODE_A, bp.angle_step, bp.n_alpha, bp.overflow_bin, alpha_m, alpha_s);
            qp = q + (d < 0 ? 1u : 0u);
        }
    }
    // the kernel's own integer path, address arithmetic included (pitch 32, row 5)
    const uint32_t pitch = 32u, unit = 4u * pitch;
    const uint32_t off = shifted_offset(hot_word(bp, pitch, 5u, u), unit * qp, unit * bp.n_turn);
    const uint32_t b = off / unit;
    return b >= bp.n_alpha ? bp.overflow_bin : b;
}

__host__ __device__ __forceinline__ float peak_theta(int mode, float angle_step, uint32_t bin) {
    const double PI_D = 3.14159265358979323846;
    if (mode == ALPHA_MODE_B) {
        float k = (float)((double)bin - floor(PI_D / (double)angle_step));
        return k * angle_step;
    }
    float k = (float)((double)bin + 0.5);
    return (float)((double)(k * angle_step) - PI_D);
}

// pose = T_sg^-1 * Rx(theta) * T_mg, 3x4 row-major ([PCL] impl/ppf_registration.hpp)
__host__ __device__ __forceinline__ void compose_pose(const Frame &sg, float theta, const Frame &mg, float *P) {
    float Rx[9];
    angle_axis_matrix(theta, make_v3(1.0f, 0.0f, 0.0f), Rx);
    float Ri[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ri[r * 3 + c] = sg.R[c * 3 + r];
    V3 ti = matvec(Ri, make_v3(sg.t[0], sg.t[1], sg.t[2]));
    ti = make_v3(-ti.x, -ti.y, -ti.z);
    float A[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            A[r * 3 + c] = (Ri[r * 3 + 0] * Rx[0 * 3 + c] + Ri[r * 3 + 1] * Rx[1 * 3 + c]) +
                           Ri[r * 3 + 2] * Rx[2 * 3 + c];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            P[r * 4 + c] = (A[r * 3 + 0] * mg.R[0 * 3 + c] + A[r * 3 + 1] * mg.R[1 * 3 + c]) +
                           A[r * 3 + 2] * mg.R[2 * 3 + c];
    V3 t = matvec(A, make_v3(mg.t[0], mg.t[1], mg.t[2]));
    P[3] = t.x + ti.x;
    P[7] = t.y + ti.y;
    P[11] = t.z + ti.z;
}

// Eigen::Quaternionf(Matrix3f) (Eigen Geometry/Quaternion.h); q = {x,y,z,w}; R row-major 3x3
__host__ __device__ __forceinline__ void quat_from_matrix(const float *R, float *q) {
    float t = R[0] + R[4] + R[8];
    if (t > 0.0f) {
        t = sqrtf(t + 1.0f);
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
        t = sqrtf(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0f);
        float qi = 0.5f * t;
        t = 0.5f / t;
        float qw = (R[k * 3 + j] - R[j * 3 + k]) * t;
        float qj = (R[j * 3 + i] + R[i * 3 + j]) * t;
        float qk = (R[k * 3 + i] + R[i * 3 + k]) * t;
        q[3] = qw;
        // runtime-indexed stores kept explicit so the array stays in registers
        q[0] = i == 0 ? qi : (j == 0 ? qj : qk);
        q[1] = i == 1 ? qi : (j == 1 ? qj : qk);
        q[2] = i == 2 ? qi : (j == 2 ? qj : qk);
    }
}

__host__ __device__ __forceinline__ void quat_to_matrix(const float *q, float *R) {
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

// |AngleAxisf(R).angle()| (Eigen: matrix -> quaternion -> 2*atan2(|vec|, |w|))
__host__ __device__ __forceinline__ float rotation_angle(const float *R) {
    float q[4];
    quat_from_matrix(R, q);
    float n = sqrtf((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    if (n < 1.1920929e-07f) {
        float m = fmaxf(fabsf(q[0]), fmaxf(fabsf(q[1]), fabsf(q[2])));
        if (m > 0.0f) {
            float a = q[0] / m, b = q[1] / m, c = q[2] / m;
            n = m * sqrtf((a * a + b * b) + c * c);
        } else {
            n = 0.0f;
        }
    }
    if (n != 0.0f) return fabsf(2.0f * atan2f(n, fabsf(q[3])));
    return 0.0f;
}

// posesWithinErrorBounds on 3x4 row-major poses
__host__ __device__ __forceinline__ bool poses_within(const float *a, const float *b, float pos_thr, float rot_thr) {
    V3 dt = make_v3(a[3] - b[3], a[7] - b[7], a[11] - b[11]);
    if (!(norm3(dt) < pos_thr)) return false;
    float M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            M[r * 3 + c] = (a[0 * 4 + r] * b[0 * 4 + c] + a[1 * 4 + r] * b[1 * 4 + c]) + a[2 * 4 + r] * b[2 * 4 + c];
    return rotation_angle(M) < rot_thr;
}

// ---- key quantisation ([PCL] registration/src/ppf_registration.cpp) ---------------------------
struct KeyParams {
    float angle_step, dist_step;
    int lo[4];    // lower bound of each quantised component
    int size[4];  // extent of each component
    uint32_t key_space;   // size[0]*size[1]*size[2]*size[3]
    uint32_t slice_rows;  // model rows per accumulator slice
    uint32_t n_slices;
    uint32_t row_pitch;   // slice_rows rounded up to a multiple of 32: words between two alpha columns of the accumulator
    uint32_t merge_bits;  // bits of alpha_m's bin appended to the sort key (phase-sorted tables with merged votes), else 0
};

__host__ __device__ __forceinline__ void quantise(const KeyParams &kp, const float *f, int *d) {
    d[0] = (int)floorf(f[0] / kp.angle_step);
    d[1] = (int)floorf(f[1] / kp.angle_step);
    d[2] = (int)floorf(f[2] / kp.angle_step);
    d[3] = (int)floorf(f[3] / kp.dist_step);
}

// dense packed key; returns false when a component leaves the table's range (no such bucket)
__host__ __device__ __forceinline__ bool pack_key(const KeyParams &kp, const int *d, uint32_t &key) {
    int a = d[0] - kp.lo[0], b = d[1] - kp.lo[1], c = d[2] - kp.lo[2], e = d[3] - kp.lo[3];
    if ((unsigned)a >= (unsigned)kp.size[0] || (unsigned)b >= (unsigned)kp.size[1] ||
        (unsigned)c >= (unsigned)kp.size[2] || (unsigned)e >= (unsigned)kp.size[3])
        return false;
    key = (((uint32_t)e * kp.size[0] + a) * kp.size[1] + b) * kp.size[2] + c;
    return true;
}

}  // namespace b200ppf
