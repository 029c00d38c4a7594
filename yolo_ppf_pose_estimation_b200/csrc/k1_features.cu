// k1_features.cu — K1 ppf_model_features: all ordered model pairs (i, j) -> (f1..f4, alpha_m).
//
// Replaces PPFEstimation<PointNormal,PointNormal,PPFSignature>::computeFeature
// ([PCL] features/include/pcl/features/impl/ppf.hpp): output row-major
// [i*N + j], invalid pairs (i == j, coincident points, normal parallel to d) all-NaN.
//
// One block = TI reference rows x 256 consecutive j.  The TI reference points and their rigid
// frames (which depend on i only; PCL recomputes them for every j) are staged in shared memory,
// the j point is a coalesced float4 SoA load held in registers across the TI rows, and the
// 20-byte signatures are transposed through shared memory so that the N^2 * 20 B output — the
// only HBM traffic that matters here — leaves as fully coalesced 4-byte-lane stores.
// Roofline: HBM write, 20 B per ordered pair.
#include "ppf_common.cuh"

namespace b200ppf {

namespace {

constexpr int TI = 4;
constexpr int TJ = 256;

struct RefStage {
    float px, py, pz, nx, ny, nz;
    Frame F;
};

__global__ void __launch_bounds__(TJ)
ppf_model_features_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm, uint32_t n,
                          int feature_mode, float *__restrict__ out /* n*n*5 */) {
    __shared__ RefStage ref[TI];
    __shared__ float stage[TJ * 5];
    const uint32_t i0 = blockIdx.y * TI;
    const uint32_t j0 = blockIdx.x * TJ;
    const uint32_t j = j0 + threadIdx.x;
    if (threadIdx.x < TI && i0 + threadIdx.x < n) {
        float4 p = pos[i0 + threadIdx.x], q = nrm[i0 + threadIdx.x];
        RefStage &r = ref[threadIdx.x];
        r.px = p.x; r.py = p.y; r.pz = p.z;
        r.nx = q.x; r.ny = q.y; r.nz = q.z;
        ref_frame(make_v3(p.x, p.y, p.z), make_v3(q.x, q.y, q.z), r.F);
    }
    float4 pj = make_float4(0, 0, 0, 0), nj = make_float4(0, 0, 0, 0);
    if (j < n) {
        pj = pos[j];
        nj = nrm[j];
    }
    __syncthreads();
    const uint32_t tile_w = min((uint32_t)TJ, n - j0);
    const float qnan = CUDART_NAN_F;
#pragma unroll 1
    for (int r = 0; r < TI; ++r) {
        const uint32_t i = i0 + r;
        if (i >= n) break;
        float sig[5] = {qnan, qnan, qnan, qnan, qnan};
        if (j < n && i != j) {
            float f[4];
            V3 pi = make_v3(ref[r].px, ref[r].py, ref[r].pz), ni = make_v3(ref[r].nx, ref[r].ny, ref[r].nz);
            if (pair_features(feature_mode, pi, ni, v3_of(pj), v3_of(nj), f)) {
                sig[0] = f[0]; sig[1] = f[1]; sig[2] = f[2]; sig[3] = f[3];
                sig[4] = planar_alpha(ref[r].F, v3_of(pj));
            }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) stage[threadIdx.x * 5 + k] = sig[k];
        __syncthreads();
        float *dst = out + ((size_t)i * n + j0) * 5;
        for (uint32_t e = threadIdx.x; e < tile_w * 5; e += TJ) dst[e] = stage[e];
        __syncthreads();
    }
}

}  // namespace

int k1_features_compute(b200ppf_ctx *ctx, const b200ppf_cloud *model, b200ppf_signature *out) {
    const uint32_t n = (uint32_t)model->n;
    if (n == 0) return B200PPF_OK;
    dim3 grid((n + TJ - 1) / TJ, (n + TI - 1) / TI);
    PPF_LAUNCH(ctx, ppf_model_features_kernel, grid, TJ, 0, model->pos, model->nrm, n, ctx->feature_mode,
               reinterpret_cast<float *>(out));
    return B200PPF_OK;
}

}  // namespace b200ppf
