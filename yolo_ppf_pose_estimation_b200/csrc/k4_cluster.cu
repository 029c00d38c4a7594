// k4_cluster.cu — K4 ppf_cluster_average: greedy pose clustering and averaging on the device.
//
// Replaces PPFRegistration::clusterPoses / posesWithinErrorBounds
// ([PCL] registration/include/pcl/registration/impl/ppf_registration.hpp):
//   sort hypotheses by votes (descending; ties keep reference order, A.8 rule 1);
//   each hypothesis joins the FIRST cluster (creation order) whose leader pose is within
//   (position, rotation) bounds, else it founds a new cluster; cluster votes are summed;
//   the three best clusters (votes descending, ties by creation order) are averaged
//   (mean translation, mean quaternion coefficients, normalised).
//
// The greedy rule is sequential in the sorted order, but only through the set of leaders.  The
// device version walks the sorted list in batches of BATCH poses:
//   match    every (batch pose, existing leader) pair is tested in parallel — leaders are staged
//            in shared memory a tile at a time, the lowest matching leader wins via atomicMin;
//   resolve  one CTA settles the poses no earlier leader claimed: the first still-unclaimed pose
//            of the batch becomes a leader, every later unclaimed pose tests against it in
//            parallel, repeat.  The serial depth is the number of NEW leaders per batch.
// Exactly the PCL assignment results, with O(P*C) tests spread over the whole chip.
#include <algorithm>

#include "ppf_common.cuh"

namespace b200ppf {

namespace {

constexpr int BATCH = 1024;
constexpr int LEADER_TILE = 128;
constexpr uint32_t NONE = 0xFFFFFFFFu;

struct PoseRows {  // 3x4 row-major pose as three float4 rows
    float4 r0, r1, r2;
};

__device__ __forceinline__ void rows_to_array(const PoseRows &p, float *a) {
    a[0] = p.r0.x; a[1] = p.r0.y; a[2] = p.r0.z; a[3] = p.r0.w;
    a[4] = p.r1.x; a[5] = p.r1.y; a[6] = p.r1.z; a[7] = p.r1.w;
    a[8] = p.r2.x; a[9] = p.r2.y; a[10] = p.r2.z; a[11] = p.r2.w;
}

__global__ void cluster_keys_kernel(const b200ppf_hypothesis *__restrict__ hyps, uint32_t n,
                                    uint32_t *__restrict__ keys) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) keys[p] = ~hyps[p].votes;  // ascending ~votes == descending votes
}

__global__ void cluster_gather_kernel(const b200ppf_hypothesis *__restrict__ hyps, const uint32_t *__restrict__ order,
                                      uint32_t n, PoseRows *__restrict__ poses, uint32_t *__restrict__ votes) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const b200ppf_hypothesis &h = hyps[order[k]];
    PoseRows p;
    p.r0 = make_float4(h.pose[0], h.pose[1], h.pose[2], h.pose[3]);
    p.r1 = make_float4(h.pose[4], h.pose[5], h.pose[6], h.pose[7]);
    p.r2 = make_float4(h.pose[8], h.pose[9], h.pose[10], h.pose[11]);
    poses[k] = p;
    votes[k] = h.votes;
}

// batch poses [b0, b1) against the leaders that existed before the batch
__global__ void __launch_bounds__(256)
cluster_match_kernel(const PoseRows *__restrict__ poses, uint32_t b0, uint32_t b1,
                     const PoseRows *__restrict__ leader_pose, const uint32_t *__restrict__ n_leaders_ptr,
                     float pos_thr, float rot_thr, uint32_t *__restrict__ match /* [BATCH] */) {
    __shared__ PoseRows tile[LEADER_TILE];
    const uint32_t n_leaders = *n_leaders_ptr;
    const uint32_t l0 = blockIdx.x * LEADER_TILE;
    if (l0 >= n_leaders) return;
    const uint32_t tl = min((uint32_t)LEADER_TILE, n_leaders - l0);
    for (uint32_t t = threadIdx.x; t < tl; t += blockDim.x) tile[t] = leader_pose[l0 + t];
    __syncthreads();
    for (uint32_t k = b0 + threadIdx.x; k < b1; k += blockDim.x) {
        if (match[k - b0] <= l0) continue;  // an earlier tile already claimed it (benign race: only a shortcut)
        float a[12], b[12];
        rows_to_array(poses[k], a);
        for (uint32_t t = 0; t < tl; ++t) {
            // cheap translation reject first
            const float dx = a[3] - tile[t].r0.w, dy = a[7] - tile[t].r1.w, dz = a[11] - tile[t].r2.w;
            if (!(sqrtf((dx * dx + dy * dy) + dz * dz) < pos_thr)) continue;
            rows_to_array(tile[t], b);
            if (poses_within(a, b, pos_thr, rot_thr)) {
                atomicMin(&match[k - b0], l0 + t);
                break;
            }
        }
    }
}

// settle one batch (single CTA of BATCH threads)
__global__ void __launch_bounds__(BATCH)
cluster_resolve_kernel(const PoseRows *__restrict__ poses, uint32_t b0, uint32_t b1, PoseRows *__restrict__ leader_pose,
                       uint32_t *__restrict__ n_leaders_ptr, float pos_thr, float rot_thr,
                       uint32_t *__restrict__ match, uint32_t *__restrict__ assign_sorted) {
    __shared__ uint32_t open_mask[BATCH / 32];
    __shared__ PoseRows cur;
    __shared__ uint32_t s_leaders;
    const uint32_t t = threadIdx.x, k = b0 + t;
    const bool live = k < b1;
    uint32_t m = live ? match[t] : 0u;
    float a[12];
    if (live) rows_to_array(poses[k], a);
    if (t == 0) s_leaders = *n_leaders_ptr;
    const uint32_t open = __ballot_sync(0xFFFFFFFFu, live && m == NONE);
    if ((t & 31) == 0) open_mask[t >> 5] = open;
    __syncthreads();
    uint32_t w0 = 0;
    for (;;) {
        // first still-unclaimed pose of the batch (every thread scans the same shared words)
        uint32_t first = NONE;
        for (uint32_t w = w0; w < BATCH / 32; ++w) {
            const uint32_t bits = open_mask[w];
            if (bits) {
                first = w * 32 + (__ffs(bits) - 1);
                w0 = w;
                break;
            }
        }
        if (first == NONE) break;
        __syncthreads();  // everyone has read open_mask before it changes
        if (t == first) {
            m = s_leaders;
            cur = poses[k];
            leader_pose[m] = cur;
            s_leaders = m + 1;
            atomicAnd(&open_mask[t >> 5], ~(1u << (t & 31)));
        }
        __syncthreads();
        if (live && m == NONE && t > first) {
            float b[12];
            rows_to_array(cur, b);
            if (poses_within(a, b, pos_thr, rot_thr)) {
                m = s_leaders - 1;
                atomicAnd(&open_mask[t >> 5], ~(1u << (t & 31)));
            }
        }
        __syncthreads();
    }
    if (live) assign_sorted[k] = m;
    match[t] = NONE;  // ready for the next batch
    __syncthreads();
    if (t == 0) *n_leaders_ptr = s_leaders;
}

__global__ void cluster_votes_kernel(const uint32_t *__restrict__ assign_sorted, const uint32_t *__restrict__ votes,
                                     const uint32_t *__restrict__ order, uint32_t n,
                                     uint32_t *__restrict__ cluster_votes, uint32_t *__restrict__ cluster_size,
                                     uint32_t *__restrict__ assign_input) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t c = assign_sorted[k];
    atomicAdd(&cluster_votes[c], votes[k]);
    atomicAdd(&cluster_size[c], 1u);
    assign_input[order[k]] = c;
}

// three best clusters: votes descending, creation index ascending (one CTA)
__global__ void __launch_bounds__(1024)
cluster_top3_kernel(const uint32_t *__restrict__ cluster_votes, const uint32_t *__restrict__ n_leaders_ptr,
                    uint32_t *__restrict__ top /* [3] cluster ids, NONE if absent */) {
    __shared__ unsigned long long s_best[32];
    __shared__ uint32_t chosen[3];
    const uint32_t nc = *n_leaders_ptr;
    for (int round = 0; round < 3; ++round) {
        unsigned long long best = 0;
        for (uint32_t c = threadIdx.x; c < nc; c += blockDim.x) {
            bool taken = false;
            for (int q = 0; q < round; ++q) taken |= (chosen[q] == c);
            if (taken) continue;
            // +1 so that a zero-vote cluster still beats "nothing"
            unsigned long long key = (((unsigned long long)cluster_votes[c] + 1ull) << 32) | (0xFFFFFFFFu - c);
            best = max(best, key);
        }
        for (int o = 16; o > 0; o >>= 1) best = max(best, (unsigned long long)__shfl_xor_sync(0xFFFFFFFFu, best, o));
        if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 32; ++w) best = max(best, s_best[w]);
            chosen[round] = best ? 0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFu) : NONE;
            top[round] = chosen[round];
        }
        __syncthreads();
    }
}

// mean translation + mean quaternion of one top cluster per CTA (deterministic tree reduction)
__global__ void __launch_bounds__(256)
cluster_average_kernel(const PoseRows *__restrict__ poses, const uint32_t *__restrict__ assign_sorted, uint32_t n,
                       const uint32_t *__restrict__ top, const uint32_t *__restrict__ cluster_votes,
                       const uint32_t *__restrict__ cluster_size, float *__restrict__ out_poses /* [3][16] */,
                       uint32_t *__restrict__ out_votes /* [3] */) {
    __shared__ double red[256][7];
    const uint32_t c = top[blockIdx.x];
    if (c == NONE) {
        if (threadIdx.x == 0) out_votes[blockIdx.x] = 0;
        return;
    }
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        if (assign_sorted[k] != c) continue;
        float a[12], q[4];
        rows_to_array(poses[k], a);
        const float R[9] = {a[0], a[1], a[2], a[4], a[5], a[6], a[8], a[9], a[10]};
        quat_from_matrix(R, q);
        acc[0] += a[3]; acc[1] += a[7]; acc[2] += a[11];
        acc[3] += q[0]; acc[4] += q[1]; acc[5] += q[2]; acc[6] += q[3];
    }
    for (int v = 0; v < 7; ++v) red[threadIdx.x][v] = acc[v];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int v = 0; v < 7; ++v) red[threadIdx.x][v] += red[threadIdx.x + s][v];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float cnt = (float)cluster_size[c];
        float t[3], q[4];
        for (int v = 0; v < 3; ++v) t[v] = (float)red[0][v] / cnt;
        for (int v = 0; v < 4; ++v) q[v] = (float)red[0][3 + v] / cnt;
        const float qn = sqrtf(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
        if (qn > 0.0f)
            for (int v = 0; v < 4; ++v) q[v] /= qn;
        float R[9];
        quat_to_matrix(q, R);
        float *M = out_poses + 16 * blockIdx.x;
        M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = t[0];
        M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = t[1];
        M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = t[2];
        M[12] = 0.0f; M[13] = 0.0f; M[14] = 0.0f; M[15] = 1.0f;
        out_votes[blockIdx.x] = cluster_votes[c];
    }
}

__global__ void fill_u32_kernel(uint32_t *p, uint32_t n, uint32_t v) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

int k4_cluster(b200ppf_ctx *ctx, const b200ppf_hypothesis *hyps, size_t n_, float pos_thr, float rot_thr,
               float *poses16, uint32_t *votes_out, size_t *n_out) {
    *n_out = 0;
    if (n_ == 0) return B200PPF_OK;
    if (n_ >= 0x7FFFFFFFull) return fail_msg(ctx, B200PPF_ERR_UNSUPPORTED, "cluster: too many hypotheses");
    const uint32_t n = (uint32_t)n_;
    cudaStream_t st = ctx->stream;
    uint32_t *keys[2] = {nullptr, nullptr}, *order[2] = {nullptr, nullptr};
    PoseRows *poses = nullptr, *leader_pose = nullptr;
    uint32_t *votes = nullptr, *assign_sorted = nullptr, *match = nullptr, *cl_votes = nullptr, *cl_size = nullptr;
    uint32_t *small = nullptr;  // n_leaders, top[3], out_votes[3]
    float *d_out = nullptr;
    cudaEventRecord(ctx->ev[0], st);
    for (int b = 0; b < 2; ++b) {
        PPF_CUDA(ctx, cudaMallocAsync(&keys[b], n * sizeof(uint32_t), st));
        PPF_CUDA(ctx, cudaMallocAsync(&order[b], n * sizeof(uint32_t), st));
    }
    PPF_CUDA(ctx, cudaMallocAsync(&poses, n * sizeof(PoseRows), st));
    PPF_CUDA(ctx, cudaMallocAsync(&leader_pose, n * sizeof(PoseRows), st));
    PPF_CUDA(ctx, cudaMallocAsync(&votes, n * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMallocAsync(&assign_sorted, n * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMallocAsync(&match, BATCH * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMallocAsync(&cl_votes, n * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMallocAsync(&cl_size, n * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMallocAsync(&small, 8 * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMallocAsync(&d_out, 3 * 16 * sizeof(float), st));
    PPF_CUDA(ctx, cudaMemsetAsync(cl_votes, 0, n * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMemsetAsync(cl_size, 0, n * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMemsetAsync(small, 0, 8 * sizeof(uint32_t), st));
    PPF_CUDA(ctx, cudaMemsetAsync(d_out, 0, 3 * 16 * sizeof(float), st));
    if (ctx->assign_n < n) {
        if (ctx->d_assign) cudaFree(ctx->d_assign);
        ctx->d_assign = nullptr;
        ctx->assign_n = 0;
        PPF_CUDA(ctx, cudaMalloc(&ctx->d_assign, n * sizeof(uint32_t)));
    }
    ctx->assign_n = n;

    const unsigned g = (n + 255) / 256;
    PPF_LAUNCH(ctx, cluster_keys_kernel, g, 256, 0, hyps, n, keys[0]);
    bool in_alt = false;
    int rc = radix_sort_u32(ctx, keys[0], keys[1], order[0], order[1], nullptr, nullptr, n, 32, /*v0_iota=*/true, &in_alt);
    if (rc) return rc;
    const uint32_t *ord = order[in_alt ? 1 : 0];
    PPF_LAUNCH(ctx, cluster_gather_kernel, g, 256, 0, hyps, ord, n, poses, votes);
    PPF_LAUNCH(ctx, fill_u32_kernel, (BATCH + 255) / 256, 256, 0, match, (uint32_t)BATCH, NONE);
    uint32_t *n_leaders = small;
    for (uint32_t b0 = 0; b0 < n; b0 += BATCH) {
        const uint32_t b1 = std::min(n, b0 + BATCH);
        if (b0 > 0) {
            const unsigned tiles = (b0 + LEADER_TILE - 1) / LEADER_TILE;  // upper bound on existing leaders
            PPF_LAUNCH(ctx, cluster_match_kernel, tiles, 256, 0, poses, b0, b1, leader_pose, n_leaders, pos_thr, rot_thr,
                       match);
        }
        PPF_LAUNCH(ctx, cluster_resolve_kernel, 1, BATCH, 0, poses, b0, b1, leader_pose, n_leaders, pos_thr, rot_thr,
                   match, assign_sorted);
    }
    PPF_LAUNCH(ctx, cluster_votes_kernel, g, 256, 0, assign_sorted, votes, ord, n, cl_votes, cl_size, ctx->d_assign);
    PPF_LAUNCH(ctx, cluster_top3_kernel, 1, 1024, 0, cl_votes, n_leaders, small + 1);
    PPF_LAUNCH(ctx, cluster_average_kernel, 3, 256, 0, poses, assign_sorted, n, small + 1, cl_votes, cl_size, d_out,
               small + 4);
    uint32_t h_small[8];
    float h_out[48];
    PPF_CUDA(ctx, cudaMemcpyAsync(h_small, small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
    PPF_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, st));
    cudaEventRecord(ctx->ev[1], st);
    PPF_CUDA(ctx, cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->timings.cluster_ms, ctx->ev[0], ctx->ev[1]);
    ctx->n_clusters = h_small[0];
    size_t k = 0;
    for (int r = 0; r < 3; ++r) {
        if (h_small[1 + r] == NONE) break;
        memcpy(poses16 + 16 * k, h_out + 16 * r, 16 * sizeof(float));
        votes_out[k] = h_small[4 + r];
        ++k;
    }
    *n_out = k;
    for (int b = 0; b < 2; ++b) {
        cudaFreeAsync(keys[b], st);
        cudaFreeAsync(order[b], st);
    }
    cudaFreeAsync(poses, st);
    cudaFreeAsync(leader_pose, st);
    cudaFreeAsync(votes, st);
    cudaFreeAsync(assign_sorted, st);
    cudaFreeAsync(match, st);
    cudaFreeAsync(cl_votes, st);
    cudaFreeAsync(cl_size, st);
    cudaFreeAsync(small, st);
    cudaFreeAsync(d_out, st);
    return B200PPF_OK;
}

}  // namespace b200ppf
