// radix_sort.cu — hand-written stable LSD radix sort (8-bit digits) for sm_100a.
//
// Used by K2 (order the N_m^2 model pairs by packed feature key -> CSR buckets, replacing the
// unordered_multimap of [PCL] registration/src/ppf_registration.cpp), by the scene grid and by K4
// (order pose hypotheses by votes, replacing the std::sort of clusterPoses).
//
// One pass = tile histogram -> column scan -> tile scatter.  A tile is WARPS x ROUNDS x 32 consecutive
// elements (4096 for large inputs, 512 for small ones so that every SM has a block to run); element
// order inside a tile is (warp, round, lane).
//   hist     per-tile digit counts with shared-memory atomics -> hist[tile][256] (coalesced 1 KB rows)
//   scan     column scan over tiles (three small kernels) -> global base of every (tile, digit)
//   scatter  the tile is sorted locally first: per-warp digit counts (match.any), a (digit, warp) scan in
//            shared memory, stable ranks by (warp, round, lane), keys and payloads staged in shared memory
//            at their tile-local sorted position; then the block copies the tile out digit run by digit
//            run — consecutive threads write consecutive addresses instead of 32 scattered 4-byte stores.
// No global atomics; stable.  HBM traffic per pass: keys read twice, payloads read once, everything
// written once, in runs of (tile / 256 ... tile) elements.
#include "ppf_common.cuh"

namespace b200ppf {

namespace {

constexpr int RADIX = 256;
constexpr int WARPS = 8;            // warps per block / tile
constexpr int ROUNDS_BIG = 16;      // 4096-element tiles
constexpr int ROUNDS_SMALL = 2;     // 512-element tiles
constexpr int SCAN_CHUNK = 256;     // histogram rows per column-scan chunk

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

template <int ROUNDS>
__global__ void __launch_bounds__(WARPS * 32)
radix_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, int shift, uint32_t *__restrict__ hist) {
    constexpr uint32_t TILE = WARPS * ROUNDS * 32;
    __shared__ uint32_t cnt[RADIX];
    for (int d = threadIdx.x; d < RADIX; d += WARPS * 32) cnt[d] = 0;
    __syncthreads();
    const uint32_t begin = blockIdx.x * TILE;
    const uint32_t end = min(n, begin + TILE);  // begin < n, no overflow: n < 2^32 - TILE
#pragma unroll 4
    for (uint32_t idx = begin + threadIdx.x; idx < end; idx += WARPS * 32)
        atomicAdd(&cnt[(keys[idx] >> shift) & (RADIX - 1)], 1u);
    __syncthreads();
    uint32_t *row = hist + (size_t)blockIdx.x * RADIX;
    for (int d = threadIdx.x; d < RADIX; d += WARPS * 32) row[d] = cnt[d];
}

// column scan, step 1: per chunk of rows, per digit totals
__global__ void __launch_bounds__(RADIX)
radix_col_reduce_kernel(const uint32_t *__restrict__ hist, uint32_t nseg, uint32_t *__restrict__ chunk_tot) {
    const uint32_t d = threadIdx.x;
    const uint32_t r0 = blockIdx.x * SCAN_CHUNK, r1 = min(nseg, r0 + SCAN_CHUNK);
    uint32_t s = 0;
    for (uint32_t r = r0; r < r1; ++r) s += hist[(size_t)r * RADIX + d];
    chunk_tot[(size_t)blockIdx.x * RADIX + d] = s;
}

// step 2 (one block): exclusive scan over chunks per digit, then over digits
__global__ void __launch_bounds__(RADIX)
radix_col_scan_chunks_kernel(uint32_t *__restrict__ chunk_tot, uint32_t nchunks, uint32_t *__restrict__ digit_base) {
    __shared__ uint32_t tot[RADIX];
    const uint32_t d = threadIdx.x;
    uint32_t run = 0;
    for (uint32_t c = 0; c < nchunks; ++c) {
        uint32_t t = chunk_tot[(size_t)c * RADIX + d];
        chunk_tot[(size_t)c * RADIX + d] = run;
        run += t;
    }
    tot[d] = run;
    __syncthreads();
    // exclusive scan of 256 totals: warp 0, 8 values per lane
    if (d < 32) {
        uint32_t v[8], s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = tot[d * 8 + k];
            s += v[k];
        }
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (d >= (uint32_t)o) incl += t;
        }
        uint32_t excl = incl - s;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            digit_base[d * 8 + k] = excl;
            excl += v[k];
        }
    }
}

// step 3: rewrite every histogram row as global bases
__global__ void __launch_bounds__(RADIX)
radix_col_apply_kernel(uint32_t *__restrict__ hist, uint32_t nseg, const uint32_t *__restrict__ chunk_tot,
                       const uint32_t *__restrict__ digit_base) {
    const uint32_t d = threadIdx.x;
    const uint32_t r0 = blockIdx.x * SCAN_CHUNK, r1 = min(nseg, r0 + SCAN_CHUNK);
    uint32_t run = chunk_tot[(size_t)blockIdx.x * RADIX + d] + digit_base[d];
    for (uint32_t r = r0; r < r1; ++r) {
        uint32_t t = hist[(size_t)r * RADIX + d];
        hist[(size_t)r * RADIX + d] = run;
        run += t;
    }
}

template <int ROUNDS, bool IOTA, bool HAS_V0, bool HAS_V1>
__global__ void __launch_bounds__(WARPS * 32)
radix_scatter_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ v0,
                     const uint32_t *__restrict__ v1, uint32_t n, int shift, const uint32_t *__restrict__ hist,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ v0_out, uint32_t *__restrict__ v1_out) {
    constexpr uint32_t TILE = WARPS * ROUNDS * 32;
    extern __shared__ uint32_t smem[];
    uint32_t *s_key = smem;                                  // [TILE] tile-local sorted order
    uint32_t *s_v0 = s_key + TILE;                           // [TILE]
    uint32_t *s_v1 = s_v0 + (HAS_V0 ? TILE : 0);             // [TILE]
    uint32_t *run = s_v1 + (HAS_V1 ? TILE : 0);              // [WARPS][RADIX] next local slot of (warp, digit)
    uint32_t *dstart = run + WARPS * RADIX;                  // [RADIX] first local slot of a digit
    uint32_t *gbase = dstart + RADIX;                        // [RADIX] global base of (this tile, digit)
    __shared__ uint32_t s_tot[8];

    const uint32_t tid = threadIdx.x, w = tid >> 5, lane = lane_id();
    const uint32_t begin = blockIdx.x * TILE;
    const uint32_t end = min(n, begin + TILE);
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t k = tid; k < WARPS * RADIX; k += WARPS * 32) run[k] = 0;
    gbase[tid] = hist[(size_t)blockIdx.x * RADIX + tid];
    __syncthreads();

    // 1. keys into registers; per-warp digit counts
    uint32_t key[ROUNDS];
    uint32_t *mine = run + w * RADIX;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t idx = begin + (w * ROUNDS + r) * 32 + lane;
        const bool valid = idx < end;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid);
        key[r] = 0;
        if (valid) {
            key[r] = keys[idx];
            const uint32_t digit = (key[r] >> shift) & (RADIX - 1);
            const uint32_t peers = __match_any_sync(mask, digit);
            if ((uint32_t)(__ffs(peers) - 1) == lane) mine[digit] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // 2. (digit, warp) exclusive scan: thread d owns digit d
    {
        uint32_t c[WARPS], tot = 0;
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) {
            c[ww] = run[ww * RADIX + tid];
            tot += c[ww];
        }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_tot[w] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) before += (ww < (int)w) ? s_tot[ww] : 0u;
        uint32_t excl = before + incl - tot;
        dstart[tid] = excl;
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) {
            run[ww * RADIX + tid] = excl;
            excl += c[ww];
        }
    }
    __syncthreads();
    // 3. stable tile-local positions; stage keys and payloads in sorted order
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t idx = begin + (w * ROUNDS + r) * 32 + lane;
        const bool valid = idx < end;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid);
        uint32_t digit = 0, peers = 0, pos = 0;
        if (valid) {
            digit = (key[r] >> shift) & (RADIX - 1);
            peers = __match_any_sync(mask, digit);
            pos = mine[digit] + __popc(peers & lt);
        }
        __syncwarp();
        if (valid && (uint32_t)(__ffs(peers) - 1) == lane) mine[digit] += __popc(peers);
        __syncwarp();
        if (valid) {
            s_key[pos] = key[r];
            if (HAS_V0) s_v0[pos] = IOTA ? idx : v0[idx];
            if (HAS_V1) s_v1[pos] = v1[idx];
        }
    }
    __syncthreads();
    // 4. copy out: local slot p of digit d goes to gbase[d] + (p - dstart[d]); runs are contiguous
    const uint32_t count = end - begin;
    for (uint32_t p = tid; p < count; p += WARPS * 32) {
        const uint32_t k = s_key[p];
        const uint32_t dg = (k >> shift) & (RADIX - 1);
        const uint32_t o = gbase[dg] + (p - dstart[dg]);
        keys_out[o] = k;
        if (HAS_V0) v0_out[o] = s_v0[p];
        if (HAS_V1) v1_out[o] = s_v1[p];
    }
}

}  // namespace

// v0_iota: on entry v0 is taken to be the identity permutation 0..n-1 (its contents are not read).
namespace {

template <int ROUNDS>
int radix_sort_impl(b200ppf_ctx *ctx, uint32_t *keys, uint32_t *keys_alt, uint32_t *v0, uint32_t *v0_alt, uint32_t *v1,
                    uint32_t *v1_alt, size_t n, int bits, bool v0_iota, bool *result_in_alt) {
    constexpr uint32_t TILE = WARPS * ROUNDS * 32;
    const uint32_t ntiles = (uint32_t)((n + TILE - 1) / TILE);
    const uint32_t nchunks = (ntiles + SCAN_CHUNK - 1) / SCAN_CHUNK;
    uint32_t *hist = nullptr, *chunk_tot = nullptr, *digit_base = nullptr;
    PPF_CUDA(ctx, cudaMallocAsync(&hist, (size_t)ntiles * RADIX * sizeof(uint32_t), ctx->stream));
    PPF_CUDA(ctx, cudaMallocAsync(&chunk_tot, (size_t)nchunks * RADIX * sizeof(uint32_t), ctx->stream));
    PPF_CUDA(ctx, cudaMallocAsync(&digit_base, RADIX * sizeof(uint32_t), ctx->stream));

    const bool has_v0 = v0 != nullptr && v0_alt != nullptr, has_v1 = v1 != nullptr && v1_alt != nullptr;
    bool iota = has_v0 && v0_iota;
    const size_t smem = ((size_t)TILE * (1 + (has_v0 ? 1 : 0) + (has_v1 ? 1 : 0)) + WARPS * RADIX + 2 * RADIX) * sizeof(uint32_t);
    uint32_t *ki = keys, *ko = keys_alt, *v0i = v0, *v0o = v0_alt, *v1i = v1, *v1o = v1_alt;
    bool in_alt = false;
    for (int shift = 0; shift < bits; shift += 8) {
        PPF_LAUNCH(ctx, radix_hist_kernel<ROUNDS>, ntiles, WARPS * 32, 0, ki, (uint32_t)n, shift, hist);
        PPF_LAUNCH(ctx, radix_col_reduce_kernel, nchunks, RADIX, 0, hist, ntiles, chunk_tot);
        PPF_LAUNCH(ctx, radix_col_scan_chunks_kernel, 1, RADIX, 0, chunk_tot, nchunks, digit_base);
        PPF_LAUNCH(ctx, radix_col_apply_kernel, nchunks, RADIX, 0, hist, ntiles, chunk_tot, digit_base);
#define SCATTER(I, A, B)                                                                                                \
    do {                                                                                                                 \
        PPF_CUDA(ctx, cudaFuncSetAttribute(radix_scatter_kernel<ROUNDS, I, A, B>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                     \
        PPF_LAUNCH(ctx, (radix_scatter_kernel<ROUNDS, I, A, B>), ntiles, WARPS * 32, smem, ki, v0i, v1i, (uint32_t)n,    \
                   shift, hist, ko, v0o, v1o);                                                                           \
    } while (0)
        if (has_v0 && has_v1) {
            if (iota) SCATTER(true, true, true); else SCATTER(false, true, true);
        } else if (has_v0) {
            if (iota) SCATTER(true, true, false); else SCATTER(false, true, false);
        } else if (has_v1) {
            SCATTER(false, false, true);
        } else {
            SCATTER(false, false, false);
        }
#undef SCATTER
        iota = false;
        uint32_t *t;
        t = ki; ki = ko; ko = t;
        t = v0i; v0i = v0o; v0o = t;
        t = v1i; v1i = v1o; v1o = t;
        in_alt = !in_alt;
    }
    PPF_CUDA(ctx, cudaFreeAsync(hist, ctx->stream));
    PPF_CUDA(ctx, cudaFreeAsync(chunk_tot, ctx->stream));
    PPF_CUDA(ctx, cudaFreeAsync(digit_base, ctx->stream));
    *result_in_alt = in_alt;
    return B200PPF_OK;
}

}  // namespace

int radix_sort_u32(b200ppf_ctx *ctx, uint32_t *keys, uint32_t *keys_alt, uint32_t *v0, uint32_t *v0_alt,
                   uint32_t *v1, uint32_t *v1_alt, size_t n, int bits, bool v0_iota, bool *result_in_alt) {
    *result_in_alt = false;
    if (n == 0 || bits <= 0) return B200PPF_OK;
    if (n >= 0xFFFFFFFFull - WARPS * ROUNDS_BIG * 32)
        return fail_msg(ctx, B200PPF_ERR_UNSUPPORTED, "radix sort: more than 2^32 elements");
    // 4096-element tiles once every SM gets a few of them, 512-element tiles for small inputs
    if (n >= (size_t)ctx->sm_count * 4 * WARPS * ROUNDS_BIG * 32)
        return radix_sort_impl<ROUNDS_BIG>(ctx, keys, keys_alt, v0, v0_alt, v1, v1_alt, n, bits, v0_iota, result_in_alt);
    return radix_sort_impl<ROUNDS_SMALL>(ctx, keys, keys_alt, v0, v0_alt, v1, v1_alt, n, bits, v0_iota, result_in_alt);
}

}  // namespace b200ppf
