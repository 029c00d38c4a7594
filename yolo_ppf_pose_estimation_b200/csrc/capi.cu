// capi.cu — the extern "C" surface declared in include/b200ppf.h.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "ppf_common.cuh"

namespace b200ppf {

static std::mutex g_err_mutex;
static std::string g_error;

int fail_msg(b200ppf_ctx *ctx, int code, const char *msg) {
    if (ctx) ctx->error = msg;
    std::lock_guard<std::mutex> lock(g_err_mutex);
    g_error = msg;
    return code;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace b200ppf

using namespace b200ppf;

#define CHECK_CTX(ctx)                                   \
    if (!(ctx)) return fail_msg(nullptr, B200PPF_ERR_INVALID, "null context"); \
    DeviceGuard _guard((ctx)->device)

extern "C" {

int b200ppf_version(void) { return B200PPF_VERSION; }

int b200ppf_create(int device, b200ppf_ctx **out) {
    if (!out) return fail_msg(nullptr, B200PPF_ERR_INVALID, "b200ppf_create: null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail_msg(nullptr, B200PPF_ERR_CUDA,
                        "b200ppf_create: no CUDA device is visible; this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail_msg(nullptr, B200PPF_ERR_INVALID, "b200ppf_create: device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return fail_msg(nullptr, B200PPF_ERR_CUDA, "b200ppf_create: cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail_msg(nullptr, B200PPF_ERR_UNSUPPORTED,
                        "b200ppf_create: kernels are built for sm_100a (Blackwell B200) only");
    b200ppf_ctx *ctx = new (std::nothrow) b200ppf_ctx();
    if (!ctx) return fail_msg(nullptr, B200PPF_ERR_NOMEM, "b200ppf_create: out of host memory");
    ctx->device = device;
    DeviceGuard guard(device);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail_msg(nullptr, B200PPF_ERR_CUDA, "b200ppf_create: cudaStreamCreate failed");
    }
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    for (auto &ev : ctx->ev_vote) cudaEventCreate(&ev);
    {   // keep freed scratch and table arrays in the device's stream-ordered pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = ctx;
    return B200PPF_OK;
}

void b200ppf_destroy(b200ppf_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->d_peaks) cudaFree(ctx->d_peaks);
    if (ctx->d_hyps) cudaFree(ctx->d_hyps);
    if (ctx->d_assign) cudaFree(ctx->d_assign);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->ev_vote)
        if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *b200ppf_last_error(const b200ppf_ctx *ctx) {
    if (ctx) return ctx->error.c_str();
    std::lock_guard<std::mutex> lock(g_err_mutex);
    static thread_local std::string copy;
    copy = g_error;
    return copy.c_str();
}

int b200ppf_set_feature_mode(b200ppf_ctx *ctx, int mode) {
    if (!ctx) return fail_msg(nullptr, B200PPF_ERR_INVALID, "null context");
    if (mode < 0 || mode > 2) return fail_msg(ctx, B200PPF_ERR_INVALID, "unknown feature mode");
    ctx->feature_mode = mode;
    return B200PPF_OK;
}

int b200ppf_set_alpha_mode(b200ppf_ctx *ctx, int mode) {
    if (!ctx) return fail_msg(nullptr, B200PPF_ERR_INVALID, "null context");
    if (mode < 0 || mode > 1) return fail_msg(ctx, B200PPF_ERR_INVALID, "unknown alpha mode");
    ctx->alpha_mode = mode;
    return B200PPF_OK;
}

int b200ppf_set_nalpha_rule(b200ppf_ctx *ctx, int rule) {
    if (!ctx) return fail_msg(nullptr, B200PPF_ERR_INVALID, "null context");
    if (rule < 0 || rule > 2) return fail_msg(ctx, B200PPF_ERR_INVALID, "unknown alpha-column rule");
    ctx->nalpha_rule = rule;
    return B200PPF_OK;
}

int b200ppf_get_device(const b200ppf_ctx *ctx) { return ctx ? ctx->device : -1; }
void *b200ppf_get_stream(const b200ppf_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t b200ppf_launch_count(const b200ppf_ctx *ctx) { return ctx ? ctx->launches : 0; }

int b200ppf_synchronize(b200ppf_ctx *ctx) {
    CHECK_CTX(ctx);
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200PPF_OK;
}

int b200ppf_get_timings(b200ppf_ctx *ctx, b200ppf_timings *out) {
    if (!ctx || !out) return fail_msg(nullptr, B200PPF_ERR_INVALID, "null argument");
    if (ctx->vote_timed) {  // the vote is asynchronous: resolve its events on demand
        DeviceGuard guard(ctx->device);
        if (cudaEventSynchronize(ctx->ev_vote[3]) == cudaSuccess) {
            cudaEventElapsedTime(&ctx->timings.grid_ms, ctx->ev_vote[0], ctx->ev_vote[1]);
            cudaEventElapsedTime(&ctx->timings.vote_ms, ctx->ev_vote[1], ctx->ev_vote[2]);
            cudaEventElapsedTime(&ctx->timings.pose_ms, ctx->ev_vote[2], ctx->ev_vote[3]);
        }
    }
    *out = ctx->timings;
    return B200PPF_OK;
}

/* ---- clouds ------------------------------------------------------------------------------- */

int b200ppf_cloud_upload(b200ppf_ctx *ctx, const float *host, size_t n, size_t stride, size_t noff,
                         b200ppf_cloud **out) {
    CHECK_CTX(ctx);
    if (!out) return fail_msg(ctx, B200PPF_ERR_INVALID, "cloud upload: null output pointer");
    *out = nullptr;
    if (n && !host) return fail_msg(ctx, B200PPF_ERR_INVALID, "cloud upload: null host pointer");
    if (stride < 6 || noff < 3 || noff + 3 > stride)
        return fail_msg(ctx, B200PPF_ERR_INVALID, "cloud upload: stride/normal offset do not describe [x y z .. nx ny nz]");
    // AoS -> float4 SoA staging in the context's grow-only pinned buffer, dropping non-finite points
    const size_t need = std::max<size_t>(1, 2 * n) * sizeof(float4);
    if (ctx->stage_bytes < need) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr;
        ctx->stage_bytes = 0;
        PPF_CUDA(ctx, cudaMallocHost(&ctx->stage, need));
        ctx->stage_bytes = need;
    }
    float4 *stage = static_cast<float4 *>(ctx->stage);
    b200ppf_cloud *c = new (std::nothrow) b200ppf_cloud();
    if (!c) return fail_msg(ctx, B200PPF_ERR_NOMEM, "cloud upload: out of host memory");
    c->ctx = ctx;
    size_t m = 0;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float4 *sp = stage, *sn = stage + n;
    for (size_t i = 0; i < n; ++i) {
        const float *p = host + i * stride, *q = p + noff;
        // non-finite rows are dropped: NaN (SURVEY.md A.8 rule 5) and +-Inf alike (depth-derived clouds carry both;
        // an infinite coordinate would make the bounding box, and every grid sized from it, meaningless)
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2]) || !std::isfinite(q[0]) ||
            !std::isfinite(q[1]) || !std::isfinite(q[2]))
            continue;
        sp[m] = make_float4(p[0], p[1], p[2], 1.0f);
        sn[m] = make_float4(q[0], q[1], q[2], 0.0f);
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], p[k]);
            hi[k] = std::max(hi[k], p[k]);
        }
        ++m;
    }
    c->n = m;
    for (int k = 0; k < 3; ++k) {
        c->bbox_min[k] = m ? lo[k] : 0.0f;
        c->bbox_max[k] = m ? hi[k] : 0.0f;
    }
    // one stream-ordered allocation holds both arrays (pos | nrm)
    cudaError_t e = cudaMallocAsync(&c->pos, std::max<size_t>(1, 2 * m) * sizeof(float4), ctx->stream);
    if (e == cudaSuccess) c->nrm = c->pos + m;
    cudaEventRecord(ctx->ev[0], ctx->stream);
    if (e == cudaSuccess && m) e = cudaMemcpyAsync(c->pos, sp, m * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && m) e = cudaMemcpyAsync(c->nrm, sn, m * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream);
    cudaEventRecord(ctx->ev[1], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // the staging buffer is reused by the next upload
    if (e != cudaSuccess) {
        b200ppf_cloud_free(c);
        return fail_msg(ctx, e == cudaErrorMemoryAllocation ? B200PPF_ERR_NOMEM : B200PPF_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaEventElapsedTime(&ctx->timings.upload_ms, ctx->ev[0], ctx->ev[1]);
    *out = c;
    return B200PPF_OK;
}

size_t b200ppf_cloud_size(const b200ppf_cloud *cloud) { return cloud ? cloud->n : 0; }

void b200ppf_cloud_free(b200ppf_cloud *c) {
    if (!c) return;
    DeviceGuard guard(c->ctx ? c->ctx->device : 0);
    if (c->pos) {  // pos | nrm share one stream-ordered allocation
        if (c->ctx) cudaFreeAsync(c->pos, c->ctx->stream);
        else cudaFree(c->pos);
    }
    delete c;
}

/* ---- K1 ----------------------------------------------------------------------------------- */

int b200ppf_features_compute(b200ppf_ctx *ctx, const b200ppf_cloud *model, b200ppf_features **out) {
    CHECK_CTX(ctx);
    if (!model || !out) return fail_msg(ctx, B200PPF_ERR_INVALID, "features compute: null argument");
    *out = nullptr;
    if (model->n > 65535) return fail_msg(ctx, B200PPF_ERR_UNSUPPORTED, "features compute: more than 65535 model points");
    b200ppf_features *f = new (std::nothrow) b200ppf_features();
    if (!f) return fail_msg(ctx, B200PPF_ERR_NOMEM, "features compute: out of host memory");
    f->ctx = ctx;
    f->count = model->n * model->n;
    cudaError_t e = cudaMalloc(&f->d, std::max<size_t>(1, f->count) * sizeof(b200ppf_signature));
    if (e != cudaSuccess) {
        delete f;
        return fail_msg(ctx, B200PPF_ERR_NOMEM, "features compute: device allocation of the N*N*20-byte feature cloud failed");
    }
    cudaEventRecord(ctx->ev[0], ctx->stream);
    int rc = k1_features_compute(ctx, model, f->d);
    cudaEventRecord(ctx->ev[1], ctx->stream);
    if (rc == B200PPF_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        rc = fail_msg(ctx, B200PPF_ERR_CUDA, "features compute: kernel failed");
    if (rc != B200PPF_OK) {
        b200ppf_features_free(f);
        return rc;
    }
    cudaEventElapsedTime(&ctx->timings.features_ms, ctx->ev[0], ctx->ev[1]);
    *out = f;
    return B200PPF_OK;
}

int b200ppf_features_upload(b200ppf_ctx *ctx, const b200ppf_signature *host, size_t count, b200ppf_features **out) {
    CHECK_CTX(ctx);
    if (!out || (count && !host)) return fail_msg(ctx, B200PPF_ERR_INVALID, "features upload: null argument");
    *out = nullptr;
    b200ppf_features *f = new (std::nothrow) b200ppf_features();
    if (!f) return fail_msg(ctx, B200PPF_ERR_NOMEM, "features upload: out of host memory");
    f->ctx = ctx;
    f->count = count;
    cudaError_t e = cudaMalloc(&f->d, std::max<size_t>(1, count) * sizeof(b200ppf_signature));
    if (e == cudaSuccess && count)
        e = cudaMemcpyAsync(f->d, host, count * sizeof(b200ppf_signature), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        b200ppf_features_free(f);
        return fail_msg(ctx, e == cudaErrorMemoryAllocation ? B200PPF_ERR_NOMEM : B200PPF_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = f;
    return B200PPF_OK;
}

int b200ppf_features_download(b200ppf_ctx *ctx, const b200ppf_features *f, size_t first, size_t count,
                              b200ppf_signature *host) {
    CHECK_CTX(ctx);
    if (!f || (count && !host)) return fail_msg(ctx, B200PPF_ERR_INVALID, "features download: null argument");
    if (first > f->count || count > f->count - first)
        return fail_msg(ctx, B200PPF_ERR_INVALID, "features download: range exceeds the feature cloud");
    if (count) {
        PPF_CUDA(ctx, cudaMemcpyAsync(host, f->d + first, count * sizeof(b200ppf_signature), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return B200PPF_OK;
}

size_t b200ppf_features_count(const b200ppf_features *f) { return f ? f->count : 0; }

void b200ppf_features_free(b200ppf_features *f) {
    if (!f) return;
    DeviceGuard guard(f->ctx ? f->ctx->device : 0);
    if (f->d) cudaFree(f->d);
    delete f;
}

/* ---- K2 ----------------------------------------------------------------------------------- */

int b200ppf_table_build(b200ppf_ctx *ctx, const b200ppf_features *f, float angle_step, float dist_step,
                        b200ppf_table **out) {
    CHECK_CTX(ctx);
    if (!f || !out) return fail_msg(ctx, B200PPF_ERR_INVALID, "table build: null argument");
    *out = nullptr;
    return k2_build(ctx, f, nullptr, angle_step, dist_step, out);
}

int b200ppf_table_build_from_cloud(b200ppf_ctx *ctx, const b200ppf_cloud *model, float angle_step, float dist_step,
                                   b200ppf_table **out) {
    CHECK_CTX(ctx);
    if (!model || !out) return fail_msg(ctx, B200PPF_ERR_INVALID, "table build: null argument");
    *out = nullptr;
    return k2_build(ctx, nullptr, model, angle_step, dist_step, out);
}

int b200ppf_table_get_info(const b200ppf_table *t, b200ppf_table_info *info) {
    if (!t || !info) return fail_msg(nullptr, B200PPF_ERR_INVALID, "table info: null argument");
    *info = t->info;
    return B200PPF_OK;
}

int b200ppf_table_query_key(b200ppf_ctx *ctx, const b200ppf_table *t, const int32_t *d4, uint64_t *pairs, size_t cap,
                            size_t *n_found) {
    CHECK_CTX(ctx);
    if (!t || !d4 || !n_found || (cap && !pairs)) return fail_msg(ctx, B200PPF_ERR_INVALID, "table query: null argument");
    return k2_query_key(ctx, t, d4, pairs, cap, n_found);
}

int b200ppf_table_query(b200ppf_ctx *ctx, const b200ppf_table *t, float f1, float f2, float f3, float f4,
                        uint64_t *pairs, size_t cap, size_t *n_found) {
    CHECK_CTX(ctx);
    if (!t || !n_found || (cap && !pairs)) return fail_msg(ctx, B200PPF_ERR_INVALID, "table query: null argument");
    *n_found = 0;
    const float f[4] = {f1, f2, f3, f4};
    if (std::isnan(f1) || std::isnan(f2) || std::isnan(f3) || std::isnan(f4)) return B200PPF_OK;
    int d[4];
    quantise(t->kp, f, d);  // same IEEE divide + floor as the device
    return k2_query_key(ctx, t, d, pairs, cap, n_found);
}

int b200ppf_table_alpha_m(b200ppf_ctx *ctx, const b200ppf_table *t, float *host) {
    CHECK_CTX(ctx);
    if (!t || !host) return fail_msg(ctx, B200PPF_ERR_INVALID, "alpha_m export: null argument");
    return k2_alpha_m(ctx, t, host);
}

int b200ppf_table_export(b200ppf_ctx *ctx, const b200ppf_table *t, uint32_t *offsets, uint32_t *entry_i,
                         uint32_t *entry_j, float *entry_alpha_m) {
    CHECK_CTX(ctx);
    if (!t) return fail_msg(ctx, B200PPF_ERR_INVALID, "table export: null table");
    const size_t total = (size_t)t->info.key_space * t->info.n_slices + 1, ne = t->info.n_entries;
    if (offsets) PPF_CUDA(ctx, cudaMemcpyAsync(offsets, t->offsets, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint32_t> idx;
    std::vector<float> ent;
    if ((entry_i || entry_j) && ne) {
        idx.resize(ne);
        PPF_CUDA(ctx, cudaMemcpyAsync(idx.data(), t->entry_idx, ne * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (entry_alpha_m && ne) {
        ent.resize(ne);
        PPF_CUDA(ctx, cudaMemcpyAsync(ent.data(), t->entry_alpha, ne * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t n = (uint32_t)t->info.n_model;
    if (t->info.phase_cells > 1 && ne) {
        // buckets are stored in (phase cell, i, j) order: report them in the canonical (i, j) order
        std::vector<uint32_t> off(total), all_idx(ne);
        PPF_CUDA(ctx, cudaMemcpyAsync(off.data(), t->offsets, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        PPF_CUDA(ctx, cudaMemcpyAsync(all_idx.data(), t->entry_idx, ne * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::vector<std::pair<uint32_t, uint32_t>> order;
        for (size_t k = 0; k + 1 < total; ++k) {
            const uint32_t b = off[k], e = off[k + 1];
            if (e - b < 2) continue;
            order.resize(e - b);
            for (uint32_t p = b; p < e; ++p) order[p - b] = std::make_pair(all_idx[p], p);
            std::sort(order.begin(), order.end());
            for (uint32_t p = b; p < e; ++p) {
                if (!idx.empty()) idx[p] = order[p - b].first;
                if (entry_alpha_m) entry_alpha_m[p] = ent[order[p - b].second];
            }
            if (entry_alpha_m)
                for (uint32_t p = b; p < e; ++p) ent[p] = entry_alpha_m[p];
        }
    }
    for (size_t e = 0; e < idx.size(); ++e) {
        if (entry_i) entry_i[e] = idx[e] / n;
        if (entry_j) entry_j[e] = idx[e] % n;
    }
    for (size_t e = 0; e < ent.size(); ++e) entry_alpha_m[e] = ent[e];
    return B200PPF_OK;
}

/* ---- table file ------------------------------------------------------------------------------- */
namespace {

constexpr uint64_t TABLE_MAGIC = 0x4C42543030325042ull;  // "BP200TBL" little-endian
constexpr uint32_t TABLE_FORMAT = 4;                      // 4: sub-phase byte above the hot word (3: alpha-column rule, bin-major hot words)

struct TableFileHeader {
    uint64_t magic;
    uint32_t format, header_bytes;
    uint32_t feature_mode, alpha_mode;
    uint64_t n_offsets, n_sub_offsets, n_entries;  // array lengths in 32-bit words (entries: without the padding)
    uint64_t n_merged;                              // merged-vote words (its cell bounds have n_sub_offsets words)
    uint64_t checksum;                              // over info, kp, bp and the six arrays, in file order
    b200ppf_table_info info;
    b200ppf::KeyParams kp;
    b200ppf::BinParams bp;
};

// 64-bit multiplicative checksum over 8-byte words (tail bytes zero-extended)
uint64_t mix_bytes(uint64_t h, const void *data, size_t bytes) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    size_t k = 0;
    for (; k + 8 <= bytes; k += 8) {
        uint64_t w;
        memcpy(&w, p + k, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    }
    if (k < bytes) {
        uint64_t w = 0;
        memcpy(&w, p + k, bytes - k);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    }
    return h;
}

struct FileCloser {
    FILE *f;
    ~FileCloser() {
        if (f) fclose(f);
    }
};

// Everything the voting kernel takes on trust from a table, re-derived from the file's own parameters and entries:
// a file that passes cannot make the kernel read outside its arrays or add outside its accumulator slice, whoever
// wrote it (the checksum only catches accidents).  Returns nullptr or what is wrong.
const char *validate_table_file(const TableFileHeader &h, const std::vector<std::vector<uint32_t>> &a) {
    const b200ppf_table_info &info = h.info;
    const b200ppf::KeyParams &kp = h.kp;
    if (info.n_model < 1) return "no model points";
    if (!(info.angle_step > 0.0f) || !(info.dist_step > 0.0f) || kp.angle_step != info.angle_step || kp.dist_step != info.dist_step)
        return "discretisation steps";
    if (h.feature_mode > 2 || h.alpha_mode > 1 || info.nalpha_rule > 2) return "unknown mode";
    unsigned __int128 ks = 1;
    for (int k = 0; k < 4; ++k) {
        if (kp.size[k] < 1 || kp.size[k] != info.size[k] || kp.lo[k] != info.lo[k]) return "key ranges";
        ks *= (unsigned __int128)(uint32_t)kp.size[k];
    }
    if (ks != (unsigned __int128)info.key_space || kp.key_space != info.key_space) return "key space is not the product of the key ranges";
    if (info.n_slices < 1 || info.slice_rows < 1 || (uint64_t)info.n_slices * info.slice_rows < info.n_model ||
        (uint64_t)(info.n_slices - 1) * info.slice_rows >= info.n_model)
        return "accumulator slices do not tile the model rows";
    // the binning parameters are a function of (angle step, alpha mode, column rule); only the number of phase
    // cells may be lower than that function's choice (large key spaces halve it)
    b200ppf::BinParams ref = b200ppf::make_bin_params(info.angle_step, (int)h.alpha_mode, (int)info.nalpha_rule);
    if (h.bp.cells_log2 > ref.cells_log2) return "phase cells";
    ref.cells_log2 = h.bp.cells_log2;
    if (ref.cells_log2 == 0) ref.bulk = 0;
    if (memcmp(&ref, &h.bp, sizeof(ref)) != 0) return "binning parameters do not follow from the angle step";
    if (info.n_alpha != ref.n_alpha || info.phase_cells != (1u << ref.cells_log2)) return "alpha columns";
    const std::vector<uint32_t> &off = a[0], &sub = a[1], &ew = a[2], &eam = a[3], &ealpha = a[4], &eidx = a[5];
    const uint64_t ne = h.n_entries, total_keys = (uint64_t)info.key_space * info.n_slices;
    if (off.empty() || off[0] != 0 || off.back() != ne) return "bucket offsets do not cover the entries";
    for (size_t k = 0; k + 1 < off.size(); ++k)
        if (off[k] > off[k + 1]) return "bucket offsets decrease";
    if (!sub.empty()) {
        if (sub.back() != ne) return "phase-cell offsets do not cover the entries";
        for (size_t k = 0; k + 1 < sub.size(); ++k)
            if (sub[k] > sub[k + 1]) return "phase-cell offsets decrease";
        for (uint64_t k = 0; k <= total_keys; ++k)
            if (sub[k << ref.cells_log2] != off[k]) return "phase-cell offsets leave their bucket";
    }
    const uint64_t n = info.n_model, nn = n * n;
    if (kp.slice_rows != info.slice_rows || kp.n_slices != info.n_slices || kp.row_pitch != (info.slice_rows + 31u) / 32u * 32u)
        return "slice geometry";
    for (uint32_t slice = 0; slice < info.n_slices; ++slice) {
        const uint64_t p0 = off[(uint64_t)slice * info.key_space], p1 = off[(uint64_t)(slice + 1) * info.key_space];
        const uint64_t row0 = (uint64_t)slice * info.slice_rows, rows = std::min<uint64_t>(info.slice_rows, n - row0);
        for (uint64_t p = p0; p < p1; ++p) {
            if (eidx[p] >= nn) return "entry pair index out of range";
            const uint64_t i = eidx[p] / n;
            if (i < row0 || i >= row0 + rows) return "entry filed under another accumulator slice";
            float alpha;
            memcpy(&alpha, &ealpha[p], sizeof(float));
            if (!(alpha >= -3.14159274f && alpha <= 3.14159274f)) return "entry alpha_m outside [-pi, pi]";
            const uint32_t a_fix = b200ppf::alpha_to_fix(alpha), local = (uint32_t)(i - row0);
            if (eam[p] != a_fix) return "entry fixed-point alpha_m does not match its float";
            if (ew[p] != (ref.bulk ? b200ppf::entry_word(ref, kp.row_pitch, local, alpha) : 4u * local))
                return "entry hot word does not match its pair";
        }
    }
    // merged votes: cell bounds monotone inside the array; every word a 4-aligned offset inside the accumulator slice
    // with a count >= 1; the counts of a cell add up to the cell's entries (the votes cast are those of the entries)
    const std::vector<uint32_t> &mw = a[6], &msub = a[7];
    if (!mw.empty() || !msub.empty()) {
        if (msub.size() != sub.size() || msub.empty() || msub[0] != 0 || msub.back() != mw.size()) return "merged cell offsets";
        const uint32_t acc_bytes = ref.acc_cols * kp.row_pitch * 4u;
        for (size_t c = 0; c + 1 < msub.size(); ++c) {
            if (msub[c] > msub[c + 1]) return "merged cell offsets decrease";
            uint64_t votes = 0;
            for (uint64_t p = msub[c]; p < msub[c + 1]; ++p) {
                const uint32_t off = mw[p] & 0xFFFFFFu, cnt = mw[p] >> 24;
                if (cnt == 0 || (off & 3u) || off >= acc_bytes) return "merged vote word outside the accumulator slice";
                votes += cnt;
            }
            if (votes != (uint64_t)sub[c + 1] - sub[c]) return "merged vote counts do not add up to the cell's entries";
        }
    }
    return nullptr;
}

}  // namespace

int b200ppf_table_save(b200ppf_ctx *ctx, const b200ppf_table *t, const char *path) {
    CHECK_CTX(ctx);
    if (!t || !path) return fail_msg(ctx, B200PPF_ERR_INVALID, "table save: null argument");
    DeviceGuard guard(ctx->device);
    TableFileHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = TABLE_MAGIC;
    h.format = TABLE_FORMAT;
    h.header_bytes = (uint32_t)sizeof(h);
    h.feature_mode = (uint32_t)t->feature_mode;
    h.alpha_mode = (uint32_t)t->bp.mode;
    h.info = t->info;
    h.kp = t->kp;
    h.bp = t->bp;
    const uint64_t total_keys = (uint64_t)t->info.key_space * t->info.n_slices;
    h.n_offsets = total_keys + 1;
    h.n_sub_offsets = t->sub_offsets ? (total_keys << t->bp.cells_log2) + 1 : 0;
    h.n_entries = t->info.n_entries;
    h.n_merged = t->merged_w ? t->n_merged : 0;
    constexpr int NA = 8;
    const uint32_t *arrays[NA] = {t->offsets, t->sub_offsets, t->entry_w, t->entry_am,
                                  reinterpret_cast<const uint32_t *>(t->entry_alpha), t->entry_idx, t->merged_w, t->msub_offsets};
    const uint64_t lengths[NA] = {h.n_offsets, h.n_sub_offsets, h.n_entries, h.n_entries, h.n_entries, h.n_entries, h.n_merged,
                                  t->msub_offsets ? h.n_sub_offsets : 0};
    std::vector<std::vector<uint32_t>> host(NA);
    for (int a = 0; a < NA; ++a) {
        host[a].resize(lengths[a]);
        if (lengths[a])
            PPF_CUDA(ctx, cudaMemcpyAsync(host[a].data(), arrays[a], lengths[a] * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                          ctx->stream));
    }
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // the parameter block (info, kp, bp: everything behind the checksum field) and then the six arrays
    uint64_t sum = mix_bytes(0x42323030u, reinterpret_cast<const char *>(&h) + offsetof(TableFileHeader, info),
                             sizeof(h) - offsetof(TableFileHeader, info));
    for (int a = 0; a < NA; ++a) sum = mix_bytes(sum, host[a].data(), host[a].size() * sizeof(uint32_t));
    h.checksum = sum;
    FileCloser fc{fopen(path, "wb")};
    if (!fc.f) return fail_msg(ctx, B200PPF_ERR_IO, "table save: cannot open the file for writing");
    bool ok = fwrite(&h, sizeof(h), 1, fc.f) == 1;
    for (int a = 0; a < NA && ok; ++a)
        if (!host[a].empty()) ok = fwrite(host[a].data(), sizeof(uint32_t), host[a].size(), fc.f) == host[a].size();
    ok = ok && fflush(fc.f) == 0;
    if (!ok) return fail_msg(ctx, B200PPF_ERR_IO, "table save: short write");
    return B200PPF_OK;
}

int b200ppf_table_load(b200ppf_ctx *ctx, const char *path, b200ppf_table **out) {
    CHECK_CTX(ctx);
    if (!path || !out) return fail_msg(ctx, B200PPF_ERR_INVALID, "table load: null argument");
    *out = nullptr;
    DeviceGuard guard(ctx->device);
    FileCloser fc{fopen(path, "rb")};
    if (!fc.f) return fail_msg(ctx, B200PPF_ERR_IO, "table load: cannot open the file");
    TableFileHeader h;
    if (fread(&h, sizeof(h), 1, fc.f) != 1) return fail_msg(ctx, B200PPF_ERR_IO, "table load: truncated header");
    if (h.magic != TABLE_MAGIC) return fail_msg(ctx, B200PPF_ERR_IO, "table load: not a b200ppf table file");
    if (h.format != TABLE_FORMAT || h.header_bytes != sizeof(h))
        return fail_msg(ctx, B200PPF_ERR_IO, "table load: file written by another format version");
    if ((int)h.feature_mode != ctx->feature_mode || (int)h.alpha_mode != ctx->alpha_mode)
        return fail_msg(ctx, B200PPF_ERR_STATE, "table load: the file's feature / alpha mode differs from the context's");
    const uint64_t total_keys = (uint64_t)h.info.key_space * h.info.n_slices;
    const bool sane = h.n_offsets == total_keys + 1 && total_keys < (1ull << 31) && h.bp.cells_log2 <= 8 &&
                      (total_keys << h.bp.cells_log2) < (1ull << 31) &&
                      h.n_sub_offsets == (h.bp.cells_log2 ? (total_keys << h.bp.cells_log2) + 1 : 0) &&
                      h.n_entries == h.info.n_entries && h.n_entries <= 0xFFFFFFFFull && h.info.n_model >= 1 &&
                      h.info.n_model <= 65535 && h.n_entries <= h.info.n_model * h.info.n_model &&
                      h.info.n_alpha >= 1 && h.bp.n_alpha == h.info.n_alpha &&
                      h.kp.slice_rows == h.info.slice_rows && h.kp.n_slices == h.info.n_slices;
    if (!sane) return fail_msg(ctx, B200PPF_ERR_IO, "table load: inconsistent header");
    constexpr int NA = 8;
    const bool has_merged = h.bp.cells_log2 != 0 && h.n_entries != 0;
    if (h.n_merged > h.n_entries || (has_merged ? h.n_merged == 0 : h.n_merged != 0) || h.info.n_merged != h.n_merged)
        return fail_msg(ctx, B200PPF_ERR_IO, "table load: inconsistent header");
    const uint64_t lengths[NA] = {h.n_offsets, h.n_sub_offsets, h.n_entries, h.n_entries, h.n_entries, h.n_entries, h.n_merged,
                                  has_merged ? h.n_sub_offsets : 0};
    std::vector<std::vector<uint32_t>> host(NA);
    // the parameter block (info, kp, bp: everything behind the checksum field) and then the six arrays
    uint64_t sum = mix_bytes(0x42323030u, reinterpret_cast<const char *>(&h) + offsetof(TableFileHeader, info),
                             sizeof(h) - offsetof(TableFileHeader, info));
    for (int a = 0; a < NA; ++a) {
        host[a].resize(lengths[a]);
        if (lengths[a] && fread(host[a].data(), sizeof(uint32_t), lengths[a], fc.f) != lengths[a])
            return fail_msg(ctx, B200PPF_ERR_IO, "table load: truncated file");
        sum = mix_bytes(sum, host[a].data(), host[a].size() * sizeof(uint32_t));
    }
    if (fgetc(fc.f) != EOF) return fail_msg(ctx, B200PPF_ERR_IO, "table load: trailing bytes");
    if (sum != h.checksum) return fail_msg(ctx, B200PPF_ERR_IO, "table load: checksum mismatch (corrupt file)");
    if (const char *why = validate_table_file(h, host)) {
        char msg[200];
        snprintf(msg, sizeof(msg), "table load: invalid table (%s)", why);
        return fail_msg(ctx, B200PPF_ERR_IO, msg);
    }

    b200ppf_table *t = new (std::nothrow) b200ppf_table();
    if (!t) return fail_msg(ctx, B200PPF_ERR_NOMEM, "table load: out of host memory");
    t->ctx = ctx;
    t->info = h.info;
    t->kp = h.kp;
    t->bp = h.bp;
    t->feature_mode = (int)h.feature_mode;
    t->n_merged = h.n_merged;
    uint32_t **dev[NA] = {&t->offsets, &t->sub_offsets, &t->entry_w, &t->entry_am,
                          reinterpret_cast<uint32_t **>(&t->entry_alpha), &t->entry_idx, &t->merged_w, &t->msub_offsets};
    for (int a = 0; a < NA; ++a) {
        if (a == 1 && !h.n_sub_offsets) continue;
        if (a >= 6 && !has_merged) continue;
        const size_t pad = (a == 2 || a == 3 || a == 6) ? b200ppf::ENTRY_PAD : 0;
        const size_t words = std::max<size_t>(1, lengths[a] + pad);
        cudaError_t e = cudaMallocAsync(dev[a], words * sizeof(uint32_t), ctx->stream);
        if (e == cudaSuccess && pad) e = cudaMemsetAsync(*dev[a] + lengths[a], 0, pad * sizeof(uint32_t), ctx->stream);
        if (e == cudaSuccess && lengths[a])
            e = cudaMemcpyAsync(*dev[a], host[a].data(), lengths[a] * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            b200ppf_table_free(t);
            return fail_msg(ctx, e == cudaErrorMemoryAllocation ? B200PPF_ERR_NOMEM : B200PPF_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        b200ppf_table_free(t);
        return fail_msg(ctx, B200PPF_ERR_CUDA, "table load: upload failed");
    }
    *out = t;
    return B200PPF_OK;
}

int b200ppf_table_clone(b200ppf_ctx *dst, const b200ppf_table *src, b200ppf_table **out) {
    CHECK_CTX(dst);
    if (!src || !out) return fail_msg(dst, B200PPF_ERR_INVALID, "table clone: null argument");
    *out = nullptr;
    if (src->ctx) cudaStreamSynchronize(src->ctx->stream);  // the source table's build has finished (device guard is dst's)
    b200ppf_table *t = new (std::nothrow) b200ppf_table();
    if (!t) return fail_msg(dst, B200PPF_ERR_NOMEM, "table clone: out of host memory");
    t->ctx = dst;
    t->info = src->info;
    t->kp = src->kp;
    t->bp = src->bp;
    t->feature_mode = src->feature_mode;
    t->n_merged = src->n_merged;
    const uint64_t total_keys = (uint64_t)src->info.key_space * src->info.n_slices;
    const size_t n_sub = src->sub_offsets ? (size_t)(total_keys << src->bp.cells_log2) + 1 : 0, ne = src->info.n_entries;
    const int src_dev = src->ctx ? src->ctx->device : dst->device;
    struct Arr {
        uint32_t *const *from;
        uint32_t **to;
        size_t words, pad;
    } arrs[8] = {{&src->offsets, &t->offsets, (size_t)total_keys + 1, 0},
                 {&src->sub_offsets, &t->sub_offsets, n_sub, 0},
                 {&src->entry_w, &t->entry_w, ne, b200ppf::ENTRY_PAD},
                 {&src->entry_am, &t->entry_am, ne, b200ppf::ENTRY_PAD},
                 {reinterpret_cast<uint32_t *const *>(&src->entry_alpha), reinterpret_cast<uint32_t **>(&t->entry_alpha), std::max<size_t>(1, ne), 0},
                 {&src->entry_idx, &t->entry_idx, std::max<size_t>(1, ne), 0},
                 {&src->merged_w, &t->merged_w, src->merged_w ? src->n_merged : 0, b200ppf::ENTRY_PAD},
                 {&src->msub_offsets, &t->msub_offsets, src->msub_offsets ? n_sub : 0, 0}};
    for (const Arr &a : arrs) {
        if (!*a.from) continue;
        cudaError_t e = cudaMallocAsync(a.to, (a.words + a.pad) * sizeof(uint32_t), dst->stream);
        if (e == cudaSuccess && a.pad) e = cudaMemsetAsync(*a.to + a.words, 0, a.pad * sizeof(uint32_t), dst->stream);
        if (e == cudaSuccess && a.words)
            e = cudaMemcpyPeerAsync(*a.to, dst->device, *a.from, src_dev, a.words * sizeof(uint32_t), dst->stream);
        if (e != cudaSuccess) {
            b200ppf_table_free(t);
            return fail_msg(dst, e == cudaErrorMemoryAllocation ? B200PPF_ERR_NOMEM : B200PPF_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    if (cudaStreamSynchronize(dst->stream) != cudaSuccess) {
        b200ppf_table_free(t);
        return fail_msg(dst, B200PPF_ERR_CUDA, "table clone: copy failed");
    }
    *out = t;
    return B200PPF_OK;
}

void b200ppf_table_free(b200ppf_table *t) {
    if (!t) return;
    DeviceGuard guard(t->ctx ? t->ctx->device : 0);
    // the arrays come from the device's stream-ordered pool (k2_build, load, clone): they go back to it, so that the
    // next build re-uses the memory instead of asking the driver for gigabytes again
    void *arrays[8] = {t->offsets, t->sub_offsets, t->entry_w, t->entry_am, t->entry_idx, t->entry_alpha, t->merged_w, t->msub_offsets};
    for (void *p : arrays) {
        if (!p) continue;
        if (t->ctx) cudaFreeAsync(p, t->ctx->stream);
        else cudaFree(p);
    }
    delete t;
}

/* ---- K3 ----------------------------------------------------------------------------------- */

int b200ppf_vote_device(b200ppf_ctx *ctx, const b200ppf_cloud *model, const b200ppf_table *t,
                        const b200ppf_cloud *scene, size_t ref_first, size_t ref_step, size_t ref_count,
                        b200ppf_hypothesis *hyps_device) {
    CHECK_CTX(ctx);
    if (ref_count && !hyps_device) return fail_msg(ctx, B200PPF_ERR_INVALID, "vote: null output pointer");
    return k3_vote(ctx, model, t, scene, ref_first, ref_step, ref_count, &hyps_device, 1, 0, 1);
}

/* ---- multi-GPU exchange fused into the vote epilogue -------------------------------------------- */
int b200ppf_vote_scatter_device(b200ppf_ctx *ctx, const b200ppf_cloud *model, const b200ppf_table *t,
                                const b200ppf_cloud *scene, size_t ref_first, size_t ref_step, size_t ref_count,
                                b200ppf_hypothesis *const *peer_buffers, int n_peers, size_t slot_first,
                                size_t slot_step) {
    CHECK_CTX(ctx);
    if (!peer_buffers || n_peers < 1) return fail_msg(ctx, B200PPF_ERR_INVALID, "vote scatter: no output buffers");
    for (int g = 0; g < n_peers; ++g)
        if (!peer_buffers[g]) return fail_msg(ctx, B200PPF_ERR_INVALID, "vote scatter: null peer buffer");
    if (slot_step == 0) return fail_msg(ctx, B200PPF_ERR_INVALID, "vote scatter: slot step must be >= 1");
    return k3_vote(ctx, model, t, scene, ref_first, ref_step, ref_count, peer_buffers, n_peers, slot_first, slot_step);
}

int b200ppf_hyp_buffer_create(b200ppf_ctx *ctx, size_t n_records, b200ppf_hypothesis **buffer,
                              unsigned char ipc_handle[64]) {
    CHECK_CTX(ctx);
    if (!buffer) return fail_msg(ctx, B200PPF_ERR_INVALID, "hypothesis buffer: null argument");
    *buffer = nullptr;
    DeviceGuard guard(ctx->device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle travels as 64 bytes");
    b200ppf_hypothesis *p = nullptr;
    PPF_CUDA(ctx, cudaMalloc(&p, std::max<size_t>(1, n_records) * sizeof(b200ppf_hypothesis)));  // cudaMalloc: IPC-exportable
    PPF_CUDA(ctx, cudaMemset(p, 0, std::max<size_t>(1, n_records) * sizeof(b200ppf_hypothesis)));
    if (ipc_handle) {
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) {
            cudaFree(p);
            return fail_msg(ctx, B200PPF_ERR_CUDA, cudaGetErrorString(e));
        }
        memcpy(ipc_handle, &h, 64);
    }
    *buffer = p;
    return B200PPF_OK;
}

int b200ppf_hyp_buffer_open(b200ppf_ctx *ctx, const unsigned char ipc_handle[64], b200ppf_hypothesis **buffer) {
    CHECK_CTX(ctx);
    if (!ipc_handle || !buffer) return fail_msg(ctx, B200PPF_ERR_INVALID, "hypothesis buffer: null argument");
    *buffer = nullptr;
    DeviceGuard guard(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    void *p = nullptr;
    PPF_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *buffer = static_cast<b200ppf_hypothesis *>(p);
    return B200PPF_OK;
}

int b200ppf_hyp_buffer_download(b200ppf_ctx *ctx, const b200ppf_hypothesis *buffer, size_t first, size_t count,
                                b200ppf_hypothesis *host) {
    CHECK_CTX(ctx);
    if (!buffer || (count && !host)) return fail_msg(ctx, B200PPF_ERR_INVALID, "hypothesis buffer: null argument");
    DeviceGuard guard(ctx->device);
    if (count)
        PPF_CUDA(ctx, cudaMemcpyAsync(host, buffer + first, count * sizeof(b200ppf_hypothesis), cudaMemcpyDeviceToHost,
                                      ctx->stream));
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200PPF_OK;
}

int b200ppf_hyp_buffer_release(b200ppf_ctx *ctx, b200ppf_hypothesis *buffer, int opened_from_handle) {
    CHECK_CTX(ctx);
    if (!buffer) return B200PPF_OK;
    DeviceGuard guard(ctx->device);
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (opened_from_handle) PPF_CUDA(ctx, cudaIpcCloseMemHandle(buffer));
    else PPF_CUDA(ctx, cudaFree(buffer));
    return B200PPF_OK;
}

int b200ppf_vote(b200ppf_ctx *ctx, const b200ppf_cloud *model, const b200ppf_table *t, const b200ppf_cloud *scene,
                 size_t ref_first, size_t ref_step, size_t ref_count, b200ppf_hypothesis *hyps_host) {
    CHECK_CTX(ctx);
    if (ref_count && !hyps_host) return fail_msg(ctx, B200PPF_ERR_INVALID, "vote: null output pointer");
    if (ctx->hyps_cap < ref_count) {
        if (ctx->d_hyps) cudaFree(ctx->d_hyps);
        ctx->d_hyps = nullptr;
        ctx->hyps_cap = 0;
        PPF_CUDA(ctx, cudaMalloc(&ctx->d_hyps, ref_count * sizeof(b200ppf_hypothesis)));
        ctx->hyps_cap = ref_count;
    }
    b200ppf_hypothesis *const target = ctx->d_hyps;
    int rc = k3_vote(ctx, model, t, scene, ref_first, ref_step, ref_count, &target, 1, 0, 1);
    if (rc) return rc;
    if (ref_count)
        PPF_CUDA(ctx, cudaMemcpyAsync(hyps_host, ctx->d_hyps, ref_count * sizeof(b200ppf_hypothesis),
                                      cudaMemcpyDeviceToHost, ctx->stream));
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200PPF_OK;
}

int b200ppf_vote_stats(b200ppf_ctx *ctx, uint64_t *stats4) {
    CHECK_CTX(ctx);
    if (!stats4) return fail_msg(ctx, B200PPF_ERR_INVALID, "vote stats: null output pointer");
    stats4[0] = stats4[1] = stats4[2] = stats4[3] = 0;
    if (!ctx->d_stats) return B200PPF_OK;
    unsigned long long h[4];
    PPF_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 4; ++k) stats4[k] = h[k];
    return B200PPF_OK;
}

int b200ppf_vote_debug_pairs(b200ppf_ctx *ctx, const b200ppf_table *t, const b200ppf_cloud *scene, size_t s_r,
                             uint8_t *in_radius, int32_t *d4, float *alpha_s) {
    CHECK_CTX(ctx);
    if (!in_radius || !d4 || !alpha_s) return fail_msg(ctx, B200PPF_ERR_INVALID, "debug pairs: null output pointer");
    return k3_debug_pairs(ctx, t, scene, s_r, in_radius, d4, alpha_s);
}

int b200ppf_vote_debug_accumulator(b200ppf_ctx *ctx, const b200ppf_table *t, const b200ppf_cloud *scene, size_t s_r,
                                   uint32_t *acc) {
    CHECK_CTX(ctx);
    if (!acc) return fail_msg(ctx, B200PPF_ERR_INVALID, "debug accumulator: null output pointer");
    return k3_debug_accumulator(ctx, t, scene, s_r, acc);
}

int b200ppf_debug_alpha_bins(b200ppf_ctx *ctx, float angle_step, int alpha_mode, int nalpha_rule, const float *alpha_m,
                             const float *alpha_s, size_t n, uint32_t *fast, uint32_t *exact) {
    if (!(angle_step > 0.0f) || alpha_mode < 0 || alpha_mode > 1 || nalpha_rule < 0 || nalpha_rule > 2 ||
        (n && (!alpha_m || !alpha_s || !fast || !exact)))
        return fail_msg(ctx, B200PPF_ERR_INVALID, "debug alpha bins: bad argument");
    if (n == 0) return B200PPF_OK;
    if (!ctx) return k3_debug_alpha_bins(nullptr, angle_step, alpha_mode, nalpha_rule, alpha_m, alpha_s, n, fast, exact);
    DeviceGuard guard(ctx->device);
    return k3_debug_alpha_bins(ctx, angle_step, alpha_mode, nalpha_rule, alpha_m, alpha_s, n, fast, exact);
}

int b200ppf_microbench_atoms(b200ppf_ctx *ctx, int pattern, double *atoms_per_sec) {
    CHECK_CTX(ctx);
    if (!atoms_per_sec || pattern < 0 || pattern > 15) return fail_msg(ctx, B200PPF_ERR_INVALID, "microbench: bad argument");
    return microbench_atoms(ctx, pattern, atoms_per_sec);
}

/* ---- K4 ----------------------------------------------------------------------------------- */

int b200ppf_cluster_device(b200ppf_ctx *ctx, const b200ppf_hypothesis *hyps_device, size_t n, float pos_thr,
                           float rot_thr, float *poses16, uint32_t *votes, size_t *n_out) {
    CHECK_CTX(ctx);
    if (!poses16 || !votes || !n_out || (n && !hyps_device)) return fail_msg(ctx, B200PPF_ERR_INVALID, "cluster: null argument");
    return k4_cluster(ctx, hyps_device, n, pos_thr, rot_thr, poses16, votes, n_out);
}

int b200ppf_cluster(b200ppf_ctx *ctx, const b200ppf_hypothesis *hyps_host, size_t n, float pos_thr, float rot_thr,
                    float *poses16, uint32_t *votes, size_t *n_out) {
    CHECK_CTX(ctx);
    if (!poses16 || !votes || !n_out || (n && !hyps_host)) return fail_msg(ctx, B200PPF_ERR_INVALID, "cluster: null argument");
    *n_out = 0;
    if (n == 0) return B200PPF_OK;
    StreamBuf<b200ppf_hypothesis> d(ctx);
    PPF_CUDA(ctx, d.alloc(n));
    PPF_CUDA(ctx, cudaMemcpyAsync(d, hyps_host, n * sizeof(b200ppf_hypothesis), cudaMemcpyHostToDevice, ctx->stream));
    return k4_cluster(ctx, d, n, pos_thr, rot_thr, poses16, votes, n_out);
}

int b200ppf_cluster_assignment(b200ppf_ctx *ctx, uint32_t *assignment, size_t n, size_t *n_clusters) {
    CHECK_CTX(ctx);
    if (n_clusters) *n_clusters = ctx->n_clusters;
    if (!assignment) return B200PPF_OK;
    if (n != ctx->assign_n || !ctx->d_assign)
        return fail_msg(ctx, B200PPF_ERR_STATE, "cluster assignment: size does not match the last cluster call");
    PPF_CUDA(ctx, cudaMemcpyAsync(assignment, ctx->d_assign, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200PPF_OK;
}

/* ---- K5 ----------------------------------------------------------------------------------- */

int b200ppf_transform(b200ppf_ctx *ctx, const b200ppf_cloud *cloud, const float *pose16, float *out_host,
                      size_t out_stride_floats) {
    CHECK_CTX(ctx);
    if (!cloud || !pose16 || (cloud->n && !out_host)) return fail_msg(ctx, B200PPF_ERR_INVALID, "transform: null argument");
    return k5_transform(ctx, cloud, pose16, out_host, out_stride_floats);
}

/* ---- scene pre-processing (prep.cu) ------------------------------------------------------ */

int b200ppf_cloud_upload_xyz(b200ppf_ctx *ctx, const float *host, size_t n, size_t stride, b200ppf_cloud **out) {
    CHECK_CTX(ctx);
    if (!out) return fail_msg(ctx, B200PPF_ERR_INVALID, "cloud upload: null output pointer");
    *out = nullptr;
    if (n && !host) return fail_msg(ctx, B200PPF_ERR_INVALID, "cloud upload: null host pointer");
    if (stride < 3) return fail_msg(ctx, B200PPF_ERR_INVALID, "cloud upload: rows must hold at least x y z");
    if (n > 0xFFFFFFF0ull) return fail_msg(ctx, B200PPF_ERR_UNSUPPORTED, "cloud upload: more than 2^32 points");
    return prep_upload_xyz(ctx, host, n, stride, out);
}

int b200ppf_cloud_download(b200ppf_ctx *ctx, const b200ppf_cloud *cloud, float *host, size_t stride, size_t noff,
                           size_t coff) {
    CHECK_CTX(ctx);
    if (!cloud || (cloud->n && !host)) return fail_msg(ctx, B200PPF_ERR_INVALID, "cloud download: null argument");
    if (stride < 3 || (noff && (noff < 3 || noff + 3 > stride)) || (coff && (coff < 3 || coff >= stride)) ||
        (noff && coff && coff >= noff && coff < noff + 3))
        return fail_msg(ctx, B200PPF_ERR_INVALID, "cloud download: stride / offsets do not describe [x y z .. nx ny nz .. curvature]");
    return prep_download(ctx, cloud, host, stride, noff, coff);
}

int b200ppf_voxel_grid(b200ppf_ctx *ctx, const b200ppf_cloud *in, const float *leaf3, b200ppf_cloud **out) {
    CHECK_CTX(ctx);
    if (!in || !leaf3 || !out) return fail_msg(ctx, B200PPF_ERR_INVALID, "voxel grid: null argument");
    *out = nullptr;
    for (int k = 0; k < 3; ++k)
        if (!(leaf3[k] > 0.0f) || !std::isfinite(leaf3[k])) return fail_msg(ctx, B200PPF_ERR_INVALID, "voxel grid: leaf size must be positive");
    ctx->error.clear();
    return prep_voxel_grid(ctx, in, leaf3, out);
}

int b200ppf_knn(b200ppf_ctx *ctx, const b200ppf_cloud *cloud, int k, uint32_t *idx_host, float *d2_host) {
    CHECK_CTX(ctx);
    if (!cloud) return fail_msg(ctx, B200PPF_ERR_INVALID, "knn: null cloud");
    if (k < 1 || k > 128 || (size_t)k > std::max<size_t>(1, cloud->n))
        return fail_msg(ctx, B200PPF_ERR_INVALID, "knn: k must be in 1 .. min(n, 128)");
    return prep_knn(ctx, cloud, k, idx_host, d2_host);
}

int b200ppf_statistical_outlier_removal(b200ppf_ctx *ctx, const b200ppf_cloud *in, int mean_k, double stddev_mul,
                                        b200ppf_cloud **out, uint32_t *kept_host, float *distances_host, double *threshold) {
    CHECK_CTX(ctx);
    if (!in || !out) return fail_msg(ctx, B200PPF_ERR_INVALID, "statistical outlier removal: null argument");
    *out = nullptr;
    if (mean_k < 1 || mean_k > 127) return fail_msg(ctx, B200PPF_ERR_INVALID, "statistical outlier removal: mean_k must be in 1 .. 127");
    if (in->n < (size_t)mean_k + 1)
        return fail_msg(ctx, B200PPF_ERR_INVALID, "statistical outlier removal: the cloud needs more than mean_k points");
    return prep_sor(ctx, in, mean_k, stddev_mul, out, kept_host, distances_host, threshold);
}

int b200ppf_normal_estimation(b200ppf_ctx *ctx, b200ppf_cloud *cloud, int k, const float *viewpoint3, int cov_mode) {
    CHECK_CTX(ctx);
    if (!cloud) return fail_msg(ctx, B200PPF_ERR_INVALID, "normal estimation: null cloud");
    if (k < 1 || k > 128) return fail_msg(ctx, B200PPF_ERR_INVALID, "normal estimation: k must be in 1 .. 128");
    if (cov_mode != B200PPF_COVARIANCE_SHIFTED && cov_mode != B200PPF_COVARIANCE_RAW)
        return fail_msg(ctx, B200PPF_ERR_INVALID, "normal estimation: unknown covariance mode");
    return prep_normals(ctx, cloud, k, viewpoint3, cov_mode);
}

int b200ppf_curvature_edges(b200ppf_ctx *ctx, const b200ppf_cloud *in, float threshold, b200ppf_cloud **out) {
    CHECK_CTX(ctx);
    if (!in || !out) return fail_msg(ctx, B200PPF_ERR_INVALID, "curvature edges: null argument");
    *out = nullptr;
    return prep_curvature_edges(ctx, in, threshold, out);
}

int b200ppf_normalize_normals(b200ppf_ctx *ctx, b200ppf_cloud *cloud) {
    CHECK_CTX(ctx);
    if (!cloud) return fail_msg(ctx, B200PPF_ERR_INVALID, "normalize normals: null cloud");
    return prep_renormalize(ctx, cloud);
}

int b200ppf_frustum_corners(const float *depth, int rows, int cols, int box_x, int box_y, int box_w, int box_h, double fx,
                            double fy, double ppx, double ppy, float *corners12) {
    if (!depth || !corners12 || rows < 1 || cols < 1 || box_w < 0 || box_h < 0 || box_x >= cols || box_y >= rows ||
        box_x + box_w + 30 < 0 || box_y + box_h + 30 < 0 || !(fx != 0.0) || !(fy != 0.0))
        return fail_msg(nullptr, B200PPF_ERR_INVALID, "frustum corners: bad argument");
    prep_frustum_corners(depth, rows, cols, box_x, box_y, box_w, box_h, fx, fy, ppx, ppy, corners12);
    return B200PPF_OK;
}

int b200ppf_crop_pyramid(b200ppf_ctx *ctx, const b200ppf_cloud *in, const float *corners12, b200ppf_cloud **out,
                         uint32_t *kept_host) {
    CHECK_CTX(ctx);
    if (!in || !corners12 || !out) return fail_msg(ctx, B200PPF_ERR_INVALID, "crop: null argument");
    *out = nullptr;
    for (int k = 0; k < 12; ++k)
        if (!std::isfinite(corners12[k])) return fail_msg(ctx, B200PPF_ERR_INVALID, "crop: non-finite corner");
    return prep_crop_pyramid(ctx, in, corners12, out, kept_host);
}

int b200ppf_debug_knn_host(const float *xyz, size_t n, size_t stride, int k, int mode, float cell_edge, const float *viewpoint3,
                           int cov_mode, uint32_t *idx, float *d2, float *mean_dist, float *normals4) {
    if ((n && !xyz) || stride < 3 || k < 1 || k > 128 || (size_t)k > std::max<size_t>(1, n) || mode < 0 || mode > 2 ||
        (mode == 0 && (!idx || !d2)) || (mode == 1 && (!mean_dist || k < 2)) || (mode == 2 && !normals4))
        return fail_msg(nullptr, B200PPF_ERR_INVALID, "debug knn: bad argument");
    return prep_debug_knn_host(xyz, n, stride, k, mode, cell_edge, viewpoint3, cov_mode, idx, d2, mean_dist, normals4);
}

/* ---- align -------------------------------------------------------------------------------- */

int b200ppf_icp_refine(b200ppf_ctx *ctx, const b200ppf_cloud *model, const b200ppf_cloud *scene,
                       const b200ppf_icp_params *params, double *poses16, size_t n_poses, double *residuals,
                       uint64_t *iterations) {
    CHECK_CTX(ctx);
    if (!model || !scene || (n_poses && !poses16)) return fail_msg(ctx, B200PPF_ERR_INVALID, "icp: null argument");
    DeviceGuard guard(ctx->device);
    b200ppf_icp_params p = {100, 0.005f, 2.5f, 8};  // the reference's ICP(100, 0.005f, 2.5f, 8)
    if (params) p = *params;
    return k6_icp_refine(ctx, model, scene, p.max_iterations, p.tolerance, p.rejection_scale, p.num_levels, poses16,
                         n_poses, residuals, iterations);
}

int b200ppf_register(b200ppf_ctx *ctx, const b200ppf_cloud *model, const b200ppf_table *t,
                     const b200ppf_cloud *scene, size_t ref_rate, float pos_thr, float rot_thr, float *final16,
                     float *poses16, uint32_t *votes, size_t *n_out) {
    CHECK_CTX(ctx);
    if (!scene || !poses16 || !votes || !n_out) return fail_msg(ctx, B200PPF_ERR_INVALID, "register: null argument");
    *n_out = 0;
    if (ref_rate == 0) ref_rate = 1;
    const size_t ref_count = (scene->n + ref_rate - 1) / ref_rate;
    if (ref_count == 0) return fail_msg(ctx, B200PPF_ERR_STATE, "register: empty scene");
    if (ctx->hyps_cap < ref_count) {
        if (ctx->d_hyps) cudaFree(ctx->d_hyps);
        ctx->d_hyps = nullptr;
        ctx->hyps_cap = 0;
        PPF_CUDA(ctx, cudaMalloc(&ctx->d_hyps, ref_count * sizeof(b200ppf_hypothesis)));
        ctx->hyps_cap = ref_count;
    }
    b200ppf_hypothesis *const reg_target = ctx->d_hyps;
    int rc = k3_vote(ctx, model, t, scene, 0, ref_rate, ref_count, &reg_target, 1, 0, 1);
    if (rc) return rc;
    rc = k4_cluster(ctx, ctx->d_hyps, ref_count, pos_thr, rot_thr, poses16, votes, n_out);
    if (rc) return rc;
    if (*n_out && final16) memcpy(final16, poses16, 16 * sizeof(float));
    return B200PPF_OK;
}

void b200ppf_object_params_default(b200ppf_object_params *p) {
    if (!p) return;
    p->leaf = 0.01f;
    p->sor_mean_k = 50;       // OutlierProcessing(50, thresh), src/YOLO_cropping_ppf_test.cpp:98
    p->sor_stddev_mul = 1.0;
    p->normal_k = 30;         // NormalEstimation(30), :101
    p->edge_curvature = 0.03f;  // EdgeExtraction(0.03), :103
    p->ref_rate = 20;         // relSceneSampleStep = 0.05, include/CloudProcessing.h:481
    p->pos_thr = 0.01f;       // PCL defaults of PPFRegistration
    p->rot_thr = 20.0f / 180.0f * 3.14159265358979f;
    p->icp_poses = 5;         // N = 5, include/CloudProcessing.h:508
    p->icp = b200ppf_icp_params{100, 0.005f, 2.5f, 8};
}

int b200ppf_match_object(b200ppf_ctx *ctx, const b200ppf_cloud *scene, const float *corners12, const b200ppf_cloud *model,
                         const b200ppf_table *table, const b200ppf_object_params *params, b200ppf_object_result *res,
                         b200ppf_cloud **object_out, b200ppf_cloud **edges_out) {
    if (!ctx) return fail_msg(nullptr, B200PPF_ERR_INVALID, "null context");
    if (object_out) *object_out = nullptr;
    if (edges_out) *edges_out = nullptr;
    if (!scene || !corners12 || !model || !table || !res) return fail_msg(ctx, B200PPF_ERR_INVALID, "match object: null argument");
    b200ppf_object_params p;
    b200ppf_object_params_default(&p);
    if (params) p = *params;
    memset(res, 0, sizeof(*res));
    const auto t0 = std::chrono::steady_clock::now();
    struct Owned {  // intermediate clouds live until the call returns
        b200ppf_cloud *c = nullptr;
        ~Owned() { if (c) b200ppf_cloud_free(c); }
        b200ppf_cloud *release() { b200ppf_cloud *r = c; c = nullptr; return r; }
    } cropped, sampled, object, edges;
    int rc = b200ppf_crop_pyramid(ctx, scene, corners12, &cropped.c, nullptr);
    if (rc) return rc;
    res->crop_ms = ctx->timings.prep_ms;
    res->n_cropped = (uint32_t)cropped.c->n;
    if (cropped.c->n == 0) return fail_msg(ctx, B200PPF_ERR_STATE, "match object: no scene point inside the box's frustum");
    const float leaf3[3] = {p.leaf, p.leaf, p.leaf};
    if ((rc = b200ppf_voxel_grid(ctx, cropped.c, leaf3, &sampled.c))) return rc;
    res->voxel_ms = ctx->timings.prep_ms;
    res->n_sampled = (uint32_t)sampled.c->n;
    if ((rc = b200ppf_statistical_outlier_removal(ctx, sampled.c, p.sor_mean_k, p.sor_stddev_mul, &object.c, nullptr, nullptr, nullptr)))
        return rc;
    res->outlier_ms = ctx->timings.prep_ms;
    res->n_filtered = (uint32_t)object.c->n;
    if ((rc = b200ppf_normal_estimation(ctx, object.c, p.normal_k, nullptr, B200PPF_COVARIANCE_SHIFTED))) return rc;
    res->normals_ms = ctx->timings.prep_ms;
    if (p.edge_curvature > 0.0f && edges_out) {  // the edge cloud does not enter the PCL matching: extracted on request only
        if ((rc = b200ppf_curvature_edges(ctx, object.c, p.edge_curvature, &edges.c))) return rc;
        res->edges_ms = ctx->timings.prep_ms;
        res->n_edges = (uint32_t)edges.c->n;
    }
    if ((rc = b200ppf_normalize_normals(ctx, object.c))) return rc;
    if (edges.c && (rc = b200ppf_normalize_normals(ctx, edges.c))) return rc;
    float final16[16], poses16[3 * 16];
    uint32_t votes[3] = {0, 0, 0};
    size_t n_poses = 0;
    if ((rc = b200ppf_register(ctx, model, table, object.c, p.ref_rate, p.pos_thr, p.rot_thr, final16, poses16, votes, &n_poses)))
        return rc;
    b200ppf_timings tm;
    b200ppf_get_timings(ctx, &tm);
    res->match_ms = tm.grid_ms + tm.vote_ms + tm.pose_ms + tm.cluster_ms;
    res->n_poses = (uint32_t)n_poses;
    if (n_poses == 0) return fail_msg(ctx, B200PPF_ERR_STATE, "match object: the matching returned no pose");  // reference: exit(0), :502-507
    double refined[3 * 16], residuals[3] = {0, 0, 0};
    for (size_t k = 0; k < n_poses * 16; ++k) refined[k] = (double)poses16[k];
    const size_t n_icp = std::min<size_t>(n_poses, p.icp_poses > 0 ? (size_t)p.icp_poses : 0);
    if (n_icp) {
        if ((rc = b200ppf_icp_refine(ctx, model, object.c, &p.icp, refined, n_icp, residuals, nullptr))) return rc;
        res->icp_ms = ctx->timings.icp_ms;
    }
    memcpy(res->pose, refined, 16 * sizeof(double));  // resultsSub[0]
    res->residual = residuals[0];
    res->votes = votes[0];
    res->total_wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (object_out) *object_out = object.release();
    if (edges_out) *edges_out = edges.release();
    return B200PPF_OK;
}

}  // extern "C"
