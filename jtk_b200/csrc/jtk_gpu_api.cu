// jtk_gpu_api.cu -- C ABI (include/jtk_gpu.h) over the sm_100a kernels.  No CPU fallback: every compute
// entry point needs a CUDA device and fails with JTK_ECUDA otherwise.
#include "../../include/jtk_gpu.h"
#include "phmm_dev.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace jtk {
int cols_per_lane_for_radius(int radius);
cudaError_t launch_modtable(const KParams &p, int C, int grid, cudaStream_t st);
cudaError_t launch_likelihood(const KParams &p, int C, int grid, cudaStream_t st);
int warps_per_cta();
} // namespace jtk

using namespace jtk;

namespace {

thread_local std::string g_create_error;

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0; // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T> struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMallocHost((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

inline uint8_t base_code(uint8_t c) {
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}

void pack_model(const jtk_hmm_params *h, float *m) {
    std::memset(m, 0, sizeof(float) * kModelFloats);
    const double tr[9] = { h->mat_mat, h->mat_ins, h->mat_del, h->ins_mat, h->ins_ins, h->ins_del,
                           h->del_mat, h->del_ins, h->del_del };
    for (int k = 0; k < 9; k++) m[k] = (float)tr[k];
    for (int t = 0; t < 4; t++)
        for (int q = 0; q < 4; q++) {
            m[kOffEM + t * 8 + q] = (float)h->mat_emit[4 * t + q];
            m[kOffEMT + q * 4 + t] = (float)h->mat_emit[4 * t + q];
        }
    for (int c = 0; c < 5; c++)
        for (int q = 0; q < 4; q++) m[kOffEI + c * 8 + q] = (float)h->ins_emit[4 * c + q];
}

// banded global edit-distance alignment: the guide of the "bootstrap" likelihood
// (likelihood_gains.rs:27-28).  Band |i-j| <= radius + |Lr-Lt|; traceback prefers
// diagonal, then Del, then Ins.  Returns false when the band cannot connect the corners.
bool edit_ops(const uint8_t *t, int Lt, const uint8_t *q, int Lr, int radius, std::vector<uint8_t> &ops,
              std::vector<int> &D) {
    const int R = radius + std::abs(Lr - Lt);
    const int INF = 1 << 29;
    const size_t W = (size_t)Lt + 1;
    D.assign((size_t)(Lr + 1) * W, INF);
    for (int i = 0; i <= Lr; i++) {
        int jlo = std::max(0, i - R), jhi = std::min(Lt, i + R);
        for (int j = jlo; j <= jhi; j++) {
            int v = INF;
            if (i == 0 && j == 0) v = 0;
            else {
                if (i > 0 && j > 0) v = std::min(v, D[(size_t)(i - 1) * W + j - 1] + (base_code(q[i - 1]) != base_code(t[j - 1])));
                if (j > 0) v = std::min(v, D[(size_t)i * W + j - 1] + 1);
                if (i > 0) v = std::min(v, D[(size_t)(i - 1) * W + j] + 1);
            }
            D[(size_t)i * W + j] = v;
        }
    }
    if (D[(size_t)Lr * W + Lt] >= INF) return false;
    ops.clear();
    int i = Lr, j = Lt;
    while (i > 0 || j > 0) {
        const int v = D[(size_t)i * W + j];
        if (i > 0 && j > 0) {
            const int mis = base_code(q[i - 1]) != base_code(t[j - 1]);
            if (D[(size_t)(i - 1) * W + j - 1] + mis == v) { ops.push_back(mis ? JTK_OP_MISMATCH : JTK_OP_MATCH); i--; j--; continue; }
        }
        if (j > 0 && D[(size_t)i * W + j - 1] + 1 == v) { ops.push_back(JTK_OP_DEL); j--; continue; }
        ops.push_back(JTK_OP_INS); i--;
    }
    std::reverse(ops.begin(), ops.end());
    return true;
}

} // namespace

struct jtk_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    uint64_t launches = 0;
    float last_ms = 0.f;
    // device buffers
    DevBuf<DevPair> d_pairs;
    DevBuf<uint8_t> d_codes;
    DevBuf<uint32_t> d_bits;
    DevBuf<float> d_models;
    DevBuf<float2> d_frows;
    DevBuf<int32_t> d_kf;
    DevBuf<float> d_delta;
    DevBuf<double> d_lk;
    DevBuf<int> d_counter;
    // pinned host staging
    PinBuf<DevPair> h_pairs;
    PinBuf<uint8_t> h_codes;
    PinBuf<uint32_t> h_bits;
    PinBuf<float> h_delta;
    PinBuf<double> h_lk;
    // host scratch
    std::vector<uint32_t> tmpl_code_off;
    std::vector<uint8_t> tmp_ops;
    std::vector<int> tmp_D;

    int fail(int code, const std::string &m) { err = m; return code; }
    int cuda_fail(cudaError_t e, const char *what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? JTK_ENOMEM : JTK_ECUDA;
    }
};

#define CU(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx->cuda_fail(e_, what); } while (0)

extern "C" {

int jtk_hmm_num_row(void) { return JTK_NUM_ROW; }
int jtk_hmm_copy_size(void) { return JTK_COPY_SIZE; }
int jtk_hmm_del_size(void) { return JTK_DEL_SIZE; }

const char *jtk_last_error(const jtk_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
uint64_t jtk_ctx_launch_count(const jtk_ctx *ctx) { return ctx ? ctx->launches : 0; }
float jtk_ctx_last_kernel_ms(const jtk_ctx *ctx) { return ctx ? ctx->last_ms : 0.f; }

int jtk_ctx_create(int device, size_t workspace_bytes, jtk_ctx **out) {
    (void)workspace_bytes;
    if (!out) { g_create_error = "out is NULL"; return JTK_EINVAL; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return JTK_ECUDA;
    }
    if (device < 0) { if ((e = cudaGetDevice(&device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return JTK_ECUDA; } }
    if (device >= n) { g_create_error = "device index out of range"; return JTK_EINVAL; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return JTK_ECUDA; }
    jtk_ctx *ctx = new jtk_ctx();
    ctx->device = device;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete ctx;
        return JTK_ECUDA;
    }
    *out = ctx;
    return JTK_OK;
}

void jtk_ctx_destroy(jtk_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->d_pairs.release(); ctx->d_codes.release(); ctx->d_bits.release(); ctx->d_models.release();
    ctx->d_frows.release(); ctx->d_kf.release(); ctx->d_delta.release(); ctx->d_lk.release(); ctx->d_counter.release();
    ctx->h_pairs.release(); ctx->h_codes.release(); ctx->h_bits.release(); ctx->h_delta.release(); ctx->h_lk.release();
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int64_t jtk_band_cell_count(const uint8_t *ops, int n_ops, int Lt, int Lr, int radius) {
    if (!ops || n_ops < 0 || Lt < 0 || Lr < 0 || radius < 0) return -1;
    int i = 0, j = 0;
    int64_t total = 0;
    auto width = [&](int cen, int s) -> int64_t {
        int lo = std::max(std::max(cen - radius, 0), s - Lt), hi = std::min(std::min(cen + radius, Lr), s);
        return hi >= lo ? hi - lo + 1 : 0;
    };
    total += width(0, 0);
    for (int k = 0; k < n_ops; k++) {
        switch (ops[k]) {
        case JTK_OP_MATCH: case JTK_OP_MISMATCH:
            if (i >= Lr || j >= Lt) return -1;
            total += width(i, i + j + 1);
            i++; j++;
            total += width(i, i + j);
            break;
        case JTK_OP_INS: if (i >= Lr) return -1; i++; total += width(i, i + j); break;
        case JTK_OP_DEL: if (j >= Lt) return -1; j++; total += width(i, i + j); break;
        default: return -1;
        }
    }
    return (i == Lr && j == Lt) ? total : -1;
}

} // extern "C"

namespace {

struct BatchShape {
    int max_nd = 0;
    uint64_t table_floats = 0;
};

// Encode templates / reads / guide paths into the device layout of phmm_dev.cuh (host staging buffers).
int pack_batch(jtk_ctx *ctx, int n_pairs, int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
               const uint8_t *read_concat, const uint32_t *read_off, const uint8_t *ops_concat,
               const uint32_t *ops_off, const uint8_t *strand, const uint32_t *tmpl_idx, int radius,
               bool want_table, BatchShape &shape, size_t &code_bytes, size_t &bit_words) {
    // sizes
    size_t cb = 0;
    for (int t = 0; t < n_tmpl; t++) {
        if (tmpl_off[t + 1] < tmpl_off[t]) return ctx->fail(JTK_EINVAL, "tmpl_off is not monotone");
        cb += 2 * (size_t)kCodePad + (tmpl_off[t + 1] - tmpl_off[t]) + 1;
    }
    size_t bwords = 0;
    for (int p = 0; p < n_pairs; p++) {
        if (read_off[p + 1] < read_off[p]) return ctx->fail(JTK_EINVAL, "read_off is not monotone");
        if (tmpl_idx[p] >= (uint32_t)n_tmpl) return ctx->fail(JTK_EINVAL, "tmpl_idx out of range");
        const size_t Lr = read_off[p + 1] - read_off[p];
        const size_t Lt = tmpl_off[tmpl_idx[p] + 1] - tmpl_off[tmpl_idx[p]];
        cb += 2 * (size_t)kCodePad + Lr + 2;
        bwords += (Lt + Lr + 1 + 31) / 32 + 1;
    }
    if (cb >= (size_t)0xffffffffu) return ctx->fail(JTK_EINVAL, "batch too large: split it (code bytes exceed 4 GiB)");
    CU(ctx->h_codes.reserve(cb), "cudaMallocHost codes");
    CU(ctx->h_bits.reserve(bwords), "cudaMallocHost bits");
    CU(ctx->h_pairs.reserve((size_t)n_pairs), "cudaMallocHost pairs");
    uint8_t *codes = ctx->h_codes.p;
    uint32_t *bits = ctx->h_bits.p;
    std::memset(codes, 4, cb);
    std::memset(bits, 0, bwords * sizeof(uint32_t));
    ctx->tmpl_code_off.resize((size_t)n_tmpl);
    size_t pos = 0;
    for (int t = 0; t < n_tmpl; t++) {
        const uint8_t *s = tmpl_concat + tmpl_off[t];
        const size_t L = tmpl_off[t + 1] - tmpl_off[t];
        pos += kCodePad;
        ctx->tmpl_code_off[t] = (uint32_t)pos;
        for (size_t j = 1; j <= L; j++) codes[pos + j] = base_code(s[j - 1]);
        pos += L + 1 + kCodePad;
    }
    size_t bpos = 0;
    uint64_t tab = 0;
    shape.max_nd = 0;
    for (int p = 0; p < n_pairs; p++) {
        const uint32_t ti = tmpl_idx[p];
        const int Lt = (int)(tmpl_off[ti + 1] - tmpl_off[ti]);
        const int Lr = (int)(read_off[p + 1] - read_off[p]);
        const uint8_t *q = read_concat + read_off[p];
        pos += kCodePad;
        DevPair &dp = ctx->h_pairs.p[p];
        dp.tb_off = ctx->tmpl_code_off[ti];
        dp.rb_off = (uint32_t)pos;
        for (int i = 1; i <= Lr; i++) {
            const uint8_t qc = base_code(q[i - 1]);
            const uint8_t cx = i >= 2 ? base_code(q[i - 2]) : 4;
            codes[pos + i] = (uint8_t)((cx << 3) | qc);
        }
        pos += (size_t)Lr + 2 + kCodePad;
        // guide path -> centre increments
        const uint8_t *ops;
        int n_ops;
        if (ops_concat) {
            if (ops_off[p + 1] < ops_off[p]) return ctx->fail(JTK_EINVAL, "ops_off is not monotone");
            ops = ops_concat + ops_off[p];
            n_ops = (int)(ops_off[p + 1] - ops_off[p]);
        } else {
            const uint8_t *t = tmpl_concat + tmpl_off[ti];
            if (!edit_ops(t, Lt, q, Lr, radius, ctx->tmp_ops, ctx->tmp_D))
                return ctx->fail(JTK_EINVAL, "bootstrap alignment: band cannot connect the corners (pair " + std::to_string(p) + ")");
            ops = ctx->tmp_ops.data();
            n_ops = (int)ctx->tmp_ops.size();
        }
        uint32_t *bw = bits + bpos;
        int i = 0, j = 0, s = 0;
        for (int k = 0; k < n_ops; k++) {
            const uint8_t op = ops[k];
            if (op <= JTK_OP_MISMATCH) { i++; j++; bw[(s + 1) >> 5] |= 1u << ((s + 1) & 31); s += 2; }
            else if (op == JTK_OP_INS) { i++; bw[s >> 5] |= 1u << (s & 31); s += 1; }
            else if (op == JTK_OP_DEL) { j++; s += 1; }
            else return ctx->fail(JTK_EINVAL, "invalid op code in pair " + std::to_string(p));
            if (i > Lr || j > Lt) break;
        }
        if (i != Lr || j != Lt)
            return ctx->fail(JTK_EINVAL, "ops of pair " + std::to_string(p) + " do not span (template, read): consumed (" +
                                             std::to_string(j) + "," + std::to_string(i) + ") of (" + std::to_string(Lt) + "," +
                                             std::to_string(Lr) + ")");
        dp.bits_off = (uint32_t)bpos;
        bpos += (size_t)(Lt + Lr + 1 + 31) / 32 + 1;
        dp.Lt = Lt; dp.Lr = Lr;
        dp.model = strand[p] ? 0 : 1;
        dp.pad_ = 0;
        dp.tab_off = tab;
        if (want_table) tab += (uint64_t)(Lt + 1) * kNumRow;
        shape.max_nd = std::max(shape.max_nd, Lt + Lr + 1);
    }
    shape.table_floats = tab;
    code_bytes = cb;
    bit_words = bwords;
    return JTK_OK;
}

int run_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_pairs, int n_tmpl,
              const uint8_t *tmpl_concat, const uint32_t *tmpl_off, const uint8_t *read_concat,
              const uint32_t *read_off, const uint8_t *ops_concat, const uint32_t *ops_off, const uint8_t *strand,
              const uint32_t *tmpl_idx, int radius, bool table, int rows, double *out_lk, double *out_table,
              const uint64_t *table_off) {
    if (!ctx) return JTK_EINVAL;
    if (!fwd || !rev || !tmpl_concat || !tmpl_off || !read_concat || !read_off || !strand || !tmpl_idx || !out_lk)
        return ctx->fail(JTK_EINVAL, "null argument");
    if (n_pairs < 0 || n_tmpl < 0) return ctx->fail(JTK_EINVAL, "negative count");
    if (table && out_table && !table_off) return ctx->fail(JTK_EINVAL, "table_off is NULL");
    if (n_pairs == 0) return JTK_OK;
    const int C = cols_per_lane_for_radius(radius);
    if (radius < 0 || C == 0 || C > 4) return ctx->fail(JTK_EINVAL, "radius out of range (0..62)");
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    BatchShape shape;
    size_t code_bytes = 0, bit_words = 0;
    const bool want_table = table && out_table;
    int rc = pack_batch(ctx, n_pairs, n_tmpl, tmpl_concat, tmpl_off, read_concat, read_off, ops_concat, ops_off, strand,
                        tmpl_idx, radius, want_table, shape, code_bytes, bit_words);
    if (rc) return rc;
    float models[2 * kModelFloats];
    pack_model(fwd, models);
    pack_model(rev, models + kModelFloats);

    const int wpc = warps_per_cta();
    int grid = (n_pairs + wpc - 1) / wpc;
    const int max_grid = ctx->sm_count * 4;
    if (grid > max_grid) grid = max_grid;
    const int NSLOT = 32 * C;
    KParams kp{};
    CU(ctx->d_pairs.reserve((size_t)n_pairs), "cudaMalloc pairs");
    CU(ctx->d_codes.reserve(code_bytes), "cudaMalloc codes");
    CU(ctx->d_bits.reserve(bit_words), "cudaMalloc bits");
    CU(ctx->d_models.reserve(2 * kModelFloats), "cudaMalloc models");
    CU(ctx->d_lk.reserve((size_t)n_pairs), "cudaMalloc lk");
    CU(ctx->d_counter.reserve(1), "cudaMalloc counter");
    CU(ctx->h_lk.reserve((size_t)n_pairs), "cudaMallocHost lk");
    cudaStream_t st = ctx->stream;
    if (table) {
        const size_t slots = (size_t)grid * wpc;
        kp.frow_stride = (size_t)(shape.max_nd + 6) * NSLOT;
        kp.kf_stride = (size_t)shape.max_nd + 6;
        CU(ctx->d_frows.reserve(slots * kp.frow_stride), "cudaMalloc forward rows");
        CU(ctx->d_kf.reserve(slots * kp.kf_stride), "cudaMalloc scale exponents");
        if (want_table) {
            CU(ctx->d_delta.reserve((size_t)shape.table_floats), "cudaMalloc table");
            CU(ctx->h_delta.reserve((size_t)shape.table_floats), "cudaMallocHost table");
        } else {
            // the kernel always writes its table; give it a scratch area
            uint64_t tab = 0;
            for (int p = 0; p < n_pairs; p++) { ctx->h_pairs.p[p].tab_off = tab; tab += (uint64_t)(ctx->h_pairs.p[p].Lt + 1) * kNumRow; }
            CU(ctx->d_delta.reserve((size_t)tab), "cudaMalloc table");
        }
    }
    CU(cudaMemcpyAsync(ctx->d_pairs.p, ctx->h_pairs.p, sizeof(DevPair) * (size_t)n_pairs, cudaMemcpyHostToDevice, st), "H2D pairs");
    CU(cudaMemcpyAsync(ctx->d_codes.p, ctx->h_codes.p, code_bytes, cudaMemcpyHostToDevice, st), "H2D codes");
    CU(cudaMemcpyAsync(ctx->d_bits.p, ctx->h_bits.p, bit_words * sizeof(uint32_t), cudaMemcpyHostToDevice, st), "H2D bits");
    CU(cudaMemcpyAsync(ctx->d_models.p, models, sizeof(models), cudaMemcpyHostToDevice, st), "H2D models");
    CU(cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(int), st), "memset counter");
    kp.pairs = ctx->d_pairs.p; kp.n_pairs = n_pairs;
    kp.codes = ctx->d_codes.p; kp.bits = ctx->d_bits.p; kp.models = ctx->d_models.p;
    kp.radius = radius; kp.rows = rows;
    kp.frows = ctx->d_frows.p; kp.kf = ctx->d_kf.p;
    kp.out_delta = ctx->d_delta.p; kp.out_lk = ctx->d_lk.p; kp.counter = ctx->d_counter.p;
    CU(cudaEventRecord(ctx->ev0, st), "event");
    CU(table ? launch_modtable(kp, C, grid, st) : launch_likelihood(kp, C, grid, st), "kernel launch");
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev1, st), "event");
    CU(cudaMemcpyAsync(ctx->h_lk.p, ctx->d_lk.p, sizeof(double) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st), "D2H lk");
    if (want_table)
        CU(cudaMemcpyAsync(ctx->h_delta.p, ctx->d_delta.p, sizeof(float) * (size_t)shape.table_floats, cudaMemcpyDeviceToHost, st), "D2H table");
    CU(cudaStreamSynchronize(st), "kernel execution");
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    std::memcpy(out_lk, ctx->h_lk.p, sizeof(double) * (size_t)n_pairs);
    if (want_table) {
        for (int p = 0; p < n_pairs; p++) {
            const DevPair &dp = ctx->h_pairs.p[p];
            const float *src = ctx->h_delta.p + dp.tab_off;
            double *dst = out_table + table_off[p];
            const double lk = out_lk[p];
            const size_t n = (size_t)(dp.Lt + 1) * kNumRow;
            for (size_t k = 0; k < n; k++) dst[k] = (src[k] <= -1.0e9f || !(lk > -INFINITY)) ? JTK_TABLE_NEG : lk + (double)src[k];
        }
    }
    return JTK_OK;
}

} // namespace

extern "C" {

int jtk_hmm_modtable_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_pairs,
                           int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                           const uint8_t *read_concat, const uint32_t *read_off,
                           const uint8_t *ops_concat, const uint32_t *ops_off,
                           const uint8_t *strand, const uint32_t *tmpl_idx, int radius,
                           double *out_lk, double *out_table, const uint64_t *table_off) {
    if (ctx && (!ops_concat || !ops_off)) return ctx->fail(JTK_EINVAL, "modification table needs guide ops");
    return run_batch(ctx, fwd, rev, n_pairs, n_tmpl, tmpl_concat, tmpl_off, read_concat, read_off, ops_concat, ops_off,
                     strand, tmpl_idx, radius, true, 14, out_lk, out_table, table_off);
}

int jtk_hmm_likelihood_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_pairs,
                             int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                             const uint8_t *read_concat, const uint32_t *read_off,
                             const uint8_t *ops_concat, const uint32_t *ops_off,
                             const uint8_t *strand, const uint32_t *tmpl_idx, int radius, double *out_lk) {
    return run_batch(ctx, fwd, rev, n_pairs, n_tmpl, tmpl_concat, tmpl_off, read_concat, read_off, ops_concat, ops_off,
                     strand, tmpl_idx, radius, false, 0, out_lk, nullptr, nullptr);
}

} // extern "C"
