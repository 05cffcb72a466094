// jtk_gpu_api.cu -- C ABI (include/jtk_gpu.h) over the sm_100a kernels.  No CPU fallback: every compute
// entry point needs a CUDA device and fails with JTK_ECUDA otherwise.
#include "../../include/jtk_gpu.h"
#include "phmm_dev.cuh"
#include "lc_host.h"
#include "mcmc_dev.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

namespace jtk {
int cols_per_lane_for_radius(int radius);
cudaError_t launch_modtable(const KParams &p, int C, int grid_fwd, int grid_bwd, cudaStream_t st);
cudaError_t launch_modtable_fused(const KParams &p, int C, int grid, cudaStream_t st);
int fused_dyn_smem(int C, int smem_rb);
int fused_grid(int C, int rows, int dyn, int sm_count);
size_t fused_ckpt_bytes(int C, int max_nd);
size_t fused_kb_ints(int max_nd);
int fwdrows_ctas_per_sm(int C);
int fwdinfo_words();
int fwd_pad_rows(int C);
cudaError_t launch_likelihood(const KParams &p, int C, int grid, cudaStream_t st);
int warps_per_cta();
int frow_slots_per_row(int C);
int frow_extra_rows();
int modtable_ctas_per_sm(int C, int rows);
int modtable_warps_per_cta(int C);
int likelihood_pairs2_max_radius();
int likelihood_pairs2_pad_rows();
cudaError_t launch_likelihood_pairs2(const KParams &p, int grid, cudaStream_t st);
int modtable_dyn_smem(int C, int smem_rb);
cudaError_t launch_fit(const KParams &p, int C, int grid, double *acc90, cudaStream_t st);
cudaError_t launch_fp32_peak(int mode, int blocks, int threads, int iters, float *sink, cudaStream_t st);
} // namespace jtk

using namespace jtk;

namespace {

thread_local std::string g_create_error;

// Device blocks released by a batch are kept by the context and handed to the next batch: a chunk pipeline creates and
// destroys batches of similar size all the time, and cudaMalloc / cudaFree (which synchronises) would dominate.
// All work of a context runs on one stream, so reuse needs no extra ordering.
struct DevPool {
    struct Block { void *p; size_t bytes; };
    std::vector<Block> free_blocks;
    size_t cached = 0;
    std::mutex mu; // take / give run on the owning context's thread, trim_all() on whichever thread ran out of memory
    static constexpr size_t kMaxCached = (size_t)24 << 30;
    static std::mutex &reg_mu() { static std::mutex m; return m; }
    static std::vector<DevPool *> &registry() { static std::vector<DevPool *> r; return r; }
    DevPool() { std::lock_guard<std::mutex> g(reg_mu()); registry().push_back(this); }
    ~DevPool() {
        std::lock_guard<std::mutex> g(reg_mu());
        auto &r = registry();
        r.erase(std::remove(r.begin(), r.end(), this), r.end());
    }
    DevPool(const DevPool &) = delete;
    DevPool &operator=(const DevPool &) = delete;
    void *take(size_t bytes, size_t &got) {
        std::lock_guard<std::mutex> g(mu);
        size_t best = free_blocks.size();
        for (size_t k = 0; k < free_blocks.size(); k++)
            if (free_blocks[k].bytes >= bytes && free_blocks[k].bytes <= 2 * bytes + (1 << 20) &&
                (best == free_blocks.size() || free_blocks[k].bytes < free_blocks[best].bytes)) best = k;
        if (best == free_blocks.size()) return nullptr;
        Block blk = free_blocks[best];
        free_blocks.erase(free_blocks.begin() + (ptrdiff_t)best);
        cached -= blk.bytes;
        got = blk.bytes;
        return blk.p;
    }
    void give(void *p, size_t bytes) {
        std::lock_guard<std::mutex> g(mu);
        if (cached + bytes > kMaxCached) { cudaFree(p); return; }
        free_blocks.push_back({ p, bytes });
        cached += bytes;
    }
    void clear() {
        std::lock_guard<std::mutex> g(mu);
        for (auto &b : free_blocks) cudaFree(b.p);
        free_blocks.clear();
        cached = 0;
    }
    // an allocation failed somewhere: every context of the process gives its cached blocks back to the driver
    static bool trim_all() {
        std::lock_guard<std::mutex> g(reg_mu());
        bool any = false;
        for (DevPool *p : registry()) { any = any || p->cached > 0; p->clear(); }
        return any;
    }
};

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0; // elements
    DevPool *pool = nullptr;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        const size_t want = n + n / 8 + 64;
        if (pool) {
            size_t got = 0;
            if (void *q = pool->take(want * sizeof(T), got)) { p = (T *)q; cap = got / sizeof(T); return cudaSuccess; }
        }
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e != cudaSuccess) { // give the caches of all contexts back to the driver and retry
            cudaGetLastError();
            if (DevPool::trim_all()) e = cudaMalloc((void **)&p, want * sizeof(T));
        }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() {
        if (p) { if (pool) pool->give(p, cap * sizeof(T)); else cudaFree(p); }
        p = nullptr; cap = 0;
    }
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

// A/a -> 0, C/c -> 1, G/g -> 2, T/t -> 3, anything else -> 0.  A table, not a switch: on random bases the switch
// mispredicted on most calls and the read encoder spent 8 of its 9 ms per 4 800 pairs there.
struct BaseCodeLut {
    uint8_t v[256];
    constexpr BaseCodeLut() : v() {
        for (int k = 0; k < 256; k++) v[k] = 0;
        v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3;
    }
};
static constexpr BaseCodeLut kBaseCode{};
inline uint8_t base_code(uint8_t c) { return kBaseCode.v[c]; }

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

// Host worker threads of one context, parked between calls (jtk_batch_create encodes every batch on them; spawning 2 x 15
// threads per batch cost ~1 ms of a 10 ms create).  run(n, grain, fn) executes fn(k) for k in [0, n) in dynamic chunks of
// `grain` items on the workers and the calling thread, and returns when all are done.
class HostPool {
  public:
    explicit HostPool(int n_threads) {
        for (int t = 1; t < n_threads; t++) workers_.emplace_back([this]() { loop(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size() + 1; }
    template <typename F> void run(int n, int grain, F fn) {
        if (n <= 0) return;
        if (workers_.empty() || n <= grain) { for (int k = 0; k < n; k++) fn(k); return; }
        std::function<void(int)> f = fn;
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &f; n_ = n; grain_ = grain; next_.store(0); active_ = (int)workers_.size(); gen_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this]() { return active_ == 0; });
        fn_ = nullptr;
    }

  private:
    void work() {
        for (;;) {
            const int k0 = next_.fetch_add(grain_);
            if (k0 >= n_) break;
            const int k1 = std::min(n_, k0 + grain_);
            for (int k = k0; k < k1; k++) (*fn_)(k);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&]() { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--active_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, grain_ = 1, active_ = 0;
    std::atomic<int> next_{0};
    uint64_t gen_ = 0;
    bool stop_ = false;
};

// Small host -> device parameter blocks (models, thresholds, offsets) are uploaded only when their bytes change: a
// cudaMemcpyAsync from pageable memory synchronises the stream first, so repeating it every call stops the host from
// queueing launches ahead of the GPU (each step then pays a host wake-up; 7.3 -> 11.3 ms per step on a busy host).
struct SmallUpload {
    std::vector<uint8_t> last;
    const void *dst = nullptr;
    cudaError_t put(void *d, const void *src, size_t bytes, cudaStream_t st) {
        if (d == dst && last.size() == bytes && std::memcmp(last.data(), src, bytes) == 0) return cudaSuccess;
        cudaError_t e = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) { last.assign((const uint8_t *)src, (const uint8_t *)src + bytes); dst = d; }
        else { last.clear(); dst = nullptr; }
        return e;
    }
};

struct jtk_ctx {
    int device = 0;
    int sm_count = 0;
    int last_variant = 0; // modification-table variant of the last table run: 1 fused (no DP matrix in HBM), 2 forward rows in HBM
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tm0 = nullptr, tm1 = nullptr;
    static constexpr int kRing = 256;
    cudaEvent_t ring0[kRing] = {}, ring1[kRing] = {};
    int ring_n = 0;
    std::string err;
    uint64_t launches = 0;
    float last_ms = 0.f;
    bool timing_pending = false;
    DevPool pool;             // device blocks recycled between batches
    int host_threads = 1;     // threads of the host-side encoders (JTK_HOST_THREADS, default: all cores, at most 32)
    std::unique_ptr<HostPool> hpool;
    SmallUpload up_models, up_minreq;
    // device buffers
    DevBuf<float> d_models;
    DevBuf<float2> d_frows;   // per-warp forward rows (scratch shared by all batches of this ctx)
    DevBuf<int32_t> d_kf;
    DevBuf<float4> d_raw;         // per pair slot: raw column sums (backward -> finalize kernel)
    DevBuf<unsigned> d_fwdinfo;   // per pair slot: end sums, total exponent, rescale-event map (forward -> backward kernel)
    size_t scratch_bytes = (size_t)16 << 30; // forward-row scratch budget of one wave (JTK_SCRATCH_MB)
    DevBuf<int> d_counter;
    DevBuf<float> d_minreq;
    DevBuf<uint32_t> d_cols;
    DevBuf<double> d_gather;
    DevBuf<double> d_tabs;        // p-value / prior / expected-gain tables of the candidate kernel
    DevBuf<uint32_t> d_tab_off;
    DevBuf<jtk_candidate> d_cand;
    DevBuf<uint32_t> d_pick;      // pick_probes_kernel: chunk offsets / copy numbers in, picks out
    DevBuf<double> d_var;         // pick_probes_kernel: the picked columns of every pair
    // k-means / MCMC restarts on the device (jtk_mcmc_restarts_batch)
    DevBuf<McmcChain> d_mc_chains;
    DevBuf<double> d_mc_f64, d_mc_lk;
    DevBuf<uint32_t> d_mc_u32;
    DevBuf<uint8_t> d_mc_u8, d_mc_asn;
    DevBuf<uint64_t> d_mc_rng, d_mc_asn_off;
    DevBuf<int> d_mc_err;
    // pinned host staging
    PinBuf<DevPair> h_pairs;
    PinBuf<uint8_t> h_codes;
    PinBuf<uint32_t> h_bits;
    PinBuf<float> h_delta;
    PinBuf<double> h_lk;
    PinBuf<uint8_t> h_homop;
    PinBuf<uint8_t> h_raw;                 // raw reads | raw ops of the batch being created (device encoder)
    DevBuf<uint8_t> d_rawin;
    DevBuf<uint32_t> d_raw_off;            // read_off | ops_off (| bootstrap: ops region offsets | ops_start | ops_len)
    DevBuf<uint32_t> d_boot;               // bootstrap_ops_kernel: 2 traceback bits per band cell
    DevBuf<int32_t> d_enc_status;          // EncStatus per pair | the cell count (8 bytes)
    PinBuf<uint32_t> h_raw_off;
    PinBuf<int32_t> h_enc_status;
    PinBuf<jtk_candidate> h_cand;
    PinBuf<double> h_gather;
    PinBuf<uint32_t> h_pick;
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

namespace jtk { int ctx_fail(jtk_ctx *ctx, int code, const char *msg) { return ctx ? ctx->fail(code, msg) : code; } }
#define CU(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx->cuda_fail(e_, what); } while (0)

extern "C" {

int jtk_hmm_num_row(void) { return JTK_NUM_ROW; }
int jtk_hmm_copy_size(void) { return JTK_COPY_SIZE; }
int jtk_hmm_del_size(void) { return JTK_DEL_SIZE; }

const char *jtk_last_error(const jtk_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
uint64_t jtk_ctx_launch_count(const jtk_ctx *ctx) { return ctx ? ctx->launches : 0; }
float jtk_ctx_last_kernel_ms(const jtk_ctx *ctx) { return ctx ? ctx->last_ms : 0.f; }
int jtk_ctx_last_modtable_variant(const jtk_ctx *ctx) { return ctx ? ctx->last_variant : 0; }

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
    {
        int nt = (int)std::thread::hardware_concurrency();
        if (const char *env = std::getenv("JTK_HOST_THREADS")) nt = std::atoi(env);
        ctx->host_threads = std::max(1, std::min(nt, 32));
        ctx->hpool.reset(new HostPool(ctx->host_threads));
        if (const char *env = std::getenv("JTK_SCRATCH_MB")) ctx->scratch_bytes = (size_t)std::max(64, std::atoi(env)) << 20;
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->tm0)) != cudaSuccess || (e = cudaEventCreate(&ctx->tm1)) != cudaSuccess) {
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
    ctx->d_models.release(); ctx->d_frows.release(); ctx->d_kf.release(); ctx->d_fwdinfo.release(); ctx->d_raw.release(); ctx->d_counter.release();
    ctx->d_minreq.release(); ctx->d_cols.release(); ctx->d_gather.release();
    ctx->h_raw.release(); ctx->h_raw_off.release(); ctx->h_enc_status.release(); ctx->d_rawin.release(); ctx->d_raw_off.release(); ctx->d_boot.release(); ctx->d_enc_status.release();
    ctx->d_mc_chains.release(); ctx->d_mc_f64.release(); ctx->d_mc_lk.release(); ctx->d_mc_u32.release(); ctx->d_mc_u8.release();
    ctx->d_mc_asn.release(); ctx->d_mc_rng.release(); ctx->d_mc_asn_off.release(); ctx->d_mc_err.release();
    ctx->d_tabs.release(); ctx->d_tab_off.release(); ctx->d_cand.release(); ctx->h_cand.release(); ctx->h_gather.release();
    ctx->h_pairs.release(); ctx->h_codes.release(); ctx->h_bits.release(); ctx->h_delta.release(); ctx->h_lk.release();
    ctx->h_homop.release();
    ctx->pool.clear();
    for (int k = 0; k < jtk_ctx::kRing; k++) { if (ctx->ring0[k]) cudaEventDestroy(ctx->ring0[k]); if (ctx->ring1[k]) cudaEventDestroy(ctx->ring1[k]); }
    if (ctx->tm0) cudaEventDestroy(ctx->tm0);
    if (ctx->tm1) cudaEventDestroy(ctx->tm1);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// global alignment of a read to a (window of a) consensus: the library's banded edit-distance aligner (the guide of the
// bootstrap likelihood) behind the call sites where the reference uses edlib in global mode
// (haplotyper/src/consensus/mod.rs:424-436 `global_align`; edge cases as there: empty query -> all Del, empty target -> all Ins)
int jtk_align_global(const uint8_t *tmpl, int Lt, const uint8_t *read, int Lr, int radius, uint8_t *out_ops, int cap) {
    if (Lt < 0 || Lr < 0 || radius < 0 || !out_ops || (Lt > 0 && !tmpl) || (Lr > 0 && !read)) return JTK_EINVAL;
    if (Lt + Lr > cap) return JTK_EINVAL;
    if (Lr == 0) { std::memset(out_ops, JTK_OP_DEL, (size_t)Lt); return Lt; }
    if (Lt == 0) { std::memset(out_ops, JTK_OP_INS, (size_t)Lr); return Lr; }
    std::vector<uint8_t> ops;
    std::vector<int> D;
    if (!edit_ops(tmpl, Lt, read, Lr, radius, ops, D)) return JTK_EINVAL;
    std::memcpy(out_ops, ops.data(), ops.size());
    return (int)ops.size();
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


// ---------------------------------------------------------------------------------------------------
// device-resident batch
// ---------------------------------------------------------------------------------------------------
struct jtk_batch {
    jtk_ctx *ctx = nullptr;
    int n_pairs = 0, n_tmpl = 0, radius = 0, C = 0, max_nd = 0, max_lt = 0, max_lr = 0;
    uint64_t table_floats = 0, cell_updates = 0, h2d_bytes = 0;
    bool has_profiles = false;
    bool has_order = false;               // ragged batch: the kernels take the pairs longest first (order follows d_pairs)
    std::vector<DevPair> pairs;           // host copy
    std::vector<uint32_t> tmpl_len;
    std::vector<uint32_t> tp_start, tp_ids; // CSR template -> pairs
    std::vector<uint32_t> homop_off;
    DevBuf<DevPair> d_pairs;
    DevBuf<uint8_t> d_codes;
    DevBuf<uint32_t> d_bits;
    DevBuf<float> d_delta;
    DevBuf<double> d_lk;
    DevBuf<uint32_t> d_tp_start, d_tp_ids, d_tmpl_len, d_homop_off;
    DevBuf<uint8_t> d_homop;
    DevBuf<uint32_t> d_tmpl_code_off; // codes[tmpl_code_off[t] + j + 1] = code of t[j]
    std::vector<uint32_t> tmpl_code_off;
    DevBuf<unsigned long long> d_stat_off;
    SmallUpload up_stat_off;
    DevBuf<jtk_colstat> d_stats;
    void attach(DevPool *pool) {
        d_pairs.pool = pool; d_codes.pool = pool; d_bits.pool = pool; d_delta.pool = pool; d_lk.pool = pool;
        d_tp_start.pool = pool; d_tp_ids.pool = pool; d_tmpl_len.pool = pool; d_homop_off.pool = pool; d_homop.pool = pool;
        d_stat_off.pool = pool; d_stats.pool = pool; d_tmpl_code_off.pool = pool;
    }
    void release() {
        d_tmpl_code_off.release();
        d_pairs.release(); d_codes.release(); d_bits.release(); d_delta.release(); d_lk.release();
        d_tp_start.release(); d_tp_ids.release(); d_tmpl_len.release(); d_homop_off.release(); d_homop.release();
        d_stat_off.release(); d_stats.release();
    }
};

namespace {

// DevPair elements that hold n uint32 queue-order entries behind the pairs
inline size_t order_pairs(size_t n) { return (n * sizeof(uint32_t) + sizeof(DevPair) - 1) / sizeof(DevPair); }

// JTK_TIMING=1: wall-clock phases of jtk_batch_create on stderr (tools/e2e_phases.py)
struct PhaseTimer {
    bool on; std::chrono::steady_clock::time_point t0; std::string out;
    PhaseTimer() : on(std::getenv("JTK_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        char buf[96];
        std::snprintf(buf, sizeof buf, " %s=%.3fms", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        out += buf; t0 = t1;
    }
    ~PhaseTimer() { if (on && !out.empty()) std::fprintf(stderr, "[jtk timing]%s\n", out.c_str()); }
};

// Run fn(k) for k in [0, n) on the context's host threads (dynamic chunks of `grain` items).
template <typename F> void parallel_for(int n_threads, int n, int grain, F fn) {
    if (n <= 0) return;
    const int nt = std::max(1, std::min(n_threads, (n + grain - 1) / grain));
    if (nt == 1) { for (int k = 0; k < n; k++) fn(k); return; }
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            const int k0 = next.fetch_add(grain);
            if (k0 >= n) break;
            const int k1 = std::min(n, k0 + grain);
            for (int k = k0; k < k1; k++) fn(k);
        }
    };
    std::vector<std::thread> th;
    th.reserve((size_t)nt - 1);
    for (int t = 1; t < nt; t++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
}

// ---- the pair encoder on the device ------------------------------------------------------------------------------
// What pack_batch's per-pair loop does on the host threads (read codes, guide bits, in-band cell count, validation of the
// ops), one warp per pair, from the caller's raw reads and ops: the host then only lays the arrays out.  Same bytes, same
// count, same first error as the host encoder.  The kernel takes 0.1 ms per 4 800 pairs, but behind another context's
// persistent table kernels it waits for milliseconds, so it is used only where the host is the short resource (a context
// with fewer than 6 encoder threads, e.g. 4 or 8 ranks on a 32-core box: bench.py --gpus 4 / 8), or under JTK_DEVICE_ENCODE=1;
// JTK_DEVICE_ENCODE=0 forces the host encoder, which the bootstrap path always uses.
struct EncStatus { int32_t bad, i, j, pad_; }; // bad: 0 ok, 1 the ops do not span (template, read), 2 invalid op code

__device__ __forceinline__ unsigned enc_base_code(unsigned c) { // A/a 0, C/c 1, G/g 2, T/t 3, anything else 0
    c &= 0xdfu; // upper case
    return c == 'C' ? 1u : (c == 'G' ? 2u : (c == 'T' ? 3u : 0u));
}


// ------------------------------------------------------------------------------------------------------------------
// Device twin of edit_ops(): the guide of the "bootstrap" likelihood (kiley likelihood_antidiagonal_bootstrap,
// likelihood_gains.rs:27-28,282-283,301-302 -- the 1.8e5 / 1e6 calibration pairs of ~100 bp), SURVEY.md 8f N3.
// One THREAD per pair: banded global edit distance (band |i-j| <= R = radius + |Lr-Lt|), row by row with two rolling rows in
// local memory; the traceback preference of the host (diagonal, then Del, then Ins) is decided when a cell is filled -- its
// three predecessors are final then -- and kept as 2 bits per band cell (16 cells per word); the traceback writes the ops
// backwards into the pair's region of `out_ops` and reports where they start.  Same ops as the host, byte for byte
// (tests/test_gpu_parity.py::test_device_bootstrap_equals_host_bootstrap).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBootMaxR = 127;
constexpr int kBootBand = 2 * kBootMaxR + 3;

__global__ void __launch_bounds__(128) bootstrap_ops_kernel(const DevPair *__restrict__ pairs, int n_pairs,
                                                            const uint8_t *__restrict__ raw_reads, const uint32_t *__restrict__ read_off,
                                                            const uint8_t *__restrict__ codes, int radius,
                                                            uint32_t *__restrict__ tb, const uint32_t *__restrict__ tb_off,
                                                            uint8_t *__restrict__ out_ops, const uint32_t *__restrict__ cap_off,
                                                            uint32_t *__restrict__ ops_start, uint32_t *__restrict__ ops_len) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const DevPair dp = pairs[p];
    const int Lt = dp.Lt, Lr = dp.Lr;
    const uint8_t *q = raw_reads + read_off[p];
    const uint8_t *tcode = codes + dp.tb_off; // tcode[j] = code of t[j-1]
    const int R = radius + abs(Lr - Lt), Wb = 2 * R + 1;
    const int INF = 1 << 29;
    int rowa[kBootBand], rowb[kBootBand];
    int *prev = rowa, *cur = rowb; // prev[b] = D[i-1][j] with b = j - (i-1) + R, one extra cell on both sides
    for (int b = 0; b < Wb + 2; b++) { prev[b] = INF; cur[b] = INF; }
    uint32_t *tbw = tb + tb_off[p];
    uint32_t word = 0u;
    size_t idx = 0;
    for (int i = 0; i <= Lr; i++) {
        const unsigned qc = i > 0 ? enc_base_code(q[i - 1]) : 0u;
        for (int b = 0; b < Wb; b++, idx++) {
            const int j = i - R + b;
            int v = INF;
            unsigned dir = 0u; // 1 diagonal, 2 Del (left), 3 Ins (up)
            if (j >= 0 && j <= Lt) {
                if (i == 0 && j == 0) v = 0;
                else {
                    // cells of the row above sit one band slot to the right: (i-1, j-1) -> b, (i-1, j) -> b+1 (index + 1 for the pad)
                    int vd = INF, vl = INF, vu = INF;
                    if (i > 0 && j > 0) vd = prev[b + 1] + (qc != (unsigned)tcode[j] ? 1 : 0);
                    if (j > 0) vl = cur[b] + 1;       // (i, j-1) -> b-1
                    if (i > 0) vu = prev[b + 2] + 1;  // (i-1, j) -> b+1
                    v = min(min(min(v, vd), vl), vu);
                    dir = (i > 0 && j > 0 && vd == v) ? 1u : ((j > 0 && vl == v) ? 2u : 3u);
                }
            }
            cur[b + 1] = v;
            word |= dir << (2 * (idx & 15));
            if ((idx & 15) == 15) { tbw[idx >> 4] = word; word = 0u; }
        }
        cur[0] = INF; cur[Wb + 1] = INF;
        int *t = prev; prev = cur; cur = t;
    }
    if (idx & 15) tbw[idx >> 4] = word;
    // traceback, written backwards from the end of the pair's region
    uint8_t *out = out_ops + cap_off[p];
    const int cap = Lr + Lt;
    int i = Lr, j = Lt, n = 0;
    while (i > 0 || j > 0) {
        const size_t c = (size_t)i * Wb + (size_t)(j - i + R);
        const unsigned dir = (tbw[c >> 4] >> (2 * (c & 15))) & 3u;
        uint8_t op;
        if (dir == 1u) { op = enc_base_code(q[i - 1]) != (unsigned)tcode[j] ? JTK_OP_MISMATCH : JTK_OP_MATCH; i--; j--; }
        else if (dir == 2u) { op = JTK_OP_DEL; j--; }
        else { op = JTK_OP_INS; i--; }
        out[cap - 1 - n] = op;
        n++;
    }
    ops_start[p] = cap_off[p] + (uint32_t)(cap - n);
    ops_len[p] = (uint32_t)n;
}

__global__ void __launch_bounds__(128) encode_pairs_kernel(const DevPair *__restrict__ pairs, int n_pairs,
                                                           const uint8_t *__restrict__ raw_reads, const uint32_t *__restrict__ read_off,
                                                           const uint8_t *__restrict__ raw_ops, const uint32_t *__restrict__ ops_off,
                                                           uint8_t *__restrict__ codes, uint32_t *__restrict__ bits, int radius,
                                                           int words_per_warp, EncStatus *__restrict__ status,
                                                           unsigned long long *__restrict__ cells,
                                                           const uint32_t *__restrict__ ops_start = nullptr,
                                                           const uint32_t *__restrict__ ops_len = nullptr) {
    extern __shared__ uint32_t enc_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + warp;
    if (p >= n_pairs) return;
    const DevPair dp = pairs[p];
    const int Lt = dp.Lt, Lr = dp.Lr;
    uint32_t *w = enc_sm + (size_t)warp * words_per_warp;
    const int nwords = (Lt + Lr + 1 + 31) / 32 + 1;
    for (int k = lane; k < nwords; k += 32) w[k] = 0u;
    // read rows: byte = ctx<<5 | qc<<2, 16 = no base; pads on both sides (four bytes per lane and store)
    {
        const uint8_t *q = raw_reads + read_off[p];
        const int rlen = (Lr + 2 + 3) & ~3;
        uint32_t *out = reinterpret_cast<uint32_t *>(codes + dp.rb_off - kCodePad);
        const int n4 = (2 * kCodePad + rlen) >> 2;
        for (int t = lane; t < n4; t += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int idx = 4 * t + b - kCodePad;
                unsigned byte = 16u;
                if (idx >= 1 && idx <= Lr) {
                    const unsigned qc = enc_base_code(q[idx - 1]);
                    const unsigned cx = idx >= 2 ? enc_base_code(q[idx - 2]) : 4u;
                    byte = (cx << 5) | (qc << 2);
                }
                v |= byte << (8 * b);
            }
            out[t] = v;
        }
    }
    __syncwarp();
    // guide ops of the caller, or the ones bootstrap_ops_kernel left (start / length per pair)
    const uint8_t *ops = raw_ops + (ops_start ? ops_start[p] : ops_off[p]);
    const int n_ops = ops_len ? (int)ops_len[p] : (int)(ops_off[p + 1] - ops_off[p]);
    const unsigned long long full = (unsigned long long)(2 * radius + 1);
    auto width = [&](int cen, int d) -> unsigned long long {
        const int lo = max(max(cen - radius, 0), d - Lt), hi = min(min(cen + radius, Lr), d);
        return hi >= lo ? (unsigned long long)(hi - lo + 1) : 0ull;
    };
    unsigned long long c = lane == 0 ? width(0, 0) : 0ull;
    int i = 0, j = 0, s = 0; // warp-uniform running coordinates
    int bad = 0, bi = 0, bj = 0;
    for (int base = 0; base < n_ops && !bad; base += 32) {
        const int k = base + lane;
        const bool act = k < n_ops;
        const unsigned op = act ? ops[k] : 0u;
        const bool invalid = act && op > JTK_OP_DEL;
        const int diag = (act && !invalid && op <= JTK_OP_MISMATCH) ? 1 : 0;
        const int ni = (act && !invalid && op != JTK_OP_DEL) ? 1 : 0, nj = (act && !invalid && op != JTK_OP_INS) ? 1 : 0;
        const int adv = (act && !invalid) ? 1 + diag : 0;
        int sa = adv, si = ni, sj = nj; // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(0xffffffffu, sa, o), ti = __shfl_up_sync(0xffffffffu, si, o), tj = __shfl_up_sync(0xffffffffu, sj, o);
            if (lane >= o) { sa += ta; si += ti; sj += tj; }
        }
        const int s_b = s + sa - adv, i_b = i + si - ni, j_b = j + sj - nj; // coordinates before this op
        const bool span = act && !invalid && (i_b + ni > Lr || j_b + nj > Lt);
        const unsigned fm = __ballot_sync(0xffffffffu, invalid || span);
        if (fm) { // the host encoder stops at the first offending op and reports what was consumed up to it
            const int first = __ffs((int)fm) - 1;
            bad = __shfl_sync(0xffffffffu, invalid ? 2 : 1, first);
            bi = __shfl_sync(0xffffffffu, i_b, first); bj = __shfl_sync(0xffffffffu, j_b, first);
            break;
        }
        if (act) {
            if (i_b > radius && i_b + 1 + radius < Lr && j_b > radius && j_b + 1 + radius < Lt) c += full * (unsigned long long)(1 + diag);
            else if (diag) c += width(i_b, s_b + 1) + width(i_b + 1, s_b + 2);
            else c += width(i_b + ni, s_b + 1);
            if (ni) { const int pos = s_b + diag; atomicOr(&w[pos >> 5], 1u << (pos & 31)); } // Match: bit s+1, Ins: bit s
        }
        s += __shfl_sync(0xffffffffu, sa, 31); i += __shfl_sync(0xffffffffu, si, 31); j += __shfl_sync(0xffffffffu, sj, 31);
    }
    if (!bad && (i != Lr || j != Lt)) { bad = 1; bi = i; bj = j; }
    __syncwarp();
    if (lane == 0) status[p] = EncStatus{ bad, bi, bj, 0 };
    if (bad) return;
    for (int k = lane; k < nwords; k += 32) bits[dp.bits_off + k] = w[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) atomicAdd(cells, 2ull * c);
}

// Encode templates / reads / guide paths into the device layout of phmm_dev.cuh (pinned staging of ctx).
// Pass 1 lays out every array (sequential, O(pairs)); pass 2 encodes templates and pairs on the host threads: every
// template / pair owns its own byte ranges.
int pack_batch(jtk_ctx *ctx, jtk_batch *b, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
               const uint8_t *read_concat, const uint32_t *read_off, const uint8_t *ops_concat,
               const uint32_t *ops_off, const uint8_t *strand, const uint32_t *tmpl_idx, size_t &code_bytes,
               size_t &bit_words, size_t &homop_bytes, bool device_encode, size_t &tmpl_code_bytes) {
    const int n_pairs = b->n_pairs, n_tmpl = b->n_tmpl, radius = b->radius;
    PhaseTimer pt;
    size_t cb = 0, hb = 0;
    b->tmpl_len.resize((size_t)n_tmpl);
    b->homop_off.resize((size_t)n_tmpl + 1);
    b->tmpl_code_off.resize((size_t)n_tmpl);
    for (int t = 0; t < n_tmpl; t++) {
        if (tmpl_off[t + 1] < tmpl_off[t]) return ctx->fail(JTK_EINVAL, "tmpl_off is not monotone");
        const size_t L = tmpl_off[t + 1] - tmpl_off[t];
        b->tmpl_len[t] = (uint32_t)L;
        b->homop_off[t] = (uint32_t)hb;
        b->tmpl_code_off[t] = (uint32_t)(cb + kCodePad);
        cb += 2 * (size_t)kCodePad + ((L + 1 + 3) & ~(size_t)3);
        hb += L + 1;
        b->max_lt = std::max(b->max_lt, (int)L);
    }
    b->homop_off[n_tmpl] = (uint32_t)hb;
    tmpl_code_bytes = cb; // the template codes come first; the read rows follow
    size_t bwords = 0;
    uint64_t tab = 0;
    std::vector<uint32_t> cnt((size_t)n_tmpl + 1, 0);
    b->pairs.resize((size_t)n_pairs);
    b->max_nd = 0;
    int min_nd = INT32_MAX;
    for (int p = 0; p < n_pairs; p++) {
        if (read_off[p + 1] < read_off[p]) return ctx->fail(JTK_EINVAL, "read_off is not monotone");
        if (tmpl_idx[p] >= (uint32_t)n_tmpl) return ctx->fail(JTK_EINVAL, "tmpl_idx out of range");
        if (ops_concat && ops_off[p + 1] < ops_off[p]) return ctx->fail(JTK_EINVAL, "ops_off is not monotone");
        const size_t Lr = read_off[p + 1] - read_off[p];
        const size_t Lt = b->tmpl_len[tmpl_idx[p]];
        DevPair &dp = b->pairs[p];
        dp.tb_off = b->tmpl_code_off[tmpl_idx[p]];
        dp.rb_off = (uint32_t)(cb + kCodePad);
        dp.bits_off = (uint32_t)bwords;
        dp.Lt = (int32_t)Lt; dp.Lr = (int32_t)Lr;
        dp.model = strand[p] ? 0 : 1;
        dp.pad_ = 0;
        dp.tab_off = tab;
        tab += (uint64_t)(Lt + 1) * kNumRow;
        cb += 2 * (size_t)kCodePad + ((Lr + 2 + 3) & ~(size_t)3);
        bwords += (Lt + Lr + 1 + 31) / 32 + 1;
        cnt[tmpl_idx[p] + 1]++;
        b->max_nd = std::max(b->max_nd, (int)(Lt + Lr + 1));
        min_nd = std::min(min_nd, (int)(Lt + Lr + 1));
        b->max_lr = std::max(b->max_lr, (int)Lr);
    }
    // length-sorted work queue: when the pairs of a batch differ in length by more than an eighth, the persistent warps pull
    // them longest first (no long pair starts when the others are finishing); uniform batches keep the batch order
    b->has_order = n_pairs > 1 && (size_t)(b->max_nd - min_nd) * 8 > (size_t)b->max_nd;
    if (cb >= (size_t)0xffffffffu) return ctx->fail(JTK_EINVAL, "batch too large: split it (code bytes exceed 4 GiB)");
    b->tp_start.assign((size_t)n_tmpl + 1, 0);
    for (int t = 0; t < n_tmpl; t++) b->tp_start[t + 1] = b->tp_start[t] + cnt[t + 1];
    b->tp_ids.resize((size_t)n_pairs);
    {
        std::vector<uint32_t> fill(b->tp_start.begin(), b->tp_start.end() - 1);
        for (int p = 0; p < n_pairs; p++) b->tp_ids[fill[tmpl_idx[p]]++] = (uint32_t)p;
    }
    CU(ctx->h_codes.reserve(cb), "cudaMallocHost codes");
    CU(ctx->h_bits.reserve(bwords), "cudaMallocHost bits");
    CU(ctx->h_pairs.reserve((size_t)n_pairs + order_pairs((size_t)n_pairs)), "cudaMallocHost pairs");
    CU(ctx->h_homop.reserve(hb), "cudaMallocHost homop");
    uint8_t *codes = ctx->h_codes.p;
    uint32_t *bits = ctx->h_bits.p;
    uint8_t *homop = ctx->h_homop.p;
    pt.mark("layout");

    ctx->hpool->run(n_tmpl, 4, [&](int t) {
        const uint8_t *s = tmpl_concat + tmpl_off[t];
        const size_t L = b->tmpl_len[t];
        uint8_t *c = codes + b->tmpl_code_off[t]; // c[j] = code of t[j-1]; 4 = none
        std::memset(c - kCodePad, 4, 2 * (size_t)kCodePad + ((L + 1 + 3) & ~(size_t)3));
        for (size_t j = 1; j <= L; j++) c[j] = base_code(s[j - 1]);
        // homopolymer run length of every base (pseudo_mcmc.rs:195-211), 1 past the end (:153)
        uint8_t *h = homop + b->homop_off[t];
        size_t k = 0;
        while (k < L) {
            size_t e = k;
            while (e < L && s[e] == s[k]) e++; // raw bytes, as the reference compares them
            for (size_t x = k; x < e; x++) h[x] = (uint8_t)std::min<size_t>(e - k, 255);
            k = e;
        }
        h[L] = 1;
    });

    pt.mark("templates");
    std::atomic<int> first_bad(n_pairs); // smallest failing pair so far (the error reported is the first in batch order)
    std::atomic<unsigned long long> cells_total(0);
    std::mutex bad_mu;
    std::string bad_text;
    int bad_code = 0;
    if (!device_encode) ctx->hpool->run(n_pairs, 8, [&](int p) {
        if (first_bad.load(std::memory_order_relaxed) < p) return;
        const DevPair &dp = b->pairs[p];
        const uint32_t ti = tmpl_idx[p];
        const int Lt = dp.Lt, Lr = dp.Lr;
        const uint8_t *q = read_concat + read_off[p];
        // read rows: byte = ctx<<5 | qc<<2 (the byte offset into the eI table; bits 2-4 index eM rows); 16 = no base
        const size_t rlen = ((size_t)Lr + 2 + 3) & ~(size_t)3;
        uint8_t *rc = codes + dp.rb_off;
        std::memset(rc - kCodePad, 16, 2 * (size_t)kCodePad + rlen);
        // row i: ctx = code of q[i-2] (4 = none for the first row), qc = code of q[i-1]; no loop-carried value
        if (Lr >= 1) rc[1] = (uint8_t)((4 << 5) | (base_code(q[0]) << 2));
        for (int i = 2; i <= Lr; i++) rc[i] = (uint8_t)((base_code(q[i - 2]) << 5) | (base_code(q[i - 1]) << 2));
        const uint8_t *ops;
        int n_ops;
        static thread_local std::vector<uint8_t> tl_ops;
        static thread_local std::vector<int> tl_D;
        auto fail = [&](int code, const std::string &msg) {
            std::lock_guard<std::mutex> lk(bad_mu);
            if (p < first_bad.load()) { first_bad.store(p); bad_text = msg; bad_code = code; }
        };
        if (ops_concat) {
            ops = ops_concat + ops_off[p];
            n_ops = (int)(ops_off[p + 1] - ops_off[p]);
        } else {
            const uint8_t *t = tmpl_concat + tmpl_off[ti];
            if (!edit_ops(t, Lt, q, Lr, radius, tl_ops, tl_D)) {
                fail(JTK_EINVAL, "bootstrap alignment: band cannot connect the corners (pair " + std::to_string(p) + ")");
                return;
            }
            ops = tl_ops.data();
            n_ops = (int)tl_ops.size();
        }
        // guide path -> centre increments, and the in-band cell count on the way.  The window of a diagonal is the full
        // 2r+1 cells unless the path cell is within r of a matrix edge; the guide bits are assembled in a register.
        const size_t nwords = (size_t)(Lt + Lr + 1 + 31) / 32 + 1;
        uint32_t *bw = bits + dp.bits_off;
        int i = 0, j = 0, s = 0;
        const uint64_t full = (uint64_t)(2 * radius + 1);
        auto width = [&](int cen, int d) -> uint64_t {
            const int jc = d - cen; // template coordinate of the centre cell
            if (cen > radius && cen + radius < Lr && jc > radius && jc + radius < Lt) return full;
            const int lo = std::max(std::max(cen - radius, 0), d - Lt), hi = std::min(std::min(cen + radius, Lr), d);
            return hi >= lo ? (uint64_t)(hi - lo + 1) : 0;
        };
        uint64_t c = width(0, 0);
        bool bad = false;
        // guide bits collect in a 64-bit window [base, base + 64) that is flushed a word at a time; away from the matrix
        // edges (the usual case) every diagonal holds the full window and an op costs a handful of branch-free operations
        uint64_t acc = 0;
        int base = 0;
        size_t wi = 0;
        const int i_lo = radius + 1, i_hi = Lr - radius - 2, j_lo = radius + 1, j_hi = Lt - radius - 2;
        for (int k = 0; k < n_ops; k++) {
            // eight diagonal ops (Match / Mismatch) in a row, all away from the matrix edges -- two thirds of all ops at 8 %
            // error: 16 anti-diagonals, guide bits 10101010 10101010 from s+1, no per-op work
            if (k + 8 <= n_ops && i >= i_lo && i + 7 <= i_hi && j >= j_lo && j + 7 <= j_hi) {
                uint64_t w8;
                std::memcpy(&w8, ops + k, 8);
                if ((w8 & 0xfefefefefefefefeULL) == 0) {
                    acc |= 0xAAAAULL << (s - base);
                    c += full * 16;
                    i += 8; j += 8; s += 16; k += 7;
                    if (s - base >= 32) { bw[wi++] = (uint32_t)acc; acc >>= 32; base += 32; }
                    continue;
                }
            }
            const uint8_t op = ops[k];
            if (op > JTK_OP_DEL) { fail(JTK_EINVAL, "invalid op code in pair " + std::to_string(p)); return; }
            const int diag = op <= JTK_OP_MISMATCH, ni = op != JTK_OP_DEL, nj = op != JTK_OP_INS;
            if (i + ni > Lr || j + nj > Lt) { bad = true; break; }
            if ((i >= i_lo) & (i <= i_hi) & (j >= j_lo) & (j <= j_hi)) c += full * (uint64_t)(1 + diag);
            else if (diag) c += width(i, s + 1) + width(i + 1, s + 2);
            else c += width(i + ni, s + 1);
            acc |= (uint64_t)ni << (s + diag - base); // Match: bit s+1, Ins: bit s, Del: none
            i += ni; j += nj; s += 1 + diag;
            if (s - base >= 32) { bw[wi++] = (uint32_t)acc; acc >>= 32; base += 32; }
        }
        if (wi < nwords) bw[wi++] = (uint32_t)acc;
        if (wi < nwords) bw[wi++] = (uint32_t)(acc >> 32);
        for (size_t k = wi; k < nwords; k++) bw[k] = 0;
        if (bad || i != Lr || j != Lt) {
            fail(JTK_EINVAL, "ops of pair " + std::to_string(p) + " do not span (template, read): consumed (" + std::to_string(j) +
                                 "," + std::to_string(i) + ") of (" + std::to_string(Lt) + "," + std::to_string(Lr) + ")");
            return;
        }
        cells_total.fetch_add(2 * c, std::memory_order_relaxed);
    });
    pt.mark("pairs");
    if (first_bad.load() < n_pairs) return ctx->fail(bad_code ? bad_code : JTK_EINVAL, bad_text);
    std::memcpy(ctx->h_pairs.p, b->pairs.data(), sizeof(DevPair) * (size_t)n_pairs);
    if (b->has_order) {
        uint32_t *ord = reinterpret_cast<uint32_t *>(ctx->h_pairs.p + n_pairs);
        for (int p = 0; p < n_pairs; p++) ord[p] = (uint32_t)p;
        const DevPair *pp = b->pairs.data();
        std::stable_sort(ord, ord + n_pairs, [pp](uint32_t x, uint32_t y) { return pp[x].Lt + pp[x].Lr > pp[y].Lt + pp[y].Lr; });
    }
    ctx->tmpl_code_off = b->tmpl_code_off;
    b->table_floats = tab;
    b->cell_updates = cells_total.load();
    b->h2d_bytes = (device_encode ? tmpl_code_bytes + (size_t)read_off[n_pairs] + (ops_off ? (size_t)ops_off[n_pairs] : 0) + 8 * ((size_t)n_pairs + 1)
                                  : cb + bwords * sizeof(uint32_t)) +
                   hb + sizeof(DevPair) * (size_t)n_pairs + sizeof(uint32_t) * (4 * (size_t)n_tmpl + 3 + (size_t)n_pairs);
    code_bytes = cb; bit_words = bwords; homop_bytes = hb;
    return JTK_OK;
}

int batch_create(jtk_ctx *ctx, int n_pairs, int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                 const uint8_t *read_concat, const uint32_t *read_off, const uint8_t *ops_concat, const uint32_t *ops_off,
                 const uint8_t *strand, const uint32_t *tmpl_idx, int radius, bool allow_bootstrap, jtk_batch **out) {
    if (!ctx) return JTK_EINVAL;
    if (!out) return ctx->fail(JTK_EINVAL, "out is NULL");
    *out = nullptr;
    if (n_pairs < 0 || n_tmpl < 0) return ctx->fail(JTK_EINVAL, "negative count");
    if (n_pairs > 0 && (!tmpl_concat || !tmpl_off || !read_concat || !read_off || !strand || !tmpl_idx))
        return ctx->fail(JTK_EINVAL, "null argument");
    if (n_pairs > 0 && !allow_bootstrap && (!ops_concat || !ops_off)) return ctx->fail(JTK_EINVAL, "guide ops are required");
    const int C = cols_per_lane_for_radius(radius);
    if (radius < 0 || C == 0 || C > 8) return ctx->fail(JTK_EINVAL, "radius out of range (0..126)");
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    jtk_batch *b = new jtk_batch();
    b->ctx = ctx; b->n_pairs = n_pairs; b->n_tmpl = n_tmpl; b->radius = radius; b->C = C;
    b->attach(&ctx->pool);
    if (n_pairs == 0) { *out = b; return JTK_OK; }
    size_t code_bytes = 0, bit_words = 0, homop_bytes = 0, tmpl_code_bytes = 0;
    // Guide ops given and few encoder threads in this context (or JTK_DEVICE_ENCODE=1): the pairs are encoded on the device
    // from the raw reads and ops; otherwise on the host threads (see encode_pairs_kernel)
    bool device_encode = ops_concat != nullptr && ctx->host_threads < 6;
    if (const char *env = std::getenv("JTK_DEVICE_ENCODE")) device_encode = ops_concat != nullptr && std::atoi(env) != 0;
    // No guide ops (bootstrap likelihood, the calibration pairs of likelihood_gains.rs): the banded edit-distance alignment runs
    // on the device too (bootstrap_ops_kernel, one thread per pair) when every pair fits its limits, and feeds the device
    // encoder; JTK_DEVICE_BOOTSTRAP=0 keeps the alignment on the host threads.
    bool device_bootstrap = false;
    std::vector<uint32_t> boot_off; // per pair: first traceback word, then (second half) first byte of its ops region
    size_t boot_words = 0, boot_ops = 0;
    if (!ops_concat && n_pairs > 0) {
        device_bootstrap = true;
        if (const char *env = std::getenv("JTK_DEVICE_BOOTSTRAP")) device_bootstrap = std::atoi(env) != 0;
        boot_off.resize(2 * ((size_t)n_pairs + 1));
        for (int p = 0; p < n_pairs && device_bootstrap; p++) {
            if (read_off[p + 1] < read_off[p] || tmpl_idx[p] >= (uint32_t)n_tmpl) { device_bootstrap = false; break; } // pack_batch reports it
            const long Lr = (long)(read_off[p + 1] - read_off[p]), Lt = (long)(tmpl_off[tmpl_idx[p] + 1] - tmpl_off[tmpl_idx[p]]);
            const long R = radius + std::labs(Lr - Lt);
            if (R > kBootMaxR || Lr + Lt > 16384) { device_bootstrap = false; break; }
            boot_off[p] = (uint32_t)boot_words;
            boot_off[(size_t)n_pairs + 1 + p] = (uint32_t)boot_ops;
            boot_words += ((size_t)(Lr + 1) * (size_t)(2 * R + 1) + 15) / 16;
            boot_ops += (size_t)(Lr + Lt);
            if (boot_words > ((size_t)1 << 29) || boot_ops > ((size_t)1 << 31)) device_bootstrap = false; // 2 GiB of traceback bits
        }
        if (device_bootstrap) {
            boot_off[n_pairs] = (uint32_t)boot_words;
            boot_off[2 * (size_t)n_pairs + 1] = (uint32_t)boot_ops;
            device_encode = true;
        }
    }
    int rc = pack_batch(ctx, b, tmpl_concat, tmpl_off, read_concat, read_off, ops_concat, ops_off, strand, tmpl_idx,
                        code_bytes, bit_words, homop_bytes, device_encode, tmpl_code_bytes);
    if (rc) { b->release(); delete b; return rc; }
    PhaseTimer pt;
    cudaStream_t st = ctx->stream;
    cudaError_t e;
#define CB(call, what) do { if ((e = (call)) != cudaSuccess) { b->release(); delete b; return ctx->cuda_fail(e, what); } } while (0)
    const size_t pair_elems = (size_t)n_pairs + (b->has_order ? order_pairs((size_t)n_pairs) : 0);
    CB(b->d_pairs.reserve(pair_elems), "cudaMalloc pairs");
    CB(b->d_codes.reserve(code_bytes), "cudaMalloc codes");
    CB(b->d_bits.reserve(bit_words), "cudaMalloc bits");
    CB(b->d_lk.reserve((size_t)n_pairs), "cudaMalloc lk");
    CB(b->d_tp_start.reserve(b->tp_start.size()), "cudaMalloc csr");
    CB(b->d_tp_ids.reserve(b->tp_ids.size()), "cudaMalloc csr");
    CB(b->d_tmpl_len.reserve(b->tmpl_len.size()), "cudaMalloc tmpl_len");
    CB(b->d_homop_off.reserve(b->homop_off.size()), "cudaMalloc homop_off");
    CB(b->d_homop.reserve(homop_bytes), "cudaMalloc homop");
    CB(b->d_tmpl_code_off.reserve(b->tmpl_code_off.size()), "cudaMalloc tmpl_code_off");
    pt.mark("reserve");
    CB(cudaMemcpyAsync(b->d_pairs.p, ctx->h_pairs.p, sizeof(DevPair) * pair_elems, cudaMemcpyHostToDevice, st), "H2D pairs");
    if (!device_encode) {
        CB(cudaMemcpyAsync(b->d_codes.p, ctx->h_codes.p, code_bytes, cudaMemcpyHostToDevice, st), "H2D codes");
        CB(cudaMemcpyAsync(b->d_bits.p, ctx->h_bits.p, bit_words * sizeof(uint32_t), cudaMemcpyHostToDevice, st), "H2D bits");
    } else {
        const size_t rb = read_off[n_pairs], ob = device_bootstrap ? boot_ops : ops_off[n_pairs];
        const size_t rb_pad = (rb + 255) & ~(size_t)255;
        CB(ctx->h_raw.reserve(rb_pad + (device_bootstrap ? 0 : ob) + 64), "cudaMallocHost raw reads / ops");
        CB(ctx->d_rawin.reserve(rb_pad + ob + 64), "cudaMalloc raw reads / ops");
        CB(ctx->d_raw_off.reserve((device_bootstrap ? 5 : 2) * ((size_t)n_pairs + 1)), "cudaMalloc raw offsets");
        CB(ctx->d_enc_status.reserve((size_t)4 * n_pairs + 4), "cudaMalloc encoder status");
        { // the caller's buffers -> pinned staging, on the host threads (1 MiB pieces)
            const size_t piece = (size_t)1 << 20, nr = (rb + piece - 1) / piece, no = device_bootstrap ? 0 : (ob + piece - 1) / piece;
            uint8_t *hr = ctx->h_raw.p;
            ctx->hpool->run((int)(nr + no), 1, [&](int k) {
                if ((size_t)k < nr) { const size_t o = (size_t)k * piece; std::memcpy(hr + o, read_concat + o, std::min(piece, rb - o)); }
                else { const size_t o = ((size_t)k - nr) * piece; std::memcpy(hr + rb_pad + o, ops_concat + o, std::min(piece, ob - o)); }
            });
        }
        unsigned long long *d_cells = reinterpret_cast<unsigned long long *>(ctx->d_enc_status.p + (size_t)4 * n_pairs);
        CB(cudaMemcpyAsync(b->d_codes.p, ctx->h_codes.p, tmpl_code_bytes, cudaMemcpyHostToDevice, st), "H2D template codes");
        CB(cudaMemcpyAsync(ctx->d_rawin.p, ctx->h_raw.p, rb_pad + (device_bootstrap ? 0 : ob), cudaMemcpyHostToDevice, st), "H2D raw reads / ops");
        // offsets through pinned staging too: a copy from pageable memory synchronises the stream before it starts
        // layout: read_off | ops_off (or the bootstrap's traceback offsets) | ops region offsets | ops_start | ops_len
        const size_t n1 = (size_t)n_pairs + 1;
        CB(ctx->h_raw_off.reserve(3 * n1), "cudaMallocHost raw offsets");
        std::memcpy(ctx->h_raw_off.p, read_off, sizeof(uint32_t) * n1);
        if (device_bootstrap) std::memcpy(ctx->h_raw_off.p + n1, boot_off.data(), sizeof(uint32_t) * 2 * n1);
        else std::memcpy(ctx->h_raw_off.p + n1, ops_off, sizeof(uint32_t) * n1);
        CB(cudaMemcpyAsync(ctx->d_raw_off.p, ctx->h_raw_off.p, sizeof(uint32_t) * (device_bootstrap ? 3 : 2) * n1, cudaMemcpyHostToDevice, st), "H2D raw offsets");
        CB(cudaMemsetAsync(d_cells, 0, sizeof(unsigned long long), st), "memset cell count");
        const uint32_t *d_ops_start = nullptr, *d_ops_len = nullptr;
        if (device_bootstrap) {
            CB(ctx->d_boot.reserve(boot_words + 1), "cudaMalloc traceback bits");
            uint32_t *start = ctx->d_raw_off.p + 3 * n1, *len = ctx->d_raw_off.p + 4 * n1;
            bootstrap_ops_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(
                b->d_pairs.p, n_pairs, ctx->d_rawin.p, ctx->d_raw_off.p, b->d_codes.p, b->radius, ctx->d_boot.p, ctx->d_raw_off.p + n1,
                ctx->d_rawin.p + rb_pad, ctx->d_raw_off.p + 2 * n1, start, len);
            CB(cudaGetLastError(), "bootstrap launch");
            ctx->launches++;
            d_ops_start = start; d_ops_len = len;
        }
        const int words = ((b->max_nd + 31) / 32 + 2 + 3) & ~3, warps = 4;
        const size_t dyn = (size_t)warps * words * sizeof(uint32_t);
        if (dyn > 48 * 1024) CB(cudaFuncSetAttribute(encode_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn), "encoder shared memory");
        encode_pairs_kernel<<<(n_pairs + warps - 1) / warps, warps * 32, dyn, st>>>(
            b->d_pairs.p, n_pairs, ctx->d_rawin.p, ctx->d_raw_off.p, ctx->d_rawin.p + rb_pad, ctx->d_raw_off.p + n_pairs + 1, b->d_codes.p,
            b->d_bits.p, b->radius, words, reinterpret_cast<EncStatus *>(ctx->d_enc_status.p), d_cells, d_ops_start, d_ops_len);
        CB(cudaGetLastError(), "encoder launch");
        ctx->launches++;
    }
    CB(cudaMemcpyAsync(b->d_homop.p, ctx->h_homop.p, homop_bytes, cudaMemcpyHostToDevice, st), "H2D homop");
    CB(cudaMemcpyAsync(b->d_tp_start.p, b->tp_start.data(), sizeof(uint32_t) * b->tp_start.size(), cudaMemcpyHostToDevice, st), "H2D csr");
    CB(cudaMemcpyAsync(b->d_tp_ids.p, b->tp_ids.data(), sizeof(uint32_t) * b->tp_ids.size(), cudaMemcpyHostToDevice, st), "H2D csr");
    CB(cudaMemcpyAsync(b->d_tmpl_len.p, b->tmpl_len.data(), sizeof(uint32_t) * b->tmpl_len.size(), cudaMemcpyHostToDevice, st), "H2D tmpl_len");
    CB(cudaMemcpyAsync(b->d_homop_off.p, b->homop_off.data(), sizeof(uint32_t) * b->homop_off.size(), cudaMemcpyHostToDevice, st), "H2D homop_off");
    CB(cudaMemcpyAsync(b->d_tmpl_code_off.p, b->tmpl_code_off.data(), sizeof(uint32_t) * b->tmpl_code_off.size(), cudaMemcpyHostToDevice, st), "H2D tmpl_code_off");
    pt.mark("h2d_issue");
    if (device_encode) {
        CB(ctx->h_enc_status.reserve((size_t)4 * n_pairs + 4), "cudaMallocHost encoder status");
        CB(cudaMemcpyAsync(ctx->h_enc_status.p, ctx->d_enc_status.p, sizeof(int32_t) * ((size_t)4 * n_pairs + 4), cudaMemcpyDeviceToHost, st), "D2H encoder status");
    }
    CB(cudaStreamSynchronize(st), "upload");
    pt.mark("h2d_sync");
#undef CB
    if (device_encode) {
        const EncStatus *es = reinterpret_cast<const EncStatus *>(ctx->h_enc_status.p);
        for (int p = 0; p < n_pairs; p++) {
            if (es[p].bad == 0) continue;
            const DevPair &dp = b->pairs[(size_t)p];
            std::string msg = es[p].bad == 2 ? "invalid op code in pair " + std::to_string(p)
                                             : "ops of pair " + std::to_string(p) + " do not span (template, read): consumed (" +
                                                   std::to_string(es[p].j) + "," + std::to_string(es[p].i) + ") of (" + std::to_string(dp.Lt) +
                                                   "," + std::to_string(dp.Lr) + ")";
            b->release(); delete b;
            return ctx->fail(JTK_EINVAL, msg);
        }
        unsigned long long cells = 0;
        std::memcpy(&cells, ctx->h_enc_status.p + (size_t)4 * n_pairs, sizeof cells);
        b->cell_updates = cells;
    }
    *out = b;
    return JTK_OK;
}

int batch_run(jtk_batch *b, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, bool table, int rows) {
    jtk_ctx *ctx = b->ctx;
    if (!fwd || !rev) return ctx->fail(JTK_EINVAL, "null model");
    if (table && rows != 14 && rows != 9) return ctx->fail(JTK_EINVAL, "rows must be 14 or 9");
    if (b->n_pairs == 0) return JTK_OK;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    float models[2 * kModelFloats];
    pack_model(fwd, models);
    pack_model(rev, models + kModelFloats);
    const int wpc = warps_per_cta();
    cudaStream_t st = ctx->stream;
    KParams kp{};
    CU(ctx->d_models.reserve(2 * kModelFloats), "cudaMalloc models");
    CU(ctx->d_counter.reserve(2), "cudaMalloc counter");
    // The modification table runs as waves of (forward kernel, backward kernel); the forward rows of every pair of a wave
    // live in HBM between the two kernels (2.3 MB per 2 kbp pair), so a wave holds as many pairs as the scratch budget allows.
    int per_wave = b->n_pairs;
    // Two bit-identical variants of the modification table:
    //   fused  one kernel, forward rows recomputed per segment in shared memory: the DP matrices never touch HBM, scratch is
    //          ~260 KB per resident WARP (checkpoints);
    //   rows   forward kernel + backward kernel, forward rows parked in HBM in between: 2.3 MB of scratch per 2 kbp PAIR of a
    //          wave and 40x the algorithmic HBM traffic, but ~25 % fewer instructions -- faster while HBM is not the bound.
    // JTK_MODTABLE=fused|rows picks one; otherwise rows when the whole batch fits the scratch budget as ONE wave, else fused.
    bool legacy = false;
    if (table) {
        const char *env = std::getenv("JTK_MODTABLE");
        if (std::getenv("JTK_MODTABLE_LEGACY") || (env && std::strcmp(env, "rows") == 0)) legacy = true;
        else if (env && std::strcmp(env, "fused") == 0) legacy = false;
        else {
            const size_t per_pair = (size_t)(b->max_nd + frow_extra_rows()) * frow_slots_per_row(b->C) * sizeof(float2) +
                                    ((size_t)b->max_nd + 6) * sizeof(int32_t) + (size_t)4 * (b->max_lt + 1) * sizeof(float4);
            const size_t need = (size_t)b->n_pairs * per_pair;
            legacy = need <= ctx->scratch_bytes;
            if (legacy && need > ctx->d_frows.cap * sizeof(float2)) { // the scratch would have to grow: does the device have it?
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                legacy = need <= (free_b + ctx->d_frows.cap * sizeof(float2)) / 2;
            }
        }
        ctx->last_variant = legacy ? 2 : 1;
    }
    int fused_blocks = 0;
    if (table && !legacy) {
        // v10: one fused kernel, forward rows recomputed in shared memory.  Per WARP SLOT: checkpoints + block exponents; per
        // pair of a wave: raw column sums (64 B per template column) + forward info for finalize_kernel.
        kp.smem_rb = ((b->max_lr + 2 * fwd_pad_rows(b->C)) + 15) & ~15;
        kp.smem_tb = 0;
        kp.fwdinfo_stride = (size_t)fwdinfo_words();
        kp.raw_stride = (size_t)4 * (b->max_lt + 1);
        kp.max_lt = b->max_lt;
        const int dyn = fused_dyn_smem(b->C, kp.smem_rb);
        if ((size_t)dyn > (size_t)227 * 1024 - 4096)
            return ctx->fail(JTK_EINVAL, "pair too long for the modification-table kernel: the read must stay below ~50 000 bases");
        const int per_sm = fused_grid(b->C, rows, dyn, 1);
        if (per_sm <= 0) return ctx->fail(JTK_ECUDA, "modification-table kernel does not fit an SM");
        const size_t per_pair = kp.fwdinfo_stride * 4 + kp.raw_stride * sizeof(float4);
        per_wave = (int)std::max<size_t>(1, std::min<size_t>((size_t)b->n_pairs, ctx->scratch_bytes / per_pair));
        per_wave = std::min(per_wave, 65535); // finalize_kernel puts the pair of a wave on grid.y
        per_wave = (int)std::max<size_t>(1, std::min<size_t>((size_t)per_wave, (size_t)0xffffffffu / kp.raw_stride - 1)); // 32-bit raw offsets
        fused_blocks = std::min((per_wave + wpc - 1) / wpc, ctx->sm_count * per_sm);
        kp.frow_stride = (fused_ckpt_bytes(b->C, b->max_nd) + 15) & ~(size_t)15;                  // BYTES per warp slot
        kp.kf_stride = (fused_kb_ints(b->max_nd) + 3) & ~(size_t)3;
        const size_t slots = (size_t)fused_blocks * wpc;
        CU(ctx->d_frows.reserve(slots * kp.frow_stride / sizeof(float2)), "cudaMalloc checkpoints");
        CU(ctx->d_kf.reserve(slots * kp.kf_stride), "cudaMalloc scale exponents");
        CU(ctx->d_fwdinfo.reserve((size_t)per_wave * kp.fwdinfo_stride), "cudaMalloc forward info");
        CU(ctx->d_raw.reserve((size_t)per_wave * kp.raw_stride), "cudaMalloc raw column sums");
        CU(b->d_delta.reserve((size_t)b->table_floats), "cudaMalloc profiles");
    } else if (table) {
        kp.frow_stride = (size_t)(b->max_nd + frow_extra_rows()) * frow_slots_per_row(b->C);
        kp.kf_stride = (size_t)b->max_nd + 6;
        kp.fwdinfo_stride = (size_t)fwdinfo_words();
        kp.smem_rb = ((b->max_lr + 2 * fwd_pad_rows(b->C)) + 15) & ~15;
        kp.smem_tb = 0;
        kp.raw_stride = (size_t)4 * (b->max_lt + 1);
        kp.max_lt = b->max_lt;
        const size_t per_pair = kp.frow_stride * sizeof(float2) + kp.kf_stride * sizeof(int32_t) + kp.fwdinfo_stride * 4 +
                                kp.raw_stride * sizeof(float4);
        const size_t have = ctx->d_frows.cap * sizeof(float2);
        size_t budget = ctx->scratch_bytes;
        if ((size_t)b->n_pairs * per_pair > have) { // the scratch has to grow: see what the device can give
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            budget = std::min(budget, (free_b + have) / 2);
        }
        per_wave = (int)std::max<size_t>(1, std::min<size_t>((size_t)b->n_pairs, budget / per_pair));
        per_wave = std::min(per_wave, 65535); // finalize_kernel puts the pair of a wave on grid.y
        // both kernels stage the read codes of one pair per warp in shared memory (the backward kernel next to its ring)
        if ((size_t)modtable_dyn_smem(b->C, kp.smem_rb) > (size_t)227 * 1024 - 6144)
            return ctx->fail(JTK_EINVAL, "pair too long for the modification-table kernels: the read must stay below ~40 000 bases");
        CU(ctx->d_frows.reserve((size_t)per_wave * kp.frow_stride), "cudaMalloc forward rows");
        CU(ctx->d_kf.reserve((size_t)per_wave * kp.kf_stride), "cudaMalloc scale exponents");
        CU(ctx->d_fwdinfo.reserve((size_t)per_wave * kp.fwdinfo_stride), "cudaMalloc forward info");
        CU(ctx->d_raw.reserve((size_t)per_wave * kp.raw_stride), "cudaMalloc raw column sums");
        CU(b->d_delta.reserve((size_t)b->table_floats), "cudaMalloc profiles");
    }
    CU(ctx->up_models.put(ctx->d_models.p, models, sizeof(models), st), "H2D models");
    kp.pairs = b->d_pairs.p; kp.n_pairs = b->n_pairs;
    kp.codes = b->d_codes.p; kp.bits = b->d_bits.p; kp.models = ctx->d_models.p;
    kp.radius = b->radius; kp.rows = rows;
    kp.frows = ctx->d_frows.p; kp.kf = ctx->d_kf.p; kp.fwdinfo = ctx->d_fwdinfo.p; kp.raw = ctx->d_raw.p;
    kp.out_delta = b->d_delta.p; kp.out_lk = b->d_lk.p; kp.counter = ctx->d_counter.p; kp.counter2 = ctx->d_counter.p + 1;
    kp.order = b->has_order ? reinterpret_cast<const uint32_t *>(b->d_pairs.p + b->n_pairs) : nullptr;
    cudaEvent_t r0 = nullptr, r1 = nullptr;
    if (ctx->ring_n < jtk_ctx::kRing) {
        const int k = ctx->ring_n;
        if (!ctx->ring0[k]) { CU(cudaEventCreate(&ctx->ring0[k]), "event"); CU(cudaEventCreate(&ctx->ring1[k]), "event"); }
        r0 = ctx->ring0[k]; r1 = ctx->ring1[k];
        ctx->ring_n++;
    }
    CU(cudaEventRecord(ctx->ev0, st), "event");
    if (r0) CU(cudaEventRecord(r0, st), "event");
    for (int lo = 0; lo < b->n_pairs; lo += per_wave) {
        const int hi = std::min(b->n_pairs, lo + per_wave);
        const int ctas = (hi - lo + wpc - 1) / wpc;
        CU(cudaMemsetAsync(ctx->d_counter.p, 0, 2 * sizeof(int), st), "memset counter");
        kp.pair_lo = lo; kp.pair_hi = hi;
        if (table && !legacy) {
            CU(launch_modtable_fused(kp, b->C, std::min(ctas, fused_blocks), st), "kernel launch");
            ctx->launches += 2;
        } else if (table) {
            // persistent CTAs: one wave of resident CTAs pulls pairs from the queue
            int cap_f = fwdrows_ctas_per_sm(b->C), cap_b = modtable_ctas_per_sm(b->C, rows);
            if (const char *v = std::getenv("JTK_GRID_FWD")) cap_f = std::max(1, std::min(cap_f, std::atoi(v))); // (tuning: CTAs per SM)
            if (const char *v = std::getenv("JTK_GRID_BWD")) cap_b = std::max(1, std::min(cap_b, std::atoi(v)));
            const int gf = std::min(ctas, ctx->sm_count * cap_f);
            const int wpb = modtable_warps_per_cta(b->C);
            const int gb = std::min((hi - lo + wpb - 1) / wpb, ctx->sm_count * cap_b);
            CU(launch_modtable(kp, b->C, gf, gb, st), "kernel launch");
            ctx->launches += 3;
        } else {
            kp.n_pairs = b->n_pairs;
            // at a small radius two pairs share a warp (calibration batches: 1e5..1e6 pairs of ~100 bp)
            kp.smem_rb = ((b->max_lr + 2 * likelihood_pairs2_pad_rows()) + 15) & ~15;
            const bool packed = b->radius <= likelihood_pairs2_max_radius() && (size_t)wpc * 2 * kp.smem_rb <= (size_t)200 * 1024 &&
                                !std::getenv("JTK_LIKELIHOOD_UNPACKED");
            if (packed) {
                const int ctas2 = ((hi - lo + 1) / 2 + wpc - 1) / wpc;
                CU(launch_likelihood_pairs2(kp, std::min(ctas2, ctx->sm_count * 8), st), "kernel launch");
            } else {
                CU(launch_likelihood(kp, b->C, std::min(ctas, ctx->sm_count * 4), st), "kernel launch");
            }
            ctx->launches++;
        }
    }
    if (r1) CU(cudaEventRecord(r1, st), "event");
    CU(cudaEventRecord(ctx->ev1, st), "event");
    ctx->timing_pending = true;
    if (table) b->has_profiles = true;
    return JTK_OK;
}

int batch_sync(jtk_batch *b) {
    jtk_ctx *ctx = b->ctx;
    CU(cudaStreamSynchronize(ctx->stream), "kernel execution");
    if (ctx->timing_pending) { cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1); ctx->timing_pending = false; }
    return JTK_OK;
}

// --- column statistics over the reads of each chunk (compress_small_gains + column_sum + strand counts) ---
__device__ __forceinline__ float min_req_of(const float *min_req, int H, int row, int hl) {
    const int type = row < 4 ? 0 : (row < 8 + JTK_COPY_SIZE ? 2 : 1); // Subst, Ins (incl. copy), Del
    const int h = min(max(hl, 1), H);
    return min_req[type * H + h - 1];
}

// The reads of one chunk as (table offset, strand model), staged in shared memory by the whole CTA: the per-entry loops
// below then issue independent, coalesced loads instead of a tp_ids -> pairs -> delta chain per read.
constexpr int kReadTile = 256;
struct ReadRef { unsigned long long tab_off; int model; int pad_; };
__device__ __forceinline__ int stage_reads(ReadRef *sh, const DevPair *__restrict__ pairs, const uint32_t *__restrict__ tp_ids,
                                           uint32_t k0, uint32_t k1) {
    const int n = (int)min(k1 - k0, (uint32_t)kReadTile);
    __syncthreads(); // the previous tile is no longer read
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const DevPair &p = pairs[tp_ids[k0 + k]];
        sh[k].tab_off = p.tab_off; sh[k].model = p.model;
    }
    __syncthreads();
    return n;
}

__global__ void colstats_kernel(const float *__restrict__ delta, const DevPair *__restrict__ pairs,
                                const uint32_t *__restrict__ tp_start, const uint32_t *__restrict__ tp_ids,
                                const uint32_t *__restrict__ tmpl_len, const uint8_t *__restrict__ homop,
                                const uint32_t *__restrict__ homop_off, const unsigned long long *__restrict__ stat_off,
                                const float *__restrict__ min_req, int H, float pos_thr, jtk_colstat *__restrict__ out) {
    __shared__ ReadRef rr[kReadTile];
    const int t = blockIdx.y;
    const uint32_t n_ent = (tmpl_len[t] + 1) * kNumRow;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x * blockDim.x >= n_ent) return;
    const bool live = e < n_ent;
    const uint32_t j = live ? e / kNumRow : 0u, row = e % kNumRow;
    const float thr = min_req_of(min_req, H, row, homop[homop_off[t] + j]);
    double sum = 0.0;
    int cnt = 0;
    unsigned sc[4] = { 0, 0, 0, 0 };
    const uint32_t k_end = tp_start[t + 1];
    for (uint32_t k0 = tp_start[t]; k0 < k_end; k0 += kReadTile) {
        const int n = stage_reads(rr, pairs, tp_ids, k0, k_end);
        if (!live) continue;
#pragma unroll 4
        for (int k = 0; k < n; k++) {
            float x = delta[rr[k].tab_off + e];
            if (fabsf(x) < thr) x = 0.f;
            if (x > pos_thr) { sum += (double)x; cnt++; }
            // (strand, sign) counters as four predicated adds: a computed index put sc[] into local memory
            const bool big = fabsf(x) > 1e-4f, pos = x > 0.f, m0 = rr[k].model == 0;
            sc[0] += (unsigned)(big & !m0 & !pos); sc[1] += (unsigned)(big & !m0 & pos);
            sc[2] += (unsigned)(big & m0 & !pos);  sc[3] += (unsigned)(big & m0 & pos);
        }
    }
    if (!live) return;
    jtk_colstat o;
    o.sum = sum; o.count = cnt;
    o.sc[0] = (uint16_t)min(sc[0], 65535u); o.sc[1] = (uint16_t)min(sc[1], 65535u);
    o.sc[2] = (uint16_t)min(sc[2], 65535u); o.sc[3] = (uint16_t)min(sc[3], 65535u);
    o.pad_ = 0;
    out[stat_off[t] + e] = o;
}

__global__ void gather_kernel(const float *__restrict__ delta, const DevPair *__restrict__ pairs,
                              const uint32_t *__restrict__ tp_ids, uint32_t first, uint32_t n_reads,
                              const uint8_t *__restrict__ homop, const float *__restrict__ min_req, int H,
                              const uint32_t *__restrict__ cols, int D, double *__restrict__ out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_reads * (uint32_t)D) return;
    const uint32_t r = idx / D, d = idx % D;
    const uint32_t e = cols[d];
    const DevPair &p = pairs[tp_ids[first + r]];
    float x = delta[p.tab_off + e];
    const float thr = min_req_of(min_req, H, e % kNumRow, homop[e / kNumRow]);
    if (fabsf(x) < thr) x = 0.f;
    out[idx] = (double)x;
}


// --- filter_profiles on the device (pseudo_mcmc.rs:426-474): per-column statistics + every per-column test -----------
// One thread per table entry: compress_small_gains + column_sum + strand/sign counts over the reads of the chunk (the
// loop of colstats_kernel), then mask / row filter (:443-447), is_in_short_homopolymer (:497-514), has_small_pvalue
// (:476-495), is_explainable_by_strandedness (:314-339) and the Poisson prior (:457-465).  Survivors are appended to one
// global list.  The p-value and prior tables are computed by the host restatement and only looked up here, and the f64
// arithmetic uses explicit round-to-nearest operations in the host's order, so both sides take the same decisions.
struct CandArgs {
    const float *delta; const DevPair *pairs;
    const uint32_t *tp_start, *tp_ids, *tmpl_len;
    const uint8_t *homop; const uint32_t *homop_off;
    const uint8_t *codes; const uint32_t *tmpl_code_off;
    const float *min_req;      // 3*H: Gains::expected * MIN_REQ_FRACTION (fp32, as colstats_kernel)
    const double *expected;    // 3*H: Gains::expected
    int H;
    const double *pv; const uint32_t *pv_off;       // per template: 3*H*(n+1) p-values
    const double *prior; const uint32_t *prior_off; // per template: n+1 Poisson priors
    float pos_thr;
    jtk_candidate *out; int cap; int *counter;
};

__global__ void candidates_kernel(CandArgs a) {
    __shared__ ReadRef rr[kReadTile];
    const int t = blockIdx.y;
    const uint32_t Lt = a.tmpl_len[t];
    const uint32_t n_ent = (Lt + 1) * kNumRow;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x * blockDim.x >= n_ent) return;
    const uint32_t j = min(e, n_ent - 1) / kNumRow, row = e % kNumRow;
    const uint32_t temp_len = Lt + 1; // profile length in positions (pseudo_mcmc.rs:439)
    bool live = e < n_ent;
    if (!(7u <= j && j + 7u <= temp_len)) live = false;   // MASK_LENGTH (:443-446)
    if (!(row < 8u || row == 8u + JTK_COPY_SIZE)) live = false; // (:447)
    const uint8_t *hp = a.homop + a.homop_off[t];
    const uint8_t *tc = a.codes + a.tmpl_code_off[t] + 1; // tc[x] = code of t[x]
    const int type = row < 4 ? 0 : (row < 8 + JTK_COPY_SIZE ? 2 : 1); // Subst, Ins (incl. copy), Del
    // is_in_short_homopolymer (:497-514); j < Lt holds for every unmasked position
    if (live) {
        if (type == 2) {
            const uint32_t r = row - 4;
            const uint32_t prev_len = hp[j - 1] + (r < 4u && tc[j - 1] == r ? 1u : 0u);
            const uint32_t next_len = hp[j] + (r < 4u && tc[j] == r ? 1u : 0u);
            if (!(prev_len <= 2u && next_len <= 2u)) live = false;
        } else if (type == 1) {
            if (!(hp[j] <= 2u)) live = false;
        }
    }
    const int hl = min(max((int)hp[j], 1), a.H);
    const float thr = a.min_req[type * a.H + hl - 1];
    double sum = 0.0;
    unsigned cnt = 0;
    unsigned sc[4] = { 0, 0, 0, 0 };
    const uint32_t k0 = a.tp_start[t], k1 = a.tp_start[t + 1];
    for (uint32_t kt = k0; kt < k1; kt += kReadTile) {
        const int nr = stage_reads(rr, a.pairs, a.tp_ids, kt, k1);
        if (!live) continue;
#pragma unroll 4
        for (int k = 0; k < nr; k++) {
            float x = a.delta[rr[k].tab_off + e];
            if (fabsf(x) < thr) x = 0.f;
            if (x > a.pos_thr) { sum = __dadd_rn(sum, (double)x); cnt++; }
            const bool big = fabsf(x) > 1e-4f, pos = x > 0.f, m0 = rr[k].model == 0;
            sc[0] += (unsigned)(big & !m0 & !pos); sc[1] += (unsigned)(big & !m0 & pos);
            sc[2] += (unsigned)(big & m0 & !pos);  sc[3] += (unsigned)(big & m0 & pos);
        }
    }
    if (!live) return;
    const uint32_t n = k1 - k0;
    // has_small_pvalue (:476-495)
    const double pvalue = a.pv[a.pv_off[t] + (size_t)(type * a.H + hl - 1) * (n + 1) + cnt];
    const double expt = __dmul_rn(a.expected[type * a.H + hl - 1], 0.8);
    if (!(__dmul_rn((double)cnt, expt) < sum && __dmul_rn((double)temp_len, pvalue) < __ddiv_rn(0.05, (double)temp_len))) return;
    // is_explainable_by_strandedness (:314-339): chi-square of the 2x2 (strand, sign) table
    {
        const unsigned strand_count[2] = { sc[0] + sc[1], sc[2] + sc[3] };
        const unsigned sign_count[2] = { sc[0] + sc[2], sc[1] + sc[3] };
        const unsigned tot = strand_count[0] + strand_count[1];
        if (tot == 0) return;
        double chisq = 0.0; // summed per strand row, then over the rows, like the reference's nested .sum()
#pragma unroll
        for (int st = 0; st < 2; st++) {
            double row = 0.0;
#pragma unroll
            for (int sg = 0; sg < 2; sg++) {
                const double expected = __ddiv_rn((double)((unsigned long long)strand_count[st] * sign_count[sg]), (double)tot);
                const double d = __dadd_rn((double)sc[st * 2 + sg], -expected);
                row = __dadd_rn(row, __ddiv_rn(__dmul_rn(d, d), expected)); // 0/0 -> NaN, NaN < 10 is false (as in Rust)
            }
            chisq = __dadd_rn(chisq, row);
        }
        if (!(chisq < 10.0)) return;
    }
    const double total_lk = __dadd_rn(a.prior[a.prior_off[t] + cnt], sum);
    if (!(0.0 < total_lk)) return;
    const int idx = atomicAdd(a.counter, 1);
    if (idx < a.cap) {
        jtk_candidate c;
        c.tmpl = (uint32_t)t; c.pos = e; c.count = cnt; c.pad_ = 0; c.sum = sum; c.lk = total_lk;
        a.out[idx] = c;
    }
}

// compressed profile values of every candidate column: candidate m of template t, read k of t ->
// out[c_base[m] + k * c_stride[m]]  (filter_by, pseudo_mcmc.rs:70-75, for all chunks at once)
__global__ void gather_all_kernel(const float *__restrict__ delta, const DevPair *__restrict__ pairs,
                                  const uint32_t *__restrict__ tp_start, const uint32_t *__restrict__ tp_ids,
                                  const uint8_t *__restrict__ homop, const uint32_t *__restrict__ homop_off,
                                  const float *__restrict__ min_req, int H, const uint32_t *__restrict__ c_tmpl,
                                  const uint32_t *__restrict__ c_pos, const uint32_t *__restrict__ c_base,
                                  const uint32_t *__restrict__ c_stride, double *__restrict__ out) {
    const uint32_t m = blockIdx.x;
    const uint32_t t = c_tmpl[m], e = c_pos[m];
    const uint32_t first = tp_start[t], n = tp_start[t + 1] - first;
    const float thr = min_req_of(min_req, H, e % kNumRow, homop[homop_off[t] + e / kNumRow]);
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        float x = delta[pairs[tp_ids[first + k]].tab_off + e];
        if (fabsf(x) < thr) x = 0.f;
        out[c_base[m] + (size_t)k * c_stride[m]] = (double)x;
    }
}

// pick_filtered_profiles on the device (pseudo_mcmc.rs:516-575): ONE WARP PER CHUNK.  The chunk's candidates are
// cand[first_m[t] ..] in position order, their values vals[val_off[t] + r * M + m] (read r of the chunk, gather_all_kernel).
// ROUND = 3 rounds of max(copy_num, 2) picks: find_next_variants = the unselected candidate with the largest score, the LAST one
// among equals (Iterator::max_by); then every candidate still in play is removed for good when it lies within MASK_LENGTH = 7 bp
// of the pick, or suppressed for the round when its Sokal-Michener or |cosine| similarity with the pick exceeds 0.8.  A lane
// computes the two similarities of one candidate sequentially over the reads, in the reference's order, with explicit
// round-to-nearest f64 operations (no contraction): the same decisions as the host twin, bit for bit.
// Out: n_pick[t], pick_pos[t * probe_cap + d] (flat positions, ascending) and the picked columns of every read of the chunk,
// variants[pair * probe_cap + d] (filter_by, :70-75).  status[t] = 1 when the chunk has more candidates than the kernel's
// shared-memory flags hold (the host twin takes that chunk), 2 when more columns were picked than probe_cap.
constexpr int kPickMaxCand = 2048;
__global__ void __launch_bounds__(32) pick_probes_kernel(const jtk_candidate *__restrict__ cand, const uint32_t *__restrict__ first_m,
                                                         const uint32_t *__restrict__ val_off, const double *__restrict__ vals,
                                                         const int32_t *__restrict__ copy_num, const uint32_t *__restrict__ tp_start,
                                                         const uint32_t *__restrict__ tp_ids, int probe_cap, uint32_t *__restrict__ n_pick,
                                                         uint32_t *__restrict__ pick_pos, double *__restrict__ variants,
                                                         uint32_t *__restrict__ status) {
    __shared__ unsigned char sel[kPickMaxCand];
    const uint32_t t = blockIdx.x, lane = threadIdx.x;
    const uint32_t m0 = first_m[t], M = first_m[t + 1] - m0;
    if (lane == 0) { n_pick[t] = 0; status[t] = 0; }
    if (M == 0 || copy_num[t] < 2) return; // pseudo_mcmc.rs:86-88: single-copy chunks are not searched
    if (M > (uint32_t)kPickMaxCand) { if (lane == 0) status[t] = 1; return; }
    const uint32_t r0 = tp_start[t], nr = tp_start[t + 1] - r0;
    const double *v = vals + val_off[t];
    const jtk_candidate *c = cand + m0;
    for (uint32_t m = lane; m < M; m += 32) sel[m] = 0;
    __syncwarp();
    const int picks = max(copy_num[t], 2);
    for (int round = 0; round < 3; round++) {
        for (uint32_t m = lane; m < M; m += 32) if (sel[m] == 3) sel[m] = 0;
        __syncwarp();
        for (int it = 0; it < picks; it++) {
            // find_next_variants (:590-600)
            double best = 0.0; uint32_t next = 0xffffffffu;
            for (uint32_t m = lane; m < M; m += 32)
                if (sel[m] == 0 && (next == 0xffffffffu || !(c[m].lk < best))) { next = m; best = c[m].lk; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const uint32_t on = __shfl_xor_sync(0xffffffffu, next, o);
                // the later candidate wins among equal scores
                if (on != 0xffffffffu && (next == 0xffffffffu || best < ob || (!(ob < best) && on > next))) { best = ob; next = on; }
            }
            if (next == 0xffffffffu) break;
            const uint32_t picked_bp = c[next].pos / (uint32_t)kNumRow;
            __syncwarp();
            if (lane == 0) sel[next] = 1;
            __syncwarp();
            for (uint32_t m = lane; m < M; m += 32) {
                if (!(sel[m] == 0 || sel[m] == 3)) continue;
                const uint32_t bp = c[m].pos / (uint32_t)kNumRow;
                const uint32_t diff = max(bp, picked_bp) - min(bp, picked_bp);
                if (diff < 7u) { sel[m] = 2; continue; }
                double ip = 0.0, isq = 0.0, jsq = 0.0;
                uint32_t mat = 0, mism = 0;
                for (uint32_t r = 0; r < nr; r++) {
                    const double x = v[(size_t)r * M + next], y = v[(size_t)r * M + m];
                    if (0.00001 < fabs(x) && 0.00001 < fabs(y)) {
                        const double xy = __dmul_rn(x, y);
                        ip = __dadd_rn(ip, xy); isq = __dadd_rn(isq, __dmul_rn(x, x)); jsq = __dadd_rn(jsq, __dmul_rn(y, y));
                        if (0.0 < xy) mat++; else mism++;
                    }
                }
                const uint32_t total = mat + mism;
                const double sok = total == 0 ? 0.0 : __ddiv_rn((double)max(mism, mat), (double)total);
                const double cs = isq == 0.0 ? 0.0 : __ddiv_rn(__ddiv_rn(ip, __dsqrt_rn(isq)), __dsqrt_rn(jsq));
                if (0.8 < sok || 0.8 < fabs(cs)) sel[m] = 3;
            }
            __syncwarp();
        }
    }
    // the selected candidates in candidate (position) order; their columns for every read of the chunk
    uint32_t base = 0;
    for (uint32_t m_lo = 0; m_lo < M; m_lo += 32) {
        const uint32_t m = m_lo + lane;
        const bool on = m < M && sel[m] == 1;
        const unsigned ball = __ballot_sync(0xffffffffu, on);
        const uint32_t d = base + __popc(ball & ((1u << lane) - 1u));
        if (on && d < (uint32_t)probe_cap) {
            pick_pos[(size_t)t * probe_cap + d] = c[m].pos;
            for (uint32_t r = 0; r < nr; r++) variants[(size_t)tp_ids[r0 + r] * probe_cap + d] = v[(size_t)r * M + m];
        }
        base += __popc(ball);
    }
    if (lane == 0) { n_pick[t] = min(base, (uint32_t)probe_cap); if (base > (uint32_t)probe_cap) status[t] = 2; }
}

// plain per-column sums over the first `take` reads of each template (polish loop)
__global__ void colsums_kernel(const float *__restrict__ delta, const DevPair *__restrict__ pairs,
                               const uint32_t *__restrict__ tp_start, const uint32_t *__restrict__ tp_ids,
                               const uint32_t *__restrict__ tmpl_len, const unsigned long long *__restrict__ stat_off,
                               int take, double *__restrict__ out) {
    const int t = blockIdx.y;
    const uint32_t n_ent = (tmpl_len[t] + 1) * kNumRow;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ent) return;
    const uint32_t first = tp_start[t];
    const uint32_t last = min(tp_start[t + 1], first + (uint32_t)take);
    double sum = 0.0;
    for (uint32_t k = first; k < last; k++) sum += (double)delta[pairs[tp_ids[k]].tab_off + e];
    out[stat_off[t] + e] = sum;
}


// The inner loop of the polish edit selection (csrc/polish.cpp select_edits): for column j of template t, the row with the
// largest summed gain over the first `take` reads (> min_gain, first maximum wins) among the rows that are valid at j, or
// -1.  One thread per column; the sums are formed per (column, row) in the order colsums_kernel uses, so the choice is the
// one the host makes on the full table -- but one byte per column crosses PCIe instead of 112.
__global__ void best_edit_kernel(const float *__restrict__ delta, const DevPair *__restrict__ pairs,
                                 const uint32_t *__restrict__ tp_start, const uint32_t *__restrict__ tp_ids,
                                 const uint32_t *__restrict__ tmpl_len, const uint8_t *__restrict__ codes,
                                 const uint32_t *__restrict__ tmpl_code_off, const unsigned long long *__restrict__ col_off,
                                 int take, int ignore_edge, double min_gain, int8_t *__restrict__ out,
                                 double *__restrict__ out_gain) {
    __shared__ ReadRef rr[kReadTile];
    const int t = blockIdx.y;
    const int L = (int)tmpl_len[t];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(blockIdx.x * blockDim.x) > L) return;
    const bool live = j <= L;
    double sum[kNumRow];
#pragma unroll
    for (int r = 0; r < kNumRow; r++) sum[r] = 0.0;
    const uint32_t first = tp_start[t];
    const uint32_t last = min(tp_start[t + 1], first + (uint32_t)max(take, 0));
    for (uint32_t k0 = first; k0 < last; k0 += kReadTile) {
        const int n = stage_reads(rr, pairs, tp_ids, k0, last);
        if (!live) continue;
        for (int k = 0; k < n; k++) {
            const float *row = delta + rr[k].tab_off + (size_t)j * kNumRow;
#pragma unroll
            for (int r = 0; r < kNumRow; r++) sum[r] += (double)row[r];
        }
    }
    if (!live) return;
    int best = -1;
    double bg = min_gain;
    const int own = j < L ? (int)(codes[tmpl_code_off[t] + 1 + j] & 3) : -1; // template code of t[j]
    if (j >= ignore_edge && j <= L - ignore_edge) {
#pragma unroll
        for (int r = 0; r < kNumRow; r++) {
            if (r == own) continue;
            if (r < 4 && j >= L - ignore_edge) continue;
            if (r >= 8 && r < 8 + JTK_COPY_SIZE && j + (r - 7) > L - ignore_edge) continue;
            if (r >= 8 + JTK_COPY_SIZE && j + (r - 7 - JTK_COPY_SIZE) > L - ignore_edge) continue;
            if (sum[r] > bg) { bg = sum[r]; best = r; }
        }
    }
    out[col_off[t] + j] = (int8_t)best;
    if (out_gain) out_gain[col_off[t] + j] = best >= 0 ? bg : 0.0;
}

} // namespace

extern "C" {

int jtk_batch_create(jtk_ctx *ctx, int n_pairs, int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                     const uint8_t *read_concat, const uint32_t *read_off, const uint8_t *ops_concat,
                     const uint32_t *ops_off, const uint8_t *strand, const uint32_t *tmpl_idx, int radius, jtk_batch **out) {
    return batch_create(ctx, n_pairs, n_tmpl, tmpl_concat, tmpl_off, read_concat, read_off, ops_concat, ops_off, strand,
                        tmpl_idx, radius, false, out);
}

void jtk_batch_destroy(jtk_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    b->release();
    delete b;
}

uint64_t jtk_batch_cell_updates(const jtk_batch *b) { return b ? b->cell_updates : 0; }
uint64_t jtk_batch_h2d_bytes(const jtk_batch *b) { return b ? b->h2d_bytes : 0; }

int jtk_batch_modtable(jtk_batch *b, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int rows) {
    if (!b) return JTK_EINVAL;
    return batch_run(b, fwd, rev, true, rows);
}

int jtk_batch_sync(jtk_batch *b) { return b ? batch_sync(b) : JTK_EINVAL; }

int jtk_batch_fetch_lk(jtk_batch *b, double *out_lk) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (!out_lk) return ctx->fail(JTK_EINVAL, "out_lk is NULL");
    if (b->n_pairs == 0) return JTK_OK;
    CU(ctx->h_lk.reserve((size_t)b->n_pairs), "cudaMallocHost lk");
    CU(cudaMemcpyAsync(ctx->h_lk.p, b->d_lk.p, sizeof(double) * (size_t)b->n_pairs, cudaMemcpyDeviceToHost, ctx->stream), "D2H lk");
    int rc = batch_sync(b);
    if (rc) return rc;
    std::memcpy(out_lk, ctx->h_lk.p, sizeof(double) * (size_t)b->n_pairs);
    return JTK_OK;
}

int jtk_batch_fetch_profile(jtk_batch *b, int pair, float *out) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (!out || pair < 0 || pair >= b->n_pairs) return ctx->fail(JTK_EINVAL, "bad pair index");
    if (!b->has_profiles) return ctx->fail(JTK_ESTATE, "jtk_batch_modtable has not run");
    const DevPair &dp = b->pairs[(size_t)pair];
    CU(cudaMemcpyAsync(out, b->d_delta.p + dp.tab_off, sizeof(float) * (size_t)(dp.Lt + 1) * kNumRow, cudaMemcpyDeviceToHost, ctx->stream), "D2H profile");
    return batch_sync(b);
}

int jtk_batch_colstats(jtk_batch *b, const float *min_req, int H, float pos_thr, jtk_colstat *out, const uint64_t *stat_off) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (!min_req || H < 1 || !stat_off) return ctx->fail(JTK_EINVAL, "null argument");
    if (!b->has_profiles) return ctx->fail(JTK_ESTATE, "jtk_batch_modtable has not run");
    if (b->n_tmpl == 0) return JTK_OK;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    uint64_t total = 0;
    for (int t = 0; t < b->n_tmpl; t++) total = std::max<uint64_t>(total, stat_off[t] + (uint64_t)(b->tmpl_len[t] + 1) * kNumRow);
    CU(ctx->d_minreq.reserve((size_t)3 * H), "cudaMalloc min_req");
    CU(b->d_stat_off.reserve((size_t)b->n_tmpl), "cudaMalloc stat_off");
    CU(b->d_stats.reserve((size_t)total), "cudaMalloc stats");
    CU(ctx->up_minreq.put(ctx->d_minreq.p, min_req, sizeof(float) * 3 * (size_t)H, st), "H2D min_req");
    CU(b->up_stat_off.put(b->d_stat_off.p, stat_off, sizeof(uint64_t) * (size_t)b->n_tmpl, st), "H2D stat_off");
    if (b->n_tmpl > 65535) return ctx->fail(JTK_EINVAL, "more than 65535 templates in one batch (the per-chunk kernels put the template on grid.y): split the call");
    dim3 grid((unsigned)(((size_t)(b->max_lt + 1) * kNumRow + 255) / 256), (unsigned)b->n_tmpl);
    colstats_kernel<<<grid, 256, 0, st>>>(b->d_delta.p, b->d_pairs.p, b->d_tp_start.p, b->d_tp_ids.p, b->d_tmpl_len.p,
                                         b->d_homop.p, b->d_homop_off.p, b->d_stat_off.p, ctx->d_minreq.p, H, pos_thr,
                                         b->d_stats.p);
    CU(cudaGetLastError(), "colstats launch");
    ctx->launches++;
    if (!out) return JTK_OK;
    CU(cudaMemcpyAsync(out, b->d_stats.p, sizeof(jtk_colstat) * (size_t)total, cudaMemcpyDeviceToHost, st), "D2H stats");
    return batch_sync(b);
}

int jtk_mcmc_restarts_batch(jtk_ctx *ctx, int n_chains, const double *data_concat, const uint64_t *data_off,
                            const uint32_t *n_rows, const uint32_t *n_cols, const uint32_t *n_clusters,
                            const double *size_to_lk_concat, int restarts, uint64_t *rng_state, uint8_t *out_asn,
                            double *out_lk, int *out_err) {
    if (!ctx) return JTK_EINVAL;
    if (n_chains < 0 || restarts < 0) return ctx->fail(JTK_EINVAL, "negative count");
    if (n_chains == 0) return JTK_OK;
    if (!data_concat || !data_off || !n_rows || !n_cols || !n_clusters || !size_to_lk_concat || !rng_state || !out_asn || !out_lk || !out_err)
        return ctx->fail(JTK_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    const auto t_begin = std::chrono::steady_clock::now();
    std::vector<McmcChain> chains((size_t)n_chains);
    std::vector<uint64_t> asn_off((size_t)n_chains);
    size_t f64 = 0, u8 = 0, asn = 0, s2l = 0, smem = 0;
    for (int c = 0; c < n_chains; c++) {
        const uint32_t n = n_rows[c], D = n_cols[c], k = n_clusters[c];
        if (n == 0 || D == 0 || D > (uint32_t)kMcmcMaxD || k == 0 || k > (uint32_t)kMcmcMaxK || k > n)
            return ctx->fail(JTK_EINVAL, "chain needs 1 <= k <= min(n, 8) and 1 <= D <= 32");
        McmcChain &m = chains[(size_t)c];
        m.n = n; m.D = D; m.k = k; m.pad_ = 0; m.cov = 0.0;
        m.off_f64 = f64; m.off_u32 = 0; m.off_u8 = u8;
        asn_off[(size_t)c] = asn;
        f64 += mcmc_f64_words(n, D); u8 += (mcmc_u8_bytes(n, D) + 127) & ~(size_t)127;
        smem = std::max(smem, mcmc_smem_bytes(n, D, k));
        asn += n;
    }
    if (smem > 200 * 1024) return ctx->fail(JTK_EINVAL, "chain too large for the shared memory of one SM");
    // workspace: flat data and size_to_lk go in, everything else is scratch the kernel initialises itself
    std::vector<double> wf(f64, 0.0);
    for (int c = 0; c < n_chains; c++) {
        const McmcChain &m = chains[(size_t)c];
        std::memcpy(wf.data() + m.off_f64, data_concat + data_off[c], sizeof(double) * (size_t)m.n * m.D);
        std::memcpy(wf.data() + m.off_f64 + (size_t)m.n * m.D, size_to_lk_concat + s2l, sizeof(double) * (m.n + 1));
        s2l += m.n + 1;
    }
    CU(ctx->d_mc_chains.reserve((size_t)n_chains), "cudaMalloc mcmc chains");
    CU(ctx->d_mc_f64.reserve(f64), "cudaMalloc mcmc workspace");
    CU(ctx->d_mc_u8.reserve(u8), "cudaMalloc mcmc workspace");
    CU(ctx->d_mc_asn.reserve(asn), "cudaMalloc mcmc assignments");
    CU(ctx->d_mc_asn_off.reserve((size_t)n_chains), "cudaMalloc mcmc offsets");
    CU(ctx->d_mc_rng.reserve((size_t)4 * n_chains), "cudaMalloc mcmc rng");
    CU(ctx->d_mc_lk.reserve((size_t)n_chains), "cudaMalloc mcmc lk");
    CU(ctx->d_mc_err.reserve((size_t)n_chains), "cudaMalloc mcmc status");
    CU(cudaMemcpyAsync(ctx->d_mc_chains.p, chains.data(), sizeof(McmcChain) * (size_t)n_chains, cudaMemcpyHostToDevice, st), "H2D mcmc chains");
    CU(cudaMemcpyAsync(ctx->d_mc_f64.p, wf.data(), sizeof(double) * f64, cudaMemcpyHostToDevice, st), "H2D mcmc data");
    CU(cudaMemcpyAsync(ctx->d_mc_asn_off.p, asn_off.data(), sizeof(uint64_t) * (size_t)n_chains, cudaMemcpyHostToDevice, st), "H2D mcmc offsets");
    CU(cudaMemcpyAsync(ctx->d_mc_rng.p, rng_state, sizeof(uint64_t) * 4 * (size_t)n_chains, cudaMemcpyHostToDevice, st), "H2D mcmc rng");
    // chain indices grouped by kernel class, each group sorted by read count (the groups of a warp then run the same trip counts)
    std::vector<int> ids((size_t)n_chains);
    int class_count[5] = { 0, 0, 0, 0, 0 };
    {
        std::vector<int> cls((size_t)n_chains);
        for (int c = 0; c < n_chains; c++) { cls[(size_t)c] = mcmc_class_of(n_rows[c], n_cols[c], n_clusters[c]); class_count[cls[(size_t)c]]++; }
        for (int c = 0; c < n_chains; c++) ids[(size_t)c] = c;
        std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) {
            if (cls[(size_t)a] != cls[(size_t)b]) return cls[(size_t)a] < cls[(size_t)b];
            return n_rows[a] < n_rows[b];
        });
    }
    CU(ctx->d_mc_u32.reserve((size_t)n_chains), "cudaMalloc mcmc ids");
    CU(cudaMemcpyAsync(ctx->d_mc_u32.p, ids.data(), sizeof(int) * (size_t)n_chains, cudaMemcpyHostToDevice, st), "H2D mcmc ids");
    CU(launch_mcmc_restarts(ctx->d_mc_chains.p, chains.data(), reinterpret_cast<const int *>(ctx->d_mc_u32.p), ids.data(), class_count, n_chains,
                            ctx->d_mc_f64.p, ctx->d_mc_u8.p, ctx->d_mc_rng.p, ctx->d_mc_asn.p, ctx->d_mc_asn_off.p, ctx->d_mc_lk.p,
                            ctx->d_mc_err.p, restarts, smem, st), "mcmc launch");
    for (int c = 0; c < 5; c++) ctx->launches += class_count[c] > 0 ? 1 : 0;
    CU(cudaMemcpyAsync(rng_state, ctx->d_mc_rng.p, sizeof(uint64_t) * 4 * (size_t)n_chains, cudaMemcpyDeviceToHost, st), "D2H mcmc rng");
    CU(cudaMemcpyAsync(out_asn, ctx->d_mc_asn.p, asn, cudaMemcpyDeviceToHost, st), "D2H mcmc assignments");
    CU(cudaMemcpyAsync(out_lk, ctx->d_mc_lk.p, sizeof(double) * (size_t)n_chains, cudaMemcpyDeviceToHost, st), "D2H mcmc lk");
    CU(cudaMemcpyAsync(out_err, ctx->d_mc_err.p, sizeof(int) * (size_t)n_chains, cudaMemcpyDeviceToHost, st), "D2H mcmc status");
    CU(cudaStreamSynchronize(st), "mcmc execution");
    if (std::getenv("JTK_MCMC_DEBUG"))
        std::fprintf(stderr, "[jtk] mcmc_restarts: %d chains, classes (2,4,6,8 columns, general) = %d %d %d %d %d, %.3f s from upload to results\n",
                     n_chains, class_count[0], class_count[1], class_count[2], class_count[3], class_count[4],
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count());
    return JTK_OK;
}

int jtk_batch_gather(jtk_batch *b, int tmpl, const float *min_req, int H, const uint32_t *cols, int D, double *out) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (tmpl < 0 || tmpl >= b->n_tmpl || !min_req || H < 1 || D < 0 || (D > 0 && (!cols || !out)))
        return ctx->fail(JTK_EINVAL, "bad argument");
    if (!b->has_profiles) return ctx->fail(JTK_ESTATE, "jtk_batch_modtable has not run");
    const uint32_t first = b->tp_start[(size_t)tmpl], n_reads = b->tp_start[(size_t)tmpl + 1] - first;
    if (D == 0 || n_reads == 0) return JTK_OK;
    const uint32_t n_ent = (b->tmpl_len[(size_t)tmpl] + 1) * kNumRow;
    for (int d = 0; d < D; d++)
        if (cols[d] >= n_ent) return ctx->fail(JTK_EINVAL, "column index out of range");
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    CU(ctx->d_minreq.reserve((size_t)3 * H), "cudaMalloc min_req");
    CU(ctx->d_cols.reserve((size_t)D), "cudaMalloc cols");
    CU(ctx->d_gather.reserve((size_t)D * n_reads), "cudaMalloc gather");
    CU(ctx->up_minreq.put(ctx->d_minreq.p, min_req, sizeof(float) * 3 * (size_t)H, st), "H2D min_req");
    CU(cudaMemcpyAsync(ctx->d_cols.p, cols, sizeof(uint32_t) * (size_t)D, cudaMemcpyHostToDevice, st), "H2D cols");
    const uint32_t n = n_reads * (uint32_t)D;
    gather_kernel<<<(n + 127) / 128, 128, 0, st>>>(b->d_delta.p, b->d_pairs.p, b->d_tp_ids.p, first, n_reads,
                                                   b->d_homop.p + b->homop_off[(size_t)tmpl], ctx->d_minreq.p, H,
                                                   ctx->d_cols.p, D, ctx->d_gather.p);
    CU(cudaGetLastError(), "gather launch");
    ctx->launches++;
    CU(cudaMemcpyAsync(out, ctx->d_gather.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st), "D2H gather");
    return batch_sync(b);
}

} // extern "C"

// ---- filter_profiles / search_variants for every chunk of a batch ---------------------------------------------------
namespace {
struct CandScratch {
    std::vector<double> pv, prior, expected;
    std::vector<uint32_t> pv_off, prior_off;
    std::vector<float> min_req;
};

int run_candidates(jtk_batch *b, const jtk_gains *gains, const int32_t *copy_num, double coverage, int cap, int *out_n) {
    jtk_ctx *ctx = b->ctx;
    const int H = gains->homop_len, n_tmpl = b->n_tmpl;
    cudaStream_t st = ctx->stream;
    CandScratch sc;
    sc.min_req.resize((size_t)3 * H);
    sc.expected.assign(gains->gain, gains->gain + (size_t)3 * H);
    for (int k = 0; k < 3 * H; k++) sc.min_req[k] = (float)(gains->gain[k] * 0.5); // MIN_REQ_FRACTION (:140)
    // tables per distinct read count / (copy number, read count)
    sc.pv_off.resize((size_t)n_tmpl); sc.prior_off.resize((size_t)n_tmpl);
    std::vector<std::pair<uint32_t, uint32_t>> pv_seen;                 // (n, offset)
    std::vector<std::pair<std::pair<uint32_t, int32_t>, uint32_t>> pr_seen; // ((n, copy_num), offset)
    std::vector<double> tmp;
    for (int t = 0; t < n_tmpl; t++) {
        const uint32_t n = b->tp_start[(size_t)t + 1] - b->tp_start[(size_t)t];
        if (copy_num[t] < 1) return ctx->fail(JTK_EINVAL, "copy_num must be >= 1");
        uint32_t off = UINT32_MAX;
        for (auto &kv : pv_seen) if (kv.first == n) off = kv.second;
        if (off == UINT32_MAX) {
            off = (uint32_t)sc.pv.size();
            jtk::host::pvalue_tables(gains, n, tmp);
            sc.pv.insert(sc.pv.end(), tmp.begin(), tmp.end());
            pv_seen.push_back({ n, off });
        }
        sc.pv_off[t] = off;
        off = UINT32_MAX;
        for (auto &kv : pr_seen) if (kv.first.first == n && kv.first.second == copy_num[t]) off = kv.second;
        if (off == UINT32_MAX) {
            off = (uint32_t)sc.prior.size();
            jtk::host::poisson_prior_table(coverage, (size_t)copy_num[t], n, tmp);
            sc.prior.insert(sc.prior.end(), tmp.begin(), tmp.end());
            pr_seen.push_back({ { n, copy_num[t] }, off });
        }
        sc.prior_off[t] = off;
    }
    CU(ctx->d_minreq.reserve((size_t)3 * H), "cudaMalloc min_req");
    CU(ctx->d_tabs.reserve(sc.pv.size() + sc.prior.size() + sc.expected.size()), "cudaMalloc tables");
    CU(ctx->d_tab_off.reserve((size_t)2 * n_tmpl), "cudaMalloc table offsets");
    CU(ctx->d_cand.reserve((size_t)cap), "cudaMalloc candidates");
    CU(ctx->d_counter.reserve(1), "cudaMalloc counter");
    double *d_pv = ctx->d_tabs.p, *d_prior = d_pv + sc.pv.size(), *d_exp = d_prior + sc.prior.size();
    CU(ctx->up_minreq.put(ctx->d_minreq.p, sc.min_req.data(), sizeof(float) * 3 * (size_t)H, st), "H2D min_req");
    CU(cudaMemcpyAsync(d_pv, sc.pv.data(), sizeof(double) * sc.pv.size(), cudaMemcpyHostToDevice, st), "H2D pvalues");
    CU(cudaMemcpyAsync(d_prior, sc.prior.data(), sizeof(double) * sc.prior.size(), cudaMemcpyHostToDevice, st), "H2D prior");
    CU(cudaMemcpyAsync(d_exp, sc.expected.data(), sizeof(double) * sc.expected.size(), cudaMemcpyHostToDevice, st), "H2D expected");
    CU(cudaMemcpyAsync(ctx->d_tab_off.p, sc.pv_off.data(), sizeof(uint32_t) * (size_t)n_tmpl, cudaMemcpyHostToDevice, st), "H2D pv_off");
    CU(cudaMemcpyAsync(ctx->d_tab_off.p + n_tmpl, sc.prior_off.data(), sizeof(uint32_t) * (size_t)n_tmpl, cudaMemcpyHostToDevice, st), "H2D prior_off");
    CU(cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(int), st), "memset counter");
    CandArgs a{};
    a.delta = b->d_delta.p; a.pairs = b->d_pairs.p; a.tp_start = b->d_tp_start.p; a.tp_ids = b->d_tp_ids.p;
    a.tmpl_len = b->d_tmpl_len.p; a.homop = b->d_homop.p; a.homop_off = b->d_homop_off.p; a.codes = b->d_codes.p;
    a.tmpl_code_off = b->d_tmpl_code_off.p; a.min_req = ctx->d_minreq.p; a.expected = d_exp; a.H = H;
    a.pv = d_pv; a.pv_off = ctx->d_tab_off.p; a.prior = d_prior; a.prior_off = ctx->d_tab_off.p + n_tmpl;
    a.pos_thr = 1e-5f; a.out = ctx->d_cand.p; a.cap = cap; a.counter = ctx->d_counter.p;
    if (b->n_tmpl > 65535) return ctx->fail(JTK_EINVAL, "more than 65535 templates in one batch (the per-chunk kernels put the template on grid.y): split the call");
    dim3 grid((unsigned)(((size_t)(b->max_lt + 1) * kNumRow + 255) / 256), (unsigned)n_tmpl);
    candidates_kernel<<<grid, 256, 0, st>>>(a);
    CU(cudaGetLastError(), "candidates launch");
    ctx->launches++;
    // the scratch vectors must outlive the asynchronous copies: wait for the count right here
    int n = 0;
    CU(cudaMemcpyAsync(&n, ctx->d_counter.p, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H candidate count");
    int rc = batch_sync(b);
    if (rc) return rc;
    *out_n = n;
    return JTK_OK;
}
} // namespace

extern "C" {

int jtk_batch_candidates(jtk_batch *b, const jtk_gains *gains, const int32_t *copy_num, double coverage,
                         jtk_candidate *out, int cap, int *out_n) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (!gains || !gains->gain || !gains->prob || gains->homop_len < 1 || !copy_num || !out_n || cap < 0 || (cap > 0 && !out))
        return ctx->fail(JTK_EINVAL, "null / bad argument");
    if (!b->has_profiles) return ctx->fail(JTK_ESTATE, "jtk_batch_modtable has not run");
    *out_n = 0;
    if (b->n_tmpl == 0 || b->n_pairs == 0) return JTK_OK;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    const int dev_cap = std::max(cap, 1);
    int n = 0;
    int rc = run_candidates(b, gains, copy_num, coverage, dev_cap, &n);
    if (rc) return rc;
    *out_n = n;
    if (n > cap) return ctx->fail(JTK_EINVAL, "candidate buffer too small: " + std::to_string(n) + " candidates");
    if (n == 0) return JTK_OK;
    CU(ctx->h_cand.reserve((size_t)n), "cudaMallocHost candidates");
    CU(cudaMemcpyAsync(ctx->h_cand.p, ctx->d_cand.p, sizeof(jtk_candidate) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream), "D2H candidates");
    rc = batch_sync(b);
    if (rc) return rc;
    std::sort(ctx->h_cand.p, ctx->h_cand.p + n, [](const jtk_candidate &x, const jtk_candidate &y) {
        return x.tmpl != y.tmpl ? x.tmpl < y.tmpl : x.pos < y.pos;
    });
    std::memcpy(out, ctx->h_cand.p, sizeof(jtk_candidate) * (size_t)n);
    return JTK_OK;
}

int jtk_batch_search_variants(jtk_batch *b, const jtk_gains *gains, const int32_t *copy_num, double coverage, int probe_cap,
                              uint32_t *out_n_probes, uint32_t *out_probe_pos, double *out_variants) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (!gains || !gains->gain || !gains->prob || gains->homop_len < 1 || !copy_num || probe_cap < 1 || !out_n_probes ||
        !out_probe_pos || !out_variants)
        return ctx->fail(JTK_EINVAL, "null / bad argument");
    if (!b->has_profiles) return ctx->fail(JTK_ESTATE, "jtk_batch_modtable has not run");
    const int n_tmpl = b->n_tmpl;
    for (int t = 0; t < n_tmpl; t++) {
        out_n_probes[t] = 0;
        if (copy_num[t] >= 1 && 3 * std::max(copy_num[t], 2) > probe_cap)
            return ctx->fail(JTK_EINVAL, "probe_cap must be at least 3 * max(copy_num, 2)");
    }
    if (n_tmpl == 0 || b->n_pairs == 0) return JTK_OK;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    // 1. candidate columns of every chunk (device)
    int cap = 256 * n_tmpl + 4096, n = 0;
    int rc = run_candidates(b, gains, copy_num, coverage, cap, &n);
    if (rc) return rc;
    if (n > cap) { // rare: rerun with the exact size
        cap = n;
        rc = run_candidates(b, gains, copy_num, coverage, cap, &n);
        if (rc) return rc;
    }
    std::fill(out_variants, out_variants + (size_t)b->n_pairs * probe_cap, 0.0);
    if (n == 0) return JTK_OK;
    CU(ctx->h_cand.reserve((size_t)n), "cudaMallocHost candidates");
    CU(cudaMemcpyAsync(ctx->h_cand.p, ctx->d_cand.p, sizeof(jtk_candidate) * (size_t)n, cudaMemcpyDeviceToHost, st), "D2H candidates");
    rc = batch_sync(b);
    if (rc) return rc;
    jtk_candidate *cand = ctx->h_cand.p;
    std::sort(cand, cand + n, [](const jtk_candidate &x, const jtk_candidate &y) {
        return x.tmpl != y.tmpl ? x.tmpl < y.tmpl : x.pos < y.pos;
    });
    // 2. their values for every read of the chunk (device gather), laid out as cand[r * M_t + m] per chunk
    std::vector<uint32_t> meta((size_t)4 * n); // c_tmpl | c_pos | c_base | c_stride
    std::vector<uint32_t> first_m((size_t)n_tmpl + 1, 0), val_off((size_t)n_tmpl + 1, 0);
    for (int m = 0; m < n; m++) first_m[cand[m].tmpl + 1]++;
    for (int t = 0; t < n_tmpl; t++) {
        first_m[t + 1] += first_m[t];
        const uint32_t Mt = first_m[t + 1] - first_m[t];
        val_off[t + 1] = val_off[t] + Mt * (b->tp_start[(size_t)t + 1] - b->tp_start[(size_t)t]);
    }
    for (int m = 0; m < n; m++) {
        const uint32_t t = cand[m].tmpl;
        meta[m] = t; meta[(size_t)n + m] = cand[m].pos;
        meta[(size_t)2 * n + m] = val_off[t] + ((uint32_t)m - first_m[t]);
        meta[(size_t)3 * n + m] = first_m[t + 1] - first_m[t];
    }
    const size_t n_val = val_off[n_tmpl];
    CU(ctx->d_cols.reserve(meta.size()), "cudaMalloc candidate meta");
    CU(ctx->d_gather.reserve(n_val), "cudaMalloc candidate values");
    CU(ctx->h_gather.reserve(n_val), "cudaMallocHost candidate values");
    CU(cudaMemcpyAsync(ctx->d_cols.p, meta.data(), sizeof(uint32_t) * meta.size(), cudaMemcpyHostToDevice, st), "H2D candidate meta");
    gather_all_kernel<<<(unsigned)n, 64, 0, st>>>(b->d_delta.p, b->d_pairs.p, b->d_tp_start.p, b->d_tp_ids.p, b->d_homop.p,
                                                 b->d_homop_off.p, ctx->d_minreq.p, gains->homop_len, ctx->d_cols.p,
                                                 ctx->d_cols.p + n, ctx->d_cols.p + 2 * (size_t)n, ctx->d_cols.p + 3 * (size_t)n,
                                                 ctx->d_gather.p);
    CU(cudaGetLastError(), "gather launch");
    ctx->launches++;
    // 3. greedy pick per chunk (pseudo_mcmc.rs:516-575) and filter_by (:70-75): on the device, one warp per chunk -- only the
    //    picked positions and the n x D variant columns cross PCIe.  JTK_HOST_PICK=1 keeps the host twin.
    const bool host_pick = std::getenv("JTK_HOST_PICK") != nullptr;
    std::vector<uint8_t> on_host((size_t)n_tmpl, host_pick ? 1 : 0);
    bool any_host = host_pick;
    if (!host_pick) {
        // sorted candidates back to the device, chunk offsets and copy numbers behind the gather meta
        std::vector<uint32_t> aux((size_t)3 * n_tmpl + 2);
        std::copy(first_m.begin(), first_m.end(), aux.begin());
        std::copy(val_off.begin(), val_off.end(), aux.begin() + n_tmpl + 1);
        for (int t = 0; t < n_tmpl; t++) aux[(size_t)2 * n_tmpl + 2 + t] = (uint32_t)copy_num[t];
        const size_t n_out = (size_t)2 * n_tmpl + (size_t)n_tmpl * probe_cap; // n_pick | status | pick_pos
        CU(ctx->d_pick.reserve(aux.size() + n_out), "cudaMalloc pick buffers");
        CU(ctx->h_pick.reserve(n_out), "cudaMallocHost pick buffers");
        CU(ctx->d_var.reserve((size_t)b->n_pairs * probe_cap), "cudaMalloc variants");
        CU(ctx->h_gather.reserve(std::max(n_val, (size_t)b->n_pairs * probe_cap)), "cudaMallocHost variants");
        CU(cudaMemcpyAsync(ctx->d_cand.p, cand, sizeof(jtk_candidate) * (size_t)n, cudaMemcpyHostToDevice, st), "H2D sorted candidates");
        CU(cudaMemcpyAsync(ctx->d_pick.p, aux.data(), sizeof(uint32_t) * aux.size(), cudaMemcpyHostToDevice, st), "H2D pick meta");
        CU(cudaMemsetAsync(ctx->d_var.p, 0, sizeof(double) * (size_t)b->n_pairs * probe_cap, st), "memset variants");
        uint32_t *d_aux = ctx->d_pick.p, *d_out = ctx->d_pick.p + aux.size();
        pick_probes_kernel<<<(unsigned)n_tmpl, 32, 0, st>>>(ctx->d_cand.p, d_aux, d_aux + n_tmpl + 1, ctx->d_gather.p,
                                                           reinterpret_cast<const int32_t *>(d_aux + 2 * (size_t)n_tmpl + 2),
                                                           b->d_tp_start.p, b->d_tp_ids.p, probe_cap, d_out, d_out + 2 * (size_t)n_tmpl,
                                                           ctx->d_var.p, d_out + n_tmpl);
        CU(cudaGetLastError(), "pick launch");
        ctx->launches++;
        CU(cudaMemcpyAsync(ctx->h_pick.p, d_out, sizeof(uint32_t) * n_out, cudaMemcpyDeviceToHost, st), "D2H picks");
        CU(cudaMemcpyAsync(ctx->h_gather.p, ctx->d_var.p, sizeof(double) * (size_t)b->n_pairs * probe_cap, cudaMemcpyDeviceToHost, st), "D2H variants");
        rc = batch_sync(b);
        if (rc) return rc;
        std::memcpy(out_variants, ctx->h_gather.p, sizeof(double) * (size_t)b->n_pairs * probe_cap);
        const uint32_t *npk = ctx->h_pick.p, *stt = ctx->h_pick.p + n_tmpl, *ppos = ctx->h_pick.p + 2 * (size_t)n_tmpl;
        for (int t = 0; t < n_tmpl; t++) {
            if (stt[t] == 2) return ctx->fail(JTK_EINVAL, "probe_cap too small");
            if (stt[t] == 1) { on_host[(size_t)t] = 1; any_host = true; continue; } // more candidates than the kernel's flags hold
            out_n_probes[t] = npk[t];
            for (uint32_t d = 0; d < npk[t]; d++) out_probe_pos[(size_t)t * probe_cap + d] = ppos[(size_t)t * probe_cap + d];
        }
    }
    if (!any_host) return JTK_OK;
    CU(ctx->h_gather.reserve(n_val), "cudaMallocHost candidate values");
    CU(cudaMemcpyAsync(ctx->h_gather.p, ctx->d_gather.p, sizeof(double) * n_val, cudaMemcpyDeviceToHost, st), "D2H candidate values");
    rc = batch_sync(b);
    if (rc) return rc;
    std::vector<uint32_t> pos;
    std::vector<double> lk;
    try {
        for (int t = 0; t < n_tmpl; t++) {
            const uint32_t m0 = first_m[t], Mt = first_m[t + 1] - m0;
            if (!on_host[(size_t)t] || Mt == 0 || copy_num[t] < 2) continue; // pseudo_mcmc.rs:86-88: single-copy chunks are not searched
            const uint32_t r0 = b->tp_start[(size_t)t], nr = b->tp_start[(size_t)t + 1] - r0;
            pos.resize(Mt); lk.resize(Mt);
            for (uint32_t m = 0; m < Mt; m++) { pos[m] = cand[m0 + m].pos; lk[m] = cand[m0 + m].lk; }
            const double *vals = ctx->h_gather.p + val_off[t];
            const std::vector<size_t> picked = jtk::host::pick_probes(pos.data(), lk.data(), Mt, vals, nr, (size_t)copy_num[t]);
            if ((int)picked.size() > probe_cap) return ctx->fail(JTK_EINVAL, "probe_cap too small");
            out_n_probes[t] = (uint32_t)picked.size();
            for (size_t d = 0; d < picked.size(); d++) {
                out_probe_pos[(size_t)t * probe_cap + d] = pos[picked[d]];
                for (uint32_t r = 0; r < nr; r++)
                    out_variants[(size_t)b->tp_ids[r0 + r] * probe_cap + d] = vals[(size_t)r * Mt + picked[d]];
            }
        }
    } catch (const std::exception &e) {
        return ctx->fail(JTK_EINVAL, std::string("pick_filtered_profiles: ") + e.what());
    }
    return JTK_OK;
}

} // extern "C"

extern "C" {

int jtk_ctx_timer_start(jtk_ctx *ctx) {
    if (!ctx) return JTK_EINVAL;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    CU(cudaEventRecord(ctx->tm0, ctx->stream), "event");
    return JTK_OK;
}

int jtk_ctx_timer_stop(jtk_ctx *ctx, float *ms) {
    if (!ctx || !ms) return JTK_EINVAL;
    CU(cudaEventRecord(ctx->tm1, ctx->stream), "event");
    CU(cudaStreamSynchronize(ctx->stream), "sync");
    CU(cudaEventElapsedTime(ms, ctx->tm0, ctx->tm1), "elapsed");
    return JTK_OK;
}

int jtk_ctx_kernel_times(jtk_ctx *ctx, float *ms, int cap) {
    if (!ctx || (!ms && cap > 0)) return JTK_EINVAL;
    CU(cudaStreamSynchronize(ctx->stream), "sync");
    const int n = std::min(cap, ctx->ring_n);
    for (int k = 0; k < n; k++) CU(cudaEventElapsedTime(&ms[k], ctx->ring0[k], ctx->ring1[k]), "elapsed");
    ctx->ring_n = 0;
    return n;
}

int jtk_ctx_measure_fp32_peak(jtk_ctx *ctx, double *tflops_ffma, double *tflops_ffma2) {
    if (!ctx || !tflops_ffma || !tflops_ffma2) return JTK_EINVAL;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    const int iters = 4096, threads = 256, blocks = ctx->sm_count * 8;
    float *sink = nullptr;
    CU(cudaMalloc((void **)&sink, sizeof(float) * (size_t)blocks * threads), "cudaMalloc sink");
    double best[2] = { 0, 0 };
    for (int mode = 0; mode < 2; mode++)
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(ctx->tm0, ctx->stream);
            cudaError_t e = launch_fp32_peak(mode, blocks, threads, iters, sink, ctx->stream);
            cudaEventRecord(ctx->tm1, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { cudaFree(sink); return ctx->cuda_fail(e, "fp32 peak kernel"); }
            ctx->launches++;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->tm0, ctx->tm1);
            // 16 independent accumulators per thread, 2 flops per FMA (packed: 2 FMAs per instruction, 8 registers pairs)
            const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
            if (rep > 0) best[mode] = std::max(best[mode], flops / (ms * 1e-3) / 1e12);
        }
    cudaFree(sink);
    *tflops_ffma = best[0];
    *tflops_ffma2 = best[1];
    return JTK_OK;
}

int jtk_batch_colsums(jtk_batch *b, int take_num, double *out, const uint64_t *stat_off) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (!out || !stat_off || take_num < 0) return ctx->fail(JTK_EINVAL, "bad argument");
    if (!b->has_profiles) return ctx->fail(JTK_ESTATE, "jtk_batch_modtable has not run");
    if (b->n_tmpl == 0) return JTK_OK;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    uint64_t total = 0;
    for (int t = 0; t < b->n_tmpl; t++) total = std::max<uint64_t>(total, stat_off[t] + (uint64_t)(b->tmpl_len[t] + 1) * kNumRow);
    CU(b->d_stat_off.reserve((size_t)b->n_tmpl), "cudaMalloc stat_off");
    CU(ctx->d_gather.reserve((size_t)total), "cudaMalloc sums");
    CU(b->up_stat_off.put(b->d_stat_off.p, stat_off, sizeof(uint64_t) * (size_t)b->n_tmpl, st), "H2D stat_off");
    if (b->n_tmpl > 65535) return ctx->fail(JTK_EINVAL, "more than 65535 templates in one batch (the per-chunk kernels put the template on grid.y): split the call");
    dim3 grid((unsigned)(((size_t)(b->max_lt + 1) * kNumRow + 255) / 256), (unsigned)b->n_tmpl);
    colsums_kernel<<<grid, 256, 0, st>>>(b->d_delta.p, b->d_pairs.p, b->d_tp_start.p, b->d_tp_ids.p, b->d_tmpl_len.p,
                                        b->d_stat_off.p, take_num, ctx->d_gather.p);
    CU(cudaGetLastError(), "colsums launch");
    ctx->launches++;
    CU(cudaMemcpyAsync(out, ctx->d_gather.p, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, st), "D2H sums");
    return batch_sync(b);
}

int jtk_batch_best_edits(jtk_batch *b, int take_num, int ignore_edge, double min_gain, int8_t *out, double *out_gain,
                         const uint64_t *col_off) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (!out || !col_off || take_num < 0 || ignore_edge < 0) return ctx->fail(JTK_EINVAL, "bad argument");
    if (!b->has_profiles) return ctx->fail(JTK_ESTATE, "jtk_batch_modtable has not run");
    if (b->n_tmpl == 0) return JTK_OK;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    uint64_t total = 0;
    for (int t = 0; t < b->n_tmpl; t++) total = std::max<uint64_t>(total, col_off[t] + (uint64_t)b->tmpl_len[t] + 1);
    CU(b->d_stat_off.reserve((size_t)b->n_tmpl), "cudaMalloc col_off");
    CU(ctx->d_mc_asn.reserve((size_t)total), "cudaMalloc best edits");
    if (out_gain) CU(ctx->d_gather.reserve((size_t)total), "cudaMalloc best gains");
    CU(b->up_stat_off.put(b->d_stat_off.p, col_off, sizeof(uint64_t) * (size_t)b->n_tmpl, st), "H2D col_off");
    if (b->n_tmpl > 65535) return ctx->fail(JTK_EINVAL, "more than 65535 templates in one batch (the per-chunk kernels put the template on grid.y): split the call");
    dim3 grid((unsigned)((b->max_lt + 1 + 127) / 128), (unsigned)b->n_tmpl);
    best_edit_kernel<<<grid, 128, 0, st>>>(b->d_delta.p, b->d_pairs.p, b->d_tp_start.p, b->d_tp_ids.p, b->d_tmpl_len.p, b->d_codes.p,
                                           b->d_tmpl_code_off.p, b->d_stat_off.p, take_num, ignore_edge, min_gain,
                                           reinterpret_cast<int8_t *>(ctx->d_mc_asn.p), out_gain ? ctx->d_gather.p : nullptr);
    CU(cudaGetLastError(), "best_edit launch");
    ctx->launches++;
    CU(cudaMemcpyAsync(out, ctx->d_mc_asn.p, (size_t)total, cudaMemcpyDeviceToHost, st), "D2H best edits");
    if (out_gain) CU(cudaMemcpyAsync(out_gain, ctx->d_gather.p, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, st), "D2H best gains");
    return batch_sync(b);
}

int jtk_batch_expected_counts(jtk_batch *b, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, double *acc90) {
    if (!b) return JTK_EINVAL;
    jtk_ctx *ctx = b->ctx;
    if (!fwd || !rev || !acc90) return ctx->fail(JTK_EINVAL, "null argument");
    std::memset(acc90, 0, sizeof(double) * 90);
    if (b->n_pairs == 0) return JTK_OK;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    float models[2 * kModelFloats];
    pack_model(fwd, models);
    pack_model(rev, models + kModelFloats);
    const int wpc = warps_per_cta();
    int grid = (b->n_pairs + wpc - 1) / wpc;
    const int max_grid = ctx->sm_count * 4;
    if (grid > max_grid) grid = max_grid;
    cudaStream_t st = ctx->stream;
    KParams kp{};
    const size_t slots = (size_t)grid * wpc;
    kp.frow_stride = (size_t)2 * (b->max_nd + 1) * 32 * b->C; // one float4 (F_M, F_I, F_D, -) per slot and row
    kp.kf_stride = (size_t)b->max_nd + 6;
    CU(ctx->d_models.reserve(2 * kModelFloats), "cudaMalloc models");
    CU(ctx->d_counter.reserve(1), "cudaMalloc counter");
    CU(ctx->d_frows.reserve(slots * kp.frow_stride), "cudaMalloc forward rows");
    CU(ctx->d_kf.reserve(slots * kp.kf_stride), "cudaMalloc scale exponents");
    CU(ctx->d_gather.reserve(90), "cudaMalloc counts");
    CU(ctx->up_models.put(ctx->d_models.p, models, sizeof(models), st), "H2D models");
    CU(cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(int), st), "memset counter");
    CU(cudaMemsetAsync(ctx->d_gather.p, 0, sizeof(double) * 90, st), "memset counts");
    kp.pairs = b->d_pairs.p; kp.n_pairs = b->n_pairs;
    kp.codes = b->d_codes.p; kp.bits = b->d_bits.p; kp.models = ctx->d_models.p;
    kp.radius = b->radius; kp.rows = 0;
    kp.frows = ctx->d_frows.p; kp.kf = ctx->d_kf.p;
    kp.out_delta = nullptr; kp.out_lk = b->d_lk.p; kp.counter = ctx->d_counter.p;
    CU(launch_fit(kp, b->C, grid, ctx->d_gather.p, st), "fit kernel launch");
    ctx->launches++;
    CU(cudaMemcpyAsync(acc90, ctx->d_gather.p, sizeof(double) * 90, cudaMemcpyDeviceToHost, st), "D2H counts");
    return batch_sync(b);
}

static void mstep(jtk_hmm_params *h, const double *acc) {
    double *tr[9] = { &h->mat_mat, &h->mat_ins, &h->mat_del, &h->ins_mat, &h->ins_ins, &h->ins_del,
                      &h->del_mat, &h->del_ins, &h->del_del };
    for (int st = 0; st < 3; st++) {
        const double tot = acc[3 * st] + acc[3 * st + 1] + acc[3 * st + 2];
        if (tot > 0) for (int k = 0; k < 3; k++) *tr[3 * st + k] = acc[3 * st + k] / tot;
    }
    for (int r = 0; r < 4; r++) {
        double tot = 0;
        for (int x = 0; x < 4; x++) tot += acc[9 + 4 * r + x];
        if (tot > 0) for (int x = 0; x < 4; x++) h->mat_emit[4 * r + x] = acc[9 + 4 * r + x] / tot;
    }
    for (int c = 0; c < 5; c++) {
        double tot = 0;
        for (int x = 0; x < 4; x++) tot += acc[25 + 4 * c + x];
        if (tot > 0) for (int x = 0; x < 4; x++) h->ins_emit[4 * c + x] = acc[25 + 4 * c + x] / tot;
    }
}

int jtk_hmm_fit_batch(jtk_ctx *ctx, jtk_hmm_params *fwd, jtk_hmm_params *rev, int n_pairs, int n_tmpl,
                      const uint8_t *tmpl_concat, const uint32_t *tmpl_off, const uint8_t *read_concat,
                      const uint32_t *read_off, const uint8_t *ops_concat, const uint32_t *ops_off,
                      const uint8_t *strand, const uint32_t *tmpl_idx, int radius) {
    if (!ctx) return JTK_EINVAL;
    if (!fwd || !rev) return ctx->fail(JTK_EINVAL, "null model");
    jtk_batch *b = nullptr;
    int rc = batch_create(ctx, n_pairs, n_tmpl, tmpl_concat, tmpl_off, read_concat, read_off, ops_concat, ops_off, strand,
                          tmpl_idx, radius, false, &b);
    if (rc) return rc;
    double acc[90];
    rc = jtk_batch_expected_counts(b, fwd, rev, acc);
    jtk_batch_destroy(b);
    if (rc) return rc;
    mstep(fwd, acc);      // strand 1 reads -> model 0 (forward)
    mstep(rev, acc + 45); // strand 0 reads -> model 1 (reverse)
    return JTK_OK;
}

// ---- level 1 on top of the batch ---------------------------------------------------------------------
int jtk_hmm_modtable_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_pairs,
                           int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                           const uint8_t *read_concat, const uint32_t *read_off,
                           const uint8_t *ops_concat, const uint32_t *ops_off,
                           const uint8_t *strand, const uint32_t *tmpl_idx, int radius,
                           double *out_lk, double *out_table, const uint64_t *table_off) {
    if (!ctx) return JTK_EINVAL;
    if (!out_lk) return ctx->fail(JTK_EINVAL, "out_lk is NULL");
    if (out_table && !table_off) return ctx->fail(JTK_EINVAL, "table_off is NULL");
    jtk_batch *b = nullptr;
    int rc = batch_create(ctx, n_pairs, n_tmpl, tmpl_concat, tmpl_off, read_concat, read_off, ops_concat, ops_off, strand,
                          tmpl_idx, radius, false, &b);
    if (rc) return rc;
    rc = batch_run(b, fwd, rev, true, 14);
    if (!rc) rc = jtk_batch_fetch_lk(b, out_lk);
    if (!rc && out_table && n_pairs > 0) {
        cudaError_t e = ctx->h_delta.reserve((size_t)b->table_floats);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(ctx->h_delta.p, b->d_delta.p, sizeof(float) * (size_t)b->table_floats, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = ctx->cuda_fail(e, "D2H table");
        else
            for (int p = 0; p < n_pairs; p++) {
                const DevPair &dp = b->pairs[(size_t)p];
                const float *src = ctx->h_delta.p + dp.tab_off;
                double *dst = out_table + table_off[p];
                const double lk = out_lk[p];
                const size_t n = (size_t)(dp.Lt + 1) * kNumRow;
                for (size_t k = 0; k < n; k++)
                    dst[k] = (src[k] <= -1.0e9f || !(lk > -INFINITY)) ? JTK_TABLE_NEG : lk + (double)src[k];
            }
    }
    jtk_batch_destroy(b);
    return rc;
}

int jtk_hmm_likelihood_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_pairs,
                             int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                             const uint8_t *read_concat, const uint32_t *read_off,
                             const uint8_t *ops_concat, const uint32_t *ops_off,
                             const uint8_t *strand, const uint32_t *tmpl_idx, int radius, double *out_lk) {
    if (!ctx) return JTK_EINVAL;
    if (!out_lk) return ctx->fail(JTK_EINVAL, "out_lk is NULL");
    jtk_batch *b = nullptr;
    int rc = batch_create(ctx, n_pairs, n_tmpl, tmpl_concat, tmpl_off, read_concat, read_off, ops_concat, ops_off, strand,
                          tmpl_idx, radius, true, &b);
    if (rc) return rc;
    rc = batch_run(b, fwd, rev, false, 0);
    if (!rc) rc = jtk_batch_fetch_lk(b, out_lk);
    jtk_batch_destroy(b);
    return rc;
}

} // extern "C"
