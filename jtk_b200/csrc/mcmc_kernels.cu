// mcmc_kernels.cu -- the k-means + MCMC restarts of pseudo_mcmc::mcmc_clustering on the GPU, one chain per warp.
//
// Reference: /root/reference/haplotyper/src/local_clustering/pseudo_mcmc.rs:649-670 (mcmc_clustering: 20 restarts of
// misc::kmeans + mcmc_with_filter on ONE generator), :704-762 (mcmc_with_filter), :764-795 (flip, get_lk), :797-869
// (LKCount, get_used_columns); misc.rs:231-341 (kmeans, suggest_first, update_assignments, get_dist).
//
// Why: once the pair-HMM is on the GPU, the 2.4 M sequential proposals per chunk (20 restarts x 2000 x n reads) are the
// cost of phasing a chunk (0.1 s of one host core; 72 % of local_clustering_selected at 640 chunks).
// The chain of one chunk cannot be split -- all restarts draw from one Xoshiro256** stream and every decision feeds the
// next -- but chunks are independent: every chunk gets one warp (all lanes run the scalar part of the chain redundantly,
// lane d owns variant column d and keeps its statistics in registers), so thousands of chains run side by side at the
// latency of one.  A first version with the chain state in global memory took 23 s per chain: every load after a store
// went to L2.
//
// Parity: the code below is the device twin of the host restatement in local_clustering.cpp, statement by statement: the
// same generator (rand 0.8.5 sampling as published), the same f64 operations in the same order (explicit round-to-nearest
// intrinsics, no FMA contraction).  The one library call is exp() in the acceptance test; CUDA's exp is within 1 ulp of
// glibc's, which moves the 64-bit Bernoulli threshold by < 2^12 of 2^64: a decision differs with probability < 2^-52 per
// draw.  tests/test_gpu_clustering.py compares assignments, scores and generator states with the host twin.
#include "../../include/jtk_gpu.h"
#include "mcmc_dev.cuh"

#include <cstdint>
#include <cuda_runtime.h>

namespace jtk {

constexpr double kPosThr = 0.00001; // pseudo_mcmc.rs:5

struct DevRng { // rand_xoshiro::Xoshiro256StarStar + the rand 0.8.5 samplers used by the reference
    uint64_t s0, s1, s2, s3;
    __device__ __forceinline__ static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    __device__ __forceinline__ uint64_t next_u64() {
        const uint64_t result = rotl(s1 * 5, 7) * 9;
        const uint64_t t = s1 << 17;
        s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3;
        s2 ^= t;
        s3 = rotl(s3, 45);
        return result;
    }
    __device__ __forceinline__ uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    __device__ __forceinline__ uint64_t gen_range(uint64_t range) { // UniformInt<usize>::sample_single_inclusive(0, n-1)
        const uint64_t zone = (range << __clzll((long long)range)) - 1;
        const uint32_t r32 = (uint32_t)range; // every range here (reads, clusters) fits 32 bits: two 32 x 32 -> 64 products
        for (;;) {
            const uint64_t v = next_u64();
            uint64_t lo, hi;
            if (range <= 0xffffffffULL) {
                const uint64_t t = (uint64_t)(uint32_t)v * r32;
                const uint64_t u = (uint64_t)(uint32_t)(v >> 32) * r32 + (t >> 32);
                lo = (u << 32) | (uint32_t)t; hi = u >> 32;
            } else { lo = v * range; hi = __umul64hi(v, range); }
            if (lo <= zone) return hi;
        }
    }
    __device__ __forceinline__ uint64_t gen_index(uint64_t ubound) { // rand::seq::gen_index
        if (ubound <= 0xffffffffULL) {
            const uint32_t range = (uint32_t)ubound;
            const uint32_t zone = (range << __clz((int)range)) - 1;
            for (;;) {
                const uint32_t v = next_u32();
                const uint64_t m = (uint64_t)v * range;
                if ((uint32_t)m <= zone) return m >> 32;
            }
        }
        return gen_range(ubound);
    }
    // Bernoulli::new(p).sample for p in [0, 1); the caller handles p == 1 and invalid p
    __device__ __forceinline__ bool bernoulli(double p) {
        const uint64_t p_int = __double2ull_rz(__dmul_rn(p, 18446744073709551616.0));
        return next_u64() < p_int;
    }
    __device__ __forceinline__ double uniform0(double total) { // Uniform<f64>::new(0, total).sample
        const uint64_t bits = (next_u64() >> 12) | (1023ULL << 52);
        const double v12 = __longlong_as_double((long long)bits);
        return __dadd_rn(__dmul_rn(__dadd_rn(v12, -1.0), total), 0.0);
    }
    __device__ __forceinline__ uint32_t choose_other(uint32_t k, uint32_t old) { // IteratorRandom::choose, one draw per element
        uint32_t consumed = 0, result = 0xffffffffu;
        for (uint32_t x = 0; x < k; x++) {
            if (x == old) continue;
            consumed++;
            if (gen_index(consumed) == 0) result = x;
        }
        return result;
    }
};

// error codes left in out_err (0 = ok); the host turns them into the reference's panics
enum { kMcmcOk = 0, kMcmcKmeansDiverged = 1, kMcmcBadProb = 2, kMcmcLkMismatch = 3, kMcmcBadWeights = 4, kMcmcNoOther = 5 };

constexpr unsigned kFullMask = 0xffffffffu;

struct ChainView {
    uint32_t n, D, k;
    const double *flat; const double *size_to_lk;          // global, read-only
    const uint8_t *pinc, *ninc, *ratio_ok;                   // global, written once at start
    double *centers, *dists, *cum; uint32_t *counts;         // shared: k-means scratch (lane 0)
    uint8_t *assign, *argmax, *best;                         // shared: assignments
};

__device__ __forceinline__ double dist2(const double *x, const double *y, uint32_t D) { // sum (x - y).powi(2)
    double s = 0.0;
    for (uint32_t d = 0; d < D; d++) { const double t = __dsub_rn(x[d], y[d]); s = __dadd_rn(s, __dmul_rn(t, t)); }
    return s;
}

__device__ void update_assignments(const ChainView &c, uint32_t n_centers) { // misc.rs: min_by keeps the first minimum
    for (uint32_t i = 0; i < c.n; i++) {
        uint32_t best = 0; double bd = 0.0;
        for (uint32_t m = 0; m < n_centers; m++) {
            const double d = dist2(c.flat + (size_t)i * c.D, c.centers + (size_t)m * c.D, c.D);
            if (m == 0 || d < bd) { best = m; bd = d; }
        }
        c.assign[i] = (uint8_t)best;
    }
}

__device__ double get_dist(const ChainView &c) {
    double s = 0.0;
    for (uint32_t i = 0; i < c.n; i++) s = __dadd_rn(s, dist2(c.flat + (size_t)i * c.D, c.centers + (size_t)c.assign[i] * c.D, c.D));
    return s;
}

// misc::kmeans (misc.rs:231-341), run by ONE lane: leaves the assignment in c.assign
__device__ int kmeans(const ChainView &c, DevRng &rng) {
    const uint32_t n = c.n, D = c.D, k = c.k;
    if (rng.bernoulli(0.5)) {
        for (uint32_t i = 0; i < n; i++) c.assign[i] = (uint8_t)rng.gen_range(k);
    } else { // suggest_first: k-means++ seeding, centres are data rows
        const uint32_t first = (uint32_t)rng.gen_index(n);
        for (uint32_t d = 0; d < D; d++) c.centers[d] = c.flat[(size_t)first * D + d];
        for (uint32_t it = 0; it + 1 < k; it++) {
            const uint32_t nc = it + 1;
            for (uint32_t i = 0; i < n; i++) {
                double m = 0.0;
                for (uint32_t q = 0; q < nc; q++) {
                    const double d = dist2(c.flat + (size_t)i * D, c.centers + (size_t)q * D, D);
                    if (q == 0 || d < m) m = d;
                }
                c.dists[i] = m;
            }
            // SliceRandom::choose_weighted: cumulative weights, Uniform(0, total), first cumulative weight > chosen
            double total = c.dists[0];
            if (!(total >= 0.0)) return kMcmcBadWeights;
            for (uint32_t i = 1; i < n; i++) {
                if (!(c.dists[i] >= 0.0)) return kMcmcBadWeights;
                c.cum[i - 1] = total;
                total = __dadd_rn(total, c.dists[i]);
            }
            if (total == 0.0) return kMcmcBadWeights;
            const double chosen = rng.uniform0(total);
            uint32_t lo = 0, hi = n - 1;
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2;
                if (c.cum[mid] <= chosen) lo = mid + 1; else hi = mid;
            }
            for (uint32_t d = 0; d < D; d++) c.centers[(size_t)nc * D + d] = c.flat[(size_t)lo * D + d];
        }
        update_assignments(c, k);
    }
    for (uint32_t m = 0; m < k * D; m++) c.centers[m] = 0.0;
    for (uint32_t m = 0; m < k; m++) c.counts[m] = 0;
    double dist = get_dist(c);
    for (;;) {
        for (uint32_t m = 0; m < k * D; m++) c.centers[m] = 0.0;
        for (uint32_t m = 0; m < k; m++) c.counts[m] = 0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t a = c.assign[i];
            for (uint32_t d = 0; d < D; d++) c.centers[(size_t)a * D + d] = __dadd_rn(c.centers[(size_t)a * D + d], c.flat[(size_t)i * D + d]);
            c.counts[a]++;
        }
        for (uint32_t m = 0; m < k; m++)
            if (0 < c.counts[m])
                for (uint32_t d = 0; d < D; d++) c.centers[(size_t)m * D + d] = __ddiv_rn(c.centers[(size_t)m * D + d], (double)c.counts[m]);
        update_assignments(c, k);
        const double nd = get_dist(c);
        if (!(nd < __dadd_rn(dist, 0.00000001))) return kMcmcKmeansDiverged;
        if (__dsub_rn(dist, nd) < 0.00000001) break;
        dist = nd;
    }
    return kMcmcOk;
}

// ------------------------------------------------------------------------------------------------
// mcmc_with_filter (:704-762), the whole warp on one chain.  Every lane runs the scalar part of the chain redundantly (the
// generator, the acceptance test: same integers and the same f64 operations in every lane, so the lanes never disagree and
// nothing has to be broadcast); lane d additionally owns variant column d: its LKCount statistics (total gain, positive /
// negative counts) for all clusters sit in registers.  A flip is then a handful of register operations per lane, and
// get_lk's sum runs over the lanes' terms in the reference's order (cluster-major, column-minor) through shuffles.
// ------------------------------------------------------------------------------------------------
template <int K> struct ColStats { double tot[K]; uint32_t np[K], nn[K]; };

__device__ __forceinline__ double shfl_f64(double v, int src) {
    const long long b = __double_as_longlong(v);
    const int lo = __shfl_sync(kFullMask, (int)(b & 0xffffffffLL), src), hi = __shfl_sync(kFullMask, (int)(b >> 32), src);
    return __longlong_as_double(((long long)hi << 32) | (unsigned)lo);
}

template <int K>
__device__ __forceinline__ void stats_add(ColStats<K> &s, uint32_t (&clus)[K], uint32_t a, double x, uint32_t pi, uint32_t ni, bool col) {
#pragma unroll
    for (int m = 0; m < K; m++) {
        const bool hit = (uint32_t)m == a;
        clus[m] += hit ? 1u : 0u;
        if (hit && col) { s.tot[m] = __dadd_rn(s.tot[m], x); s.np[m] += pi; s.nn[m] += ni; }
    }
}
template <int K>
__device__ __forceinline__ void stats_sub(ColStats<K> &s, uint32_t (&clus)[K], uint32_t a, double x, uint32_t pi, uint32_t ni, bool col) {
#pragma unroll
    for (int m = 0; m < K; m++) {
        const bool hit = (uint32_t)m == a;
        clus[m] -= hit ? 1u : 0u;
        if (hit && col) { s.tot[m] = __dsub_rn(s.tot[m], x); s.np[m] -= pi; s.nn[m] -= ni; }
    }
}

template <int K>
__device__ void build_stats(const ChainView &c, ColStats<K> &s, uint32_t (&clus)[K], int lane) {
    const bool col = (uint32_t)lane < c.D;
#pragma unroll
    for (int m = 0; m < K; m++) { s.tot[m] = 0.0; s.np[m] = 0; s.nn[m] = 0; clus[m] = 0; }
    for (uint32_t i = 0; i < c.n; i++) {
        const size_t e = (size_t)i * c.D + (col ? lane : 0);
        stats_add(s, clus, c.assign[i], c.flat[e], c.pinc[e], c.ninc[e], col);
    }
}

// get_lk (:785-795) with get_used_columns (:847-869); identical value in every lane
template <int K>
__device__ __forceinline__ double current_lk(const ChainView &c, const ColStats<K> &s, const uint32_t (&clus)[K], int lane) {
    const uint32_t D = c.D, n1 = c.n + 1;
    unsigned u = 0;
    uint32_t in_use = 0, in_neg = 0;
#pragma unroll
    for (int m = 0; m < K; m++) {
        const double g = s.tot[m];
        const uint32_t p = s.np[m];
        const bool pos = 0.0 < g;
        u |= (unsigned)pos & c.ratio_ok[(size_t)p * n1 + s.nn[m]];
        in_use += pos ? p : 0u;
        in_neg += (g <= 0.0) ? p : 0u;
    }
    const bool use = (uint32_t)lane < D && (u & (unsigned)(__dmul_rn((double)in_neg, 2.0) < (double)in_use)) != 0u;
    double lk = 0.0;
#pragma unroll
    for (int m = 0; m < K; m++) lk = __dadd_rn(lk, c.size_to_lk[clus[m]]);
#pragma unroll
    for (int m = 0; m < K; m++) {
        // a column that is not in use adds +0.0: lk is never -0.0 here (size_to_lk < 0), so the sum is the reference's
        const double term = use ? fmax(s.tot[m], 0.0) : 0.0;
        for (uint32_t d = 0; d < D; d++) lk = __dadd_rn(lk, shfl_f64(term, (int)d));
    }
    return lk;
}

template <int K>
__device__ __forceinline__ int mcmc_with_filter(const ChainView &c, DevRng &rng, double &out_lk, int lane) {
    const uint32_t n = c.n, k = K;
    const bool col = (uint32_t)lane < c.D;
    ColStats<K> s; uint32_t clus[K];
    build_stats(c, s, clus, lane);
    double lk = current_lk(c, s, clus, lane);
    double mx = lk;
    for (uint32_t i = lane; i < n; i += 32) c.argmax[i] = c.assign[i];
    __syncwarp();
    const uint64_t total = 2000ull * n;
    for (uint64_t t = 0; t < total; t++) {
        const uint32_t idx = (uint32_t)rng.gen_range(n);
        const uint32_t old = c.assign[idx];
        const uint32_t nw = rng.choose_other(k, old);
        if (nw == 0xffffffffu) return kMcmcNoOther;
        const size_t e = (size_t)idx * c.D + (col ? lane : 0);
        const double x = c.flat[e];
        const uint32_t pi = c.pinc[e], ni = c.ninc[e];
        __syncwarp();                                  // every lane has read assign[idx]
        stats_sub(s, clus, old, x, pi, ni, col);       // flip (:764-783)
        stats_add(s, clus, nw, x, pi, ni, col);
        const double proposed = current_lk(c, s, clus, lane);
        const double diff = __dsub_rn(proposed, lk);
        bool accept;
        if (0.0 < diff) accept = true;
        else if (diff < -45.0) { (void)rng.next_u64(); accept = false; } // threshold 0: the draw is consumed, never accepted
        else {
            const double p = exp(diff);
            if (p == 1.0) accept = true;                       // Bernoulli ALWAYS_TRUE: no draw
            else if (!(p >= 0.0 && p < 1.0)) return kMcmcBadProb; // NaN: the reference panics
            else accept = rng.bernoulli(p);
        }
        if (accept) {
            if (lane == 0) c.assign[idx] = (uint8_t)nw;
            __syncwarp();
            lk = proposed;
            if (mx < lk) {
                mx = proposed;
                for (uint32_t i = lane; i < n; i += 32) c.argmax[i] = c.assign[i];
                __syncwarp();
            }
        } else { // flip back: the same subtraction / addition the reference performs
            stats_sub(s, clus, nw, x, pi, ni, col);
            stats_add(s, clus, old, x, pi, ni, col);
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) c.assign[i] = c.argmax[i];
    __syncwarp();
    build_stats(c, s, clus, lane);
    const double chk = current_lk(c, s, clus, lane);
    if (!(fabs(__dsub_rn(mx, chk)) < 0.0001)) return kMcmcLkMismatch;
    out_lk = mx;
    return kMcmcOk;
}

// One warp per chain.  rng_state: 4 words per chain, in/out.  out_asn: best assignment (bytes, n per chain at asn_off),
// out_lk: its likelihood, out_err: 0 or the first failure.  Dynamic shared memory: smem_per_chain bytes per warp.
__global__ void __launch_bounds__(128, 4) mcmc_restarts_kernel(const McmcChain *__restrict__ chains, int n_chains, double *wf64,
                                                            uint8_t *wu8, uint64_t *rng_state, uint8_t *out_asn,
                                                            const uint64_t *__restrict__ asn_off, double *out_lk, int *out_err,
                                                            int restarts, int smem_per_chain) {
    extern __shared__ __align__(16) unsigned char mcmc_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + warp;
    if (chain >= n_chains) return;
    const McmcChain ch = chains[chain];
    ChainView c;
    c.n = ch.n; c.D = ch.D; c.k = ch.k;
    const double *f = wf64 + ch.off_f64;
    c.flat = f; c.size_to_lk = f + (size_t)c.n * c.D;
    uint8_t *b = wu8 + ch.off_u8;
    uint8_t *pinc = b, *ninc = b + (size_t)c.n * c.D, *ratio_ok = b + 2 * (size_t)c.n * c.D;
    c.pinc = pinc; c.ninc = ninc; c.ratio_ok = ratio_ok;
    unsigned char *sm = mcmc_smem + (size_t)warp * smem_per_chain;
    c.centers = reinterpret_cast<double *>(sm); sm += sizeof(double) * c.k * c.D;
    c.dists = reinterpret_cast<double *>(sm); sm += sizeof(double) * c.n;
    c.cum = reinterpret_cast<double *>(sm); sm += sizeof(double) * c.n;
    c.counts = reinterpret_cast<uint32_t *>(sm); sm += sizeof(uint32_t) * ((c.k + 1) & ~1u);
    c.assign = sm; sm += c.n;
    c.argmax = sm; sm += c.n;
    c.best = sm;
    // sign classes of the entries (LKCount::add / sub) and the is_informative ratio table, as the host twin builds them
    for (size_t i = lane; i < (size_t)c.n * c.D; i += 32) {
        const double x = c.flat[i];
        pinc[i] = kPosThr < x ? 1 : 0;
        ninc[i] = (!(kPosThr < x) && x < -kPosThr) ? 1 : 0;
    }
    const uint32_t n1 = c.n + 1;
    for (uint32_t p = lane; p <= c.n; p += 32)
        for (uint32_t q = 0; q + p <= c.n; q++)
            ratio_ok[(size_t)p * n1 + q] = 0.70 < __ddiv_rn((double)p, __dadd_rn((double)(p + q), 0.0000001));
    __syncwarp();
    DevRng rng{ rng_state[4 * chain], rng_state[4 * chain + 1], rng_state[4 * chain + 2], rng_state[4 * chain + 3] };
    double best_lk = 0.0;
    bool any = false;
    int err = kMcmcOk;
    for (int t = 0; t < restarts && err == kMcmcOk; t++) { // mcmc_clustering (:649-670): max_by keeps the last maximum
        if (lane == 0) err = kmeans(c, rng);          // sequential and small: one lane, then every lane takes its generator
        __syncwarp();
        err = __shfl_sync(kFullMask, err, 0);
        rng.s0 = __shfl_sync(kFullMask, rng.s0, 0); rng.s1 = __shfl_sync(kFullMask, rng.s1, 0);
        rng.s2 = __shfl_sync(kFullMask, rng.s2, 0); rng.s3 = __shfl_sync(kFullMask, rng.s3, 0);
        if (err != kMcmcOk) break;
        double lk = 0.0;
        switch (c.k) { // one instantiation per cluster number: the per-cluster loops are unrolled over registers
        case 1: err = mcmc_with_filter<1>(c, rng, lk, lane); break;
        case 2: err = mcmc_with_filter<2>(c, rng, lk, lane); break;
        case 3: err = mcmc_with_filter<3>(c, rng, lk, lane); break;
        case 4: err = mcmc_with_filter<4>(c, rng, lk, lane); break;
        case 5: err = mcmc_with_filter<5>(c, rng, lk, lane); break;
        case 6: err = mcmc_with_filter<6>(c, rng, lk, lane); break;
        case 7: err = mcmc_with_filter<7>(c, rng, lk, lane); break;
        default: err = mcmc_with_filter<8>(c, rng, lk, lane); break;
        }
        if (err != kMcmcOk) break;
        if (!any || !(lk < best_lk)) { for (uint32_t i = lane; i < c.n; i += 32) c.best[i] = c.assign[i]; best_lk = lk; any = true; }
        __syncwarp();
    }
    __syncwarp();
    if (lane == 0) {
        rng_state[4 * chain] = rng.s0; rng_state[4 * chain + 1] = rng.s1; rng_state[4 * chain + 2] = rng.s2; rng_state[4 * chain + 3] = rng.s3;
        out_lk[chain] = best_lk;
        out_err[chain] = err;
    }
    uint8_t *oa = out_asn + asn_off[chain];
    for (uint32_t i = lane; i < c.n; i += 32) oa[i] = c.best[i];
}

size_t mcmc_smem_bytes(uint32_t n, uint32_t D, uint32_t k) {
    return (sizeof(double) * ((size_t)k * D + 2 * (size_t)n) + sizeof(uint32_t) * ((k + 1) & ~1u) + 3 * (size_t)n + 15) & ~(size_t)15;
}

cudaError_t launch_mcmc_restarts(const McmcChain *chains, int n_chains, double *wf64, uint8_t *wu8, uint64_t *rng_state,
                                 uint8_t *out_asn, const uint64_t *asn_off, double *out_lk, int *out_err, int restarts,
                                 size_t smem_per_chain, cudaStream_t st) {
    if (n_chains <= 0) return cudaSuccess;
    int warps = 4;
    while (warps > 1 && warps * smem_per_chain > 200 * 1024) warps >>= 1;
    const size_t dyn = warps * smem_per_chain;
    cudaError_t e = cudaFuncSetAttribute(mcmc_restarts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    mcmc_restarts_kernel<<<(n_chains + warps - 1) / warps, warps * 32, dyn, st>>>(chains, n_chains, wf64, wu8, rng_state, out_asn, asn_off,
                                                                                   out_lk, out_err, restarts, (int)smem_per_chain);
    return cudaGetLastError();
}

} // namespace jtk
