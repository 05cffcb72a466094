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

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
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
    uint32_t ld;                                             // row stride of flat (doubles): D, or the padded width of the sub-warp kernel
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
            const double d = dist2(c.flat + (size_t)i * c.ld, c.centers + (size_t)m * c.D, c.D);
            if (m == 0 || d < bd) { best = m; bd = d; }
        }
        c.assign[i] = (uint8_t)best;
    }
}

__device__ double get_dist(const ChainView &c) {
    double s = 0.0;
    for (uint32_t i = 0; i < c.n; i++) s = __dadd_rn(s, dist2(c.flat + (size_t)i * c.ld, c.centers + (size_t)c.assign[i] * c.D, c.D));
    return s;
}

// misc::kmeans (misc.rs:231-341), run by ONE lane: leaves the assignment in c.assign
__device__ int kmeans(const ChainView &c, DevRng &rng) {
    const uint32_t n = c.n, D = c.D, k = c.k;
    if (rng.bernoulli(0.5)) {
        for (uint32_t i = 0; i < n; i++) c.assign[i] = (uint8_t)rng.gen_range(k);
    } else { // suggest_first: k-means++ seeding, centres are data rows
        const uint32_t first = (uint32_t)rng.gen_index(n);
        for (uint32_t d = 0; d < D; d++) c.centers[d] = c.flat[(size_t)first * c.ld + d];
        for (uint32_t it = 0; it + 1 < k; it++) {
            const uint32_t nc = it + 1;
            for (uint32_t i = 0; i < n; i++) {
                double m = 0.0;
                for (uint32_t q = 0; q < nc; q++) {
                    const double d = dist2(c.flat + (size_t)i * c.ld, c.centers + (size_t)q * D, D);
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
            for (uint32_t d = 0; d < D; d++) c.centers[(size_t)nc * D + d] = c.flat[(size_t)lo * c.ld + d];
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
            for (uint32_t d = 0; d < D; d++) c.centers[(size_t)a * D + d] = __dadd_rn(c.centers[(size_t)a * D + d], c.flat[(size_t)i * c.ld + d]);
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
__global__ void __launch_bounds__(128, 4) mcmc_restarts_kernel(const McmcChain *__restrict__ chains, const int *__restrict__ ids, int n_chains, double *wf64,
                                                            uint8_t *wu8, uint64_t *rng_state, uint8_t *out_asn,
                                                            const uint64_t *__restrict__ asn_off, double *out_lk, int *out_err,
                                                            int restarts, int smem_per_chain) {
    extern __shared__ __align__(16) unsigned char mcmc_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
    if (slot >= n_chains) return;
    const int chain = ids[slot];
    const McmcChain ch = chains[chain];
    ChainView c;
    c.n = ch.n; c.D = ch.D; c.k = ch.k; c.ld = ch.D;
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

// ------------------------------------------------------------------------------------------------
// Two clusters, at most eight variant columns (every diploid chunk: D <= 3 * max(copy_num, 2) = 6): FOUR chains per warp,
// eight lanes each, everything a proposal touches in shared memory, and a proposal latency of a few hundred cycles instead
// of ~2 000.  Same stream, same decisions, same f64 sums in the same order as mcmc_with_filter<2> above / the host twin:
//   * lane g of a group owns column g (columns >= D hold +0.0 and are never in use: adding +0.0 is exact, so the sums run
//     over the padded width DP without a branch);
//   * flip = `tot0 += s; tot1 -= s` with s = -x or x by the old cluster (a - b == a + (-b) exactly); the reject path
//     replays the reference's flip back (the round trip (a - x) + x is not the identity in floating point);
//   * the per-column terms of get_lk go through a double-buffered shared-memory exchange: one group barrier per proposal,
//     then every lane adds the 2 + 2*DP terms in the reference's order;
//   * the acceptance draw `next_u64() < trunc(exp(diff) * 2^64)` is decided from a cheap bracket of exp(diff) (ex2.approx,
//     relative error < 1e-5, bracket +-1e-4) whenever the draw falls outside the bracket -- which is all but ~1e-4 of the
//     time; only then is the correctly rounded exp evaluated.  The decision is the reference's in every case.
// ------------------------------------------------------------------------------------------------
constexpr int kSubLanes = 8;
struct SubLayout { // byte offsets inside one chain's shared-memory block
    uint32_t x, cl, s2l, ratio, assign, argmax, best, xch, centers, dists, cum, counts, total, rwords;
};
__host__ __device__ inline SubLayout sub_layout(uint32_t n, uint32_t DP) {
    SubLayout L;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { const uint32_t at = o; o += (bytes + 15u) & ~15u; return at; };
    L.rwords = (n + 1 + 31) / 32;
    L.x = take(8u * n * DP);
    L.s2l = take(8u * (n + 1));
    L.xch = take(8u * 2 * 2 * DP);
    L.centers = take(8u * 2 * DP);
    L.dists = take(8u * n);
    L.cum = take(8u * n);
    L.ratio = take(4u * (n + 1) * L.rwords);
    L.counts = take(16);
    L.cl = take(n * DP);
    L.assign = take(n);
    L.argmax = take(n);
    L.best = take(n);
    L.total = o;
    return L;
}

template <int DP>
__global__ void __launch_bounds__(128, 2) mcmc_diploid_kernel(const McmcChain *__restrict__ chains, const int *__restrict__ ids,
                                                              int n_ids, const double *__restrict__ wf64, uint64_t *rng_state,
                                                              uint8_t *out_asn, const uint64_t *__restrict__ asn_off,
                                                              double *out_lk, int *out_err, int restarts, int smem_per_chain) {
    extern __shared__ __align__(16) unsigned char mcmc_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane & (kSubLanes - 1), grp = lane / kSubLanes, lead = grp * kSubLanes;
    const unsigned gmask = 0xffu << lead;
    const int slot = (blockIdx.x * (blockDim.x >> 5) + warp) * (32 / kSubLanes) + grp;
    if (slot >= n_ids) return; // the eight lanes of a group leave together
    const int chain = ids[slot];
    const McmcChain ch = chains[chain];
    const uint32_t n = ch.n, D = ch.D;
    unsigned char *sm = mcmc_smem + (size_t)(warp * (32 / kSubLanes) + grp) * smem_per_chain;
    const SubLayout L = sub_layout(n, DP);
    double *X = reinterpret_cast<double *>(sm + L.x);
    double *S2L = reinterpret_cast<double *>(sm + L.s2l);
    double *XCH = reinterpret_cast<double *>(sm + L.xch);
    uint32_t *RAT = reinterpret_cast<uint32_t *>(sm + L.ratio);
    uint8_t *CL = sm + L.cl;
    const uint32_t RW = L.rwords;
    ChainView c;
    c.n = n; c.D = D; c.k = 2; c.ld = DP;
    c.flat = X; c.size_to_lk = S2L; c.pinc = nullptr; c.ninc = nullptr; c.ratio_ok = nullptr;
    c.centers = reinterpret_cast<double *>(sm + L.centers);
    c.dists = reinterpret_cast<double *>(sm + L.dists);
    c.cum = reinterpret_cast<double *>(sm + L.cum);
    c.counts = reinterpret_cast<uint32_t *>(sm + L.counts);
    c.assign = sm + L.assign; c.argmax = sm + L.argmax; c.best = sm + L.best;
    { // stage the chain: data (padded to DP columns), sign classes, size prior, the is_informative ratio table as bits
        const double *f = wf64 + ch.off_f64;
        for (uint32_t e = g; e < n * DP; e += kSubLanes) {
            const uint32_t i = e / DP, d = e % DP;
            const double x = d < D ? f[(size_t)i * D + d] : 0.0;
            X[e] = x;
            CL[e] = (uint8_t)((kPosThr < x ? 1 : 0) | ((!(kPosThr < x) && x < -kPosThr) ? 2 : 0));
        }
        for (uint32_t i = g; i <= n; i += kSubLanes) S2L[i] = f[(size_t)n * D + i];
        for (uint32_t e = g; e < (n + 1) * RW; e += kSubLanes) {
            const uint32_t p = e / RW, w = e % RW;
            uint32_t bits = 0;
            for (uint32_t b = 0; b < 32; b++) {
                const uint32_t q = w * 32 + b;
                if (q + p <= n && 0.70 < __ddiv_rn((double)p, __dadd_rn((double)(p + q), 0.0000001))) bits |= 1u << b;
            }
            RAT[e] = bits;
        }
    }
    __syncwarp(gmask);
    DevRng rng{ rng_state[4 * chain], rng_state[4 * chain + 1], rng_state[4 * chain + 2], rng_state[4 * chain + 3] };
    const bool col = (uint32_t)g < D;
    auto ratio_ok = [&](uint32_t p, uint32_t q) -> unsigned { return (RAT[p * RW + (q >> 5)] >> (q & 31u)) & 1u; };
    // column statistics of the two clusters (lane g: column g) and the cluster sizes (every lane)
    double tot0, tot1; uint32_t np0, np1, nn0, nn1, c0, c1;
    auto build = [&]() {
        tot0 = tot1 = 0.0; np0 = np1 = nn0 = nn1 = c0 = c1 = 0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t a = c.assign[i];
            const double x = X[i * DP + g];
            const uint32_t cl = CL[i * DP + g];
            if (a == 0) { c0++; tot0 = __dadd_rn(tot0, x); np0 += cl & 1u; nn0 += cl >> 1; }
            else { c1++; tot1 = __dadd_rn(tot1, x); np1 += cl & 1u; nn1 += cl >> 1; }
        }
    };
    unsigned buf = 0;
    auto current_lk = [&]() -> double { // get_lk (:785-795) with get_used_columns (:847-869)
        const bool pos0 = 0.0 < tot0, pos1 = 0.0 < tot1;
        const unsigned u = ((unsigned)pos0 & ratio_ok(np0, nn0)) | ((unsigned)pos1 & ratio_ok(np1, nn1));
        const uint32_t in_use = (pos0 ? np0 : 0u) + (pos1 ? np1 : 0u);
        const uint32_t in_neg = ((tot0 <= 0.0) ? np0 : 0u) + ((tot1 <= 0.0) ? np1 : 0u);
        const bool use = col && u != 0u && 2u * in_neg < in_use; // the reference compares the same integers as f64
        double *T = XCH + buf * (2 * DP);
        buf ^= 1u;
        T[g] = use ? (tot0 < 0.0 ? 0.0 : tot0) : 0.0;
        T[DP + g] = use ? (tot1 < 0.0 ? 0.0 : tot1) : 0.0;
        __syncwarp(gmask);
        double lk = __dadd_rn(S2L[c0], S2L[c1]);
#pragma unroll
        for (int d = 0; d < 2 * DP; d++) lk = __dadd_rn(lk, T[d]);
        return lk;
    };
    double best_lk = 0.0;
    bool any = false;
    int err = kMcmcOk;
    for (int t = 0; t < restarts && err == kMcmcOk; t++) { // mcmc_clustering (:649-670): max_by keeps the last maximum
        if (g == 0) err = kmeans(c, rng);
        __syncwarp(gmask);
        err = __shfl_sync(gmask, err, lead);
        rng.s0 = __shfl_sync(gmask, rng.s0, lead); rng.s1 = __shfl_sync(gmask, rng.s1, lead);
        rng.s2 = __shfl_sync(gmask, rng.s2, lead); rng.s3 = __shfl_sync(gmask, rng.s3, lead);
        if (err != kMcmcOk) break;
        // ---- mcmc_with_filter (:704-762), k = 2 ----
        build();
        double lk = current_lk();
        double mx = lk;
        for (uint32_t i = g; i < n; i += kSubLanes) c.argmax[i] = c.assign[i];
        const uint64_t total = 2000ull * n;
        for (uint64_t it = 0; it < total; it++) {
            const uint32_t idx = (uint32_t)rng.gen_range(n);
            const uint32_t old = c.assign[idx];
            // choose_other(2, old): one candidate, gen_index(1) draws 32-bit words until one is <= 0x7fffffff
            while (rng.next_u32() > 0x7fffffffu) { }
            const double x = X[idx * DP + g];
            const uint32_t cl = CL[idx * DP + g];
            const double s = old == 0u ? -x : x;
            const int dp = old == 0u ? -(int)(cl & 1u) : (int)(cl & 1u);
            const int dn = old == 0u ? -(int)(cl >> 1) : (int)(cl >> 1);
            const int dc = old == 0u ? -1 : 1;
            tot0 = __dadd_rn(tot0, s); tot1 = __dadd_rn(tot1, -s);      // flip (:764-783)
            np0 += dp; np1 -= dp; nn0 += dn; nn1 -= dn; c0 += dc; c1 -= dc;
            const double proposed = current_lk();
            const double diff = __dsub_rn(proposed, lk);
            bool accept;
            if (0.0 < diff) accept = true;
            else if (diff < -45.0) { (void)rng.next_u64(); accept = false; } // threshold 0: the draw is consumed, never accepted
            else {
                bool decided = false;
                accept = false;
                uint64_t r = 0;
                if (diff < -0.001) { // exp(diff) < 1: exactly one draw, compared with trunc(exp(diff) * 2^64)
                    r = rng.next_u64();
                    float pf;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pf) : "f"((float)diff * 1.4426950408889634f));
                    const double rh = (double)(uint32_t)(r >> 32);
                    const double ub = __dadd_rn(__dmul_rn((double)(pf * 1.0001f), 4294967296.0), 2.0);
                    const double lb = __dadd_rn(__dmul_rn((double)(pf * 0.9999f), 4294967296.0), -2.0);
                    if (rh >= ub) { decided = true; accept = false; }
                    else if (__dadd_rn(rh, 1.0) <= lb) { decided = true; accept = true; }
                }
                if (!decided) {
                    const double p = exp(diff);
                    if (p == 1.0) accept = true;                            // Bernoulli ALWAYS_TRUE: no draw
                    else if (!(p >= 0.0 && p < 1.0)) { err = kMcmcBadProb; break; } // NaN: the reference panics
                    else {
                        if (!(diff < -0.001)) r = rng.next_u64();
                        accept = r < __double2ull_rz(__dmul_rn(p, 18446744073709551616.0));
                    }
                }
            }
            if (accept) {
                c.assign[idx] = (uint8_t)(1u - old); // every lane of the group stores the same byte: no barrier needed to read it back
                lk = proposed;
                if (mx < lk) {
                    mx = proposed;
                    for (uint32_t i = g; i < n; i += kSubLanes) c.argmax[i] = c.assign[i];
                }
            } else { // flip back: the same subtraction / addition the reference performs
                tot0 = __dadd_rn(tot0, -s); tot1 = __dadd_rn(tot1, s);
                np0 -= dp; np1 += dp; nn0 -= dn; nn1 += dn; c0 -= dc; c1 += dc;
            }
        }
        if (err != kMcmcOk) break;
        __syncwarp(gmask);
        for (uint32_t i = g; i < n; i += kSubLanes) c.assign[i] = c.argmax[i];
        __syncwarp(gmask);
        build();
        const double chk = current_lk();
        if (!(fabs(__dsub_rn(mx, chk)) < 0.0001)) { err = kMcmcLkMismatch; break; }
        if (!any || !(mx < best_lk)) { for (uint32_t i = g; i < n; i += kSubLanes) c.best[i] = c.assign[i]; best_lk = mx; any = true; }
        __syncwarp(gmask);
    }
    __syncwarp(gmask);
    if (g == 0) {
        rng_state[4 * chain] = rng.s0; rng_state[4 * chain + 1] = rng.s1; rng_state[4 * chain + 2] = rng.s2; rng_state[4 * chain + 3] = rng.s3;
        out_lk[chain] = best_lk;
        out_err[chain] = err;
    }
    uint8_t *oa = out_asn + asn_off[chain];
    for (uint32_t i = g; i < n; i += kSubLanes) oa[i] = c.best[i];
}

// ------------------------------------------------------------------------------------------------
// Two clusters, at most eight columns, SPECULATIVE: one chain per warp PAIR.  The evaluator warp runs up to four proposals of
// the chain side by side on its four 8-lane groups and commits up to and including the first acceptance; the helper warp
// runs the chain's generator AHEAD of it.  mcmc_with_filter rejects almost every proposal once the chain has settled
// (3.8-4.0 of 4 speculated proposals commit on diploid chunks), and a rejected proposal leaves the chain where it was except
// for the rounding of its flip / flip-back round trip (a + s) + (-s):
//   * helper: blocks of 16 draws into a 128-entry ring (every lane advances the state, lane i keeps s1 of step i and computes
//     draw i, its gen_range(n) result and two flags: could it end gen_range(n) -- low word of v*n inside the zone --, could it
//     end gen_index(1) -- top bit clear), the generator state at the last eight block starts kept for the rewind at the end of
//     the chain (k-means of the next restart continues at the stream position).  Hand-off through release / acquire words in
//     shared memory: GEN (draws published), HEAD (draws consumed: the helper stays < 128 ahead), CMD / ACK (start, stop, exit);
//   * evaluator, lane L: "a proposal that starts at draw head+L ends where?" from two ballots of the flags and two bit scans
//     (index draw(s), choose draw(s), ONE acceptance draw -- exactly what a rejected proposal consumes); four dependent
//     shuffles chain the proposals of the round;
//   * every lane replays the round trips of proposals 0..3 on its column (group j starts from the state after j of them);
//     group j then evaluates proposal j like mcmc_diploid_kernel (same f64 operations in the same order);
//   * first acceptance j*: the state of group j* is broadcast, the stream moves to behind its acceptance draw (or to the
//     draw itself when the acceptance consumed none: diff > 0 or exp(diff) == 1), j* + 1 proposals are done.  None: the
//     state after all round trips, all proposals done.
// The schedule is restated on the host (local_clustering.cpp, mcmc_with_filter_spec2) and checked against the sequential
// chain there; this kernel is checked against the sequential host twin (tests/test_gpu_clustering.py).  A window that
// does not hold one whole proposal (never at 32 draws; forced by a small `window` in the tests) is scanned draw by draw.
// ------------------------------------------------------------------------------------------------
#ifndef JTK_SPEC_NARROW
#define JTK_SPEC_NARROW 0
#endif
constexpr bool kSpecNarrow = JTK_SPEC_NARROW != 0; // eight proposals per round for chains of <= 4 columns (measured: see DESIGN 3.4)
constexpr int kSpecMax = 8, kSpecRing = 128, kSpecBlock = 16, kSpecSnaps = kSpecRing / kSpecBlock;
enum { kCmdStart = 1, kCmdStop = 2, kCmdExit = 3 };
struct SpecLayout { uint32_t x, cl, s2l, ratio, assign, argmax, best, xch, centers, dists, cum, counts, ring, flags, snap, start, ctrl, total, rwords; };
__host__ __device__ inline SpecLayout spec_layout(uint32_t n, uint32_t DP) {
    SpecLayout L;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { const uint32_t at = o; o += (bytes + 15u) & ~15u; return at; };
    L.rwords = (n + 1 + 31) / 32;
    L.x = take(8u * (n + 1) * DP);            // one spare row: lanes past the last column read (and ignore) the next entries
    L.s2l = take(8u * (n + 1));
    L.xch = take(8u * 2 * kSpecMax * 2 * DP);
    L.ring = take(8u * kSpecRing);
    L.flags = take(4u * kSpecRing);
    L.snap = take(32u * kSpecSnaps);
    L.start = take(32u);
    L.ctrl = take(32u);                       // GEN, HEAD, CMD, ACK, LIVE
    L.centers = take(8u * 2 * DP);
    L.dists = take(8u * n);
    L.cum = take(8u * n);
    L.ratio = take(4u * (n + 1) * L.rwords);
    L.counts = take(16);
    L.cl = take((n + 1) * DP);
    L.assign = take(n);
    L.argmax = take(n);
    L.best = take(n);
    L.total = o;
    return L;
}
__device__ __forceinline__ uint32_t ld_acquire_smem(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_smem(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_smem(uint32_t *p, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// A helper warp: lane group q (4 lanes) runs the generator of chain 8h + q of the CTA.  The eight groups poll their chains'
// control words and generate a block whenever one of them has room: the state update is the same instruction stream for
// all of them (a chain that has no room computes and discards), so one block costs the warp what it cost for one chain.
// Commands in CTRL[2] = (sequence << 2) | type; lane l of a group computes draws l, l + 4, l + 8 and l + 12 of a block.
constexpr int kHelpLanes = 4, kHelpChains = 32 / kHelpLanes, kHelpDraws = kSpecBlock / kHelpLanes;
struct SpecShared { uint64_t *RING, *SNAP, *START; uint32_t *FLAGS, *CTRL; };
__device__ __forceinline__ SpecShared spec_shared(unsigned char *sm, const SpecLayout &L) {
    SpecShared S;
    S.RING = reinterpret_cast<uint64_t *>(sm + L.ring); S.SNAP = reinterpret_cast<uint64_t *>(sm + L.snap);
    S.START = reinterpret_cast<uint64_t *>(sm + L.start); S.FLAGS = reinterpret_cast<uint32_t *>(sm + L.flags);
    S.CTRL = reinterpret_cast<uint32_t *>(sm + L.ctrl);
    return S;
}
__device__ __noinline__ void spec_helper(const SpecShared S, uint32_t n, bool valid, int lane) {
    const int l8 = lane & (kHelpLanes - 1);
    const uint32_t r32 = n;
    const uint64_t zone = ((uint64_t)n << __clzll((long long)n)) - 1;
    uint32_t seen = 0, gen = 0;
    int phase = valid ? 0 : 2;             // 0 idle, 1 running, 2 done
    DevRng rng{ 0, 0, 0, 0 };
    for (;;) {
        if (phase != 2) {
            const uint32_t cmd = ld_acquire_smem(S.CTRL + 2);
            if (cmd != seen) {
                if (phase == 1) { // stop (or exit) while running: leave the live state and how far it got, acknowledge
                    if (l8 == 0) {
                        S.START[0] = rng.s0; S.START[1] = rng.s1; S.START[2] = rng.s2; S.START[3] = rng.s3; S.CTRL[4] = gen;
                        st_release_smem(S.CTRL + 3, cmd);
                    }
                    phase = 0;
                }
                seen = cmd;
                if ((cmd & 3u) == kCmdStart) { rng.s0 = S.START[0]; rng.s1 = S.START[1]; rng.s2 = S.START[2]; rng.s3 = S.START[3]; gen = 0; phase = 1; }
                else if ((cmd & 3u) == kCmdExit) phase = 2;
            }
        }
        bool run = false;
        if (phase == 1) {
            const uint32_t hp = ld_acquire_smem(S.CTRL + 1);
            run = gen + kSpecBlock <= (hp & ~(uint32_t)(kSpecBlock - 1)) + kSpecRing;
        }
        if (__all_sync(kFullMask, phase == 2)) return;
        if (!__any_sync(kFullMask, run)) { __nanosleep(64); continue; }
        if (run && l8 < 4) S.SNAP[((gen / kSpecBlock) & (kSpecSnaps - 1)) * 4 + l8] = l8 == 0 ? rng.s0 : l8 == 1 ? rng.s1 : l8 == 2 ? rng.s2 : rng.s3;
        DevRng t = rng;
        uint64_t mine[kHelpDraws];
#pragma unroll
        for (int i = 0; i < kHelpDraws; i++) mine[i] = 0;
#pragma unroll
        for (int i = 0; i < kSpecBlock; i++) {
            if (l8 == (i & (kHelpLanes - 1))) mine[i / kHelpLanes] = t.s1;
            const uint64_t sh = t.s1 << 17;
            t.s2 ^= t.s0; t.s3 ^= t.s1; t.s1 ^= t.s2; t.s0 ^= t.s3;
            t.s2 ^= sh;
            t.s3 = DevRng::rotl(t.s3, 45);
        }
        if (run) {
            rng = t;
#pragma unroll
            for (int h = 0; h < kHelpDraws; h++) {
                const uint64_t v = DevRng::rotl(mine[h] * 5, 7) * 9;
                const uint64_t t0 = (uint64_t)(uint32_t)v * r32;
                const uint64_t u = (uint64_t)(uint32_t)(v >> 32) * r32 + (t0 >> 32);
                const uint64_t lo = (u << 32) | (uint32_t)t0;
                const uint32_t at = (gen + kHelpLanes * h + l8) & (kSpecRing - 1);
                S.RING[at] = v;
                S.FLAGS[at] = (uint32_t)(u >> 32) | (lo <= zone ? 0x100u : 0u) | (!(v >> 63) ? 0x200u : 0u);
            }
            gen += kSpecBlock;
        }
        __syncwarp();
        if (run && l8 == 0) st_release_smem(S.CTRL, gen);
    }
}

template <int DP>
__global__ void __maxnreg__(112) mcmc_speculative_kernel(const McmcChain *__restrict__ chains, const int *__restrict__ ids,
                                                                  int n_ids, const double *__restrict__ wf64, uint64_t *rng_state,
                                                                  uint8_t *out_asn, const uint64_t *__restrict__ asn_off,
                                                                  double *out_lk, int *out_err, int restarts, int smem_per_chain,
                                                                  int win_arg, int n_eval) {
    extern __shared__ __align__(16) unsigned char mcmc_smem[];
    // warps 0..n_eval-1 evaluate one chain each; helper warp n_eval + h runs the generators of chains 8h..8h+7 (4 lanes each)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool helper = warp >= n_eval;
    const int mine = helper ? kHelpChains * (warp - n_eval) + lane / kHelpLanes : warp;   // the chain of the CTA this lane works for
    // SPEC proposals side by side on groups of GL lanes: eight 4-lane groups for chains of <= 4 columns (they accept 10-20 % of
    // their proposals once settled: 4.0 of 8 commit per round where 3.0 of 4 did), four 8-lane groups otherwise
    constexpr int SPEC = (kSpecNarrow && DP <= 4) ? 8 : 4, GL = 32 / SPEC;
    const int g = lane & (GL - 1), grp = lane / GL;
    const int slot = blockIdx.x * n_eval + mine;
    const bool active = mine < n_eval && slot < n_ids;        // (idle evaluators / lane groups stay for the barrier)
    const int chain = ids[active ? slot : 0];
    const McmcChain ch = chains[chain];
    const uint32_t n = ch.n, D = ch.D;
    unsigned char *sm = mcmc_smem + (size_t)(mine < n_eval ? mine : 0) * smem_per_chain;
    const SpecLayout L = spec_layout(n, DP);
    const SpecShared S = spec_shared(sm, L);
    uint64_t *RING = S.RING, *SNAP = S.SNAP, *START = S.START;
    uint32_t *FLAGS = S.FLAGS, *CTRL = S.CTRL;
    if (!helper && lane < 8) CTRL[lane] = 0;                  // the control words of the chain, before its helper polls them
    __syncthreads();
    if (helper) { spec_helper(S, n, active, lane); return; }
    if (!active) return;
    double *X = reinterpret_cast<double *>(sm + L.x);
    double *S2L = reinterpret_cast<double *>(sm + L.s2l);
    double *XCH = reinterpret_cast<double *>(sm + L.xch);
    uint32_t *RAT = reinterpret_cast<uint32_t *>(sm + L.ratio);
    uint8_t *CL = sm + L.cl;
    const uint32_t RW = L.rwords;
    ChainView c;
    c.n = n; c.D = D; c.k = 2; c.ld = DP;
    c.flat = X; c.size_to_lk = S2L; c.pinc = nullptr; c.ninc = nullptr; c.ratio_ok = nullptr;
    c.centers = reinterpret_cast<double *>(sm + L.centers);
    c.dists = reinterpret_cast<double *>(sm + L.dists);
    c.cum = reinterpret_cast<double *>(sm + L.cum);
    c.counts = reinterpret_cast<uint32_t *>(sm + L.counts);
    c.assign = sm + L.assign; c.argmax = sm + L.argmax; c.best = sm + L.best;
    { // stage the chain: data (padded to DP columns, one spare zero row), sign classes, size prior, the ratio table as bits
        const double *f = wf64 + ch.off_f64;
        for (uint32_t e = lane; e < (n + 1) * DP; e += 32) {
            const uint32_t i = e / DP, d = e % DP;
            const double x = (d < D && i < n) ? f[(size_t)i * D + d] : 0.0;
            X[e] = x;
            CL[e] = (uint8_t)((kPosThr < x ? 1 : 0) | ((!(kPosThr < x) && x < -kPosThr) ? 2 : 0));
        }
        for (uint32_t i = lane; i <= n; i += 32) S2L[i] = f[(size_t)n * D + i];
        for (uint32_t e = lane; e < (n + 1) * RW; e += 32) {
            const uint32_t p = e / RW, w = e % RW;
            uint32_t bits = 0;
            for (uint32_t b = 0; b < 32; b++) {
                const uint32_t q = w * 32 + b;
                if (q + p <= n && 0.70 < __ddiv_rn((double)p, __dadd_rn((double)(p + q), 0.0000001))) bits |= 1u << b;
            }
            RAT[e] = bits;
        }
    }
    __syncwarp();
    DevRng rng{ rng_state[4 * chain], rng_state[4 * chain + 1], rng_state[4 * chain + 2], rng_state[4 * chain + 3] };
    const bool col = (uint32_t)g < D;        // lanes g >= DP shadow column 0 and never publish a term
    const bool pub = g < DP;
    const int gg = pub ? g : 0;
    auto ratio_ok = [&](uint32_t p, uint32_t q) -> unsigned { return (RAT[p * RW + (q >> 5)] >> (q & 31u)) & 1u; };
    uint32_t head = 0, seq = 0;              // stream position relative to the start of this restart's chain; command counter
    double tot0 = 0.0, tot1 = 0.0; uint32_t np0 = 0, np1 = 0, nn0 = 0, nn1 = 0, c0 = 0, c1 = 0;
    auto build = [&]() {
        tot0 = tot1 = 0.0; np0 = np1 = nn0 = nn1 = c0 = c1 = 0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t a = c.assign[i];
            const double x = X[i * DP + gg];
            const uint32_t cl = CL[i * DP + gg];
            if (a == 0) { c0++; tot0 = __dadd_rn(tot0, x); np0 += cl & 1u; nn0 += cl >> 1; }
            else { c1++; tot1 = __dadd_rn(tot1, x); np1 += cl & 1u; nn1 += cl >> 1; }
        }
    };
    unsigned buf = 0;
    // get_lk (:785-795) with get_used_columns (:847-869) of the state the GROUP holds (one column per lane)
    auto group_lk = [&](double a0, double a1, uint32_t p0, uint32_t p1, uint32_t q0, uint32_t q1, uint32_t k0, uint32_t k1) -> double {
        const bool pos0 = 0.0 < a0, pos1 = 0.0 < a1;
        const unsigned u = ((unsigned)pos0 & ratio_ok(p0, q0)) | ((unsigned)pos1 & ratio_ok(p1, q1));
        const uint32_t in_use = (pos0 ? p0 : 0u) + (pos1 ? p1 : 0u);
        const uint32_t in_neg = ((a0 <= 0.0) ? p0 : 0u) + ((a1 <= 0.0) ? p1 : 0u);
        const bool use = col && u != 0u && 2u * in_neg < in_use; // the reference compares the same integers as f64
        double *T = XCH + (buf * SPEC + grp) * (2 * DP);
        buf ^= 1u;
        if (pub) {
            T[g] = use ? (a0 < 0.0 ? 0.0 : a0) : 0.0;
            T[DP + g] = use ? (a1 < 0.0 ? 0.0 : a1) : 0.0;
        }
        __syncwarp();
        double lk = __dadd_rn(S2L[k0], S2L[k1]);
#pragma unroll
        for (int d = 0; d < 2 * DP; d++) lk = __dadd_rn(lk, T[d]);
        return lk;
    };
    auto command = [&](uint32_t type) { seq++; __syncwarp(); if (lane == 0) st_release_smem(CTRL + 2, (seq << 2) | type); };
    double best_lk = 0.0;
    bool any = false;
    int err = kMcmcOk;
    for (int rs = 0; rs < restarts && err == kMcmcOk; rs++) { // mcmc_clustering (:649-670): max_by keeps the last maximum
        if (lane == 0) err = kmeans(c, rng);
        __syncwarp();
        err = __shfl_sync(kFullMask, err, 0);
        rng.s0 = __shfl_sync(kFullMask, rng.s0, 0); rng.s1 = __shfl_sync(kFullMask, rng.s1, 0);
        rng.s2 = __shfl_sync(kFullMask, rng.s2, 0); rng.s3 = __shfl_sync(kFullMask, rng.s3, 0);
        if (err != kMcmcOk) break;
        // ---- mcmc_with_filter (:704-762), k = 2, speculative schedule ----
        head = 0;
        if (lane == 0) { START[0] = rng.s0; START[1] = rng.s1; START[2] = rng.s2; START[3] = rng.s3; CTRL[0] = 0; CTRL[1] = 0; }
        command(kCmdStart);
        build();
        double lk = group_lk(tot0, tot1, np0, np1, nn0, nn1, c0, c1);
        double mx = lk;
        for (uint32_t i = lane; i < n; i += 32) c.argmax[i] = c.assign[i];
        const uint64_t total = 2000ull * n;
        uint64_t t = 0;
        while (t < total) {
            // the draws the helper has published: 24 are enough to go on (four proposals take 16.4 on average), 32 are looked at
            uint32_t avail;
            while ((avail = ld_acquire_smem(CTRL) - head) < (uint32_t)(win_arg < 24 ? win_arg : 24)) __nanosleep(100); // (leave the issue slots to the helper)
            const int window = (int)avail < win_arg ? (int)avail : win_arg;
            // ---- lane L: a proposal that starts at draw head + L ends where? ----
            const uint32_t fl = FLAGS[(head + lane) & (kSpecRing - 1)];
            const uint32_t maskA = __ballot_sync(kFullMask, lane < window && (fl & 0x100u)), maskB = __ballot_sync(kFullMask, lane < window && (fl & 0x200u));
            uint32_t word;
            {
                const uint32_t ma = maskA >> lane;
                const int a = lane + __ffs((int)ma) - 1;                 // (ma == 0: a = lane - 1, not used)
                const uint32_t mb = (ma != 0u && a + 1 < 32) ? maskB >> (a + 1) : 0u;
                const int b = a + __ffs((int)mb);
                const uint32_t my_idx = __shfl_sync(kFullMask, fl & 0xffu, a & 31);
                word = (mb != 0u && b + 1 < window) ? (my_idx | ((uint32_t)(b + 2) << 8) | 0x10000u) : 0u; // index | next start | valid
            }
            const int want = (int)(total - t < (uint64_t)SPEC ? total - t : (uint64_t)SPEC);
            int nvalid = 0;
            uint32_t info[SPEC];                                 // per proposal: read index | (acceptance draw + 1) << 8
            {
                uint32_t alive = 1u; int p = 0;
#pragma unroll
                for (int j = 0; j < SPEC; j++) {
                    uint32_t w = __shfl_sync(kFullMask, word, p & 31);
                    alive &= (uint32_t)(j < want) & (uint32_t)(p < window) & (w >> 16);
                    w = alive ? w : 0x100u;                  // (not valid: index 0, acceptance draw 0 -- reads nobody uses)
                    info[j] = w & 0xffffu;
                    p = (int)((w >> 8) & 0xffu);
                    nvalid += (int)alive;
                }
            }
            if (nvalid == 0) { // the first proposal does not end inside the window: scan it draw by draw
                uint32_t hi = 0;
                for (;;) {
                    while (ld_acquire_smem(CTRL) == head) { }
                    const uint32_t f1 = FLAGS[head & (kSpecRing - 1)]; head++;
                    __syncwarp();
                    if (lane == 0) st_release_smem(CTRL + 1, head);
                    if (f1 & 0x100u) { hi = f1 & 0xffu; break; }
                }
                for (;;) {
                    while (ld_acquire_smem(CTRL) == head) { }
                    const uint32_t f1 = FLAGS[head & (kSpecRing - 1)]; head++;
                    __syncwarp();
                    if (lane == 0) st_release_smem(CTRL + 1, head);
                    if (f1 & 0x200u) break;
                }
                while (ld_acquire_smem(CTRL) == head) { }
                nvalid = 1;
#pragma unroll
                for (int j = 0; j < SPEC; j++) info[j] = hi | 0x100u;
            }
            // ---- this group's proposal ----
            uint32_t my = info[0];
#pragma unroll
            for (int m = 1; m < SPEC; m++) if (m == grp) my = info[m];
            const uint32_t idx_own = my & 0xffu;
            const int pc_own = (int)(my >> 8) - 1;
            const uint32_t old_own = c.assign[idx_own], cl_own = CL[idx_own * DP + gg];
            // ---- states: every lane replays the round trips of the proposals on its column; group j starts from state j ----
            double a0 = tot0, a1 = tot1, b0 = tot0, b1 = tot1, e0 = tot0, e1 = tot1, s_own = 0.0;
#pragma unroll
            for (int m = 0; m < SPEC; m++) {
                const uint32_t im = info[m] & 0xffu;
                const double x = X[im * DP + gg];
                const double s = c.assign[im] == 0u ? -x : x;
                if (m == grp) { b0 = a0; b1 = a1; s_own = s; }
                a0 = __dadd_rn(__dadd_rn(a0, s), -s);
                a1 = __dadd_rn(__dadd_rn(a1, -s), s);
                if (m + 1 == nvalid) { e0 = a0; e1 = a1; }
            }
            // ---- proposal grp: flip (:764-783), get_lk, acceptance ----
            const int dp = old_own == 0u ? -(int)(cl_own & 1u) : (int)(cl_own & 1u);
            const int dn = old_own == 0u ? -(int)(cl_own >> 1) : (int)(cl_own >> 1);
            const int dc = old_own == 0u ? -1 : 1;
            const double f0 = __dadd_rn(b0, s_own), f1 = __dadd_rn(b1, -s_own);
            const uint32_t P0 = np0 + dp, P1 = np1 - dp, Q0 = nn0 + dn, Q1 = nn1 - dn, K0 = c0 + dc, K1 = c1 - dc;
            const double proposed = group_lk(f0, f1, P0, P1, Q0, Q1, K0, K1);
            const double diff = __dsub_rn(proposed, lk);
            bool accept = false, bad = false;
            uint32_t used = 1;
            if (0.0 < diff) { accept = true; used = 0; }
            else if (diff < -45.0) { accept = false; } // threshold 0: the draw is consumed, never accepted
            else {
                const uint64_t r = RING[(head + pc_own) & (kSpecRing - 1)];
                bool decided = false;
                if (diff < -0.001) { // exp(diff) < 1: exactly one draw, compared with trunc(exp(diff) * 2^64)
                    float pf;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pf) : "f"((float)diff * 1.4426950408889634f));
                    const double rh = (double)(uint32_t)(r >> 32);
                    const double ub = __dadd_rn(__dmul_rn((double)(pf * 1.0001f), 4294967296.0), 2.0);
                    const double lb = __dadd_rn(__dmul_rn((double)(pf * 0.9999f), 4294967296.0), -2.0);
                    if (rh >= ub) { decided = true; accept = false; }
                    else if (__dadd_rn(rh, 1.0) <= lb) { decided = true; accept = true; }
                }
                if (!decided) {
                    const double p = exp(diff);
                    if (p == 1.0) { accept = true; used = 0; }               // Bernoulli ALWAYS_TRUE: no draw
                    else if (!(p >= 0.0 && p < 1.0)) { accept = true; bad = true; } // NaN: the reference panics (if the chain gets here)
                    else accept = r < __double2ull_rz(__dmul_rn(p, 18446744073709551616.0));
                }
            }
            const uint32_t accmask = __ballot_sync(kFullMask, accept && grp < nvalid);
            if (accmask) {
                const int src = __ffs((int)accmask) - 1;   // first lane of the first accepting group (its lanes agree)
                const int js = src / GL;
                if (__shfl_sync(kFullMask, (int)bad, src)) { err = kMcmcBadProb; break; }
                tot0 = shfl_f64(f0, js * GL + g); tot1 = shfl_f64(f1, js * GL + g);
                np0 = __shfl_sync(kFullMask, P0, js * GL + g); np1 = __shfl_sync(kFullMask, P1, js * GL + g);
                nn0 = __shfl_sync(kFullMask, Q0, js * GL + g); nn1 = __shfl_sync(kFullMask, Q1, js * GL + g);
                c0 = __shfl_sync(kFullMask, K0, src); c1 = __shfl_sync(kFullMask, K1, src);
                lk = shfl_f64(proposed, src);
                const uint32_t used_s = __shfl_sync(kFullMask, used, src);
                const uint32_t idx_s = __shfl_sync(kFullMask, idx_own, src), old_s = __shfl_sync(kFullMask, old_own, src);
                const int pc_s = __shfl_sync(kFullMask, pc_own, src);
                __syncwarp();                                  // every lane has read assign[] for this round
                if (lane == 0) c.assign[idx_s] = (uint8_t)(1u - old_s);
                __syncwarp();
                if (mx < lk) {
                    mx = lk;
                    for (uint32_t i = lane; i < n; i += 32) c.argmax[i] = c.assign[i];
                }
                head += (uint32_t)pc_s + used_s;
                t += (uint64_t)js + 1;
            } else {
                tot0 = e0; tot1 = e1;
                uint32_t last = info[SPEC - 1];                // (the usual round: every proposal of the round was valid)
                if (nvalid != SPEC) {
                    last = info[0];
#pragma unroll
                    for (int m = 1; m < SPEC - 1; m++) if (m == nvalid - 1) last = info[m];
                }
                head += last >> 8;                             // behind the acceptance draw of the last proposal
                t += (uint64_t)nvalid;
            }
            if (lane == 0) st_relaxed_smem(CTRL + 1, head); // (every read of the ring fed a ballot of this round: they are done)
        }
        // the generator of the sequential chain stands at `head`: stop the helper, take the block start before head, replay
        command(kCmdStop);
        while (ld_acquire_smem(CTRL + 3) != ((seq << 2) | kCmdStop)) { }
        if (head == CTRL[4]) { rng.s0 = START[0]; rng.s1 = START[1]; rng.s2 = START[2]; rng.s3 = START[3]; }
        else {
            const uint32_t b = head / kSpecBlock, q = (b & (kSpecSnaps - 1)) * 4;
            rng.s0 = SNAP[q]; rng.s1 = SNAP[q + 1]; rng.s2 = SNAP[q + 2]; rng.s3 = SNAP[q + 3];
            for (uint32_t p = b * kSpecBlock; p < head; p++) (void)rng.next_u64();
        }
        if (err != kMcmcOk) break;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) c.assign[i] = c.argmax[i];
        __syncwarp();
        build();
        const double chk = group_lk(tot0, tot1, np0, np1, nn0, nn1, c0, c1);
        if (!(fabs(__dsub_rn(mx, chk)) < 0.0001)) { err = kMcmcLkMismatch; break; }
        if (!any || !(mx < best_lk)) { for (uint32_t i = lane; i < n; i += 32) c.best[i] = c.assign[i]; best_lk = mx; any = true; }
        __syncwarp();
    }
    command(kCmdExit);
    __syncwarp();
    if (lane == 0) {
        rng_state[4 * chain] = rng.s0; rng_state[4 * chain + 1] = rng.s1; rng_state[4 * chain + 2] = rng.s2; rng_state[4 * chain + 3] = rng.s3;
        out_lk[chain] = best_lk;
        out_err[chain] = err;
    }
    uint8_t *oa = out_asn + asn_off[chain];
    for (uint32_t i = lane; i < n; i += 32) oa[i] = c.best[i];
}

size_t mcmc_smem_bytes(uint32_t n, uint32_t D, uint32_t k) {
    return (sizeof(double) * ((size_t)k * D + 2 * (size_t)n) + sizeof(uint32_t) * ((k + 1) & ~1u) + 3 * (size_t)n + 15) & ~(size_t)15;
}

template <int DP>
static cudaError_t launch_diploid(const McmcChain *chains, const int *ids, int n_ids, uint32_t n_max, const double *wf64, uint64_t *rng_state,
                                  uint8_t *out_asn, const uint64_t *asn_off, double *out_lk, int *out_err, int restarts, cudaStream_t st) {
    const size_t per_chain = sub_layout(n_max, DP).total;
    int warps = 4;
    while (warps > 1 && (size_t)warps * (32 / kSubLanes) * per_chain > 110 * 1024) warps >>= 1; // two CTAs per SM when they fit
    const size_t dyn = (size_t)warps * (32 / kSubLanes) * per_chain;
    if (dyn > 220 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(mcmc_diploid_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    const int per_cta = warps * (32 / kSubLanes);
    mcmc_diploid_kernel<DP><<<(n_ids + per_cta - 1) / per_cta, warps * 32, dyn, st>>>(chains, ids, n_ids, wf64, rng_state, out_asn, asn_off,
                                                                                       out_lk, out_err, restarts, (int)per_chain);
    return cudaGetLastError();
}

constexpr int kSpecMaxEval = 14; // evaluator warps per CTA: 14 + 2 helpers = 16 warps x 112 registers fill the register file of an SM
static int spec_chains_per_sm(size_t per_chain) { return (int)std::min<size_t>(kSpecMaxEval, (220 * 1024) / per_chain); }
template <int DP>
static cudaError_t launch_speculative(const McmcChain *chains, const int *ids, int n_ids, uint32_t n_max, const double *wf64, uint64_t *rng_state,
                                      uint8_t *out_asn, const uint64_t *asn_off, double *out_lk, int *out_err, int restarts, int window,
                                      cudaStream_t st) {
    const size_t per_chain = spec_layout(n_max, DP).total;
    if (per_chain > 220 * 1024) return cudaErrorInvalidValue;
    // few chains: small CTAs spread them over the SMs (an evaluator alone on its scheduler has the lowest latency)
    int n_eval = n_ids > 8 * 148 ? kSpecMaxEval : n_ids > 4 * 148 ? 8 : n_ids > 2 * 148 ? 4 : n_ids > 148 ? 2 : 1;
    n_eval = std::min(n_eval, spec_chains_per_sm(per_chain));
    const size_t dyn = (size_t)n_eval * per_chain;
    cudaError_t e = cudaFuncSetAttribute(mcmc_speculative_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mcmc_speculative_kernel<DP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    const int warps = n_eval + (n_eval + kHelpChains - 1) / kHelpChains;
    mcmc_speculative_kernel<DP><<<(n_ids + n_eval - 1) / n_eval, warps * 32, dyn, st>>>(chains, ids, n_ids, wf64, rng_state, out_asn, asn_off,
                                                                                        out_lk, out_err, restarts, (int)per_chain, window, n_eval);
    return cudaGetLastError();
}

// Chains with two clusters and at most eight columns (class_of >= 0) go to the sub-warp kernel of their padded width;
// everything else runs one warp per chain.  ids: device array of n_chains ints, the chain indices grouped by class
// (class_count[c] of them for class c = 0..4: widths 2, 4, 6, 8, then the general kernel), each group sorted by read count.
int mcmc_class_of(uint32_t n, uint32_t D, uint32_t k) {
    if (k != 2 || D > 8 || n > 255) return 4;
    return (int)((D + 1) / 2) - 1;
}

cudaError_t launch_mcmc_restarts(const McmcChain *chains, const McmcChain *host_chains, const int *ids, const int *host_ids,
                                 const int class_count[5], int n_chains, double *wf64, uint8_t *wu8, uint64_t *rng_state,
                                 uint8_t *out_asn, const uint64_t *asn_off, double *out_lk, int *out_err, int restarts,
                                 size_t smem_per_chain, cudaStream_t st) {
    if (n_chains <= 0) return cudaSuccess;
    // JTK_MCMC_KERNEL = auto (default) | speculative | subwarp: which kernel takes the two-cluster chains.  The speculative
    // kernel holds 14 chains per SM (2 072 per wave), the sub-warp kernel 32 per SM (4 736) at 1.4 s per 20 restarts of 60
    // reads: auto takes the speculative kernel whenever the chains of a class fit one wave of it.
    // JTK_MCMC_WINDOW (tests) shrinks the draw window so that the draw-by-draw path runs.
    int variant = 2, window = 32;
    if (const char *v = std::getenv("JTK_MCMC_KERNEL")) variant = std::strcmp(v, "subwarp") == 0 ? 0 : std::strcmp(v, "speculative") == 0 ? 1 : 2;
    if (const char *v = std::getenv("JTK_MCMC_WINDOW")) { window = std::atoi(v); window = window < 3 ? 3 : window > 32 ? 32 : window; }
    // One kernel per class.  A chain takes the same ~0.5-1.4 s whatever the size of its launch, so the classes of a mixed
    // batch (chunks with 2, 4, 6 probes, some with three clusters) run side by side on their own streams, forked from and
    // joined to the caller's stream by events.
    int n_classes = 0;
    for (int cls = 0; cls < 5; cls++) n_classes += class_count[cls] > 0 ? 1 : 0;
    const bool fork = n_classes > 1;
    cudaEvent_t ready = nullptr;
    if (fork) {
        cudaError_t e = cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(ready, st);
        if (e != cudaSuccess) { cudaEventDestroy(ready); return e; }
    }
    cudaError_t result = cudaSuccess;
    int at = 0;
    for (int cls = 0; cls < 5 && result == cudaSuccess; cls++) {
        const int cnt = class_count[cls];
        if (cnt <= 0) continue;
        cudaStream_t cs = st;
        if (fork) {
            result = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
            if (result != cudaSuccess) break;
            result = cudaStreamWaitEvent(cs, ready, 0);
        }
        if (result == cudaSuccess && cls < 4) {
            uint32_t n_max = 0;
            for (int q = 0; q < cnt; q++) n_max = n_max > host_chains[host_ids[at + q]].n ? n_max : host_chains[host_ids[at + q]].n;
            static const int widths[4] = { 2, 4, 6, 8 };
            const size_t spec_bytes = spec_layout(n_max, (uint32_t)widths[cls]).total;
            const bool spec = variant == 1 || (variant == 2 && cnt <= 148 * spec_chains_per_sm(spec_bytes)); // one wave of it
#define JTK_MCMC_LAUNCH(DPV)                                                                                                          \
    (spec ? launch_speculative<DPV>(chains, ids + at, cnt, n_max, wf64, rng_state, out_asn, asn_off, out_lk, out_err, restarts, window, cs) \
          : launch_diploid<DPV>(chains, ids + at, cnt, n_max, wf64, rng_state, out_asn, asn_off, out_lk, out_err, restarts, cs))
            switch (cls) {
            case 0: result = JTK_MCMC_LAUNCH(2); break;
            case 1: result = JTK_MCMC_LAUNCH(4); break;
            case 2: result = JTK_MCMC_LAUNCH(6); break;
            default: result = JTK_MCMC_LAUNCH(8); break;
            }
#undef JTK_MCMC_LAUNCH
        } else if (result == cudaSuccess) {
            int warps = 4;
            while (warps > 1 && warps * smem_per_chain > 200 * 1024) warps >>= 1;
            const size_t dyn = warps * smem_per_chain;
            result = cudaFuncSetAttribute(mcmc_restarts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (result == cudaSuccess) {
                mcmc_restarts_kernel<<<(cnt + warps - 1) / warps, warps * 32, dyn, cs>>>(chains, ids + at, cnt, wf64, wu8, rng_state, out_asn, asn_off,
                                                                                        out_lk, out_err, restarts, (int)smem_per_chain);
                result = cudaGetLastError();
            }
        }
        if (fork) { // join: the caller's stream waits for this class; the stream and the event are released when their work is done
            cudaEvent_t done = nullptr;
            if (result == cudaSuccess) result = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
            if (result == cudaSuccess) result = cudaEventRecord(done, cs);
            if (result == cudaSuccess) result = cudaStreamWaitEvent(st, done, 0);
            if (done) cudaEventDestroy(done);
            cudaStreamDestroy(cs);
        }
        at += cnt;
    }
    if (ready) cudaEventDestroy(ready);
    return result;
}

} // namespace jtk
