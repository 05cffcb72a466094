// local_clustering.cpp -- host side of haplotyper::local_clustering (everything in pseudo_mcmc.rs that is not the
// pair-HMM): compress_small_gains, filter_profiles, pick_filtered_profiles, k-means++ / MCMC clustering, posteriors.
//
// Line-by-line restatement of /root/reference/haplotyper/src/local_clustering/pseudo_mcmc.rs, misc.rs:84-92,231-341 and
// likelihood_gains.rs:56-158.  In the real drop-in these loops stay in Rust; this C++ twin exists so
// that the CUDA tables can be checked end to end here: identical profiles must give identical probe columns and
// assignments.  The random streams follow rand 0.8.5 / rand_xoshiro 0.6.0 (Cargo.lock:621-663) as published:
// Xoshiro256** seeded through SplitMix64, widening-multiply range sampling, Bernoulli via 64-bit threshold,
// reservoir `IteratorRandom::choose`, cumulative-weight `choose_weighted`.  Those crates are
// not under /root/reference, so stream identity with the Rust build is unpinned; determinism and every reference
// assertion are tested.
#include "../../include/jtk_gpu.h"
#include "lc_host.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>

namespace jtk {
namespace host {

constexpr int NUM_ROW = JTK_NUM_ROW;
constexpr int COPY_SIZE = JTK_COPY_SIZE;
constexpr size_t MASK_LENGTH = 7;     // pseudo_mcmc.rs:3
constexpr size_t MAX_HOMOP_LENGTH = 2; // :4
constexpr double POS_THR = 0.00001;   // :5
constexpr double MIN_REQ_FRACTION = 0.5; // :140
constexpr double EXPT_GAIN_FACTOR = 0.8; // :286
constexpr int ROUND = 3;              // :421
constexpr double PVALUE = 0.05;       // :422

struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };
#define JTK_ASSERT(c, msg) do { if (!(c)) throw Panic(std::string("assertion failed: ") + msg); } while (0)

enum DiffType { Subst = 0, Del = 1, Ins = 2 }; // likelihood_gains.rs:195-199

// ---------------------------------------------------------------------------------------------- rand 0.8.5
struct Rng { // rand_xoshiro::Xoshiro256StarStar
    uint64_t s[4];
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    explicit Rng(uint64_t seed) { // seed_from_u64: SplitMix64 fills the state
        uint64_t x = seed;
        for (int i = 0; i < 4; i++) {
            x += 0x9e3779b97f4a7c15ULL;
            uint64_t z = x;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            s[i] = z ^ (z >> 31);
        }
    }
    uint64_t next_u64() {
        const uint64_t result = rotl(s[1] * 5, 7) * 9;
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    // UniformInt<usize>::sample_single_inclusive(0, n-1)
    size_t gen_range(size_t n) {
        JTK_ASSERT(n > 0, "gen_range: empty range");
        const uint64_t range = n;
        const uint64_t zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            const uint64_t v = next_u64();
            const unsigned __int128 m = (unsigned __int128)v * range;
            if ((uint64_t)m <= zone) return (size_t)(m >> 64);
        }
    }
    // rand::seq::gen_index
    size_t gen_index(size_t ubound) {
        if (ubound <= 0xffffffffULL) {
            const uint32_t range = (uint32_t)ubound;
            JTK_ASSERT(range > 0, "gen_index: empty range");
            const uint32_t zone = (range << __builtin_clz(range)) - 1;
            for (;;) {
                const uint32_t v = next_u32();
                const uint64_t m = (uint64_t)v * range;
                if ((uint32_t)m <= zone) return (size_t)(m >> 32);
            }
        }
        return gen_range(ubound);
    }
    // Bernoulli::new(p).sample
    bool gen_bool(double p) {
        if (!(p >= 0.0 && p < 1.0)) {
            if (p == 1.0) return true; // ALWAYS_TRUE: no draw
            throw Panic("gen_bool: p is outside range [0.0, 1.0]");
        }
        const uint64_t p_int = (uint64_t)(p * 18446744073709551616.0);
        return next_u64() < p_int;
    }
    // Uniform<f64>::new(0, total).sample
    double uniform0(double total) {
        const uint64_t bits = (next_u64() >> 12) | (1023ULL << 52);
        double v12;
        std::memcpy(&v12, &bits, 8);
        return (v12 - 1.0) * total + 0.0;
    }
    // SliceRandom::choose_weighted -> index; panics like unwrap() on invalid / all-zero weights
    size_t choose_weighted(const std::vector<double> &w) {
        JTK_ASSERT(!w.empty(), "choose_weighted: NoItem");
        double total = w[0];
        JTK_ASSERT(total >= 0.0, "choose_weighted: InvalidWeight");
        std::vector<double> cum;
        cum.reserve(w.size());
        for (size_t i = 1; i < w.size(); i++) {
            JTK_ASSERT(w[i] >= 0.0, "choose_weighted: InvalidWeight");
            cum.push_back(total);
            total += w[i];
        }
        JTK_ASSERT(total != 0.0, "choose_weighted: AllWeightsZero");
        const double chosen = uniform0(total);
        size_t lo = 0, hi = cum.size(); // first cumulative weight > chosen
        while (lo < hi) {
            const size_t mid = lo + (hi - lo) / 2;
            if (cum[mid] <= chosen) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
    // IteratorRandom::choose over (0..k).filter(|x| x != old): size_hint lower bound 0 -> one draw per element
    size_t choose_other(size_t k, size_t old) {
        size_t consumed = 0, result = (size_t)-1;
        for (size_t x = 0; x < k; x++) {
            if (x == old) continue;
            consumed++;
            if (gen_index(consumed) == 0) result = x;
        }
        JTK_ASSERT(result != (size_t)-1, "choose: empty iterator");
        return result;
    }
};

// ---------------------------------------------------------------------------------------------- Gains / Pvalues
struct Gains { // likelihood_gains.rs:56-112
    int H = 0;
    std::vector<double> gain[3], prob[3]; // indexed by DiffType
    double expected(size_t homop_len, DiffType t) const {
        JTK_ASSERT(0 < homop_len, "0 < homop_len");
        const size_t h = std::min<size_t>(homop_len, (size_t)H);
        return gain[t][h - 1];
    }
};

static double logsumexp2(double x, double y) { // likelihood_gains.rs:131-137
    return (y < x) ? x + std::log(1.0 + std::exp(y - x)) : y + std::log(1.0 + std::exp(x - y));
}
// i -> Prob(i <= X | n, prob)   (likelihood_gains.rs:115-129)
static std::vector<double> pvalues_of(double prob, size_t n) {
    const double ln = std::log(prob), in_ln = std::log(1.0 - prob);
    std::vector<double> lp{ in_ln * (double)n };
    for (size_t k = 0; k < n; k++) {
        const double prev = lp.back();
        const double offset = ln + std::log((double)(n - k)) - in_ln - std::log((double)(k + 1));
        lp.push_back(prev + offset);
    }
    for (size_t k = n; k-- > 0;) lp[k] = logsumexp2(lp[k + 1], lp[k]);
    for (double &x : lp) x = std::exp(x);
    return lp;
}
struct Pvalues {
    int H; size_t total;
    std::vector<std::vector<double>> tab[3];
    double pvalue(size_t homop_len, DiffType t, size_t count) const { // likelihood_gains.rs:148-158
        JTK_ASSERT(0 < homop_len, "0 < homop_len");
        JTK_ASSERT(count <= total, "count <= total");
        const size_t h = std::min<size_t>(homop_len, (size_t)H);
        return tab[t][h - 1][count];
    }
};
static Pvalues make_pvalues(const Gains &g, size_t total) {
    Pvalues p; p.H = g.H; p.total = total;
    for (int t = 0; t < 3; t++)
        for (int h = 0; h < g.H; h++) p.tab[t].push_back(pvalues_of(g.prob[t][h], total));
    return p;
}

// ---------------------------------------------------------------------------------------------- small helpers
static double logsumexp(const std::vector<double> &xs) { // misc.rs:84-92
    if (xs.empty()) return 0.;
    const double mx = *std::max_element(xs.begin(), xs.end());
    double sum = 0;
    for (double x : xs) sum += std::exp(x - mx);
    sum = std::log(sum);
    JTK_ASSERT(sum >= 0., "logsumexp sum >= 0");
    return mx + sum;
}

static void pos_to_bp_and_difftype(size_t pos, size_t &bp, DiffType &t) { // pseudo_mcmc.rs:168-178
    bp = pos / NUM_ROW;
    const size_t op = pos % NUM_ROW;
    t = op < 4 ? Subst : (op < 8 + (size_t)COPY_SIZE ? Ins : Del);
}

std::vector<size_t> homopolymer_length(const uint8_t *xs, size_t n) { // pseudo_mcmc.rs:195-211
    std::vector<size_t> homop;
    if (n == 0) return homop;
    uint8_t current = xs[0];
    size_t length = 0;
    for (size_t i = 0; i < n; i++) {
        if (xs[i] == current) length++;
        else { homop.insert(homop.end(), length, length); current = xs[i]; length = 1; }
    }
    homop.insert(homop.end(), length, length);
    JTK_ASSERT(homop.size() == n, "homop.len() == xs.len()");
    return homop;
}

static double poisson_lk(size_t x, double lambda) { // :636-638
    double s = 0;
    for (size_t c = 1; c < x + 1; c++) s += std::log((double)c);
    return (double)x * std::log(lambda) - lambda - s;
}
static double max_poisson_lk(size_t x, double lambda, size_t c_start, size_t c_end) { // :641-645
    double best = -std::numeric_limits<double>::infinity();
    for (size_t c = std::max<size_t>(c_start, 1); c <= c_end; c++) best = std::max(best, poisson_lk(x, lambda * (double)c));
    return best;
}

struct ColStat { double sum; size_t count; size_t sc[4]; }; // sc[strand*2 + positive]

// is_explainable_by_strandedness (:314-339) from the 2x2 counts
static bool explainable_by_strandedness(const ColStat &c) {
    size_t strand_count[2] = { c.sc[0] + c.sc[1], c.sc[2] + c.sc[3] };
    size_t sign_count[2] = { c.sc[0] + c.sc[2], c.sc[1] + c.sc[3] };
    const size_t sum = strand_count[0] + strand_count[1];
    if (sum == 0) return false;
    double chisq = 0; // per strand row first, then over the rows: the reference's nested .sum() (:328-337)
    for (int st = 0; st < 2; st++) {
        double row = 0;
        for (int sg = 0; sg < 2; sg++) {
            const double expected = (double)(strand_count[st] * sign_count[sg]) / (double)sum;
            const double obs = (double)c.sc[st * 2 + sg];
            row += (obs - expected) * (obs - expected) / expected; // 0/0 -> NaN, NaN < 10 is false (as in Rust)
        }
        chisq += row;
    }
    return chisq < 10.0;
}

static bool is_in_short_homopolymer(size_t pos, const std::vector<size_t> &homop, const uint8_t *tmpl, size_t Lt) { // :497-514
    size_t x; DiffType t;
    pos_to_bp_and_difftype(pos, x, t);
    if (t == Ins) {
        const size_t r = pos % NUM_ROW - 4;
        const uint8_t base = r < 4 ? (uint8_t)"ACGT"[r] : 0;
        JTK_ASSERT(x >= 1, "index out of bounds: template[x - 1]");
        const size_t prev_len = (0 < x) ? homop[x - 1] + (tmpl[x - 1] == base) : (size_t)(tmpl[x - 1] == base);
        JTK_ASSERT(x < Lt, "index out of bounds: template[x]");
        const size_t next_len = (x < Lt) ? homop[x] + (tmpl[x] == base) : (size_t)(tmpl[x] == base);
        return prev_len <= MAX_HOMOP_LENGTH && next_len <= MAX_HOMOP_LENGTH;
    }
    if (t == Del && x < Lt) return homop[x] <= MAX_HOMOP_LENGTH;
    return true;
}

static bool has_small_pvalue(size_t pos, double gain, size_t count, const std::vector<size_t> &homop, const Pvalues &pv,
                             const Gains &gains, size_t template_len) { // :476-495
    size_t bp; DiffType t;
    pos_to_bp_and_difftype(pos, bp, t);
    const size_t homop_len = bp < homop.size() ? homop[bp] : 0;
    double pvalue = pv.pvalue(homop_len, t, count);
    const double expt = gains.expected(homop_len, t) * EXPT_GAIN_FACTOR;
    pvalue = (double)template_len * pvalue;
    return (double)count * expt < gain && pvalue < PVALUE / (double)template_len;
}

// candidate columns after the per-column filters of filter_profiles (:440-466), before the greedy pick
struct Probe { size_t pos; double lk; };

static std::vector<Probe> candidate_probes(const uint8_t *tmpl, size_t Lt, const std::vector<ColStat> &stats, size_t n_reads,
                                           const Gains &gains, size_t cluster_num, double coverage) {
    const Pvalues pv = make_pvalues(gains, n_reads);
    const std::vector<size_t> homop = homopolymer_length(tmpl, Lt);
    const size_t temp_len = stats.size() / NUM_ROW;
    std::vector<Probe> probes;
    for (size_t pos = 0; pos < stats.size(); pos++) {
        const size_t bp = pos / NUM_ROW, row = pos % NUM_ROW;
        if (!(MASK_LENGTH <= bp && bp <= temp_len - MASK_LENGTH)) continue;
        if (!(row < 8 || row == 8 + (size_t)COPY_SIZE)) continue;
        if (!is_in_short_homopolymer(pos, homop, tmpl, Lt)) continue;
        if (!has_small_pvalue(pos, stats[pos].sum, stats[pos].count, homop, pv, gains, temp_len)) continue;
        if (!explainable_by_strandedness(stats[pos])) continue;
        double max_lk = -std::numeric_limits<double>::infinity();
        bool any = false;
        for (size_t k = 1; k < cluster_num + 1; k++) {
            const double v = poisson_lk(stats[pos].count, coverage * (double)k);
            if (!any || !(v < max_lk)) max_lk = v; // Iterator::max_by keeps the last of equal maxima
            any = true;
        }
        JTK_ASSERT(any, "cluster_num");
        const double total_lk = max_lk + stats[pos].sum;
        if (0.0 < total_lk) probes.push_back({ pos, total_lk });
    }
    return probes;
}

// values of the candidate columns: cand[r * M + m] = compressed profile of read r at probes[m].pos
static double cosine_similarity(const std::vector<double> &cand, size_t n, size_t M, size_t i, size_t j) { // :602-615
    double ip = 0, isq = 0, jsq = 0;
    for (size_t r = 0; r < n; r++) {
        const double x = cand[r * M + i], y = cand[r * M + j];
        if (POS_THR < std::fabs(x) && POS_THR < std::fabs(y)) { ip += x * y; isq += x * x; jsq += y * y; }
    }
    if (isq == 0.0) return 0.0;
    return ip / std::sqrt(isq) / std::sqrt(jsq);
}
static double sokal_michener(const std::vector<double> &cand, size_t n, size_t M, size_t i, size_t j) { // :618-633
    size_t mat = 0, mism = 0;
    for (size_t r = 0; r < n; r++) {
        const double x = cand[r * M + i], y = cand[r * M + j];
        if (POS_THR < std::fabs(x) && POS_THR < std::fabs(y)) { if (0.0 < x * y) mat++; else mism++; }
    }
    const size_t total = mat + mism;
    return total == 0 ? 0.0 : (double)std::max(mism, mat) / (double)total;
}

// pick_filtered_profiles (:516-575): returns indices into probes, in probe order
static std::vector<size_t> pick_filtered_profiles(const std::vector<Probe> &probes, const std::vector<double> &cand, size_t n,
                                                  size_t cluster_num) {
    const size_t M = probes.size();
    std::vector<uint8_t> sel(M, 0);
    for (int round = 0; round < ROUND; round++) {
        for (auto &b : sel) if (b == 3) b = 0;
        for (size_t it = 0; it < std::max<size_t>(cluster_num, 2); it++) {
            // find_next_variants (:590-600): max_by keeps the last maximum
            bool found = false; size_t next = 0; double best = 0;
            for (size_t m = 0; m < M; m++)
                if (sel[m] == 0 && (!found || !(probes[m].lk < best))) { found = true; next = m; best = probes[m].lk; }
            if (!found) break;
            const size_t picked_bp = probes[next].pos / NUM_ROW;
            sel[next] = 1;
            for (size_t m = 0; m < M; m++) {
                if (!(sel[m] == 0 || sel[m] == 3)) continue;
                const size_t bp = probes[m].pos / NUM_ROW;
                const size_t diff = std::max(bp, picked_bp) - std::min(bp, picked_bp);
                if (diff < MASK_LENGTH) sel[m] = 2;
                else {
                    const double sok = sokal_michener(cand, n, M, next, m);
                    const double cs = cosine_similarity(cand, n, M, next, m);
                    if (0.8 < sok || 0.8 < std::fabs(cs)) sel[m] = 3;
                }
            }
        }
    }
    std::vector<size_t> out;
    for (size_t m = 0; m < M; m++) if (sel[m] == 1) out.push_back(m);
    return out;
}

// ---------------------------------------------------------------------------------------------- clustering
typedef std::vector<std::vector<double>> Mat;

struct LKCount { // :797-845
    double total_gain = 0; size_t num_pos = 0, num_neg = 0, num_zero = 0;
    bool is_informative() const {
        const double cov = (double)(num_pos + num_neg) + 0.0000001;
        return 0.0 < total_gain && 0.70 < (double)num_pos / cov;
    }
    void add(double x) {
        total_gain += x;
        if (POS_THR < x) num_pos++; else if (x < -POS_THR) num_neg++; else { JTK_ASSERT(std::fabs(x) < POS_THR, "x.abs() < POS_THR"); num_zero++; }
    }
    void sub(double x) {
        total_gain -= x;
        if (POS_THR < x) num_pos--; else if (x < -POS_THR) num_neg--; else { JTK_ASSERT(std::fabs(x) < POS_THR, "x.abs() < POS_THR"); num_zero--; }
    }
};
typedef std::vector<std::vector<LKCount>> LKs;

static std::vector<bool> get_used_columns(const LKs &lks) { // :847-869
    std::vector<bool> use(lks[0].size(), false);
    for (const auto &row : lks) {
        JTK_ASSERT(use.size() == row.size(), "to_uses.len() == lks.len()");
        for (size_t d = 0; d < row.size(); d++) use[d] = use[d] | row[d].is_informative();
    }
    for (size_t d = 0; d < use.size(); d++) {
        size_t in_use = 0, in_neg = 0;
        for (const auto &row : lks) { if (0.0 < row[d].total_gain) in_use += row[d].num_pos; if (row[d].total_gain <= 0.0) in_neg += row[d].num_pos; }
        use[d] = use[d] & ((double)in_neg * 2.0 < (double)in_use);
    }
    return use;
}

static double get_lk(const LKs &lks, const std::vector<size_t> &clusters, const std::vector<double> &size_to_lk) { // :785-795
    const std::vector<bool> use = get_used_columns(lks);
    double lk = 0;
    for (size_t sz : clusters) lk += size_to_lk[sz];
    for (const auto &row : lks)
        for (size_t d = 0; d < row.size(); d++) if (use[d]) lk += std::max(row[d].total_gain, 0.0);
    return lk;
}

static void build_lks(const Mat &data, const std::vector<size_t> &assign, size_t k, LKs &lks, std::vector<size_t> &clusters) {
    clusters.assign(k, 0);
    lks.assign(k, std::vector<LKCount>(data[0].size()));
    for (size_t i = 0; i < data.size(); i++) {
        clusters[assign[i]]++;
        for (size_t d = 0; d < data[i].size(); d++) lks[assign[i]][d].add(data[i][d]);
    }
}

static void flip(const Mat &data, std::vector<size_t> &assign, size_t idx, size_t to, LKs &lks, std::vector<size_t> &clusters) { // :764-783
    const size_t from = assign[idx];
    clusters[from]--;
    for (size_t d = 0; d < data[idx].size(); d++) lks[from][d].sub(data[idx][d]);
    assign[idx] = to;
    clusters[to]++;
    for (size_t d = 0; d < data[idx].size(); d++) lks[to][d].add(data[idx][d]);
}

// mcmc_with_filter (:704-762).  Same proposals, same arithmetic in the same order as the reference; the state lives in
// flat arrays that are allocated once per call (the per-proposal get_lk of the restatement above allocates nothing here).
static double mcmc_with_filter(const Mat &data, std::vector<size_t> &assign, size_t k, double cov, Rng &rng) {
    const size_t n = data.size(), D = n ? data[0].size() : 0;
    std::vector<double> size_to_lk;
    for (size_t x = 0; x <= n; x++) size_to_lk.push_back(max_poisson_lk(x, cov, 1, k));
    for (double x : size_to_lk) JTK_ASSERT(!std::isnan(x), "is_valid_lk");
    std::vector<double> flat(n * D);
    for (size_t i = 0; i < n; i++) { JTK_ASSERT(data[i].size() == D, "ragged variants"); for (size_t d = 0; d < D; d++) flat[i * D + d] = data[i][d]; }
    // LKCount (:797-845) as flat arrays: total gain and the counts of positive / negative entries per (cluster, column).
    // The sign class of every entry is fixed, so it is tabulated once and a flip updates the counters without a branch
    // (the per-entry `if POS_THR < x ... else if x < -POS_THR` of add / sub mispredicted on most proposals).
    std::vector<double> total(k * D, 0.0);
    std::vector<uint32_t> npos(k * D, 0), nneg(k * D, 0);
    std::vector<uint32_t> pinc(n * D), ninc(n * D);
    std::vector<size_t> clusters(k, 0);
    std::vector<uint8_t> use(D);
    for (size_t i = 0; i < n * D; i++) {
        const double x = flat[i];
        pinc[i] = POS_THR < x ? 1u : 0u;
        ninc[i] = (!(POS_THR < x) && x < -POS_THR) ? 1u : 0u;
        JTK_ASSERT(pinc[i] || ninc[i] || std::fabs(x) < POS_THR, "x.abs() < POS_THR");
    }
    for (size_t i = 0; i < n; i++) {
        clusters[assign[i]]++;
        for (size_t d = 0; d < D; d++) {
            total[assign[i] * D + d] += flat[i * D + d];
            npos[assign[i] * D + d] += pinc[i * D + d];
            nneg[assign[i] * D + d] += ninc[i * D + d];
        }
    }
    // LKCount::is_informative's ratio test (0.70 < num_pos / (num_pos + num_neg + 1e-7)) for every pair of counts,
    // evaluated once with the reference's own expression: the proposal loop looks it up instead of dividing
    std::vector<uint8_t> ratio_ok((n + 1) * (n + 1));
    for (size_t p = 0; p <= n; p++)
        for (size_t q = 0; q + p <= n; q++) ratio_ok[p * (n + 1) + q] = 0.70 < (double)p / ((double)(p + q) + 0.0000001);
    // raw views for the proposal loop (no reloads of the vectors' data pointers after every store)
    double *__restrict__ const tot_p = total.data();
    uint32_t *__restrict__ const npos_p = npos.data();
    uint32_t *__restrict__ const nneg_p = nneg.data();
    const uint32_t *__restrict__ const pinc_p = pinc.data();
    const uint32_t *__restrict__ const ninc_p = ninc.data();
    const double *__restrict__ const flat_p = flat.data();
    const uint8_t *__restrict__ const ratio_p = ratio_ok.data();
    const double *__restrict__ const s2l_p = size_to_lk.data();
    size_t *__restrict__ const clus_p = clusters.data();
    uint8_t *__restrict__ const use_p = use.data();
    size_t *__restrict__ const asn_p = assign.data();
    auto current_lk = [&]() -> double { // get_lk (:785-795) with get_used_columns (:847-869)
        for (size_t d = 0; d < D; d++) {
            unsigned u = 0;
            uint32_t in_use = 0, in_neg = 0;
            for (size_t c = 0; c < k; c++) {
                const double g = tot_p[c * D + d];
                const uint32_t p = npos_p[c * D + d];
                const bool pos = 0.0 < g;
                u |= (unsigned)pos & ratio_p[(size_t)p * (n + 1) + nneg_p[c * D + d]];
                in_use += pos ? p : 0u;
                in_neg += (g <= 0.0) ? p : 0u;
            }
            use_p[d] = (uint8_t)(u & (unsigned)((double)in_neg * 2.0 < (double)in_use));
        }
        double lk = 0;
        for (size_t c = 0; c < k; c++) lk += s2l_p[clus_p[c]];
        for (size_t c = 0; c < k; c++)
            for (size_t d = 0; d < D; d++) if (use_p[d]) lk += std::max(tot_p[c * D + d], 0.0);
        return lk;
    };
    auto do_flip = [&](size_t idx, size_t to) { // flip (:764-783)
        const size_t from = asn_p[idx];
        const double *row = &flat_p[idx * D];
        const uint32_t *pi = &pinc_p[idx * D], *ni = &ninc_p[idx * D];
        clus_p[from]--;
        for (size_t d = 0; d < D; d++) { tot_p[from * D + d] -= row[d]; npos_p[from * D + d] -= pi[d]; nneg_p[from * D + d] -= ni[d]; }
        asn_p[idx] = to;
        clus_p[to]++;
        for (size_t d = 0; d < D; d++) { tot_p[to * D + d] += row[d]; npos_p[to * D + d] += pi[d]; nneg_p[to * D + d] += ni[d]; }
    };
    double lk = current_lk();
    double mx = lk;
    std::vector<size_t> argmax = assign;
    const size_t n_proposals = 2000 * n;
    for (size_t t = 0; t < n_proposals; t++) {
        const size_t idx = rng.gen_range(n);
        const size_t old = assign[idx];
        const size_t nw = rng.choose_other(k, old);
        do_flip(idx, nw);
        const double proposed = current_lk();
        const double diff = proposed - lk;
        // exp(diff) * 2^64 < 1 for diff < -45 (e^-45 = 2^-64.9): Bernoulli's integer threshold is 0, the draw is consumed and
        // the proposal rejected -- the same stream and the same decision as gen_bool(exp(diff)), without the exp
        bool accept;
        if (0.0 < diff) accept = true;
        else if (diff < -45.0) { (void)rng.next_u64(); accept = false; }
        else accept = rng.gen_bool(std::exp(diff));
        if (accept) {
            lk = proposed;
            if (mx < lk) { mx = proposed; std::copy(assign.begin(), assign.end(), argmax.begin()); }
        } else do_flip(idx, old);
    }
    assign = argmax;
    LKs chk_lks; std::vector<size_t> chk_clusters;
    build_lks(data, assign, k, chk_lks, chk_clusters);
    const double chk = get_lk(chk_lks, chk_clusters, size_to_lk);
    JTK_ASSERT(std::fabs(mx - chk) < 0.0001, "(max - lk).abs() < 0.0001");
    return mx;
}

// ---- misc::kmeans (misc.rs:231-341)
static double dist2(const std::vector<double> &x, const std::vector<double> &y) {
    JTK_ASSERT(x.size() == y.size(), "xs.len() == ys.len()");
    double s = 0;
    for (size_t i = 0; i < x.size(); i++) s += (x[i] - y[i]) * (x[i] - y[i]); // powi(2)
    return s;
}
static void update_assignments(const Mat &data, const Mat &centers, std::vector<size_t> &asn) {
    for (size_t i = 0; i < data.size(); i++) {
        size_t best = 0; double bd = 0;
        for (size_t c = 0; c < centers.size(); c++) {
            const double d = dist2(data[i], centers[c]);
            if (c == 0 || d < bd) { best = c; bd = d; } // min_by keeps the first minimum
        }
        asn[i] = best;
    }
}
static std::vector<size_t> suggest_first(const Mat &data, size_t k, Rng &rng) {
    JTK_ASSERT(k <= data.size(), "k <= data.len()");
    Mat centers{ data[rng.gen_index(data.size())] };
    for (size_t it = 0; it + 1 < k; it++) {
        std::vector<double> dists(data.size());
        for (size_t i = 0; i < data.size(); i++) {
            double m = 0;
            for (size_t c = 0; c < centers.size(); c++) { const double d = dist2(data[i], centers[c]); if (c == 0 || d < m) m = d; }
            dists[i] = m;
        }
        centers.push_back(data[rng.choose_weighted(dists)]);
    }
    std::vector<size_t> asn(data.size(), 0);
    update_assignments(data, centers, asn);
    return asn;
}
static double get_dist(const Mat &data, const Mat &centers, const std::vector<size_t> &asn) {
    double s = 0;
    for (size_t i = 0; i < data.size(); i++) s += dist2(data[i], centers[asn[i]]);
    return s;
}
static std::vector<size_t> kmeans(const Mat &data, size_t k, Rng &rng) {
    JTK_ASSERT(1 <= k, "1 <= k");
    const size_t dim = data[0].size();
    JTK_ASSERT(0 < dim, "0 < dim");
    std::vector<size_t> asn;
    if (rng.gen_bool(0.5)) { asn.resize(data.size()); for (auto &a : asn) a = rng.gen_range(k); }
    else asn = suggest_first(data, k, rng);
    Mat centers(k, std::vector<double>(dim, 0.0));
    std::vector<size_t> counts(k, 0);
    double dist = get_dist(data, centers, asn);
    for (;;) {
        for (auto &c : centers) std::fill(c.begin(), c.end(), 0.0);
        std::fill(counts.begin(), counts.end(), 0);
        for (size_t i = 0; i < data.size(); i++) { for (size_t d = 0; d < dim; d++) centers[asn[i]][d] += data[i][d]; counts[asn[i]]++; }
        for (size_t c = 0; c < k; c++) if (0 < counts[c]) for (double &x : centers[c]) x /= (double)counts[c];
        update_assignments(data, centers, asn);
        const double nd = get_dist(data, centers, asn);
        JTK_ASSERT(nd < dist + 0.00000001, "new_dist < dist + UPDATE_THR");
        if (dist - nd < 0.00000001) break;
        dist = nd;
    }
    return asn;
}

struct ClusterOut { std::vector<size_t> asn; double score; std::vector<double> read_gains; std::vector<bool> used; };

static void get_read_lk_gains(const Mat &variants, const std::vector<size_t> &asn, size_t k, std::vector<bool> &use, std::vector<double> &gain) { // :381-408
    LKs lks; std::vector<size_t> clusters;
    build_lks(variants, asn, k, lks, clusters);
    use = get_used_columns(lks);
    gain.assign(variants.size(), 0.0);
    for (size_t i = 0; i < variants.size(); i++) {
        double s = 0;
        for (size_t d = 0; d < variants[i].size(); d++)
            if (use[d] && POS_THR < lks[asn[i]][d].total_gain) s += variants[i][d];
        gain[i] = s;
    }
}

// mcmc_clustering (:649-670) in two halves: the 20 restarts on one generator (the part jtk_mcmc_restarts_batch runs on the
// GPU for many chunks at once) and the bookkeeping on the best assignment.
static void mcmc_restarts(const Mat &data, size_t k, double cov, Rng &rng, int restarts, std::vector<size_t> &best, double &best_lk) {
    bool any = false;
    best.clear(); best_lk = 0;
    for (int t = 0; t < restarts; t++) {
        std::vector<size_t> asn = kmeans(data, k, rng);
        const double lk = mcmc_with_filter(data, asn, k, cov, rng);
        if (!any || !(lk < best_lk)) { best = asn; best_lk = lk; any = true; } // max_by keeps the last maximum
    }
}
// ---- Host twin of the SPECULATIVE schedule of mcmc_speculative_kernel (mcmc_kernels.cu), two clusters.
// mcmc_with_filter (:704-762) rejects almost every proposal once the chain has settled, and a rejected proposal leaves
// the chain where it was except for the rounding of its flip / flip-back round trip ((a - x) + x).  So up to kSpec
// proposals are evaluated side by side: the draws are generated ahead into a ring, parsed into proposals under the
// assumption that every earlier proposal of the round is rejected (then each consumes exactly one acceptance draw),
// proposal j starts from the state after the round trips of proposals 0..j-1, and the round commits up to and including
// the first accepted proposal.  The stream position moves by exactly the draws the sequential chain would have consumed,
// so assignments, likelihoods and the generator state are the sequential chain's, bit for bit
// (tests/test_local_clustering_host.py::test_speculative_schedule_equals_the_sequential_chain).
struct SpecStats { uint64_t rounds = 0, proposals = 0, slow = 0; };
constexpr int kSpec = 8, kSpecRing = 64, kSpecBlock = 16; // kSpec: the widest round (the kernel runs 4 or 8 proposals side by side)
struct SpecStream { // draws generated ahead in blocks of kSpecBlock; snapshots of the generator at the last four block starts
    Rng &rng;
    uint64_t ring[kSpecRing];
    uint64_t snap[4][4];
    uint64_t head = 0, gen = 0; // absolute stream positions: next draw to consume / next draw to generate
    explicit SpecStream(Rng &r) : rng(r) {}
    void gen_block() {
        std::memcpy(snap[(gen / kSpecBlock) & 3], rng.s, 32);
        for (int i = 0; i < kSpecBlock; i++) ring[(gen + i) & (kSpecRing - 1)] = rng.next_u64();
        gen += kSpecBlock;
    }
    uint64_t at(uint64_t pos) const { return ring[pos & (kSpecRing - 1)]; }
    void rewind_generator() { // leaves rng where the sequential chain's generator is: at `head`
        if (head == gen) return;
        const uint64_t b = head / kSpecBlock;
        std::memcpy(rng.s, snap[b & 3], 32);
        for (uint64_t p = b * kSpecBlock; p < head; p++) (void)rng.next_u64();
        gen = head;
    }
};
static double mcmc_with_filter_spec2(const Mat &data, std::vector<size_t> &assign, double cov, Rng &rng, SpecStats &stats, int kSpecWindow, int spec) {
    const size_t n = data.size(), D = n ? data[0].size() : 0;
    std::vector<double> s2l;
    for (size_t x = 0; x <= n; x++) s2l.push_back(max_poisson_lk(x, cov, 1, 2));
    std::vector<double> X(n * D);
    std::vector<uint8_t> CL(n * D);
    for (size_t i = 0; i < n; i++)
        for (size_t d = 0; d < D; d++) {
            const double x = data[i][d];
            X[i * D + d] = x;
            CL[i * D + d] = (uint8_t)((POS_THR < x ? 1 : 0) | ((!(POS_THR < x) && x < -POS_THR) ? 2 : 0));
        }
    auto ratio_ok = [&](int64_t p, int64_t q) -> bool { return 0.70 < (double)p / ((double)(p + q) + 0.0000001); };
    struct St { std::vector<double> t0, t1; };
    std::vector<double> tot0(D, 0.0), tot1(D, 0.0);
    std::vector<int64_t> np0(D, 0), np1(D, 0), nn0(D, 0), nn1(D, 0);
    int64_t c0 = 0, c1 = 0;
    for (size_t i = 0; i < n; i++)
        for (size_t d = 0; d < D; d++) {
            const double x = X[i * D + d]; const unsigned cl = CL[i * D + d];
            if (assign[i] == 0) { if (d == 0) c0++; tot0[d] += x; np0[d] += cl & 1u; nn0[d] += cl >> 1; }
            else { if (d == 0) c1++; tot1[d] += x; np1[d] += cl & 1u; nn1[d] += cl >> 1; }
        }
    // get_lk (:785-795) with get_used_columns (:847-869) of a state given as column arrays
    auto lk_of = [&](const std::vector<double> &a0, const std::vector<double> &a1, const std::vector<int64_t> &p0, const std::vector<int64_t> &p1,
                     const std::vector<int64_t> &q0, const std::vector<int64_t> &q1, int64_t k0, int64_t k1) -> double {
        double lk = s2l[(size_t)k0] + s2l[(size_t)k1];
        std::vector<uint8_t> use(D);
        for (size_t d = 0; d < D; d++) {
            const bool pos0 = 0.0 < a0[d], pos1 = 0.0 < a1[d];
            const bool u = (pos0 && ratio_ok(p0[d], q0[d])) || (pos1 && ratio_ok(p1[d], q1[d]));
            const int64_t in_use = (pos0 ? p0[d] : 0) + (pos1 ? p1[d] : 0), in_neg = ((a0[d] <= 0.0) ? p0[d] : 0) + ((a1[d] <= 0.0) ? p1[d] : 0);
            use[d] = u && 2 * in_neg < in_use;
        }
        for (size_t d = 0; d < D; d++) lk += use[d] ? (a0[d] < 0.0 ? 0.0 : a0[d]) : 0.0;
        for (size_t d = 0; d < D; d++) lk += use[d] ? (a1[d] < 0.0 ? 0.0 : a1[d]) : 0.0;
        return lk;
    };
    double lk = lk_of(tot0, tot1, np0, np1, nn0, nn1, c0, c1);
    double mx = lk;
    std::vector<size_t> argmax = assign;
    SpecStream S(rng);
    const uint64_t range = n, zone = (range << __builtin_clzll(range)) - 1;
    const uint64_t total = 2000ull * n;
    uint64_t t = 0;
    while (t < total) {
        while (S.gen - S.head < (uint64_t)kSpecWindow) S.gen_block();
        // the window of kSpecWindow draws: which could end a gen_range (maskA), which a gen_index(1) (maskB)
        uint32_t maskA = 0, maskB = 0;
        for (int L = 0; L < kSpecWindow; L++) {
            const uint64_t v = S.at(S.head + L);
            const unsigned __int128 m = (unsigned __int128)v * range;
            if ((uint64_t)m <= zone) maskA |= 1u << L;
            if (!(v >> 63)) maskB |= 1u << L;
        }
        const int want = (int)std::min<uint64_t>((uint64_t)spec, total - t);
        int nvalid = 0, pos_idx[kSpec], pos_acc[kSpec];
        for (int j = 0, p = 0; j < want; j++) {
            if (p >= kSpecWindow) break;
            const uint32_t ma = maskA >> p;
            if (!ma) break;
            const int a = p + __builtin_ctz(ma);
            if (a + 1 >= kSpecWindow) break;
            const uint32_t mb = maskB >> (a + 1);
            if (!mb) break;
            const int b = a + 1 + __builtin_ctz(mb);
            if (b + 1 >= kSpecWindow) break;
            pos_idx[j] = a; pos_acc[j] = b + 1; p = b + 2; nvalid++;
        }
        size_t idx[kSpec];
        if (nvalid == 0) { // the first proposal does not end inside the window: scan it draw by draw (never in practice)
            stats.slow++;
            auto take = [&]() -> uint64_t { if (S.head == S.gen) S.gen_block(); return S.at(S.head++); };
            for (;;) { const unsigned __int128 m = (unsigned __int128)take() * range; if ((uint64_t)m <= zone) { idx[0] = (size_t)(m >> 64); break; } }
            while (take() >> 63) { }
            if (S.head == S.gen) S.gen_block();
            pos_acc[0] = 0; nvalid = 1;
        } else {
            for (int j = 0; j < nvalid; j++) idx[j] = (size_t)(((unsigned __int128)S.at(S.head + pos_idx[j]) * range) >> 64);
        }
        stats.rounds++;
        // state j = the chain after the round trips of proposals 0..j-1 (every one of them rejected)
        std::vector<St> st(nvalid + 1);
        st[0].t0 = tot0; st[0].t1 = tot1;
        std::vector<std::vector<double>> s(nvalid, std::vector<double>(D));
        size_t old[kSpec];
        for (int m = 0; m < nvalid; m++) {
            old[m] = assign[idx[m]];
            st[m + 1].t0.resize(D); st[m + 1].t1.resize(D);
            for (size_t d = 0; d < D; d++) {
                const double x = X[idx[m] * D + d];
                s[m][d] = old[m] == 0 ? -x : x;
                st[m + 1].t0[d] = (st[m].t0[d] + s[m][d]) + (-s[m][d]);
                st[m + 1].t1[d] = (st[m].t1[d] + (-s[m][d])) + s[m][d];
            }
        }
        int jstar = -1, used = 0;
        double proposed_star = 0;
        for (int j = 0; j < nvalid && jstar < 0; j++) { // (side by side on the device; only the first acceptance counts)
            std::vector<double> f0(D), f1(D);
            std::vector<int64_t> p0 = np0, p1 = np1, q0 = nn0, q1 = nn1;
            const int64_t sg = old[j] == 0 ? -1 : 1;
            for (size_t d = 0; d < D; d++) {
                const unsigned cl = CL[idx[j] * D + d];
                f0[d] = st[j].t0[d] + s[j][d]; f1[d] = st[j].t1[d] + (-s[j][d]);
                p0[d] += sg * (int64_t)(cl & 1u); p1[d] -= sg * (int64_t)(cl & 1u);
                q0[d] += sg * (int64_t)(cl >> 1); q1[d] -= sg * (int64_t)(cl >> 1);
            }
            const double proposed = lk_of(f0, f1, p0, p1, q0, q1, c0 + sg, c1 - sg);
            const double diff = proposed - lk;
            bool accept; int consumed;
            if (0.0 < diff) { accept = true; consumed = 0; }
            else if (diff < -45.0) { accept = false; consumed = 1; }
            else {
                const double p = std::exp(diff);
                if (p == 1.0) { accept = true; consumed = 0; }
                else {
                    if (!(p >= 0.0 && p < 1.0)) throw Panic("gen_bool: p is outside range [0.0, 1.0]");
                    accept = S.at(S.head + pos_acc[j]) < (uint64_t)(p * 18446744073709551616.0);
                    consumed = 1;
                }
            }
            if (accept) {
                jstar = j; used = consumed; proposed_star = proposed;
                tot0 = f0; tot1 = f1; np0 = p0; np1 = p1; nn0 = q0; nn1 = q1; c0 += sg; c1 -= sg;
            }
        }
        if (jstar >= 0) {
            assign[idx[jstar]] = 1 - old[jstar];
            lk = proposed_star;
            if (mx < lk) { mx = proposed_star; argmax = assign; }
            S.head += (uint64_t)pos_acc[jstar] + used;
            t += jstar + 1; stats.proposals += jstar + 1;
        } else {
            tot0 = st[nvalid].t0; tot1 = st[nvalid].t1;
            S.head += (uint64_t)pos_acc[nvalid - 1] + 1;
            t += nvalid; stats.proposals += nvalid;
        }
    }
    S.rewind_generator();
    assign = argmax;
    LKs chk_lks; std::vector<size_t> chk_clusters;
    build_lks(data, assign, 2, chk_lks, chk_clusters);
    const double chk = get_lk(chk_lks, chk_clusters, s2l);
    JTK_ASSERT(std::fabs(mx - chk) < 0.0001, "(max - lk).abs() < 0.0001");
    return mx;
}
static void mcmc_restarts_spec2(const Mat &data, double cov, Rng &rng, int restarts, std::vector<size_t> &best, double &best_lk, SpecStats &stats, int window, int spec) {
    bool any = false;
    best.clear(); best_lk = 0;
    for (int t = 0; t < restarts; t++) {
        std::vector<size_t> asn = kmeans(data, 2, rng);
        const double lk = mcmc_with_filter_spec2(data, asn, cov, rng, stats, window, spec);
        if (!any || !(lk < best_lk)) { best = asn; best_lk = lk; any = true; }
    }
}

static ClusterOut mcmc_finish(const Mat &data, size_t k, double cov, const std::vector<size_t> &best, double best_lk) {
    ClusterOut o; o.asn = best;
    get_read_lk_gains(data, o.asn, k, o.used, o.read_gains);
    std::vector<size_t> counts(k, 0);
    for (size_t a : o.asn) counts[a]++;
    double cluster_lk = 0;
    for (size_t c : counts) cluster_lk += max_poisson_lk(c, cov, 1, k);
    o.score = best_lk - cluster_lk;
    return o;
}
static ClusterOut mcmc_clustering(const Mat &data, size_t k, double cov, Rng &rng) { // :649-670
    std::vector<size_t> best; double best_lk = 0;
    mcmc_restarts(data, k, cov, rng, 20, best, best_lk);
    return mcmc_finish(data, k, cov, best, best_lk);
}

static ClusterOut use_highest_gain(const Mat &data) { // :673-693
    const size_t dim = data[0].size();
    std::vector<double> gains(dim, 0.0);
    for (const auto &xs : data) for (size_t d = 0; d < dim; d++) gains[d] += std::max(xs[d], 0.0);
    size_t mi = 0;
    for (size_t d = 0; d < dim; d++) if (!(gains[d] < gains[mi])) mi = d; // last maximum
    ClusterOut o;
    for (const auto &xs : data) o.asn.push_back((size_t)(0.0 < xs[mi]));
    get_read_lk_gains(data, o.asn, 2, o.used, o.read_gains);
    o.score = 0;
    for (double g : o.read_gains) o.score += g;
    return o;
}

typedef std::vector<std::pair<size_t, DiffType>> VarTypes;

static double min_gain(const Gains &g, const VarTypes &vt, const std::vector<bool> &used) { // :276-284
    bool any = false; double m = 0;
    for (size_t d = 0; d < vt.size() && d < used.size(); d++)
        if (used[d]) { const double v = g.expected(vt[d].first, vt[d].second) / 3.0; if (!any || v < m) { m = v; any = true; } }
    return any ? m : 1.0;
}
static double expected_gains(const Gains &g, const VarTypes &vt, const std::vector<bool> &prev, const std::vector<bool> &used) { // :287-306
    JTK_ASSERT(vt.size() == used.size(), "variant_type.len() == used_columns.len()");
    const bool no_new = prev == used;
    bool any = false; double m = 0;
    for (size_t d = 0; d < vt.size(); d++) {
        const bool check = ((!prev[d]) & used[d]) | no_new;
        const double v = check ? g.expected(vt[d].first, vt[d].second) : 0.0000001;
        if (!any || !(v < m)) { m = v; any = true; }
    }
    if (!any) m = 0.0;
    return std::max(EXPT_GAIN_FACTOR * m, 0.1);
}

static Mat get_likelihood_gain(const Mat &variants, const std::vector<size_t> &asn, size_t k) { // :353-379
    LKs lks; std::vector<size_t> clusters;
    build_lks(variants, asn, k, lks, clusters);
    const std::vector<bool> use = get_used_columns(lks);
    Mat out;
    for (const auto &vars : variants) {
        std::vector<double> row;
        for (const auto &slots : lks) {
            double s = 0;
            for (size_t d = 0; d < vars.size(); d++) if (use[d] && POS_THR < slots[d].total_gain) s += vars[d];
            row.push_back(s);
        }
        out.push_back(row);
    }
    return out;
}

struct DevResult { std::vector<size_t> asn; Mat gains; double score; size_t k; };

static DevResult cluster_filtered_variants(const Mat &variants, const VarTypes &vt, size_t copy_num, double coverage,
                                           double per_cluster_cov, const Gains &gains, Rng &rng) { // :213-274
    bool all_empty = true;
    for (const auto &xs : variants) if (!xs.empty()) all_empty = false;
    if (copy_num <= 1 || all_empty || variants.size() <= copy_num)
        return { std::vector<size_t>(variants.size(), 0), Mat(variants.size(), std::vector<double>(1, 0.0)), 0.0, 1 };
    const size_t n = variants.size();
    std::vector<size_t> assignments(n, 0);
    double mx = 0; size_t max_k = 1;
    std::vector<double> read_lk_gains(n, 0.0);
    std::vector<bool> prev_used(variants[0].size(), false);
    const size_t end = std::min(copy_num, 1 + 2 * vt.size());
    const size_t start = std::max<size_t>(end, 5) - 3;
    for (size_t k = start; k <= end; k++) {
        ClusterOut c = mcmc_clustering(variants, k, coverage, rng);
        if (k == 2) {
            ClusterOut h = use_highest_gain(variants);
            if (c.score < h.score) c = h;
        }
        (void)min_gain(gains, vt, c.used); // count_improved_reads only feeds a trace! line
        const double expected_gain = expected_gains(gains, vt, prev_used, c.used) * per_cluster_cov + 0.1;
        if (expected_gain < c.score - mx) {
            assignments = c.asn; mx = c.score; max_k = k; read_lk_gains = c.read_gains; prev_used = c.used;
        } else break;
    }
    return { assignments, get_likelihood_gain(variants, assignments, max_k), mx, max_k };
}

// exact_clustering::cluster_filtered_variants_exact (haplotyper/src/local_clustering/exact_clustering.rs:7-77): every cluster
// picks a subset of the D columns (a bit mask), a read scores the sum of its values over the subset of its best cluster;
// all non-increasing tuples of copy_num masks are enumerated (increment_one, :66-77) and the first strict maximum is kept.
// Used by sandbox/src/bin/benchmark_mcmc.rs:112-118 as the score the MCMC is compared with (SURVEY 8c pin P7).
static double exact_selection_score(size_t selection, const std::vector<double> &xs) { // get_exact_score :49-54
    double s = 0;
    for (size_t i = 0; i < xs.size(); i++) if ((selection >> i) & 1u) s += xs[i];
    return s;
}
static double exact_calc_score(const std::vector<size_t> &vars, const Mat &variants) { // calc_score :56-64
    double total = 0;
    for (const auto &xs : variants) {
        double best = exact_selection_score(vars[0], xs);
        for (size_t c = 1; c < vars.size(); c++) { const double v = exact_selection_score(vars[c], xs); if (!(v < best)) best = v; }
        total += best;
    }
    return total;
}
static DevResult exact_get_result(const std::vector<size_t> &vars, const Mat &variants) { // get_result :28-47
    DevResult r;
    r.score = exact_calc_score(vars, variants);
    r.k = vars.size();
    for (const auto &xs : variants) {
        std::vector<double> lk_gain;
        for (size_t sel : vars) lk_gain.push_back(exact_selection_score(sel, xs));
        size_t bi = 0;
        for (size_t c = 0; c < lk_gain.size(); c++) if (!(lk_gain[c] < lk_gain[bi])) bi = c; // max_by: last maximum
        r.asn.push_back(bi);
        r.gains.push_back(lk_gain);
    }
    return r;
}
static DevResult cluster_filtered_variants_exact(const Mat &variants, size_t feature_dim, size_t copy_num) {
    JTK_ASSERT(copy_num >= 1 && feature_dim < 24, "exact clustering: copy_num >= 1 and fewer than 24 columns");
    std::vector<size_t> selected(copy_num, 0);
    const size_t choices = (size_t)1 << feature_dim;
    const std::vector<size_t> last_loop(copy_num, choices - 1);
    double mx = 0;
    DevResult argmax = exact_get_result(selected, variants);
    while (selected != last_loop) { // the last tuple itself is never scored, as in the reference (:17-24)
        const double score = exact_calc_score(selected, variants);
        if (mx < score) { argmax = exact_get_result(selected, variants); mx = score; }
        size_t idx = 0; // increment_one
        while (choices == selected[idx] + 1) idx++;
        selected[idx]++;
        for (size_t j = 0; j < idx; j++) selected[j] = selected[idx];
        for (size_t w = 0; w + 1 < selected.size(); w++) JTK_ASSERT(selected[w + 1] <= selected[w], "increment_one: tuple not sorted");
    }
    return argmax;
}

// pseudo_mcmc::clustering (:77-107) after search_variants produced (variants, variant types)
static DevResult clustering_tail(const Mat &variants, const VarTypes &vt, size_t copy_num, double coverage,
                                 double local_coverage, const Gains &gains, Rng &rng) {
    DevResult r = cluster_filtered_variants(variants, vt, copy_num, coverage, local_coverage, gains, rng);
    for (size_t i = 0; i < r.asn.size(); i++) {
        const auto &lks = r.gains[i];
        size_t bi = 0;
        for (size_t c = 0; c < lks.size(); c++) if (!(lks[c] < lks[bi])) bi = c; // last maximum
        if (lks[r.asn[i]] + 0.001 < lks[bi]) r.asn[i] = bi;
    }
    for (auto &xs : r.gains) { const double tot = logsumexp(xs); for (double &x : xs) x -= tot; }
    return r;
}

static Gains to_gains(const jtk_gains *g) {
    Gains out; out.H = g->homop_len;
    for (int t = 0; t < 3; t++) { out.gain[t].assign(g->gain + t * g->homop_len, g->gain + (t + 1) * g->homop_len); out.prob[t].assign(g->prob + t * g->homop_len, g->prob + (t + 1) * g->homop_len); }
    return out;
}


// ---------------------------------------------------------------------------------------------- device-path helpers
void pvalue_tables(const jtk_gains *g, size_t total, std::vector<double> &out) {
    const Gains gains = to_gains(g);
    const Pvalues pv = make_pvalues(gains, total);
    out.assign((size_t)3 * gains.H * (total + 1), 0.0);
    for (int t = 0; t < 3; t++)
        for (int h = 0; h < gains.H; h++)
            for (size_t c = 0; c <= total; c++) out[((size_t)t * gains.H + h) * (total + 1) + c] = pv.tab[t][h][c];
}

void poisson_prior_table(double coverage, size_t cluster_num, size_t total, std::vector<double> &out) {
    out.assign(total + 1, 0.0);
    for (size_t count = 0; count <= total; count++) {
        double max_lk = -std::numeric_limits<double>::infinity();
        bool any = false;
        for (size_t k = 1; k < cluster_num + 1; k++) {
            const double v = poisson_lk(count, coverage * (double)k);
            if (!any || !(v < max_lk)) max_lk = v;
            any = true;
        }
        out[count] = max_lk;
    }
}

std::vector<size_t> pick_probes(const uint32_t *pos, const double *lk, size_t M, const double *cand, size_t n,
                                size_t cluster_num) {
    std::vector<Probe> probes(M);
    for (size_t m = 0; m < M; m++) probes[m] = Probe{ (size_t)pos[m], lk[m] };
    const std::vector<double> c(cand, cand + n * M);
    return pick_filtered_profiles(probes, c, n, cluster_num);
}

} // namespace host
} // namespace jtk

using namespace jtk::host;

namespace {
thread_local std::string g_lc_error;

int write_result(const DevResult &r, const std::vector<Probe> &probes, const std::vector<size_t> &picked, uint64_t *out_asn,
                 double *out_post, int post_stride, double *out_score, int *out_k, uint32_t *out_probe_pos, int probe_cap,
                 int *out_n_probes) {
    if ((int)r.k > post_stride) { g_lc_error = "post_stride smaller than the cluster number"; return JTK_EINVAL; }
    for (size_t i = 0; i < r.asn.size(); i++) {
        out_asn[i] = r.asn[i];
        for (size_t c = 0; c < r.k; c++) out_post[i * (size_t)post_stride + c] = r.gains[i][c];
    }
    *out_score = r.score; *out_k = (int)r.k;
    if (out_n_probes) *out_n_probes = (int)picked.size();
    if (out_probe_pos)
        for (size_t m = 0; m < picked.size() && (int)m < probe_cap; m++) out_probe_pos[m] = (uint32_t)probes[picked[m]].pos;
    return JTK_OK;
}
} // namespace

extern "C" {

const char *jtk_lc_last_error(void) { return g_lc_error.c_str(); }

// pseudo_mcmc::clustering on host-resident profiles (pseudo_mcmc.rs:77-138).
int jtk_lc_clustering_profiles(const double *profiles, int n_reads, const uint8_t *tmpl, int Lt, const uint8_t *strands,
                               const jtk_gains *gains_c, const jtk_clustering_config *cfg, uint64_t seed, uint64_t *out_asn,
                               double *out_post, int post_stride, double *out_score, int *out_k, uint32_t *out_probe_pos,
                               int probe_cap, int *out_n_probes) {
    try {
        if (!profiles || !tmpl || !strands || !gains_c || !cfg || !out_asn || !out_post || !out_score || !out_k || n_reads < 0 || Lt < 1) {
            g_lc_error = "null / bad argument"; return JTK_EINVAL;
        }
        const size_t n = (size_t)n_reads, ncol = (size_t)(Lt + 1) * NUM_ROW;
        const size_t copy_num = (size_t)cfg->copy_num;
        if (copy_num < 2) { // :86-88
            for (size_t i = 0; i < n; i++) { out_asn[i] = 0; out_post[i * (size_t)post_stride] = 0.0; }
            *out_score = 0.0; *out_k = 1; if (out_n_probes) *out_n_probes = 0;
            return JTK_OK;
        }
        const Gains gains = to_gains(gains_c);
        Rng rng(seed);
        // compress_small_gains (:141-165) + column_sum (:577-588) + strand/sign counts (:314-322)
        const std::vector<size_t> homop = homopolymer_length(tmpl, (size_t)Lt);
        std::vector<double> min_req(ncol);
        for (size_t pos = 0; pos < ncol; pos++) {
            size_t bp; DiffType t;
            pos_to_bp_and_difftype(pos, bp, t);
            const size_t hl = bp < homop.size() ? homop[bp] : 1;
            min_req[pos] = gains.expected(hl, t) * MIN_REQ_FRACTION;
        }
        std::vector<ColStat> stats(ncol, ColStat{ 0.0, 0, { 0, 0, 0, 0 } });
        auto value = [&](size_t r, size_t pos) { const double x = profiles[r * ncol + pos]; return std::fabs(x) < min_req[pos] ? 0.0 : x; };
        for (size_t r = 0; r < n; r++)
            for (size_t pos = 0; pos < ncol; pos++) {
                const double x = value(r, pos);
                if (POS_THR < x) { stats[pos].sum += x; stats[pos].count++; }
                if (std::fabs(x) > 0.0001) stats[pos].sc[(strands[r] ? 2 : 0) + (std::signbit(x) ? 0 : 1)]++;
            }
        const std::vector<Probe> probes = candidate_probes(tmpl, (size_t)Lt, stats, n, gains, copy_num, cfg->coverage);
        const size_t M = probes.size();
        std::vector<double> cand(n * M);
        for (size_t r = 0; r < n; r++) for (size_t m = 0; m < M; m++) cand[r * M + m] = value(r, probes[m].pos);
        const std::vector<size_t> picked = pick_filtered_profiles(probes, cand, n, copy_num);
        Mat variants(n);
        VarTypes vt;
        for (size_t m : picked) {
            size_t bp; DiffType t;
            pos_to_bp_and_difftype(probes[m].pos, bp, t);
            vt.push_back({ bp < homop.size() ? homop[bp] : 0, t });
        }
        for (size_t r = 0; r < n; r++) for (size_t m : picked) variants[r].push_back(cand[r * M + m]);
        const DevResult res = clustering_tail(variants, vt, copy_num, cfg->coverage, cfg->local_coverage, gains, rng);
        return write_result(res, probes, picked, out_asn, out_post, post_stride, out_score, out_k, out_probe_pos, probe_cap, out_n_probes);
    } catch (const std::exception &e) {
        g_lc_error = e.what();
        return JTK_EINVAL;
    }
}

// The same call with the profiles resident on the device: per-column statistics and the candidate columns come from
// jtk_batch_colstats / jtk_batch_gather; everything after that is the same host code.
int jtk_lc_clustering_batch(jtk_batch *b, int tmpl_index, const uint8_t *tmpl, int Lt, int n_reads, const jtk_colstat *stats_c,
                            const jtk_gains *gains_c, const jtk_clustering_config *cfg, uint64_t seed, uint64_t *out_asn,
                            double *out_post, int post_stride, double *out_score, int *out_k, uint32_t *out_probe_pos,
                            int probe_cap, int *out_n_probes) {
    try {
        if (!b || !tmpl || !stats_c || !gains_c || !cfg || !out_asn || !out_post || !out_score || !out_k || n_reads < 0 || Lt < 1) {
            g_lc_error = "null / bad argument"; return JTK_EINVAL;
        }
        const size_t n = (size_t)n_reads, ncol = (size_t)(Lt + 1) * NUM_ROW;
        const size_t copy_num = (size_t)cfg->copy_num;
        if (copy_num < 2) {
            for (size_t i = 0; i < n; i++) { out_asn[i] = 0; out_post[i * (size_t)post_stride] = 0.0; }
            *out_score = 0.0; *out_k = 1; if (out_n_probes) *out_n_probes = 0;
            return JTK_OK;
        }
        const Gains gains = to_gains(gains_c);
        Rng rng(seed);
        const std::vector<size_t> homop = homopolymer_length(tmpl, (size_t)Lt);
        std::vector<ColStat> stats(ncol);
        for (size_t pos = 0; pos < ncol; pos++)
            stats[pos] = ColStat{ stats_c[pos].sum, (size_t)stats_c[pos].count,
                                  { stats_c[pos].sc[0], stats_c[pos].sc[1], stats_c[pos].sc[2], stats_c[pos].sc[3] } };
        const std::vector<Probe> probes = candidate_probes(tmpl, (size_t)Lt, stats, n, gains, copy_num, cfg->coverage);
        const size_t M = probes.size();
        std::vector<double> cand(n * M);
        if (M > 0) {
            std::vector<uint32_t> cols(M);
            for (size_t m = 0; m < M; m++) cols[m] = (uint32_t)probes[m].pos;
            std::vector<float> min_req((size_t)3 * gains.H);
            for (int t = 0; t < 3; t++) for (int h = 0; h < gains.H; h++) min_req[(size_t)t * gains.H + h] = (float)(gains.gain[t][h] * MIN_REQ_FRACTION);
            const int rc = jtk_batch_gather(b, tmpl_index, min_req.data(), gains.H, cols.data(), (int)M, cand.data());
            if (rc) { g_lc_error = "jtk_batch_gather failed"; return rc; }
        }
        const std::vector<size_t> picked = pick_filtered_profiles(probes, cand, n, copy_num);
        Mat variants(n);
        VarTypes vt;
        for (size_t m : picked) {
            size_t bp; DiffType t;
            pos_to_bp_and_difftype(probes[m].pos, bp, t);
            vt.push_back({ bp < homop.size() ? homop[bp] : 0, t });
        }
        for (size_t r = 0; r < n; r++) for (size_t m : picked) variants[r].push_back(cand[r * M + m]);
        const DevResult res = clustering_tail(variants, vt, copy_num, cfg->coverage, cfg->local_coverage, gains, rng);
        return write_result(res, probes, picked, out_asn, out_post, post_stride, out_score, out_k, out_probe_pos, probe_cap, out_n_probes);
    } catch (const std::exception &e) {
        g_lc_error = e.what();
        return JTK_EINVAL;
    }
}

// pseudo_mcmc::clustering (:77-107) from the output of search_variants: variants[r * stride + d] is the compressed
// profile value of read r at the selected flat position probe_pos[d] (jtk_batch_search_variants).
static int clustering_variants_impl(const double *variants, int n_reads, int n_probes, int stride, const uint32_t *probe_pos,
                                    const uint8_t *tmpl, int Lt, const jtk_gains *gains_c, const jtk_clustering_config *cfg,
                                    Rng &rng, uint64_t *out_asn, double *out_post, int post_stride, double *out_score, int *out_k) {
    try {
        if (!tmpl || !gains_c || !cfg || !out_asn || !out_post || !out_score || !out_k || n_reads < 0 || Lt < 1 || n_probes < 0 ||
            (n_probes > 0 && (!variants || !probe_pos || stride < n_probes))) {
            g_lc_error = "null / bad argument"; return JTK_EINVAL;
        }
        const size_t n = (size_t)n_reads, copy_num = (size_t)cfg->copy_num;
        if (copy_num < 2) { // :86-88
            for (size_t i = 0; i < n; i++) { out_asn[i] = 0; out_post[i * (size_t)post_stride] = 0.0; }
            *out_score = 0.0; *out_k = 1;
            return JTK_OK;
        }
        const Gains gains = to_gains(gains_c);
        const std::vector<size_t> homop = homopolymer_length(tmpl, (size_t)Lt);
        Mat vars(n);
        VarTypes vt;
        for (int d = 0; d < n_probes; d++) {
            size_t bp; DiffType t;
            pos_to_bp_and_difftype(probe_pos[d], bp, t);
            vt.push_back({ bp < homop.size() ? homop[bp] : 0, t });
        }
        for (size_t r = 0; r < n; r++) vars[r].assign(variants + r * (size_t)stride, variants + r * (size_t)stride + n_probes);
        const DevResult res = clustering_tail(vars, vt, copy_num, cfg->coverage, cfg->local_coverage, gains, rng);
        if ((int)res.k > post_stride) { g_lc_error = "post_stride smaller than the cluster number"; return JTK_EINVAL; }
        for (size_t i = 0; i < n; i++) {
            out_asn[i] = res.asn[i];
            for (size_t c = 0; c < res.k; c++) out_post[i * (size_t)post_stride + c] = res.gains[i][c];
        }
        *out_score = res.score; *out_k = (int)res.k;
        return JTK_OK;
    } catch (const std::exception &e) {
        g_lc_error = e.what();
        return JTK_EINVAL;
    }
}

int jtk_lc_clustering_variants(const double *variants, int n_reads, int n_probes, int stride, const uint32_t *probe_pos,
                               const uint8_t *tmpl, int Lt, const jtk_gains *gains_c, const jtk_clustering_config *cfg,
                               uint64_t seed, uint64_t *out_asn, double *out_post, int post_stride, double *out_score, int *out_k) {
    Rng rng(seed);
    return clustering_variants_impl(variants, n_reads, n_probes, stride, probe_pos, tmpl, Lt, gains_c, cfg, rng, out_asn, out_post,
                                    post_stride, out_score, out_k);
}

// The same with the caller's generator state (four Xoshiro256** words, in/out): clustering_recursive threads ONE rng
// through every level of the recursion (local_clustering/mod.rs:97,114,143,159).
void jtk_lc_rng_seed(uint64_t seed, uint64_t *state4) {
    Rng rng(seed);
    std::memcpy(state4, rng.s, 32);
}
int jtk_lc_clustering_variants_rng(const double *variants, int n_reads, int n_probes, int stride, const uint32_t *probe_pos,
                                   const uint8_t *tmpl, int Lt, const jtk_gains *gains_c, const jtk_clustering_config *cfg,
                                   uint64_t *state4, uint64_t *out_asn, double *out_post, int post_stride, double *out_score,
                                   int *out_k) {
    if (!state4) { g_lc_error = "null rng state"; return JTK_EINVAL; }
    Rng rng(0);
    std::memcpy(rng.s, state4, 32);
    const int rc = clustering_variants_impl(variants, n_reads, n_probes, stride, probe_pos, tmpl, Lt, gains_c, cfg, rng, out_asn,
                                            out_post, post_stride, out_score, out_k);
    std::memcpy(state4, rng.s, 32);
    return rc;
}

// exact_clustering::cluster_filtered_variants_exact (exact_clustering.rs:7-77) on a dense n_reads x n_probes matrix
int jtk_lc_cluster_filtered_variants_exact(const double *variants, int n_reads, int n_probes, int stride, int copy_num,
                                           uint64_t *out_asn, double *out_gains, double *out_score) {
    try {
        if (n_reads < 0 || n_probes < 0 || copy_num < 1 || !out_asn || !out_gains || !out_score ||
            (n_reads > 0 && n_probes > 0 && (!variants || stride < n_probes))) { g_lc_error = "null / bad argument"; return JTK_EINVAL; }
        Mat vars((size_t)n_reads);
        for (int r = 0; r < n_reads; r++)
            if (n_probes > 0) vars[(size_t)r].assign(variants + (size_t)r * stride, variants + (size_t)r * stride + n_probes);
        const DevResult res = cluster_filtered_variants_exact(vars, (size_t)n_probes, (size_t)copy_num);
        for (int r = 0; r < n_reads; r++) {
            out_asn[r] = res.asn[(size_t)r];
            for (int c = 0; c < copy_num; c++) out_gains[(size_t)r * copy_num + c] = res.gains[(size_t)r][(size_t)c];
        }
        *out_score = res.score;
        return JTK_OK;
    } catch (const std::exception &e) { g_lc_error = e.what(); return JTK_EINVAL; }
}

// Host twin of one chain of jtk_mcmc_restarts_batch (tests: device == host, bit for bit)
int jtk_lc_mcmc_restarts_host(const double *data, int n, int D, int k, double cov, int restarts, uint64_t *state4,
                              uint8_t *out_asn, double *out_lk) {
    try {
        if (!data || !state4 || !out_asn || !out_lk || n < 1 || D < 1 || k < 1) { g_lc_error = "bad argument"; return JTK_EINVAL; }
        Mat m((size_t)n);
        for (int i = 0; i < n; i++) m[(size_t)i].assign(data + (size_t)i * D, data + (size_t)(i + 1) * D);
        Rng rng(0);
        std::memcpy(rng.s, state4, 32);
        std::vector<size_t> best; double best_lk = 0;
        mcmc_restarts(m, (size_t)k, cov, rng, restarts, best, best_lk);
        std::memcpy(state4, rng.s, 32);
        for (int i = 0; i < n; i++) out_asn[i] = (uint8_t)best[(size_t)i];
        *out_lk = best_lk;
        return JTK_OK;
    } catch (const std::exception &e) { g_lc_error = e.what(); return JTK_EINVAL; }
}
// Host twin of the speculative schedule of mcmc_speculative_kernel (two clusters): the same results as
// jtk_lc_mcmc_restarts_host; out_stats = { rounds, proposals, slow-path rounds } (proposals / rounds = the speed-up of a round)
int jtk_lc_mcmc_restarts_spec_host(const double *data, int n, int D, double cov, int restarts, int window, int spec, uint64_t *state4,
                                   uint8_t *out_asn, double *out_lk, uint64_t *out_stats) {
    try {
        if (!data || !state4 || !out_asn || !out_lk || n < 1 || D < 1 || window < 3 || window > 32 || spec < 1 || spec > kSpec) { g_lc_error = "bad argument"; return JTK_EINVAL; }
        Mat m((size_t)n);
        for (int i = 0; i < n; i++) m[(size_t)i].assign(data + (size_t)i * D, data + (size_t)(i + 1) * D);
        Rng rng(0);
        std::memcpy(rng.s, state4, 32);
        std::vector<size_t> best; double best_lk = 0;
        SpecStats stats;
        mcmc_restarts_spec2(m, cov, rng, restarts, best, best_lk, stats, window, spec);
        std::memcpy(state4, rng.s, 32);
        for (int i = 0; i < n; i++) out_asn[i] = (uint8_t)best[(size_t)i];
        *out_lk = best_lk;
        if (out_stats) { out_stats[0] = stats.rounds; out_stats[1] = stats.proposals; out_stats[2] = stats.slow; }
        return JTK_OK;
    } catch (const std::exception &e) { g_lc_error = e.what(); return JTK_EINVAL; }
}
void jtk_lc_size_to_lk(int n, double cov, int k, double *out) { // max_poisson_lk(x, cov, 1, k), x = 0..n
    for (int x = 0; x <= n; x++) out[x] = max_poisson_lk((size_t)x, cov, 1, (size_t)k);
}

// Independent per-chunk host work of the batch call below on a few threads (the first exception is rethrown on the caller).
static void parallel_chunks(size_t n, const std::function<void(size_t)> &fn) {
    unsigned hw = std::thread::hardware_concurrency();
    if (const char *v = std::getenv("JTK_CLUSTER_THREADS")) { const int t = std::atoi(v); if (t > 0) hw = (unsigned)t; }
    const size_t T = std::min<size_t>({ (size_t)(hw ? hw : 1), (size_t)8, (n + 63) / 64 });
    if (T <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::vector<std::thread> th;
    std::vector<std::string> errs(T);
    std::vector<char> failed(T, 0);
    for (size_t t = 0; t < T; t++)
        th.emplace_back([&, t]() {
            try { for (size_t i = t; i < n; i += T) fn(i); }
            catch (const std::exception &e) { errs[t] = e.what(); failed[t] = 1; }
        });
    for (auto &x : th) x.join();
    for (size_t t = 0; t < T; t++) if (failed[t]) throw Panic(errs[t]);
}

// pseudo_mcmc::clustering after search_variants (clustering_variants_impl) for MANY chunks, with the k-means / MCMC
// restarts of every chunk on the GPU (jtk_mcmc_restarts_batch): cluster_filtered_variants (:213-274) is walked level by
// level -- all chunks that still try cluster number k run their 20 restarts side by side, then each decides on the host
// whether to go on -- so every chunk sees exactly the generator stream and the decisions of the per-chunk call.
int jtk_lc_clustering_variants_batch(jtk_ctx *ctx, int n_chunks, const double *variants_concat, const uint64_t *var_off,
                                     const int32_t *n_reads, const int32_t *n_probes, const uint32_t *probe_pos_concat,
                                     const uint64_t *ppos_off, const uint8_t *tmpl_concat, const uint64_t *tmpl_off,
                                     const jtk_gains *gains_c, const jtk_clustering_config *cfgs, uint64_t *states,
                                     uint64_t *out_asn_concat, const uint64_t *asn_off, double *out_post_concat,
                                     const uint64_t *post_off, int post_stride, double *out_score, int32_t *out_k) {
    try {
        if (!ctx || n_chunks < 0 || !var_off || !n_reads || !n_probes || !ppos_off || !tmpl_concat || !tmpl_off || !gains_c || !cfgs ||
            !states || !out_asn_concat || !asn_off || !out_post_concat || !post_off || !out_score || !out_k || post_stride < 1) {
            g_lc_error = "null / bad argument"; return JTK_EINVAL;
        }
        const Gains gains = to_gains(gains_c);
        struct Job {
            Mat vars; VarTypes vt; size_t copy_num = 0; double coverage = 0, per_cluster = 0;
            size_t k = 0, end = 0, max_k = 1; double mx = 0; bool active = false;
            std::vector<size_t> assignments; std::vector<bool> prev_used;
            DevResult res;
        };
        std::vector<Job> jobs((size_t)n_chunks);
        parallel_chunks((size_t)n_chunks, [&](size_t gi) {
            const int g = (int)gi;
            Job &j = jobs[(size_t)g];
            const size_t n = (size_t)n_reads[g], D = (size_t)n_probes[g];
            j.copy_num = (size_t)cfgs[g].copy_num; j.coverage = cfgs[g].coverage; j.per_cluster = cfgs[g].local_coverage;
            if (j.copy_num < 2) { // clustering (:86-88)
                j.res = { std::vector<size_t>(n, 0), Mat(n, std::vector<double>(1, 0.0)), 0.0, 1 };
                return;
            }
            const uint8_t *tmpl = tmpl_concat + tmpl_off[g];
            const size_t Lt = (size_t)(tmpl_off[g + 1] - tmpl_off[g]);
            const std::vector<size_t> homop = homopolymer_length(tmpl, Lt);
            for (size_t d = 0; d < D; d++) {
                size_t bp; DiffType t;
                pos_to_bp_and_difftype(probe_pos_concat[ppos_off[g] + d], bp, t);
                j.vt.push_back({ bp < homop.size() ? homop[bp] : 0, t });
            }
            j.vars.assign(n, std::vector<double>());
            for (size_t r = 0; r < n; r++) j.vars[r].assign(variants_concat + var_off[g] + r * D, variants_concat + var_off[g] + (r + 1) * D);
            // cluster_filtered_variants (:213-274), the part before the loop
            if (D == 0 || n <= j.copy_num) {
                j.res = { std::vector<size_t>(n, 0), Mat(n, std::vector<double>(1, 0.0)), 0.0, 1 };
                return;
            }
            j.assignments.assign(n, 0);
            j.prev_used.assign(D, false);
            j.end = std::min(j.copy_num, 1 + 2 * j.vt.size());
            j.k = std::max<size_t>(j.end, 5) - 3;
            j.active = j.k <= j.end;
        });
        for (;;) { // one level: every active chunk runs the restarts for its current k
            std::vector<int> act;
            for (int g = 0; g < n_chunks; g++) if (jobs[(size_t)g].active) act.push_back(g);
            if (act.empty()) break;
            std::vector<double> data, s2l, lk(act.size());
            std::vector<uint64_t> doff, rng(4 * act.size());
            std::vector<uint32_t> nr, nc, kk;
            std::vector<int> err(act.size());
            size_t n_asn = 0;
            for (size_t a = 0; a < act.size(); a++) {
                const Job &j = jobs[(size_t)act[a]];
                doff.push_back(data.size());
                for (const auto &row : j.vars) data.insert(data.end(), row.begin(), row.end());
                nr.push_back((uint32_t)j.vars.size()); nc.push_back((uint32_t)j.vars[0].size()); kk.push_back((uint32_t)j.k);
                for (size_t x = 0; x <= j.vars.size(); x++) s2l.push_back(max_poisson_lk(x, j.coverage, 1, j.k));
                std::memcpy(&rng[4 * a], states + 4 * (size_t)act[a], 32);
                n_asn += j.vars.size();
            }
            std::vector<uint8_t> asn(n_asn);
            const int rc = jtk_mcmc_restarts_batch(ctx, (int)act.size(), data.data(), doff.data(), nr.data(), nc.data(), kk.data(),
                                                   s2l.data(), 20, rng.data(), asn.data(), lk.data(), err.data());
            if (rc) { g_lc_error = std::string("jtk_mcmc_restarts_batch: ") + jtk_last_error(ctx); return rc; }
            std::vector<size_t> apos_of(act.size() + 1, 0);
            for (size_t a = 0; a < act.size(); a++) apos_of[a + 1] = apos_of[a] + jobs[(size_t)act[a]].vars.size();
            parallel_chunks(act.size(), [&](size_t a) {
                Job &j = jobs[(size_t)act[a]];
                const size_t n = j.vars.size(), apos = apos_of[a];
                if (err[a] != 0) throw Panic("assertion failed on the device (mcmc status " + std::to_string(err[a]) + ")");
                std::memcpy(states + 4 * (size_t)act[a], &rng[4 * a], 32);
                std::vector<size_t> best(n);
                for (size_t i = 0; i < n; i++) best[i] = asn[apos + i];
                ClusterOut c = mcmc_finish(j.vars, j.k, j.coverage, best, lk[a]);
                if (j.k == 2) {
                    ClusterOut h = use_highest_gain(j.vars);
                    if (c.score < h.score) c = h;
                }
                (void)min_gain(gains, j.vt, c.used); // as the reference (its result only feeds a trace! line)
                const double expected_gain = expected_gains(gains, j.vt, j.prev_used, c.used) * j.per_cluster + 0.1;
                if (expected_gain < c.score - j.mx) {
                    j.assignments = c.asn; j.mx = c.score; j.max_k = j.k; j.prev_used = c.used;
                    j.k++;
                    j.active = j.k <= j.end;
                } else j.active = false;
                if (!j.active) j.res = { j.assignments, get_likelihood_gain(j.vars, j.assignments, j.max_k), j.mx, j.max_k };
            });
        }
        for (int g = 0; g < n_chunks; g++)
            if ((int)jobs[(size_t)g].res.k > post_stride) { g_lc_error = "post_stride smaller than the cluster number"; return JTK_EINVAL; }
        parallel_chunks((size_t)n_chunks, [&](size_t gi) { // clustering (:77-107): re-assignment to the best cluster, log-posteriors
            const int g = (int)gi;
            Job &j = jobs[(size_t)g];
            DevResult &r = j.res;
            if (j.copy_num >= 2) {
                for (size_t i = 0; i < r.asn.size(); i++) {
                    const auto &lks = r.gains[i];
                    size_t bi = 0;
                    for (size_t c = 0; c < lks.size(); c++) if (!(lks[c] < lks[bi])) bi = c; // last maximum
                    if (lks[r.asn[i]] + 0.001 < lks[bi]) r.asn[i] = bi;
                }
                for (auto &xs : r.gains) { const double tot = logsumexp(xs); for (double &x : xs) x -= tot; }
            }
            for (size_t i = 0; i < r.asn.size(); i++) {
                out_asn_concat[asn_off[g] + i] = r.asn[i];
                for (size_t c = 0; c < r.k; c++) out_post_concat[post_off[g] + i * (size_t)post_stride + c] = r.gains[i][c];
            }
            out_score[g] = r.score; out_k[g] = (int32_t)r.k;
        });
        return JTK_OK;
    } catch (const std::exception &e) {
        g_lc_error = e.what();
        return JTK_EINVAL;
    }
}

size_t jtk_abi_sizeof(const char *name) {
    if (!name) return 0;
    const std::string n(name);
    if (n == "jtk_hmm_params") return sizeof(jtk_hmm_params);
    if (n == "jtk_colstat") return sizeof(jtk_colstat);
    if (n == "jtk_candidate") return sizeof(jtk_candidate);
    if (n == "jtk_gains") return sizeof(jtk_gains);
    if (n == "jtk_clustering_config") return sizeof(jtk_clustering_config);
    if (n == "jtk_polish_config") return sizeof(jtk_polish_config);
    return 0;
}

// test hooks for the reference's own unit tests on these files (pseudo_mcmc.rs:876-904)
double jtk_lc_cosine_similarity(const double *profiles, int n, int ncol, int i, int j) {
    std::vector<double> cand((size_t)n * 2);
    for (int r = 0; r < n; r++) { cand[(size_t)r * 2] = profiles[(size_t)r * ncol + i]; cand[(size_t)r * 2 + 1] = profiles[(size_t)r * ncol + j]; }
    return cosine_similarity(cand, (size_t)n, 2, 0, 1);
}
int jtk_lc_nonmatch_columns(const uint8_t *ops, int n_ops, const uint8_t *read, int Lr, const uint8_t *tmpl, int Lt) {
    if (!ops || n_ops < 0 || Lr < 0 || Lt < 0 || (Lr > 0 && !read) || (Lt > 0 && !tmpl)) return -1;
    int i = 0, j = 0, bad = 0;
    for (int k = 0; k < n_ops; k++) {
        const uint8_t op = ops[k];
        if (op <= JTK_OP_MISMATCH) {
            if (i >= Lr || j >= Lt) return -1;
            bad += std::toupper(read[i]) != std::toupper(tmpl[j]); // Node::recover compares to_ascii_uppercase (definitions/src/lib.rs:777-783)
            i++; j++;
        } else if (op == JTK_OP_INS) { if (i >= Lr) return -1; i++; bad++; }
        else if (op == JTK_OP_DEL) { if (j >= Lt) return -1; j++; bad++; }
        else return -1;
    }
    return bad;
}

int jtk_lc_nonmatch_columns_batch(int n, const uint8_t *ops_concat, const uint64_t *ops_off, const uint8_t *read_concat,
                                  const uint64_t *read_off, const uint8_t *tmpl_concat, const uint64_t *tmpl_off,
                                  const uint32_t *tmpl_idx, int32_t *out) {
    if (n < 0 || (n > 0 && (!ops_concat || !ops_off || !read_concat || !read_off || !tmpl_concat || !tmpl_off || !tmpl_idx || !out))) return JTK_EINVAL;
    // nodes are independent: the reference computes these keys inside sort_by_cached_key on its rayon workers
    auto range = [&](int lo, int hi) {
        for (int k = lo; k < hi; k++) {
            const uint32_t t = tmpl_idx[k];
            out[k] = jtk_lc_nonmatch_columns(ops_concat + ops_off[k], (int)(ops_off[k + 1] - ops_off[k]), read_concat + read_off[k],
                                             (int)(read_off[k + 1] - read_off[k]), tmpl_concat + tmpl_off[t],
                                             (int)(tmpl_off[t + 1] - tmpl_off[t]));
        }
    };
    int nt = (int)std::thread::hardware_concurrency();
    if (const char *env = std::getenv("JTK_CLUSTER_THREADS")) nt = std::atoi(env);
    else if (const char *env2 = std::getenv("JTK_HOST_THREADS")) nt = std::atoi(env2);
    nt = std::max(1, std::min(std::min(nt, 32), n / 256));
    if (nt <= 1) { range(0, n); return JTK_OK; }
    std::vector<std::thread> th;
    const int per = (n + nt - 1) / nt;
    for (int t = 1; t < nt; t++) th.emplace_back(range, std::min(n, t * per), std::min(n, (t + 1) * per));
    range(0, std::min(n, per));
    for (auto &x : th) x.join();
    return JTK_OK;
}

// SURVEY 8a K6 (kiley::gen_seq `Generate::gen`; kiley is absent, so the order in which it consumes the generator is unknown:
// this is OUR sampler of the same model, jtk_b200/likelihood_gains.py::gen_reads on host threads).  One read per source
// sequence: start in Match, transitions row-normalised; Match emits by mat_emit[template base], Ins by ins_emit[previous
// read base] (4 = none), Del emits nothing; the read ends when the template is consumed.  Read k is drawn from source
// src_idx[k] and has its own generator
// (Xoshiro256** seeded with seed + k), so the result does not depend on the thread count.
int jtk_lc_gen_reads(const double *hmm45, int n, const uint8_t *src_concat, const uint64_t *src_off, const uint32_t *src_idx,
                     uint64_t seed, int cap, uint8_t *out, uint32_t *out_len) {
    if (!hmm45 || n < 0 || cap <= 0 || (n > 0 && (!src_concat || !src_off || !src_idx || !out || !out_len))) return JTK_EINVAL;
    double tc[3][3], mc[4][4], ic[5][4];
    auto cum = [](const double *row, int m, double *dst) {
        double tot = 0; for (int k = 0; k < m; k++) tot += row[k];
        double acc = 0; for (int k = 0; k < m; k++) { acc += row[k] / tot; dst[k] = acc; }
    };
    for (int r = 0; r < 3; r++) cum(hmm45 + 3 * r, 3, tc[r]);
    for (int r = 0; r < 4; r++) cum(hmm45 + 9 + 4 * r, 4, mc[r]);
    for (int r = 0; r < 5; r++) cum(hmm45 + 25 + 4 * r, 4, ic[r]);
    static const uint8_t kAcgt[4] = { 'A', 'C', 'G', 'T' };
    auto code = [](uint8_t b) -> int { switch (b | 0x20) { case 'a': return 0; case 'c': return 1; case 'g': return 2; default: return 3; } };
    auto pick = [](const double *c, int m, double u) -> int { int k = 0; while (k < m - 1 && u > c[k]) k++; return k; };
    auto range = [&](int lo, int hi) {
        for (int k = lo; k < hi; k++) {
            Rng rng(seed + (uint64_t)k);
            auto unif = [&]() -> double { return (double)(rng.next_u64() >> 11) * (1.0 / 9007199254740992.0); };
            const uint32_t si = src_idx[k];
            const uint8_t *t = src_concat + src_off[si];
            const size_t L = (size_t)(src_off[si + 1] - src_off[si]);
            uint8_t *o = out + (size_t)k * (size_t)cap;
            size_t j = 0; int state = 0, prev = 4; uint32_t len = 0;
            while (j < L) {
                const int nxt = pick(tc[state], 3, unif());
                const double v = unif();
                if (nxt == 0) { const int e = pick(mc[code(t[j])], 4, v); if (len < (uint32_t)cap) { o[len++] = kAcgt[e]; prev = e; } j++; }
                else if (nxt == 1) { const int e = pick(ic[prev], 4, v); if (len < (uint32_t)cap) { o[len++] = kAcgt[e]; prev = e; } }
                else j++;
                state = nxt;
            }
            out_len[k] = len;
        }
    };
    int nt = (int)std::thread::hardware_concurrency();
    if (const char *env = std::getenv("JTK_CLUSTER_THREADS")) nt = std::atoi(env);
    else if (const char *env2 = std::getenv("JTK_HOST_THREADS")) nt = std::atoi(env2);
    nt = std::max(1, std::min(std::min(nt, 32), n / 256));
    if (nt <= 1) { range(0, n); return JTK_OK; }
    std::vector<std::thread> th;
    const int per = (n + nt - 1) / nt;
    for (int t = 1; t < nt; t++) th.emplace_back(range, std::min(n, t * per), std::min(n, (t + 1) * per));
    range(0, std::min(n, per));
    for (auto &x : th) x.join();
    return JTK_OK;
}

int jtk_lc_homopolymer_length(const uint8_t *xs, int n, uint32_t *out) {
    const std::vector<size_t> h = homopolymer_length(xs, (size_t)n);
    for (int i = 0; i < n; i++) out[i] = (uint32_t)h[(size_t)i];
    return 0;
}
// raw generator words (test hook: Xoshiro256** reference vector, SplitMix64 seeding)
void jtk_lc_rng_words(uint64_t seed, int use_state, const uint64_t *state, int n, uint64_t *out) {
    Rng r(seed);
    if (use_state) std::memcpy(r.s, state, 32);
    for (int i = 0; i < n; i++) out[i] = r.next_u64();
}

} // extern "C"
