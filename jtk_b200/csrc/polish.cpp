// polish.cpp -- K3: PairHiddenMarkovModelOnStrands::polish_until_converge_antidiagonal, batched over chunks.
// Reference call sites: haplotyper/src/local_clustering/mod.rs:105-106,154-156, model_tune.rs:141-143,
// consensus/mod.rs:476-483.  kiley's loop body is not observable from the reference (SURVEY A.5 [FREE]); the
// definition implemented here is the one written down in DESIGN.md section 2 and restated in f64 by the oracle:
// the tables and per-column sums come from the GPU, edit selection and the local patch of the guide ops run here.
#include "../../include/jtk_gpu.h"
#include "lc_host.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <memory>
#include <vector>

namespace {

constexpr double kMinGain = 0.1;
constexpr int kMaxIter = 20;
constexpr int NUM_ROW = JTK_NUM_ROW, COPY = JTK_COPY_SIZE;

struct Edit { int j, row; };

constexpr double kTieMargin = 0.01;
constexpr int kSuppress = 10;          // non-maximum suppression radius (columns)
constexpr int kDuplicateSpan = 40;     // an indel of the same row and about the same gain this close to a taken edit is its tandem-repeat twin
constexpr double kTwinTolerance = 0.1; // relative gain difference of twins (the band edges of the reads make them differ by ~1 %)

// Pick over the per-column best rows (jtk_batch_best_edits: the row with the largest summed gain > kMinGain among the rows
// valid at j, first maximum, or -1, and that gain).  An edit is taken iff
//   (1) no candidate within kSuppress columns gains more (ties within kTieMargin: the leftmost wins), and
//   (2) it is not the twin of an edit already taken: an insertion / copy / deletion of the same row with about the same gain
//       (within kTwinTolerance), at most kDuplicateSpan columns to the right -- in a tandem repeat "delete one unit" scores
//       the same at every unit boundary, and applying two of them overshoots.
// Edits taken in one round are therefore more than kSuppress columns apart and do not interact through the band of a
// read.  A scan that took the FIRST column with any positive gain picked a +2-nat insertion two columns left of a
// +100-nat substitution, skipped the real fix, and oscillated between spurious indels for all 20 rounds on ~10 % of the
// 2 kbp drafts with 20 planted errors; whatever is passed over here gets its turn in the next round.
void select_edits(const int8_t *best_row, const double *best_gain, const std::vector<uint8_t> &tmpl, int ignore_edge,
                  std::vector<Edit> &ed) {
    const int L = (int)tmpl.size();
    ed.clear();
    std::vector<double> taken_gain;
    for (int j = ignore_edge; j <= L - ignore_edge; j++) {
        const int best = best_row[j];
        if (best < 0) continue;
        const double g = best_gain[j];
        bool take = true;
        for (int a = std::max(j - kSuppress, ignore_edge); a <= j + kSuppress && a <= L - ignore_edge && take; a++) {
            if (a == j || best_row[a] < 0) continue;
            if (a < j ? best_gain[a] >= g - kTieMargin : best_gain[a] > g + kTieMargin) take = false;
        }
        for (size_t e = ed.size(); take && e-- > 0 && j - ed[e].j <= kDuplicateSpan;)
            if (best >= 4 && ed[e].row == best && std::fabs(taken_gain[e] - g) <= kTwinTolerance * std::max(taken_gain[e], g)) take = false;
        if (take) { ed.push_back({ j, best }); taken_gain.push_back(g); }
    }
}

void apply_edits(std::vector<uint8_t> &tmpl, const std::vector<Edit> &ed) {
    std::vector<uint8_t> next;
    next.reserve(tmpl.size() + 3 * ed.size());
    size_t src = 0;
    for (const Edit &e : ed) {
        while (src < (size_t)e.j) next.push_back(tmpl[src++]);
        if (e.row < 4) { next.push_back((uint8_t)"ACGT"[e.row]); src++; }
        else if (e.row < 8) next.push_back((uint8_t)"ACGT"[e.row - 4]);
        else if (e.row < 8 + COPY) { for (int x = 0; x < e.row - 7; x++) next.push_back(tmpl[(size_t)e.j + x]); }
        else src += (size_t)(e.row - 7 - COPY);
    }
    while (src < tmpl.size()) next.push_back(tmpl[src++]);
    tmpl.swap(next);
}

// rewrite one read's guide path after the template edits (sorted, old coordinates)
bool update_ops(const uint8_t *ops, int n, const std::vector<Edit> &ed, std::vector<uint8_t> &out) {
    out.clear();
    int k = 0, tpos = 0, del_left = 0;
    size_t e = 0;
    for (;;) {
        while (e < ed.size() && ed[e].j == tpos && del_left == 0) {
            const int row = ed[e].row;
            e++;
            if (row < 4) continue;
            if (row < 8 + COPY) {
                const int nb = row < 8 ? 1 : row - 7;
                for (int x = 0; x < nb; x++) {
                    if (k < n && ops[k] == JTK_OP_INS) { out.push_back(JTK_OP_MATCH); k++; }
                    else out.push_back(JTK_OP_DEL);
                }
            } else del_left = row - 7 - COPY;
        }
        if (k >= n) break;
        const uint8_t op = ops[k++];
        if (op == JTK_OP_INS) { out.push_back(JTK_OP_INS); continue; }
        if (del_left > 0) { if (op != JTK_OP_DEL) out.push_back(JTK_OP_INS); del_left--; }
        else out.push_back(op);
        tpos++;
    }
    return true;
}

} // namespace

extern "C" int jtk_polish_until_converge_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_chunks,
                                               const uint8_t *draft_concat, const uint32_t *draft_off, int n_pairs,
                                               const uint8_t *read_concat, const uint32_t *read_off, uint8_t *ops_buf,
                                               const uint64_t *ops_pos, const uint32_t *ops_cap, uint32_t *n_ops,
                                               const uint8_t *strand, const uint32_t *tmpl_idx, const jtk_polish_config *cfg,
                                               uint8_t *out_cons, const uint64_t *cons_pos, const uint32_t *cons_cap,
                                               uint32_t *out_len, int32_t *out_iters) {
    if (!ctx || !fwd || !rev || !cfg || n_chunks < 0 || n_pairs < 0) return jtk::ctx_fail(ctx, JTK_EINVAL, "polish: null argument or negative count");
    if (n_chunks == 0) return JTK_OK;
    if (!draft_concat || !draft_off || !read_concat || !read_off || !ops_buf || !ops_pos || !ops_cap || !n_ops || !strand ||
        !tmpl_idx || !out_cons || !cons_pos || !cons_cap || !out_len)
        return jtk::ctx_fail(ctx, JTK_EINVAL, "polish: null argument");
    std::vector<std::vector<uint8_t>> tmpl((size_t)n_chunks);
    std::vector<std::vector<uint32_t>> members((size_t)n_chunks);
    for (int c = 0; c < n_chunks; c++) tmpl[(size_t)c].assign(draft_concat + draft_off[c], draft_concat + draft_off[c + 1]);
    for (int p = 0; p < n_pairs; p++) {
        if (tmpl_idx[p] >= (uint32_t)n_chunks) return jtk::ctx_fail(ctx, JTK_EINVAL, "polish: tmpl_idx out of range");
        members[tmpl_idx[p]].push_back((uint32_t)p);
    }
    std::vector<int> iters((size_t)n_chunks, 0);
    std::vector<uint8_t> active((size_t)n_chunks, 1);
    std::vector<uint8_t> t_cat, s_vec;
    struct RawBuf { // staging bytes that are always overwritten: no value-initialisation (a resize() of 0.5 GB is a memset)
        std::unique_ptr<uint8_t[]> p; size_t cap = 0;
        uint8_t *data() { return p.get(); }
        void need(size_t n) { if (n > cap) { cap = n + n / 8 + 64; p.reset(new uint8_t[cap]); } }
    } r_cat, o_cat;
    std::vector<uint32_t> t_off, r_off, o_off, t_idx, chunk_of;
    std::vector<uint64_t> stat_off;
    std::vector<int8_t> best_rows;
    std::vector<double> best_gains;
    const bool timing = std::getenv("JTK_TIMING") != nullptr; // wall-clock phases of the loop on stderr
    double t_pack = 0, t_create = 0, t_table = 0, t_pick = 0, t_patch = 0;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    int rounds_run = 0;
    for (int round = 0; round < kMaxIter; round++) {
        const auto p0 = now();
        // batch of the voting reads of every chunk that is still changing
        t_cat.clear(); s_vec.clear();
        if (round == 0) { // the first round carries every chunk: size the staging vectors once (growth by doubling copied ~3x the bytes)
            size_t tb = 0, ob = 0;
            for (int c = 0; c < n_chunks; c++) tb += tmpl[(size_t)c].size();
            for (int p = 0; p < n_pairs; p++) ob += n_ops[p];
            t_cat.reserve(tb + 64 * (size_t)n_chunks); r_cat.need(read_off[n_pairs]); o_cat.need(ob + 64 * (size_t)n_pairs);
            s_vec.reserve((size_t)n_pairs); t_idx.reserve((size_t)n_pairs); r_off.reserve((size_t)n_pairs + 1); o_off.reserve((size_t)n_pairs + 1);
        }
        t_off.assign(1, 0); r_off.assign(1, 0); o_off.assign(1, 0); t_idx.clear(); chunk_of.clear(); stat_off.clear();
        uint64_t so = 0;
        // pass 1: the layout (offsets of every template / read / ops run of the batch); pass 2 copies the bytes on a few threads
        std::vector<uint32_t> src_pair;
        for (int c = 0; c < n_chunks; c++) {
            if (!active[(size_t)c]) continue;
            const uint32_t bt = (uint32_t)chunk_of.size();
            chunk_of.push_back((uint32_t)c);
            t_off.push_back(t_off.back() + (uint32_t)tmpl[(size_t)c].size());
            stat_off.push_back(so);
            so += (uint64_t)tmpl[(size_t)c].size() + 1;
            const size_t take = std::min<size_t>((size_t)std::max(cfg->take_num, 0), members[(size_t)c].size());
            for (size_t m = 0; m < take; m++) {
                const uint32_t p = members[(size_t)c][m];
                src_pair.push_back(p);
                r_off.push_back(r_off.back() + (read_off[p + 1] - read_off[p]));
                o_off.push_back(o_off.back() + n_ops[p]);
                s_vec.push_back(strand[p]);
                t_idx.push_back(bt);
            }
        }
        t_cat.resize(t_off.back()); r_cat.need(r_off.back()); o_cat.need(o_off.back());
        {
            const size_t np = src_pair.size(), nc = chunk_of.size();
            auto copy_range = [&](size_t lo, size_t hi) {
                for (size_t k = lo; k < hi; k++) {
                    const uint32_t p = src_pair[k];
                    std::memcpy(r_cat.data() + r_off[k], read_concat + read_off[p], r_off[k + 1] - r_off[k]);
                    std::memcpy(o_cat.data() + o_off[k], ops_buf + ops_pos[p], o_off[k + 1] - o_off[k]);
                }
            };
            for (size_t bt = 0; bt < nc; bt++) std::memcpy(t_cat.data() + t_off[bt], tmpl[chunk_of[bt]].data(), t_off[bt + 1] - t_off[bt]);
            const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
            const size_t nt = std::min<size_t>(hw, (np + 2047) / 2048);
            if (nt <= 1) copy_range(0, np);
            else {
                std::vector<std::thread> th;
                const size_t per = (np + nt - 1) / nt;
                for (size_t t = 1; t < nt; t++) th.emplace_back(copy_range, std::min(np, t * per), std::min(np, (t + 1) * per));
                copy_range(0, std::min(np, per));
                for (auto &x : th) x.join();
            }
        }
        if (chunk_of.empty()) break;
        rounds_run++;
        const auto p1 = now();
        t_pack += ms(p0, p1);
        jtk_batch *b = nullptr;
        int rc = jtk_batch_create(ctx, (int)t_idx.size(), (int)chunk_of.size(), t_cat.data(), t_off.data(), r_cat.data(),
                                  r_off.data(), o_cat.data(), o_off.data(), s_vec.data(), t_idx.data(), cfg->radius, &b);
        if (rc) return rc;
        const auto p2 = now();
        t_create += ms(p1, p2);
        best_rows.assign((size_t)so, (int8_t)-1);
        best_gains.assign((size_t)so, 0.0);
        if (!t_idx.empty()) {
            rc = jtk_batch_modtable(b, fwd, rev, 14);
            if (!rc) rc = jtk_batch_best_edits(b, cfg->take_num, cfg->ignore_edge, kMinGain, best_rows.data(), best_gains.data(), stat_off.data());
        }
        jtk_batch_destroy(b);
        if (rc) return rc;
        const auto p3 = now();
        t_table += ms(p2, p3);
        // chunks are independent (their reads' ops live in disjoint buffers): patch them on a few threads
        {
            std::atomic<size_t> next(0);
            std::atomic<int> failed(0);
            auto worker = [&]() {
                std::vector<Edit> my_ed;
                std::vector<uint8_t> my_patched;
                for (;;) {
                    const size_t bt = next.fetch_add(1);
                    if (bt >= chunk_of.size()) break;
                    const int c = (int)chunk_of[bt];
                    select_edits(best_rows.data() + stat_off[bt], best_gains.data() + stat_off[bt], tmpl[(size_t)c], cfg->ignore_edge, my_ed);
                    if (my_ed.empty()) { active[(size_t)c] = 0; continue; }
                    iters[(size_t)c]++;
                    for (uint32_t p : members[(size_t)c]) {
                        update_ops(ops_buf + ops_pos[p], (int)n_ops[p], my_ed, my_patched);
                        if (my_patched.size() > ops_cap[p]) { failed.store(1); return; }
                        std::memcpy(ops_buf + ops_pos[p], my_patched.data(), my_patched.size());
                        n_ops[p] = (uint32_t)my_patched.size();
                    }
                    apply_edits(tmpl[(size_t)c], my_ed);
                }
            };
            const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            const size_t nt = std::min<size_t>(hw, (chunk_of.size() + 15) / 16);
            std::vector<std::thread> th;
            for (size_t t = 1; t < nt; t++) th.emplace_back(worker);
            worker();
            for (auto &t : th) t.join();
            if (failed.load()) return jtk::ctx_fail(ctx, JTK_EINVAL, "polish: a read's patched ops exceed ops_cap (ops_buf / n_ops are undefined after a failed call)");
        }
        t_patch += ms(p3, now());
    }
    (void)t_pick;
    if (timing)
        std::fprintf(stderr, "[jtk timing] polish: %d rounds, pack=%.1fms create=%.1fms tables+pick=%.1fms patch=%.1fms\n", rounds_run,
                     t_pack, t_create, t_table, t_patch);
    for (int c = 0; c < n_chunks; c++) {
        if (tmpl[(size_t)c].size() > cons_cap[c]) return jtk::ctx_fail(ctx, JTK_EINVAL, "polish: polished consensus exceeds cons_cap (ops_buf / n_ops are undefined after a failed call)");
        std::memcpy(out_cons + cons_pos[c], tmpl[(size_t)c].data(), tmpl[(size_t)c].size());
        out_len[c] = (uint32_t)tmpl[(size_t)c].size();
        if (out_iters) out_iters[c] = iters[(size_t)c];
    }
    return JTK_OK;
}

// Moves n byte runs (run k: len[k] bytes at buf + pos[k], pos ascending, runs disjoint) to the front of buf, back to back;
// out_off[k] = new start of run k, out_off[n] = total.  The host mirror uses it to hand the patched guide paths of
// jtk_polish_until_converge_batch back as one compact array instead of n padded slots.
extern "C" int jtk_compact_runs(uint8_t *buf, const uint64_t *pos, const uint32_t *len, int n, uint64_t *out_off) {
    if (n < 0 || (n > 0 && (!buf || !pos || !len)) || !out_off) return JTK_EINVAL;
    uint64_t w = 0;
    for (int k = 0; k < n; k++) {
        if (pos[k] < w) return JTK_EINVAL;
        out_off[k] = w;
        if (pos[k] != w) std::memmove(buf + w, buf + pos[k], len[k]);
        w += len[k];
    }
    out_off[n] = w;
    return JTK_OK;
}

// The inverse staging step: run k (src[src_off[k] .. src_off[k + 1])) is copied to buf + pos[k] (the padded per-read ops slots
// of jtk_polish_until_converge_batch), on a few threads -- 1e5 reads x 2 KB per call; the gaps are left as they are.
extern "C" int jtk_scatter_runs(const uint8_t *src, const uint32_t *src_off, int n, uint8_t *buf, const uint64_t *pos) {
    if (n < 0 || (n > 0 && (!src || !src_off || !buf || !pos))) return JTK_EINVAL;
    unsigned hw = std::thread::hardware_concurrency();
    const int T = std::max(1, std::min({ (int)(hw ? hw : 1), 8, (n + 4095) / 4096 }));
    auto work = [&](int t) {
        const int lo = (int)((long long)n * t / T), hi = (int)((long long)n * (t + 1) / T);
        for (int k = lo; k < hi; k++) std::memcpy(buf + pos[k], src + src_off[k], (size_t)(src_off[k + 1] - src_off[k]));
    };
    if (T == 1) { work(0); return JTK_OK; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(work, t);
    for (auto &x : th) x.join();
    return JTK_OK;
}

// Guide ops (values 0..3) at 2 bits per column, four per byte, lowest bits first -- the form in which the per-chunk results
// of a rank travel to rank 0 (the reference's Node.cigar is run-length coded for the same reason).  n_ops ops in, (n_ops + 3) / 4
// bytes out, and back.  `out` of jtk_ops_unpack2 needs room for 4 * ((n_ops + 3) / 4) bytes.
extern "C" int jtk_ops_pack2(const uint8_t *ops, uint64_t n_ops, uint8_t *out) {
    if (n_ops > 0 && (!ops || !out)) return JTK_EINVAL;
    const uint64_t full = n_ops / 4;
    for (uint64_t i = 0; i < full; i++) {
        const uint8_t *o = ops + 4 * i;
        out[i] = (uint8_t)((o[0] & 3) | ((o[1] & 3) << 2) | ((o[2] & 3) << 4) | ((o[3] & 3) << 6));
    }
    if (n_ops % 4) {
        uint8_t b = 0;
        for (uint64_t k = 0; k < n_ops % 4; k++) b |= (uint8_t)((ops[4 * full + k] & 3) << (2 * k));
        out[full] = b;
    }
    return JTK_OK;
}
extern "C" int jtk_ops_unpack2(const uint8_t *packed, uint64_t n_ops, uint8_t *out) {
    if (n_ops > 0 && (!packed || !out)) return JTK_EINVAL;
    static uint32_t lut[256];
    static bool ready = false;
    if (!ready) { // (idempotent: two threads racing here write the same values)
        for (uint32_t b = 0; b < 256; b++) lut[b] = (b & 3u) | (((b >> 2) & 3u) << 8) | (((b >> 4) & 3u) << 16) | (((b >> 6) & 3u) << 24);
        ready = true;
    }
    const uint64_t bytes = (n_ops + 3) / 4;
    for (uint64_t i = 0; i < bytes; i++) std::memcpy(out + 4 * i, &lut[packed[i]], 4);
    return JTK_OK;
}
