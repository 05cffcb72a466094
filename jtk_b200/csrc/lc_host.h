// lc_host.h -- pieces of the host restatement of pseudo_mcmc.rs (local_clustering.cpp) that the device path of
// filter_profiles / search_variants needs: the tables the candidate kernel looks up (so that device and host decisions
// come from the very same f64 values) and the greedy probe pick.
#pragma once
#include "../../include/jtk_gpu.h"

#include <cstddef>
#include <cstdint>
#include <vector>

namespace jtk {
// records `msg` as the context's last error (jtk_last_error) and returns `code`: error paths outside jtk_gpu_api.cu
int ctx_fail(jtk_ctx *ctx, int code, const char *msg);
namespace host {

// Pvalues::pvalue tables (likelihood_gains.rs:115-129,148-158) for `total` reads: out[(type * H + h) * (total + 1) + count]
void pvalue_tables(const jtk_gains *g, size_t total, std::vector<double> &out);
// max over k = 1..cluster_num of poisson_lk(count, coverage * k), count = 0..total (pseudo_mcmc.rs:457-462,636-638)
void poisson_prior_table(double coverage, size_t cluster_num, size_t total, std::vector<double> &out);
// pick_filtered_profiles (pseudo_mcmc.rs:516-575) over M candidate columns (flat positions pos[m], scores lk[m],
// values cand[r * M + m] of n reads); returns indices into the candidates, ascending
std::vector<size_t> pick_probes(const uint32_t *pos, const double *lk, size_t M, const double *cand, size_t n,
                                size_t cluster_num);

} // namespace host
} // namespace jtk
