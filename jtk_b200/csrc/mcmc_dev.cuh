// mcmc_dev.cuh -- layout shared by mcmc_kernels.cu and the C-ABI host code (jtk_mcmc_restarts_batch).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace jtk {

struct McmcChain {            // one chain = one (chunk, cluster number) problem; offsets into the workspace arrays
    uint32_t n, D, k;         // reads, variant columns, clusters
    uint32_t pad_;
    double cov;               // haploid coverage (kept for diagnostics: size_to_lk comes from the host)
    uint64_t off_f64;         // doubles (global, read-only): flat[n*D] | size_to_lk[n+1]
    uint64_t off_u32;         // unused
    uint64_t off_u8;          // bytes (global, written once by the kernel): pinc[n*D] | ninc[n*D] | ratio_ok[(n+1)^2]
};
constexpr int kMcmcMaxK = 8;  // clusters per chain (mcmc_kernels.cu keeps the column statistics of all clusters in registers)
constexpr int kMcmcMaxD = 32; // variant columns per chain (one lane each)
inline size_t mcmc_f64_words(uint32_t n, uint32_t D) { return (size_t)n * D + (n + 1); }
inline size_t mcmc_u8_bytes(uint32_t n, uint32_t D) { return 2 * (size_t)n * D + (size_t)(n + 1) * (n + 1); }
size_t mcmc_smem_bytes(uint32_t n, uint32_t D, uint32_t k); // shared memory of one chain: k-means scratch + assignments

// kernel class of a chain: 0..3 = two clusters and <= 2 / 4 / 6 / 8 columns (four chains per warp), 4 = one warp per chain
int mcmc_class_of(uint32_t n, uint32_t D, uint32_t k);
// ids / host_ids: the chain indices grouped by class (class_count[c] of them), device and host copies
cudaError_t launch_mcmc_restarts(const McmcChain *chains, const McmcChain *host_chains, const int *ids, const int *host_ids,
                                 const int class_count[5], int n_chains, double *wf64, uint8_t *wu8, uint64_t *rng_state,
                                 uint8_t *out_asn, const uint64_t *asn_off, double *out_lk, int *out_err, int restarts,
                                 size_t smem_per_chain, cudaStream_t st);

} // namespace jtk
