/*
 * jtk_gpu.h -- C ABI of libjtkgpu.so: the B200 (sm_100a) drop-in for the per-chunk pair-HMM hot path
 * of ban-m/jtk.  The reference has no FFI of its own: the hot path is plain Rust
 * method calls into the un-vendored crate kiley 0.3.0.  Each entry point below names the reference call
 * site (file:line under /root/reference) whose kiley call it replaces; INTEGRATION.md shows the Rust
 * `extern "C"` shim a maintainer would add.
 *
 * Conventions (SURVEY.md section 8b "Data conventions"):
 *   sequences  ASCII ACGT, reads already in template orientation (definitions/src/lib.rs:678)
 *   ops        one byte per alignment column: 0 Match, 1 Mismatch, 2 Ins (read-only base),
 *              3 Del (template-only base)          (haplotyper/src/misc.rs:167-186)
 *   table      row-minor t[j*NUM_ROW + row], j in 0..=Lt (pseudo_mcmc.rs:125,169,439)
 *   params     the 9+16+20 doubles of HMMParam in declaration order (definitions/src/lib.rs:102-126)
 *   strand     1 = forward model, 0 = reverse model (pseudo_mcmc.rs:58-61)
 * All pointers are HOST pointers unless the name says `dev`.  The caller owns every host buffer; the
 * library never retains a host pointer past the call.  Every function returns 0 on success or a
 * negative JTK_E* code; jtk_last_error(ctx) gives the message.  Functions are thread-safe per distinct
 * ctx, not re-entrant on one ctx.  There is NO CPU fallback: without a CUDA device every compute call
 * fails with JTK_ECUDA.
 */
#ifndef JTK_GPU_H
#define JTK_GPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define JTK_NUM_ROW 14   /* kiley::hmm::NUM_ROW   (pseudo_mcmc.rs:7)   */
#define JTK_COPY_SIZE 3  /* kiley::hmm::COPY_SIZE (pseudo_mcmc.rs:172) */
#define JTK_DEL_SIZE 3
#define JTK_TABLE_NEG (-1.0e10) /* absolute value stored for an impossible edit */

enum { JTK_OP_MATCH = 0, JTK_OP_MISMATCH = 1, JTK_OP_INS = 2, JTK_OP_DEL = 3 };
enum {
    JTK_OK = 0,
    JTK_EINVAL = -1,  /* bad argument (null pointer, ops that do not span the pair, radius out of range) */
    JTK_ECUDA = -2,   /* CUDA runtime error, or no device */
    JTK_ENOMEM = -3,  /* device or host allocation failed */
    JTK_ESTATE = -4   /* handle used in the wrong state */
};

/* definitions/src/lib.rs:102-126 (HMMParam) == kiley::hmm::PairHiddenMarkovModel (model_tune.rs:36-62) */
typedef struct {
    double mat_mat, mat_ins, mat_del;
    double ins_mat, ins_ins, ins_del;
    double del_mat, del_ins, del_del;
    double mat_emit[16]; /* 4*ref + query                         */
    double ins_emit[20]; /* 4*prev_read_base + query, prev=4: none */
} jtk_hmm_params;

typedef struct jtk_ctx jtk_ctx;

/* likelihood_gains::Gains (likelihood_gains.rs:56-62): expected gain and null probability per (DiffType, homopolymer
 * length 1..homop_len); rows in DiffType order Subst, Del, Ins (likelihood_gains.rs:195-199). */
typedef struct {
    int homop_len;
    const double *gain; /* 3 * homop_len */
    const double *prob; /* 3 * homop_len */
} jtk_gains;
/* pseudo_mcmc::ClusteringConfig (pseudo_mcmc.rs:18-43) */
typedef struct {
    int band_width;
    int copy_num;
    double coverage;       /* haploid coverage */
    double local_coverage; /* per-cluster coverage */
} jtk_clustering_config;


/* ---- context ----------------------------------------------------------------------------------- */
/* device < 0: use the current CUDA device.  workspace_bytes = 0: size scratch on demand. */
int jtk_ctx_create(int device, size_t workspace_bytes, jtk_ctx **out);
void jtk_ctx_destroy(jtk_ctx *ctx);
const char *jtk_last_error(const jtk_ctx *ctx); /* ctx may be NULL: last error of jtk_ctx_create */
int jtk_hmm_num_row(void);   /* kiley::hmm::NUM_ROW   */
int jtk_hmm_copy_size(void); /* kiley::hmm::COPY_SIZE */
int jtk_hmm_del_size(void);
/* number of kernels launched by this ctx so far (bench.py's gpu_launches) */
uint64_t jtk_ctx_launch_count(const jtk_ctx *ctx);
/* device time (ms, CUDA events on the ctx stream) of the dominant kernel in the last batch call */
float jtk_ctx_last_kernel_ms(const jtk_ctx *ctx);
/* which modification-table variant the last table call ran: 1 = fused kernel (forward rows recomputed in shared memory, the DP
 * matrices never touch HBM), 2 = forward rows parked in HBM between a forward and a backward kernel (bit-identical results;
 * chosen when the whole batch fits the scratch budget as one wave, or by JTK_MODTABLE=rows|fused), 0 = none yet */
int jtk_ctx_last_modtable_variant(const jtk_ctx *ctx);
/* measurement helpers (bench.py): bracket a region of work on the ctx stream with CUDA events */
int jtk_ctx_timer_start(jtk_ctx *ctx);
int jtk_ctx_timer_stop(jtk_ctx *ctx, float *ms); /* synchronises the stream */
/* per-launch device times (ms) of the pair-HMM kernels launched since the last call (ring of 256);
 * synchronises the stream; returns the number written (<= cap) or a negative error */
int jtk_ctx_kernel_times(jtk_ctx *ctx, float *ms, int cap);
/* FP32 FMA peak of this device measured with a register-resident FFMA loop on every SM (TFLOP/s, 2 flops/FMA):
 * scalar fma.rn.f32 and packed fma.rn.f32x2 (the roofline denominators SURVEY.md 8d asks for) */
int jtk_ctx_measure_fp32_peak(jtk_ctx *ctx, double *tflops_ffma, double *tflops_ffma2);

/* ---- level 1: kiley-shaped batch calls --------------------------------------------------------- */
/*
 * Batched PairHiddenMarkovModel::modification_table_antidiagonal(template, read, ops, band) -> (table, lk)
 *   reference call site: haplotyper/src/local_clustering/pseudo_mcmc.rs:62-63 (one call per read).
 * Pair p uses template tmpl_idx[p] (bytes tmpl_concat[tmpl_off[t] .. tmpl_off[t+1])), read
 * read_concat[read_off[p] .. read_off[p+1]), ops ops_concat[ops_off[p] .. ops_off[p+1]) and the model
 * chosen by strand[p].  out_lk[p] = ln P(read | template).  out_table (may be NULL) receives, at
 * table_off[p], (Lt+1)*JTK_NUM_ROW absolute log-likelihoods (kiley's return value; the caller subtracts lk,
 * pseudo_mcmc.rs:64).
 */
int jtk_hmm_modtable_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_pairs,
                           int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                           const uint8_t *read_concat, const uint32_t *read_off,
                           const uint8_t *ops_concat, const uint32_t *ops_off,
                           const uint8_t *strand, const uint32_t *tmpl_idx, int radius,
                           double *out_lk, double *out_table, const uint64_t *table_off);

/*
 * Batched forward likelihood with guide ops: the arithmetic of
 * PairHiddenMarkovModel::likelihood_antidiagonal_bootstrap(template, read, band)
 *   reference call sites: haplotyper/src/likelihood_gains.rs:27-28,282-283,301-302.
 * ops_concat == NULL selects the bootstrap path: the guide is a banded global edit-distance alignment
 * computed inside the library.
 */
int jtk_hmm_likelihood_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_pairs,
                             int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                             const uint8_t *read_concat, const uint32_t *read_off,
                             const uint8_t *ops_concat, const uint32_t *ops_off,
                             const uint8_t *strand, const uint32_t *tmpl_idx, int radius, double *out_lk);


/* ---- level 2: haplotyper-shaped, device-resident chunk batches ---------------------------------- */
/*
 * A batch is a set of chunks (templates) with their reads, resident in HBM.  It replaces the per-read loop
 * of pseudo_mcmc::modification_table (haplotyper/src/local_clustering/pseudo_mcmc.rs:45-68) for many chunks
 * at once: profiles (table - lk, fp32) stay on the device and only per-column statistics or the selected
 * probe columns cross PCIe (SURVEY.md section 7 "PCIe").
 */
typedef struct jtk_batch jtk_batch;

/* Encode, upload.  Same input arrays as jtk_hmm_modtable_batch; ops are required. */
int jtk_batch_create(jtk_ctx *ctx, int n_pairs, int n_tmpl, const uint8_t *tmpl_concat, const uint32_t *tmpl_off,
                     const uint8_t *read_concat, const uint32_t *read_off,
                     const uint8_t *ops_concat, const uint32_t *ops_off,
                     const uint8_t *strand, const uint32_t *tmpl_idx, int radius, jtk_batch **out);
void jtk_batch_destroy(jtk_batch *b);
/* sum over pairs of 2*C: forward + backward cell updates of one modification-table pass */
uint64_t jtk_batch_cell_updates(const jtk_batch *b);
/* bytes jtk_batch_create copied host->device */
uint64_t jtk_batch_h2d_bytes(const jtk_batch *b);

/* Run forward + backward + table reduction for every pair.  rows = 14 (all) or 9 (substitution, insertion
 * and one-base deletion rows only: the rows filter_profiles reads, pseudo_mcmc.rs:447; other rows are
 * written as impossible).  Profiles and likelihoods stay on the device.  Asynchronous w.r.t. the host
 * until a fetch call or jtk_batch_sync. */
int jtk_batch_modtable(jtk_batch *b, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int rows);
int jtk_batch_sync(jtk_batch *b);
int jtk_batch_fetch_lk(jtk_batch *b, double *out_lk /* n_pairs */);
/* one pair's profile (table - lk), (Lt+1)*JTK_NUM_ROW floats; impossible edits are <= -1e9 */
int jtk_batch_fetch_profile(jtk_batch *b, int pair, float *out);

/*
 * Per-column statistics of one chunk's profiles after compress_small_gains (pseudo_mcmc.rs:141-165):
 * an entry x of read p is zeroed when |x| < min_req[type][min(homop,H)-1] (type 0 Subst, 1 Del, 2 Ins as
 * likelihood_gains::DiffType; min_req = Gains::expected * MIN_REQ_FRACTION), then
 *   sum, count  over entries > pos_thr                                   (column_sum, pseudo_mcmc.rs:577-588)
 *   sc[strand][positive]  counts over entries with |x| > 1e-4           (is_explainable_by_strandedness, :314-339)
 */
typedef struct {
    double sum;
    int32_t count;
    uint16_t sc[4]; /* [strand*2 + is_sign_positive] */
    int32_t pad_;
} jtk_colstat;
/* out receives, for template t at out + stat_off[t], (Lt+1)*JTK_NUM_ROW entries.  out == NULL: compute on the
 * device only (no device->host copy, no synchronisation). */
int jtk_batch_colstats(jtk_batch *b, const float *min_req /* 3*H */, int H, float pos_thr,
                       jtk_colstat *out, const uint64_t *stat_off);
/* compressed profile values of the reads of template t at flat positions cols[0..D): out[n_reads_t * D],
 * reads in batch order (filter_by, pseudo_mcmc.rs:70-75) */
int jtk_batch_gather(jtk_batch *b, int tmpl, const float *min_req, int H, const uint32_t *cols, int D, double *out);

/*
 * filter_profiles (pseudo_mcmc.rs:426-474) for every chunk of the batch on the device: compress_small_gains, column_sum,
 * position mask, row filter, is_in_short_homopolymer, has_small_pvalue, is_explainable_by_strandedness and the Poisson
 * prior; only the surviving candidate columns cross PCIe.  copy_num[t] is the chunk's copy number (cluster_num of
 * filter_profiles), coverage the haploid coverage (ClusteringConfig, pseudo_mcmc.rs:18-43).  Candidates are returned
 * sorted by (tmpl, pos); *out_n receives their number (if it exceeds cap the call fails and *out_n says how many).
 */
typedef struct {
    uint32_t tmpl;  /* chunk index in the batch */
    uint32_t pos;   /* flat position j*NUM_ROW + row */
    uint32_t count; /* reads with a gain > POS_THR in this column */
    uint32_t pad_;
    double sum;     /* their summed gain (column_sum) */
    double lk;      /* sum + max_k poisson_lk(count, coverage*k): the score pick_filtered_profiles ranks by */
} jtk_candidate;
int jtk_batch_candidates(jtk_batch *b, const jtk_gains *gains, const int32_t *copy_num /* n_tmpl */, double coverage,
                         jtk_candidate *out, int cap, int *out_n);
/*
 * pseudo_mcmc::search_variants (pseudo_mcmc.rs:109-138) for every chunk of the batch, all of it on the device: candidates
 * (filter_profiles, :426-474), their values, and the greedy pick_filtered_profiles (:516-575: one warp per chunk, same
 * decisions as the host twin bit for bit; JTK_HOST_PICK=1 keeps the pick on the host).  Only the picks and the n x D
 * variant columns come back.
 * out_n_probes[t] selected columns of chunk t, their flat positions at out_probe_pos[t*probe_cap ..], and for every pair p
 * (batch order) the compressed profile values at out_variants[p*probe_cap ..] (filter_by, :70-75); probe_cap >=
 * 3*max(copy_num, 2).  Chunks with copy_num < 2 are skipped (pseudo_mcmc.rs:86-88).
 */
int jtk_batch_search_variants(jtk_batch *b, const jtk_gains *gains, const int32_t *copy_num, double coverage, int probe_cap,
                              uint32_t *out_n_probes /* n_tmpl */, uint32_t *out_probe_pos /* n_tmpl*probe_cap */,
                              double *out_variants /* n_pairs*probe_cap */);

/* ---- consensus polishing (K3) ------------------------------------------------------------------------- */
/* kiley::hmm::HMMPolishConfig::new(radius, take_num, ignore_edge)
 *   (local_clustering/mod.rs:105,154; model_tune.rs:142; consensus/mod.rs:476) */
typedef struct {
    int radius;
    int take_num;    /* only the first take_num reads vote (pile-ups are pre-sorted, mod.rs:47-50) */
    int ignore_edge; /* bases at both ends of the template that are never edited */
} jtk_polish_config;

/*
 * Batched PairHiddenMarkovModelOnStrands::polish_until_converge_antidiagonal(draft, seqs, &mut ops, strands, cfg)
 *   reference call sites: local_clustering/mod.rs:106,155-156; model_tune.rs:143; consensus/mod.rs:477-483.
 * Every chunk c has a draft (draft_concat[draft_off[c]..draft_off[c+1])) and the reads p with tmpl_idx[p] == c, in
 * batch order.  Loop (the oracle's definition): modification tables of the first take_num
 * reads on the GPU, per-column sums and best rows on the GPU, left-to-right pick of the locally best positive-gain edits on
 * the host (an edit is taken unless one of the next five columns gains more), local patch of every read's ops, until no
 * chunk changes (<= 20 rounds).
 * ops are in/out: pair p owns ops_buf[ops_pos[p] .. ops_pos[p] + ops_cap[p]) and n_ops[p] holds its length.  On a non-zero
 * return ops_buf / n_ops / out_cons are UNDEFINED (some chunks may already be patched): the caller keeps its own copy if it
 * wants to continue after a failure; jtk_last_error(ctx) names the cause.
 * out_cons: chunk c owns out_cons[cons_pos[c] .. cons_pos[c] + cons_cap[c]); out_len[c] is the polished length.
 * out_iters (may be NULL): rounds in which chunk c changed.
 */
int jtk_polish_until_converge_batch(jtk_ctx *ctx, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, int n_chunks,
                                    const uint8_t *draft_concat, const uint32_t *draft_off, int n_pairs,
                                    const uint8_t *read_concat, const uint32_t *read_off, uint8_t *ops_buf,
                                    const uint64_t *ops_pos, const uint32_t *ops_cap, uint32_t *n_ops,
                                    const uint8_t *strand, const uint32_t *tmpl_idx, const jtk_polish_config *cfg,
                                    uint8_t *out_cons, const uint64_t *cons_pos, const uint32_t *cons_cap,
                                    uint32_t *out_len, int32_t *out_iters);
/* Host helper of the call above: moves n byte runs (run k: len[k] bytes at buf + pos[k], pos ascending, disjoint) to the
 * front of buf, back to back; out_off[k] = new start of run k, out_off[n] = total bytes.  Turns the padded per-read ops slots
 * into one compact array. */
int jtk_compact_runs(uint8_t *buf, const uint64_t *pos, const uint32_t *len, int n, uint64_t *out_off);
/* The inverse: run k (src[src_off[k] .. src_off[k + 1])) is copied to buf + pos[k] -- fills the padded per-read ops slots of
 * jtk_polish_until_converge_batch from one compact array, on a few host threads. */
int jtk_scatter_runs(const uint8_t *src, const uint32_t *src_off, int n, uint8_t *buf, const uint64_t *pos);
/* Guide ops (0..3) at 2 bits per column, four per byte, lowest bits first: the form in which a rank's per-chunk results travel
 * to rank 0 in the host gather.  pack2 writes (n_ops + 3) / 4 bytes; unpack2 needs room for 4 * ((n_ops + 3) / 4) bytes. */
int jtk_ops_pack2(const uint8_t *ops, uint64_t n_ops, uint8_t *out);
int jtk_ops_unpack2(const uint8_t *packed, uint64_t n_ops, uint8_t *out);
/* per-column sums over the first take_num reads of every template of a batch (device reduction used by the
 * polish loop): out[stat_off[t] + e] = sum_r profile_r[e] */
int jtk_batch_colsums(jtk_batch *b, int take_num, double *out, const uint64_t *stat_off);
/* The per-column choice of the polish loop on the device: out[col_off[t] + j] = the row (0..13) with the largest gain summed
 * over the first take_num reads of template t, among the rows valid at column j (not the template's own base, within
 * ignore_edge of neither end) and above min_gain, or -1; out_gain (may be NULL) receives that summed gain (0 where the row
 * is -1).  Same sums and tie-break (first maximum) as a host scan of jtk_batch_colsums; tmpl_len[t] + 1 entries per template. */
int jtk_batch_best_edits(jtk_batch *b, int take_num, int ignore_edge, double min_gain, int8_t *out, double *out_gain,
                         const uint64_t *col_off);

/* ---- HMM fit (K4) --------------------------------------------------------------------------------------- */
/* Expected transition / emission counts of every pair of a batch under (fwd, rev), summed per strand:
 * acc90[0..45) forward-strand reads, acc90[45..90) reverse-strand reads; each block is 9 transitions, 16 mat_emit,
 * 20 ins_emit in HMMParam order.  The E-step of one Baum-Welch round. */
int jtk_batch_expected_counts(jtk_batch *b, const jtk_hmm_params *fwd, const jtk_hmm_params *rev, double *acc90);
/*
 * PairHiddenMarkovModelOnStrands::fit_antidiagonal_par_multiple(&mut self, &[TrainingDataPack], radius)
 *   reference call site: haplotyper/src/model_tune.rs:145-151 (TrainingDataPack::new(cons, strands, seqs, ops)).
 * One EM update of both strand models in place: E-step on the GPU over all packs (= templates of the batch inputs),
 * M-step (row normalisation; rows without counts keep their values) on the host.
 */
int jtk_hmm_fit_batch(jtk_ctx *ctx, jtk_hmm_params *fwd, jtk_hmm_params *rev, int n_pairs, int n_tmpl,
                      const uint8_t *tmpl_concat, const uint32_t *tmpl_off, const uint8_t *read_concat,
                      const uint32_t *read_off, const uint8_t *ops_concat, const uint32_t *ops_off,
                      const uint8_t *strand, const uint32_t *tmpl_idx, int radius);

/* ---- host side of local_clustering: everything in pseudo_mcmc.rs that is not the pair HMM ---------------- */
/*
 * pseudo_mcmc::clustering (pseudo_mcmc.rs:77-107) given the per-read profiles (table - lk) of one chunk:
 * compress_small_gains, filter_profiles, pick_filtered_profiles, cluster_filtered_variants, re-assignment and
 * posteriors.  profiles: n_reads x (Lt+1)*JTK_NUM_ROW, row-major.  The generator is
 * Xoshiro256StarStar::seed_from_u64(seed) (local_clustering/mod.rs:97 uses chunk.id * 3490).
 * out_asn[n_reads]; out_post[n_reads * post_stride] holds log-posteriors of the first *out_k clusters per read;
 * out_probe_pos (may be NULL) receives the selected flat positions j*NUM_ROW+row.
 */
int jtk_lc_clustering_profiles(const double *profiles, int n_reads, const uint8_t *tmpl, int Lt, const uint8_t *strands,
                               const jtk_gains *gains, const jtk_clustering_config *cfg, uint64_t seed, uint64_t *out_asn,
                               double *out_post, int post_stride, double *out_score, int *out_k, uint32_t *out_probe_pos,
                               int probe_cap, int *out_n_probes);
/* The same with the chunk's profiles resident in a batch: stats = this template's slice of jtk_batch_colstats;
 * candidate columns are fetched with jtk_batch_gather. */
int jtk_lc_clustering_batch(jtk_batch *b, int tmpl_index, const uint8_t *tmpl, int Lt, int n_reads, const jtk_colstat *stats,
                            const jtk_gains *gains, const jtk_clustering_config *cfg, uint64_t seed, uint64_t *out_asn,
                            double *out_post, int post_stride, double *out_score, int *out_k, uint32_t *out_probe_pos,
                            int probe_cap, int *out_n_probes);
/* pseudo_mcmc::clustering (pseudo_mcmc.rs:77-107) from the output of search_variants: variants[r*stride + d] is the value
 * of read r at the selected position probe_pos[d] (one chunk's rows of jtk_batch_search_variants). */
int jtk_lc_clustering_variants(const double *variants, int n_reads, int n_probes, int stride, const uint32_t *probe_pos,
                               const uint8_t *tmpl, int Lt, const jtk_gains *gains, const jtk_clustering_config *cfg,
                               uint64_t seed, uint64_t *out_asn, double *out_post, int post_stride, double *out_score,
                               int *out_k);
/* the same with the caller's generator (four Xoshiro256** state words, in/out): clustering_recursive threads one rng
 * through every level of the recursion (local_clustering/mod.rs:97,114,143,159) */
void jtk_lc_rng_seed(uint64_t seed, uint64_t *state4);
int jtk_lc_clustering_variants_rng(const double *variants, int n_reads, int n_probes, int stride, const uint32_t *probe_pos,
                                   const uint8_t *tmpl, int Lt, const jtk_gains *gains, const jtk_clustering_config *cfg,
                                   uint64_t *state4, uint64_t *out_asn, double *out_post, int post_stride, double *out_score,
                                   int *out_k);
/* exact_clustering::cluster_filtered_variants_exact (haplotyper/src/local_clustering/exact_clustering.rs:7-77): exhaustive
 * search over one column subset per cluster (2^n_probes masks, non-increasing tuples of copy_num masks); the score the MCMC
 * is compared with in sandbox/src/bin/benchmark_mcmc.rs:112-118.  out_gains[r * copy_num + c] = read r's summed values over
 * cluster c's subset; out_asn[r] its best cluster (last maximum); *out_score = sum over reads of the best value. */
int jtk_lc_cluster_filtered_variants_exact(const double *variants, int n_reads, int n_probes, int stride, int copy_num,
                                           uint64_t *out_asn, double *out_gains, double *out_score);
const char *jtk_lc_last_error(void);
/* hooks for the reference's unit tests on these files (pseudo_mcmc.rs:876-904) and the generator */
double jtk_lc_cosine_similarity(const double *profiles, int n, int ncol, int i, int j);
int jtk_lc_homopolymer_length(const uint8_t *xs, int n, uint32_t *out);
void jtk_lc_rng_words(uint64_t seed, int use_state, const uint64_t *state, int n, uint64_t *out);
/* sort key of pileup_nodes (haplotyper/src/local_clustering/mod.rs:47-50): alignment columns of Node::recover that are not
 * '|' = indel columns + diagonal columns whose bases differ; <0 if the ops do not span (read, template) */
int jtk_lc_nonmatch_columns(const uint8_t *ops, int n_ops, const uint8_t *read, int Lr, const uint8_t *tmpl, int Lt);
/* the same for n nodes in one call (concatenated ops / reads / templates with n+1 offsets each, node k on template
 * tmpl_idx[k]); out[k] = the key or -1 */
int jtk_lc_nonmatch_columns_batch(int n, const uint8_t *ops_concat, const uint64_t *ops_off, const uint8_t *read_concat,
                                  const uint64_t *read_off, const uint8_t *tmpl_concat, const uint64_t *tmpl_off,
                                  const uint32_t *tmpl_idx, int32_t *out);

/* K6 (kiley::gen_seq `Generate::gen`, call sites haplotyper/src/likelihood_gains.rs:20-21,276-279): one read
 * per source sequence sampled from the pair HMM `hmm45` (HMMParam order) on the host threads; kiley is absent, so this is our
 * own sampler of the same model, not kiley's generator stream.  Read k is drawn from source s = src_idx[k] =
 * src_concat[src_off[s] .. src_off[s+1]) with its own Xoshiro256** stream seeded with seed + k; it goes to out[k * cap ..]
 * (at most cap bases), its length to out_len[k]. */
int jtk_lc_gen_reads(const double *hmm45, int n, const uint8_t *src_concat, const uint64_t *src_off, const uint32_t *src_idx,
                     uint64_t seed, int cap, uint8_t *out, uint32_t *out_len);

/* The k-means + MCMC restarts of pseudo_mcmc::mcmc_clustering (haplotyper/src/local_clustering/pseudo_mcmc.rs:649-670:
 * `restarts` = 20 x (misc::kmeans + mcmc_with_filter) on one generator, keeping the last maximum) for n_chains independent
 * problems at once, one warp per chain.  Chain c: variants data_concat[data_off[c] ..] as n_rows[c] x
 * n_cols[c] doubles (row-major), n_clusters[c] clusters, size_to_lk (max_poisson_lk(x, cov, 1, k), x = 0..n_rows[c]:
 * n_rows[c]+1 doubles, concatenated), generator state rng_state[4c..4c+3] (Xoshiro256**, in/out).  Out: best assignment
 * (one byte per read, chains concatenated), its likelihood, and a status per chain (0 ok; non-zero = an assertion of the
 * reference failed: 1 k-means diverged, 2 invalid acceptance probability, 3 likelihood bookkeeping, 4 weights, 5 k < 2). */
int jtk_mcmc_restarts_batch(jtk_ctx *ctx, int n_chains, const double *data_concat, const uint64_t *data_off,
                            const uint32_t *n_rows, const uint32_t *n_cols, const uint32_t *n_clusters,
                            const double *size_to_lk_concat, int restarts, uint64_t *rng_state, uint8_t *out_asn,
                            double *out_lk, int *out_err);

/* Host twin of one chain of jtk_mcmc_restarts_batch (parity tests) and the size_to_lk table both use. */
int jtk_lc_mcmc_restarts_host(const double *data, int n, int D, int k, double cov, int restarts, uint64_t *state4,
                              uint8_t *out_asn, double *out_lk);
/* Host twin of the speculative schedule of mcmc_speculative_kernel (two clusters; up to `spec` = 4 or 8 proposals of one chain
 * evaluated side by side, committed up to the first acceptance): the results of jtk_lc_mcmc_restarts_host, bit for bit.
 * window = draws looked at per round (32 in the kernel; 3..31 makes rounds take the draw-by-draw path: tests).
 * out_stats (may be NULL) = { rounds, proposals, rounds that took the draw-by-draw path }. */
int jtk_lc_mcmc_restarts_spec_host(const double *data, int n, int D, double cov, int restarts, int window, int spec, uint64_t *state4,
                                   uint8_t *out_asn, double *out_lk, uint64_t *out_stats);
void jtk_lc_size_to_lk(int n, double cov, int k, double *out);
/* jtk_lc_clustering_variants_rng for many chunks with the restarts of every chunk on the GPU: chunk g has n_reads[g] x
 * n_probes[g] variants (dense, row-major, at var_off[g]), probe positions at ppos_off[g], template bytes
 * tmpl_off[g]..tmpl_off[g+1], its own config and generator state (states[4g..], in/out).  Out per chunk: assignments at
 * asn_off[g], log-posteriors at post_off[g] (n_reads[g] x post_stride), score, cluster number.  Results and generator
 * streams are those of the per-chunk call (pseudo_mcmc.rs:77-107,213-274). */
int jtk_lc_clustering_variants_batch(jtk_ctx *ctx, int n_chunks, const double *variants_concat, const uint64_t *var_off,
                                     const int32_t *n_reads, const int32_t *n_probes, const uint32_t *probe_pos_concat,
                                     const uint64_t *ppos_off, const uint8_t *tmpl_concat, const uint64_t *tmpl_off,
                                     const jtk_gains *gains, const jtk_clustering_config *cfgs, uint64_t *states,
                                     uint64_t *out_asn_concat, const uint64_t *asn_off, double *out_post_concat,
                                     const uint64_t *post_off, int post_stride, double *out_score, int32_t *out_k);

/* ---- layout pins of every struct that crosses the ABI (LP64): the Rust `#[repr(C)]` mirrors in jtk-gpu-sys/src/lib.rs
 * and the ctypes / numpy mirrors in jtk_b200/ carry the same numbers (tests/test_abi_exports.py) ---------------------- */
#ifdef __cplusplus
#define JTK_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define JTK_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
JTK_STATIC_ASSERT(sizeof(jtk_hmm_params) == 45 * 8 && offsetof(jtk_hmm_params, mat_emit) == 72 &&
                  offsetof(jtk_hmm_params, ins_emit) == 200, "jtk_hmm_params: 9 + 16 + 20 doubles, HMMParam order");
JTK_STATIC_ASSERT(sizeof(jtk_colstat) == 24 && offsetof(jtk_colstat, count) == 8 && offsetof(jtk_colstat, sc) == 12,
                  "jtk_colstat layout");
JTK_STATIC_ASSERT(sizeof(jtk_candidate) == 32 && offsetof(jtk_candidate, sum) == 16 && offsetof(jtk_candidate, lk) == 24,
                  "jtk_candidate layout");
JTK_STATIC_ASSERT(sizeof(jtk_gains) == 24 && offsetof(jtk_gains, gain) == 8 && offsetof(jtk_gains, prob) == 16, "jtk_gains layout");
JTK_STATIC_ASSERT(sizeof(jtk_clustering_config) == 24 && offsetof(jtk_clustering_config, coverage) == 8, "jtk_clustering_config layout");
JTK_STATIC_ASSERT(sizeof(jtk_polish_config) == 12, "jtk_polish_config layout");
/* sizeof() of the struct called `name` ("jtk_hmm_params", "jtk_colstat", ...) as this library was compiled; 0 = unknown name */
size_t jtk_abi_sizeof(const char *name);

/* Global alignment of `read` to `tmpl` (banded edit distance, band |i-j| <= radius + |Lr-Lt|; traceback prefers diagonal,
 * Del, Ins): the host aligner behind consensus::global_align (haplotyper/src/consensus/mod.rs:424-436, edlib global mode in
 * the reference) and the bootstrap guide of the likelihood call.  Writes one op per alignment column (<= Lt + Lr <= cap);
 * returns the number of columns or a negative JTK_E* code (band cannot connect the corners). */
int jtk_align_global(const uint8_t *tmpl, int Lt, const uint8_t *read, int Lr, int radius, uint8_t *out_ops, int cap);

/* in-band cell count C = sum_d w(d) of one pair (SURVEY.md section 8d work unit); <0 if ops are invalid */
int64_t jtk_band_cell_count(const uint8_t *ops, int n_ops, int Lt, int Lr, int radius);

#ifdef __cplusplus
}
#endif
#endif
