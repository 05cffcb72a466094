/*
 * phmm_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Double-precision restatement of the four kiley 0.3.0 pair-HMM entry points that
 * ban-m/jtk calls on its per-chunk hot path.  kiley is an un-vendored git dependency
 * (Cargo.lock:452-454, git rev 34ebbda0cb358335e22e20b054d357ea34d8326d); its source is
 * absent from /root/reference and no Rust toolchain exists here, so this file follows the
 * *specification implied by the reference call sites* (SURVEY.md Appendix A):
 *
 *   K1 modification_table_antidiagonal   haplotyper/src/local_clustering/pseudo_mcmc.rs:62-63
 *   K2 likelihood_antidiagonal_bootstrap haplotyper/src/likelihood_gains.rs:27-28,282-283,301-302
 *   K3 polish_until_converge_antidiagonal haplotyper/src/local_clustering/mod.rs:105-106,154-156
 *                                         haplotyper/src/model_tune.rs:141-143
 *   K4 fit_antidiagonal_par_multiple      haplotyper/src/model_tune.rs:145-151
 *   parameters (9 transitions, 16+20 emissions, declaration order)
 *                                         definitions/src/lib.rs:102-126
 *
 * PARITY UNPINNED: the reference holds no golden vector, known-answer test or fixture for
 * any of these calls and kiley itself cannot be run here.  The
 * oracle is pinned instead by self-checking invariants (tests/test_oracle_pins.py):
 * P1 table == likelihood of the explicitly edited template, P2 forward == backward,
 * P3 identity substitution == lk, P4 band -> infinity converges, P5 normalisation.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may call into this library.
 */
#ifndef PHMM_ORACLE_H
#define PHMM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUM_ROW 14   /* 4 subst + 4 ins + COPY_SIZE + DEL_SIZE  (pseudo_mcmc.rs:168-177,447) */
#define ORC_COPY_SIZE 3
#define ORC_DEL_SIZE 3
#define ORC_TABLE_NEG (-1.0e10) /* value of an impossible / out-of-range edit */

/* ops: one byte per alignment column (misc.rs:167-186, definitions/src/lib.rs:819-822) */
enum { ORC_OP_MATCH = 0, ORC_OP_MISMATCH = 1, ORC_OP_INS = 2 /* read only */, ORC_OP_DEL = 3 /* template only */ };

/* HMMParam in declaration order (definitions/src/lib.rs:102-126) */
typedef struct {
    double mat_mat, mat_ins, mat_del;
    double ins_mat, ins_ins, ins_del;
    double del_mat, del_ins, del_del;
    double mat_emit[16]; /* 4*ref + query */
    double ins_emit[20]; /* 4*prev_read_base + query, prev = 4 when the read has no previous base */
} orc_hmm;

void orc_hmm_default(orc_hmm *h); /* definitions/src/lib.rs:128-147 */

/* Band centre per anti-diagonal d = i + j.  centre has Lt+Lr+1 entries.
 * Returns 0, or -1 when ops do not span exactly (Lt, Lr). */
int orc_band_centres(const uint8_t *ops, int n_ops, int Lt, int Lr, int32_t *centre);

/* number of in-band cells C = sum_d w(d) (SURVEY 8d work unit) */
int64_t orc_cell_count(const uint8_t *ops, int n_ops, int Lt, int Lr, int radius);

/* ln P(read | template), banded around ops. */
double orc_likelihood(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr,
                      const uint8_t *ops, int n_ops, int radius);

/* same value obtained from the backward recursion (pin P2) */
double orc_likelihood_backward(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr,
                               const uint8_t *ops, int n_ops, int radius);

/* K1: table[(Lt+1)*14], absolute log-likelihoods; *lk = ln P(read|template).  0 on success. */
int orc_modification_table(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr,
                           const uint8_t *ops, int n_ops, int radius, double *table, double *lk);

/* apply edit (j,row) to a template; out must hold Lt+3 bytes; returns new length or -1 if invalid */
int orc_apply_edit(const uint8_t *t, int Lt, int j, int row, uint8_t *out);

/* banded global edit-distance alignment (the "bootstrap" path of K2). ops_out holds Lt+Lr bytes.
 * returns number of ops, or -1 if the band cannot connect the corners. */
int orc_edit_ops(const uint8_t *t, int Lt, const uint8_t *q, int Lr, int radius, uint8_t *ops_out);

/* K2 */
double orc_likelihood_bootstrap(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr, int radius);

/* batch K1 over pairs sharing nothing; n_threads OS threads over pairs (the rayon decomposition).
 * tables may be NULL (then only lks and column sums of (table-lk) are produced into colsum if not NULL) */
typedef struct {
    const uint8_t *t; int Lt;
    const uint8_t *q; int Lr;
    const uint8_t *ops; int n_ops;
    int strand; /* 1 = forward model, 0 = reverse model (pseudo_mcmc.rs:58-61) */
    double *table; /* (Lt+1)*14 or NULL */
    double lk;
} orc_pair;
int orc_modification_table_batch(const orc_hmm *fwd, const orc_hmm *rev, orc_pair *pairs, int n_pairs,
                                 int radius, int n_threads);

/* K3.  draft -> polished consensus.  ops[r] (length n_ops[r], capacity ops_cap) rewritten in place.
 * out_cons capacity cons_cap.  returns polished length or <0 on error. */
typedef struct { int radius; int take_num; int ignore_edge; } orc_polish_cfg;
int orc_polish_until_converge(const orc_hmm *fwd, const orc_hmm *rev, const uint8_t *draft, int Ld,
                              int n_reads, const uint8_t *const *reads, const int *read_len,
                              uint8_t **ops, int *n_ops, int ops_cap, const uint8_t *strands,
                              const orc_polish_cfg *cfg, uint8_t *out_cons, int cons_cap, int *n_iter);

/* K4: one Baum-Welch update of both strand models from packs of (template, reads, ops, strands). */
typedef struct {
    const uint8_t *t; int Lt; int n_reads;
    const uint8_t *const *reads; const int *read_len;
    const uint8_t *const *ops; const int *n_ops; const uint8_t *strands;
} orc_pack;
int orc_fit(orc_hmm *fwd, orc_hmm *rev, const orc_pack *packs, int n_packs, int radius);
/* expected counts of one pair (45 doubles: 9 transitions, 16 mat_emit, 20 ins_emit), added into acc */
int orc_expected_counts(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr,
                        const uint8_t *ops, int n_ops, int radius, double *acc45);

#ifdef __cplusplus
}
#endif
#endif
