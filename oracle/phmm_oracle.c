/*
 * phmm_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See phmm_oracle.h.
 *
 * Three-state (Match / Ins / Del) global pair HMM, probability space, f64, one scale factor per
 * anti-diagonal.  Recurrences follow SURVEY.md Appendix A.2 (the specification implied by the
 * reference call sites; kiley's own source is absent -- PARITY UNPINNED):
 *
 *   F_M(i,j) = eM(t[j-1], q[i-1]) * toM(i-1,j-1)      toM(c) = mm*F_M(c) + im*F_I(c) + dm*F_D(c)
 *   F_I(i,j) = eI(q[i-2], q[i-1]) * toI(i-1,j)        toI(c) = mi*F_M(c) + ii*F_I(c) + di*F_D(c)
 *   F_D(i,j) =                      toD(i,  j-1)      toD(c) = md*F_M(c) + id*F_I(c) + dd*F_D(c)
 *   F_M(0,0) = 1,  lk = ln( F_M + F_I + F_D )(Lr,Lt)
 *
 * Band (Appendix A.3): cell (i,j) on anti-diagonal d=i+j is filled iff |i - centre(d)| <= radius,
 * centre(d) = read coordinate of the guide-path cell on d, or of the path cell on d-1 when a
 * diagonal step of the guide skips d.
 *
 * Modification table (Appendix A.4, call site pseudo_mcmc.rs:62-63, row map :168-177,447,500-511):
 * every global path consumes the first template base right of an edit exactly once, by a Match or a
 * Del move, so every entry is a cut between a forward column and a backward column:
 *   X(jf,jb) = sum_i toM(i,jf) * eM(t[jb],q[i]) * B_M(i+1,jb+1) + toD(i,jf) * B_D(i,jb+1)
 *   subst b at j    : the same cut with eM(b, q[i]) in place of eM(t[j], q[i])          (jf=jb=j)
 *   insert b before j: sum_i toM(i,j) * eM(b,q[i]) * B_M(i+1,j) + toD(i,j) * B_D(i,j)
 *   copy c at j     : X(j+c, j)      delete d at j: X(j, j+d)   (X(j,Lt) = sum_s F_s(Lr,j))
 * all evaluated with the band of the unedited template.
 */
#include "phmm_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline int base2(uint8_t c) {
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}

void orc_hmm_default(orc_hmm *h) {
    /* definitions/src/lib.rs:128-147 */
    h->mat_mat = 0.97; h->mat_ins = 0.01; h->mat_del = 0.01;
    h->ins_mat = 0.97; h->ins_ins = 0.01; h->ins_del = 0.01;
    h->del_mat = 0.97; h->del_ins = 0.01; h->del_del = 0.01;
    for (int r = 0; r < 4; r++)
        for (int q = 0; q < 4; q++) h->mat_emit[4 * r + q] = (r == q) ? 0.97 : 0.01;
    for (int k = 0; k < 20; k++) h->ins_emit[k] = 0.25;
}

/* ---------------------------------------------------------------- band geometry */
typedef struct {
    int Lt, Lr, r, nd, W;
    int32_t *c, *lo, *hi;
} band_t;

int orc_band_centres(const uint8_t *ops, int n_ops, int Lt, int Lr, int32_t *centre) {
    int i = 0, j = 0;
    centre[0] = 0;
    for (int k = 0; k < n_ops; k++) {
        switch (ops[k]) {
        case ORC_OP_MATCH: case ORC_OP_MISMATCH:
            if (i >= Lr || j >= Lt) return -1;
            centre[i + j + 1] = i; /* skipped anti-diagonal keeps the previous path cell's row */
            i++; j++;
            centre[i + j] = i;
            break;
        case ORC_OP_INS:
            if (i >= Lr) return -1;
            i++; centre[i + j] = i;
            break;
        case ORC_OP_DEL:
            if (j >= Lt) return -1;
            j++; centre[i + j] = i;
            break;
        default: return -1;
        }
    }
    return (i == Lr && j == Lt) ? 0 : -1;
}

static int band_init(band_t *b, const uint8_t *ops, int n_ops, int Lt, int Lr, int r) {
    b->Lt = Lt; b->Lr = Lr; b->r = r; b->nd = Lt + Lr + 1; b->W = 2 * r + 1;
    b->c = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)b->nd);
    b->lo = b->c + b->nd; b->hi = b->lo + b->nd;
    if (orc_band_centres(ops, n_ops, Lt, Lr, b->c) != 0) { free(b->c); b->c = NULL; return -1; }
    for (int d = 0; d < b->nd; d++) {
        int lo = b->c[d] - r, hi = b->c[d] + r;
        if (lo < 0) lo = 0;
        if (lo < d - Lt) lo = d - Lt;
        if (hi > Lr) hi = Lr;
        if (hi > d) hi = d;
        b->lo[d] = lo; b->hi[d] = hi;
    }
    return 0;
}
static void band_free(band_t *b) { free(b->c); b->c = NULL; }
static inline int inband(const band_t *b, int d, int i) {
    return d >= 0 && d < b->nd && i >= b->lo[d] && i <= b->hi[d];
}
static inline size_t bidx(const band_t *b, int d, int i) { return (size_t)d * b->W + (size_t)(i - b->c[d] + b->r); }

int64_t orc_cell_count(const uint8_t *ops, int n_ops, int Lt, int Lr, int radius) {
    band_t b;
    if (band_init(&b, ops, n_ops, Lt, Lr, radius)) return -1;
    int64_t n = 0;
    for (int d = 0; d < b.nd; d++)
        if (b.hi[d] >= b.lo[d]) n += b.hi[d] - b.lo[d] + 1;
    band_free(&b);
    return n;
}

/* ---------------------------------------------------------------- DP matrices */
typedef struct {
    double *m, *i, *d; /* nd * W each, scaled */
    double *ls;        /* ln scale per anti-diagonal: true = stored * exp(ls[d]) */
} mat_t;

static int mat_alloc(mat_t *x, const band_t *b) {
    size_t n = (size_t)b->nd * b->W;
    x->m = (double *)calloc(3 * n + b->nd, sizeof(double));
    if (!x->m) return -1;
    x->i = x->m + n; x->d = x->i + n; x->ls = x->d + n;
    return 0;
}
static void mat_free(mat_t *x) { free(x->m); x->m = NULL; }

typedef struct { const orc_hmm *h; const uint8_t *t, *q; } seqs_t; /* t,q already 2-bit codes */

static inline double eM(const orc_hmm *h, int tb, int qb) { return h->mat_emit[4 * tb + qb]; }
/* insertion of q[i] (0-based read index i): context = previous read base, 4 if none */
static inline double eI(const orc_hmm *h, const uint8_t *q, int i) {
    int ctx = i >= 1 ? q[i - 1] : 4;
    return h->ins_emit[4 * ctx + q[i]];
}

static void forward(const band_t *b, const seqs_t *s, mat_t *F) {
    const orc_hmm *h = s->h;
    for (int d = 0; d < b->nd; d++) {
        double sum = 0.0;
        double w2 = (d >= 2) ? exp(F->ls[d - 2] - F->ls[d - 1]) : 0.0;
        for (int i = b->lo[d]; i <= b->hi[d]; i++) {
            int j = d - i;
            double M = 0, I = 0, D = 0;
            if (d == 0) M = 1.0;
            else {
                if (i >= 1 && j >= 1 && inband(b, d - 2, i - 1)) {
                    size_t p = bidx(b, d - 2, i - 1);
                    double toM = h->mat_mat * F->m[p] + h->ins_mat * F->i[p] + h->del_mat * F->d[p];
                    M = eM(h, s->t[j - 1], s->q[i - 1]) * toM * w2;
                }
                if (i >= 1 && inband(b, d - 1, i - 1)) {
                    size_t p = bidx(b, d - 1, i - 1);
                    double toI = h->mat_ins * F->m[p] + h->ins_ins * F->i[p] + h->del_ins * F->d[p];
                    I = eI(h, s->q, i - 1) * toI;
                }
                if (j >= 1 && inband(b, d - 1, i)) {
                    size_t p = bidx(b, d - 1, i);
                    D = h->mat_del * F->m[p] + h->ins_del * F->i[p] + h->del_del * F->d[p];
                }
            }
            size_t p = bidx(b, d, i);
            F->m[p] = M; F->i[p] = I; F->d[p] = D;
            sum += M + I + D;
        }
        double base = d ? F->ls[d - 1] : 0.0;
        if (sum > 0.0) {
            double inv = 1.0 / sum;
            for (int i = b->lo[d]; i <= b->hi[d]; i++) {
                size_t p = bidx(b, d, i);
                F->m[p] *= inv; F->i[p] *= inv; F->d[p] *= inv;
            }
            F->ls[d] = base + log(sum);
        } else F->ls[d] = base;
    }
}

static double forward_lk(const band_t *b, const mat_t *F) {
    int d = b->nd - 1;
    if (!inband(b, d, b->Lr)) return -INFINITY;
    size_t p = bidx(b, d, b->Lr);
    double v = F->m[p] + F->i[p] + F->d[p];
    return v > 0 ? log(v) + F->ls[d] : -INFINITY;
}

/* B_s(i,j) = s_mat*gM + s_ins*gI + s_del*gD with
 * gM(i,j)=eM(t[j],q[i])*B_M(i+1,j+1), gI(i,j)=eI(q[i])*B_I(i+1,j), gD(i,j)=B_D(i,j+1) */
static void backward(const band_t *b, const seqs_t *s, mat_t *B) {
    const orc_hmm *h = s->h;
    int last = b->nd - 1;
    for (int d = last; d >= 0; d--) {
        double sum = 0.0;
        double w2 = (d + 2 <= last) ? exp(B->ls[d + 2] - B->ls[d + 1]) : 0.0;
        for (int i = b->lo[d]; i <= b->hi[d]; i++) {
            int j = d - i;
            double M, I, D;
            if (d == last) { M = I = D = 1.0; }
            else {
                double gM = 0, gI = 0, gD = 0;
                if (i < b->Lr && j < b->Lt && inband(b, d + 2, i + 1))
                    gM = eM(h, s->t[j], s->q[i]) * B->m[bidx(b, d + 2, i + 1)] * w2;
                if (i < b->Lr && inband(b, d + 1, i + 1)) gI = eI(h, s->q, i) * B->i[bidx(b, d + 1, i + 1)];
                if (j < b->Lt && inband(b, d + 1, i)) gD = B->d[bidx(b, d + 1, i)];
                M = h->mat_mat * gM + h->mat_ins * gI + h->mat_del * gD;
                I = h->ins_mat * gM + h->ins_ins * gI + h->ins_del * gD;
                D = h->del_mat * gM + h->del_ins * gI + h->del_del * gD;
            }
            size_t p = bidx(b, d, i);
            B->m[p] = M; B->i[p] = I; B->d[p] = D;
            sum += M + I + D;
        }
        double base = (d < last) ? B->ls[d + 1] : 0.0;
        if (sum > 0.0) {
            double inv = 1.0 / sum;
            for (int i = b->lo[d]; i <= b->hi[d]; i++) {
                size_t p = bidx(b, d, i);
                B->m[p] *= inv; B->i[p] *= inv; B->d[p] *= inv;
            }
            B->ls[d] = base + log(sum);
        } else B->ls[d] = base;
    }
}

static uint8_t *to_codes(const uint8_t *s, int n) {
    uint8_t *c = (uint8_t *)malloc((size_t)n + 1);
    for (int k = 0; k < n; k++) c[k] = (uint8_t)base2(s[k]);
    return c;
}

double orc_likelihood(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr,
                      const uint8_t *ops, int n_ops, int radius) {
    band_t b;
    if (band_init(&b, ops, n_ops, Lt, Lr, radius)) return NAN;
    mat_t F;
    if (mat_alloc(&F, &b)) { band_free(&b); return NAN; }
    uint8_t *tc = to_codes(t, Lt), *qc = to_codes(q, Lr);
    seqs_t s = { h, tc, qc };
    forward(&b, &s, &F);
    double lk = forward_lk(&b, &F);
    free(tc); free(qc); mat_free(&F); band_free(&b);
    return lk;
}

double orc_likelihood_backward(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr,
                               const uint8_t *ops, int n_ops, int radius) {
    band_t b;
    if (band_init(&b, ops, n_ops, Lt, Lr, radius)) return NAN;
    mat_t B;
    if (mat_alloc(&B, &b)) { band_free(&b); return NAN; }
    uint8_t *tc = to_codes(t, Lt), *qc = to_codes(q, Lr);
    seqs_t s = { h, tc, qc };
    backward(&b, &s, &B);
    double v = B.m[bidx(&b, 0, 0)]; /* start state is Match */
    double lk = v > 0 ? log(v) + B.ls[0] : -INFINITY;
    free(tc); free(qc); mat_free(&B); band_free(&b);
    return lk;
}

/* ---------------------------------------------------------------- K1 */
static int modtable_core(const orc_hmm *h, const uint8_t *tc, int Lt, const uint8_t *qc, int Lr,
                         const band_t *b, double *table, double *lk_out) {
    mat_t F, B;
    if (mat_alloc(&F, b)) return -2;
    if (mat_alloc(&B, b)) { mat_free(&F); return -2; }
    seqs_t s = { h, tc, qc };
    forward(b, &s, &F);
    backward(b, &s, &B);
    double lk = forward_lk(b, &F);
    *lk_out = lk;
    size_t nent = (size_t)(Lt + 1) * ORC_NUM_ROW;
    if (!table) { mat_free(&F); mat_free(&B); return 0; }
    if (!(lk > -INFINITY)) {
        for (size_t k = 0; k < nent; k++) table[k] = ORC_TABLE_NEG;
        mat_free(&F); mat_free(&B);
        return 0;
    }
    double *acc = (double *)calloc(nent, sizeof(double));
    /* EF[d][k+3] = exp(lsF[d] - lk + lsB[d+k]), k = -3..5 : scale of a forward cell on d times a backward cell on d+k */
    double *EF = (double *)calloc((size_t)b->nd * 9, sizeof(double));
    for (int d = 0; d < b->nd; d++)
        for (int k = -3; k <= 5; k++)
            if (d + k >= 0 && d + k < b->nd) EF[(size_t)d * 9 + k + 3] = exp(F.ls[d] - lk + B.ls[d + k]);
    /* rows of each column that are in band */
    int *cmin = (int *)malloc(sizeof(int) * 2 * (size_t)(Lt + 1)), *cmax = cmin + Lt + 1;
    for (int j = 0; j <= Lt; j++) { cmin[j] = Lr + 1; cmax[j] = -1; }
    for (int d = 0; d < b->nd; d++)
        for (int i = b->lo[d]; i <= b->hi[d]; i++) {
            int j = d - i;
            if (i < cmin[j]) cmin[j] = i;
            if (i > cmax[j]) cmax[j] = i;
        }
    for (int j = 0; j <= Lt; j++) {
        double *row = acc + (size_t)j * ORC_NUM_ROW;
        for (int i = cmin[j]; i <= cmax[j]; i++) {
            int d = i + j;
            if (!inband(b, d, i)) continue;
            size_t p = bidx(b, d, i);
            double toM = h->mat_mat * F.m[p] + h->ins_mat * F.i[p] + h->del_mat * F.d[p];
            double toD = h->mat_del * F.m[p] + h->ins_del * F.i[p] + h->del_del * F.d[p];
            const double *ef = EF + (size_t)d * 9 + 3; /* ef[k], k=-3..5 */
            /* substitution at j */
            if (j < Lt) {
                if (i < Lr && inband(b, d + 2, i + 1)) {
                    double U = toM * B.m[bidx(b, d + 2, i + 1)] * ef[2];
                    for (int x = 0; x < 4; x++) row[x] += U * eM(h, x, qc[i]);
                }
                if (inband(b, d + 1, i)) {
                    double V = toD * B.d[bidx(b, d + 1, i)] * ef[1];
                    for (int x = 0; x < 4; x++) row[x] += V;
                }
            }
            /* insertion before j */
            {
                if (i < Lr && inband(b, d + 1, i + 1)) {
                    double U = toM * B.m[bidx(b, d + 1, i + 1)] * ef[1];
                    for (int x = 0; x < 4; x++) row[4 + x] += U * eM(h, x, qc[i]);
                }
                double V = toD * B.d[p] * ef[0];
                for (int x = 0; x < 4; x++) row[4 + x] += V;
            }
            /* cuts X(jf=j, jb) for jb = j-3..j+3, jb != j */
            for (int e = -3; e <= 3; e++) {
                if (e == 0) continue;
                int jb = j + e;
                if (jb < 0 || jb > Lt) continue;
                double x = 0.0;
                if (jb == Lt) {
                    if (i == Lr) x = (F.m[p] + F.i[p] + F.d[p]) * exp(F.ls[d] - lk);
                } else {
                    int db = d + e;
                    if (i < Lr && inband(b, db + 2, i + 1))
                        x += toM * eM(h, tc[jb], qc[i]) * B.m[bidx(b, db + 2, i + 1)] * ef[e + 2];
                    if (inband(b, db + 1, i)) x += toD * B.d[bidx(b, db + 1, i)] * ef[e + 1];
                }
                if (e > 0) row[8 + ORC_COPY_SIZE + (e - 1)] += x;                      /* delete e bases at j   */
                else acc[(size_t)jb * ORC_NUM_ROW + 8 + (-e - 1)] += x;               /* copy -e bases at jb  */
            }
        }
    }
    free(EF); free(cmin);
    for (int j = 0; j <= Lt; j++) {
        for (int r = 0; r < ORC_NUM_ROW; r++) {
            int valid = 1;
            if (r < 4) valid = j < Lt;
            else if (r >= 8 && r < 8 + ORC_COPY_SIZE) valid = j + (r - 7) <= Lt;
            else if (r >= 8 + ORC_COPY_SIZE) valid = j + (r - 7 - ORC_COPY_SIZE) <= Lt;
            double a = acc[(size_t)j * ORC_NUM_ROW + r];
            table[(size_t)j * ORC_NUM_ROW + r] = (valid && a > 0.0) ? lk + log(a) : ORC_TABLE_NEG;
        }
    }
    free(acc); mat_free(&F); mat_free(&B);
    return 0;
}

int orc_modification_table(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr,
                           const uint8_t *ops, int n_ops, int radius, double *table, double *lk) {
    band_t b;
    if (band_init(&b, ops, n_ops, Lt, Lr, radius)) return -1;
    uint8_t *tc = to_codes(t, Lt), *qc = to_codes(q, Lr);
    int rc = modtable_core(h, tc, Lt, qc, Lr, &b, table, lk);
    free(tc); free(qc); band_free(&b);
    return rc;
}

int orc_apply_edit(const uint8_t *t, int Lt, int j, int row, uint8_t *out) {
    static const char B4[4] = { 'A', 'C', 'G', 'T' };
    if (j < 0 || j > Lt) return -1;
    if (row < 4) {
        if (j >= Lt) return -1;
        memcpy(out, t, (size_t)Lt); out[j] = (uint8_t)B4[row];
        return Lt;
    } else if (row < 8) {
        memcpy(out, t, (size_t)j); out[j] = (uint8_t)B4[row - 4];
        memcpy(out + j + 1, t + j, (size_t)(Lt - j));
        return Lt + 1;
    } else if (row < 8 + ORC_COPY_SIZE) {
        int c = row - 7;
        if (j + c > Lt) return -1;
        memcpy(out, t, (size_t)(j + c));
        memcpy(out + j + c, t + j, (size_t)(Lt - j));
        return Lt + c;
    } else if (row < ORC_NUM_ROW) {
        int d = row - 7 - ORC_COPY_SIZE;
        if (j + d > Lt) return -1;
        memcpy(out, t, (size_t)j);
        memcpy(out + j, t + j + d, (size_t)(Lt - j - d));
        return Lt - d;
    }
    return -1;
}

/* ---------------------------------------------------------------- K2 bootstrap */
int orc_edit_ops(const uint8_t *t, int Lt, const uint8_t *q, int Lr, int radius, uint8_t *ops_out) {
    int diff = Lr > Lt ? Lr - Lt : Lt - Lr;
    int R = radius + diff;
    const int INF = 1 << 29;
    size_t W = (size_t)Lt + 1;
    int *D = (int *)malloc(sizeof(int) * (size_t)(Lr + 1) * W);
    if (!D) return -1;
    for (int i = 0; i <= Lr; i++)
        for (int j = 0; j <= Lt; j++) {
            int v = INF;
            int off = i > j ? i - j : j - i;
            if (off <= R) {
                if (i == 0 && j == 0) v = 0;
                else {
                    if (i > 0 && j > 0) {
                        int c = D[(size_t)(i - 1) * W + j - 1] + (base2(q[i - 1]) != base2(t[j - 1]));
                        if (c < v) v = c;
                    }
                    if (j > 0) { int c = D[(size_t)i * W + j - 1] + 1; if (c < v) v = c; }
                    if (i > 0) { int c = D[(size_t)(i - 1) * W + j] + 1; if (c < v) v = c; }
                }
            }
            D[(size_t)i * W + j] = v;
        }
    if (D[(size_t)Lr * W + Lt] >= INF) { free(D); return -1; }
    int i = Lr, j = Lt, n = 0;
    while (i > 0 || j > 0) {
        int v = D[(size_t)i * W + j];
        if (i > 0 && j > 0) {
            int mis = base2(q[i - 1]) != base2(t[j - 1]);
            if (D[(size_t)(i - 1) * W + j - 1] + mis == v) {
                ops_out[n++] = mis ? ORC_OP_MISMATCH : ORC_OP_MATCH; i--; j--; continue;
            }
        }
        if (j > 0 && D[(size_t)i * W + j - 1] + 1 == v) { ops_out[n++] = ORC_OP_DEL; j--; continue; }
        ops_out[n++] = ORC_OP_INS; i--;
    }
    for (int a = 0, z = n - 1; a < z; a++, z--) { uint8_t x = ops_out[a]; ops_out[a] = ops_out[z]; ops_out[z] = x; }
    free(D);
    return n;
}

double orc_likelihood_bootstrap(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr, int radius) {
    uint8_t *ops = (uint8_t *)malloc((size_t)Lt + Lr + 1);
    int n = orc_edit_ops(t, Lt, q, Lr, radius, ops);
    double lk = n < 0 ? NAN : orc_likelihood(h, t, Lt, q, Lr, ops, n, radius);
    free(ops);
    return lk;
}

/* ---------------------------------------------------------------- batch K1 */
typedef struct {
    const orc_hmm *fwd, *rev; orc_pair *pairs; int n_pairs, radius;
    volatile int *next; int rc;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *job = (batch_job *)arg;
    for (;;) {
        int k = __sync_fetch_and_add(job->next, 1);
        if (k >= job->n_pairs) break;
        orc_pair *p = &job->pairs[k];
        int rc = orc_modification_table(p->strand ? job->fwd : job->rev, p->t, p->Lt, p->q, p->Lr,
                                        p->ops, p->n_ops, job->radius, p->table, &p->lk);
        if (rc) job->rc = rc;
    }
    return NULL;
}

int orc_modification_table_batch(const orc_hmm *fwd, const orc_hmm *rev, orc_pair *pairs, int n_pairs,
                                 int radius, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    volatile int next = 0;
    batch_job *jobs = (batch_job *)calloc((size_t)n_threads, sizeof(batch_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int k = 0; k < n_threads; k++) {
        jobs[k] = (batch_job){ fwd, rev, pairs, n_pairs, radius, &next, 0 };
        pthread_create(&th[k], NULL, batch_worker, &jobs[k]);
    }
    int rc = 0;
    for (int k = 0; k < n_threads; k++) { pthread_join(th[k], NULL); if (jobs[k].rc) rc = jobs[k].rc; }
    free(jobs); free(th);
    return rc;
}

/* ---------------------------------------------------------------- K3 polish */
#define POLISH_MIN_GAIN 0.1
#define POLISH_SUPPRESS 10
#define POLISH_DUP_SPAN 40
#define POLISH_TWIN_TOL 0.1
#define POLISH_MAX_ITER 20
#define POLISH_TIE_MARGIN 0.01

typedef struct { int j, row; } edit_t;

/* Rewrite one read's ops after the template edits (sorted, old coordinates). */
static int update_ops(const uint8_t *ops, int n, const edit_t *ed, int n_ed, uint8_t *out, int cap) {
    int m = 0, k = 0, tpos = 0, e = 0;
    int del_left = 0; /* template bases still to delete */
#define PUSH(x) do { if (m >= cap) return -1; out[m++] = (uint8_t)(x); } while (0)
    for (;;) {
        /* edits anchored at tpos fire before anything else at this template position */
        while (e < n_ed && ed[e].j == tpos && del_left == 0) {
            int row = ed[e].row; e++;
            if (row < 4) continue;
            if (row < 8 + ORC_COPY_SIZE) {
                int nb = row < 8 ? 1 : row - 7;
                for (int x = 0; x < nb; x++) {
                    if (k < n && ops[k] == ORC_OP_INS) { PUSH(ORC_OP_MATCH); k++; }
                    else PUSH(ORC_OP_DEL);
                }
            } else del_left = row - 7 - ORC_COPY_SIZE;
        }
        if (k >= n) break;
        uint8_t op = ops[k++];
        if (op == ORC_OP_INS) { PUSH(ORC_OP_INS); continue; }
        /* consumes template base tpos */
        if (del_left > 0) {
            if (op != ORC_OP_DEL) PUSH(ORC_OP_INS);
            del_left--;
        } else PUSH(op);
        tpos++;
    }
#undef PUSH
    return m;
}

int orc_polish_until_converge(const orc_hmm *fwd, const orc_hmm *rev, const uint8_t *draft, int Ld,
                              int n_reads, const uint8_t *const *reads, const int *read_len,
                              uint8_t **ops, int *n_ops, int ops_cap, const uint8_t *strands,
                              const orc_polish_cfg *cfg, uint8_t *out_cons, int cons_cap, int *n_iter) {
    int L = Ld;
    uint8_t *tmpl = (uint8_t *)malloc((size_t)cons_cap + 8);
    uint8_t *next = (uint8_t *)malloc((size_t)cons_cap + 8);
    uint8_t *obuf = (uint8_t *)malloc((size_t)ops_cap);
    if (Ld > cons_cap) return -1;
    memcpy(tmpl, draft, (size_t)Ld);
    int take = cfg->take_num < n_reads ? cfg->take_num : n_reads;
    int iters = 0, rc = 0;
    for (; iters < POLISH_MAX_ITER; iters++) {
        size_t nent = (size_t)(L + 1) * ORC_NUM_ROW;
        double *sum = (double *)calloc(nent, sizeof(double));
        double *tab = (double *)malloc(nent * sizeof(double));
        for (int r = 0; r < take && rc == 0; r++) {
            double lk;
            if (orc_modification_table(strands[r] ? fwd : rev, tmpl, L, reads[r], read_len[r], ops[r], n_ops[r],
                                       cfg->radius, tab, &lk)) { rc = -3; break; }
            for (size_t k = 0; k < nent; k++) sum[k] += tab[k] - lk;
        }
        free(tab);
        if (rc) { free(sum); break; }
        /* per-column best row and gain */
        edit_t *ed = (edit_t *)malloc(sizeof(edit_t) * (size_t)(L + 1));
        int *brow = (int *)malloc(sizeof(int) * (size_t)(L + 1));
        double *bgain = (double *)malloc(sizeof(double) * (size_t)(L + 1));
        for (int j = 0; j <= L; j++) {
            int best = -1; double bg = POLISH_MIN_GAIN;
            int own = j < L ? base2(tmpl[j]) : -1;
            if (j >= cfg->ignore_edge && j <= L - cfg->ignore_edge) {
                for (int row = 0; row < ORC_NUM_ROW; row++) {
                    if (row == own) continue;
                    if (row < 4 && j >= L - cfg->ignore_edge) continue;
                    if (row >= 8 && row < 8 + ORC_COPY_SIZE && j + (row - 7) > L - cfg->ignore_edge) continue;
                    if (row >= 8 + ORC_COPY_SIZE && j + (row - 7 - ORC_COPY_SIZE) > L - cfg->ignore_edge) continue;
                    double g = sum[(size_t)j * ORC_NUM_ROW + row];
                    if (g > bg) { bg = g; best = row; }
                }
            }
            brow[j] = best; bgain[j] = best >= 0 ? bg : 0.0;
        }
        /* an edit is taken iff (1) no candidate within POLISH_SUPPRESS columns gains more (ties within the margin: the
         * leftmost wins) and (2) it is not the tandem-repeat twin of an edit already taken (an indel of the same row with
         * a gain within POLISH_TWIN_TOL, at most POLISH_DUP_SPAN columns to the right) */
        int n_ed = 0;
        double *tgain = (double *)malloc(sizeof(double) * (size_t)(L + 1));
        for (int j = cfg->ignore_edge; j <= L - cfg->ignore_edge; j++) {
            int best = brow[j];
            if (best < 0) continue;
            double g = bgain[j];
            int take = 1;
            int lo = j - POLISH_SUPPRESS < cfg->ignore_edge ? cfg->ignore_edge : j - POLISH_SUPPRESS;
            for (int a = lo; a <= j + POLISH_SUPPRESS && a <= L - cfg->ignore_edge && take; a++) {
                if (a == j || brow[a] < 0) continue;
                if (a < j ? bgain[a] >= g - POLISH_TIE_MARGIN : bgain[a] > g + POLISH_TIE_MARGIN) take = 0;
            }
            for (int e = n_ed - 1; take && e >= 0 && j - ed[e].j <= POLISH_DUP_SPAN; e--)
                if (best >= 4 && ed[e].row == best && fabs(tgain[e] - g) <= POLISH_TWIN_TOL * (tgain[e] > g ? tgain[e] : g)) take = 0;
            if (take) { ed[n_ed].j = j; ed[n_ed].row = best; tgain[n_ed] = g; n_ed++; }
        }
        free(tgain);
        if (getenv("ORC_POLISH_DEBUG")) {
            fprintf(stderr, "iter %d L %d:", iters, L);
            for (int e = 0; e < n_ed; e++) fprintf(stderr, " (%d,%d,%.2f)", ed[e].j, ed[e].row, bgain[ed[e].j]);
            fprintf(stderr, "\n");
        }
        free(brow); free(bgain);
        free(sum);
        if (n_ed == 0) { free(ed); break; }
        /* apply to the template */
        int m = 0, src = 0;
        for (int e = 0; e < n_ed; e++) {
            int ej = ed[e].j, row = ed[e].row;
            while (src < ej) next[m++] = tmpl[src++];
            if (row < 4) { next[m++] = (uint8_t)"ACGT"[row]; src++; }
            else if (row < 8) next[m++] = (uint8_t)"ACGT"[row - 4];
            else if (row < 8 + ORC_COPY_SIZE) { for (int x = 0; x < row - 7; x++) next[m++] = tmpl[ej + x]; }
            else src += row - 7 - ORC_COPY_SIZE;
            if (m + 8 > cons_cap) { rc = -4; break; }
        }
        while (src < L && rc == 0) { if (m >= cons_cap) { rc = -4; break; } next[m++] = tmpl[src++]; }
        if (rc) { free(ed); break; }
        for (int r = 0; r < n_reads; r++) {
            int nn = update_ops(ops[r], n_ops[r], ed, n_ed, obuf, ops_cap);
            if (nn < 0) { rc = -5; break; }
            memcpy(ops[r], obuf, (size_t)nn); n_ops[r] = nn;
        }
        free(ed);
        if (rc) break;
        { uint8_t *x = tmpl; tmpl = next; next = x; }
        L = m;
    }
    if (rc == 0) { memcpy(out_cons, tmpl, (size_t)L); rc = L; }
    if (n_iter) *n_iter = iters;
    free(tmpl); free(next); free(obuf);
    return rc;
}

/* ---------------------------------------------------------------- K4 fit */
int orc_expected_counts(const orc_hmm *h, const uint8_t *t, int Lt, const uint8_t *q, int Lr,
                        const uint8_t *ops, int n_ops, int radius, double *acc) {
    band_t b;
    if (band_init(&b, ops, n_ops, Lt, Lr, radius)) return -1;
    mat_t F, B;
    if (mat_alloc(&F, &b)) { band_free(&b); return -2; }
    if (mat_alloc(&B, &b)) { mat_free(&F); band_free(&b); return -2; }
    uint8_t *tc = to_codes(t, Lt), *qc = to_codes(q, Lr);
    seqs_t s = { h, tc, qc };
    forward(&b, &s, &F);
    backward(&b, &s, &B);
    double lk = forward_lk(&b, &F);
    if (lk > -INFINITY) {
        const double tr[9] = { h->mat_mat, h->mat_ins, h->mat_del, h->ins_mat, h->ins_ins, h->ins_del,
                               h->del_mat, h->del_ins, h->del_del };
        for (int d = 0; d < b.nd; d++)
            for (int i = b.lo[d]; i <= b.hi[d]; i++) {
                int j = d - i;
                size_t p = bidx(&b, d, i);
                double f[3] = { F.m[p], F.i[p], F.d[p] };
                double lsF = F.ls[d] - lk;
                if (i < Lr && j < Lt && inband(&b, d + 2, i + 1)) {
                    double g = eM(h, tc[j], qc[i]) * B.m[bidx(&b, d + 2, i + 1)] * exp(lsF + B.ls[d + 2]);
                    double tot = 0;
                    for (int st = 0; st < 3; st++) { double x = f[st] * tr[3 * st + 0] * g; acc[3 * st + 0] += x; tot += x; }
                    acc[9 + 4 * tc[j] + qc[i]] += tot;
                }
                if (i < Lr && inband(&b, d + 1, i + 1)) {
                    double g = eI(h, qc, i) * B.i[bidx(&b, d + 1, i + 1)] * exp(lsF + B.ls[d + 1]);
                    double tot = 0;
                    for (int st = 0; st < 3; st++) { double x = f[st] * tr[3 * st + 1] * g; acc[3 * st + 1] += x; tot += x; }
                    int ctx = i >= 1 ? qc[i - 1] : 4;
                    acc[25 + 4 * ctx + qc[i]] += tot;
                }
                if (j < Lt && inband(&b, d + 1, i)) {
                    double g = B.d[bidx(&b, d + 1, i)] * exp(lsF + B.ls[d + 1]);
                    for (int st = 0; st < 3; st++) acc[3 * st + 2] += f[st] * tr[3 * st + 2] * g;
                }
            }
    }
    free(tc); free(qc); mat_free(&F); mat_free(&B); band_free(&b);
    return 0;
}

static void mstep(orc_hmm *h, const double *acc) {
    double *tr[9] = { &h->mat_mat, &h->mat_ins, &h->mat_del, &h->ins_mat, &h->ins_ins, &h->ins_del,
                      &h->del_mat, &h->del_ins, &h->del_del };
    for (int st = 0; st < 3; st++) {
        double tot = acc[3 * st] + acc[3 * st + 1] + acc[3 * st + 2];
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

int orc_fit(orc_hmm *fwd, orc_hmm *rev, const orc_pack *packs, int n_packs, int radius) {
    double accf[45] = { 0 }, accr[45] = { 0 };
    for (int p = 0; p < n_packs; p++)
        for (int r = 0; r < packs[p].n_reads; r++) {
            int st = packs[p].strands[r];
            int rc = orc_expected_counts(st ? fwd : rev, packs[p].t, packs[p].Lt, packs[p].reads[r], packs[p].read_len[r],
                                         packs[p].ops[r], packs[p].n_ops[r], radius, st ? accf : accr);
            if (rc) return rc;
        }
    mstep(fwd, accf);
    mstep(rev, accr);
    return 0;
}
