// phmm_kernels.cu -- sm_100a kernels of the per-chunk pair-HMM path.
//
// Replaces the arithmetic behind kiley::hmm::PairHiddenMarkovModel::modification_table_antidiagonal
// (call site /root/reference/haplotyper/src/local_clustering/pseudo_mcmc.rs:62-63) and
// likelihood_antidiagonal_bootstrap (haplotyper/src/likelihood_gains.rs:27-28,282-283,301-302).
// Recurrences, band and table semantics are those of oracle/phmm_oracle.c.
//
// Layout: one warp per (read, template) pair, anti-diagonal wavefront, lanes own
// template columns.  Scaled fp32: state is multiplied by an exact power of two whenever the largest
// value on a diagonal drops below 2^-24; the backward pass mirrors the forward schedule so that every
// forward x backward product carries the same exponent and the table is a plain ratio of sums.
#include "phmm_dev.cuh"
#include <type_traits>
#include <algorithm>

namespace jtk {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerCta = 4;
constexpr int kRescaleEvery = 4;
// Scaled fp32 with one exact power-of-two scale per anti-diagonal.
//  * forward: the largest value on a diagonal is kept in [2^kScaleLow, 2^kScaleTarget] (checked every
//    kRescaleEvery steps, one rescale moves the exponent by at most kScaleStep).  Keeping the maximum HIGH leaves
//    ~2^180 of range BELOW it: after a long indel the cells that carry the eventual likelihood can be 1e-40 of
//    the locally largest (dead-end) cell.
//  * backward: mirrors the forward schedule, B(Lr,Lt) = 2^(kProductExp - exponent(fin)), so every forward x
//    backward product of one cell is (posterior mass) x 2^kProductExp: B = 2^kProductExp / F on the cells that
//    matter, inside the fp32 range on both sides as long as the dead-end advantage stays below ~2^180.
//  * a product that pairs rows on the two sides of a rescale gets its exact power-of-two correction.
constexpr int kScaleLow = 64;
constexpr int kScaleTarget = 100;
constexpr int kScaleStep = 64;
constexpr int kProductExp = 0;
constexpr int kEvBits = 4096;   // rescale-event map: one bit per block of four anti-diagonals (pairs up to ~16 k diagonals)
constexpr int kEvWords = kEvBits / 32 + 2;
// backward pass: per-warp shared-memory ring of forward rows, filled by bulk async copies
constexpr int kRingRows = 16;   // row slots (four groups of four rows)
// copies of the rows at the other end of the ring: as many as the cuts reach below (e = -1..-3: deletions) and above (copies)
constexpr int ring_below(int rows) { return rows == 14 ? 3 : 1; }
constexpr int ring_above(int rows) { return rows == 14 ? 3 : 0; }
constexpr int kRowShift = 4;    // ring / group index of anti-diagonal s is rho = s + kRowShift (four zero rows below s = 0)
constexpr int kRowsAbove = 8;   // rows past the last anti-diagonal that exist in the scratch (the first three read as zero)
#ifndef JTK_BWD_UNROLL
#define JTK_BWD_UNROLL 4
#endif
constexpr int kBwdUnroll = JTK_BWD_UNROLL; // steps of the fast backward block unrolled by the compiler
// resident CTAs per SM of bwdtable_kernel<2, ROWS> (register cap 65536 / (128 * this)): the 14-row kernel carries 44
// accumulators + 12 reused forward-row entries per lane and spills at 128 registers (152 without a cap -> 3 CTAs);
// the 9-row kernel fits 128 registers (4 CTAs).  Measured: profiles/README.md, v9.
#ifndef JTK_BWD_CTAS14
#define JTK_BWD_CTAS14 3
#endif
#ifndef JTK_BWD_CTAS9
#define JTK_BWD_CTAS9 4
#endif
// C = 4 (radius 31..62): a 107 KB ring per CTA of four warps, two CTAs per SM
// C = 8 (radius 63..126): a ring is 53 KB per WARP: CTAs of two warps, two CTAs per SM
constexpr int bwd_ctas_per_sm(int C, int rows) { return C == 2 ? (rows == 14 ? JTK_BWD_CTAS14 : JTK_BWD_CTAS9) : 2; }
constexpr int bwd_warps_per_cta(int C) { return C == 8 ? 2 : 4; } // kWarpsPerCta, except where four rings do not fit an SM
constexpr int kHalo = 4;        // (single-kernel forward_pass, STORE == 1) replicated slots on both sides of a row
// Forward rows of the modification table (v11): lane l owns the slots l + 32c; a row is C/2 pair planes, plane p holds per
// entry k the 16 bytes (toM[k+64p], toM[k+64p+32], toD[k+64p], toD[k+64p+32]), 32 entries + 3 replicated entries on both sides
// (k = -3..34, slot indices mod NSLOT).  A backward lane finds row s+e, slots sigma+e of both cells of a pair as ONE 128-bit
// load at lane stride 16 bytes (conflict-free) and a compile-time offset from one per-lane pointer.
constexpr int kPlaneHalo = 3;
constexpr int kPlane = 32 + 2 * kPlaneHalo; // 16-byte entries per pair plane; a row is C/2 planes = C * kPlane f2

// ---- small helpers --------------------------------------------------------------------------------------
typedef unsigned long long f2; // two packed fp32 (lo, hi) in one 64-bit register pair

__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi2(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// acc += a * b, accumulator tied in place (keeps loop-carried sums out of the register allocator's copy lists)
__device__ __forceinline__ void acc2(f2 &acc, f2 a, f2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 bc2(float v) { return mk2(v, v); }
// x = +0 for a finite x >= 0, as ONE instruction on the FMA pipe (FMUL2 / FMUL with RZ)
__device__ __forceinline__ void clear2(f2 &x) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(0ull)); }
__device__ __forceinline__ void clear1(float &x) { asm volatile("mul.rn.f32 %0, %0, 0f00000000;" : "+f"(x)); }

__device__ __forceinline__ float lds_f32(unsigned a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void lds_f32x4(unsigned a, f2 &x, f2 &y) {
    asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a));
}

__device__ __forceinline__ float pow2i(int k) { // exact 2^k, k in [-126, 127]
    return __uint_as_float((unsigned)(127 + k) << 23);
}
// band window of an anti-diagonal, NOT clamped to the matrix: cells outside the matrix evaluate to zero by themselves
// (sentinel codes have zero emissions), see the note above the forward pass
__device__ __forceinline__ int band_lo(int cen, int r, int s, int Lt) { (void)s; (void)Lt; return cen - r; }
__device__ __forceinline__ int band_hi(int cen, int r, int s, int Lr) { (void)s; (void)Lr; return cen + r; }

// four consecutive read-row codes Rb[i..i+3] as one register (byte 0 = row i); RbP = Rb - kCodePad is 4-aligned
__device__ __forceinline__ unsigned win_up(const uint8_t *__restrict__ RbP, int i) {
    const unsigned off = (unsigned)(i + kCodePad);
    const unsigned *wp = reinterpret_cast<const unsigned *>(RbP) + (off >> 2);
    return __funnelshift_r(wp[0], wp[1], (off & 3u) * 8u);
}
// four read-row codes Rb[a], Rb[a-1], Rb[a-2], Rb[a-3] (byte 0 = row a): the backward pass walks rows downwards
__device__ __forceinline__ unsigned win_down(const uint8_t *__restrict__ RbP, int a) {
    return __byte_perm(win_up(RbP, a - 3), 0u, 0x0123u);
}

// transition coefficients, paired the way the cell update consumes them
struct Coef {
    f2 fM, fI, fD;        // forward:  (toM, toD) = fM*M + fI*I + fD*D       = ((mm,md), (im,id), (dm,dd))
    float f_mi, f_ii, f_di; //         toI      = mi*M + ii*I + di*D
    f2 bM, bI, bD;        // backward: (B_M, B_D) = bM*gM + bI*gI + bD*gD   = ((mm,dm), (mi,di), (md,dd))
    float b_im, b_ii, b_id; //         B_I      = im*gM + ii*gI + id*gD
};
__device__ __forceinline__ Coef load_coef(const float *t, const unsigned long long *cp) {
    // t: mat_mat, mat_ins, mat_del, ins_mat, ...; cp: the six packed pairs, read as 64-bit values so that they stay
    // in aligned register pairs (no re-packing at every use)
    const float mi = t[1], im = t[3], ii = t[4], id = t[5], di = t[7];
    Coef c;
    c.fM = cp[0]; c.fI = cp[1]; c.fD = cp[2];
    c.f_mi = mi; c.f_ii = ii; c.f_di = di;
    c.bM = cp[3]; c.bI = cp[4]; c.bD = cp[5];
    c.b_im = im; c.b_ii = ii; c.b_id = id;
    return c;
}

// per-model tables in shared memory, aligned so that a table address is (base | code bits): one LOP3 per lookup
struct __align__(256) SmemLayout {
    float em[2][64];  // [tc*8 + qc]   byte offset tc*32 + qc*4
    float ei[2][64];  // [ctx*8 + qc]  byte offset = the read-row code byte (ctx<<5 | qc<<2)
    float emt[2][32]; // [qc*4 + b]    eM(ref b, query qc)
    float trans[2][12];
    unsigned long long cpair[2][6]; // (mm,md) (im,id) (dm,dd) | (mm,dm) (mi,di) (md,dd) as packed fp32 pairs
    float stage[kWarpsPerCta][kStageCols * kStageStride];
    float ftot[kWarpsPerCta][4];
    unsigned evw[kWarpsPerCta][kEvWords];          // blocks of four anti-diagonals in which the forward pass rescaled
    unsigned long long bar[kWarpsPerCta][4];       // mbarriers of the four ring groups
};

__device__ __forceinline__ void fill_tables(SmemLayout &sh, const float *__restrict__ models) {
    for (int k = threadIdx.x; k < 2 * 64; k += blockDim.x) {
        const int m = k >> 6, e = k & 63;
        sh.em[m][e] = models[m * kModelFloats + kOffEM + e];
        sh.ei[m][e] = models[m * kModelFloats + kOffEI + e];
        if (e < 32) sh.emt[m][e] = models[m * kModelFloats + kOffEMT + e];
        if (e < 12) sh.trans[m][e] = models[m * kModelFloats + e];
        if (e < 6) {
            const int lo_i[6] = { 0, 3, 6, 0, 1, 2 }, hi_i[6] = { 2, 5, 8, 6, 7, 8 };
            sh.cpair[m][e] = mk2(models[m * kModelFloats + lo_i[e]], models[m * kModelFloats + hi_i[e]]);
        }
    }
    __syncthreads();
}

// pair behind queue position k of the current wave (longest pairs first when the host gives an order)
__device__ __forceinline__ int pair_index(const KParams &p, int k) {
    return p.order ? (int)p.order[p.pair_lo + k] : p.pair_lo + k;
}

struct PairCtx { // warp-uniform view of one pair
    const uint8_t *Tb;  // Tb[j] = code of t[j-1]
    const uint8_t *RbP; // Rb - kCodePad
    const uint32_t *bw;
    int Lt, Lr, nd, r;
    unsigned sEM, sEI, sEC;  // shared-space addresses of this pair's model tables (sEC: paired substitution emissions)
};

// ------------------------------------------------------------------------------------------------
// Band bookkeeping.  A slot tracks x = i - (centre(s) - r), the offset of its cell inside the band window of
// anti-diagonal s; the cell is in band iff 0 <= x <= W = 2r.  The matrix edges need no clamps: codes outside
// [1, L] are sentinels whose emissions are zero, so every cell outside the matrix evaluates to zero by itself
// (forward: nothing flows in from i < 0 / j < 0; the D-state garbage in columns > Lt never flows back and meets
// g = 0 in every table sum; backward: the terminal value is injected at (Lr, Lt) only).
// Moving to the next anti-diagonal changes x only when the centre stays (guide bit 0): all x move by one and
// exactly one slot leaves the band -- a warp-uniform event, no vote needed.
// ------------------------------------------------------------------------------------------------
// Forward pass.  STORE = 1: write (toM, toD) of every cell to frow[s*RS + kHalo + slot] and the cumulative scale
// exponent to kf[s], and mark rescale events in evw.  STORE = 2 (fit): write (F_M, F_I, F_D) per slot.
// s_ftot[d] (d = 0..3) receives (F_M+F_I+F_D)(Lr, Lt-d) * 2^Ktot; s_ftot[0] is the value the likelihood is read from.
// ------------------------------------------------------------------------------------------------
template <int C> struct FwdState {
    int x[C], j[C];
    unsigned tcB[C], win[C];
    unsigned tcn[C]; // template code of the column this slot takes next (j + NSLOT), loaded one column-life ahead
    float toI[C], inD[C], inMa[C], inMb[C];
};

template <int C, int STORE, bool SPECIAL>
__device__ __forceinline__ void fwd_step(const PairCtx &pc, const Coef &a, FwdState<C> &st, const int s, const int W,
                                         int &K, f2 *__restrict__ &wrow, const int halo, int32_t *__restrict__ kf,
                                         volatile float *s_ftot, unsigned *evw) {
    constexpr int NSLOT = 32 * C;
    constexpr int RS = NSLOT + 2 * kHalo;
    const int lane = threadIdx.x & 31;
    f2 tMD[C];
    float Fm[C], Fi[C], Fd[C]; // forward states, kept only for the fit kernel (STORE == 2)
#pragma unroll
    for (int c = 0; c < C; c++) {
        const bool valid = (unsigned)st.x[c] <= (unsigned)W;
        const unsigned w = st.win[c];
        st.win[c] = w >> 8;
        float em = lds_f32(st.tcB[c] | (w & 0x1cu));
        float ei = lds_f32(pc.sEI | (w & 0xffu));
        em = valid ? em : 0.f;
        ei = valid ? ei : 0.f;
        float M = em * st.inMb[c];
        const float I = ei * st.toI[c];
        const float D = valid ? st.inD[c] : 0.f;
        if (SPECIAL && s == 0 && st.j[c] == 0) M = 1.f;
        tMD[c] = fma2(a.fD, bc2(D), fma2(a.fI, bc2(I), mul2(a.fM, bc2(M))));
        st.toI[c] = fmaf(a.f_di, D, fmaf(a.f_ii, I, a.f_mi * M));
        if (STORE == 2) { Fm[c] = M; Fi[c] = I; Fd[c] = D; }
        if (SPECIAL && s >= pc.nd - 4 && valid && s - st.j[c] == pc.Lr && st.j[c] >= pc.Lt - 3)
            s_ftot[pc.Lt - st.j[c]] = M + I + D;
    }
    if ((s & (kRescaleEvery - 1)) == kRescaleEvery - 1 && s < pc.nd - 8) {
        float v = lo2(tMD[0]);
#pragma unroll
        for (int c = 1; c < C; c++) v = fmaxf(v, lo2(tMD[c]));
        const unsigned m = __reduce_max_sync(kFull, __float_as_uint(v));
        const int e = (int)(m >> 23) - 127;
        if (m != 0u && e < kScaleLow) {
            const int k = min(kScaleTarget - e, kScaleStep);
            const float sc = pow2i(k);
#pragma unroll
            for (int c = 0; c < C; c++) { tMD[c] = mul2(tMD[c], bc2(sc)); st.toI[c] *= sc; st.inMa[c] *= sc; }
            if (STORE == 2) {
#pragma unroll
                for (int c = 0; c < C; c++) { Fm[c] *= sc; Fi[c] *= sc; Fd[c] *= sc; }
            }
            K += k;
            if (STORE == 1) { // the backward pass takes the exact-correction path around this row
                const unsigned b = (unsigned)(s >> 2) + 2u;
                if (lane == 0 && b < (unsigned)(kEvWords * 32)) evw[b >> 5] |= 1u << (b & 31u);
            }
        }
    }
    if (STORE == 2) { // fit: the three forward states of every cell, one float4 per slot, no halo
        if (lane == 0) kf[s] = K;
        float4 *srow = reinterpret_cast<float4 *>(wrow);
#pragma unroll
        for (int c = 0; c < C; c++) srow[c] = make_float4(Fm[c], Fi[c], Fd[c], 0.f);
        wrow = reinterpret_cast<f2 *>(srow + NSLOT);
    }
    if (STORE == 1) {
        if (lane == 0) kf[s] = K;
#pragma unroll
        for (int c = 0; c < C; c += 2)
            asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(wrow + c), "l"(tMD[c]), "l"(tMD[c + 1]) : "memory");
        if (halo != 0) { // the first / last three slots are replicated past the other end of the row
#pragma unroll
            for (int c = 0; c < C; c += 2)
                asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(wrow + halo + c), "l"(tMD[c]), "l"(tMD[c + 1]) : "memory");
        }
        wrow += RS;
    }
    // hand (toM, toD) to the right-hand neighbour column (slot+1, wrapping)
    const float rM = __shfl_sync(kFull, lo2(tMD[C - 1]), (lane + 31) & 31);
    const float rD = __shfl_sync(kFull, hi2(tMD[C - 1]), (lane + 31) & 31);
#pragma unroll
    for (int c = C - 1; c >= 1; c--) { st.inMb[c] = st.inMa[c]; st.inMa[c] = lo2(tMD[c - 1]); st.inD[c] = hi2(tMD[c - 1]); }
    st.inMb[0] = st.inMa[0]; st.inMa[0] = rM; st.inD[0] = rD;
}

template <int C, int STORE>
__device__ __forceinline__ void forward_pass(const PairCtx &pc, const Coef &a, float2 *__restrict__ frow,
                                             int32_t *__restrict__ kf, volatile float *s_ftot, int &Ktot, unsigned *evw) {
    constexpr int NSLOT = 32 * C;
    const int lane = threadIdx.x & 31;
    const int nd = pc.nd, W = 2 * pc.r;
    f2 *wrow = (STORE == 2) ? reinterpret_cast<f2 *>(reinterpret_cast<float4 *>(frow) + lane * C)
                            : reinterpret_cast<f2 *>(frow) + kHalo + lane * C; // this lane's slots in row 0
    const int halo = (lane * C < 3) ? NSLOT : ((lane * C + C > NSLOT - 3) ? -NSLOT : 0);
    FwdState<C> st;
#pragma unroll
    for (int c = 0; c < C; c++) {
        st.j[c] = lane * C + c;
        st.x[c] = pc.r - st.j[c]; // row i = -j on anti-diagonal 0, window starts at row -r
        st.tcB[c] = pc.sEM + ((unsigned)pc.Tb[st.j[c]] << 5);
        st.tcn[c] = pc.Tb[st.j[c] + NSLOT];
        st.win[c] = win_up(pc.RbP, -st.j[c]);
        st.toI[c] = st.inD[c] = st.inMa[c] = st.inMb[c] = 0.f;
    }
    if (lane < 4) s_ftot[lane] = 0.f;
    int K = 0;
    unsigned bword = pc.bw[0];
    // anti-diagonal s -> s+1: when the centre stays (guide bit 0) every cell moves one row up inside the window
    auto transition = [&](int s) {
        const int up = (int)(((bword >> (s & 31)) & 1u) ^ 1u);
        if (((s + 1) & 31) == 0) bword = pc.bw[(s + 1) >> 5];
#pragma unroll
        for (int c = 0; c < C; c++) st.x[c] += up;
    };
    // A slot whose cell has left the band at the top (x > W) moves on to column j + NSLOT.  It may be done up to two
    // steps late: the ring has NSLOT - (W+1) >= 3 spare slots, and a cell outside the band is masked anyway.
    auto retarget = [&](int s_next) {
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (st.x[c] > W) {
                st.j[c] += NSLOT;
                st.x[c] -= NSLOT;
                // no load on this path: the template code was fetched when the slot took its previous column, and the
                // read-row window of a retargeted slot is not needed before the next regular reload (the slot stays
                // outside the band for at least as many steps as that is away)
                st.tcB[c] = pc.sEM + (st.tcn[c] << 5);
                st.tcn[c] = pc.Tb[st.j[c] + NSLOT];
                (void)s_next;
            }
        }
    };
    auto reload = [&](int s) {
#pragma unroll
        for (int c = 0; c < C; c++) st.win[c] = win_up(pc.RbP, s - st.j[c]);
    };
    int s = 0;
    // prologue: the first four anti-diagonals (the start cell is injected at s = 0)
    for (; s < 4 && s < nd; ++s) {
        fwd_step<C, STORE, true>(pc, a, st, s, W, K, wrow, halo, kf, s_ftot, evw);
        if (s < nd - 1) { transition(s); retarget(s + 1); }
    }
    // main loop: four steps per read-row window
    for (; s + 3 < nd - 4; s += 4) {
        reload(s);
        fwd_step<C, STORE, false>(pc, a, st, s, W, K, wrow, halo, kf, s_ftot, evw); transition(s);
        fwd_step<C, STORE, false>(pc, a, st, s + 1, W, K, wrow, halo, kf, s_ftot, evw); transition(s + 1);
        retarget(s + 2);
        fwd_step<C, STORE, false>(pc, a, st, s + 2, W, K, wrow, halo, kf, s_ftot, evw); transition(s + 2);
        fwd_step<C, STORE, false>(pc, a, st, s + 3, W, K, wrow, halo, kf, s_ftot, evw); transition(s + 3);
        retarget(s + 4);
    }
    // epilogue: the last anti-diagonals also record the delete-to-end terms
    for (; s < nd; ++s) {
        if ((s & 3) == 0) reload(s);
        fwd_step<C, STORE, true>(pc, a, st, s, W, K, wrow, halo, kf, s_ftot, evw);
        if (s < nd - 1) { transition(s); retarget(s + 1); }
    }
    Ktot = K;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Backward pass fused with the modification-table reduction.  ROWS = 14 (all rows) or 9 (rows 0-7 and
// the one-base deletion, the only rows local_clustering reads: pseudo_mcmc.rs:447).
//
// The forward rows come back from HBM through a per-warp shared-memory ring filled by bulk async copies
// (cp.async.bulk, completion on an mbarrier), one group of four rows per copy, issued one block of four
// anti-diagonals ahead of their first use: the inner loop sees shared-memory loads with immediate offsets only.
// Ring geometry: row index rho = s + kRowShift; 16 row slots + 3 margin slots on both sides that hold copies of the
// rows at the other end, so that rows rho-3 .. rho+3 are always contiguous around slot (rho & 15) + 3.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one lane of the (converged) warp, chosen by the hardware: tells ptxas that what follows runs in ONE thread, so the bulk
// copy's uniform-register operands need no "for every distinct value among the active lanes" loop
__device__ __forceinline__ bool elect_one() {
    unsigned p;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(p));
    return p != 0u;
}


// Exact power-of-two corrections of one step of a block of four anti-diagonals (block q = rows 4q-4 .. 4q-1) next to forward
// rescales.  The forward pass rescales on the LAST row of a block, so inside the rows 4q-7 .. 4q+2 a step can touch there are
// two boundaries: 4q-6 | 4q-5 (amount kA, end of forward block q-2) and 4q-2 | 4q-1 (amount kB, block q-1).  Which of the
// three zones row s = 4q-1-k and row s+e fall into is known at compile time: ce[e+3] = 2^(K(s) - K(s+e)) is one of
// 1, fB = 2^kB, fBi = 2^-kB, fA = 2^kA.
__device__ __forceinline__ void block_corrections(const int k, const float fA, const float fB, const float fBi, float (&ce)[7]) {
    const int zs = (k == 0) ? 2 : 1;
#pragma unroll
    for (int e = -3; e <= 3; e++) {
        const int r = -1 - k + e; // row s+e relative to 4q
        const int zt = r <= -6 ? 0 : (r <= -2 ? 1 : 2);
        ce[e + 3] = zs == zt ? 1.f : (zs == 2 ? fB : (zt == 2 ? fBi : fA));
    }
}
__device__ __forceinline__ float pow2c(int k) { return pow2i(max(-126, min(126, k))); }

// ------------------------------------------------------------------------------------------------
// Backward / table state (v11).  Lane l owns the slots sigma = l + 32c (c = 0..C-1); the cells of slots l + 64p and
// l + 64p + 32 are packed into ONE fp32 pair (lo = the lower slot), so every arithmetic instruction of the step is a packed
// one (FFMA2 / FMUL2) over two cells, the nine transition coefficients are plain scalars (the packed instructions take a
// broadcast operand) and nothing is ever moved between the halves of a pair.  A forward row is laid out to match: per pair
// p one plane of 16-byte entries, entry k = (toM[k+64p], toM[k+64p+32], toD[k+64p], toD[k+64p+32]), three replicated
// entries past both ends -- so row s+e, slots sigma+e of BOTH cells is one conflict-free 128-bit load at a compile-time
// offset from one per-lane pointer, for every e = -3..3.
// ------------------------------------------------------------------------------------------------
template <int C> struct BwdState {
    static constexpr int P = C / 2;
    int x[C], j[C];
    unsigned tcB[C];             // shared address of the eM row of the cell's column
    unsigned tcn[C];             // template code of the column this slot takes next (j - NSLOT), loaded one column-life ahead
    const unsigned char *rbp[C]; // staged code byte of read row s - j + 1 for the first step of the current block
    f2 msk[P];                   // 1.0 where the cell is inside the band
    f2 BI[P], BMo[P], inD[P], inMa[P], inMb[P];
    f2 S01[C], S23[C], N01[C], N23[C]; // substitution / insertion sums over the four bases, per CELL: (b0, b1), (b2, b3)
    f2 Vs[P], Vn[P];
    f2 Xp[P][3], Xm[P][3];       // copy / deletion cuts (sum toM*gM + toD*gD)
};
struct BCoef { float mm, mi, md, im, ii, id, dm, di, dd; };
__device__ __forceinline__ BCoef load_bcoef(const float *t) {
    BCoef c;
    c.mm = t[0]; c.mi = t[1]; c.md = t[2]; c.im = t[3]; c.ii = t[4]; c.id = t[5]; c.dm = t[6]; c.di = t[7]; c.dd = t[8];
    return c;
}
__device__ __forceinline__ float half2f(f2 v, int h) { return h ? hi2(v) : lo2(v); }
// zero one half of a pair with a multiply (the sums are finite and >= 0): one FMA-pipe instruction, no constant moves
__device__ __forceinline__ void clear_half(f2 &v, int h) {
    float lo = lo2(v), hi = hi2(v);
    if (h) clear1(hi); else clear1(lo);
    v = mk2(lo, hi);
}

template <int C> __device__ __forceinline__ void bwd_masks(BwdState<C> &st, const int W) {
#pragma unroll
    for (int p = 0; p < C / 2; p++)
        st.msk[p] = mk2((unsigned)st.x[2 * p] <= (unsigned)W ? 1.f : 0.f, (unsigned)st.x[2 * p + 1] <= (unsigned)W ? 1.f : 0.f);
}

template <int C> __device__ __forceinline__ void bwd_init(BwdState<C> &st, const PairCtx &pc, const unsigned char *rb0) {
    constexpr int NSLOT = 32 * C;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int sigma = lane + 32 * c;
        const int d = (pc.Lt - sigma) & (NSLOT - 1);
        st.j[c] = pc.Lt - d;   // largest column <= Lt owned by this slot
        st.x[c] = d + pc.r;    // row i = Lr + d on the last anti-diagonal, window starts at row Lr - r
        st.tcB[c] = pc.sEM + ((unsigned)pc.Tb[st.j[c] + 1] << 5); // code of t[j]
        st.tcn[c] = pc.Tb[st.j[c] + 1 - NSLOT];
        st.rbp[c] = rb0 + (pc.nd - st.j[c]); // row s - j + 1 at s = nd - 1
    }
#pragma unroll
    for (int p = 0; p < C / 2; p++) {
        st.BI[p] = st.BMo[p] = st.inD[p] = st.inMa[p] = st.inMb[p] = 0ull;
        st.Vs[p] = st.Vn[p] = 0ull;
#pragma unroll
        for (int e = 0; e < 3; e++) { st.Xp[p][e] = 0ull; st.Xm[p][e] = 0ull; }
    }
#pragma unroll
    for (int c = 0; c < C; c++) st.S01[c] = st.S23[c] = st.N01[c] = st.N23[c] = 0ull;
    bwd_masks<C>(st, 2 * pc.r);
}

// One anti-diagonal.  rq = entry `lane` of pair plane 0 of forward row s (16-byte entries, RQ per row): row s+e, slots
// sigma+e of pair p is rq[e * RQ + p * kPlane + e].
// CORR: a rescale lies within rows s-2 .. s+3 (products that pair two rows get their exact power-of-two correction);
// FIRST: s = nd-1, the terminal cell is injected (and nothing is carried yet).
template <int C, int ROWS, bool CORR, bool FIRST>
__device__ __forceinline__ void bwd_step(const PairCtx &pc, const BCoef &a, BwdState<C> &st, const ulonglong2 *rq, const float *ce,
                                         const float boff, f2 (&bM)[C / 2], f2 (&bD)[C / 2], const int kk) {
    constexpr int P = C / 2, RQ = P * kPlane;
    constexpr int NXM = (ROWS == 14) ? 3 : 1;
    constexpr int NXP = (ROWS == 14) ? 3 : 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        const unsigned w0 = st.rbp[2 * p][-kk], w1 = st.rbp[2 * p + 1][-kk];
        const f2 em = mk2(lds_f32(st.tcB[2 * p] | (w0 & 0x1cu)), lds_f32(st.tcB[2 * p + 1] | (w1 & 0x1cu)));
        const f2 ei = mk2(lds_f32(pc.sEI + w0), lds_f32(pc.sEI + w1));
        // eM(ref b, read base), b = 0..3, of each cell: the substitution / insertion sums are packed over b, not over the cells
        f2 ea01, ea23, eb01, eb23;
        lds_f32x4(pc.sEC | ((w0 & 0x1cu) << 2), ea01, ea23);
        lds_f32x4(pc.sEC | ((w1 & 0x1cu) << 2), eb01, eb23);
        // a cell outside the band contributes nothing: its in-sums are masked (its forward values are zero already)
        const f2 m = st.msk[p];
        const f2 gD = mul2(st.inD[p], m);
        const f2 gM = mul2(mul2(em, m), st.inMb[p]);
        const f2 gI = mul2(mul2(ei, m), st.BI[p]);
        f2 m_ = fma2(bc2(a.md), gD, fma2(bc2(a.mi), gI, mul2(bc2(a.mm), gM)));
        f2 d_ = fma2(bc2(a.dd), gD, fma2(bc2(a.di), gI, mul2(bc2(a.dm), gM)));
        f2 i_ = fma2(bc2(a.id), gD, fma2(bc2(a.ii), gI, mul2(bc2(a.im), gM)));
        if (FIRST) { // B(Lr, Lt) = boff, everything else on the last anti-diagonal is outside the matrix
            m_ = mk2(st.j[2 * p] == pc.Lt ? boff : 0.f, st.j[2 * p + 1] == pc.Lt ? boff : 0.f);
            d_ = m_; i_ = m_;
        }
        // ---- table reduction for the cells (i, j) of the pair ----
        const ulonglong2 F0 = rq[p * kPlane];
        const f2 U = mul2(F0.x, st.inMb[p]);
        acc2(st.S01[2 * p], bc2(lo2(U)), ea01); acc2(st.S23[2 * p], bc2(lo2(U)), ea23);
        acc2(st.S01[2 * p + 1], bc2(hi2(U)), eb01); acc2(st.S23[2 * p + 1], bc2(hi2(U)), eb23);
        acc2(st.Vs[p], F0.y, st.inD[p]);
        const f2 U2 = mul2(F0.x, st.BMo[p]);
        acc2(st.N01[2 * p], bc2(lo2(U2)), ea01); acc2(st.N23[2 * p], bc2(lo2(U2)), ea23);
        acc2(st.N01[2 * p + 1], bc2(hi2(U2)), eb01); acc2(st.N23[2 * p + 1], bc2(hi2(U2)), eb23);
        acc2(st.Vn[p], F0.y, d_);
        // cuts pairing this column's backward in-sums with forward columns j-1..j-3 (deletions) and j+1..j+3
        // (copies).  g is zero unless the cell is in band, and then slot+e of row s+e holds column j+e.
#pragma unroll
        for (int e = 1; e <= NXM; e++) {
            const ulonglong2 Fe = rq[-e * RQ + p * kPlane - e];
            if (CORR) { acc2(st.Xm[p][e - 1], mul2(Fe.x, gM), bc2(ce[3 - e])); acc2(st.Xm[p][e - 1], mul2(Fe.y, gD), bc2(ce[3 - e])); }
            else { acc2(st.Xm[p][e - 1], Fe.x, gM); acc2(st.Xm[p][e - 1], Fe.y, gD); }
        }
#pragma unroll
        for (int e = 1; e <= NXP; e++) {
            const ulonglong2 Fe = rq[e * RQ + p * kPlane + e];
            if (CORR) { acc2(st.Xp[p][e - 1], mul2(Fe.x, gM), bc2(ce[3 + e])); acc2(st.Xp[p][e - 1], mul2(Fe.y, gD), bc2(ce[3 + e])); }
            else { acc2(st.Xp[p][e - 1], Fe.x, gM); acc2(st.Xp[p][e - 1], Fe.y, gD); }
        }
        bM[p] = m_; bD[p] = d_;
        st.BI[p] = i_; st.BMo[p] = m_;
    }
}

// hand (B_M, B_D) to the left-hand neighbour column: slot sigma receives from slot sigma+1 = the same cell of lane+1; lane 31
// receives from the NEXT cell of lane 0 (slot 32c + 31 -> 32(c+1), wrapping)
template <int C> __device__ __forceinline__ void bwd_hand_off(BwdState<C> &st, const f2 (&bM)[C / 2], const f2 (&bD)[C / 2]) {
    const int lane = threadIdx.x & 31;
    const bool seam = lane == 31;
    float rM[C], rD[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        rM[c] = __shfl_sync(kFull, half2f(bM[c / 2], c & 1), (lane + 1) & 31);
        rD[c] = __shfl_sync(kFull, half2f(bD[c / 2], c & 1), (lane + 1) & 31);
    }
#pragma unroll
    for (int p = 0; p < C / 2; p++) {
        st.inMb[p] = st.inMa[p];
        st.inMa[p] = mk2(seam ? rM[(2 * p + 1) % C] : rM[2 * p], seam ? rM[(2 * p + 2) % C] : rM[2 * p + 1]);
        st.inD[p] = mk2(seam ? rD[(2 * p + 1) % C] : rD[2 * p], seam ? rD[(2 * p + 2) % C] : rD[2 * p + 1]);
    }
}
template <int C> __device__ __forceinline__ void bwd_scale(BwdState<C> &st, const float sc) {
#pragma unroll
    for (int p = 0; p < C / 2; p++) {
        st.BI[p] = mul2(st.BI[p], bc2(sc)); st.BMo[p] = mul2(st.BMo[p], bc2(sc)); st.inD[p] = mul2(st.inD[p], bc2(sc));
        st.inMa[p] = mul2(st.inMa[p], bc2(sc)); st.inMb[p] = mul2(st.inMb[p], bc2(sc));
    }
}
// anti-diagonal s -> s-1 when the centre stays (guide bit 0, dec = 1): every cell moves one row down inside the window
template <int C> __device__ __forceinline__ void bwd_band_down(BwdState<C> &st, const int dec, const int W) {
#pragma unroll
    for (int c = 0; c < C; c++) st.x[c] -= dec;
    bwd_masks<C>(st, W);
}
// a finished column leaves its 16 raw sums in the pair's scratch (64 B, one lane); finalize_kernel turns them into the 14
// log-ratios of the table.  raw[j] = { S0 S1 S2 S3 | N0 N1 N2 N3 | Vs Vn Xp0 Xp1 | Xp2 Xm0 Xm1 Xm2 }
template <int C> __device__ __forceinline__ void bwd_flush_col(const BwdState<C> &st, const int c, float4 *raw) {
    const int p = c / 2, h = c & 1;
    f2 *sg = reinterpret_cast<f2 *>(raw);
    asm volatile("st.global.cg.v2.b64 [%0], {%1, %2};" ::"l"(sg + 0), "l"(st.S01[c]), "l"(st.S23[c]) : "memory");
    asm volatile("st.global.cg.v2.b64 [%0], {%1, %2};" ::"l"(sg + 2), "l"(st.N01[c]), "l"(st.N23[c]) : "memory");
    __stcg(raw + 2, make_float4(half2f(st.Vs[p], h), half2f(st.Vn[p], h), half2f(st.Xp[p][0], h), half2f(st.Xp[p][1], h)));
    __stcg(raw + 3, make_float4(half2f(st.Xp[p][2], h), half2f(st.Xm[p][0], h), half2f(st.Xm[p][1], h), half2f(st.Xm[p][2], h)));
}
// Retirement, once per block of four anti-diagonals: a slot whose cell left the band at the bottom (x < 0) has seen its last
// in-band cell.  It is masked like any cell outside the band until then; the ring has NSLOT - (W+1) >= 3 spare slots, and the
// column j - NSLOT that the slot takes over cannot enter the band before the first step of the next block.  The (on average
// two) slots that retire in one block are neighbouring lanes with the same c, so one pass through the flush code serves both.
// RAW(j) = address of the column's raw sums.
template <int C, int ROWS, typename RAW>
__device__ __forceinline__ void bwd_retire(BwdState<C> &st, const PairCtx &pc, RAW raw_of) {
    constexpr int NSLOT = 32 * C;
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (st.x[c] < 0) {
            const int p = c / 2, h = c & 1;
            if (st.j[c] >= 0) bwd_flush_col<C>(st, c, raw_of(st.j[c]));
            st.j[c] -= NSLOT;
            st.x[c] += NSLOT;
            st.rbp[c] += NSLOT;
            st.tcB[c] = pc.sEM + (st.tcn[c] << 5); // fetched one column-life ago
            st.tcn[c] = pc.Tb[st.j[c] + 1 - NSLOT];
            clear2(st.S01[c]); clear2(st.S23[c]); clear2(st.N01[c]); clear2(st.N23[c]);
            clear_half(st.Vs[p], h); clear_half(st.Vn[p], h);
#pragma unroll
            for (int e = 0; e < 3; e++) { // the 9-row kernel only accumulates Xm[0]: the others stay compile-time zeros
                if (ROWS == 14) clear_half(st.Xp[p][e], h);
                if (ROWS == 14 || e == 0) clear_half(st.Xm[p][e], h);
            }
        }
    }
    bwd_masks<C>(st, 2 * pc.r);
}

// ------------------------------------------------------------------------------------------------
// Backward pass fused with the modification-table reduction, forward rows in HBM (rows variant).  ROWS = 14 (all rows) or
// 9 (rows 0-7 and the one-base deletion, the only rows local_clustering reads: pseudo_mcmc.rs:447).
//
// The forward rows come back from HBM through a per-warp shared-memory ring filled by bulk async copies
// (cp.async.bulk, completion on an mbarrier), one group of four rows per copy, issued one block of four
// anti-diagonals ahead of their first use: the inner loop sees shared-memory loads with immediate offsets only.
// Ring geometry: row index rho = s + kRowShift; 16 row slots + 3 margin slots on both sides that hold copies of the
// rows at the other end, so that rows rho-3 .. rho+3 are always contiguous around slot (rho & 15) + 3.
// ------------------------------------------------------------------------------------------------
template <int C, int ROWS>
__device__ __forceinline__ void backward_pass(const PairCtx &pc, const BCoef &a, const float2 *__restrict__ frow,
                                              const int32_t *__restrict__ kb, float4 *__restrict__ raw,
                                              volatile float *s_ftot, f2 *ring, const unsigned bars, unsigned &phase,
                                              const unsigned char *rb0) {
    constexpr int P = C / 2;
    constexpr int RS = C * kPlane;    // f2 per forward row
    constexpr int RQ = P * kPlane;    // 16-byte entries per forward row
    constexpr unsigned RSB = RS * 8u; // bytes per forward row
    constexpr unsigned MB = ring_below(ROWS), MA = ring_above(ROWS);
    const int lane = threadIdx.x & 31;
    const int nd = pc.nd, W = 2 * pc.r;
    // B(Lr,Lt) = boff puts sum_cells F*B = fin * boff at about 2^kProductExp
    const float fin_raw = s_ftot[0];
    const int e_fin = (int)(__float_as_uint(fin_raw) >> 23) - 127;
    const float boff = fin_raw > 0.f ? pow2i(max(-120, min(120, kProductExp - e_fin))) : 1.f;

    // ---- ring of forward rows ------------------------------------------------------------------------
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    const ulonglong2 *ring_q = reinterpret_cast<const ulonglong2 *>(ring);
    const float2 *grow0 = frow - (ptrdiff_t)kRowShift * RS; // global row rho = 0
    auto issue_group = [&](int q) { // rows rho = 4q .. 4q+3
        if (elect_one()) {
            const unsigned g = (unsigned)q & 3u;
            const unsigned bar = bars + 8u * g;
            const float2 *src = grow0 + (size_t)(4 * q) * RS;
            mbar_expect_tx(bar, (4u + (g == 3u ? MB : 0u) + (g == 0u ? MA : 0u)) * RSB);
            bulk_g2s(ring_s + (4u * g + MB) * RSB, src, 4u * RSB, bar);
            if (g == 3u) bulk_g2s(ring_s, src + (4 - (int)MB) * RS, MB * RSB, bar);                    // rows 16-MB..15 below slot 0
            if (MA > 0u && g == 0u) bulk_g2s(ring_s + (kRingRows + MB) * RSB, src, MA * RSB, bar);    // rows 0..MA-1 above slot 15
        }
    };
    auto wait_group = [&](int q) {
        const unsigned g = (unsigned)q & 3u;
        mbar_wait(bars + 8u * g, (phase >> g) & 1u);
        phase ^= 1u << g;
    };
    const int q_top = (nd + 3) >> 2; // block of the last anti-diagonal (rho = nd + 3)
    // the forward rows were written through the generic proxy: order them before the async-proxy reads
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncwarp();
    issue_group(q_top + 1);
    issue_group(q_top);
    issue_group(q_top - 1);

    BwdState<C> st;
    bwd_init<C>(st, pc, rb0);
    auto raw_of = [&](int j) -> float4 * { return raw + (size_t)j * 4; };

    // generic step: any s, exact corrections, ring slot computed from s
    auto slow_step = [&](int s) {
        const ulonglong2 *rq = ring_q + (size_t)(((s + kRowShift) & (kRingRows - 1)) + MB) * RQ + kPlaneHalo + lane;
        // cumulative scale exponent of forward row t: rescales happen on the last row of a block of four, kb[q] = exponent
        // after block q
        auto kfat = [&](int t) -> int { return kb[(t - 3) >> 2]; };
        const int kcur = kfat(s);
        const int kstep = kcur - kfat(s - 1); // exponent the forward pass added at step s (mirrored below)
        float ce[7];
#pragma unroll
        for (int e = -3; e <= 3; e++) ce[e + 3] = pow2i(max(-126, min(126, kcur - kfat(s + e))));
        f2 bM[P], bD[P];
        if (s == nd - 1) bwd_step<C, ROWS, true, true>(pc, a, st, rq, ce, boff, bM, bD, 0);
        else bwd_step<C, ROWS, true, false>(pc, a, st, rq, ce, boff, bM, bD, 0);
#pragma unroll
        for (int c = 0; c < C; c++) st.rbp[c] -= 1;
        bwd_hand_off<C>(st, bM, bD);
        if (s > 0) {
            if (kstep != 0) bwd_scale<C>(st, pow2i(kstep)); // mirror of the forward rescale at step s
            const int dec = (int)(((pc.bw[(s - 1) >> 5] >> ((s - 1) & 31)) & 1u) ^ 1u);
            bwd_band_down<C>(st, dec, W);
        }
    };

    // block q (rows 4q-4 .. 4q-1 = forward block q-1) touches forward rows 4q-7 .. 4q+2: it needs exact corrections iff the
    // forward pass rescaled at the end of forward block q-1 or q-2, i.e. unless kb[q-1] == kb[q-2] == kb[q-3]
    int kb1 = kb[q_top - 1], kb2 = kb[q_top - 2], kb3 = kb[q_top - 3];
    wait_group(q_top + 1);
    wait_group(q_top);
    auto load_nib = [&](int q) -> unsigned { // guide bits 4q-5 .. 4q-2 of block q >= 2
        const int b0 = 4 * q - 5;
        return __funnelshift_r(pc.bw[b0 >> 5], pc.bw[(b0 >> 5) + 1], b0 & 31);
    };
    unsigned nib_cur = 0u, nib_nxt = q_top >= 2 ? load_nib(q_top) : 0u;
    // per-block preamble: guide bits one block ahead, the ring group this block reads below itself has landed, the next
    // copy is issued; returns whether block q takes the fast path
    int kA = 0, kB = 0; // rescale amounts at the end of forward blocks q-2 and q-1 of the current block q
    auto preamble = [&](int q) -> int { // 0: generic steps, 1: fast block, 2: fast block with exact corrections
        nib_cur = nib_nxt;
        if (q >= 3) nib_nxt = load_nib(q - 1);
        wait_group(q - 1);
        __syncwarp(); // every lane is done with the rows that the next copy overwrites
        if (q >= 2) issue_group(q - 2);
        kB = kb1 - kb2; kA = kb2 - kb3;
        kb1 = kb2; kb2 = kb3; kb3 = kb[q - 4];
        if (!(4 * q - 1 <= nd - 2 && q >= 2)) return 0;
        return (kA | kB) == 0 ? 1 : 2;
    };
    for (int q = q_top; q >= 1; --q) {
        const int mode = preamble(q);
        if (mode == 2) { // a rescale within reach: the fast block with the exact corrections of its cut products
            const ulonglong2 *rq = ring_q + (size_t)(4 * (q & 3) + 3 + MB) * RQ + kPlaneHalo + lane;
            const unsigned nib = nib_cur;
            const float fA = pow2c(kA), fB = pow2c(kB), fBi = pow2c(-kB);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                f2 bM[P], bD[P];
                float ce[7];
                block_corrections(k, fA, fB, fBi, ce);
                bwd_step<C, ROWS, true, false>(pc, a, st, rq - k * RQ, ce, boff, bM, bD, k);
                bwd_hand_off<C>(st, bM, bD);
                if (k == 0 && kB != 0) bwd_scale<C>(st, pow2i(kB)); // mirror of the forward rescale on row 4q-1
                bwd_band_down<C>(st, (int)(((nib >> (3 - k)) & 1u) ^ 1u), W);
            }
#pragma unroll
            for (int c = 0; c < C; c++) st.rbp[c] -= 4;
        } else if (mode == 1) {
            const ulonglong2 *rq = ring_q + (size_t)(4 * (q & 3) + 3 + MB) * RQ + kPlaneHalo + lane;
            // guide bits 4q-5 .. 4q-2 (fetched one block ahead): step k (s = 4q-1-k) moves on with bit 4q-2-k
            const unsigned nib = nib_cur;
#pragma unroll kBwdUnroll
            for (int k = 0; k < 4; k++) {
                f2 bM[P], bD[P];
                bwd_step<C, ROWS, false, false>(pc, a, st, rq - k * RQ, nullptr, boff, bM, bD, k);
                bwd_hand_off<C>(st, bM, bD);
                bwd_band_down<C>(st, (int)(((nib >> (3 - k)) & 1u) ^ 1u), W);
            }
#pragma unroll
            for (int c = 0; c < C; c++) st.rbp[c] -= 4;
        } else {
            for (int s = min(4 * q - 1, nd - 1); s >= 4 * q - 4; --s) slow_step(s);
        }
        bwd_retire<C, ROWS>(st, pc, raw_of);
    }
    // the columns still in the window after s = 0 (0 .. r) are complete
#pragma unroll
    for (int c = 0; c < C; c++)
        if (st.j[c] >= 0) bwd_flush_col<C>(st, c, raw_of(st.j[c]));
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
template <typename SM>
__device__ __forceinline__ PairCtx make_pair_ctx(const KParams &p, const DevPair &P, SM &sh) {
    PairCtx pc;
    pc.Tb = p.codes + P.tb_off;
    pc.RbP = p.codes + P.rb_off - kCodePad;
    pc.bw = p.bits + P.bits_off;
    pc.Lt = P.Lt; pc.Lr = P.Lr; pc.nd = P.Lt + P.Lr + 1; pc.r = p.radius;
    pc.sEM = (unsigned)__cvta_generic_to_shared(&sh.em[P.model][0]);
    pc.sEI = (unsigned)__cvta_generic_to_shared(&sh.ei[P.model][0]);
    pc.sEC = 0u;
    return pc;
}
// ... of a kernel that runs the backward / table step (its tables hold the paired substitution emissions)
template <typename SM>
__device__ __forceinline__ PairCtx make_pair_ctx_bwd(const KParams &p, const DevPair &P, SM &sh) {
    PairCtx pc = make_pair_ctx(p, P, sh);
    pc.sEC = (unsigned)__cvta_generic_to_shared(&sh.ecp[P.model][0]);
    return pc;
}

template <int C, int ROWS> constexpr int ring_floats2() { return (kRingRows + ring_below(ROWS) + ring_above(ROWS)) * (C * kPlane); }

// Kernel 2: backward pass fused with the table reduction, reading the forward rows of kernel 1.  Finished columns leave
// their 16 raw sums in the pair's scratch; kernel 3 (finalize) turns them into the table.
struct __align__(256) BwdSmem {
    float ecp[2][32];  // [qc*4 + b] = eM(ref b, read base qc)
    float em[2][64];   // [tc*8 + qc]   byte offset tc*32 + qc*4
    float ei[2][64];   // [ctx*8 + qc]  byte offset = the read-row code byte (ctx<<5 | qc<<2)
    float trans[2][12];
    float ftot[kWarpsPerCta][4];
    unsigned long long bar[kWarpsPerCta][4];       // mbarriers of the four ring groups
};
template <typename SM> __device__ __forceinline__ void fill_ecp(SM &sh, const float *__restrict__ models) {
    for (int k = threadIdx.x; k < 2 * 32; k += blockDim.x) sh.ecp[k >> 5][k & 31] = models[(k >> 5) * kModelFloats + kOffEMT + (k & 31)];
}
__device__ __forceinline__ void fill_bwd_tables(BwdSmem &sh, const float *__restrict__ models) {
    fill_ecp(sh, models);
    for (int k = threadIdx.x; k < 2 * 64; k += blockDim.x) {
        const int m = k >> 6, e = k & 63;
        sh.em[m][e] = models[m * kModelFloats + kOffEM + e];
        sh.ei[m][e] = models[m * kModelFloats + kOffEI + e];
        if (e < 12) sh.trans[m][e] = models[m * kModelFloats + e];
    }
    __syncthreads();
}

template <int C, int ROWS>
__global__ void __launch_bounds__(bwd_warps_per_cta(C) * 32, bwd_ctas_per_sm(C, ROWS)) bwdtable_kernel(KParams p) {
    __shared__ BwdSmem sh;
    extern __shared__ __align__(128) unsigned char dyn_smem[]; // kWarpsPerCta rings of forward rows + staged read codes
    fill_bwd_tables(sh, p.models);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int RS = C * kPlane, PADR = 32 * C + 16;
    f2 *ring = reinterpret_cast<f2 *>(dyn_smem) + (size_t)warp * ring_floats2<C, ROWS>();
    unsigned char *rb_s = dyn_smem + (size_t)bwd_warps_per_cta(C) * ring_floats2<C, ROWS>() * sizeof(f2) + (size_t)warp * p.smem_rb; // rb_s[i + PADR]
    const unsigned bars = (unsigned)__cvta_generic_to_shared(&sh.bar[warp][0]);
    unsigned phase = 0u;
    if (lane < 4) mbar_init(bars + 8u * lane, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(p.counter2, 1);
        k = __shfl_sync(kFull, k, 0);
        if (p.pair_lo + k >= p.pair_hi) break;
        const int pi = pair_index(p, k);
        const DevPair P = p.pairs[pi];
        const PairCtx pc = make_pair_ctx_bwd(p, P, sh);
        const BCoef a = load_bcoef(sh.trans[P.model]);
        const float2 *frow = p.frows + (size_t)k * p.frow_stride + kRowShift * RS;
        const int32_t *kb = p.kf + (size_t)k * p.kf_stride + 4;
        const unsigned *info = p.fwdinfo + (size_t)k * p.fwdinfo_stride;
        if (lane < 4) sh.ftot[warp][lane] = __uint_as_float(info[lane]);
        {   // stage the read-row codes of the pair
            const uint8_t *Rb = p.codes + P.rb_off;
            const int nr = P.Lr + 2 * PADR;
            for (int w = lane; w < nr; w += 32) rb_s[w] = Rb[w - PADR];
        }
        __syncwarp();
        backward_pass<C, ROWS>(pc, a, frow, kb, p.raw + (size_t)k * p.raw_stride, sh.ftot[warp], ring, bars, phase, rb_s + PADR);
        __syncwarp();
    }
}

// Kernel 3: raw column sums -> the 14 log-ratios of the table (table - lk).  One CTA per 128 columns of one pair: the raw
// sums (64 B per column) come in and the table rows (56 B per column) go out as contiguous runs through shared memory,
// one thread computes one column in between.
// raw[j] = { S[0..3] | N[0..3] | Vs, Vn, Xp[0..1] | Xp[2], Xm[0..2] } (DESIGN.md 3.2, table identities).
constexpr int kFinCols = 128;
constexpr int kFinRawStride = 5;  // float4 per staged column (4 used): 80-byte stride, conflict-free 16-byte reads
constexpr int kFinOutStride = 15; // floats per staged output column (14 used)
__global__ void __launch_bounds__(kFinCols) finalize_kernel(KParams p) {
    __shared__ float4 sraw[(kFinCols + 3) * kFinRawStride];
    __shared__ float sout[kFinCols * kFinOutStride];
    const int k = blockIdx.y;
    const int pi = pair_index(p, k);
    const DevPair P = p.pairs[pi];
    const int Lt = P.Lt;
    const int j0 = blockIdx.x * kFinCols;
    if (j0 > Lt) return;
    const int ncol = min(kFinCols, Lt + 1 - j0);       // columns of this CTA
    const int nraw = min(kFinCols + 3, Lt + 1 - j0);   // + the three columns to the right (multi-base deletions)
    const float4 *sg = p.raw + (size_t)k * p.raw_stride + (size_t)j0 * 4;
    for (int t = threadIdx.x; t < nraw * 4; t += kFinCols) sraw[(t >> 2) * kFinRawStride + (t & 3)] = __ldcg(sg + t);
    const unsigned *info = p.fwdinfo + (size_t)k * p.fwdinfo_stride;
    const float fin_raw = __uint_as_float(info[0]);
    const int e_fin = (int)(__float_as_uint(fin_raw) >> 23) - 127;
    const float boff = fin_raw > 0.f ? pow2i(max(-120, min(120, kProductExp - e_fin))) : 1.f;
    const float fin = fin_raw * boff;
    __syncthreads();
    const int c = threadIdx.x, jj = j0 + c;
    if (c < ncol) {
        const float4 q0 = sraw[c * kFinRawStride], q1 = sraw[c * kFinRawStride + 1], q2 = sraw[c * kFinRawStride + 2],
                     q3 = sraw[c * kFinRawStride + 3];
        const float s4[4] = { q0.x, q0.y, q0.z, q0.w };
        const float n4[4] = { q1.x, q1.y, q1.z, q1.w };
        const float vs = q2.x, vn = q2.y;
        const float xp[3] = { q2.z, q2.w, q3.x };
        const int tcode = p.codes[P.tb_off + jj + 1];
        float ref = fin;
        if (jj < Lt) ref = s4[tcode & 3] + vs;
        auto dlog = [](float num, float den) -> float {
            return (num > 0.f && den > 0.f) ? __logf(num / den) : kDeltaNeg; // lg2.approx * ln2: abs. error < 1e-6 here, log(1) = 0 exactly
        };
        float *o = sout + c * kFinOutStride;
        const bool all_rows = p.rows == 14;
#pragma unroll
        for (int b = 0; b < 4; b++) o[b] = (jj < Lt) ? dlog(s4[b] + vs, ref) : kDeltaNeg;
#pragma unroll
        for (int b = 0; b < 4; b++) o[4 + b] = dlog(n4[b] + vn, ref);
#pragma unroll
        for (int e = 1; e <= 3; e++) {
            o[7 + e] = (all_rows && jj + e <= Lt) ? dlog(xp[e - 1], ref) : kDeltaNeg;
            float v = kDeltaNeg;
            if ((all_rows || e == 1) && jj + e <= Lt) {
                const float4 qe = sraw[(c + e) * kFinRawStride + 3];
                float acc = e == 1 ? qe.y : (e == 2 ? qe.z : qe.w);
                if (jj + e == Lt) acc += __uint_as_float(info[e]) * boff;
                v = dlog(acc, ref);
            }
            o[10 + e] = v;
        }
    }
    __syncthreads();
    float *og = p.out_delta + P.tab_off + (size_t)j0 * kNumRow;
    for (int n = threadIdx.x; n < ncol * kNumRow; n += kFinCols) og[n] = sout[(n / kNumRow) * kFinOutStride + (n % kNumRow)];
}

template <int C>
__global__ void __launch_bounds__(kWarpsPerCta * 32) likelihood_kernel(KParams p) {
    __shared__ SmemLayout sh;
    fill_tables(sh, p.models);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (;;) {
        int pi = 0;
        if (lane == 0) pi = atomicAdd(p.counter, 1);
        pi = __shfl_sync(kFull, pi, 0);
        if (pi >= p.n_pairs) break;
        const DevPair P = p.pairs[pi];
        const PairCtx pc = make_pair_ctx(p, P, sh);
        const Coef a = load_coef(sh.trans[P.model], sh.cpair[P.model]);
        int Ktot;
        forward_pass<C, 0>(pc, a, nullptr, nullptr, sh.ftot[warp], Ktot, nullptr);
        const float fin = sh.ftot[warp][0];
        if (lane == 0)
            p.out_lk[pi] = fin > 0.f ? log((double)fin) - (double)Ktot * 0.6931471805599453 : -INFINITY;
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// K4: expected counts of one Baum-Welch step (kiley fit_antidiagonal_par_multiple, model_tune.rs:151).
// Forward pass stores (F_M, F_I, F_D) of every cell; the backward pass recomputes the in-sums gM, gI, gD of the
// same cell and accumulates, per lane,
//   T[s][move]      = sum F_s(i,j) * g_move(i,j)              (9 sums; times a[s][move] / P on the way out)
//   binM[t[j]][q[i]] += toM(i,j) * gM(i,j)      binI[ctx][q[i]] += toI(i,j) * gI(i,j)
// bins live in a lane-private column of shared memory (bin*32 + lane: no conflicts, no atomics).
// ------------------------------------------------------------------------------------------------
constexpr int kFitBins = 40; // [tc or ctx 0..4][qc 0..7]

template <int C>
__device__ __forceinline__ void fit_backward(const PairCtx &pc, const Coef &a, const float *trans, const float4 *__restrict__ srow0,
                                             const int32_t *__restrict__ kf, volatile float *s_ftot, float *binM, float *binI,
                                             double *__restrict__ acc45) {
    constexpr int NSLOT = 32 * C;
    const int lane = threadIdx.x & 31;
    const int Lt = pc.Lt, Lr = pc.Lr, nd = pc.nd;
    const float fin_raw = s_ftot[0];
    if (!(fin_raw > 0.f)) return;
    const int e_fin = (int)(__float_as_uint(fin_raw) >> 23) - 127;
    const float boff = pow2i(max(-120, min(120, kProductExp - e_fin)));
    const float inv_p = 1.f / (fin_raw * boff);
    const float mm = trans[0], mi = trans[1], md = trans[2], im = trans[3], ii = trans[4], id = trans[5], dm = trans[6],
                di = trans[7], dd = trans[8];
    for (int k = 0; k < kFitBins; k++) { binM[k * 32 + lane] = 0.f; binI[k * 32 + lane] = 0.f; }
    int j[C];
    unsigned tcB[C], win[C];
    float BI[C], inD[C], inMa[C], inMb[C];
    float T[9];
#pragma unroll
    for (int k = 0; k < 9; k++) T[k] = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int sigma = lane * C + c;
        j[c] = Lt - ((Lt - sigma) & (NSLOT - 1));
        tcB[c] = pc.sEM + ((unsigned)pc.Tb[j[c] + 1] << 5);
        win[c] = 0u;
        BI[c] = inD[c] = inMa[c] = inMb[c] = 0.f;
    }
    int cen = Lr;
    unsigned bword = pc.bw[(nd - 1) >> 5];
    const float4 *rp = srow0 + (ptrdiff_t)(nd - 1) * NSLOT + lane * C;
    int kcur = kf[nd - 1];
    for (int s = nd - 1; s >= 0; --s, rp -= NSLOT) {
        const int lo = band_lo(cen, pc.r, s, Lt), hi = band_hi(cen, pc.r, s, Lr);
        const int W = hi - lo, A = s - lo;
        if (s == nd - 1 || (s & 3) == 3) {
#pragma unroll
            for (int c = 0; c < C; c++) win[c] = win_down(pc.RbP, s - j[c] + 1);
        }
        float bM[C], bD[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int x = A - j[c];
            const bool valid = (unsigned)x <= (unsigned)W;
            const unsigned w = win[c];
            win[c] = w >> 8;
            const float em = lds_f32(tcB[c] | (w & 0x1cu));
            const float ei = lds_f32(pc.sEI | (w & 0xffu));
            const float gM = em * inMb[c], gI = ei * BI[c], gD = inD[c];
            float m_ = mm * gM + mi * gI + md * gD;
            float i_ = im * gM + ii * gI + id * gD;
            float d_ = dm * gM + di * gI + dd * gD;
            if (s == nd - 1) { const float t = (j[c] == Lt) ? boff : 0.f; m_ = t; i_ = t; d_ = t; } // B(Lr, Lt) only
            if (!valid) { m_ = 0.f; i_ = 0.f; d_ = 0.f; }
            const float4 F = rp[c]; // zero outside the band
            T[0] = fmaf(F.x, gM, T[0]); T[1] = fmaf(F.x, gI, T[1]); T[2] = fmaf(F.x, gD, T[2]);
            T[3] = fmaf(F.y, gM, T[3]); T[4] = fmaf(F.y, gI, T[4]); T[5] = fmaf(F.y, gD, T[5]);
            T[6] = fmaf(F.z, gM, T[6]); T[7] = fmaf(F.z, gI, T[7]); T[8] = fmaf(F.z, gD, T[8]);
            const float toM = mm * F.x + im * F.y + dm * F.z;
            const float toI = mi * F.x + ii * F.y + di * F.z;
            const unsigned bm = (((tcB[c] - pc.sEM) >> 2) | ((w >> 2) & 7u)) * 32u + lane;
            const unsigned bi = ((w >> 2) & 63u) * 32u + lane;
            binM[bm] += toM * gM;
            binI[bi] += toI * gI;
            bM[c] = m_; bD[c] = d_; BI[c] = i_;
            if (x < 0) { // above the band for good
                j[c] -= NSLOT;
                tcB[c] = pc.sEM + ((unsigned)pc.Tb[j[c] + 1] << 5);
                win[c] = win_down(pc.RbP, (s - 1) - j[c] + 1);
            }
        }
        const float rM = __shfl_sync(kFull, bM[0], (lane + 1) & 31);
        const float rD = __shfl_sync(kFull, bD[0], (lane + 1) & 31);
#pragma unroll
        for (int c = 0; c < C - 1; c++) { inMb[c] = inMa[c]; inMa[c] = bM[c + 1]; inD[c] = bD[c + 1]; }
        inMb[C - 1] = inMa[C - 1]; inMa[C - 1] = rM; inD[C - 1] = rD;
        if (s > 0) {
            const int kprev = kf[s - 1];
            if (kcur != kprev) {
                const float sc = pow2i(kcur - kprev);
#pragma unroll
                for (int c = 0; c < C; c++) { BI[c] *= sc; inD[c] *= sc; inMa[c] *= sc; inMb[c] *= sc; }
            }
            kcur = kprev;
            cen -= (bword >> ((s - 1) & 31)) & 1u;
            if (((s - 1) & 31) == 0 && s > 1) bword = pc.bw[(s - 2) >> 5];
        }
    }
    // reduce over lanes, scale by a[s][move] / P, add into the strand's accumulators
    const float tr[9] = { mm, mi, md, im, ii, id, dm, di, dd };
#pragma unroll
    for (int k = 0; k < 9; k++) {
        float v = T[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) atomicAdd(&acc45[k], (double)(v * tr[k]) * (double)inv_p);
    }
    __syncwarp();
    for (int k = 0; k < kFitBins; k++) {
        float vm = binM[k * 32 + lane], vi = binI[k * 32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { vm += __shfl_xor_sync(kFull, vm, o); vi += __shfl_xor_sync(kFull, vi, o); }
        const int hi8 = k >> 3, q = k & 7;
        if (lane == 0 && q < 4) {
            if (hi8 < 4) atomicAdd(&acc45[9 + 4 * hi8 + q], (double)vm * (double)inv_p);
            atomicAdd(&acc45[25 + 4 * hi8 + q], (double)vi * (double)inv_p);
        }
    }
    __syncwarp();
}

struct FitSmem {
    float binM[kWarpsPerCta][kFitBins * 32];
    float binI[kWarpsPerCta][kFitBins * 32];
};

template <int C>
__global__ void __launch_bounds__(kWarpsPerCta * 32) fit_kernel(KParams p, double *acc90) {
    __shared__ SmemLayout sh;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    FitSmem &fs = *reinterpret_cast<FitSmem *>(dyn_smem);
    fill_tables(sh, p.models);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wslot = (size_t)blockIdx.x * kWarpsPerCta + warp;
    // the same scratch, viewed as one float4 per slot: stride (nd+6) * (NSLOT+8) float2 >= nd * NSLOT float4 / 2 is NOT
    // guaranteed, so the host sizes frow_stride for this kernel as 2 * (max_nd + 1) * NSLOT float2
    float2 *frow = p.frows + wslot * p.frow_stride;
    int32_t *kf = p.kf + wslot * p.kf_stride + 3;
    for (;;) {
        int pi = 0;
        if (lane == 0) pi = atomicAdd(p.counter, 1);
        pi = __shfl_sync(kFull, pi, 0);
        if (pi >= p.n_pairs) break;
        const DevPair P = p.pairs[pi];
        const PairCtx pc = make_pair_ctx(p, P, sh);
        const Coef a = load_coef(sh.trans[P.model], sh.cpair[P.model]);
        int Ktot;
        forward_pass<C, 2>(pc, a, frow, kf, sh.ftot[warp], Ktot, nullptr);
        __syncwarp();
        const float fin = sh.ftot[warp][0];
        if (lane == 0)
            p.out_lk[pi] = fin > 0.f ? log((double)fin) - (double)Ktot * 0.6931471805599453 : -INFINITY;
        fit_backward<C>(pc, a, sh.trans[P.model], reinterpret_cast<const float4 *>(frow), kf, sh.ftot[warp],
                        fs.binM[warp], fs.binI[warp], acc90 + 45 * P.model);
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// The lean forward pass (all three modification-table kernels) and the fused kernel (v10 / v11): the modification table as
// ONE kernel whose DP matrices never touch HBM (north star: "the backward pass is fused with the modification-table
// reduction so the full DP matrix never touches HBM").
//
//   pass 1   lean forward pass (no row stores): likelihood, one scale exponent per block of four anti-diagonals (kb, 1 int /
//            4 rows) and a CHECKPOINT of the forward state every SEG anti-diagonals (inMb, inMa, inD, toI of every slot + the
//            band offset: C*512 + 16 bytes, i.e. 65 B per row instead of the 608 B of a forward row);
//   pass 2   top-down over segments of SEG anti-diagonals: the forward rows of the segment are RECOMPUTED from its
//            checkpoint into a per-warp shared-memory buffer (same row layout as the ring of the rows variant), then the
//            backward / table step runs over them.  The cuts reach three rows up and down, so the buffer keeps the eight
//            lowest rows of the segment above (moved up by one shared-memory copy per segment) and a segment's rows start four
//            rows below its first backward row.  Checkpoints are fetched one segment ahead by a bulk async copy (mbarrier).
// The recomputation replays pass 1 instruction for instruction (explicit fma / mul intrinsics) and takes its rescale decisions
// from kb, so the rows are the ones pass 1 saw.  The rows variant runs the same forward code once, with the rows going to HBM.
// ------------------------------------------------------------------------------------------------
template <int C> __host__ __device__ constexpr int seg_rows() { return C == 2 ? 16 : 8; }
constexpr int kSegKeep = 8;   // rows of the segment above that stay in the buffer
constexpr int kSegBelow = 4;  // a segment's rows start this far below its first backward row
template <int C> __host__ __device__ constexpr int seg_buf_rows() { return seg_rows<C>() + kSegKeep; }
template <int C> __host__ __device__ constexpr int ckpt_bytes() { return C * 512 + 16; }
// landing zone of a checkpoint in shared memory, followed by the warp's mbarrier (8 B at +0) and its four end sums (16 B at +16)
template <int C> __host__ __device__ constexpr int ckpt_stage_bytes() { return (ckpt_bytes<C>() + 32 + 127) & ~127; }
constexpr int fused_ctas_per_sm(int C) { return C == 2 ? 3 : (C == 4 ? 2 : 1); }

// shared tables of the forward kernel: emissions are looked up with the raw read-row code byte w = ctx<<5 | q<<2, eM at
// (row of the column's base) | (w & 0x1c), eI at base + w: every distinct address of a warp sits in its own bank
struct __align__(256) FwdTabSmem {
    float em[2][64];  // [tc*8 + qc]
    float ei[2][64];  // [ctx*8 + qc], byte offset = the read-row code byte
    unsigned long long cdup[2][9]; // (t[k], t[k]): the nine transitions as broadcast pairs
    float ftot[kWarpsPerCta][4];
};
// ... of the fused kernel: the same + what the backward / table step needs
struct __align__(256) LeanSmem {
    float ecp[2][32];  // backward: [qc*4 + b] = eM(ref b, read base qc)
    float em[2][64];
    float ei[2][64];
    float trans[2][12];
    unsigned long long cdup[2][9];
    float ftot[kWarpsPerCta][4];
    unsigned long long bar[kWarpsPerCta];
};
template <typename SM> __device__ __forceinline__ void fill_fwd_part(SM &sh, const float *__restrict__ models) {
    for (int k = threadIdx.x; k < 2 * 64; k += blockDim.x) {
        const int m = k >> 6, e = k & 63;
        sh.em[m][e] = models[m * kModelFloats + kOffEM + e];
        sh.ei[m][e] = models[m * kModelFloats + kOffEI + e];
        if (e < 9) sh.cdup[m][e] = mk2(models[m * kModelFloats + e], models[m * kModelFloats + e]);
    }
}
__device__ __forceinline__ void fill_lean_tables(LeanSmem &sh, const float *__restrict__ models) {
    fill_fwd_part(sh, models);
    fill_ecp(sh, models);
    for (int k = threadIdx.x; k < 24; k += blockDim.x) sh.trans[k / 12][k % 12] = models[(k / 12) * kModelFloats + k % 12];
    __syncthreads();
}

// kb[q] of one warp slot as (kernel-parameter base, 32-bit offset): no 64-bit pointer has to stay in registers
struct KbRef {
    int32_t *base;
    int off;
    __device__ __forceinline__ int32_t &operator[](int i) const { return base[off + i]; }
};

struct LeanPair { // warp-uniform view of one pair for the lean forward pass
    const uint8_t *Tb;        // Tb[j] = code of t[j-1]
    const uint32_t *bw;       // guide bits
    const unsigned char *rb0; // staged read codes in shared memory: rb0[i] = code byte of read row i
    unsigned sEM, sEI;        // shared-space addresses of the pair's eM / eI tables (256-byte aligned)
    int Lt, Lr, nd, r;
};

// Forward state: lane l owns the slots l + 32c; the cells of slots l + 64p and l + 64p + 32 are packed into one fp32 pair
// (lo = the lower slot): every arithmetic instruction of the step is a packed one, the coefficients are (v, v) pairs and
// nothing moves between the halves of a pair (the hand-off to the right-hand neighbour column is lane -> lane+1 for both).
template <int C> struct LeanFwd {
    static constexpr int P = C / 2;
    int x[C];
    const unsigned char *rbp[C]; // staged code of the slot's cell in the first row of the current block
    unsigned emrow[C];           // shared address of the eM row of the slot's column
    const uint8_t *tnext[C];     // &Tb[j + 2*NSLOT]: code of the column after the next one (the slot's column is tnext - Tb - 2*NSLOT)
    unsigned tcn[C];             // template code of the column the slot takes next, fetched one column-life ahead
    f2 msk[P], toI[P], inD[P], inMa[P], inMb[P];
};
struct LeanCoef { f2 mm, im, dm, md, id, dd, mi, ii, di; }; // (v, v) pairs
__device__ __forceinline__ LeanCoef load_lean_coef(const unsigned long long *d) {
    LeanCoef c; // d[k] = (t[k], t[k]), t = mat_mat, mat_ins, mat_del, ins_mat, ins_ins, ins_del, del_mat, del_ins, del_del
    c.mm = d[0]; c.mi = d[1]; c.md = d[2]; c.im = d[3]; c.ii = d[4]; c.id = d[5]; c.dm = d[6]; c.di = d[7]; c.dd = d[8];
    return c;
}
__device__ __forceinline__ void set_lo(f2 &v, float x) { v = mk2(x, hi2(v)); }
__device__ __forceinline__ void set_hi(f2 &v, float x) { v = mk2(lo2(v), x); }

// slot state of anti-diagonal s0 (a multiple of four) from the number of band moves so far: the slot holds the one column
// j == sigma (mod NSLOT) with x <= W, which is what the per-block retargeting of the forward pass leaves at block boundaries
template <int C>
__device__ __forceinline__ void lean_seed(LeanFwd<C> &st, const LeanPair &lp, const int s0, const int ups) {
    constexpr int NSLOT = 32 * C;
    const int lane = threadIdx.x & 31, W = 2 * lp.r;
    float m[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int sigma = lane + 32 * c;
        const int t = lp.r - sigma + ups;
        const int w = t > W ? (t - W + NSLOT - 1) / NSLOT : 0;
        const int j = sigma + w * NSLOT;
        st.x[c] = t - w * NSLOT;
        st.rbp[c] = lp.rb0 + (s0 - j);
        st.emrow[c] = lp.sEM + ((unsigned)lp.Tb[j] << 5);
        st.tcn[c] = lp.Tb[j + NSLOT];
        st.tnext[c] = lp.Tb + j + 2 * NSLOT;
        m[c] = (unsigned)st.x[c] <= (unsigned)W ? 1.f : 0.f;
    }
#pragma unroll
    for (int p = 0; p < C / 2; p++) st.msk[p] = mk2(m[2 * p], m[2 * p + 1]);
}

// Rows [s_begin, s_end) of the forward pass, s_begin a multiple of four, s_end a multiple of four or nd.
// MODE 0 (fused, pass 1): rescale decisions by warp maximum, kb[q] and checkpoints written, end sums recorded.
// MODE 1 (fused, recompute): rescales replayed from kb, (toM, toD) of every cell written to the shared-memory rows at `wrow`
// MODE 2 (rows variant): decisions and kb as MODE 0, no checkpoints, rows written to global memory at `wrow`
// (wrow = entry `lane` of pair plane 0 of row s_begin, as f2 *).
template <int C, int MODE>
__device__ __forceinline__ void lean_forward(const LeanCoef &a, LeanFwd<C> &st, const LeanPair &lp, const int s_begin,
                                             const int s_end, int &K, int &ups, const KbRef kb,
                                             unsigned char *__restrict__ ckpt_g, f2 *wrow, volatile float *s_ftot, unsigned &ev) {
    constexpr int NSLOT = 32 * C, RS = C * kPlane, SEG = seg_rows<C>(), CKB = ckpt_bytes<C>(), P = C / 2;
    constexpr bool DECIDE = MODE != 1; // this pass takes the rescale decisions (and records the end sums)
    const int lane = threadIdx.x & 31, W = 2 * lp.r, nd = lp.nd;
    // the replicated entries past the ends of a plane: lanes 0..2 also write entry 32 + lane, lanes 29..31 entry lane - 32
    const int halo = (lane < kPlaneHalo) ? 64 : ((lane >= 32 - kPlaneHalo) ? -64 : 0); // in f2
    // the cells of one anti-diagonal: out-sums (toM, toD, toI) of the slot pairs
    auto cells = [&](const int s, const int kk, const bool special, f2 (&tM)[P], f2 (&tD)[P], f2 (&nI)[P]) {
#pragma unroll
        for (int p = 0; p < P; p++) {
            const unsigned w0 = st.rbp[2 * p][kk], w1 = st.rbp[2 * p + 1][kk];
            const f2 em = mk2(lds_f32(st.emrow[2 * p] | (w0 & 0x1cu)), lds_f32(st.emrow[2 * p + 1] | (w1 & 0x1cu)));
            const f2 ei = mk2(lds_f32(lp.sEI + w0), lds_f32(lp.sEI + w1));
            f2 M = mul2(em, st.inMb[p]);
            const f2 I = mul2(ei, st.toI[p]);
            const f2 D = st.inD[p];
            if (special && s == 0 && p == 0 && lane == 0) set_lo(M, 1.f); // F_M(0, 0) = 1
            tM[p] = mul2(fma2(a.dm, D, fma2(a.im, I, mul2(a.mm, M))), st.msk[p]);
            tD[p] = mul2(fma2(a.dd, D, fma2(a.id, I, mul2(a.md, M))), st.msk[p]);
            nI[p] = mul2(fma2(a.di, D, fma2(a.ii, I, mul2(a.mi, M))), st.msk[p]);
            if (DECIDE && special && s >= nd - 4) { // the end sums F(Lr, Lt - d), d = 0..3
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int j = (int)(st.tnext[2 * p + h] - lp.Tb) - 2 * NSLOT;
                    const float mk = h ? hi2(st.msk[p]) : lo2(st.msk[p]);
                    if (mk != 0.f && s - j == lp.Lr && j >= lp.Lt - 3)
                        s_ftot[lp.Lt - j] = h ? hi2(M) + hi2(I) + hi2(D) : lo2(M) + lo2(I) + lo2(D);
                }
            }
        }
    };
    auto rescale_by = [&](const int kx, f2 (&tM)[P], f2 (&tD)[P], f2 (&nI)[P]) {
        const f2 sc = bc2(pow2i(kx));
#pragma unroll
        for (int p = 0; p < P; p++) { tM[p] = mul2(tM[p], sc); tD[p] = mul2(tD[p], sc); nI[p] = mul2(nI[p], sc); st.inMa[p] = mul2(st.inMa[p], sc); }
    };
    auto finish = [&](const int kk, f2 (&tM)[P], f2 (&tD)[P], f2 (&nI)[P]) {
        auto cM = [&](int c) -> float { c = (c + C) % C; return (c & 1) ? hi2(tM[c / 2]) : lo2(tM[c / 2]); };
        auto cD = [&](int c) -> float { c = (c + C) % C; return (c & 1) ? hi2(tD[c / 2]) : lo2(tD[c / 2]); };
        if (MODE >= 1) { // pair plane p, entry lane = (toM, toD) of the slots lane + 64p and lane + 64p + 32
            f2 *w = wrow + kk * RS;
#pragma unroll
            for (int p = 0; p < P; p++) *reinterpret_cast<ulonglong2 *>(w + p * 2 * kPlane) = make_ulonglong2(tM[p], tD[p]);
            // entry 32 + t holds the slots (32 + t + 64p, 64 + t + 64p) = cells (2p+1, 2p+2) of lane t; entry -1 - t the
            // slots (64p - 1 - t, 64p + 31 - t) = cells (2p-1, 2p) of lane 31 - t
            if (C == 2) {
                if (halo != 0)
                    *reinterpret_cast<ulonglong2 *>(w + halo) = make_ulonglong2(mk2(cM(1), cM(0)), mk2(cD(1), cD(0)));
            } else {
                if (lane < kPlaneHalo) {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        *reinterpret_cast<ulonglong2 *>(w + p * 2 * kPlane + 64) =
                            make_ulonglong2(mk2(cM(2 * p + 1), cM(2 * p + 2)), mk2(cD(2 * p + 1), cD(2 * p + 2)));
                }
                if (lane >= 32 - kPlaneHalo) {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        *reinterpret_cast<ulonglong2 *>(w + p * 2 * kPlane - 64) =
                            make_ulonglong2(mk2(cM(2 * p - 1), cM(2 * p)), mk2(cD(2 * p - 1), cD(2 * p)));
                }
            }
        }
        // hand (toM, toD) to the right-hand neighbour column: slot sigma+1 = the same cell of lane+1; lane 0 receives from the
        // PREVIOUS cell of lane 31 (slot 32c - 1, wrapping)
        float rM[C], rD[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            rM[c] = __shfl_sync(kFull, cM(c), (lane + 31) & 31);
            rD[c] = __shfl_sync(kFull, cD(c), (lane + 31) & 31);
        }
        const bool seam = lane == 0;
#pragma unroll
        for (int p = 0; p < P; p++) {
            st.inMb[p] = st.inMa[p];
            st.inMa[p] = mk2(seam ? rM[(2 * p - 1 + C) % C] : rM[2 * p], seam ? rM[2 * p] : rM[2 * p + 1]);
            st.inD[p] = mk2(seam ? rD[(2 * p - 1 + C) % C] : rD[2 * p], seam ? rD[2 * p] : rD[2 * p + 1]);
            st.toI[p] = nI[p];
        }
    };
    // anti-diagonal s -> s+1: when the centre stays (guide bit 0, up = 1) every cell moves one row up inside the window.
    // Unconditional: a predicated version issues its dozen instructions whether or not the band moves.
    auto masks = [&]() {
        float m[C];
#pragma unroll
        for (int c = 0; c < C; c++) m[c] = (unsigned)st.x[c] <= (unsigned)W ? 1.f : 0.f;
#pragma unroll
        for (int p = 0; p < P; p++) st.msk[p] = mk2(m[2 * p], m[2 * p + 1]);
    };
    // a slot whose cell left the band at the top (x > W) moves on to column j + NSLOT (at the end of a block of four: the
    // slot ring has NSLOT - (W+1) >= 3 spare slots, and until then its mask is zero)
    auto band_up = [&](const int up, const bool retarget) {
#pragma unroll
        for (int c = 0; c < C; c++) {
            st.x[c] += up;
            if (retarget && st.x[c] > W) {
                st.x[c] -= NSLOT; st.rbp[c] -= NSLOT;
                st.emrow[c] = lp.sEM + (st.tcn[c] << 5);
                st.tcn[c] = *st.tnext[c];
                st.tnext[c] += NSLOT;
            }
        }
        masks();
    };
    unsigned bword = lp.bw[s_begin >> 5];
    for (int s = s_begin; s < s_end; s += 4) {
        if ((s & 31) == 0) bword = lp.bw[s >> 5];
        const unsigned nib = bword >> (s & 31);
        if (MODE == 0 && ((s + kSegBelow) & (SEG - 1)) == 0 && s > 0) { // checkpoint: the state before row s
            unsigned char *ck = ckpt_g + (size_t)((s + kSegBelow) / SEG) * CKB;
#pragma unroll
            for (int p = 0; p < P; p++) {
                reinterpret_cast<float4 *>(ck)[(2 * p) * 32 + lane] = make_float4(lo2(st.inMb[p]), hi2(st.inMb[p]), lo2(st.inMa[p]), hi2(st.inMa[p]));
                reinterpret_cast<float4 *>(ck)[(2 * p + 1) * 32 + lane] = make_float4(lo2(st.inD[p]), hi2(st.inD[p]), lo2(st.toI[p]), hi2(st.toI[p]));
            }
            if (lane == 0) *reinterpret_cast<int4 *>(ck + C * 512) = make_int4(ups, K, 0, 0);
        }
        if (MODE == 1) ev &= ~(1u << ((s >> 2) & 31)); // ev bit (b & 31): the forward pass rescaled at the end of block b
        if (s >= 4 && s + 3 < nd - 4) {
            const int Knew = (MODE == 1) ? kb[s >> 2] : 0;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                f2 tM[P], tD[P], nI[P];
                cells(s + kk, kk, false, tM, tD, nI);
                if (kk == 3) {
                    if (DECIDE) {
                        if (s + 3 < nd - 8) {
                            float v = fmaxf(lo2(tM[0]), hi2(tM[0]));
#pragma unroll
                            for (int p = 1; p < P; p++) v = fmaxf(v, fmaxf(lo2(tM[p]), hi2(tM[p])));
                            const unsigned mx = __reduce_max_sync(kFull, __float_as_uint(v));
                            const int e = (int)(mx >> 23) - 127;
                            if (mx != 0u && e < kScaleLow) {
                                const int kx = min(kScaleTarget - e, kScaleStep);
                                rescale_by(kx, tM, tD, nI);
                                K += kx;
                            }
                        }
                    } else if (Knew != K) {
                        rescale_by(Knew - K, tM, tD, nI);
                        K = Knew;
                        ev |= 1u << ((s >> 2) & 31);
                    }
                }
                finish(kk, tM, tD, nI);
                band_up((int)((~nib >> kk) & 1u), kk == 3);
            }
            ups += __popc(~nib & 15u);
        } else { // the first four and the last anti-diagonals: start cell, end sums, no rescale
            for (int kk = 0; kk < 4 && s + kk < s_end; kk++) {
                f2 tM[P], tD[P], nI[P];
                cells(s + kk, kk, true, tM, tD, nI);
                finish(kk, tM, tD, nI);
                if (s + kk < nd - 1) {
                    const int up = (int)((~nib >> kk) & 1u);
                    band_up(up, true);
                    ups += up;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; c++) st.rbp[c] += 4;
        if (MODE >= 1) wrow += 4 * RS;
        if (DECIDE && lane == 0) kb[s >> 2] = K;
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 1 of the rows variant: the forward pass of every pair of the wave.  Writes the forward rows ((toM, toD) of every
// cell, 608 B per anti-diagonal at C = 2), kb, the four end sums and the likelihood.  HBM-write bound.
// ------------------------------------------------------------------------------------------------
constexpr int fwd_ctas_per_sm(int C) { return C == 2 ? 6 : (C == 4 ? 3 : 2); }
template <int C>
__global__ void __launch_bounds__(kWarpsPerCta * 32, fwd_ctas_per_sm(C)) fwdrows_kernel(KParams p) {
    __shared__ FwdTabSmem sh;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    fill_fwd_part(sh, p.models);
    __syncthreads();
    constexpr int NSLOT = 32 * C, RS = C * kPlane, PADR = NSLOT + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *rb_s = dyn_smem + (size_t)warp * p.smem_rb; // rb_s[i + PADR] = code byte of read row i
    volatile float *s_ftot = sh.ftot[warp];
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(p.counter, 1);
        k = __shfl_sync(kFull, k, 0);
        if (p.pair_lo + k >= p.pair_hi) break;
        const int pi = pair_index(p, k);
        const DevPair P = p.pairs[pi];
        const int Lt = P.Lt, Lr = P.Lr, nd = Lt + Lr + 1;
        float2 *frow = p.frows + (size_t)k * p.frow_stride + kRowShift * RS; // row 0 (kRowShift zero rows below)
        const KbRef kb = { p.kf, (int)((size_t)k * p.kf_stride) + 4 };      // kb[q], q >= -4
        unsigned *info = p.fwdinfo + (size_t)k * p.fwdinfo_stride;
        {   // stage the read codes of this pair; rows -kRowShift..-1 (before the first anti-diagonal) read as zero, exponent 0
            const uint8_t *Rb = p.codes + P.rb_off;
            const int nr = Lr + 2 * PADR;
            for (int w = lane; w < nr; w += 32) rb_s[w] = Rb[w - PADR];
            for (int w = lane; w < kRowShift * RS; w += 32) frow[w - kRowShift * RS] = make_float2(0.f, 0.f);
            if (lane < 4) { kb[lane - 4] = 0; s_ftot[lane] = 0.f; }
        }
        __syncwarp();
        LeanPair lp;
        lp.Tb = p.codes + P.tb_off; lp.bw = p.bits + P.bits_off; lp.rb0 = rb_s + PADR;
        lp.Lt = Lt; lp.Lr = Lr; lp.nd = nd; lp.r = p.radius;
        lp.sEM = (unsigned)__cvta_generic_to_shared(&sh.em[P.model][0]);
        lp.sEI = (unsigned)__cvta_generic_to_shared(&sh.ei[P.model][0]);
        int K = 0;
        {
            LeanFwd<C> st;
#pragma unroll
            for (int q = 0; q < C / 2; q++) st.inMb[q] = st.inMa[q] = st.inD[q] = st.toI[q] = 0ull;
            int ups = 0;
            lean_seed<C>(st, lp, 0, 0);
            const LeanCoef la = load_lean_coef(sh.cdup[P.model]);
            unsigned ev_unused = 0u;
            lean_forward<C, 2>(la, st, lp, 0, nd, K, ups, kb, nullptr, reinterpret_cast<f2 *>(frow) + 2 * (kPlaneHalo + lane), s_ftot, ev_unused);
        }
        // rows just past the last anti-diagonal read as zero; the exponent stays at K
        for (int w = lane; w < kRowsAbove * RS; w += 32) frow[(size_t)nd * RS + w] = make_float2(0.f, 0.f);
        for (int q = ((nd + 3) >> 2) + lane; q <= ((nd + 3) >> 2) + 2; q += 32) kb[q] = K;
        __syncwarp();
        const float fin = s_ftot[0];
        if (lane < 4) info[lane] = __float_as_uint(s_ftot[lane]);
        if (lane == 4) info[4] = (unsigned)K;
        if (lane == 0)
            p.out_lk[pi] = fin > 0.f ? log((double)fin) - (double)K * 0.6931471805599453 : -INFINITY;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// likelihood of many SHORT pairs at a small radius (the calibrations of likelihood_gains.rs:6-39,253-315:
// 1.8e5 / 1e6 pairs of ~100 bp at radius 10 / 12), TWO pairs per warp.  At radius <= 14 the slot ring needs 2r + 4 <= 32 slots,
// so lane l owns slot l of pair A AND slot l of pair B: the lo half of every fp32 pair belongs to A, the hi half to B, every
// arithmetic instruction is a packed one, the transition coefficients are (A's model, B's model) pairs, the hand-off to the
// right-hand neighbour column is lane -> lane+1 for both (no seam).  Band bits, rescale decisions and end sums are per pair; a
// pair that is shorter than its partner stops moving once its last anti-diagonal is done.  Same arithmetic per cell as
// lean_forward (fwdrows_kernel); likelihood_kernel may rescale one block earlier, which moves ln(fin) - K ln 2 in its last bits.
// ------------------------------------------------------------------------------------------------
constexpr int kPackRadiusMax = 14;
__global__ void __launch_bounds__(kWarpsPerCta * 32) likelihood_pairs2_kernel(KParams p) {
    __shared__ FwdTabSmem sh;
    __shared__ float ftot2[kWarpsPerCta][8];
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    fill_fwd_part(sh, p.models);
    __syncthreads();
    constexpr int NSLOT = 32, PADR = NSLOT + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *rbs = dyn_smem + (size_t)warp * 2 * p.smem_rb;
    volatile float *s_ftot = ftot2[warp];
    const int W = 2 * p.radius, r = p.radius;
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(p.counter, 1);
        k = __shfl_sync(kFull, k, 0);
        if (2 * k >= p.n_pairs) break;
        const bool two = 2 * k + 1 < p.n_pairs;
        int pi[2];
        pi[0] = p.order ? (int)p.order[2 * k] : 2 * k;
        pi[1] = two ? (p.order ? (int)p.order[2 * k + 1] : 2 * k + 1) : pi[0];
        int nd[2], Lt[2], Lr[2], x[2], K[2] = { 0, 0 };
        const unsigned char *rbp[2];
        const uint8_t *Tb[2], *tnext[2];
        const uint32_t *bw[2];
        unsigned emrow[2], tcn[2], sEM[2], sEI[2];
        LeanCoef a;
        {
            const DevPair PA = p.pairs[pi[0]], PB = p.pairs[pi[1]];
            const DevPair *PP[2] = { &PA, &PB };
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const DevPair &P = *PP[h];
                Lt[h] = P.Lt; Lr[h] = P.Lr; nd[h] = P.Lt + P.Lr + 1;
                Tb[h] = p.codes + P.tb_off; bw[h] = p.bits + P.bits_off;
                sEM[h] = (unsigned)__cvta_generic_to_shared(&sh.em[P.model][0]);
                sEI[h] = (unsigned)__cvta_generic_to_shared(&sh.ei[P.model][0]);
                const uint8_t *Rb = p.codes + P.rb_off;
                unsigned char *dst = rbs + (size_t)h * p.smem_rb;
                const int nr = P.Lr + 2 * PADR;
                for (int w = lane; w < nr; w += 32) dst[w] = Rb[w - PADR];
                // slot state of anti-diagonal 0 (lean_seed): slot sigma = lane holds column j = lane
                x[h] = r - lane;
                rbp[h] = dst + PADR - lane;
                emrow[h] = sEM[h] + ((unsigned)Tb[h][lane] << 5);
                tcn[h] = Tb[h][lane + NSLOT];
                tnext[h] = Tb[h] + lane + 2 * NSLOT;
            }
            const float *tA = p.models + PA.model * kModelFloats, *tB = p.models + PB.model * kModelFloats;
            a.mm = mk2(tA[0], tB[0]); a.mi = mk2(tA[1], tB[1]); a.md = mk2(tA[2], tB[2]);
            a.im = mk2(tA[3], tB[3]); a.ii = mk2(tA[4], tB[4]); a.id = mk2(tA[5], tB[5]);
            a.dm = mk2(tA[6], tB[6]); a.di = mk2(tA[7], tB[7]); a.dd = mk2(tA[8], tB[8]);
        }
        if (lane < 8) s_ftot[lane] = 0.f;
        __syncwarp();
        f2 toI = 0ull, inD = 0ull, inMa = 0ull, inMb = 0ull;
        auto mask2 = [&]() -> f2 { return mk2((unsigned)x[0] <= (unsigned)W ? 1.f : 0.f, (unsigned)x[1] <= (unsigned)W ? 1.f : 0.f); };
        f2 msk = mask2();
        const int nd_max = max(nd[0], nd[1]);
        for (int s0 = 0; s0 < nd_max; s0 += 4) {
            unsigned nib[2];
#pragma unroll
            for (int h = 0; h < 2; h++) nib[h] = s0 < nd[h] ? bw[h][s0 >> 5] >> (s0 & 31) : 0xfu;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const int s = s0 + kk;
                const unsigned w0 = rbp[0][kk], w1 = rbp[1][kk];
                const f2 em = mk2(lds_f32(emrow[0] | (w0 & 0x1cu)), lds_f32(emrow[1] | (w1 & 0x1cu)));
                const f2 ei = mk2(lds_f32(sEI[0] + w0), lds_f32(sEI[1] + w1));
                f2 M = mul2(em, inMb);
                const f2 I = mul2(ei, toI);
                const f2 D = inD;
                if (s == 0 && lane == 0) M = mk2(1.f, 1.f); // F_M(0, 0) = 1
                f2 tM = mul2(fma2(a.dm, D, fma2(a.im, I, mul2(a.mm, M))), msk);
                f2 tD = mul2(fma2(a.dd, D, fma2(a.id, I, mul2(a.md, M))), msk);
                f2 nI = mul2(fma2(a.di, D, fma2(a.ii, I, mul2(a.mi, M))), msk);
#pragma unroll
                for (int h = 0; h < 2; h++) { // the end sums F(Lr, Lt - d), d = 0..3
                    if (s >= nd[h] - 4 && s < nd[h]) {
                        const int j = (int)(tnext[h] - Tb[h]) - 2 * NSLOT;
                        if (half2f(msk, h) != 0.f && s - j == Lr[h] && j >= Lt[h] - 3 && (h == 0 || two))
                            s_ftot[4 * h + Lt[h] - j] = half2f(M, h) + half2f(I, h) + half2f(D, h);
                    }
                }
                if (kk == 3 && s0 >= 4) { // rescale decisions, one per pair (as lean_forward: never in the first block, never in the last eight rows)
                    float sc[2] = { 1.f, 1.f };
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (s < nd[h] - 8) {
                            const unsigned mx = __reduce_max_sync(kFull, __float_as_uint(half2f(tM, h)));
                            const int e = (int)(mx >> 23) - 127;
                            if (mx != 0u && e < kScaleLow) {
                                const int kx = min(kScaleTarget - e, kScaleStep);
                                sc[h] = pow2i(kx);
                                K[h] += kx;
                            }
                        }
                    }
                    if (sc[0] != 1.f || sc[1] != 1.f) {
                        const f2 sc2 = mk2(sc[0], sc[1]);
                        tM = mul2(tM, sc2); tD = mul2(tD, sc2); nI = mul2(nI, sc2); inMa = mul2(inMa, sc2);
                    }
                }
                // hand (toM, toD) to the right-hand neighbour column (slot+1 = lane+1, wrapping) of both pairs
                const float rM0 = __shfl_sync(kFull, lo2(tM), (lane + 31) & 31), rM1 = __shfl_sync(kFull, hi2(tM), (lane + 31) & 31);
                const float rD0 = __shfl_sync(kFull, lo2(tD), (lane + 31) & 31), rD1 = __shfl_sync(kFull, hi2(tD), (lane + 31) & 31);
                inMb = inMa; inMa = mk2(rM0, rM1); inD = mk2(rD0, rD1); toI = nI;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (s < nd[h] - 1) {
                        x[h] += (int)((~nib[h] >> kk) & 1u);
                        if (x[h] > W) { // the slot's cell left the band at the top: on to column j + NSLOT
                            x[h] -= NSLOT; rbp[h] -= NSLOT;
                            emrow[h] = sEM[h] + (tcn[h] << 5);
                            tcn[h] = *tnext[h];
                            tnext[h] += NSLOT;
                        }
                    }
                }
                msk = mask2();
            }
#pragma unroll
            for (int h = 0; h < 2; h++)
                if (s0 + 4 < nd[h] + 4) rbp[h] += 4; // a finished pair stops moving
        }
        __syncwarp();
        if (lane < 2 && (lane == 0 || two)) {
            const float fin = s_ftot[4 * lane];
            const int Kl = lane ? K[1] : K[0];
            p.out_lk[lane ? pi[1] : pi[0]] = fin > 0.f ? log((double)fin) - (double)Kl * 0.6931471805599453 : -INFINITY;
        }
        __syncwarp();
    }
}

template <int C, int ROWS>
__device__ __forceinline__ void backward_fused(const PairCtx &pc, const BCoef &a, const LeanSmem &fsh, const int model, const LeanPair &lp,
                                               const KbRef kb, const KParams &p, const unsigned wslot, const unsigned rawk,
                                               unsigned char *wsm, unsigned &phase) {
    // per-warp shared memory: [row buffer][checkpoint landing zone | mbarrier | end sums][staged read codes]
    f2 *buf = reinterpret_cast<f2 *>(wsm);
    const ulonglong2 *buf_q = reinterpret_cast<const ulonglong2 *>(wsm);
    unsigned char *ckstage = wsm + seg_buf_rows<C>() * (C * kPlane) * 8;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(ckstage + ckpt_bytes<C>());
    volatile float *s_ftot = reinterpret_cast<volatile float *>(ckstage + ckpt_bytes<C>() + 16);
    float4 *raw_base = p.raw;
    unsigned ev = 0u; // rescale events of the forward blocks in the buffer (set by the recomputation)
    constexpr int P = C / 2;
    constexpr int RS = C * kPlane, RQ = P * kPlane;
    constexpr int SEG = seg_rows<C>(), CKB = ckpt_bytes<C>();
    constexpr int QSEG = SEG / 4; // backward blocks per segment
    const int lane = threadIdx.x & 31;
    const int nd = pc.nd, W = 2 * pc.r;
    // B(Lr, Lt) = boff puts sum_cells F*B = fin * boff at about 2^kProductExp (read where it is used: the first step only)
    auto boff_of = [&]() -> float {
        const float fin_raw = s_ftot[0];
        const int e_fin = (int)(__float_as_uint(fin_raw) >> 23) - 127;
        return fin_raw > 0.f ? pow2i(max(-120, min(120, kProductExp - e_fin))) : 1.f;
    };

    // ---- forward rows of one segment, recomputed into the buffer ---------------------------------------
    const unsigned ck_s = (unsigned)__cvta_generic_to_shared(ckstage);
    auto prefetch_ckpt = [&](int k) { // forward state before row SEG*k - kSegBelow
        if (k >= 1 && SEG * k - kSegBelow < nd) {
            if (elect_one()) {
                mbar_expect_tx(bar, (unsigned)CKB);
                bulk_g2s(ck_s, reinterpret_cast<const unsigned char *>(p.frows) + (size_t)wslot * p.frow_stride + (size_t)k * CKB, (unsigned)CKB, bar);
            }
        }
    };
    auto load_segment = [&](int k) { // rows SEG*k - 4 .. SEG*k + SEG - 5 -> buffer rows 0 .. SEG-1
        __syncwarp();
        {   // the eight lowest rows (of the segment above) move to the top of the buffer
            const float4 *src = reinterpret_cast<const float4 *>(buf);
            float4 *dst = reinterpret_cast<float4 *>(buf + SEG * RS);
            for (int w = lane; w < kSegKeep * RS / 2; w += 32) dst[w] = src[w];
        }
        __syncwarp();
        const int s0 = SEG * k - kSegBelow;
        int s_from = max(s0, 0);
        int s_to = min(s0 + SEG, nd);
        if (s_from < s_to) {
            LeanFwd<C> st;
            int K = 0, ups = 0;
            if (k >= 1) {
                mbar_wait(bar, phase & 1u);
                phase ^= 1u;
                const float4 *ck = reinterpret_cast<const float4 *>(ckstage);
#pragma unroll
                for (int q = 0; q < P; q++) {
                    const float4 v = ck[(2 * q) * 32 + lane], u = ck[(2 * q + 1) * 32 + lane];
                    st.inMb[q] = mk2(v.x, v.y); st.inMa[q] = mk2(v.z, v.w); st.inD[q] = mk2(u.x, u.y); st.toI[q] = mk2(u.z, u.w);
                }
                ups = reinterpret_cast<const int *>(ckstage + C * 512)[0];
                K = kb[(s0 >> 2) - 1];
            } else {
#pragma unroll
                for (int q = 0; q < P; q++) st.inMb[q] = st.inMa[q] = st.inD[q] = st.toI[q] = 0ull;
            }
            lean_seed<C>(st, lp, s_from, ups);
            const LeanCoef la = load_lean_coef(fsh.cdup[model]); // live during the recomputation only
            lean_forward<C, 1>(la, st, lp, s_from, s_to, K, ups, kb, nullptr, buf + (size_t)(s_from - s0) * RS + 2 * (kPlaneHalo + lane), s_ftot, ev);
        } else {
            s_from = s_to = s0; // nothing to compute: every row of the segment reads as zero
        }
        // rows below anti-diagonal 0 and past the last one read as zero
        for (int w = lane; w < (s_from - s0) * RS; w += 32) buf[w] = 0ull;
        {
            const int z0 = max(s_to - s0, 0);
            for (int w = z0 * RS + lane; w < SEG * RS; w += 32) buf[w] = 0ull;
        }
        __syncwarp();
        prefetch_ckpt(k - 1);
    };

    BwdState<C> st;
    bwd_init<C>(st, pc, lp.rb0);
    // 32-bit offset: no 64-bit pointer lives across the steps
    auto raw_of = [&](int j) -> float4 * { return raw_base + (size_t)(rawk + (unsigned)j * 4u); };
    int k_have = 0; // segment in the buffer; buffer row 0 holds anti-diagonal SEG * k_have - kSegBelow
    auto slow_step = [&](int s) {
        const ulonglong2 *rq = buf_q + (size_t)(s - (SEG * k_have - kSegBelow)) * RQ + kPlaneHalo + lane;
        auto kfat = [&](int t) -> int { return kb[(t - 3) >> 2]; };
        const int kcur = kfat(s);
        const int kstep = kcur - kfat(s - 1);
        float ce[7];
#pragma unroll
        for (int e = -3; e <= 3; e++) ce[e + 3] = pow2i(max(-126, min(126, kcur - kfat(s + e))));
        f2 bM[P], bD[P];
        if (s == nd - 1) bwd_step<C, ROWS, true, true>(pc, a, st, rq, ce, boff_of(), bM, bD, 0);
        else bwd_step<C, ROWS, true, false>(pc, a, st, rq, ce, 0.f, bM, bD, 0);
#pragma unroll
        for (int c = 0; c < C; c++) st.rbp[c] -= 1;
        bwd_hand_off<C>(st, bM, bD);
        if (s > 0) {
            if (kstep != 0) bwd_scale<C>(st, pow2i(kstep));
            const int dec = (int)(((pc.bw[(s - 1) >> 5] >> ((s - 1) & 31)) & 1u) ^ 1u);
            bwd_band_down<C>(st, dec, W);
        }
    };

    const int q_top = (nd + 3) >> 2; // block of the last anti-diagonal
    auto load_nib = [&](int q) -> unsigned { // guide bits 4q-5 .. 4q-2 of block q >= 2
        const int b0 = 4 * q - 5;
        return __funnelshift_r(pc.bw[b0 >> 5], pc.bw[(b0 >> 5) + 1], b0 & 31);
    };
    unsigned nib_cur = 0u, nib_nxt = q_top >= 2 ? load_nib(q_top) : 0u;
    auto preamble = [&](int q) -> int { // 0: generic steps, 1: fast block, 2: fast block with exact corrections
        nib_cur = nib_nxt;
        if (q >= 3) nib_nxt = load_nib(q - 1);
        if (!(4 * q - 1 <= nd - 2 && q >= 2)) return 0;
        // block q touches forward rows 4q-7 .. 4q+2: exact corrections iff the forward pass rescaled at the end of block q-1 or q-2
        return (__funnelshift_r(ev, ev, (q - 2) & 31) & 3u) == 0u ? 1 : 2;
    };
    // segment of block q: blocks QSEG*k + 1 .. QSEG*k + QSEG are the rows SEG*k .. SEG*k + SEG - 1
    k_have = (q_top - 1) / QSEG + 2;
    prefetch_ckpt(k_have - 1);
    for (int q = q_top; q >= 1;) {
        const int kq = (q - 1) / QSEG;
        if (k_have > kq) { // one call site: the two segments at the top first, then one per QSEG blocks
            --k_have;
            load_segment(k_have);
            continue;
        }
        const int mode = preamble(q);
        if (mode == 1) {
            const ulonglong2 *rq = buf_q + (size_t)(4 * q - 1 - (SEG * k_have - kSegBelow)) * RQ + kPlaneHalo + lane;
            const unsigned nib = nib_cur;
#pragma unroll kBwdUnroll
            for (int k = 0; k < 4; k++) {
                f2 bM[P], bD[P];
                bwd_step<C, ROWS, false, false>(pc, a, st, rq - k * RQ, nullptr, 0.f, bM, bD, k);
                bwd_hand_off<C>(st, bM, bD);
                bwd_band_down<C>(st, (int)(((nib >> (3 - k)) & 1u) ^ 1u), W);
            }
#pragma unroll
            for (int c = 0; c < C; c++) st.rbp[c] -= 4;
        } else if (mode == 2) { // a rescale within reach: the same block with the exact corrections of its cut products
            const ulonglong2 *rq = buf_q + (size_t)(4 * q - 1 - (SEG * k_have - kSegBelow)) * RQ + kPlaneHalo + lane;
            const unsigned nib = nib_cur;
            const int k1 = kb[q - 1], k2 = kb[q - 2], k3 = kb[q - 3];
            const float fA = pow2c(k2 - k3), fB = pow2c(k1 - k2), fBi = pow2c(k2 - k1);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                f2 bM[P], bD[P];
                float ce[7];
                block_corrections(k, fA, fB, fBi, ce);
                bwd_step<C, ROWS, true, false>(pc, a, st, rq - k * RQ, ce, 0.f, bM, bD, k);
                bwd_hand_off<C>(st, bM, bD);
                if (k == 0 && k1 != k2) bwd_scale<C>(st, pow2i(k1 - k2)); // mirror of the forward rescale on row 4q-1
                bwd_band_down<C>(st, (int)(((nib >> (3 - k)) & 1u) ^ 1u), W);
            }
#pragma unroll
            for (int c = 0; c < C; c++) st.rbp[c] -= 4;
        } else {
            for (int s = min(4 * q - 1, nd - 1); s >= 4 * q - 4; --s) slow_step(s);
        }
        bwd_retire<C, ROWS>(st, pc, raw_of);
        --q;
    }
#pragma unroll
    for (int c = 0; c < C; c++)
        if (st.j[c] >= 0) bwd_flush_col<C>(st, c, raw_of(st.j[c]));
}

template <int C, int ROWS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, fused_ctas_per_sm(C)) modtable_fused_kernel(KParams p) {
    __shared__ LeanSmem fsh;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    fill_lean_tables(fsh, p.models);
    constexpr int NSLOT = 32 * C, RS = C * kPlane, PADR = NSLOT + 16;
    constexpr int BUFB = seg_buf_rows<C>() * RS * 8, CKS = ckpt_stage_bytes<C>();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wsm = dyn_smem + (size_t)warp * (BUFB + CKS + p.smem_rb);
    unsigned char *ckstage = wsm + BUFB;
    unsigned char *rb_s = wsm + BUFB + CKS; // rb_s[i + PADR] = code byte of read row i
    const unsigned bar = (unsigned)__cvta_generic_to_shared(ckstage + ckpt_bytes<C>());
    unsigned phase = 0u;
    if (lane == 0) mbar_init(bar, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const size_t wslot = (size_t)blockIdx.x * kWarpsPerCta + warp;
    unsigned char *ckpt_g = reinterpret_cast<unsigned char *>(p.frows) + wslot * p.frow_stride; // frow_stride: BYTES per warp slot here
    const KbRef kb = { p.kf, (int)(wslot * p.kf_stride) + 4 };                                 // kb[q], q >= -4
    volatile float *s_ftot = reinterpret_cast<volatile float *>(ckstage + ckpt_bytes<C>() + 16);
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(p.counter, 1);
        k = __shfl_sync(kFull, k, 0);
        if (p.pair_lo + k >= p.pair_hi) break;
        const int pi = pair_index(p, k);
        const DevPair P = p.pairs[pi];
        const int Lt = P.Lt, Lr = P.Lr, nd = Lt + Lr + 1;
        const BCoef a = load_bcoef(fsh.trans[P.model]);
        unsigned *info = p.fwdinfo + (size_t)k * p.fwdinfo_stride;
        {   // stage the read codes of this pair
            const uint8_t *Rb = p.codes + P.rb_off;
            const int nr = Lr + 2 * PADR;
            for (int w = lane; w < nr; w += 32) rb_s[w] = Rb[w - PADR];
            if (lane < 4) { kb[lane - 4] = 0; s_ftot[lane] = 0.f; }
        }
        __syncwarp();
        LeanPair lp;
        lp.Tb = p.codes + P.tb_off; lp.bw = p.bits + P.bits_off; lp.rb0 = rb_s + PADR;
        lp.Lt = Lt; lp.Lr = Lr; lp.nd = nd; lp.r = p.radius;
        lp.sEM = (unsigned)__cvta_generic_to_shared(&fsh.em[P.model][0]);
        lp.sEI = (unsigned)__cvta_generic_to_shared(&fsh.ei[P.model][0]);
        int K = 0;
        {   // ---- pass 1 ----
            LeanFwd<C> st;
#pragma unroll
            for (int q = 0; q < C / 2; q++) st.inMb[q] = st.inMa[q] = st.inD[q] = st.toI[q] = 0ull;
            int ups = 0;
            lean_seed<C>(st, lp, 0, 0);
            const LeanCoef la = load_lean_coef(fsh.cdup[P.model]);
            unsigned ev_unused = 0u;
            lean_forward<C, 0>(la, st, lp, 0, nd, K, ups, kb, ckpt_g, nullptr, s_ftot, ev_unused);
            for (int q = ((nd + 3) >> 2) + lane; q <= ((nd + 3) >> 2) + 2; q += 32) kb[q] = K; // blocks past the last row
        }
        __syncwarp();
        const float fin = s_ftot[0];
        if (lane < 4) info[lane] = __float_as_uint(s_ftot[lane]);
        if (lane == 4) info[4] = (unsigned)K;
        if (lane == 0) p.out_lk[pi] = fin > 0.f ? log((double)fin) - (double)K * 0.6931471805599453 : -INFINITY;
        // the checkpoints were written through the generic proxy: order them before the async-proxy reads
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __syncwarp();
        // ---- pass 2 ----
        const PairCtx pc = make_pair_ctx_bwd(p, P, fsh);
        backward_fused<C, ROWS>(pc, a, fsh, P.model, lp, kb, p, (unsigned)wslot, (unsigned)k * (unsigned)p.raw_stride, wsm, phase);
        __syncwarp();
    }
}

// host-callable launchers ---------------------------------------------------------------------------
int cols_per_lane_for_radius(int radius) {
    // the slot ring must satisfy 2r + 4 <= 32*C
    for (int c = 2; c <= 8; c *= 2)
        if (2 * radius + 4 <= 32 * c) return c;
    return 0;
}

// dynamic shared memory of one CTA of the backward kernel: per warp the ring of forward rows and the staged read codes
template <int C, int ROWS> static int bwd_dyn_c(int smem_rb) { return bwd_warps_per_cta(C) * (ring_floats2<C, ROWS>() * (int)sizeof(f2) + smem_rb); }
template <int C, int ROWS>
static cudaError_t launch_modtable_cr(const KParams &p, int grid_fwd, int grid_bwd, cudaStream_t st) {
    const int dyn = bwd_dyn_c<C, ROWS>(p.smem_rb);
    cudaError_t e = cudaFuncSetAttribute(bwdtable_kernel<C, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    const int dyn_f = kWarpsPerCta * p.smem_rb;
    e = cudaFuncSetAttribute(fwdrows_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_f);
    if (e != cudaSuccess) return e;
    fwdrows_kernel<C><<<grid_fwd, kWarpsPerCta * 32, dyn_f, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    bwdtable_kernel<C, ROWS><<<grid_bwd, bwd_warps_per_cta(C) * 32, dyn, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    finalize_kernel<<<dim3((unsigned)((p.max_lt + kFinCols) / kFinCols), (unsigned)(p.pair_hi - p.pair_lo)), kFinCols, 0, st>>>(p);
    return cudaGetLastError();
}
template <int C>
static cudaError_t launch_modtable_c(const KParams &p, int rows, int grid_fwd, int grid_bwd, cudaStream_t st) {
    return rows == 14 ? launch_modtable_cr<C, 14>(p, grid_fwd, grid_bwd, st) : launch_modtable_cr<C, 9>(p, grid_fwd, grid_bwd, st);
}

template <int C, int ROWS>
static cudaError_t launch_fused_cr(const KParams &p, int grid, int dyn, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(modtable_fused_kernel<C, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    modtable_fused_kernel<C, ROWS><<<grid, kWarpsPerCta * 32, dyn, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    finalize_kernel<<<dim3((unsigned)((p.max_lt + kFinCols) / kFinCols), (unsigned)(p.pair_hi - p.pair_lo)), kFinCols, 0, st>>>(p);
    return cudaGetLastError();
}
template <int C> static int fused_dyn_c(int smem_rb) { return kWarpsPerCta * (seg_buf_rows<C>() * C * kPlane * 8 + ckpt_stage_bytes<C>() + smem_rb); }
// dynamic shared memory of one CTA of the fused kernel: per warp the row buffer, the checkpoint landing zone and the staged read codes
int fused_dyn_smem(int C, int smem_rb) { return C == 2 ? fused_dyn_c<2>(smem_rb) : (C == 4 ? fused_dyn_c<4>(smem_rb) : fused_dyn_c<8>(smem_rb)); }
// resident CTAs of the fused kernel on sm_count SMs (registers and shared memory decide; at most fused_ctas_per_sm(C) per SM)
int fused_grid(int C, int rows, int dyn, int sm_count) {
    int n = 0;
    auto occ = [&](auto kern) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess) { n = 0; return; }
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kWarpsPerCta * 32, (size_t)dyn) != cudaSuccess) n = 0;
    };
    switch (C) {
    case 2: if (rows == 14) occ(modtable_fused_kernel<2, 14>); else occ(modtable_fused_kernel<2, 9>); break;
    case 4: if (rows == 14) occ(modtable_fused_kernel<4, 14>); else occ(modtable_fused_kernel<4, 9>); break;
    case 8: if (rows == 14) occ(modtable_fused_kernel<8, 14>); else occ(modtable_fused_kernel<8, 9>); break;
    default: break;
    }
    return std::min(n, fused_ctas_per_sm(C)) * sm_count;
}
// bytes of checkpoints / ints of block exponents one warp slot needs for pairs of up to max_nd anti-diagonals
size_t fused_ckpt_bytes(int C, int max_nd) {
    const size_t seg = C == 2 ? seg_rows<2>() : seg_rows<4>(), ckb = C == 2 ? ckpt_bytes<2>() : (C == 4 ? ckpt_bytes<4>() : ckpt_bytes<8>());
    return (((size_t)max_nd + kSegBelow) / seg + 2) * ckb;
}
size_t fused_kb_ints(int max_nd) { return (size_t)((max_nd + 3) >> 2) + 12; }

// the fused modification table (v10): pairs [p.pair_lo, p.pair_hi), p.frows / p.frow_stride = checkpoint scratch (bytes per warp slot)
cudaError_t launch_modtable_fused(const KParams &p, int C, int grid, cudaStream_t st) {
    const int dyn = fused_dyn_smem(C, p.smem_rb);
    switch (C) {
    case 2: return p.rows == 14 ? launch_fused_cr<2, 14>(p, grid, dyn, st) : launch_fused_cr<2, 9>(p, grid, dyn, st);
    case 4: return p.rows == 14 ? launch_fused_cr<4, 14>(p, grid, dyn, st) : launch_fused_cr<4, 9>(p, grid, dyn, st);
    case 8: return p.rows == 14 ? launch_fused_cr<8, 14>(p, grid, dyn, st) : launch_fused_cr<8, 9>(p, grid, dyn, st);
    default: return cudaErrorInvalidValue;
    }
}

// one wave of the modification table: forward kernel, then backward kernel, pairs [p.pair_lo, p.pair_hi)
cudaError_t launch_modtable(const KParams &p, int C, int grid_fwd, int grid_bwd, cudaStream_t st) {
    switch (C) {
    case 2: return launch_modtable_c<2>(p, p.rows, grid_fwd, grid_bwd, st);
    case 4: return launch_modtable_c<4>(p, p.rows, grid_fwd, grid_bwd, st);
    case 8: return launch_modtable_c<8>(p, p.rows, grid_fwd, grid_bwd, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_likelihood(const KParams &p, int C, int grid, cudaStream_t st) {
    switch (C) {
    case 2: likelihood_kernel<2><<<grid, kWarpsPerCta * 32, 0, st>>>(p); break;
    case 4: likelihood_kernel<4><<<grid, kWarpsPerCta * 32, 0, st>>>(p); break;
    case 8: likelihood_kernel<8><<<grid, kWarpsPerCta * 32, 0, st>>>(p); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// two short pairs per warp (radius <= likelihood_pairs2_max_radius()); p.smem_rb = staged read codes per PAIR
int likelihood_pairs2_max_radius() { return kPackRadiusMax; }
int likelihood_pairs2_pad_rows() { return 32 + 16; }
cudaError_t launch_likelihood_pairs2(const KParams &p, int grid, cudaStream_t st) {
    const int dyn = kWarpsPerCta * 2 * p.smem_rb;
    cudaError_t e = cudaFuncSetAttribute(likelihood_pairs2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    likelihood_pairs2_kernel<<<grid, kWarpsPerCta * 32, dyn, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fit(const KParams &p, int C, int grid, double *acc90, cudaStream_t st) {
    const int dyn = (int)sizeof(FitSmem);
    cudaError_t e = cudaSuccess;
    switch (C) {
    case 2:
        e = cudaFuncSetAttribute(fit_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        if (e == cudaSuccess) fit_kernel<2><<<grid, kWarpsPerCta * 32, dyn, st>>>(p, acc90);
        break;
    case 4:
        e = cudaFuncSetAttribute(fit_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        if (e == cudaSuccess) fit_kernel<4><<<grid, kWarpsPerCta * 32, dyn, st>>>(p, acc90);
        break;
    case 8:
        e = cudaFuncSetAttribute(fit_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        if (e == cudaSuccess) fit_kernel<8><<<grid, kWarpsPerCta * 32, dyn, st>>>(p, acc90);
        break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

int warps_per_cta() { return kWarpsPerCta; }
int frow_slots_per_row(int C) { return C * kPlane; }
int frow_extra_rows() { return kRowShift + kRowsAbove; }
int modtable_ctas_per_sm(int C, int rows) { return bwd_ctas_per_sm(C, rows); }
int modtable_warps_per_cta(int C) { return bwd_warps_per_cta(C); }
// dynamic shared memory of one CTA of the backward kernel of the rows variant
int modtable_dyn_smem(int C, int smem_rb) { return C == 2 ? bwd_dyn_c<2, 14>(smem_rb) : (C == 4 ? bwd_dyn_c<4, 14>(smem_rb) : bwd_dyn_c<8, 14>(smem_rb)); }
int fwdrows_ctas_per_sm(int C) { return fwd_ctas_per_sm(C); }
int fwdinfo_words() { return 16; }
int fwd_pad_rows(int C) { return 32 * C + 16; }

// ---- FP32 peak micro-benchmark (roofline denominator) ---------------------------------------------------
// 16 independent accumulators per thread so the 4-cycle FMA latency is covered at 8 warps per scheduler.
template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(int iters, float *sink) {
    float a = 1.0f + 1e-7f * threadIdx.x, b = 1e-9f * (blockIdx.x + 1);
    if (MODE == 0) {
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; k++) acc[k] = (float)k;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; k++) acc[k] = fmaf(acc[k], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k++) s += acc[k];
        sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        unsigned long long acc[8], aa, bb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
        for (int k = 0; k < 8; k++) { float lo = (float)k, hi = (float)(k + 8); asm("mov.b64 %0, {%1, %2};" : "=l"(acc[k]) : "f"(lo), "f"(hi)); }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(aa), "l"(bb));
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k])); s += lo + hi; }
        sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

cudaError_t launch_fp32_peak(int mode, int blocks, int threads, int iters, float *sink, cudaStream_t st) {
    if (mode == 0) fp32_peak_kernel<0><<<blocks, threads, 0, st>>>(iters, sink);
    else fp32_peak_kernel<1><<<blocks, threads, 0, st>>>(iters, sink);
    return cudaGetLastError();
}

} // namespace jtk
