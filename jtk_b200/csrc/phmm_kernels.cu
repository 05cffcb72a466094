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
constexpr int kScaleLow = 88;
constexpr int kScaleTarget = 100;
constexpr int kScaleStep = 64;
constexpr int kProductExp = 0;

struct Trans { float mm, mi, md, im, ii, id, dm, di, dd; };

__device__ __forceinline__ Trans load_trans(const float *m) {
    Trans a;
    a.mm = m[0]; a.mi = m[1]; a.md = m[2];
    a.im = m[3]; a.ii = m[4]; a.id = m[5];
    a.dm = m[6]; a.di = m[7]; a.dd = m[8];
    return a;
}

__device__ __forceinline__ float pow2i(int k) { // exact 2^k, k in [-126, 127]
    return __uint_as_float((unsigned)(127 + k) << 23);
}

__device__ __forceinline__ int band_lo(int cen, int r, int s, int Lt) { return max(max(cen - r, 0), s - Lt); }
__device__ __forceinline__ int band_hi(int cen, int r, int s, int Lr) { return min(min(cen + r, Lr), s); }

// ------------------------------------------------------------------------------------------------
// Forward pass.  STORE: write (toM, toD) of every cell to frow[s*NSLOT + slot] and the cumulative scale
// exponent to kf[s].  Returns the stored final value fin = (F_M+F_I+F_D)(Lr,Lt) * 2^Ktot.
// s_ftot[d] (d = 0..3) receives (F_M+F_I+F_D)(Lr, Lt-d) * 2^Ktot (the delete-to-end rows).
// ------------------------------------------------------------------------------------------------
template <int C, bool STORE>
__device__ __forceinline__ void forward_pass(const DevPair &P, const uint8_t *__restrict__ codes,
                                             const uint32_t *__restrict__ bits, const float *sm, int r,
                                             float2 *__restrict__ frow, int32_t *__restrict__ kf,
                                             volatile float *s_ftot, int &Ktot) {
    constexpr int NSLOT = 32 * C;
    const int lane = threadIdx.x & 31;
    const uint8_t *Tb = codes + P.tb_off;
    const uint8_t *Rb = codes + P.rb_off;
    const uint32_t *bw = bits + P.bits_off;
    const int Lt = P.Lt, Lr = P.Lr, nd = Lt + Lr + 1;
    const float *sEM = sm + kOffEM;
    const float *sEI = sm + kOffEI;
    const Trans a = load_trans(sm);

    int j[C], tc8[C];
    float toI[C], inD[C], inMa[C], inMb[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        j[c] = lane * C + c;
        tc8[c] = (int)Tb[j[c]] << 3;
        toI[c] = inD[c] = inMa[c] = inMb[c] = 0.f;
    }
    if (lane < 4) s_ftot[lane] = 0.f;
    int cen = 0, K = 0;
    for (int s = 0; s < nd; ++s) {
        const int lo = band_lo(cen, r, s, Lt), hi = band_hi(cen, r, s, Lr);
        const int W = hi - lo;
        const bool tail = s >= nd - 4;
        float tM[C], tD[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int i = s - j[c];
            const int x = i - lo;
            const bool valid = (unsigned)x <= (unsigned)W;
            const int rb = Rb[i];
            const float em = sEM[tc8[c] | (rb & 7)];
            const float ei = sEI[rb & 63];
            float M = em * inMb[c], I = ei * toI[c], D = inD[c];
            if (s == 0 && j[c] == 0) M = 1.f;
            if (!valid) { M = 0.f; I = 0.f; D = 0.f; }
            tM[c] = a.mm * M + a.im * I + a.dm * D;
            toI[c] = a.mi * M + a.ii * I + a.di * D;
            tD[c] = a.md * M + a.id * I + a.dd * D;
            if (tail && valid && i == Lr && j[c] >= Lt - 3) s_ftot[Lt - j[c]] = M + I + D;
            if (x > W) { // below the band for good: the slot moves on to column j + NSLOT
                j[c] += NSLOT;
                tc8[c] = (int)Tb[j[c]] << 3;
            }
        }
        if (((s & (kRescaleEvery - 1)) == kRescaleEvery - 1) && !tail && s < nd - 8) {
            float v = tM[0];
#pragma unroll
            for (int c = 1; c < C; c++) v = fmaxf(v, tM[c]);
            const unsigned m = __reduce_max_sync(kFull, __float_as_uint(v));
            const int e = (int)(m >> 23) - 127;
            if (m != 0u && e < kScaleLow) {
                const int k = min(kScaleTarget - e, kScaleStep);
                const float sc = pow2i(k);
#pragma unroll
                for (int c = 0; c < C; c++) { tM[c] *= sc; tD[c] *= sc; toI[c] *= sc; inMa[c] *= sc; }
                K += k;
            }
        }
        if (STORE) {
            if (lane == 0) kf[s] = K;
            float2 *row = frow + (size_t)s * NSLOT + lane * C;
            if constexpr (C == 2) {
                *reinterpret_cast<float4 *>(row) = make_float4(tM[0], tD[0], tM[1], tD[1]);
            } else {
#pragma unroll
                for (int c = 0; c < C; c++) row[c] = make_float2(tM[c], tD[c]);
            }
        }
        // hand toM / toD to the right-hand neighbour column (slot+1, wrapping)
        const float rM = __shfl_sync(kFull, tM[C - 1], (lane + 31) & 31);
        const float rD = __shfl_sync(kFull, tD[C - 1], (lane + 31) & 31);
#pragma unroll
        for (int c = C - 1; c >= 1; c--) { inMb[c] = inMa[c]; inMa[c] = tM[c - 1]; inD[c] = tD[c - 1]; }
        inMb[0] = inMa[0]; inMa[0] = rM; inD[0] = rD;
        if (s < nd - 1) cen += (bw[s >> 5] >> (s & 31)) & 1u;
    }
    Ktot = K;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Backward pass fused with the modification-table reduction.  ROWS = 14 (all rows) or 9 (rows 0-7 and
// the one-base deletion, the only rows local_clustering reads: pseudo_mcmc.rs:447).
// ------------------------------------------------------------------------------------------------
template <int C, int ROWS>
__device__ __forceinline__ void backward_pass(const DevPair &P, const uint8_t *__restrict__ codes,
                                              const uint32_t *__restrict__ bits, const float *sm, int r,
                                              const float2 *__restrict__ frow, const int32_t *__restrict__ kf,
                                              float *stage, volatile float *s_ftot, float *__restrict__ out) {
    constexpr int NSLOT = 32 * C;
    constexpr int NXM = (ROWS == 14) ? 3 : 1; // deletion lengths accumulated
    constexpr int NXP = (ROWS == 14) ? 3 : 0; // copy lengths accumulated
    const int lane = threadIdx.x & 31;
    const uint8_t *Tb = codes + P.tb_off;
    const uint8_t *Rb = codes + P.rb_off;
    const uint32_t *bw = bits + P.bits_off;
    const int Lt = P.Lt, Lr = P.Lr, nd = Lt + Lr + 1;
    const float *sEM = sm + kOffEM;
    const float *sEI = sm + kOffEI;
    const float4 *sEMT = reinterpret_cast<const float4 *>(sm + kOffEMT);
    const Trans a = load_trans(sm);
    // B(Lr,Lt) = boff puts sum_cells F*B = fin * boff at about 2^kProductExp
    const float fin_raw = s_ftot[0];
    const int e_fin = (int)(__float_as_uint(fin_raw) >> 23) - 127;
    const float boff = fin_raw > 0.f ? pow2i(max(-120, min(120, kProductExp - e_fin))) : 1.f;
    const float fin = fin_raw * boff;

    int j[C], tc8[C];
    float BI[C], BMo[C], inD[C], inMa[C], inMb[C];
    float S[C][4], N[C][4], Vs[C], Vn[C], Xp[C][3], Xm[C][3];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int sigma = lane * C + c;
        j[c] = Lt - ((Lt - sigma) & (NSLOT - 1)); // largest column <= Lt owned by this slot
        tc8[c] = (int)Tb[j[c] + 1] << 3;           // code of t[j]
        BI[c] = BMo[c] = inD[c] = inMa[c] = inMb[c] = 0.f;
        Vs[c] = Vn[c] = 0.f;
#pragma unroll
        for (int b = 0; b < 4; b++) { S[c][b] = 0.f; N[c][b] = 0.f; }
#pragma unroll
        for (int e = 0; e < 3; e++) { Xp[c][e] = 0.f; Xm[c][e] = 0.f; }
    }
    int cen = Lr;
    int blk_lo = (Lt >> 5) << 5;

    auto flush_col = [&](int c) {
        float4 *st = reinterpret_cast<float4 *>(stage + (j[c] & (kStageCols - 1)) * kStageStride);
        st[0] = make_float4(S[c][0], S[c][1], S[c][2], S[c][3]);
        st[1] = make_float4(Vs[c], N[c][0], N[c][1], N[c][2]);
        st[2] = make_float4(N[c][3], Vn[c], Xp[c][0], Xp[c][1]);
        st[3] = make_float4(Xp[c][2], Xm[c][0], Xm[c][1], Xm[c][2]);
    };
    auto dlog = [](float num, float ref) -> float {
        return (num > 0.f && ref > 0.f) ? logf(num / ref) : kDeltaNeg;
    };
    auto emit_block = [&](int lo_col) {
        __syncwarp();
        const int jj = lo_col + lane;
        if (jj >= 0 && jj <= Lt) {
            const float4 *st = reinterpret_cast<const float4 *>(stage + (jj & (kStageCols - 1)) * kStageStride);
            const float4 q0 = st[0], q1 = st[1], q2 = st[2], q3 = st[3];
            const float s4[4] = { q0.x, q0.y, q0.z, q0.w };
            const float n4[4] = { q1.y, q1.z, q1.w, q2.x };
            const float vs = q1.x, vn = q2.y;
            const float xp[3] = { q2.z, q2.w, q3.x };
            const int tcode = Tb[jj + 1];
            float ref = fin;
            if (jj < Lt) ref = s4[tcode & 3] + vs;
            float *o = out + (size_t)jj * kNumRow;
#pragma unroll
            for (int b = 0; b < 4; b++) o[b] = (jj < Lt) ? dlog(s4[b] + vs, ref) : kDeltaNeg;
#pragma unroll
            for (int b = 0; b < 4; b++) o[4 + b] = dlog(n4[b] + vn, ref);
#pragma unroll
            for (int e = 1; e <= 3; e++) {
                o[7 + e] = (ROWS == 14 && jj + e <= Lt) ? dlog(xp[e - 1], ref) : kDeltaNeg;
                float v = kDeltaNeg;
                if ((ROWS == 14 || e == 1) && jj + e <= Lt) {
                    float acc = stage[((jj + e) & (kStageCols - 1)) * kStageStride + 13 + (e - 1)];
                    if (jj + e == Lt) acc += s_ftot[e] * boff;
                    v = dlog(acc, ref);
                }
                o[10 + e] = v;
            }
        }
        __syncwarp();
    };

    for (int s = nd - 1; s >= 0; --s) {
        const int lo = band_lo(cen, r, s, Lt), hi = band_hi(cen, r, s, Lr);
        const int W = hi - lo;
        // forward x backward products pair rows s-3 .. s+3: exponents differ only next to a rescale
        const bool corr = (ROWS == 14) ? (kf[s + 3] != kf[s - 3]) : (kf[s] != kf[s - 1]);
        float ce[7];
        if (corr) {
            const int k0 = kf[s];
#pragma unroll
            for (int e = -3; e <= 3; e++) ce[e + 3] = pow2i(max(-126, min(126, k0 - kf[s + e])));
        }
        float bM[C], bD[C];
        bool any_dead = false;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int i = s - j[c];
            const int x = i - lo;
            const bool valid = (unsigned)x <= (unsigned)W;
            const int rb = Rb[i + 1]; // q[i]
            const float em = sEM[tc8[c] | (rb & 7)];
            const float ei = sEI[rb & 63];
            const float gM = em * inMb[c], gI = ei * BI[c], gD = inD[c];
            float m_ = a.mm * gM + a.mi * gI + a.md * gD;
            float i_ = a.im * gM + a.ii * gI + a.id * gD;
            float d_ = a.dm * gM + a.di * gI + a.dd * gD;
            if (s == nd - 1) { m_ = boff; i_ = boff; d_ = boff; }
            if (!valid) { m_ = 0.f; i_ = 0.f; d_ = 0.f; }
            // ---- table reduction for cell (i, j) ----
            const int slot = lane * C + c;
            const float2 F0 = frow[(ptrdiff_t)s * NSLOT + slot];
            const float4 ec = sEMT[rb & 7];
            const float U = F0.x * inMb[c];
            S[c][0] += U * ec.x; S[c][1] += U * ec.y; S[c][2] += U * ec.z; S[c][3] += U * ec.w;
            Vs[c] += F0.y * inD[c];
            const float U2 = F0.x * BMo[c];
            N[c][0] += U2 * ec.x; N[c][1] += U2 * ec.y; N[c][2] += U2 * ec.z; N[c][3] += U2 * ec.w;
            Vn[c] += F0.y * d_;
            // cuts pairing this column's backward terms with forward columns j-1..j-3 (deletions) and
            // j+1..j+3 (copies).  The masks keep a slot from aliasing the column NSLOT away.
            const bool okm = (unsigned)(x + 1) <= (unsigned)(W + 1);
            const float gMm = okm ? gM : 0.f, gDm = okm ? gD : 0.f;
#pragma unroll
            for (int e = 1; e <= NXM; e++) {
                const float2 Fe = frow[(ptrdiff_t)(s - e) * NSLOT + ((slot - e) & (NSLOT - 1))];
                // next to a rescale the two rows carry different exponents: the exact correction goes on the product
                if (corr) Xm[c][e - 1] += (Fe.x * gMm + Fe.y * gDm) * ce[3 - e];
                else Xm[c][e - 1] += Fe.x * gMm + Fe.y * gDm;
            }
            if (NXP > 0) {
                const bool okp = (unsigned)x <= (unsigned)(W + 1);
                const float gMp = okp ? gM : 0.f, gDp = okp ? gD : 0.f;
#pragma unroll
                for (int e = 1; e <= NXP; e++) {
                    const float2 Fe = frow[(ptrdiff_t)(s + e) * NSLOT + ((slot + e) & (NSLOT - 1))];
                    if (corr) Xp[c][e - 1] += (Fe.x * gMp + Fe.y * gDp) * ce[3 + e];
                    else Xp[c][e - 1] += Fe.x * gMp + Fe.y * gDp;
                }
            }
            bM[c] = m_; bD[c] = d_;
            BI[c] = i_; BMo[c] = m_;
            any_dead |= (x < 0);
        }
        if (__any_sync(kFull, any_dead)) {
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int x = (s - j[c]) - lo;
                if (x < 0) { // above the band for good: stage the finished sums, move to column j - NSLOT
                    if (j[c] >= 0 && j[c] <= Lt) flush_col(c);
                    j[c] -= NSLOT;
                    tc8[c] = (int)Tb[j[c] + 1] << 3;
                    Vs[c] = Vn[c] = 0.f;
#pragma unroll
                    for (int b = 0; b < 4; b++) { S[c][b] = 0.f; N[c][b] = 0.f; }
#pragma unroll
                    for (int e = 0; e < 3; e++) { Xp[c][e] = 0.f; Xm[c][e] = 0.f; }
                }
            }
            const int jhi = s - lo;
            while (blk_lo > jhi) { emit_block(blk_lo); blk_lo -= 32; }
        }
        // hand B_M / B_D to the left-hand neighbour column (slot-1, wrapping)
        const float rM = __shfl_sync(kFull, bM[0], (lane + 1) & 31);
        const float rD = __shfl_sync(kFull, bD[0], (lane + 1) & 31);
#pragma unroll
        for (int c = 0; c < C - 1; c++) { inMb[c] = inMa[c]; inMa[c] = bM[c + 1]; inD[c] = bD[c + 1]; }
        inMb[C - 1] = inMa[C - 1]; inMa[C - 1] = rM; inD[C - 1] = rD;
        if (s > 0) {
            const int k = kf[s] - kf[s - 1]; // mirror of the forward rescale at step s
            if (k != 0) {
                const float sc = pow2i(k);
#pragma unroll
                for (int c = 0; c < C; c++) { BI[c] *= sc; BMo[c] *= sc; inD[c] *= sc; inMa[c] *= sc; inMb[c] *= sc; }
            }
            cen -= (bw[(s - 1) >> 5] >> ((s - 1) & 31)) & 1u;
        }
    }
    // columns still alive after s = 0 (column 0), then the remaining blocks
#pragma unroll
    for (int c = 0; c < C; c++)
        if (j[c] >= 0 && j[c] <= Lt) flush_col(c);
    while (blk_lo >= 0) { emit_block(blk_lo); blk_lo -= 32; }
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
struct __align__(16) SmemLayout {
    float models[2 * kModelFloats];
    float stage[kWarpsPerCta][kStageCols * kStageStride];
    float ftot[kWarpsPerCta][4];
};

template <int C, int ROWS>
__global__ void __launch_bounds__(kWarpsPerCta * 32) modtable_kernel(KParams p) {
    __shared__ SmemLayout sh;
    for (int k = threadIdx.x; k < 2 * kModelFloats; k += blockDim.x) sh.models[k] = p.models[k];
    __syncthreads();
    constexpr int NSLOT = 32 * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wslot = (size_t)blockIdx.x * kWarpsPerCta + warp;
    float2 *frow = p.frows + wslot * p.frow_stride + 3 * NSLOT; // row 0 (3 zero rows below)
    int32_t *kf = p.kf + wslot * p.kf_stride + 3;
    // rows -3..-1 (before the first anti-diagonal) read as zero, exponent 0
    for (int k = lane; k < 3 * NSLOT; k += 32) frow[k - 3 * NSLOT] = make_float2(0.f, 0.f);
    if (lane < 3) kf[lane - 3] = 0;
    __syncwarp();
    for (;;) {
        int pi = 0;
        if (lane == 0) pi = atomicAdd(p.counter, 1);
        pi = __shfl_sync(kFull, pi, 0);
        if (pi >= p.n_pairs) break;
        const DevPair P = p.pairs[pi];
        const int nd = P.Lt + P.Lr + 1;
        const float *sm = sh.models + P.model * kModelFloats;
        int Ktot;
        forward_pass<C, true>(P, p.codes, p.bits, sm, p.radius, frow, kf, sh.ftot[warp], Ktot);
        // rows / exponents just past the last anti-diagonal read as zero / Ktot
        for (int k = lane; k < 3 * NSLOT; k += 32) frow[(size_t)nd * NSLOT + k] = make_float2(0.f, 0.f);
        if (lane < 3) kf[nd + lane] = Ktot;
        __syncwarp();
        const float fin = sh.ftot[warp][0];
        if (lane == 0)
            p.out_lk[pi] = fin > 0.f ? log((double)fin) - (double)Ktot * 0.6931471805599453 : -INFINITY;
        backward_pass<C, ROWS>(P, p.codes, p.bits, sm, p.radius, frow, kf, sh.stage[warp], sh.ftot[warp],
                               p.out_delta + P.tab_off);
        __syncwarp();
    }
}

template <int C>
__global__ void __launch_bounds__(kWarpsPerCta * 32) likelihood_kernel(KParams p) {
    __shared__ SmemLayout sh;
    for (int k = threadIdx.x; k < 2 * kModelFloats; k += blockDim.x) sh.models[k] = p.models[k];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (;;) {
        int pi = 0;
        if (lane == 0) pi = atomicAdd(p.counter, 1);
        pi = __shfl_sync(kFull, pi, 0);
        if (pi >= p.n_pairs) break;
        const DevPair P = p.pairs[pi];
        const float *sm = sh.models + P.model * kModelFloats;
        int Ktot;
        forward_pass<C, false>(P, p.codes, p.bits, sm, p.radius, nullptr, nullptr, sh.ftot[warp], Ktot);
        const float fin = sh.ftot[warp][0];
        if (lane == 0)
            p.out_lk[pi] = fin > 0.f ? log((double)fin) - (double)Ktot * 0.6931471805599453 : -INFINITY;
        __syncwarp();
    }
}

// host-callable launchers ---------------------------------------------------------------------------
int cols_per_lane_for_radius(int radius) {
    // the slot ring must satisfy 2r + 4 <= 32*C
    for (int c = 1; c <= 8; c *= 2)
        if (2 * radius + 4 <= 32 * c) return c;
    return 0;
}

template <int C>
static cudaError_t launch_modtable_c(const KParams &p, int rows, int grid, cudaStream_t st) {
    if (rows == 14) modtable_kernel<C, 14><<<grid, kWarpsPerCta * 32, 0, st>>>(p);
    else modtable_kernel<C, 9><<<grid, kWarpsPerCta * 32, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_modtable(const KParams &p, int C, int grid, cudaStream_t st) {
    switch (C) {
    case 1: return launch_modtable_c<1>(p, p.rows, grid, st);
    case 2: return launch_modtable_c<2>(p, p.rows, grid, st);
    case 4: return launch_modtable_c<4>(p, p.rows, grid, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_likelihood(const KParams &p, int C, int grid, cudaStream_t st) {
    switch (C) {
    case 1: likelihood_kernel<1><<<grid, kWarpsPerCta * 32, 0, st>>>(p); break;
    case 2: likelihood_kernel<2><<<grid, kWarpsPerCta * 32, 0, st>>>(p); break;
    case 4: likelihood_kernel<4><<<grid, kWarpsPerCta * 32, 0, st>>>(p); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int warps_per_cta() { return kWarpsPerCta; }

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
