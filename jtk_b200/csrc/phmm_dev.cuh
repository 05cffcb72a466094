// phmm_dev.cuh -- device-side data layout shared by the kernels and the C-ABI host code.
//
// One warp scores one (read, template) pair.  Lanes own template COLUMNS ("j-owner" layout): lane l holds
// C adjacent column slots sigma = l*C + c, and slot sigma holds the one column j == sigma (mod NSLOT=32*C)
// that is currently inside the band window.  On anti-diagonal s the cell of column j is (i, j), i = s - j.
// Per-column sums of the modification table therefore never leave their lane; only the D- and M-moves
// cross a lane boundary (2 shuffles per lane and step).  See DESIGN.md section 3.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace jtk {

constexpr int kNumRow = 14;
constexpr int kCodePad = 544;        // sentinel bytes on both sides of every code array (kernels reach at most 2*NSLOT+16 = 528 past an end, NSLOT <= 256)
constexpr int kStageCols = 64;       // per-warp staging ring (columns) for finished column sums
constexpr int kStageStride = 16;     // floats per staged column
// per-model float block: [0..8] transitions (HMMParam order), [12..75] eM[tc*8+qc], [76..139] eI[ctx*8+qc],
// [140..171] eMT[qc*4+b] = eM(ref b, query qc).  Codes 4..7 (no base / padding) hold zeros.
constexpr int kOffEM = 12, kOffEI = 76, kOffEMT = 140;
constexpr int kModelFloats = 172;
constexpr float kDeltaNeg = -1.0e10f;

// One pair as the kernels see it.  All offsets are into the batch-wide device arrays.
struct DevPair {
    uint32_t tb_off;   // codes[tb_off + j]   = code of t[j-1] (4 = none), j in -pad .. Lt+pad
    uint32_t rb_off;   // codes[rb_off + i]   = (ctx<<3 | q[i-1]) of read row i (qc 4 = none)
    uint32_t bits_off; // bits[bits_off + (s>>5)] bit (s&31) = centre(s+1) - centre(s)
    int32_t Lt, Lr;
    int32_t model;     // 0: forward-strand model, 1: reverse-strand model
    uint32_t pad_;
    uint64_t tab_off;  // float offset of this pair's (Lt+1)*14 delta table
};

struct KParams {
    const DevPair *pairs;
    int n_pairs;
    const uint8_t *codes;
    const uint32_t *bits;
    const float *models;    // 2 * kModelFloats
    int radius;
    int rows;               // 14: all rows, 9: clustering rows only (copy / multi-base deletion skipped)
    // per-warp scratch (forward rows + per-row scale exponents)
    float2 *frows;          // [n_warp_slots][frow_stride] float2
    size_t frow_stride;     // (max_nd + 6) * NSLOT
    int32_t *kf;            // [n_warp_slots][kf_stride]
    size_t kf_stride;       // max_nd + 6
    float *out_delta;       // table - lk, fp32, kDeltaNeg for impossible edits
    double *out_lk;
    int *counter;           // work queue head (forward kernel / single-kernel paths)
    // modification table = two kernels (forward rows, then backward + table): the scratch is per PAIR of a wave
    int pair_lo, pair_hi;   // pairs [pair_lo, pair_hi) of this wave; scratch slot = pair index - pair_lo
    unsigned *fwdinfo;      // per pair slot: [0..3] ftot (float bits), [4] Ktot, [8..] rescale-event map
    size_t fwdinfo_stride;  // words per pair slot
    int *counter2;          // work queue head of the backward kernel
    float4 *raw;            // per pair slot: 16 raw column sums per template column (backward -> finalize kernel)
    size_t raw_stride;      // float4 per pair slot = 4 * (max_lt + 1)
    int max_lt;
    int smem_rb, smem_tb;   // bytes of staged read / template codes per warp in the forward kernel (multiples of 16)
    const uint32_t *order;  // queue position -> pair index (longest pairs first), or null: pairs in batch order
};

} // namespace jtk
