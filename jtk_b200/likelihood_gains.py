"""Mirror of haplotyper::likelihood_gains (reference: haplotyper/src/likelihood_gains.rs).

`estimate_gain` / `estimate_gain_default` (likelihood_gains.rs:162-192,253-315) calibrate, per (DiffType, homopolymer
length), the expected log-likelihood gain of a true variant and the probability that a read from the null model shows
it; `estimate_minimum_gain` (:6-39) is the 10^6-pair variant used by phmm_likelihood_correction.rs:118.  Every
`likelihood_antidiagonal_bootstrap` call of a calibration (1.8e5 for the default one) goes to the GPU as ONE
`jtk_hmm_likelihood_batch` (bootstrap guide computed inside the library on the host threads).

What cannot be reproduced bit for bit: the reference draws its templates and reads with `kiley::gen_seq::generate_seq`
and `Generate::gen` from Xoshiro256** streams; kiley is not under /root/reference, so the order in
which those functions consume the generator is unknown.  The sampler below follows the model's own definition (start in
Match, row-normalised transitions, emissions as in the oracle) with numpy's PCG64 seeded by the reference's seeds; the
statistics (median, 10th / 67th percentile positions, floors) are the reference's.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .hmm import PairHiddenMarkovModel, PairHiddenMarkovModelOnStrands, default_context
from .local_clustering import Gains

SUBST, DEL, INS = 0, 1, 2  # likelihood_gains.rs:195-199 (row order of Gains)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

SEED, SEQ_LEN, BAND, HOMOP_LEN = 309423, 100, 10, 3  # likelihood_gains.rs:186-189


def generate_seq(rng: np.random.Generator, length: int) -> np.ndarray:
    return ACGT[rng.integers(0, 4, size=length)]


def sample_triple(rng) -> Tuple[int, int, int]:
    """(right, homop, left): three distinct bases (likelihood_gains.rs:213-222)."""
    homop = int(rng.integers(0, 4))
    right = int(rng.choice([b for b in range(4) if b != homop]))
    left = int(rng.choice([b for b in range(4) if b != homop and b != right]))
    return right, homop, left


def gen_diff_haplotypes(rng, length: int, diff_type: int) -> Tuple[np.ndarray, np.ndarray]:
    """likelihood_gains.rs:224-251: [right, homop x len, left] and the same with one edit."""
    right, center, left = sample_triple(rng)
    c1 = [center] * length
    c2 = list(c1)
    if diff_type == SUBST:
        c2[0] = int(rng.choice([b for b in range(4) if b != center]))
    elif diff_type == DEL:
        c2.pop(0)
    else:
        c2.insert(1, int(rng.choice([b for b in range(4) if b != center])))
    return ACGT[np.array([right] + c1 + [left])], ACGT[np.array([right] + c2 + [left])]


def gen_reads(model: PairHiddenMarkovModel, templates: Sequence[np.ndarray], rng: np.random.Generator) -> List[np.ndarray]:
    """Sample one read per template from the pair HMM (kiley `Generate::gen`), all templates at once: the state machine
    is stepped with numpy over the whole set.  Start state Match; transitions row-normalised; Match emits by
    mat_emit[template base], Ins by ins_emit[previous read base] (4 = none), Del emits nothing; a read ends when the
    template is consumed."""
    n = len(templates)
    lens = np.array([len(t) for t in templates], dtype=np.int64)
    maxl = int(lens.max()) if n else 0
    tc = np.full((n, maxl + 1), 0, dtype=np.int64)
    lut = np.zeros(256, dtype=np.int64)
    lut[ACGT] = np.arange(4)
    for k, t in enumerate(templates):
        tc[k, :len(t)] = lut[np.asarray(t, dtype=np.uint8)]
    a = model.as_array()
    trans = a[:9].reshape(3, 3)
    trans_c = np.cumsum(trans / trans.sum(axis=1, keepdims=True), axis=1)
    mat_c = np.cumsum(a[9:25].reshape(4, 4) / a[9:25].reshape(4, 4).sum(axis=1, keepdims=True), axis=1)
    ins_c = np.cumsum(a[25:45].reshape(5, 4) / a[25:45].reshape(5, 4).sum(axis=1, keepdims=True), axis=1)
    state = np.zeros(n, dtype=np.int64)  # 0 Match, 1 Ins, 2 Del
    j = np.zeros(n, dtype=np.int64)
    prev = np.full(n, 4, dtype=np.int64)
    out = np.zeros((n, 3 * maxl + 32), dtype=np.uint8)
    olen = np.zeros(n, dtype=np.int64)
    active = j < lens
    rows = np.arange(n)
    while active.any():
        u = rng.random(n)
        nxt = (u[:, None] > trans_c[state]).sum(axis=1).clip(0, 2)
        v = rng.random(n)
        tb = tc[rows, np.minimum(j, maxl)]
        emit_m = (v[:, None] > mat_c[tb]).sum(axis=1).clip(0, 3)
        emit_i = (v[:, None] > ins_c[prev]).sum(axis=1).clip(0, 3)
        room = olen < out.shape[1]
        is_m = active & (nxt == 0)
        is_i = active & (nxt == 1) & room
        is_d = active & (nxt == 2)
        em = np.where(is_m, emit_m, emit_i)
        wr = (is_m | is_i) & room
        out[rows[wr], olen[wr]] = ACGT[em[wr]]
        prev = np.where(wr, em, prev)
        olen = olen + wr
        j = j + (is_m | is_d)
        state = np.where(active, nxt, state)
        active = j < lens
    return [out[k, :olen[k]].copy() for k in range(n)]


def _flatten(out: np.ndarray, lens: np.ndarray, fallback: Sequence[np.ndarray], fb_idx: np.ndarray):
    """(uint8[n, cap], lengths) -> (concatenated reads, uint32 offsets[n+1]); an empty read is replaced by the first base of
    its source (a zero-length read has no alignment)."""
    lens = lens.astype(np.int64)
    empty = np.flatnonzero(lens == 0)
    for k in empty:
        out[k, 0] = fallback[int(fb_idx[k])][0]
        lens[k] = 1
    mask = np.arange(out.shape[1])[None, :] < lens[:, None]
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return out[mask], off.astype(np.uint32)


def _lk_bootstrap(hmm: PairHiddenMarkovModelOnStrands, templates, reads, strands, tmpl_idx, band: int, ctx) -> np.ndarray:
    """likelihood_antidiagonal_bootstrap for many (template, read) pairs in ONE jtk_hmm_likelihood_batch; templates is a list
    of the distinct templates, pair k uses templates[tmpl_idx[k]]; reads a list or a (concat, offsets) tuple."""
    f, r = hmm.forward().to_c(), hmm.reverse().to_c()
    return ctx.likelihood_batch(f, r, templates, reads, None, strands, tmpl_idx, band)


def estimate_gain(hmm: PairHiddenMarkovModelOnStrands, seed: int, seq_len: int, band: int, homop_len: int,
                  ctx: Optional[_lib.Context] = None, sample_num: int = 100, seq_num: int = 50) -> Gains:
    """likelihood_gains.rs:162-184 with gain_of (:253-315) for every (type, length): the 2 * sample_num * seq_num * 9 * ...
    likelihood calls of the calibration (1.8e5 by default) are ONE GPU batch (two pairs per warp); the reads
    are sampled on the host threads (jtk_lc_gen_reads)."""
    ctx = ctx or default_context()
    gain_pos, prob_pos = sample_num // 10, sample_num * 2 // 3
    meta = []  # (type, len, sample) in generation order
    seqs = []  # sequences: 2m = template of sample m, 2m + 1 = the same with the variant
    for dt in (SUBST, DEL, INS):
        for length in range(1, homop_len + 1):
            for i in range(sample_num):
                rng = np.random.default_rng(i + seed)  # one stream per sample, as the reference seeds them (:269)
                seg1, seg2 = generate_seq(rng, seq_len // 2), generate_seq(rng, seq_len // 2)
                hap1, hap2 = gen_diff_haplotypes(rng, length, dt)
                seqs += [np.concatenate([seg1, hap1, seg2]), np.concatenate([seg1, hap2, seg2])]
                meta.append((dt, length, i))
    n_s = len(meta)
    # reads: per sample seq_num from the variant haplotype (expected gain) then seq_num from the template (null), even t from
    # the forward model, odd t from the reverse model (:276-279).  Read index = ((m * 2 + which) * seq_num + t)
    m_i, which, t_i = np.meshgrid(np.arange(n_s), np.arange(2), np.arange(seq_num), indexing="ij")
    src = (2 * m_i + (1 - which)).ravel().astype(np.uint32)       # which = 0: reads of the variant sequence 2m + 1
    is_fwd = (t_i.ravel() % 2 == 0)
    cap = int(max(len(x) for x in seqs)) * 3 + 32
    out = np.empty((len(src), cap), dtype=np.uint8)
    lens = np.empty(len(src), dtype=np.uint32)
    for model, sel, sd in ((hmm.forward(), is_fwd, seed), (hmm.reverse(), ~is_fwd, seed + 0x9E3779B9)):
        idx = np.flatnonzero(sel)
        o, l = _lib.gen_reads(model.as_array(), seqs, src[idx], sd, cap)
        out[idx] = o
        lens[idx] = l
    # every read is scored against the template (2m) and against the variant (2m + 1)
    rcat, roff = _flatten(np.repeat(out, 2, axis=0), np.repeat(lens, 2), seqs, np.repeat(2 * m_i.ravel(), 2))
    tmpl_idx = (2 * np.repeat(m_i.ravel(), 2) + np.tile(np.arange(2), len(src))).astype(np.uint32)
    strands = np.repeat(is_fwd, 2).astype(np.uint8)
    lk = _lk_bootstrap(hmm, seqs, (rcat, roff), strands, tmpl_idx, band, ctx).reshape(n_s, 2, seq_num, 2)
    gain = np.zeros((3, homop_len))
    prob = np.zeros((3, homop_len))
    per = {}
    for m, (dt, length, _) in enumerate(meta):
        d = lk[m, 0, :, 1] - lk[m, 0, :, 0]                 # lk_diff - lk_base on reads from the variant haplotype
        expected = float(np.sort(d)[seq_num // 2])          # select_nth_unstable_by(SEQ_NUM / 2)
        min_gain = expected / 10.0 if dt == SUBST else 0.0001
        null = float(np.mean(lk[m, 1, :, 0] + min_gain < lk[m, 1, :, 1]))
        per.setdefault((dt, length), []).append((expected, null))
    for (dt, length), xs in per.items():
        med = np.sort([x[0] for x in xs])
        prs = np.sort([x[1] for x in xs])
        gain[dt, length - 1] = med[gain_pos]
        prob[dt, length - 1] = max(prs[prob_pos], 1e-9)
    return Gains(gain=gain, prob=prob)


def estimate_gain_default(hmm: PairHiddenMarkovModelOnStrands, ctx: Optional[_lib.Context] = None) -> Gains:
    """likelihood_gains.rs:190-192."""
    return estimate_gain(hmm, SEED, SEQ_LEN, BAND, HOMOP_LEN, ctx=ctx)


def estimate_minimum_gain(hmm: PairHiddenMarkovModelOnStrands, ctx: Optional[_lib.Context] = None, sample_num: int = 1000,
                          seq_num: int = 500) -> float:
    """likelihood_gains.rs:6-39: 1000 templates x 500 reads x 2 likelihoods at 100 bp, band 25 -- 1e6 pairs, in slices of
    100 templates (1e5 pairs per GPU batch)."""
    ctx = ctx or default_context()
    seed0, length, band, min_req = 23908, 100, 25, 1.0
    medians = []
    block = 100
    for s0 in range(0, sample_num, block):
        n_here = min(block, sample_num - s0)
        seqs = []
        for s in range(s0, s0 + n_here):
            rng = np.random.default_rng(seed0 + s)
            hap1 = generate_seq(rng, length)
            pos = int(rng.integers(0, length))
            seqs += [hap1, np.delete(hap1, pos)]  # introduce_errors(hap1, rng, 0, 1, 0): one deletion
        m_i, t_i = np.meshgrid(np.arange(n_here), np.arange(seq_num), indexing="ij")
        src = (2 * m_i).ravel().astype(np.uint32)  # reads come from hap1
        is_fwd = t_i.ravel() < (seq_num + 1) // 2
        cap = length * 3 + 32
        out = np.empty((len(src), cap), dtype=np.uint8)
        lens = np.empty(len(src), dtype=np.uint32)
        for model, sel, sd in ((hmm.forward(), is_fwd, seed0 + s0), (hmm.reverse(), ~is_fwd, seed0 + s0 + 0x9E3779B9)):
            idx = np.flatnonzero(sel)
            o, l = _lib.gen_reads(model.as_array(), seqs, src[idx], sd, cap)
            out[idx] = o
            lens[idx] = l
        rcat, roff = _flatten(np.repeat(out, 2, axis=0), np.repeat(lens, 2), seqs, np.repeat(src, 2))
        tmpl_idx = (np.repeat(src, 2) + np.tile(np.arange(2, dtype=np.uint32), len(src))).astype(np.uint32)
        lk = _lk_bootstrap(hmm, seqs, (rcat, roff), np.repeat(is_fwd, 2).astype(np.uint8), tmpl_idx, band, ctx).reshape(n_here, seq_num, 2)
        d = lk[:, :, 0] - lk[:, :, 1]
        medians += [float(np.sort(row)[seq_num // 2]) for row in d]
    medians.sort()
    return max(medians[2], min_req)
