"""Window polish of a contig: the pair-HMM part of haplotyper::consensus (reference:
haplotyper/src/consensus/mod.rs).

  allocate_on_windows  consensus/mod.rs:270-298   alignments -> per-window pile-ups (`split`, :620-706)
  polish               consensus/mod.rs:300-371   round loop: allocate, polish every window, stitch, re-anchor alignments
  polish_seg           consensus/mod.rs:445-496   one window: length filter, (bialignment bootstrap), HMM polish, re-align

Reference work per round: `par_chunks(window)` over rayon threads, one `polish_until_converge_antidiagonal` per window.  Here
ALL windows of a round go to the GPU as one `jtk_polish_until_converge_batch` call (`polish_windows`); splitting, stitching
and the edit-distance re-alignment of the reads outside the length range stay on the host.

What is not restated (out of scope): the edit-distance (kiley::bialignment) polish of round 0
(`polish_until_converge_with`, :471-475) and the re-alignment of the partial-window tips of an alignment (`fix_alignment`,
:498-561, edlib infix + global_guided).  An alignment therefore keeps only its whole windows after a round: its tips are
trimmed, which the reference does not do.  Everything else follows the reference line by line.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .hmm import HMMPolishConfig, PairHiddenMarkovModelOnStrands, polish_chunks

OP_MATCH, OP_MISMATCH, OP_INS, OP_DEL = 0, 1, 2, 3
EDGE = 100  # consensus/mod.rs:616


def polish_radius(band_width: int) -> int:
    """assemble/mod.rs:190: radius = band_width(window).max(20) - 10."""
    return max(band_width, 20) - 10


@dataclass
class PolishConfig:
    """consensus::PolishConfig (consensus/mod.rs:37-44)."""
    min_coverage: int
    max_coverage: int
    window_size: int
    radius: int
    round_num: int
    seed: int = 0


@dataclass
class Alignment:
    """The fields of consensus::Alignment that the polish loop touches: a read segment aligned globally to
    contig[contig_start:contig_end]; ops: one byte per column (0 Match, 1 Mismatch, 2 Ins, 3 Del)."""
    query: np.ndarray
    ops: np.ndarray
    contig_start: int
    contig_end: int
    is_forward: bool = True


def global_align(query, target) -> np.ndarray:
    """consensus::global_align (:424-436): edlib global alignment in the reference, the library's edit-distance aligner here."""
    L = _lib.lib()
    L.jtk_align_global.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    q, t = _lib._u8(query), _lib._u8(target)
    out = np.zeros(len(q) + len(t) + 1, dtype=np.uint8)
    n = L.jtk_align_global(_lib._ptr(t), len(t), _lib._ptr(q), len(q), max(len(q), len(t)), _lib._ptr(out), len(out))
    if n < 0:
        raise _lib.JtkError(n, "jtk_align_global failed")
    return out[:n].copy()


def split(aln: Alignment, window: int, window_num: int, contig_len: int):
    """consensus::split (:620-706).  Returns (start, chunks, end): start / end = (query position, window index) of the part
    that was allocated, chunks = [(window index, query start, query end, ops of the window)].  A window is taken when the
    alignment covers it completely; the last window of the contig is also taken when the alignment stops within EDGE of
    the contig end (the missing columns become deletions); insertions that sit on a window boundary open the next window."""
    ops = aln.ops
    assert int(np.count_nonzero(ops != OP_DEL)) == len(aln.query), "ops do not span the query"
    assert int(np.count_nonzero(ops != OP_INS)) == aln.contig_end - aln.contig_start, "ops do not span the contig range"
    start_pos_in_contig = aln.contig_start if aln.contig_start % window == 0 else (aln.contig_start // window + 1) * window
    qstep = (ops != OP_DEL).astype(np.int64)
    cstep = (ops != OP_INS).astype(np.int64)
    cpos_after = aln.contig_start + np.cumsum(cstep)   # contig position after column k
    qpos_after = np.cumsum(qstep)
    n = len(ops)
    # seek: consume columns while cpos < start_pos_in_contig
    k = int(np.searchsorted(cpos_after, start_pos_in_contig, side="left")) + 1 if start_pos_in_contig > aln.contig_start else 0
    k = min(k, n)
    cpos = int(cpos_after[k - 1]) if k > 0 else aln.contig_start
    qpos = int(qpos_after[k - 1]) if k > 0 else 0
    start_chunk_id = cpos // window
    start = (qpos, start_chunk_id)
    if cpos < start_pos_in_contig:
        return start, [], start
    chunks = []
    cur = start_pos_in_contig // window
    end_pos = qpos
    while True:
        target = (cur + 1) * window
        # columns k.. while cpos < target: up to and including the column that brings cpos to target
        k2 = int(np.searchsorted(cpos_after, target, side="left")) + 1
        if k2 > n:  # the ops run out inside this window
            cpos_end = int(cpos_after[n - 1]) if n > 0 else aln.contig_start
            if cur == window_num - 1 and contig_len - cpos_end < EDGE and k <= n:
                qend = int(qpos_after[n - 1]) if n > 0 else 0
                assert qend == len(aln.query) and cpos_end <= contig_len
                w_ops = np.concatenate([ops[k:n], np.full(contig_len - cpos_end, OP_DEL, dtype=np.uint8)])
                chunks.append((cur, qpos, qend, w_ops))
                end_pos = qend
                cur += 1
            break
        qend = int(qpos_after[k2 - 1])
        chunks.append((cur, qpos, qend, ops[k:k2].copy()))
        end_pos = qend
        k, qpos = k2, qend
        cur += 1
    return start, chunks, (end_pos, cur)


def allocate_on_windows(alignments: Sequence[Alignment], round_: int, window: int, draft_len: int):
    """consensus::allocate_on_windows (:270-298): slots[w] = [(alignment index, piece index, is_forward, seq, ops)], and the
    used range of every alignment.  After round 0 the pile-up of a window is sorted by its number of non-Match columns
    (stable, as sort_by_cached_key), so that the cleanest max_coverage reads vote in the HMM polish."""
    num_slot = draft_len // window + (1 if draft_len % window != 0 else 0)
    slots: List[list] = [[] for _ in range(num_slot)]
    used_range: Dict[int, tuple] = {}
    for ai, aln in enumerate(alignments):
        start, chunks, end = split(aln, window, num_slot, draft_len)
        if not chunks:
            assert start[0] == end[0] or start[0] + 1 == end[0]
        for idx, (pos, q0, q1, ops) in enumerate(chunks):
            slots[pos].append((ai, idx, aln.is_forward, aln.query[q0:q1], ops))
        used_range[ai] = (start, end)
    if round_ != 0:
        for pile in slots:
            pile.sort(key=lambda x: int(np.count_nonzero(x[4] != OP_MATCH)))
    return slots, used_range


def _within_range(median: int, length: int, frac: float) -> bool:
    return (max(median, length) - min(median, length)) / median < frac   # :378-381


def polish_windows(models: PairHiddenMarkovModelOnStrands, drafts: Sequence[np.ndarray], windows: Sequence[Tuple[list, list, list]],
                   radius: int, max_cov: int, ctx=None) -> Tuple[List[np.ndarray], List[list]]:
    """The HMM step of `polish_seg` (:476-483) for every window of a round in ONE batch.  windows[w] = (seqs, ops, strands)
    of the reads that vote in window w.  Returns the polished window sequences and the rewritten ops (the reference asserts
    that they span the polished sequence, :484-488)."""
    if not windows:
        return [], []
    reads = [s for w in windows for s in w[0]]
    ops = [o for w in windows for o in w[1]]
    strands = [s for w in windows for s in w[2]]
    tidx = np.repeat(np.arange(len(windows), dtype=np.uint32), [len(w[0]) for w in windows])
    cfg = HMMPolishConfig.new(radius // 2, max_cov, 0)   # consensus/mod.rs:476
    cons, new_ops, iters = polish_chunks(models, list(drafts), reads, ops, strands, tidx, cfg, ctx=ctx)
    polish_windows.last_iters = iters
    out_ops, k = [], 0
    for w in windows:
        out_ops.append(new_ops[k:k + len(w[0])])
        k += len(w[0])
    for c, w_ops, w in zip(cons, out_ops, windows):
        for o, s in zip(w_ops, w[0]):
            assert np.count_nonzero(o != 2) == len(c) and np.count_nonzero(o != 3) == len(s)   # consensus/mod.rs:461-464,487-488
    return cons, out_ops


polish_windows.last_iters = None


def polish(draft, alignments: List[Alignment], models: PairHiddenMarkovModelOnStrands, config: PolishConfig, ctx=None,
           stats: Optional[dict] = None) -> np.ndarray:
    """consensus::polish (:300-371).  Mutates `alignments` (ops, contig range, query trimmed to whole windows) and returns
    the polished contig.  stats (optional) receives per-round window counts and polish rounds."""
    window = config.window_size
    polished = _lib._u8(draft).copy()
    for round_ in range(config.round_num):
        slots, used = allocate_on_windows(alignments, round_, window, len(polished))
        segs = [polished[w * window:(w + 1) * window] for w in range(len(slots))]
        out_seg: List[Optional[np.ndarray]] = [None] * len(slots)
        new_piece_ops: Dict[Tuple[int, int], np.ndarray] = {}
        todo, todo_drafts, todo_meta = [], [], []
        for w, pile in enumerate(slots):
            if len(pile) < config.min_coverage:                           # :327-329
                out_seg[w] = segs[w].copy()
                for (ai, idx, _, _, ops) in pile:
                    new_piece_ops[(ai, idx)] = ops
                continue
            # ---- polish_seg (:445-496) up to the HMM call ----
            lens = sorted(len(p[3]) for p in pile)
            median = lens[len(pile) // 2]                                  # length_median (:372-376)
            if median == 0:                                                # remove_reference (:383-390)
                out_seg[w] = np.zeros(0, dtype=np.uint8)
                for (ai, idx, _, seq, ops) in pile:
                    new_piece_ops[(ai, idx)] = np.full(len(seq), OP_INS, dtype=np.uint8)
                continue
            use = [p for p in pile if _within_range(median, len(p[3]), 0.15)]       # split_sequences (:399-422)
            rest = [p for p in pile if not _within_range(median, len(p[3]), 0.15)]
            seg = segs[w]
            use_ops = [p[4] for p in use]
            if not _within_range(len(seg), median, 0.2):                  # bootstrap_consensus (:437-443), without the
                seg = use[0][3].copy()                                    # edit-distance polish (out of scope)
                use_ops = [global_align(p[3], seg) for p in use]
            todo.append(([p[3] for p in use], use_ops, [p[2] for p in use]))
            todo_drafts.append(seg)
            todo_meta.append((w, use, rest))
        cons, ops_out = polish_windows(models, todo_drafts, todo, config.radius, config.max_coverage, ctx=ctx)
        for (w, use, rest), c, o in zip(todo_meta, cons, ops_out):
            out_seg[w] = c
            for p, po in zip(use, o):
                new_piece_ops[(p[0], p[1])] = po
            for p in rest:                                                 # :489-493
                new_piece_ops[(p[0], p[1])] = global_align(p[3], c)
        acc_len = np.concatenate([[0], np.cumsum([len(s) for s in out_seg])]).astype(np.int64)
        polished = np.concatenate(out_seg) if out_seg else polished
        # re-anchor the alignments on the polished contig (the whole-window part of fix_alignment, :498-561)
        for ai, aln in enumerate(alignments):
            (q0, w0), (q1, w1) = used[ai]
            n_pieces = w1 - w0 if q1 > q0 or w1 > w0 else 0
            pieces = [new_piece_ops[(ai, i)] for i in range(n_pieces) if (ai, i) in new_piece_ops]
            if not pieces:
                aln.query = aln.query[q0:q0]
                aln.ops = np.zeros(0, dtype=np.uint8)
                aln.contig_start = aln.contig_end = int(acc_len[min(w0, len(acc_len) - 1)])
                continue
            aln.query = aln.query[q0:q1]
            aln.ops = np.concatenate(pieces)
            aln.contig_start, aln.contig_end = int(acc_len[w0]), int(acc_len[w1])
            assert int(np.count_nonzero(aln.ops != OP_INS)) == aln.contig_end - aln.contig_start
            assert int(np.count_nonzero(aln.ops != OP_DEL)) == len(aln.query)
        if stats is not None:
            it = polish_windows.last_iters
            stats.setdefault("rounds", []).append({"windows": len(slots), "polished_windows": len(todo),
                                                   "hmm_rounds_max": int(max(it)) if it is not None and len(it) else 0,
                                                   "length": int(len(polished))})
    return polished
