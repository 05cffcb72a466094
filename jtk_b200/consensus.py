"""Mirror of the pair-HMM part of haplotyper::consensus (secondary target; reference: haplotyper/src/consensus/mod.rs).

`polish_seg` (consensus/mod.rs:445-496) polishes one 2 kbp window of a contig: after the edit-distance bootstrap (round 0,
out of scope: kiley::bialignment) it calls `polish_until_converge_antidiagonal` with
`HMMPolishConfig::new(radius / 2, max_cov, 0)` (:476-483).  `polish` (:300-371) does that for every window of a contig
(`par_chunks`, :316-331).  Here all windows of a round go to the GPU as ONE batch; window allocation, chaining and
`fix_alignment` stay on the host."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .hmm import HMMPolishConfig, PairHiddenMarkovModelOnStrands, polish_chunks


def polish_radius(band_width: int) -> int:
    """assemble/mod.rs:190: radius = band_width(window).max(20) - 10."""
    return max(band_width, 20) - 10


def polish_windows(models: PairHiddenMarkovModelOnStrands, drafts: Sequence[np.ndarray], windows: Sequence[Tuple[list, list, list]],
                   radius: int, max_cov: int, ctx=None) -> Tuple[List[np.ndarray], List[list]]:
    """The HMM step of `polish_seg` for every window of a round.  windows[w] = (seqs, ops, strands) of the reads allocated
    to window w (`allocate_on_windows`, consensus/mod.rs:270-298: at most max_cov reads, cleanest first).  Returns the
    polished window sequences and the rewritten ops (the reference asserts that they span the polished sequence,
    consensus/mod.rs:484-488)."""
    reads = [s for w in windows for s in w[0]]
    ops = [o for w in windows for o in w[1]]
    strands = [s for w in windows for s in w[2]]
    tidx = np.repeat(np.arange(len(windows), dtype=np.uint32), [len(w[0]) for w in windows])
    cfg = HMMPolishConfig.new(radius // 2, max_cov, 0)   # consensus/mod.rs:476
    cons, new_ops, _ = polish_chunks(models, list(drafts), reads, ops, strands, tidx, cfg, ctx=ctx)
    out_ops, k = [], 0
    for w in windows:
        out_ops.append(new_ops[k:k + len(w[0])])
        k += len(w[0])
    for c, w_ops, w in zip(cons, out_ops, windows):
        for o, s in zip(w_ops, w[0]):
            assert np.count_nonzero(o != 2) == len(c) and np.count_nonzero(o != 3) == len(s)   # consensus/mod.rs:461-464,487-488
    return cons, out_ops
