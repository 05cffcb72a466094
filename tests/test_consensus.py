"""Window polish of a contig (jtk_b200/consensus.py; reference haplotyper/src/consensus/mod.rs:270-371,445-496,620-706):
host logic on CPU, the round loop on the GPU."""
import numpy as np
import pytest

from jtk_b200 import consensus as CS
from jtk_b200 import synth


def make_alignment(rng, contig, start, end, err):
    q, ops = synth.mutate_read(rng, contig[start:end], err)
    return CS.Alignment(query=q, ops=ops, contig_start=start, contig_end=end, is_forward=bool(rng.integers(0, 2)))


def test_split_takes_whole_windows_and_pads_the_contig_end():
    """consensus::split (:620-706): an alignment from 1 500 to 5 250 on a 5 300 bp contig with 2 000 bp windows skips the
    partial first window, takes window 1 whole, and takes the last window because it stops within EDGE of the contig end
    (the 50 missing columns become deletions)."""
    rng = np.random.default_rng(1)
    contig = synth.random_template(rng, 5300)
    aln = make_alignment(rng, contig, 1500, 5250, 0.1)
    start, chunks, end = CS.split(aln, 2000, 3, 5300)
    assert [c[0] for c in chunks] == [1, 2]
    assert start[1] == 1 and end == (len(aln.query), 3)
    lens = {1: 2000, 2: 1300}
    q_prev = start[0]
    for w, q0, q1, ops in chunks:
        assert q0 == q_prev
        assert np.count_nonzero(ops != CS.OP_INS) == lens[w]
        assert np.count_nonzero(ops != CS.OP_DEL) == q1 - q0
        q_prev = q1
    assert (chunks[1][3][-50:] == CS.OP_DEL).all()
    # an alignment that stops 500 bp before the contig end loses its partial last window
    aln2 = make_alignment(rng, contig, 0, 4800, 0.05)
    s2, c2, e2 = CS.split(aln2, 2000, 3, 5300)
    assert [c[0] for c in c2] == [0, 1] and e2[1] == 2 and s2 == (0, 0)
    # an alignment inside one window is not allocated at all
    aln3 = make_alignment(rng, contig, 2100, 3900, 0.05)
    s3, c3, e3 = CS.split(aln3, 2000, 3, 5300)
    assert c3 == [] and s3 == e3


def test_split_sends_boundary_insertions_to_the_next_window():
    contig = np.frombuffer(b"ACGTACGTAC", np.uint8)
    # 10 template columns, window 5: an insertion right after column 5 belongs to window 1
    ops = np.array([0, 0, 0, 0, 0, 2, 2, 0, 0, 0, 0, 0], dtype=np.uint8)
    q = np.frombuffer(b"ACGTATTCGTAC", np.uint8)
    aln = CS.Alignment(q, ops, 0, 10)
    _, chunks, end = CS.split(aln, 5, 2, 10)
    assert [list(c[3]) for c in chunks] == [[0] * 5, [2, 2, 0, 0, 0, 0, 0]]
    assert [(c[1], c[2]) for c in chunks] == [(0, 5), (5, 12)] and end == (12, 2)


def test_allocate_on_windows_sorts_cleanest_first_after_round_zero():
    rng = np.random.default_rng(5)
    contig = synth.random_template(rng, 1200)
    alns = [make_alignment(rng, contig, 0, 1200, e) for e in (0.2, 0.02, 0.1, 0.0)]
    slots0, used0 = CS.allocate_on_windows(alns, 0, 400, 1200)
    assert [len(s) for s in slots0] == [4, 4, 4]
    assert [p[0] for p in slots0[0]] == [0, 1, 2, 3]                     # round 0: input order
    slots1, _ = CS.allocate_on_windows(alns, 1, 400, 1200)
    for pile in slots1:
        bad = [int(np.count_nonzero(p[4] != CS.OP_MATCH)) for p in pile]
        assert bad == sorted(bad) and pile[0][0] == 3
    assert set(used0) == {0, 1, 2, 3}


def test_global_align_matches_the_oracle_aligner():
    import oracle_lib as O
    rng = np.random.default_rng(9)
    t = synth.random_template(rng, 300)
    q, _ = synth.mutate_read(rng, t, 0.15)
    ops = CS.global_align(q, t)
    assert np.count_nonzero(ops != CS.OP_INS) == len(t) and np.count_nonzero(ops != CS.OP_DEL) == len(q)
    assert np.array_equal(ops, O.edit_ops(t, q, max(len(t), len(q))))
    assert (CS.global_align(np.zeros(0, np.uint8), t) == CS.OP_DEL).all()
    assert (CS.global_align(q, np.zeros(0, np.uint8)) == CS.OP_INS).all()


@pytest.mark.gpu
def test_polish_rounds_recover_the_contig():
    """consensus::polish (:300-371) on a 1 700 bp contig with 600 bp windows: a draft with 24 substitution errors, 40 reads
    over random ranges; after the rounds the contig equals the truth and every alignment still spans its contig range."""
    from jtk_b200.hmm import PairHiddenMarkovModelOnStrands
    rng = np.random.default_rng(77)
    truth = synth.random_template(rng, 1700)
    draft = truth.copy()
    pos = rng.choice(np.arange(10, 1690), size=24, replace=False)
    draft[pos] = synth.ACGT[(np.searchsorted(synth.ACGT, draft[pos]) + rng.integers(1, 4, size=24)) % 4]
    alns = []
    for k in range(40):
        a = int(rng.integers(0, 500)) if k % 3 else 0
        b = int(rng.integers(1300, 1700)) if k % 4 else 1700
        alns.append(make_alignment(rng, truth, a, b, 0.08))   # substitutions only in the draft: the true path is a valid guide
    cfg = CS.PolishConfig(min_coverage=3, max_coverage=25, window_size=600, radius=40, round_num=2)
    stats = {}
    out = CS.polish(draft, alns, PairHiddenMarkovModelOnStrands.default(), cfg, stats=stats)
    assert len(stats["rounds"]) == 2 and stats["rounds"][0]["windows"] == 3
    ident = (out == truth).mean() if len(out) == len(truth) else 0.0
    assert len(out) == len(truth) and ident == 1.0, ident
    for aln in alns:
        assert np.count_nonzero(aln.ops != CS.OP_INS) == aln.contig_end - aln.contig_start
        assert np.count_nonzero(aln.ops != CS.OP_DEL) == len(aln.query)
        assert aln.contig_start % 600 == 0 and (aln.contig_end % 600 == 0 or aln.contig_end == len(out))
