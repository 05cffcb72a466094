"""K3 polish loop: oracle behaviour on CPU; GPU loop vs oracle (identical consensus and guide ops)."""
import numpy as np
import pytest

import oracle_lib as O
from jtk_b200 import synth


def make_case(seed, L=400, n=24, err=0.08, draft_err=0.03):
    rng = np.random.default_rng(seed)
    truth = synth.random_template(rng, L)
    draft, _ = synth.mutate_read(rng, truth, draft_err)
    reads = [synth.mutate_read(rng, truth, err)[0] for _ in range(n)]
    ops = [O.edit_ops(draft, r, 30) for r in reads]
    strands = (rng.random(n) < 0.5).astype(np.uint8)
    return truth, draft, reads, ops, strands


def check_ops_span(ops, Lt, Lr):
    ops = np.asarray(ops)
    assert (ops != 2).sum() == Lt and (ops != 3).sum() == Lr


def test_oracle_polish_recovers_truth():
    truth, draft, reads, ops, strands = make_case(3)
    h = O.default_hmm()
    cons, new_ops, iters = O.polish_until_converge(h, h, draft, reads, ops, strands, 20, len(reads), 3)
    assert 1 <= iters < 20
    assert bytes(cons[3:-3]) in bytes(truth) or bytes(cons) == bytes(truth)
    for o, r in zip(new_ops, reads):
        check_ops_span(o, len(cons), len(r))  # consensus/mod.rs:461-464 invariant
    # a converged consensus is a fixed point
    cons2, _, it2 = O.polish_until_converge(h, h, cons, reads, new_ops, strands, 20, len(reads), 3)
    assert it2 == 0 and bytes(cons2) == bytes(cons)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,take,edge", [(5, 24, 3), (6, 12, 0), (7, 24, 0)])
def test_gpu_polish_matches_oracle(seed, take, edge):
    from jtk_b200 import hmm
    truth, draft, reads, ops, strands = make_case(seed)
    h = O.default_hmm()
    want_cons, want_ops, want_it = O.polish_until_converge(h, h, draft, reads, ops, strands, 20, take, edge)
    m = hmm.PairHiddenMarkovModelOnStrands.default()
    gops = [o.copy() for o in ops]
    cons = m.polish_until_converge_antidiagonal(draft, reads, gops, strands, hmm.HMMPolishConfig.new(20, take, edge))
    assert bytes(cons) == bytes(want_cons)
    for a, b in zip(gops, want_ops):
        assert a.tolist() == b.tolist()


@pytest.mark.gpu
def test_gpu_polish_many_chunks():
    from jtk_b200 import hmm
    cases = [make_case(20 + c, L=300 + 40 * c, n=16) for c in range(4)]
    h = O.default_hmm()
    drafts = [c[1] for c in cases]
    reads = [r for c in cases for r in c[2]]
    ops = [o for c in cases for o in c[3]]
    strands = np.concatenate([c[4] for c in cases])
    tidx = np.repeat(np.arange(4, dtype=np.uint32), 16)
    m = hmm.PairHiddenMarkovModelOnStrands.default()
    cons, new_ops, iters = hmm.polish_chunks(m, drafts, reads, ops, strands, tidx, hmm.HMMPolishConfig.new(20, 16, 0))
    for c, case in enumerate(cases):
        want_cons, want_ops, want_it = O.polish_until_converge(h, h, case[1], case[2], case[3], case[4], 20, 16, 0)
        assert bytes(cons[c]) == bytes(want_cons) and iters[c] == want_it
        for k in range(16):
            assert new_ops[c * 16 + k].tolist() == want_ops[k].tolist()


@pytest.mark.gpu
def test_consensus_window_polish_repairs_drafts():
    """consensus::polish_seg's HMM step for several windows in one batch (consensus/mod.rs:445-496): drafts with planted
    errors converge to the truth, ops keep spanning (draft, read)."""
    from jtk_b200 import consensus, hmm
    rng = np.random.default_rng(12)
    models = hmm.PairHiddenMarkovModelOnStrands.default()
    drafts, windows, truths = [], [], []
    for w in range(4):
        truth = synth.random_template(rng, 500 + 20 * w)
        d = truth.copy()
        for p in (60, 200, 333):
            d[p] = synth.ACGT[(np.searchsorted(synth.ACGT, d[p]) + 1 + w % 3) % 4]
        d = np.delete(d, 410)
        seqs, ops, strands = [], [], []
        for _ in range(14):
            q, _o = synth.mutate_read(rng, truth, 0.07)
            seqs.append(q); ops.append(O.edit_ops(d, q, 40)); strands.append(bool(rng.random() < 0.5))
        drafts.append(d); windows.append((seqs, ops, strands)); truths.append(truth)
    radius = consensus.polish_radius(60)
    assert radius == 50
    cons, new_ops = consensus.polish_windows(models, drafts, windows, radius, 50)
    for c, t in zip(cons, truths):
        assert np.array_equal(c, t)


def clustered_error_case(seed, L=700, n=20, n_err=12):
    """A draft whose errors come in clusters (pairs 2-6 columns apart) and that contains a tandem repeat with one unit missing:
    the shapes on which a pick of the FIRST positive-gain column oscillates (spurious indels next to a substitution, two
    equivalent unit deletions applied at once)."""
    rng = np.random.default_rng(seed)
    truth = synth.random_template(rng, L)
    unit = synth.random_template(rng, 6)
    truth[300:336] = np.tile(unit, 6)
    draft = truth.copy()
    pos = []
    for k in range(n_err // 2):
        a = 30 + k * 100 + int(rng.integers(0, 20))
        pos += [a, a + int(rng.integers(2, 7))]
    pos = np.array(pos)
    draft[pos] = synth.ACGT[(np.searchsorted(synth.ACGT, draft[pos]) + rng.integers(1, 4, size=len(pos))) % 4]
    draft = np.delete(draft, np.arange(312, 318))   # one repeat unit missing
    reads, ops = [], []
    for _ in range(n):
        q, _o = synth.mutate_read(rng, truth, 0.08)
        reads.append(q); ops.append(O.edit_ops(draft, q, 30))
    strands = (rng.random(n) < 0.5).astype(np.uint8)
    return truth, draft, reads, ops, strands


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_polish_converges_on_clustered_errors_and_repeats(seed):
    truth, draft, reads, ops, strands = clustered_error_case(seed)
    h = O.default_hmm()
    cons, new_ops, iters = O.polish_until_converge(h, h, draft, reads, ops, strands, 30, len(reads), 0)
    assert iters <= 10, iters
    assert bytes(cons) == bytes(truth)
    for o, r in zip(new_ops, reads):
        check_ops_span(o, len(cons), len(r))


@pytest.mark.gpu
def test_gpu_polish_converges_on_clustered_errors_like_the_oracle():
    from jtk_b200 import hmm
    h = O.default_hmm()
    m = hmm.PairHiddenMarkovModelOnStrands.default()
    for seed in (1, 2):
        truth, draft, reads, ops, strands = clustered_error_case(seed)
        want_cons, want_ops, want_it = O.polish_until_converge(h, h, draft, reads, ops, strands, 30, len(reads), 0)
        gops = [o.copy() for o in ops]
        cons = m.polish_until_converge_antidiagonal(draft, reads, gops, strands, hmm.HMMPolishConfig.new(30, len(reads), 0))
        assert bytes(cons) == bytes(want_cons) == bytes(truth)
        for a, b in zip(gops, want_ops):
            assert a.tolist() == b.tolist()
