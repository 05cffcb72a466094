"""Pins for the CPU oracle.  The reference holds no golden vector for the kiley
boundary (parity unpinned), so the oracle is anchored on definitional invariants P1-P5."""
import itertools

import numpy as np
import pytest

import oracle_lib as O
from jtk_b200 import synth


def random_hmm(seed):
    rng = np.random.default_rng(seed)
    a = np.empty(45)
    for s in range(3):
        row = rng.dirichlet([30, 2, 2]) * 0.99  # rows need not sum to 1 (defaults sum to 0.99)
        a[3 * s:3 * s + 3] = row
    for r in range(4):
        e = rng.dirichlet([1, 1, 1, 1]) * 0.2
        e[r] += 0.8
        a[9 + 4 * r:13 + 4 * r] = e
    for c in range(5):
        a[25 + 4 * c:29 + 4 * c] = rng.dirichlet([3, 3, 3, 3])
    return O.OrcHmm.from_array(a)


def brute_edit_lk(h, t, q, j, row, R):
    e = O.apply_edit(t, j, row)
    if e is None:
        return None
    if len(e) == 0 and len(q) == 0:
        return 0.0
    ops = O.edit_ops(e, q, R)
    return O.likelihood(h, e, q, ops, R)


@pytest.mark.parametrize("seed,L,err", [(1, 12, 0.2), (2, 25, 0.1), (3, 40, 0.15), (4, 8, 0.3), (5, 33, 0.0)])
@pytest.mark.parametrize("model", ["default", "random"])
def test_p1_table_is_likelihood_of_edited_template(seed, L, err, model):
    rng = np.random.default_rng(seed)
    h = O.default_hmm() if model == "default" else random_hmm(seed)
    t = synth.random_template(rng, L)
    q, ops = synth.mutate_read(rng, t, err)
    R = 4 * L + 10  # full band
    tab, lk = O.modification_table(h, t, q, ops, R)
    tab = tab.reshape(-1, 14)
    assert abs(lk - O.likelihood(h, t, q, ops, R)) < 1e-10
    for j, row in itertools.product(range(L + 1), range(14)):
        want = brute_edit_lk(h, t, q, j, row, R)
        if want is None:
            assert tab[j, row] == -1e10
        else:
            assert abs(tab[j, row] - want) < 1e-9, (j, row, tab[j, row], want)


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("R", [3, 10, 30])
def test_p2_forward_equals_backward(seed, R):
    rng = np.random.default_rng(100 + seed)
    h = random_hmm(seed)
    t = synth.random_template(rng, 300)
    q, ops = synth.mutate_read(rng, t, 0.1)
    f = O.likelihood(h, t, q, ops, R)
    b = O.likelihood(h, t, q, ops, R, backward=True)
    assert np.isfinite(f) and abs(f - b) < 1e-8 * abs(f)


@pytest.mark.parametrize("R", [5, 30])
def test_p3_identity_substitution_is_lk(R):
    rng = np.random.default_rng(9)
    h = random_hmm(9)
    t = synth.random_template(rng, 400)
    q, ops = synth.mutate_read(rng, t, 0.08)
    tab, lk = O.modification_table(h, t, q, ops, R)
    tab = tab.reshape(-1, 14)
    code = np.searchsorted(synth.ACGT, t)
    own = tab[np.arange(len(t)), code]
    assert np.max(np.abs(own - lk)) < 1e-9


def test_p4_band_converges_to_full_dp():
    rng = np.random.default_rng(11)
    h = O.default_hmm()
    t = synth.random_template(rng, 200)
    q, ops = synth.mutate_read(rng, t, 0.1)
    full = O.likelihood(h, t, q, ops, 1000)
    prev = None
    for R in (2, 4, 8, 16, 32, 64):
        lk = O.likelihood(h, t, q, ops, R)
        assert lk <= full + 1e-9  # a band only removes paths
        if prev is not None:
            assert lk >= prev - 1e-9
        prev = lk
    assert abs(prev - full) < 1e-6


def test_p5_total_probability_over_reads():
    """sum over all reads (up to a length cap) of P(read|template) approaches the mass implied by the
    transition rows.  There is no explicit end transition (SURVEY A.1 [FREE]), so after the last template
    base a path may still emit a geometric tail of insertions: with normalised rows and equal X->Ins
    probabilities the total is 1 + p_ins / (1 - ins_ins), approached from below as the cap grows."""
    a = O.default_hmm().as_array()
    for s in range(3):
        a[3 * s:3 * s + 3] /= a[3 * s:3 * s + 3].sum()
    h = O.OrcHmm.from_array(a)
    t = np.frombuffer(b"AC", dtype=np.uint8)
    tot = 0.0
    for n in range(0, 6):
        for tup in itertools.product(b"ACGT", repeat=n):
            q = np.array(tup, dtype=np.uint8)
            ops = O.edit_ops(t, q, 50) if n else np.array([3, 3], dtype=np.uint8)
            tot += np.exp(O.likelihood(h, t, q, ops, 50))
    want = 1.0 + a[1] / (1.0 - a[4])
    assert want - 1e-3 < tot <= want + 1e-9, (tot, want)


def test_ragged_and_tiny_inputs():
    h = O.default_hmm()
    # read much shorter / longer than template, and length-1 sequences
    for t, q in ((b"ACGTACGTAC", b"AC"), (b"AC", b"ACGTTTTTGA"), (b"A", b"A"), (b"A", b"C"), (b"ACG", b"")):
        t = np.frombuffer(t, dtype=np.uint8)
        q = np.frombuffer(q, dtype=np.uint8)
        ops = O.edit_ops(t, q, 20) if len(q) else np.full(len(t), 3, dtype=np.uint8)
        f = O.likelihood(h, t, q, ops, 20)
        b = O.likelihood(h, t, q, ops, 20, backward=True)
        assert np.isfinite(f) and abs(f - b) < 1e-10
        tab, lk = O.modification_table(h, t, q, ops, 20)
        assert np.isfinite(tab[tab > -1e9]).all()


def test_bad_ops_rejected():
    h = O.default_hmm()
    t = np.frombuffer(b"ACGT", dtype=np.uint8)
    with pytest.raises(ValueError):
        O.modification_table(h, t, t, np.zeros(3, dtype=np.uint8), 5)


def test_cell_count_matches_numpy_restatement():
    rng = np.random.default_rng(5)
    t = synth.random_template(rng, 500)
    q, ops = synth.mutate_read(rng, t, 0.1)
    for R in (3, 30, 100):
        assert O.cell_count(ops, len(t), len(q), R) == synth.cell_count(ops, len(t), len(q), R)


def test_expected_counts_sum_rules():
    """Posterior path counts: every path consumes Lt template bases and Lr read bases."""
    rng = np.random.default_rng(21)
    h = random_hmm(3)
    t = synth.random_template(rng, 150)
    q, ops = synth.mutate_read(rng, t, 0.1)
    acc = O.expected_counts(h, t, q, ops, 20)
    to_m = acc[0] + acc[3] + acc[6]
    to_i = acc[1] + acc[4] + acc[7]
    to_d = acc[2] + acc[5] + acc[8]
    assert abs(to_m + to_d - len(t)) < 1e-6
    assert abs(to_m + to_i - len(q)) < 1e-6
    assert abs(acc[9:25].sum() - to_m) < 1e-9 and abs(acc[25:45].sum() - to_i) < 1e-9
