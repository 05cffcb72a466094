"""K4 fit: one Baum-Welch step, GPU E-step vs the f64 oracle (expected counts and updated parameters)."""
import numpy as np
import pytest

import oracle_lib as O
from jtk_b200 import synth


def make_packs(seed, n_packs=3, L=500, n=20):
    rng = np.random.default_rng(seed)
    packs = []
    for _ in range(n_packs):
        t = synth.random_template(rng, L)
        reads, ops = zip(*[synth.mutate_read(rng, t, 0.1) for _ in range(n)])
        strands = (rng.random(n) < 0.5).astype(np.uint8)
        packs.append((t, strands, list(reads), list(ops)))
    return packs


def test_oracle_fit_moves_towards_the_generating_error_rates():
    packs = make_packs(1)
    h = O.default_hmm()
    f, r, accf, accr = O.fit(h, h, packs, 20)
    a = f.as_array()
    assert abs(a[0:3].sum() - 1) < 1e-12 and abs(a[9:13].sum() - 1) < 1e-12
    # reads were generated with ~3.3 % substitutions / insertions / deletions per base
    assert 0.9 < a[0] < 0.97 and 0.015 < a[1] < 0.06 and 0.015 < a[2] < 0.06
    assert 0.9 < a[9] < 0.99


@pytest.mark.gpu
def test_gpu_expected_counts_and_fit_match_oracle():
    from jtk_b200 import hmm
    packs = make_packs(2)
    rng = np.random.default_rng(9)
    a = O.default_hmm().as_array()
    a[25:45] = np.concatenate([rng.dirichlet([3, 3, 3, 3]) for _ in range(5)])  # context-dependent insertions
    fo = O.OrcHmm.from_array(a)
    ro = O.default_hmm()
    wf, wr, accf, accr = O.fit(fo, ro, packs, 20)
    models = hmm.PairHiddenMarkovModelOnStrands.new(hmm.PairHiddenMarkovModel.from_array(fo.as_array()),
                                                    hmm.PairHiddenMarkovModel.from_array(ro.as_array()))
    templates = [p[0] for p in packs]
    reads = [q for p in packs for q in p[2]]
    ops = [o for p in packs for o in p[3]]
    strands = np.concatenate([p[1] for p in packs])
    tidx = np.repeat(np.arange(len(packs), dtype=np.uint32), [len(p[2]) for p in packs])
    acc = hmm.expected_counts(models, templates, reads, ops, strands, tidx, 20)
    assert np.allclose(acc[0], accf, rtol=2e-4, atol=2e-3)
    assert np.allclose(acc[1], accr, rtol=2e-4, atol=2e-3)
    hmm.fit_antidiagonal_par_multiple(models, packs, 20)
    assert np.allclose(models.forward().as_array(), wf.as_array(), atol=2e-5)
    assert np.allclose(models.reverse().as_array(), wr.as_array(), atol=2e-5)
