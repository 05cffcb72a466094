"""Committed golden vectors (tests/golden/phmm_golden.npz, written by tests/golden/make_golden.py).

CPU: the oracle reproduces its committed outputs bit for bit (pins the restatement; the reference itself has no vector
for this path and kiley cannot be built here -- SURVEY.md 8c).  GPU: the CUDA path against the committed numbers."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phmm_golden.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files if not k.startswith("polish/")})


def case(name):
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    if "radius" in g:
        g["radius"] = int(g["radius"][0])
    return g


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_committed_vectors(name):
    g = case(name)
    h = O.OrcHmm.from_array(g["params"])
    table, lk = O.modification_table(h, g["template"], g["read"], g["ops"], g["radius"])
    assert lk == g["lk"][0]
    assert np.array_equal(np.asarray(table), g["table"])
    assert O.likelihood_bootstrap(h, g["template"], g["read"], g["radius"]) == g["lk_bootstrap"][0]
    assert np.array_equal(O.expected_counts(h, g["template"], g["read"], g["ops"], g["radius"]), g["expected_counts"])
    # the vectors themselves satisfy the definitional pins: identity substitution == lk, table rows are likelihoods
    t = g["table"].reshape(-1, 14)
    code = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), g["template"])
    assert np.allclose(t[np.arange(len(code)), code], g["lk"][0], rtol=0, atol=1e-9)
    j, row = len(code) // 2, 5
    edited = O.apply_edit(g["template"], j, row)
    full = O.likelihood(h, edited, g["read"], O.edit_ops(edited, g["read"], 200), 200)
    assert abs(full - t[j, row]) < 1e-6 or g["radius"] < 30  # narrow bands differ from the unbanded value by design


def test_oracle_reproduces_committed_polish():
    p = case("polish")
    n = int(p["n_reads"][0])
    h = O.default_hmm()
    reads = [p[f"read{k}"] for k in range(n)]
    ops = [p[f"ops{k}"] for k in range(n)]
    cons, new_ops, _ = O.polish_until_converge(h, h, p["draft"], reads, ops, p["strands"], 15, n, 0)
    assert np.array_equal(cons, p["consensus"])
    for k in range(n):
        assert np.array_equal(new_ops[k], p[f"new_ops{k}"])
    assert np.array_equal(p["consensus"], p["truth"])  # the three planted draft errors are repaired


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_committed_vectors(name):
    from jtk_b200 import _lib
    g = case(name)
    ctx = _lib.Context()
    try:
        hc = _lib.HmmParams.from_buffer_copy(g["params"].tobytes())
        lk, tabs = ctx.modtable_batch(hc, hc, [g["template"]], [g["read"]], [g["ops"]], [1], np.zeros(1, np.uint32), g["radius"])
        assert abs(lk[0] - g["lk"][0]) <= 2e-5 * abs(g["lk"][0])
        gd, od = tabs[0] - lk[0], g["table"] - g["lk"][0]
        ok = (tabs[0] > -1e9) & (g["table"] > -1e9)
        assert ((tabs[0] > -1e9) == (g["table"] > -1e9))[od > -60].all()
        tol = np.maximum(2e-3, 1e-3 * np.abs(od[ok]))
        assert (np.abs(gd - od)[ok] <= tol).all()
        lkb = ctx.likelihood_batch(hc, hc, [g["template"]], [g["read"]], None, [1], np.zeros(1, np.uint32), g["radius"])
        assert abs(lkb[0] - g["lk_bootstrap"][0]) <= 2e-5 * abs(g["lk_bootstrap"][0])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_polish_matches_committed_consensus():
    from jtk_b200 import hmm
    p = case("polish")
    n = int(p["n_reads"][0])
    models = hmm.PairHiddenMarkovModelOnStrands.default()
    ops = [p[f"ops{k}"].copy() for k in range(n)]
    cons = models.polish_until_converge_antidiagonal(p["draft"], [p[f"read{k}"] for k in range(n)], ops, p["strands"],
                                                     hmm.HMMPolishConfig.new(15, n, 0))
    assert np.array_equal(cons, p["consensus"])
    for k in range(n):
        assert np.array_equal(ops[k], p[f"new_ops{k}"])
