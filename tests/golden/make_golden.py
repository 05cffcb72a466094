#!/usr/bin/env python
"""Writes tests/golden/phmm_golden.npz: small seeded inputs with the f64 oracle's outputs (likelihood, 14-row
modification table, bootstrap likelihood, expected counts, polished consensus).

The reference has no golden vector for this path and kiley cannot be built here, so these vectors pin the
ORACLE: `tests/test_golden.py` checks that oracle/phmm_oracle.c still reproduces them bit for bit (a change of the
restatement shows up as a diff of this file, not as a silent move of both sides of the parity tests), and the GPU test
compares the CUDA path with the committed numbers rather than with whatever the oracle computes today.

    python tests/golden/make_golden.py        # regenerate (only when the oracle's definition changes on purpose)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
from jtk_b200 import synth  # noqa: E402


def random_hmm(seed):
    rng = np.random.default_rng(seed)
    a = np.empty(45)
    for s in range(3):
        a[3 * s:3 * s + 3] = rng.dirichlet([30, 2, 2]) * 0.99
    for r in range(4):
        e = rng.dirichlet([1, 1, 1, 1]) * 0.2
        e[r] += 0.8
        a[9 + 4 * r:13 + 4 * r] = e
    for c in range(5):
        a[25 + 4 * c:29 + 4 * c] = rng.dirichlet([3, 3, 3, 3])
    return a


CASES = [  # (name, template length, error rate, radius, model seed or None for HMMParam::default())
    ("default_r10", 120, 0.10, 10, None),
    ("random_r30", 200, 0.08, 30, 11),
    ("random_r5_noisy", 80, 0.20, 5, 12),
    ("default_r50", 150, 0.05, 50, None),
]


def main():
    out = {}
    for name, L, err, R, ms in CASES:
        rng = np.random.default_rng(sum(map(ord, name)))
        h = O.default_hmm() if ms is None else O.OrcHmm.from_array(random_hmm(ms))
        t = synth.random_template(rng, L)
        q, ops = synth.mutate_read(rng, t, err)
        table, lk = O.modification_table(h, t, q, ops, R)
        out[f"{name}/params"] = h.as_array()
        out[f"{name}/template"] = t
        out[f"{name}/read"] = q
        out[f"{name}/ops"] = ops
        out[f"{name}/radius"] = np.array([R])
        out[f"{name}/lk"] = np.array([lk])
        out[f"{name}/table"] = np.asarray(table, dtype=np.float64)
        out[f"{name}/lk_bootstrap"] = np.array([O.likelihood_bootstrap(h, t, q, R)])
        out[f"{name}/expected_counts"] = O.expected_counts(h, t, q, ops, R)
    # one polish case: a draft with three errors, 12 reads
    rng = np.random.default_rng(77)
    h = O.default_hmm()
    truth = synth.random_template(rng, 150)
    draft = truth.copy()
    draft[40] = synth.ACGT[(np.searchsorted(synth.ACGT, draft[40]) + 1) % 4]
    draft = np.delete(draft, 90)
    draft = np.insert(draft, 120, synth.ACGT[2])
    reads, ops = [], []
    for _ in range(12):
        q, _o = synth.mutate_read(rng, truth, 0.06)
        reads.append(q)
        ops.append(O.edit_ops(draft, q, 30))
    strands = (rng.random(12) < 0.5).astype(np.uint8)
    cons, new_ops, iters = O.polish_until_converge(h, h, draft, reads, ops, strands, 15, 12, 0)
    out["polish/draft"] = draft
    out["polish/truth"] = truth
    out["polish/strands"] = strands
    out["polish/consensus"] = np.asarray(cons, dtype=np.uint8)
    out["polish/n_reads"] = np.array([12])
    for k in range(12):
        out[f"polish/read{k}"] = reads[k]
        out[f"polish/ops{k}"] = np.asarray(ops[k], dtype=np.uint8)
        out[f"polish/new_ops{k}"] = np.asarray(new_ops[k], dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "phmm_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "phmm_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
