#!/usr/bin/env python
"""tools/cmp_modes.py -- the fused modification-table kernel against the three-kernel path (JTK_MODTABLE_LEGACY) and the oracle
on one full-size chunk: prints the worst differences and where they sit."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from jtk_b200 import _lib, synth


def main():
    ctx = _lib.Context(0)
    d = synth.diploid_chunk(7, length=2000, n_reads=60, error_rate=0.08, n_snv=5)
    h = O.default_hmm()
    hc = _lib.HmmParams.from_buffer_copy(bytes(h))
    n = 60
    tidx = np.zeros(n, np.uint32)
    out = {}
    for mode in ("fused", "legacy"):
        if mode == "legacy": os.environ["JTK_MODTABLE_LEGACY"] = "1"
        else: os.environ.pop("JTK_MODTABLE_LEGACY", None)
        lk, tabs = ctx.modtable_batch(hc, hc, [d["template"]], d["reads"], d["ops"], d["strands"], tidx, 30)
        out[mode] = (np.asarray(lk), [np.asarray(t) for t in tabs])
    otabs, olk = O.modification_table_batch(h, h, [d["template"]] * n, d["reads"], d["ops"], d["strands"], 30, n_threads=8)
    for a, b in (("fused", "legacy"), ("fused", "oracle"), ("legacy", "oracle")):
        la, ta = out[a]
        lb, tb = (np.asarray(olk), [np.asarray(t) for t in otabs]) if b == "oracle" else out[b]
        worst = (0.0, None)
        for k in range(n):
            x = ta[k] - la[k]; y = tb[k] - lb[k]
            ok = (ta[k] > -1e9) & (tb[k] > -1e9)
            dd = np.abs(x - y) * ok
            i = int(np.argmax(dd))
            if dd.flat[i] > worst[0]: worst = (float(dd.flat[i]), (k, i // 14, i % 14, float(x.flat[i]), float(y.flat[i])))
        print(f"{a} vs {b}: max |dlk| = {np.max(np.abs(la - lb)):.3e}  worst table entry {worst[0]:.3e} at (pair, col, row, a, b) = {worst[1]}")


if __name__ == "__main__":
    main()
