#!/usr/bin/env python
"""tools/e2e_phases.py -- wall-clock phases of one end-to-end step (not a bench): create (encode + H2D), modtable launch,
search_variants (device filter + gather + host pick + D2H), fetch lk, destroy."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from jtk_b200 import _lib
    ctx = _lib.Context(0)
    fwd = _lib.HmmParams.from_buffer_copy(bench._default_params())
    templates, reads, ops, strands, tidx = bench.make_workload(0, 80, 60, 2000)
    packed = _lib.pack_inputs(templates, reads, ops, strands, tidx)
    acc = {}
    for it in range(8):
        t = [time.perf_counter()]
        b = _lib.Batch(ctx, templates, reads, ops, strands, tidx, bench.RADIUS, packed=packed); t.append(time.perf_counter())
        b.modtable(fwd, fwd, 14); t.append(time.perf_counter())
        b.sync(); t.append(time.perf_counter())
        r = b.search_variants(bench.GAINS_EXPECTED.astype(np.float64), bench.GAINS_PROB, 2, 30.0); t.append(time.perf_counter())
        lk = b.lk(); t.append(time.perf_counter())
        b.close(); t.append(time.perf_counter())
        if it >= 2:
            for name, d in zip(["create", "modtable_launch", "kernels_sync", "search_variants", "fetch_lk", "destroy"], np.diff(t)):
                acc.setdefault(name, []).append(d * 1e3)
    for k, v in acc.items():
        print(f"{k:16s} {np.median(v):7.3f} ms")
    print("total", sum(np.median(v) for v in acc.values()))
    ctx.close()


if __name__ == "__main__":
    main()
