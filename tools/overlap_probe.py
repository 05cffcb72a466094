#!/usr/bin/env python
"""tools/overlap_probe.py -- can the forward kernel of one batch hide behind the backward kernel of another?  Two contexts (two
streams) loop jtk_batch_modtable on their own copy of the bench batch from two host threads; JTK_GRID_FWD / JTK_GRID_BWD cap the
CTAs per SM of the two kernels so that both fit an SM together.  Prints the aggregate ms per 4 800-pair batch."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from jtk_b200 import _lib
    fwd = _lib.HmmParams.from_buffer_copy(bench._default_params())
    w = bench.make_workload(0, 80, 60, 2000)
    n_ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    reps = 20
    ctxs = [_lib.Context(0) for _ in range(n_ctx)]
    batches = [c.batch(*w, 30) for c in ctxs]
    for b in batches:
        for _ in range(2):
            b.modtable(fwd, fwd, 14)
        b.sync()

    def loop(b):
        for _ in range(reps):
            b.modtable(fwd, fwd, 14)
        b.sync()
    t0 = time.perf_counter()
    th = [threading.Thread(target=loop, args=(b,)) for b in batches]
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"contexts={n_ctx} JTK_GRID_FWD={os.environ.get('JTK_GRID_FWD')} JTK_GRID_BWD={os.environ.get('JTK_GRID_BWD')}: "
          f"{dt * 1e3 / (reps * n_ctx):.3f} ms per batch of {len(w[1])} pairs")


if __name__ == "__main__":
    main()
