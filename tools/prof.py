#!/usr/bin/env python
"""tools/prof.py -- minimal driver for kernel timing / ncu captures of modtable_kernel (not a bench: prints kernel ms).

  python tools/prof.py [--chunks 80] [--rows 14] [--reps 5] [--check]
"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=80)
    ap.add_argument("--reads", type=int, default=60)
    ap.add_argument("--length", type=int, default=2000)
    ap.add_argument("--rows", type=int, default=14)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--radius", type=int, default=30)
    args = ap.parse_args()
    from jtk_b200 import _lib
    ctx = _lib.Context(0)
    fwd = _lib.HmmParams.from_buffer_copy(bench._default_params())
    templates, reads, ops, strands, tidx = bench.make_workload(0, args.chunks, args.reads, args.length)
    batch = ctx.batch(templates, reads, ops, strands, tidx, args.radius)
    cells = batch.cell_updates
    for _ in range(2):
        batch.modtable(fwd, fwd, args.rows)
    batch.sync()
    ctx.kernel_times()
    for _ in range(args.reps):
        batch.modtable(fwd, fwd, args.rows)
    batch.sync()
    kt = np.array(ctx.kernel_times())
    ms = float(np.median(kt))
    print(f"rows={args.rows} pairs={len(reads)} kernel_ms median={ms:.3f} min={kt.min():.3f} GCUPS={cells / (ms * 1e-3) / 1e9:.1f}", flush=True)
    batch.close()
    ctx.close()


if __name__ == "__main__":
    main()
