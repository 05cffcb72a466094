#!/usr/bin/env python
"""tools/phase_scale.py -- chunks phased per second at larger chunk counts (BASELINE.json configs[2] is ~2 000 chunks):
the whole local_clustering_selected path (bench.phase_leg) on N synthetic diploid chunks, one GPU.
  python tools/phase_scale.py --chunks 800"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=800)
    args = ap.parse_args()
    from jtk_b200 import _lib
    ctx = _lib.Context(0)
    t0 = time.perf_counter()
    w = bench.make_workload(0, args.chunks, 60, 2000)
    print(f"workload: {args.chunks} chunks in {time.perf_counter() - t0:.1f} s", flush=True)
    out = bench.phase_leg(ctx, *w, args.chunks, 30.0)
    print({k: v for k, v in out.items() if k != "what"})


if __name__ == "__main__":
    main()
