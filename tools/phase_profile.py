#!/usr/bin/env python
"""tools/phase_profile.py -- cProfile of bench.phase_leg (chunks phased per second) on one GPU: where the host time goes."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from jtk_b200 import _lib
    ctx = _lib.Context(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    w = bench.make_workload(0, n, 60, 2000)
    bench.phase_leg(ctx, *w, n, 30.0)
    pr = cProfile.Profile()
    pr.enable()
    out = bench.phase_leg(ctx, *w, n, 30.0)
    pr.disable()
    print({k: v for k, v in out.items() if k != "what"})
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()
