"""Timing of the calibrations of haplotyper::likelihood_gains on one GPU: estimate_gain_default (1.8e5
likelihood_antidiagonal_bootstrap calls, likelihood_gains.rs:253-315) and estimate_minimum_gain (1e6 calls, :6-39).
usage: python tools/calib_time.py [--min-gain]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jtk_b200 import _lib, likelihood_gains as G, hmm
ctx = _lib.Context(0)
m = hmm.PairHiddenMarkovModelOnStrands.default()
for name in ("first", "second", "third"):
    n0 = ctx.launch_count
    t0 = time.perf_counter(); g = G.estimate_gain_default(m, ctx=ctx); dt = time.perf_counter() - t0
    print(f"estimate_gain_default {name}: {dt:.3f} s, {ctx.launch_count - n0} kernel launches, kernels {sum(ctx.kernel_times()):.2f} ms", flush=True)
print("gain", np.round(g.gain, 3).tolist(), "prob", np.round(g.prob, 4).tolist())
for unpacked in ("", "1"):
    if unpacked: os.environ["JTK_LIKELIHOOD_UNPACKED"] = "1"
    t0 = time.perf_counter(); g2 = G.estimate_gain_default(m, ctx=ctx); dt = time.perf_counter() - t0
    print(f"  {'one pair per warp' if unpacked else 'two pairs per warp'}: {dt:.3f} s, kernels {sum(ctx.kernel_times()):.2f} ms", flush=True)
    assert np.allclose(g2.gain, g.gain, rtol=1e-9) and np.allclose(g2.prob, g.prob, rtol=1e-9)
os.environ.pop("JTK_LIKELIHOOD_UNPACKED", None)
if "--min-gain" in sys.argv:
    t0 = time.perf_counter(); v = G.estimate_minimum_gain(m, ctx=ctx); dt = time.perf_counter() - t0
    print(f"estimate_minimum_gain (1e6 pairs): {dt:.3f} s -> {v:.4f}, kernels {sum(ctx.kernel_times()):.2f} ms")
