import sys, time, cProfile, pstats
sys.path.insert(0, '/root/repo')
from jtk_b200 import _lib, likelihood_gains as G, hmm
ctx = _lib.Context(0)
m = hmm.PairHiddenMarkovModelOnStrands.default()
t0 = time.perf_counter(); g = G.estimate_gain_default(m, ctx=ctx); print("first", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); g = G.estimate_gain_default(m, ctx=ctx); print("second", time.perf_counter() - t0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
