import cProfile, pstats, sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import bench
from jtk_b200 import _lib
ctx = _lib.Context(0)
w = bench.make_workload(0, 1000, 60, 2000)
bench.phase_leg(ctx, *bench.make_workload(0, 80, 60, 2000), 80, 30.0)
pr = cProfile.Profile(); pr.enable()
out = bench.phase_leg(ctx, *w, 1000, 30.0)
pr.disable()
print(out['phases_s'], out['seconds'])
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
