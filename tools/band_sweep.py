"""Band sweep of bench.py (configs[3]) alone: python tools/band_sweep.py  (JTK_MODTABLE=rows|fused to force a variant)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from jtk_b200 import _lib
ctx = _lib.Context(0)
fwd = _lib.HmmParams.from_buffer_copy(bench._default_params())
sweep, _ = bench.band_sweep_leg(ctx, fwd)
print(json.dumps({k: (round(v["ms"], 3), round(v["gcups"], 1)) for k, v in sweep["radius"].items()}), "variant", ctx.last_modtable_variant if hasattr(ctx, "last_modtable_variant") else "?")
