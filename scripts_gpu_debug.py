# scratch: first GPU contact -- error breakdown per row group and a rough timing
import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, oracle_lib as O
from jtk_b200 import _lib, synth
ctx = _lib.Context()
def to_c(h): return _lib.HmmParams.from_buffer_copy(bytes(h))
h = O.default_hmm()
for (L, err, R) in [(40, 0.15, 3), (150, 0.1, 14), (300, 0.1, 30), (2000, 0.08, 30), (120, 0.1, 50)]:
    rng = np.random.default_rng(L)
    t = synth.random_template(rng, L)
    reads, ops = zip(*[synth.mutate_read(rng, t, err) for _ in range(6)])
    lk, tabs = ctx.modtable_batch(to_c(h), to_c(h), [t], list(reads), list(ops), np.ones(6, np.uint8), np.zeros(6, np.uint32), R)
    otabs, olk = O.modification_table_batch(h, h, [t] * 6, list(reads), list(ops), np.ones(6, np.uint8), R, n_threads=4)
    print(f"L={L} R={R} lk gpu {lk[:2]} orc {olk[:2]}")
    for k in range(6):
        g = tabs[k].reshape(-1, 14) - lk[k]; o = otabs[k].reshape(-1, 14) - olk[k]
        ok = (tabs[k].reshape(-1, 14) > -1e9) & (otabs[k].reshape(-1, 14) > -1e9)
        mism = ((tabs[k].reshape(-1, 14) > -1e9) != (otabs[k].reshape(-1, 14) > -1e9)).sum()
        e = np.abs(g - o) * ok
        grp = [e[:, 0:4].max(), e[:, 4:8].max(), e[:, 8:11].max(), e[:, 11:14].max()]
        print("  pair", k, "max err sub/ins/copy/del", ["%.2e" % x for x in grp], "neg-mismatch", mism, "argmax", np.unravel_index(e.argmax(), e.shape))
# timing: config1-like batch
chunks = synth.diploid_region(1, 8)
templates = [c['template'] for c in chunks]
reads = [r for c in chunks for r in c['reads']]; ops = [o for c in chunks for o in c['ops']]
strands = np.concatenate([c['strands'] for c in chunks]); tidx = np.repeat(np.arange(8), 60).astype(np.uint32)
for it in range(3):
    t0 = time.time()
    lk, _ = ctx.modtable_batch(to_c(h), to_c(h), templates, reads, ops, strands, tidx, 30, want_table=False)
    cells = sum(2 * _lib.band_cell_count(ops[k], 2000, len(reads[k]), 30) for k in range(len(reads)))
    print("batch 480 pairs: wall %.1f ms kernel %.2f ms -> %.1f GCUPS (kernel)" % ((time.time() - t0) * 1e3, ctx.last_kernel_ms, cells / ctx.last_kernel_ms / 1e6))
