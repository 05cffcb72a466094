import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, oracle_lib as O
from jtk_b200 import _lib, synth
ctx = _lib.Context()
def to_c(h): return _lib.HmmParams.from_buffer_copy(bytes(h))
fwd = rev = O.default_hmm()
cases = []
for t, q in ((b"ACGTACGTAC", b"AC"), (b"AC", b"ACGTTTTTGA"), (b"A", b"A"), (b"A", b"C"), (b"ACGTTGCA", b"ACGTTGCA")):
    t = np.frombuffer(t, np.uint8); q = np.frombuffer(q, np.uint8)
    cases.append((t, q, O.edit_ops(t, q, 20)))
rng = np.random.default_rng(5)
t = synth.random_template(rng, 400)
ins = synth.random_template(rng, 25)
q = np.concatenate([t[:200], ins, t[200:]])
o = np.concatenate([np.zeros(200, np.uint8), np.full(25, 2, np.uint8), np.zeros(200, np.uint8)])
cases.append((t, q, o))
q2 = np.concatenate([t[:150], t[180:]])
o2 = np.concatenate([np.zeros(150, np.uint8), np.full(30, 3, np.uint8), np.zeros(220, np.uint8)])
cases.append((t, q2, o2))
templates = [c[0] for c in cases]; reads = [c[1] for c in cases]; ops = [c[2] for c in cases]
n = len(cases)
lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), templates, reads, ops, np.ones(n, np.uint8), np.arange(n), 30)
otabs, olk = O.modification_table_batch(fwd, rev, templates, reads, ops, np.ones(n, np.uint8), 30)
for k in range(n):
    g = tabs[k].reshape(-1, 14); o = otabs[k].reshape(-1, 14)
    gneg, oneg = g < -1e9, o < -1e9
    dis = gneg != oneg
    print("case", k, "lk", lk[k], olk[k], "neg mismatches", int(dis.sum()))
    for (j, r) in np.argwhere(dis)[:12]:
        print("   j", j, "row", r, "gpu", g[j, r] - lk[k], "orc", o[j, r] - olk[k])
    ok = ~gneg & ~oneg
    e = np.abs((g - lk[k]) - (o - olk[k])) * ok
    print("   max err", e.max(), np.unravel_index(e.argmax(), e.shape))
