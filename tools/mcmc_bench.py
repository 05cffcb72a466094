#!/usr/bin/env python
"""tools/mcmc_bench.py -- jtk_mcmc_restarts_batch (GPU, one warp per chain) against the host twin on N synthetic diploid
variant matrices (60 reads x 6 columns, 20 restarts each).  python tools/mcmc_bench.py --chains 640"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=640)
    ap.add_argument("--host", type=int, default=4, help="chains also run on the host twin (one thread) for comparison")
    ap.add_argument("--restarts", type=int, default=20)
    ap.add_argument("--cols", type=int, default=6, help="variant columns per chain (the bench's diploid chunks have 1-2)")
    ap.add_argument("--zero-reads", type=float, default=0.0, help="fraction of reads without a call in any column")
    args = ap.parse_args()
    from jtk_b200 import _lib, pipeline as P
    from test_gpu_clustering import _host_restarts
    ctx = _lib.Context(0)
    rng = np.random.default_rng(3)
    datas, states = [], []
    for c in range(args.chains):
        hap = rng.integers(0, 2, 60)
        D = args.cols
        v = np.where(hap[:, None] == 1, rng.normal(8, 2, (60, D)), -rng.normal(8, 2, (60, D)))
        v[rng.random((60, D)) < 0.1] = 0.0
        v[rng.random(60) < args.zero_reads] = 0.0
        datas.append(v); states.append(P._rng_seed(3490 * (c + 1)))
    ks, covs = [2] * args.chains, [30.0] * args.chains
    ctx.mcmc_restarts(datas[:4], ks[:4], covs[:4], np.array(states[:4]), 1)
    t0 = time.perf_counter()
    asn, lk, err, st = ctx.mcmc_restarts(datas, ks, covs, np.array(states), args.restarts)
    dt = time.perf_counter() - t0
    print(f"gpu: {args.chains} chains x {args.restarts} restarts in {dt:.3f} s = {args.chains / dt:.0f} chains/s, status ok: {(err == 0).all()}")
    t0 = time.perf_counter()
    same = True
    for c in range(min(args.host, args.chains)):
        ha, hlk, hst = _host_restarts(datas[c], 2, 30.0, states[c], args.restarts)
        same &= np.array_equal(ha, asn[c]) and hlk == lk[c] and np.array_equal(hst, st[c])
    dh = (time.perf_counter() - t0) / max(1, min(args.host, args.chains))
    print(f"host twin: {dh:.3f} s per chain on one thread; identical to the device on {min(args.host, args.chains)} chains: {same}")


if __name__ == "__main__":
    main()
