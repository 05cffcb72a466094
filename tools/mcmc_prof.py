#!/usr/bin/env python
"""tools/mcmc_prof.py -- a short jtk_mcmc_restarts_batch run for ncu captures (diploid chains, 60 reads x 6 columns).
  python tools/mcmc_prof.py --chains 592 --restarts 1"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=592)
    ap.add_argument("--restarts", type=int, default=1)
    ap.add_argument("--cols", type=int, default=6)
    args = ap.parse_args()
    from jtk_b200 import _lib, pipeline as P
    ctx = _lib.Context(0)
    rng = np.random.default_rng(3)
    datas, states = [], []
    for c in range(args.chains):
        hap = rng.integers(0, 2, 60)
        v = np.where(hap[:, None] == 1, rng.normal(8, 2, (60, args.cols)), -rng.normal(8, 2, (60, args.cols)))
        v[rng.random((60, args.cols)) < 0.1] = 0.0
        datas.append(v); states.append(P._rng_seed(3490 * (c + 1)))
    asn, lk, err, st = ctx.mcmc_restarts(datas, [2] * args.chains, [30.0] * args.chains, np.array(states), args.restarts)
    print("ok", (err == 0).all())


if __name__ == "__main__":
    main()
