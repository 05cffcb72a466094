#!/usr/bin/env python
"""tools/phase_scale.py -- chunks phased per second at BASELINE.json configs[2] scale: the whole local_clustering_selected
path (bench.phase_leg) on N synthetic diploid chunks, on one GPU or strong-scaled under torchrun.
  python tools/phase_scale.py --chunks 2000
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/phase_scale.py --chunks 2000"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=800)
    args = ap.parse_args()
    from jtk_b200 import _lib
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
        group = dist.group.WORLD
        os.environ.setdefault("JTK_HOST_THREADS", str(max(2, (os.cpu_count() or 2) // world)))
        os.environ.setdefault("JTK_CLUSTER_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    ctx = _lib.Context(local)
    t0 = time.perf_counter()
    w = bench.make_workload(0, args.chunks, 60, 2000)
    if rank == 0:
        print(f"workload: {args.chunks} chunks in {time.perf_counter() - t0:.1f} s, {os.cpu_count()} cores, world {world}", flush=True)
    out = bench.phase_leg(ctx, *w, args.chunks, 30.0, rank=rank, world=world, group=group)
    if out is not None:
        print({k: v for k, v in out.items() if k != "what"})


if __name__ == "__main__":
    main()
