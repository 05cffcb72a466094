#!/usr/bin/env python
"""bench.py -- pair-HMM modification-table throughput of the per-chunk hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A step is one pass of the hot path over one batch: every read of every chunk of BASELINE.json configs[1]
(mock diploid 200 kbp region -> 80 chunks of 2 kbp x 60 ONT-like reads at 8 % error, radius 30) is scored
against its chunk consensus by the banded pair-HMM forward/backward, the 14-row modification table
(table - lk) is written to HBM, and the per-column statistics that `filter_profiles` consumes
(haplotyper/src/local_clustering/pseudo_mcmc.rs:426-474) are reduced on the device.

value   GCUPS with the batch already resident in HBM (kernels only).  One cell update = all three states of
        one in-band DP cell in one direction; a modification-table job is 2*C cell updates;
        checkpoint / reduction work is not credited.
e2e     the same metric through the C ABI with HOST buffers (the call sequence of INTEGRATION.md): jtk_batch_create
        (encode on the host threads + H2D of the step's inputs), jtk_batch_modtable (14 rows),
        jtk_batch_search_variants (filter_profiles on the device, gather of the candidate columns, greedy pick on the
        host, D2H of candidates and values), jtk_batch_fetch_lk, jtk_batch_destroy -- all inside the timed region.
Chunks are independent: rank r processes its own 80 chunks (weak scaling), no collective on the data path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_CELL = 17.0  # 3 mul + 6 FMA + 2 emission mul per cell update
RADIUS = 30
POS_THR = 1e-5
# Gains fixture (expected log-lik gain per (DiffType, homopolymer length)): estimate_gain_default output is an
# input fixture here; rows Subst, Del, Ins; MIN_REQ_FRACTION 0.5 (pseudo_mcmc.rs:140)
GAINS_EXPECTED = np.array([[4.0, 4.0, 4.0], [3.0, 2.0, 1.5], [3.0, 2.0, 1.5]], dtype=np.float32)
GAINS_PROB = np.array([[0.02, 0.02, 0.02], [0.05, 0.08, 0.1], [0.05, 0.08, 0.1]], dtype=np.float64)


def make_workload(rank: int, n_chunks: int, n_reads: int, length: int):
    from jtk_b200 import synth
    chunks = synth.diploid_region(20261017 + rank, n_chunks, length=length, n_reads=n_reads, error_rate=0.08)
    templates = [c["template"] for c in chunks]
    reads = [r for c in chunks for r in c["reads"]]
    ops = [o for c in chunks for o in c["ops"]]
    strands = np.concatenate([c["strands"] for c in chunks])
    tidx = np.repeat(np.arange(n_chunks, dtype=np.uint32), [len(c["reads"]) for c in chunks])
    return templates, reads, ops, strands, tidx


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    The samples are taken in-process through NVML (the library behind nvidia-smi: same counters, same reason bits).  A
    polling `nvidia-smi --query-gpu ... -lms 100` child was measured to stall the GPU for milliseconds per sample (a 150 ms
    timed region read 40 % slow with it, profiles/README.md); JTK_BENCH_SAMPLER=smi selects it anyway, =none disables."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, period_s: float = 0.01):
        self.index = index
        self.period = period_s
        self.lines = []
        self.samples = []          # (sm_mhz, reason bitmask)
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.mode = os.environ.get("JTK_BENCH_SAMPLER", "nvml")
        self.mx = None
        self.nv = None

    def start(self):
        if self.mode == "none":
            return
        if self.mode == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = self.index
                if vis:
                    try:
                        idx = int(vis.split(",")[self.index])
                    except (ValueError, IndexError):
                        idx = self.index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
                self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
                self.nv = pynvml
                self.thread = threading.Thread(target=self._poll, daemon=True)
                self.thread.start()
                return
            except Exception:
                self.nv = None
                self.mode = "smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((sm, rs))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            nv = self.nv
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            reasons = sorted(n for n, b in bits.items() if any(rs & b for _, rs in self.samples))
            sm = [x for x, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                    "samples": len(sm), "source": "nvml in-process, %d ms period" % int(self.period * 1e3)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler unavailable"], "source": self.mode}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 200"}


def ncu_traffic(n_pairs, variant):
    """DRAM bytes per modification-table launch sequence of `variant` ("fused" | "rows") from the committed `ncu --set full`
    captures (profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum and the number of pairs), scaled to
    this launch."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    if variant not in t:
        return None
    return (t[variant]["dram_bytes_read"] + t[variant]["dram_bytes_write"]) * n_pairs / t["pairs_in_launch"]


KERNELS = {"fused": "modtable_fused_kernel<2,14> (pass 1: checkpoints; pass 2: forward rows recomputed per 16-row segment in shared "
                    "memory + backward + table sums) + finalize_kernel",
           "rows": "fwdrows_kernel<2> + bwdtable_kernel<2,14> + finalize_kernel (forward rows parked in HBM; bwdtable is ~65 % of it)"}
TRAFFIC_NOTE = {"fused": "dram__bytes_read.sum + dram__bytes_write.sum of modtable_fused_kernel + finalize_kernel (ncu --set full, "
                         "profiles/r2_traffic.json): checkpoints (65 B per anti-diagonal, written once, read once, partly from "
                         "L2), the raw column sums between the two kernels and the profiles; no DP matrix",
                "rows": "dram__bytes_read.sum + dram__bytes_write.sum of the three kernels (ncu --set full, profiles/r2_traffic.json): "
                        "the forward-row scratch (2 x 2.4 MB per pair: written by fwdrows at 5.8 TB/s, read back by bwdtable), "
                        "not the algorithmic bytes, see roofline.hbm"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def cpu_baseline_run(templates, reads, ops, strands, tidx, n_pairs, threads, min_seconds=0.0):
    """Oracle (f64 restatement, kind 'port') over the chunks that hold the first n_pairs pairs.  Decomposition as in the
    reference: one task per CHUNK on `threads` OS threads (`pileups.into_par_iter()`, local_clustering/mod.rs:64-72), each
    task scoring the reads of its chunk one after the other (pseudo_mcmc.rs:53-67).  Repeated until min_seconds of wall time
    have passed.  Returns (seconds, cell updates)."""
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from jtk_b200 import _lib
    h = O.default_hmm()
    n_pairs = min(n_pairs, len(reads))
    chunk_ids = sorted({int(tidx[k]) for k in range(n_pairs)})
    members = {c: [] for c in chunk_ids}
    for k in range(n_pairs):
        members[int(tidx[k])].append(k)
    per_pass = sum(2 * _lib.band_cell_count(ops[k], len(templates[int(tidx[k])]), len(reads[k]), RADIUS) for k in range(n_pairs))

    def one_chunk(c):
        ks = members[c]
        O.modification_table_batch(h, h, [templates[c]] * len(ks), [reads[k] for k in ks], [ops[k] for k in ks],
                                   [strands[k] for k in ks], RADIUS, n_threads=1, want_tables=True)
    t0 = time.perf_counter()
    cells = 0
    with ThreadPoolExecutor(max_workers=threads) as pool:
        while True:
            list(pool.map(one_chunk, chunk_ids))
            cells += per_pass
            if time.perf_counter() - t0 >= min_seconds:
                break
    return time.perf_counter() - t0, cells


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  kiley (the crate that
    holds the arithmetic) is absent and there is no Rust toolchain, so this is the oracle port (kind 'port')."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_chunks_sample = max(8, threads)  # at least one chunk per thread
    templates, reads, ops, strands, tidx = make_workload(0, n_chunks_sample, args.reads, args.length)
    n_pairs = len(reads)
    for _ in range(args.warmup):
        cpu_baseline_run(templates, reads, ops, strands, tidx, min(n_pairs, 2 * threads), threads)
    t = 0.0
    cells = 0
    for _ in range(args.steps):
        dt, c = cpu_baseline_run(templates, reads, ops, strands, tidx, n_pairs, threads)
        t += dt
        cells += c
    gcups = cells / t / 1e9
    sample = (f"{n_chunks_sample} chunks x {args.reads} reads ({n_pairs} pairs) of the same workload per step, one task per chunk on "
              f"{threads} threads (rayon's decomposition)")
    line = {
        "impl": "reference", "metric": "pair-HMM modification-table GCUPS", "value": gcups, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def phase_leg(ctx, templates, reads, ops, strands, tidx, n_chunks, cov, rank=0, world=1, group=None):
    """'chunks phased per second' (BASELINE.json metric, second half): the whole per-chunk path of
    local_clustering_selected (local_clustering/mod.rs:56-83 without the model fit) on n_chunks chunks -- batched polish,
    9-row tables, device-side filter_profiles, greedy pick, k-means + MCMC (GPU restarts kernel + host threads), posteriors,
    normalisation -- through jtk_b200/pipeline.py.  Wall clock.  Under torchrun every rank calls this with the SAME chunks:
    the scheduler shards them over the ranks and gathers the per-chunk results on rank 0 over the host group (strong
    scaling; the host threads of all ranks share the box's cores)."""
    from jtk_b200 import pipeline as P
    from jtk_b200 import local_clustering as LC
    gains = LC.Gains(gain=GAINS_EXPECTED.astype(np.float64), prob=GAINS_PROB)

    def dataset(lo, hi):
        chunks = [P.Chunk(id=c + 1, seq=templates[c].copy(), copy_num=2) for c in range(lo, hi)]
        nodes = [P.Node(chunk=int(tidx[k]) + 1, seq=reads[k], ops=ops[k], is_forward=bool(strands[k]))
                 for k in range(len(reads)) if lo <= int(tidx[k]) < hi]
        return P.DataSet(selected_chunks=chunks, nodes=nodes, read_type="ONT")

    # warm-up = one full-size call: the pipeline calls local_clustering_selected 4+ times per run (cli/src/pipeline.rs:158,
    # 164-168,175), so the timed call below is the steady state (pinned staging buffers and device pools already sized)
    warm = dataset(0, n_chunks)
    P.local_clustering_selected(warm, {c.id for c in warm.selected_chunks}, gains=gains, ctx=ctx, fit_models=False,
                                rank=rank, world=world, group=group)
    del warm
    ds = dataset(0, n_chunks)
    if world > 1:
        import torch.distributed as dist
        dist.barrier(group=group)
    t0 = time.perf_counter()
    out = P.local_clustering_selected(ds, {c.id for c in ds.selected_chunks}, gains=gains, ctx=ctx, fit_models=False,
                                      rank=rank, world=world, group=group)
    dt = time.perf_counter() - t0
    if out is None:
        return None
    ks = [out[c][2] for c in sorted(out)]
    return {"phases_s": {k: round(v, 4) for k, v in P.LAST_TIMING.items()}, "chunks_per_s": n_chunks / dt, "chunks": n_chunks, "seconds": dt,
            "host_threads": P.host_threads(), "two_cluster_chunks": int(sum(1 for k in ks if k == 2)),
            "what": "polish + 9-row tables + device filter_profiles + pick + k-means/MCMC (GPU restarts + host threads) + normalise; "
                    f"{n_chunks} chunks sharded over {world} GPU(s), host gather on rank 0 (strong scaling); second call on warm "
                    "buffers; phases_s are rank 0's"}


def band_sweep_leg(ctx, fwd, n_chunks=4, n_reads=240, length=2000):
    """BASELINE.json configs[3]: repeat-heavy chunks (240 reads per chunk) with the band swept over the reference's range
    (radius = ceil(frac * L) / 2: 10 .. 100, definitions/src/lib.rs:173-175,201-210): 14-row modification-table GCUPS per
    radius on rank 0's GPU (kernels only), and the copy_num = 8 recursive clustering path of local_clustering/mod.rs:126-190."""
    from jtk_b200 import synth
    chunks = [synth.paralog_chunk(4242 + c, length=length, n_reads=n_reads) for c in range(n_chunks)]
    templates = [c["template"] for c in chunks]
    reads = [r for c in chunks for r in c["reads"]]
    ops = [o for c in chunks for o in c["ops"]]
    strands = np.concatenate([c["strands"] for c in chunks])
    tidx = np.repeat(np.arange(n_chunks, dtype=np.uint32), [len(c["reads"]) for c in chunks])
    out = {"what": f"{n_chunks} chunks x {n_reads} reads (4 paralogs x 2 haplotypes), 14 rows, kernels only", "radius": {}}
    for radius in (10, 20, 30, 50, 100):
        b = ctx.batch(templates, reads, ops, strands, tidx, radius)
        for _ in range(2):
            b.modtable(fwd, fwd, 14)
        b.sync()
        ctx.kernel_times()
        for _ in range(3):
            b.modtable(fwd, fwd, 14)
        b.sync()
        ms = float(np.median(ctx.kernel_times()))
        out["radius"][str(radius)] = {"ms": ms, "gcups": b.cell_updates / (ms * 1e-3) / 1e9}
        b.close()
    return out, (chunks, templates, reads, ops, strands, tidx)


def recursive_leg(ctx, chunks):
    """copy_num = 8 chunks through the driver: clustering_recursive splits into <= 4, re-polishes every sub-cluster and
    recurses (local_clustering/mod.rs:136-189).  Returns chunks/s and the pairwise agreement with the planted paralogs."""
    from jtk_b200 import pipeline as P
    from jtk_b200 import local_clustering as LC
    gains = LC.Gains(gain=GAINS_EXPECTED.astype(np.float64), prob=GAINS_PROB)
    cs = [P.Chunk(id=c + 1, seq=ch["template"].copy(), copy_num=8) for c, ch in enumerate(chunks)]
    nodes = [P.Node(chunk=c + 1, seq=r, ops=o, is_forward=bool(s)) for c, ch in enumerate(chunks)
             for r, o, s in zip(ch["reads"], ch["ops"], ch["strands"])]
    ds = P.DataSet(selected_chunks=cs, nodes=nodes, read_type="ONT")
    t0 = time.perf_counter()
    P.local_clustering_selected(ds, {c.id for c in cs}, gains=gains, ctx=ctx, fit_models=False)
    dt = time.perf_counter() - t0
    rand = []
    for c, ch in enumerate(chunks):
        lab = np.array([n.cluster for n in ds.nodes if n.chunk == c + 1])
        par = np.asarray(ch["paralog"])
        same_l = lab[:, None] == lab[None, :]
        same_p = par[:, None] == par[None, :]
        iu = np.triu_indices(len(lab), 1)
        rand.append(float((same_l[iu] == same_p[iu]).mean()))
    return {"chunks": len(chunks), "seconds": dt, "chunks_per_s": len(chunks) / dt, "cluster_num": [int(c.cluster_num) for c in cs],
            "rand_index_vs_paralogs": rand, "what": "copy_num 8, 240 reads per chunk: clustering_recursive (mod.rs:126-190) on one GPU"}


def polish_fit_leg(ctx, templates, reads, ops, strands, tidx, n_chunks):
    """BASELINE.json configs[4]: consensus polishing over the chunks of the region (drafts with 20 planted substitution errors
    each, HMMPolishConfig(30, n, 0)) and the 10-round fit on 5 chunks (model_tune.rs:94-156); plus the window polish of a
    contig stitched from the first chunks (consensus::polish, consensus/mod.rs:300-371)."""
    from jtk_b200 import consensus as CS
    from jtk_b200 import pipeline as P
    from jtk_b200 import synth
    from jtk_b200.hmm import HMMPolishConfig, PairHiddenMarkovModelOnStrands, polish_chunks
    rng = np.random.default_rng(5)
    models = PairHiddenMarkovModelOnStrands.default()
    n_chunks = min(n_chunks, len(templates))
    sel = [k for k in range(len(reads)) if int(tidx[k]) < n_chunks]
    drafts = []
    for c in range(n_chunks):
        d = templates[c].copy()
        pos = rng.choice(np.arange(5, len(d) - 5), size=20, replace=False)
        d[pos] = synth.ACGT[(np.searchsorted(synth.ACGT, d[pos]) + rng.integers(1, 4, size=20)) % 4]
        drafts.append(d)
    t0 = time.perf_counter()
    cons, _, iters = polish_chunks(models, drafts, [reads[k] for k in sel], [ops[k] for k in sel], [bool(strands[k]) for k in sel],
                                   np.asarray([tidx[k] for k in sel], dtype=np.uint32), HMMPolishConfig.new(RADIUS, 60, 0), ctx=ctx)
    dt = time.perf_counter() - t0
    ident = [float((c == t).mean()) if len(c) == len(t) else 0.0 for c, t in zip(cons, templates[:n_chunks])]
    out = {"polish": {"chunks": n_chunks, "seconds": dt, "chunks_per_s": n_chunks / dt, "iterations_mean": float(np.mean(iters)),
                      "iterations_max": int(np.max(iters)), "identity_mean": float(np.mean(ident)), "identity_min": float(np.min(ident)),
                      "what": "polish_until_converge_antidiagonal over all chunks in one batch, drafts with 20 substitution errors"}}
    # fit: TRAIN_UNIT_SIZE = 5 chunks x TRAIN_ROUND = 10 rounds of (polish + Baum-Welch)
    cs = [P.Chunk(id=c + 1, seq=templates[c].copy(), copy_num=2) for c in range(min(8, n_chunks))]
    nodes = [P.Node(chunk=int(tidx[k]) + 1, seq=reads[k], ops=ops[k], is_forward=bool(strands[k])) for k in sel if int(tidx[k]) < len(cs)]
    ds = P.DataSet(selected_chunks=cs, nodes=nodes, read_type="ONT")
    t0 = time.perf_counter()
    m = P.estimate_model_parameters_on_both_strands(ds, ctx=ctx)
    out["fit"] = {"seconds": time.perf_counter() - t0, "rounds": 10, "chunks": 5, "mat_mat_forward": float(m.forward().as_array()[0]),
                  "what": "estimate_model_parameters_on_both_strands (model_tune.rs:96-156)"}
    # window polish of a contig (3 rounds, 2 000 bp windows, radius 50, coverage cap 50: assemble/mod.rs:186-195)
    n_w = min(16, n_chunks)
    truth = np.concatenate(templates[:n_w])
    draft = np.concatenate(drafts[:n_w])
    alns, off = [], np.concatenate([[0], np.cumsum([len(t) for t in templates[:n_w]])])
    for k in sel:
        c = int(tidx[k])
        if c < n_w:
            alns.append(CS.Alignment(query=reads[k], ops=ops[k], contig_start=int(off[c]), contig_end=int(off[c + 1]), is_forward=bool(strands[k])))
    stats = {}
    t0 = time.perf_counter()
    pol = CS.polish(draft, alns, models, CS.PolishConfig(min_coverage=3, max_coverage=50, window_size=2000, radius=50, round_num=3),
                    ctx=ctx, stats=stats)
    dt = time.perf_counter() - t0
    out["consensus_windows"] = {"windows": n_w, "rounds": stats.get("rounds"), "seconds": dt, "windows_per_s": 3 * n_w / dt,
                                "identity": float((pol == truth).mean()) if len(pol) == len(truth) else 0.0,
                                "what": "consensus::polish (consensus/mod.rs:300-371), all windows of a round in one batch"}
    return out


def workload_config(args):
    return {"workload": f"BASELINE.json configs[1]: mock diploid {args.chunks * 2.5:.0f} kbp region, {args.chunks} chunks x "
                        f"{args.reads} ONT-like reads ({args.length} bp, 8% error), radius {RADIUS}, 14-row table + column stats, per GPU",
            "chunks_per_gpu": args.chunks, "reads_per_chunk": args.reads, "chunk_len": args.length, "radius": RADIUS,
            "rows": 14, "parallelism": "chunks sharded over ranks, no collective",
            "l2": "no explicit flush: each step writes 0.6 GB of raw column sums and 0.5 GB of profiles per GPU, plus 1.2 GB of "
                  "checkpoints (fused variant) or 11.7 GB of forward rows (rows variant), all read back (>> 126 MB L2), so the "
                  "25 MB of inputs are evicted between steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=80)
    ap.add_argument("--reads", type=int, default=60)
    ap.add_argument("--length", type=int, default=2000)
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="pairs in the cpu_baseline sample (0: auto)")
    ap.add_argument("--phase-chunks", type=int, default=2000,
                    help="chunks of the 'chunks phased per second' leg: BASELINE.json configs[2] scale, strong-scaled over the ranks (0: skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[3] / configs[4] legs (band sweep, recursive clustering, polish + fit)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from jtk_b200 import _lib, build
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: jtk_b200 has no CPU fallback")
    build.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    host_group = None
    if world > 1:
        # the only exchange of the path is a HOST gather of per-chunk results: it runs over a gloo group, NCCL
        # carries nothing but the timing reductions of this script
        host_group = dist.new_group(backend="gloo")
        # every rank runs two encoder contexts (e2e leg): share the box's cores instead of 2 x world x all-cores threads;
        # the clustering threads of the phase leg get the rank's full share
        os.environ.setdefault("JTK_HOST_THREADS", str(max(2, (os.cpu_count() or 2) // (2 * world))))
        os.environ.setdefault("JTK_CLUSTER_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    ctx = _lib.Context(local_rank)
    fwd = _lib.HmmParams.from_buffer_copy(_default_params())
    rev = fwd
    min_req = GAINS_EXPECTED * 0.5

    templates, reads, ops, strands, tidx = make_workload(rank, args.chunks, args.reads, args.length)
    packed = _lib.pack_inputs(templates, reads, ops, strands, tidx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg (value) ------------------------------------------------------------
    batch = ctx.batch(templates, reads, ops, strands, tidx, RADIUS)
    cells = batch.cell_updates

    def step_resident(rows=14):
        batch.modtable(fwd, rev, rows)
        batch.colstats(min_req, POS_THR, fetch=False)

    for _ in range(args.warmup):
        step_resident()
    batch.sync()
    ctx.kernel_times()
    sampler = ClockSampler(local_rank)
    if rank == 0 and sampler.mode == "smi":
        sampler.start()          # the nvidia-smi child needs a head start
        time.sleep(0.3)
    barrier()
    l0 = ctx.launch_count
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
    # the steps are queued, the GPU is working through them: sample its clocks now, i.e. DURING the timed region
    if rank == 0 and sampler.mode != "smi":
        sampler.start()
        time.sleep(min(0.25, 0.5 * args.steps * 6e-3))
    dev_ms = ctx.timer_stop()
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count - l0
    ktimes = ctx.kernel_times()
    variant = ctx.last_modtable_variant

    # the other modification-table variant on the same batch (bit-identical tables, tests/test_gpu_parity.py), as an extra
    variants = {}
    if rank == 0:
        saved = os.environ.get("JTK_MODTABLE")
        for v in ("fused", "rows"):
            os.environ["JTK_MODTABLE"] = v
            try:
                for _ in range(2):
                    batch.modtable(fwd, rev, 14)
                batch.sync()
                ctx.kernel_times()
                n_v = max(3, args.steps // 4)
                for _ in range(n_v):
                    batch.modtable(fwd, rev, 14)
                batch.sync()
                kt = ctx.kernel_times()
                ms_v = float(np.mean(kt)) if len(kt) else None
                variants[v] = {"kernel_ms": ms_v, "gcups_kernel": cells / (ms_v * 1e-3) / 1e9 if ms_v else None,
                               "dram_bytes_per_launch": ncu_traffic(len(reads), v)}
            except Exception as exc:  # e.g. not enough device memory for the forward rows of one wave
                variants[v] = {"error": str(exc)}
        if saved is None:
            os.environ.pop("JTK_MODTABLE", None)
        else:
            os.environ["JTK_MODTABLE"] = saved
        variants["default"] = variant
        variants["note"] = ("fused: the DP matrices never touch HBM (checkpoints + recomputation in shared memory); rows: forward rows "
                            "parked in HBM between two kernels (2.4 MB per pair). Same tables bit for bit; the library takes 'rows' "
                            "when the whole batch fits its scratch budget as one wave (JTK_SCRATCH_MB, JTK_MODTABLE overrides)")

    # 9-row (clustering rows only) variant, reported as an extra
    for _ in range(2):
        step_resident(9)
    batch.sync()
    ctx.timer_start()
    for _ in range(max(3, args.steps // 4)):
        step_resident(9)
    ms9 = ctx.timer_stop() / max(3, args.steps // 4)
    ctx.kernel_times()

    # ---- end-to-end leg (host buffers in, statistics out) ---------------------------------------
    cov = args.reads / 2.0

    def step_e2e(c):
        b = _lib.Batch(c, templates, reads, ops, strands, tidx, RADIUS, packed=packed)
        b.modtable(fwd, rev, 14)
        n_probes, probe_pos, variants = b.search_variants(GAINS_EXPECTED.astype(np.float64), GAINS_PROB, 2, cov)
        lk = b.lk()
        h2d = b.h2d_bytes
        b.close()
        return h2d, n_probes.nbytes + probe_pos.nbytes + variants.nbytes + lk.nbytes

    # serial: one host thread, one context, one batch after the other
    for _ in range(2):
        h2d_bytes, d2h_bytes = step_e2e(ctx)
    barrier()
    t1 = time.perf_counter()
    # at least 25 batches per mode (50 pipelined: ~0.4 s): the first encode cannot overlap anything, and with the MAX over the
    # ranks one descheduled thread in a 0.15 s region of 8 ranks x 6 threads on 32 cores halves the figure (seen once at --steps 20)
    e2e_steps = max(25, args.steps)
    for _ in range(e2e_steps):
        step_e2e(ctx)
    barrier()
    e2e_serial_ms = (time.perf_counter() - t1) * 1e3 / e2e_steps
    # pipelined: two host threads with one context each take alternate batches, the way the reference's rayon workers call
    # concurrently (a context is thread-safe per distinct ctx): the encode / copies of one batch overlap the kernels of
    # the other.  Every step still does its own encode, H2D, kernels and D2H inside the timed region.
    from concurrent.futures import ThreadPoolExecutor
    ctx2 = _lib.Context(local_rank)
    workers = [ctx, ctx2]
    with ThreadPoolExecutor(max_workers=2) as pool:
        for _ in range(2):
            list(pool.map(step_e2e, workers))  # warm the second context (its pools and pinned staging reach their size)
        barrier()
        t1 = time.perf_counter()
        futs = [pool.submit(step_e2e, workers[k % 2]) for k in range(2 * e2e_steps)]
        for f in futs:
            f.result()
        barrier()
        e2e_ms = (time.perf_counter() - t1) * 1e3 / (2 * e2e_steps)
    ctx2.close()

    # ---- chunks phased per second through the driver: configs[2] scale, the SAME chunks on every rank, sharded ----------
    phased = None
    if args.phase_chunks > 0:
        w0 = make_workload(0, args.phase_chunks, args.reads, args.length)   # rank-independent seed: every rank holds the region
        phased = phase_leg(ctx, *w0, args.phase_chunks, cov, rank=rank, world=world, group=host_group)
    # ---- configs[3] / configs[4] legs (rank 0's GPU; bounded samples) ------------------------------------------------------
    extras = {}
    if rank == 0 and not args.no_extras:
        sweep, par = band_sweep_leg(ctx, fwd)
        extras["band_sweep"] = sweep
        extras["recursive_clustering"] = recursive_leg(ctx, par[0][:2])
        extras["polish_fit"] = polish_fit_leg(ctx, templates, reads, ops, strands, tidx, args.chunks)
    if world > 1:
        dist.barrier(group=host_group)

    # ---- reduce over ranks ------------------------------------------------------------------------
    step_ms = dev_ms / args.steps
    tot_cells = float(cells)
    if world > 1:
        t = torch.tensor([step_ms, e2e_ms, wall_ms / args.steps, e2e_serial_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms, wall_step, e2e_serial_ms = [float(x) for x in t.tolist()]
        c = torch.tensor([tot_cells], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_cells = float(c.item())
    else:
        wall_step = wall_ms / args.steps
    value = tot_cells / (step_ms * 1e-3) / 1e9
    e2e_value = tot_cells / (e2e_ms * 1e-3) / 1e9

    if rank == 0:
        peaks, peak_src = measured_peaks()
        ffma, ffma2 = ctx.measure_fp32_peak()
        k_ms = float(np.mean(ktimes)) if len(ktimes) else step_ms
        ach_tflops = cells * FLOPS_PER_CELL / (k_ms * 1e-3) / 1e12
        prof_bytes = float(sum((len(templates[int(t)]) + 1) * 14 * 4 for t in tidx))
        in_bytes = float(batch.h2d_bytes)
        roof = {"bound": "fp32", "kernel": KERNELS.get(variant, variant), "variant": variant, "achieved": ach_tflops, "peak": ffma,
                "unit": "TFLOP/s", "frac": ach_tflops / ffma if ffma else None,
                "peak_source": "measured in this run: register-resident fma.rn.f32 loop on all SMs "
                               "(MEASURED_PEAKS.json holds only HBM and bf16 figures)",
                "peak_ffma2": ffma2, "flops_per_cell_update": FLOPS_PER_CELL,
                "cell_updates_per_launch": cells, "kernel_ms": k_ms, "gcups_kernel": cells / (k_ms * 1e-3) / 1e9,
                "traffic": ncu_traffic(len(reads), variant),
                "traffic_note": TRAFFIC_NOTE.get(variant, ""),
                "hbm": {"algorithmic_bytes_per_launch": prof_bytes + in_bytes,
                        "achieved_gbs": (prof_bytes + in_bytes) / (k_ms * 1e-3) / 1e9,
                        "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src}}
        # cpu baseline: bounded sample on this box's host cores, one task per chunk like the reference's rayon fan-out, plus
        # the single-thread figure (the reference's own benches pin num_threads(1), sandbox/src/bin/benchmark_clustering.rs:45-48)
        threads = os.cpu_count() or 1
        n_sample = args.cpu_sample_pairs or min(len(reads), max(threads, 16) * args.reads)
        cpu_s, cpu_cells = cpu_baseline_run(templates, reads, ops, strands, tidx, n_sample, threads, min_seconds=8.0)
        one_s, one_cells = cpu_baseline_run(templates, reads, ops, strands, tidx, args.reads, 1, min_seconds=3.0)
        cpu = {"value": cpu_cells / cpu_s / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
               "sample": f"the chunks of the first {n_sample} pairs of rank 0's workload, repeated for {cpu_s:.1f} s, f64 oracle "
                         f"(scalar restatement, not kiley), one task per chunk on {threads} threads",
               "mcups_per_thread": 1e3 * cpu_cells / cpu_s / 1e9 / threads,
               "single_thread": {"value": one_cells / one_s / 1e9, "unit": "GCUPS", "sample": f"one chunk ({args.reads} pairs), {one_s:.1f} s"}}
        line = {
            "metric": "pair-HMM modification-table GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e2e_ms,
                    "mode": "two host threads, one context each, alternating batches",
                    "serial": {"value": tot_cells / (e2e_serial_ms * 1e-3) / 1e9, "ms_per_step": e2e_serial_ms,
                               "mode": "one host thread, one context"}},
            "gpu_launches": int(launches),
            "chunks_per_s": phased["chunks_per_s"] if phased else None,
            "chunks_per_s_note": (f"chunks phased per second: {phased['chunks']} chunks (configs[2] scale) through the whole per-chunk "
                                  f"path, strong-scaled over {world} GPU(s), wall clock") if phased else None,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "extra": {"chunks_per_s": world * args.chunks / (step_ms * 1e-3),
                      "pairs_per_s": world * len(reads) / (step_ms * 1e-3),
                      "wall_ms_per_step": wall_step,
                      "rows9": {"ms_per_step": ms9, "gcups": cells / (ms9 * 1e-3) / 1e9,
                                "chunks_per_s": args.chunks / (ms9 * 1e-3), "note": "rank 0 only"},
                      "modtable_variants": variants,
                      "chunks_phased": phased, **extras},
        }
        print(json.dumps(line), flush=True)
    batch.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _default_params() -> bytes:
    """HMMParam::default() (definitions/src/lib.rs:128-147) as the 45 doubles of jtk_hmm_params."""
    a = [0.97, 0.01, 0.01] * 3 + [0.97 if r == q else 0.01 for r in range(4) for q in range(4)] + [0.25] * 20
    return np.array(a, dtype=np.float64).tobytes()


if __name__ == "__main__":
    main()
