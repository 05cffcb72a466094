"""Mirror of the per-chunk driver of haplotyper::local_clustering and of the HMM fit loop.

Reference (all under /root/reference/haplotyper/src):
  local_clustering/mod.rs:23-83    LocalClustering::{local_clustering, local_clustering_selected}, pileup_nodes
  local_clustering/mod.rs:86-123   clustering_on_pileup
  local_clustering/mod.rs:126-260  clustering_recursive, estim_copy_num, filter_sub_clusters, update_by_clusterings
  local_clustering/normalize.rs    normalize_local_clustering, reorder
  model_tune.rs:94-156             estimate_model_parameters_on_both_strands (5 chunks x 10 rounds of polish + EM)
  misc.rs:394-407                  update_coverage
  definitions/src/lib.rs:173-210   ReadType::band_width

Same names and argument meaning; the difference is the execution shape: where the reference runs one chunk per rayon
task and one kiley call per read, this driver hands ALL selected pile-ups to the GPU at once (one polish batch, one
9-row modification-table batch, one device-side filter_profiles), keeps the host loops (greedy pick, k-means, MCMC)
per chunk, and can shard the chunks over ranks (scheduler.py).  Data model: `Chunk`, `Node`, `DataSet` carry exactly the
fields the hot path reads and writes (definitions/src/lib.rs); `Node.ops` is the per-column op vector
(misc::ops_to_kiley of `Node.cigar`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from . import _lib, scheduler
from .hmm import HMMPolishConfig, PairHiddenMarkovModelOnStrands, default_context, fit_antidiagonal_par_multiple, polish_chunks
from .local_clustering import ClusteringConfig, ClusteringResult, Gains, _bind, _CConfig
import ctypes as C

UPPER_COPY_NUM = 8   # local_clustering/mod.rs:85
BRANCH_NUM = 4       # local_clustering/mod.rs:139
TRAIN_UNIT_SIZE = 5  # model_tune.rs:94
TRAIN_ROUND = 10     # model_tune.rs:95
BAND_FRAC = {"CCS": 0.01, "CLR": 0.05, "ONT": 0.03, "None": 0.05}  # definitions/src/lib.rs:173-175,201-210


def band_width(read_type: str, length: int) -> int:
    """ReadType::band_width (definitions/src/lib.rs:201-210)."""
    return int(math.ceil(length * BAND_FRAC[read_type]))


@dataclass
class Chunk:
    id: int
    seq: np.ndarray          # uint8 ASCII
    copy_num: int
    cluster_num: int = 1
    score: float = 0.0


@dataclass
class Node:
    chunk: int
    seq: np.ndarray          # uint8 ASCII, already in chunk orientation (definitions/src/lib.rs:678)
    ops: np.ndarray          # uint8 per alignment column: 0 Match, 1 Mismatch, 2 Ins, 3 Del (global over the chunk)
    is_forward: bool
    cluster: int = 0
    posterior: np.ndarray = field(default_factory=lambda: np.zeros(1))


@dataclass
class DataSet:
    selected_chunks: List[Chunk]
    nodes: List[Node]        # encoded_reads.iter().flat_map(|r| r.nodes) in read order
    read_type: str = "ONT"
    coverage: Optional[float] = None
    coverage_protected: bool = False
    model: Optional[PairHiddenMarkovModelOnStrands] = None


def update_coverage(ds: DataSet) -> None:
    """misc::update_coverage (misc.rs:394-407): haploid coverage = median node count per chunk / 2."""
    if ds.coverage_protected and ds.coverage is not None:
        return
    counts: Dict[int, int] = {}
    for n in ds.nodes:
        counts[n.chunk] = counts.get(n.chunk, 0) + 1
    cs = sorted(counts.values())
    ds.coverage = cs[len(cs) // 2] / 2.0


def nonmatch_columns(node: Node, chunk: Chunk) -> int:
    """Sort key of pileup_nodes (mod.rs:47-50): alignment columns of Node::recover that are not '|'
    (jtk_lc_nonmatch_columns; the numpy version of this loop was 17 % of a 80-chunk local_clustering_selected call)."""
    L = _bind()
    ops, q, t = _lib._u8(node.ops), _lib._u8(node.seq), _lib._u8(chunk.seq)
    f = L.jtk_lc_nonmatch_columns
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    r = f(_lib._ptr(ops), len(ops), _lib._ptr(q), len(q), _lib._ptr(t), len(t))
    if r < 0:
        raise ValueError("node ops do not span (read, chunk)")
    return int(r)


def group_nodes(ds: DataSet, selection: Set[int]) -> Dict[int, Tuple[List[Node], Chunk]]:
    """The grouping half of pileup_nodes (mod.rs:33-46): nodes per selected chunk in read order, not yet sorted."""
    pile: Dict[int, Tuple[List[Node], Chunk]] = {c.id: ([], c) for c in ds.selected_chunks if c.id in selection}
    for n in ds.nodes:
        if n.chunk in pile:
            pile[n.chunk][0].append(n)
    return pile


def pileup_nodes(ds: DataSet, selection: Set[int], pile: Optional[Dict[int, Tuple[List[Node], Chunk]]] = None,
                 only: Optional[Iterable[int]] = None) -> Dict[int, Tuple[List[Node], Chunk]]:
    """mod.rs:33-53: nodes per selected chunk, cleanest alignments first (stable sort, as sort_by_cached_key).  The sort
    keys of all nodes come from ONE call (jtk_lc_nonmatch_columns_batch).  `pile` / `only`: sort just these chunks of an
    existing grouping (a rank sorts the pile-ups it will process, not everybody's)."""
    if pile is None:
        pile = group_nodes(ds, selection)
    wanted = set(only) if only is not None else None
    cids = [cid for cid, (nodes, _) in pile.items() if nodes and (wanted is None or cid in wanted)]
    if not cids:
        return pile
    flat = [n for cid in cids for n in pile[cid][0]]
    tidx = np.repeat(np.arange(len(cids), dtype=np.uint32), [len(pile[cid][0]) for cid in cids])

    def cat64(seqs):
        c, off = _lib.concat(seqs)
        return c, off.astype(np.uint64)
    ocat, ooff = cat64([n.ops for n in flat])
    rcat, roff = cat64([n.seq for n in flat])
    tcat, toff = cat64([pile[cid][1].seq for cid in cids])
    keys = np.zeros(len(flat), dtype=np.int32)
    L = _bind()
    vp = C.c_void_p
    L.jtk_lc_nonmatch_columns_batch.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    p = _lib._ptr
    rc = L.jtk_lc_nonmatch_columns_batch(len(flat), p(ocat), p(ooff), p(rcat), p(roff), p(tcat), p(toff), p(tidx), p(keys))
    if rc != 0 or (keys < 0).any():
        raise ValueError("node ops do not span (read, chunk)")
    pos = 0
    for cid in cids:
        nodes = pile[cid][0]
        k = keys[pos:pos + len(nodes)]
        pos += len(nodes)
        order = np.argsort(k, kind="stable")
        nodes[:] = [nodes[i] for i in order]
    return pile


def estim_copy_num(asn: Sequence[int], k: int, copy_num: int, coverage: float) -> List[int]:
    """mod.rs:223-242 (max_by keeps the last of equal maxima)."""
    assert k <= copy_num, (k, copy_num)
    counts = [0.0] * k
    for x in asn:
        counts[int(x)] += 1.0
    cps = [1] * k
    for _ in range(k, copy_num):
        best, arg = None, 0
        for i in range(k):
            v = (counts[i] - coverage * cps[i]) ** 2
            if best is None or not (v < best):
                best, arg = v, i
        cps[arg] += 1
    assert sum(cps) == copy_num
    return cps


def reorder(xs, indices) -> None:
    """normalize.rs:54-63 (in place)."""
    for i in range(len(xs)):
        while int(indices[i]) != i:
            to = int(indices[i])
            xs[i], xs[to] = xs[to], xs[i]
            indices[i], indices[to] = indices[to], indices[i]


def normalize_local_clustering(ds: DataSet) -> None:
    """normalize.rs:6-51: relabel the clusters of every chunk by descending size, permute the posteriors."""
    cluster_num = {c.id: c.cluster_num for c in ds.selected_chunks}
    pile: Dict[int, List[Node]] = {}
    for n in ds.nodes:
        pile.setdefault(n.chunk, []).append(n)
    for cid, nodes in pile.items():
        if cid not in cluster_num:
            continue
        mx = cluster_num[cid]
        for n in nodes:
            assert len(n.posterior) == mx, (cid, len(n.posterior), mx)
        counts = [[c, 0] for c in range(mx)]
        for n in nodes:
            counts[int(n.cluster)][1] += 1
        counts.sort(key=lambda x: x[1])   # stable, then reversed: as sort_by_key + reverse
        counts.reverse()
        mapsto = [0] * mx
        for to, (frm, _) in enumerate(counts):
            mapsto[frm] = to
        if mapsto == list(range(mx)):
            continue   # already in descending size order: `reorder` with the identity changes nothing
        # `reorder(&mut posterior, &mut mapsto.clone())` (normalize.rs:47-49,54-63) moves entry c to position mapsto[c]: one
        # scatter for the whole pile-up instead of a swap loop per node
        post = np.stack([n.posterior for n in nodes])
        moved = np.empty_like(post)
        moved[:, mapsto] = post
        for i, n in enumerate(nodes):
            n.cluster = mapsto[int(n.cluster)]
            n.posterior = moved[i]


# ---------------------------------------------------------------------------------------------------------------------
def _rng_seed(seed: int) -> np.ndarray:
    st = np.zeros(4, dtype=np.uint64)
    L = _bind()
    L.jtk_lc_rng_seed.argtypes = [C.c_uint64, C.c_void_p]
    L.jtk_lc_rng_seed.restype = None
    L.jtk_lc_rng_seed(seed, _lib._ptr(st))
    return st


def _clustering_variants_rng(variants, probe_pos, template, config: ClusteringConfig, state: np.ndarray) -> ClusteringResult:
    """jtk_lc_clustering_variants_rng: pseudo_mcmc::clustering after search_variants, advancing the caller's generator."""
    L = _bind()
    vp = C.c_void_p
    L.jtk_lc_clustering_variants_rng.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_void_p, C.POINTER(_CConfig),
                                                 vp, vp, vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    v = np.ascontiguousarray(variants, dtype=np.float64)
    pp = np.ascontiguousarray(probe_pos, dtype=np.uint32)
    t = _lib._u8(template)
    n, stride = v.shape
    pstride = max(config.copy_num, 1)
    asn = np.zeros(n, dtype=np.uint64)
    post = np.zeros(n * pstride, dtype=np.float64)
    score, k = C.c_double(), C.c_int()
    g, c = config.gains.to_c(), config.to_c()
    rc = L.jtk_lc_clustering_variants_rng(_lib._ptr(v), n, len(pp), stride, _lib._ptr(pp), _lib._ptr(t), len(t), C.byref(g),
                                          C.byref(c), _lib._ptr(state), _lib._ptr(asn), _lib._ptr(post), pstride,
                                          C.byref(score), C.byref(k))
    if rc != 0:
        raise _lib.JtkError(rc, L.jtk_lc_last_error().decode())
    kk = int(k.value)
    return ClusteringResult(asn, post.reshape(n, pstride)[:, :kk].copy(), float(score.value), kk, pp.copy())


def _clustering_variants_batch_gpu(ctx, jobs) -> List[ClusteringResult]:
    """jtk_lc_clustering_variants_batch: the per-chunk `_clustering_variants_rng` for many chunks, with the 20 k-means +
    MCMC restarts of every chunk as one warp each on the GPU.  jobs: list of (variants[n, D], probe_pos[D],
    template, config, state[4]); the states are advanced in place exactly as by the per-chunk call."""
    L = _bind()
    vp = C.c_void_p
    L.jtk_lc_clustering_variants_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp]
    m = len(jobs)
    if m == 0:
        return []
    vs = [np.ascontiguousarray(j[0], dtype=np.float64).reshape(len(j[0]), -1) for j in jobs]
    pps = [np.ascontiguousarray(j[1], dtype=np.uint32) for j in jobs]
    for v, pp in zip(vs, pps):
        if v.shape[1] != len(pp):
            raise ValueError("variants must have one column per probe")
    ts = [_lib._u8(j[2]) for j in jobs]
    n_reads = np.array([v.shape[0] for v in vs], dtype=np.int32)
    n_probes = np.array([len(pp) for pp in pps], dtype=np.int32)

    def offs(sizes):
        o = np.zeros(len(sizes) + 1, dtype=np.uint64)
        np.cumsum(sizes, out=o[1:])
        return o
    var_off, ppos_off, tmpl_off = offs([v.size for v in vs]), offs(n_probes), offs([len(t) for t in ts])
    vcat = np.concatenate([v.ravel() for v in vs]) if var_off[-1] else np.zeros(1)
    pcat = np.concatenate(pps) if ppos_off[-1] else np.zeros(1, dtype=np.uint32)
    tcat = np.concatenate(ts)
    cfgs = (_CConfig * m)(*[j[3].to_c() for j in jobs])
    g = jobs[0][3].gains.to_c()
    states = np.ascontiguousarray(np.stack([j[4] for j in jobs]), dtype=np.uint64)
    pstride = max(max(j[3].copy_num, 1) for j in jobs)
    asn_off = offs(n_reads)
    post_off = offs(n_reads.astype(np.int64) * pstride)
    asn = np.zeros(int(asn_off[-1]), dtype=np.uint64)
    post = np.zeros(int(post_off[-1]), dtype=np.float64)
    score = np.zeros(m, dtype=np.float64)
    kk = np.zeros(m, dtype=np.int32)
    p = _lib._ptr
    rc = L.jtk_lc_clustering_variants_batch(ctx._h, m, p(vcat), p(var_off), p(n_reads), p(n_probes), p(pcat), p(ppos_off), p(tcat),
                                            p(tmpl_off), C.byref(g), C.cast(cfgs, vp), p(states), p(asn), p(asn_off), p(post),
                                            p(post_off), pstride, p(score), p(kk))
    if rc != 0:
        raise _lib.JtkError(rc, L.jtk_lc_last_error().decode())
    out = []
    for i, j in enumerate(jobs):
        j[4][:] = states[i]
        n, k = int(n_reads[i]), int(kk[i])
        a = asn[int(asn_off[i]):int(asn_off[i]) + n].copy()
        po = post[int(post_off[i]):int(post_off[i]) + n * pstride].reshape(n, pstride)[:, :k].copy()
        out.append(ClusteringResult(a, po, float(score[i]), k, pps[i].copy()))
    return out


def _search_and_cluster(ctx, hmm, cons, seqs, ops, strands, config: ClusteringConfig, state: np.ndarray) -> ClusteringResult:
    """pseudo_mcmc::clustering (pseudo_mcmc.rs:77-107) for one pile-up with the caller's generator."""
    n = len(seqs)
    if config.copy_num < 2:
        return ClusteringResult(np.zeros(n, np.uint64), np.zeros((n, 1)), 0.0, 1, np.zeros(0, np.uint32))
    b = ctx.batch([cons], list(seqs), list(ops), np.asarray(strands, dtype=np.uint8), np.zeros(n, np.uint32), config.band_width)
    try:
        b.modtable(hmm.forward().to_c(), hmm.reverse().to_c(), 9)
        n_probes, probe_pos, variants = b.search_variants(config.gains.gain, config.gains.prob, config.copy_num, config.coverage)
    finally:
        b.close()
    d = int(n_probes[0])
    return _clustering_variants_rng(variants[:, :max(d, 1)], probe_pos[0, :d], cons, config, state)


def clustering_recursive(ctx, cons, seqs, ops, strands, state: np.ndarray, hmm, config: ClusteringConfig,
                         first: Optional[ClusteringResult] = None):
    """mod.rs:126-190.  Returns (assignments list, posterior rows list, score, k).  `first`: the top-level clustering when it
    was already computed in a batch (only valid for copy_num < UPPER_COPY_NUM)."""
    if config.copy_num < UPPER_COPY_NUM:
        r = first if first is not None else _search_and_cluster(ctx, hmm, cons, seqs, ops, strands, config, state)
        return [int(a) for a in r.assignments], [list(p) for p in r.posterior], r.score, r.k
    rec = ClusteringConfig(config.band_width, BRANCH_NUM, config.coverage, config.local_coverage, config.gains)
    r = _search_and_cluster(ctx, hmm, cons, seqs, ops, strands, rec, state)
    asn, pss, score, k = [int(a) for a in r.assignments], [list(p) for p in r.posterior], r.score, r.k
    copy_numbers = estim_copy_num(asn, k, config.copy_num, config.coverage)
    if k <= 1:
        return asn, pss, score, k
    rec_results = []
    for cl, cp in enumerate(copy_numbers):
        idx = [i for i, a in enumerate(asn) if a == cl]          # filter_sub_clusters (mod.rs:200-221)
        sub_seqs = [seqs[i] for i in idx]
        sub_ops = [np.array(ops[i], copy=True) for i in idx]
        sub_str = [strands[i] for i in idx]
        pcfg = HMMPolishConfig.new(config.band_width, len(sub_seqs), 0)   # mod.rs:154
        sub_cons = hmm.polish_until_converge_antidiagonal(cons, sub_seqs, sub_ops, sub_str, pcfg, ctx=ctx)
        sub_cfg = ClusteringConfig(config.band_width, cp, config.coverage, config.local_coverage, config.gains)
        rec_results.append(clustering_recursive(ctx, sub_cons, sub_seqs, sub_ops, sub_str, state, hmm, sub_cfg))
    total_score = sum(x[2] for x in rec_results) + score
    cluster_nums = [x[3] for x in rec_results]
    total_k = sum(cluster_nums)
    offsets = list(np.cumsum([0] + cluster_nums[:-1]))
    pointers = [0] * BRANCH_NUM
    merged_asn, merged_post = [], []
    for a, ps in zip(asn, pss):
        in_asn = rec_results[a][0][pointers[a]]
        in_ps = rec_results[a][1][pointers[a]]
        pointers[a] += 1
        merged_asn.append(int(offsets[a]) + in_asn)
        posterior: List[float] = []
        for p, num in zip(ps, cluster_nums):
            posterior.extend([p - math.log(num)] * num)
        for t, p in enumerate(in_ps):
            posterior[t + int(offsets[a])] += p + math.log(cluster_nums[a])
        assert abs(1.0 - sum(math.exp(x) for x in posterior)) < 1e-4   # mod.rs:185
        merged_post.append(posterior)
    return merged_asn, merged_post, total_score, total_k


def estimate_model_parameters_on_both_strands(ds: DataSet, ctx=None, rounds: int = TRAIN_ROUND) -> PairHiddenMarkovModelOnStrands:
    """model_tune.rs:96-156: chunks with coverage within median +-2, first five by id; per round polish every pile-up
    with the current models (HMMPolishConfig(bw/2, n, 0)), then one Baum-Welch step over all of them (radius max bw / 2)."""
    ctx = ctx or default_context()
    chunks = {c.id: c for c in ds.selected_chunks}
    models = ds.model or PairHiddenMarkovModelOnStrands.default()
    models = PairHiddenMarkovModelOnStrands.new(models.forward(), models.reverse())
    pile: Dict[int, List[Node]] = {k: [] for k in chunks}
    for n in ds.nodes:
        pile.setdefault(n.chunk, []).append(n)
    covs = sorted(len(v) for v in pile.values())
    cov = covs[len(pile) // 2]
    filtered = sorted((k, v) for k, v in pile.items() if max(cov, 2) - 2 <= len(v) < cov + 2)[:TRAIN_UNIT_SIZE]
    pairs = []
    for uid, nodes in filtered:
        if uid not in chunks:
            continue
        bw = band_width(ds.read_type, len(chunks[uid].seq))
        pairs.append([np.array(chunks[uid].seq, copy=True), [n.seq for n in nodes], [np.array(n.ops, copy=True) for n in nodes],
                      [n.is_forward for n in nodes], bw])
    assert pairs, "no pile-up to train on"
    bw = max(p[4] for p in pairs)
    for _ in range(rounds):
        # the five polishes of a round are independent: one batch per distinct radius
        for radius in sorted({p[4] // 2 for p in pairs}):
            grp = [p for p in pairs if p[4] // 2 == radius]
            drafts = [p[0] for p in grp]
            reads = [r for p in grp for r in p[1]]
            ops = [o for p in grp for o in p[2]]
            strands = [s for p in grp for s in p[3]]
            tidx = np.repeat(np.arange(len(grp), dtype=np.uint32), [len(p[1]) for p in grp])
            cons, new_ops, _ = polish_chunks(models, drafts, reads, ops, strands, tidx,
                                             HMMPolishConfig.new(radius, max(len(p[1]) for p in grp), 0), ctx=ctx)
            k = 0
            for g, p in enumerate(grp):
                p[0] = cons[g]
                p[2] = new_ops[k:k + len(p[1])]
                k += len(p[1])
        packs = [(p[0], p[3], p[1], p[2]) for p in pairs]   # TrainingDataPack::new(cons, strands, seqs, ops)
        fit_antidiagonal_par_multiple(models, packs, bw // 2, ctx=ctx)
    return models


def local_clustering_selected(ds: DataSet, selection: Iterable[int], gains: Optional[Gains] = None, ctx=None,
                              fit_models: bool = True, rank: int = 0, world: int = 1, group=None) -> Optional[Dict[int, tuple]]:
    """local_clustering/mod.rs:56-83.  Mutates `ds` on rank 0 (chunk.seq / score / cluster_num, node.cluster / posterior /
    ops) and returns {chunk id: (consensus, score, cluster_num)} there; other ranks return None.
    gains=None runs likelihood_gains::estimate_gain_default on the GPU (mod.rs:60)."""
    ctx = ctx or default_context()
    selection = set(selection)
    update_coverage(ds)
    if fit_models:
        ds.model = estimate_model_parameters_on_both_strands(ds, ctx=ctx)           # mod.rs:58
    hmm = ds.model or PairHiddenMarkovModelOnStrands.default()
    if gains is None:
        from .likelihood_gains import estimate_gain_default
        gains = estimate_gain_default(hmm, ctx=ctx)
    coverage = float(ds.coverage)
    import time
    t0 = time.perf_counter()
    pile = {cid: pc for cid, pc in group_nodes(ds, selection).items() if pc[0]}
    ids = sorted(pile)
    weights = [scheduler.chunk_weight(len(pile[c][0]), len(pile[c][1].seq),
                                      sum(len(n.seq) for n in pile[c][0]) / len(pile[c][0]),
                                      band_width(ds.read_type, len(pile[c][1].seq)) // 2) for c in ids]
    t_group = time.perf_counter() - t0
    inner: Dict[str, float] = {}

    def process(my_ids: List[int]) -> Dict[int, tuple]:
        t1 = time.perf_counter()
        pileup_nodes(ds, selection, pile=pile, only=my_ids)   # the sort of mod.rs:47-50, for this rank's chunks only
        inner["pileup_sort"] = time.perf_counter() - t1
        out = _cluster_pileups(ctx, hmm, gains, coverage, ds.read_type, {c: pile[c] for c in my_ids})
        inner.update(LAST_TIMING)
        if world > 1:
            # what travels to rank 0 per chunk: assignments as one int array, log-posteriors as one n x k array, guide ops at
            # 2 bits per column (the reference's Node.cigar is run-length coded for the same reason)
            t1 = time.perf_counter()
            out = {c: (r[0], r[1], r[2], np.asarray(r[3], dtype=np.int64), np.asarray(r[4], dtype=np.float64), _pack_ops_chunk(r[5]))
                   for c, r in out.items()}
            inner["pack_ops"] = time.perf_counter() - t1
        return out

    t0 = time.perf_counter()
    merged = scheduler.run_sharded(ids, weights, process, rank, world, group=group)
    t_run = time.perf_counter() - t0
    if merged is None:
        return None
    t0 = time.perf_counter()
    out = {}
    for cid, (cons, score, k, asn, post, ops) in merged.items():
        nodes, chunk = pile[cid]
        if isinstance(ops, tuple):
            ops = _unpack_ops_chunk(ops)
        post = np.asarray(post, dtype=np.float64)
        asn = np.asarray(asn).tolist()
        for i, n in enumerate(nodes):                                               # update_by_clusterings (mod.rs:244-260)
            n.posterior = post[i]
            n.cluster = asn[i]
            n.ops = ops[i]
        chunk.seq, chunk.score, chunk.cluster_num = cons, score, k                 # mod.rs:74-81
        out[cid] = (cons, score, k)
    normalize_local_clustering(ds)
    LAST_TIMING.clear()
    LAST_TIMING.update(inner)
    LAST_TIMING.update({"group_nodes": t_group, "gather_wait": max(0.0, t_run - sum(v for k, v in inner.items() if k not in ("gpu_mcmc_chunks", "mcmc_gpu_side", "mcmc_host_side"))),
                        "write_back": time.perf_counter() - t0})
    return out


def _pack_ops(ops: np.ndarray):
    """Guide ops (values 0..3) as 2 bits per column for the host gather: (column count, packed bytes) -- jtk_ops_pack2."""
    o = np.ascontiguousarray(ops, dtype=np.uint8)
    out = np.empty((len(o) + 3) // 4, dtype=np.uint8)
    L = _lib.lib()
    L.jtk_ops_pack2.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    if L.jtk_ops_pack2(_lib._ptr(o), len(o), _lib._ptr(out)) != 0:
        raise ValueError("jtk_ops_pack2")
    return len(o), out


def _unpack_ops(packed) -> np.ndarray:
    n, b = packed
    b = np.ascontiguousarray(b, dtype=np.uint8)
    if len(b) != (n + 3) // 4:
        raise ValueError("packed ops: wrong length")
    out = np.empty(4 * len(b), dtype=np.uint8)
    L = _lib.lib()
    L.jtk_ops_unpack2.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    if L.jtk_ops_unpack2(_lib._ptr(b), n, _lib._ptr(out)) != 0:
        raise ValueError("jtk_ops_unpack2")
    return out[:n]


def _pack_ops_chunk(ops_list):
    """All guide ops of one chunk in one packed array: (lengths int32[n], packed bytes) -- one pass per chunk."""
    if isinstance(ops_list, _lib.Packed):
        raise TypeError("slice the Packed ops first")
    lens = np.fromiter((len(o) for o in ops_list), dtype=np.int32, count=len(ops_list))
    cat = np.concatenate(ops_list) if len(ops_list) else np.zeros(0, dtype=np.uint8)
    return lens, _pack_ops(cat)[1]


def _unpack_ops_chunk(packed):
    lens, b = packed
    flat = _unpack_ops((int(lens.sum()), b))
    ends = np.cumsum(lens).tolist()
    return [flat[a:e] for a, e in zip([0] + ends[:-1], ends)]   # views (np.split costs 6x as much per piece)


LAST_TIMING: Dict[str, float] = {}  # seconds spent in the phases of the last _cluster_pileups call (diagnostics)


def host_threads() -> int:
    """Threads of the per-chunk host loops (k-means / MCMC): the reference's rayon pool (`-t`, cli/src/bin/jtk.rs:396-408)."""
    import os
    v = os.environ.get("JTK_CLUSTER_THREADS") or os.environ.get("JTK_HOST_THREADS") or (os.cpu_count() or 1)
    return max(1, min(int(v), 64))


GPU_MCMC_CAPACITY = 4736   # diploid chains one B200 runs side by side (148 SMs x 8 warps x 4 chains, mcmc_diploid_kernel)
GPU_MCMC_HOST_S = 0.24     # seconds of one host thread per chunk (20 restarts x 2000 x 60 proposals: 0.18 s alone, 0.24 s with all cores busy)


GPU_MCMC_SPEC_CAPACITY = 2072  # chains mcmc_speculative_kernel holds in one wave (148 SMs x 14 evaluator warps)
_GPU_MCMC_LATENCY = ((296, 0.60), (592, 0.70), (1184, 0.85), (2072, 1.00))


def gpu_mcmc_seconds(n_chains: int) -> float:
    """Latency of jtk_mcmc_restarts_batch for n diploid chains of 60 reads x 20 restarts, as the share model uses it: the
    speculative kernel (one chain per evaluator warp, up to 2 072 chains a wave) takes 0.46 s (up to 2 chains per SM) to
    0.82 s (14 per SM) on settled six-column chains (profiles/r2_mcmc_speculative.txt) and 0.64 to 1.03 s on the one- and
    two-column chains of the bench's chunks (they accept more proposals: fewer commit per round); beyond one wave the
    sub-warp kernel (four chains per warp), 1.4 s for up to 4 736 chains."""
    if n_chains <= 0:
        return 0.0
    for cap, sec in _GPU_MCMC_LATENCY:
        if n_chains <= cap:
            return sec
    return 1.4 * -(-n_chains // GPU_MCMC_CAPACITY)


def gpu_mcmc_share(n_chunks: int) -> int:
    """How many of n_chunks go to jtk_mcmc_restarts_batch: the split that lets the GPU chains (concurrent: their latency
    hardly depends on their number) and the host threads (GPU_MCMC_HOST_S per chunk and thread) finish together; none
    for calls the host threads absorb in less than the latency of one GPU chain.  JTK_GPU_MCMC=0 / 1 forces none / all."""
    import os
    mode = os.environ.get("JTK_GPU_MCMC", "auto")
    if mode == "0":
        return 0
    if mode == "1":
        return n_chunks
    threads = host_threads()
    best, best_t = 0, -(-n_chunks // threads) * GPU_MCMC_HOST_S
    for t_gpu in (1.4,) + tuple(sec for _, sec in reversed(_GPU_MCMC_LATENCY)):   # (ties: the larger host share)
        host = int(t_gpu / GPU_MCMC_HOST_S) * threads           # chunks the host threads finish in that time
        n_gpu = max(0, n_chunks - host)
        if n_gpu == 0:
            continue
        t = max(gpu_mcmc_seconds(n_gpu), -(-(n_chunks - n_gpu) // threads) * GPU_MCMC_HOST_S)
        if t < best_t:
            best, best_t = n_gpu, t
    return min(best, GPU_MCMC_CAPACITY)


def _cluster_pileups(ctx, hmm, gains: Gains, coverage: float, read_type: str, pile: Dict[int, Tuple[List[Node], Chunk]]):
    """clustering_on_pileup (mod.rs:86-123) for a set of pile-ups: one polish batch and one table batch per radius; the
    host clustering of the chunks runs on a thread pool (the C call releases the GIL), one task per chunk as rayon does."""
    from concurrent.futures import ThreadPoolExecutor
    import time
    res: Dict[int, tuple] = {}
    tm = {"polish": 0.0, "tables_filter": 0.0, "host_clustering": 0.0, "recursive": 0.0}
    by_radius: Dict[int, List[int]] = {}
    for cid, (nodes, chunk) in pile.items():
        by_radius.setdefault(band_width(read_type, len(chunk.seq)) // 2, []).append(cid)
    for radius, cids in by_radius.items():
        cids.sort()
        drafts = [pile[c][1].seq for c in cids]
        reads = [n.seq for c in cids for n in pile[c][0]]
        ops = [n.ops for c in cids for n in pile[c][0]]
        strands = [n.is_forward for c in cids for n in pile[c][0]]
        counts = [len(pile[c][0]) for c in cids]
        tidx = np.repeat(np.arange(len(cids), dtype=np.uint32), counts)
        # HMMPolishConfig::new(band_width / 2, seqs.len(), 3): every read of a chunk votes (mod.rs:105)
        t0 = time.perf_counter()
        reads = _lib.Packed(*_lib.concat(reads))   # one concatenate of the 1e5 reads of a call, shared by the polish and table batches
        cons, new_ops, _ = polish_chunks(hmm, drafts, reads, ops, strands, tidx, HMMPolishConfig.new(radius, max(counts), 3), ctx=ctx)
        tm["polish"] += time.perf_counter() - t0
        first = np.concatenate([[0], np.cumsum(counts)])
        cfgs = []
        for g, c in enumerate(cids):
            cp, n = pile[c][1].copy_num, counts[g]
            # mod.rs:108-111: f64 division, copy_num 0 gives inf and `clustering` then returns the trivial one-cluster
            # result because copy_num < 2 (pseudo_mcmc.rs:86-88)
            ratio = n / cp if cp > 0 else float("inf")
            per_cluster = ratio if cp <= 2 else max(ratio, coverage)
            cfgs.append(ClusteringConfig.new(radius, cp, coverage, per_cluster, gains))
        # chunks below UPPER_COPY_NUM: one 9-row table batch + device-side filter_profiles for all of them
        small = [g for g, c in enumerate(cids) if 2 <= cfgs[g].copy_num < UPPER_COPY_NUM]
        firsts: Dict[int, ClusteringResult] = {}
        if small:
            t0 = time.perf_counter()
            if len(small) == len(cids):              # (the usual call: every chunk) the packed arrays as they are
                sub_reads, sub_ops, sub_str = reads, new_ops, strands
            else:
                sub_reads = [reads[k] for g in small for k in range(first[g], first[g + 1])]
                sub_ops = [new_ops[k] for g in small for k in range(first[g], first[g + 1])]
                sub_str = [strands[k] for g in small for k in range(first[g], first[g + 1])]
            sub_idx = np.repeat(np.arange(len(small), dtype=np.uint32), [counts[g] for g in small])
            b = ctx.batch([cons[g] for g in small], sub_reads, sub_ops, np.asarray(sub_str, dtype=np.uint8), sub_idx, radius)
            try:
                b.modtable(hmm.forward().to_c(), hmm.reverse().to_c(), 9)
                n_probes, probe_pos, variants = b.search_variants(gains.gain, gains.prob,
                                                                  np.array([cfgs[g].copy_num for g in small], dtype=np.int32), coverage)
            finally:
                b.close()
            off = np.concatenate([[0], np.cumsum([counts[g] for g in small])])
            tm["tables_filter"] += time.perf_counter() - t0
            t0 = time.perf_counter()

            def one(sg):
                s, g = sg
                d = int(n_probes[s])
                state = _rng_seed(pile[cids[g]][1].id * 3490)                       # mod.rs:97
                return g, _clustering_variants_rng(variants[off[s]:off[s + 1], :max(d, 1)], probe_pos[s, :d], cons[g],
                                                   cfgs[g], state)
            # The 2.4 M sequential MCMC proposals of a chunk take ~0.1 s of one host core and ~2.5 s of one GPU warp, but
            # the GPU runs ~1 800 chunks side by side (csrc/mcmc_kernels.cu): with many chunks in the call the restarts of
            # most of them go to the GPU while the host threads work through the rest.  Same streams, same results.
            todo = list(enumerate(small))
            n_gpu = gpu_mcmc_share(len(todo))
            gpu_part = [sg for sg in todo[:n_gpu] if int(n_probes[sg[0]]) >= 1]
            gpu_set = {sg[0] for sg in gpu_part}
            host_part = [sg for sg in todo if sg[0] not in gpu_set]

            def gpu_side():
                t_g = time.perf_counter()
                jobs = []
                for s, g in gpu_part:
                    d = int(n_probes[s])
                    jobs.append((variants[off[s]:off[s + 1], :d], probe_pos[s, :d], cons[g], cfgs[g],
                                 _rng_seed(pile[cids[g]][1].id * 3490)))
                out = _clustering_variants_batch_gpu(ctx, jobs)
                tm["mcmc_gpu_side"] = tm.get("mcmc_gpu_side", 0.0) + time.perf_counter() - t_g
                return out
            with ThreadPoolExecutor(max_workers=host_threads() + 1) as pool:
                fut = pool.submit(gpu_side) if gpu_part else None
                for g, r in pool.map(one, host_part):
                    firsts[g] = r
                tm["mcmc_host_side"] = tm.get("mcmc_host_side", 0.0) + time.perf_counter() - t0
                if fut is not None:
                    for (s, g), r in zip(gpu_part, fut.result()):
                        firsts[g] = r
            tm["host_clustering"] += time.perf_counter() - t0
            tm["gpu_mcmc_chunks"] = tm.get("gpu_mcmc_chunks", 0) + len(gpu_part)
        t0 = time.perf_counter()
        for g, c in enumerate(cids):
            sl = slice(first[g], first[g + 1])
            r = firsts.get(g)
            if r is not None and cfgs[g].copy_num < UPPER_COPY_NUM:   # clustered in the batch above: arrays as they are
                res[c] = (cons[g], r.score, r.k, r.assignments, r.posterior, new_ops[sl])
                continue
            state = _rng_seed(pile[c][1].id * 3490)
            asn, post, score, k = clustering_recursive(ctx, cons[g], reads[sl], new_ops[sl], strands[sl], state, hmm, cfgs[g],
                                                       first=firsts.get(g))
            res[c] = (cons[g], score, k, asn, post, new_ops[sl])
        tm["recursive"] += time.perf_counter() - t0
    LAST_TIMING.clear()
    LAST_TIMING.update(tm)
    return res
