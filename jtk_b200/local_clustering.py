"""Host-side mirror of haplotyper::local_clustering::pseudo_mcmc (reference: pseudo_mcmc.rs:18-138).

`clustering(...)` has the reference's argument meaning: template, reads, ops, strands, an RNG seed (the reference seeds
Xoshiro256StarStar with chunk.id * 3490, local_clustering/mod.rs:97), the strand models and a ClusteringConfig.  The
pair-HMM part runs on the GPU through the C ABI; the column filters / greedy pick / k-means / MCMC are the C++ host
restatement inside libjtkgpu.so (csrc/local_clustering.cpp)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import NUM_ROW, Batch, Context, _ptr

MIN_REQ_FRACTION = 0.5  # pseudo_mcmc.rs:140
POS_THR = 1e-5          # pseudo_mcmc.rs:5


class _CGains(C.Structure):
    _fields_ = [("homop_len", C.c_int), ("gain", C.c_void_p), ("prob", C.c_void_p)]


class _CConfig(C.Structure):
    _fields_ = [("band_width", C.c_int), ("copy_num", C.c_int), ("coverage", C.c_double), ("local_coverage", C.c_double)]


@dataclass
class Gains:
    """likelihood_gains::Gains: rows Subst, Del, Ins; columns homopolymer length 1..H."""
    gain: np.ndarray  # float64[3, H]
    prob: np.ndarray  # float64[3, H]

    def __post_init__(self):
        self.gain = np.ascontiguousarray(self.gain, dtype=np.float64)
        self.prob = np.ascontiguousarray(self.prob, dtype=np.float64)
        assert self.gain.shape == self.prob.shape and self.gain.shape[0] == 3

    @property
    def min_req(self) -> np.ndarray:
        return (self.gain * MIN_REQ_FRACTION).astype(np.float32)

    def to_c(self) -> _CGains:
        return _CGains(self.gain.shape[1], self.gain.ctypes.data, self.prob.ctypes.data)


@dataclass
class ClusteringConfig:
    band_width: int
    copy_num: int
    coverage: float
    local_coverage: float
    gains: Gains

    @classmethod
    def new(cls, band_width, copy_num, coverage, local_coverage, gains):
        return cls(band_width, copy_num, coverage, local_coverage, gains)

    def to_c(self) -> _CConfig:
        return _CConfig(self.band_width, self.copy_num, self.coverage, self.local_coverage)


@dataclass
class ClusteringResult:
    assignments: np.ndarray      # uint64[n]
    posterior: np.ndarray        # float64[n, k] log-posteriors
    score: float
    k: int
    probes: np.ndarray           # uint32 flat positions j*NUM_ROW+row of the selected variant columns


_bound = False


def _bind():
    global _bound
    L = _lib.lib()
    if not _bound:
        vp = C.c_void_p
        common = [vp, C.POINTER(_CGains), C.POINTER(_CConfig), C.c_uint64, vp, vp, C.c_int, C.POINTER(C.c_double),
                  C.POINTER(C.c_int), vp, C.c_int, C.POINTER(C.c_int)]
        L.jtk_lc_clustering_profiles.argtypes = [vp, C.c_int, vp, C.c_int] + common
        L.jtk_lc_clustering_batch.argtypes = [C.c_void_p, C.c_int, vp, C.c_int, C.c_int, vp, C.POINTER(_CGains),
                                              C.POINTER(_CConfig), C.c_uint64, vp, vp, C.c_int, C.POINTER(C.c_double),
                                              C.POINTER(C.c_int), vp, C.c_int, C.POINTER(C.c_int)]
        L.jtk_lc_clustering_variants.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.POINTER(_CGains),
                                                 C.POINTER(_CConfig), C.c_uint64, vp, vp, C.c_int, C.POINTER(C.c_double),
                                                 C.POINTER(C.c_int)]
        L.jtk_lc_cluster_filtered_variants_exact.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                                             C.POINTER(C.c_double)]
        L.jtk_lc_last_error.restype = C.c_char_p
        L.jtk_lc_cosine_similarity.restype = C.c_double
        L.jtk_lc_cosine_similarity.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.jtk_lc_homopolymer_length.argtypes = [vp, C.c_int, vp]
        L.jtk_lc_rng_words.argtypes = [C.c_uint64, C.c_int, vp, C.c_int, vp]
        L.jtk_lc_rng_words.restype = None
        _bound = True
    return L


def _finish(rc, n, stride, asn, post, score, k, probes, nprobe) -> ClusteringResult:
    if rc != 0:
        raise _lib.JtkError(rc, _bind().jtk_lc_last_error().decode())
    kk = int(k.value)
    return ClusteringResult(asn, post.reshape(n, stride)[:, :kk].copy(), float(score.value), kk,
                            probes[:min(int(nprobe.value), len(probes))].copy())


def clustering_on_profiles(profiles, template, strands, config: ClusteringConfig, seed: int) -> ClusteringResult:
    """pseudo_mcmc::clustering after `modification_table`: profiles float64[n, (Lt+1)*NUM_ROW] = table - lk."""
    L = _bind()
    prof = np.ascontiguousarray(profiles, dtype=np.float64)
    t = _lib._u8(template)
    st = _lib._u8(np.asarray(strands, dtype=np.uint8))
    n = prof.shape[0]
    stride = max(config.copy_num, 1)
    asn = np.zeros(n, dtype=np.uint64)
    post = np.zeros(n * stride, dtype=np.float64)
    probes = np.zeros(64, dtype=np.uint32)
    score, k, nprobe = C.c_double(), C.c_int(), C.c_int()
    g, c = config.gains.to_c(), config.to_c()
    rc = L.jtk_lc_clustering_profiles(_ptr(prof), n, _ptr(t), len(t), _ptr(st), C.byref(g), C.byref(c), seed, _ptr(asn),
                                      _ptr(post), stride, C.byref(score), C.byref(k), _ptr(probes), len(probes),
                                      C.byref(nprobe))
    return _finish(rc, n, stride, asn, post, score, k, probes, nprobe)


def clustering_on_batch(batch: Batch, tmpl: int, template, stats, config: ClusteringConfig, seed: int) -> ClusteringResult:
    """The same with the chunk's profiles resident on the device (stats: this chunk's slice of Batch.colstats)."""
    L = _bind()
    t = _lib._u8(template)
    n = int((batch.tmpl_idx == tmpl).sum())
    stride = max(config.copy_num, 1)
    asn = np.zeros(n, dtype=np.uint64)
    post = np.zeros(n * stride, dtype=np.float64)
    probes = np.zeros(64, dtype=np.uint32)
    score, k, nprobe = C.c_double(), C.c_int(), C.c_int()
    g, c = config.gains.to_c(), config.to_c()
    stats = np.ascontiguousarray(stats)
    rc = L.jtk_lc_clustering_batch(batch._h, tmpl, _ptr(t), len(t), n, _ptr(stats), C.byref(g), C.byref(c), seed,
                                   _ptr(asn), _ptr(post), stride, C.byref(score), C.byref(k), _ptr(probes), len(probes),
                                   C.byref(nprobe))
    return _finish(rc, n, stride, asn, post, score, k, probes, nprobe)


def clustering_on_variants(variants, probe_pos, template, config: ClusteringConfig, seed: int) -> ClusteringResult:
    """pseudo_mcmc::clustering after `search_variants`: variants float64[n, >=D] (one chunk's rows of
    Batch.search_variants), probe_pos uint32[D]."""
    L = _bind()
    v = np.ascontiguousarray(variants, dtype=np.float64)
    pp = np.ascontiguousarray(probe_pos, dtype=np.uint32)
    t = _lib._u8(template)
    n, stride = v.shape
    pstride = max(config.copy_num, 1)
    asn = np.zeros(n, dtype=np.uint64)
    post = np.zeros(n * pstride, dtype=np.float64)
    score, k = C.c_double(), C.c_int()
    g, c = config.gains.to_c(), config.to_c()
    rc = L.jtk_lc_clustering_variants(_ptr(v), n, len(pp), stride, _ptr(pp), _ptr(t), len(t), C.byref(g), C.byref(c), seed,
                                      _ptr(asn), _ptr(post), pstride, C.byref(score), C.byref(k))
    if rc != 0:
        raise _lib.JtkError(rc, L.jtk_lc_last_error().decode())
    kk = int(k.value)
    return ClusteringResult(asn, post.reshape(n, pstride)[:, :kk].copy(), float(score.value), kk, pp.copy())


def cluster_filtered_variants_exact(variants, copy_num: int):
    """exact_clustering::cluster_filtered_variants_exact (exact_clustering.rs:7-77): exhaustive search over one column subset
    per cluster.  Returns (assignments, per-read per-cluster gains, score, copy_num) like the reference's tuple."""
    L = _bind()
    v = np.ascontiguousarray(variants, dtype=np.float64)
    n, d = v.shape
    asn = np.zeros(n, dtype=np.uint64)
    gains = np.zeros((n, copy_num), dtype=np.float64)
    score = C.c_double()
    rc = L.jtk_lc_cluster_filtered_variants_exact(_ptr(v), n, d, d, copy_num, _ptr(asn), _ptr(gains), C.byref(score))
    if rc != 0:
        raise _lib.JtkError(rc, L.jtk_lc_last_error().decode())
    return asn, gains, float(score.value), copy_num


def local_clustering_batch(batch: Batch, templates, config_of, seeds, hmm_fwd, hmm_rev, gains: Gains, coverage: float):
    """The per-chunk loop of local_clustering_selected (local_clustering/mod.rs:64-72) after polishing, for a whole
    batch: 9-row modification tables, device-side filter_profiles, gathered candidate columns, host pick + MCMC.
    config_of(t) -> ClusteringConfig of chunk t; seeds[t] the chunk's RNG seed (chunk.id * 3490, mod.rs:97)."""
    batch.modtable(hmm_fwd, hmm_rev, 9)
    cfgs = [config_of(t) for t in range(batch.n_tmpl)]
    copy_num = np.array([max(c.copy_num, 1) for c in cfgs], dtype=np.int32)
    n_probes, probe_pos, variants = batch.search_variants(gains.gain, gains.prob, copy_num, coverage)
    out = []
    for t in range(batch.n_tmpl):
        rows = np.flatnonzero(batch.tmpl_idx == t)
        d = int(n_probes[t])
        out.append(clustering_on_variants(variants[rows][:, :max(d, 1)] if d else np.zeros((len(rows), 1)), probe_pos[t, :d],
                                          templates[t], cfgs[t], int(seeds[t])))
    return out


def clustering(template, reads: Sequence, ops: Sequence, strands: Sequence[bool], seed: int, hmm, config: ClusteringConfig,
               ctx: Optional[Context] = None) -> ClusteringResult:
    """pseudo_mcmc::clustering (pseudo_mcmc.rs:77-107) for one chunk: GPU modification tables (rows used by
    filter_profiles only), device-side column statistics, host-side filters and MCMC."""
    from .hmm import default_context
    ctx = ctx or default_context()
    n = len(reads)
    if config.copy_num < 2:
        return ClusteringResult(np.zeros(n, np.uint64), np.zeros((n, 1)), 0.0, 1, np.zeros(0, np.uint32))
    b = ctx.batch([template], list(reads), list(ops), np.asarray(strands, dtype=np.uint8), np.zeros(n, np.uint32),
                  config.band_width)
    try:
        b.modtable(hmm.forward().to_c(), hmm.reverse().to_c(), 9)
        stats = b.colstats(config.gains.min_req, POS_THR)
        return clustering_on_batch(b, 0, template, stats, config, seed)
    finally:
        b.close()


def cosine_similarity(profiles, i: int, j: int) -> float:
    p = np.ascontiguousarray(profiles, dtype=np.float64)
    return float(_bind().jtk_lc_cosine_similarity(_ptr(p), p.shape[0], p.shape[1], i, j))


def homopolymer_length(xs) -> np.ndarray:
    x = _lib._u8(xs)
    out = np.zeros(len(x), dtype=np.uint32)
    _bind().jtk_lc_homopolymer_length(_ptr(x), len(x), _ptr(out))
    return out


def rng_words(seed: int, n: int, state=None) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    st = None if state is None else np.ascontiguousarray(state, dtype=np.uint64)
    _bind().jtk_lc_rng_words(seed, 0 if st is None else 1, _ptr(st), n, _ptr(out))
    return out
