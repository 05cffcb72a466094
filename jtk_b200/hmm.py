"""Host-side mirror of the kiley::hmm call surface that ban-m/jtk uses on its per-chunk hot path.

Same names, argument meaning and error behaviour as the reference call sites (the reference panics on
failure -- `model_tune.rs:21`, `pseudo_mcmc.rs:100` -- here a `JtkError` is raised):

  PairHiddenMarkovModel            fields as `haplotyper/src/model_tune.rs:50-62`
  PairHiddenMarkovModelOnStrands   new / forward / reverse / default  (`model_tune.rs:32`, `pseudo_mcmc.rs:59-60`)
  modification_table_antidiagonal  `haplotyper/src/local_clustering/pseudo_mcmc.rs:62-63`
  likelihood_antidiagonal_bootstrap `haplotyper/src/likelihood_gains.rs:27-28`
  NUM_ROW, COPY_SIZE               `pseudo_mcmc.rs:7,172`

Every call goes through the C ABI of libjtkgpu.so into the sm_100a kernels; there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import COPY_SIZE, DEL_SIZE, NUM_ROW, Context, HmmParams, JtkError  # noqa: F401

_default_ctx: Optional[Context] = None


class PolishCfg(_lib.C.Structure):
    """jtk_polish_config == kiley::hmm::HMMPolishConfig::new(radius, take_num, ignore_edge)."""
    _fields_ = [("radius", _lib.C.c_int), ("take_num", _lib.C.c_int), ("ignore_edge", _lib.C.c_int)]


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


@dataclass
class PairHiddenMarkovModel:
    mat_mat: float = 0.97
    mat_ins: float = 0.01
    mat_del: float = 0.01
    ins_mat: float = 0.97
    ins_ins: float = 0.01
    ins_del: float = 0.01
    del_mat: float = 0.97
    del_ins: float = 0.01
    del_del: float = 0.01
    mat_emit: List[float] = field(default_factory=lambda: [0.97 if r == q else 0.01 for r in range(4) for q in range(4)])
    ins_emit: List[float] = field(default_factory=lambda: [0.25] * 20)

    def to_c(self) -> HmmParams:
        h = HmmParams()
        for n in ("mat_mat", "mat_ins", "mat_del", "ins_mat", "ins_ins", "ins_del", "del_mat", "del_ins", "del_del"):
            setattr(h, n, float(getattr(self, n)))
        for k in range(16):
            h.mat_emit[k] = float(self.mat_emit[k])
        for k in range(20):
            h.ins_emit[k] = float(self.ins_emit[k])
        return h

    @classmethod
    def from_array(cls, a) -> "PairHiddenMarkovModel":
        a = np.asarray(a, dtype=np.float64)
        assert a.size == 45
        return cls(*[float(x) for x in a[:9]], mat_emit=[float(x) for x in a[9:25]], ins_emit=[float(x) for x in a[25:45]])

    def as_array(self) -> np.ndarray:
        return np.array([self.mat_mat, self.mat_ins, self.mat_del, self.ins_mat, self.ins_ins, self.ins_del,
                         self.del_mat, self.del_ins, self.del_del] + list(self.mat_emit) + list(self.ins_emit))

    # ---- kiley-shaped single-pair calls (each is a batch of one) ------------------------------
    def modification_table_antidiagonal(self, template, read, ops, band: int, ctx: Optional[Context] = None):
        """-> (table float64[(Lt+1)*NUM_ROW] of absolute log-likelihoods, lk)."""
        ctx = ctx or default_context()
        c = self.to_c()
        lk, tabs = ctx.modtable_batch(c, c, [template], [read], [ops], [1], [0], band)
        return tabs[0], float(lk[0])

    def likelihood_antidiagonal_bootstrap(self, template, read, band: int, ctx: Optional[Context] = None) -> float:
        ctx = ctx or default_context()
        c = self.to_c()
        return float(ctx.likelihood_batch(c, c, [template], [read], None, [1], [0], band)[0])

    def likelihood_antidiagonal(self, template, read, ops, band: int, ctx: Optional[Context] = None) -> float:
        ctx = ctx or default_context()
        c = self.to_c()
        return float(ctx.likelihood_batch(c, c, [template], [read], [ops], [1], [0], band)[0])


@dataclass
class HMMPolishConfig:
    """kiley::hmm::HMMPolishConfig::new(radius, take_num, ignore_edge) (local_clustering/mod.rs:105)."""
    radius: int
    take_num: int
    ignore_edge: int

    @classmethod
    def new(cls, radius: int, take_num: int, ignore_edge: int) -> "HMMPolishConfig":
        return cls(radius, take_num, ignore_edge)


@dataclass
class PairHiddenMarkovModelOnStrands:
    _forward: PairHiddenMarkovModel = field(default_factory=PairHiddenMarkovModel)
    _reverse: PairHiddenMarkovModel = field(default_factory=PairHiddenMarkovModel)

    @classmethod
    def new(cls, forward: PairHiddenMarkovModel, reverse: PairHiddenMarkovModel) -> "PairHiddenMarkovModelOnStrands":
        return cls(forward, reverse)

    @classmethod
    def default(cls) -> "PairHiddenMarkovModelOnStrands":
        return cls()

    def forward(self) -> PairHiddenMarkovModel:
        return self._forward

    def reverse(self) -> PairHiddenMarkovModel:
        return self._reverse

    # ---- batched forms used by the haplotyper-side host code -----------------------------------
    def modification_tables(self, template, reads: Sequence, ops: Sequence, strands: Sequence[bool], band: int,
                            ctx: Optional[Context] = None):
        """All reads of one chunk against its consensus: the loop of `pseudo_mcmc.rs:53-67` as one batch.
        Returns (profiles float64[n, (Lt+1)*NUM_ROW] with lk already subtracted, lks)."""
        ctx = ctx or default_context()
        n = len(reads)
        lk, tabs = ctx.modtable_batch(self._forward.to_c(), self._reverse.to_c(), [template], list(reads), list(ops),
                                      np.asarray(strands, dtype=np.uint8), np.zeros(n, dtype=np.uint32), band)
        prof = np.stack(tabs) - lk[:, None]
        prof[np.stack(tabs) <= _lib.TABLE_NEG * 0.1] = _lib.TABLE_NEG
        return prof, lk

    def polish_until_converge_antidiagonal(self, draft, seqs: Sequence, ops: list, strands: Sequence[bool],
                                           config: HMMPolishConfig, ctx: Optional[Context] = None) -> np.ndarray:
        """Reference signature (local_clustering/mod.rs:106): returns the polished consensus; `ops` (a list of uint8
        arrays) is rewritten in place like the reference's `&mut ops`."""
        cons, new_ops, _ = polish_chunks(self, [draft], list(seqs), ops, strands, np.zeros(len(seqs), np.uint32), config,
                                         ctx=ctx)
        for k in range(len(ops)):
            ops[k] = new_ops[k]
        return cons[0]


def polish_chunks(models: PairHiddenMarkovModelOnStrands, drafts: Sequence, reads: Sequence, ops: Sequence,
                  strands: Sequence[bool], tmpl_idx, config: HMMPolishConfig, ctx: Optional[Context] = None):
    """jtk_polish_until_converge_batch: many chunks per call.  Returns (consensus list, ops list, iterations)."""
    import ctypes as C
    ctx = ctx or default_context()
    L = _lib.lib()
    n_chunks, n_pairs = len(drafts), len(reads)
    dcat, doff = _lib.concat(drafts)
    rcat, roff = _lib.concat(reads)
    tmpl_idx = np.ascontiguousarray(tmpl_idx, dtype=np.uint32)
    st = _lib._u8(np.asarray(strands, dtype=np.uint8))
    dlen = np.diff(doff.astype(np.int64))
    rlen = np.diff(roff.astype(np.int64))
    caps = (rlen + 2 * dlen[tmpl_idx] + 64).astype(np.uint32)
    pos = np.zeros(n_pairs + 1, dtype=np.uint64)
    np.cumsum(caps, out=pos[1:])
    # every read's ops at pos[k] with room to grow (caps[k]): one compact concatenate, then jtk_scatter_runs copies the runs
    # into the padded slots of a staging buffer the context keeps between calls (0.7 GB per 120 000 reads: a fresh one costs
    # more in page faults than the copy itself)
    ocat, ooff = _lib.concat(ops)
    n_ops = np.diff(ooff.astype(np.int64)).astype(np.uint32)
    if (n_ops > caps).any():
        raise ValueError("ops longer than read + 2 * draft + 64")
    need = int(pos[-1])
    buf = getattr(ctx, "_polish_ops_buf", None)
    if buf is None or len(buf) < need:
        buf = np.empty(need + need // 8 + 64, dtype=np.uint8)
        ctx._polish_ops_buf = buf
    L.jtk_scatter_runs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    ctx._check(L.jtk_scatter_runs(_lib._ptr(ocat), _lib._ptr(ooff), n_pairs, _lib._ptr(buf), _lib._ptr(pos)))
    ccap = (2 * dlen + 64).astype(np.uint32)
    cpos = np.zeros(n_chunks + 1, dtype=np.uint64)
    np.cumsum(ccap, out=cpos[1:])
    cons = np.zeros(int(cpos[-1]), dtype=np.uint8)
    clen = np.zeros(n_chunks, dtype=np.uint32)
    iters = np.zeros(n_chunks, dtype=np.int32)

    cfg = PolishCfg(config.radius, config.take_num, config.ignore_edge)
    vp = C.c_void_p
    L.jtk_polish_until_converge_batch.argtypes = [vp, C.POINTER(HmmParams), C.POINTER(HmmParams), C.c_int, vp, vp, C.c_int,
                                                  vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(PolishCfg), vp, vp, vp, vp, vp]
    f, r = models.forward().to_c(), models.reverse().to_c()
    p = _lib._ptr
    ctx._check(L.jtk_polish_until_converge_batch(ctx._h, C.byref(f), C.byref(r), n_chunks, p(dcat), p(doff), n_pairs, p(rcat),
                                                 p(roff), p(buf), p(pos), p(caps), p(n_ops), p(st), p(tmpl_idx),
                                                 C.byref(cfg), p(cons), p(cpos), p(ccap), p(clen), p(iters)))
    out_cons = [cons[int(cpos[c]):int(cpos[c]) + int(clen[c])].copy() for c in range(n_chunks)]
    # the patched guide paths as views into ONE compact copy (120 000 per-read copies cost more than the kernels of a round)
    off = np.zeros(n_pairs + 1, dtype=np.uint64)
    L.jtk_compact_runs.argtypes = [vp, vp, vp, C.c_int, vp]
    ctx._check(L.jtk_compact_runs(p(buf), p(pos), p(n_ops), n_pairs, p(off)))
    out_ops = _lib.Packed(buf[:int(off[-1])].copy(), off.astype(np.uint32))   # behaves like the list of the per-read ops
    return out_cons, out_ops, iters


def expected_counts(models: PairHiddenMarkovModelOnStrands, templates, reads, ops, strands, tmpl_idx, radius,
                    ctx: Optional[Context] = None) -> np.ndarray:
    """jtk_batch_expected_counts: float64[2, 45] (forward-strand reads, reverse-strand reads)."""
    import ctypes as C
    ctx = ctx or default_context()
    L = _lib.lib()
    b = ctx.batch(list(templates), list(reads), list(ops), np.asarray(strands, dtype=np.uint8), tmpl_idx, radius)
    try:
        acc = np.zeros(90, dtype=np.float64)
        L.jtk_batch_expected_counts.argtypes = [C.c_void_p, C.POINTER(HmmParams), C.POINTER(HmmParams), C.c_void_p]
        f, r = models.forward().to_c(), models.reverse().to_c()
        ctx._check(L.jtk_batch_expected_counts(b._h, C.byref(f), C.byref(r), _lib._ptr(acc)))
        return acc.reshape(2, 45)
    finally:
        b.close()


def fit_antidiagonal_par_multiple(models: PairHiddenMarkovModelOnStrands, packs, radius: int,
                                  ctx: Optional[Context] = None) -> None:
    """Reference signature (model_tune.rs:151): packs = [(cons, strands, seqs, ops), ...] as TrainingDataPack::new;
    both strand models are updated in place by one EM step."""
    import ctypes as C
    ctx = ctx or default_context()
    L = _lib.lib()
    templates = [p[0] for p in packs]
    reads = [r for p in packs for r in p[2]]
    ops = [o for p in packs for o in p[3]]
    strands = np.concatenate([np.asarray(p[1], dtype=np.uint8) for p in packs])
    tidx = np.repeat(np.arange(len(packs), dtype=np.uint32), [len(p[2]) for p in packs])
    tcat, toff, rcat, roff, ocat, ooff, st, ti = _lib.pack_inputs(templates, reads, ops, strands, tidx)
    f, r = models.forward().to_c(), models.reverse().to_c()
    vp = C.c_void_p
    L.jtk_hmm_fit_batch.argtypes = [vp, C.POINTER(HmmParams), C.POINTER(HmmParams), C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                    vp, vp, C.c_int]
    p = _lib._ptr
    ctx._check(L.jtk_hmm_fit_batch(ctx._h, C.byref(f), C.byref(r), len(reads), len(templates), p(tcat), p(toff), p(rcat),
                                   p(roff), p(ocat), p(ooff), p(st), p(ti), radius))
    models._forward = PairHiddenMarkovModel.from_array(np.frombuffer(bytes(f), dtype=np.float64))
    models._reverse = PairHiddenMarkovModel.from_array(np.frombuffer(bytes(r), dtype=np.float64))
