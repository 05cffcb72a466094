"""ctypes loader for libjtkgpu.so (the C ABI declared in include/jtk_gpu.h).

There is no CPU fallback: if the shared library is missing this module raises at first use, and if no
CUDA device is present `Context()` raises `JtkError` (JTK_ECUDA).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# JTK_LIB_PATH: kernel-tuning experiments load an alternative build of the SAME library (tools/prof.py); never a fallback
LIB_PATH = os.environ.get("JTK_LIB_PATH") or os.path.join(HERE, "libjtkgpu.so")

NUM_ROW = 14
COPY_SIZE = 3
DEL_SIZE = 3
TABLE_NEG = -1.0e10

EXPORTS = [
    "jtk_ctx_create", "jtk_ctx_destroy", "jtk_last_error", "jtk_hmm_num_row", "jtk_hmm_copy_size",
    "jtk_hmm_del_size", "jtk_ctx_launch_count", "jtk_ctx_last_kernel_ms", "jtk_ctx_last_modtable_variant", "jtk_hmm_modtable_batch",
    "jtk_hmm_likelihood_batch", "jtk_band_cell_count", "jtk_batch_create", "jtk_batch_destroy",
    "jtk_batch_cell_updates", "jtk_batch_h2d_bytes", "jtk_batch_modtable", "jtk_batch_sync", "jtk_batch_fetch_lk",
    "jtk_batch_fetch_profile", "jtk_batch_colstats", "jtk_batch_gather", "jtk_ctx_timer_start", "jtk_ctx_timer_stop",
    "jtk_ctx_kernel_times", "jtk_ctx_measure_fp32_peak", "jtk_batch_candidates", "jtk_batch_search_variants",
    "jtk_lc_clustering_variants", "jtk_batch_colsums", "jtk_batch_expected_counts", "jtk_hmm_fit_batch",
    "jtk_polish_until_converge_batch", "jtk_lc_clustering_profiles", "jtk_lc_clustering_batch", "jtk_lc_last_error",
]

COLSTAT_DTYPE = np.dtype([("sum", "<f8"), ("count", "<i4"), ("sc", "<u2", (4,)), ("pad", "<i4")])
assert COLSTAT_DTYPE.itemsize == 24
CANDIDATE_DTYPE = np.dtype([("tmpl", "<u4"), ("pos", "<u4"), ("count", "<u4"), ("pad", "<u4"), ("sum", "<f8"), ("lk", "<f8")])
assert CANDIDATE_DTYPE.itemsize == 32


class CGains(C.Structure):
    """jtk_gains (likelihood_gains::Gains)."""
    _fields_ = [("homop_len", C.c_int), ("gain", C.c_void_p), ("prob", C.c_void_p)]


class JtkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libjtkgpu error {code}: {msg}")
        self.code = code


class HmmParams(C.Structure):
    """jtk_hmm_params == definitions::HMMParam (definitions/src/lib.rs:102-126)."""
    _fields_ = [(n, C.c_double) for n in
                ("mat_mat", "mat_ins", "mat_del", "ins_mat", "ins_ins", "ins_del", "del_mat", "del_ins", "del_del")] + \
               [("mat_emit", C.c_double * 16), ("ins_emit", C.c_double * 20)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m jtk_b200.build` (nvcc, sm_100a). "
            "jtk_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32p, u64p = C.c_void_p, C.c_void_p, C.c_void_p
    L.jtk_ctx_create.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
    L.jtk_ctx_destroy.argtypes = [C.c_void_p]
    L.jtk_ctx_destroy.restype = None
    L.jtk_last_error.argtypes = [C.c_void_p]
    L.jtk_last_error.restype = C.c_char_p
    L.jtk_ctx_launch_count.argtypes = [C.c_void_p]
    L.jtk_ctx_launch_count.restype = C.c_uint64
    L.jtk_ctx_last_kernel_ms.argtypes = [C.c_void_p]
    L.jtk_ctx_last_kernel_ms.restype = C.c_float
    L.jtk_ctx_last_modtable_variant.argtypes = [C.c_void_p]
    L.jtk_ctx_last_modtable_variant.restype = C.c_int
    batch = [C.c_void_p, C.POINTER(HmmParams), C.POINTER(HmmParams), C.c_int, C.c_int, vp, u32p, vp, u32p, vp, u32p,
             vp, u32p, C.c_int, vp]
    L.jtk_hmm_modtable_batch.argtypes = batch + [vp, u64p]
    L.jtk_hmm_likelihood_batch.argtypes = batch
    L.jtk_band_cell_count.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.jtk_band_cell_count.restype = C.c_int64
    L.jtk_batch_create.argtypes = [C.c_void_p, C.c_int, C.c_int, vp, u32p, vp, u32p, vp, u32p, vp, u32p, C.c_int,
                                   C.POINTER(C.c_void_p)]
    L.jtk_batch_destroy.argtypes = [C.c_void_p]
    L.jtk_batch_destroy.restype = None
    L.jtk_batch_cell_updates.argtypes = [C.c_void_p]
    L.jtk_batch_cell_updates.restype = C.c_uint64
    L.jtk_batch_h2d_bytes.argtypes = [C.c_void_p]
    L.jtk_batch_h2d_bytes.restype = C.c_uint64
    L.jtk_batch_modtable.argtypes = [C.c_void_p, C.POINTER(HmmParams), C.POINTER(HmmParams), C.c_int]
    L.jtk_batch_sync.argtypes = [C.c_void_p]
    L.jtk_batch_fetch_lk.argtypes = [C.c_void_p, vp]
    L.jtk_batch_fetch_profile.argtypes = [C.c_void_p, C.c_int, vp]
    L.jtk_batch_colstats.argtypes = [C.c_void_p, vp, C.c_int, C.c_float, vp, u64p]
    L.jtk_batch_gather.argtypes = [C.c_void_p, C.c_int, vp, C.c_int, u32p, C.c_int, vp]
    L.jtk_batch_candidates.argtypes = [C.c_void_p, C.POINTER(CGains), vp, C.c_double, vp, C.c_int, C.POINTER(C.c_int)]
    L.jtk_batch_search_variants.argtypes = [C.c_void_p, C.POINTER(CGains), vp, C.c_double, C.c_int, vp, vp, vp]
    L.jtk_ctx_timer_start.argtypes = [C.c_void_p]
    L.jtk_ctx_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.jtk_ctx_kernel_times.argtypes = [C.c_void_p, vp, C.c_int]
    L.jtk_ctx_measure_fp32_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray)):
        return np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


class Packed:
    """Byte sequences that are already one array: `cat` uint8 and `off` uint32[n + 1].  Behaves like the list of its pieces
    (len, indexing, slices give views) and passes through `concat` without another 1e5-piece concatenate."""
    __slots__ = ("cat", "off", "_o")

    def __init__(self, cat: np.ndarray, off: np.ndarray):
        self.cat = np.ascontiguousarray(cat, dtype=np.uint8)
        self.off = np.ascontiguousarray(off, dtype=np.uint32)
        self._o = self.off.tolist()

    def __len__(self):
        return len(self._o) - 1

    def __getitem__(self, k):
        if isinstance(k, slice):
            o, c = self._o, self.cat
            return [c[o[i]:o[i + 1]] for i in range(*k.indices(len(o) - 1))]
        if k < 0:
            k += len(self._o) - 1
        return self.cat[self._o[k]:self._o[k + 1]]

    def __iter__(self):
        o, c = self._o, self.cat
        return (c[o[i]:o[i + 1]] for i in range(len(o) - 1))


def concat(seqs):
    """list of uint8 arrays -> (concat uint8, offsets uint32[n+1]); a Packed or a (concat, offsets) tuple passes through."""
    if isinstance(seqs, Packed):
        return seqs.cat, seqs.off
    if isinstance(seqs, tuple) and len(seqs) == 2 and isinstance(seqs[0], np.ndarray) and seqs[0].dtype == np.uint8:
        return np.ascontiguousarray(seqs[0]), np.ascontiguousarray(seqs[1], dtype=np.uint32)
    n = len(seqs)
    lens = np.fromiter(map(len, seqs), dtype=np.int64, count=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    if off[-1] >= 2 ** 32:
        raise ValueError("batch exceeds 4 GiB of sequence; split it")
    if n == 0 or off[-1] == 0:
        return np.zeros(0, dtype=np.uint8), off.astype(np.uint32)
    try:  # lists of uint8 arrays (the common case: 1e5 reads per call) go through one C-level concatenate
        cat = np.concatenate(seqs)
        if cat.dtype != np.uint8 or cat.ndim != 1:
            raise TypeError
    except (TypeError, ValueError):
        cat = np.concatenate([_u8(s) for s in seqs])
    return np.ascontiguousarray(cat), off.astype(np.uint32)


class Context:
    """Owns one jtk_ctx (one CUDA device, one stream, growable device workspaces)."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        rc = lib().jtk_ctx_create(device, 0, C.byref(self._h))
        if rc != 0:
            raise JtkError(rc, lib().jtk_last_error(None).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().jtk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise JtkError(rc, lib().jtk_last_error(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(lib().jtk_ctx_launch_count(self._h))

    @property
    def last_kernel_ms(self) -> float:
        return float(lib().jtk_ctx_last_kernel_ms(self._h))

    @property
    def last_modtable_variant(self) -> str:
        """'fused' (DP matrices stay on chip) or 'rows' (forward rows parked in HBM) for the last table call, '' before it."""
        return {1: "fused", 2: "rows"}.get(int(lib().jtk_ctx_last_modtable_variant(self._h)), "")

    def timer_start(self):
        self._check(lib().jtk_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(lib().jtk_ctx_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def kernel_times(self) -> np.ndarray:
        buf = np.zeros(256, dtype=np.float32)
        n = lib().jtk_ctx_kernel_times(self._h, _ptr(buf), 256)
        if n < 0:
            self._check(n)
        return buf[:n].copy()

    def measure_fp32_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(lib().jtk_ctx_measure_fp32_peak(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def batch(self, templates, reads, ops, strands, tmpl_idx, radius) -> "Batch":
        return Batch(self, templates, reads, ops, strands, tmpl_idx, radius)

    # ---- level 1 -------------------------------------------------------------------------------
    def modtable_batch(self, fwd: HmmParams, rev: HmmParams, templates, reads, ops, strands, tmpl_idx, radius,
                       want_table=True):
        """jtk_hmm_modtable_batch.  Returns (lk float64[n], tables: list of float64[(Lt+1)*14] or None)."""
        n = len(reads)
        tcat, toff = concat(templates)
        rcat, roff = concat(reads)
        ocat, ooff = concat(ops)
        strands = _u8(strands)
        tmpl_idx = np.ascontiguousarray(tmpl_idx, dtype=np.uint32)
        lk = np.empty(n, dtype=np.float64)
        tab = tab_off = None
        if want_table:
            sizes = np.array([(len(templates[int(t)]) + 1) * NUM_ROW for t in tmpl_idx], dtype=np.uint64)
            tab_off = np.zeros(n + 1, dtype=np.uint64)
            np.cumsum(sizes, out=tab_off[1:])
            tab = np.empty(int(tab_off[-1]), dtype=np.float64)
        self._check(lib().jtk_hmm_modtable_batch(
            self._h, C.byref(fwd), C.byref(rev), n, len(templates), _ptr(tcat), _ptr(toff), _ptr(rcat), _ptr(roff),
            _ptr(ocat), _ptr(ooff), _ptr(strands), _ptr(tmpl_idx), radius, _ptr(lk), _ptr(tab), _ptr(tab_off)))
        tables = None
        if want_table:
            tables = [tab[int(tab_off[k]):int(tab_off[k + 1])] for k in range(n)]
        return lk, tables

    def likelihood_batch(self, fwd: HmmParams, rev: HmmParams, templates, reads, ops, strands, tmpl_idx, radius):
        """jtk_hmm_likelihood_batch; ops=None selects the bootstrap guide.  templates / reads / ops: lists of uint8 arrays or
        (concatenated uint8, uint32 offsets[n+1]) tuples (large calibration batches arrive flat)."""
        tcat, toff = concat(templates)
        rcat, roff = concat(reads)
        n = len(roff) - 1
        ocat = ooff = None
        if ops is not None:
            ocat, ooff = concat(ops)
        strands = _u8(strands)
        tmpl_idx = np.ascontiguousarray(tmpl_idx, dtype=np.uint32)
        lk = np.empty(n, dtype=np.float64)
        self._check(lib().jtk_hmm_likelihood_batch(
            self._h, C.byref(fwd), C.byref(rev), n, len(toff) - 1, _ptr(tcat), _ptr(toff), _ptr(rcat), _ptr(roff),
            _ptr(ocat), _ptr(ooff), _ptr(strands), _ptr(tmpl_idx), radius, _ptr(lk)))
        return lk


    def mcmc_restarts(self, datas, ks, covs, states, restarts: int = 20):
        """jtk_mcmc_restarts_batch: the k-means + MCMC restarts of pseudo_mcmc::mcmc_clustering for many chains at once.
        datas: list of float64[n, D]; ks, covs per chain; states: uint64[n_chains, 4] (updated in place).
        Returns (list of uint8[n] assignments, float64[n_chains] likelihoods, int32[n_chains] status)."""
        L = lib()
        nch = len(datas)
        ds = [np.ascontiguousarray(d, dtype=np.float64) for d in datas]
        n_rows = np.array([d.shape[0] for d in ds], dtype=np.uint32)
        n_cols = np.array([d.shape[1] for d in ds], dtype=np.uint32)
        kk = np.ascontiguousarray(ks, dtype=np.uint32)
        off = np.zeros(nch, dtype=np.uint64)
        if nch > 1:
            off[1:] = np.cumsum([d.size for d in ds[:-1]])
        data = np.concatenate([d.ravel() for d in ds]) if nch else np.zeros(0)
        s2l = []
        L.jtk_lc_size_to_lk.argtypes = [C.c_int, C.c_double, C.c_int, C.c_void_p]
        L.jtk_lc_size_to_lk.restype = None
        for d, k, cov in zip(ds, kk, covs):
            t = np.empty(d.shape[0] + 1, dtype=np.float64)
            L.jtk_lc_size_to_lk(int(d.shape[0]), float(cov), int(k), _ptr(t))
            s2l.append(t)
        s2l = np.concatenate(s2l) if nch else np.zeros(0)
        states = np.ascontiguousarray(states, dtype=np.uint64).reshape(nch, 4)
        asn = np.zeros(int(n_rows.sum()), dtype=np.uint8)
        lk = np.zeros(nch, dtype=np.float64)
        err = np.zeros(nch, dtype=np.int32)
        vp = C.c_void_p
        L.jtk_mcmc_restarts_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
        self._check(L.jtk_mcmc_restarts_batch(self._h, nch, _ptr(data), _ptr(off), _ptr(n_rows), _ptr(n_cols), _ptr(kk), _ptr(s2l),
                                              restarts, _ptr(states), _ptr(asn), _ptr(lk), _ptr(err)))
        pos = np.concatenate([[0], np.cumsum(n_rows)]).astype(np.int64)
        return [asn[pos[c]:pos[c + 1]].copy() for c in range(nch)], lk, err, states

class Batch:
    """jtk_batch: chunks + reads resident in HBM (level 2 of include/jtk_gpu.h)."""

    def __init__(self, ctx: Context, templates, reads, ops, strands, tmpl_idx, radius, packed=None):
        self.ctx = ctx
        self.n_pairs = len(reads)
        self.n_tmpl = len(templates)
        self.tmpl_len = np.array([len(t) for t in templates], dtype=np.int64)
        self.tmpl_idx = np.ascontiguousarray(tmpl_idx, dtype=np.uint32)
        self.packed = packed if packed is not None else pack_inputs(templates, reads, ops, strands, tmpl_idx)
        tcat, toff, rcat, roff, ocat, ooff, st, ti = self.packed
        self._h = C.c_void_p()
        ctx._check(lib().jtk_batch_create(ctx._h, self.n_pairs, self.n_tmpl, _ptr(tcat), _ptr(toff), _ptr(rcat),
                                          _ptr(roff), _ptr(ocat), _ptr(ooff), _ptr(st), _ptr(ti), radius,
                                          C.byref(self._h)))
        sizes = (self.tmpl_len + 1) * NUM_ROW
        self.stat_off = np.zeros(self.n_tmpl + 1, dtype=np.uint64)
        np.cumsum(sizes, out=self.stat_off[1:])

    def close(self):
        if getattr(self, "_h", None):
            lib().jtk_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def cell_updates(self) -> int:
        return int(lib().jtk_batch_cell_updates(self._h))

    @property
    def h2d_bytes(self) -> int:
        return int(lib().jtk_batch_h2d_bytes(self._h))

    def modtable(self, fwd: HmmParams, rev: HmmParams, rows: int = 14):
        self.ctx._check(lib().jtk_batch_modtable(self._h, C.byref(fwd), C.byref(rev), rows))

    def sync(self):
        self.ctx._check(lib().jtk_batch_sync(self._h))

    def lk(self) -> np.ndarray:
        out = np.empty(self.n_pairs, dtype=np.float64)
        self.ctx._check(lib().jtk_batch_fetch_lk(self._h, _ptr(out)))
        return out

    def profile(self, pair: int) -> np.ndarray:
        t = int(self.tmpl_idx[pair])
        out = np.empty((int(self.tmpl_len[t]) + 1) * NUM_ROW, dtype=np.float32)
        self.ctx._check(lib().jtk_batch_fetch_profile(self._h, pair, _ptr(out)))
        return out

    def colstats(self, min_req: np.ndarray, pos_thr: float = 1e-5, fetch: bool = True):
        """min_req: float32[3, H] (rows Subst, Del, Ins).  Returns a structured array (COLSTAT_DTYPE) with all
        templates concatenated (template t at self.stat_off[t]), or None when fetch=False."""
        mr = np.ascontiguousarray(min_req, dtype=np.float32)
        assert mr.ndim == 2 and mr.shape[0] == 3
        out = np.empty(int(self.stat_off[-1]), dtype=COLSTAT_DTYPE) if fetch else None
        self.ctx._check(lib().jtk_batch_colstats(self._h, _ptr(mr), mr.shape[1], pos_thr, _ptr(out),
                                                 _ptr(self.stat_off)))
        return out

    def _gains_c(self, gain, prob):
        self._g_gain = np.ascontiguousarray(gain, dtype=np.float64)
        self._g_prob = np.ascontiguousarray(prob, dtype=np.float64)
        assert self._g_gain.shape == self._g_prob.shape and self._g_gain.shape[0] == 3
        return CGains(self._g_gain.shape[1], self._g_gain.ctypes.data, self._g_prob.ctypes.data)

    def candidates(self, gain, prob, copy_num, coverage: float, cap: int = 0) -> np.ndarray:
        """jtk_batch_candidates: filter_profiles on the device; structured array (CANDIDATE_DTYPE) sorted by (tmpl, pos)."""
        g = self._gains_c(gain, prob)
        cn = np.ascontiguousarray(np.broadcast_to(copy_num, (self.n_tmpl,)), dtype=np.int32)
        cap = cap or 256 * self.n_tmpl + 4096
        out = np.zeros(cap, dtype=CANDIDATE_DTYPE)
        n = C.c_int()
        self.ctx._check(lib().jtk_batch_candidates(self._h, C.byref(g), _ptr(cn), coverage, _ptr(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def search_variants(self, gain, prob, copy_num, coverage: float, probe_cap: int = 0):
        """jtk_batch_search_variants: (n_probes uint32[n_tmpl], probe_pos uint32[n_tmpl, cap], variants float64[n_pairs, cap])."""
        g = self._gains_c(gain, prob)
        cn = np.ascontiguousarray(np.broadcast_to(copy_num, (self.n_tmpl,)), dtype=np.int32)
        cap = probe_cap or 3 * max(2, int(cn.max()) if len(cn) else 2)
        n_probes = np.zeros(self.n_tmpl, dtype=np.uint32)
        probe_pos = np.zeros((self.n_tmpl, cap), dtype=np.uint32)
        variants = np.zeros((self.n_pairs, cap), dtype=np.float64)
        self.ctx._check(lib().jtk_batch_search_variants(self._h, C.byref(g), _ptr(cn), coverage, cap, _ptr(n_probes),
                                                        _ptr(probe_pos), _ptr(variants)))
        return n_probes, probe_pos, variants

    def gather(self, tmpl: int, min_req: np.ndarray, cols) -> np.ndarray:
        mr = np.ascontiguousarray(min_req, dtype=np.float32)
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        n_reads = int((self.tmpl_idx == tmpl).sum())
        out = np.zeros((n_reads, len(cols)), dtype=np.float64)
        self.ctx._check(lib().jtk_batch_gather(self._h, tmpl, _ptr(mr), mr.shape[1], _ptr(cols), len(cols), _ptr(out)))
        return out


def pack_inputs(templates, reads, ops, strands, tmpl_idx):
    """Concatenate host inputs into the flat arrays the C ABI takes."""
    tcat, toff = concat(templates)
    rcat, roff = concat(reads)
    ocat, ooff = concat(ops)
    return (tcat, toff, rcat, roff, ocat, ooff, _u8(strands), np.ascontiguousarray(tmpl_idx, dtype=np.uint32))


def band_cell_count(ops, Lt: int, Lr: int, radius: int) -> int:
    ops = _u8(ops)
    return int(lib().jtk_band_cell_count(_ptr(ops), len(ops), Lt, Lr, radius))


def gen_reads(hmm45: np.ndarray, sources, src_idx, seed: int, cap: int):
    """jtk_lc_gen_reads: read k sampled from the pair HMM `hmm45` (45 doubles, HMMParam order) along sources[src_idx[k]], own
    generator stream seed + k, on the host threads.  Returns (uint8[n, cap], uint32[n] lengths)."""
    L = lib()
    vp = C.c_void_p
    L.jtk_lc_gen_reads.argtypes = [vp, C.c_int, vp, vp, vp, C.c_uint64, C.c_int, vp, vp]
    scat, soff = concat(sources)
    soff = soff.astype(np.uint64)
    src_idx = np.ascontiguousarray(src_idx, dtype=np.uint32)
    h = np.ascontiguousarray(hmm45, dtype=np.float64)
    n = len(src_idx)
    out = np.empty((n, cap), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint32)
    rc = L.jtk_lc_gen_reads(_ptr(h), n, _ptr(scat), _ptr(soff), _ptr(src_idx), seed, cap, _ptr(out), _ptr(lens))
    if rc != 0:
        raise JtkError(rc, "jtk_lc_gen_reads")
    return out, lens
