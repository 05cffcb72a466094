"""ctypes binding of the CPU oracle (oracle/libphmm_oracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
NUM_ROW = 14


class OrcHmm(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("mat_mat", "mat_ins", "mat_del", "ins_mat", "ins_ins", "ins_del", "del_mat", "del_ins", "del_del")] + \
               [("mat_emit", C.c_double * 16), ("ins_emit", C.c_double * 20)]

    def as_array(self) -> np.ndarray:
        return np.frombuffer(bytes(self), dtype=np.float64).copy()

    @classmethod
    def from_array(cls, a) -> "OrcHmm":
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == 45
        return cls.from_buffer_copy(a.tobytes())


class OrcPair(C.Structure):
    _fields_ = [("t", C.c_void_p), ("Lt", C.c_int), ("q", C.c_void_p), ("Lr", C.c_int),
                ("ops", C.c_void_p), ("n_ops", C.c_int), ("strand", C.c_int),
                ("table", C.c_void_p), ("lk", C.c_double)]


class OrcPolishCfg(C.Structure):
    _fields_ = [("radius", C.c_int), ("take_num", C.c_int), ("ignore_edge", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "libphmm_oracle.so")
        src = os.path.join(ORACLE_DIR, "phmm_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        L = C.CDLL(so)
        u8p = C.c_void_p
        L.orc_hmm_default.argtypes = [C.POINTER(OrcHmm)]
        L.orc_likelihood.restype = C.c_double
        L.orc_likelihood.argtypes = [C.POINTER(OrcHmm), u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int]
        L.orc_likelihood_backward.restype = C.c_double
        L.orc_likelihood_backward.argtypes = L.orc_likelihood.argtypes
        L.orc_modification_table.argtypes = [C.POINTER(OrcHmm), u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int,
                                             C.c_void_p, C.POINTER(C.c_double)]
        L.orc_apply_edit.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_edit_ops.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, u8p]
        L.orc_likelihood_bootstrap.restype = C.c_double
        L.orc_likelihood_bootstrap.argtypes = [C.POINTER(OrcHmm), u8p, C.c_int, u8p, C.c_int, C.c_int]
        L.orc_cell_count.restype = C.c_int64
        L.orc_cell_count.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_band_centres.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_modification_table_batch.argtypes = [C.POINTER(OrcHmm), C.POINTER(OrcHmm), C.POINTER(OrcPair), C.c_int,
                                                   C.c_int, C.c_int]
        L.orc_expected_counts.argtypes = [C.POINTER(OrcHmm), u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int,
                                          C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def default_hmm() -> OrcHmm:
    h = OrcHmm()
    lib().orc_hmm_default(C.byref(h))
    return h


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def likelihood(h, t, q, ops, radius, backward=False) -> float:
    t, q, ops = _u8(t), _u8(q), _u8(ops)
    f = lib().orc_likelihood_backward if backward else lib().orc_likelihood
    return f(C.byref(h), _p(t), len(t), _p(q), len(q), _p(ops), len(ops), radius)


def modification_table(h, t, q, ops, radius):
    t, q, ops = _u8(t), _u8(q), _u8(ops)
    tab = np.empty((len(t) + 1) * NUM_ROW, dtype=np.float64)
    lk = C.c_double()
    rc = lib().orc_modification_table(C.byref(h), _p(t), len(t), _p(q), len(q), _p(ops), len(ops), radius,
                                      _p(tab), C.byref(lk))
    if rc:
        raise ValueError(f"orc_modification_table rc={rc}")
    return tab, lk.value


def apply_edit(t, j, row):
    t = _u8(t)
    out = np.empty(len(t) + 4, dtype=np.uint8)
    n = lib().orc_apply_edit(_p(t), len(t), j, row, _p(out))
    return None if n < 0 else out[:n].copy()


def edit_ops(t, q, radius):
    t, q = _u8(t), _u8(q)
    out = np.empty(len(t) + len(q) + 1, dtype=np.uint8)
    n = lib().orc_edit_ops(_p(t), len(t), _p(q), len(q), radius, _p(out))
    if n < 0:
        raise ValueError("band cannot connect corners")
    return out[:n].copy()


def likelihood_bootstrap(h, t, q, radius) -> float:
    t, q = _u8(t), _u8(q)
    return lib().orc_likelihood_bootstrap(C.byref(h), _p(t), len(t), _p(q), len(q), radius)


def cell_count(ops, Lt, Lr, radius) -> int:
    ops = _u8(ops)
    return lib().orc_cell_count(_p(ops), len(ops), Lt, Lr, radius)


def expected_counts(h, t, q, ops, radius) -> np.ndarray:
    t, q, ops = _u8(t), _u8(q), _u8(ops)
    acc = np.zeros(45, dtype=np.float64)
    rc = lib().orc_expected_counts(C.byref(h), _p(t), len(t), _p(q), len(q), _p(ops), len(ops), radius, _p(acc))
    if rc:
        raise ValueError(f"orc_expected_counts rc={rc}")
    return acc


def modification_table_batch(fwd, rev, templates, reads, ops, strands, radius, n_threads=1, want_tables=True):
    """templates[k], reads[k], ops[k] per pair.  Returns (list of tables or None, lks)."""
    n = len(reads)
    pairs = (OrcPair * n)()
    keep = []
    tables = []
    for k in range(n):
        t, q, o = _u8(templates[k]), _u8(reads[k]), _u8(ops[k])
        keep += [t, q, o]
        pairs[k].t, pairs[k].Lt = t.ctypes.data, len(t)
        pairs[k].q, pairs[k].Lr = q.ctypes.data, len(q)
        pairs[k].ops, pairs[k].n_ops = o.ctypes.data, len(o)
        pairs[k].strand = int(strands[k])
        if want_tables:
            tab = np.empty((len(t) + 1) * NUM_ROW, dtype=np.float64)
            tables.append(tab)
            pairs[k].table = tab.ctypes.data
        else:
            pairs[k].table = None
    rc = lib().orc_modification_table_batch(C.byref(fwd), C.byref(rev), pairs, n, radius, n_threads)
    if rc:
        raise ValueError(f"batch rc={rc}")
    lks = np.array([pairs[k].lk for k in range(n)])
    return (tables if want_tables else None), lks


def polish_until_converge(fwd, rev, draft, reads, ops, strands, radius, take_num, ignore_edge):
    """orc_polish_until_converge.  Returns (consensus uint8[], ops list, iterations)."""
    L = lib()
    draft = _u8(draft)
    n = len(reads)
    reads = [_u8(r) for r in reads]
    cap = max(len(r) for r in reads) + 2 * len(draft) + 64
    bufs = [np.zeros(cap, dtype=np.uint8) for _ in range(n)]
    n_ops = (C.c_int * n)()
    for k, o in enumerate(ops):
        o = _u8(o)
        bufs[k][:len(o)] = o
        n_ops[k] = len(o)
    rp = (C.c_void_p * n)(*[r.ctypes.data for r in reads])
    rl = (C.c_int * n)(*[len(r) for r in reads])
    op = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
    st = _u8(np.asarray(strands, dtype=np.uint8))
    cfg = OrcPolishCfg(radius, take_num, ignore_edge)
    cons_cap = 2 * len(draft) + 64
    out = np.zeros(cons_cap + 16, dtype=np.uint8)
    it = C.c_int()
    L.orc_polish_until_converge.argtypes = [C.POINTER(OrcHmm), C.POINTER(OrcHmm), C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(OrcPolishCfg),
                                            C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    rc = L.orc_polish_until_converge(C.byref(fwd), C.byref(rev), _p(draft), len(draft), n, rp, rl, op, n_ops, cap, _p(st),
                                     C.byref(cfg), _p(out), cons_cap, C.byref(it))
    if rc < 0:
        raise ValueError(f"orc_polish_until_converge rc={rc}")
    return out[:rc].copy(), [bufs[k][:n_ops[k]].copy() for k in range(n)], it.value


def fit(fwd, rev, packs, radius):
    """orc_fit on packs [(template, strands, reads, ops), ...]; returns the two updated OrcHmm."""
    f = OrcHmm.from_array(fwd.as_array())
    r = OrcHmm.from_array(rev.as_array())
    accf, accr = np.zeros(45), np.zeros(45)
    for (t, strands, reads, ops) in packs:
        for q, o, st in zip(reads, ops, strands):
            acc = expected_counts(f if st else r, t, q, o, radius)
            if st:
                accf += acc
            else:
                accr += acc

    def mstep(h, acc):
        a = h.as_array()
        for s in range(3):
            tot = acc[3 * s:3 * s + 3].sum()
            if tot > 0:
                a[3 * s:3 * s + 3] = acc[3 * s:3 * s + 3] / tot
        for k in range(4):
            tot = acc[9 + 4 * k:13 + 4 * k].sum()
            if tot > 0:
                a[9 + 4 * k:13 + 4 * k] = acc[9 + 4 * k:13 + 4 * k] / tot
        for c in range(5):
            tot = acc[25 + 4 * c:29 + 4 * c].sum()
            if tot > 0:
                a[25 + 4 * c:29 + 4 * c] = acc[25 + 4 * c:29 + 4 * c] / tot
        return OrcHmm.from_array(a)
    return mstep(f, accf), mstep(r, accr), accf, accr
