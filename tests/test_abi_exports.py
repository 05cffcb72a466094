"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/jtk_gpu.h declares, and compute calls fail loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "jtk_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jtk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from jtk_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/jtk_gpu.h but not exported"
    assert set(_lib.EXPORTS) <= set(names)
    assert L.jtk_hmm_num_row() == 14 and L.jtk_hmm_copy_size() == 3 and L.jtk_hmm_del_size() == 3


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from jtk_b200 import _lib
    with pytest.raises(_lib.JtkError) as e:
        _lib.Context()
    assert e.value.code == -2  # JTK_ECUDA


def test_band_cell_count_matches_oracle():
    import numpy as np
    import oracle_lib as O
    from jtk_b200 import _lib, synth
    rng = np.random.default_rng(3)
    t = synth.random_template(rng, 300)
    q, ops = synth.mutate_read(rng, t, 0.12)
    for R in (2, 10, 30):
        assert _lib.band_cell_count(ops, len(t), len(q), R) == O.cell_count(ops, len(t), len(q), R)
    assert _lib.band_cell_count(ops[:-1], len(t), len(q), 5) == -1
