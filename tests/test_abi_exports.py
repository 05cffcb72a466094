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


def test_struct_layouts_agree_across_header_library_and_python_mirrors():
    """Every struct that crosses the ABI: sizeof as compiled into the library == the static_assert pins of
    include/jtk_gpu.h == the ctypes / numpy mirrors the Python host uses (and the #[repr(C)] text of jtk-gpu-sys)."""
    import ctypes as C
    from jtk_b200 import _lib, hmm, local_clustering as LC
    L = _lib.lib()
    L.jtk_abi_sizeof.restype = C.c_size_t
    L.jtk_abi_sizeof.argtypes = [C.c_char_p]
    want = {"jtk_hmm_params": 360, "jtk_colstat": 24, "jtk_candidate": 32, "jtk_gains": 24, "jtk_clustering_config": 24,
            "jtk_polish_config": 12}
    for name, size in want.items():
        assert L.jtk_abi_sizeof(name.encode()) == size, name
    assert L.jtk_abi_sizeof(b"nonsense") == 0
    assert C.sizeof(_lib.HmmParams) == 360 and _lib.HmmParams.mat_emit.offset == 72 and _lib.HmmParams.ins_emit.offset == 200
    assert _lib.COLSTAT_DTYPE.itemsize == 24 and _lib.COLSTAT_DTYPE.fields["sc"][1] == 12
    assert _lib.CANDIDATE_DTYPE.itemsize == 32 and _lib.CANDIDATE_DTYPE.fields["lk"][1] == 24
    assert C.sizeof(_lib.CGains) == 24 and C.sizeof(LC._CGains) == 24 and LC._CGains.prob.offset == 16
    assert C.sizeof(LC._CConfig) == 24 and LC._CConfig.coverage.offset == 8
    assert C.sizeof(hmm.PolishCfg) == 12
    src = open(os.path.join(ROOT, "jtk-gpu-sys", "src", "lib.rs")).read()
    for name in want:
        assert re.search(r"#\[repr\(C\)\]\s*(#\[derive\([^)]*\)\]\s*)?pub struct " + "".join(w.capitalize() for w in name.split("_")[1:]), src), name


def test_rust_shim_declarations_match_the_header():
    """jtk-gpu-sys/src/lib.rs cannot be compiled here (no Rust toolchain), so it is checked mechanically: every extern
    function it declares exists in include/jtk_gpu.h with the same number of parameters, and the four kiley-named wrappers
     are written out."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "jtk_gpu.h")).read(), flags=re.S)
    rs = open(os.path.join(ROOT, "jtk-gpu-sys", "src", "lib.rs")).read()
    rs = re.sub(r"//.*", "", rs)

    def nargs(text):
        text = text.strip()
        if text in ("", "void"):
            return 0
        return len([a for a in text.split(",") if a.strip()])
    c_decl = {m.group(1): nargs(m.group(2)) for m in re.finditer(r"\b(jtk_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", hdr)}
    rust = {m.group(1): nargs(m.group(2)) for m in re.finditer(r"pub fn (jtk_[a-z0-9_]+)\s*\(([^()]*)\)", rs)}
    assert len(rust) >= 20
    for name, n in rust.items():
        assert name in c_decl, name
        assert c_decl[name] == n, (name, c_decl[name], n)
    for wrapper in ("modification_table_antidiagonal", "likelihood_antidiagonal_bootstrap",
                    "polish_until_converge_antidiagonal", "fit_antidiagonal_par_multiple"):
        assert re.search(r"pub fn " + wrapper + r"\b", rs), wrapper
