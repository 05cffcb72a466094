"""GPU parity: the CUDA path (through the C ABI) against the f64 oracle on identical seeded inputs.

Tolerances (stated per BASELINE.json north_star: |delta log-lik| <= 1e-3 relative; we hold much tighter):
  lk                      |gpu - oracle| <= 2e-5 * |lk|  (fp32 forward, f64 exponent bookkeeping)
  table - lk (per entry)  |gpu - oracle| <= 2e-3 absolute, and <= 1e-3 relative for |delta| >= 1
  identity substitution   exactly 0
"""
import numpy as np
import pytest

import oracle_lib as O
from jtk_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("modtable_variant")]

NEG = -1.0e9


@pytest.fixture(scope="module")
def ctx():
    from jtk_b200 import _lib
    c = _lib.Context()
    yield c
    c.close()


def to_c(h):
    from jtk_b200 import _lib
    return _lib.HmmParams.from_buffer_copy(bytes(h))


def random_hmm(seed):
    rng = np.random.default_rng(seed)
    a = np.empty(45)
    for s in range(3):
        a[3 * s:3 * s + 3] = rng.dirichlet([30, 2, 2]) * 0.99
    for r in range(4):
        e = rng.dirichlet([1, 1, 1, 1]) * 0.2
        e[r] += 0.8
        a[9 + 4 * r:13 + 4 * r] = e
    for c in range(5):
        a[25 + 4 * c:29 + 4 * c] = rng.dirichlet([3, 3, 3, 3])
    return O.OrcHmm.from_array(a)


def check_tables(gpu_tabs, gpu_lk, orc_tabs, orc_lk, templates, tmpl_idx):
    worst = 0.0
    for k in range(len(gpu_tabs)):
        g = gpu_tabs[k].reshape(-1, 14)
        o = orc_tabs[k].reshape(-1, 14)
        assert abs(gpu_lk[k] - orc_lk[k]) <= 2e-5 * abs(orc_lk[k]) + 1e-6, (k, gpu_lk[k], orc_lk[k])
        gneg, oneg = g < NEG, o < NEG
        # impossible edits agree, except entries whose probability underflows fp32 (delta < -80)
        dis = gneg != oneg
        if dis.any():
            od = o - orc_lk[k]
            assert (od[dis & ~oneg] < -60).all() and not (dis & oneg).any(), (k, np.argwhere(dis)[:5])
        ok = ~gneg & ~oneg
        gd = g - gpu_lk[k]
        od = o - orc_lk[k]
        err = np.abs(gd - od)[ok]
        tol = np.maximum(2e-3, 1e-3 * np.abs(od[ok]))
        assert (err <= tol).all(), (k, float(err.max()), np.argwhere(np.abs(gd - od) * ok > 2e-3)[:5])
        worst = max(worst, float(err.max()) if err.size else 0.0)
        t = templates[int(tmpl_idx[k])]
        code = np.searchsorted(synth.ACGT, t)
        assert (gd[np.arange(len(t)), code] == 0.0).all()
    return worst


@pytest.mark.parametrize("L,err,R,model", [
    (40, 0.15, 3, "default"), (40, 0.15, 14, "random"), (150, 0.1, 5, "random"), (150, 0.1, 30, "default"),
    (300, 0.08, 10, "random"), (300, 0.2, 30, "random"), (120, 0.1, 50, "random"), (90, 0.0, 30, "default"),
])
def test_modtable_small_vs_oracle(ctx, L, err, R, model):
    rng = np.random.default_rng(L * 1000 + R)
    fwd = O.default_hmm() if model == "default" else random_hmm(1)
    rev = O.default_hmm() if model == "default" else random_hmm(2)
    templates = [synth.random_template(rng, L), synth.random_template(rng, L + 17)]
    reads, ops, strands, tidx = [], [], [], []
    for k in range(12):
        ti = k % 2
        q, o = synth.mutate_read(rng, templates[ti], err)
        reads.append(q); ops.append(o); strands.append(k % 3 != 0); tidx.append(ti)
    lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), templates, reads, ops, strands, tidx, R)
    otabs, olk = O.modification_table_batch(fwd, rev, [templates[i] for i in tidx], reads, ops, strands, R)
    check_tables(tabs, lk, otabs, olk, templates, tidx)


def test_modtable_chunk_config0_vs_oracle(ctx):
    """BASELINE.json configs[0]: 2 kbp chunk, 60 reads at ~8 % error, radius 30 (ONT)."""
    d = synth.diploid_chunk(7)
    fwd, rev = random_hmm(11), random_hmm(12)
    n = 16  # the oracle needs ~30 ms per pair
    reads, ops, strands = d["reads"][:n], d["ops"][:n], d["strands"][:n]
    lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), [d["template"]], reads, ops, strands, np.zeros(n, np.uint32), 30)
    otabs, olk = O.modification_table_batch(fwd, rev, [d["template"]] * n, reads, ops, strands, 30, n_threads=4)
    worst = check_tables(tabs, lk, otabs, olk, [d["template"]], np.zeros(n, np.uint32))
    assert worst < 2e-3


@pytest.mark.parametrize("L,err,R,n", [
    (300, 0.15, 63, 6), (700, 0.12, 80, 6), (1500, 0.15, 100, 4), (4000, 0.12, 100, 2), (8000, 0.08, 120, 2),
])
def test_modtable_wide_band_vs_oracle(ctx, L, err, R, n):
    """Radius 63..126 (8 column slots per lane): the reference band is ceil(frac * L) / 2 (definitions/src/lib.rs:173-175,
    201-210; local_clustering/mod.rs:105,112) -- ONT at 8 kbp gives radius 120, CLR at 4 kbp radius 100, and BASELINE.json
    configs[3] sweeps the band to 100."""
    rng = np.random.default_rng(L + R)
    fwd, rev = random_hmm(21), random_hmm(22)
    t = synth.random_template(rng, L)
    reads, ops = [], []
    for _ in range(n):
        q, o = synth.mutate_read(rng, t, err)
        reads.append(q); ops.append(o)
    strands = (np.arange(n) % 2).astype(np.uint8)
    tidx = np.zeros(n, np.uint32)
    lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), [t], reads, ops, strands, tidx, R)
    otabs, olk = O.modification_table_batch(fwd, rev, [t] * n, reads, ops, strands, R, n_threads=4)
    check_tables(tabs, lk, otabs, olk, [t], tidx)
    # likelihood-only kernel on the same pairs (guided)
    lk2 = ctx.likelihood_batch(to_c(fwd), to_c(rev), [t], reads, ops, strands, tidx, R)
    assert np.allclose(lk2, olk, rtol=2e-5, atol=1e-6)


def test_band_sweep_config3_vs_oracle(ctx):
    """BASELINE.json configs[3] sweeps the band width: radius 10, 20, 30, 50 and 100 on the same pile-up; a wider band can
    only add paths, so the likelihood is non-decreasing in the radius, and every radius matches the oracle."""
    rng = np.random.default_rng(404)
    fwd, rev = random_hmm(31), random_hmm(32)
    t = synth.random_template(rng, 1000)
    reads, ops = [], []
    for _ in range(4):
        q, o = synth.mutate_read(rng, t, 0.15)
        reads.append(q); ops.append(o)
    strands = np.array([1, 0, 1, 0], np.uint8)
    tidx = np.zeros(4, np.uint32)
    prev = None
    for R in (10, 20, 30, 50, 100):
        lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), [t], reads, ops, strands, tidx, R)
        otabs, olk = O.modification_table_batch(fwd, rev, [t] * 4, reads, ops, strands, R, n_threads=4)
        check_tables(tabs, lk, otabs, olk, [t], tidx)
        if prev is not None:
            assert (lk >= prev - 1e-4 * np.abs(prev)).all()
        prev = lk


def test_likelihood_guided_and_bootstrap(ctx):
    rng = np.random.default_rng(77)
    fwd, rev = random_hmm(5), random_hmm(6)
    templates, reads, ops, strands = [], [], [], []
    for k in range(40):
        t = synth.random_template(rng, 103)
        q, o = synth.mutate_read(rng, t, 0.1)
        templates.append(t); reads.append(q); ops.append(o); strands.append(k % 2 == 0)
    tidx = np.arange(40, dtype=np.uint32)
    for R in (10, 25):
        lk = ctx.likelihood_batch(to_c(fwd), to_c(rev), templates, reads, ops, strands, tidx, R)
        lkb = ctx.likelihood_batch(to_c(fwd), to_c(rev), templates, reads, None, strands, tidx, R)
        for k in range(40):
            h = fwd if strands[k] else rev
            want = O.likelihood(h, templates[k], reads[k], ops[k], R)
            wantb = O.likelihood_bootstrap(h, templates[k], reads[k], R)
            assert abs(lk[k] - want) <= 2e-5 * abs(want)
            assert abs(lkb[k] - wantb) <= 2e-5 * abs(wantb)


def test_ragged_tiny_and_long_indel(ctx):
    """Edge cases the domain has: very short / very unequal sequences, a long insertion run (forces the
    rescale path), reads that are pure deletions of the template."""
    fwd = rev = O.default_hmm()
    cases = []
    for t, q in ((b"ACGTACGTAC", b"AC"), (b"AC", b"ACGTTTTTGA"), (b"A", b"A"), (b"A", b"C"), (b"ACGTTGCA", b"ACGTTGCA")):
        t = np.frombuffer(t, np.uint8); q = np.frombuffer(q, np.uint8)
        cases.append((t, q, O.edit_ops(t, q, 20)))
    rng = np.random.default_rng(5)
    t = synth.random_template(rng, 400)
    ins = synth.random_template(rng, 25)
    q = np.concatenate([t[:200], ins, t[200:]])
    o = np.concatenate([np.zeros(200, np.uint8), np.full(25, 2, np.uint8), np.zeros(200, np.uint8)])
    cases.append((t, q, o))
    q2 = np.concatenate([t[:150], t[180:]])
    o2 = np.concatenate([np.zeros(150, np.uint8), np.full(30, 3, np.uint8), np.zeros(220, np.uint8)])
    cases.append((t, q2, o2))
    templates = [c[0] for c in cases]
    reads = [c[1] for c in cases]
    ops = [c[2] for c in cases]
    n = len(cases)
    lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), templates, reads, ops, np.ones(n, np.uint8), np.arange(n), 30)
    otabs, olk = O.modification_table_batch(fwd, rev, templates, reads, ops, np.ones(n, np.uint8), 30)
    check_tables(tabs, lk, otabs, olk, templates, np.arange(n))


def test_bad_ops_and_radius_are_rejected(ctx):
    from jtk_b200 import _lib
    h = to_c(O.default_hmm())
    t = np.frombuffer(b"ACGT", np.uint8)
    with pytest.raises(_lib.JtkError) as e:
        ctx.modtable_batch(h, h, [t], [t], [np.zeros(3, np.uint8)], [1], [0], 5)
    assert e.value.code == -1
    with pytest.raises(_lib.JtkError):
        ctx.modtable_batch(h, h, [t], [t], [np.zeros(4, np.uint8)], [1], [0], 127)
    assert ctx.modtable_batch(h, h, [t], [], [], [], [], 5)[0].size == 0


def test_kiley_shaped_single_call(ctx):
    """hmm.PairHiddenMarkovModel mirrors the kiley method the reference calls (pseudo_mcmc.rs:62-63)."""
    from jtk_b200 import hmm
    rng = np.random.default_rng(2)
    t = synth.random_template(rng, 200)
    q, o = synth.mutate_read(rng, t, 0.1)
    m = hmm.PairHiddenMarkovModel()
    table, lk = m.modification_table_antidiagonal(t, q, o, 20, ctx=ctx)
    otab, olk = O.modification_table(O.default_hmm(), t, q, o, 20)
    assert table.shape == ((len(t) + 1) * hmm.NUM_ROW,)
    assert abs(lk - olk) < 2e-5 * abs(olk)
    ok = otab > NEG
    assert np.max(np.abs((table - lk) - (otab - olk))[ok]) < 2e-3
    lkb = m.likelihood_antidiagonal_bootstrap(t, q, 20, ctx=ctx)
    assert abs(lkb - O.likelihood_bootstrap(O.default_hmm(), t, q, 20)) < 2e-5 * abs(olk)


def test_very_long_pair_takes_the_generic_backward_path(ctx):
    """Maximum sizes: a 8.4 kbp chunk (4x the pipeline's chunk length) has more anti-diagonals than the rescale-event
    map covers, so every backward block takes the generic (exact-correction) step; results still match the oracle."""
    rng = np.random.default_rng(99)
    t = synth.random_template(rng, 8400)
    fwd, rev = O.default_hmm(), random_hmm(3)
    reads, ops = [], []
    for _ in range(3):
        q, o = synth.mutate_read(rng, t, 0.08)
        reads.append(q); ops.append(o)
    assert len(t) + min(len(q) for q in reads) + 1 > 4 * 4096
    strands = np.array([1, 0, 1], np.uint8)
    lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), [t], reads, ops, strands, np.zeros(3, np.uint32), 30)
    otabs, olk = O.modification_table_batch(fwd, rev, [t] * 3, reads, ops, strands, 30, n_threads=3)
    check_tables(tabs, lk, otabs, olk, [t], np.zeros(3, np.uint32))


def test_full_size_properties_without_the_oracle(ctx):
    """BASELINE.json full size (2 kbp x 60 reads x several chunks) through size-independent properties: the identity
    substitution is exactly 0 in every profile, the 9-row kernel equals the 14-row kernel on the rows it computes, both
    agree on the likelihood, and rows that are impossible by construction are flagged as such."""
    chunks = synth.diploid_region(5, 6, length=2000, n_reads=60, error_rate=0.08)
    templates = [c["template"] for c in chunks]
    reads = [r for c in chunks for r in c["reads"]]
    ops = [o for c in chunks for o in c["ops"]]
    strands = np.concatenate([c["strands"] for c in chunks])
    tidx = np.repeat(np.arange(6, dtype=np.uint32), 60)
    h = to_c(O.default_hmm())
    b = ctx.batch(templates, reads, ops, strands, tidx, 30)
    b.modtable(h, h, 14)
    lk14 = b.lk()
    prof14 = {p: b.profile(p).reshape(-1, 14) for p in (0, 59, 60, 201, 359)}
    b.modtable(h, h, 9)
    lk9 = b.lk()
    assert np.array_equal(lk14, lk9)
    assert (lk14 < -200).all() and (lk14 > -3000).all()
    for p, a in prof14.items():
        c = b.profile(p).reshape(-1, 14)
        rows = [0, 1, 2, 3, 4, 5, 6, 7, 11]
        assert np.array_equal(a[:, rows], c[:, rows])
        assert (c[:, [8, 9, 10, 12, 13]] < -1e9).all()
        t = templates[int(tidx[p])]
        code = np.searchsorted(synth.ACGT, t)
        assert (a[np.arange(len(t)), code] == 0.0).all()
        assert (a[len(t), :4] < -1e9).all() and (a[len(t) - 1, 12:] < -1e9).all()  # no base to substitute / too few to delete
        # a read of its own haplotype rarely gains more than a few nats from any single edit, and never loses nothing
        assert a[a > -1e9].max() < 40 and a[a > -1e9].min() < -5
    b.close()


def test_parameter_uploads_follow_their_contents(ctx):
    """Models, min_req and stat_off are uploaded only when their bytes change (csrc/jtk_gpu_api.cu, SmallUpload): a call
    with NEW parameters on a context that has seen others must use the new ones, and going back must reproduce the first
    answer bit for bit."""
    from jtk_b200 import _lib
    d = synth.diploid_chunk(21, length=400, n_reads=10)
    tidx = np.zeros(len(d["reads"]), np.uint32)
    ha, hb = to_c(O.default_hmm()), to_c(random_hmm(5))
    b = ctx.batch([d["template"]], d["reads"], d["ops"], d["strands"], tidx, 30)
    b.modtable(ha, ha, 14); lk_a = b.lk(); pa = b.profile(3)
    b.modtable(hb, hb, 14); lk_b = b.lk(); pb = b.profile(3)
    b.modtable(ha, ha, 14); lk_a2 = b.lk(); pa2 = b.profile(3)
    assert np.array_equal(lk_a, lk_a2) and np.array_equal(pa, pa2)
    assert not np.array_equal(lk_a, lk_b) and not np.array_equal(pa, pb)
    fresh = _lib.Context()
    fb = fresh.batch([d["template"]], d["reads"], d["ops"], d["strands"], tidx, 30)
    fb.modtable(hb, hb, 14)
    assert np.array_equal(fb.lk(), lk_b) and np.array_equal(fb.profile(3), pb)
    # thresholds: two different min_req blocks on the same batch / context, then the first one again
    m1 = np.full((3, 4), 0.5, np.float32); m2 = np.full((3, 4), 3.0, np.float32)
    s1 = b.colstats(m1); s2 = b.colstats(m2); s1b = b.colstats(m1)
    assert np.array_equal(s1, s1b)
    assert s2["count"].sum() < s1["count"].sum()
    fb.modtable(ha, ha, 14)                      # same profiles as b holds now, on a context that never saw m1
    assert np.array_equal(fb.colstats(m2), s2)
    fb.close(); fresh.close(); b.close()


def test_device_encoder_equals_host_encoder(monkeypatch):
    """encode_pairs_kernel (used when a context has few encoder threads, or under JTK_DEVICE_ENCODE=1) against the host
    encoder: same cell count, same likelihoods and tables bit for bit (identical codes and guide bits), same rejection
    of ops that do not span the pair."""
    from jtk_b200 import _lib
    d = synth.diploid_chunk(33, length=700, n_reads=12)
    rng = np.random.default_rng(2)
    t2 = synth.random_template(rng, 90)
    q2, o2 = synth.mutate_read(rng, t2, 0.2)
    templates = [d["template"], t2]
    reads, ops = d["reads"] + [q2], d["ops"] + [o2]
    strands = np.concatenate([d["strands"], [1]]).astype(np.uint8)
    tidx = np.concatenate([np.zeros(len(d["reads"]), np.uint32), [1]]).astype(np.uint32)
    h = to_c(random_hmm(3))
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("JTK_DEVICE_ENCODE", mode)
        c = _lib.Context()
        b = c.batch(templates, reads, ops, strands, tidx, 30)
        b.modtable(h, h, 14)
        out[mode] = (b.cell_updates, b.lk(), b.profile(0), b.profile(len(reads) - 1))
        b.close()
        with pytest.raises(_lib.JtkError) as e:
            c.batch(templates, reads, [o[:-2] for o in ops], strands, tidx, 30)
        assert e.value.code == -1
        bad = [o.copy() for o in ops]
        bad[3][5] = 7
        with pytest.raises(_lib.JtkError):
            c.batch(templates, reads, bad, strands, tidx, 30)
        c.close()
    assert out["0"][0] == out["1"][0] > 0
    for k in (1, 2, 3):
        assert np.array_equal(out["0"][k], out["1"][k])


def test_fused_and_rows_variants_are_bit_identical(ctx, monkeypatch):
    """The fused kernel (forward rows recomputed in shared memory, nothing of the DP matrices in HBM) replays the forward pass
    of the rows variant instruction for instruction: likelihoods and every table entry are equal BIT FOR BIT -- on a ragged
    batch (lengths 1 .. 2 600, which also exercises the longest-first work queue), at radius 30 / 50 / 100 (2, 4, 8 column
    slots per lane) and for the 9-row table."""
    rng = np.random.default_rng(77)
    h = to_c(random_hmm(5))
    reads, ops, strands, tidx, templates = [], [], [], [], []
    for k, L in enumerate((1, 2, 3, 7, 15, 16, 17, 40, 300, 900, 2600)):
        t = synth.random_template(rng, L)
        templates.append(t)
        for r in range(3):
            q, o = synth.mutate_read(rng, t, 0.1)
            reads.append(q); ops.append(o); strands.append((k + r) % 2 == 0); tidx.append(k)
    tidx = np.asarray(tidx, np.uint32)
    strands = np.asarray(strands, np.uint8)
    for radius, rows in ((30, 14), (30, 9), (50, 14), (100, 14)):
        out = {}
        for v in ("fused", "rows"):
            monkeypatch.setenv("JTK_MODTABLE", v)
            b = ctx.batch(templates, reads, ops, strands, tidx, radius)
            b.modtable(h, h, rows)
            b.sync()
            assert ctx.last_modtable_variant == v
            out[v] = (b.lk().copy(), [b.profile(k).copy() for k in range(len(reads))])
            b.close()
        assert np.array_equal(out["fused"][0], out["rows"][0]), (radius, rows)
        for k in range(len(reads)):
            assert np.array_equal(out["fused"][1][k], out["rows"][1][k]), (radius, rows, k)


def test_variant_is_chosen_by_the_scratch_budget(ctx, monkeypatch):
    """Without JTK_MODTABLE the library parks forward rows in HBM only when the whole batch fits its scratch budget as one wave."""
    monkeypatch.delenv("JTK_MODTABLE", raising=False)
    d = synth.diploid_chunk(5, length=400, n_reads=6)
    h = to_c(O.default_hmm())
    b = ctx.batch([d["template"]], d["reads"], d["ops"], d["strands"], np.zeros(6, np.uint32), 30)
    b.modtable(h, h, 14)
    b.sync()
    assert ctx.last_modtable_variant == "rows"
    b.close()


def test_device_bootstrap_equals_host_bootstrap(ctx, monkeypatch):
    """The banded edit-distance guide of likelihood_antidiagonal_bootstrap computed on the device (bootstrap_ops_kernel, one
    thread per pair) is the host's alignment op for op: the likelihoods of the two paths are equal bit for bit,
    on calibration-shaped pairs (~100 bp, radius 10), on unequal lengths and on longer pairs."""
    rng = np.random.default_rng(123)
    fwd, rev = random_hmm(8), random_hmm(9)
    templates, reads, strands = [], [], []
    for k, L in enumerate([103] * 200 + [1, 2, 5, 17, 64, 65, 300, 600, 1500]):
        t = synth.random_template(rng, L)
        q, _ = synth.mutate_read(rng, t, 0.12 if k % 3 else 0.25)
        if k % 7 == 0 and L > 20:
            q = q[: len(q) - 9]          # a read that ends early: |Lr - Lt| widens the alignment band
        templates.append(t); reads.append(q); strands.append(k % 2 == 0)
    tidx = np.arange(len(reads), dtype=np.uint32)
    for R in (10, 30):
        out = {}
        for dev in ("0", "1"):
            monkeypatch.setenv("JTK_DEVICE_BOOTSTRAP", dev)
            n0 = ctx.launch_count
            out[dev] = ctx.likelihood_batch(to_c(fwd), to_c(rev), templates, reads, None, strands, tidx, R)
            # device path: bootstrap + encoder + likelihood kernels; host path: the likelihood kernel alone
            assert ctx.launch_count - n0 == (3 if dev == "1" else 1)
        assert np.array_equal(out["0"], out["1"]), R
        assert np.isfinite(out["1"]).all()


def test_packed_likelihood_kernel_on_calibration_pairs(ctx, monkeypatch):
    """at radius <= 14 two pairs share a warp (likelihood_pairs2_kernel).  1 201 calibration-shaped pairs
    (likelihood_gains.rs:253-315: ~100 bp, radius 10; ragged lengths, both strand models, an odd pair count) against
    O.likelihood_bootstrap pair by pair, against the guided oracle likelihood, and against the one-pair-per-warp kernel
    (JTK_LIKELIHOOD_UNPACKED=1; the two place their power-of-two rescales on different anti-diagonals, so ln(fin) - K ln 2 agrees to
    the last bits of the f64 logarithm, not bit for bit)."""
    rng = np.random.default_rng(2024)
    fwd, rev = random_hmm(21), random_hmm(22)
    lens = [100 + int(rng.integers(-8, 9)) for _ in range(1100)] + [int(rng.integers(1, 260)) for _ in range(101)]
    templates, reads, ops, strands = [], [], [], []
    for k, L in enumerate(lens):
        t = synth.random_template(rng, L)
        q, o = synth.mutate_read(rng, t, 0.08 if k % 4 else 0.2)
        templates.append(t); reads.append(q); ops.append(o); strands.append(k % 2 == 0)
    tidx = np.arange(len(reads), dtype=np.uint32)
    for R in (10, 14):
        got = {}
        for unpacked in (False, True):
            if unpacked: monkeypatch.setenv("JTK_LIKELIHOOD_UNPACKED", "1")
            else: monkeypatch.delenv("JTK_LIKELIHOOD_UNPACKED", raising=False)
            got[unpacked] = (ctx.likelihood_batch(to_c(fwd), to_c(rev), templates, reads, None, strands, tidx, R),
                             ctx.likelihood_batch(to_c(fwd), to_c(rev), templates, reads, ops, strands, tidx, R))
        assert np.allclose(got[False][0], got[True][0], rtol=1e-12, atol=0) and np.allclose(got[False][1], got[True][1], rtol=1e-12, atol=0), R
        for k in range(len(reads)):
            h = fwd if strands[k] else rev
            wantb = O.likelihood_bootstrap(h, templates[k], reads[k], R)
            want = O.likelihood(h, templates[k], reads[k], ops[k], R)
            assert abs(got[False][0][k] - wantb) <= 2e-5 * abs(wantb) + 1e-6, (R, k, got[False][0][k], wantb)
            assert abs(got[False][1][k] - want) <= 2e-5 * abs(want) + 1e-6, (R, k, got[False][1][k], want)
    monkeypatch.delenv("JTK_LIKELIHOOD_UNPACKED", raising=False)
