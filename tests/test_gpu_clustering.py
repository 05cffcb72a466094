"""GPU parity for the haplotyper-shaped path: identical selected variant columns and per-read cluster assignments
whether the profiles come from the f64 oracle or from the CUDA kernels (BASELINE.json north_star), and the
device-side column statistics / gather agree with a numpy restatement of pseudo_mcmc.rs:141-165,577-588,314-339."""
import numpy as np
import pytest

import oracle_lib as O
from jtk_b200 import local_clustering as LC
from jtk_b200 import synth

pytestmark = pytest.mark.gpu

GAINS = LC.Gains(gain=np.array([[4.0, 4.0, 4.0], [3.0, 2.0, 1.5], [3.0, 2.0, 1.5]]),
                 prob=np.array([[0.02, 0.02, 0.02], [0.05, 0.08, 0.1], [0.05, 0.08, 0.1]]))


@pytest.fixture(scope="module")
def ctx():
    from jtk_b200 import _lib
    c = _lib.Context()
    yield c
    c.close()


def to_c(h):
    from jtk_b200 import _lib
    return _lib.HmmParams.from_buffer_copy(bytes(h))


def oracle_profiles(d, fwd, rev, radius):
    n = len(d["reads"])
    tabs, lks = O.modification_table_batch(fwd, rev, [d["template"]] * n, d["reads"], d["ops"], d["strands"], radius,
                                           n_threads=4)
    prof = np.stack(tabs) - lks[:, None]
    prof[np.stack(tabs) < -1e9] = -1e10
    return prof


def gpu_profiles(ctx, d, fwd, rev, radius):
    n = len(d["reads"])
    lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), [d["template"]], d["reads"], d["ops"], d["strands"],
                                  np.zeros(n, np.uint32), radius)
    prof = np.stack(tabs) - lk[:, None]
    prof[np.stack(tabs) < -1e9] = -1e10
    return prof


@pytest.mark.parametrize("seed,length,n_reads,n_snv,radius", [(21, 600, 40, 4, 18), (22, 1000, 60, 5, 30), (23, 800, 50, 2, 24)])
def test_columns_and_assignments_identical_oracle_vs_gpu(ctx, seed, length, n_reads, n_snv, radius):
    d = synth.diploid_chunk(seed, length=length, n_reads=n_reads, error_rate=0.08, n_snv=n_snv)
    h = O.default_hmm()
    cfg = LC.ClusteringConfig.new(radius, 2, n_reads / 2, n_reads / 2, GAINS)
    ro = LC.clustering_on_profiles(oracle_profiles(d, h, h, radius), d["template"], d["strands"], cfg, seed * 3490)
    rg = LC.clustering_on_profiles(gpu_profiles(ctx, d, h, h, radius), d["template"], d["strands"], cfg, seed * 3490)
    assert ro.probes.tolist() == rg.probes.tolist()
    assert ro.k == rg.k and (ro.assignments == rg.assignments).all()
    assert abs(ro.score - rg.score) < 1e-3
    assert np.allclose(ro.posterior, rg.posterior, atol=1e-3)
    agree = (rg.assignments == d["hap"]).mean()
    assert max(agree, 1 - agree) >= 0.95


def numpy_colstats(prof, template, strands, gains):
    """compress_small_gains + column_sum + strand/sign counts, restated with numpy."""
    Lt = len(template)
    homop = LC.homopolymer_length(template).astype(np.int64)
    hl = np.concatenate([homop, [1]])
    H = gains.gain.shape[1]
    row = np.arange(14)
    typ = np.where(row < 4, 0, np.where(row < 11, 2, 1))  # Subst, Ins (incl. copy), Del
    thr = gains.gain[typ[None, :], np.minimum(hl, H)[:, None] - 1] * 0.5
    x = prof.reshape(len(prof), Lt + 1, 14).copy()
    x[np.abs(x) < thr[None]] = 0.0
    pos = x > 1e-5
    s = np.where(pos, x, 0).sum(axis=0)
    cnt = pos.sum(axis=0)
    big = np.abs(x) > 1e-4
    st = np.asarray(strands).astype(bool)[:, None, None]
    sc = np.stack([(big & ~st & (x < 0)).sum(0), (big & ~st & (x > 0)).sum(0), (big & st & (x < 0)).sum(0),
                   (big & st & (x > 0)).sum(0)], axis=-1)
    return x, s.reshape(-1), cnt.reshape(-1), sc.reshape(-1, 4)


def test_batch_path_matches_profile_path(ctx):
    """Level 2 (profiles stay in HBM, 9-row kernel, colstats + gather) gives the same result as level 1."""
    chunks = [synth.diploid_chunk(31 + c, length=700, n_reads=44, error_rate=0.08, n_snv=3) for c in range(3)]
    h = O.default_hmm()
    templates = [c["template"] for c in chunks]
    reads = [r for c in chunks for r in c["reads"]]
    ops = [o for c in chunks for o in c["ops"]]
    strands = np.concatenate([c["strands"] for c in chunks])
    tidx = np.repeat(np.arange(3, dtype=np.uint32), 44)
    b = ctx.batch(templates, reads, ops, strands, tidx, 21)
    b.modtable(to_c(h), to_c(h), 9)
    stats = b.colstats(GAINS.min_req, 1e-5)
    lk = b.lk()
    cfg = LC.ClusteringConfig.new(21, 2, 22.0, 22.0, GAINS)
    for t, d in enumerate(chunks):
        sl = slice(int(b.stat_off[t]), int(b.stat_off[t + 1]))
        prof = gpu_profiles(ctx, d, h, h, 21)
        # device statistics == numpy restatement on the 14-row profiles (the rows filter_profiles reads)
        x, s, cnt, sc = numpy_colstats(prof, d["template"], d["strands"], GAINS)
        rows = (np.arange(len(s)) % 14 < 8) | (np.arange(len(s)) % 14 == 11)
        assert (stats["count"][sl][rows] == cnt[rows]).all()
        assert np.allclose(stats["sum"][sl][rows], s[rows], atol=1e-3)
        assert (stats["sc"][sl][rows] == sc[rows]).all()
        # gather == compressed profile columns
        cols = np.flatnonzero(rows & (cnt > 0))[:50].astype(np.uint32)
        g = b.gather(t, GAINS.min_req, cols)
        assert np.allclose(g, x.reshape(len(prof), -1)[:, cols], atol=1e-4)
        rb = LC.clustering_on_batch(b, t, d["template"], stats[sl], cfg, (t + 1) * 3490)
        rp = LC.clustering_on_profiles(prof, d["template"], d["strands"], cfg, (t + 1) * 3490)
        assert rb.probes.tolist() == rp.probes.tolist()
        assert rb.k == rp.k and (rb.assignments == rp.assignments).all()
        # likelihoods of the two kernels (9-row and 14-row variants share the forward pass)
        want = ctx.likelihood_batch(to_c(h), to_c(h), [d["template"]], d["reads"], d["ops"], d["strands"],
                                    np.zeros(44, np.uint32), 21)
        assert np.allclose(lk[t * 44:(t + 1) * 44], want, rtol=1e-6)
    b.close()


def test_device_filter_profiles_matches_host(ctx):
    """jtk_batch_candidates / jtk_batch_search_variants (filter_profiles on the device, only candidates cross PCIe) select
    exactly the columns the host restatement selects from the full per-column statistics, and the clustering that
    follows is identical (pseudo_mcmc.rs:109-138,426-575)."""
    chunks = [synth.diploid_chunk(51 + c, length=600 + 50 * c, n_reads=40 + 2 * c, error_rate=0.08, n_snv=2 + c) for c in range(4)]
    h = O.default_hmm()
    templates = [c["template"] for c in chunks]
    reads = [r for c in chunks for r in c["reads"]]
    ops = [o for c in chunks for o in c["ops"]]
    strands = np.concatenate([c["strands"] for c in chunks])
    tidx = np.repeat(np.arange(4, dtype=np.uint32), [len(c["reads"]) for c in chunks])
    b = ctx.batch(templates, reads, ops, strands, tidx, 20)
    b.modtable(to_c(h), to_c(h), 9)
    stats = b.colstats(GAINS.min_req, 1e-5)
    copy_num = np.array([2, 2, 3, 2], dtype=np.int32)
    cov = 21.0
    cand = b.candidates(GAINS.gain, GAINS.prob, copy_num, cov)
    n_probes, probe_pos, variants = b.search_variants(GAINS.gain, GAINS.prob, copy_num, cov)
    assert len(cand) > 0 and (np.diff(cand["tmpl"].astype(np.int64)) >= 0).all()
    for t, d in enumerate(chunks):
        sl = slice(int(b.stat_off[t]), int(b.stat_off[t + 1]))
        cfg = LC.ClusteringConfig.new(20, int(copy_num[t]), cov, cov, GAINS)
        ref = LC.clustering_on_batch(b, t, d["template"], stats[sl], cfg, (t + 7) * 3490)
        ct = cand[cand["tmpl"] == t]
        # every candidate carries the statistics of its column, and the selected probes come out of the candidates
        assert (stats["count"][sl][ct["pos"]] == ct["count"]).all()
        assert (stats["sum"][sl][ct["pos"]] == ct["sum"]).all()
        k = int(n_probes[t])
        assert probe_pos[t, :k].tolist() == ref.probes.tolist()
        assert set(probe_pos[t, :k].tolist()) <= set(ct["pos"].tolist())
        rows = np.flatnonzero(tidx == t)
        got = LC.clustering_on_variants(variants[rows], probe_pos[t, :k], d["template"], cfg, (t + 7) * 3490)
        assert got.k == ref.k and (got.assignments == ref.assignments).all()
        assert got.score == ref.score and np.array_equal(got.posterior, ref.posterior)
        if k:
            g = b.gather(t, GAINS.min_req, probe_pos[t, :k])
            assert np.array_equal(g, variants[rows][:, :k])
    b.close()


def test_full_size_chunk_oracle_vs_gpu(ctx):
    """BASELINE.json configs[0] / one chunk of configs[1] at FULL size: 2 kbp, all 60 reads, radius 30, fitted-looking
    (non-default) strand models.  Every table entry against the oracle, then pseudo_mcmc::clustering on oracle tables and on
    GPU tables: identical probe columns, cluster number and assignments."""
    from test_gpu_parity import check_tables, random_hmm
    d = synth.diploid_chunk(7, length=2000, n_reads=60, error_rate=0.08, n_snv=5)
    fwd, rev = random_hmm(41), random_hmm(42)
    n = 60
    tidx = np.zeros(n, np.uint32)
    lk, tabs = ctx.modtable_batch(to_c(fwd), to_c(rev), [d["template"]], d["reads"], d["ops"], d["strands"], tidx, 30)
    otabs, olk = O.modification_table_batch(fwd, rev, [d["template"]] * n, d["reads"], d["ops"], d["strands"], 30, n_threads=8)
    check_tables(tabs, lk, otabs, olk, [d["template"]], tidx)   # asserts 2e-3 absolute / 1e-3 relative per entry
    # the clustering decision on the default models (random emission tables blur the SNV signal at 8 % error)
    h = O.default_hmm()
    lk, tabs = ctx.modtable_batch(to_c(h), to_c(h), [d["template"]], d["reads"], d["ops"], d["strands"], tidx, 30)
    otabs, olk = O.modification_table_batch(h, h, [d["template"]] * n, d["reads"], d["ops"], d["strands"], 30, n_threads=8)
    check_tables(tabs, lk, otabs, olk, [d["template"]], tidx)
    cfg = LC.ClusteringConfig.new(30, 2, 30.0, 30.0, GAINS)

    def prof(t, l):
        p = np.stack(t) - np.asarray(l)[:, None]
        p[np.stack(t) < -1e9] = -1e10
        return p
    ro = LC.clustering_on_profiles(prof(otabs, olk), d["template"], d["strands"], cfg, 7 * 3490)
    rg = LC.clustering_on_profiles(prof(tabs, lk), d["template"], d["strands"], cfg, 7 * 3490)
    assert ro.probes.tolist() == rg.probes.tolist() and len(rg.probes) >= 3
    assert ro.k == rg.k == 2 and (ro.assignments == rg.assignments).all()
    agree = (rg.assignments == d["hap"]).mean()
    assert max(agree, 1 - agree) >= 0.95


def test_repeat_heavy_chunk_first_level_oracle_vs_gpu(ctx):
    """BASELINE.json configs[3] shape: 4 paralogs x 2 haplotypes, 240 reads.  clustering_recursive's first level clusters into
    BRANCH_NUM = 4 (local_clustering/mod.rs:139-143): the same columns and assignments from oracle tables and from GPU tables,
    and the four clusters are the four paralogs."""
    d = synth.paralog_chunk(11, length=1000, n_reads=240)
    h = O.default_hmm()
    cfg = LC.ClusteringConfig.new(30, 4, 30.0, 60.0, GAINS)
    ro = LC.clustering_on_profiles(oracle_profiles(d, h, h, 30), d["template"], d["strands"], cfg, 11 * 3490)
    rg = LC.clustering_on_profiles(gpu_profiles(ctx, d, h, h, 30), d["template"], d["strands"], cfg, 11 * 3490)
    assert ro.probes.tolist() == rg.probes.tolist()
    assert ro.k == rg.k and (ro.assignments == rg.assignments).all()
    assert rg.k == 4
    lab, par = rg.assignments.astype(int), d["paralog"]
    iu = np.triu_indices(len(lab), 1)
    rand = ((lab[:, None] == lab[None, :])[iu] == (par[:, None] == par[None, :])[iu]).mean()
    assert rand >= 0.97, rand


def test_kiley_shaped_clustering_call(ctx):
    """local_clustering.clustering mirrors pseudo_mcmc::clustering's argument list."""
    from jtk_b200 import hmm
    d = synth.diploid_chunk(41, length=600, n_reads=40, error_rate=0.08, n_snv=4)
    cfg = LC.ClusteringConfig.new(18, 2, 20.0, 20.0, GAINS)
    r = LC.clustering(d["template"], d["reads"], d["ops"], d["strands"], 41 * 3490,
                      hmm.PairHiddenMarkovModelOnStrands.default(), cfg, ctx=ctx)
    assert r.k == 2
    agree = (r.assignments == d["hap"]).mean()
    assert max(agree, 1 - agree) >= 0.95


def _host_restarts(data, k, cov, state, restarts):
    import ctypes as C
    from jtk_b200 import _lib
    L = _lib.lib()
    d = np.ascontiguousarray(data, dtype=np.float64)
    st = np.array(state, dtype=np.uint64)
    asn = np.zeros(d.shape[0], dtype=np.uint8)
    lk = C.c_double()
    L.jtk_lc_mcmc_restarts_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_double)]
    rc = L.jtk_lc_mcmc_restarts_host(_lib._ptr(d), d.shape[0], d.shape[1], int(k), float(cov), restarts, _lib._ptr(st), _lib._ptr(asn),
                                     C.byref(lk))
    assert rc == 0
    return asn, lk.value, st


@pytest.mark.parametrize("kernel", ["speculative", "speculative-window-6", "subwarp"])
def test_device_mcmc_restarts_equal_the_host_twin(ctx, monkeypatch, kernel):
    """jtk_mcmc_restarts_batch (two clusters and <= 8 columns: mcmc_speculative_kernel, one chain per warp with
    four proposals side by side -- also with a 6-draw window, which sends many rounds through its draw-by-draw path -- or
    mcmc_diploid_kernel, four chains per warp; otherwise one warp per chain) against the host restatement of mcmc_clustering's
    restart loop on the same variants and generator states: assignments, likelihood and the generator state after the
    restarts must be identical, bit for bit (the generator state proves that every accept / reject went the same way)."""
    from jtk_b200 import pipeline as P
    rng = np.random.default_rng(17)
    datas, ks, covs, states = [], [], [], []
    for c, (n, D, k) in enumerate([(60, 6, 2), (24, 1, 2), (40, 3, 3), (12, 5, 2), (60, 4, 4), (33, 2, 2), (60, 6, 2), (18, 8, 3),
                                   (64, 7, 2), (100, 8, 2), (31, 2, 2), (63, 4, 2), (60, 9, 2)]):
        hap = rng.integers(0, k, n)
        sign = np.where(rng.random((k, D)) < 0.5, 1.0, -1.0)
        v = sign[hap] * rng.normal(6, 2, (n, D))
        v[rng.random((n, D)) < 0.15] = 0.0
        if c in (6, 11):
            v = rng.normal(0, 1.5, (n, D))  # no structure: the chain wanders, many acceptances through exp()
        datas.append(v); ks.append(k); covs.append(n / k); states.append(P._rng_seed(1000 + 7 * c))
    restarts = 3
    monkeypatch.setenv("JTK_MCMC_KERNEL", kernel.split("-")[0])
    if kernel.endswith("window-6"): monkeypatch.setenv("JTK_MCMC_WINDOW", "6")
    else: monkeypatch.delenv("JTK_MCMC_WINDOW", raising=False)
    asn, lk, err, st = ctx.mcmc_restarts(datas, ks, covs, np.array(states), restarts)
    assert (err == 0).all(), err
    for c in range(len(datas)):
        ha, hlk, hst = _host_restarts(datas[c], ks[c], covs[c], states[c], restarts)
        assert np.array_equal(asn[c], ha), c
        assert lk[c] == hlk, (c, lk[c], hlk)
        assert np.array_equal(st[c], hst), c


def test_device_pick_filtered_profiles_equals_the_host_twin(ctx, monkeypatch):
    """pick_filtered_profiles on the device (pick_probes_kernel, one warp per chunk: pseudo_mcmc.rs:516-575 with
    find_next_variants' last-maximum rule, the 7 bp mask, Sokal-Michener and cosine sweeps in the reference's summation order)
    against the host twin on the same candidates: identical probe counts, positions and variant columns -- diploid chunks at
    configs[1] size, a triploid one and two repeat-heavy copy_num 8 chunks (hundreds of candidates)."""
    chunks = [synth.diploid_chunk(700 + c, length=2000, n_reads=60, error_rate=0.08, n_snv=3 + 2 * c) for c in range(4)]
    chunks += [synth.paralog_chunk(4300 + c, length=1200, n_reads=160) for c in range(2)]
    h = O.default_hmm()
    templates = [c["template"] for c in chunks]
    reads = [r for c in chunks for r in c["reads"]]
    ops = [o for c in chunks for o in c["ops"]]
    strands = np.concatenate([c["strands"] for c in chunks])
    tidx = np.repeat(np.arange(len(chunks), dtype=np.uint32), [len(c["reads"]) for c in chunks])
    copy_num = np.array([2, 2, 3, 2, 8, 8], dtype=np.int32)
    b = ctx.batch(templates, reads, ops, strands, tidx, 30)
    b.modtable(to_c(h), to_c(h), 9)
    out = {}
    for host in (False, True):
        if host: monkeypatch.setenv("JTK_HOST_PICK", "1")
        else: monkeypatch.delenv("JTK_HOST_PICK", raising=False)
        n0 = ctx.launch_count
        out[host] = b.search_variants(GAINS.gain, GAINS.prob, copy_num, 25.0)
        assert ctx.launch_count - n0 == (2 if host else 3)     # candidates, gather (+ pick)
    monkeypatch.delenv("JTK_HOST_PICK", raising=False)
    assert out[False][0].sum() >= 6 and (out[False][0][4:] >= 2).all()
    for a, c in zip(out[False], out[True]):
        assert np.array_equal(a, c)
    b.close()
