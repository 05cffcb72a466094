"""Host side of local_clustering (csrc/local_clustering.cpp): the reference's own unit tests on these files restated,
generator vectors, and behaviour on oracle-made profiles (no GPU needed)."""
import math

import numpy as np
import pytest

import oracle_lib as O
from jtk_b200 import local_clustering as LC
from jtk_b200 import synth

GAINS = LC.Gains(gain=np.array([[4.0, 4.0, 4.0], [3.0, 2.0, 1.5], [3.0, 2.0, 1.5]]),
                 prob=np.array([[0.02, 0.02, 0.02], [0.05, 0.08, 0.1], [0.05, 0.08, 0.1]]))


def test_cosine_similarity_test():
    """pseudo_mcmc.rs:876-897"""
    assert abs(1 - LC.cosine_similarity([[1, 1], [2, 2]], 0, 1)) < 1e-4
    assert abs(-1 - LC.cosine_similarity([[1, -3], [1, -3]], 0, 1)) < 1e-5
    assert abs(LC.cosine_similarity([[1, 3], [1, -3]], 0, 1)) < 1e-5
    assert math.sqrt(0.5) - abs(LC.cosine_similarity([[0, 100], [1, 100]], 0, 1)) < 1e-5


def test_homop_length_test():
    """pseudo_mcmc.rs:899-904"""
    assert LC.homopolymer_length(b"ACCCCGTTTGGTT").tolist() == [1, 4, 4, 4, 4, 1, 3, 3, 3, 2, 2, 2, 2]


def test_xoshiro256starstar_reference_vector():
    """First outputs of the published xoshiro256** reference implementation for state {1,2,3,4} (the vector
    rand_xoshiro's own test uses)."""
    w = LC.rng_words(0, 4, state=[1, 2, 3, 4])
    assert w.tolist() == [11520, 0, 1509978240, 1215971899390074240]


def test_seed_from_u64_is_splitmix64():
    """seed_from_u64 fills the state with SplitMix64 outputs; check against an independent Python restatement."""
    def splitmix(x, n):
        out = []
        for _ in range(n):
            x = (x + 0x9e3779b97f4a7c15) & (2 ** 64 - 1)
            z = x
            z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & (2 ** 64 - 1)
            z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & (2 ** 64 - 1)
            out.append(z ^ (z >> 31))
        return out
    for seed in (0, 3490, 7 * 3490):
        a = LC.rng_words(seed, 6)
        b = LC.rng_words(0, 6, state=splitmix(seed, 4))
        assert a.tolist() == b.tolist()


def oracle_profiles(d, fwd, rev, radius):
    n = len(d["reads"])
    tabs, lks = O.modification_table_batch(fwd, rev, [d["template"]] * n, d["reads"], d["ops"], d["strands"], radius,
                                           n_threads=4)
    prof = np.stack(tabs) - lks[:, None]
    prof[np.stack(tabs) < -1e9] = -1e10
    return prof


@pytest.fixture(scope="module")
def chunk_profiles():
    d = synth.diploid_chunk(11, length=600, n_reads=40, error_rate=0.08, n_snv=4)
    h = O.default_hmm()
    return d, oracle_profiles(d, h, h, 18)


def test_diploid_chunk_is_phased(chunk_profiles):
    """P8: on benchmark_clustering-style data the adjusted Rand index is ~1 and the selected columns are
    the planted SNVs."""
    d, prof = chunk_profiles
    cfg = LC.ClusteringConfig.new(18, 2, 20.0, 20.0, GAINS)
    r = LC.clustering_on_profiles(prof, d["template"], d["strands"], cfg, seed=5 * 3490)
    assert r.k == 2
    agree = (r.assignments == d["hap"]).mean()
    assert max(agree, 1 - agree) >= 0.95
    snv = set(d["snv_pos"][0])
    assert len(r.probes) >= 2 and {int(p) // 14 for p in r.probes} <= snv
    assert all(int(p) % 14 < 4 for p in r.probes)  # substitutions
    # posteriors are normalised log-probabilities (the reference asserts this at mod.rs:185)
    assert np.allclose(np.exp(r.posterior).sum(axis=1), 1.0, atol=1e-9)
    # determinism: same seed, same stream
    r2 = LC.clustering_on_profiles(prof, d["template"], d["strands"], cfg, seed=5 * 3490)
    assert (r.assignments == r2.assignments).all() and r.score == r2.score


def test_single_haplotype_is_not_split():
    d = synth.diploid_chunk(12, length=500, n_reads=30, error_rate=0.08, n_snv=0, n_hap=1)
    h = O.default_hmm()
    prof = oracle_profiles(d, h, h, 15)
    cfg = LC.ClusteringConfig.new(15, 2, 15.0, 15.0, GAINS)
    r = LC.clustering_on_profiles(prof, d["template"], d["strands"], cfg, seed=1)
    assert r.k == 1 and (r.assignments == 0).all() and len(r.probes) == 0


def test_copy_num_one_is_trivial(chunk_profiles):
    d, prof = chunk_profiles
    cfg = LC.ClusteringConfig.new(18, 1, 20.0, 40.0, GAINS)
    r = LC.clustering_on_profiles(prof, d["template"], d["strands"], cfg, seed=3)
    assert r.k == 1 and r.score == 0.0 and (r.assignments == 0).all()


def test_host_restarts_twin_is_deterministic_and_splits_two_haplotypes():
    """jtk_lc_mcmc_restarts_host (the host twin the GPU restarts are compared with, tests/test_gpu_clustering.py): same
    generator state in -> same assignment, likelihood and state out; a clean two-haplotype matrix is split exactly."""
    import ctypes as C
    from jtk_b200 import _lib, pipeline as P
    L = _lib.lib()
    rng = np.random.default_rng(5)
    hap = rng.integers(0, 2, 40)
    v = np.ascontiguousarray(np.where(hap[:, None] == 1, 7.0, -7.0) + rng.normal(0, 1, (40, 4)))
    L.jtk_lc_mcmc_restarts_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_double)]
    outs = []
    for _ in range(2):
        st = P._rng_seed(99)
        asn = np.zeros(40, dtype=np.uint8)
        lk = C.c_double()
        assert L.jtk_lc_mcmc_restarts_host(_lib._ptr(v), 40, 4, 2, 20.0, 2, _lib._ptr(st), _lib._ptr(asn), C.byref(lk)) == 0
        outs.append((asn.copy(), lk.value, st.copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1] and np.array_equal(outs[0][2], outs[1][2])
    assert not np.array_equal(outs[0][2], P._rng_seed(99))
    a = outs[0][0]
    assert (a == hap).all() or (a == 1 - hap).all()


def test_speculative_schedule_equals_the_sequential_chain():
    """The schedule of mcmc_speculative_kernel restated on the host (jtk_lc_mcmc_restarts_spec_host: up to four proposals of a
    chain evaluated from the same state, earlier rejections replayed as their rounding round trips, commit up to the first
    acceptance, draws parsed from a look-ahead ring) against the sequential twin of mcmc_clustering's restart loop
    (pseudo_mcmc.rs:649-670,704-762): same assignments, likelihood and generator state, bit for bit -- on structured and
    structure-free matrices, tiny and 255-read chains, with the full 32-draw window and with windows so small that many
    rounds take the draw-by-draw path."""
    import ctypes as C
    from jtk_b200 import _lib, pipeline as P
    L = _lib.lib()
    vp = C.c_void_p
    L.jtk_lc_mcmc_restarts_host.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, vp, C.POINTER(C.c_double)]
    L.jtk_lc_mcmc_restarts_spec_host.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(C.c_double), vp]
    rng = np.random.default_rng(17)
    shapes = [(60, 6), (24, 1), (12, 5), (33, 2), (60, 6), (64, 7), (100, 8), (31, 2), (63, 4), (255, 8), (3, 1), (17, 3), (60, 6)]
    slow_rounds = 0
    for c, (n, D) in enumerate(shapes):
        hap = rng.integers(0, 2, n)
        sign = np.where(rng.random((2, D)) < 0.5, 1.0, -1.0)
        v = sign[hap] * rng.normal(6, 2, (n, D))
        v[rng.random((n, D)) < 0.15] = 0.0
        if c in (4, 11):
            v = rng.normal(0, 1.5, (n, D))  # no structure: many acceptances through exp()
        v = np.ascontiguousarray(v)
        st0 = P._rng_seed(1000 + 7 * c)
        st = st0.copy(); asn = np.zeros(n, dtype=np.uint8); lk = C.c_double()
        assert L.jtk_lc_mcmc_restarts_host(_lib._ptr(v), n, D, 2, n / 2, 2, _lib._ptr(st), _lib._ptr(asn), C.byref(lk)) == 0
        for window, spec in ((32, 4), (32, 8), (7, 4), (3, 8)):
            st2 = st0.copy(); asn2 = np.zeros(n, dtype=np.uint8); lk2 = C.c_double(); stats = np.zeros(3, dtype=np.uint64)
            assert L.jtk_lc_mcmc_restarts_spec_host(_lib._ptr(v), n, D, n / 2, 2, window, spec, _lib._ptr(st2), _lib._ptr(asn2),
                                                    C.byref(lk2), _lib._ptr(stats)) == 0
            assert np.array_equal(asn, asn2) and lk.value == lk2.value and np.array_equal(st, st2), (c, window, spec)
            assert stats[1] == 2 * 2000 * n and stats[0] <= stats[1]
            if window == 32:
                assert stats[2] == 0
                if c == 0 and spec == 4: assert stats[1] / stats[0] > 3.5      # a settled diploid chain commits almost four of four
                if c == 0 and spec == 8: assert stats[1] / stats[0] > 6.0      # ... and most of eight (they have to fit 32 draws)
            else:
                slow_rounds += int(stats[2])
    assert slow_rounds > 10000


def _brute_exact(v, k):
    """Independent restatement of exact_clustering.rs:7-77 in plain Python (tiny sizes only)."""
    n, d = v.shape
    sel = [0] * k
    choices = 1 << d
    score_of = lambda sel: sum(max(sum(x for i, x in enumerate(row) if (s >> i) & 1) for s in sel) for row in v)
    best, best_sel = 0.0, list(sel)
    while sel != [choices - 1] * k:
        sc = score_of(sel)
        if best < sc:
            best, best_sel = sc, list(sel)
        idx = 0
        while choices == sel[idx] + 1:
            idx += 1
        sel[idx] += 1
        for j in range(idx):
            sel[j] = sel[idx]
    return best, best_sel


def test_exact_clustering_matches_a_python_restatement():
    """exact_clustering::cluster_filtered_variants_exact (exact_clustering.rs:7-77) against a plain-Python loop."""
    rng = np.random.default_rng(4)
    for n, d, k in ((7, 3, 2), (9, 4, 2), (6, 3, 3), (5, 1, 2)):
        v = rng.normal(0, 3, (n, d))
        asn, gains, score, kk = LC.cluster_filtered_variants_exact(v, k)
        want, sel = _brute_exact(v, k)
        assert kk == k and abs(score - want) < 1e-12
        for r in range(n):
            g = [sum(x for i, x in enumerate(v[r]) if (s >> i) & 1) for s in sel]
            assert np.allclose(gains[r], g)
            assert gains[r, int(asn[r])] == max(g)


def test_exact_score_bounds_the_mcmc_score():
    """SURVEY 8c pin P7 (sandbox/src/bin/benchmark_mcmc.rs:112-118 prints both): the exhaustive search maximises the data
    term of the MCMC objective with free column subsets per cluster and no size prior (max_poisson_lk <= 0), so on the
    same variants its score can never be below what cluster_filtered_variants returns; on a clean two-haplotype matrix
    both find the planted split."""
    rng = np.random.default_rng(17)
    t = synth.random_template(rng, 300)
    for trial in range(6):
        n, d = 30 + 4 * trial, 2 + trial % 4
        hap = rng.integers(0, 2, n)
        signal = np.where(hap[:, None] == 1, 6.0, -6.0) * rng.choice([1.0, -1.0], d)[None, :]
        v = signal + rng.normal(0, 1.5, (n, d))
        v[np.abs(v) < 1.0] = 0.0  # compress_small_gains leaves exact zeros
        pos = (np.arange(d) * 23 + 40) * 14 + rng.integers(0, 4, d)
        cfg = LC.ClusteringConfig.new(15, 2, n / 2, n / 2, GAINS)
        r = LC.clustering_on_variants(v, pos.astype(np.uint32), t, cfg, seed=3490 * (trial + 1))
        asn, gains, score, _ = LC.cluster_filtered_variants_exact(v, 2)
        assert score >= r.score - 1e-9, (trial, score, r.score)
        assert r.k == 2
        a = r.assignments.astype(int)
        assert max((a == hap).mean(), (a != hap).mean()) >= 0.95
        e = asn.astype(int)
        assert max((e == hap).mean(), (e != hap).mean()) >= 0.95


def test_gen_reads_samples_the_pair_hmm():
    """jtk_lc_gen_reads (our sampler of kiley's Generate::gen): deterministic per (seed, read index) whatever the
    thread count, reads follow their source (length and identity as the model's error rates imply), and a model that never
    leaves Match with an identity emission matrix copies the source."""
    import os
    from jtk_b200 import _lib, hmm
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    srcs = [acgt[rng.integers(0, 4, size=L)] for L in (100, 101, 99, 37, 1)]
    idx = np.arange(3000, dtype=np.uint32) % len(srcs)
    model = hmm.PairHiddenMarkovModelOnStrands.default().forward().as_array()
    out, lens = _lib.gen_reads(model, srcs, idx, 77, 400)
    os.environ["JTK_CLUSTER_THREADS"] = "1"
    try:
        out1, lens1 = _lib.gen_reads(model, srcs, idx, 77, 400)
    finally:
        del os.environ["JTK_CLUSTER_THREADS"]
    assert np.array_equal(lens, lens1) and all(np.array_equal(out[k, :lens[k]], out1[k, :lens1[k]]) for k in range(0, 3000, 97))
    src_len = np.array([len(srcs[i]) for i in idx])
    assert abs(float(np.mean(lens / src_len)) - 1.0) < 0.05
    assert len({bytes(out[k, :lens[k]]) for k in range(0, 3000, 5)}) > 300          # not one read repeated
    same = [np.mean(out[k, :min(lens[k], 100)][:20] == srcs[idx[k]][:20]) for k in range(0, 3000, 5) if src_len[k] >= 99]
    assert np.mean(same) > 0.7                                                        # the first bases mostly agree
    exact = np.zeros(45)
    exact[[0, 3, 6]] = 1.0
    exact[9:25] = np.eye(4).ravel()
    exact[25:45] = 0.25
    o2, l2 = _lib.gen_reads(exact, srcs, idx[:50], 1, 400)
    assert all(np.array_equal(o2[k, :l2[k]], srcs[idx[k]]) for k in range(50))
