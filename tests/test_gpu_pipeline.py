"""GPU: the per-chunk driver (jtk_b200/pipeline.py) end to end on synthetic diploid chunks -- the shape of
BASELINE.json configs[0]/[1] at test size -- plus gain calibration and the recursive path for high copy numbers."""
import copy

import numpy as np
import pytest

from jtk_b200 import likelihood_gains as LG
from jtk_b200 import local_clustering as LC
from jtk_b200 import pipeline as P
from jtk_b200 import synth
from jtk_b200.hmm import PairHiddenMarkovModelOnStrands

pytestmark = pytest.mark.gpu

GAINS = LC.Gains(gain=np.array([[4.0, 4.0, 4.0], [3.0, 2.0, 1.5], [3.0, 2.0, 1.5]]),
                 prob=np.array([[0.02, 0.02, 0.02], [0.05, 0.08, 0.1], [0.05, 0.08, 0.1]]))


@pytest.fixture(scope="module")
def ctx():
    from jtk_b200 import _lib
    c = _lib.Context()
    yield c
    c.close()


def make_dataset(n_chunks, length, n_reads, seed0=700, n_snv=3, copy_num=2, draft_errors=3):
    chunks, nodes, truth = [], [], {}
    for c in range(n_chunks):
        d = synth.diploid_chunk(seed0 + c, length=length, n_reads=n_reads, error_rate=0.08, n_snv=n_snv)
        draft = d["template"].copy()
        chunks.append(P.Chunk(id=c + 1, seq=draft, copy_num=copy_num))
        order = np.random.default_rng(c).permutation(len(d["reads"]))
        for k in order:
            nodes.append(P.Node(chunk=c + 1, seq=d["reads"][k], ops=d["ops"][k], is_forward=bool(d["strands"][k])))
        truth[c + 1] = {id(nodes[-len(order) + i]): int(d["hap"][k]) for i, k in enumerate(order)}
    return P.DataSet(selected_chunks=chunks, nodes=nodes, read_type="ONT"), truth


def agreement(ds, truth, cid):
    lab = np.array([n.cluster for n in ds.nodes if n.chunk == cid])
    hap = np.array([truth[cid][id(n)] for n in ds.nodes if n.chunk == cid])
    a = (lab == hap).mean()
    return max(a, 1 - a)


def test_local_clustering_selected_phases_diploid_chunks(ctx):
    ds, truth = make_dataset(5, 700, 40)
    out = P.local_clustering_selected(ds, {c.id for c in ds.selected_chunks}, gains=GAINS, ctx=ctx, fit_models=False)
    assert set(out) == {1, 2, 3, 4, 5}
    assert ds.coverage == 20.0
    for c in ds.selected_chunks:
        assert c.cluster_num == 2 and c.score > 0
        assert agreement(ds, truth, c.id) >= 0.95
        nodes = [n for n in ds.nodes if n.chunk == c.id]
        sizes = np.bincount([n.cluster for n in nodes], minlength=2)
        assert sizes[0] >= sizes[1]  # normalize_local_clustering
        for n in nodes:
            assert len(n.posterior) == 2 and abs(np.exp(n.posterior).sum() - 1.0) < 1e-6
            assert np.count_nonzero(n.ops != 2) == len(c.seq) and np.count_nonzero(n.ops != 3) == len(n.seq)


def test_copy_number_zero_and_one_take_the_trivial_path(ctx):
    """Chunks with copy_num 0 or 1 occur after multiplicity estimation: the reference divides in f64 (n / 0 = inf,
    local_clustering/mod.rs:108-111) and `clustering` returns one cluster because copy_num < 2 (pseudo_mcmc.rs:86-88)."""
    ds, _ = make_dataset(3, 400, 20, seed0=1300)
    ds.selected_chunks[0].copy_num = 0
    ds.selected_chunks[1].copy_num = 1
    out = P.local_clustering_selected(ds, {1, 2, 3}, gains=GAINS, ctx=ctx, fit_models=False)
    assert set(out) == {1, 2, 3}
    for c in ds.selected_chunks[:2]:
        assert c.cluster_num == 1 and c.score == 0.0
        for n in (n for n in ds.nodes if n.chunk == c.id):
            assert n.cluster == 0 and list(n.posterior) == [0.0]
    assert ds.selected_chunks[2].cluster_num == 2


def test_sharded_driver_equals_single_rank(ctx):
    """Chunks are independent: clustering the two halves of a partition separately (what two ranks do) gives exactly the
    single-rank result."""
    ds, _ = make_dataset(4, 500, 30, seed0=800)
    P.update_coverage(ds)
    hmm = PairHiddenMarkovModelOnStrands.default()
    pile = P.pileup_nodes(copy.deepcopy(ds), {1, 2, 3, 4})
    whole = P._cluster_pileups(ctx, hmm, GAINS, ds.coverage, "ONT", pile)
    from jtk_b200 import scheduler
    parts = scheduler.partition_chunks([1.0] * 4, 2)
    ids = sorted(pile)
    merged = {}
    for part in parts:
        merged.update(P._cluster_pileups(ctx, hmm, GAINS, ds.coverage, "ONT", {ids[c]: pile[ids[c]] for c in part}))
    assert set(merged) == set(whole)
    for cid in whole:
        a, b = whole[cid], merged[cid]
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2] and np.array_equal(np.asarray(a[3]), np.asarray(b[3]))
        assert np.array_equal(np.array(a[4]), np.array(b[4]))


def test_gpu_mcmc_path_equals_host_path(ctx, monkeypatch):
    """with the k-means / MCMC restarts of every chunk on the GPU (jtk_lc_clustering_variants_batch) the
    driver returns exactly what it returns with the per-chunk host calls: consensus, score, cluster number, assignments
    and log-posteriors of every read."""
    ds, _ = make_dataset(4, 500, 30, seed0=900)
    ds.selected_chunks[3].copy_num = 3  # two levels of cluster_filtered_variants for one chunk
    P.update_coverage(ds)
    hmm = PairHiddenMarkovModelOnStrands.default()
    pile = P.pileup_nodes(copy.deepcopy(ds), {1, 2, 3, 4})
    monkeypatch.setenv("JTK_GPU_MCMC", "0")
    host = P._cluster_pileups(ctx, hmm, GAINS, ds.coverage, "ONT", pile)
    assert P.LAST_TIMING.get("gpu_mcmc_chunks", 0) == 0
    monkeypatch.setenv("JTK_GPU_MCMC", "1")
    dev = P._cluster_pileups(ctx, hmm, GAINS, ds.coverage, "ONT", pile)
    assert P.LAST_TIMING["gpu_mcmc_chunks"] >= 3
    for cid in host:
        a, b = host[cid], dev[cid]
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2], cid
        assert np.array_equal(np.array(a[3]), np.array(b[3])), cid
        assert np.array_equal(np.array(a[4]), np.array(b[4])), cid


def test_fit_loop_and_default_gain_calibration(ctx):
    ds, _ = make_dataset(6, 300, 24, seed0=900)
    models = P.estimate_model_parameters_on_both_strands(ds, ctx=ctx, rounds=2)
    f = models.forward().as_array()
    assert 0.85 < f[0] < 1.0 and abs(f[:3].sum() - 1.0) < 1e-6  # mat_mat learned from 8 % error reads, rows normalised
    gains = LG.estimate_gain(models, LG.SEED, LG.SEQ_LEN, LG.BAND, 2, ctx=ctx, sample_num=20, seq_num=20)
    assert gains.gain.shape == (3, 2) and (gains.gain > 0.5).all() and (gains.prob > 0).all() and (gains.prob <= 1).all()
    assert gains.gain[LG.SUBST, 0] > gains.gain[LG.DEL, 1]  # a substitution is easier to see than an indel in a homopolymer


def test_recursive_clustering_for_high_copy_number(ctx):
    """copy_num >= 8 takes clustering_recursive's split / re-polish / recurse path (local_clustering/mod.rs:136-189)."""
    rng = np.random.default_rng(3)
    base = synth.random_template(rng, 600)
    groups = []
    for g in range(2):  # two paralog groups, 12 substitutions apart, two haplotypes each (3 SNVs)
        t = base.copy()
        pos = rng.choice(np.arange(30, 570, 12), size=12, replace=False) if g else []
        for p in pos:
            t[p] = synth.ACGT[(np.searchsorted(synth.ACGT, t[p]) + 1) % 4]
        groups.append(t)
    nodes = []
    for g, t in enumerate(groups):
        for _ in range(32):
            r, o = synth.mutate_read(rng, t, 0.06)
            nodes.append(P.Node(chunk=1, seq=r, ops=o, is_forward=bool(rng.random() < 0.5)))
    ds = P.DataSet(selected_chunks=[P.Chunk(id=1, seq=base, copy_num=8)], nodes=nodes)
    out = P.local_clustering_selected(ds, {1}, gains=GAINS, ctx=ctx, fit_models=False)
    k = out[1][2]
    assert 2 <= k <= 8 and ds.selected_chunks[0].cluster_num == k
    lab = np.array([n.cluster for n in ds.nodes])
    grp = np.array([0] * 32 + [1] * 32)
    # no cluster mixes the two paralog groups
    for c in range(k):
        members = grp[lab == c]
        assert len(members) == 0 or members.min() == members.max()
    for n in ds.nodes:
        assert len(n.posterior) == k and abs(np.exp(n.posterior).sum() - 1.0) < 1e-4
