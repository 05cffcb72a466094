"""Host logic of the per-chunk driver (jtk_b200/pipeline.py, scheduler.py): restated reference unit tests and invariants.
No GPU: nothing here calls a compute entry point."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jtk_b200 import pipeline as P  # noqa: E402
from jtk_b200 import scheduler as S  # noqa: E402
from jtk_b200 import synth  # noqa: E402


def test_reorder_reference_vector():
    """haplotyper/src/local_clustering/normalize.rs:68-74 (reorder_test)."""
    arr, idx = [50, 40, 70, 60, 90], [3, 0, 4, 1, 2]
    P.reorder(arr, idx)
    assert arr == [40, 60, 90, 50, 70] and idx == [0, 1, 2, 3, 4]


def test_band_width_matches_reference_fractions():
    """definitions/src/lib.rs:173-175,201-210: ceil(frac * len)."""
    assert P.band_width("ONT", 2000) == 60 and P.band_width("CCS", 2000) == 20 and P.band_width("CLR", 2000) == 100
    assert P.band_width("ONT", 2001) == 61 and P.band_width("None", 10) == 1


def _toy_dataset(n_chunks=3, n_reads=(10, 12, 14)):
    chunks, nodes = [], []
    for c in range(n_chunks):
        d = synth.diploid_chunk(100 + c, length=60, n_reads=n_reads[c], error_rate=0.05, n_snv=1)
        chunks.append(P.Chunk(id=c, seq=d["template"], copy_num=2))
        for r, o, s in zip(d["reads"], d["ops"], d["strands"]):
            nodes.append(P.Node(chunk=c, seq=r, ops=o, is_forward=bool(s)))
    return P.DataSet(selected_chunks=chunks, nodes=nodes)


def test_update_coverage_is_half_the_median_node_count():
    ds = _toy_dataset()
    P.update_coverage(ds)
    assert ds.coverage == 6.0  # counts 10, 12, 14 -> median 12 -> haploid 6 (misc.rs:394-407)
    ds.coverage, ds.coverage_protected = 3.5, True
    P.update_coverage(ds)
    assert ds.coverage == 3.5


def test_pileup_nodes_sorts_by_non_match_columns():
    ds = _toy_dataset()
    pile = P.pileup_nodes(ds, {0, 2})
    assert set(pile) == {0, 2}
    for nodes, chunk in pile.values():
        keys = [P.nonmatch_columns(n, chunk) for n in nodes]
        assert keys == sorted(keys)
        for n in nodes:  # the key is what Node::recover shows: indel columns + substituted bases
            ops = n.ops
            i = j = bad = 0
            for op in ops:
                if op <= 1:
                    bad += int(n.seq[i] != chunk.seq[j]); i += 1; j += 1
                elif op == 2:
                    bad += 1; i += 1
                else:
                    bad += 1; j += 1
            assert bad == P.nonmatch_columns(n, chunk)


def test_estim_copy_num_sums_and_follows_coverage():
    """local_clustering/mod.rs:223-242."""
    asn = [0] * 30 + [1] * 10
    cps = P.estim_copy_num(asn, 2, 4, 10.0)
    assert sum(cps) == 4 and cps == [3, 1]
    assert P.estim_copy_num([0, 1, 2], 3, 3, 1.0) == [1, 1, 1]
    with pytest.raises(AssertionError):
        P.estim_copy_num([0], 3, 2, 1.0)


def test_normalize_relabels_by_descending_size():
    ds = _toy_dataset(1, (9,))
    ds.selected_chunks[0].cluster_num = 3
    labels = [2, 2, 2, 2, 0, 0, 1, 1, 1]
    for n, a in zip(ds.nodes, labels):
        n.cluster = a
        n.posterior = np.array([10.0 + a, 20.0 + a, 30.0 + a])  # column c = "posterior of old cluster c"
    P.normalize_local_clustering(ds)
    assert [n.cluster for n in ds.nodes] == [0, 0, 0, 0, 2, 2, 1, 1, 1]
    for n, a in zip(ds.nodes, labels):  # old cluster 2 -> 0, 1 -> 1, 0 -> 2: posterior columns move with the labels
        assert n.posterior.tolist() == [30.0 + a, 20.0 + a, 10.0 + a]


def test_partition_is_a_balanced_cover():
    rng = np.random.default_rng(5)
    w = rng.uniform(1, 10, size=37)
    for world in (1, 2, 4, 8):
        parts = S.partition_chunks(w, world)
        assert sorted(c for p in parts for c in p) == list(range(37))
        loads = [sum(w[c] for c in p) for p in parts]
        assert max(loads) - min(loads) <= max(w) + 1e-9  # LPT bound
    assert S.partition_chunks(w, 4) == S.partition_chunks(w, 4)


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = list(range(100, 123))
        weights = [1.0 + (i % 5) for i in ids]

        def process(my):  # stand-in for the GPU work of a rank: a deterministic function of the chunk id
            return {c: (c * 3490, np.arange(c % 7, dtype=np.uint64), f"rank{rank}") for c in my}

        merged = S.run_sharded(ids, weights, process, rank, world)
        # a caller that asks for ONE rank inside a process that has a group (bench.py's rank-0-only legs) must not enter a
        # collective the other ranks never join
        if rank == 0:
            solo = S.run_sharded(ids[:5], weights[:5], process, 0, 1)
            ret["solo"] = sorted(solo)
        if rank == 0:
            ret["ids"] = sorted(merged)
            ret["seeds"] = [merged[c][0] for c in sorted(merged)]
            ret["ranks"] = sorted({merged[c][2] for c in merged})
            ret["lens"] = [len(merged[c][1]) for c in sorted(merged)]
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_sharded_run_gathers_every_chunk_once_gloo_world2():
    """chunks partitioned over ranks, no data-path collective, one host gather on rank 0."""
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret["ids"] == list(range(100, 123))
        assert ret["solo"] == list(range(100, 105))
        assert ret["seeds"] == [c * 3490 for c in range(100, 123)]
        assert ret["ranks"] == ["rank0", "rank1"]
        assert ret["lens"] == [c % 7 for c in range(100, 123)]


def test_ops_travel_packed_two_bits_per_column():
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 5, 2047):
        o = rng.integers(0, 4, n).astype(np.uint8)
        assert np.array_equal(P._unpack_ops(P._pack_ops(o)), o)
        assert len(P._pack_ops(o)[1]) == (n + 3) // 4
    ops = [rng.integers(0, 4, n).astype(np.uint8) for n in (5, 0, 2047, 3)]
    back = P._unpack_ops_chunk(P._pack_ops_chunk(ops))
    assert len(back) == 4 and all(np.array_equal(a, b) for a, b in zip(ops, back))


def test_packed_sequences_behave_like_the_list_of_their_pieces():
    """_lib.Packed: the reads / guide ops of a call as ONE array + offsets (what the C ABI takes), indexable like the list."""
    from jtk_b200 import _lib
    rng = np.random.default_rng(3)
    xs = [rng.integers(0, 4, n).astype(np.uint8) for n in (3, 0, 5, 2, 0, 7)]
    pk = _lib.Packed(*_lib.concat(xs))
    assert len(pk) == len(xs) and [len(a) for a in pk] == [len(x) for x in xs]
    assert all(np.array_equal(a, x) for a, x in zip(pk, xs))
    assert np.array_equal(pk[-1], xs[-1]) and np.array_equal(pk[2], xs[2])
    assert [a.tolist() for a in pk[1:4]] == [x.tolist() for x in xs[1:4]]
    cat, off = _lib.concat(pk)            # passes through: no second concatenate
    assert cat is pk.cat and off is pk.off
    assert off.tolist() == np.concatenate([[0], np.cumsum([len(x) for x in xs])]).tolist()


def test_scatter_and_compact_runs_are_inverse():
    """jtk_scatter_runs spreads the compact ops of a call over the padded per-read slots of the polish staging buffer (the gaps
    stay untouched); jtk_compact_runs brings the patched runs back to the front."""
    import ctypes as C
    from jtk_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(4)
    runs = [rng.integers(0, 4, n).astype(np.uint8) for n in rng.integers(0, 300, 9000)]   # > 4096 runs: the threaded path
    cat, off = _lib.concat(runs)
    caps = np.array([len(r) + int(g) for r, g in zip(runs, rng.integers(0, 50, len(runs)))], dtype=np.uint64)
    pos = np.zeros(len(runs) + 1, dtype=np.uint64)
    np.cumsum(caps, out=pos[1:])
    buf = np.full(int(pos[-1]) + 7, 9, dtype=np.uint8)
    L.jtk_scatter_runs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    assert L.jtk_scatter_runs(_lib._ptr(cat), _lib._ptr(off), len(runs), _lib._ptr(buf), _lib._ptr(pos)) == 0
    for k in (0, 1, 17, 4095, 4096, len(runs) - 1):
        a = int(pos[k])
        assert np.array_equal(buf[a:a + len(runs[k])], runs[k])
        assert (buf[a + len(runs[k]):int(pos[k + 1])] == 9).all()      # the gap was not written
    lens = np.array([len(r) for r in runs], dtype=np.uint32)
    out_off = np.zeros(len(runs) + 1, dtype=np.uint64)
    L.jtk_compact_runs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    assert L.jtk_compact_runs(_lib._ptr(buf), _lib._ptr(pos), _lib._ptr(lens), len(runs), _lib._ptr(out_off)) == 0
    assert out_off.tolist() == off.astype(np.uint64).tolist()
    assert np.array_equal(buf[:int(out_off[-1])], cat)
    assert L.jtk_scatter_runs(None, None, 0, None, None) == 0


def test_gpu_mcmc_share_balances_the_device_and_the_host_threads(monkeypatch):
    """pipeline.gpu_mcmc_share: no GPU chains for calls the host threads absorb faster than one GPU chain runs; otherwise the
    host keeps what it finishes in the GPU's latency and the rest goes to the device (one wave of the speculative kernel, the
    sub-warp kernel beyond)."""
    monkeypatch.delenv("JTK_GPU_MCMC", raising=False)
    monkeypatch.setenv("JTK_CLUSTER_THREADS", "16")
    assert P.gpu_mcmc_share(20) == 0
    for n in (80, 250, 2000, 3750):
        g = P.gpu_mcmc_share(n)
        assert 0 < g <= n and g <= P.GPU_MCMC_CAPACITY
        host = n - g
        assert -(-host // 16) * P.GPU_MCMC_HOST_S <= max(P.gpu_mcmc_seconds(g), 1.4) + 1e-9   # the host side does not finish last
    assert P.gpu_mcmc_share(2000) <= P.GPU_MCMC_SPEC_CAPACITY          # configs[2]: one wave of the speculative kernel
    assert P.gpu_mcmc_seconds(100) < P.gpu_mcmc_seconds(2000) < P.gpu_mcmc_seconds(3000)
    monkeypatch.setenv("JTK_GPU_MCMC", "0"); assert P.gpu_mcmc_share(5000) == 0
    monkeypatch.setenv("JTK_GPU_MCMC", "1"); assert P.gpu_mcmc_share(5000) == 5000
