"""CPU tests of the region-sharded (multi-GPU) path's host logic: the bit-exact partition, halo
sets and local CSR slices, and -- with world_size-2 gloo groups -- the exchange step (flat gradient
all-reduce, all-gather + reassembly of regional outputs).  The compute inside the ranks is the CPU
oracle (test infrastructure); the CUDA kernels are covered by the -m gpu tests."""
import os

import numpy as np
import pytest
import torch

from oracle import regt_oracle as O
from regt_b200 import shard as S
from regt_b200 import workloads as W


def test_lpt_partition_matches_oracle_and_is_deterministic():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        for _ in range(5):
            sizes = rng.integers(1, 50, size=int(rng.integers(1, 40))).tolist()
            a, b = S.lpt_partition(sizes, world), O.lpt_partition(sizes, world)
            assert a.dtype == np.int32 and np.array_equal(a, b)
    assert S.lpt_partition([4, 4, 4, 4], 4).tolist() == [0, 1, 2, 3]
    assert S.lpt_partition([5, 9, 3, 9, 1, 4], 2).tolist() == [0, 0, 1, 1, 0, 1]   # 9->r0, 9->r1, 5->r0, 4->r1, 3->r1, 1->r0


@pytest.mark.parametrize("cfg", [1, 3])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_shards_partition_the_graph(cfg, world):
    w = W.make_workload(cfg)
    shards = [S.make_shard(w.N, w.edge_index, w.reg_edge_index, r, world) for r in range(world)]
    own_all = np.concatenate([s.own for s in shards])
    assert np.array_equal(np.sort(own_all), np.arange(w.N))            # every node owned exactly once
    ro = S.region_of_nodes(w.reg_edge_index, w.N)
    assert np.array_equal(ro, O.region_of_nodes([e.numpy() for e in w.reg_edge_index], w.N))
    src, dst = w.edge_index.numpy()
    for s in shards:
        assert np.array_equal(s.region_owner, shards[0].region_owner)   # same on every rank
        assert np.array_equal(s.counts, [len(t.own) for t in shards])
        for r in np.unique(ro[s.own]):                                   # whole regions
            if r >= 0:
                assert s.region_owner[r] == s.rank
        need = np.unique(src[np.isin(dst, s.own)])
        assert np.array_equal(np.sort(np.concatenate([np.intersect1d(need, s.own), s.halo])), need)
        assert not np.isin(s.halo, s.own).any()
    # LPT balance: the largest rank is within one largest-region of the mean
    sizes = np.bincount(ro[ro >= 0])
    assert max(len(s.own) for s in shards) <= w.N / world + sizes.max()


def test_random_decomposition_is_rejected():
    w = W.tiny_workload("RegionalTemporalGCN", N=23, T=4, H=8, O=3, R=3, B=1, seed=22, adversarial=True)
    with pytest.raises(ValueError, match="regional decomposition"):
        S.make_shard(w.N, w.edge_index, w.reg_edge_index, 0, 2)


@pytest.mark.parametrize("world", [2, 4])
def test_local_csr_slice_is_the_owned_rows_of_the_canonical_csr(world):
    w = W.make_workload(3)
    rowptr, col, eid = O.canonical_gcn_csr(w.edge_index.numpy(), w.N)
    _, _, what = O.gcn_norm(w.edge_index, None, w.N, torch.float64)
    val = what[torch.from_numpy(eid)]
    x = torch.rand(w.N, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    full = torch.zeros(w.N, 5, dtype=torch.float64)
    for n in range(w.N):
        for e in range(rowptr[n], rowptr[n + 1]):
            full[n] += val[e] * x[col[e]]
    for r in range(world):
        s = S.make_shard(w.N, w.edge_index, w.reg_edge_index, r, world)
        lut = S.local_lut(s, "cpu")
        rp, lc, lv = S.slice_csr(torch.from_numpy(rowptr), torch.from_numpy(col), val, torch.from_numpy(s.own), lut)
        assert rp.dtype == torch.int32 and lc.dtype == torch.int32
        # bit-exact structure: same row lengths, same entry order, columns map back to the global ids
        assert np.array_equal(np.diff(rp.numpy()), np.diff(rowptr)[s.own])
        perm = s.perm
        for i, n in enumerate(s.own):
            assert np.array_equal(perm[lc[rp[i]:rp[i + 1]].numpy()], col[rowptr[n]:rowptr[n + 1]])
        xl = x[torch.from_numpy(perm)]
        loc = torch.zeros(s.n_own, 5, dtype=torch.float64)
        for i in range(s.n_own):
            for e in range(int(rp[i]), int(rp[i + 1])):
                loc[i] += lv[e] * xl[int(lc[e])]
        assert torch.equal(loc, full[torch.from_numpy(s.own)])


# ---------------------------------------------------------------------------------------------
# world_size-2 (and 3) gloo: the exchange step with the oracle as the per-rank compute
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 3])
def test_gloo_exchange_equals_single_process(world, tmp_path):
    import subprocess
    import sys
    import shard_gloo_worker as G
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shard_gloo_worker.py")
    port = 29500 + os.getpid() % 2000 + world
    out_path = str(tmp_path / "rank0.pt")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), str(port), out_path]) for r in range(world)]
    try:
        for p in procs:
            assert p.wait(timeout=120) == 0
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    got = torch.load(out_path)
    w, m, x, y = G.make_case()
    out, hid, ref_loss = O.batched_step(m, x, y, w.graph_args())
    assert abs(got["loss"] - ref_loss) <= 1e-6 * abs(ref_loss)
    for g, p in zip(got["grads"], G.live_params(m)):
        ref = p.grad if p.grad is not None else torch.zeros_like(p)
        assert float((g.double() - ref).abs().max()) <= 1e-6 * max(1e-12, float(ref.abs().max()))
    both = got["both"]
    assert torch.allclose(both[..., :w.O], out, atol=1e-12) and torch.allclose(both[..., w.O:], hid, atol=1e-12)
