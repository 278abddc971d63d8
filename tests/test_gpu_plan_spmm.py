"""GPU parity of K1 (edge_index -> CSR + normalisation) and the F-wide SpMM, through the C-ABI."""
import numpy as np
import pytest
import torch

from parity_util import W, relerr

pytestmark = pytest.mark.gpu


def _graphs():
    gs = []
    for adv in (False, True):
        gs.append(W.tiny_workload("RegionalTemporalGCN", N=20, T=2, H=8, O=1, R=3, B=1, seed=11, adversarial=adv))
        gs.append(W.tiny_workload("TemporalGCN", N=17, T=2, H=8, O=1, R=0, B=1, seed=5, adversarial=adv))
    gs.append(W.make_workload(1))
    gs.append(W.make_workload(2))
    gs.append(W.make_workload(3))
    gs.append(W.make_workload(4))      # 10 k nodes, 32 regions, 61 k edges
    gs.append(W.make_workload(5))      # 100 k nodes, 256 regions, 610 k edges (the region-sharded config)
    return gs


@pytest.mark.parametrize("w", _graphs(), ids=lambda w: w.name)
def test_gcn_plan_bit_exact(w):
    from oracle import regt_oracle as O
    from regt_b200.plan import GraphPlanTensors, build_gcn
    dev = torch.device("cuda:0")
    ew = w.edge_attr
    plan = GraphPlanTensors(dev, w.N)
    build_gcn(plan, w.edge_index.to(dev), None if ew is None else ew.to(dev))
    rowptr, col, eid = O.canonical_gcn_csr(w.edge_index.numpy(), w.N)
    assert plan.nnz_gcn == len(col)
    assert np.array_equal(plan.t["g_rowptr"].cpu().numpy(), rowptr)
    assert np.array_equal(plan.t["g_col"].cpu().numpy(), col)
    assert np.array_equal(plan.t["g_eid"].cpu().numpy().astype(np.int64), eid)
    _, _, what = O.gcn_norm(w.edge_index, ew, w.N, torch.float64)
    val_ref = what[torch.from_numpy(eid)]
    assert relerr(plan.t["g_val"], val_ref) <= 1e-6
    # against the fp32 twin the sequential-order sums are reproduced to the last bit or two
    _, _, w32 = O.gcn_norm(w.edge_index, ew, w.N, torch.float32)
    assert relerr(plan.t["g_val"], w32[torch.from_numpy(eid)]) <= 3e-7


@pytest.mark.parametrize("w", _graphs(), ids=lambda w: w.name)
def test_cheb_plan_bit_exact(w):
    from oracle import regt_oracle as O
    from regt_b200.plan import GraphPlanTensors, build_cheb
    dev = torch.device("cuda:0")
    if w.model == "TemporalGCN":
        eis, eas = [w.edge_index], [w.edge_attr]
    else:
        eis, eas = w.reg_edge_index, w.reg_edge_attr
    plan = GraphPlanTensors(dev, w.N)
    build_cheb(plan, [e.to(dev) for e in eis], [None if a is None else a.to(dev) for a in eas])
    rp, col, reg, eid, segp = O.canonical_cheb_csr([e.numpy() for e in eis], w.N)
    assert plan.nnz_cheb == len(col) and plan.nseg == segp[-1] and plan.R == len(eis)
    assert np.array_equal(plan.t["c_rowptr"].cpu().numpy(), rp)
    assert np.array_equal(plan.t["c_col"].cpu().numpy()[:len(col)], col)
    assert np.array_equal(plan.t["c_reg"].cpu().numpy()[:len(col)], reg)
    assert np.array_equal(plan.t["c_eid"].cpu().numpy().astype(np.int64), eid)
    assert np.array_equal(plan.t["seg_ptr"].cpu().numpy(), segp)
    assert np.array_equal(plan.t["region_of"].cpu().numpy(), O.region_of_nodes([e.numpy() for e in eis], w.N))
    # segment tables are consistent with the CSR
    seg_eptr = plan.t["seg_eptr"].cpu().numpy()
    seg_reg = plan.t["seg_reg"].cpu().numpy()[:plan.nseg]
    seg_node = plan.t["seg_node"].cpu().numpy()[:plan.nseg]
    for s in range(plan.nseg):
        a, b = seg_eptr[s], seg_eptr[s + 1]
        assert b > a and (reg[a:b] == seg_reg[s]).all()
        assert rp[seg_node[s]] <= a and b <= rp[seg_node[s] + 1]
    rseg_ptr = plan.t["rseg_ptr"].cpu().numpy()
    rseg_list = plan.t["rseg_list"].cpu().numpy()[:plan.nseg]
    assert rseg_ptr[-1] == plan.nseg
    for r in range(plan.R):
        ids = rseg_list[rseg_ptr[r]:rseg_ptr[r + 1]]
        assert (seg_reg[ids] == r).all() and (np.diff(ids) > 0).all()
    # values: per-list Laplacian entries in canonical order
    vals = []
    for e, a in zip(eis, eas):
        vals.append(O.cheb_norm(e, a, w.N, torch.float64)[2])
    val_ref = torch.cat(vals)[torch.from_numpy(eid)] if len(col) else torch.zeros(0, dtype=torch.float64)
    if len(col):
        assert relerr(plan.t["c_val"][:len(col)], val_ref) <= 1e-6


@pytest.mark.parametrize("cfg,B", [(1, 3), (2, 4), (3, 2)])
def test_spmm_f8_matches_oracle_propagate(cfg, B):
    from oracle import regt_oracle as O
    from regt_b200.plan import GraphPlanTensors, build_gcn, spmm_f8
    w = W.make_workload(cfg)
    dev = torch.device("cuda:0")
    plan = GraphPlanTensors(dev, w.N)
    build_gcn(plan, w.edge_index.to(dev), None if w.edge_attr is None else w.edge_attr.to(dev))
    x, _ = w.inputs(B)
    y = spmm_f8(plan.t["g_rowptr"], plan.t["g_col"], plan.t["g_val"], x.to(dev))
    row, col, what = O.gcn_norm(w.edge_index, w.edge_attr, w.N, torch.float64)
    for b in range(B):
        ref = O.propagate(x[b].double().reshape(w.N, -1), row, col, what).reshape(x[b].shape)
        assert relerr(y[b], ref) <= 2e-6


def test_spmm_linearity_and_empty_rows_full_size():
    """size-independent properties at config-2 full size: linearity, and rows without in-edges
    (only the self-loop) return dis^2 * x."""
    from regt_b200.plan import GraphPlanTensors, build_gcn, spmm_f8
    w = W.make_workload(2)
    dev = torch.device("cuda:0")
    plan = GraphPlanTensors(dev, w.N)
    build_gcn(plan, w.edge_index.to(dev), w.edge_attr.to(dev))
    x1, _ = w.inputs(w.B)
    x2, _ = w.inputs(w.B, seed_offset=1)
    x1, x2 = x1.to(dev), x2.to(dev)
    args = (plan.t["g_rowptr"], plan.t["g_col"], plan.t["g_val"])
    lhs = spmm_f8(*args, (2.0 * x1 + x2).contiguous())
    rhs = 2.0 * spmm_f8(*args, x1) + spmm_f8(*args, x2)
    assert relerr(lhs, rhs) <= 1e-6


def _csr_ref(rowptr, col, val, x):
    """fp64 reference of the CSR product on the GPU (index_add over the edges)."""
    N = rowptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(N, device=x.device), (rowptr[1:] - rowptr[:-1]).long())
    xd = x.double().reshape(x.shape[0], N, -1)
    out = torch.zeros_like(xd)
    out.index_add_(1, rows, xd[:, col.long()] * val.double()[None, :, None])
    return out.reshape(x.shape)


@pytest.mark.parametrize("cfg,B", [(4, 3), (5, 2)])
def test_spmm_staged_partition_bit_identical_full_graphs(cfg, B):
    """the staged kernel on the 10 k / 100 k-node graphs: the plan's row partition (regt_spmm_partition) and the uniform
    blocks give bit-identical sums (same CSR order of additions), both within fp32 rounding of the fp64 product; the
    partition covers the rows, respects the kernel's block capacity and cuts the region-ordered graph at region borders."""
    from regt_b200 import _lib
    from regt_b200.plan import GraphPlanTensors, build_gcn, spmm_f8, spmm_partition
    w = W.make_workload(cfg)
    dev = torch.device("cuda:0")
    plan = GraphPlanTensors(dev, w.N)
    build_gcn(plan, w.edge_index.to(dev), None if w.edge_attr is None else w.edge_attr.to(dev))
    rp, cj, va = plan.t["g_rowptr"], plan.t["g_col"], plan.t["g_val"]
    x, _ = w.inputs(B)
    x = x.to(dev)
    y_part = spmm_f8(rp, cj, va, x, partition=True)
    y_unif = spmm_f8(rp, cj, va, x, partition=False)
    assert torch.equal(y_part, y_unif)
    assert relerr(y_part, _csr_ref(rp, cj, va, x)) <= 2e-6
    blk, nblk = spmm_partition(rp, cj, w.N, x[0, 0].numel())
    assert nblk > 0 and blk.numel() == nblk + 1
    b = blk.cpu().numpy()
    assert b[0] == 0 and b[-1] == w.N and (np.diff(b) > 0).all()
    cap = (227 * 1024 - 2048 - 3584 * 16 - 64) // (x[0, 0].numel() * 4 + 4) // 8 * 8     # spmm_geometry, one CTA per SM
    assert np.diff(b).max() <= cap
    # edges leaving their block: far fewer than with uniform blocks of the same count
    rows = torch.repeat_interleave(torch.arange(w.N, device=dev), (rp[1:] - rp[:-1]).long()).cpu().numpy()
    cols = cj.cpu().numpy()[: len(rows)]
    def crossing(bounds):
        return int((np.searchsorted(bounds, rows, side="right") != np.searchsorted(bounds, cols, side="right")).sum())
    uniform = np.minimum(np.arange(nblk + 1) * -(-w.N // nblk), w.N)
    assert crossing(b) <= crossing(uniform)
    assert crossing(b) <= 0.2 * len(rows)      # the synthetic graphs keep > 95 % of their edges inside a region


def test_spmm_staged_hub_row_and_tail_edges():
    """a hub row with more edges than one block stages (the tail of its edge list is read from global memory), rows
    without any edge, and a partition forced down to single-row blocks around the hub."""
    from regt_b200.plan import spmm_f8
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    N, Wd, B = 6000, 96, 2
    deg = torch.randint(0, 6, (N,), generator=g)
    deg[17] = 5000                       # hub
    deg[100:140] = 0                     # empty rows
    rowptr = torch.zeros(N + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(deg, 0).int()
    E = int(rowptr[-1])
    col = torch.randint(0, N, (E,), generator=g).int()
    val = torch.rand(E, generator=g) - 0.5
    x = torch.rand(B, N, Wd, generator=g)
    rowptr, col, val, x = rowptr.to(dev), col.to(dev), val.to(dev), x.to(dev)
    ref = _csr_ref(rowptr, col, val, x)
    for part in (True, False):
        y = spmm_f8(rowptr, col, val, x, partition=part)
        assert relerr(y, ref) <= 5e-6, part
        assert (y[:, 100:140] == 0).all()
    assert torch.equal(spmm_f8(rowptr, col, val, x, partition=True), spmm_f8(rowptr, col, val, x, partition=False))
