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
