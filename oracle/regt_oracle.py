"""CPU oracle for the RegT-GCN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``regt-gcn_b200/``) never imports it and has no CPU fallback.

PARITY UNPINNED.  The arithmetic of the reference lives in third-party packages
that are absent from ``/root/reference`` and not installable offline:
``torch_geometric`` (GCNConv, ChebConv; unpinned, ``README.md:31``) and
``torch_geometric_temporal`` (containers only).  The reference ships no tests and no
golden outputs for this path.  This file therefore *restates* the published PyG >= 2.0
semantics op for op, in the order the reference calls them, and
is pinned only by (i) the three shipped checkpoints loading strictly, (ii) the one
upstream known-answer vector we could recall (``get_laplacian`` 'sym', see
tests/test_oracle.py), and (iii) an independent dense-matrix restatement in
``oracle/dense_check.py`` that must agree with the scatter formulation below.

Everything is plain torch on CPU, dtype-generic (fp64 master, fp32 twin).

Reference call sites restated here (file:line under /root/reference):
  models/utils.py:163-203                TGCN cell (gates, blend)
  models/utils.py:107-156                GCNConv x3 + Linear(2H,H) x3 construction
  models/TemporalGCN.py:75-91, 21-32     A3TGCN loop over periods + head
  models/RegionalTemporalGCN.py:114-149  regional ChebConv branch, combine, TGCN, attention
  models/RegionalTemporalGCN.py:25-39    head (relu, linear1, relu, linear2)
  run.py:180,190                         loss = mean((out-y)^2); grads accumulate over snapshots
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# [3P] graph normalisations
# --------------------------------------------------------------------------------------
def gcn_norm(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int,
             dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """PyG ``gcn_norm(improved=False, add_self_loops=True)`` as used by GCNConv
    (reference ctor models/utils.py:107-155, call models/utils.py:169,175,181).

    Returns (row, col, w_hat) in PyG's post-normalisation edge order:
    non-loop edges in input order, then one self-loop per node 0..N-1.
    """
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop_w = torch.ones(num_nodes, dtype=dtype)
    if edge_weight is not None:
        ew = edge_weight.to(dtype)
        inv = ~mask
        # existing self-loops keep their weight; on duplicates the last one wins
        # (index_put semantics on CPU are sequential).
        for i, w in zip(row[inv].tolist(), ew[inv].tolist()):
            loop_w[i] = w
        w = torch.cat([ew[mask], loop_w])
    else:
        w = torch.ones(int(mask.sum()) + num_nodes, dtype=dtype)
    loops = torch.arange(num_nodes, dtype=row.dtype)
    row2 = torch.cat([row[mask], loops])
    col2 = torch.cat([col[mask], loops])
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, col2, w)  # target degree
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    w_hat = dis[row2] * w * dis[col2]
    return row2, col2, w_hat


def cheb_norm(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int,
              dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """PyG ``ChebConv.__norm__`` for normalization='sym', lambda_max=None (-> 2.0):
    ``get_laplacian`` (self-loops removed, degree scattered on the SOURCE index) followed
    by the Chebyshev rescale 2L/lambda_max - I.  Used at
    models/RegionalTemporalGCN.py:77-80,136-140 and models/TemporalGCN.py:65-69,88.

    Returns the off-diagonal entries only: (row, col, -dis[row]*w*dis[col]); the
    diagonal of the rescaled operator is exactly 0 and is dropped.
    """
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    row2, col2 = row[mask], col[mask]
    if edge_weight is None:
        w = torch.ones(row2.numel(), dtype=dtype)
    else:
        w = edge_weight.to(dtype)[mask]
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, row2, w)  # source degree
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    lap = -(dis[row2] * w * dis[col2])  # L = I - D^-1/2 A D^-1/2, off-diagonal part
    lam = 2.0
    lap = (2.0 * lap) / lam
    lap[lap == float("inf")] = 0
    return row2, col2, lap


def propagate(x: torch.Tensor, row: torch.Tensor, col: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """MessagePassing(aggr='add', flow='source_to_target'): out[col] += w * x[row]."""
    out = torch.zeros_like(x)
    out.index_add_(0, col, x[row] * w.unsqueeze(-1))
    return out


# --------------------------------------------------------------------------------------
# [3P] layers
# --------------------------------------------------------------------------------------
def _glorot_(t: torch.Tensor) -> torch.Tensor:
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)
    return t


class _Lin(nn.Module):
    """PyG ``Linear(bias=False)``: only a ``weight`` [out,in] (checkpoint key ``lin.weight``)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.weight = nn.Parameter(_glorot_(torch.empty(cout, cin)))

    def forward(self, x):
        return x @ self.weight.t()


class GCNConv(nn.Module):
    """Appendix A.2.  Parameters: ``lin.weight`` [H,F] (glorot), ``bias`` [H] (zeros)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(cout))
        self.lin = _Lin(cin, cout)

    def forward(self, x, edge_index, edge_weight=None):
        row, col, w = gcn_norm(edge_index, edge_weight, x.size(0), x.dtype)  # cached=False: every call
        xw = self.lin(x)
        return propagate(xw, row, col, w) + self.bias


class ChebConv(nn.Module):
    """Appendix A.3, K=2.  Parameters: ``lins.{0,1}.weight`` [H,F], ``bias`` [H]."""

    def __init__(self, cin: int, cout: int, K: int = 2):
        super().__init__()
        assert K == 2
        self.bias = nn.Parameter(torch.zeros(cout))
        self.lins = nn.ModuleList([_Lin(cin, cout) for _ in range(K)])

    def forward(self, x, edge_index, edge_weight=None):
        row, col, w = cheb_norm(edge_index, edge_weight, x.size(0), x.dtype)
        out = self.lins[0](x)
        tx1 = propagate(x, row, col, w)
        out = out + self.lins[1](tx1)
        return out + self.bias


# --------------------------------------------------------------------------------------
# reference modules restated
# --------------------------------------------------------------------------------------
class TGCN(nn.Module):
    """models/utils.py:69-203 with baseblock='gcn'."""

    def __init__(self, in_channels: int, out_channels: int, baseblock: str = "gcn",
                 improved: bool = False, cached: bool = False, add_self_loops: bool = True):
        super().__init__()
        if baseblock != "gcn":
            raise NotImplementedError("Current baseblock %s is not supported." % (baseblock))
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv_z = GCNConv(in_channels, out_channels)
        self.linear_z = nn.Linear(2 * out_channels, out_channels)
        self.conv_r = GCNConv(in_channels, out_channels)
        self.linear_r = nn.Linear(2 * out_channels, out_channels)
        self.conv_h = GCNConv(in_channels, out_channels)
        self.linear_h = nn.Linear(2 * out_channels, out_channels)

    def forward(self, X, edge_index, edge_weight=None, H=None):
        if H is None:  # models/utils.py:163-166
            H = torch.zeros(X.shape[0], self.out_channels, dtype=X.dtype)
        Z = torch.sigmoid(self.linear_z(torch.cat([self.conv_z(X, edge_index, edge_weight), H], dim=1)))
        R = torch.sigmoid(self.linear_r(torch.cat([self.conv_r(X, edge_index, edge_weight), H], dim=1)))
        Ht = torch.tanh(self.linear_h(torch.cat([self.conv_h(X, edge_index, edge_weight), H * R], dim=1)))
        return Z * H + (1 - Z) * Ht


class A3TGCN(nn.Module):
    """models/TemporalGCN.py:35-91 (the dead ``linear = Linear(64, H)`` is kept for key parity)."""

    def __init__(self, in_channels: int, out_channels: int, periods: int):
        super().__init__()
        self.periods = periods
        self._base_tgcn = TGCN(in_channels, out_channels)
        self.conv = ChebConv(in_channels, out_channels, K=2)
        self.linear = nn.Linear(64, out_channels)
        self._attention = nn.Parameter(torch.empty(periods).uniform_())

    def forward(self, X, edge_index, edge_weight=None):
        acc = 0
        probs = F.softmax(self._attention, dim=0)
        for t in range(self.periods):
            h = self.conv(X[:, :, t], edge_index, edge_weight)
            acc = acc + probs[t] * self._base_tgcn(X[:, :, t], edge_index, edge_weight, h)
        return acc


class RegionalA3TGCN(nn.Module):
    """models/RegionalTemporalGCN.py:42-149, generalised from 5 to R regional edge lists."""

    def __init__(self, in_channels: int, out_channels: int, num_nodes: int, periods: int, n_regions: int = 5):
        super().__init__()
        self.periods, self.n_regions = periods, n_regions
        self._base_tgcn = TGCN(in_channels, out_channels)
        self.conv = ChebConv(in_channels, out_channels, K=2)
        self.linear = nn.Linear(out_channels * n_regions, out_channels)
        self._attention = nn.Parameter(torch.empty(periods).uniform_())
        # dead parameters (never used in forward; kept for checkpoint compatibility)
        self._weight_att1 = nn.Parameter(torch.normal(0.0, 0.1, size=(out_channels, 1)))
        self._weight_att2 = nn.Parameter(torch.normal(0.0, 0.1, size=(num_nodes, 1)))
        self._bias_att1 = nn.Parameter(torch.normal(0.0, 1.0, size=(1, 1)))
        self._bias_att2 = nn.Parameter(torch.normal(0.0, 1.0, size=(1, 1)))

    def forward(self, X, edge_index, reg_edge_index: Sequence[torch.Tensor], reg_edge_weight: Sequence[torch.Tensor]):
        acc = 0
        probs = F.softmax(self._attention, dim=0)
        for t in range(self.periods):
            hs = [self.conv(X[:, :, t], ei, ew) for ei, ew in zip(reg_edge_index, reg_edge_weight)]
            h = F.leaky_relu(self.linear(torch.cat(hs, dim=1)))
            # the wrapper never passes edge_weight: the full-graph TGCN sees None (unit weights)
            acc = acc + probs[t] * self._base_tgcn(X[:, :, t], edge_index, None, h)
        return acc


class _Head(nn.Module):
    def _head(self, h):
        out_hidden = h
        h = self.linear2(torch.relu(self.linear1(torch.relu(h))))
        return h, out_hidden


class TemporalGCN(_Head):
    """models/TemporalGCN.py:7-32 (hidden generalised; reference value 256)."""

    def __init__(self, node_features: int, periods: int, output_dim: int, hidden: int = 256):
        super().__init__()
        self.tgnn = A3TGCN(node_features, hidden, periods)
        self.linear1 = nn.Linear(hidden, 128)
        self.linear2 = nn.Linear(128, output_dim)

    def forward(self, x, edge_index, edge_attr):
        return self._head(self.tgnn(x, edge_index, edge_attr))


class RegionalTemporalGCN(_Head):
    """models/RegionalTemporalGCN.py:9-39 (hidden / n_regions generalised; reference 256 / 5)."""

    def __init__(self, node_features: int, num_nodes: int, periods: int, output_dim: int,
                 hidden: int = 256, n_regions: int = 5):
        super().__init__()
        self.tgnn = RegionalA3TGCN(node_features, hidden, num_nodes, periods, n_regions)
        self.linear1 = nn.Linear(hidden, 128)
        self.linear2 = nn.Linear(128, output_dim)

    def forward(self, x, edge_index, *regional):
        """Positional form of the reference: R edge_index tensors then R edge_attr tensors.
        Also accepts two lists."""
        if len(regional) == 2 and isinstance(regional[0], (list, tuple)):
            eis, eas = regional
        else:
            R = len(regional) // 2
            eis, eas = regional[:R], regional[R:]
        return self._head(self.tgnn(x, edge_index, list(eis), list(eas)))


# --------------------------------------------------------------------------------------
# batched step = loop over snapshots, grads accumulate (run.py:170-195)
# --------------------------------------------------------------------------------------
def batched_step(model: nn.Module, x: torch.Tensor, y: Optional[torch.Tensor], graph_args: tuple,
                 backward: bool = True):
    """x [B,N,F,T], y [B,N,O].  Returns (out [B,N,O], out_hidden [B,N,H], loss_sum).
    loss_sum = sum_b mean((out_b - y_b)^2); parameter .grad fields hold the summed gradient."""
    outs, hids, total = [], [], 0.0
    for b in range(x.size(0)):
        out, hid = model(x[b], *graph_args)
        outs.append(out.detach())
        hids.append(hid.detach())
        if y is not None:
            loss = torch.mean((out - y[b]) ** 2)
            if backward:
                loss.backward()
            total += float(loss.detach())
    return torch.stack(outs), torch.stack(hids), total


# --------------------------------------------------------------------------------------
# canonical integer structures (bit-exact contract)
# --------------------------------------------------------------------------------------
def canonical_gcn_csr(edge_index: np.ndarray, num_nodes: int):
    """CSR by destination of PyG's post-gcn_norm edge list (non-loop edges in input order
    followed by N self-loops), stable-sorted by destination.  Duplicates kept.
    Returns rowptr int32 [N+1], col int32 [nnz] (sources), eid int64 [nnz] (position in the
    post-normalisation list; >= E' marks the appended self-loop of node eid-E')."""
    row, col = edge_index[0].astype(np.int64), edge_index[1].astype(np.int64)
    keep = row != col
    r2 = np.concatenate([row[keep], np.arange(num_nodes)])
    c2 = np.concatenate([col[keep], np.arange(num_nodes)])
    order = np.argsort(c2, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(c2, minlength=num_nodes), out=rowptr[1:])
    return rowptr, r2[order].astype(np.int32), order


def canonical_cheb_csr(edge_lists: Sequence[np.ndarray], num_nodes: int):
    """CSR by destination of the concatenation of R regional edge lists with self-loops
    removed, stable-sorted by destination (so inside a row, edges are grouped by region
    in list order).  Returns rowptr int32 [N+1], col int32 [nnz], reg int32 [nnz], eid int64
    [nnz] (index into the loop-free concatenation), seg_ptr int32 [N+1] (number of
    (node, region) segments before node n)."""
    rows, cols, regs = [], [], []
    for r, ei in enumerate(edge_lists):
        row, col = ei[0].astype(np.int64), ei[1].astype(np.int64)
        keep = row != col
        rows.append(row[keep]); cols.append(col[keep]); regs.append(np.full(int(keep.sum()), r, dtype=np.int32))
    row = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    col = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    reg = np.concatenate(regs) if regs else np.zeros(0, np.int32)
    order = np.argsort(col, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(col, minlength=num_nodes), out=rowptr[1:])
    reg_s = reg[order]
    seg_cnt = np.zeros(num_nodes, dtype=np.int32)
    for n in range(num_nodes):
        seg = reg_s[rowptr[n]:rowptr[n + 1]]
        seg_cnt[n] = 0 if seg.size == 0 else 1 + int(np.count_nonzero(seg[1:] != seg[:-1]))
    seg_ptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(seg_cnt, out=seg_ptr[1:])
    return rowptr, row[order].astype(np.int32), reg_s.astype(np.int32), order, seg_ptr


def region_of_nodes(edge_lists: Sequence[np.ndarray], num_nodes: int) -> np.ndarray:
    """node -> region derived from the regional edge lists (the reference has no such tensor:
    a region is only an edge list, load_dataset.py:318-356).  -1: node touched by no list;
    -2: node touched by more than one list ('random' decomposition)."""
    out = np.full(num_nodes, -1, dtype=np.int32)
    for r, ei in enumerate(edge_lists):
        touched = np.unique(np.concatenate([ei[0], ei[1]])).astype(np.int64)
        cur = out[touched]
        out[touched] = np.where(cur == -1, r, np.where(cur == r, r, -2))
    return out


def lpt_partition(region_sizes: Sequence[int], world: int) -> np.ndarray:
    """regions -> ranks by longest-processing-time bin packing: regions
    sorted by (-size, id); each goes to the least-loaded rank, ties -> lowest rank id."""
    order = sorted(range(len(region_sizes)), key=lambda r: (-int(region_sizes[r]), r))
    load = [0] * world
    owner = np.zeros(len(region_sizes), dtype=np.int32)
    for r in order:
        k = min(range(world), key=lambda i: (load[i], i))
        owner[r] = k
        load[k] += int(region_sizes[r])
    return owner
