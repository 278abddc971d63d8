"""SURVEY 8f.1 / 8f.2 on the device (csrc/loop.cu, regt_b200/loop.py) against the CPU restatement of run.py / predict.py /
load_dataset.py (oracle/loop_oracle.py)."""
import copy

import numpy as np
import pytest
import torch

from parity_util import W, build_cuda, build_oracle, is_dead, relerr, to_dev

pytestmark = pytest.mark.gpu


def _series(N, F, T_total, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(N, F, T_total, generator=g)          # MinMax-scaled features (load_dataset.py:430)


def test_window_gather_bit_exact():
    from oracle import loop_oracle as LO
    from regt_b200.loop import SlidingWindows
    nd = _series(37, 8, 61, 1)
    feats, targ = LO.windows(nd, 12, 6)
    sw = SlidingWindows(nd.cuda(), 12, 6)
    assert len(sw) == len(feats) == 61 - 18 + 1
    starts = torch.tensor([0, 5, 43, 17, 17], device="cuda")
    x, y = sw.gather(starts)
    for b, s in enumerate(starts.tolist()):
        assert torch.equal(x[b].cpu(), feats[s]) and torch.equal(y[b].cpu(), targ[s])


@pytest.mark.parametrize("skip_one", [False, True])
@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_flat_rmsprop_matches_torch(wd, skip_one):
    """incl. a parameter that never receives a gradient (torch skips it, weight decay and all: the reference's dead parameters)"""
    from regt_b200.loop import FlatRMSprop
    torch.manual_seed(3)
    ref = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5))
    mine = copy.deepcopy(ref).cuda()
    opt_ref = torch.optim.RMSprop(ref.parameters(), lr=1e-3, weight_decay=wd)
    skipped = [list(mine.parameters())[1]] if skip_one else []
    opt = FlatRMSprop(list(mine.parameters()), lr=1e-3, weight_decay=wd, skip=skipped)
    for it in range(4):
        g = torch.Generator().manual_seed(10 + it)
        for i, (pr, pm) in enumerate(zip(ref.parameters(), mine.parameters())):
            gr = torch.randn(pr.shape, generator=g)
            if skip_one and i == 1:
                pr.grad = None
                continue
            pr.grad = gr.clone()
            pm.grad.copy_(gr)
        opt_ref.step()
        opt.step()
    for pr, pm in zip(ref.parameters(), mine.parameters()):
        assert relerr(pm, pr) <= 1e-6


@pytest.mark.parametrize("n,zero_snapshot", [(624, False), (3900, True), (100_003, False)])
def test_eval_metrics_match_predict_py(n, zero_snapshot):
    """MAE / RMSE / p95-normalised MAPE incl. numpy's percentile interpolation and the 'leave out inf' rule."""
    from oracle import loop_oracle as LO
    from regt_b200 import _lib
    from regt_b200.loop import reduce_metrics
    lib = _lib.load()
    g = torch.Generator().manual_seed(n)
    S = 5
    y = torch.rand(S, n, generator=g)
    y[:, ::7] = 0.0                                   # ties and exact zeros
    if zero_snapshot:
        y[2] = 0.0                                    # p95 == 0 -> normalised errors are inf -> snapshot left out of the MAPE
    out = y + 0.1 * (torch.rand(S, n, generator=g) - 0.5)
    sums = torch.empty(S, 4, device="cuda", dtype=torch.float64)
    out_d, y_d = out.cuda(), y.cuda()
    rc = lib.regt_eval_metrics(out_d.data_ptr(), y_d.data_ptr(), S, n, 95.0, sums.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "regt_eval_metrics")
    for b in range(S):
        assert abs(float(sums[b, 2]) - float(np.percentile(y[b].numpy(), q=95))) <= 1e-6
    mae, rmse, mape = reduce_metrics(sums.cpu(), n)
    rmae, rrmse, rmape = LO.predict_metrics(list(out), list(y))
    assert abs(mae - rmae) <= 1e-5 * rmae and abs(rmse - rrmse) <= 1e-5 * rrmse and abs(mape - rmape) <= 1e-5 * abs(rmape)


@pytest.mark.parametrize("model_name", ["RegionalTemporalGCN", "TemporalGCN"])
def test_epoch_and_evaluate_match_run_py(model_name):
    """one epoch of run.py:163-199 (gradients of all snapshots accumulate, one RMSprop step) and predict.py's metrics,
    device-resident windows vs the oracle loop on the same series."""
    from oracle import loop_oracle as LO
    from regt_b200.loop import FlatRMSprop, SlidingWindows, evaluate, train_epoch
    w = W.tiny_workload(model_name, N=30, T=6, H=32, O=3, R=3 if model_name == "RegionalTemporalGCN" else 0, B=1, seed=11)
    nd = _series(w.N, 8, 6 + 3 + 8, 5)               # 9 windows
    ref = build_oracle(w, torch.float64)
    state = copy.deepcopy(ref.state_dict())
    feats, targ = LO.windows(nd.double(), 6, 3)
    opt_ref = torch.optim.RMSprop(ref.parameters(), lr=1e-3)
    # gradients of the epoch before the step (the step zeroes them)
    ref2 = build_oracle(w, torch.float64)
    for x, y in zip(feats, targ):
        o, _ = ref2(x, *w.graph_args())
        torch.mean((o - y) ** 2).backward()
    last_ref, total_ref = LO.train_epoch(ref, feats, targ, w.graph_args(), opt_ref)

    m = build_cuda(w, state, "cuda", precision="fp32")
    gargs = to_dev(w.graph_args(), "cuda")
    opt = FlatRMSprop([p for p in m.parameters()], lr=1e-3)
    sw = SlidingWindows(nd.cuda(), 6, 3)
    grads_before = {}
    last, total = train_epoch(m, sw, gargs, _Spy(opt, m, grads_before), batch=4)
    assert abs(float(last) - float(last_ref)) <= 1e-5 * abs(float(last_ref))
    assert abs(float(total) - float(total_ref)) <= 1e-5 * abs(float(total_ref))
    for k, p in ref2.named_parameters():
        if not is_dead(w.model, k) and p.grad is not None:
            assert relerr(grads_before[k], p.grad) <= 1e-5, k
    # the update itself: torch's RMSprop applied to OUR gradients (the first RMSprop step is +-10 lr per element wherever
    # |g| >> eps, so it is compared against the same gradients, not across implementations)
    chk = {k: torch.nn.Parameter(v.clone().float()) for k, v in state.items()}
    o2 = torch.optim.RMSprop(chk.values(), lr=1e-3)
    for k, p in chk.items():
        p.grad = grads_before[k].cpu() if k in grads_before else torch.zeros_like(p)
    o2.step()
    for k, p in m.named_parameters():
        assert relerr(p, chk[k]) <= 1e-6, k
    # evaluation on the updated model vs predict.py's arithmetic on the same outputs
    mae, rmse, mape = evaluate(m, sw, gargs, batch=4)
    with torch.no_grad():
        x, y = sw.gather(torch.arange(len(sw), device="cuda"))
        outs = m(x, *gargs)[0]
    rmae, rrmse, rmape = LO.predict_metrics(list(outs.cpu()), list(y.cpu()))
    assert abs(mae - rmae) <= 1e-5 * rmae and abs(rmse - rrmse) <= 1e-5 * rrmse and abs(mape - rmape) <= 1e-5 * abs(rmape)


class _Spy:
    """records the accumulated gradients right before the optimizer step"""

    def __init__(self, opt, model, store):
        self.opt, self.model, self.store = opt, model, store

    def zero_grad(self):
        self.opt.zero_grad()

    def step(self):
        for k, p in self.model.named_parameters():
            if p.grad is not None:
                self.store[k] = p.grad.detach().clone()
        self.opt.step()
