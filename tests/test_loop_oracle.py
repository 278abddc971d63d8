"""CPU checks of the loop restatement (oracle/loop_oracle.py) and of the host-side metric reduction."""
import numpy as np
import torch

from parity_util import W  # noqa: F401  (sys.path)


def test_windows_follow_load_dataset():
    from oracle import loop_oracle as LO
    nd = torch.arange(2 * 3 * 10, dtype=torch.float32).reshape(2, 3, 10)
    f, t = LO.windows(nd, 4, 2)
    assert len(f) == 10 - 6 + 1
    assert torch.equal(f[3], nd[:, :, 3:7]) and torch.equal(t[3], nd[:, 2, 7:9])


def test_reduce_metrics_equals_predict_py_on_host_sums():
    """reduce_metrics over per-snapshot sums (what regt_eval_metrics returns) == predict.py's concatenate-and-mean."""
    from oracle import loop_oracle as LO
    from regt_b200.loop import reduce_metrics
    g = torch.Generator().manual_seed(0)
    S, n = 6, 500
    y = torch.rand(S, n, generator=g)
    y[1] = 0.0                                   # p95 == 0: left out of the MAPE
    out = y + 0.05 * torch.randn(S, n, generator=g)
    sums = torch.zeros(S, 4, dtype=torch.float64)
    for b in range(S):
        e = (y[b] - out[b])
        p = float(np.percentile(y[b].numpy(), q=95))
        sums[b, 0] = float(e.abs().double().sum())
        sums[b, 1] = float((e * e).double().sum())
        sums[b, 2] = p
        sums[b, 3] = sums[b, 0] / p if p != 0 else float("inf")
    mae, rmse, mape = reduce_metrics(sums, n)
    rmae, rrmse, rmape = LO.predict_metrics(list(out), list(y))
    assert abs(mae - rmae) <= 1e-5 * rmae and abs(rmse - rrmse) <= 1e-5 * rrmse and abs(mape - rmape) <= 1e-5 * rmape
