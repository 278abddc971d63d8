"""TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this).

CPU restatement of the reference's callers on either side of the hot path, op for op:
  windows()          load_dataset.py:451-457   sliding windows of node_data [N, F, T_total]; the LAST feature is the target
  train_epoch()      run.py:163-199            per-snapshot forward / mean-squared loss / backward, ONE optimizer step
                                               (torch.optim.RMSprop, run.py:145); returns the last snapshot's loss (run.py:197)
  predict_metrics()  predict.py:142-194        MAE, RMSE, MAPE normalised by each snapshot's 95th percentile (numpy.percentile),
                                               snapshots whose normalised errors contain an inf are left out of the MAPE
Parity unpinned for the same reason as oracle/regt_oracle.py (the reference ships no tests or golden outputs and its model
code needs torch_geometric); the restatement follows the cited lines literally, including the numpy calls."""
import numpy as np
import torch


def windows(node_data: torch.Tensor, t_in: int, t_out: int):
    """load_dataset.py:446-457 -> (features [S][N,F,t_in], target [S][N,t_out])"""
    indices = [(i, i + (t_in + t_out)) for i in range(node_data.shape[2] - (t_in + t_out) + 1)]
    features, target = [], []
    for i, j in indices:
        features.append(node_data[:, :, i:i + t_in])
        target.append(node_data[:, -1, i + t_in:j])
    return features, target


def train_epoch(model, features, target, graph_args, optimizer):
    """run.py:163-199 (the RegionalTemporalGCN / TemporalGCN branches)"""
    model.train()
    total_loss = 0
    loss = None
    for x, y in zip(features, target):
        out, _ = model(x, *graph_args)
        loss = torch.mean((out - y) ** 2)
        loss.backward()
        total_loss += loss.detach()
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach(), total_loss


@torch.no_grad()
def predict_metrics(outs, ys):
    """predict.py:142-194 with the model outputs given (fp32 tensors, one per snapshot)"""
    mae, mse, mape = [], [], []
    for out, y in zip(outs, ys):
        # predict.py hands torch tensors to numpy (np.abs(tensor), np.percentile(tensor)); the explicit .numpy() is the same
        # arithmetic on the same fp32 array without numpy 2's __array_wrap__ deprecation noise
        err = (y - out).cpu().numpy()
        mae.append(torch.from_numpy(np.abs(err)))
        mse.append(((y - out) ** 2).cpu())
        with np.errstate(divide="ignore", invalid="ignore"):
            p95 = np.percentile(y.cpu().numpy(), q=95)
            if np.isinf(np.abs(err / p95)).any() == 0:
                mape.append(torch.from_numpy(np.abs(err / p95)))
    return (float(torch.cat(mae, dim=0).mean()), float(torch.cat(mse, dim=0).mean().sqrt()),
            float(torch.cat(mape, dim=0).mean()) * 100 if mape else float("nan"))
