"""one rank of the world_size-N gloo test of the exchange step (tests/test_shard_cpu.py).
usage: python shard_gloo_worker.py RANK WORLD PORT OUT.pt -- the per-rank compute is the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

from oracle import regt_oracle as O
from regt_b200 import shard as S
from regt_b200 import workloads as W


def live_params(m):
    return [p for n, p in m.named_parameters() if not n.split(".")[-1].startswith(("_weight_att", "_bias_att"))]


def make_case():
    w = W.tiny_workload("RegionalTemporalGCN", N=30, T=3, H=8, O=2, R=4, B=2, seed=9)
    m = O.RegionalTemporalGCN(8, w.N, w.T, w.O, hidden=w.H, n_regions=w.R).double()
    W.init_params_synthetic(m, 5)
    x, y = w.inputs(2)
    return w, m, x.double(), y.double()


def main(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        w, m, x, y = make_case()
        sh = S.make_shard(w.N, w.edge_index, w.reg_edge_index, rank, world)
        live = live_params(m)
        ex = S.GradExchange(live, world)
        own = torch.from_numpy(sh.own)
        outs, hids = [], []
        for b in range(x.shape[0]):
            out, hid = m(x[b], *w.graph_args())
            # this rank's share of the snapshot's loss: owned rows, mean over the FULL graph (loss_nodes = N)
            loss = ((out[own] - y[b, own]) ** 2).sum() / (w.N * w.O)
            for p, g in zip(live, torch.autograd.grad(loss, live, allow_unused=True)):
                if g is not None:
                    p.grad += g
            ex.add_loss(loss.detach())
            outs.append(out[own].detach())
            hids.append(hid[own].detach())
        loss = ex.sync()
        both = S.all_gather_nodes(torch.cat([torch.stack(outs), torch.stack(hids)], dim=2), sh)
        if rank == 0:
            torch.save({"loss": float(loss), "grads": [p.grad.clone() for p in live], "both": both}, out_path)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
