"""a few fused steps of config 4 at a reduced batch (1250 tiles, ~8 items per CTA): the command profiled by ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200")):
    sys.path.insert(0, p)
import torch
from models import RegionalTemporalGCN
from regt_b200 import workloads as W

B = int(os.environ.get("PB", "16"))
w = W.make_workload(os.environ.get("PW", "4"), B)
dev = torch.device("cuda:0")
m = RegionalTemporalGCN(8, w.N, w.T, w.O, hidden=w.H, n_regions=w.R, precision="tf32x3")
W.init_params_synthetic(m, 1234)
m = m.to(dev)
g = tuple(None if a is None else a.to(dev) for a in w.graph_args())
x, y = w.inputs(B)
x, y = x.to(dev), y.to(dev)
for i in range(int(os.environ.get("PN", "3"))):
    loss = m.fused_step(x, y, *g)[0]
torch.cuda.synchronize()
print("loss", float(loss))
