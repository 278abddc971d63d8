"""the stand-alone F-wide SpMM (regt_spmm_f8) on config 5's graph at B = PB snapshots: the command profiled by ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200")):
    sys.path.insert(0, p)
import torch
from regt_b200 import plan as P, workloads as W
B = int(os.environ.get("PB", "16"))
PART = os.environ.get("PPART", "1") == "1"
w = W.make_workload(os.environ.get("PW", "5"), B)
dev = torch.device("cuda:0")
x, _ = w.inputs(B)
x = x.to(dev)
ei = w.edge_index.to(dev)
plan = P.get_plan(w.N, dev, ei, None, [], [], need_cheb=False)
t = plan.t
for _ in range(3):
    y = P.spmm_f8(t["g_rowptr"], t["g_col"], t["g_val"], x, partition=PART)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    y = P.spmm_f8(t["g_rowptr"], t["g_col"], t["g_val"], x, partition=PART)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"spmm B={B} partition={PART}: {ms:.3f} ms, {W.spmm_bytes(w, B) / ms / 1e6:.0f} GB/s (L2 warm)")
