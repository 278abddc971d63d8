"""prints the normwise relative error of every output / gradient vs the fp64 oracle, beside the fp32 twin's own error"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from parity_util import W, build_cuda, is_dead, oracle_step, relerr, to_dev

cases = {"cfg5s": (W.cfg5_shaped(), 2), "cfg4": (W.make_workload(4), 1), "cfg3": (W.make_workload(3), 1)}
for name in sys.argv[1:] or list(cases):
    w, B = cases[name]
    ref = oracle_step(w, B)
    twin = oracle_step(w, B, dtype=torch.float32)
    x, y = w.inputs(B)
    res = {}
    for prec in ("fp32", "tf32x3"):
        m = build_cuda(w, ref["state"], precision=prec)
        loss, out, hid = m.fused_step(x.cuda(), y.cuda(), *to_dev(w.graph_args(), "cuda"))
        res[prec] = dict(out=relerr(out, ref["out"]), hid=relerr(hid, ref["hid"]), loss=abs(float(loss) - ref["loss"]) / abs(ref["loss"]),
                         **{k: relerr(m.get_parameter(k).grad, g) for k, g in ref["grads"].items() if not is_dead(w.model, k)})
    tw = dict(out=relerr(twin["out"], ref["out"]), hid=relerr(twin["hid"], ref["hid"]), loss=abs(twin["loss"] - ref["loss"]) / abs(ref["loss"]),
              **{k: relerr(twin["grads"][k], g) for k, g in ref["grads"].items() if not is_dead(w.model, k)})
    print(f"== {name} ({w.name}, B={B})   {'key':45s} {'fp32':>10s} {'tf32x3':>10s} {'twin':>10s}")
    for k in tw:
        print(f"   {k:60s} {res['fp32'][k]:10.2e} {res['tf32x3'][k]:10.2e} {tw[k]:10.2e}")
