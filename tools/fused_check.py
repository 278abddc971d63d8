"""bring-up check of the fused 3xTF32 cell (csrc/cell_f.cu) against the fp64 oracle: prints every error; exit code 2 when an output or the
loss is off by more than 1e-4 (the GPU scripts stop there: a wrong kernel should not go on to be benchmarked)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from parity_util import W, build_cuda, is_dead, oracle_step, relerr, to_dev, twin_limits

WORST = [0.0]

cases = [
    ("RegionalTemporalGCN", dict(N=60, T=3, H=128, O=6, R=3, B=2, seed=8)),
    ("RegionalTemporalGCN", dict(N=60, T=12, H=128, O=6, R=3, B=2, seed=8, adversarial=True)),
    ("TemporalGCN", dict(N=150, T=12, H=64, O=12, R=0, B=3, seed=7, k_intra=5)),
    ("RegionalTemporalGCN", dict(N=700, T=4, H=128, O=4, R=5, B=30, seed=9)),      # 165 tiles > 148 SMs: two items on some CTAs
]
only = os.environ.get("CASE")
for ci, (model, kw) in enumerate(cases):
    if only is not None and int(only) != ci:
        continue
    w = W.tiny_workload(model, **kw)
    B = kw["B"]
    t0 = time.time()
    ref = oracle_step(w, B)
    m = build_cuda(w, ref["state"], "cuda:0", precision="tf32x3")
    x, y = w.inputs(B)
    fwd_only = os.environ.get("FWD_ONLY") == "1"
    if fwd_only:
        with torch.no_grad():
            out, hid = m(x.cuda(), *to_dev(w.graph_args(), "cuda:0"))
        torch.cuda.synchronize()
        print(f"[{ci}] {w.name} fwd-only: out {relerr(out, ref['out']):.2e} hid {relerr(hid, ref['hid']):.2e}", flush=True)
        continue
    loss, out, hid = m.fused_step(x.cuda(), y.cuda(), *to_dev(w.graph_args(), "cuda:0"))
    torch.cuda.synchronize()
    print(f"[{ci}] {w.name}: out {relerr(out, ref['out']):.2e} hid {relerr(hid, ref['hid']):.2e} "
          f"loss {abs(float(loss) - ref['loss']) / abs(ref['loss']):.2e}  ({time.time() - t0:.1f}s)", flush=True)
    WORST[0] = max(WORST[0], float(relerr(out, ref['out'])), float(relerr(hid, ref['hid'])), abs(float(loss) - ref['loss']) / abs(ref['loss']))
    _, twin = twin_limits(w, B, ref)
    for k, g in ref["grads"].items():
        if not is_dead(w.model, k):
            print(f"      grad {k:45s} {relerr(m.get_parameter(k).grad, g):.2e}   (fp32 twin of the oracle: {twin.get(k, float('nan')):.2e})", flush=True)
            if not k.endswith("_attention"):
                WORST[0] = max(WORST[0], float(relerr(m.get_parameter(k).grad, g)) / 10.0)

if WORST[0] > 1e-4:
    print(f"FUSED CHECK: worst error {WORST[0]:.2e} > 1e-4")
    sys.exit(2)
