"""GPU parity at the widths of the large configs: BASELINE configs[3] (cfg4: N = 10 000, R = 32,
H = 128) at full graph size with B = 1, and a config-5-shaped case (R = 256 regional lists, H = 128) at a node count
whose literal [N, R*H] regional concat the fp64 oracle can hold -- through fused_step in fp32 and tf32x3, and as the
sum over 8 region shards.  These are the regimes where the per-region weight-block tables (R*F*H), the dM1 chunk table
and the partial buffers change shape.  Tolerance: 1e-5 normwise relative vs the fp64 oracle (north_star) for activations and
loss; gradients follow SURVEY 8(c)'s rule: <= 1e-5, or within 4x of the fp32 twin's own error where the twin -- the reference's
own fp32 arithmetic -- is itself further than 1e-5 from the fp64 master (parity_util.twin_limits)."""
import functools

import pytest
import torch

from parity_util import W, twin_limits, build_cuda, is_dead, oracle_step, relerr, to_dev

pytestmark = pytest.mark.gpu
TOL = 1e-5


@functools.lru_cache(maxsize=None)
def _case(name):
    if name == "cfg4":
        w, B = W.make_workload(4), 1
    else:
        w, B = W.cfg5_shaped(), 2
    ref = oracle_step(w, B)                                   # fp64 master
    lims, twin_errs = twin_limits(w, B, ref)                  # the fp32 twin's own error bounds what fp32 can deliver
    return w, B, ref, (lims, twin_errs)


def _check(w, out, hid, loss, grads, ref, twin, tag):
    lims, twin_errs = twin
    assert relerr(hid, ref["hid"]) <= TOL, f"{tag} out_hidden {relerr(hid, ref['hid']):.3e}"
    assert relerr(out, ref["out"]) <= TOL, f"{tag} out {relerr(out, ref['out']):.3e}"
    assert abs(float(loss) - ref["loss"]) <= TOL * abs(ref["loss"]), f"{tag} loss"
    for k, g in ref["grads"].items():
        if is_dead(w.model, k):
            continue
        e = relerr(grads[k], g)
        print(f"{tag} grad {k}: {e:.3e} (fp32 twin: {twin_errs[k]:.3e}, limit {lims[k]:.3e})")
        assert e <= lims[k], f"{tag} grad {k}: {e:.3e} > {lims[k]:.3e} (fp32 twin {twin_errs[k]:.3e})"


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", ["cfg4", "cfg5_shaped"])
def test_large_config_fused_step_matches_oracle(name, precision):
    w, B, ref, twin = _case(name)
    m = build_cuda(w, ref["state"], precision=precision)
    x, y = w.inputs(B)
    loss, out, hid = m.fused_step(x.cuda(), y.cuda(), *to_dev(w.graph_args(), "cuda"))
    _check(w, out, hid, loss, {k: p.grad for k, p in m.named_parameters()}, ref, twin, f"{name}/{precision}")


@pytest.mark.parametrize("name,world", [("cfg5_shaped", 8), ("cfg4", 4)])
def test_large_config_sum_of_region_shards(name, world):
    """the R = 256 case as 8 region shards (32 owned lists + 224 EMPTY ones per rank) run one after the other on one GPU:
    outputs scatter back to global order, losses and gradients add up to the unsharded oracle."""
    from regt_b200 import shard as S
    w, B, ref, twin = _case(name)
    x, y = w.inputs(B)
    x, y = x.cuda(), y.cuda()
    ei, reis, reas = w.edge_index.cuda(), [e.cuda() for e in w.reg_edge_index], [a.cuda() for a in w.reg_edge_attr]
    loss_sum, grad_sum = 0.0, {}
    out_full = torch.zeros(B, w.N, w.O, device="cuda")
    hid_full = torch.zeros(B, w.N, w.H, device="cuda")
    for rank in range(world):
        m = build_cuda(w, ref["state"], precision="tf32x3")
        sm = S.RegionShardedModel(m, ei, reis, reas, rank, world)
        loss, out, hid = sm.fused_step(x, y, sync=False)
        loss_sum += float(loss)
        S.scatter_rows(out, sm.own, out_full)
        S.scatter_rows(hid, sm.own, hid_full)
        for k, p in m.named_parameters():
            if p.grad is not None:
                grad_sum[k] = p.grad.double().cpu() + grad_sum.get(k, 0.0)
    _check(w, out_full, hid_full, loss_sum, grad_sum, ref, twin, f"{name}/shards{world}")
