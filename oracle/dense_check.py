"""Independent dense-matrix restatement of the hot path (numpy fp64) -- TEST INFRASTRUCTURE.

Written from the closed-form operators of SURVEY.md Appendix A/B, *not* from the scatter
code in regt_oracle.py, so that the two can be checked against each other:
  A_hat = D~^-1/2 (A_noloop + diag(loop_w)) D~^-1/2   with D~ = in-degree (column sums of
          the [src,dst] matrix) incl. the self-loop                      (gcn_norm)
  L_hat = - D^-1/2 A_noloop D^-1/2                    with D  = out-degree (Cheb, K=2)
Message passing is  out = M^T x  with M[src,dst].
Follows models/RegionalTemporalGCN.py:114-149, models/TemporalGCN.py:75-91,
models/utils.py:163-203 of the reference.
"""
import numpy as np


def _dense(ei, w, n):
    m = np.zeros((n, n))
    for s, d, v in zip(ei[0], ei[1], w):
        m[s, d] += v
    return m


def dense_gcn(ei, ew, n):
    ei = np.asarray(ei); ew = np.ones(ei.shape[1]) if ew is None else np.asarray(ew, dtype=np.float64)
    off = ei[0] != ei[1]
    a = _dense(ei[:, off], ew[off], n)
    loop = np.ones(n)
    for s, v in zip(ei[0][~off], ew[~off]):
        loop[s] = v
    a = a + np.diag(loop)
    deg = a.sum(axis=0)
    with np.errstate(divide="ignore"):
        dis = np.where(deg > 0, deg ** -0.5, 0.0)
    return dis[:, None] * a * dis[None, :]


def dense_cheb(ei, ew, n):
    ei = np.asarray(ei); ew = np.ones(ei.shape[1]) if ew is None else np.asarray(ew, dtype=np.float64)
    off = ei[0] != ei[1]
    a = _dense(ei[:, off], ew[off], n)
    deg = a.sum(axis=1)
    with np.errstate(divide="ignore"):
        dis = np.where(deg > 0, deg ** -0.5, 0.0)
    return -(dis[:, None] * a * dis[None, :])


def _sig(v):
    return 1.0 / (1.0 + np.exp(-v))


def forward(sd, x, ei, ew_full, reg_eis, reg_ews, regional):
    """sd: dict name -> float64 ndarray (reference state_dict keys); x [N,F,T].
    regional=True: RegionalTemporalGCN (full-graph TGCN gets unit weights);
    regional=False: TemporalGCN (reg_* ignored, ew_full goes to both operators)."""
    n, f, t_in = x.shape
    g = lambda k: np.asarray(sd[k], dtype=np.float64)
    if regional:
        ahat = dense_gcn(ei, None, n)
        lhats = [dense_cheb(e, w, n) for e, w in zip(reg_eis, reg_ews)]
    else:
        ahat = dense_gcn(ei, ew_full, n)
        lhats = [dense_cheb(ei, ew_full, n)]
    a = g("tgnn._attention"); p = np.exp(a - a.max()); p /= p.sum()
    w0, w1, cb = g("tgnn.conv.lins.0.weight"), g("tgnn.conv.lins.1.weight"), g("tgnn.conv.bias")
    acc = 0.0
    for t in range(t_in):
        xt = x[:, :, t]
        hs = [xt @ w0.T + (lh.T @ xt) @ w1.T + cb for lh in lhats]
        if regional:
            h = np.concatenate(hs, axis=1) @ g("tgnn.linear.weight").T + g("tgnn.linear.bias")
            h = np.where(h > 0, h, 0.01 * h)
        else:
            h = hs[0]
        s = ahat.T @ xt
        def conv(k):
            return s @ g(f"tgnn._base_tgcn.conv_{k}.lin.weight").T + g(f"tgnn._base_tgcn.conv_{k}.bias")
        def lin(k, v):
            return v @ g(f"tgnn._base_tgcn.linear_{k}.weight").T + g(f"tgnn._base_tgcn.linear_{k}.bias")
        z = _sig(lin("z", np.concatenate([conv("z"), h], axis=1)))
        r = _sig(lin("r", np.concatenate([conv("r"), h], axis=1)))
        ht = np.tanh(lin("h", np.concatenate([conv("h"), h * r], axis=1)))
        acc = acc + p[t] * (z * h + (1 - z) * ht)
    hid = acc
    o = np.maximum(hid, 0) @ g("linear1.weight").T + g("linear1.bias")
    o = np.maximum(o, 0) @ g("linear2.weight").T + g("linear2.bias")
    return o, hid
