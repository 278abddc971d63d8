import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "regt-gcn_b200"))
import torch
from regt_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for (M, K, N, splits) in ((5000, 512, 128, 7), (5000, 512, 128, 7), (5000, 512, 128, 1), (5000, 512, 128, 3), (5000, 512, 128, 1), (5000, 512, 128, 3), (5000, 512, 128, 1), (5000, 512, 128, 3), (20000, 1024, 256, 10), (20000, 1024, 256, 10), (20000, 1024, 256, 10)):
    g = torch.Generator().manual_seed(1)
    A = (torch.rand(M, K, generator=g) - 0.5).cuda()
    B = (torch.rand(M, N, generator=g) - 0.5).cuda()
    B2 = (torch.rand(M, 32, generator=g) - 0.5).cuda()
    Cp = torch.full((splits, K, N), float("nan"), device="cuda")
    Cp2 = torch.full((splits, K, 32), float("nan"), device="cuda")
    rc = lib.regt_debug_gemm_tn_tma(A.data_ptr(), K, B.data_ptr(), N, Cp.data_ptr(), M, K, N, splits, B2.data_ptr(), 32, Cp2.data_ptr(), st)
    torch.cuda.synchronize()
    chunk = ((M + splits - 1) // splits + 31) // 32 * 32
    print(M, K, N, splits)
    Ad, Bd = A.double(), B.double()
    for z in range(splits):
        r0, r1 = z * chunk, min(M, (z + 1) * chunk)
        d = Cp[z].double() - Ad[r0:r1].t() @ Bd[r0:r1]           # [K][N] error of this split
        badrows = (d.abs().max(1).values > 1e-3).nonzero().flatten().tolist()
        badcols = (d.abs().max(0).values > 1e-3).nonzero().flatten().tolist()
        if not badrows: continue
        print("  split", z, "bad rows", badrows[:24], "ncols bad", len(badcols))
        if len(badcols) < N:   # column-type corruption (B side)
            print("     B-side: bad cols", badcols[:24])
        # A-side hypothesis test on the first 4-row group
        grp = [r for r in badrows if r // 4 == badrows[0] // 4]
        nch = (r1 - r0 + 31) // 32
        def rows(c):
            lo, hi = r0 + c * 32, min(r1, r0 + c * 32 + 32)
            a = torch.zeros(32, K, dtype=torch.float64, device="cuda"); b = torch.zeros(32, N, dtype=torch.float64, device="cuda")
            if 0 <= c < nch: a[:hi - lo] = Ad[lo:hi]; b[:hi - lo] = Bd[lo:hi]
            return a, b
        best = []
        for c in range(nch):
            a, b = rows(c)
            for name, off in (("stale-2", -2), ("stale-1", -1), ("future+1", 1), ("future+2", 2), ("zero", None)):
                aw = torch.zeros_like(a) if off is None else rows(c + off)[0]
                e = (aw - a)[:, grp].t() @ b                           # predicted error if A rows of chunk c were replaced
                resid = float((d[grp] - e).abs().max())
                best.append((resid, c, name))
                # B stays right; also try the whole chunk (A and B) replaced for these rows: (aw^T bw - a^T b)
                if off is not None:
                    bw = rows(c + off)[1]
                    e2 = aw[:, grp].t() @ bw - a[:, grp].t() @ b
                    best.append((float((d[grp] - e2).abs().max()), c, name + "(A and B)"))
        best.sort()
        print("     rows", grp, "err", float(d[grp].abs().max()), "best hypotheses:", best[:3])
