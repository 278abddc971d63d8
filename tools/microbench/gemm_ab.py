"""A/B of the register-fed (gemm_tc.cu) and TMA-fed (gemm_tma.cu) 3xTF32 GEMMs on the cell's shapes (CUDA events)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "regt-gcn_b200"))
import torch
from regt_b200 import _lib

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=5):
    flush = torch.empty(64 << 20, device="cuda", dtype=torch.float32)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for (M, H) in ((499200, 256), (3840000, 128), (158976, 64)):
    A = torch.rand(M, H, device="cuda") - 0.5
    D = torch.rand(M, 4 * H, device="cuda") - 0.5
    W = torch.rand(H, 2 * H, device="cuda") - 0.5
    C = torch.empty(M, H, device="cuda")
    scratch = torch.empty(4 * max(H, 128) * 2 * H, device="cuda")
    for K, name in ((H, "nt K=H "), (2 * H, "nt K=2H")):
        Aop, lda = (A, H) if K == H else (D, 4 * H)
        t0 = timeit(lambda: _lib.check(lib.regt_debug_gemm_nt(Aop.data_ptr(), lda, W.data_ptr(), 2 * H, C.data_ptr(), H, M, H, K, st), "nt"))
        t1 = timeit(lambda: _lib.check(lib.regt_debug_gemm_nt_tma(Aop.data_ptr(), lda, W.data_ptr(), 2 * H, C.data_ptr(), H, M, H, K, scratch.data_ptr(), st), "nt_tma"))
        gb = M * (K + H) * 4 / 1e9
        tf = 2 * 3 * M * H * K / 1e12
        print(f"M={M} H={H} {name}: legacy {t0:.3f} ms, tma {t1:.3f} ms  ({gb / t1 * 1e3:.0f} GB/s, {tf / t1 * 1e3:.0f} TF/s issued)")
    h = A
    hR = torch.rand(M, H, device="cuda") - 0.5
    F = torch.rand(M, 32, device="cuda")
    for splits in (37, 64) if H % 128 == 0 else (64,):
        C0 = torch.empty(splits, 2 * H, H, device="cuda"); C1 = torch.empty(splits, H, H, device="cuda"); C2 = torch.empty(splits, 4 * H, 32, device="cuda")
        def legacy():
            lib.regt_debug_gemm_tn2(D.data_ptr(), 4 * H, h.data_ptr(), H, C0.data_ptr(), M, 2 * H, H, splits, F.data_ptr(), 32, C2.data_ptr(), st)
            lib.regt_debug_gemm_tn2(D.data_ptr() + 8 * H, 4 * H, hR.data_ptr(), H, C1.data_ptr(), M, H, H, splits, F.data_ptr(), 32, C2.data_ptr(), st)
            lib.regt_debug_gemm_tn2(D.data_ptr() + 12 * H, 4 * H, None, 0, None, M, H, 0, splits, F.data_ptr(), 32, C2.data_ptr(), st)
        def tma3():
            lib.regt_debug_gemm_tn_tma(D.data_ptr(), 4 * H, h.data_ptr(), H, C0.data_ptr(), M, 2 * H, H, splits, F.data_ptr(), 32, C2.data_ptr(), st)
            lib.regt_debug_gemm_tn_tma(D.data_ptr() + 8 * H, 4 * H, hR.data_ptr(), H, C1.data_ptr(), M, H, H, splits, F.data_ptr(), 32, C2.data_ptr(), st)
            lib.regt_debug_gemm_tn_tma(D.data_ptr() + 12 * H, 4 * H, None, 0, None, M, H, 0, splits, F.data_ptr(), 32, C2.data_ptr(), st)
        t0, t1 = timeit(legacy), timeit(tma3)
        gb = M * (4 * H + 2 * H + 32) * 4 / 1e9
        msg = f"M={M} H={H} tn x3 splits={splits}: legacy {t0:.3f} ms, tma {t1:.3f} ms ({gb / t1 * 1e3:.0f} GB/s)"
        if H % 128 == 0:
            t2 = timeit(lambda: _lib.check(lib.regt_debug_gemm_tn_multi(D.data_ptr(), 4 * H, M, H, h.data_ptr(), hR.data_ptr(), C0.data_ptr(), C1.data_ptr(), splits, F.data_ptr(), C2.data_ptr(), st), "multi"))
            msg += f", merged {t2:.3f} ms ({gb / t2 * 1e3:.0f} GB/s)"
            ref = torch.zeros(2 * H, H, device="cuda", dtype=torch.float64)
            for r in range(0, M, 1 << 19):
                ref += D[r:r + (1 << 19), :2 * H].double().t() @ h[r:r + (1 << 19)].double()
            got = C0.double().sum(0)
            msg += f"; dB_zr normwise err vs fp64 {float((got - ref).abs().max() / ref.abs().max()):.2e}"
        print(msg)
    del A, D, C, hR, F
