"""probe: MN-major no-swizzle (interleaved 16-byte groups) tf32 / bf16 operands, selftest variants 6 and 7."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200")):
    sys.path.insert(0, p)
import torch
from regt_b200 import _lib
lib = _lib.load()
g = torch.Generator().manual_seed(1)
for fmt in (2, 1):
    for variant in (6, 7):
        for N, K in ((64, 128), (128, 64), (128, 32)):
            A = (torch.randint(-8, 9, (K, 128), generator=g).float() / 4.0)
            B = (torch.randint(-8, 9, (K, N), generator=g).float() / 4.0)
            ref = A.t() @ B
            D = torch.full((128, N), float("nan"), device="cuda")
            rc = lib.regt_debug_umma_selftest(fmt, variant, A.cuda().data_ptr(), B.cuda().data_ptr(), D.data_ptr(), N, K,
                                              torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            err = (D.cpu() - ref).abs().max().item()
            print(f"fmt {fmt} variant {variant} N={N} K={K}: rc={rc} max err {err}")
