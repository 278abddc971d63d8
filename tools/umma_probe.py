"""which MN-major tf32 operand layouts does tcgen05 accept?  prints the error of every (variant, N, K) of the UMMA self-test."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "regt-gcn_b200")):
    sys.path.insert(0, p)
import torch
from regt_b200 import _lib
lib = _lib.load()
def ints(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(-8, 9, shape, generator=g).float() / 4.0)
for fmt in (2,):
    for variant in (2, 4, 5):
        for N, K in ((32, 8), (32, 32), (64, 128), (128, 64)):
            A = ints((K, 128), 1); B = ints((K, N), 2)
            ref = A.t() @ B
            D = torch.full((128, N), float("nan"), device="cuda")
            Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()
            rc = lib.regt_debug_umma_selftest(fmt, variant, Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            err = (D.cpu() - ref).abs().max().item()
            print(f"fmt {fmt} variant {variant} N={N} K={K}: rc={rc} max err {err}", flush=True)
