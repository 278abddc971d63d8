"""The generic 3xTF32 tcgen05 GEMMs (csrc/gemm_tc.cu) against fp64 matmul: fp32-equivalent accuracy
(1e-5 normwise would already fail a single-pass tf32 product, whose error is ~5e-4)."""
import pytest
import torch

from parity_util import relerr

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("M,N,K,pad", [(128, 128, 32, 0), (300, 256, 128, 0), (1000, 64, 256, 8), (77, 512, 64, 4), (129, 16, 96, 0),
                                       (4096, 128, 136, 0)])
def test_gemm_nt_tf32x3(M, N, K, pad):
    from regt_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.rand(M, K + pad, generator=g) - 0.5).cuda()
    Bt = (torch.rand(N, K + pad, generator=g) - 0.5).cuda()
    C = torch.full((M, N + pad), float("nan"), device="cuda")
    rc = lib.regt_debug_gemm_nt(A.data_ptr(), K + pad, Bt.data_ptr(), K + pad, C.data_ptr(), N + pad, M, N, K, _st())
    _lib.check(rc, "regt_debug_gemm_nt")
    ref = A[:, :K].double() @ Bt[:, :K].double().t()
    assert relerr(C[:, :N], ref) <= 1e-5
    if pad:
        assert torch.isnan(C[:, N:]).all()      # nothing written outside the N columns


@pytest.mark.parametrize("M,K,N,splits", [(32, 32, 16, 1), (256, 128, 128, 1), (1000, 256, 64, 4), (5000, 512, 128, 7), (70, 64, 32, 3),
                                          (4097, 128, 272, 5)])
def test_gemm_tn_tf32x3(M, K, N, splits):
    from regt_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.rand(M, K + 4, generator=g) - 0.5).cuda()
    B = (torch.rand(M, N + 8, generator=g) - 0.5).cuda()
    Cp = torch.full((splits, K, N), float("nan"), device="cuda")
    rc = lib.regt_debug_gemm_tn(A.data_ptr(), K + 4, B.data_ptr(), N + 8, Cp.data_ptr(), M, K, N, splits, _st())
    _lib.check(rc, "regt_debug_gemm_tn")
    ref = A[:, :K].double().t() @ B[:, :N].double()
    assert relerr(Cp.double().sum(0), ref) <= 1e-5


@pytest.mark.parametrize("M,K,N,splits", [(1000, 256, 128, 4), (3000, 128, 256, 3), (777, 64, 0, 2)])
def test_gemm_tn_with_second_operand(M, K, N, splits):
    """one pass over A contracts it with B [M,N] AND with a 32-column plane B2 (the F-wide weight gradients)."""
    from regt_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.rand(M, K + 4, generator=g) - 0.5).cuda()
    B = (torch.rand(M, max(N, 4), generator=g) - 0.5).cuda()
    B2 = (torch.rand(M, 32, generator=g) - 0.5).cuda()
    Cp = torch.full((splits, K, max(N, 1)), float("nan"), device="cuda")
    Cp2 = torch.full((splits, K, 32), float("nan"), device="cuda")
    rc = lib.regt_debug_gemm_tn2(A.data_ptr(), K + 4, B.data_ptr() if N else None, max(N, 4), Cp.data_ptr() if N else None, M, K, N,
                                 splits, B2.data_ptr(), 32, Cp2.data_ptr(), _st())
    _lib.check(rc, "regt_debug_gemm_tn2")
    if N:
        assert relerr(Cp.double().sum(0), A[:, :K].double().t() @ B[:, :N].double()) <= 1e-5
    assert relerr(Cp2.double().sum(0), A[:, :K].double().t() @ B2.double()) <= 1e-5
