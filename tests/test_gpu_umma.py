"""tcgen05 building blocks: the hand-written UMMA operand layouts and descriptors (csrc/tc_common.cuh)
reproduce an exact small-integer GEMM in every role the fused kernels use them in."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ints(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(-8, 9, shape, generator=g).float() / 4.0)   # exact in bf16 and tf32


@pytest.mark.parametrize("fmt", [1, 2], ids=["bf16", "tf32"])
@pytest.mark.parametrize("variant,N,K", [(0, 128, 64), (0, 64, 128), (1, 64, 0), (2, 64, 128), (2, 128, 64), (3, 32, 128)])
def test_umma_layout_roles(fmt, variant, N, K):
    from regt_b200 import _lib
    lib = _lib.load()
    uk = 8 if fmt == 2 else 16
    if fmt == 2 and variant >= 2:
        pytest.skip("MN-major tiles are only used with bf16 operands; the 3xTF32 row contraction (gemm_tn) transposes "
                    "into K-major tiles instead (a first guess at the 32-byte-atom swizzle did not reproduce the GEMM)")
    if variant == 1:
        K = uk
    if variant >= 2:
        A = _ints((K, 128), 1)        # D = A^T B, contraction over the tile rows
        B = _ints((K, N), 2)
        ref = A.t() @ B
    else:
        A = _ints((128, K), 1)
        B = _ints((N, K), 2)
        ref = A @ B.t()
    Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.regt_debug_umma_selftest(fmt, variant, Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K,
                                      torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "regt_debug_umma_selftest")
    torch.cuda.synchronize()
    assert torch.equal(D.cpu(), ref), f"max err {(D.cpu() - ref).abs().max()}"
