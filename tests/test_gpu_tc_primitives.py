"""GPU tests of the tcgen05 building blocks (descriptors, swizzle, TMEM) against torch fp32 matmul."""
import pytest
import torch

from nvp_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [64, 128, 256])
def test_umma_kmajor_tile(K):
    g = torch.Generator().manual_seed(K)
    A = (torch.randn(128, K, generator=g)).half().cuda()
    B = (torch.randn(128, K, generator=g)).half().cuda()
    D = torch.full((128, 128), float("nan"), device="cuda")
    _lib.check(_lib.load().nvp_selftest_umma(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, 0,
                                             torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert float((D - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


@pytest.mark.parametrize("K", [16, 64, 128])
def test_umma_mnmajor_tile(K):
    g = torch.Generator().manual_seed(K + 1)
    A = (torch.randn(K, 128, generator=g)).half().cuda()
    B = (torch.randn(K, 128, generator=g)).half().cuda()
    D = torch.full((128, 128), float("nan"), device="cuda")
    _lib.check(_lib.load().nvp_selftest_umma(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, 1,
                                             torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    ref = A.float().t() @ B.float()
    assert float((D - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


@pytest.mark.parametrize("x", [0, 1, 2, 3])
def test_umma_skinny_column_block(x):
    """N=16 MMA whose B operand starts 16*x columns into the swizzle atom (small-gradient reductions)."""
    K = 128
    g = torch.Generator().manual_seed(40 + x)
    A = (torch.randn(K, 128, generator=g)).half().cuda()
    B = (torch.randn(K, 128, generator=g)).half().cuda()
    D = torch.full((128, 128), float("nan"), device="cuda")
    _lib.check(_lib.load().nvp_selftest_umma(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, 2 + x,
                                             torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    ref = A.float().t() @ B.float()[:, 16 * x:16 * x + 16]
    assert float((D[:, :16] - ref).abs().max()) <= 1e-3 * float(ref.abs().max())
