"""fp32 GEMM of the autocast-off islands (Linear, src/tinyedm/networks.py:46-64, :164, :255, :319) against torch fp32
(TF32 off): every transpose combination, ragged sizes, split-K shapes and the beta path."""
import pytest
import torch

from tests.helpers import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (they fail, not skip, without one)"
    return torch.device("cuda:0")


@pytest.mark.parametrize("M,N,K", [(256, 5376, 256), (5376, 256, 256), (256, 256, 5376), (128, 65, 257), (1, 64, 300),
                                   (7, 3, 5), (64, 1, 128)])
@pytest.mark.parametrize("transA", [False, True])
@pytest.mark.parametrize("transB", [False, True])
def test_sgemm_matches_torch(dev, M, N, K, transA, transB):
    from tinyedm_b200 import ops
    ops.ensure_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(M + 3 * N + 7 * K)
    A = torch.randn((K, M) if transA else (M, K), device=dev)
    Bm = torch.randn((N, K) if transB else (K, N), device=dev)
    C = torch.full((M, N), float("nan"), device=dev)
    ops.sgemm(A, Bm, C, M, N, K, A.shape[1], Bm.shape[1], N, transA, transB)
    ref = (A.t() if transA else A).double() @ (Bm.t() if transB else Bm).double()
    assert rel(C, ref) < 2e-6
