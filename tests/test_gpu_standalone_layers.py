"""The reference's exported layer classes used ON THEIR OWN (`from tinyedm import Conv2d, Linear`,
src/tinyedm/__init__.py:9): forward, the training-mode in-place weight normalisation (networks.py:32-34, :55-57) and
autograd against the oracle's mp_conv2d / mp_linear — including channel counts that are not multiples of 64 (the
reference's own ScaleLong convolutions are 257 -> 16 and 16 -> 256, networks.py:109-110)."""
import pytest
import torch

from oracle import edm2_oracle as O
from tests.helpers import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.mark.parametrize("cin,cout,k,B,H", [(64, 128, 3, 2, 16), (257, 16, 1, 3, 1), (16, 256, 1, 3, 1), (40, 72, 3, 2, 8),
                                            (3, 64, 3, 2, 32), (128, 3, 1, 2, 16),
                                            # widths the weight-gradient kernel cannot tile: zero columns are added around it
                                            (64, 64, 3, 3, 9), (24, 70, 3, 2, 17), (64, 128, 1, 2, 13)])
def test_standalone_conv2d_vs_oracle(dev, cin, cout, k, B, H):
    import tinyedm_b200 as T
    torch.manual_seed(cin + cout)
    conv = T.Conv2d(cin, cout, k).to(dev)
    w0 = conv.weight.detach().clone()
    x = torch.randn(B, cin, H, H, device=dev).to(torch.bfloat16).float().requires_grad_(True)
    conv.eval()
    y = conv(x)
    assert y.shape == (B, cout, H, H)
    assert torch.equal(conv.weight.detach(), w0)                       # eval leaves the parameter alone
    wo = w0.clone().requires_grad_(True)
    xo = x.detach().clone().requires_grad_(True)
    yo = O.mp_conv2d(xo, wo)
    assert rel(y, yo) < 1e-2
    g = torch.randn_like(yo).to(torch.bfloat16).float()
    y.backward(g.to(y.dtype))
    yo.backward(g)
    assert rel(x.grad, xo.grad) < 1e-2
    assert rel(conv.weight.grad, wo.grad) < 1e-2
    conv.train()
    conv(x.detach())
    assert rel(conv.weight.detach(), O.normalize_weight(w0)) < 1e-6     # forced weight normalisation, in place


def test_standalone_linear_vs_oracle(dev):
    import tinyedm_b200 as T
    torch.manual_seed(5)
    lin = T.Linear(65, 48).to(dev).eval()
    x = torch.randn(7, 65, device=dev, requires_grad=True)
    y = lin(x)
    wo = lin.weight.detach().clone().requires_grad_(True)
    xo = x.detach().clone().requires_grad_(True)
    yo = O.mp_linear(xo, wo)
    assert rel(y, yo) < 1e-5
    g = torch.randn_like(yo)
    y.backward(g); yo.backward(g)
    assert rel(x.grad, xo.grad) < 1e-5 and rel(lin.weight.grad, wo.grad) < 1e-4
    with pytest.raises(RuntimeError):
        T.Conv2d(64, 64, 3)(torch.zeros(1, 64, 8, 8))                  # CPU tensor: no fallback
