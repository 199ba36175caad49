"""conv_out fused with the output preconditioning, D = c_skip * noisy + c_out * gain_out * conv1x1(x)
(src/tinyedm/networks.py:579-581, :602-603) and its adjoint, against fp32 torch arithmetic on the same bf16 inputs.
Covers the channel counts of the three shipped configs (256 / 128 / 192), the generic kernel (C > 256), ragged pixel
counts (not a multiple of the 8- and 32-pixel trips) and both per-image and shared sigma."""
import pytest
import torch

from tests.helpers import rel

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (they fail, not skip, without one)"
    return torch.device("cuda:0")


def _reference(x, w, gain_out, noisy, sigma, sd):
    """fp32 autograd reference on the bf16-rounded operands (x is NHWC, noisy / D are NCHW)."""
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    g = gain_out.clone().requires_grad_(True)
    s = sigma.reshape(-1, 1, 1, 1) if sigma.numel() > 1 else sigma.reshape(1, 1, 1, 1)
    c_skip = sd * sd / (s * s + sd * sd)
    c_out = s * sd / torch.sqrt(s * s + sd * sd)
    f = torch.einsum("bhwc,oc->bohw", xf, wf)
    D = f * g * c_out + noisy * c_skip
    return xf, wf, g, f, D


@pytest.mark.parametrize("B,H,W,C,Co,shared_sigma", [
    (3, 32, 32, 256, 3, False),
    (2, 28, 28, 128, 1, False),
    (2, 16, 16, 192, 4, True),
    (5, 7, 5, 256, 3, False),      # 35 pixels per image: ragged against the 8- and 32-pixel trips
    (1, 3, 3, 64, 2, True),
    (2, 8, 8, 320, 3, False),      # C > 256: generic kernel
])
def test_conv_out_forward_backward(dev, B, H, W, C, Co, shared_sigma):
    from tinyedm_b200 import ops
    ops.ensure_device(dev)
    torch.manual_seed(B * 1000 + C + Co)
    sd = 0.5
    x = torch.randn(B, H, W, C, device=dev).to(BF)
    w = (torch.randn(Co, C, device=dev) / C ** 0.5).to(BF)
    gain_out = torch.tensor(0.8, device=dev)
    noisy = torch.randn(B, Co, H, W, device=dev)
    sigma = (torch.rand(1, device=dev) * 3 + 0.1) if shared_sigma else (torch.rand(B, device=dev) * 3 + 0.1)

    D, f_raw = ops.conv_out_forward(x, w, gain_out, noisy, sigma, sd, keep_raw=True)
    xf, wf, g, f_ref, D_ref = _reference(x, w, gain_out, noisy, sigma, sd)
    assert rel(f_raw, f_ref) < 1e-5
    assert rel(D, D_ref) < 1e-5
    D2, none = ops.conv_out_forward(x, w, gain_out, noisy, sigma, sd, keep_raw=False)
    assert none is None and torch.equal(D2, D)

    g_D = torch.randn_like(D)
    D_ref.backward(g_D)
    g_w = torch.zeros(Co, C, device=dev)
    g_gain = torch.zeros((), device=dev)
    g_x = ops.conv_out_backward(g_D, f_raw, x, w, gain_out, sigma, sd, g_w, g_gain)
    assert g_x.dtype == BF and g_x.shape == x.shape
    assert rel(g_x, xf.grad) < 6e-3            # bf16 rounding of the stored gradient
    assert rel(g_w, wf.grad) < 1e-4
    assert abs(g_gain.item() - g.grad.item()) <= 1e-4 * max(1.0, abs(g.grad.item()))
