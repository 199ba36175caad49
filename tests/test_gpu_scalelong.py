"""ScaleLong (src/tinyedm/networks.py:106-118: sigmoid(W2 mp_silu(W1 [mean_HW(skip), 1]))) — the three fp32 kernels
(tedm_scalelong_forward / _backward / _wgrad) against the oracle's `scale_long` and torch autograd, at the configs' own
batch sizes and skip widths (CIFAR 256 @ B=256, MNIST 128/512 @ B=128, ImageNet-latent 192/768 @ micro-batch 176) and at
ragged sizes (batch not a multiple of the rows per CTA, R not a multiple of the vector width)."""
import pytest
import torch

from oracle import edm2_oracle as O
from tests.helpers import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from tinyedm_b200 import ops
    d = torch.device("cuda:0")
    ops.ensure_device(d)
    return d


@pytest.mark.parametrize("B,C,R", [(256, 256, 16), (128, 128, 8), (128, 512, 32), (176, 192, 12), (176, 768, 48), (64, 768, 48),
                                   (1, 64, 4), (5, 64, 4), (3, 320, 20), (7, 96, 6), (33, 1024, 64)])
def test_scalelong_forward_backward_wgrad_vs_oracle(dev, B, C, R):
    from tinyedm_b200 import ops
    g = torch.Generator().manual_seed(B + C)
    w1 = torch.randn(R, C + 1, 1, 1, generator=g)
    w2 = torch.randn(C, R, 1, 1, generator=g)
    w1h, w2h = O.effective_weight(w1).to(dev), O.effective_weight(w2).to(dev)       # what the weight bank hands the kernels
    mean = torch.randn(B, C, generator=g).to(dev)
    aug, h_pre, h, gain = ops.scalelong_forward(mean, w1h.view(R, C + 1).contiguous(), w2h.view(C, R).contiguous(), R)
    # oracle: the module's own graph on a (B,C,1,1) "skip" whose spatial mean is `mean`
    p = {"layer1.weight": w1.to(dev).requires_grad_(True), "layer2.weight": w2.to(dev).requires_grad_(True)}
    skip = mean.view(B, C, 1, 1).clone().requires_grad_(True)
    gain_o = O.scale_long(p, "", skip)
    assert rel(gain, gain_o.view(B, C)) < 1e-5
    assert torch.equal(aug[:, :C], mean) and bool((aug[:, C] == 1).all())
    d_gain = torch.randn(B, C, generator=g).to(dev)
    g_skip, g_w1, g_w2 = torch.autograd.grad(gain_o, [skip, p["layer1.weight"], p["layer2.weight"]], d_gain.view(B, C, 1, 1))
    for times_gain in (False, True):
        dg = d_gain * gain if times_gain else d_gain
        d_pre2, d_hpre, d_mean = ops.scalelong_backward(dg.contiguous(), gain, h_pre, w1h.view(R, C + 1), w2h.view(C, R),
                                                        d_gain_times_gain=times_gain)
        assert rel(d_mean, g_skip.view(B, C)) < 1e-4
    # dL/dw_hat of both layers (accumulated onto what is already there), pushed through the weight-norm Jacobian by autograd
    dw2 = torch.full((C, R), 0.5, device=dev)
    dw1 = torch.full((R, C + 1), -0.25, device=dev)
    ops.scalelong_wgrad(d_pre2, h, d_hpre, aug, dw2, dw1)
    for w, dw, off, gw in ((w1, dw1, -0.25, g_w1), (w2, dw2, 0.5, g_w2)):
        wl = w.to(dev).requires_grad_(True)
        (want,) = torch.autograd.grad(O.effective_weight(wl), wl, (dw - off).view_as(wl))
        assert rel(want, gw) < 1e-4
