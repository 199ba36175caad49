"""The decoder skip blocks' backward with the concat split, the ScaleLong gain and the d(gain) reduction fused into the
epilogue of conv_3x3_1's data gradient (tedm_conv2d_dgrad_split + tedm_bias_add_bc) against the separate kernels it
replaces (TEDM_SPLIT_EPILOGUE=0: dgrad -> g_cat, channel_dot, block_prep_backward), same inputs, every gradient.
Autograd of src/tinyedm/networks.py:106-118 (ScaleLong) and :309-316 (concat)."""
import os

import pytest
import torch

from tests.helpers import rel

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _grads(model, batch):
    model.zero_grad(set_to_none=True)
    loss = model.training_step(batch, 0)
    loss.backward()
    return float(loss.detach()), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("B", [4, 40])
def test_split_epilogue_matches_separate_kernels(dev, B, monkeypatch):
    import tinyedm_b200 as T
    from tinyedm_b200.configs import CIFAR10, build_edm
    torch.manual_seed(3)
    model = build_edm(CIFAR10, dropout_rate=0.0, use_uncertainty=True).to(dev).eval()   # eval: weights stay put between the two runs
    with torch.no_grad():
        model.denoiser.gain_out.fill_(1.0)
        for p in model.parameters():
            if p.ndim == 0 and float(p) == 0.0:
                p.fill_(0.7)
    x = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1)
    y = torch.zeros(B, dtype=torch.long, device=dev)
    gen_state = torch.cuda.get_rng_state(dev)
    monkeypatch.setenv("TEDM_SPLIT_EPILOGUE", "1")
    torch.cuda.set_rng_state(gen_state, dev)
    loss_new, g_new = _grads(model, (x, y))
    monkeypatch.setenv("TEDM_SPLIT_EPILOGUE", "0")
    torch.cuda.set_rng_state(gen_state, dev)
    loss_old, g_old = _grads(model, (x, y))
    assert abs(loss_new - loss_old) < 1e-5           # the forward is untouched (the loss reduction uses fp32 atomics)
    assert set(g_new) == set(g_old)
    worst = ("", 0.0)
    for k in g_old:
        assert torch.isfinite(g_new[k]).all(), k
        if g_old[k].ndim == 0:
            continue
        r = rel(g_new[k], g_old[k])
        if r > worst[1]:
            worst = (k, r)
        # ScaleLong weights see the reduction most directly; everything else only through bf16 rounding of g_skip / g_in
        assert r < (2e-2 if "cat_factor" in k else 1e-2), (k, r)
    print("largest difference between the fused and the separate backward:", worst)


def test_bias_add_bc_matches_torch(dev):
    from tinyedm_b200 import ops
    ops.ensure_device(dev)
    torch.manual_seed(0)
    for (B, H, W, C) in [(3, 8, 8, 256), (2, 7, 5, 64), (5, 16, 16, 192)]:
        g = torch.randn(B, H, W, C, device=dev).to(BF)
        bias = torch.randn(B, C, device=dev)
        ref = (g.float() + 0.125 * bias[:, None, None, :]).to(BF)
        ops.bias_add_bc(g, bias, 0.125)
        assert torch.equal(g, ref)
