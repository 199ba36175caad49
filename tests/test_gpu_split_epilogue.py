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
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (they fail, not skip, without one)"
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
    monkeypatch.setenv("TEDM_SPLIT_EPILOGUE", "1")
    model.diffuser.seed(77, 0)               # the same (seed, step) -> the same diffusion noise in both runs
    loss_new, g_new = _grads(model, (x, y))
    monkeypatch.setenv("TEDM_SPLIT_EPILOGUE", "0")
    model.diffuser.seed(77, 0)
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


@pytest.mark.parametrize("B", [3, 48])
def test_skip_mean_from_conv_epilogue_matches_channel_dot(dev, B, monkeypatch):
    """ScaleLong's spatial mean taken from the producing conv's epilogue (fixed-order partial sums) against the separate
    reduction (TEDM_FUSED_SKIP_MEAN=0); B = 48 is large enough for the CTA-pair kernel at every level of the CIFAR net.
    Also: the fused path is bit-reproducible (no atomics)."""
    from tinyedm_b200.configs import CIFAR10, build_edm
    torch.manual_seed(5)
    model = build_edm(CIFAR10, num_classes=10, dropout_rate=0.0).to(dev).eval()
    with torch.no_grad():
        model.denoiser.gain_out.fill_(1.0)
    x = torch.randn(B, 3, 32, 32, device=dev)
    sigma = torch.full((B,), 1.3, device=dev)
    y = torch.randint(0, 10, (B,), device=dev)
    with torch.no_grad():
        monkeypatch.setenv("TEDM_FUSED_SKIP_MEAN", "1")
        d1 = model(x, sigma, y).clone()
        d1b = model(x, sigma, y).clone()
        monkeypatch.setenv("TEDM_FUSED_SKIP_MEAN", "0")
        d0 = model(x, sigma, y).clone()
    assert torch.equal(d1, d1b)
    # Self-consistency of two equally valid bf16 paths, NOT a parity claim: the means differ in summation order (fixed-order
    # partial sums vs one CTA per channel group), which moves a few bf16 roundings downstream; measured 1.6-2.0e-3, once
    # 2.015e-3 — hence 4e-3 rather than the 2e-3 this assert started with. Parity of the fused mean itself is pinned
    # against fp32 torch arithmetic in test_colsum_partials_equal_the_column_sums (<= 1e-5), and the networks that use it
    # against the oracle in tests/test_gpu_parity.py / test_gpu_block_parity.py.
    assert rel(d1, d0) < 4e-3, rel(d1, d0)


def test_colsum_partials_equal_the_column_sums(dev):
    """col_partial of a PLAIN and an AXPBY conv: summed per image it equals the spatial sum of the stored bf16 output."""
    from tinyedm_b200 import ops
    from tinyedm_b200.ops import EPI_AXPBY, EPI_PLAIN
    ops.ensure_device(dev)
    torch.manual_seed(1)
    for (B, H, W, Cin, Cout, ks) in [(40, 32, 32, 64, 256, 1), (150, 16, 16, 128, 256, 3), (300, 8, 8, 64, 128, 3)]:
        x = torch.randn(B, H, W, Cin, device=dev).to(BF)
        w = (torch.randn(Cout, ks * ks * Cin, device=dev) / (ks * ks * Cin) ** 0.5).to(BF)
        res = torch.randn(B, H, W, Cout, device=dev).to(BF)
        for epi, kw in ((EPI_PLAIN, {}), (EPI_AXPBY, dict(alpha=0.6, beta=0.8, res=res))):
            slots = ops.conv2d_colsum_slots(B, H, W, Cin, Cout, ks, epi)
            assert slots > 0, (B, H, W, Cin, Cout, ks)
            part = torch.full((B * slots, Cout), float("nan"), device=dev)
            out = ops.conv2d(x, w, ks, Cout, epi=epi, col_partial=part, **kw)
            assert torch.isfinite(part).all()                      # every entry written
            mean = ops.colsum_mean(part, B, slots, Cout, 1.0 / (H * W))
            ref = out.float().mean(dim=(1, 2))
            assert rel(mean, ref) < 1e-5
            assert torch.equal(out, ops.conv2d(x, w, ks, Cout, epi=epi, **kw))   # the conv result itself is unchanged
