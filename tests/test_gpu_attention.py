"""Cosine attention kernels against CosineAttention.forward's arithmetic (src/tinyedm/networks.py:194-202) in fp32 torch:
the kernels specialised for head_dim 64 / S in {64, 256} (csrc/attention_tc.cu), the generic tcgen05 pair on pre-normalised
q, k, v for every other shape of the configs (csrc/attention_gen.cu), and the warp-MMA kernels behind tedm_attention_forward
for shapes neither covers."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (they fail, not skip, without one)"
    return torch.device("cuda:0")


def reference(qkv: torch.Tensor, heads: int):
    """qkv (B,S,3C) with channel = {q,k,v}*C + head*hd + d  ->  y (B,S,C), fp32, differentiable w.r.t. qkv."""
    B, S, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    t = qkv.view(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4)                  # (3,B,heads,S,hd)
    n = t.norm(dim=-1, keepdim=True) / math.sqrt(hd)
    t = t / (1e-4 + n)                                                        # pixel_norm over hd (networks.py:9-14)
    y = F.scaled_dot_product_attention(t[0], t[1], t[2])                      # (B,heads,S,hd), scale 1/sqrt(hd)
    return y.permute(0, 2, 1, 3).reshape(B, S, C)


def reference_bf16(qkv: torch.Tensor, heads: int, g_y: torch.Tensor):
    """The reference's OWN arithmetic under bf16 autocast on this GPU (networks.py:194-202 with a bf16 qkv: norm in fp32,
    cast to bf16 before the divide (:14), fused bf16 SDPA and torch autograd): returns (y, d qkv). Its distance from the
    fp32 evaluation of the same graph is the noise floor of the path this library replaces."""
    B, S, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    x = qkv.detach().clone().requires_grad_(True)
    t = x.view(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    n = torch.linalg.vector_norm(t, dim=-1, keepdim=True, dtype=torch.float32) / math.sqrt(hd)
    t = t / (1e-4 + n).to(t.dtype)
    y = F.scaled_dot_product_attention(t[0], t[1], t[2]).permute(0, 2, 1, 3).reshape(B, S, C)
    (g,) = torch.autograd.grad(y, x, g_y.to(y.dtype))
    return y.detach(), g


NORTH_STAR_TOL = 1e-2     # bf16 per-layer relative L2


def _assert_within(r: float, floor: float, what):
    """<= 1e-2 (north_star), or — where bf16 itself cannot do that — no worse than the reference's own bf16 path on the
    same inputs (both numbers are printed)."""
    assert r < NORTH_STAR_TOL or r <= floor, (what, f"ours {r:.3e}", f"reference bf16 path {floor:.3e}")


@pytest.mark.parametrize("B,H,W,heads,hd", [(3, 16, 16, 4, 64), (5, 8, 8, 4, 64), (3, 8, 8, 3, 64), (1, 8, 8, 1, 64),
                                            (2, 14, 14, 4, 64), (2, 7, 7, 4, 128), (2, 16, 16, 4, 144), (2, 8, 8, 4, 192),
                                            # the configs' own batch sizes: CIFAR train 256 / sampling 128, MNIST 128
                                            (256, 16, 16, 4, 64), (256, 8, 8, 4, 64), (128, 14, 14, 4, 64), (128, 7, 7, 4, 128),
                                            (64, 16, 16, 4, 144), (64, 8, 8, 4, 192)])
def test_attention_forward_backward_vs_torch(dev, B, H, W, heads, hd):
    from tinyedm_b200 import ops
    ops.ensure_device(dev)
    torch.manual_seed(B * 1000 + H * 10 + heads)
    C = heads * hd
    qkv = (torch.randn(B, H, W, 3 * C, device=dev) * 1.3).to(torch.bfloat16)
    y, lse = ops.attention_forward(qkv, heads, need_lse=True)
    x = qkv.float().view(B, H * W, 3 * C).requires_grad_(True)
    with torch.backends.cuda.sdp_kernel(enable_flash=False, enable_mem_efficient=False, enable_math=True):
        ref = reference(x, heads)
    r = rel(y.view(B, H * W, C), ref)
    assert r < 1e-2, r                                    # north_star: bf16 per-layer relative L2 <= 1e-2
    assert torch.isfinite(lse).all()
    # log-sum-exp of the scaled scores of the normalised q, k
    t = x.detach().view(B, H * W, 3, heads, hd).permute(2, 0, 3, 1, 4)
    t = t / (1e-4 + t.norm(dim=-1, keepdim=True) / math.sqrt(hd))
    lse_ref = torch.logsumexp(t[0] @ t[1].transpose(-1, -2) / math.sqrt(hd), dim=-1)    # (B,heads,S)
    assert rel(lse.view(B, heads, H * W), lse_ref) < 2e-3
    g_y = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    g_qkv = ops.attention_backward(qkv, y, g_y, lse, heads)
    (g_ref,) = torch.autograd.grad(ref, x, g_y.float().view(B, H * W, C))
    rg = rel(g_qkv.view(B, H * W, 3 * C), g_ref)
    y16, g16 = reference_bf16(qkv.view(B, H * W, 3 * C), heads, g_y.view(B, H * W, C))
    floor_f, floor_b = rel(y16, ref), rel(g16, g_ref)
    print(f"attention B={B} S={H * W} hd={hd}: forward {r:.2e} (reference bf16 path {floor_f:.2e}), "
          f"backward {rg:.2e} (reference bf16 path {floor_b:.2e})")
    _assert_within(rg, floor_b, "d qkv")


@pytest.mark.parametrize("B,heads", [(2, 4), (3, 1), (41, 4), (5, 3)])
def test_attention_backward_s256_each_gradient(dev, B, heads):
    """S = 256, head_dim 64 (the fused one-CTA-per-head backward): dq, dk and dv separately against fp32 autograd,
    more (image, head) pairs than SMs in one case so that CTAs are reused."""
    from tinyedm_b200 import ops
    ops.ensure_device(dev)
    torch.manual_seed(7 * B + heads)
    hd, H, W = 64, 16, 16
    C = heads * hd
    qkv = (torch.randn(B, H, W, 3 * C, device=dev) * 1.3).to(torch.bfloat16)
    y, lse = ops.attention_forward(qkv, heads, need_lse=True)
    g_y = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    g_qkv = ops.attention_backward(qkv, y, g_y, lse, heads).view(B, H * W, 3, C).float()
    x = qkv.float().view(B, H * W, 3 * C).requires_grad_(True)
    with torch.backends.cuda.sdp_kernel(enable_flash=False, enable_mem_efficient=False, enable_math=True):
        ref = reference(x, heads)
    (g_ref,) = torch.autograd.grad(ref, x, g_y.float().view(B, H * W, C))
    g_ref = g_ref.view(B, H * W, 3, C)
    assert torch.isfinite(g_qkv).all()
    _, g16 = reference_bf16(qkv.view(B, H * W, 3 * C), heads, g_y.view(B, H * W, C))
    g16 = g16.view(B, H * W, 3, C)
    for part, name in enumerate("qkv"):
        r = rel(g_qkv[:, :, part], g_ref[:, :, part])
        floor = rel(g16[:, :, part], g_ref[:, :, part])
        print(f"attention backward S=256 B={B} heads={heads}: d{name} {r:.2e} (reference bf16 path {floor:.2e})")
        _assert_within(r, floor, "d" + name)
    # deterministic: no atomics anywhere in the backward
    again = ops.attention_backward(qkv, y, g_y, lse, heads).view(B, H * W, 3, C).float()
    assert torch.equal(again, g_qkv)


GEN_SHAPES = [  # B, H, W, heads, hd — the configs' own (head_dim, S) pairs first, at small and at full batch
    (2, 14, 14, 4, 64), (2, 7, 7, 4, 128), (2, 16, 16, 4, 144), (2, 8, 8, 4, 192),
    (128, 14, 14, 4, 64), (128, 7, 7, 4, 128), (64, 16, 16, 4, 144), (64, 8, 8, 4, 192),
    # the CIFAR shapes through the generic kernels too, ragged S / head_dim, a single head, one image
    (3, 16, 16, 4, 64), (5, 8, 8, 4, 64), (3, 5, 5, 3, 16), (1, 9, 9, 1, 80), (2, 13, 11, 2, 48), (1, 16, 16, 2, 128), (3, 1, 1, 4, 64)]


@pytest.mark.parametrize("B,H,W,heads,hd", GEN_SHAPES)
def test_generic_tcgen05_attention_forward_backward_vs_torch(dev, B, H, W, heads, hd):
    """qkv_normalize + the tcgen05 forward / backward pair (csrc/attention_gen.cu) against fp32 torch arithmetic of
    networks.py:194-202 and its autograd, <= 1e-2 (north_star) with the reference's own bf16 path printed beside it."""
    from tinyedm_b200 import ops
    ops.ensure_device(dev)
    torch.manual_seed(B * 1000 + H * 10 + heads + hd)
    C, S = heads * hd, H * W
    qkv = (torch.randn(B, H, W, 3 * C, device=dev) * 1.3).to(torch.bfloat16)
    qn, norms = ops.qkv_normalize(qkv, heads)
    t = qkv.float().view(B, S, 3 * heads, hd)
    n_ref = 1e-4 + t.norm(dim=-1) / math.sqrt(hd)
    assert rel(norms.view(B, S, 3 * heads), n_ref) < 1e-5
    assert rel(qn.view(B, S, 3 * heads, hd), t / n_ref[..., None]) < 4e-3          # one bf16 rounding
    y, lse = ops.attention_forward_normalized(qn, heads, need_lse=True)
    x = qkv.float().view(B, S, 3 * C).requires_grad_(True)
    with torch.backends.cuda.sdp_kernel(enable_flash=False, enable_mem_efficient=False, enable_math=True):
        ref = reference(x, heads)
    r = rel(y.view(B, S, C), ref)
    tt = x.detach().view(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    tt = tt / (1e-4 + tt.norm(dim=-1, keepdim=True) / math.sqrt(hd))
    lse_ref = torch.logsumexp(tt[0] @ tt[1].transpose(-1, -2) / math.sqrt(hd), dim=-1)
    assert torch.isfinite(lse).all() and rel(lse.view(B, heads, S), lse_ref) < 2e-3
    g_y = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    g_qkv = ops.attention_backward_normalized(qn, norms, y, g_y, lse, heads)
    (g_ref,) = torch.autograd.grad(ref, x, g_y.float().view(B, S, C))
    g_ours = g_qkv.view(B, S, 3, C).float()
    g_ref = g_ref.view(B, S, 3, C)
    # (S = 1: the softmax is the constant 1, d q and d k are exactly zero in the reference — measure those against the
    # size of d v instead of against zero)
    floor_norm = 1e-3 * float(g_ref[:, :, 2].norm())
    parts = {name: float((g_ours[:, :, i] - g_ref[:, :, i]).norm()) / max(float(g_ref[:, :, i].norm()), floor_norm)
             for i, name in enumerate("qkv")}
    y16, g16 = reference_bf16(qkv.view(B, S, 3 * C), heads, g_y.view(B, S, C))
    floor_f, floor_b = rel(y16, ref), rel(g16, g_ref.reshape(B, S, 3 * C))
    print(f"generic attention B={B} S={S} hd={hd}: forward {r:.2e} (reference bf16 path {floor_f:.2e}), backward "
          + ", ".join(f"d{k} {v:.2e}" for k, v in parts.items()) + f" (reference bf16 path {floor_b:.2e})")
    assert torch.isfinite(g_qkv).all()
    _assert_within(r, floor_f, "y")
    for k, v in parts.items():
        _assert_within(v, floor_b, "d" + k)
    # no atomics anywhere: bit-reproducible
    y2, lse2 = ops.attention_forward_normalized(qn, heads, need_lse=True)
    assert torch.equal(y2, y) and torch.equal(lse2, lse)
    assert torch.equal(ops.attention_backward_normalized(qn, norms, y, g_y, lse, heads), g_qkv)
