"""The CTA-pair (tcgen05 cta_group::2) conv / wgrad kernels, which a launch only takes when its pair tiles cover at least
half of the 74 CTA pairs: shapes here are large enough for that. Every fused epilogue is checked against fp32 torch
arithmetic of what it replaces (src/tinyedm/networks.py:37, :206, :255-263 and their autograd), and against the
single-CTA kernel (block_n override) on the same inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (they fail, not skip, without one)"
    return torch.device("cuda:0")


def _ref_conv(x, w):
    return F.conv2d(x.float().permute(0, 3, 1, 2), w.to(BF).float(), padding="same").permute(0, 2, 3, 1).contiguous()


def _wq(w):
    return w.permute(0, 2, 3, 1).contiguous().to(BF).reshape(w.shape[0], -1)


def _silu(x):
    return F.silu(x) / 0.596


def _silu_grad(x):
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s)) / 0.596


# B, H, W, Cin, Cout, ks — pair tiles: (B*H*W/256) * ceil(Cout/256) > 37
CASES = [(40, 32, 32, 256, 256, 3), (41, 16, 16, 256, 256, 3), (150, 8, 8, 256, 256, 3), (40, 16, 16, 256, 768, 1),
         (24, 28, 28, 128, 128, 3), (90, 14, 14, 256, 256, 1), (12, 64, 64, 192, 192, 3), (161, 7, 7, 512, 512, 1),
         # tail split: 100 resp. 150 pair tiles on 74 CTA pairs leave 26 resp. 2 tiles, cut into half-N work items
         (100, 16, 16, 256, 256, 3), (50, 16, 16, 256, 768, 1),
         # BASELINE.json's own batch sizes, element-wise: training B = 256 at every CIFAR level (the
         # headline launch: 1 024 pair tiles in 14 waves), the concatenated decoder input, the qkv conv; sampling B = 128
         (256, 32, 32, 256, 256, 3), (256, 16, 16, 256, 256, 3), (256, 8, 8, 256, 256, 3), (256, 16, 16, 512, 256, 3),
         (256, 16, 16, 256, 768, 1), (128, 32, 32, 256, 256, 3), (128, 16, 16, 256, 256, 3), (128, 8, 8, 256, 256, 3),
         # ImageNet-latent micro-batch 176 (imagenet.yaml:14) at its first level, MNIST batch 128
         (176, 64, 64, 192, 192, 3), (128, 28, 28, 128, 128, 3)]


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks", CASES)
def test_pair_conv_epilogues_vs_torch(dev, B, H, W, Cin, Cout, ks):
    from tinyedm_b200 import ops
    from tinyedm_b200.ops import EPI_AXPBY, EPI_MODSILU, EPI_MODSILU_BWD, EPI_SILU_BWD
    ops.ensure_device(dev)
    torch.manual_seed(B + H + Cout)
    x = torch.randn(B, H, W, Cin, device=dev).to(BF)
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / math.sqrt(Cin * ks * ks)
    wq = _wq(w)
    acc = _ref_conv(x, w)
    bn_single = 256 if Cout >= 256 else 128
    # plain
    y = ops.conv2d(x, wq, ks, Cout, alpha=0.7)
    assert rel(y, 0.7 * acc) < 6e-3
    assert rel(y, ops.conv2d(x, wq, ks, Cout, alpha=0.7, block_n=bn_single)) < 1e-4
    # mp_add
    res = torch.randn(B, H, W, Cout, device=dev).to(BF)
    y = ops.conv2d(x, wq, ks, Cout, epi=EPI_AXPBY, alpha=0.4, beta=0.9, res=res)
    assert rel(y, 0.4 * acc + 0.9 * res.float()) < 6e-3
    # mp_add whose residual is the UN-normalised block input, divided by its pixel norm in the epilogue (networks.py:249, :263)
    rn = (torch.rand(B, H, W, device=dev) + 0.5).contiguous()
    want_n = 0.4 * acc + 0.9 * res.float() / rn[..., None]
    y = ops.conv2d(x, wq, ks, Cout, epi=EPI_AXPBY, alpha=0.4, beta=0.9, res=res, nrm=rn)
    assert rel(y, want_n) < 6e-3
    assert rel(ops.conv2d(x, wq, ks, Cout, epi=EPI_AXPBY, alpha=0.4, beta=0.9, res=res, nrm=rn, block_n=bn_single), want_n) < 6e-3
    # modulation * mp_silu (+ raw copy), no dropout
    mod = (torch.randn(B, Cout, device=dev) * 0.3 + 1).contiguous()
    raw = torch.zeros(B, H, W, Cout, device=dev, dtype=BF)
    y = ops.conv2d(x, wq, ks, Cout, epi=EPI_MODSILU, mod=mod, raw=raw)
    assert rel(raw, acc) < 6e-3
    assert rel(y, _silu(raw.float() * mod[:, None, None, :])) < 6e-3
    # dropout statistics + determinism of the counter-based mask
    yd = ops.conv2d(x, wq, ks, Cout, epi=EPI_MODSILU, mod=mod, drop_p=0.13, seed=9)
    keep = (yd != 0).float().mean().item()
    assert abs(keep - 0.87) < 0.01
    assert torch.equal(yd, ops.conv2d(x, wq, ks, Cout, epi=EPI_MODSILU, mod=mod, drop_p=0.13, seed=9))
    # adjoint of modulation * mp_silu: g_raw and d_mod
    rawt = torch.randn(B, H, W, Cout, device=dev).to(BF)
    d_mod = torch.zeros(B, Cout, device=dev)
    g_raw = ops.conv2d(x, wq, ks, Cout, epi=EPI_MODSILU_BWD, alpha=0.5, aux=rawt, mod=mod, d_mod=d_mod)
    gz = 0.5 * acc * _silu_grad(rawt.float() * mod[:, None, None, :])
    assert rel(g_raw, gz * mod[:, None, None, :]) < 6e-3
    assert rel(d_mod, (gz * rawt.float()).sum(dim=(1, 2))) < 5e-3
    # adjoint of mp_silu + residual share (+ accumulation)
    xs = torch.randn(B, H, W, Cout, device=dev).to(BF)
    want = 1.1 * acc * _silu_grad(xs.float()) + 0.6 * res.float()
    y = ops.conv2d(x, wq, ks, Cout, epi=EPI_SILU_BWD, alpha=1.1, aux=xs, res=res, beta=0.6)
    assert rel(y, want) < 6e-3
    old = torch.randn(B, H, W, Cout, device=dev).to(BF)
    out = old.clone()
    ops.conv2d(x, wq, ks, Cout, epi=EPI_SILU_BWD, alpha=1.1, aux=xs, res=res, beta=0.6, out=out, accumulate_out=True)
    assert rel(out, want + old.float()) < 6e-3
    if Cout <= 256:
        # fused pixel-norm adjoint: g_u = g/n - x (g.x) / ((n - eps) C)
        nrm = (torch.rand(B, H, W, device=dev) + 0.5).contiguous()
        y = ops.conv2d(x, wq, ks, Cout, epi=EPI_SILU_BWD, alpha=1.1, aux=xs, res=res, beta=0.6, nrm=nrm)
        n = nrm[..., None]
        ref = want / n - xs.float() * (want * xs.float()).sum(-1, keepdim=True) / ((n - 1e-4) * Cout)
        assert rel(y, ref) < 6e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks", [(40, 32, 32, 256, 256, 3), (33, 16, 16, 512, 256, 3), (9, 16, 16, 256, 768, 1),
                                               (70, 8, 8, 128, 256, 3), (5, 14, 14, 256, 256, 3), (3, 7, 7, 512, 512, 3),
                                               # transposed pair kernel (k slabs on M, output channels on N): the ImageNet-latent
                                               # and MNIST channel counts, N = 192 blocks, padded last N block, padded k slab
                                               (3, 64, 64, 192, 192, 3), (5, 32, 32, 384, 384, 3), (4, 16, 16, 576, 576, 3),
                                               (4, 32, 32, 576, 384, 3), (6, 16, 16, 576, 1728, 1), (2, 28, 28, 128, 128, 3),
                                               (3, 16, 16, 1344, 768, 3), (2, 8, 8, 768, 576, 1), (9, 8, 8, 64, 128, 3),
                                               (3, 16, 16, 128, 320, 3), (12, 64, 64, 192, 192, 3), (5, 7, 7, 128, 384, 3),
                                               # BASELINE.json's training batch, element-wise
                                               (256, 32, 32, 256, 256, 3), (256, 16, 16, 512, 256, 3), (256, 8, 8, 256, 256, 3),
                                               (256, 16, 16, 256, 768, 1), (176, 64, 64, 192, 192, 3)])
def test_pair_wgrad_vs_torch_and_single(dev, B, H, W, Cin, Cout, ks):
    from tinyedm_b200 import ops
    ops.ensure_device(dev)
    torch.manual_seed(B + Cin)
    x = torch.randn(B, H, W, Cin, device=dev).to(BF)
    g = torch.randn(B, H, W, Cout, device=dev).to(BF)
    K = B * H * W                      # contraction length of the weight gradient
    big = K > 100_000
    dt = torch.float64 if big else torch.float32     # the fp32 torch reference itself drifts at K ~ 10^5..10^6: use fp64 there
    w0 = torch.zeros(Cout, Cin, ks, ks, device=dev, dtype=dt, requires_grad=True)
    (gw,) = torch.autograd.grad(F.conv2d(x.to(dt).permute(0, 3, 1, 2), w0, padding="same"), w0, g.to(dt).permute(0, 3, 1, 2))
    ref = gw.permute(0, 2, 3, 1).reshape(Cout, ks * ks, Cin)
    # fp32 accumulation in the tensor core (products exact, running sum rounded once per k-block of 16 pixels) plus fp32
    # split-K reduce-adds: the error grows with the contraction length. 1e-4 up to K = 50 000, then 2e-9 per pixel
    # (measured 1.7e-4 at K = 262 144 and 3.9e-4 at K = 720 896 against fp64)
    tol = max(1e-4, 2e-9 * K)
    for acc in (False, True):
        dw = torch.full((Cout, ks * ks, Cin), 3.0, device=dev)
        ops.conv2d_wgrad(g, x, dw, ks, alpha=0.5, accumulate=acc)                 # CTA-pair kernel (TMA reduce-add)
        assert rel(dw, 0.5 * ref + (3.0 if acc else 0.0)) < tol
        if big and acc:
            continue                       # the single-CTA comparison once per big case is enough
        dws = torch.full((Cout, ks * ks, Cin), 3.0, device=dev)
        ops.conv2d_wgrad(g, x, dws, ks, alpha=0.5, accumulate=acc, splits=-1)      # single-CTA kernel (vector atomics)
        assert rel(dw, dws) < 2 * tol


@pytest.mark.parametrize("B,H,W,Cin,C1,C2", [(40, 32, 32, 256, 256, 256), (150, 8, 8, 256, 256, 256), (256, 32, 32, 256, 256, 256),
                                             (256, 16, 16, 256, 256, 256)])
def test_dgrad_split_epilogue_vs_torch(dev, B, H, W, Cin, C1, C2):
    """tedm_conv2d_dgrad_split element-wise against fp32 torch: g_cat = alpha * dgrad * mp_silu'(x) + beta * res split into
    g_in (+ accumulation + per-(image, channel) bias), g_skip * gain and the reduction sum_pixels g_cat * x (autograd of
    networks.py:309-316 through :106-118's gain)."""
    from tinyedm_b200 import ops
    ops.ensure_device(dev)
    if not ops.conv2d_dgrad_split_supported(B, H, W, Cin, C1, C2, 3):
        pytest.skip("launch too small for the CTA-pair kernel")
    torch.manual_seed(B + H)
    Ct = C1 + C2
    g = torch.randn(B, H, W, Cin, device=dev).to(BF)
    w = torch.randn(Ct, Cin, 3, 3, device=dev) / math.sqrt(Cin * 9)
    acc = _ref_conv(g, w)
    x = torch.randn(B, H, W, Ct, device=dev).to(BF)
    res = torch.randn(B, H, W, Ct, device=dev).to(BF)
    gain = torch.rand(B, C2, device=dev) * 0.8 + 0.1
    bias = torch.randn(B, C1, device=dev)
    old = torch.randn(B, H, W, C1, device=dev).to(BF)
    g_cat = 0.9 * acc * _silu_grad(x.float()) + 0.7 * res.float()
    for accumulate in (False, True):
        g_in = old.clone()
        g_skip = torch.empty(B, H, W, C2, device=dev, dtype=BF)
        d_gx = torch.zeros(B, C2, device=dev)
        ops.conv2d_dgrad_split(g, _wq(w), 3, x=x, res=res, beta=0.7, gain=gain, g_in=g_in, g_skip=g_skip, d_gx=d_gx,
                               accumulate_in=accumulate, alpha=0.9, in_bias=bias if accumulate else None, in_bias_scale=0.25)
        want_in = g_cat[..., :C1] + ((old.float() + 0.25 * bias[:, None, None, :]) if accumulate else 0.0)
        assert rel(g_in, want_in) < 6e-3
        assert rel(g_skip, g_cat[..., C1:] * gain[:, None, None, :]) < 6e-3
        assert rel(d_gx, (g_cat[..., C1:] * x[..., C1:].float()).sum(dim=(1, 2))) < 5e-3
