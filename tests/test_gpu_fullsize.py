"""BASELINE.json's full sizes (CIFAR config, per-GPU batch 256 / sampling batch 128), where the CPU oracle would take
minutes: size-independent properties instead of element-wise comparison (task brief §3).
  * bilinear identities of the convolution triple: <conv(x,w), g> == <x, dgrad(g,w)> == <w, wgrad(g,x)>
  * forced weight normalisation: the parameter gradient is orthogonal to the weight row
  * data linearity of the denoiser's skip path and determinism / graph-replay identity of the 32-step sampler
"""
import math

import pytest
import torch

from tests.helpers import rel

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _eager(solver, model, x0, labels):
    """The solver's eager loop exactly as `solve` runs it (under no_grad: with gradients enabled the Denoiser takes its
    training-style path, which stores the pixel-normalised block inputs and is not bit-identical to the inference path)."""
    with torch.no_grad():
        return solver._solve_eager(model, x0, labels)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (they fail, not skip, without one)"
    return torch.device("cuda:0")


# CIFAR shapes at B = 256; ImageNet-latent shapes at B = 64 (channel counts that are multiples of 192: N = 192 conv tiles,
# the transposed CTA-pair weight-gradient kernel, a concatenated decoder input, the qkv 1x1 conv)
@pytest.mark.parametrize("B,H,Cin,Cout,ks", [(256, 32, 256, 256, 3), (256, 16, 512, 256, 3), (256, 16, 256, 768, 1),
                                             (256, 8, 256, 256, 3), (64, 64, 192, 192, 3), (64, 32, 576, 384, 3),
                                             (64, 16, 576, 1728, 1)])
def test_conv_triple_bilinear_identities_at_batch_256(dev, B, H, Cin, Cout, ks):
    from tinyedm_b200 import ops
    from tinyedm_b200.engine import WeightBank, conv_slot
    ops.ensure_device(dev)
    torch.manual_seed(H + Cin)
    p = torch.nn.Parameter(torch.randn(Cout, Cin, ks, ks, device=dev))
    bank = WeightBank([conv_slot("w", p)])
    bank.materialise(dev)
    bank.prepare(False)
    s = bank.slots[0]
    x = torch.randn(B, H, H, Cin, device=dev).to(BF)
    g = torch.randn(B, H, H, Cout, device=dev).to(BF)
    y = ops.conv2d(x, s.fwd, ks, Cout)                       # forward
    gx = ops.conv2d(g, s.dgrad, ks, Cin)                     # data gradient (flipped / transposed operand)
    dw = torch.zeros(Cout, ks * ks, Cin, device=dev)
    ops.conv2d_wgrad(g, x, dw, ks)                           # weight gradient, [Cout][tap][Cin]
    a = (y.double() * g.double()).sum()
    b = (gx.double() * x.double()).sum()
    c = (dw.double() * s.fwd.double().view(Cout, ks * ks, Cin)).sum()
    scale = math.sqrt(float(B * H * H * Cout)) * float(y.float().std()) * float(g.float().std())   # ~ std of the sum
    # y and gx carry one bf16 rounding each (2^-9 relative, random sign): the sums agree to a few ulp x sqrt(N)
    assert abs(a - c) < 0.05 * scale, (float(a), float(c), scale)
    assert abs(b - c) < 0.05 * scale, (float(b), float(c), scale)


def test_full_size_training_step_properties(dev):
    import tinyedm_b200 as T
    from tinyedm_b200.configs import CIFAR10, build_edm
    torch.manual_seed(1)
    model = build_edm(CIFAR10).to(dev).train()
    with torch.no_grad():
        model.denoiser.gain_out.fill_(1.0)
    B = 256
    clean = (0.5 * torch.randn(B, 3, 32, 32, device=dev)).clamp(-1, 1)
    labels = torch.zeros(B, dtype=torch.long, device=dev)
    loss = model.training_step((clean, labels), 0)
    loss.backward()
    assert loss.shape == (1,) and torch.isfinite(loss).all()
    n_checked = 0
    for name, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        if name.endswith("weight") and p.dim() == 4 and p.shape[1] >= 64:
            # after the training-mode forward the stored weight is on the norm sphere and dL/dw is orthogonal to each
            # filter (the weight-norm Jacobian projects the radial component out, up to eps)
            w, g = p.detach().flatten(1).double(), p.grad.flatten(1).double()
            assert rel(w.norm(dim=1), torch.full_like(w[:, 0], math.sqrt(w.shape[1]))) < 1e-3, name
            cos = (w * g).sum(1).abs() / (w.norm(dim=1) * g.norm(dim=1) + 1e-30)
            assert float(cos.max()) < 2e-3, (name, float(cos.max()))
            n_checked += 1
    assert n_checked > 40
    # one optimiser step changes every weight and keeps everything finite
    opt = model.configure_optimizers()["optimizer"]
    before = model.denoiser.encoder_blocks[0].conv_3x3_1.weight.detach().clone()
    opt.step()
    assert not torch.equal(before, model.denoiser.encoder_blocks[0].conv_3x3_1.weight)
    assert all(torch.isfinite(p).all() for p in model.parameters())


def test_full_size_sampler_is_deterministic_and_graph_replay_matches_eager(dev):
    import tinyedm_b200 as T
    from tinyedm_b200.configs import CIFAR10, build_edm
    torch.manual_seed(2)
    model = build_edm(CIFAR10, num_classes=10, dropout_rate=0.0).to(dev).eval()
    with torch.no_grad():
        model.denoiser.gain_out.fill_(1.0)
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(128, 3, 32, 32, generator=g).to(dev)
    labels = torch.randint(0, 10, (128, 1), generator=g).to(dev)
    solver = T.DeterministicSolver(num_steps=32)
    eager = _eager(solver, model, x0, labels)
    a = solver.solve(model, x0, labels)      # captures the 63-evaluation trajectory
    b = solver.solve(model, x0, labels)      # replays it
    assert torch.isfinite(a).all()
    assert torch.equal(a, eager) and torch.equal(b, eager)
    # batch sharding (how the 8 GPUs split the work) does not change an image: the network has no cross-sample op
    half = _eager(solver, model, x0[:64].contiguous(), labels[:64].contiguous())
    assert rel(half, eager[:64]) < 2e-2      # different tile shapes -> different bf16 summation order, not bit-identical


@pytest.mark.parametrize("name,B", [("cifar", 256), ("mnist", 128), ("imagenet", 16)])
def test_full_size_forward_backward_vs_fp32_oracle_on_the_gpu(dev, name, B):
    """BASELINE.json's own training batches (CIFAR 256 x 3 x 32 x 32 on the 35.6 M net, MNIST 128 x 1 x 28 x 28 on the
    87 M net; 16 latents of 4 x 64 x 64 on the 273 M ImageNet net — its micro-batch of 176 would need > 150 GB for the fp32
    autograd graph), element-wise: D, the loss and the gradient of every parameter tensor against the fp32 oracle
    evaluated ON THE GPU with TF32 off (seconds instead of the CPU's minutes). Eval mode / no dropout so that both sides
    see the same weights and masks; the bounds are the accumulated-drift bounds of the small-batch config tests
    (tests/test_gpu_parity.py)."""
    import dataclasses
    import tinyedm_b200 as T
    from oracle import edm2_oracle as O
    from tests.helpers import build_modules, cifar_cfg, seeded_params
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if name == "cifar":
        cfg, img = cifar_cfg(num_classes=None), (3, 32, 32)
    elif name == "mnist":
        cfg, img = dict(O.MNIST), (1, 28, 28)
        cfg["denoiser"] = dataclasses.replace(cfg["denoiser"], dropout_rate=0.0)
    else:
        cfg, img = dict(O.IMAGENET), (4, 64, 64)
    dp, ep, _ = seeded_params(cfg, seed=17)
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.eval(); emb_m.eval()
    g = torch.Generator().manual_seed(6)
    clean = (0.5 * torch.randn(B, *img, generator=g)).clamp(-1, 1).to(dev)
    sigma = torch.exp(torch.randn(B, generator=g) * 1.2 - 1.2).to(dev)
    noisy = clean + torch.randn(B, *img, generator=g).to(dev) * sigma.view(-1, 1, 1, 1)
    ncls = cfg["embedding"].num_classes
    labels = torch.randint(0, ncls, (B,), generator=g).to(dev) if ncls else None
    # oracle, fp32, on the device
    dpo = {k: v.to(dev).requires_grad_(True) for k, v in dp.items()}
    epo = {k: v.to(dev) for k, v in ep.items()}
    _, emb_o = O.embedding_forward(epo, cfg["embedding"], sigma, labels)
    D_o = O.denoiser_forward(dpo, cfg["denoiser"], noisy, sigma, emb_o)
    loss_o = O.training_loss(O.loss_weight(sigma, 0.5), D_o, clean)
    loss_o.backward()
    D_ref, loss_ref = D_o.detach(), loss_o.detach()
    g_ref = {k: v.grad for k, v in dpo.items()}
    del D_o, loss_o
    torch.cuda.empty_cache()
    # this library
    _, e = emb_m(sigma, labels)
    D = den(noisy, sigma, e)
    loss = T.fused_edm_loss(D, clean, sigma, 0.5)
    loss.backward()
    c_skip, c_out, _ = O.precond_coeffs(sigma, 0.5)
    F_net, F_ref = (D.detach() - noisy * c_skip) / c_out, (D_ref - noisy * c_skip) / c_out     # the network branch alone
    r_D, r_F, r_loss = rel(D, D_ref), rel(F_net, F_ref), rel(loss, loss_ref)
    worst = ("", 0.0)
    for k, p in den.named_parameters():
        if p.ndim == 0:
            continue
        r = rel(p.grad, g_ref[k])
        if r > worst[1]:
            worst = (k, r)
    print(f"{name} B = {B} vs fp32 oracle on the GPU: D {r_D:.2e}, network branch {r_F:.2e}, loss {r_loss:.2e}, worst tensor gradient "
          f"{worst[1]:.2e} ({worst[0]})")
    assert r_D < 4e-2 and r_F < 4e-2 and r_loss < 4e-2
    assert worst[1] < 6e-2, worst
