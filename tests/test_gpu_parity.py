"""GPU parity tests (`-m gpu`): the CUDA path vs the oracle / the reference's golden vectors.

Tolerances (north_star): bf16 network kernels — per-layer relative L2 <= 1e-2 against an fp32 oracle fed the SAME
(bf16-rounded) inputs; fp32 kernels (embedding, loss, sampler arithmetic, optimiser) — <= 1e-5; whole-network
accumulated drift of a bf16 network vs the fp32 oracle is reported against the reference's own bf16 noise floor
(1-3e-2), so those end-to-end checks use 4e-2.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import edm2_oracle as O
from tests.helpers import SMALL, build_modules, cifar_cfg, nhwc_to_nchw, rel, seeded_params, small_inputs, small_params

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2
DRIFT_TOL = 4e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from tinyedm_b200 import ops
    d = torch.device("cuda:0")
    ops.ensure_device(d)
    return d


def _conv_ref(x_nhwc, w_oihw):
    """fp32 conv of bf16-rounded operands (what a bf16 tensor-core kernel with fp32 accumulation computes)."""
    xr = x_nhwc.float().permute(0, 3, 1, 2)
    wr = w_oihw.to(torch.bfloat16).float()
    return F.conv2d(xr, wr, padding="same").permute(0, 2, 3, 1).contiguous()


CONV_CASES = [  # B,H,W,Cin,Cout,ks
    (2, 32, 32, 64, 64, 1), (2, 32, 32, 64, 64, 3), (4, 32, 32, 256, 256, 3), (4, 16, 16, 256, 256, 3),
    (5, 8, 8, 256, 256, 3), (4, 16, 16, 256, 768, 1), (3, 8, 8, 512, 256, 3), (3, 8, 8, 512, 256, 1),
    (2, 28, 28, 128, 128, 3), (3, 14, 14, 256, 256, 3), (5, 7, 7, 512, 512, 3), (2, 64, 64, 192, 192, 3),
    (2, 16, 16, 576, 1728, 1), (1, 4, 4, 128, 128, 3), (7, 4, 4, 64, 128, 3),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks", CONV_CASES)
def test_conv_forward(dev, B, H, W, Cin, Cout, ks):
    from tinyedm_b200 import ops
    torch.manual_seed(B * 1000 + H + Cin)
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / math.sqrt(Cin * ks * ks)
    wq = w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    y = ops.conv2d(x, wq, ks, Cout)
    assert rel(y, _conv_ref(x, w)) < 6e-3   # one bf16 output rounding (2^-9 relative ~ 2e-3 rms)


WGRAD_CASES = [(2, 32, 32, 64, 64, 1), (2, 32, 32, 64, 128, 3), (4, 32, 32, 256, 256, 3), (4, 16, 16, 256, 256, 3),
               (6, 8, 8, 256, 256, 3), (3, 8, 8, 512, 256, 3), (4, 16, 16, 256, 768, 1), (2, 28, 28, 128, 128, 3),
               (3, 14, 14, 256, 256, 3), (5, 7, 7, 512, 512, 3), (2, 64, 64, 192, 192, 3), (3, 16, 16, 64, 256, 1),
               (5, 4, 4, 128, 128, 3)]


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks", WGRAD_CASES)
def test_conv_wgrad(dev, B, H, W, Cin, Cout, ks):
    from tinyedm_b200 import ops
    torch.manual_seed(B * 77 + H + Cout)
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    g = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
    dw = torch.full((Cout, ks * ks, Cin), 7.0, device=dev)
    ops.conv2d_wgrad(g, x, dw, ks)
    xr = x.float().permute(0, 3, 1, 2).contiguous()
    w0 = torch.zeros(Cout, Cin, ks, ks, device=dev, requires_grad=True)
    (gw,) = torch.autograd.grad(F.conv2d(xr, w0, padding="same"), w0, g.float().permute(0, 3, 1, 2))
    assert rel(dw, gw.permute(0, 2, 3, 1).reshape(Cout, ks * ks, Cin)) < 1e-4


def test_conv_dgrad_is_the_adjoint(dev):
    """<conv(x, w), g> == <x, dgrad(g, w)> with the flipped/transposed layout emitted by the weight bank."""
    from tinyedm_b200.engine import WeightBank, conv_slot
    from tinyedm_b200 import ops
    torch.manual_seed(5)
    for (Cin, Cout, ks) in [(128, 256, 3), (256, 64, 1)]:
        p = torch.nn.Parameter(torch.randn(Cout, Cin, ks, ks, device=dev))
        bank = WeightBank([conv_slot("w", p)])
        bank.materialise(dev)
        bank.prepare(False)
        s = bank.slots[0]
        w_hat = O.effective_weight(p.detach().cpu())
        assert rel(s.fwd.float().view(Cout, ks, ks, Cin).permute(0, 3, 1, 2), w_hat) < 3e-3
        x = torch.randn(3, 16, 16, Cin, device=dev).to(torch.bfloat16)
        g = torch.randn(3, 16, 16, Cout, device=dev).to(torch.bfloat16)
        gx = ops.conv2d(g, s.dgrad, ks, Cin)
        wq = s.fwd.float().view(Cout, ks, ks, Cin).permute(0, 3, 1, 2).contiguous()
        xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
        (gx_ref,) = torch.autograd.grad(F.conv2d(xr, wq, padding="same"), xr, g.float().permute(0, 3, 1, 2))
        assert rel(gx, gx_ref.permute(0, 2, 3, 1)) < 6e-3


@pytest.mark.parametrize("shapes", [
    [(64, 64, 3, 3), (128, 64, 1, 1), (16, 257, 1, 1), (96, 40)],
    # long rows (ImageNet-latent decoder: 9 x 1536 floats): several fan-in tiles per 16-row group strided over blockIdx.y,
    # the 256-thread backward with > 48 KB of staging, a row count that is not a multiple of 16
    [(32, 1536, 3, 3), (40, 768, 3, 3), (24, 1000), (48, 320, 3, 3)],
])
def test_weight_prep_forward_backward_and_forced_rewrite(dev, shapes):
    from tinyedm_b200.engine import WeightBank, conv_slot, f32_slot
    torch.manual_seed(11)
    params = [torch.nn.Parameter(torch.randn(*s, device=dev) * (0.5 + i)) for i, s in enumerate(shapes)]
    slots = [conv_slot("a", params[0]), conv_slot("b", params[1]), f32_slot("c", params[2]),
             conv_slot("d", params[3]) if len(shapes[3]) == 4 else f32_slot("d", params[3])]
    bank = WeightBank(slots)
    bank.materialise(dev)
    before = [p.detach().cpu().clone() for p in params]
    bank.prepare(False)                                     # eval: parameters untouched
    for p, b in zip(params, before):
        assert torch.equal(p.detach().cpu(), b)
    for s, b in zip(slots, before):
        w_hat = O.effective_weight(b)
        if s.want_f32:
            assert rel(s.f32.view_as(b.reshape(b.shape[0], -1)), w_hat.reshape(b.shape[0], -1)) < 1e-6
        else:
            got = s.fwd.float().view(b.shape[0], b.shape[2], b.shape[3], b.shape[1]).permute(0, 3, 1, 2)
            assert rel(got, w_hat) < 3e-3
            if s.dgrad is not None:   # data-gradient operand: [cin][tap flipped][rows]
                got_d = s.dgrad.float().view(b.shape[1], b.shape[2], b.shape[3], b.shape[0])
                assert rel(got_d, w_hat.flip(2, 3).permute(1, 2, 3, 0)) < 3e-3
    # backward through the normalisation vs autograd of the oracle
    bank.ensure_grad_buffers()
    bank.prepare(False)
    gs = []
    for s, b in zip(slots, before):
        g = torch.randn(s.rows, s.kpad, device=dev)
        s.ghat.copy_(g)
        gs.append(g.cpu())
    bank.backward()
    for s, b, g in zip(slots, before, gs):
        w = b.clone().requires_grad_(True)
        w_hat = O.effective_weight(w)
        if s.taps > 1:   # g_hat is [rows][tap][cin]
            g_oihw = g.view(s.rows, 3, 3, s.cin).permute(0, 3, 1, 2)
        else:
            g_oihw = g[:, :w_hat[0].numel()].reshape(w_hat.shape)
        (gw,) = torch.autograd.grad(w_hat, w, g_oihw)
        assert rel(s.grad, gw) < 1e-5, s.name
    # training: in-place forced re-normalisation (networks.py:32-34)
    v0 = params[0]._version
    bank.prepare(True)
    assert params[0]._version > v0
    for p, b in zip(params, before):
        assert rel(p, O.normalize_weight(b)) < 1e-6


def test_embedding_uncertainty_loss_fp32(dev, golden):
    import tinyedm_b200 as T
    cfg = SMALL
    dp, ep, up = small_params()
    _, emb_m, unc_m = build_modules(cfg, dp, ep, up, dev)
    emb_m.eval(); unc_m.eval()
    sigma = torch.from_numpy(golden["sigma"]).to(dev)
    labels = torch.from_numpy(golden["labels"]).to(dev)
    four, emb = emb_m(sigma, labels)
    assert rel(four, golden["fourier"]) < 1e-5 and rel(emb, golden["embedding"]) < 1e-5
    # 0-d sigma (sampling): fourier has one row, embedding broadcasts over the labels (networks.py:139, :175)
    f0, e0 = emb_m(sigma[0], labels)
    fo, eo = O.embedding_forward(ep, cfg["embedding"], sigma[0].cpu(), labels.cpu())
    assert f0.shape == fo.shape and e0.shape == eo.shape
    assert rel(f0, fo) < 1e-5 and rel(e0, eo) < 1e-5
    with pytest.raises(ValueError):
        T.Embedding(32, 64, None).to(dev)(sigma, labels)
    # loss (+ uncertainty) and every gradient of this fp32 sub-graph vs the oracle
    D = torch.from_numpy(golden["D"]).to(dev).requires_grad_(True)
    clean = torch.from_numpy(golden["clean"]).to(dev)
    four, emb = emb_m(sigma, labels)
    u = unc_m(four).flatten()
    loss = T.fused_edm_loss(D, clean, sigma, 0.5, u)
    assert rel(loss, golden["loss_unc"]) < 1e-5
    assert rel(T.fused_edm_loss(D, clean, sigma, 0.5), golden["loss_plain"]) < 1e-5
    loss.backward()
    upo = {k: v.clone().requires_grad_(True) for k, v in up.items()}
    Do = torch.from_numpy(golden["D"]).requires_grad_(True)
    lo = O.training_loss(O.loss_weight(torch.from_numpy(golden["sigma"]), 0.5), Do, torch.from_numpy(golden["clean"]),
                         O.uncertainty_forward(upo, torch.from_numpy(golden["fourier"])))
    lo.backward()
    assert rel(D.grad, Do.grad) < 1e-5
    assert rel(unc_m.linear1.weight.grad, upo["linear1.weight"].grad) < 1e-4
    assert rel(unc_m.linear2.weight.grad, upo["linear2.weight"].grad) < 1e-4
    assert rel(unc_m.gain.grad, upo["gain"].grad) < 1e-4


def test_weighted_mse_metric_identity(dev):
    """The reference's own numerical test (tests/test_weighted_mean_squared_error.py:6-21) on the CUDA path."""
    import tinyedm_b200 as T
    metric = T.WeightedMeanSquaredError()
    torch.manual_seed(0)
    w = torch.rand(8, device=dev)
    preds = torch.randn(8, 3, 32, 32, device=dev)
    target = torch.randn(8, 3, 32, 32, device=dev)
    val = metric(w, preds, target)
    expect = torch.mean(w[:, None, None, None] * (preds - target) ** 2)
    assert val.shape == (1,) and torch.allclose(val.squeeze(), expect, rtol=1e-5)
    assert torch.allclose(metric.compute().squeeze(), expect, rtol=1e-5) and int(metric.total) == 8
    metric(w, preds, target)
    assert int(metric.total) == 16 and torch.allclose(metric.compute().squeeze(), expect, rtol=1e-5)
    with pytest.raises(RuntimeError):
        metric(w[:4], preds, target)


def _run_denoiser(den, emb_m, noisy, sigma, labels, taps=None):
    _, e = emb_m(sigma, labels)
    D, _ = den.engine.forward(noisy.float().contiguous(), sigma.reshape(-1).float().contiguous(), e.detach(),
                              training=False, save=False, taps=taps)
    return D


def test_small_denoiser_forward_vs_reference_golden(dev, golden):
    cfg = SMALL
    dp, ep, _ = small_params()
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.eval(); emb_m.eval()
    noisy = torch.from_numpy(golden["noisy"]).to(dev)
    sigma = torch.from_numpy(golden["sigma"]).to(dev)
    labels = torch.from_numpy(golden["labels"]).to(dev)
    taps = {}
    with torch.no_grad():
        D = _run_denoiser(den, emb_m, noisy, sigma, labels, taps)
    errs = {k[4:]: rel(nhwc_to_nchw(taps[k[4:]]), golden[k].astype(np.float32)) for k in golden.files if k.startswith("tap/")}
    print("block drift vs fp32 reference:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["conv_in"] < BF16_TOL
    assert max(errs.values()) < DRIFT_TOL
    assert rel(D, golden["D"]) < DRIFT_TOL
    # module path == engine path, and eval leaves the parameters untouched
    with torch.no_grad():
        _, e = emb_m(sigma, labels)
        D2 = den(noisy, sigma, e)
    assert torch.equal(D, D2)
    assert rel(den.conv_in.weight, dp["conv_in.weight"]) == 0


def test_forward_hooks_on_blocks_fire_with_reference_arguments(dev, golden):
    """activation taps (`register_forward_hook` on EncoderBlock / DecoderBlock — what make_golden.py and
    wandb-style watchers use) keep working although the blocks run fused: (input, embedding[, skip]) in, block output out,
    NCHW, equal to the reference's own taps within the accumulated bf16 drift."""
    cfg = SMALL
    dp, ep, _ = small_params()
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.eval(); emb_m.eval()
    seen = {}
    def tap(name):
        def hook(mod, args, out):
            seen[name] = (tuple(a.shape for a in args), out.float().clone())
        return hook
    handles = [den.encoder_blocks[1].register_forward_hook(tap("encoder_blocks.1")),
               den.decoder_blocks[2].register_forward_hook(tap("decoder_blocks.2")),
               den.decoder_blocks[3].register_forward_hook(tap("decoder_blocks.3"))]
    noisy = torch.from_numpy(golden["noisy"]).to(dev)
    sigma = torch.from_numpy(golden["sigma"]).to(dev)
    labels = torch.from_numpy(golden["labels"]).to(dev)
    with torch.no_grad():
        _, e = emb_m(sigma, labels)
        den(noisy, sigma, e)
    assert set(seen) == {"encoder_blocks.1", "decoder_blocks.2", "decoder_blocks.3"}
    for name, (shapes, out) in seen.items():
        ref = golden["tap/" + name].astype(np.float32)
        assert out.shape == ref.shape and rel(out, ref) < DRIFT_TOL, name
        assert shapes[1] == tuple(e.shape)
    assert len(seen["encoder_blocks.1"][0]) == 2 and len(seen["decoder_blocks.2"][0]) == 3     # the skip block gets (x, emb, skip)
    assert len(seen["decoder_blocks.3"][0]) == 2
    for h in handles:
        h.remove()
    seen.clear()
    with torch.no_grad():
        den(noisy, sigma, e)
    assert not seen


def test_small_denoiser_teacher_forced_blocks(dev, golden):
    """Per-layer criterion: every block fed the ORACLE's (bf16-rounded) input must match the oracle's output <= 1e-2."""
    cfg = SMALL
    dp, ep, _ = small_params()
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.eval(); emb_m.eval()
    clean, eps, noise, labels, _ = small_inputs()
    noisy, sigma = O.diffuse(clean, eps, noise, -1.2, 1.2)
    _, emb = O.embedding_forward(ep, cfg["embedding"], sigma, labels)
    taps = {}
    O.denoiser_forward(dp, cfg["denoiser"], noisy, sigma, emb, taps=taps)
    eng = den.engine
    with torch.no_grad():
        eng._ensure_device(dev)
        eng.bank.prepare(False)
        e = emb.to(dev)
        from tinyedm_b200 import ops
        lin = torch.empty((e.shape[0], eng.n_mod), device=dev)
        ops.sgemm(e, eng.w_embed_all, lin, e.shape[0], eng.n_mod, e.shape[1], e.shape[1], e.shape[1], eng.n_mod, False, True)
        mod = ops.mod_finish_forward(lin, eng.gain_ptrs, eng.col_block)
        to_dev = lambda t: t.permute(0, 2, 3, 1).contiguous().to(dev).to(torch.bfloat16)
        names = ["conv_in"] + [bp.name for bp in eng.blocks]
        skips = [taps["conv_in"]] + [taps[f"encoder_blocks.{i}"] for i in range(eng.n_enc)]
        worst = 0.0
        for j, bp in enumerate(eng.blocks):
            xin = to_dev(taps[names[j]])
            skip = to_dev(skips[bp.skip_src]) if bp.cskip > 0 else None
            out, _, _ = eng._block_forward(bp, xin, skip, mod, eng.n_mod, 0.0, False)
            r = rel(nhwc_to_nchw(out), taps[bp.name])
            worst = max(worst, r)
            assert r < BF16_TOL, (bp.name, r)
    print(f"worst teacher-forced block error {worst:.2e}")


def test_small_denoiser_backward_vs_reference_golden(dev, golden):
    import tinyedm_b200 as T
    cfg = SMALL
    dp, ep, up = small_params()
    den, emb_m, unc_m = build_modules(cfg, dp, ep, up, dev)
    den.eval(); emb_m.eval(); unc_m.eval()      # eval: no in-place rewrite, dropout 0 — the fixture's setting
    noisy = torch.from_numpy(golden["noisy"]).to(dev)
    sigma = torch.from_numpy(golden["sigma"]).to(dev)
    labels = torch.from_numpy(golden["labels"]).to(dev)
    clean = torch.from_numpy(golden["clean"]).to(dev)
    four, e = emb_m(sigma, labels)
    D = den(noisy, sigma, e)
    loss = T.fused_edm_loss(D, clean, sigma, 0.5, unc_m(four).flatten())
    assert rel(loss, golden["loss_unc"]) < DRIFT_TOL
    loss.backward()
    named = {**{f"denoiser.{k}": v for k, v in den.named_parameters()},
             **{f"embedding.{k}": v for k, v in emb_m.named_parameters()}, **{f"u.{k}": v for k, v in unc_m.named_parameters()}}
    errs = {}
    for k in golden.files:
        if k.startswith("grad/"):
            assert named[k[5:]].grad is not None, k
            errs[k[5:]] = rel(named[k[5:]].grad, golden[k])
    print("gradient error vs fp32 reference:", {k: f"{v:.2e}" for k, v in errs.items()})
    norm_errs = {}
    for k in golden.files:
        if k.startswith("gradnorm/"):
            g = named[k[9:]].grad
            assert g is not None, k
            norm_errs[k[9:]] = abs(float(g.norm()) - float(golden[k])) / (float(golden[k]) + 1e-12)
    print("worst gradient-norm error:", max(norm_errs.items(), key=lambda kv: kv[1]))
    # tensors: accumulated bf16 noise of a full forward+backward; 0-d gains are signed sums with heavy cancellation
    scalar = lambda k: named[k].ndim == 0
    assert max(v for k, v in errs.items() if not scalar(k)) < DRIFT_TOL, max(errs.items(), key=lambda kv: kv[1])
    assert max(v for k, v in norm_errs.items() if not scalar(k)) < DRIFT_TOL
    assert max(v for k, v in norm_errs.items() if scalar(k)) < 0.15


def test_cifar_config_forward_vs_oracle(dev):
    """The headline architecture (cifar10.yaml:39-43, 35.6 M parameters), class-conditional, batch 3, vs the fp32 oracle."""
    cfg = cifar_cfg(num_classes=10)
    dp, ep, _ = seeded_params(cfg, seed=3)
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.eval(); emb_m.eval()
    assert abs(sum(p.numel() for p in den.parameters()) - 35.6e6) < 0.1e6     # README.md:64-69 "35.6M"
    g = torch.Generator().manual_seed(9)
    noisy = torch.randn(3, 3, 32, 32, generator=g) * 1.5
    sigma = torch.tensor([0.05, 0.7, 12.0])
    labels = torch.randint(0, 10, (3,), generator=g)
    _, emb = O.embedding_forward(ep, cfg["embedding"], sigma, labels)
    taps_o = {}
    with torch.no_grad():
        D_o = O.denoiser_forward(dp, cfg["denoiser"], noisy, sigma, emb, taps=taps_o)
        taps = {}
        D = _run_denoiser(den, emb_m, noisy.to(dev), sigma.to(dev), labels.to(dev), taps)
    errs = {k: rel(nhwc_to_nchw(v), taps_o[k]) for k, v in taps.items()}
    print("CIFAR block drift:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < DRIFT_TOL
    c_skip, c_out, _ = O.precond_coeffs(sigma, 0.5)
    F_net = (D.cpu() - noisy * c_skip) / c_out            # network branch, the part that is not c_skip * x
    F_ref = (D_o - noisy * c_skip) / c_out
    assert rel(F_net, F_ref) < DRIFT_TOL
    # gain_out == 0 at init hides the network exactly
    with torch.no_grad():
        den.gain_out.zero_()
        D0 = _run_denoiser(den, emb_m, noisy.to(dev), sigma.to(dev), labels.to(dev))
    assert rel(D0, noisy * c_skip) < 1e-6


def test_sampler_arithmetic_fp32_with_shared_denoiser(dev):
    """fp32 sampler trajectory <= 1e-4 over 32 steps (north_star), isolating the solver arithmetic with an analytic
    denoiser shared by both sides (a bf16 network cannot meet 1e-4 over 63 evaluations)."""
    import tinyedm_b200 as T
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(16, 3, 32, 32, generator=g)
    mu = torch.randn(1, 3, 32, 32, generator=g) * 0.5

    def ideal(x, s, lab=None):     # posterior mean of N(mu, 0.5^2 I) data under noise level s
        s = s.to(x.device)
        return (x * 0.25 + mu.to(x.device) * s ** 2) / (0.25 + s ** 2)

    calls = []
    def ideal_counted(x, s, lab=None):
        calls.append(1)
        return ideal(x, s, lab)
    solver = T.DeterministicSolver(num_steps=32)
    out = solver.solve(ideal_counted, x0.to(dev))
    assert len(calls) == 63
    ref = O.heun_solve(ideal, x0, None, num_steps=32)
    assert rel(out, ref) < 1e-4
    np.testing.assert_allclose(solver.t_steps.numpy(), O.t_schedule(32).numpy(), rtol=0, atol=0)
    with pytest.raises(ValueError):
        T.DeterministicSolver(num_steps=8, dtype="float16")


def test_small_sampler_with_network_vs_reference_golden(dev, golden):
    import tinyedm_b200 as T
    cfg = SMALL
    dp, ep, _ = small_params()
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    model = T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb_m, denoiser=den, use_ema=False, use_uncertainty=False,
                  steady_steps=1, rampup_steps=1, scheduler_interval="step").eval()
    x0 = torch.from_numpy(golden["x0"]).to(dev)
    labels = torch.from_numpy(golden["labels"]).to(dev)
    out = T.DeterministicSolver(num_steps=int(golden["sampler_steps"])).solve(model, x0, labels)
    r = rel(out, golden["sampler_out"])
    print(f"6-step sampler with the bf16 network vs fp32 reference: {r:.2e}")
    assert r < DRIFT_TOL


def test_dropout_statistics_and_train_mode_side_effects(dev):
    cfg = cifar_cfg(dropout=0.13)
    dp, ep, _ = seeded_params(cfg, seed=5)
    # un-normalised weights: training forward must project them (networks.py:32-34)
    dp = {k: (v * 1.7 if k.endswith("weight") else v) for k, v in dp.items()}
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.train(); emb_m.eval()
    B = 4
    noisy = torch.randn(B, 3, 32, 32, device=dev)
    sigma = torch.rand(B, device=dev) + 0.3
    _, e = emb_m(sigma)
    D1, ctx = den.engine.forward(noisy, sigma, e.detach(), training=True, save=True)
    h = ctx["blocks"][0]["h"]
    frac = float((h == 0).float().mean())
    assert abs(frac - 0.13) < 0.01, frac
    w = den.encoder_blocks[0].conv_3x3_1.weight.detach().cpu()
    assert rel(w, O.normalize_weight(dp["encoder_blocks.0.conv_3x3_1.weight"])) < 1e-5
    D2, ctx2 = den.engine.forward(noisy, sigma, e.detach(), training=True, save=True)
    assert not torch.equal(ctx2["blocks"][0]["h"] == 0, h == 0), "dropout mask must change between steps"


def test_fused_adam_ema_vs_torch(dev):
    import tinyedm_b200 as T
    torch.manual_seed(0)
    shapes = [(256, 256, 3, 3), (257,), (), (16, 257, 1, 1), (100003,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = T.FusedAdamEMA(ps, lr=0.02, betas=(0.9, 0.999), ema_length=0.13)
    ref = torch.optim.Adam(qs, lr=0.02, betas=(0.9, 0.999))
    gamma = T.sigma_rel_to_gamma(0.13)
    ema = [q.detach().clone() for q in qs]
    for step in range(4):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad = g.clone(); q.grad = g.clone()
        opt.step(); ref.step()
        decay = (1 - 1 / (step + 1)) ** (gamma + 1)
        for e, q in zip(ema, qs):
            e.mul_(decay).add_(q.detach(), alpha=1 - decay)
    for p, q in zip(ps, qs):
        assert rel(p, q) < 5e-6    # powf-based bias correction in fp32 vs torch's double
    for e, e2 in zip(ema, opt.ema_params):
        assert rel(e2, e) < 5e-6


def test_unsupported_inputs_fail_loudly(dev):
    import tinyedm_b200 as T
    from tinyedm_b200 import ops
    with pytest.raises(RuntimeError):
        ops.conv2d(torch.zeros(1, 8, 8, 48, device=dev, dtype=torch.bfloat16), torch.zeros(64, 48, device=dev, dtype=torch.bfloat16), 1, 64)
    with pytest.raises(RuntimeError):
        T.Denoiser(encoder_block_types=("Enc",), decoder_block_types=("Dec", "Dec"), encoder_out_channels=(48,),
                   decoder_out_channels=(48, 48), skip_connections=(True, True), embedding_dim=64).engine
    cfg = SMALL
    dp, ep, _ = small_params()
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    with pytest.raises(RuntimeError):
        den(torch.zeros(2, 3, 16, 16), torch.ones(2), torch.zeros(2, 64))     # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        den(torch.zeros(2, 5, 16, 16, device=dev), torch.ones(2, device=dev), torch.zeros(2, 64, device=dev))


def test_mnist_config_forward_backward_vs_oracle(dev):
    """BASELINE configs[0]: the MNIST arch (mnist.yaml:39-43, 87.2 M parameters; 28/14/7 feature maps, channel changes with
    conv_1x1 in the encoder, head_dim 64 at S=196 and 128 at S=49), class-conditional, batch 2, forward and backward."""
    import dataclasses
    import tinyedm_b200 as T
    cfg = dict(O.MNIST)
    cfg["denoiser"] = dataclasses.replace(cfg["denoiser"], dropout_rate=0.0)
    dp, ep, _ = seeded_params(cfg, seed=7)
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.eval(); emb_m.eval()
    assert sum(p.numel() for p in den.parameters()) + sum(p.numel() for p in emb_m.parameters()) == 87_194_724
    g = torch.Generator().manual_seed(2)
    clean = (0.5 * torch.randn(2, 1, 28, 28, generator=g)).clamp(-1, 1)
    noisy = clean + torch.randn(2, 1, 28, 28, generator=g) * torch.tensor([0.3, 2.0]).view(2, 1, 1, 1)
    sigma = torch.tensor([0.3, 2.0])
    labels = torch.tensor([3, 8])
    dpo = {k: v.clone().requires_grad_(True) for k, v in dp.items()}
    _, emb = O.embedding_forward(ep, cfg["embedding"], sigma, labels)
    taps_o = {}
    D_o = O.denoiser_forward(dpo, cfg["denoiser"], noisy, sigma, emb, taps=taps_o)
    loss_o = O.training_loss(O.loss_weight(sigma, 0.5), D_o, clean)
    loss_o.backward()
    _, e = emb_m(sigma.to(dev), labels.to(dev))
    D = den(noisy.to(dev), sigma.to(dev), e)
    loss = T.fused_edm_loss(D, clean.to(dev), sigma.to(dev), 0.5)
    loss.backward()
    assert rel(D, D_o) < DRIFT_TOL and rel(loss, loss_o) < DRIFT_TOL
    worst = ("", 0.0)
    for k, p in den.named_parameters():
        if p.ndim == 0:
            continue
        r = rel(p.grad, dpo[k].grad)
        if r > worst[1]:
            worst = (k, r)
    print(f"MNIST config: D rel {rel(D, D_o):.2e}, worst tensor-gradient error {worst[1]:.2e} ({worst[0]})")
    assert worst[1] < 6e-2, worst


def test_imagenet_latent_config_forward_backward_vs_oracle(dev):
    """BASELINE configs[3]/[4]: the ImageNet-512 latent architecture (Denoiser defaults, 272 M parameters; 64/32/16/8 maps,
    C = 192..768, head_dim 144 at S=256 and 192 at S=64, concat widths up to 1536), batch 1, forward and backward."""
    import tinyedm_b200 as T
    cfg = dict(O.IMAGENET)
    dp, ep, _ = seeded_params(cfg, seed=11)
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.eval(); emb_m.eval()
    g = torch.Generator().manual_seed(4)
    clean = (0.5 * torch.randn(1, 4, 64, 64, generator=g)).clamp(-1, 1)
    sigma = torch.tensor([0.8])
    noisy = clean + torch.randn(1, 4, 64, 64, generator=g) * 0.8
    labels = torch.tensor([417])
    dpo = {k: v.clone().requires_grad_(True) for k, v in dp.items()}
    _, emb = O.embedding_forward(ep, cfg["embedding"], sigma, labels)
    D_o = O.denoiser_forward(dpo, cfg["denoiser"], noisy, sigma, emb)
    loss_o = O.training_loss(O.loss_weight(sigma, 0.5), D_o, clean)
    loss_o.backward()
    _, e = emb_m(sigma.to(dev), labels.to(dev))
    D = den(noisy.to(dev), sigma.to(dev), e)
    loss = T.fused_edm_loss(D, clean.to(dev), sigma.to(dev), 0.5)
    loss.backward()
    assert rel(D, D_o) < DRIFT_TOL and rel(loss, loss_o) < DRIFT_TOL
    worst = ("", 0.0)
    for k, p in den.named_parameters():
        if p.ndim == 0:
            continue
        r = rel(p.grad, dpo[k].grad)
        if r > worst[1]:
            worst = (k, r)
    print(f"ImageNet-latent config: D rel {rel(D, D_o):.2e}, worst tensor-gradient error {worst[1]:.2e} ({worst[0]})")
    assert worst[1] < 6e-2, worst


def test_sample_postprocessing_uint8_bit_exact(dev):
    """denormalise -> clamp -> NHWC -> uint8 on the device, byte-identical to callbacks.py:152-154."""
    import tinyedm_b200 as T
    g = torch.Generator().manual_seed(11)
    for (B, C, H, W, mean, std) in [(16, 3, 32, 32, (0.4914, 0.4822, 0.4465), (0.247, 0.243, 0.261)), (5, 1, 28, 28, (0.1307,), (0.3081,)),
                                    (3, 4, 64, 64, (0.1, 0.2, 0.3, 0.4), (0.5, 0.6, 0.7, 0.8))]:
        x = torch.randn(B, C, H, W, generator=g) * 1.5
        x.view(-1)[:8] = torch.tensor([0.0, 1e9, -1e9, 2.0, -2.0, 0.5, -0.5, 1.0])
        got = T.to_uint8_images(x.to(dev), torch.tensor(mean), torch.tensor(std))
        want = O.to_uint8_images(x, mean, std)
        assert got.dtype == torch.uint8 and got.shape == (B, H, W, C)
        assert torch.equal(got.cpu(), want)
