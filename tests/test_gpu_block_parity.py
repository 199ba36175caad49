"""Per-layer criterion of north_star on the REAL architectures (CIFAR-10, MNIST, ImageNet-512 latent), forward AND
backward: every block is fed the oracle's own (bf16-rounded) input and the oracle's own output gradient, and its output,
its input / skip / embedding gradients and the gradient of every one of its parameters must match the fp32 oracle to
relative L2 <= 1e-2 (bf16). Teacher forcing isolates the block: nothing accumulates across layers.

The oracle (oracle/edm2_oracle.py: networks.py:246-265 EncoderBlock, :306-329 DecoderBlock, :191-207 CosineAttention, and
torch autograd for the adjoint) runs here on the GPU in fp32 with TF32 disabled: it is the checker, not the product.
The modulation gradient d m (B, C) — the last per-element quantity before the embed Linear — is compared element-wise
like everything else. The 0-d block gain's gradient is the signed sum sum_{b,c} d m * lin over 10^2..10^3 terms that
largely cancel: it is held to 1e-2 of the sum's UN-CANCELLED magnitude sum |d m * lin| (an error of 1e-2 on every term can
do no worse), and its plain relative error is printed. Every failure of a config is collected and reported together.
"""
import dataclasses

import pytest
import torch

from oracle import edm2_oracle as O
from tests.helpers import build_modules, cifar_cfg, nhwc_to_nchw, rel, seeded_params

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2
BF = torch.bfloat16


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from tinyedm_b200 import ops
    d = torch.device("cuda:0")
    ops.ensure_device(d)
    return d


def _cfg(name):
    if name == "cifar":
        return cifar_cfg(num_classes=10), 2, (3, 32, 32)
    if name == "mnist":
        cfg = dict(O.MNIST)
        cfg["denoiser"] = dataclasses.replace(cfg["denoiser"], dropout_rate=0.0)
        return cfg, 2, (1, 28, 28)
    return dict(O.IMAGENET), 1, (4, 64, 64)


def _bf(t):
    return t.to(BF).float()


@pytest.mark.parametrize("name", ["cifar", "mnist", "imagenet"])
def test_teacher_forced_blocks_forward_and_backward(dev, name):
    from tinyedm_b200 import ops
    cfg, B, img = _cfg(name)
    spec = cfg["denoiser"]
    dp, ep, _ = seeded_params(cfg, seed=21)
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    den.eval(); emb_m.eval()
    g = torch.Generator().manual_seed(5)
    sigma = torch.tensor([0.4, 3.0][:B])
    noisy = ((0.5 * torch.randn(B, *img, generator=g)).clamp(-1, 1) + torch.randn(B, *img, generator=g) * sigma.view(-1, 1, 1, 1)).to(dev)
    labels = torch.randint(0, cfg["embedding"].num_classes, (B,), generator=g).to(dev)
    sigma = sigma.to(dev)
    dpd = {k: v.to(dev) for k, v in dp.items()}
    epd = {k: v.to(dev) for k, v in ep.items()}
    with torch.no_grad():
        _, emb = O.embedding_forward(epd, cfg["embedding"], sigma, labels)
        taps = {}
        O.denoiser_forward(dpd, spec, noisy, sigma, emb, taps=taps)
    enc, dec = O.block_plan(spec)
    eng = den.engine
    eng._ensure_device(dev)
    eng.bank.prepare(False)
    eng.bank.ensure_grad_buffers()
    Be, E = emb.shape
    lin = torch.empty((Be, eng.n_mod), device=dev)
    ops.sgemm(emb, eng.w_embed_all, lin, Be, eng.n_mod, E, E, E, eng.n_mod, False, True)
    mod = ops.mod_finish_forward(lin, eng.gain_ptrs, eng.col_block)
    ctx = dict(mod=mod, lin=lin, emb=emb, drop_p=0.0)
    nb = len(eng.blocks)
    names = ["conv_in"] + [bp.name for bp in eng.blocks]
    skip_taps = [taps["conv_in"]] + [taps[f"encoder_blocks.{i}"] for i in range(eng.n_enc)]
    to_nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(BF)
    gen = torch.Generator(device=dev).manual_seed(77)
    worst: dict[str, tuple] = {}
    failures: list = []
    captured: dict = {}
    orig_modulation = O._modulation

    def modulation_tap(p, pre, e):           # the oracle's m = embed(emb) * gain + 1, kept so that d m can be asked for
        captured["m"] = orig_modulation(p, pre, e)
        return captured["m"]

    def note(kind, where, r, tol=BF16_TOL):
        if r > worst.get(kind, ("", 0.0))[1]:
            worst[kind] = (where, r)
        if not r < tol:
            failures.append((kind, where, f"{r:.3e}"))

    for j, bp in enumerate(eng.blocks):
        pre = bp.name + "."
        blk = enc[j] if bp.kind == "enc" else dec[j - eng.n_enc]
        xin_o = _bf(taps[names[j]]).requires_grad_(True)
        skip_o = _bf(skip_taps[bp.skip_src]).requires_grad_(True) if bp.cskip > 0 else None
        emb_o = emb.clone().requires_grad_(True)
        p_loc = {k: v.clone().requires_grad_(True) for k, v in dpd.items() if k.startswith(pre)}
        O._modulation = modulation_tap
        try:
            if bp.kind == "enc":
                out_o = O.encoder_block(p_loc, pre, blk, xin_o, emb_o, spec)
            else:
                out_o = O.decoder_block(p_loc, pre, blk, xin_o, emb_o, skip_o, spec)
        finally:
            O._modulation = orig_modulation
        g_out = _bf(torch.randn(out_o.shape, device=dev, generator=gen))
        wrt = [xin_o, emb_o, captured["m"]] + ([skip_o] if skip_o is not None else []) + list(p_loc.values())
        grads = torch.autograd.grad(out_o, wrt, g_out)
        g_xin_o, g_emb_o, g_m_o = grads[0], grads[1], grads[2]
        g_skip_o = grads[3] if skip_o is not None else None
        g_par_o = dict(zip(p_loc.keys(), grads[4 if skip_o is not None else 3:]))

        # ---- this library: the same block, same inputs, through the engine's per-block forward and adjoint ----
        with torch.no_grad():
            out, S, _ = eng._block_forward(bp, to_nhwc(xin_o), to_nhwc(skip_o) if skip_o is not None else None, mod,
                                           eng.n_mod, 0.0, True)
            note("output", bp.name, rel(nhwc_to_nchw(out), out_o))
            eng.bank._ghat_flat.zero_()
            sg = torch.zeros(nb + 1, device=dev)
            d_mod = torch.zeros((B, eng.n_mod), device=dev)
            pending: dict = {}
            g_in = eng._block_backward(bp, S, to_nhwc(g_out), ctx, d_mod, pending)
            g_emb = eng._modulation_backward(ctx, d_mod, sg, True)
            eng.bank.backward()
            note("d input", bp.name, rel(nhwc_to_nchw(g_in), g_xin_o))
            note("d modulation", bp.name, rel(d_mod[:, bp.col0:bp.col0 + bp.cout], g_m_o))
            if skip_o is not None:
                g_skip, share, scale = pending[bp.skip_src]
                g_skip = g_skip.float()
                if share is not None:     # the gradient through ScaleLong's spatial mean travels separately (engine.py)
                    g_skip = g_skip + scale * share[:, None, None, :]
                note("d skip", bp.name, rel(g_skip.permute(0, 3, 1, 2), g_skip_o))
            note("d embedding", bp.name, rel(g_emb, g_emb_o))
            for key, s in bp.w.items():
                note("d weight", s.name, rel(s.grad, g_par_o[s.name]))
            # 0-d gain: error against the un-cancelled magnitude of its sum (see the module docstring)
            lin_blk = lin[:, bp.col0:bp.col0 + bp.cout]
            uncancelled = float((g_m_o * lin_blk).abs().sum())
            err = abs(float(sg[j]) - float(g_par_o[pre + "gain"]))
            note("d gain (0-d) / un-cancelled sum", bp.name, err / (uncancelled + 1e-30))
            note("d gain (0-d), plain relative [printed only]", bp.name, rel(sg[j], g_par_o[pre + "gain"]), tol=float("inf"))
    print(f"{name}: worst teacher-forced errors over {nb} blocks: " + ", ".join(f"{k} {v[1]:.2e} ({v[0]})" for k, v in worst.items()))
    assert not failures, (name, f"{len(failures)} quantities above {BF16_TOL}", failures[:20])
