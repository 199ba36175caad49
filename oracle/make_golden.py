"""Pins oracle/edm2_oracle.py against the REAL reference and writes tests/golden/*.npz — TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference exists):   python oracle/make_golden.py

It imports /root/reference/src/tinyedm/{networks,solvers}.py by file path (they need only torch + numpy;
`import tinyedm` itself needs lightning/torchmetrics/hydra which are absent), builds the reference modules
for a small architecture that exercises every block flavour (plain / down / up / attention / skip+ScaleLong /
conv_1x1 / class embedding / uncertainty head), loads the oracle's seeded parameters into them, runs BOTH on
the same inputs and asserts agreement before anything is stored. The stored vectors are the REFERENCE's
outputs (fp32, CPU). The GPU box has no /root/reference: tests only read the committed .npz files.

The training-step glue that needs absent packages (edm.py:205-219, metric.py:8-18) is restated inline here
and cross-checked against the identity the reference's own test pins
(tests/test_weighted_mean_squared_error.py:18-21).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import edm2_oracle as O  # noqa: E402

REF = os.environ.get("TINYEDM_REFERENCE", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_ref(name: str):
    spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF, "src", "tinyedm", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


SMALL = dict(
    denoiser=O.DenoiserSpec(
        in_channels=3, out_channels=3, sigma_data=0.5, embedding_dim=64, num_heads=4,
        encoder_block_types=("Enc", "EncD", "EncA"),
        decoder_block_types=("DecA", "Dec", "DecA", "DecU", "Dec", "Dec"),
        encoder_out_channels=(64, 128, 128), decoder_out_channels=(128, 128, 128, 64, 64, 64),
        skip_connections=(False, True, True, False, True, True), dropout_rate=0.0),
    embedding=O.EmbeddingSpec(fourier_dim=32, embedding_dim=64, num_classes=5, add_factor=0.5),
    image=(3, 16, 16), batch=3, seed=1234)


def small_params(cfg=SMALL, gain_out=1.0, u_gain=1.0):
    """Seeded parameters (shared by this script and the tests; checksums are stored in the golden file)."""
    gen = torch.Generator().manual_seed(cfg["seed"])
    dp = O.init_denoiser_params(cfg["denoiser"], gen, gain_out=gain_out)
    ep = O.init_embedding_params(cfg["embedding"], gen)
    up = O.init_uncertainty_params(cfg["embedding"].fourier_dim, gen, gain=u_gain)
    # non-trivial block gains so that d(gain) is exercised
    for k in dp:
        if k.endswith(".gain"):
            dp[k] = torch.tensor(0.5 + torch.rand((), generator=gen).item())
    O.project_weights_(dp)
    O.project_weights_(ep)
    O.project_weights_(up)
    return dp, ep, up


def small_inputs(cfg=SMALL):
    gen = torch.Generator().manual_seed(cfg["seed"] + 1)
    B = cfg["batch"]
    clean = (0.5 * torch.randn(B, *cfg["image"], generator=gen)).clamp(-1, 1)
    eps = torch.randn(B, generator=gen)
    noise = torch.randn(B, *cfg["image"], generator=gen)
    labels = torch.randint(0, cfg["embedding"].num_classes, (B,), generator=gen)
    x0 = torch.randn(B, *cfg["image"], generator=gen)
    return clean, eps, noise, labels, x0


def checksum(params: dict) -> np.ndarray:
    return np.array([float(sum(v.double().sum() for v in params.values())),
                     float(sum(v.double().abs().sum() for v in params.values()))])


def build_reference(net, cfg, dp, ep, up):
    s = cfg["denoiser"]
    den = net.Denoiser(in_channels=s.in_channels, out_channels=s.out_channels,
                       encoder_block_types=s.encoder_block_types, decoder_block_types=s.decoder_block_types,
                       encoder_out_channels=s.encoder_out_channels, decoder_out_channels=s.decoder_out_channels,
                       skip_connections=s.skip_connections, dropout_rate=s.dropout_rate, sigma_data=s.sigma_data,
                       encoder_add_factor=s.encoder_add_factor, decoder_add_factor=s.decoder_add_factor,
                       embedding_dim=s.embedding_dim, num_heads=s.num_heads)
    e = cfg["embedding"]
    emb = net.Embedding(e.fourier_dim, e.embedding_dim, e.num_classes, e.add_factor)
    unc = net.UncertaintyNet(e.fourier_dim, e.fourier_dim)
    den.load_state_dict({k: v.clone() for k, v in dp.items()}, strict=True)
    emb.load_state_dict({k: v.clone() for k, v in ep.items()}, strict=True)
    unc.load_state_dict({k: v.clone() for k, v in up.items()}, strict=True)
    return den, emb, unc


def ref_training_loss(den, emb, unc, clean, eps, noise, labels, P_mean, P_std, sigma_data):
    """edm.py:205-219 with Diffuser's two random draws supplied (edm.py:84-93) and metric.py:8-18 inline."""
    sigma = (P_mean + eps * P_std).exp()
    noisy = clean + noise * sigma.view(-1, 1, 1, 1)
    four, e = emb(sigma, labels)
    D = den(noisy, sigma, e)
    weight = (sigma ** 2 + sigma_data ** 2) / (sigma * sigma_data) ** 2
    N = clean.shape[0]
    out = {}
    for name, u in (("plain", None), ("unc", unc(four).flatten())):
        w = weight if u is None else weight / u.exp()
        diff = D.view(N, -1) - clean.view(N, -1)
        wsse = torch.mean(w.view(N, 1) * diff * diff, dim=1).sum()
        loss = (wsse / N).reshape(1)
        if u is not None:
            loss = loss + u.mean()
        out[name] = loss
    return noisy, sigma, four, e, D, out


def maxrel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    net = load_ref("networks")
    solvers = load_ref("solvers")
    os.makedirs(GOLDEN, exist_ok=True)
    cfg = SMALL
    dp, ep, up = small_params(cfg)
    clean, eps, noise, labels, x0 = small_inputs(cfg)
    den, emb, unc = build_reference(net, cfg, dp, ep, up)
    den.eval(); emb.eval(); unc.eval()
    P_mean, P_std, sd = -1.2, 1.2, cfg["denoiser"].sigma_data

    # ---------------- reference forward / backward (eval mode: no in-place forced WN, dropout 0) -------------
    noisy, sigma, four_r, emb_r, D_r, loss_r = ref_training_loss(den, emb, unc, clean, eps, noise, labels, P_mean, P_std, sd)
    grads_r = {}
    all_params = [(f"denoiser.{k}", v) for k, v in den.named_parameters()] + \
                 [(f"embedding.{k}", v) for k, v in emb.named_parameters()] + \
                 [(f"u.{k}", v) for k, v in unc.named_parameters()]
    g = torch.autograd.grad(loss_r["unc"], [v for _, v in all_params], retain_graph=True)
    for (k, _), gv in zip(all_params, g):
        grads_r[k] = gv.detach()
    taps_r = {}
    hooks = []
    for name, m in list(den.encoder_blocks.named_children()):
        hooks.append(m.register_forward_hook(lambda mod, i, o, n=f"encoder_blocks.{name}": taps_r.__setitem__(n, o.detach())))
    for name, m in list(den.decoder_blocks.named_children()):
        hooks.append(m.register_forward_hook(lambda mod, i, o, n=f"decoder_blocks.{name}": taps_r.__setitem__(n, o.detach())))
    hooks.append(den.conv_in.register_forward_hook(lambda mod, i, o: taps_r.__setitem__("conv_in", o.detach())))
    with torch.no_grad():
        den(noisy, sigma, emb_r)
    for h in hooks:
        h.remove()

    # ---------------- oracle on the same inputs ----------------
    dpo = {k: v.clone().requires_grad_(True) for k, v in dp.items()}
    epo = {k: (v.clone().requires_grad_(True) if k.endswith("weight") else v.clone()) for k, v in ep.items()}
    upo = {k: v.clone().requires_grad_(True) for k, v in up.items()}
    noisy_o, sigma_o = O.diffuse(clean, eps, noise, P_mean, P_std)
    four_o, emb_o = O.embedding_forward(epo, cfg["embedding"], sigma_o, labels)
    taps_o = {}
    D_o = O.denoiser_forward(dpo, cfg["denoiser"], noisy_o, sigma_o, emb_o, taps=taps_o)
    w_o = O.loss_weight(sigma_o, sd)
    loss_plain_o = O.training_loss(w_o, D_o, clean)
    loss_unc_o = O.training_loss(w_o, D_o, clean, O.uncertainty_forward(upo, four_o))
    names_o = [(f"denoiser.{k}", v) for k, v in dpo.items()] + \
              [(f"embedding.{k}", v) for k, v in epo.items() if v.requires_grad] + [(f"u.{k}", v) for k, v in upo.items()]
    g_o = torch.autograd.grad(loss_unc_o, [v for _, v in names_o])
    grads_o = {k: gv for (k, _), gv in zip(names_o, g_o)}

    report = []
    def check(name, a, b, tol):
        r = maxrel(a, b)
        report.append((name, r))
        assert r <= tol, f"oracle disagrees with the reference on {name}: rel L2 {r:.3e} > {tol}"
    check("noisy", noisy_o, noisy, 0)
    check("fourier", four_o, four_r, 1e-6)
    check("embedding", emb_o, emb_r, 1e-6)
    check("D", D_o, D_r, 1e-5)
    check("loss_plain", loss_plain_o, loss_r["plain"], 1e-6)
    check("loss_unc", loss_unc_o, loss_r["unc"], 1e-6)
    for k, v in taps_r.items():
        check(f"tap {k}", taps_o[k], v, 1e-5)
    assert set(grads_o) == set(grads_r), (set(grads_o) ^ set(grads_r))
    for k, v in grads_r.items():
        check(f"grad {k}", grads_o[k], v, 2e-4)

    # training-mode side effect: forced weight normalisation rewrites the parameter (networks.py:32-34)
    conv = net.Conv2d(8, 4, 3)
    w_before = conv.weight.detach().clone()
    conv.train()
    conv(torch.randn(1, 8, 5, 5))
    w_forced = w_before.clone()
    O.forced_weight_norm_(w_forced)
    check("forced_wn", w_forced, conv.weight.detach(), 1e-7)

    # metric identity the reference's own test pins (tests/test_weighted_mean_squared_error.py:18-21)
    gen = torch.Generator().manual_seed(7)
    wt, pr, tg = torch.rand(8, generator=gen), torch.randn(8, 3, 32, 32, generator=gen), torch.randn(8, 3, 32, 32, generator=gen)
    ident = torch.mean(wt[:, None, None, None] * (pr - tg) ** 2)
    check("wmse identity", O.weighted_mse(wt, pr, tg).squeeze(), ident, 1e-6)

    # ---------------- sampler: reference DeterministicSolver vs oracle heun_solve ----------------
    class RefEDM(torch.nn.Module):  # edm.py:280-286
        def forward(self, x, s, lab=None):
            _, e = emb(s, lab)
            return den(x, s, e)
    steps = 6
    solver = solvers.DeterministicSolver(num_steps=steps)
    with torch.no_grad():
        xs_r = solver.solve(RefEDM(), x0, labels)
        traj = []
        model_o = lambda x, s, lab: O.edm_forward(dp, cfg["denoiser"], ep, cfg["embedding"], x, s, lab)
        xs_o = O.heun_solve(model_o, x0, labels, num_steps=steps, trajectory=traj)
    check("t_steps", O.t_schedule(steps), solver.t_steps, 0)
    check("sampler", xs_o, xs_r, 1e-5)
    s32 = solvers.DeterministicSolver(num_steps=32)
    check("t_steps32", O.t_schedule(32), s32.t_steps, 0)

    for n, r in report:
        print(f"  {n:55s} rel L2 = {r:.3e}")

    # ---------------- store ----------------
    sel_grads = ["denoiser.gain_out", "denoiser.conv_in.weight", "denoiser.conv_out.weight",
                 "denoiser.encoder_blocks.0.conv_3x3_1.weight", "denoiser.encoder_blocks.1.conv_1x1.weight",
                 "denoiser.encoder_blocks.2.attention.qkv_conv.weight", "denoiser.encoder_blocks.2.embed.weight",
                 "denoiser.encoder_blocks.2.gain", "denoiser.decoder_blocks.1.cat_factor.layer1.weight",
                 "denoiser.decoder_blocks.1.cat_factor.layer2.weight", "denoiser.decoder_blocks.1.conv_1x1.weight",
                 "denoiser.decoder_blocks.4.conv_3x3_1.weight", "denoiser.decoder_blocks.0.attention.out_conv.weight",
                 "embedding.sigma_embed.weight", "embedding.class_embed.linear.weight",
                 "u.linear1.weight", "u.linear2.weight", "u.gain"]
    out = {
        "param_checksum": np.concatenate([checksum(dp), checksum(ep), checksum(up)]),
        "clean": clean.numpy(), "eps": eps.numpy(), "noise": noise.numpy(), "labels": labels.numpy(), "x0": x0.numpy(),
        "noisy": noisy.numpy(), "sigma": sigma.numpy(), "fourier": four_r.detach().numpy(), "embedding": emb_r.detach().numpy(),
        "D": D_r.detach().numpy(), "loss_plain": loss_r["plain"].detach().numpy(), "loss_unc": loss_r["unc"].detach().numpy(),
        "sampler_steps": np.array(steps), "sampler_out": xs_r.numpy(), "t_steps32": s32.t_steps.numpy(),
        "forced_wn_before": w_before.numpy(), "forced_wn_after": conv.weight.detach().numpy(),
    }
    for k, v in taps_r.items():
        out[f"tap/{k}"] = v.numpy().astype(np.float16)  # block outputs are O(1); fp16 keeps the fixture small
    for k in sel_grads:
        out[f"grad/{k}"] = grads_r[k].numpy()
    for k, v in grads_r.items():
        out[f"gradnorm/{k}"] = np.array(float(v.norm()))
    path = os.path.join(GOLDEN, "small_edm2.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); oracle == reference on {len(report)} checks")


if __name__ == "__main__":
    main()
