"""Does it LEARN like the reference? 300 optimiser steps on a fixed structured data set, same initial weights and
hyper-parameters on both sides:

* this library: `GraphedTrainStep` replays (Philox diffuser -> embedding -> denoiser -> weighted MSE -> backward -> forced
  weight normalisation -> fused Adam), i.e. the path bench.py times;
* the oracle (oracle/edm2_oracle.py, fp32 torch ops evaluated on the GPU so that it takes seconds): edm.py:205-236 with
  `torch.optim.Adam` and networks.py:32-34's in-place normalisation.

The noise draws differ (in-kernel Philox vs torch.randn), so the curves are compared statistically: both must come down
from their starting level, and the means of the last 60 steps must agree to 25 % (a batch of 64 sigma draws makes a single
step's loss noisy; a broken gradient, optimiser or normalisation path does not learn at all). Afterwards the weights are
still on the norm sphere and a 12-step Heun solve with the trained weights gives finite images of the data's scale."""
import math

import pytest
import torch

from oracle import edm2_oracle as O
from tests.helpers import SMALL, build_modules, small_params

pytestmark = pytest.mark.gpu

STEPS, B, LR = 300, 64, 5e-3


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from tinyedm_b200 import ops
    d = torch.device("cuda:0")
    ops.ensure_device(d)
    return d


def _dataset(n_classes, gen):
    """256 smooth 3x16x16 images, one family of patterns per class, standard deviation about sigma_data = 0.5."""
    ys, xs = torch.meshgrid(torch.arange(16.0), torch.arange(16.0), indexing="ij")
    imgs, labels = [], []
    for i in range(256):
        k = i % n_classes
        ph = torch.rand(3, generator=gen) * 2 * math.pi
        amp = 0.85 + 0.2 * torch.rand(3, generator=gen)
        ch = [amp[c] * torch.sin(2 * math.pi * (k + 1) * xs / 16 + ph[c]) * torch.cos(2 * math.pi * (c + 1) * ys / 16 + ph[c]) for c in range(3)]
        imgs.append(torch.stack(ch))
        labels.append(k)
    return torch.stack(imgs).clamp(-1, 1), torch.tensor(labels)


def test_training_run_learns_like_the_oracle(dev):
    import tinyedm_b200 as T
    cfg = SMALL
    sd = cfg["denoiser"].sigma_data
    gen = torch.Generator().manual_seed(99)
    data, labels = _dataset(cfg["embedding"].num_classes, gen)
    data, labels = data.to(dev), labels.to(dev)
    order = [torch.randperm(256, generator=gen)[:B].to(dev) for _ in range(STEPS)]
    dp, ep, _ = small_params()

    # ---------------- oracle ----------------
    torch.manual_seed(5)
    dpo = {k: v.clone().to(dev).requires_grad_(True) for k, v in dp.items()}
    epo = {k: v.clone().to(dev) for k, v in ep.items()}
    train_e = [v.requires_grad_(True) for k, v in epo.items() if k.endswith("weight")]
    weights = [v for k, v in list(dpo.items()) + list(epo.items()) if k.endswith("weight")]
    opt_o = torch.optim.Adam(list(dpo.values()) + train_e, lr=LR, betas=(0.9, 0.999))
    loss_o = []
    for it in range(STEPS):
        with torch.no_grad():
            for w in weights:
                O.forced_weight_norm_(w)
        clean, lab = data[order[it]], labels[order[it]]
        noisy, sigma = O.diffuse(clean, torch.randn(B, device=dev), torch.randn_like(clean), -1.2, 1.2)
        _, e = O.embedding_forward(epo, cfg["embedding"], sigma, lab)
        D = O.denoiser_forward(dpo, cfg["denoiser"], noisy, sigma, e)
        loss = O.training_loss(O.loss_weight(sigma, sd), D, clean)
        opt_o.zero_grad(set_to_none=True)
        loss.backward()
        opt_o.step()
        loss_o.append(float(loss.detach()))

    # ---------------- this library ----------------
    den, emb_m, _ = build_modules(cfg, dp, ep, None, dev)
    model = T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb_m, denoiser=den, use_ema=False, use_uncertainty=False,
                  steady_steps=10 ** 9, rampup_steps=1, scheduler_interval="step", lr=LR).to(dev).train()
    opt = model.configure_optimizers()["optimizer"]
    for g in opt.param_groups:
        g["lr"] = LR
    step = T.GraphedTrainStep(model, opt, (data[order[0]], labels[order[0]]))
    assert step.graph is not None, step.error
    loss_g = [step((data[order[it]], labels[order[it]])).clone() for it in range(STEPS)]
    loss_g = [float(v) for v in torch.cat([v.reshape(1) for v in loss_g])]

    head = lambda v: sum(v[:10]) / 10
    tail = lambda v: sum(v[-60:]) / 60
    print(f"loss, first 10 -> last 60 steps: oracle {head(loss_o):.3f} -> {tail(loss_o):.3f}; this library {head(loss_g):.3f} -> {tail(loss_g):.3f}")
    assert all(math.isfinite(v) for v in loss_g)
    assert tail(loss_o) < 0.7 * head(loss_o), "the oracle itself did not learn: the test set-up is wrong"
    assert tail(loss_g) < 0.7 * head(loss_g)
    assert abs(tail(loss_g) - tail(loss_o)) < 0.25 * tail(loss_o)
    # forced weight normalisation kept every weight row on the sphere: ||w_row|| = sqrt(fan_in) (networks.py:32-34)
    for name, p in list(den.named_parameters()) + list(emb_m.named_parameters()):
        if p.ndim >= 2:
            w = p.detach().flatten(1)
            n = w.norm(dim=1) / math.sqrt(w.shape[1])
            assert float((n - 1).abs().max()) < 2e-2, (name, float((n - 1).abs().max()))      # one Adam step off the sphere
    # the trained weights sample: finite images at the data's scale, closer to the data than noise is
    model.eval()
    x0 = torch.randn(32, 3, 16, 16, generator=torch.Generator().manual_seed(1)).to(dev)
    lab = (torch.arange(32) % cfg["embedding"].num_classes).to(dev)
    with torch.no_grad():
        imgs = T.DeterministicSolver(num_steps=12).solve(model, x0, lab)
    assert torch.isfinite(imgs).all()
    assert 0.1 < float(imgs.std()) < 1.0 and float(imgs.abs().max()) < 3.0, (float(imgs.std()), float(imgs.abs().max()))
