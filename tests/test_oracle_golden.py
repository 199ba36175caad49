"""CPU: the oracle reproduces the REFERENCE's stored outputs (tests/golden/small_edm2.npz was written by
oracle/make_golden.py from /root/reference/src/tinyedm/{networks,solvers}.py)."""
import numpy as np
import torch

from oracle import edm2_oracle as O
from tests.helpers import SMALL, checksum, rel, small_inputs, small_params


def test_seeded_parameters_match_the_fixture(golden):
    dp, ep, up = small_params()
    got = np.concatenate([checksum(dp), checksum(ep), checksum(up)])
    np.testing.assert_allclose(got, golden["param_checksum"], rtol=1e-9)
    clean, eps, noise, labels, x0 = small_inputs()
    np.testing.assert_array_equal(clean.numpy(), golden["clean"])
    np.testing.assert_array_equal(labels.numpy(), golden["labels"])


def test_forward_loss_and_taps(golden):
    cfg = SMALL
    dp, ep, up = small_params()
    clean, eps, noise, labels, _ = small_inputs()
    noisy, sigma = O.diffuse(clean, eps, noise, -1.2, 1.2)
    assert rel(noisy, golden["noisy"]) == 0 and rel(sigma, golden["sigma"]) == 0
    four, emb = O.embedding_forward(ep, cfg["embedding"], sigma, labels)
    assert rel(four, golden["fourier"]) < 1e-6 and rel(emb, golden["embedding"]) < 1e-6
    taps = {}
    D = O.denoiser_forward(dp, cfg["denoiser"], noisy, sigma, emb, taps=taps)
    assert rel(D, golden["D"]) < 1e-5
    for k in golden.files:
        if k.startswith("tap/"):
            assert rel(taps[k[4:]], golden[k].astype(np.float32)) < 2e-3, k   # fixture taps are stored in fp16
    w = O.loss_weight(sigma, cfg["denoiser"].sigma_data)
    assert rel(O.training_loss(w, D, clean), golden["loss_plain"]) < 1e-6
    assert rel(O.training_loss(w, D, clean, O.uncertainty_forward(up, four)), golden["loss_unc"]) < 1e-6


def test_gradients(golden):
    cfg = SMALL
    dp, ep, up = small_params()
    dp = {k: v.requires_grad_(True) for k, v in dp.items()}
    ep = {k: (v.requires_grad_(True) if k.endswith("weight") else v) for k, v in ep.items()}
    up = {k: v.requires_grad_(True) for k, v in up.items()}
    clean, eps, noise, labels, _ = small_inputs()
    noisy, sigma = O.diffuse(clean, eps, noise, -1.2, 1.2)
    four, emb = O.embedding_forward(ep, cfg["embedding"], sigma, labels)
    D = O.denoiser_forward(dp, cfg["denoiser"], noisy, sigma, emb)
    loss = O.training_loss(O.loss_weight(sigma, 0.5), D, clean, O.uncertainty_forward(up, four))
    loss.backward()
    named = {**{f"denoiser.{k}": v for k, v in dp.items()}, **{f"embedding.{k}": v for k, v in ep.items()},
             **{f"u.{k}": v for k, v in up.items()}}
    n = 0
    for k in golden.files:
        if k.startswith("grad/"):
            assert rel(named[k[5:]].grad, golden[k]) < 2e-4, k
            n += 1
        elif k.startswith("gradnorm/"):
            assert abs(float(named[k[9:]].grad.norm()) - float(golden[k])) <= 2e-4 * float(golden[k]) + 1e-9, k
    assert n >= 15


def test_sampler_schedule_and_trajectory(golden):
    cfg = SMALL
    dp, ep, _ = small_params()
    _, _, _, labels, x0 = small_inputs()
    ts = O.t_schedule(32)
    np.testing.assert_array_equal(ts.numpy(), golden["t_steps32"])
    assert abs(float(ts[0]) - 80.0) < 1e-4 and abs(float(ts[31]) - 0.002) < 1e-7 and ts[32] == 0.0
    calls = []
    def model(x, s, lab):
        calls.append(float(s))
        return O.edm_forward(dp, cfg["denoiser"], ep, cfg["embedding"], x, s, lab)
    with torch.no_grad():
        out = O.heun_solve(model, x0, labels, num_steps=int(golden["sampler_steps"]))
    assert len(calls) == 2 * int(golden["sampler_steps"]) - 1      # N steps => 2N-1 network evaluations
    assert rel(out, golden["sampler_out"]) < 1e-5


def test_forced_weight_norm_and_known_answers(golden):
    w = torch.from_numpy(golden["forced_wn_before"]).clone()
    O.forced_weight_norm_(w)
    assert rel(w, golden["forced_wn_after"]) < 1e-7
    # gain_out == 0 hides the network: D == c_skip * x exactly (SURVEY.md §0 parity trap)
    cfg = SMALL
    dp, ep, _ = small_params(gain_out=0.0)
    clean, eps, noise, labels, _ = small_inputs()
    noisy, sigma = O.diffuse(clean, eps, noise, -1.2, 1.2)
    _, emb = O.embedding_forward(ep, cfg["embedding"], sigma, labels)
    D = O.denoiser_forward(dp, cfg["denoiser"], noisy, sigma, emb)
    c_skip, _, _ = O.precond_coeffs(sigma, 0.5)
    assert torch.equal(D, noisy * c_skip)
    # the reference's own metric test identity (tests/test_weighted_mean_squared_error.py:18-21)
    g = torch.Generator().manual_seed(3)
    wt, p, t = torch.rand(8, generator=g), torch.randn(8, 3, 32, 32, generator=g), torch.randn(8, 3, 32, 32, generator=g)
    assert torch.allclose(O.weighted_mse(wt, p, t).squeeze(), torch.mean(wt[:, None, None, None] * (p - t) ** 2))
