"""CUDA-graph replay (tinyedm_b200.graphs / DeterministicSolver) must compute exactly what the eager launches compute."""
import pytest
import torch

from tests.helpers import SMALL, build_modules, rel, small_params

pytestmark = pytest.mark.gpu


def _eager(solver, model, x0, labels):
    """The solver's eager loop exactly as `solve` runs it (under no_grad: with gradients enabled the Denoiser takes its
    training-style path, which stores the pixel-normalised block inputs and is not bit-identical to the inference path)."""
    with torch.no_grad():
        return solver._solve_eager(model, x0, labels)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (they fail, not skip, without one)"
    return torch.device("cuda:0")


def _small_edm(dev, T, **kw):
    dp, ep, _ = small_params()
    den, emb_m, _ = build_modules(SMALL, dp, ep, None, dev)
    return T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb_m, denoiser=den, use_ema=False, use_uncertainty=False,
                 steady_steps=1, rampup_steps=1, scheduler_interval="step", **kw)


def test_sampler_graph_replay_is_bit_identical_to_eager(dev, golden):
    import tinyedm_b200 as T
    model = _small_edm(dev, T).eval()
    x0 = torch.from_numpy(golden["x0"]).to(dev)
    labels = torch.from_numpy(golden["labels"]).to(dev)
    solver = T.DeterministicSolver(num_steps=6)
    eager = _eager(solver, model, x0.float().contiguous(), labels)
    first = solver.solve(model, x0, labels)            # captures
    ent = next(iter(solver._graphs.values()))
    assert ent.get("graph") is not None, ent.get("error")
    again = solver.solve(model, x0, labels)            # replays
    assert torch.equal(first, eager) and torch.equal(again, eager)
    # different inputs through the same graph
    x1 = torch.randn_like(x0)
    assert torch.equal(solver.solve(model, x1, labels), _eager(solver, model, x1.float().contiguous(), labels))
    # a parameter update must reach the replayed graph (weights are re-normalised outside the graph)
    with torch.no_grad():
        model.denoiser.encoder_blocks[0].conv_3x3_1.weight.mul_(-1.0)
    changed = solver.solve(model, x0, labels)
    assert not torch.equal(changed, eager)
    assert torch.equal(changed, _eager(solver, model, x0.float().contiguous(), labels))


def test_graphed_train_step_matches_eager_step(dev):
    """From a bit-identical parameter state and the same noise, the replayed graph must produce the gradients of the eager
    step (up to the summation order of the atomically reduced sums). The state has to be restored in between because a
    training-mode forward rewrites every weight in place (forced weight normalisation, networks.py:32-34), which is not
    exactly idempotent in fp32 and perturbs the bf16 operands of the next pass."""
    import tinyedm_b200 as T
    torch.manual_seed(0)
    B = 8
    clean = (0.5 * torch.randn(B, 3, 16, 16, device=dev)).clamp(-1, 1)
    labels = torch.randint(0, 10, (B,), device=dev)
    fixed_sigma = torch.exp(torch.randn(B, device=dev) * 1.2 - 1.2)
    fixed_noise = torch.randn_like(clean)
    m = _small_edm(dev, T).train()
    with torch.no_grad():
        m.denoiser.gain_out.fill_(1.0)
    m.diffuser.forward = lambda x: (x + fixed_sigma.view(-1, 1, 1, 1) * fixed_noise, fixed_sigma)
    opt = T.FusedAdamEMA(m.parameters(), lr=1e-3)
    step = T.GraphedTrainStep(m, opt, (clean, labels), warmup=1)
    assert step.graph is not None, step.error
    assert step.launches_per_step > 50
    params = dict(m.named_parameters())
    state0 = {n: p.detach().clone() for n, p in params.items()}

    def restore():
        with torch.no_grad():
            for n, p in params.items():
                p.copy_(state0[n])

    from tinyedm_b200.engine import bump_weights_epoch
    opt.zero_grad(set_to_none=True)
    bump_weights_epoch()      # like the captured step: its first forward always re-normalises the weights
    loss_e = m.training_step((clean, labels), 0)
    loss_e.backward()
    g_e = {n: p.grad.clone() for n, p in params.items() if p.grad is not None}
    after_e = {n: p.detach().clone() for n, p in params.items()}
    restore()
    step.graph.replay()
    g_g = {n: p.grad.clone() for n, p in params.items() if p.grad is not None}
    assert g_e.keys() == g_g.keys() and len(g_e) == len(params)
    assert abs(float(loss_e.detach()) - float(step.loss)) <= 1e-5 * abs(float(loss_e.detach()))
    for k in g_e:
        if float(g_e[k].norm()) > 0:
            assert rel(g_g[k], g_e[k]) < 1e-4, (k, rel(g_g[k], g_e[k]))
    for n, p in params.items():          # the in-place forced weight normalisation happened inside the graph too
        assert torch.equal(p.detach(), after_e[n]), n
    # full graphed steps (graph + optimiser launch) train
    losses = [float(step((clean, labels))) for _ in range(6)]
    assert all(l == l and l < 1e4 for l in losses)
    assert losses[-1] < losses[0]
    assert any(not torch.equal(p.detach(), after_e[n]) for n, p in params.items())
