"""The step-level surfaces around the denoiser, CUDA path vs the oracle / the reference's golden vectors:
`Diffuser.forward` (edm.py:84-93), `EDM.validation_step` (:238-248), `EDM.predict_step` (:288-295), gradient accumulation
(imagenet.yaml:7 `accumulate_grad_batches: 3`), the optimiser's checkpoint round trip (ema.py:326-348) and the EMA weight
swap (ema.py:293-317) with the eval-mode cache of normalised weights."""
import copy

import pytest
import torch

from oracle import edm2_oracle as O
from tests.helpers import SMALL, build_modules, rel, small_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from tinyedm_b200 import ops
    d = torch.device("cuda:0")
    ops.ensure_device(d)
    return d


def _small_edm(dev, use_uncertainty=False, **kw):
    import tinyedm_b200 as T
    dp, ep, up = small_params()
    den, emb_m, unc = build_modules(SMALL, dp, ep, up, dev)
    args = dict(use_ema=False, use_uncertainty=use_uncertainty, steady_steps=1, rampup_steps=1, scheduler_interval="step")
    args.update(kw)
    model = T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb_m, denoiser=den, **args)
    if use_uncertainty:
        model.u = unc
    return model.to(dev), (dp, ep, up)


def test_diffuse_kernel_vs_reference_golden(dev, golden):
    """tedm_diffuse against the reference's own numbers (oracle/make_golden.py stores edm.py:84-93 evaluated by torch on
    the committed eps / noise draws): fp32, <= 1e-6."""
    from tinyedm_b200 import ops
    clean, eps, noise = (torch.from_numpy(golden[k]).to(dev) for k in ("clean", "eps", "noise"))
    noisy, sigma = ops.diffuse(clean, eps, noise, -1.2, 1.2)
    assert noisy.shape == clean.shape and sigma.shape == (clean.shape[0],)
    assert rel(sigma, golden["sigma"]) < 1e-6
    assert rel(noisy, golden["noisy"]) < 1e-6
    no, so = O.diffuse(clean.cpu(), eps.cpu(), noise.cpu(), -1.2, 1.2)
    assert rel(noisy, no) < 1e-6 and rel(sigma, so) < 1e-6
    # other (P_mean, P_std) and a batch that is not a multiple of anything
    g = torch.Generator().manual_seed(0)
    c2, e2, n2 = torch.randn(7, 4, 64, 64, generator=g), torch.randn(7, generator=g), torch.randn(7, 4, 64, 64, generator=g)
    noisy2, sigma2 = ops.diffuse(c2.to(dev), e2.to(dev), n2.to(dev), -0.4, 1.0)
    no2, so2 = O.diffuse(c2, e2, n2, -0.4, 1.0)
    assert rel(noisy2, no2) < 1e-6 and rel(sigma2, so2) < 1e-6


def test_diffuser_module_statistics_and_freshness(dev):
    """`Diffuser.forward` (in-kernel Philox draws): ln(sigma) ~ N(P_mean, P_std), noise ~ N(0, sigma^2)
    independent of the image, new draws on every call, no gradient (edm.py:84 `@torch.no_grad`), reproducible from
    (seed, step)."""
    import tinyedm_b200 as T
    torch.manual_seed(123)
    diff = T.Diffuser(-1.2, 1.2)
    clean = torch.zeros(4096, 3, 8, 8, device=dev)
    noisy, sigma = diff(clean)
    assert not noisy.requires_grad and sigma.shape == (4096,) and noisy.dtype == torch.float32
    ls = sigma.log()
    assert abs(float(ls.mean()) + 1.2) < 0.08 and abs(float(ls.std()) - 1.2) < 0.08
    z = noisy / sigma.view(-1, 1, 1, 1)                       # unit normal draws
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert abs(float((z ** 3).mean())) < 2e-2 and abs(float((z ** 4).mean()) - 3.0) < 5e-2     # skewness 0, kurtosis 3
    assert abs(float((z[:, 0] * z[:, 1]).mean())) < 1e-2      # channels uncorrelated
    assert abs(float((z[:-1] * z[1:]).mean())) < 1e-2         # images uncorrelated
    assert abs(float((z[..., :-1] * z[..., 1:]).mean())) < 1e-2   # neighbouring pixels (they share a Philox call) uncorrelated
    assert float(z.abs().max()) > 4.0                          # tails are there (786 432 draws)
    noisy2, sigma2 = diff(clean)
    assert not torch.equal(sigma, sigma2) and not torch.equal(noisy, noisy2)
    diff.seed(diff.rng_state()[0], 0)                          # rewind: the same (seed, step) gives the same draws
    noisy3, sigma3 = diff(clean)
    assert torch.equal(noisy3, noisy) and torch.equal(sigma3, sigma)


@pytest.mark.parametrize("B,C,H,W", [(6, 3, 32, 32), (5, 1, 28, 28), (3, 4, 64, 64), (2, 3, 16, 16)])
def test_fused_diffuser_equals_its_parts(dev, B, C, H, W):
    """The fused kernel against the separate steps it replaces, on ITS OWN draws: noisy/sigma == edm.py:84-93 evaluated by
    the oracle on (epsilon, n); the hand-over operand == tedm_conv_in_im2col(noisy, sigma) bit for bit; the Denoiser
    consumes the hand-over (same output as when it gathers the patches itself)."""
    import tinyedm_b200 as T
    from tinyedm_b200 import ops
    g = torch.Generator().manual_seed(B + H)
    clean = (0.5 * torch.randn(B, C, H, W, generator=g)).clamp(-1, 1).to(dev)
    diff = T.Diffuser(-0.4, 1.0)
    diff._fuse_sigma_data = 0.5
    diff.seed(1234567, 41)
    noisy, sigma = diff(clean)
    eps, noise = diff.draws(B, C * H * W, dev)
    no, so = O.diffuse(clean.cpu(), eps.cpu(), noise.view(B, C, H, W).cpu(), -0.4, 1.0)
    assert rel(sigma, so) < 1e-6 and rel(noisy, no) < 1e-6
    n2, s2 = ops.diffuse(clean, eps, noise.view(B, C, H, W).contiguous(), -0.4, 1.0)     # the two-input kernel on the same draws
    assert rel(noisy, n2) < 1e-6 and rel(sigma, s2) < 1e-6
    xcol, sd, ptr = noisy._tedm_xcol
    assert sd == 0.5 and ptr == sigma.data_ptr()
    assert torch.equal(xcol, ops.conv_in_im2col(noisy, sigma, 0.5))


def test_training_step_uses_the_fused_input_block_and_graph_replays_draw_fresh_noise(dev):
    import tinyedm_b200 as T
    from tinyedm_b200 import ops
    model, _ = _small_edm(dev)
    model.train()
    assert model.diffuser._fuse_sigma_data == 0.5
    g = torch.Generator().manual_seed(4)
    clean = (0.5 * torch.randn(8, 3, 16, 16, generator=g)).clamp(-1, 1).to(dev)
    labels = torch.randint(0, 5, (8,), generator=g).to(dev)
    calls = []
    orig = ops.conv_in_im2col
    ops.conv_in_im2col = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        model.diffuser.seed(9, 0)
        l1 = float(model.training_step((clean, labels), 0))
        assert not calls, "the engine re-gathered the patches although the diffuser handed them over"
        # a caller-made noisy image (no hand-over) still works and gives the same loss for the same noise
        noisy, sigma = model.diffuser(clean)          # step 2
        del noisy._tedm_xcol
        _, e = model.embedding(sigma, labels)
        D = model.denoiser(noisy, sigma, e)
        assert calls and torch.isfinite(D).all()
    finally:
        ops.conv_in_im2col = orig
    # (an autograd graph that is still alive pins the parameters' AccumulateGrad nodes to the stream it was built on — the
    # default stream here — and a capture on another stream would have to synchronise with it)
    del D, e, noisy, sigma
    opt = T.FusedAdamEMA(model.parameters(), lr=1e-4)
    step = T.GraphedTrainStep(model, opt, (clean, labels), warmup=1)
    assert step.graph is not None, getattr(step, "error_traceback", step.error)
    losses = [float(step((clean, labels))) for _ in range(4)]
    assert len({round(l, 6) for l in losses}) == 4, losses     # same batch, different sigma / noise on every replay
    assert abs(losses[0] - l1) > 0


def test_validation_step_vs_oracle(dev):
    """edm.py:238-248 on the CUDA path: the (noisy, sigma) the diffuser produced are taken from a forward hook and pushed
    through the fp32 oracle (embedding -> denoiser -> lambda(sigma)-weighted MSE); the bf16 network's accumulated drift
    bounds the loss error. The running metric state follows metric.py:33-36."""
    model, (dp, ep, _) = _small_edm(dev)
    model.eval()
    seen = {}
    h = model.diffuser.register_forward_hook(lambda m, i, o: seen.update(noisy=o[0].detach().cpu(), sigma=o[1].detach().cpu()))
    g = torch.Generator().manual_seed(3)
    clean = (0.5 * torch.randn(6, 3, 16, 16, generator=g)).clamp(-1, 1)
    labels = torch.randint(0, 5, (6,), generator=g)
    torch.manual_seed(11)
    with torch.no_grad():
        loss = model.validation_step((clean.to(dev), labels.to(dev)), 0)
    h.remove()
    assert loss.shape == (1,)
    # the diffuser's output is edm.py:84-93 of SOME draws: sigma positive, noisy - clean = sigma * unit normal
    z = (seen["noisy"] - clean) / seen["sigma"].view(-1, 1, 1, 1)
    assert abs(float(z.std()) - 1.0) < 0.05
    _, emb = O.embedding_forward(ep, SMALL["embedding"], seen["sigma"], labels)
    D_o = O.denoiser_forward(dp, SMALL["denoiser"], seen["noisy"], seen["sigma"], emb)
    loss_o = O.training_loss(O.loss_weight(seen["sigma"], 0.5), D_o, clean)
    r = rel(loss, loss_o)
    print(f"validation_step loss vs fp32 oracle: {r:.2e}")
    assert r < 4e-2
    assert int(model.val_mse.total) == 6 and rel(model.val_mse.compute(), loss_o) < 4e-2
    # an unconditional model ignores the labels it is handed (edm.py:240)
    import tinyedm_b200 as T
    dp2, ep2, _ = small_params()
    ep2 = {k: v for k, v in ep2.items() if not k.startswith("class_embed")}
    import dataclasses
    cfg_u = dict(SMALL)
    cfg_u["embedding"] = dataclasses.replace(SMALL["embedding"], num_classes=None)
    den2, emb2, _ = build_modules(cfg_u, dp2, ep2, None, dev)
    m2 = T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb2, denoiser=den2, use_ema=False, use_uncertainty=False,
               steady_steps=1, rampup_steps=1, scheduler_interval="step").eval()
    with torch.no_grad():
        assert torch.isfinite(m2.validation_step((clean.to(dev), labels.to(dev)), 0)).all()


def test_predict_step_vs_reference_golden(dev, golden):
    """edm.py:288-295: `predict_step(batch)` == `solver.solve(self, x0, labels)`; against the reference's 6-step sample."""
    import tinyedm_b200 as T
    model, _ = _small_edm(dev)
    model.eval()
    model.solver = T.DeterministicSolver(num_steps=int(golden["sampler_steps"]))
    x0 = torch.from_numpy(golden["x0"]).to(dev)
    labels = torch.from_numpy(golden["labels"]).to(dev)
    out = model.predict_step((x0, labels), 0)
    assert out.shape == x0.shape and out.dtype == torch.float32
    r = rel(out, golden["sampler_out"])
    print(f"predict_step (6-step Heun, bf16 network) vs fp32 reference: {r:.2e}")
    assert r < 4e-2
    assert torch.equal(out, model.solver.solve(model, x0, labels))


def _grads_of(model):
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def test_gradient_accumulation_equals_one_big_batch(dev):
    """3 micro-batches with deferred hand-out (one Jacobian, one exchange) == the gradient of the mean loss over the three,
    computed micro-batch by micro-batch through plain autograd accumulation, and == Lightning's semantics (loss / k)."""
    import tinyedm_b200 as T
    model, _ = _small_edm(dev, use_uncertainty=True)
    model.eval()                                # eval: no dropout, weights stay put -> the two routes see identical inputs
    g = torch.Generator().manual_seed(8)
    k, mb = 3, 4
    clean = (0.5 * torch.randn(k * mb, 3, 16, 16, generator=g)).clamp(-1, 1).to(dev)
    labels = torch.randint(0, 5, (k * mb,), generator=g).to(dev)

    def run(deferred: bool):
        model.zero_grad(set_to_none=True)
        model.diffuser.seed(5, 0)                # the same diffusion noise in every run
        for j in range(k):
            batch = (clean[j * mb:(j + 1) * mb], labels[j * mb:(j + 1) * mb])
            with model.denoiser.accumulate_grads(deferred and j < k - 1):
                loss = model.training_step(batch, j) / k
                loss.backward()
            if deferred and j < k - 1:
                assert model.denoiser.conv_in.weight.grad is None      # nothing handed out yet
        return _grads_of(model)

    g_plain = run(False)      # autograd accumulates three handed-out gradients
    g_defer = run(True)       # g_hat accumulates, one hand-out
    assert set(g_plain) == set(g_defer)
    worst = max((rel(g_defer[n], g_plain[n]), n) for n in g_plain)
    print("accumulation: deferred vs per-micro-batch hand-out, worst", worst)
    assert worst[0] < 2e-5
    # and a second optimiser step starts from a clean slate (the accumulator is reset by the hand-out)
    g_again = run(True)
    assert max(rel(g_again[n], g_defer[n]) for n in g_defer) < 2e-5


def test_optimizer_state_round_trip_resumes_exactly(dev):
    """save -> load -> step must continue the Adam moments, the EMA copies and the bias-correction step,
    in the reference's EMAOptimizer layout (ema.py:326-348)."""
    import tinyedm_b200 as T
    torch.manual_seed(0)
    shapes = [(64, 64, 3, 3), (130,), (), (16, 65, 1, 1)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    opt = T.FusedAdamEMA(ps, lr=0.02, betas=(0.9, 0.999), ema_length=0.13)
    grads = [[torch.randn_like(p) for p in ps] for _ in range(5)]
    for step in range(3):
        for p, gr in zip(ps, grads[step]):
            p.grad = gr.clone()
        opt.step()
    sd = copy.deepcopy(opt.state_dict())
    assert set(sd) == {"opt", "ema", "current_step", "gamma", "every_n_steps"} and sd["current_step"] == 3
    assert len(sd["ema"]) == len(ps) and set(sd["opt"]["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    # the Adam part loads into the reference's own optimiser (edm.py:250-253)
    ref_ps = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref = torch.optim.Adam(ref_ps, lr=0.02, betas=(0.9, 0.999))
    ref.load_state_dict(sd["opt"])
    assert float(ref.state[ref_ps[0]]["step"]) == 3.0
    # resume in a fresh optimiser over fresh parameter objects
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt2 = T.FusedAdamEMA(qs, lr=0.5, betas=(0.8, 0.9), ema_length=0.13)
    opt2.load_state_dict(sd)
    assert opt2.current_step == 3 and opt2.param_groups[0]["lr"] == 0.02 and tuple(opt2.param_groups[0]["betas"]) == (0.9, 0.999)
    for e1, e2 in zip(opt.ema_params, opt2.ema_params):
        assert torch.equal(e1, e2)
    for step in range(3, 5):
        for p, q, gr in zip(ps, qs, grads[step]):
            p.grad = gr.clone(); q.grad = gr.clone()
        opt.step(); opt2.step()
    for p, q in zip(ps, qs):
        assert torch.equal(p, q)
    for e1, e2 in zip(opt.ema_params, opt2.ema_params):
        assert torch.equal(e1, e2)
    assert torch.equal(opt.state[ps[0]]["exp_avg_sq"], opt2.state[qs[0]]["exp_avg_sq"])
    # every_n_steps > 1: the EMA moves only on steps whose 0-based index is a multiple of it (ema.py:262-269)
    rs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt3 = T.FusedAdamEMA(rs, lr=0.02, ema_length=0.13, every_n_steps=2)
    snaps = []
    for step in range(4):
        for r_, gr in zip(rs, grads[step]):
            r_.grad = gr.clone()
        opt3.step()
        snaps.append(opt3.ema_params[0].clone())
    assert torch.equal(snaps[0], snaps[1]) and not torch.equal(snaps[1], snaps[2]) and torch.equal(snaps[2], snaps[3])


def test_ema_swap_invalidates_cached_normalised_weights(dev):
    """`swap_tensors(param.data, ema)` bumps no version counter; the eval-mode weight cache must still
    notice (ema.py:293-296 is how the reference validates and samples with EMA weights)."""
    import tinyedm_b200 as T
    model, _ = _small_edm(dev)
    model.eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 3, 16, 16, generator=g).to(dev)
    sigma = torch.tensor([0.5, 1.0, 2.0], device=dev)
    labels = torch.tensor([0, 1, 2], device=dev)
    with torch.no_grad():
        d_a = model(x, sigma, labels).clone()
        other = [torch.randn_like(p) if p.ndim > 1 else p.detach().clone() for p in model.parameters()]
        for p, o in zip(model.parameters(), other):
            T.swap_tensors(p.data, o)
        d_b = model(x, sigma, labels).clone()
        assert rel(d_b, d_a) > 1e-2, "the swapped-in weights were ignored (stale cache)"
        for p, o in zip(model.parameters(), other):
            T.swap_tensors(p.data, o)
        assert torch.equal(model(x, sigma, labels), d_a)
        # raw `.data` writes + the explicit hook
        for p in model.parameters():
            if p.ndim > 1:
                p.data.mul_(-1.0)
        model.invalidate_weights()
        d_c = model(x, sigma, labels)
        assert rel(d_c, d_a) > 1e-2
    # the optimiser's own swap (ema.py:299-317)
    model.train()
    opt = T.FusedAdamEMA(model.parameters(), lr=1e-2, ema_length=0.13)
    clean = torch.randn(4, 3, 16, 16, generator=g).clamp(-1, 1).to(dev)
    for i in range(2):
        opt.zero_grad(set_to_none=True)
        model.training_step((clean, torch.zeros(4, dtype=torch.long, device=dev)), i).backward()
        opt.step()
    model.eval()
    with torch.no_grad():
        d_w = model(x, sigma, labels).clone()
        with opt.swap_ema_weights():
            d_e = model(x, sigma, labels).clone()
        assert not torch.equal(d_e, d_w)
        assert torch.equal(model(x, sigma, labels), d_w)


def test_graphed_accumulated_step_matches_eager_accumulation_and_leaves_the_optimizer_alone(dev):
    """`GraphedTrainStep(accumulate=3)` (imagenet.yaml:7): the captured three-micro-batch step computes what the eager
    deferred accumulation computes from the same state and noise; building it performs NO optimiser step (the
    warm-up used to take two hidden Adam steps)."""
    import tinyedm_b200 as T
    model, _ = _small_edm(dev, use_uncertainty=True)
    model.train()
    g = torch.Generator().manual_seed(21)
    k, mb = 3, 4
    clean = (0.5 * torch.randn(k * mb, 3, 16, 16, generator=g)).clamp(-1, 1).to(dev)
    labels = torch.randint(0, 5, (k * mb,), generator=g).to(dev)
    opt = T.FusedAdamEMA(model.parameters(), lr=1e-3, ema_length=0.13)
    model.training_step((clean[:mb], labels[:mb]), 0)          # one training forward: every weight is on its norm sphere now
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    step = T.GraphedTrainStep(model, opt, (clean, labels), accumulate=k, warmup=1)
    assert step.graph is not None, getattr(step, "error_traceback", step.error)
    assert opt.current_step == 0 and opt._flat is None                       # no optimiser step was taken
    for n, p in model.named_parameters():                                    # (re-normalising normalised weights: ~1e-7)
        assert rel(p, before[n]) < 1e-5, n
    # the captured step first (its `.grad` tensors are the ones bound at capture time) ...
    from tinyedm_b200.engine import bump_weights_epoch
    state0 = {n: p.detach().clone() for n, p in model.named_parameters()}
    model.diffuser.seed(3, 0)
    step.graph.replay()
    g_graph = _grads_of(model)
    # ... then the eager deferred accumulation from the same parameter state and the same noise (a training forward rewrites
    # the weights in place, which is not exactly idempotent in fp32: restore them)
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(state0[n])
    model.diffuser.seed(3, 0)
    opt.zero_grad(set_to_none=True)
    bump_weights_epoch()            # like the captured step: the first forward of a step always re-normalises
    for j in range(k):
        with model.denoiser.accumulate_grads(j < k - 1):
            (model.training_step((clean[j * mb:(j + 1) * mb], labels[j * mb:(j + 1) * mb]), j) / k).backward()
    g_eager = _grads_of(model)
    assert set(g_eager) == set(g_graph) and len(g_graph) == len(before)
    worst = max((rel(g_graph[n], g_eager[n]), n) for n in g_eager if float(g_eager[n].norm()) > 0)
    print("graphed accumulated step vs eager accumulation, worst gradient difference:", worst)
    assert worst[0] < 1e-3          # atomically reduced sums in a different order
    losses = [float(step((clean, labels))) for _ in range(5)]
    assert opt.current_step == 5 and all(l == l for l in losses)
    # the replayed graph re-normalises the weights the optimiser moved (networks.py:32-34): after a replay every filter is
    # back on its norm sphere, and the prepared operand is the normalised CURRENT weight
    w = model.denoiser.encoder_blocks[0].conv_3x3_1.weight
    with torch.no_grad():
        w.mul_(1.5)                             # what an optimiser step does, exaggerated: off the sphere ...
    off = rel(w.detach().flatten(1).norm(dim=1), torch.full((w.shape[0],), float(w[0].numel()) ** 0.5))
    assert off > 0.4, off
    step.graph.replay()                         # ... and the next step's forward pulls it back
    torch.cuda.synchronize()
    assert rel(w.detach().flatten(1).norm(dim=1), torch.full((w.shape[0],), float(w[0].numel()) ** 0.5)) < 2e-4
    slot = model.denoiser.engine.blocks[0].w["conv_3x3_1"]
    w_hat = O.effective_weight(w.detach().cpu())
    assert rel(slot.fwd.float().view(w.shape[0], 3, 3, w.shape[1]).permute(0, 3, 1, 2), w_hat) < 3e-3


def test_unchanged_weights_are_not_renormalised_twice_but_updated_ones_are(dev):
    """WeightBank.prepare: the 2nd / 3rd micro-batch of an accumulated step finds the operands of the 1st (no parameter
    moved), an optimiser step or an EMA swap invalidates them."""
    import tinyedm_b200 as T
    from tinyedm_b200 import ops
    model, _ = _small_edm(dev)
    model.train()
    g = torch.Generator().manual_seed(2)
    clean = (0.5 * torch.randn(4, 3, 16, 16, generator=g)).clamp(-1, 1).to(dev)
    labels = torch.randint(0, 5, (4,), generator=g).to(dev)
    calls = []
    orig = ops.weight_prep_forward
    ops.weight_prep_forward = lambda *a, **kw: (calls.append(a[-1]), orig(*a, **kw))[1]
    try:
        opt = T.FusedAdamEMA(model.parameters(), lr=1e-2)
        def fwd_bwd():
            opt.zero_grad(set_to_none=True)
            model.training_step((clean, labels), 0).backward()
        fwd_bwd()
        n1 = len(calls)
        assert n1 >= 2                      # denoiser + embedding banks
        w = model.denoiser.encoder_blocks[0].conv_3x3_1.weight
        assert rel(w.detach().flatten(1).norm(dim=1), torch.full((w.shape[0],), float(w[0].numel()) ** 0.5)) < 2e-4   # sqrt(n) (1 - eps)
        fwd_bwd()
        assert len(calls) == n1             # nothing changed: no second pass
        opt.step()
        fwd_bwd()
        assert len(calls) == 2 * n1         # the optimiser moved the weights off the sphere: re-normalised
        assert rel(w.detach().flatten(1).norm(dim=1), torch.full((w.shape[0],), float(w[0].numel()) ** 0.5)) < 2e-4
        model.eval()
        with torch.no_grad():
            model(clean, torch.ones(4, device=dev), labels)
        assert len(calls) == 2 * n1         # eval right after a training pass: operands are current
    finally:
        ops.weight_prep_forward = orig
