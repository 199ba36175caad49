"""the reference's checkpoint layout (hyper_parameters = deinstantiate tree, state_dict with the
reference's names, optimizer_states[0]["ema"] = positional EMA tuple) loads into the new modules. CPU only (module
construction and state handling; no kernels). Mirrors the reference's tests/test_deinstantiate.py."""
import torch

import tinyedm_b200 as T
from tinyedm_b200.configs import CIFAR10, MNIST, build_edm


def _as_reference_tree(tree):
    """What the reference itself would have written: `_target_` paths inside the `tinyedm` package."""
    if isinstance(tree, dict):
        out = {k: _as_reference_tree(v) for k, v in tree.items()}
        if "_target_" in out:
            name = out["_target_"].rsplit(".", 1)[1]
            out["_target_"] = {"EDM": "tinyedm.edm.EDM", "Diffuser": "tinyedm.edm.Diffuser"}.get(name, f"tinyedm.networks.{name}")
        return out
    if isinstance(tree, list):
        return [_as_reference_tree(v) for v in tree]
    return tree


def test_deinstantiate_round_trip_like_the_reference_test():
    model = build_edm(CIFAR10)
    cfg = T.deinstantiate(model)
    assert cfg["_target_"].endswith(".EDM") and cfg["denoiser"]["encoder_block_types"][0] == "Enc"
    assert isinstance(cfg["denoiser"]["encoder_out_channels"], list)          # tuples become lists (utils.py:21-22)
    again = T.instantiate(cfg)
    assert isinstance(again, type(model))
    again.load_state_dict(model.state_dict(), strict=True)
    assert T.deinstantiate(again) == cfg


def test_reference_format_checkpoint_loads(tmp_path):
    torch.manual_seed(3)
    model = build_edm(MNIST)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.01 * torch.randn_like(p))
    ema = tuple(p.detach() * 0.5 + 1.0 for p in model.parameters())
    ckpt = {"hyper_parameters": _as_reference_tree(model.save_config()), "state_dict": model.state_dict(),
            "optimizer_states": [{"opt": {}, "ema": ema, "current_step": 7, "gamma": 6.94, "every_n_steps": 1}]}
    path = tmp_path / "ref.ckpt"
    torch.save(ckpt, path)
    loaded = T.EDM.load_from_checkpoint(path)
    assert isinstance(loaded, T.EDM) and loaded.num_classes == 10
    for (n1, p1), (n2, p2) in zip(model.named_parameters(), loaded.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)
    loaded_ema = T.EDM.load_from_checkpoint(path, load_ema=True)
    for p, e in zip(loaded_ema.parameters(), ema):
        assert torch.equal(p, e)
    del ckpt["optimizer_states"][0]["ema"]
    try:
        T.load_reference_checkpoint(ckpt, load_ema=True)
        assert False, "missing EMA weights must raise"
    except ValueError as e:
        assert "EMA weights not found" in str(e)


def test_swap_tensors():
    a, b = torch.arange(4.0), torch.ones(4)
    T.swap_tensors(a, b)
    assert torch.equal(a, torch.ones(4)) and torch.equal(b, torch.arange(4.0))


def test_optimizer_state_dict_layouts_round_trip_on_cpu():
    """`FusedAdamEMA.state_dict()` is the reference's EMAOptimizer layout (ema.py:326-336) around torch's Adam
    layout; `load_state_dict` takes it back (and a plain Adam state dict too) without touching a device — the flat
    kernel buffers are built from the loaded tensors at the first step on the GPU."""
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(()))]
    ref = torch.optim.Adam(ps, lr=0.02, betas=(0.9, 0.999))
    for p in ps:
        p.grad = torch.randn_like(p)
    ref.step(); ref.step()
    ema = tuple(p.detach() * 0.5 for p in ps)
    ref_sd = {"opt": ref.state_dict(), "ema": ema, "current_step": 2, "gamma": 6.94, "every_n_steps": 1}
    opt = T.FusedAdamEMA(ps, lr=1.0, ema_length=0.13)
    opt.load_state_dict(ref_sd)
    assert opt.current_step == 2 and abs(opt.gamma - 6.94) < 1e-12 and opt.param_groups[0]["lr"] == 0.02
    sd = opt.state_dict()
    assert set(sd) == {"opt", "ema", "current_step", "gamma", "every_n_steps"}
    for i, p in enumerate(ps):
        assert torch.equal(sd["opt"]["state"][i]["exp_avg"], ref.state[p]["exp_avg"])
        assert torch.equal(sd["opt"]["state"][i]["exp_avg_sq"], ref.state[p]["exp_avg_sq"])
        assert float(sd["opt"]["state"][i]["step"]) == 2.0
        assert torch.equal(sd["ema"][i], ema[i])
    torch.optim.Adam(ps, lr=0.1).load_state_dict(sd["opt"])        # the reference's optimiser accepts the Adam part
    # plain Adam layout into an optimiser without EMA: step count recovered from the per-parameter `step`
    plain = T.FusedAdamEMA(ps, lr=1.0)
    plain.load_state_dict(ref.state_dict())
    assert plain.current_step == 2 and set(plain.state_dict()) == {"state", "param_groups"}
    try:
        opt.load_state_dict({**ref_sd, "ema": ema[:1]})
        assert False, "a wrong number of EMA tensors must raise"
    except RuntimeError as e:
        assert "EMA tensors" in str(e)
