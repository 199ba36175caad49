"""CPU tests (`-m "not gpu"`): the C-ABI library loads and exports everything the header declares, and the host-side
mirror of the reference interface (module tree, state_dict names/order, builders, schedules) is intact.
No compute call is made here: there is no GPU in the build container."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import tinyedm_b200 as T
from oracle import edm2_oracle as O
from tests.helpers import SMALL, cifar_cfg, small_params, spec_kwargs
from tinyedm_b200 import _lib, networks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    sigs = _lib.header_signatures()
    assert len(sigs) >= 30
    for name in sigs:
        assert hasattr(lib, name), f"{name} is declared in include/tinyedm_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (tedm_\w+)", out))
    assert exported == set(sigs), exported ^ set(sigs)
    assert _lib.call_int("tedm_version") >= 100
    assert _lib.call_int("tedm_adam_chunk_elems") > 0


def test_descriptor_structs_match_the_header():
    assert ctypes.sizeof(_lib.WeightDesc) == 7 * 8 + 8 * 4
    from tinyedm_b200.optim import AdamDesc
    assert ctypes.sizeof(AdamDesc) == 6 * 8


def test_sm100a_sass_contains_tcgen05_and_tma():
    """The conv kernels must be real Blackwell tensor-core code (UTCHMMA = tcgen05.mma, UTMALDG = TMA load)."""
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UTCHMMA" in out and "UTMALDG" in out


def test_no_cpu_fallback():
    den = T.Denoiser(**spec_kwargs(SMALL["denoiser"]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        den(torch.zeros(1, 3, 16, 16), torch.ones(1), torch.zeros(1, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.DeterministicSolver(4).solve(lambda *a: None, torch.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError):
        T.Diffuser(-1.2, 1.2)(torch.zeros(2, 3, 8, 8))
    with pytest.raises(RuntimeError, match="fused inside"):
        den.encoder_blocks[0](torch.zeros(1), torch.zeros(1))


def test_product_package_never_imports_the_oracle():
    for fn in os.listdir(os.path.join(ROOT, "tinyedm_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "tinyedm_b200", fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn


def test_unet_builder_lengths():
    """The reference's tests/test_unet_builder.py:9-30 on this package's builders."""
    enc_t, dec_t = networks.get_encoder_blocks_types(), networks.get_decoder_blocks_types()
    enc_c, dec_c = networks.get_encoder_out_channels(), networks.get_decoder_out_channels()
    skips = networks.get_skip_connections()
    assert len(enc_t) == 15 and len(enc_c) == 15
    assert len(dec_t) == 21 and len(dec_c) == 21 and len(skips) == 21
    sc = networks.get_skip_channels(enc_c, dec_c, skips)
    assert len(sc) == 21
    assert sc == (0, 0, 768, 768, 768, 576, 0, 576, 576, 576, 384, 0, 384, 384, 384, 192, 0, 192, 192, 192, 192)
    assert enc_t == ("Enc",) * 3 + ("EncD", "Enc", "Enc", "Enc", "EncD", "EncA", "EncA", "EncA", "EncD", "EncA", "EncA", "EncA")
    assert dec_t[:7] == ("DecA", "Dec", "DecA", "DecA", "DecA", "DecA", "DecU") and dec_t[-5:] == ("DecU", "Dec", "Dec", "Dec", "Dec")


@pytest.mark.parametrize("cfg", [SMALL, cifar_cfg(num_classes=10), dict(O.MNIST), dict(O.IMAGENET)])
def test_state_dict_names_shapes_and_order_match_the_reference(cfg):
    den = T.Denoiser(**spec_kwargs(cfg["denoiser"]))
    ref = O.init_denoiser_params(cfg["denoiser"], torch.Generator().manual_seed(0))
    sd = den.state_dict()
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert sd[k].shape == v.shape and sd[k].dtype == torch.float32, k
    e = cfg["embedding"]
    emb = T.Embedding(e.fourier_dim, e.embedding_dim, e.num_classes, e.add_factor)
    assert set(emb.state_dict()) == set(O.init_embedding_params(e, torch.Generator().manual_seed(0)))
    # parameters() order: a module's own parameters come first, then its children in registration order
    names = [n for n, _ in den.named_parameters()]
    assert names[:3] == ["gain_out", "conv_in.weight", "conv_out.weight"]
    assert names[3] == "encoder_blocks.0.gain"
    first_dec_skip = next(i for i, s in enumerate(cfg["denoiser"].skip_connections) if s)
    blk = [n for n in names if n.startswith(f"decoder_blocks.{first_dec_skip}.")]
    assert blk[0].endswith(".gain") and blk[1].endswith("cat_factor.layer1.weight") and blk[2].endswith("cat_factor.layer2.weight")
    # constructor arguments mirrored as attributes (utils.deinstantiate, utils.py:15-25)
    for k, v in spec_kwargs(cfg["denoiser"]).items():
        assert getattr(den, k) == (tuple(v) if isinstance(v, (tuple, list)) else v), k


def test_imagenet_defaults_are_the_reference_defaults():
    """Denoiser() without arguments is the ImageNet-latent architecture (networks.py:332-432, 272.0 M parameters)."""
    from tinyedm_b200.configs import IMAGENET, build_edm
    m = build_edm(IMAGENET)
    s = O.IMAGENET["denoiser"]
    assert m.denoiser.encoder_block_types == s.encoder_block_types and m.denoiser.decoder_block_types == s.decoder_block_types
    assert m.denoiser.encoder_out_channels == s.encoder_out_channels and m.denoiser.decoder_out_channels == s.decoder_out_channels
    assert m.denoiser.skip_connections == s.skip_connections
    n_den = sum(p.numel() for p in m.denoiser.parameters())
    n_emb = sum(p.numel() for p in m.embedding.parameters())
    assert abs(n_den - 271.997e6) < 2e3 and abs(n_emb - 0.915e6) < 1e3, (n_den, n_emb)


def test_cifar_parameter_count():
    den = T.Denoiser(**spec_kwargs(O.CIFAR10["denoiser"]))
    assert sum(p.numel() for p in den.parameters()) == 35_603_990 or abs(sum(p.numel() for p in den.parameters()) - 35.6e6) < 5e4
    assert float(den.gain_out) == 0.0 and all(float(b.gain) == 1.0 for b in den.encoder_blocks)


def test_engine_plan_matches_the_reference_skip_pairing():
    """dec2<-enc7, dec3<-enc6, dec4<-enc5, dec6<-enc4, dec7<-enc3, dec8<-enc2, dec10<-enc1,
    dec11<-enc0, dec12<-conv_in (skip list index = encoder index + 1)."""
    den = T.Denoiser(**spec_kwargs(O.CIFAR10["denoiser"]))
    eng = den.engine
    pairing = {bp.name: bp.skip_src for bp in eng.blocks if bp.cskip > 0}
    assert pairing == {"decoder_blocks.2": 8, "decoder_blocks.3": 7, "decoder_blocks.4": 6, "decoder_blocks.6": 5,
                       "decoder_blocks.7": 4, "decoder_blocks.8": 3, "decoder_blocks.10": 2, "decoder_blocks.11": 1,
                       "decoder_blocks.12": 0}
    assert eng.n_mod == 21 * 256 and len(eng.bank.slots) == sum(1 for n, _ in den.named_parameters() if n.endswith("weight"))
    enc_o, dec_o = O.block_plan(O.CIFAR10["denoiser"])
    for bp, ref in zip(eng.blocks, enc_o + dec_o):
        assert (bp.cin, bp.cout) == (ref["cin"], ref["cout"]) and bp.attn == ref["attn"]
        assert bp.cskip == ref.get("cskip", 0)
    # the skip tensors whose ScaleLong mean is taken at production time are exactly the ones a decoder block pops
    assert eng._skips_consumed() == [True] * 9
    for cfg in (O.MNIST, O.IMAGENET):
        e2 = T.Denoiser(**spec_kwargs(cfg["denoiser"])).engine
        popped = {bp.skip_src for bp in e2.blocks if bp.cskip > 0}
        assert [i for i, u in enumerate(e2._skips_consumed()) if u] == sorted(popped)


def test_lr_schedule_and_ema_exponent():
    f = T.EDM.lr_factor
    assert f(0, 200, 200) == pytest.approx(1e-8) and f(100, 200, 200) == pytest.approx(0.5, rel=1e-6)
    assert f(200, 200, 200) == 1.0 and f(399, 200, 200) == 1.0
    assert f(400 + 600, 200, 200) == pytest.approx(0.5)
    g = T.sigma_rel_to_gamma(0.13)
    assert g == pytest.approx(4.6036, abs=2e-3)
    t = 0.13 ** -2
    assert abs(g ** 3 + 7 * g ** 2 + (16 - t) * g + 12 - t) < 1e-6       # ema.py:29-32
    with pytest.raises(ValueError):
        T.FusedAdamEMA([torch.nn.Parameter(torch.zeros(1))], ema_length=0.5)


def test_solver_schedule_matches_reference_formula():
    s = T.DeterministicSolver(num_steps=32)
    np.testing.assert_array_equal(s.t_steps.numpy(), O.t_schedule(32).numpy())
    assert len(s.t_steps) == 33 and float(s.t_steps[-1]) == 0.0


def test_edm_module_surface():
    cfg = SMALL
    den = T.Denoiser(**spec_kwargs(cfg["denoiser"]))
    emb = T.Embedding(32, 64, 5)
    with pytest.raises(ValueError):
        T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb, denoiser=den, use_ema=True, use_uncertainty=False,
              steady_steps=1, rampup_steps=1, scheduler_interval="step")
    m = T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb, denoiser=den, use_ema=True, use_uncertainty=True,
              steady_steps=200, rampup_steps=200, scheduler_interval="epoch", lr=0.02, ema_length=0.13)
    assert m.conditional and m.num_classes == 5 and m.sigma_data == 0.5
    names = [n for n, _ in m.named_parameters()]
    assert names[0] == "embedding.sigma_embed.weight" and names[1] == "embedding.class_embed.linear.weight"
    assert names[2] == "denoiser.gain_out" and names[-3:] == ["u.gain", "u.linear1.weight", "u.linear2.weight"]
    assert "embedding.fourier_embed.freqs" in m.state_dict() and "train_mse.total" not in m.state_dict()
    cfgd = m.configure_optimizers()
    assert isinstance(cfgd["optimizer"], T.FusedAdamEMA) and cfgd["lr_scheduler"]["interval"] == "epoch"
    assert cfgd["optimizer"].param_groups[0]["lr"] == pytest.approx(0.02 * 1e-8)   # LambdaLR applied factor(0)


def test_wgrad_width_padding_rule_matches_the_kernel_geometry():
    """networks.py `_wgrad_width` (stand-alone Conv2d) against the rule of csrc/conv_wgrad.cu `wgrad_geometry`: a pixel tile is
    rh whole rows of nb images, rh * W * nb a multiple of 16 and at most 128."""
    from tinyedm_b200.networks import _wgrad_width

    def tileable(H, W):
        return any((rh * W * nb) % 16 == 0 for rh in range(1, H + 1) for nb in range(1, 128 // (rh * W) + 1) if rh * W <= 128)

    for W in range(1, 129):
        Wp = _wgrad_width(W)
        assert W <= Wp <= 128 and tileable(1, Wp) and tileable(Wp, Wp)
        assert Wp == W or not tileable(W, W)                       # pads only what the kernel cannot tile ...
        assert all(not tileable(w, w) for w in range(W, Wp))       # ... and by as little as possible
    for W in (64, 32, 16, 8, 28, 14, 7):                           # every feature map of the three configs runs unpadded
        assert _wgrad_width(W) == W
