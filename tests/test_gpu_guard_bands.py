"""Out-of-bounds WRITE check without a sanitizer: every CUDA buffer the package allocates while the guard is active
(torch.empty / zeros / empty_like / zeros_like from tinyedm_b200's own modules: kernel outputs, workspaces, the flat weight
and gradient banks) is carved out of a larger allocation with a 64 KiB band of 0xA5 bytes on either side (the tail band
starts at the buffer's last byte, not at the allocator's 512-byte rounding). After whole training steps, eval forwards and a
Heun solve on RAGGED shapes (odd batch, 7x7 / 14x14 feature maps, 192-wide layers, attention from head_dim 16 at S = 196 to
head_dim 64 at S = 49), and after the attention / diffuser / stand-alone conv kernels called directly on further shapes,
every band must still be intact — a kernel that stores one element past (or before) any of its outputs fails here. Results are compared with the oracle in the other test files;
this file only checks where the kernels write."""
import dataclasses
import math

import pytest
import torch

from tests.helpers import SMALL, build_modules, seeded_params, small_params

pytestmark = pytest.mark.gpu

GUARD = 1 << 16
PATTERN = 0xA5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from tinyedm_b200 import ops
    d = torch.device("cuda:0")
    ops.ensure_device(d)
    return d


class GuardedAllocs:
    """Context manager: routes the package's CUDA allocations through guarded buffers (see the module docstring)."""

    def __init__(self):
        self.tracked = []            # (whole uint8 buffer, payload bytes, what)
        self._real = {}

    # -- allocation ------------------------------------------------------------------------------------------
    def _carve(self, shape, dtype, device, zero, what):
        shape = tuple(int(s) for s in (shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else shape))
        n = math.prod(shape) if shape else 1
        nbytes = n * dtype.itemsize
        buf = self._real["empty"](2 * GUARD + nbytes, dtype=torch.uint8, device=device)
        buf.fill_(PATTERN)
        self.tracked.append((buf, nbytes, what))
        if nbytes == 0:
            return self._real["empty"](shape, dtype=dtype, device=device)
        view = buf[GUARD:GUARD + nbytes].view(dtype).view(shape)
        if zero:
            view.zero_()
        return view

    @staticmethod
    def _is_cuda(device):
        return device is not None and torch.device(device).type == "cuda"

    def _wrap_new(self, name, zero):
        real = self._real[name]

        def alloc(*shape, dtype=None, device=None, **kw):
            if not self._is_cuda(device) or not shape or any(kw.get(k) for k in ("pin_memory", "requires_grad", "size", "layout")) or \
                    kw.get("out") is not None or kw.get("memory_format") not in (None, torch.contiguous_format):
                return real(*shape, dtype=dtype, device=device, **kw)
            return self._carve(shape, dtype or torch.get_default_dtype(), device, zero, name)
        return alloc

    def _wrap_like(self, name, zero):
        real = self._real[name]

        def alloc(t, dtype=None, device=None, **kw):
            device = device if device is not None else t.device
            if not self._is_cuda(device) or kw or not t.is_contiguous():
                return real(t, dtype=dtype, device=device, **kw)
            return self._carve(tuple(t.shape), dtype or t.dtype, device, zero, name)
        return alloc

    def __enter__(self):
        for name in ("empty", "zeros", "empty_like", "zeros_like"):
            self._real[name] = getattr(torch, name)
        torch.empty = self._wrap_new("empty", False)
        torch.zeros = self._wrap_new("zeros", True)
        torch.empty_like = self._wrap_like("empty_like", False)
        torch.zeros_like = self._wrap_like("zeros_like", True)
        return self

    def __exit__(self, *exc):
        for name, fn in self._real.items():
            setattr(torch, name, fn)

    # -- verdict ---------------------------------------------------------------------------------------------
    def violations(self):
        torch.cuda.synchronize()
        bad = []
        for buf, nbytes, what in self.tracked:
            head, tail = buf[:GUARD], buf[GUARD + nbytes:]
            for side, band in (("before", head), ("after", tail)):
                hit = (band != PATTERN).nonzero()
                if hit.numel():
                    first = int(hit[0]) if side == "after" else int(hit[-1]) - GUARD
                    bad.append(f"{what} of {nbytes} B: {hit.numel()} guard bytes overwritten {side} the buffer (nearest at offset {first:+d})")
        return bad


def test_the_guard_itself_sees_a_one_element_overrun(dev):
    with GuardedAllocs() as g:
        t = torch.empty(3, 5, device=dev, dtype=torch.bfloat16)
        z = torch.zeros(7, device=dev)
        assert t.shape == (3, 5) and t.is_contiguous() and float(z.abs().sum()) == 0.0
        assert g.violations() == []
        torch.as_strided(t, (16,), (1,)).fill_(1.0)          # one bf16 element past the end
    v = g.violations()
    assert len(v) == 1 and "after" in v[0] and "2 guard bytes" in v[0], v
    assert torch.empty is g._real["empty"]                   # the patch is gone


def _ragged_cfg():
    """A small net whose shapes are as awkward as the configs get: 14x14 -> 7x7 maps, 64 / 192 channels (N = 192 tiles, the
    transposed weight-gradient kernel), attention with head_dim 64 at S = 49 (packed generic kernels)."""
    den = dataclasses.replace(
        SMALL["denoiser"], encoder_out_channels=(64, 192, 192), decoder_out_channels=(192, 192, 192, 64, 64, 64), num_heads=3,
        dropout_rate=0.1)
    return dict(denoiser=den, embedding=SMALL["embedding"], image=(3, 14, 14), batch=5, seed=77)


def _model(dev, cfg, params, use_uncertainty=True):
    import tinyedm_b200 as T
    dp, ep, up = params
    den, emb, unc = build_modules(cfg, dp, ep, up, dev)
    m = T.EDM(diffuser=T.Diffuser(-1.2, 1.2), embedding=emb, denoiser=den, use_ema=True, ema_length=0.1, use_uncertainty=use_uncertainty,
              steady_steps=1, rampup_steps=1, scheduler_interval="step")
    if use_uncertainty:
        m.u = unc
    return m.to(dev)


@pytest.mark.parametrize("which", ["small_16x16", "ragged_14x14_192ch", "attention_at_full_resolution"])
def test_whole_steps_write_only_inside_their_buffers(dev, which, monkeypatch):
    import tinyedm_b200 as T
    monkeypatch.setenv("TEDM_CUDA_GRAPHS", "0")               # graph-private pools would bypass the guarded allocations
    if which == "small_16x16":
        cfg, params = SMALL, small_params()
    elif which == "ragged_14x14_192ch":
        cfg = _ragged_cfg(); params = seeded_params(cfg, seed=5)
    else:       # every block attends: head_dim 16 at S = 196 and head_dim 32 at S = 49, batch 2
        den = dataclasses.replace(SMALL["denoiser"], encoder_block_types=("EncA", "EncD", "EncA"),
                                  decoder_block_types=("DecA", "DecA", "DecA", "DecU", "DecA", "DecA"))
        cfg = dict(denoiser=den, embedding=SMALL["embedding"], image=(3, 14, 14), batch=2, seed=78)
        params = seeded_params(cfg, seed=6)
    B = cfg["batch"]
    gen = torch.Generator().manual_seed(cfg["seed"])
    clean = (0.5 * torch.randn(B, *cfg["image"], generator=gen)).clamp(-1, 1).to(dev)
    labels = torch.randint(0, cfg["embedding"].num_classes, (B,), generator=gen).to(dev)
    x0 = torch.randn(B, *cfg["image"], generator=gen).to(dev)
    with GuardedAllocs() as g:
        model = _model(dev, cfg, params)
        opt = model.configure_optimizers()["optimizer"]
        model.train()
        for it in range(2):                                   # two optimiser steps, the second with accumulation
            opt.zero_grad(set_to_none=True)
            if it == 1:
                with model.denoiser.accumulate_grads():
                    (model.training_step((clean, labels), 0) / 2).backward()
                (model.training_step((clean.flip(0), labels), 1) / 2).backward()
            else:
                model.training_step((clean, labels), 0).backward()
            opt.step()
        model.eval()
        with torch.no_grad():
            sigma = torch.rand(B, 1, 1, 1, device=dev) + 0.2
            D = model(clean, sigma, labels)
            imgs = T.DeterministicSolver(num_steps=3).solve(model, x0, labels)
        assert torch.isfinite(D).all() and torch.isfinite(imgs).all()
        bad = g.violations()
    assert len(g.tracked) > 100, len(g.tracked)               # the guard really was in the allocation path
    assert bad == [], "\n".join(bad[:20])


def test_standalone_kernels_write_only_inside_their_buffers(dev):
    """The round-2 kernels called directly on shapes the nets above do not reach: generic attention at head_dim 128 / 144 /
    192 / ragged, ScaleLong, the fused Philox diffuser with its im2col operand, the padded stand-alone Conv2d."""
    import tinyedm_b200 as T
    from tinyedm_b200 import ops
    torch.manual_seed(3)
    with GuardedAllocs() as g:
        for (B, H, W, heads, hd) in [(2, 7, 7, 4, 128), (3, 16, 16, 2, 144), (3, 8, 8, 2, 192), (1, 9, 9, 1, 80), (2, 13, 11, 2, 48),
                                     (3, 1, 1, 4, 64), (5, 14, 14, 4, 64), (3, 5, 5, 3, 16)]:
            C = heads * hd
            qkv = (torch.randn(B, H, W, 3 * C, device=dev) * 1.3).to(torch.bfloat16)
            qn, norms = ops.qkv_normalize(qkv, heads)
            y, lse = ops.attention_forward_normalized(qn, heads, True)
            ops.attention_backward_normalized(qn, norms, y, torch.randn_like(y.float()).to(torch.bfloat16), lse, heads)
        for (B, H, W, heads) in [(3, 8, 8, 4), (2, 16, 16, 2), (5, 8, 8, 1)]:
            qkv = torch.randn(B, H, W, 3 * heads * 64, device=dev).to(torch.bfloat16)
            y, lse = ops.attention_forward(qkv, heads, True)
            ops.attention_backward(qkv, y, torch.randn_like(y.float()).to(torch.bfloat16), lse, heads)
        for (B, C, H, W) in [(5, 3, 32, 32), (3, 1, 28, 28), (2, 4, 64, 64), (7, 3, 6, 10)]:
            d = T.Diffuser(-1.2, 1.2)
            d._fuse_sigma_data = 0.5
            noisy, sigma = d(torch.randn(B, C, H, W, device=dev))
            assert noisy.shape == (B, C, H, W)
        for (cin, cout, k, hw) in [(3, 40, 3, 9), (70, 130, 1, 5), (64, 64, 3, 7)]:
            conv = T.networks.Conv2d(cin, cout, k).to(dev).train()
            x = torch.randn(3, cin, hw, hw, device=dev, requires_grad=True)
            conv(x).square().mean().backward()
        bad = g.violations()
    assert len(g.tracked) > 40
    assert bad == [], "\n".join(bad[:20])


@pytest.mark.parametrize("name,B", [("cifar", 256), ("mnist", 128), ("imagenet", 16)])
def test_full_size_steps_write_only_inside_their_buffers(dev, name, B, monkeypatch):
    """The three real architectures at their own feature-map sizes and (CIFAR, MNIST) batch sizes: one training step with
    dropout and the optimiser, one eval forward, under the guard."""
    from tinyedm_b200 import configs
    monkeypatch.setenv("TEDM_CUDA_GRAPHS", "0")
    cfg = {"cifar": configs.CIFAR10, "mnist": configs.MNIST, "imagenet": configs.IMAGENET}[name]
    C, H, W = {"cifar": (3, 32, 32), "mnist": (1, 28, 28), "imagenet": (4, 64, 64)}[name]
    torch.manual_seed(11)
    clean = (0.5 * torch.randn(B, C, H, W, device=dev)).clamp(-1, 1)
    with GuardedAllocs() as g:
        model = configs.build_edm(cfg).to(dev)
        ncls = model.embedding.num_classes
        ncls = ncls if ncls and ncls > 0 else None
        labels = torch.randint(0, ncls, (B,), device=dev) if ncls else torch.zeros(B, dtype=torch.long, device=dev)
        with torch.no_grad():
            model.denoiser.gain_out.fill_(1.0)
        opt = model.configure_optimizers()["optimizer"]
        model.train()
        loss = model.training_step((clean, labels), 0)
        loss.backward()
        opt.step()
        model.eval()
        with torch.no_grad():
            D = model(clean, torch.rand(B, 1, 1, 1, device=dev) + 0.2, labels if ncls else None)
        assert torch.isfinite(loss) and torch.isfinite(D).all()
        bad = g.violations()
        n = len(g.tracked)
        del model, opt, loss, D
    g.tracked.clear()
    torch.cuda.empty_cache()
    assert n > 300, n
    assert bad == [], "\n".join(bad[:20])
