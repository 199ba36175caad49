"""Drop-in module surface of the reference's `tinyedm.networks` on top of the B200 engine.

What is kept EXACTLY (so that Hydra `_target_` configs, `deinstantiate`, checkpoints and the EMA's positional
parameter zip keep working — SURVEY.md §8b): class names, constructor keyword names mirrored as attributes,
the module tree / registration order and therefore `state_dict()` keys, shapes, dtypes (fp32) and
`parameters()` order. Reference: /root/reference/src/tinyedm/networks.py (cited per class below).

What changes: no module here computes with torch ops. `Denoiser.forward`, `Embedding.forward`,
`UncertaintyNet.forward`, `Conv2d.forward` and `Linear.forward` hand raw device pointers to
libtinyedm_b200.so (hand-written sm_100a kernels) through `torch.autograd.Function`s; the inner block
modules only OWN parameters — their arithmetic runs fused inside the Denoiser plan (engine.py). There is no
CPU path: CPU tensors raise.
"""
from __future__ import annotations

import contextlib
import math

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .engine import DenoiserEngine, WeightBank, conv_slot, f32_slot
from .ops import BF16, F32


# ---------------------------------------------------------------------------------------------------------
# weight holders
# ---------------------------------------------------------------------------------------------------------
class _FusedOnly(nn.Module):
    """Parameter holder whose arithmetic is executed by the enclosing Denoiser's fused plan. Forward hooks registered on
    an EncoderBlock / DecoderBlock are fired by the engine with (input, embedding[, skip]) and the block's output
    (engine.DenoiserEngine.forward); the finer-grained holders (attention, ScaleLong, resamplers) have no separate
    activations to show — their arithmetic lives in epilogues."""

    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__} runs fused inside tinyedm_b200.Denoiser (engine.py); call the Denoiser instead")


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, mod):
        y, saved = mod._run(x)
        ctx.mod, ctx.saved = mod, saved
        return y

    @staticmethod
    def backward(ctx, gy):
        gx, gw = ctx.mod._run_backward(ctx.saved, gy, ctx.needs_input_grad[0])
        return gx, gw, None


class Conv2d(nn.Module):
    """MPConv: force-normalised stride-1 'same' convolution, no bias (networks.py:22-43).

    Inside a Denoiser it is only a weight holder. Called on its own (it is part of the reference's public exports) it runs
    the same tensor-core kernel on an NCHW tensor; the NCHW<->NHWC relayout is torch plumbing. Channel counts that are
    not multiples of 64 (e.g. ScaleLong's 257 -> 16, networks.py:109-110) are zero-padded to the next multiple around the
    kernel: the weight bank then emits the fp32 w_hat in parameter layout and the padded bf16 operands are laid out here.
    """

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size))
        self._bank: WeightBank | None = None

    @property
    def _padded(self) -> bool:
        return self.in_channels % 64 != 0 or self.out_channels % 64 != 0

    def _prepared(self, dev):
        if self.kernel_size not in (1, 3):
            raise RuntimeError("tinyedm_b200.Conv2d (standalone): kernel size 1 or 3 is required")
        ops.ensure_device(dev)
        if self._bank is None or self._bank.slots[0].param is not self.weight:
            self._bank = WeightBank([f32_slot("weight", self.weight) if self._padded else conv_slot("weight", self.weight)])
        if self._bank.device != dev:
            self._bank.materialise(dev)
        self._bank.prepare(self.training)
        return self._bank.slots[0]

    def _padded_operands(self, slot):
        """bf16 forward [Cout_p][tap][Cin_p] and data-gradient [Cin_p][tap flipped][Cout_p] operands from the fp32 w_hat."""
        k, ci, co = self.kernel_size, self.in_channels, self.out_channels
        cip, cop = (ci + 63) // 64 * 64, (co + 63) // 64 * 64
        w_hat = slot.f32.view(co, ci, k, k)
        wp = torch.zeros((cop, cip, k, k), device=w_hat.device, dtype=F32)
        wp[:co, :ci] = w_hat
        fwd = wp.permute(0, 2, 3, 1).reshape(cop, k * k * cip).to(BF16).contiguous()
        dgrad = wp.flip(2, 3).permute(1, 2, 3, 0).reshape(cip, k * k * cop).to(BF16).contiguous()
        return fwd, dgrad, cip, cop

    def _run(self, x: Tensor):
        slot = self._prepared(x.device)
        out_dtype = x.dtype if x.dtype in (BF16, torch.float16) else BF16
        xh = x.permute(0, 2, 3, 1).to(BF16)
        if not self._padded:
            xh = xh.contiguous()
            y = ops.conv2d(xh, slot.fwd, self.kernel_size, self.out_channels)
            return y.permute(0, 3, 1, 2).to(out_dtype), (xh, x.dtype, None)
        fwd, dgrad, cip, cop = self._padded_operands(slot)
        xp = torch.zeros(xh.shape[:3] + (cip,), device=x.device, dtype=BF16)
        xp[..., :self.in_channels] = xh
        y = ops.conv2d(xp, fwd, self.kernel_size, cop)[..., :self.out_channels]
        return y.permute(0, 3, 1, 2).to(out_dtype), (xp, x.dtype, (dgrad, cip, cop))

    def _run_backward(self, saved, gy: Tensor, need_gx: bool):
        xh, in_dtype, pad = saved
        bank = self._bank
        bank.begin_backward()
        slot = bank.slots[0]
        k, ci, co = self.kernel_size, self.in_channels, self.out_channels
        gh = gy.permute(0, 2, 3, 1).to(BF16)
        Wp = _wgrad_width(gh.shape[2])
        if pad is None:
            gh = gh.contiguous()
            ops.conv2d_wgrad(_pad_width(gh, Wp), _pad_width(xh, Wp), slot.ghat, k)
            bank.backward()
            gx = None
            if need_gx:
                gx = ops.conv2d(gh, slot.dgrad, k, ci).permute(0, 3, 1, 2).to(in_dtype)
            return gx, bank.autograd_grads()[0]
        dgrad, cip, cop = pad
        gp = torch.zeros(gh.shape[:3] + (cop,), device=gy.device, dtype=BF16)
        gp[..., :co] = gh
        dw = torch.zeros((cop, k * k, cip), device=gy.device, dtype=F32)
        ops.conv2d_wgrad(_pad_width(gp, Wp), _pad_width(xh, Wp), dw, k)
        # dL/dw_hat back in parameter layout (Cout, Cin, k, k) for the weight-norm Jacobian
        slot.ghat.copy_(dw.view(cop, k, k, cip)[:co, :, :, :ci].permute(0, 3, 1, 2).reshape(co, ci * k * k))
        bank.backward()
        gx = None
        if need_gx:
            gx = ops.conv2d(gp, dgrad, k, cip)[..., :ci].permute(0, 3, 1, 2).to(in_dtype)
        return gx, bank.autograd_grads()[0]

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("tinyedm_b200.Conv2d runs on CUDA (sm_100a) only; there is no CPU fallback")
        return _ConvFn.apply(x, self.weight, self)

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}"


def _wgrad_width(W: int) -> int:
    """Smallest width >= W the weight-gradient kernels can tile: their pixel tiles are whole rows, a multiple of 16 pixels
    and at most 128 (csrc/conv_wgrad.cu `wgrad_geometry`), i.e. lcm(16, W') <= 128. Every feature map of the three
    configs qualifies as it is (64, 32, 16, 8, 28, 14, 7); 9, 13, 17, 18 ... do not."""
    w = W
    while w < 128 and math.lcm(16, w) > 128:
        w += 1
    return w


def _pad_width(t: Tensor, Wp: int) -> Tensor:
    """(B, H, W, C) -> (B, H, Wp, C) with zero columns on the right. Exact for the weight gradient: the new output-gradient
    columns are zero, and the new input columns are the zeros the convolution's own padding already stands for."""
    if t.shape[2] == Wp:
        return t
    out = torch.zeros((t.shape[0], t.shape[1], Wp, t.shape[3]), device=t.device, dtype=t.dtype)
    out[:, :, :t.shape[2]] = t
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, mod):
        slot = mod._prepared(x.device)
        x2 = ops.check(x.reshape(-1, mod.in_features).float().contiguous(), F32, "x")
        y = torch.empty((x2.shape[0], mod.out_features), device=x.device, dtype=F32)
        ops.sgemm(x2, slot.f32, y, x2.shape[0], mod.out_features, mod.in_features, mod.in_features, mod.in_features,
                  mod.out_features, False, True)
        ctx.mod, ctx.x2, ctx.shape = mod, x2, x.shape
        return y.reshape(*x.shape[:-1], mod.out_features)

    @staticmethod
    def backward(ctx, gy):
        mod, x2 = ctx.mod, ctx.x2
        bank = mod._bank
        bank.begin_backward()
        slot = bank.slots[0]
        g2 = gy.reshape(-1, mod.out_features).float().contiguous()
        M = x2.shape[0]
        ops.sgemm(g2, x2, slot.ghat, mod.out_features, mod.in_features, M, mod.out_features, mod.in_features,
                  mod.in_features, True, False)
        bank.backward()
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x2)
            ops.sgemm(g2, slot.f32, gx, M, mod.in_features, mod.out_features, mod.out_features, mod.in_features,
                      mod.in_features, False, False)
            gx = gx.reshape(ctx.shape)
        return gx, bank.autograd_grads()[0], None


class Linear(nn.Module):
    """Force-normalised fp32 linear layer, no bias (networks.py:46-64)."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.randn(out_features, in_features))
        self._bank: WeightBank | None = None

    def _prepared(self, dev):
        ops.ensure_device(dev)
        if self._bank is None or self._bank.slots[0].param is not self.weight:
            self._bank = WeightBank([f32_slot("weight", self.weight)])
        if self._bank.device != dev:
            self._bank.materialise(dev)
        self._bank.prepare(self.training)
        return self._bank.slots[0]

    def forward(self, x: Tensor):
        return _LinearFn.apply(x, self.weight, self)

    def extra_repr(self) -> str:
        return f"{self.in_features}, {self.out_features}"


class UpSample(_FusedOnly):
    """Nearest-exact x2 (networks.py:67-72): folded into block_prep's loader."""


class DownSample(_FusedOnly):
    """2x2 average pool (networks.py:75-80): folded into block_prep's loader."""


# ---------------------------------------------------------------------------------------------------------
# heads and embeddings
# ---------------------------------------------------------------------------------------------------------
class _UncertaintyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fourier, w1, w2, gain, mod):
        bank = mod._bank_for(fourier.device)
        bank.prepare(mod.training)
        s1, s2 = bank.slots
        f = ops.check(fourier.float().contiguous(), F32, "fourier")
        aug, h_pre, h, u_raw, u = ops.uncertainty_forward(f, s1.f32, s2.f32, gain)
        ctx.mod = mod
        ctx.saved = (aug, h_pre, h, u_raw, gain)
        return u.view(-1, 1)

    @staticmethod
    def backward(ctx, g_u):
        mod = ctx.mod
        aug, h_pre, h, u_raw, gain = ctx.saved
        bank = mod._bank
        bank.begin_backward()
        if mod._g_gain is None or mod.gain.grad is not None or mod._g_gain.device != g_u.device:
            mod._g_gain = torch.empty((1,), device=g_u.device, dtype=F32)
        s1, s2 = bank.slots
        B, F_ = h_pre.shape
        g_u = g_u.reshape(-1).float().contiguous()
        g_uraw, g_hpre = ops.uncertainty_backward(g_u, gain, s2.f32, h_pre)
        ops.sgemm(g_uraw, h, s2.ghat, 1, F_, B, 1, F_, F_, True, False)
        ops.sgemm(g_hpre, aug, s1.ghat, F_, F_ + 1, B, F_, F_ + 1, F_ + 1, True, False)
        g_gain = mod._g_gain
        ops.sgemm(g_u, u_raw, g_gain, 1, 1, B, 1, 1, 1, True, False)
        bank.backward()
        g1, g2 = bank.autograd_grads()
        # fourier features carry no learnable dependency upstream (freqs/phases are buffers, sigma is data)
        return None, g1, g2, g_gain.view(()), None


class UncertaintyNet(nn.Module):
    """Learned log-variance head u(sigma) (networks.py:91-103)."""

    def __init__(self, in_features: int, hidden_features: int):
        super().__init__()
        if in_features != hidden_features:
            raise RuntimeError("tinyedm_b200.UncertaintyNet: in_features must equal hidden_features (as in edm.py:142-146)")
        self.linear1 = Linear(in_features + 1, hidden_features)
        self.linear2 = Linear(hidden_features, 1)
        self.gain = nn.Parameter(torch.zeros(()))
        self._bank: WeightBank | None = None
        self._g_gain: Tensor | None = None

    def _bank_for(self, dev):
        ops.ensure_device(dev)
        if self._bank is None or self._bank.slots[0].param is not self.linear1.weight:
            self._bank = WeightBank([f32_slot("linear1.weight", self.linear1.weight), f32_slot("linear2.weight", self.linear2.weight)])
        if self._bank.device != dev:
            self._bank.materialise(dev)
        return self._bank

    def forward(self, x: Tensor):
        return _UncertaintyFn.apply(x, self.linear1.weight, self.linear2.weight, self.gain, self)


class ScaleLong(_FusedOnly):
    """Learned squeeze-excite skip gain (networks.py:106-118); runs inside the decoder plan."""

    def __init__(self, dim, r=16):
        super().__init__()
        self.layer1 = Conv2d(dim + 1, int(dim // r), 1)
        self.layer2 = Conv2d(int(dim // r), dim, 1)


class ClassEmbedding(_FusedOnly):
    """one_hot(label)*sqrt(N) -> Linear, i.e. a scaled column gather of w_hat (networks.py:121-129)."""

    def __init__(self, num_embeddings, embedding_dim):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.linear = Linear(num_embeddings, embedding_dim)


class FourierEmbedding(_FusedOnly):
    """cos(outer(c_noise, freqs) + phases) * sqrt(2) with frozen random freqs/phases (networks.py:132-141)."""

    def __init__(self, embedding_dim: int):
        super().__init__()
        self.register_buffer("freqs", 2 * np.pi * torch.randn(embedding_dim))
        self.register_buffer("phases", 2 * np.pi * torch.rand(embedding_dim))


class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma, labels, w_sigma, w_class, mod):
        dev = sigma.device
        bank = mod._bank_for(dev)
        bank.prepare(mod.training)
        s_sig = bank.slots[0]
        s_cls = bank.slots[1] if len(bank.slots) > 1 else None
        sig = ops.check(sigma.reshape(-1).float().contiguous(), F32, "sigmas")
        lab = None
        rows = sig.numel()
        if labels is not None:
            lab = ops.check(labels.reshape(-1).to(torch.int64).contiguous(), torch.int64, "class_labels")
            if rows not in (1, lab.numel()):
                raise RuntimeError("tinyedm_b200.Embedding: sigmas and class_labels disagree on the batch size")
            rows = lab.numel()
        fourier, pre, emb = ops.embedding_forward(sig, mod.fourier_embed.freqs, mod.fourier_embed.phases, s_sig.f32,
                                                  s_cls.f32 if (s_cls is not None and lab is not None) else None, lab,
                                                  rows, mod.num_classes or 0, float(mod.add_factor))
        if sig.numel() == 1 and lab is not None:
            fourier = fourier[:1]  # the reference's fourier embedding has one row for a 0-d sigma (networks.py:139)
        ctx.mod, ctx.saved = mod, (fourier, pre, lab, sig.numel())
        ctx.mark_non_differentiable(fourier)
        return fourier, emb

    @staticmethod
    def backward(ctx, _g_fourier, g_emb):
        mod = ctx.mod
        fourier, pre, lab, n_sigma = ctx.saved
        bank = mod._bank
        bank.begin_backward()
        s_sig = bank.slots[0]
        s_cls = bank.slots[1] if len(bank.slots) > 1 else None
        B, E = pre.shape
        if n_sigma != B:
            raise RuntimeError("tinyedm_b200.Embedding: backward needs one sigma per sample")
        g_w_class = None
        if s_cls is not None:
            s_cls.ghat.zero_()
            g_w_class = s_cls.ghat if lab is not None else None
        g_sig = ops.embedding_backward(g_emb.float().contiguous(), pre, lab, g_w_class, mod.num_classes or 0,
                                       float(mod.add_factor))
        F_ = fourier.shape[1]
        ops.sgemm(g_sig, fourier, s_sig.ghat, E, F_, B, E, F_, F_, True, False)
        bank.backward()
        grads = bank.autograd_grads()
        return None, None, grads[0], (grads[1] if s_cls is not None else None), None


class Embedding(nn.Module):
    """Noise (+ class) conditioning: c_noise = ln(sigma)/4 -> Fourier features -> Linear [-> mp_add with the class
    embedding] -> mp_silu, all in fp32 (networks.py:144-178). Returns `(fourier_embedding, embedding)`.
    One fused kernel (tedm_embedding_forward); the autocast-off island of the reference is implicit."""

    def __init__(self, fourier_dim: int, embedding_dim: int, num_classes: int | None = None, add_factor: float = 0.5):
        super().__init__()
        self.fourier_dim = fourier_dim
        self.add_factor = add_factor
        self.embedding_dim = embedding_dim
        self.num_classes = num_classes
        self.fourier_embed = FourierEmbedding(fourier_dim)
        self.sigma_embed = Linear(fourier_dim, embedding_dim)
        self.class_embed = None
        if num_classes is not None and num_classes != -1:
            self.class_embed = ClassEmbedding(num_classes, embedding_dim)
        self._bank: WeightBank | None = None

    def _bank_for(self, dev):
        ops.ensure_device(dev)
        if self._bank is None or self._bank.slots[0].param is not self.sigma_embed.weight:
            slots = [f32_slot("sigma_embed.weight", self.sigma_embed.weight)]
            if self.class_embed is not None:
                slots.append(f32_slot("class_embed.linear.weight", self.class_embed.linear.weight))
            self._bank = WeightBank(slots)
        if self._bank.device != dev:
            self._bank.materialise(dev)
        return self._bank

    def prepare_weights(self, device) -> None:
        """(Re)normalises the weights if a parameter changed since the last call (eval-mode cache refresh)."""
        self._bank_for(torch.device(device)).prepare(self.training)

    def forward(self, sigmas, class_labels=None):
        if class_labels is not None and self.class_embed is None:
            raise ValueError("class_labels is not None, but num_classes is None. ")
        w_class = self.class_embed.linear.weight if self.class_embed is not None else None
        return _EmbeddingFn.apply(sigmas, class_labels, self.sigma_embed.weight, w_class, self)


# ---------------------------------------------------------------------------------------------------------
# blocks (parameter owners; executed by engine.DenoiserEngine)
# ---------------------------------------------------------------------------------------------------------
class CosineAttention(_FusedOnly):
    """qkv 1x1 conv -> pixel-normalised q,k,v -> SDPA -> out conv -> mp_add(x, y, 0.5) (networks.py:181-207)."""

    def __init__(self, embedding_dim: int, num_heads):
        super().__init__()
        assert embedding_dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = embedding_dim // num_heads
        self.embedding_dim = embedding_dim
        self.qkv_conv = Conv2d(embedding_dim, 3 * embedding_dim, 1)
        self.out_conv = Conv2d(embedding_dim, embedding_dim, 1)


def _residual_core(self, width_in: int, out_channels: int, embedding_dim: int, attention: bool, num_heads: int,
                   dropout_rate: float) -> None:
    """Members shared by both block flavours, registered in the reference's order."""
    self.conv_3x3_1 = Conv2d(width_in, out_channels, 3)
    self.conv_3x3_2 = Conv2d(out_channels, out_channels, 3)
    self.dropout = nn.Dropout(dropout_rate)
    self.attention = CosineAttention(out_channels, num_heads) if attention else nn.Identity()
    self.embed = Linear(embedding_dim, out_channels)
    self.gain = nn.Parameter(torch.ones(()))


class EncoderBlock(_FusedOnly):
    """resample -> conv_1x1 -> pixel_norm -> [mp_silu -> conv3x3 -> modulation -> mp_silu -> dropout -> conv3x3]
    -> mp_add(x, res, t) -> attention (networks.py:210-265)."""

    def __init__(self, in_channels: int, out_channels: int, embedding_dim: int, down: bool, attention: bool,
                 num_heads: int = 4, dropout_rate: float = 0.0, add_factor: float = 0.3):
        super().__init__()
        self.dropout_rate = dropout_rate
        self.add_factor = add_factor
        self.resample = DownSample() if down else nn.Identity()
        self.conv_1x1 = Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()
        _residual_core(self, out_channels, out_channels, embedding_dim, attention, num_heads, dropout_rate)


class DecoderBlock(_FusedOnly):
    """[cat(input, skip*ScaleLong(skip))] -> resample -> conv_1x1 | residual branch as in the encoder (no
    pixel_norm) -> mp_add -> attention (networks.py:268-329)."""

    def __init__(self, in_channels: int, out_channels: int, embedding_dim: int, up: bool, attention: bool,
                 num_heads: int = 4, skip_channels: int = 0, dropout_rate: float = 0.0, add_factor: float = 0.3):
        super().__init__()
        self.add_factor = add_factor
        self.cat_factor = ScaleLong(skip_channels) if skip_channels > 0 else None
        self.resample = UpSample() if up else nn.Identity()
        width = in_channels + skip_channels
        self.conv_1x1 = Conv2d(width, out_channels, 1) if width != out_channels else nn.Identity()
        _residual_core(self, width, out_channels, embedding_dim, attention, num_heads, dropout_rate)


# ---------------------------------------------------------------------------------------------------------
# architecture defaults (the ImageNet-latent arch, networks.py:332-432) and builders (:435-487)
# ---------------------------------------------------------------------------------------------------------
def _level(kind: str, n: int, first: str = "") -> tuple[str, ...]:
    return ((first,) if first else ()) + (kind,) * n


def get_encoder_blocks_types() -> tuple[str, ...]:
    return _level("Enc", 3) + _level("Enc", 3, "EncD") + _level("EncA", 3, "EncD") + _level("EncA", 3, "EncD")


def get_decoder_blocks_types() -> tuple[str, ...]:
    return (("DecA", "Dec") + _level("DecA", 4) + _level("DecA", 4, "DecU") + _level("Dec", 4, "DecU")
            + _level("Dec", 4, "DecU"))


def get_encoder_out_channels() -> tuple[int, ...]:
    return (192,) * 4 + (384,) * 4 + (576,) * 4 + (768,) * 3


def get_decoder_out_channels() -> tuple[int, ...]:
    return (768,) * 6 + (576,) * 5 + (384,) * 6 + (192,) * 4


def get_skip_connections() -> tuple[bool, ...]:
    """Which decoder blocks take a skip connection."""
    return (False, False) + (True,) * 4 + ((False,) + (True,) * 4) * 3


def get_skip_channels(encoder_out_channels, decoder_out_channels, skip_connections) -> tuple[int, ...]:
    """Skip widths scattered into the decoder slots that take one: encoder outputs deepest-first, then the
    input block (networks.py:435-444)."""
    feed = iter(list(encoder_out_channels[::-1]) + [encoder_out_channels[0]])
    n_true = sum(bool(s) for s in skip_connections)
    if n_true != len(encoder_out_channels) + 1:
        raise ValueError(f"skip_connections must mark exactly {len(encoder_out_channels) + 1} decoder blocks, got {n_true}")
    return tuple(int(next(feed)) if has else 0 for has, _ in zip(skip_connections, decoder_out_channels))


def build_encoder_blocks(block_types, out_channels, **kwargs):
    blocks, width = nn.ModuleList(), out_channels[0]
    for kind, cout in zip(block_types, out_channels):
        blocks.append(EncoderBlock(in_channels=width, out_channels=cout, down=kind.endswith("D"),
                                   attention=kind.endswith("A"), **kwargs))
        width = cout
    return blocks


def build_decoder_blocks(block_types, out_channels, skip_channels, **kwargs):
    blocks, width = nn.ModuleList(), out_channels[0]
    for kind, cout, cskip in zip(block_types, out_channels, skip_channels):
        blocks.append(DecoderBlock(in_channels=width, out_channels=cout, skip_channels=cskip, up=kind.endswith("U"),
                                   attention=kind.endswith("A"), **kwargs))
        width = cout
    return blocks


# ---------------------------------------------------------------------------------------------------------
# Denoiser
# ---------------------------------------------------------------------------------------------------------
class _DenoiserFn(torch.autograd.Function):
    """One autograd node for the whole preconditioned U-Net (forward plan + hand-derived adjoint)."""

    @staticmethod
    def forward(ctx, noisy, sigma, emb, mod, *params):
        D, saved = mod.engine.forward(noisy, sigma, emb, training=mod.training, save=True)
        ctx.mod, ctx.saved, ctx.params = mod, saved, params
        return D

    @staticmethod
    def backward(ctx, g_D):
        eng = ctx.mod.engine
        if ctx.saved is None:
            raise RuntimeError("tinyedm_b200.Denoiser: backward through the same forward twice is not supported")
        eng.begin_backward()     # gradient buffers handed out earlier and still alive are not overwritten
        g_emb, sg = eng.backward(ctx.saved, g_D.float().contiguous(), need_g_emb=ctx.needs_input_grad[2],
                                 defer=eng.defer_grads)
        ctx.saved = None
        if sg is None:           # gradient accumulation: this micro-batch was only summed into g_hat (engine.backward)
            return (None, None, g_emb, None) + (None,) * len(ctx.params)
        by_param = eng.grads_by_param(sg)
        grads = tuple(by_param[id(p)] if p.requires_grad else None for p in ctx.params)
        return (None, None, g_emb, None) + grads


class Denoiser(nn.Module):
    """EDM2 magnitude-preserving U-Net with the EDM preconditioning folded in (networks.py:490-605).

    `forward(noisy_image, sigma, embedding)`: noisy_image (B,C,H,W) fp32, sigma (B,) or 0-d, embedding (B|1, E)
    fp32 -> denoised image (B,C,H,W) fp32.
    """

    def __init__(self, in_channels: int = 3, out_channels: int = 3,
                 encoder_block_types: tuple[str, ...] = get_encoder_blocks_types(),
                 decoder_block_types: tuple[str, ...] = get_decoder_blocks_types(),
                 encoder_out_channels: tuple[int, ...] = get_encoder_out_channels(),
                 decoder_out_channels: tuple[int, ...] = get_decoder_out_channels(),
                 skip_connections: tuple[bool, ...] = get_skip_connections(),
                 dropout_rate: float = 0.0, sigma_data: float = 0.5, encoder_add_factor: float = 0.3,
                 decoder_add_factor: float = 0.3, embedding_dim: int = 768, num_heads: int = 4):
        super().__init__()
        lists = dict(encoder_block_types=encoder_block_types, decoder_block_types=decoder_block_types,
                     encoder_out_channels=encoder_out_channels, decoder_out_channels=decoder_out_channels,
                     skip_connections=skip_connections)
        lists = {k: tuple(v) for k, v in lists.items()}
        for a, b in (("encoder_block_types", "encoder_out_channels"), ("decoder_block_types", "decoder_out_channels"),
                     ("skip_connections", "decoder_out_channels")):
            assert len(lists[a]) == len(lists[b]), f"{a} and {b} must have the same length, got {len(lists[a])} and {len(lists[b])}"

        self.conv_in = Conv2d(in_channels + 1, lists["encoder_out_channels"][0], 3)
        self.conv_out = Conv2d(lists["decoder_out_channels"][-1], out_channels, 1)
        self.gain_out = nn.Parameter(torch.zeros(()))
        shared = dict(embedding_dim=embedding_dim, dropout_rate=dropout_rate, num_heads=num_heads)
        self.encoder_blocks = build_encoder_blocks(lists["encoder_block_types"], lists["encoder_out_channels"],
                                                   add_factor=encoder_add_factor, **shared)
        skip_channels = get_skip_channels(lists["encoder_out_channels"], lists["decoder_out_channels"], lists["skip_connections"])
        self.decoder_blocks = build_decoder_blocks(lists["decoder_block_types"], lists["decoder_out_channels"], skip_channels,
                                                   add_factor=decoder_add_factor, **shared)
        # every constructor argument mirrored as an attribute (utils.deinstantiate relies on it, utils.py:15-25)
        self.in_channels, self.out_channels = in_channels, out_channels
        for k, v in lists.items():
            setattr(self, k, v)
        self.dropout_rate, self.sigma_data = dropout_rate, sigma_data
        self.encoder_add_factor, self.decoder_add_factor = encoder_add_factor, decoder_add_factor
        self.embedding_dim, self.num_heads = embedding_dim, num_heads
        self._engine: DenoiserEngine | None = None

    @property
    def engine(self) -> DenoiserEngine:
        if self._engine is None:
            self._engine = DenoiserEngine(self)
        return self._engine

    def forward(self, noisy_image: Tensor, sigma: Tensor, embedding: Tensor):
        if not noisy_image.is_cuda:
            raise RuntimeError("tinyedm_b200.Denoiser runs on CUDA (sm_100a) only; there is no CPU fallback")
        noisy = noisy_image.float().contiguous()
        sig = sigma.to(device=noisy.device, dtype=F32).reshape(-1).contiguous()
        emb = embedding.float().contiguous()
        params = tuple(self.parameters())
        if torch.is_grad_enabled() and (emb.requires_grad or any(p.requires_grad for p in params)):
            return _DenoiserFn.apply(noisy, sig, emb, self, *params)
        D, _ = self.engine.forward(noisy, sig, emb, training=self.training, save=False)
        return D

    def prepare_weights(self, device) -> None:
        """(Re)normalises the weights if a parameter changed since the last call (eval-mode cache refresh)."""
        eng = self.engine
        eng._ensure_device(torch.device(device))
        eng.bank.prepare(self.training)

    @contextlib.contextmanager
    def accumulate_grads(self, enabled: bool = True):
        """Backwards inside this context only add to the flat dL/dw_hat buffer; the first backward after it hands out
        the gradients of the whole sum (Lightning's `accumulate_grad_batches`, imagenet.yaml:7). Parameter `.grad`s of
        the denoiser stay None inside the context."""
        eng = self.engine
        old = eng.defer_grads
        eng.defer_grads = bool(enabled) or old
        try:
            yield
        finally:
            eng.defer_grads = old

    def __getstate__(self):  # the engine holds device buffers and ctypes tables: rebuild lazily after copy/unpickle
        state = self.__dict__.copy()
        state["_engine"] = None
        return state


class DenoiserWrapper(nn.Module):
    """Generic EDM preconditioning around an arbitrary network (networks.py:608-646). Unused by every shipped config;
    kept importable. The four coefficient maps are (B,1,1,1) scalars, the network is the caller's: nothing here is a
    hot path, so this stays plain tensor arithmetic."""

    def __init__(self, net: nn.Module, sigma_data: float):
        super().__init__()
        self.net = net
        self._sigma_data = sigma_data

    @property
    def sigma_data(self) -> float:
        return self._sigma_data

    def forward(self, noisy_image: Tensor, sigma: Tensor, embedding: Tensor | None = None) -> Tensor:
        s = sigma.view(-1, 1, 1, 1)
        sd = self.sigma_data
        tot = s ** 2 + sd ** 2
        F_x = self.net(noisy_image / tot.sqrt(), (s.log() / 4).flatten(), embedding)
        return noisy_image * (sd ** 2 / tot) + F_x * (s * sd / tot.sqrt())
