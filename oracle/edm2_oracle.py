"""CPU oracle for the tinyedm denoiser hot path — TEST INFRASTRUCTURE ONLY.

A functional (parameter-dict driven) fp32 restatement of the reference algorithm in plain torch ops.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product package `tinyedm_b200` never does.

Pinned against the real reference: `oracle/make_golden.py` imports /root/reference/src/tinyedm/
{networks,solvers}.py by file path, runs both on identical seeded weights/inputs, asserts agreement and
stores the vectors in tests/golden/*.npz (checked by tests/test_oracle_golden.py on every run).
Parts that need packages absent from this image (lightning, torchmetrics) are restated from
edm.py / metric.py and pinned by the reference's own test identity
(tests/test_weighted_mean_squared_error.py:18-21); Lightning's autocast/DDP wrapping: parity unpinned.

Every function cites the reference lines (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

SILU_DIV = 0.596
EPS = 1e-4


# ------------------------------------------------------------------------------------------------
# Magnitude-preserving primitives
# ------------------------------------------------------------------------------------------------
def pixel_norm(x: Tensor, dim=1, eps: float = EPS) -> Tensor:
    """src/tinyedm/networks.py:9-14 — x / (eps + ||x||_dim / sqrt(N_dim)); the norm is taken in fp32 and
    cast to x.dtype before the divide."""
    n = torch.linalg.vector_norm(x, dim=dim, keepdim=True, dtype=torch.float32)
    scale = np.sqrt(n.numel() / x.numel(), dtype=np.float32)
    n = eps + n * float(scale)
    return x / n.to(x.dtype)


def normalize_weight(w: Tensor, eps: float = EPS) -> Tensor:
    """src/tinyedm/networks.py:17-19 — per-output-row pixel norm over all remaining dims."""
    return pixel_norm(w, dim=list(range(1, w.ndim)), eps=eps)


def effective_weight(w: Tensor) -> Tensor:
    """src/tinyedm/networks.py:35-36 / :58-59 — normalize(w) / sqrt(fan_in)."""
    fan_in = w[0].numel()
    return normalize_weight(w) / float(np.sqrt(fan_in, dtype=np.float32))


def forced_weight_norm_(w: Tensor) -> None:
    """src/tinyedm/networks.py:32-34 / :55-57 — training-mode in-place w <- normalize(w)."""
    with torch.no_grad():
        w.copy_(normalize_weight(w))


def mp_conv2d(x: Tensor, w: Tensor) -> Tensor:
    """src/tinyedm/networks.py:31-38 (eval branch) — stride-1 'same' conv with the effective weight."""
    return F.conv2d(x, effective_weight(w).to(x.dtype), padding="same")


def mp_linear(x: Tensor, w: Tensor) -> Tensor:
    """src/tinyedm/networks.py:54-61."""
    return F.linear(x, effective_weight(w).to(x.dtype))


def mp_silu(x: Tensor) -> Tensor:
    """src/tinyedm/networks.py:83-84."""
    return F.silu(x) / SILU_DIV


def mp_add(a: Tensor, b: Tensor, t: float = 0.5) -> Tensor:
    """src/tinyedm/networks.py:87-88."""
    return torch.lerp(a, b, t) / math.sqrt((1 - t) ** 2 + t ** 2)


def upsample2(x: Tensor) -> Tensor:
    """src/tinyedm/networks.py:67-72 — nearest-exact x2."""
    return F.interpolate(x, scale_factor=2, mode="nearest-exact")


def downsample2(x: Tensor) -> Tensor:
    """src/tinyedm/networks.py:75-80 — 2x2 average pooling."""
    return F.avg_pool2d(x, kernel_size=2, stride=2)


# ------------------------------------------------------------------------------------------------
# Architecture description (mirrors the Denoiser constructor arguments, networks.py:490-575)
# ------------------------------------------------------------------------------------------------
@dataclass
class DenoiserSpec:
    in_channels: int = 3
    out_channels: int = 3
    encoder_block_types: tuple = ()
    decoder_block_types: tuple = ()
    encoder_out_channels: tuple = ()
    decoder_out_channels: tuple = ()
    skip_connections: tuple = ()
    dropout_rate: float = 0.0
    sigma_data: float = 0.5
    encoder_add_factor: float = 0.3
    decoder_add_factor: float = 0.3
    embedding_dim: int = 256
    num_heads: int = 4

    def skip_channels(self) -> tuple:
        """src/tinyedm/networks.py:435-444."""
        out = [0] * len(self.decoder_out_channels)
        src = list(self.encoder_out_channels[::-1]) + [self.encoder_out_channels[0]]
        it = iter(src)
        for i, has in enumerate(self.skip_connections):
            if has:
                out[i] = next(it)
        return tuple(out)


@dataclass
class EmbeddingSpec:
    fourier_dim: int = 64
    embedding_dim: int = 256
    num_classes: int | None = None
    add_factor: float = 0.5


CIFAR10 = dict(  # experiments/conf/cifar10.yaml:20-56
    denoiser=DenoiserSpec(
        in_channels=3, out_channels=3, sigma_data=0.5, embedding_dim=256,
        encoder_block_types=("Enc", "Enc", "EncD", "EncA", "EncA", "EncD", "EncA", "EncA"),
        decoder_block_types=("DecA", "Dec", "DecA", "DecA", "DecA", "DecU", "DecA", "DecA", "DecA", "DecU", "Dec", "Dec", "Dec"),
        encoder_out_channels=(256,) * 8, decoder_out_channels=(256,) * 13,
        skip_connections=(False, False, True, True, True, False, True, True, True, False, True, True, True),
        dropout_rate=0.13),
    embedding=EmbeddingSpec(fourier_dim=64, embedding_dim=256, num_classes=None),
    P_mean=-1.2, P_std=1.2, image=(3, 32, 32), batch=256)

MNIST = dict(  # experiments/conf/mnist.yaml:20-56
    denoiser=DenoiserSpec(
        in_channels=1, out_channels=1, sigma_data=0.5, embedding_dim=256,
        encoder_block_types=("Enc", "Enc", "Enc", "EncD", "EncA", "EncA", "EncA", "EncD", "EncA", "EncA", "EncA"),
        decoder_block_types=("DecA", "Dec", "DecA", "DecA", "DecA", "DecA", "DecU", "DecA", "DecA", "DecA", "DecA", "DecU", "Dec", "Dec", "Dec", "Dec"),
        encoder_out_channels=(128,) * 4 + (256,) * 4 + (512,) * 3,
        decoder_out_channels=(512,) * 7 + (256,) * 5 + (128,) * 4,
        skip_connections=(False, False, True, True, True, True, False, True, True, True, True, False, True, True, True, True),
        dropout_rate=0.1),
    embedding=EmbeddingSpec(fourier_dim=64, embedding_dim=256, num_classes=10),
    P_mean=-1.2, P_std=1.2, image=(1, 28, 28), batch=128)


IMAGENET = dict(  # experiments/conf/imagenet.yaml:20-51 + the Denoiser defaults (networks.py:332-432)
    denoiser=DenoiserSpec(
        in_channels=4, out_channels=4, sigma_data=0.5, embedding_dim=768,
        encoder_block_types=("Enc", "Enc", "Enc", "EncD", "Enc", "Enc", "Enc", "EncD", "EncA", "EncA", "EncA", "EncD", "EncA", "EncA", "EncA"),
        decoder_block_types=("DecA", "Dec", "DecA", "DecA", "DecA", "DecA", "DecU", "DecA", "DecA", "DecA", "DecA", "DecU",
                             "Dec", "Dec", "Dec", "Dec", "DecU", "Dec", "Dec", "Dec", "Dec"),
        encoder_out_channels=(192,) * 4 + (384,) * 4 + (576,) * 4 + (768,) * 3,
        decoder_out_channels=(768,) * 6 + (576,) * 5 + (384,) * 6 + (192,) * 4,
        skip_connections=(False, False, True, True, True, True, False, True, True, True, True, False, True, True, True, True,
                          False, True, True, True, True),
        dropout_rate=0.0),
    embedding=EmbeddingSpec(fourier_dim=192, embedding_dim=768, num_classes=1000),
    P_mean=-0.4, P_std=1.0, image=(4, 64, 64), batch=176)


def block_plan(spec: DenoiserSpec):
    """Static description of every block (in/out/skip channels, flags) — networks.py:447-487."""
    enc, dec = [], []
    cin = spec.encoder_out_channels[0]
    for typ, cout in zip(spec.encoder_block_types, spec.encoder_out_channels):
        enc.append(dict(cin=cin, cout=cout, down=typ.endswith("D"), attn=typ.endswith("A")))
        cin = cout
    cin = spec.decoder_out_channels[0]
    for typ, cout, cs in zip(spec.decoder_block_types, spec.decoder_out_channels, spec.skip_channels()):
        dec.append(dict(cin=cin, cout=cout, cskip=int(cs), up=typ.endswith("U"), attn=typ.endswith("A")))
        cin = cout
    return enc, dec


# ------------------------------------------------------------------------------------------------
# Parameter initialisation with the reference's state_dict names (networks.py:28, :49, :96, :135-136,
# :244, :538)
# ------------------------------------------------------------------------------------------------
def init_denoiser_params(spec: DenoiserSpec, gen: torch.Generator, gain_out: float = 0.0) -> dict:
    p: dict[str, Tensor] = {}
    rn = lambda *s: torch.randn(*s, generator=gen)
    enc, dec = block_plan(spec)
    E = spec.embedding_dim
    p["conv_in.weight"] = rn(spec.encoder_out_channels[0], spec.in_channels + 1, 3, 3)
    p["conv_out.weight"] = rn(spec.out_channels, spec.decoder_out_channels[-1], 1, 1)
    p["gain_out"] = torch.tensor(float(gain_out))

    def attn(prefix, c):
        p[prefix + "attention.qkv_conv.weight"] = rn(3 * c, c, 1, 1)
        p[prefix + "attention.out_conv.weight"] = rn(c, c, 1, 1)

    for i, b in enumerate(enc):
        pre = f"encoder_blocks.{i}."
        if b["cin"] != b["cout"]:
            p[pre + "conv_1x1.weight"] = rn(b["cout"], b["cin"], 1, 1)
        p[pre + "conv_3x3_1.weight"] = rn(b["cout"], b["cout"], 3, 3)
        p[pre + "conv_3x3_2.weight"] = rn(b["cout"], b["cout"], 3, 3)
        if b["attn"]:
            attn(pre, b["cout"])
        p[pre + "embed.weight"] = rn(b["cout"], E)
        p[pre + "gain"] = torch.ones(())
    for i, b in enumerate(dec):
        pre = f"decoder_blocks.{i}."
        tot = b["cin"] + b["cskip"]
        if b["cskip"] > 0:
            p[pre + "cat_factor.layer1.weight"] = rn(b["cskip"] // 16, b["cskip"] + 1, 1, 1)
            p[pre + "cat_factor.layer2.weight"] = rn(b["cskip"], b["cskip"] // 16, 1, 1)
        if tot != b["cout"]:
            p[pre + "conv_1x1.weight"] = rn(b["cout"], tot, 1, 1)
        p[pre + "conv_3x3_1.weight"] = rn(b["cout"], tot, 3, 3)
        p[pre + "conv_3x3_2.weight"] = rn(b["cout"], b["cout"], 3, 3)
        if b["attn"]:
            attn(pre, b["cout"])
        p[pre + "embed.weight"] = rn(b["cout"], E)
        p[pre + "gain"] = torch.ones(())
    return p


def init_embedding_params(spec: EmbeddingSpec, gen: torch.Generator) -> dict:
    p = {
        "fourier_embed.freqs": 2 * math.pi * torch.randn(spec.fourier_dim, generator=gen),
        "fourier_embed.phases": 2 * math.pi * torch.rand(spec.fourier_dim, generator=gen),
        "sigma_embed.weight": torch.randn(spec.embedding_dim, spec.fourier_dim, generator=gen),
    }
    if spec.num_classes is not None and spec.num_classes != -1:
        p["class_embed.linear.weight"] = torch.randn(spec.embedding_dim, spec.num_classes, generator=gen)
    return p


def init_uncertainty_params(fourier_dim: int, gen: torch.Generator, gain: float = 0.0) -> dict:
    """src/tinyedm/networks.py:91-96 with in=hidden=fourier_dim (edm.py:142-146)."""
    return {
        "linear1.weight": torch.randn(fourier_dim, fourier_dim + 1, generator=gen),
        "linear2.weight": torch.randn(1, fourier_dim, generator=gen),
        "gain": torch.tensor(float(gain)),
    }


def project_weights_(params: dict) -> None:
    """One training-mode forward leaves every weight on its norm sphere (networks.py:32-34)."""
    for k, v in params.items():
        if k.endswith(".weight"):
            forced_weight_norm_(v)


# ------------------------------------------------------------------------------------------------
# Embedding / heads
# ------------------------------------------------------------------------------------------------
def embedding_forward(p: dict, spec: EmbeddingSpec, sigma: Tensor, labels: Tensor | None = None):
    """src/tinyedm/networks.py:163-178 (+ :132-141 Fourier, :121-129 class embedding). fp32 throughout.
    Returns (fourier_embedding, embedding)."""
    c_noise = sigma.float().log() / 4
    four = torch.outer(c_noise.flatten(), p["fourier_embed.freqs"]) + p["fourier_embed.phases"]
    four = four.cos() * float(np.sqrt(2, dtype=np.float32))
    emb = mp_linear(four, p["sigma_embed.weight"])
    if labels is not None:
        if "class_embed.linear.weight" not in p:
            raise ValueError("class_labels is not None, but num_classes is None. ")
        n = p["class_embed.linear.weight"].shape[1]
        onehot = F.one_hot(labels.flatten(), n) * float(np.sqrt(n, dtype=np.float32))
        cls = mp_linear(onehot, p["class_embed.linear.weight"])
        emb = mp_add(emb, cls, spec.add_factor)
    return four, mp_silu(emb)


def uncertainty_forward(p: dict, fourier: Tensor) -> Tensor:
    """src/tinyedm/networks.py:98-103."""
    x = torch.cat((fourier, torch.ones_like(fourier[:, 0:1])), dim=1)
    x = mp_silu(mp_linear(x, p["linear1.weight"]))
    return p["gain"] * mp_linear(x, p["linear2.weight"])


def scale_long(p: dict, pre: str, skip: Tensor) -> Tensor:
    """src/tinyedm/networks.py:112-118."""
    x = torch.cat((skip, torch.ones_like(skip[:, 0:1])), dim=1).mean(dim=[2, 3], keepdim=True)
    x = mp_silu(mp_conv2d(x, p[pre + "layer1.weight"]))
    return torch.sigmoid(mp_conv2d(x, p[pre + "layer2.weight"]))


def cosine_attention(p: dict, pre: str, x: Tensor, heads: int) -> Tensor:
    """src/tinyedm/networks.py:191-207. Channel c of qkv = head*3*hd + d*3 + {q,k,v}."""
    b, c, h, w = x.shape
    qkv = mp_conv2d(x, p[pre + "qkv_conv.weight"]).view(b, heads, -1, 3, h * w)
    qkv = pixel_norm(qkv, dim=2)
    q, k, v = (t.transpose(2, 3) for t in qkv.unbind(3))
    y = F.scaled_dot_product_attention(q, k, v)
    y = y.transpose(2, 3).reshape(b, -1, h, w)
    y = mp_conv2d(y, p[pre + "out_conv.weight"])
    return mp_add(x, y, 0.5)


def _modulation(p: dict, pre: str, emb: Tensor) -> Tensor:
    """src/tinyedm/networks.py:255-258 / :319-322 — (embed(emb) * gain + 1), fp32."""
    return mp_linear(emb.float(), p[pre + "embed.weight"]) * p[pre + "gain"] + 1


def encoder_block(p, pre, blk, x, emb, spec: DenoiserSpec, taps=None, drop: Callable | None = None):
    """src/tinyedm/networks.py:246-265."""
    if blk["down"]:
        x = downsample2(x)
    if blk["cin"] != blk["cout"]:
        x = mp_conv2d(x, p[pre + "conv_1x1.weight"])
    x = pixel_norm(x)
    r = mp_conv2d(mp_silu(x), p[pre + "conv_3x3_1.weight"])
    r = mp_silu(r * _modulation(p, pre, emb)[:, :, None, None])
    if drop is not None:
        r = drop(r)
    r = mp_conv2d(r, p[pre + "conv_3x3_2.weight"])
    out = mp_add(x, r, spec.encoder_add_factor)
    if taps is not None:
        taps[pre + "resid"] = out
    if blk["attn"]:
        out = cosine_attention(p, pre + "attention.", out, spec.num_heads)
    return out


def decoder_block(p, pre, blk, x, emb, skip, spec: DenoiserSpec, taps=None, drop: Callable | None = None):
    """src/tinyedm/networks.py:306-329."""
    if skip is not None:
        x = torch.cat((x, skip * scale_long(p, pre + "cat_factor.", skip)), dim=1)
    if blk["up"]:
        x = upsample2(x)
    r = x
    if blk["cin"] + blk["cskip"] != blk["cout"]:
        x = mp_conv2d(x, p[pre + "conv_1x1.weight"])
    r = mp_conv2d(mp_silu(r), p[pre + "conv_3x3_1.weight"])
    r = mp_silu(r * _modulation(p, pre, emb)[:, :, None, None])
    if drop is not None:
        r = drop(r)
    r = mp_conv2d(r, p[pre + "conv_3x3_2.weight"])
    out = mp_add(x, r, spec.decoder_add_factor)
    if taps is not None:
        taps[pre + "resid"] = out
    if blk["attn"]:
        out = cosine_attention(p, pre + "attention.", out, spec.num_heads)
    return out


def precond_coeffs(sigma: Tensor, sigma_data: float):
    """src/tinyedm/networks.py:578-581."""
    s = sigma.view(-1, 1, 1, 1)
    c_skip = sigma_data ** 2 / (s ** 2 + sigma_data ** 2)
    c_out = s * sigma_data / (s ** 2 + sigma_data ** 2).sqrt()
    c_in = 1 / (sigma_data ** 2 + s ** 2).sqrt()
    return c_skip, c_out, c_in


def denoiser_forward(p: dict, spec: DenoiserSpec, noisy: Tensor, sigma: Tensor, emb: Tensor,
                     taps: dict | None = None, drop: Callable | None = None) -> Tensor:
    """src/tinyedm/networks.py:577-605. `taps` (optional) collects every block output (NCHW)."""
    c_skip, c_out, c_in = precond_coeffs(sigma, spec.sigma_data)
    enc, dec = block_plan(spec)
    x = c_in * noisy
    x = torch.cat((x, torch.ones_like(x[:, 0:1])), dim=1)
    x = mp_conv2d(x, p["conv_in.weight"])
    if taps is not None:
        taps["conv_in"] = x
    skips = [x]
    for i, blk in enumerate(enc):
        x = encoder_block(p, f"encoder_blocks.{i}.", blk, x, emb, spec, taps, drop)
        skips.append(x)
        if taps is not None:
            taps[f"encoder_blocks.{i}"] = x
    for i, blk in enumerate(dec):
        skip = skips.pop() if spec.skip_connections[i] else None
        x = decoder_block(p, f"decoder_blocks.{i}.", blk, x, emb, skip, spec, taps, drop)
        if taps is not None:
            taps[f"decoder_blocks.{i}"] = x
    f = mp_conv2d(x, p["conv_out.weight"]) * p["gain_out"]
    if taps is not None:
        taps["conv_out"] = f
    return f * c_out + noisy * c_skip


# ------------------------------------------------------------------------------------------------
# Training-step arithmetic
# ------------------------------------------------------------------------------------------------
def diffuse(clean: Tensor, eps: Tensor, noise: Tensor, P_mean: float, P_std: float):
    """src/tinyedm/edm.py:84-93 with the two random draws passed in (eps: (B,), noise: like clean)."""
    sigma = (P_mean + eps * P_std).exp()
    return clean + noise * sigma.view(-1, 1, 1, 1), sigma


def loss_weight(sigma: Tensor, sigma_data: float) -> Tensor:
    """src/tinyedm/edm.py:212."""
    return (sigma ** 2 + sigma_data ** 2) / (sigma * sigma_data) ** 2


def weighted_mse(weight: Tensor, preds: Tensor, target: Tensor) -> Tensor:
    """src/tinyedm/metric.py:8-18 + Metric.forward batch value (sum / N), shape (1,)."""
    n = target.shape[0]
    d = preds.reshape(n, -1) - target.reshape(n, -1)
    s = torch.mean(weight.view(n, 1) * d * d, dim=1).sum()
    return (s / n).reshape(1)


def training_loss(weight: Tensor, denoised: Tensor, clean: Tensor, uncertainty: Tensor | None = None) -> Tensor:
    """src/tinyedm/edm.py:213-219 / :228."""
    if uncertainty is None:
        return weighted_mse(weight, denoised, clean)
    u = uncertainty.flatten()
    return weighted_mse(weight / u.exp(), denoised, clean) + u.mean()


def edm_forward(dp, dspec, ep, espec, x, sigma, labels=None):
    """src/tinyedm/edm.py:280-286."""
    labels = labels if (espec.num_classes is not None) else None
    _, emb = embedding_forward(ep, espec, sigma, labels)
    return denoiser_forward(dp, dspec, x, sigma, emb)


# ------------------------------------------------------------------------------------------------
# Sampler
# ------------------------------------------------------------------------------------------------
def t_schedule(num_steps: int, sigma_min=0.002, sigma_max=80.0, rho=7.0) -> Tensor:
    """src/tinyedm/solvers.py:34-41 (fp32 arithmetic, trailing zero)."""
    i = torch.arange(num_steps, dtype=torch.float32)
    t = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros(1)])


def heun_solve(model: Callable, x0: Tensor, labels=None, num_steps=18, sigma_min=0.002, sigma_max=80.0, rho=7.0,
               trajectory: list | None = None) -> Tensor:
    """src/tinyedm/solvers.py:43-59 — EDM Algorithm 1 (deterministic Heun), fp32."""
    ts = t_schedule(num_steps, sigma_min, sigma_max, rho).to(x0.device)
    x = x0.float() * ts[0]
    for i in range(num_steps):
        t0, t1 = ts[i], ts[i + 1]
        cur = x
        d = (cur - model(cur, t0, labels).float()) / t0
        x = cur + (t1 - t0) * d
        if i < num_steps - 1:
            d2 = (x - model(x, t1, labels).float()) / t1
            x = cur + (t1 - t0) * (0.5 * d + 0.5 * d2)
        if trajectory is not None:
            trajectory.append(x.clone())
    return x


def to_uint8_images(x: Tensor, mean, std) -> Tensor:
    """src/tinyedm/callbacks.py:152-154 (PreditionWriter.write_on_batch_end): denormalise, clamp, NHWC, * 255, uint8."""
    mean = torch.as_tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
    std = torch.as_tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
    images = x * std * 2 + mean
    images = torch.clamp(images, 0, 1).permute(0, 2, 3, 1) * 255
    return images.to(torch.uint8)
