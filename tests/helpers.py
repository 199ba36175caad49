"""Shared by the CPU and GPU tests: seeded parameters/inputs (same generator as oracle/make_golden.py) and model builders."""
from __future__ import annotations

import numpy as np
import torch

from oracle import edm2_oracle as O
from oracle.make_golden import SMALL, checksum, small_inputs, small_params  # noqa: F401


def rel(a, b) -> float:
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def spec_kwargs(s: O.DenoiserSpec) -> dict:
    return dict(in_channels=s.in_channels, out_channels=s.out_channels, encoder_block_types=s.encoder_block_types,
                decoder_block_types=s.decoder_block_types, encoder_out_channels=s.encoder_out_channels,
                decoder_out_channels=s.decoder_out_channels, skip_connections=s.skip_connections,
                dropout_rate=s.dropout_rate, sigma_data=s.sigma_data, encoder_add_factor=s.encoder_add_factor,
                decoder_add_factor=s.decoder_add_factor, embedding_dim=s.embedding_dim, num_heads=s.num_heads)


def build_modules(cfg, dp, ep, up=None, device="cuda"):
    """tinyedm_b200 modules loaded with the oracle's parameter dicts (state_dict names are the reference's)."""
    import tinyedm_b200 as T
    den = T.Denoiser(**spec_kwargs(cfg["denoiser"]))
    e = cfg["embedding"]
    emb = T.Embedding(e.fourier_dim, e.embedding_dim, e.num_classes, e.add_factor)
    den.load_state_dict({k: v.clone() for k, v in dp.items()}, strict=True)
    emb.load_state_dict({k: v.clone() for k, v in ep.items()}, strict=True)
    unc = None
    if up is not None:
        unc = T.UncertaintyNet(e.fourier_dim, e.fourier_dim)
        unc.load_state_dict({k: v.clone() for k, v in up.items()}, strict=True)
        unc = unc.to(device)
    return den.to(device), emb.to(device), unc


def nhwc_to_nchw(t: torch.Tensor) -> torch.Tensor:
    return t.permute(0, 3, 1, 2).float()


def cifar_cfg(num_classes=None, dropout=0.0):
    cfg = dict(O.CIFAR10)
    import dataclasses
    cfg["denoiser"] = dataclasses.replace(cfg["denoiser"], dropout_rate=dropout)
    cfg["embedding"] = dataclasses.replace(cfg["embedding"], num_classes=num_classes)
    return cfg


def seeded_params(cfg, seed=0, gain_out=1.0, u_gain=1.0):
    gen = torch.Generator().manual_seed(seed)
    dp = O.init_denoiser_params(cfg["denoiser"], gen, gain_out=gain_out)
    ep = O.init_embedding_params(cfg["embedding"], gen)
    up = O.init_uncertainty_params(cfg["embedding"].fourier_dim, gen, gain=u_gain)
    for k in dp:
        if k.endswith(".gain"):
            dp[k] = torch.tensor(0.5 + torch.rand((), generator=gen).item())
    for p in (dp, ep, up):
        O.project_weights_(p)
    return dp, ep, up
