"""`WeightedMeanSquaredError` — drop-in for the reference's torchmetrics metric (src/tinyedm/metric.py:21-49).

Same call surface: `metric(weight, preds, target)` returns the differentiable batch value
`sum_i mean_chw(w_i (p_i - t_i)^2) / N` with shape (1,) (torchmetrics `Metric.forward` with
`full_state_update=False`) while accumulating the running states `weighted_sum_squared_error` / `total`
(`dist_reduce_fx="sum"`, metric.py:33-36); `compute()` = their ratio; `reset()` clears them.
The reduction, the running-state update and the backward seed are one fused kernel each
(tedm_wmse_forward / tedm_wmse_backward); torchmetrics itself is not required.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .ops import F32


class _WmseFn(torch.autograd.Function):
    """loss = sum_b w_b mean((D-y)^2)/B [+ mean(u)] with w_b given explicitly or derived from (sigma, u) in-kernel."""

    @staticmethod
    def forward(ctx, preds, target, weight, sigma, u, sigma_data, wsum):
        ops.ensure_device(preds.device)
        D = ops.check(preds.float().contiguous(), F32, "preds")
        y = ops.check(target.float().contiguous(), F32, "target")
        B = y.shape[0]
        w = ops.check(weight.reshape(-1).float().contiguous(), F32, "weight") if weight is not None else None
        s = ops.check(sigma.reshape(-1).float().contiguous(), F32, "sigma") if sigma is not None else None
        uu = ops.check(u.reshape(-1).float().contiguous(), F32, "uncertainty") if u is not None else None
        for name, t in (("weight", w), ("sigma", s), ("uncertainty", uu)):
            if t is not None and t.numel() != B:
                raise RuntimeError(f"tinyedm_b200: `{name}` must have one entry per sample ({B}), got {t.numel()}")
        if D.shape != y.shape:
            raise RuntimeError(f"tinyedm_b200: preds {tuple(D.shape)} and target {tuple(y.shape)} differ")
        loss, mse = ops.wmse_forward(D, y, s, uu, w, float(sigma_data), wsum)
        ctx.saved = (D, y, s, uu, w, mse, float(sigma_data))
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        D, y, s, uu, w, mse, sd = ctx.saved
        need_w = w is not None and ctx.needs_input_grad[2]
        need_u = uu is not None and ctx.needs_input_grad[4]
        g_D, g_u, g_w = ops.wmse_backward(D, y, s, uu, w, mse, g_loss.reshape(1).float().contiguous(), sd, need_u, need_w)
        g_t = -g_D if ctx.needs_input_grad[1] else None
        return (g_D if ctx.needs_input_grad[0] else None), g_t, g_w, None, g_u, None, None


def fused_edm_loss(denoised: Tensor, clean: Tensor, sigma: Tensor, sigma_data: float, uncertainty: Tensor | None = None,
                   wsum: Tensor | None = None) -> Tensor:
    """edm.py:212-219 / :228 in one reduction: lambda(sigma) [* exp(-u)] weighted MSE [+ mean(u)], shape (1,)."""
    return _WmseFn.apply(denoised, clean, None, sigma, uncertainty, sigma_data, wsum)


class WeightedMeanSquaredError(nn.Module):
    is_differentiable = True
    higher_is_better = False
    full_state_update = False

    def __init__(self, **kwargs) -> None:
        super().__init__()
        self.register_buffer("weighted_sum_squared_error", torch.zeros(1), persistent=False)
        self.register_buffer("total", torch.tensor(0), persistent=False)

    def update(self, weight: Tensor, preds: Tensor, target: Tensor) -> None:
        with torch.no_grad():
            self.forward(weight, preds, target)

    def forward(self, weight: Tensor, preds: Tensor, target: Tensor) -> Tensor:
        if self.weighted_sum_squared_error.device != preds.device:
            self.to(preds.device)
        out = _WmseFn.apply(preds, target, weight, None, None, 0.0, self.weighted_sum_squared_error)
        self.total += target.shape[0]
        return out

    def edm_loss(self, denoised: Tensor, clean: Tensor, sigma: Tensor, sigma_data: float, uncertainty: Tensor | None = None):
        """Fused training loss that also feeds this metric's running state (used by EDM.training_step)."""
        if self.weighted_sum_squared_error.device != denoised.device:
            self.to(denoised.device)
        out = fused_edm_loss(denoised, clean, sigma, sigma_data, uncertainty, self.weighted_sum_squared_error)
        self.total += clean.shape[0]
        return out

    def compute(self) -> Tensor:
        return self.weighted_sum_squared_error / self.total

    def reset(self) -> None:
        self.weighted_sum_squared_error.zero_()
        self.total.zero_()
