"""`FusedAdamEMA` — the optimiser step of the training hot path in ONE kernel launch.

Replaces `optim.Adam(fused=True)` (src/tinyedm/edm.py:250-253) plus the reference's EMA wrapper
(`EMAOptimizer.update`, src/tinyedm/ema.py:137-140 with decay `(1 - 1/(t+1))^(gamma+1)` :273 and
`gamma = sigma_rel_to_gamma(ema_length)` :29-32). It is a regular `torch.optim.Optimizer` (so `LambdaLR` and
Lightning drive it unchanged): `param_groups[0]["lr"]` is read every step. The forced weight re-normalisation
that EDM2 applies after each step happens at the next forward's weight-prep launch (networks.py:32-34).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops
from .ops import F32


class AdamDesc(ctypes.Structure):
    """Mirror of `tedm_adam_desc` (include/tinyedm_b200.h)."""

    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("ema", ctypes.c_void_p), ("n", ctypes.c_int64)]


def sigma_rel_to_gamma(sigma_rel: float) -> float:
    """EDM2 power-function EMA exponent for a relative std `sigma_rel` (ema.py:29-32)."""
    t = sigma_rel ** -2
    return float(np.roots([1, 7, 16 - t, 12 - t]).real.max())


class FusedAdamEMA(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, ema_length: float | None = None):
        if ema_length is not None and not (0 <= ema_length <= 0.2886):
            raise ValueError("EMA length value must be between 0 and 0.2886")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self.gamma = sigma_rel_to_gamma(ema_length) if ema_length is not None else -1.0
        self.current_step = 0
        self._flat = None   # (m, v, ema) flat state buffers
        self._table = None
        self._table_key = None

    # ---- state ----
    def _params(self):
        return [p for g in self.param_groups for p in g["params"] if p.requires_grad]

    def _ensure_state(self, params) -> None:
        if self._flat is not None:
            return
        dev = params[0].device
        ops.ensure_device(dev)
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 127) // 128 * 128
        m = torch.zeros(n, device=dev, dtype=F32)
        v = torch.zeros(n, device=dev, dtype=F32)
        ema = torch.zeros(n, device=dev, dtype=F32) if self.gamma >= 0 else None
        self._flat, self._offs = (m, v, ema), offs
        for p, o in zip(params, offs):
            st = self.state[p]
            st["exp_avg"] = m[o:o + p.numel()].view_as(p)
            st["exp_avg_sq"] = v[o:o + p.numel()].view_as(p)
            if ema is not None:
                st["ema"] = ema[o:o + p.numel()].view_as(p)

    @property
    def ema_params(self) -> tuple[Tensor, ...]:
        """EMA copies in `parameters()` order (the reference checkpoints them as `optimizer_states[0]["ema"]`)."""
        return tuple(self.state[p]["ema"] for p in self._params()) if self.gamma >= 0 else ()

    def ema_state_dict(self) -> dict:
        """The reference's `EMAOptimizer.state_dict()` layout (ema.py:326-336): what Lightning stores as
        `optimizer_states[0]` and what `EDM.load_from_checkpoint(load_ema=True)` reads back."""
        return {"opt": super().state_dict(), "ema": tuple(t.detach().clone() for t in self.ema_params),
                "current_step": self.current_step, "gamma": self.gamma, "every_n_steps": 1}

    def _ensure_table(self, params) -> None:
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        if key == self._table_key:
            return
        m, v, ema = self._flat
        chunk = _lib.call_int("tedm_adam_chunk_elems")
        arr = (AdamDesc * len(params))()
        chunks = []
        for i, (d, p, o) in enumerate(zip(arr, params, self._offs)):
            if not (p.is_cuda and p.dtype == F32 and p.is_contiguous() and p.grad.is_contiguous() and p.grad.dtype == F32):
                raise RuntimeError("tinyedm_b200.FusedAdamEMA: parameters and gradients must be contiguous fp32 CUDA tensors")
            d.p, d.g, d.n = p.data_ptr(), p.grad.data_ptr(), p.numel()
            d.m, d.v = m.data_ptr() + 4 * o, v.data_ptr() + 4 * o
            d.ema = ema.data_ptr() + 4 * o if ema is not None else None
            chunks += [(i, c) for c in range((p.numel() + chunk - 1) // chunk)]
        dev = params[0].device
        # pinned + non_blocking: a pageable upload would block the host until the stream drains
        self._host_table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory()
        self._host_chunks = torch.tensor(chunks, dtype=torch.int32).pin_memory()
        self._table = self._host_table.to(dev, non_blocking=True)
        self._chunks = self._host_chunks.to(dev, non_blocking=True)
        self._n_chunks = len(chunks)
        self._table_key = key

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        params = [p for p in self._params() if p.grad is not None]
        if not params:
            return loss
        if len(params) != len(self._params()):
            raise RuntimeError("tinyedm_b200.FusedAdamEMA: every parameter must have a gradient")
        self._ensure_state(params)
        self._ensure_table(params)
        group = self.param_groups[0]
        self.current_step += 1
        b1, b2 = group["betas"]
        _lib.call("tedm_adam_ema_step", self._table.data_ptr(), self._chunks.data_ptr(), self._n_chunks, float(group["lr"]),
                  float(self.current_step), None, float(b1), float(b2), float(group["eps"]), float(self.gamma),
                  torch.cuda.current_stream().cuda_stream)
        torch.autograd.graph.increment_version(params)
        return loss
