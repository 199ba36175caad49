"""`FusedAdamEMA` — the optimiser step of the training hot path in ONE kernel launch.

Replaces `optim.Adam(fused=True)` (src/tinyedm/edm.py:250-253) plus the reference's EMA wrapper
(`EMAOptimizer.update`, src/tinyedm/ema.py:137-140 with decay `(1 - 1/(t+1))^(gamma+1)` :273 and
`gamma = sigma_rel_to_gamma(ema_length)` :29-32). It is a regular `torch.optim.Optimizer` (so `LambdaLR` and
Lightning drive it unchanged): `param_groups[0]["lr"]` is read every step. The forced weight re-normalisation
that EDM2 applies after each step happens at the next forward's weight-prep launch (networks.py:32-34).
"""
from __future__ import annotations

import contextlib
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
    """`every_n_steps`: the EMA is updated on steps whose 0-based index is a multiple of it (ema.py:262-269)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, ema_length: float | None = None,
                 every_n_steps: int = 1):
        if ema_length is not None and not (0 <= ema_length <= 0.2886):
            raise ValueError("EMA length value must be between 0 and 0.2886")
        if every_n_steps < 1:
            raise ValueError("every_n_steps must be >= 1")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self.gamma = sigma_rel_to_gamma(ema_length) if ema_length is not None else -1.0
        self.every_n_steps = int(every_n_steps)
        self.current_step = 0
        self._flat = None   # (m, v, ema) flat state buffers: the ONLY copy the kernel reads and writes
        self._table = None
        self._table_key = None

    # ---- state ----
    def _params(self):
        return [p for g in self.param_groups for p in g["params"] if p.requires_grad]

    def _ensure_state(self, params) -> None:
        """Allocates the flat m / v / ema buffers and binds `self.state[p]` to views of them. Tensors already present in
        `self.state` (put there by `load_state_dict`) are copied in, so a resumed run continues from the loaded moments."""
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
        for p, o in zip(params, offs):
            st = self.state[p]
            st.pop("step", None)
            for key, buf in (("exp_avg", m), ("exp_avg_sq", v), ("ema", ema)):
                if buf is None:
                    st.pop(key, None)
                    continue
                view = buf[o:o + p.numel()].view_as(p)
                old = st.get(key)
                if old is not None:
                    view.copy_(old.to(device=dev, dtype=F32))
                elif key == "ema" and self.current_step > 0:
                    view.copy_(p.detach())      # resumed from a state without EMA copies: start the average at the weights
                st[key] = view
        self._flat, self._offs = (m, v, ema), offs
        self._table_key = None

    @property
    def ema_params(self) -> tuple[Tensor, ...]:
        """EMA copies in `parameters()` order (the reference checkpoints them as `optimizer_states[0]["ema"]`)."""
        if self.gamma < 0:
            return ()
        params = self._params()
        if self._flat is None and params and params[0].is_cuda:
            self._ensure_state(params)
        return tuple(self.state[p]["ema"] if "ema" in self.state[p] else p.detach().clone() for p in params)

    # ---- checkpointing: the reference's layouts, both ways ----
    def _adam_state_dict(self) -> dict:
        """torch.optim.Adam's own layout ({"state": {i: {step, exp_avg, exp_avg_sq}}, "param_groups"}), loadable by the
        reference's `optim.Adam(fused=True)` (edm.py:250-253)."""
        sd = super().state_dict()
        step = torch.tensor(float(self.current_step))
        # (the packed per-parameter dicts ARE the live `self.state` entries: build new ones instead of editing them)
        sd["state"] = {i: {"step": step.clone(), **{k: v for k, v in st.items() if k in ("exp_avg", "exp_avg_sq")}}
                       for i, st in sd["state"].items()}
        sd["param_groups"] = [dict(g) for g in sd["param_groups"]]
        for g in sd["param_groups"]:
            for k, dflt in (("weight_decay", 0), ("amsgrad", False), ("maximize", False), ("foreach", None),
                            ("capturable", False), ("differentiable", False), ("fused", True)):
                g.setdefault(k, dflt)
        return sd

    def state_dict(self) -> dict:
        """With EMA: the reference's `EMAOptimizer.state_dict()` layout (ema.py:326-336) — what Lightning stores as
        `optimizer_states[0]` and what `EDM.load_from_checkpoint(load_ema=True)` reads back. Without EMA: Adam's layout."""
        opt = self._adam_state_dict()
        if self.gamma < 0:
            return opt
        return {"opt": opt, "ema": tuple(t.detach() for t in self.ema_params), "current_step": self.current_step,
                "gamma": self.gamma, "every_n_steps": self.every_n_steps}

    def ema_state_dict(self) -> dict:
        """Always the EMAOptimizer layout, with cloned EMA tensors."""
        return {"opt": self._adam_state_dict(), "ema": tuple(t.detach().clone() for t in self.ema_params),
                "current_step": self.current_step, "gamma": self.gamma, "every_n_steps": self.every_n_steps}

    def load_state_dict(self, state_dict: dict) -> None:
        """Accepts the EMAOptimizer layout (ema.py:338-348) or a plain Adam state dict. The Adam moments, the EMA copies
        and the step count all continue from the loaded values (the flat kernel buffers are rebuilt from them)."""
        ema = None
        if "opt" in state_dict:
            opt_sd, ema = state_dict["opt"], state_dict.get("ema")
            self.current_step = int(state_dict["current_step"])
            if self.gamma >= 0 and state_dict.get("gamma") is not None:
                self.gamma = float(state_dict["gamma"])
            self.every_n_steps = int(state_dict.get("every_n_steps", self.every_n_steps))
        else:
            opt_sd = state_dict
            steps = [float(st["step"]) for st in opt_sd["state"].values() if "step" in st]
            self.current_step = int(max(steps)) if steps else 0
        opt_sd = {"state": {k: dict(v) for k, v in opt_sd["state"].items()},
                  "param_groups": [{k: v for k, v in g.items() if k in ("params", "lr", "betas", "eps")} for g in opt_sd["param_groups"]]}
        super().load_state_dict(opt_sd)
        params = self._params()
        if ema is not None and self.gamma >= 0:
            if len(ema) != len(params):
                raise RuntimeError(f"tinyedm_b200.FusedAdamEMA: state holds {len(ema)} EMA tensors for {len(params)} parameters")
            for p, e in zip(params, ema):
                self.state[p]["ema"] = e.detach().clone()
        for g in self.param_groups:
            g["betas"] = tuple(g["betas"])
        self._flat = None          # rebuilt from self.state at the next step (or now, if the parameters are on the GPU)
        self._table_key = None
        if params and params[0].is_cuda:
            self._ensure_state(params)

    # ---- EMA weight swap (ema.py:293-317) ----
    def switch_main_parameter_weights(self) -> None:
        """Swaps every parameter with its EMA copy in place and invalidates the cached normalised weights."""
        from .engine import bump_weights_epoch
        from .utils import swap_tensors
        with torch.no_grad():
            for p, e in zip(self._params(), self.ema_params):
                swap_tensors(p.data, e)
        bump_weights_epoch()

    @contextlib.contextmanager
    def swap_ema_weights(self, enabled: bool = True):
        if enabled:
            self.switch_main_parameter_weights()
        try:
            yield
        finally:
            if enabled:
                self.switch_main_parameter_weights()

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
        update_ema = self.gamma >= 0 and self.current_step % self.every_n_steps == 0      # ema.py:262-269
        self.current_step += 1
        b1, b2 = group["betas"]
        _lib.call("tedm_adam_ema_step", self._table.data_ptr(), self._chunks.data_ptr(), self._n_chunks, float(group["lr"]),
                  float(self.current_step), None, float(b1), float(b2), float(group["eps"]),
                  float(self.gamma) if update_ema else -1.0, torch.cuda.current_stream().cuda_stream)
        torch.autograd.graph.increment_version(params)
        return loss
