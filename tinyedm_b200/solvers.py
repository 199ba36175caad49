"""`DeterministicSolver` — EDM Algorithm 1 (Heun, sigma(t)=t, s(t)=1), drop-in for src/tinyedm/solvers.py:4-59.

Same constructor / `solve(model, x0, class_labels)` surface and the same rho-schedule with a trailing zero
(solvers.py:34-41). The per-step arithmetic is one fused kernel per stage (tedm_heun_step): the Euler predictor
(derivative + step, solvers.py:49-50) and the trapezoidal corrector (:56-57) each read their operands once and
write the next state once; the schedule lives on the device, so the loop never synchronises with the host.
State and arithmetic are fp32 (the reference's only working dtype: any string dtype reaches `.to("float32")`
there, which torch parses as a device and raises — solvers.py:41).
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import ops
from .graphs import graphs_enabled
from .ops import F32


class DeterministicSolver:
    def __init__(self, num_steps: int = 18, sigma_min: float = 0.002, sigma_max: float = 80.0, rho: float = 7.0,
                 dtype: str | None = None):
        if dtype not in (None, "float32"):
            raise ValueError("tinyedm_b200.DeterministicSolver keeps the sampler state in float32 (dtype=None)")
        self.num_steps = num_steps
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.rho = rho
        self.dtype = torch.float32
        i = torch.arange(num_steps, dtype=self.dtype)
        t = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
        self.t_steps = torch.cat([t, torch.zeros(1)])
        self._dev_ts: dict = {}
        self._graphs: dict = {}

    def _schedule_on(self, device) -> Tensor:
        ts = self._dev_ts.get(device)
        if ts is None:
            ts = self._dev_ts[device] = self.t_steps.to(device)
        return ts

    @torch.no_grad()
    def solve(self, model, x0: Tensor, class_labels=None) -> Tensor:
        if not x0.is_cuda:
            raise RuntimeError("tinyedm_b200.DeterministicSolver runs on CUDA (sm_100a) only; there is no CPU fallback")
        ops.ensure_device(x0.device)
        x_in = ops.check(x0.to(F32).contiguous(), F32, "x0")
        # Whole-trajectory CUDA graph: the 2N-1 network evaluations and 2N stage kernels (~9 000 launches for N = 32)
        # replay as ONE graph launch. Needs a model whose prepared weights can be refreshed outside the graph
        # (`prepare_weights`, i.e. tinyedm_b200.EDM) in eval mode; anything else runs the eager loop.
        if graphs_enabled() and hasattr(model, "prepare_weights") and not getattr(model, "training", False):
            out = self._solve_graphed(model, x_in, class_labels)
            if out is not None:
                return out.to(x0.dtype)
        return self._solve_eager(model, x_in, class_labels).to(x0.dtype)

    def _solve_eager(self, model, x_in: Tensor, class_labels) -> Tensor:
        ts = self._schedule_on(x_in.device)
        x = torch.empty_like(x_in)
        ops.heun_step(x_in, None, None, None, x, None, ts, 0, ops.HEUN_SCALE)          # x = x0 * t_0      (:45)
        d = torch.empty_like(x)
        x1 = torch.empty_like(x)
        for i in range(self.num_steps):
            D = ops.check(model(x, ts[i], class_labels).to(F32).contiguous(), F32, "denoised")
            ops.heun_step(x, None, D, None, x1, d, ts, i, ops.HEUN_EULER)                # (:49-50)
            if i < self.num_steps - 1:
                D2 = ops.check(model(x1, ts[i + 1], class_labels).to(F32).contiguous(), F32, "denoised")
                ops.heun_step(x, x1, D2, d, x1, None, ts, i, ops.HEUN_CORRECT)           # (:56-57)
            x, x1 = x1, x
        return x

    def _solve_graphed(self, model, x_in: Tensor, class_labels) -> Tensor | None:
        lab_key = None if class_labels is None else (tuple(class_labels.shape), class_labels.dtype)
        key = (id(model), tuple(x_in.shape), lab_key, x_in.device.index)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 8:
                self._graphs.clear()
            ent = self._graphs[key] = self._capture(model, x_in, class_labels)
        if ent.get("graph") is None:
            return None
        model.prepare_weights(x_in.device)          # re-normalises the weights only if a parameter changed
        ent["x0"].copy_(x_in, non_blocking=True)
        if class_labels is not None:
            ent["labels"].copy_(class_labels, non_blocking=True)
        ent["graph"].replay()
        return ent["out"].clone()

    def _capture(self, model, x_in: Tensor, class_labels) -> dict:
        ent: dict = {"graph": None}
        try:
            ent["x0"] = x_in.clone()
            ent["labels"] = None if class_labels is None else class_labels.to(x_in.device).clone()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):             # eager warm-up: builds weight banks / descriptor tables
                model.prepare_weights(x_in.device)
                self._solve_eager(model, ent["x0"], ent["labels"])
            cur.wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._solve_eager(model, ent["x0"], ent["labels"])
            ent["graph"], ent["out"] = g, out
        except Exception as e:  # noqa: BLE001 - capture failure degrades to the eager loop
            ent = {"graph": None, "error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
        return ent
