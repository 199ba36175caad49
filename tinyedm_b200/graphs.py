"""CUDA-graph replay of the hot path: one graph launch instead of ~480 kernel launches per training step.

The reference drives `EDM.training_step` -> `loss.backward()` -> `optimizer.step()` eagerly (Lightning automatic
optimisation around src/tinyedm/edm.py:205-236); each of its ~5 000 ops is a separate launch. Here the step is already
down to ~480 launches of the C ABI, but at 24 ms of device time per step the Python/ctypes issue rate (~45 us per call
incl. torch allocator work) still leaves the GPU idle for ~2.5 ms per step. `GraphedTrainStep` captures
zero_grad + training_step + backward (+ the data-parallel gradient exchange) once and replays it; the fused
Adam/EMA kernel stays outside the graph (ONE launch) so the learning-rate schedule and the Adam bias-correction step are
ordinary host scalars, exactly as in the eager path.

Capture is safe because the engine never synchronises with the host, takes all memory from torch's allocator (graph
private pool during capture), keeps dropout seeds in a device-resident step counter, and draws the diffusion noise from
torch's graph-aware Philox generator.
"""
from __future__ import annotations

import os

import torch
from torch import Tensor

from . import _lib
from .engine import bump_weights_epoch


def graphs_enabled() -> bool:
    return os.environ.get("TEDM_CUDA_GRAPHS", "1") != "0"


class GraphedTrainStep:
    """`step = GraphedTrainStep(model, optimizer, example_batch[, ddp][, accumulate=k]); loss = step(batch)`.

    `batch` = (clean_image (k*B,C,H,W) fp32, class_label (k*B,) int64), host (pinned) or device tensors of the example's
    shapes; they are copied into static device buffers. With `accumulate=k` (Lightning's `accumulate_grad_batches`,
    imagenet.yaml:7) the batch is cut into k micro-batches of B whose gradients are summed (loss/k each, like Lightning)
    with ONE gradient exchange and ONE weight-norm Jacobian, on the last micro-batch. Returns the static loss tensor of
    shape (1,) (mean over the micro-batches; overwritten by the next call). Falls back to the eager step when capture is
    disabled (TEDM_CUDA_GRAPHS=0) or fails.

    The warm-up that precedes the capture runs forward + backward only (no optimiser step): parameters move only
    through the training-mode forced weight normalisation (networks.py:32-34), exactly as in any training forward;
    Adam moments, EMA and `current_step` are untouched.
    """

    def __init__(self, model, optimizer, example_batch, ddp=None, warmup: int = 2, accumulate: int = 1):
        self.model, self.opt, self.ddp = model, optimizer, ddp
        dev = next(model.parameters()).device
        x, y = example_batch
        if accumulate < 1 or x.shape[0] % accumulate != 0:
            raise ValueError(f"batch of {x.shape[0]} cannot be cut into {accumulate} micro-batches")
        self.accumulate = accumulate
        self.x = torch.empty(x.shape, device=dev, dtype=torch.float32)
        self.y = torch.empty(y.shape, device=dev, dtype=y.dtype)
        self.x.copy_(x)
        self.y.copy_(y)
        self.graph: torch.cuda.CUDAGraph | None = None
        self.loss: Tensor | None = None
        self.launches_per_step = 0       # C-ABI calls captured in the graph (+ the optimiser launch)
        self.error: str | None = None
        self._warmup = warmup
        if graphs_enabled():
            try:
                self._capture()
            except Exception as e:  # noqa: BLE001 - any capture failure degrades to the eager step
                import traceback
                self.graph = None
                self.error = f"{type(e).__name__}: {e}"
                self.error_traceback = traceback.format_exc()
                torch.cuda.synchronize()

    def _fwd_bwd(self) -> Tensor:
        self.opt.zero_grad(set_to_none=True)
        # The weight banks decide on the HOST whether the normalised operands are current (parameter versions); a captured
        # graph freezes that decision. Every step must re-normalise (the optimiser moved the weights), so the first
        # forward of a step always finds its banks invalidated — the later micro-batches of the same step then skip.
        bump_weights_epoch()
        k = self.accumulate
        if k == 1:
            loss = self.model.training_step((self.x, self.y), 0)
            loss.backward()
            total = loss
        else:
            total = None
            mb = self.x.shape[0] // k
            for j in range(k):
                with self.model.denoiser.accumulate_grads(j < k - 1):
                    loss = self.model.training_step((self.x[j * mb:(j + 1) * mb], self.y[j * mb:(j + 1) * mb]), j) / k
                    loss.backward()
                total = loss.detach() if total is None else total + loss.detach()
        if self.ddp is not None:
            self.ddp.finish_backward()
        return total

    def _capture(self) -> None:
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self._warmup):     # builds descriptor tables / gradient buffers outside the capture
                self._fwd_bwd()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = _lib.n_calls()
        with torch.cuda.graph(g):
            loss = self._fwd_bwd()
        self.launches_per_step = _lib.n_calls() - n0 + 1
        self.graph, self.loss = g, loss.detach()   # (capture records, it does not execute: no optimiser step is owed)
        # the replayed kernels write the gradients into THESE tensors; an eager backward in between would re-bind `.grad`
        self._grads = [(p, p.grad) for p in self.model.parameters() if p.grad is not None]

    def __call__(self, batch) -> Tensor:
        x, y = batch
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.graph is None:
            loss = self._fwd_bwd().detach()
        else:
            self.graph.replay()
            loss = self.loss
            for p, g in self._grads:
                if p.grad is not g:
                    p.grad = g
        self.opt.step()
        return loss
