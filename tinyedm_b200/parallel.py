"""Data-parallel training and batch-sharded sampling across the GPUs of one box.

One process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch) for the plumbing. The reference gets its
data parallelism implicitly from Lightning's DDP (25 MiB buckets of fp32 parameter gradients, all-reduced from
autograd hooks). Here the engine owns the backward, so the exchange is planned instead of hooked:

* what is exchanged is dL/dw_hat (`g_hat`, fp32): the weight-norm Jacobian applied afterwards is linear in g_hat
  and identical on every rank, so averaging g_hat == averaging the parameter gradients, and the single
  multi-tensor Jacobian launch runs once on the reduced buffer;
* the engine completes blocks in reverse order and the bank lays g_hat out in forward order, so finished regions
  grow from the END of one flat buffer: buckets are contiguous slices, launched on a side stream as soon as their
  last block is done, overlapping the remaining dgrad/wgrad work;
* the gradients of the 0-d parameters (block gains, gain_out) sit at the front of the same flat buffer and ride in its
  last message; the last unit to complete (embed weights, conv_in, conv_out) travels alone so that the only message
  with nothing left to hide under is small; the few tensors produced outside the Denoiser adjoint (embedding /
  uncertainty weights) follow in one small message (`finish_backward`).

Sampling needs no collective: rank r takes a contiguous slice of the batch (`shard_slice`).
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass

import torch
import torch.distributed as dist


def shard_slice(n: int, rank: int, world: int) -> slice:
    """Contiguous, balanced partition of `n` items: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


@dataclass
class Bucket:
    start: int        # element offsets into the flat g_hat buffer
    end: int
    ready_after: int  # index (in BACKWARD completion order) of the unit whose completion makes the bucket ready


def plan_buckets(unit_ranges: list[tuple[int, int]], bucket_elems: int, lone_tail: bool = False) -> list[Bucket]:
    """unit_ranges[i] = (start, end) of the i-th unit to COMPLETE during backward; units must tile a contiguous region
    from its end towards its start. Greedy: close a bucket once it holds >= bucket_elems. `lone_tail`: the last unit
    travels alone (a bucket is closed right before it), so the only message with no compute left to hide under is as
    small as the plan allows."""
    buckets: list[Bucket] = []
    cur_end = None      # end of the bucket being filled
    prev_start = None   # start of the most recent unit: the next one must end exactly there
    cur_start = None
    for i, (s, e) in enumerate(unit_ranges):
        if e <= s:
            continue
        if prev_start is not None and e != prev_start:
            raise ValueError(f"unit {i} [{s},{e}) is not adjacent to the previous units (expected end {prev_start})")
        if cur_end is None:
            cur_end = e
        cur_start = prev_start = s
        if cur_end - cur_start >= bucket_elems or (lone_tail and i == len(unit_ranges) - 2):
            buckets.append(Bucket(cur_start, cur_end, i))
            cur_end = None
    if cur_end is not None:
        buckets.append(Bucket(cur_start, cur_end, len(unit_ranges) - 1))
    return buckets


class GradReducer:
    """Averages slices of a flat buffer over the process group, asynchronously on a side stream (CUDA) or inline (CPU)."""

    def __init__(self, group=None, inline: bool = False):
        self.group = group
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.inline = inline      # CUDA: the compute stream waits for every reduction right away (no concurrent kernels)
        self._stream = None
        self._pending = []

    def _avg(self, t: torch.Tensor):
        if self.backend == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        w.wait()
        t.mul_(1.0 / self.world)
        return None

    def launch(self, t: torch.Tensor) -> None:
        if t.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=t.device)
            cur = torch.cuda.current_stream(t.device)
            self._stream.wait_stream(cur)           # the slice was produced on the compute stream
            with torch.cuda.stream(self._stream):
                w = self._avg(t)
            self._pending.append(w)
            if self.inline:
                self.wait(t.device)
        else:
            self._avg(t)

    def wait(self, device=None) -> None:
        """Makes the current stream wait for every launched reduction."""
        for w in self._pending:
            if w is not None:
                w.wait()                            # stream-level wait for NCCL work objects
        self._pending.clear()
        if self._stream is not None:
            torch.cuda.current_stream(device).wait_stream(self._stream)


class DistributedEDM:
    """Data-parallel wrapper around `tinyedm_b200.EDM` (the DDP-equivalent; edm.py is wrapped by Lightning's DDP in the
    reference). Usage per step:  loss = model.training_step(batch, i); loss.backward(); ddp.finish_backward(); opt.step()
    """

    def __init__(self, model, group=None, bucket_mb: float | None = None, broadcast: bool = True, inline: bool | None = None):
        """`bucket_mb` (default 25, or TEDM_DDP_BUCKET_MB): size at which a run of completed blocks is sent. `inline`
        (default off, or TEDM_DDP_INLINE=1): the compute stream waits for each message instead of running under it —
        NCCL's CTAs and the persistent one-CTA-per-SM conv kernels cannot share an SM, so an overlapped message delays
        the kernels it runs beside (profiles/r2_ddp_timeline_n8.txt); inline trades that for the bare transfer time."""
        import os
        if bucket_mb is None:
            bucket_mb = float(os.environ.get("TEDM_DDP_BUCKET_MB", "25"))
        if inline is None:
            inline = os.environ.get("TEDM_DDP_INLINE", "0") == "1"
        self.model = model
        self.group = group
        self.reducer = GradReducer(group, inline=inline)
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        eng = model.denoiser.engine
        eng.grad_sync = self
        self._eng = eng
        self._plan = None
        self._done_units = 0
        if broadcast:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0, group=group)
            from .engine import bump_weights_epoch
            bump_weights_epoch()      # `.data` writes bump no version: every bank (denoiser, embedding, head) is stale

    # ---- engine callbacks (see DenoiserEngine.backward) ----
    def _ensure_plan(self):
        if self._plan is None:
            units = self._eng.ghat_units_backward_order()
            # the trailing unit (embed weights of all blocks, conv_in, conv_out and the 0-d gradients at the front of the
            # buffer) completes when nothing is left to overlap with: it travels alone, everything before it is in flight
            self._plan = plan_buckets([u for u in units], self.bucket_elems, lone_tail=True)
        return self._plan

    def backward_started(self) -> None:
        self._done_units = 0
        self._next_bucket = 0
        self._ensure_plan()

    def unit_done(self) -> None:
        plan = self._plan
        i = self._done_units
        self._done_units += 1
        flat = self._eng.bank._ghat_flat
        while self._next_bucket < len(plan) and plan[self._next_bucket].ready_after <= i:
            b = plan[self._next_bucket]
            self.reducer.launch(flat[b.start:b.end])
            self._next_bucket += 1

    def before_weight_jacobian(self) -> None:
        """Called by the engine right before the multi-tensor weight-norm backward: everything must be reduced."""
        assert self._next_bucket == len(self._plan), "a g_hat bucket was never launched"
        self.reducer.wait()

    # ---- gradient accumulation ----
    @contextlib.contextmanager
    def no_sync(self, enabled: bool = True):
        """DDP's `no_sync()` for this wrapper (Lightning enters it for all but the last micro-batch when
        `accumulate_grad_batches > 1`, imagenet.yaml:7): backwards inside the context only add to the local flat g_hat
        buffer; the first backward outside it exchanges the SUM once and hands out the gradients. Call
        `finish_backward()` only after that last backward."""
        with self.model.denoiser.accumulate_grads(enabled):
            yield

    # ---- after loss.backward() ----
    def finish_backward(self) -> None:
        """Averages the gradients produced outside the Denoiser adjoint (embedding / uncertainty head)."""
        rest = [p for n, p in self.model.named_parameters() if not n.startswith("denoiser.") and p.grad is not None]
        if not rest:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in rest])
        self.reducer.launch(flat)
        self.reducer.wait()
        o = 0
        for p in rest:
            p.grad.copy_(flat[o:o + p.numel()].view_as(p))
            o += p.numel()
