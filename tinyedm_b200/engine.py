"""Execution engine of the B200 denoiser hot path: which kernel runs when, on which buffers.

The reference executes `Denoiser.forward` (src/tinyedm/networks.py:577-605) as ~1 400 eager torch ops per
call and leaves the backward to autograd (~3 000 more). Here the whole network is ONE statically planned
sequence of C-ABI kernel launches (see include/tinyedm_b200.h):

    weight bank    one multi-tensor launch normalises every weight (networks.py:32-36) into the bf16
                   tensor-core operand layouts (forward and data-gradient) / fp32 for the tiny layers
    modulation     one fp32 GEMM for the `embed` Linear of ALL blocks + one finishing kernel (:255-258)
    conv_in        preconditioning c_in, ones channel and 3x3 patch gather fused, then a K=64 GEMM (:578-587)
    blocks         block_prep (resample + concat*ScaleLong + pixel_norm + mp_silu, one pass) ->
                   conv3x3 [epilogue: modulation * mp_silu * dropout] -> conv3x3 [epilogue: mp_add with x]
                   -> (qkv conv -> attention -> out conv [epilogue: mp_add])            (:246-265, :306-329)
    conv_out       1x1 conv fused with D = c_skip x + c_out gain_out F                  (:602-603)

The backward is the hand-derived adjoint of the same plan: data gradients are the
same implicit-GEMM kernel with the flipped/transposed weight layout, weight gradients a pixel-contraction
GEMM into fp32 `g_hat` buffers, and one multi-tensor launch turns every `g_hat` into the gradient of the
raw parameter through the weight-norm Jacobian.

Activations are NHWC bf16. All memory comes from torch's caching allocator; the engine is CUDA-graph
capturable (no host synchronisation, no data-dependent host branching).
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field

import torch
from torch import Tensor

from . import _lib, ops
from .ops import (BF16, F32, EPI_AXPBY, EPI_MODSILU, EPI_MODSILU_BWD, EPI_SILU_BWD, RESAMPLE_DOWN, RESAMPLE_NONE,
                  RESAMPLE_UP)

_ALIGN = 128  # elements; keeps every slice of the flat buffers 256-byte aligned for TMA


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


# Writes through `param.data` (the reference's EMA swap: ema.py:293-296 via utils.swap_tensors; dist.broadcast of
# `t.data`) do not bump `param._version`, so the eval-mode cache of prepared weights is also keyed on this process-wide
# epoch: `swap_tensors`, `FusedAdamEMA.switch_main_parameter_weights`, `DistributedEDM` and `EDM.invalidate_weights` bump it.
_WEIGHTS_EPOCH = [0]


def bump_weights_epoch() -> None:
    """Invalidates every cached set of normalised weights in this process (they are rebuilt at the next forward)."""
    _WEIGHTS_EPOCH[0] += 1


# =========================================================================================================
# Weight bank
# =========================================================================================================
@dataclass
class WeightSlot:
    name: str
    param: torch.nn.Parameter
    rows: int
    cin: int
    taps: int
    kpad: int
    want_fwd: bool = False      # bf16 [rows][kpad], k = tap*cin + ci
    want_dgrad: bool = False    # bf16 [cin][taps flipped][rows]
    want_f32: bool = False      # fp32 [rows][cin*taps]
    # filled by WeightBank.materialise
    fwd: Tensor | None = None
    dgrad: Tensor | None = None
    f32: Tensor | None = None
    ghat: Tensor | None = None  # fp32 [rows][kpad]  dL/dw_hat
    grad: Tensor | None = None  # fp32, shaped like the parameter
    row_start: int = 0
    qkv_head_dim: int = 0       # != 0: rows permuted to the q | k | v plane layout (see tedm_weight_desc)


class WeightBank:
    """All force-normalised weights of one module group, prepared by ONE kernel launch.

    Reference behaviour reproduced (networks.py:32-36, :55-59): in training mode every forward first rewrites
    the fp32 parameter in place with `normalize(weight)`; the effective weight is `normalize(w)/sqrt(fan_in)`.
    In eval mode the prepared operands are cached until a parameter changes (`_version` / `data_ptr`), which is
    what makes the 63 network evaluations of a sampling run pay for weight normalisation once instead of 63x.
    """

    def __init__(self, slots: list[WeightSlot]):
        self.slots = slots
        self.device: torch.device | None = None
        self._key = None
        self._table: Tensor | None = None
        self._table_ptrs = None
        self._have_grad_buffers = False
        self.aux_elems = 0      # fp32 scratch at the FRONT of the flat g_hat buffer (the engine keeps its 0-d gradients there)
        self.total_rows = sum(s.rows for s in slots)
        self.total_groups = sum((s.rows + 15) // 16 for s in slots)     # 16 prepared rows per CTA of weight_prep_fwd
        self.max_row_floats = max(s.taps * (s.cin + 4) for s in slots)  # staging of one dL/dw_hat row in weight_prep_bwd

    # ---- buffers ----
    def materialise(self, device: torch.device) -> None:
        self.device = device
        n_fwd = n_dg = n_f32 = 0
        row = 0
        offs = []
        for s in self.slots:
            s.row_start = row
            row += s.rows
            o = (n_fwd, n_dg, n_f32)
            if s.want_fwd:
                n_fwd += _round_up(s.rows * s.kpad, _ALIGN)
            if s.want_dgrad:
                n_dg += _round_up(s.rows * s.cin * s.taps, _ALIGN)
            if s.want_f32:
                n_f32 += _round_up(s.rows * s.cin * s.taps, _ALIGN)
            offs.append(o)
        self._fwd_flat = torch.zeros(max(n_fwd, 1), device=device, dtype=BF16)
        self._dg_flat = torch.zeros(max(n_dg, 1), device=device, dtype=BF16)
        self._f32_flat = torch.zeros(max(n_f32, 1), device=device, dtype=F32)
        self.stats = torch.zeros(4 * self.total_rows, device=device, dtype=F32)
        for s, (of, od, o3) in zip(self.slots, offs):
            n = s.rows * s.cin * s.taps
            if s.want_fwd:
                s.fwd = self._fwd_flat[of:of + s.rows * s.kpad].view(s.rows, s.kpad)
            if s.want_dgrad:
                s.dgrad = self._dg_flat[od:od + n].view(s.cin, s.taps * s.rows)
            if s.want_f32:
                s.f32 = self._f32_flat[o3:o3 + n].view(s.rows, s.cin * s.taps)
        self._key = None
        self._table = None
        self._have_grad_buffers = False

    def ensure_grad_buffers(self) -> None:
        if self._have_grad_buffers:
            return
        aux = _round_up(self.aux_elems, _ALIGN)
        n_gh = aux + sum(_round_up(s.rows * s.kpad, _ALIGN) for s in self.slots)
        n_gr = sum(_round_up(s.param.numel(), _ALIGN) for s in self.slots)
        self._ghat_flat = torch.zeros(n_gh, device=self.device, dtype=F32)
        self.grad_flat = torch.zeros(n_gr, device=self.device, dtype=F32)
        og, orr = aux, 0
        for s in self.slots:
            s.ghat = self._ghat_flat[og:og + s.rows * s.kpad].view(s.rows, s.kpad)
            s.grad = self.grad_flat[orr:orr + s.param.numel()].view_as(s.param)
            og += _round_up(s.rows * s.kpad, _ALIGN)
            orr += _round_up(s.param.numel(), _ALIGN)
        self._have_grad_buffers = True
        self._table = None

    def begin_backward(self) -> None:
        """Call before a backward writes into the gradient buffers. `param.grad` tensors handed out by an earlier
        backward are views of the current buffer; if any is still alive (gradient accumulation without zero_grad), the
        bank moves to a fresh buffer instead of overwriting them."""
        self.ensure_grad_buffers()
        if any(s.param.grad is not None for s in self.slots):
            self.fresh_grad_buffer()

    def autograd_grads(self) -> list[Tensor]:
        """One NEW view object per slot (sole owner -> autograd's AccumulateGrad adopts it instead of cloning; the
        underlying pointers stay the same from step to step, so the optimiser's descriptor table stays valid)."""
        return [s.grad.view(s.grad.shape) for s in self.slots]

    def fresh_grad_buffer(self) -> None:
        """Detaches the current parameter-gradient buffer (views of it may live on as `param.grad`) and allocates a
        new one for the next backward."""
        if not self._have_grad_buffers:
            return
        self.grad_flat = torch.zeros_like(self.grad_flat)
        orr = 0
        for s in self.slots:
            s.grad = self.grad_flat[orr:orr + s.param.numel()].view_as(s.param)
            orr += _round_up(s.param.numel(), _ALIGN)
        self._table = None

    # ---- descriptor table ----
    def _upload(self, raw: bytes) -> Tensor:
        """Host struct array -> device, from pinned memory and asynchronously (a pageable copy would stall the host
        until the stream drains and destroy its run-ahead over the GPU)."""
        host = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
        self._pinned = host   # keep alive until the copy has executed
        return host.to(self.device, non_blocking=True)

    def _build_table(self) -> None:
        arr = (_lib.WeightDesc * len(self.slots))()
        group = 0
        for d, s in zip(arr, self.slots):
            d.group_start = group
            group += (s.rows + 15) // 16
            p = s.param
            if not (p.is_cuda and p.dtype == F32 and p.is_contiguous()):
                raise RuntimeError(f"tinyedm_b200: parameter {s.name} must be a contiguous fp32 CUDA tensor")
            d.w = p.data_ptr()
            d.grad = s.grad.data_ptr() if s.grad is not None else None
            d.g_hat = s.ghat.data_ptr() if s.ghat is not None else None
            d.out_fwd = s.fwd.data_ptr() if s.fwd is not None else None
            d.out_dgrad = s.dgrad.data_ptr() if s.dgrad is not None else None
            d.out_f32 = s.f32.data_ptr() if s.f32 is not None else None
            d.stats = self.stats.data_ptr() + 16 * s.row_start
            d.rows, d.cin, d.taps, d.kpad, d.row_start = s.rows, s.cin, s.taps, s.kpad, s.row_start
            d.qkv_head_dim = s.qkv_head_dim
        self._table = self._upload(bytes(arr))
        self._table_ptrs = tuple(s.param.data_ptr() for s in self.slots)

    def _table_current(self) -> bool:
        return self._table is not None and self._table_ptrs == tuple(s.param.data_ptr() for s in self.slots)

    # ---- launches ----
    def prepare(self, training: bool) -> None:
        """Normalises every weight (one launch). Training mode always runs (and rewrites the parameters in place)."""
        if not self._table_current():
            self._build_table()
            self._key = None
        if training:
            # Forced normalisation of an already normalised weight is the identity (up to fp32 rounding): when no parameter
            # changed since the last training-mode pass — the 2nd and 3rd micro-batch of an accumulated step — the
            # prepared operands and the in-place rewrite are already what another pass would produce
            key = ("train", _WEIGHTS_EPOCH[0]) + tuple(s.param._version for s in self.slots)
            if key == self._key:
                return
            ops.weight_prep_forward(self._table, len(self.slots), self.total_groups, True)
            # the kernel rewrote every parameter in place (networks.py:32-34): let autograd / EMA code see the mutation
            torch.autograd.graph.increment_version([s.param for s in self.slots])
            self._key = ("train", _WEIGHTS_EPOCH[0]) + tuple(s.param._version for s in self.slots)
            return
        key = (_WEIGHTS_EPOCH[0],) + tuple(s.param._version for s in self.slots)
        if key != self._key and self._key != ("train",) + key:      # (operands prepared by a training pass are current too)
            ops.weight_prep_forward(self._table, len(self.slots), self.total_groups, False)
        self._key = key

    def invalidate(self) -> None:
        self._key = None

    def backward(self) -> None:
        """g_hat -> gradient of the raw parameter for every slot (one launch)."""
        if not self._table_current():
            self._build_table()
        ops.weight_prep_backward(self._table, len(self.slots), self.total_rows, self.max_row_floats)


def conv_slot(name: str, param, *, dgrad: bool = True) -> WeightSlot:
    rows, cin, kh, kw = param.shape
    taps = kh * kw
    return WeightSlot(name, param, rows, cin, taps, kpad=taps * cin, want_fwd=True, want_dgrad=dgrad)


def f32_slot(name: str, param) -> WeightSlot:
    rows = param.shape[0]
    fan_in = param[0].numel()
    return WeightSlot(name, param, rows, fan_in, 1, kpad=fan_in, want_f32=True)


# =========================================================================================================
# Denoiser plan
# =========================================================================================================
@dataclass
class BlockPlan:
    name: str
    kind: str             # "enc" | "dec"
    index: int            # block id over enc+dec (modulation column block, dropout salt)
    cin: int
    cout: int
    cskip: int = 0
    resample: int = RESAMPLE_NONE
    attn: bool = False
    col0: int = 0         # first modulation column
    add_t: float = 0.3
    skip_src: int = -1    # index into the skip list this decoder block pops
    in_src: int = -1      # index into the skip list that is this block's input (if its input is a skip source)
    w: dict = field(default_factory=dict)  # name -> WeightSlot
    gain: torch.nn.Parameter | None = None
    module: torch.nn.Module | None = None  # the EncoderBlock / DecoderBlock that owns the parameters (forward hooks)


def fused_skip_mean_enabled() -> bool:
    """TEDM_FUSED_SKIP_MEAN=0: ScaleLong's mean always through channel_dot (A/B switch)."""
    return os.environ.get("TEDM_FUSED_SKIP_MEAN", "1") != "0"


def generic_attention_enabled() -> bool:
    """TEDM_ATTN_GENERIC=0: shapes outside (head_dim 64, S in {64, 256}) run the warp-MMA kernels (A/B switch)."""
    return os.environ.get("TEDM_ATTN_GENERIC", "1") != "0"


def split_epilogue_enabled() -> bool:
    """TEDM_SPLIT_EPILOGUE=0 keeps the separate concat-split / gain-gradient kernels in the backward (A/B switch)."""
    return os.environ.get("TEDM_SPLIT_EPILOGUE", "1") != "0"


class DenoiserEngine:
    """Plans and runs `Denoiser.forward` / its adjoint for one `tinyedm_b200.networks.Denoiser` module."""

    def __init__(self, module):
        self.m = module
        self.bank: WeightBank | None = None
        self.blocks: list[BlockPlan] = []
        self._aux_key = None
        self.grad_sync = None   # set by parallel.DistributedEDM: overlaps the gradient all-reduce with this backward
        # gradient accumulation (imagenet.yaml:7 `accumulate_grad_batches: 3`): while `defer_grads` is set a backward only
        # ADDS its dL/dw_hat to the flat g_hat buffer (which also holds the 0-d gradients); the first backward with it
        # cleared applies the weight-norm Jacobian to the sum, exchanges it between ranks and hands out the gradients
        self.defer_grads = False
        self._accum = 0         # micro-batches already summed into g_hat since the last hand-out
        self._sg_alive = False  # 0-d `.grad`s of an earlier backward are still referenced (plain autograd accumulation)
        self._build_plan()

    # ---- static plan ----
    def _build_plan(self) -> None:
        m = self.m
        slots: list[WeightSlot] = []
        if 9 * (m.in_channels + 1) > 64:
            raise RuntimeError("tinyedm_b200: conv_in supports at most 6 image channels")
        w_in = m.conv_in.weight
        self.s_in = WeightSlot("conv_in.weight", w_in, w_in.shape[0], w_in.shape[1], 9, kpad=64, want_fwd=True)
        w_out = m.conv_out.weight
        self.s_out = WeightSlot("conv_out.weight", w_out, w_out.shape[0], w_out.shape[1], 1, kpad=w_out.shape[1], want_fwd=True)
        slots += [self.s_in, self.s_out]
        embed_slots = []
        col = 0
        idx = 0
        n_enc = len(m.encoder_blocks)

        def common(bp: BlockPlan, blk, pre: str) -> None:
            nonlocal col, idx
            for nm in ("conv_1x1", "conv_3x3_1", "conv_3x3_2"):
                mod = getattr(blk, nm)
                if hasattr(mod, "weight"):
                    bp.w[nm] = conv_slot(pre + nm + ".weight", mod.weight)
            if bp.attn:
                bp.w["qkv"] = conv_slot(pre + "attention.qkv_conv.weight", blk.attention.qkv_conv.weight)
                bp.w["qkv"].qkv_head_dim = bp.cout // m.num_heads
                bp.w["out"] = conv_slot(pre + "attention.out_conv.weight", blk.attention.out_conv.weight)
            bp.w["embed"] = f32_slot(pre + "embed.weight", blk.embed.weight)
            bp.gain = blk.gain
            bp.module = blk
            bp.col0 = col
            bp.index = idx
            col += bp.cout
            idx += 1

        cin = m.encoder_out_channels[0]
        for i, (blk, typ, cout) in enumerate(zip(m.encoder_blocks, m.encoder_block_types, m.encoder_out_channels)):
            bp = BlockPlan(f"encoder_blocks.{i}", "enc", 0, cin, cout, resample=RESAMPLE_DOWN if typ.endswith("D") else RESAMPLE_NONE,
                           attn=typ.endswith("A"), add_t=m.encoder_add_factor, in_src=i)
            common(bp, blk, f"encoder_blocks.{i}.")
            self.blocks.append(bp)
            cin = cout
        from .networks import get_skip_channels
        skip_ch = get_skip_channels(m.encoder_out_channels, m.decoder_out_channels, m.skip_connections)
        cin = m.decoder_out_channels[0]
        next_skip = n_enc  # skips = [conv_in, enc0, ..., enc_{n-1}], popped from the end
        for i, (blk, typ, cout, cs) in enumerate(zip(m.decoder_blocks, m.decoder_block_types, m.decoder_out_channels, skip_ch)):
            bp = BlockPlan(f"decoder_blocks.{i}", "dec", 0, cin, cout, cskip=int(cs),
                           resample=RESAMPLE_UP if typ.endswith("U") else RESAMPLE_NONE, attn=typ.endswith("A"),
                           add_t=m.decoder_add_factor, in_src=n_enc if i == 0 else -1)
            if cs > 0:
                bp.skip_src = next_skip
                next_skip -= 1
                bp.w["sl1"] = f32_slot(f"decoder_blocks.{i}.cat_factor.layer1.weight", blk.cat_factor.layer1.weight)
                bp.w["sl2"] = f32_slot(f"decoder_blocks.{i}.cat_factor.layer2.weight", blk.cat_factor.layer2.weight)
            common(bp, blk, f"decoder_blocks.{i}.")
            self.blocks.append(bp)
            cin = cout
        self.n_mod = col
        self.n_enc = n_enc
        # embed weights are laid out back to back (in block order) so that ONE GEMM serves every block
        for bp in self.blocks:
            embed_slots.append(bp.w["embed"])
        for bp in self.blocks:
            for k, s in bp.w.items():
                if k != "embed":
                    slots.append(s)
        self.embed_slots = embed_slots
        if m.embedding_dim % 2 != 0:
            raise RuntimeError("tinyedm_b200: embedding_dim must be even")
        self.bank = WeightBank(embed_slots + slots)
        # the gradients of the 0-d parameters (block gains..., gain_out) live at the front of the flat g_hat buffer: the
        # one memset of a backward clears them and the trailing data-parallel message carries them (no collective of their own)
        self.bank.aux_elems = len(self.blocks) + 1
        for c in (list(m.encoder_out_channels) + list(m.decoder_out_channels)):
            if c % 64 != 0:
                raise RuntimeError(f"tinyedm_b200: channel counts must be multiples of 64 (got {c})")

    # ---- device-side auxiliary tables (pointers to the block gains, column -> block maps) ----
    def _ensure_device(self, dev: torch.device) -> None:
        ops.ensure_device(dev)
        if self.bank.device != dev:
            self.bank.materialise(dev)
            self._aux_key = None
        key = tuple(bp.gain.data_ptr() for bp in self.blocks)
        if key != self._aux_key:
            self.gain_ptrs = torch.tensor(list(key), dtype=torch.int64).to(dev)
            cb, starts = [], [0]
            for j, bp in enumerate(self.blocks):
                cb += [j] * bp.cout
                starts.append(starts[-1] + bp.cout)
            self.col_block = torch.tensor(cb, dtype=torch.int32).to(dev)
            self.blk_start = torch.tensor(starts, dtype=torch.int32).to(dev)
            self.step_counter = torch.zeros(1, dtype=torch.int64, device=dev)
            self._aux_key = key

    @property
    def w_embed_all(self) -> Tensor:
        """fp32 w_hat of every block's `embed` Linear, rows back to back: (sum C, E)."""
        first = self.embed_slots[0]
        E = first.cin
        base = first.f32.storage_offset() - self.bank._f32_flat.storage_offset()
        return self.bank._f32_flat[base:base + self.n_mod * E].view(self.n_mod, E)

    # =====================================================================================================
    # forward
    # =====================================================================================================
    def forward(self, noisy: Tensor, sigma: Tensor, emb: Tensor, *, training: bool, save: bool,
                taps: dict | None = None):
        """Returns (D, ctx). `ctx` (only when save) holds what the adjoint needs."""
        m = self.m
        ops.check(noisy, F32, "noisy_image")
        ops.check(emb, F32, "embedding")
        sigma = sigma.reshape(-1)
        ops.check(sigma, F32, "sigma")
        B, Ci, H, W = noisy.shape
        if Ci != m.in_channels:
            raise RuntimeError(f"tinyedm_b200: expected {m.in_channels} image channels, got {Ci}")
        if sigma.numel() not in (1, B):
            raise RuntimeError(f"tinyedm_b200: sigma must have 1 or {B} elements, got {sigma.numel()}")
        Be, E = emb.shape
        if Be not in (1, B) or (save and Be != B):
            raise RuntimeError(f"tinyedm_b200: embedding must have {B} rows (or 1 without grad), got {Be}")
        if E != m.embedding_dim:
            raise RuntimeError(f"tinyedm_b200: embedding width {E} != embedding_dim {m.embedding_dim}")
        dev = noisy.device
        self._ensure_device(dev)
        self.bank.prepare(training)
        if save:
            self.bank.ensure_grad_buffers()
        drop_p = float(m.dropout_rate) if training else 0.0
        seed_t = self.step_counter
        if training and drop_p > 0:
            self.step_counter += 1
            # per-forward snapshot: the backward regenerates THIS forward's masks even if another training forward ran
            # in between (two micro-batches, then their backwards)
            seed_t = self.step_counter.clone() if save else self.step_counter

        # ---- modulation for every block: m = embed(emb) * gain + 1 (networks.py:255-258, :319-322) ----
        N = self.n_mod
        lin = torch.empty((Be, N), device=dev, dtype=F32)
        ops.sgemm(emb, self.w_embed_all, lin, Be, N, E, E, E, N, False, True)
        mod = ops.mod_finish_forward(lin, self.gain_ptrs, self.col_block)
        mod_stride = N if Be == B else 0

        # ---- input block (networks.py:578-587) ----
        # the Diffuser's kernel already wrote this operand when it produced (noisy, sigma) (edm.Diffuser, row N2)
        pre = getattr(noisy, "_tedm_xcol", None)
        if (pre is not None and pre[1] == float(m.sigma_data) and pre[2] == sigma.data_ptr()
                and tuple(pre[0].shape) == (B, H, W, 64) and pre[0].device == dev):
            xcol = pre[0]
        else:
            xcol = ops.conv_in_im2col(noisy, sigma, float(m.sigma_data))
        C0 = m.encoder_out_channels[0]
        used = self._skips_consumed()
        # ScaleLong's per-(image, channel) mean of a skip tensor (networks.py:112) comes out of the epilogue of the conv that
        # produces the tensor (fixed-order partial sums, no atomics) when that launch runs on the CTA-pair kernel; otherwise
        # it is reduced by channel_dot right after the tensor exists, while it is still (partly) in L2
        x, mean0 = self._conv_with_mean(xcol, self.s_in.fwd, 1, C0, used[0])
        if taps is not None:
            taps["conv_in"] = x
        ctx = None
        if save:
            ctx = dict(B=B, H=H, W=W, noisy=noisy, sigma=sigma, emb=emb, lin=lin, mod=mod, xcol=xcol, blocks=[],
                       drop_p=drop_p, Be=Be, seed=seed_t)
        skips = [x]
        means = [mean0]
        n_pushed = 1
        for bp in self.blocks:
            skip = mean = None
            if bp.kind == "dec" and bp.cskip > 0:
                skip = skips.pop()
                mean = means.pop()
            want_mean = bp.kind == "enc" and used[n_pushed]
            x_in = x
            x, saved, out_mean = self._block_forward(bp, x, skip, mod, mod_stride, drop_p, save, mean, want_mean, seed_t)
            if bp.module is not None and bp.module._forward_hooks:
                # The block ran fused, not through its module's __call__: forward hooks registered on the EncoderBlock /
                # DecoderBlock (activation logging, `wandb.watch`-style taps) still fire, with the reference's argument
                # order (input, embedding[, skip]) and the block output as NCHW views of the bf16 NHWC tensors
                nchw = lambda t: t.permute(0, 3, 1, 2)
                hook_args = (nchw(x_in), emb) + ((nchw(skip),) if skip is not None else ())
                for hook in list(bp.module._forward_hooks.values()):
                    hook(bp.module, hook_args, nchw(x))
            if bp.kind == "enc":
                skips.append(x)
                means.append(out_mean)
                n_pushed += 1
            if save:
                ctx["blocks"].append(saved)
            if taps is not None:
                taps[bp.name] = x
        D, f_raw = ops.conv_out_forward(x, self.s_out.fwd, m.gain_out, noisy, sigma, float(m.sigma_data), keep_raw=save)
        if save:
            ctx["x_last"] = x
            ctx["f_raw"] = f_raw
        return D, ctx

    def _skips_consumed(self) -> list[bool]:
        """used[i]: the i-th tensor pushed on the skip stack (conv_in output, then every encoder block's) is popped by a
        decoder block with a skip connection."""
        if getattr(self, "_skip_used", None) is None:
            stack, used, n = [0], {}, 1
            for bp in self.blocks:
                if bp.kind == "dec" and bp.cskip > 0:
                    used[stack.pop()] = True
                if bp.kind == "enc":
                    stack.append(n)
                    n += 1
            self._skip_used = [used.get(i, False) for i in range(n)]
        return self._skip_used

    @staticmethod
    def _conv_with_mean(x: Tensor, w: Tensor, ksize: int, cout: int, want_mean: bool, **kw):
        """conv2d (epilogue PLAIN or AXPBY) that also returns the spatial mean of its output per (image, channel)."""
        if not want_mean:
            return ops.conv2d(x, w, ksize, cout, **kw), None
        B, H, W, cin = x.shape
        slots = ops.conv2d_colsum_slots(B, H, W, cin, cout, ksize, kw.get("epi", 0)) if fused_skip_mean_enabled() else 0
        if slots == 0:
            out = ops.conv2d(x, w, ksize, cout, **kw)
            return out, DenoiserEngine._skip_mean(out)
        part = torch.empty((B * slots, cout), device=x.device, dtype=F32)
        out = ops.conv2d(x, w, ksize, cout, col_partial=part, **kw)
        return out, ops.colsum_mean(part, B, slots, cout, 1.0 / (H * W))

    @staticmethod
    def _skip_mean(x: Tensor) -> Tensor:
        B, H, W, C = x.shape
        mean = torch.zeros((B, C), device=x.device, dtype=F32)
        ops.channel_dot(x, None, mean, C, 0, 1.0 / (H * W))
        return mean

    def _block_forward(self, bp: BlockPlan, xin: Tensor, skip: Tensor | None, mod: Tensor, mod_stride: int,
                       drop_p: float, save: bool, skip_mean: Tensor | None = None, want_mean: bool = False,
                       seed_t: Tensor | None = None):
        """Returns (block output, tensors saved for backward, spatial mean of the output per (image, channel) or None)."""
        S: dict = {}
        B, Hin, Win, _ = xin.shape
        S["in_shape"] = (B, Hin, Win)
        res_nrm = None
        if bp.kind == "enc":
            if "conv_1x1" not in bp.w and bp.resample == RESAMPLE_NONE and not save:
                # inference: x = pixel_norm(xin) is never stored — only mp_silu(x) and the per-pixel norm; conv_3x3_2's
                # mp_add epilogue takes the residual as xin / nrm (one activation-sized write less per encoder block)
                _, a, res_nrm = ops.block_prep(xin, pixelnorm=True, want_x=False, want_nrm=True)
                x, nrm = xin, None
            elif "conv_1x1" not in bp.w:
                x, a, nrm = ops.block_prep(xin, resample=bp.resample, pixelnorm=True, want_nrm=save)
            else:
                r = xin
                if bp.resample != RESAMPLE_NONE:
                    r, _, _ = ops.block_prep(xin, resample=bp.resample, want_a=False)
                u = ops.conv2d(r, bp.w["conv_1x1"].fwd, 1, bp.cout)
                x, a, nrm = ops.block_prep(u, pixelnorm=True, want_nrm=save)
                S["r"] = r
            xr = x
            S["nrm"] = nrm
        else:
            if skip is not None:
                Cs = bp.cskip
                mean = skip_mean if skip_mean is not None else self._skip_mean(skip)
                aug, h_pre, hh, gain = ops.scalelong_forward(mean, bp.w["sl1"].f32, bp.w["sl2"].f32, bp.w["sl1"].rows)
                x, a, _ = ops.block_prep(xin, skip=skip, gain=gain, resample=bp.resample)
                S.update(skip=skip, aug=aug, h_pre=h_pre, hh=hh, gain=gain)
            elif bp.resample != RESAMPLE_NONE:
                x, a, _ = ops.block_prep(xin, resample=bp.resample)
            else:
                x = xin
                _, a, _ = ops.block_prep(xin, want_x=False)
            xr = ops.conv2d(x, bp.w["conv_1x1"].fwd, 1, bp.cout) if "conv_1x1" in bp.w else x
        wa, wb = ops.mp_add_coeffs(bp.add_t)
        raw = torch.empty((x.shape[0], x.shape[1], x.shape[2], bp.cout), device=x.device, dtype=BF16) if save else None
        h = ops.conv2d(a, bp.w["conv_3x3_1"].fwd, 3, bp.cout, epi=EPI_MODSILU, mod=mod, mod_off=bp.col0,
                       mod_stride=mod_stride, drop_p=drop_p, seed=0x5EED0000 + bp.index,
                       seed_ptr=seed_t if seed_t is not None else self.step_counter, raw=raw)
        out, out_mean = self._conv_with_mean(h, bp.w["conv_3x3_2"].fwd, 3, bp.cout, want_mean and not bp.attn, epi=EPI_AXPBY,
                                             alpha=wb, beta=wa, res=xr, nrm=res_nrm)
        if save:
            S.update(x=x, a=a, raw=raw, h=h)
        if bp.attn:
            c5 = 1.0 / math.sqrt(2.0)
            qkv = ops.conv2d(out, bp.w["qkv"].fwd, 1, 3 * bp.cout)
            heads = self.m.num_heads
            S_att = qkv.shape[1] * qkv.shape[2]
            specialised = ops.attention_specialised(S_att, bp.cout // heads)
            if specialised and save and S_att == 64 and os.environ.get("TEDM_ATTN_S64_GENERIC", "1") != "0":
                # training at S = 64 (CIFAR 8x8): the packed generic backward (two (image, head) pairs per tile, two CTAs per
                # SM) beats the specialised two-kernel one by more than the normalisation pass costs the forward
                specialised = False
            if specialised or not generic_attention_enabled():
                y, lse = ops.attention_forward(qkv, heads, need_lse=save)      # normalises q, k, v inside the kernel
                if save:
                    S.update(qkv=qkv)
            else:
                qn, norms = ops.qkv_normalize(qkv, heads)                       # once per layer; the backward reuses it
                y, lse = ops.attention_forward_normalized(qn, heads, need_lse=save)
                if save:
                    S.update(qn=qn, norms=norms)
            out2, out_mean = self._conv_with_mean(y, bp.w["out"].fwd, 1, bp.cout, want_mean, epi=EPI_AXPBY, alpha=c5, beta=c5,
                                                  res=out)
            if save:
                S.update(mid=out, lse=lse, y=y)
            out = out2
        return out, S, out_mean

    # =====================================================================================================
    # backward
    # =====================================================================================================
    def ghat_units_backward_order(self) -> list[tuple[int, int]]:
        """(start, end) element ranges of the flat g_hat buffer in the order `backward` completes them: the blocks from
        last to first, then one trailing unit with everything in front of the first block (embed weights of all
        blocks, conv_in, conv_out). Together they tile the whole buffer from its end to its start."""
        self.bank.ensure_grad_buffers()
        base = self.bank._ghat_flat.data_ptr()
        def span(slots):
            starts = [(s.ghat.data_ptr() - base) // 4 for s in slots]
            return min(starts), max(st + _round_up(s.rows * s.kpad, _ALIGN) for st, s in zip(starts, slots))
        units = []
        for bp in reversed(self.blocks):
            units.append(span([s for k, s in bp.w.items() if k != "embed"]))
        first_block_start = units[-1][0]
        units.append((0, first_block_start))
        assert units[0][1] == self.bank._ghat_flat.numel()
        return units

    def backward(self, ctx: dict, g_D: Tensor, *, need_g_emb: bool = True, defer: bool = False):
        """Adjoint of `forward`. Fills every WeightSlot.grad, returns (g_emb, scalar_grads).

        scalar_grads: fp32 tensor [n_blocks + 1] = d(block gains..., gain_out).
        With `self.grad_sync` set (parallel.DistributedEDM) the g_hat regions are handed to the gradient all-reduce
        as soon as their block is done, so the exchange overlaps the rest of this backward.
        `defer`: accumulate into g_hat / the scalar gradients only and return (g_emb, None): no Jacobian, no exchange.
        The weight-norm Jacobian is linear in g_hat and the weights do not move between the micro-batches of one
        optimiser step (the forced re-normalisation of an already normalised weight is the identity up to fp32
        rounding), so J(sum g_hat) == sum J(g_hat): one Jacobian launch and one exchange per step instead of one per
        micro-batch (SURVEY.md §8e "with accumulation reduce only on the last micro-batch").
        """
        m = self.m
        ops.check(g_D, F32, "grad of denoised_image")
        B, H, W = ctx["B"], ctx["H"], ctx["W"]
        dev = g_D.device
        sigma = ctx["sigma"]
        nb = len(self.blocks)
        sg = self.bank._ghat_flat[:nb + 1]          # 0-d gradients: front of the flat g_hat buffer (see _build_plan)
        sync = None if defer else self.grad_sync
        if sync is not None:
            sync.backward_started()
        if self._accum == 0:
            # ONE memset for every dL/dw_hat of the network (and the 0-d gradients); all weight-gradient kernels then
            # accumulate (split-K partial sums via TMA reduce-add / atomics) instead of zeroing their own slice with
            # 70+ small memsets
            self.bank._ghat_flat.zero_()
        g = ops.conv_out_backward(g_D, ctx["f_raw"], ctx["x_last"], self.s_out.fwd, m.gain_out, sigma,
                                  float(m.sigma_data), self.s_out.ghat, sg[nb:])
        d_mod = torch.zeros((B, self.n_mod), device=dev, dtype=F32)
        pending: dict[int, tuple] = {}   # skip index -> (g_skip, pending per-(image, channel) share or None, its scale)
        for bp, S in zip(reversed(self.blocks), reversed(ctx["blocks"])):
            g = self._block_backward(bp, S, g, ctx, d_mod, pending)
            if sync is not None:
                sync.unit_done()
        # conv_in weight gradient (its input is the image: no data gradient needed)
        ops.conv2d_wgrad(g, ctx["xcol"], self.s_in.ghat, 1, accumulate=True)
        g_emb = self._modulation_backward(ctx, d_mod, sg, need_g_emb)
        if defer:
            self._accum += 1
            return g_emb, None
        self._accum = 0
        if sync is not None:
            sync.unit_done()                 # trailing unit: embed weights, conv_in, conv_out
            sync.before_weight_jacobian()
        self.bank.backward()
        if self._sg_alive:
            sg = sg.clone()      # 0-d `.grad`s handed out by an earlier backward still alias the buffer: do not share it
        return g_emb, sg

    def _modulation_backward(self, ctx: dict, d_mod: Tensor, sg: Tensor, need_g_emb: bool):
        """Adjoint of `m = embed(emb) * gain + 1` for every block at once: d_mod -> block gains (into sg), the embed
        weights' dL/dw_hat (added to the flat g_hat buffer) and the embedding gradient (returned)."""
        nb = len(self.blocks)
        d_lin = ops.mod_finish_backward(ctx["lin"], d_mod, self.gain_ptrs, self.blk_start, sg, nb)
        emb = ctx["emb"]
        B, E = emb.shape
        N = self.n_mod
        g_emb = None
        if need_g_emb:
            g_emb = torch.empty((B, E), device=emb.device, dtype=F32)
            ops.sgemm(d_lin, self.w_embed_all, g_emb, B, E, N, N, E, E, False, False)
        first = self.embed_slots[0]
        base = first.ghat.storage_offset() - self.bank._ghat_flat.storage_offset()
        gh_all = self.bank._ghat_flat[base:base + N * E].view(N, E)
        ops.sgemm(d_lin, emb, gh_all, N, E, B, N, E, E, True, False, 1.0, 1.0)   # += : g_hat was zeroed by the memset
        return g_emb

    def _take_g_in(self, bp: BlockPlan, pending: dict, shape, dev):
        """Buffer receiving the gradient w.r.t. this block's input. If the input is also a skip source whose
        decoder already deposited its share, accumulate into that buffer."""
        g, acc, bias, scale = self._take_g_in_bias(bp, pending, shape, dev)
        if bias is not None:
            ops.bias_add_bc(g, bias, scale)      # this consumer's kernel cannot fold the per-(image, channel) share in
        return g, acc

    def _take_g_in_bias(self, bp: BlockPlan, pending: dict, shape, dev):
        """As `_take_g_in`, for consumers whose kernel adds a pending per-(image, channel) share itself (`out_bias` of the
        SILU_BWD conv epilogue): returns (buffer, accumulate, bias or None, bias scale). The share is the gradient through
        ScaleLong's spatial mean, deposited with the decoder's g_skip instead of being added by a pass of its own."""
        if bp.in_src >= 0 and bp.in_src in pending:
            g, bias, scale = pending.pop(bp.in_src)
            return g, True, bias, scale
        return torch.empty(shape, device=dev, dtype=BF16), False, None, 1.0

    def _attn_backward(self, bp: BlockPlan, S: dict, g_out: Tensor) -> Tensor:
        c5 = 1.0 / math.sqrt(2.0)
        g_y = ops.conv2d(g_out, bp.w["out"].dgrad, 1, bp.cout, alpha=c5)
        ops.conv2d_wgrad(g_out, S["y"], bp.w["out"].ghat, 1, alpha=c5, accumulate=True)
        if "qn" in S:
            g_qkv = ops.attention_backward_normalized(S["qn"], S["norms"], S["y"], g_y, S["lse"], self.m.num_heads)
        else:
            g_qkv = ops.attention_backward(S["qkv"], S["y"], g_y, S["lse"], self.m.num_heads)
        g_mid = ops.conv2d(g_qkv, bp.w["qkv"].dgrad, 1, bp.cout, epi=EPI_AXPBY, alpha=1.0, beta=c5, res=g_out)
        ops.conv2d_wgrad(g_qkv, S["mid"], bp.w["qkv"].ghat, 1, accumulate=True)
        return g_mid

    def _block_backward(self, bp: BlockPlan, S: dict, g_out: Tensor, ctx: dict, d_mod: Tensor, pending: dict) -> Tensor:
        dev = g_out.device
        B, Hin, Win = S["in_shape"]
        wa, wb = ops.mp_add_coeffs(bp.add_t)
        w1, w2 = bp.w["conv_3x3_1"], bp.w["conv_3x3_2"]
        g_mid = self._attn_backward(bp, S, g_out) if bp.attn else g_out
        # out = wb * conv2(h) + wa * xr, h = drop(mp_silu(raw * m)): the adjoint of the modulation/silu/dropout epilogue
        # (incl. the d_mod reduction) runs in the epilogue of conv2's data gradient
        g_raw = ops.conv2d(g_mid, w2.dgrad, 3, bp.cout, epi=EPI_MODSILU_BWD, alpha=wb, aux=S["raw"], mod=ctx["mod"],
                           mod_off=bp.col0, d_mod=d_mod, drop_p=ctx["drop_p"], seed=0x5EED0000 + bp.index,
                           seed_ptr=ctx.get("seed", self.step_counter))
        ops.conv2d_wgrad(g_mid, S["h"], w2.ghat, 3, alpha=wb, accumulate=True)
        ops.conv2d_wgrad(g_raw, S["a"], w1.ghat, 3, accumulate=True)
        x = S["x"]
        if bp.kind == "enc":
            Hh, Ww = ops.resampled_hw(Hin, Win, bp.resample)
            has_1x1 = "conv_1x1" in bp.w
            fuse_pn = bp.cout <= 256          # the fused pixel-norm adjoint needs all channels in one N tile
            if fuse_pn and not has_1x1 and bp.resample == RESAMPLE_NONE:
                # conv1's data gradient, mp_silu', the residual share and the pixel-norm adjoint in ONE kernel
                g_in, acc, bias, bscale = self._take_g_in_bias(bp, pending, (B, Hin, Win, bp.cin), dev)
                ops.conv2d(g_raw, w1.dgrad, 3, bp.cout, epi=EPI_SILU_BWD, aux=x, res=g_mid, beta=wa, nrm=S["nrm"], out=g_in,
                           accumulate_out=acc, out_bias=bias, out_bias_scale=bscale)
                return g_in
            if fuse_pn:
                g_u = ops.conv2d(g_raw, w1.dgrad, 3, bp.cout, epi=EPI_SILU_BWD, aux=x, res=g_mid, beta=wa, nrm=S["nrm"])
            else:
                g_a = ops.conv2d(g_raw, w1.dgrad, 3, bp.cout)
                g_u = torch.empty((B, Hh, Ww, bp.cout), device=dev, dtype=BF16)
                ops.block_prep_backward(g_res=g_mid, beta=wa, g_a=g_a, x=x, nrm=S["nrm"], gain=None, d_mean=None, g_in=g_u,
                                        g_skip=None, accumulate_in=False, accumulate_skip=False, B=B, Hin=Hh, Win=Ww,
                                        C1=bp.cout, C2=0, resample=RESAMPLE_NONE, pixelnorm=True)
            if has_1x1:
                g_r = ops.conv2d(g_u, bp.w["conv_1x1"].dgrad, 1, bp.cin)
                ops.conv2d_wgrad(g_u, S["r"], bp.w["conv_1x1"].ghat, 1, accumulate=True)
            else:
                g_r = g_u
            g_in, acc = self._take_g_in(bp, pending, (B, Hin, Win, bp.cin), dev)
            if bp.resample == RESAMPLE_NONE and not acc:
                return g_r
            ops.block_prep_backward(g_res=g_r, beta=1.0, g_a=None, x=None, nrm=None, gain=None, d_mean=None, g_in=g_in,
                                    g_skip=None, accumulate_in=acc, accumulate_skip=False, B=B, Hin=Hin, Win=Win,
                                    C1=bp.cin, C2=0, resample=bp.resample, pixelnorm=False)
            return g_in
        # ---- decoder: x = resample(cat(in, skip * gain)), a = mp_silu(x), xr = conv_1x1(x) | x ----
        ctot = bp.cin + bp.cskip
        if "conv_1x1" in bp.w:
            g_res = ops.conv2d(g_mid, bp.w["conv_1x1"].dgrad, 1, ctot, alpha=wa)
            ops.conv2d_wgrad(g_mid, x, bp.w["conv_1x1"].ghat, 1, alpha=wa, accumulate=True)
            beta = 1.0
        else:
            g_res, beta = g_mid, wa
        if bp.cskip == 0 and bp.resample == RESAMPLE_NONE:
            g_in, acc, bias, bscale = self._take_g_in_bias(bp, pending, (B, Hin, Win, bp.cin), dev)
            ops.conv2d(g_raw, w1.dgrad, 3, ctot, epi=EPI_SILU_BWD, aux=x, res=g_res, beta=beta, out=g_in, accumulate_out=acc,
                       out_bias=bias, out_bias_scale=bscale)
            return g_in
        if (bp.cskip > 0 and bp.resample == RESAMPLE_NONE and split_epilogue_enabled()
                and ops.conv2d_dgrad_split_supported(B, Hin, Win, bp.cout, bp.cin, bp.cskip, 3)):
            # conv1's data gradient, mp_silu', the residual share, the concat split, the ScaleLong gain and the reduction
            # for d(gain) in ONE kernel: g_cat is never materialised (it used to be written once and read twice)
            Cs = bp.cskip
            g_in, acc, bias, bscale = self._take_g_in_bias(bp, pending, (B, Hin, Win, bp.cin), dev)
            g_skip = torch.empty((B, Hin, Win, Cs), device=dev, dtype=BF16)
            d_gx = torch.zeros((B, Cs), device=dev, dtype=F32)
            ops.conv2d_dgrad_split(g_raw, w1.dgrad, 3, x=x, res=g_res, beta=beta, gain=S["gain"], g_in=g_in, g_skip=g_skip,
                                   d_gx=d_gx, accumulate_in=acc, in_bias=bias, in_bias_scale=bscale)
            s1, s2 = bp.w["sl1"], bp.w["sl2"]
            d_pre2, d_hpre, d_mean = ops.scalelong_backward(d_gx, S["gain"], S["h_pre"], s1.f32, s2.f32, d_gain_times_gain=True)
            ops.scalelong_wgrad(d_pre2, S["hh"], d_hpre, S["aug"], s2.ghat, s1.ghat)   # into the zeroed g_hat buffer
            # the gradient through the spatial mean of [skip, 1] (d_mean / HW on every pixel) travels with g_skip and is
            # added by the kernel that accumulates the encoder side's gradient into it
            pending[bp.skip_src] = (g_skip, d_mean, 1.0 / (Hin * Win))
            return g_in
        # gradient w.r.t. x on the post-resample grid: conv1's data gradient * mp_silu'(x) + the residual share
        g_x = ops.conv2d(g_raw, w1.dgrad, 3, ctot, epi=EPI_SILU_BWD, aux=x, res=g_res, beta=beta)
        if bp.cskip == 0:
            g_in, acc = self._take_g_in(bp, pending, (B, Hin, Win, bp.cin), dev)
            ops.block_prep_backward(g_res=g_x, beta=1.0, g_a=None, x=None, nrm=None, gain=None, d_mean=None, g_in=g_in,
                                    g_skip=None, accumulate_in=acc, accumulate_skip=False, B=B, Hin=Hin, Win=Win,
                                    C1=bp.cin, C2=0, resample=bp.resample, pixelnorm=False)
            return g_in
        # gradient of the concatenated tensor on the pre-resample grid, then the ScaleLong adjoint, then the split
        Cs = bp.cskip
        if bp.resample != RESAMPLE_NONE:
            g_cat = torch.empty((B, Hin, Win, ctot), device=dev, dtype=BF16)
            ops.block_prep_backward(g_res=g_x, beta=1.0, g_a=None, x=None, nrm=None, gain=None, d_mean=None, g_in=g_cat,
                                    g_skip=None, accumulate_in=False, accumulate_skip=False, B=B, Hin=Hin, Win=Win, C1=ctot,
                                    C2=0, resample=bp.resample, pixelnorm=False)
        else:
            g_cat = g_x
        d_gain = torch.zeros((B, Cs), device=dev, dtype=F32)
        ops.channel_dot(g_cat, S["skip"], d_gain, Cs, bp.cin, 1.0)
        s1, s2 = bp.w["sl1"], bp.w["sl2"]
        d_pre2, d_hpre, d_mean = ops.scalelong_backward(d_gain, S["gain"], S["h_pre"], s1.f32, s2.f32)
        ops.scalelong_wgrad(d_pre2, S["hh"], d_hpre, S["aug"], s2.ghat, s1.ghat)   # into the zeroed g_hat buffer
        g_in, acc = self._take_g_in(bp, pending, (B, Hin, Win, bp.cin), dev)
        g_skip = torch.empty((B, Hin, Win, Cs), device=dev, dtype=BF16)
        ops.block_prep_backward(g_res=g_cat, beta=1.0, g_a=None, x=None, nrm=None, gain=S["gain"], d_mean=d_mean,
                                g_in=g_in, g_skip=g_skip, accumulate_in=acc, accumulate_skip=False, B=B, Hin=Hin, Win=Win,
                                C1=bp.cin, C2=Cs, resample=RESAMPLE_NONE, pixelnorm=False)
        pending[bp.skip_src] = (g_skip, None, 1.0)
        return g_in

    # ---- parameter / gradient enumeration in `module.parameters()` order ----
    def begin_backward(self) -> None:
        """See WeightBank.begin_backward; also protects the scalar-gradient buffer."""
        self.bank.begin_backward()
        scalars = [bp.gain for bp in self.blocks] + [self.m.gain_out]
        self._sg_alive = any(p.grad is not None for p in scalars)
        if self._sg_alive and self._accum == 0:
            # plain autograd accumulation (no zero_grad since the last hand-out): the 0-d `.grad`s may alias the front of
            # the g_hat buffer, which this backward is about to clear
            with torch.no_grad():
                for p in scalars:
                    if p.grad is not None:
                        p.grad = p.grad.clone()

    def grads_by_param(self, sg: Tensor) -> dict[int, Tensor]:
        """id(param) -> a NEW view object into the bank's flat gradient buffer / the scalar gradient vector."""
        out = {id(s.param): g for s, g in zip(self.bank.slots, self.bank.autograd_grads())}
        for j, bp in enumerate(self.blocks):
            out[id(bp.gain)] = sg[j].view(())
        out[id(self.m.gain_out)] = sg[len(self.blocks)].view(())
        return out
