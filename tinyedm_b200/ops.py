"""Thin torch-tensor wrappers over the C ABI (include/tinyedm_b200.h).

Tensors are only used as owners of device memory; every function launches on torch's current stream.
Activations are NHWC bf16 `(B, H, W, C)`; see DESIGN.md for the layouts.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from . import _lib

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Tensor | None):
    return None if t is None else t.data_ptr()


def _check(t: Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"tinyedm_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"tinyedm_b200: `{name}` must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"tinyedm_b200: `{name}` must be contiguous")


def ensure_device(t: Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("tinyedm_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    _lib.init_device(t.device.index if t.device.index is not None else torch.cuda.current_device())


def mp_add_coeffs(t: float) -> tuple[float, float]:
    """(weight of a, weight of b) of mp_add(a, b, t) — networks.py:87-88."""
    c = math.sqrt((1 - t) ** 2 + t ** 2)
    return (1 - t) / c, t / c


EPI_PLAIN, EPI_MODSILU, EPI_AXPBY = 0, 1, 2


def conv2d(x: Tensor, w: Tensor, ksize: int, cout: int, *, epi: int = EPI_PLAIN, alpha: float = 1.0, beta: float = 0.0,
           res: Tensor | None = None, raw: Tensor | None = None, mod: Tensor | None = None, mod_off: int = 0,
           drop_p: float = 0.0, seed: int = 0, out: Tensor | None = None, block_n: int = 0) -> Tensor:
    """Implicit-GEMM MPConv (forward or data gradient). `w` is the prepared bf16 weight [cout][k*k][cin]."""
    _check(x, BF16, "x")
    B, H, W, cin = x.shape
    if out is None:
        out = torch.empty((B, H, W, cout), device=x.device, dtype=BF16)
    mod_ptr = None
    mod_stride = 0
    if mod is not None:
        mod_ptr = mod.data_ptr() + 4 * mod_off
        mod_stride = mod.shape[1]
    _lib.call("tedm_conv2d_forward", x.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, W, cin, cout, ksize, epi, alpha,
              _p(raw), _p(res), beta, mod_ptr, mod_stride, drop_p, seed, block_n, _stream())
    return out


def conv2d_wgrad(g: Tensor, x: Tensor, dw: Tensor, ksize: int, *, alpha: float = 1.0, accumulate: bool = False,
                 splits: int = 0) -> None:
    """dw[cout][k*k][cin] (fp32) (+)= alpha * sum_pixels g x."""
    B, H, W, cout = g.shape
    cin = x.shape[3]
    _lib.call("tedm_conv2d_wgrad", g.data_ptr(), x.data_ptr(), dw.data_ptr(), B, H, W, cin, cout, ksize, alpha,
              1 if accumulate else 0, splits, _stream())


def block_prep(inp: Tensor, *, skip: Tensor | None = None, gain: Tensor | None = None, resample: int = 0,
               pixelnorm: bool = False, want_x: bool = True, want_a: bool = True, want_nrm: bool = False):
    B, Hin, Win, C1 = inp.shape
    C2 = 0 if skip is None else skip.shape[3]
    H = Hin // 2 if resample == 1 else (Hin * 2 if resample == 2 else Hin)
    W = Win // 2 if resample == 1 else (Win * 2 if resample == 2 else Win)
    x = torch.empty((B, H, W, C1 + C2), device=inp.device, dtype=BF16) if want_x else None
    a = torch.empty((B, H, W, C1 + C2), device=inp.device, dtype=BF16) if want_a else None
    nrm = torch.empty((B, H, W), device=inp.device, dtype=torch.float32) if (want_nrm and pixelnorm) else None
    _lib.call("tedm_block_prep_forward", inp.data_ptr(), _p(skip), _p(gain), _p(x), _p(a), _p(nrm), B, Hin, Win, C1, C2,
              resample, 1 if pixelnorm else 0, _stream())
    return x, a, nrm


def block_prep_backward(*, g_res: Tensor | None, beta: float, g_a: Tensor | None, x: Tensor | None, nrm: Tensor | None,
                        gain: Tensor | None, d_mean: Tensor | None, g_in: Tensor | None, g_skip: Tensor | None,
                        accumulate_in: bool, accumulate_skip: bool, B: int, Hin: int, Win: int, C1: int, C2: int,
                        resample: int, pixelnorm: bool) -> None:
    _lib.call("tedm_block_prep_backward", _p(g_res), beta, _p(g_a), _p(x), _p(nrm), _p(gain), _p(d_mean), _p(g_in),
              _p(g_skip), 1 if accumulate_in else 0, 1 if accumulate_skip else 0, B, Hin, Win, C1, C2, resample,
              1 if pixelnorm else 0, _stream())


def modsilu_backward(g_h: Tensor, raw: Tensor, mod: Tensor, mod_off: int, d_mod: Tensor, drop_p: float, seed: int) -> Tensor:
    B, H, W, C = g_h.shape
    g_raw = torch.empty_like(g_h)
    _lib.call("tedm_modsilu_backward", g_h.data_ptr(), raw.data_ptr(), mod.data_ptr() + 4 * mod_off,
              d_mod.data_ptr() + 4 * mod_off, g_raw.data_ptr(), B, H * W, C, mod.shape[1], drop_p, seed, _stream())
    return g_raw


def channel_dot(A: Tensor, Bm: Tensor | None, out: Tensor, C: int, a_off: int, scale: float) -> None:
    """out[b,c] += scale * sum_hw A[b,hw,a_off+c] * (Bm[b,hw,c] or 1). `out` must be zeroed by the caller."""
    B, H, W, CA = A.shape
    _lib.call("tedm_channel_dot", A.data_ptr(), _p(Bm), out.data_ptr(), B, H * W, C, CA, a_off, scale, _stream())


def attention_forward(qkv: Tensor, heads: int, need_lse: bool):
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    S = H * W
    qkvn = torch.empty((3, B, heads, S, hd), device=qkv.device, dtype=BF16)
    y = torch.empty((B, H, W, C), device=qkv.device, dtype=BF16)
    lse = torch.empty((B * heads * S,), device=qkv.device, dtype=torch.float32) if need_lse else None
    _lib.call("tedm_attention_forward", qkv.data_ptr(), qkvn.data_ptr(), y.data_ptr(), _p(lse), B, S, heads, hd, _stream())
    return y, qkvn, lse


def attention_backward(qkv: Tensor, qkvn: Tensor, y: Tensor, g_y: Tensor, lse: Tensor, heads: int) -> Tensor:
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    S = H * W
    delta = torch.empty_like(lse)
    g_qkvn = torch.empty_like(qkvn)
    g_qkv = torch.empty_like(qkv)
    _lib.call("tedm_attention_backward", qkv.data_ptr(), qkvn.data_ptr(), y.data_ptr(), g_y.data_ptr(), lse.data_ptr(),
              delta.data_ptr(), g_qkvn.data_ptr(), g_qkv.data_ptr(), B, S, heads, hd, _stream())
    return g_qkv


def sgemm(A: Tensor, B: Tensor, C: Tensor, M: int, N: int, K: int, lda: int, ldb: int, ldc: int, transA: bool,
          transB: bool, alpha: float = 1.0, beta: float = 0.0) -> None:
    _lib.call("tedm_sgemm", A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, lda, ldb, ldc, int(transA), int(transB),
              alpha, beta, _stream())


def heun_step(x0, x1, D, d_prev, x_out, d_out, ts: Tensor, step: int, mode: int) -> None:
    n = x_out.numel()
    _lib.call("tedm_heun_step", _p(x0), _p(x1), _p(D), _p(d_prev), x_out.data_ptr(), _p(d_out), ts.data_ptr(), step, mode,
              n, _stream())
