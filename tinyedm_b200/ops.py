"""Thin torch-tensor wrappers over the C ABI (include/tinyedm_b200.h).

Tensors are only used as owners of device memory; every function launches on torch's current stream.
Activations are NHWC bf16 `(B, H, W, C)`; see DESIGN.md for the layouts. Nothing here computes on the host
and nothing falls back to torch ops: a missing library or an unsupported shape raises RuntimeError.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32

EPI_PLAIN, EPI_MODSILU, EPI_AXPBY, EPI_MODSILU_BWD, EPI_SILU_BWD = 0, 1, 2, 3, 4
RESAMPLE_NONE, RESAMPLE_DOWN, RESAMPLE_UP = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Tensor | None):
    return None if t is None else t.data_ptr()


def check(t: Tensor, dtype, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"tinyedm_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"tinyedm_b200: `{name}` must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"tinyedm_b200: `{name}` must be contiguous")
    return t


def ensure_device(dev: torch.device) -> None:
    if dev.type != "cuda":
        raise RuntimeError("tinyedm_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    _lib.init_device(dev.index if dev.index is not None else torch.cuda.current_device())


def mp_add_coeffs(t: float) -> tuple[float, float]:
    """(weight of a, weight of b) of mp_add(a, b, t) — networks.py:87-88."""
    c = math.sqrt((1 - t) ** 2 + t ** 2)
    return (1 - t) / c, t / c


# ---------------------------------------------------------------------------------------------------------
# weight normalisation
# ---------------------------------------------------------------------------------------------------------
def weight_prep_forward(table: Tensor, n: int, total_groups: int, training: bool) -> None:
    _lib.call("tedm_weight_prep_forward", table.data_ptr(), n, total_groups, 1 if training else 0, _stream())


def weight_prep_backward(table: Tensor, n: int, total_rows: int, max_row_floats: int) -> None:
    _lib.call("tedm_weight_prep_backward", table.data_ptr(), n, total_rows, max_row_floats, _stream())


# ---------------------------------------------------------------------------------------------------------
# MPConv
# ---------------------------------------------------------------------------------------------------------
def conv2d(x: Tensor, w: Tensor, ksize: int, cout: int, *, epi: int = EPI_PLAIN, alpha: float = 1.0, beta: float = 0.0,
           res: Tensor | None = None, raw: Tensor | None = None, mod: Tensor | None = None, mod_off: int = 0,
           mod_stride: int | None = None, drop_p: float = 0.0, seed: int = 0, seed_ptr: Tensor | None = None,
           out: Tensor | None = None, block_n: int = 0, aux: Tensor | None = None, d_mod: Tensor | None = None,
           nrm: Tensor | None = None, accumulate_out: bool = False, col_partial: Tensor | None = None,
           out_bias: Tensor | None = None, out_bias_scale: float = 1.0) -> Tensor:
    """Implicit-GEMM MPConv (forward or data gradient). `w` is the prepared bf16 weight [cout][k*k][cin]."""
    B, H, W, cin = x.shape
    if out is None:
        out = torch.empty((B, H, W, cout), device=x.device, dtype=BF16)
    mod_ptr = None
    if mod is not None:
        mod_ptr = mod.data_ptr() + 4 * mod_off
        if mod_stride is None:
            mod_stride = mod.shape[1]
    d_mod_ptr = None if d_mod is None else d_mod.data_ptr() + 4 * mod_off
    _lib.call("tedm_conv2d_forward", x.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, W, cin, cout, ksize, epi, alpha,
              _p(raw), _p(res), beta, mod_ptr, mod_stride or 0, drop_p, seed, _p(seed_ptr), block_n, _p(aux), d_mod_ptr,
              _p(nrm), 1 if accumulate_out else 0, _p(col_partial), _p(out_bias), out_bias_scale, _stream())
    return out


def conv2d_colsum_slots(B: int, H: int, W: int, cin: int, cout: int, ksize: int, epi: int) -> int:
    """Rows per image of the `col_partial` output of `conv2d` for this launch (0: not available, reduce with channel_dot)."""
    return _lib.call_int("tedm_conv2d_colsum_slots", B, H, W, cin, cout, ksize, epi)


def colsum_mean(col_partial: Tensor, B: int, slots: int, C: int, scale: float) -> Tensor:
    mean = torch.empty((B, C), device=col_partial.device, dtype=F32)
    _lib.call("tedm_colsum_mean", col_partial.data_ptr(), mean.data_ptr(), B, slots, C, scale, _stream())
    return mean


def conv2d_dgrad_split_supported(B: int, H: int, W: int, cin: int, c1: int, c2: int, ksize: int) -> bool:
    return _lib.call_int("tedm_conv2d_dgrad_split_supported", B, H, W, cin, c1, c2, ksize) == 1


def conv2d_dgrad_split(g: Tensor, w: Tensor, ksize: int, *, x: Tensor, res: Tensor, beta: float, gain: Tensor, g_in: Tensor,
                       g_skip: Tensor, d_gx: Tensor, accumulate_in: bool, alpha: float = 1.0, in_bias: Tensor | None = None,
                       in_bias_scale: float = 1.0) -> None:
    """Data gradient of a skip-decoder block's first conv with the concat split, the ScaleLong gain and the reduction for
    d(gain) fused into the epilogue (see tedm_conv2d_dgrad_split). d_gx (B,C2) fp32 must be zeroed by the caller."""
    B, H, W, cin = g.shape
    c1, c2 = g_in.shape[3], g_skip.shape[3]
    _lib.call("tedm_conv2d_dgrad_split", g.data_ptr(), w.data_ptr(), g_in.data_ptr(), g_skip.data_ptr(), B, H, W, cin, c1, c2,
              ksize, alpha, x.data_ptr(), res.data_ptr(), beta, gain.data_ptr(), d_gx.data_ptr(), 1 if accumulate_in else 0,
              _p(in_bias), in_bias_scale, _stream())


def bias_add_bc(g: Tensor, bias: Tensor, scale: float) -> None:
    """g[b,h,w,c] += scale * bias[b,c] in place (bf16 NHWC, fp32 bias)."""
    B, H, W, C = g.shape
    _lib.call("tedm_bias_add_bc", g.data_ptr(), bias.data_ptr(), scale, B, H * W, C, _stream())


def conv2d_wgrad(g: Tensor, x: Tensor, dw: Tensor, ksize: int, *, alpha: float = 1.0, accumulate: bool = False,
                 splits: int = 0) -> None:
    """dw[cout][k*k][cin] (fp32) (+)= alpha * sum_pixels g x."""
    B, H, W, cout = g.shape
    cin = x.shape[3]
    _lib.call("tedm_conv2d_wgrad", g.data_ptr(), x.data_ptr(), dw.data_ptr(), B, H, W, cin, cout, ksize, alpha,
              1 if accumulate else 0, splits, _stream())


# ---------------------------------------------------------------------------------------------------------
# bandwidth-bound block kernels
# ---------------------------------------------------------------------------------------------------------
def resampled_hw(Hin: int, Win: int, resample: int) -> tuple[int, int]:
    if resample == RESAMPLE_DOWN:
        return Hin // 2, Win // 2
    if resample == RESAMPLE_UP:
        return Hin * 2, Win * 2
    return Hin, Win


def block_prep(inp: Tensor, *, skip: Tensor | None = None, gain: Tensor | None = None, resample: int = 0,
               pixelnorm: bool = False, want_x: bool = True, want_a: bool = True, want_nrm: bool = False):
    B, Hin, Win, C1 = inp.shape
    C2 = 0 if skip is None else skip.shape[3]
    H, W = resampled_hw(Hin, Win, resample)
    x = torch.empty((B, H, W, C1 + C2), device=inp.device, dtype=BF16) if want_x else None
    a = torch.empty((B, H, W, C1 + C2), device=inp.device, dtype=BF16) if want_a else None
    nrm = torch.empty((B, H, W), device=inp.device, dtype=F32) if (want_nrm and pixelnorm) else None
    _lib.call("tedm_block_prep_forward", inp.data_ptr(), _p(skip), _p(gain), _p(x), _p(a), _p(nrm), B, Hin, Win, C1, C2,
              resample, 1 if pixelnorm else 0, _stream())
    return x, a, nrm


def block_prep_backward(*, g_res: Tensor | None, beta: float, g_a: Tensor | None, x: Tensor | None, nrm: Tensor | None,
                        gain: Tensor | None, d_mean: Tensor | None, g_in: Tensor, g_skip: Tensor | None,
                        accumulate_in: bool, accumulate_skip: bool, B: int, Hin: int, Win: int, C1: int, C2: int,
                        resample: int, pixelnorm: bool) -> None:
    _lib.call("tedm_block_prep_backward", _p(g_res), beta, _p(g_a), _p(x), _p(nrm), _p(gain), _p(d_mean), _p(g_in),
              _p(g_skip), 1 if accumulate_in else 0, 1 if accumulate_skip else 0, B, Hin, Win, C1, C2, resample,
              1 if pixelnorm else 0, _stream())


def modsilu_backward(g_h: Tensor, raw: Tensor, mod: Tensor, mod_off: int, d_mod: Tensor, drop_p: float, seed: int,
                     seed_ptr: Tensor | None) -> Tensor:
    B, H, W, C = g_h.shape
    g_raw = torch.empty_like(g_h)
    _lib.call("tedm_modsilu_backward", g_h.data_ptr(), raw.data_ptr(), mod.data_ptr() + 4 * mod_off,
              d_mod.data_ptr() + 4 * mod_off, g_raw.data_ptr(), B, H * W, C, mod.shape[1], drop_p, seed, _p(seed_ptr),
              _stream())
    return g_raw


def channel_dot(A: Tensor, Bm: Tensor | None, out: Tensor, C: int, a_off: int, scale: float) -> None:
    """out[b,c] += scale * sum_hw A[b,hw,a_off+c] * (Bm[b,hw,c] or 1). `out` must be zeroed by the caller."""
    B, H, W, CA = A.shape
    _lib.call("tedm_channel_dot", A.data_ptr(), _p(Bm), out.data_ptr(), B, H * W, C, CA, a_off, scale, _stream())


# ---------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------
def attention_forward(qkv: Tensor, heads: int, need_lse: bool):
    """qkv: (B,H,W,3C) with channel = {q,k,v}*C + head*hd + d. Returns (y, lse)."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    S = H * W
    y = torch.empty((B, H, W, C), device=qkv.device, dtype=BF16)
    lse = torch.empty((B * heads * S,), device=qkv.device, dtype=F32) if need_lse else None
    _lib.call("tedm_attention_forward", qkv.data_ptr(), y.data_ptr(), _p(lse), B, S, heads, hd, _stream())
    return y, lse


def attention_backward(qkv: Tensor, y: Tensor, g_y: Tensor, lse: Tensor, heads: int) -> Tensor:
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    S = H * W
    delta = torch.empty_like(lse)
    g_qkv = torch.empty_like(qkv)
    _lib.call("tedm_attention_backward", qkv.data_ptr(), y.data_ptr(), g_y.data_ptr(), lse.data_ptr(), delta.data_ptr(),
              g_qkv.data_ptr(), B, S, heads, hd, _stream())
    return g_qkv


def attention_specialised(S: int, hd: int) -> bool:
    """Shapes served by the kernels that normalise q, k, v themselves (head_dim 64, S in {64, 256}: the CIFAR config);
    every other shape goes through qkv_normalize + the *_normalized pair."""
    return hd == 64 and S in (64, 256)


def qkv_normalize(qkv: Tensor, heads: int):
    """(qn, norms): pixel_norm over head_dim of q, k and v (networks.py:195) and n = eps + rms per row (fp32)."""
    B, H, W, C3 = qkv.shape
    hd = C3 // 3 // heads
    qn = torch.empty_like(qkv)
    norms = torch.empty((B * H * W, 3 * heads), device=qkv.device, dtype=F32)
    _lib.call("tedm_qkv_normalize", qkv.data_ptr(), qn.data_ptr(), norms.data_ptr(), B * H * W, heads, hd, _stream())
    return qn, norms


def attention_forward_normalized(qn: Tensor, heads: int, need_lse: bool):
    B, H, W, C3 = qn.shape
    C = C3 // 3
    S = H * W
    y = torch.empty((B, H, W, C), device=qn.device, dtype=BF16)
    lse = torch.empty((B * heads * S,), device=qn.device, dtype=F32) if need_lse else None
    _lib.call("tedm_attention_forward_normalized", qn.data_ptr(), y.data_ptr(), _p(lse), B, S, heads, C // heads, _stream())
    return y, lse


def attention_backward_normalized(qn: Tensor, norms: Tensor, y: Tensor, g_y: Tensor, lse: Tensor, heads: int) -> Tensor:
    B, H, W, C3 = qn.shape
    C = C3 // 3
    delta = torch.empty_like(lse)
    g_qkv = torch.empty_like(qn)
    _lib.call("tedm_attention_backward_normalized", qn.data_ptr(), norms.data_ptr(), y.data_ptr(), g_y.data_ptr(), lse.data_ptr(),
              delta.data_ptr(), g_qkv.data_ptr(), B, H * W, heads, C // heads, _stream())
    return g_qkv


# ---------------------------------------------------------------------------------------------------------
# small fp32 layers
# ---------------------------------------------------------------------------------------------------------
def sgemm(A: Tensor, B: Tensor, C: Tensor, M: int, N: int, K: int, lda: int, ldb: int, ldc: int, transA: bool,
          transB: bool, alpha: float = 1.0, beta: float = 0.0) -> None:
    """C[M,N] = alpha * op(A) op(B) + beta * C, row-major fp32."""
    _lib.call("tedm_sgemm", A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, lda, ldb, ldc, int(transA), int(transB),
              alpha, beta, _stream())


def embedding_forward(sigma: Tensor, freqs: Tensor, phases: Tensor, w_sigma: Tensor, w_class: Tensor | None,
                      labels: Tensor | None, rows: int, n_classes: int, add_factor: float):
    F_, E = freqs.numel(), w_sigma.shape[0]
    dev = sigma.device
    fourier = torch.empty((rows, F_), device=dev, dtype=F32)
    pre = torch.empty((rows, E), device=dev, dtype=F32)
    emb = torch.empty((rows, E), device=dev, dtype=F32)
    _lib.call("tedm_embedding_forward", sigma.data_ptr(), 0 if sigma.numel() == 1 else 1, freqs.data_ptr(),
              phases.data_ptr(), w_sigma.data_ptr(), _p(w_class), _p(labels), fourier.data_ptr(), pre.data_ptr(),
              emb.data_ptr(), rows, F_, E, n_classes, add_factor, _stream())
    return fourier, pre, emb


def embedding_backward(g_emb: Tensor, pre: Tensor, labels: Tensor | None, g_w_class: Tensor | None, n_classes: int,
                       add_factor: float) -> Tensor:
    B, E = pre.shape
    g_sig = torch.empty_like(pre)
    _lib.call("tedm_embedding_backward", g_emb.data_ptr(), pre.data_ptr(), _p(labels), g_sig.data_ptr(), _p(g_w_class),
              B, E, n_classes, add_factor, _stream())
    return g_sig


def mod_finish_forward(lin: Tensor, gains: Tensor, col_block: Tensor) -> Tensor:
    B, N = lin.shape
    m = torch.empty_like(lin)
    _lib.call("tedm_mod_finish_forward", lin.data_ptr(), gains.data_ptr(), col_block.data_ptr(), m.data_ptr(), B, N,
              _stream())
    return m


def mod_finish_backward(lin: Tensor, dm: Tensor, gains: Tensor, blk_start: Tensor, d_gain: Tensor, n_blocks: int) -> Tensor:
    B, N = lin.shape
    d_lin = torch.empty_like(lin)
    _lib.call("tedm_mod_finish_backward", lin.data_ptr(), dm.data_ptr(), gains.data_ptr(), blk_start.data_ptr(),
              d_lin.data_ptr(), d_gain.data_ptr(), B, N, n_blocks, _stream())
    return d_lin


def scalelong_forward(mean: Tensor, w1: Tensor, w2: Tensor, R: int):
    B, C = mean.shape
    dev = mean.device
    aug = torch.empty((B, C + 1), device=dev, dtype=F32)
    h_pre = torch.empty((B, R), device=dev, dtype=F32)
    h = torch.empty((B, R), device=dev, dtype=F32)
    gain = torch.empty((B, C), device=dev, dtype=F32)
    _lib.call("tedm_scalelong_forward", mean.data_ptr(), w1.data_ptr(), w2.data_ptr(), aug.data_ptr(), h_pre.data_ptr(),
              h.data_ptr(), gain.data_ptr(), B, C, R, _stream())
    return aug, h_pre, h, gain


def scalelong_backward(d_gain: Tensor, gain: Tensor, h_pre: Tensor, w1: Tensor, w2: Tensor, d_gain_times_gain: bool = False):
    B, C = gain.shape
    R = h_pre.shape[1]
    d_pre2 = torch.empty_like(gain)
    d_hpre = torch.empty_like(h_pre)
    d_mean = torch.empty_like(gain)
    _lib.call("tedm_scalelong_backward", d_gain.data_ptr(), gain.data_ptr(), h_pre.data_ptr(), w1.data_ptr(),
              w2.data_ptr(), d_pre2.data_ptr(), d_hpre.data_ptr(), d_mean.data_ptr(), B, C, R,
              1 if d_gain_times_gain else 0, _stream())
    return d_pre2, d_hpre, d_mean


def scalelong_wgrad(d_pre2: Tensor, h: Tensor, d_hpre: Tensor, aug: Tensor, dw2: Tensor, dw1: Tensor) -> None:
    """dw2[C][R] += d_pre2^T h ; dw1[R][C+1] += d_hpre^T aug (both accumulated: zero them first)."""
    B, C = d_pre2.shape
    R = h.shape[1]
    _lib.call("tedm_scalelong_wgrad", d_pre2.data_ptr(), h.data_ptr(), d_hpre.data_ptr(), aug.data_ptr(), dw2.data_ptr(),
              dw1.data_ptr(), B, C, R, _stream())


def uncertainty_forward(fourier: Tensor, w1: Tensor, w2: Tensor, gain: Tensor):
    B, F_ = fourier.shape
    dev = fourier.device
    aug = torch.empty((B, F_ + 1), device=dev, dtype=F32)
    h_pre = torch.empty((B, F_), device=dev, dtype=F32)
    h = torch.empty((B, F_), device=dev, dtype=F32)
    u_raw = torch.empty((B,), device=dev, dtype=F32)
    u = torch.empty((B,), device=dev, dtype=F32)
    _lib.call("tedm_uncertainty_forward", fourier.data_ptr(), w1.data_ptr(), w2.data_ptr(), gain.data_ptr(),
              aug.data_ptr(), h_pre.data_ptr(), h.data_ptr(), u_raw.data_ptr(), u.data_ptr(), B, F_, _stream())
    return aug, h_pre, h, u_raw, u


def uncertainty_backward(g_u: Tensor, gain: Tensor, w2: Tensor, h_pre: Tensor):
    B, F_ = h_pre.shape
    g_uraw = torch.empty((B,), device=h_pre.device, dtype=F32)
    g_hpre = torch.empty_like(h_pre)
    _lib.call("tedm_uncertainty_backward", g_u.data_ptr(), gain.data_ptr(), w2.data_ptr(), h_pre.data_ptr(),
              g_uraw.data_ptr(), g_hpre.data_ptr(), B, F_, _stream())
    return g_uraw, g_hpre


# ---------------------------------------------------------------------------------------------------------
# image-sized kernels
# ---------------------------------------------------------------------------------------------------------
def conv_in_im2col(noisy: Tensor, sigma: Tensor, sigma_data: float) -> Tensor:
    B, Ci, H, W = noisy.shape
    out = torch.empty((B, H, W, 64), device=noisy.device, dtype=BF16)
    _lib.call("tedm_conv_in_im2col", noisy.data_ptr(), sigma.data_ptr(), 0 if sigma.numel() == 1 else 1, sigma_data,
              out.data_ptr(), B, Ci, H, W, _stream())
    return out


def conv_out_forward(x: Tensor, w: Tensor, gain_out: Tensor, noisy: Tensor, sigma: Tensor, sigma_data: float,
                     keep_raw: bool):
    B, H, W, C = x.shape
    Co = noisy.shape[1]
    D = torch.empty_like(noisy)
    f_raw = torch.empty_like(noisy) if keep_raw else None
    _lib.call("tedm_conv_out_forward", x.data_ptr(), w.data_ptr(), gain_out.data_ptr(), noisy.data_ptr(), sigma.data_ptr(),
              0 if sigma.numel() == 1 else 1, sigma_data, _p(f_raw), D.data_ptr(), B, H * W, C, Co, _stream())
    return D, f_raw


def conv_out_backward(g_D: Tensor, f_raw: Tensor, x: Tensor, w: Tensor, gain_out: Tensor, sigma: Tensor,
                      sigma_data: float, g_w: Tensor, g_gain_out: Tensor) -> Tensor:
    """g_w and g_gain_out are accumulated atomically: zero them first."""
    B, H, W, C = x.shape
    Co = g_D.shape[1]
    g_x = torch.empty_like(x)
    _lib.call("tedm_conv_out_backward", g_D.data_ptr(), f_raw.data_ptr(), x.data_ptr(), w.data_ptr(), gain_out.data_ptr(),
              sigma.data_ptr(), 0 if sigma.numel() == 1 else 1, sigma_data, g_x.data_ptr(), g_w.data_ptr(),
              g_gain_out.data_ptr(), B, H * W, C, Co, _stream())
    return g_x


def wmse_forward(D: Tensor, y: Tensor, sigma: Tensor | None, u: Tensor | None, weight: Tensor | None, sigma_data: float,
                 wsum: Tensor | None = None):
    B = D.shape[0]
    n = D.numel() // B
    mse = torch.empty((B,), device=D.device, dtype=F32)
    loss = torch.empty((1,), device=D.device, dtype=F32)
    _lib.call("tedm_wmse_forward", D.data_ptr(), y.data_ptr(), _p(sigma), _p(u), _p(weight), sigma_data, mse.data_ptr(),
              _p(wsum), loss.data_ptr(), B, n, _stream())
    return loss, mse


def wmse_backward(D: Tensor, y: Tensor, sigma: Tensor | None, u: Tensor | None, weight: Tensor | None, mse: Tensor,
                  g_loss: Tensor, sigma_data: float, want_g_u: bool, want_g_weight: bool):
    B = D.shape[0]
    n = D.numel() // B
    g_D = torch.empty_like(D)
    g_u = torch.empty((B,), device=D.device, dtype=F32) if want_g_u else None
    g_w = torch.empty((B,), device=D.device, dtype=F32) if want_g_weight else None
    _lib.call("tedm_wmse_backward", D.data_ptr(), y.data_ptr(), _p(sigma), _p(u), _p(weight), mse.data_ptr(),
              g_loss.data_ptr(), sigma_data, g_D.data_ptr(), _p(g_u), _p(g_w), B, n, _stream())
    return g_D, g_u, g_w


HEUN_EULER, HEUN_CORRECT, HEUN_SCALE = 0, 1, 2


def heun_step(x0, x1, D, d_prev, x_out, d_out, ts: Tensor, step: int, mode: int) -> None:
    n = x_out.numel()
    _lib.call("tedm_heun_step", _p(x0), _p(x1), _p(D), _p(d_prev), x_out.data_ptr(), _p(d_out), ts.data_ptr(), step, mode,
              n, _stream())


def diffuse(clean: Tensor, eps: Tensor, noise: Tensor, P_mean: float, P_std: float):
    B = clean.shape[0]
    n = clean.numel() // B
    noisy = torch.empty_like(clean)
    sigma = torch.empty((B,), device=clean.device, dtype=F32)
    _lib.call("tedm_diffuse", clean.data_ptr(), eps.data_ptr(), noise.data_ptr(), P_mean, P_std, noisy.data_ptr(),
              sigma.data_ptr(), B, n, _stream())
    return noisy, sigma


def diffuse_philox(clean: Tensor, state: Tensor, P_mean: float, P_std: float, sigma_data: float | None):
    """Diffuser.forward with in-kernel Philox draws; with `sigma_data` also conv_in's im2col operand of c_in * noisy.
    Returns (noisy, sigma, xcol or None). `state`: device int64 tensor {seed, step}."""
    B, Ci, H, W = clean.shape
    noisy = torch.empty_like(clean)
    sigma = torch.empty((B,), device=clean.device, dtype=F32)
    xcol = torch.empty((B, H, W, 64), device=clean.device, dtype=BF16) if sigma_data is not None else None
    _lib.call("tedm_diffuse_philox", clean.data_ptr(), state.data_ptr(), P_mean, P_std,
              float(sigma_data or 0.0), noisy.data_ptr(), sigma.data_ptr(), _p(xcol), B, Ci, H, W, _stream())
    return noisy, sigma, xcol


def philox_normal_draws(state: Tensor, B: int, n: int):
    """(eps (B,), noise (B, n)): the draws `diffuse_philox` makes at state = {seed, step}."""
    eps = torch.empty((B,), device=state.device, dtype=F32)
    noise = torch.empty((B, n), device=state.device, dtype=F32)
    _lib.call("tedm_philox_normal_draws", state.data_ptr(), eps.data_ptr(), noise.data_ptr(), B, n, _stream())
    return eps, noise


def to_uint8_images(x: Tensor, mean: Tensor, std: Tensor) -> Tensor:
    """(B,C,H,W) fp32 sampler output -> (B,H,W,C) uint8: clamp(x * std * 2 + mean, 0, 1) * 255 (callbacks.py:152-154)."""
    check(x, F32, "images")
    B, C, H, W = x.shape
    mean = check(mean.to(x.device, F32).reshape(-1).contiguous(), F32, "mean")
    std = check(std.to(x.device, F32).reshape(-1).contiguous(), F32, "std")
    if mean.numel() != C or std.numel() != C:
        raise RuntimeError(f"tinyedm_b200: mean/std must have {C} entries")
    out = torch.empty((B, H, W, C), device=x.device, dtype=torch.uint8)
    _lib.call("tedm_to_uint8_images", x.data_ptr(), mean.data_ptr(), std.data_ptr(), out.data_ptr(), B, C, H * W, _stream())
    return out
