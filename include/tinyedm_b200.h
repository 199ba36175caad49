/* tinyedm_b200 — C ABI of the B200-native EDM2 denoiser hot path (libtinyedm_b200.so).
 *
 * The reference (YichengDWu/tinyedm) is pure Python: its "FFI" for this path is the set of torch
 * functional calls inside src/tinyedm/networks.py, solvers.py, edm.py and metric.py. Each entry point
 * below names the reference lines it replaces. All pointers are raw DEVICE pointers owned by the
 * caller; kernels never allocate; every call is asynchronous on `stream` and CUDA-graph capturable.
 * Return value: 0 on success, non-zero on failure (message via tedm_last_error(), thread local).
 *
 * Internal activation layout: NHWC bf16, (B,H,W,C) with C % 64 == 0 for tensor-core convolutions.
 * Prepared ("normalised") weights: bf16 [Cout][kh][kw][Cin]; weight gradients: fp32, same layout.
 */
#ifndef TINYEDM_B200_H_
#define TINYEDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tedm_stream_t; /* cudaStream_t */

/* ---- runtime ---- */
int tedm_version(void);
const char* tedm_last_error(void);
/* Binds the calling thread to `device` and checks it is an sm_100 part. */
int tedm_init(int device);

/* ---- MPConv: F.conv2d(x, w_hat, padding="same")  (networks.py:22-43, :37) ---- */
/* epilogue: 0 plain (out = alpha*conv)
 *           1 modulation + mp_silu + dropout: out = drop(mp_silu(conv * mod[b,c]))   (networks.py:255-260, :319-324)
 *             raw (optional) receives the un-modulated conv output for backward
 *           2 mp_add: out = ((1-t)*res + t*conv) / sqrt((1-t)^2+t^2)                  (networks.py:87-88, :263, :327) */
int tedm_conv2d_forward(const void* x, const void* w, void* out, int B, int H, int W, int Cin, int Cout, int ksize,
                        int epilogue, float alpha, void* raw, const void* res, float t, const float* mod,
                        int mod_stride, float drop_p, uint64_t seed, int block_n, tedm_stream_t stream);
/* dL/dw_hat of the convolution above: dw[co][tap][ci] (=|+=) alpha * sum_p g[p,co] * x[p+tap,ci]  (autograd of :37) */
int tedm_conv2d_wgrad(const void* g, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, int ksize,
                      float alpha, int accumulate, int splits, tedm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TINYEDM_B200_H_ */
