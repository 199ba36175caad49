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

/* ---- forced weight normalisation, all weight tensors in one launch (networks.py:17-19, :32-36, :55-59) ---- */
typedef struct tedm_weight_desc {
  void* w;           /* fp32 parameter, [rows][cin*taps] (OIHW flattened)                      */
  void* grad;        /* fp32 dL/dw, same layout (written by tedm_weight_prep_backward)         */
  const void* g_hat; /* fp32 dL/dw_hat, [rows][kpad] with k = tap*cin + ci                     */
  void* out_fwd;     /* bf16 w_hat [rows][kpad], k = tap*cin + ci (zero padded)   or NULL      */
  void* out_dgrad;   /* bf16 w_hat [cin][taps flipped][rows]                      or NULL      */
  void* out_f32;     /* fp32 w_hat [rows][cin*taps]                               or NULL      */
  void* stats;       /* fp32 [rows][4] = {1/(eps*sqrt(n)+||w||), ||w||, training rescale, 0}; required, 16-byte aligned */
  int32_t rows, cin, taps, kpad, row_start;
  int32_t qkv_head_dim; /* != 0: qkv conv of CosineAttention — prepared rows are permuted from the reference's
                           head*3*hd + d*3 + {q,k,v} (networks.py:194) to {q,k,v}*C + head*hd + d, so the conv emits
                           de-interleaved q | k | v planes; out_fwd / out_dgrad / g_hat use the permuted row index */
  int32_t group_start; /* index of this tensor's first 16-row group: sum over earlier descriptors of ceil(rows/16) */
  int32_t reserved;
} tedm_weight_desc;
/* table: DEVICE array of n descriptors ordered by row_start / group_start; total_groups = sum(ceil(rows/16)).
 * training != 0 additionally rewrites every parameter in place: w <- normalize(w)  (networks.py:32-34). */
int tedm_weight_prep_forward(const tedm_weight_desc* table, int n, int total_groups, int training, tedm_stream_t stream);
/* dL/dw = g/s - w (w.g)/(s^2 ||w||) for every descriptor with g_hat and grad set (autograd of :35-36).
 * total_rows = sum(rows); max_row_floats = max over descriptors of taps*(cin+4) (shared-memory staging of one row). */
int tedm_weight_prep_backward(const tedm_weight_desc* table, int n, int total_rows, int max_row_floats,
                              tedm_stream_t stream);

/* ---- MPConv: F.conv2d(x, w_hat, padding="same")  (networks.py:22-43, :37) ---- */
/* The same entry computes the data gradient: pass g as x and the out_dgrad weight layout (Cin/Cout swapped).
 * epilogue: 0 plain (out = alpha*conv)
 *           1 modulation + mp_silu + dropout: out = drop(mp_silu(alpha*conv * mod[b,c]))  (networks.py:255-260, :319-324)
 *             raw (optional) receives the un-modulated conv output for backward; the dropout mask is a counter-based
 *             Philox stream keyed by seed (+ *seed_ptr, an optional DEVICE step counter so that a captured CUDA graph
 *             draws a fresh mask on every replay)
 *           2 axpby: out = alpha*conv + beta*res; mp_add(x, conv, t) is alpha = t/c, beta = (1-t)/c,
 *             c = sqrt((1-t)^2+t^2)                                                     (networks.py:87-88, :263, :327).
 *             With nrm (eps + rms per pixel): res is the UN-normalised block input and enters as res / nrm[pixel],
 *             i.e. the pixel-normalised residual of an encoder block (networks.py:249, :263) without storing it
 *         Backward epilogues (the call is then the DATA GRADIENT of a conv, `conv` = alpha * dgrad):
 *           3 adjoint of epilogue 1 fused into the dgrad of conv_3x3_2: conv = dL/dh; out = dL/draw;
 *             d_mod[b,c] += sum_pixels (needs aux = raw, mod, d_mod zeroed by the caller, the same seed/seed_ptr)
 *           4 adjoint of a = mp_silu(x) fused into the dgrad of conv_3x3_1: out = conv*mp_silu'(aux=x) + beta*res,
 *             then, when nrm (eps + rms per pixel, from tedm_block_prep_forward) is given, the pixel_norm adjoint
 *             g/n - x*sum_c(g*x)/((n-eps)*C) (needs Cout <= 256), then out += previous out when accumulate_out,
 *             then out[b,p,c] += out_bias_scale * out_bias[b,c] when out_bias (fp32 (B,Cout)) is given: a pending
 *             per-(image, channel) share of the same gradient, e.g. ScaleLong's mean gradient (networks.py:112)  */
int tedm_conv2d_forward(const void* x, const void* w, void* out, int B, int H, int W, int Cin, int Cout, int ksize,
                        int epilogue, float alpha, void* raw, const void* res, float beta, const float* mod,
                        int mod_stride, float drop_p, uint64_t seed, const uint64_t* seed_ptr, int block_n,
                        const void* aux, float* d_mod, const float* nrm, int accumulate_out, float* col_partial,
                        const float* out_bias, float out_bias_scale, tedm_stream_t stream);
/* Epilogues 0 and 2 can also emit per-(image, channel) partial sums of the stored output, so that ScaleLong's spatial
 * mean of a skip tensor (networks.py:112) needs no pass over the tensor: col_partial is fp32 [B * slots][Cout] with
 * slots = tedm_conv2d_colsum_slots(...) rows per image (0: not available for this launch - pass NULL and reduce the tensor
 * with tedm_channel_dot instead). Every entry is written exactly once, without atomics (bit-reproducible).
 * tedm_colsum_mean: mean[b,c] = scale * sum_s col_partial[b*slots + s, c]. */
int tedm_conv2d_colsum_slots(int B, int H, int W, int Cin, int Cout, int ksize, int epilogue);
int tedm_colsum_mean(const float* col_partial, float* mean, int B, int slots, int C, float scale, tedm_stream_t stream);
/* Data gradient of conv_3x3_1 of a decoder block WITH a skip connection, whose input is cat(in, skip*gain)
 * (networks.py:309-316 and autograd): epilogue 4 over the C1+C2 concatenated channels, g_cat = alpha*dgrad *
 * mp_silu'(x) + beta*res, split in the epilogue instead of materialising g_cat:
 *   g_in  (B,H,W,C1)  (= | += when accumulate_in)  g_cat[..., :C1]
 *   g_skip(B,H,W,C2)   =  g_cat[..., C1:] * gain[b,:]          (the ScaleLong mean-gradient share is added later with
 *                                                                tedm_bias_add_bc once d gain is known)
 *   d_gx  (B,C2)      +=  sum_pixels g_cat[..., C1:] * x[..., C1:]   = d gain * gain (x holds skip*gain); zero it first
 * g: gradient w.r.t. the conv output (B,H,W,Cin); w: the out_dgrad weight layout [C1+C2][k*k][Cin]; x, res: (B,H,W,C1+C2).
 * in_bias (fp32 (B,C1), may be NULL): g_in[b,p,c] += in_bias_scale * in_bias[b,c], as out_bias of tedm_conv2d_forward.
 * Only the CTA-pair kernel implements it: ask tedm_conv2d_dgrad_split_supported (returns 1 / 0) first. */
int tedm_conv2d_dgrad_split_supported(int B, int H, int W, int Cin, int C1, int C2, int ksize);
int tedm_conv2d_dgrad_split(const void* g, const void* w, void* g_in, void* g_skip, int B, int H, int W, int Cin, int C1,
                            int C2, int ksize, float alpha, const void* x, const void* res, float beta, const float* gain,
                            float* d_gx, int accumulate_in, const float* in_bias, float in_bias_scale, tedm_stream_t stream);
/* g[b,p,c] += scale * bias[b,c]  (bf16 NHWC tensor, fp32 per-(image, channel) bias): the gradient of ScaleLong's
 * spatial mean (networks.py:112) with scale = 1/HW */
int tedm_bias_add_bc(void* g, const float* bias, float scale, int B, int HW, int C, tedm_stream_t stream);
/* dL/dw_hat of the convolution above: dw[co][tap][ci] (=|+=) alpha * sum_p g[p,co] * x[p+tap,ci]  (autograd of :37) */
int tedm_conv2d_wgrad(const void* g, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, int ksize,
                      float alpha, int accumulate, int splits, tedm_stream_t stream);

/* ---- bandwidth-bound block kernels (NHWC bf16) ---- */
/* resample (0 none / 1 avg-pool 2x2 / 2 nearest-exact x2), skip concat with ScaleLong gain, pixel_norm, mp_silu in one
 * pass: networks.py:9-14, :67-88, :246-252 (encoder), :306-316 (decoder). Any of x_out / a_out / nrm_out may be NULL. */
int tedm_block_prep_forward(const void* in, const void* skip, const float* gain, void* x_out, void* a_out, float* nrm_out,
                            int B, int Hin, int Win, int C1, int C2, int resample, int pixelnorm, tedm_stream_t stream);
/* adjoint of the above: g_x = beta*g_res + g_a*mp_silu'(x), pixel_norm / resample / concat adjoints. */
int tedm_block_prep_backward(const void* g_res, float beta, const void* g_a, const void* x, const float* nrm,
                             const float* gain, const float* d_mean, void* g_in, void* g_skip, int accumulate_in,
                             int accumulate_skip, int B, int Hin, int Win, int C1, int C2, int resample, int pixelnorm,
                             tedm_stream_t stream);
/* backward of dropout(mp_silu(raw * mod[b,c])) (networks.py:255-261): g_raw, and d_mod[b,c] += sum_hw (zero d_mod first). */
int tedm_modsilu_backward(const void* g_h, const void* raw, const float* mod, float* d_mod, void* g_raw, int B, int HW,
                          int C, int mod_stride, float drop_p, uint64_t seed, const uint64_t* seed_ptr,
                          tedm_stream_t stream);
/* out[b,c] += scale * sum_p A[b,p,a_off+c] * (Bm ? Bm[b,p,c] : 1)   (ScaleLong mean networks.py:115 and its adjoint) */
int tedm_channel_dot(const void* A, const void* Bm, float* out, int B, int HW, int C, int CA, int a_off, float scale,
                     tedm_stream_t stream);

/* ---- cosine attention core (networks.py:194-202): pixel_norm over head_dim of q, k and v, then
 *      softmax(q k^T / sqrt(hd)) v, flash-style (no S x S matrix in memory) ----
 * qkv: (B,S,3C) bf16 with channel = {q,k,v}*C + head*hd + d (the layout the qkv conv emits when its weight rows are
 * permuted by the weight bank, see tedm_weight_desc.qkv_head_dim); y: (B,S,C), channel = head*hd + d (networks.py:202);
 * lse: (B*heads*S) fp32 log-sum-exp kept for backward (may be NULL in inference). The normalisation and its
 * backward are fused into the kernels: no normalised copy of q,k,v exists in HBM. delta_ws: (B*heads*S) fp32 scratch
 * (rowsum(dO o O); the fused S = 256, head_dim 64 backward keeps it in shared memory and leaves the buffer untouched).
 * The backward uses no atomics: results are bit-reproducible. */
int tedm_attention_forward(const void* qkv, void* y, float* lse, int B, int S, int heads, int head_dim,
                           tedm_stream_t stream);
int tedm_attention_backward(const void* qkv, const void* y, const void* g_y, const float* lse, float* delta_ws,
                            void* g_qkv, int B, int S, int heads, int head_dim, tedm_stream_t stream);

/* The same attention core split in two for every other (head_dim, S) of the reference's configs — MNIST (64, 196) and
 * (128, 49), ImageNet-latent (144, 256) and (192, 64); any head_dim % 16 == 0 up to 192 and S <= 256 — on tcgen05/TMEM:
 *   tedm_qkv_normalize: qn = pixel_norm over head_dim of every q, k and v row (networks.py:195; bf16 like the reference's
 *     cast) and norms[(row*3 + plane)*heads + head] = eps + rms of the raw row (fp32), rows = B*S;
 *   tedm_attention_forward_normalized: y, lse from qn (networks.py:201-202);
 *   tedm_attention_backward_normalized: d qkv (w.r.t. the RAW qkv: the pixel-norm adjoint runs in the kernels' epilogue)
 *     from qn, norms, y, d y, lse; delta_ws as above. No atomics: bit-reproducible. */
int tedm_qkv_normalize(const void* qkv, void* qn, float* norms, int64_t rows, int heads, int head_dim, tedm_stream_t stream);
int tedm_attention_forward_normalized(const void* qn, void* y, float* lse, int B, int S, int heads, int head_dim,
                                      tedm_stream_t stream);
int tedm_attention_backward_normalized(const void* qn, const float* norms, const void* y, const void* g_y, const float* lse,
                                       float* delta_ws, void* g_qkv, int B, int S, int heads, int head_dim, tedm_stream_t stream);

/* ---- small fp32 layers ---- */
/* C[M,N] = alpha*op(A)*op(B) + beta*C, row-major fp32 (F.linear of the autocast-off islands, networks.py:46-64) */
int tedm_sgemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA,
               int transB, float alpha, float beta, tedm_stream_t stream);
/* Embedding.forward (networks.py:163-178): sigma (B,) or 0-d (sigma_stride 0), labels int64 (B,) or NULL */
int tedm_embedding_forward(const float* sigma, int sigma_stride, const float* freqs, const float* phases,
                           const float* w_sigma, const float* w_class, const int64_t* labels, float* fourier, float* pre,
                           float* emb, int B, int F, int E, int n_classes, float add_factor, tedm_stream_t stream);
int tedm_embedding_backward(const float* g_emb, const float* pre, const int64_t* labels, float* g_sig, float* g_w_class,
                            int B, int E, int n_classes, float add_factor, tedm_stream_t stream);
/* m[b,col] = lin[b,col]*gain[block(col)] + 1 for all blocks (networks.py:255-258): gains = device array of pointers */
int tedm_mod_finish_forward(const float* lin, const void* gains, const int32_t* col_block, float* m, int B, int N,
                            tedm_stream_t stream);
/* adjoint: d_lin = dm*gain; d_gain[block] += sum dm*lin (accumulated atomically: zero d_gain first) */
int tedm_mod_finish_backward(const float* lin, const float* dm, const void* gains, const int32_t* blk_start, float* d_lin,
                             float* d_gain, int B, int N, int n_blocks, tedm_stream_t stream);
/* ScaleLong (networks.py:106-118) on the spatial mean: gain = sigmoid(W2 mp_silu(W1 [mean,1])) */
int tedm_scalelong_forward(const float* mean, const float* w1, const float* w2, float* aug, float* h_pre, float* h,
                           float* gain, int B, int C, int R, tedm_stream_t stream);
int tedm_scalelong_backward(const float* d_gain, const float* gain, const float* h_pre, const float* w1, const float* w2,
                            float* d_pre2, float* d_hpre, float* d_mean, int B, int C, int R, int d_gain_times_gain,
                            tedm_stream_t stream);
/* sample post-processing (callbacks.py:152-154, datamodule.denormalize): out[b,h,w,c] = uint8(clamp(x[b,c,h,w] * std[c] * 2
 * + mean[c], 0, 1) * 255), x fp32 NCHW (the sampler's output), out uint8 NHWC (what PIL / the PNG writer consumes) */
int tedm_to_uint8_images(const float* x, const float* mean, const float* std, void* out, int B, int C, int HW,
                         tedm_stream_t stream);
/* dL/dw_hat of both ScaleLong layers, ACCUMULATED into dw2 [C][R] and dw1 [R][C+1] (zero them first):
 * dw2 += d_pre2^T h,  dw1 += d_hpre^T aug  (autograd of networks.py:112-115) */
int tedm_scalelong_wgrad(const float* d_pre2, const float* h, const float* d_hpre, const float* aug, float* dw2, float* dw1,
                         int B, int C, int R, tedm_stream_t stream);
/* UncertaintyNet (networks.py:91-103) */
int tedm_uncertainty_forward(const float* fourier, const float* w1, const float* w2, const float* gain, float* aug,
                             float* h_pre, float* h, float* u_raw, float* u, int B, int F, tedm_stream_t stream);
int tedm_uncertainty_backward(const float* g_u, const float* gain, const float* w2, const float* h_pre, float* g_uraw,
                              float* g_hpre, int B, int F, tedm_stream_t stream);

/* ---- image-sized kernels ---- */
/* c_in*x, ones channel, 3x3 patch gather -> (B,H,W,64) bf16 (networks.py:578-587) */
int tedm_conv_in_im2col(const float* noisy, const float* sigma, int sigma_stride, float sigma_data, void* out, int B,
                        int Ci, int H, int W, tedm_stream_t stream);
/* D = conv_out(x)*gain_out*c_out + noisy*c_skip (networks.py:602-603); f_raw (optional) keeps conv_out(x) */
int tedm_conv_out_forward(const void* x, const void* w, const float* gain_out, const float* noisy, const float* sigma,
                          int sigma_stride, float sigma_data, float* f_raw, float* D, int B, int HW, int C, int Co,
                          tedm_stream_t stream);
int tedm_conv_out_backward(const float* g_D, const float* f_raw, const void* x, const void* w, const float* gain_out,
                           const float* sigma, int sigma_stride, float sigma_data, void* g_x, float* g_w,
                           float* g_gain_out, int B, int HW, int C, int Co, tedm_stream_t stream);
/* loss = sum_b w_b * mean_chw((D-y)^2) / B [+ mean(u)]  (edm.py:212-219, metric.py:8-18), one fused reduction.
 * w_b = weight[b] when `weight` is given (the WeightedMeanSquaredError(weight, preds, target) surface), otherwise
 * lambda(sigma_b) [* exp(-u_b) when u is given]. mse: (B,) per-sample mean squared error (kept for backward);
 * wsum (optional): += sum_b w_b*mse_b, the metric's running state (metric.py:44). */
int tedm_wmse_forward(const float* D, const float* y, const float* sigma, const float* u, const float* weight,
                      float sigma_data, float* mse, float* wsum, float* loss, int B, int n, tedm_stream_t stream);
int tedm_wmse_backward(const float* D, const float* y, const float* sigma, const float* u, const float* weight,
                       const float* mse, const float* g_loss, float sigma_data, float* g_D, float* g_u, float* g_weight,
                       int B, int n, tedm_stream_t stream);
/* Heun stages (solvers.py:45-57): mode 0 Euler predictor, 1 trapezoid corrector, 2 x*t_0; ts = device schedule */
int tedm_heun_step(const float* x0, const float* x1, const float* D, const float* d_prev, float* x_out, float* d_out,
                   const float* ts, int step, int mode, int64_t n, tedm_stream_t stream);
/* Diffuser.forward (edm.py:84-93) with the two normal draws supplied by the caller */
int tedm_diffuse(const float* clean, const float* eps, const float* noise, float P_mean, float P_std, float* noisy,
                 float* sigma, int B, int n, tedm_stream_t stream);
/* Diffuser.forward (edm.py:84-93) with BOTH normal draws generated in the kernel, fused with Denoiser.forward's input
 * block (networks.py:578-587) — SURVEY.md §8f row N2. Counter-based Philox4x32-10: the draw for element e is a pure
 * function of (seed, step, e); state = DEVICE int64[2] = {seed, step}, the caller advances step per call (both are read
 * on the device, so replays of a captured graph draw fresh noise and see a re-seed).
 * Writes noisy (B,Ci,H,W) fp32, sigma (B,) and, when xcol != NULL, conv_in's im2col operand of c_in*noisy with the ones
 * channel, (B,H,W,64) bf16 — bit-identical to tedm_conv_in_im2col(noisy, sigma) (Ci <= 6). */
int tedm_diffuse_philox(const float* clean, const int64_t* state, float P_mean, float P_std, float sigma_data,
                        float* noisy, float* sigma, void* xcol, int B, int Ci, int H, int W, tedm_stream_t stream);
/* The draws alone — eps (B,), noise (B,n) — exactly as tedm_diffuse_philox generates them at state = {seed, step}: lets a
 * caller (or a test) reproduce edm.py:84-93 from explicit draws with tedm_diffuse. */
int tedm_philox_normal_draws(const int64_t* state, float* eps, float* noise, int B, int64_t n, tedm_stream_t stream);

/* ---- optimiser step ("next" row N1): fused multi-tensor Adam + power-function EMA ----
 * Replaces optim.Adam(fused=True) (edm.py:250-253) + EMAOptimizer.update (ema.py:137-140, :273) with ONE launch.
 * table: DEVICE array of descriptors; chunks: DEVICE array of n_chunks (tensor index, chunk index) int32 pairs, one CTA
 * per chunk of tedm_adam_chunk_elems() elements. step counts from 1 (Adam bias correction); the EMA decay is
 * (1 - 1/step)^(gamma+1); gamma < 0 disables the EMA. hyper (optional, DEVICE {lr, step}) overrides lr/step so a
 * captured CUDA graph can be replayed with new values. */
typedef struct tedm_adam_desc {
  void* p;       /* fp32 parameter            */
  const void* g; /* fp32 gradient             */
  void* m;       /* fp32 first moment         */
  void* v;       /* fp32 second moment        */
  void* ema;     /* fp32 EMA copy or NULL     */
  int64_t n;     /* elements                  */
} tedm_adam_desc;
int tedm_adam_chunk_elems(void);
int tedm_adam_ema_step(const tedm_adam_desc* table, const int32_t* chunks, int n_chunks, float lr, float step,
                       const float* hyper, float beta1, float beta2, float eps, float gamma, tedm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TINYEDM_B200_H_ */
