// extern "C" entry points of libtinyedm_b200.so (declared in include/tinyedm_b200.h).
#include "../../include/tinyedm_b200.h"

#include "common.cuh"
#include "kernels.h"

using namespace tedm;

extern "C" {

int tedm_version(void) { return 100; }

const char* tedm_last_error(void) { return last_error(); }

int tedm_init(int device) {
  TEDM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TEDM_CUDA(cudaGetDeviceProperties(&prop, device));
  TEDM_CHECK(prop.major == 10, "tinyedm_b200 requires an sm_100-class GPU (found sm_%d%d, %s)", prop.major, prop.minor,
             prop.name);
  return 0;
}

int tedm_conv2d_forward(const void* x, const void* w, void* out, int B, int H, int W, int Cin, int Cout, int ksize,
                        int epilogue, float alpha, void* raw, const void* res, float beta, const float* mod,
                        int mod_stride, float drop_p, uint64_t seed, const uint64_t* seed_ptr, int block_n,
                        const void* aux, float* d_mod, const float* nrm, int accumulate_out, float* col_partial,
                        const float* out_bias, float out_bias_scale, tedm_stream_t stream) {
  ConvGemmArgs a{};
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.w = static_cast<const __nv_bfloat16*>(w);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize;
  a.epi = epilogue; a.alpha = alpha;
  a.out2 = static_cast<__nv_bfloat16*>(raw);
  a.res = static_cast<const __nv_bfloat16*>(res);
  a.beta = beta;
  a.mod = mod; a.mod_stride = mod_stride; a.drop_p = drop_p; a.seed = seed;
  a.seed_ptr = reinterpret_cast<const unsigned long long*>(seed_ptr);
  a.block_n_override = block_n;
  a.aux = static_cast<const __nv_bfloat16*>(aux);
  a.d_mod = d_mod; a.nrm = nrm; a.accumulate_out = accumulate_out;
  a.col_partial = col_partial;
  a.out_bias = out_bias; a.out_bias_scale = out_bias_scale;
  return conv_gemm_launch(a, static_cast<cudaStream_t>(stream));
}
int tedm_conv2d_colsum_slots(int B, int H, int W, int Cin, int Cout, int ksize, int epilogue) {
  ConvGemmArgs a{};
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize; a.epi = epilogue;
  return conv_colsum_slots(a);
}
int tedm_colsum_mean(const float* col_partial, float* mean, int B, int slots, int C, float scale, tedm_stream_t stream) {
  return colsum_mean(col_partial, mean, B, slots, C, scale, static_cast<cudaStream_t>(stream));
}

static ConvGemmArgs dgrad_split_args(int B, int H, int W, int Cin, int C1, int C2, int ksize) {
  ConvGemmArgs a{};
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = C1 + C2; a.ksize = ksize;
  a.epi = EPI_SILU_BWD;
  a.split_c = C1;
  return a;
}
int tedm_conv2d_dgrad_split_supported(int B, int H, int W, int Cin, int C1, int C2, int ksize) {
  if (C1 <= 0 || C2 <= 0) return 0;
  return conv_split_supported(dgrad_split_args(B, H, W, Cin, C1, C2, ksize)) ? 1 : 0;
}
int tedm_conv2d_dgrad_split(const void* g, const void* w, void* g_in, void* g_skip, int B, int H, int W, int Cin, int C1,
                            int C2, int ksize, float alpha, const void* x, const void* res, float beta, const float* gain,
                            float* d_gx, int accumulate_in, const float* in_bias, float in_bias_scale, tedm_stream_t stream) {
  ConvGemmArgs a = dgrad_split_args(B, H, W, Cin, C1, C2, ksize);
  a.x = static_cast<const __nv_bfloat16*>(g);
  a.w = static_cast<const __nv_bfloat16*>(w);
  a.out = static_cast<__nv_bfloat16*>(g_in);
  a.out2 = static_cast<__nv_bfloat16*>(g_skip);
  a.alpha = alpha;
  a.aux = static_cast<const __nv_bfloat16*>(x);
  a.res = static_cast<const __nv_bfloat16*>(res);
  a.beta = beta;
  a.mod = gain; a.mod_stride = C2; a.d_mod = d_gx;
  a.accumulate_out = accumulate_in;
  a.out_bias = in_bias; a.out_bias_scale = in_bias_scale;
  return conv_gemm_launch(a, static_cast<cudaStream_t>(stream));
}
int tedm_bias_add_bc(void* g, const float* bias, float scale, int B, int HW, int C, tedm_stream_t stream) {
  return bias_add_bc(static_cast<__nv_bfloat16*>(g), bias, scale, B, HW, C, static_cast<cudaStream_t>(stream));
}

int tedm_conv2d_wgrad(const void* g, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, int ksize,
                      float alpha, int accumulate, int splits, tedm_stream_t stream) {
  ConvWgradArgs a{};
  a.g = static_cast<const __nv_bfloat16*>(g);
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.dw = dw;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize;
  a.alpha = alpha; a.accumulate = accumulate; a.splits_override = splits;
  return conv_wgrad_launch(a, static_cast<cudaStream_t>(stream));
}

#define BF(p) static_cast<__nv_bfloat16*>(p)
#define CBF(p) static_cast<const __nv_bfloat16*>(p)
#define ST(s) static_cast<cudaStream_t>(s)

int tedm_weight_prep_forward(const tedm_weight_desc* table, int n, int total_groups, int training, tedm_stream_t stream) {
  return weight_prep_forward(table, n, total_groups, training, ST(stream));
}
int tedm_weight_prep_backward(const tedm_weight_desc* table, int n, int total_rows, int max_row_floats,
                              tedm_stream_t stream) {
  return weight_prep_backward(table, n, total_rows, max_row_floats, ST(stream));
}

int tedm_block_prep_forward(const void* in, const void* skip, const float* gain, void* x_out, void* a_out, float* nrm_out,
                            int B, int Hin, int Win, int C1, int C2, int resample, int pixelnorm, tedm_stream_t stream) {
  PrepArgs a{CBF(in), CBF(skip), gain, BF(x_out), BF(a_out), nrm_out, B, Hin, Win, C1, C2, resample, pixelnorm};
  return block_prep_forward(a, ST(stream));
}
int tedm_block_prep_backward(const void* g_res, float beta, const void* g_a, const void* x, const float* nrm,
                             const float* gain, const float* d_mean, void* g_in, void* g_skip, int accumulate_in,
                             int accumulate_skip, int B, int Hin, int Win, int C1, int C2, int resample, int pixelnorm,
                             tedm_stream_t stream) {
  PrepBwdArgs a{CBF(g_res), beta, CBF(g_a), CBF(x), nrm, gain, d_mean, BF(g_in), BF(g_skip), accumulate_in,
                accumulate_skip, B, Hin, Win, C1, C2, resample, pixelnorm};
  return block_prep_backward(a, ST(stream));
}
int tedm_modsilu_backward(const void* g_h, const void* raw, const float* mod, float* d_mod, void* g_raw, int B, int HW,
                          int C, int mod_stride, float drop_p, uint64_t seed, const uint64_t* seed_ptr,
                          tedm_stream_t stream) {
  ModSiluBwdArgs a{CBF(g_h), CBF(raw), mod, d_mod, BF(g_raw), B, HW, C, mod_stride, drop_p, (uint32_t)seed,
                   (uint32_t)(seed >> 32), reinterpret_cast<const unsigned long long*>(seed_ptr)};
  return modsilu_backward(a, ST(stream));
}
int tedm_channel_dot(const void* A, const void* Bm, float* out, int B, int HW, int C, int CA, int a_off, float scale,
                     tedm_stream_t stream) {
  ChannelDotArgs a{CBF(A), CBF(Bm), out, B, HW, C, CA, a_off, scale};
  return channel_dot(a, ST(stream));
}
int tedm_attention_forward(const void* qkv, void* y, float* lse, int B, int S, int heads, int head_dim,
                           tedm_stream_t stream) {
  return attention_forward(CBF(qkv), BF(y), lse, B, S, heads, head_dim, ST(stream));
}
int tedm_qkv_normalize(const void* qkv, void* qn, float* norms, int64_t rows, int heads, int head_dim, tedm_stream_t stream) {
  return qkv_normalize(CBF(qkv), BF(qn), norms, (long long)rows, heads, head_dim, ST(stream));
}
int tedm_attention_forward_normalized(const void* qn, void* y, float* lse, int B, int S, int heads, int head_dim,
                                      tedm_stream_t stream) {
  return attention_forward_normalized(CBF(qn), BF(y), lse, B, S, heads, head_dim, ST(stream));
}
int tedm_attention_backward_normalized(const void* qn, const float* norms, const void* y, const void* g_y, const float* lse,
                                       float* delta_ws, void* g_qkv, int B, int S, int heads, int head_dim, tedm_stream_t stream) {
  return attention_backward_normalized(CBF(qn), norms, CBF(y), CBF(g_y), lse, delta_ws, BF(g_qkv), B, S, heads, head_dim, ST(stream));
}
int tedm_attention_backward(const void* qkv, const void* y, const void* g_y, const float* lse, float* delta_ws,
                            void* g_qkv, int B, int S, int heads, int head_dim, tedm_stream_t stream) {
  return attention_backward(CBF(qkv), CBF(y), CBF(g_y), lse, delta_ws, BF(g_qkv), B, S, heads, head_dim, ST(stream));
}
int tedm_sgemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA,
               int transB, float alpha, float beta, tedm_stream_t stream) {
  return sgemm(A, B, C, M, N, K, lda, ldb, ldc, transA, transB, alpha, beta, ST(stream));
}
int tedm_embedding_forward(const float* sigma, int sigma_stride, const float* freqs, const float* phases,
                           const float* w_sigma, const float* w_class, const int64_t* labels, float* fourier, float* pre,
                           float* emb, int B, int F, int E, int n_classes, float add_factor, tedm_stream_t stream) {
  EmbeddingArgs a{sigma, sigma_stride, freqs, phases, w_sigma, w_class, reinterpret_cast<const long long*>(labels),
                  fourier, pre, emb, B, F, E, n_classes, add_factor};
  return embedding_forward(a, ST(stream));
}
int tedm_embedding_backward(const float* g_emb, const float* pre, const int64_t* labels, float* g_sig, float* g_w_class,
                            int B, int E, int n_classes, float add_factor, tedm_stream_t stream) {
  EmbeddingBwdArgs a{g_emb, pre, reinterpret_cast<const long long*>(labels), g_sig, g_w_class, B, E, n_classes, add_factor};
  return embedding_backward(a, ST(stream));
}
int tedm_mod_finish_forward(const float* lin, const void* gains, const int32_t* col_block, float* m, int B, int N,
                            tedm_stream_t stream) {
  return mod_finish_forward(lin, static_cast<const float* const*>(gains), col_block, m, B, N, ST(stream));
}
int tedm_mod_finish_backward(const float* lin, const float* dm, const void* gains, const int32_t* blk_start, float* d_lin,
                             float* d_gain, int B, int N, int n_blocks, tedm_stream_t stream) {
  return mod_finish_backward(lin, dm, static_cast<const float* const*>(gains), blk_start, d_lin, d_gain, B, N, n_blocks,
                             ST(stream));
}
int tedm_scalelong_forward(const float* mean, const float* w1, const float* w2, float* aug, float* h_pre, float* h,
                           float* gain, int B, int C, int R, tedm_stream_t stream) {
  ScaleLongArgs a{mean, w1, w2, aug, h_pre, h, gain, B, C, R};
  return scalelong_forward(a, ST(stream));
}
int tedm_scalelong_backward(const float* d_gain, const float* gain, const float* h_pre, const float* w1, const float* w2,
                            float* d_pre2, float* d_hpre, float* d_mean, int B, int C, int R, int d_gain_times_gain,
                            tedm_stream_t stream) {
  ScaleLongBwdArgs a{d_gain, gain, h_pre, w1, w2, d_pre2, d_hpre, d_mean, B, C, R, d_gain_times_gain};
  return scalelong_backward(a, ST(stream));
}
int tedm_to_uint8_images(const float* x, const float* mean, const float* std, void* out, int B, int C, int HW,
                         tedm_stream_t stream) {
  return to_uint8_images(x, mean, std, static_cast<uint8_t*>(out), B, C, HW, ST(stream));
}
int tedm_scalelong_wgrad(const float* d_pre2, const float* h, const float* d_hpre, const float* aug, float* dw2, float* dw1,
                         int B, int C, int R, tedm_stream_t stream) {
  return scalelong_wgrad(d_pre2, h, d_hpre, aug, dw2, dw1, B, C, R, ST(stream));
}
int tedm_uncertainty_forward(const float* fourier, const float* w1, const float* w2, const float* gain, float* aug,
                             float* h_pre, float* h, float* u_raw, float* u, int B, int F, tedm_stream_t stream) {
  UncertaintyArgs a{fourier, w1, w2, gain, aug, h_pre, h, u_raw, u, B, F};
  return uncertainty_forward(a, ST(stream));
}
int tedm_uncertainty_backward(const float* g_u, const float* gain, const float* w2, const float* h_pre, float* g_uraw,
                              float* g_hpre, int B, int F, tedm_stream_t stream) {
  UncertaintyBwdArgs a{g_u, gain, w2, h_pre, g_uraw, g_hpre, B, F};
  return uncertainty_backward(a, ST(stream));
}
int tedm_conv_in_im2col(const float* noisy, const float* sigma, int sigma_stride, float sigma_data, void* out, int B,
                        int Ci, int H, int W, tedm_stream_t stream) {
  return conv_in_im2col(noisy, sigma, sigma_stride, sigma_data, BF(out), B, Ci, H, W, ST(stream));
}
int tedm_conv_out_forward(const void* x, const void* w, const float* gain_out, const float* noisy, const float* sigma,
                          int sigma_stride, float sigma_data, float* f_raw, float* D, int B, int HW, int C, int Co,
                          tedm_stream_t stream) {
  ConvOutArgs a{CBF(x), CBF(w), gain_out, noisy, sigma, sigma_stride, sigma_data, f_raw, D, B, HW, C, Co};
  return conv_out_forward(a, ST(stream));
}
int tedm_conv_out_backward(const float* g_D, const float* f_raw, const void* x, const void* w, const float* gain_out,
                           const float* sigma, int sigma_stride, float sigma_data, void* g_x, float* g_w,
                           float* g_gain_out, int B, int HW, int C, int Co, tedm_stream_t stream) {
  ConvOutBwdArgs a{g_D, f_raw, CBF(x), CBF(w), gain_out, sigma, sigma_stride, sigma_data, BF(g_x), g_w, g_gain_out,
                   B, HW, C, Co};
  return conv_out_backward(a, ST(stream));
}
int tedm_wmse_forward(const float* D, const float* y, const float* sigma, const float* u, const float* weight,
                      float sigma_data, float* mse, float* wsum, float* loss, int B, int n, tedm_stream_t stream) {
  return wmse_forward(D, y, sigma, u, weight, sigma_data, mse, wsum, loss, B, n, ST(stream));
}
int tedm_wmse_backward(const float* D, const float* y, const float* sigma, const float* u, const float* weight,
                       const float* mse, const float* g_loss, float sigma_data, float* g_D, float* g_u, float* g_weight,
                       int B, int n, tedm_stream_t stream) {
  return wmse_backward(D, y, sigma, u, weight, mse, g_loss, sigma_data, g_D, g_u, g_weight, B, n, ST(stream));
}
int tedm_heun_step(const float* x0, const float* x1, const float* D, const float* d_prev, float* x_out, float* d_out,
                   const float* ts, int step, int mode, int64_t n, tedm_stream_t stream) {
  return heun_step(x0, x1, D, d_prev, x_out, d_out, ts, step, mode, (long long)n, ST(stream));
}
int tedm_diffuse(const float* clean, const float* eps, const float* noise, float P_mean, float P_std, float* noisy,
                 float* sigma, int B, int n, tedm_stream_t stream) {
  return diffuse(clean, eps, noise, P_mean, P_std, noisy, sigma, B, n, ST(stream));
}

int tedm_diffuse_philox(const float* clean, const int64_t* state, float P_mean, float P_std, float sigma_data,
                        float* noisy, float* sigma, void* xcol, int B, int Ci, int H, int W, tedm_stream_t stream) {
  return diffuse_philox(clean, reinterpret_cast<const long long*>(state), P_mean, P_std, sigma_data, noisy, sigma,
                        BF(xcol), B, Ci, H, W, ST(stream));
}
int tedm_philox_normal_draws(const int64_t* state, float* eps, float* noise, int B, int64_t n, tedm_stream_t stream) {
  return philox_normal_draws(reinterpret_cast<const long long*>(state), eps, noise, B, (long long)n, ST(stream));
}

int tedm_adam_chunk_elems(void) { return adam_chunk_elems(); }
int tedm_adam_ema_step(const tedm_adam_desc* table, const int32_t* chunks, int n_chunks, float lr, float step,
                       const float* hyper, float beta1, float beta2, float eps, float gamma, tedm_stream_t stream) {
  return adam_ema_step(table, chunks, n_chunks, lr, step, hyper, beta1, beta2, eps, gamma, ST(stream));
}

}  // extern "C"
