// Internal (C++) launch interfaces shared between the .cu translation units and api.cu.
// The public C ABI is include/tinyedm_b200.h.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tinyedm_b200.h"

namespace tedm {

typedef tedm_weight_desc WeightDesc;
int weight_prep_forward(const WeightDesc* table_dev, int n_tensors, int total_groups, int training, cudaStream_t stream);
int weight_prep_backward(const WeightDesc* table_dev, int n_tensors, int total_rows, int max_row_floats, cudaStream_t stream);

enum ConvEpilogue : int {
  EPI_PLAIN = 0,    // out = alpha * acc
  EPI_MODSILU = 1,  // out = dropout(mp_silu(acc * mod[b, c])); optional raw copy of acc in out2
  EPI_AXPBY = 2,    // out = alpha * acc + beta * res   (mp_add, networks.py:87-88: alpha = t/c, beta = (1-t)/c)
  EPI_MODSILU_BWD = 3,  // acc = dL/dh of h = drop(mp_silu(raw*mod)): out = dL/draw, d_mod[b,c] += sum_pix (autograd of :255-261)
  EPI_SILU_BWD = 4,     // acc = dL/da of a = mp_silu(x): out = alpha*acc*mp_silu'(x) + beta*res, optional pixel-norm adjoint
                        // (given nrm) and accumulation into out (autograd of :249-252, :263)
};

struct ConvGemmArgs {
  const __nv_bfloat16* x;  // (B,H,W,Cin) NHWC
  const __nv_bfloat16* w;  // [Cout][ksize*ksize][Cin]
  __nv_bfloat16* out;      // (B,H,W,Cout)
  int B, H, W, Cin, Cout, ksize;
  int epi;
  float alpha;
  __nv_bfloat16* out2;       // MODSILU: raw conv output (may be null)
  const __nv_bfloat16* res;  // AXPBY: residual (B,H,W,Cout)
  float beta;
  const float* mod;          // MODSILU: (B, mod_stride) fp32, column offset already applied
  int mod_stride;
  float drop_p;
  uint64_t seed;
  const unsigned long long* seed_ptr;  // optional device step counter mixed into the seed
  int block_n_override;      // 0 = heuristic
  const __nv_bfloat16* aux;  // MODSILU_BWD: raw conv output; SILU_BWD: x          (B,H,W,Cout)
  float* d_mod;              // MODSILU_BWD: (B, mod_stride) fp32, column offset applied, accumulated atomically
  const float* nrm;          // SILU_BWD: (B*H*W) eps + rms of the pixel norm whose adjoint is fused, or null.
                             // AXPBY: when given, `res` is the UN-normalised tensor and enters as res / nrm[pixel]
                             // (the pixel-normalised residual of an encoder block without ever storing it)
  int accumulate_out;        // SILU_BWD: out += result
  // SILU_BWD, CTA-pair kernel only: split the Cout = split_c + C2 output channels of a decoder block's concatenated
  // input gradient in the epilogue. Channels < split_c go to `out` ((B,H,W,split_c), accumulate_out applies to it);
  // channels >= split_c are multiplied by gain = mod[b, c - split_c] and go to `out2` ((B,H,W,C2)), and
  // d_mod[b, c - split_c] += sum_pixels g * x (= d gain * gain, because x holds skip * gain). 0 = off.
  int split_c;
  // PLAIN / AXPBY, CTA-pair kernel only: per-warp column sums of the (bf16-rounded) output, written without atomics to
  // col_partial[(m_tile * 4 + warp quarter)][Cout] (fp32); conv_colsum_slots() consecutive rows belong to one image, so
  // ScaleLong's spatial mean of a skip tensor (networks.py:112) needs no extra pass over it. null = off.
  float* col_partial;
  // SILU_BWD: out[b,p,c] += out_bias_scale * out_bias[b,c] (fp32 (B, channels of `out`)): a pending per-(image, channel)
  // gradient share - ScaleLong's mean gradient - folded into the kernel that accumulates into the tensor anyway. null = off.
  const float* out_bias;
  float out_bias_scale;
};

// Device-side parameter block of the implicit-GEMM kernel.
struct ConvGemmParams {
  int B, H, W, Cin, Cout, taps;
  int RH, NB, tiles_h, m_tiles, n_tiles, block_n, k_blocks;
  int epi;
  float alpha;
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
  const __nv_bfloat16* res;
  float beta;
  const float* mod;
  int mod_stride;
  float drop_p;
  uint32_t seed_lo, seed_hi;
  const unsigned long long* seed_ptr;
  const __nv_bfloat16* aux;
  float* d_mod;
  const float* nrm;
  int accumulate_out;
  // conv_pair.cu only: work items [0, split_from) are whole pair tiles, items [split_from, work_items) are HALF-N
  // tiles (two consecutive items = the two 128-channel halves of one pair tile); split_from == work_items: no split
  int split_from, work_items;
  int split_c;   // see ConvGemmArgs::split_c
  float* col_partial;   // see ConvGemmArgs::col_partial
  const float* out_bias;   // see ConvGemmArgs::out_bias
  float out_bias_scale;
};

int conv_tile_geometry(int H, int W, int* RH, int* NB);
int conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t stream);
bool conv_split_supported(const ConvGemmArgs& a);
// rows of col_partial per image if this launch takes the CTA-pair kernel and its tile geometry allows the sums, else 0
int conv_colsum_slots(const ConvGemmArgs& a);
int colsum_mean(const float* partial, float* mean, int B, int slots, int C, float scale, cudaStream_t stream);   // ConvGemmArgs::split_c > 0 can be honoured (CTA-pair kernel)
// conv_pair.cu: the CTA-pair (tcgen05 cta_group::2) version with the TMA-staged epilogue
bool conv_pair_supported(const ConvGemmArgs& a);
int conv_pair_tiles(const ConvGemmArgs& a);
int conv_pair_colsum_slots(const ConvGemmArgs& a);
int conv_pair_launch(const ConvGemmArgs& a, cudaStream_t stream);

struct ConvWgradArgs {
  const __nv_bfloat16* g;  // (B,H,W,Cout) gradient w.r.t. the conv output
  const __nv_bfloat16* x;  // (B,H,W,Cin) conv input
  float* dw;               // [Cout][ksize*ksize][Cin] fp32
  int B, H, W, Cin, Cout, ksize;
  float alpha;
  int accumulate;          // 0: dw = result, 1: dw += result
  int splits_override;     // 0 = heuristic
};
int conv_wgrad_launch(const ConvWgradArgs& a, cudaStream_t stream);

// ---- elementwise.cu ----
struct PrepArgs {
  const __nv_bfloat16* in;    // (B,Hin,Win,C1)
  const __nv_bfloat16* skip;  // (B,Hin,Win,C2) or null
  const float* gain;          // (B,C2) ScaleLong gain or null
  __nv_bfloat16* x_out;       // (B,H,W,C1+C2) or null
  __nv_bfloat16* a_out;       // mp_silu(x) or null
  float* nrm_out;             // (B,H,W) eps + rms_C, written when pixelnorm (may be null)
  int B, Hin, Win, C1, C2;
  int resample;               // 0 none, 1 avg-pool 2x2, 2 nearest-exact x2
  int pixelnorm;
};
int block_prep_forward(const PrepArgs& a, cudaStream_t stream);

struct PrepBwdArgs {
  const __nv_bfloat16* g_res;  // gradient reaching x through the residual path (scaled by beta) or null
  float beta;
  const __nv_bfloat16* g_a;    // gradient w.r.t. mp_silu(x) or null
  const __nv_bfloat16* x;      // saved x (needed with g_a or pixelnorm)
  const float* nrm;            // saved eps + rms (pixelnorm)
  const float* gain;           // (B,C2) or null
  const float* d_mean;         // (B,C2) gradient w.r.t. the spatial mean of skip (ScaleLong path) or null
  __nv_bfloat16* g_in;         // (B,Hin,Win,C1)
  __nv_bfloat16* g_skip;       // (B,Hin,Win,C2) or null
  int accumulate_in, accumulate_skip;
  int B, Hin, Win, C1, C2, resample, pixelnorm;
};
int block_prep_backward(const PrepBwdArgs& a, cudaStream_t stream);

struct ModSiluBwdArgs {
  const __nv_bfloat16* g_h;   // (B,HW,C)
  const __nv_bfloat16* raw;   // (B,HW,C) un-modulated conv output
  const float* mod;           // (B, mod_stride), column offset applied
  float* d_mod;               // same indexing as mod; accumulated atomically (zero it first)
  __nv_bfloat16* g_raw;       // (B,HW,C)
  int B, HW, C, mod_stride;
  float drop_p;
  uint32_t seed_lo, seed_hi;
  const unsigned long long* seed_ptr;
};
int modsilu_backward(const ModSiluBwdArgs& a, cudaStream_t stream);

struct ChannelDotArgs {
  const __nv_bfloat16* A;   // (B,HW,CA), columns [a_off, a_off + C) used
  const __nv_bfloat16* Bm;  // (B,HW,C) or null (plain channel sum)
  float* out;               // (B,C) accumulated atomically (zero it first)
  int B, HW, C, CA, a_off;
  float scale;
  int group_c;              // set by the launcher: > 0 = one CTA per group of this many channels (deterministic mode)
};
int channel_dot(const ChannelDotArgs& a, cudaStream_t stream);

// ---- attention_tc.cu: tcgen05/TMEM forward for head_dim 64, S in {64, 256} ----
bool attention_tc_supported(int S, int hd);
int attention_forward_tc(const __nv_bfloat16* qkv, __nv_bfloat16* y, float* lse, int B, int S, int heads, cudaStream_t stream);
int attention_backward_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* g_y, const float* lse,
                          float* delta, __nv_bfloat16* g_qkv, int B, int S, int heads, cudaStream_t stream);

// ---- attention_gen.cu: tcgen05/TMEM attention for any head_dim % 16 == 0 (<= 192), S <= 256, on pre-normalised q, k, v ----
int qkv_normalize(const __nv_bfloat16* qkv, __nv_bfloat16* qn, float* norms, long long rows, int heads, int hd, cudaStream_t stream);
int attention_forward_normalized(const __nv_bfloat16* qn, __nv_bfloat16* y, float* lse, int B, int S, int heads, int hd,
                                 cudaStream_t stream);
int attention_backward_normalized(const __nv_bfloat16* qn, const float* norms, const __nv_bfloat16* y, const __nv_bfloat16* g_y,
                                  const float* lse, float* delta, __nv_bfloat16* g_qkv, int B, int S, int heads, int hd,
                                  cudaStream_t stream);

// ---- attention.cu ----
int attention_forward(const __nv_bfloat16* qkv, __nv_bfloat16* y, float* lse, int B, int S, int heads, int hd,
                      cudaStream_t stream);
int attention_backward(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* g_y, const float* lse,
                       float* delta, __nv_bfloat16* g_qkv, int B, int S, int heads, int hd, cudaStream_t stream);

// ---- small.cu ----
int sgemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
          float alpha, float beta, cudaStream_t stream);

struct EmbeddingArgs {
  const float* sigma; int sigma_stride;      // (B,) or 0-d (stride 0)
  const float* freqs; const float* phases;   // (F,)
  const float* w_sigma;                      // w_hat fp32 (E,F)
  const float* w_class;                      // w_hat fp32 (E,n_classes) or null
  const long long* labels;                   // (B,) int64 or null
  float* fourier; float* pre; float* emb;    // (B,F), (B,E), (B,E)
  int B, F, E, n_classes;
  float add_factor;
};
int embedding_forward(const EmbeddingArgs& a, cudaStream_t stream);
struct EmbeddingBwdArgs {
  const float* g_emb; const float* pre; const long long* labels;
  float* g_sig;        // (B,E): gradient w.r.t. sigma_embed output
  float* g_w_class;    // (E,n_classes) accumulated atomically (zero first) or null
  int B, E, n_classes;
  float add_factor;
};
int embedding_backward(const EmbeddingBwdArgs& a, cudaStream_t stream);
int mod_finish_forward(const float* lin, const float* const* gains, const int* col_block, float* m, int B, int N,
                       cudaStream_t stream);
int mod_finish_backward(const float* lin, const float* dm, const float* const* gains, const int* blk_start, float* d_lin,
                        float* d_gain, int B, int N, int n_blocks, cudaStream_t stream);

struct ScaleLongArgs {
  const float* mean;   // (B,C) spatial mean of skip
  const float* w1;     // w_hat fp32 (R, C+1)
  const float* w2;     // w_hat fp32 (C, R)
  float* aug_out;      // (B,C+1) = [mean, 1]
  float* h_pre;        // (B,R)
  float* h_out;        // (B,R) mp_silu(h_pre)
  float* gain;         // (B,C)
  int B, C, R;
};
int scalelong_forward(const ScaleLongArgs& a, cudaStream_t stream);
struct ScaleLongBwdArgs {
  const float* d_gain; const float* gain; const float* h_pre; const float* w1; const float* w2;
  float* d_pre2;   // (B,C)
  float* d_hpre;   // (B,R)
  float* d_mean;   // (B,C)
  int B, C, R;
  int d_gain_times_gain;   // d_gain holds (d gain) * gain (the split conv epilogue's reduction over skip * gain)
};
int scalelong_backward(const ScaleLongBwdArgs& a, cudaStream_t stream);
int scalelong_wgrad(const float* d_pre2, const float* h, const float* d_hpre, const float* aug, float* dw2, float* dw1, int B,
                    int C, int R, cudaStream_t stream);

struct UncertaintyArgs {
  const float* fourier; const float* w1; const float* w2; const float* gain;
  float* aug_out; float* h_pre; float* h_out; float* u_raw; float* u;
  int B, F;
};
int uncertainty_forward(const UncertaintyArgs& a, cudaStream_t stream);
struct UncertaintyBwdArgs {
  const float* g_u; const float* gain; const float* w2; const float* h_pre;
  float* g_uraw; float* g_hpre;
  int B, F;
};
int uncertainty_backward(const UncertaintyBwdArgs& a, cudaStream_t stream);

int conv_in_im2col(const float* noisy, const float* sigma, int sigma_stride, float sigma_data, __nv_bfloat16* out, int B,
                   int Ci, int H, int W, cudaStream_t stream);
struct ConvOutArgs {
  const __nv_bfloat16* x;   // (B,HW,C)
  const __nv_bfloat16* w;   // w_hat bf16 [Co][C]
  const float* gain_out;    // 0-d
  const float* noisy;       // (B,Co,H,W) fp32
  const float* sigma; int sigma_stride;
  float sigma_data;
  float* f_raw;             // (B,Co,H,W) raw conv output (saved for backward) or null
  float* D;                 // (B,Co,H,W)
  int B, HW, C, Co;
};
int conv_out_forward(const ConvOutArgs& a, cudaStream_t stream);
struct ConvOutBwdArgs {
  const float* g_D; const float* f_raw; const __nv_bfloat16* x; const __nv_bfloat16* w; const float* gain_out;
  const float* sigma; int sigma_stride; float sigma_data;
  __nv_bfloat16* g_x;   // (B,HW,C)
  float* g_w;           // [Co][C] accumulated atomically (zero first)
  float* g_gain_out;    // scalar accumulated atomically (zero first)
  int B, HW, C, Co;
};
int conv_out_backward(const ConvOutBwdArgs& a, cudaStream_t stream);
int bias_add_bc(__nv_bfloat16* g, const float* bias, float scale, int B, int HW, int C, cudaStream_t stream);
int to_uint8_images(const float* x, const float* mean, const float* std, uint8_t* out, int B, int C, int HW, cudaStream_t stream);
int wmse_forward(const float* D, const float* y, const float* sigma, const float* u, const float* weight, float sigma_data,
                 float* mse, float* wsum, float* loss, int B, int n, cudaStream_t stream);
int wmse_backward(const float* D, const float* y, const float* sigma, const float* u, const float* weight, const float* mse,
                  const float* g_loss, float sigma_data, float* g_D, float* g_u, float* g_weight, int B, int n,
                  cudaStream_t stream);
int heun_step(const float* x0, const float* x1, const float* D, const float* d_prev, float* x_out, float* d_out,
              const float* ts, int step, int mode, long long n, cudaStream_t stream);
int diffuse(const float* clean, const float* eps, const float* noise, float P_mean, float P_std, float* noisy,
            float* sigma, int B, int n, cudaStream_t stream);
int diffuse_philox(const float* clean, const long long* state, float P_mean, float P_std,
                   float sigma_data, float* noisy, float* sigma, __nv_bfloat16* xcol, int B, int Ci, int H, int W,
                   cudaStream_t stream);
int philox_normal_draws(const long long* state, float* eps, float* noise, int B, long long n, cudaStream_t stream);

// ---- optim.cu ----
int adam_chunk_elems();
int adam_ema_step(const tedm_adam_desc* table, const int32_t* chunks, int n_chunks, float lr, float step, const float* hyper,
                  float beta1, float beta2, float eps, float gamma, cudaStream_t stream);

}  // namespace tedm
