// Multi-tensor forced weight normalisation (EDM2 "MPConv" weights), forward and backward.
//
// Reference: src/tinyedm/networks.py:17-19 (normalize), :32-36 / :55-59 (Conv2d / Linear forward):
//   training:  w <- w / (eps + ||w||/sqrt(n))                    (in place, no grad)
//   always:    w_hat = w / (eps + ||w||/sqrt(n)) / sqrt(n) = w / (eps*sqrt(n) + ||w||)
// One launch handles EVERY weight tensor of the model: one CTA per output row (filter); a device-side
// descriptor table says where each tensor's raw fp32 rows live and which operand layouts to emit:
//   out_fwd   bf16 [rows][kpad]            k = tap*Cin + ci            (implicit-GEMM B operand, forward)
//   out_dgrad bf16 [Cin][taps][rows]       tap flipped                 (B operand of the data gradient)
//   out_f32   fp32 [rows][fan_in]          same layout as the parameter (small fp32 layers)
// Backward (closed form verified against autograd):
//   dL/dw = g/s - w (w.g) / (s^2 ||w||),  s = eps*sqrt(n) + ||w||,   g = dL/dw_hat
#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr int kThreads = 128;
constexpr float kEps = 1e-4f;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += red[i];
  __syncthreads();
  return t;
}

// Output row of parameter row `row`. For the qkv conv of CosineAttention the reference's channel order is
// head*3*hd + d*3 + {q,k,v} (networks.py:194); the prepared operands use {q,k,v}*C + head*hd + d instead, so the
// convolution itself emits de-interleaved q | k | v planes (a free re-layout: only weight rows move).
__device__ __forceinline__ int out_row(const WeightDesc& d, int row) {
  const int hd = d.qkv_head_dim;
  if (hd == 0) return row;
  const int head = row / (3 * hd), rem = row - head * 3 * hd;
  const int dd = rem / 3, j = rem - dd * 3;
  return j * (d.rows / 3) + head * hd + dd;
}

__device__ __forceinline__ int find_tensor(const WeightDesc* table, int n, int row) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (table[mid].row_start <= row) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kThreads)
weight_prep_fwd_kernel(const WeightDesc* __restrict__ table, int n_tensors, int training) {
  __shared__ float red[kThreads / 32];
  const int ti = find_tensor(table, n_tensors, blockIdx.x);
  const WeightDesc d = table[ti];
  const int row = blockIdx.x - d.row_start;
  const int fan_in = d.cin * d.taps;
  float* w = static_cast<float*>(d.w) + (size_t)row * fan_in;
  float* stats = static_cast<float*>(d.stats);
  float* out_f32 = static_cast<float*>(d.out_f32);
  __nv_bfloat16* out_fwd = static_cast<__nv_bfloat16*>(d.out_fwd);
  __nv_bfloat16* out_dgrad = static_cast<__nv_bfloat16*>(d.out_dgrad);

  float ss = 0.f;
  for (int j = threadIdx.x; j < fan_in; j += kThreads) {
    float v = w[j];
    ss += v * v;
  }
  ss = block_sum(ss, red);
  float norm = sqrtf(ss);
  const float sqrt_n = sqrtf((float)fan_in);
  float s1 = 1.0f;
  if (training) {
    s1 = 1.0f / (kEps + norm / sqrt_n);
    norm *= s1;
  }
  const float inv_s = 1.0f / (kEps * sqrt_n + norm);
  if (threadIdx.x == 0 && stats != nullptr) {
    stats[2 * row + 0] = inv_s;
    stats[2 * row + 1] = norm;
  }
  const int taps = d.taps, cin = d.cin;
  const int orow = out_row(d, row);
  for (int j = threadIdx.x; j < fan_in; j += kThreads) {
    float v = w[j] * s1;
    if (training) w[j] = v;
    const float wh = v * inv_s;
    const int ci = j / taps, tap = j - ci * taps;
    if (out_f32 != nullptr) out_f32[(size_t)row * fan_in + j] = wh;
    if (out_fwd != nullptr) out_fwd[(size_t)orow * d.kpad + tap * cin + ci] = __float2bfloat16_rn(wh);
    if (out_dgrad != nullptr)
      out_dgrad[((size_t)ci * taps + (taps - 1 - tap)) * d.rows + orow] = __float2bfloat16_rn(wh);
  }
  if (out_fwd != nullptr) {
    for (int j = fan_in + threadIdx.x; j < d.kpad; j += kThreads)
      out_fwd[(size_t)orow * d.kpad + j] = __float2bfloat16_rn(0.f);
  }
}

__global__ void __launch_bounds__(kThreads)
weight_prep_bwd_kernel(const WeightDesc* __restrict__ table, int n_tensors) {
  __shared__ float red[kThreads / 32];
  const int ti = find_tensor(table, n_tensors, blockIdx.x);
  const WeightDesc d = table[ti];
  if (d.g_hat == nullptr || d.grad == nullptr) return;
  const int row = blockIdx.x - d.row_start;
  const int taps = d.taps, cin = d.cin;
  const int fan_in = cin * taps;
  const float* w = static_cast<const float*>(d.w) + (size_t)row * fan_in;
  const float* g = static_cast<const float*>(d.g_hat) + (size_t)out_row(d, row) * d.kpad;
  float* out = static_cast<float*>(d.grad) + (size_t)row * fan_in;
  const float* stats = static_cast<const float*>(d.stats);
  float dot = 0.f;
  for (int j = threadIdx.x; j < fan_in; j += kThreads) {
    const int ci = j / taps, tap = j - ci * taps;
    dot += w[j] * g[tap * cin + ci];
  }
  dot = block_sum(dot, red);
  const float inv_s = stats[2 * row + 0];
  const float norm = stats[2 * row + 1];
  const float c2 = dot * inv_s * inv_s / fmaxf(norm, 1e-30f);
  for (int j = threadIdx.x; j < fan_in; j += kThreads) {
    const int ci = j / taps, tap = j - ci * taps;
    out[j] = g[tap * cin + ci] * inv_s - w[j] * c2;
  }
}

}  // namespace

int weight_prep_forward(const WeightDesc* table_dev, int n_tensors, int total_rows, int training, cudaStream_t stream) {
  if (total_rows <= 0) return 0;
  weight_prep_fwd_kernel<<<total_rows, kThreads, 0, stream>>>(table_dev, n_tensors, training);
  TEDM_LAUNCH_CHECK();
  return 0;
}

int weight_prep_backward(const WeightDesc* table_dev, int n_tensors, int total_rows, cudaStream_t stream) {
  if (total_rows <= 0) return 0;
  weight_prep_bwd_kernel<<<total_rows, kThreads, 0, stream>>>(table_dev, n_tensors);
  TEDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace tedm
