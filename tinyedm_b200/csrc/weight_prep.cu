// Multi-tensor forced weight normalisation (EDM2 "MPConv" weights), forward and backward.
//
// Reference: src/tinyedm/networks.py:17-19 (normalize), :32-36 / :55-59 (Conv2d / Linear forward):
//   training:  w <- w / (eps + ||w||/sqrt(n))                    (in place, no grad)
//   always:    w_hat = w / (eps + ||w||/sqrt(n)) / sqrt(n) = w / (eps*sqrt(n) + ||w||)
// One launch handles EVERY weight tensor of the model: one CTA per group of 16 prepared rows (filters); a device-side
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

// Inverse of out_row: which parameter row feeds prepared row `orow`.
__device__ __forceinline__ int src_row(const WeightDesc& d, int orow) {
  const int hd = d.qkv_head_dim;
  if (hd == 0) return orow;
  const int plane = d.rows / 3;
  const int j = orow / plane, rem = orow - j * plane;
  const int head = rem / hd, dd = rem - head * hd;
  return head * 3 * hd + dd * 3 + j;
}

__device__ __forceinline__ int find_group_tensor(const WeightDesc* table, int n, int group) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (table[mid].group_start <= group) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int kGroupRows = 16;     // prepared rows per CTA: 16 bf16 = one 32-byte sector of the data-gradient layout
constexpr int kFwdThreads = 256;
// fan-in elements staged per row and pass (64 input channels of a 3x3 filter). 288 (18.5 KB tiles, 8 CTAs per SM, the whole
// CIFAR net in one wave) was measured slower: 267 us vs 210 us — the extra barrier round trips cost more than the tail.
constexpr int kTileElems = 576;
constexpr int kTileStride = kTileElems + 1;   // odd stride: conflict-free column reads

// One CTA = 16 consecutive PREPARED rows of one tensor.
//   pass A  row norms (one warp per 2 rows, 16-byte loads)
//   pass B  fan-in tiles of whole input channels: coalesced read (+ in-place rewrite in training mode, + fp32 w_hat)
//           into padded shared memory, then the two bf16 operand layouts are written with the fastest-varying thread
//           index running along THEIR contiguous axis (input channel for out_fwd, output row for out_dgrad) — the
//           row-per-CTA version wrote 2-byte elements 512 bytes apart and reached 11-16 % of HBM bandwidth.
__global__ void __launch_bounds__(kFwdThreads)
weight_prep_fwd_kernel(const WeightDesc* __restrict__ table, int n_tensors, int training) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  __shared__ float tile[kGroupRows * kTileStride];
  __shared__ float s_s1[kGroupRows], s_inv[kGroupRows];
  const int ti = find_group_tensor(table, n_tensors, blockIdx.x);
  const WeightDesc d = table[ti];
  const int o0 = (blockIdx.x - d.group_start) * kGroupRows;
  const int nrows = d.rows - o0 < kGroupRows ? d.rows - o0 : kGroupRows;
  const int taps = d.taps, cin = d.cin;
  const int fan_in = cin * taps;
  float* wbase = static_cast<float*>(d.w);
  float* stats = static_cast<float*>(d.stats);
  float* out_f32 = static_cast<float*>(d.out_f32);
  __nv_bfloat16* out_fwd = static_cast<__nv_bfloat16*>(d.out_fwd);
  __nv_bfloat16* out_dgrad = static_cast<__nv_bfloat16*>(d.out_dgrad);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sqrt_n = sqrtf((float)fan_in);

  // ---- pass A: norms ----
  for (int r = warp; r < nrows; r += kFwdThreads / 32) {
    const int row = src_row(d, o0 + r);
    const float* w = wbase + (size_t)row * fan_in;
    float ss = 0.f;
    if ((fan_in & 3) == 0) {
      const float4* w4 = reinterpret_cast<const float4*>(w);
      for (int j = lane; j < fan_in / 4; j += 32) {
        const float4 v = w4[j];
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    } else {
      for (int j = lane; j < fan_in; j += 32) ss += w[j] * w[j];
    }
    ss = warp_sum(ss);
    if (lane == 0) {
      float norm = sqrtf(ss);
      float s1 = 1.0f;
      if (training) {
        s1 = 1.0f / (kEps + norm / sqrt_n);
        norm *= s1;
      }
      const float inv_s = 1.0f / (kEps * sqrt_n + norm);
      s_s1[r] = s1;
      s_inv[r] = inv_s;
      if (stats != nullptr) {
        stats[2 * row + 0] = inv_s;
        stats[2 * row + 1] = norm;
      }
    }
  }
  __syncthreads();

  // ---- pass B: tiles of `ci_tile` whole input channels ----
  const int ci_tile = kTileElems / taps < cin ? kTileElems / taps : cin;
  for (int c0 = 0; c0 < cin; c0 += ci_tile) {
    const int nci = cin - c0 < ci_tile ? cin - c0 : ci_tile;
    const int nel = nci * taps;   // contiguous elements [c0*taps, c0*taps + nel) of every row
    for (int r = warp; r < nrows; r += kFwdThreads / 32) {
      const int row = src_row(d, o0 + r);
      float* w = wbase + (size_t)row * fan_in + (size_t)c0 * taps;
      const float s1 = s_s1[r], inv_s = s_inv[r];
      if ((nel & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 && out_f32 == nullptr) {
        float4* w4 = reinterpret_cast<float4*>(w);
        for (int j4 = lane; j4 < nel / 4; j4 += 32) {
          float4 v = w4[j4];
          v.x *= s1; v.y *= s1; v.z *= s1; v.w *= s1;
          if (training) w4[j4] = v;
          float* t = tile + r * kTileStride + 4 * j4;
          t[0] = v.x * inv_s; t[1] = v.y * inv_s; t[2] = v.z * inv_s; t[3] = v.w * inv_s;
        }
      } else {
        for (int j = lane; j < nel; j += 32) {
          const float v = w[j] * s1;
          if (training) w[j] = v;
          const float wh = v * inv_s;
          if (out_f32 != nullptr) out_f32[(size_t)row * fan_in + (size_t)c0 * taps + j] = wh;
          tile[r * kTileStride + j] = wh;
        }
      }
    }
    __syncthreads();
    if (out_fwd != nullptr) {
      // one warp per (row, tap): lanes run along ci, two channels per lane (4-byte stores, 128 contiguous bytes per warp);
      // nested loops instead of per-element div/mod (the element-indexed version was integer-instruction bound)
      for (int rt = warp; rt < nrows * taps; rt += kFwdThreads / 32) {
        const int r = rt / taps, tap = rt - r * taps;
        const float* trow = tile + r * kTileStride + tap;
        __nv_bfloat16* orow_p = out_fwd + (size_t)(o0 + r) * d.kpad + tap * cin + c0;
        if ((nci & 1) == 0 && ((reinterpret_cast<uintptr_t>(orow_p) & 3) == 0)) {
          for (int cl = 2 * lane; cl < nci; cl += 64)
            *reinterpret_cast<uint32_t*>(orow_p + cl) = pack_bf16(trow[cl * taps], trow[(cl + 1) * taps]);
        } else {
          for (int cl = lane; cl < nci; cl += 32) orow_p[cl] = __float2bfloat16_rn(trow[cl * taps]);
        }
      }
    }
    if (out_dgrad != nullptr) {
      // (ci, tap) pairs x 16 rows: 8 lanes cover the 16 rows of a pair with 4-byte stores (one 32-byte sector), so a
      // warp handles 4 pairs per pass; pairs are walked as j = cl * taps + tap without div/mod in the inner loop
      const int rp = (lane & 7) * 2;           // this lane's two rows
      const int sub = lane >> 3;               // which of the warp's 4 pairs
      if (nrows == kGroupRows && (d.rows & 1) == 0) {
        for (int j = warp * 4 + sub; j < nel; j += (kFwdThreads / 32) * 4) {
          const int cl = j / taps, tap = j - cl * taps;
          const uint32_t v = pack_bf16(tile[rp * kTileStride + j], tile[(rp + 1) * kTileStride + j]);
          *reinterpret_cast<uint32_t*>(out_dgrad + ((size_t)(c0 + cl) * taps + (taps - 1 - tap)) * d.rows + o0 + rp) = v;
        }
      } else {
        for (int idx = threadIdx.x; idx < nel * kGroupRows; idx += kFwdThreads) {
          const int r = idx & (kGroupRows - 1), j = idx >> 4;
          if (r < nrows) {
            const int cl = j / taps, tap = j - cl * taps;
            out_dgrad[((size_t)(c0 + cl) * taps + (taps - 1 - tap)) * d.rows + o0 + r] =
                __float2bfloat16_rn(tile[r * kTileStride + j]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (out_fwd != nullptr && d.kpad > fan_in) {
    const int pad = d.kpad - fan_in;
    for (int idx = threadIdx.x; idx < nrows * pad; idx += kFwdThreads) {
      const int r = idx / pad, j = idx - r * pad;
      out_fwd[(size_t)(o0 + r) * d.kpad + fan_in + j] = __float2bfloat16_rn(0.f);
    }
  }
}

// Backward: one CTA per parameter row. The dL/dw_hat row ([tap][cin], fp32) is staged in shared memory with one float
// of padding per tap so that the gather g[tap*cin + ci] for consecutive j = ci*taps + tap is conflict-free and every
// global access is coalesced.
__global__ void __launch_bounds__(kThreads)
weight_prep_bwd_kernel(const WeightDesc* __restrict__ table, int n_tensors) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  extern __shared__ float grow[];   // [taps][cin + 1]
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
  for (int k = threadIdx.x; k < fan_in; k += kThreads) {
    const int tap = k / cin, ci = k - tap * cin;
    grow[tap * (cin + 1) + ci] = g[k];
  }
  __syncthreads();
  // j = ci * taps + tap walks the parameter row; (ci, tap) advance incrementally (no div/mod per element)
  const int step_ci = kThreads / taps, step_tap = kThreads - step_ci * taps;
  const int ci0 = threadIdx.x / taps, tap0 = threadIdx.x - ci0 * taps;
  float dot = 0.f;
  {
    int ci = ci0, tap = tap0;
    for (int j = threadIdx.x; j < fan_in; j += kThreads) {
      dot += w[j] * grow[tap * (cin + 1) + ci];
      ci += step_ci; tap += step_tap;
      if (tap >= taps) { tap -= taps; ++ci; }
    }
  }
  dot = block_sum(dot, red);
  const float inv_s = stats[2 * row + 0];
  const float norm = stats[2 * row + 1];
  const float c2 = dot * inv_s * inv_s / fmaxf(norm, 1e-30f);
  {
    int ci = ci0, tap = tap0;
    for (int j = threadIdx.x; j < fan_in; j += kThreads) {
      out[j] = grow[tap * (cin + 1) + ci] * inv_s - w[j] * c2;
      ci += step_ci; tap += step_tap;
      if (tap >= taps) { tap -= taps; ++ci; }
    }
  }
}

}  // namespace

int weight_prep_forward(const WeightDesc* table_dev, int n_tensors, int total_groups, int training, cudaStream_t stream) {
  if (total_groups <= 0) return 0;
  launch_pdl(weight_prep_fwd_kernel, total_groups, kFwdThreads, 0, stream, table_dev, n_tensors, training);
  TEDM_LAUNCH_CHECK();
  return 0;
}

int weight_prep_backward(const WeightDesc* table_dev, int n_tensors, int total_rows, int max_row_floats, cudaStream_t stream) {
  if (total_rows <= 0) return 0;
  const size_t smem = (size_t)max_row_floats * sizeof(float);
  TEDM_CHECK(smem <= 200 * 1024, "weight_prep_bwd: fan-in too large (%d floats)", max_row_floats);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    TEDM_CUDA(cudaFuncSetAttribute(weight_prep_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_pdl(weight_prep_bwd_kernel, total_rows, kThreads, smem, stream, table_dev, n_tensors);
  TEDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace tedm
