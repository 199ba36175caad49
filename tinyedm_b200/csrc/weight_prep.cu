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
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr float kEps = 1e-4f;

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

// Two launches. weight_norm_kernel: one warp per parameter row, 16-byte loads four deep, writes
// stats[row] = {1/(eps*sqrt(n)+||w||), ||w||, s1, 0} (s1 = the in-place rescale of training mode, else 1).
// weight_prep_fwd_kernel: one CTA per (group of 16 consecutive PREPARED rows, fan-in tile of whole input channels);
// blockIdx.y strides over the tiles of long rows, so the CIFAR net is ~6 000 short CTAs instead of 2 000 long ones whose
// ragged last wave cost a third of the kernel (30 % of the HBM peak, profiles/r1p_elementwise_ncu.txt):
//   coalesced read (+ in-place rewrite in training mode, + fp32 w_hat) into padded shared memory, then the two bf16
//   operand layouts are written with the fastest-varying thread index running along THEIR contiguous axis (input channel
//   for out_fwd, output row for out_dgrad) — a row-per-CTA version wrote 2-byte elements 512 bytes apart and reached
//   11-16 % of HBM bandwidth.
constexpr int kTileSplit = 4;   // gridDim.y of weight_prep_fwd_kernel

__global__ void __launch_bounds__(kFwdThreads)
weight_norm_kernel(const WeightDesc* __restrict__ table, int n_tensors, int training) {
  pdl_trigger();
  pdl_wait();
  const int group = blockIdx.x >> 1;
  const int ti = find_group_tensor(table, n_tensors, group);
  const WeightDesc d = table[ti];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = (group - d.group_start) * kGroupRows + (blockIdx.x & 1) * 8 + warp;   // parameter row
  if (r >= d.rows || d.stats == nullptr) return;
  const int fan_in = d.cin * d.taps;
  const float* w = static_cast<const float*>(d.w) + (size_t)r * fan_in;
  float ss = 0.f;
  if ((fan_in & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const int n4 = fan_in / 4;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int j = lane;
    for (; j + 96 < n4; j += 128) {
      const float4 a = w4[j], b = w4[j + 32], c = w4[j + 64], e = w4[j + 96];
      s0 += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      s1 += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
      s2 += c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
      s3 += e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w;
    }
    for (; j < n4; j += 32) {
      const float4 a = w4[j];
      s0 += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    ss = (s0 + s1) + (s2 + s3);
  } else {
    for (int j = lane; j < fan_in; j += 32) ss += w[j] * w[j];
  }
  ss = warp_sum(ss);
  if (lane == 0) {
    const float sqrt_n = sqrtf((float)fan_in);
    float norm = sqrtf(ss);
    float s1 = 1.0f;
    if (training) {
      s1 = 1.0f / (kEps + norm / sqrt_n);
      norm *= s1;
    }
    const float inv_s = 1.0f / (kEps * sqrt_n + norm);
    reinterpret_cast<float4*>(d.stats)[r] = make_float4(inv_s, norm, s1, 0.f);
  }
}

__global__ void __launch_bounds__(kFwdThreads)
weight_prep_fwd_kernel(const WeightDesc* __restrict__ table, int n_tensors, int training) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs weight_norm_kernel complete
  __shared__ float tile[kGroupRows * kTileStride];
  const int ti = find_group_tensor(table, n_tensors, blockIdx.x);
  const WeightDesc d = table[ti];
  const int o0 = (blockIdx.x - d.group_start) * kGroupRows;
  const int nrows = d.rows - o0 < kGroupRows ? d.rows - o0 : kGroupRows;
  const int taps = d.taps, cin = d.cin;
  const int fan_in = cin * taps;
  float* wbase = static_cast<float*>(d.w);
  const float4* stats4 = static_cast<const float4*>(d.stats);
  float* out_f32 = static_cast<float*>(d.out_f32);
  __nv_bfloat16* out_fwd = static_cast<__nv_bfloat16*>(d.out_fwd);
  __nv_bfloat16* out_dgrad = static_cast<__nv_bfloat16*>(d.out_dgrad);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tiles of `ci_tile` whole input channels; this CTA takes tiles blockIdx.y, blockIdx.y + gridDim.y, ...
  const int ci_tile = kTileElems / taps < cin ? kTileElems / taps : cin;
  for (int c0 = blockIdx.y * ci_tile; c0 < cin; c0 += gridDim.y * ci_tile) {
    const int nci = cin - c0 < ci_tile ? cin - c0 : ci_tile;
    const int nel = nci * taps;   // contiguous elements [c0*taps, c0*taps + nel) of every row
    for (int r = warp; r < nrows; r += kFwdThreads / 32) {
      const int row = src_row(d, o0 + r);
      float* w = wbase + (size_t)row * fan_in + (size_t)c0 * taps;
      const float4 st = stats4[row];
      const float s1 = st.z, inv_s = st.x;
      if ((nel & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 && out_f32 == nullptr) {
        float4* w4 = reinterpret_cast<float4*>(w);
        const int n4 = nel / 4;
        for (int j0 = lane; j0 < n4; j0 += 128) {   // four 16-byte loads in flight per lane
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j0 + 32 * u < n4) v[u] = w4[j0 + 32 * u];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j4 = j0 + 32 * u;
            if (j4 < n4) {
              v[u].x *= s1; v[u].y *= s1; v[u].z *= s1; v[u].w *= s1;
              if (training) w4[j4] = v[u];
              float* t = tile + r * kTileStride + 4 * j4;
              t[0] = v[u].x * inv_s; t[1] = v[u].y * inv_s; t[2] = v[u].z * inv_s; t[3] = v[u].w * inv_s;
            }
          }
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
      __nv_bfloat16* obase = out_fwd + (size_t)o0 * d.kpad + c0;
      if ((nci & 7) == 0 && ((cin | d.kpad) & 7) == 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0) {
        // 16-byte stores: a lane owns 8 consecutive input channels of one (row, tap); nci / 8 lanes cover the pair and a
        // warp store instruction writes 512 contiguous-by-128 bytes (4 pairs at nci = 64) instead of 128 — a quarter of the
        // store instructions of the 4-byte version, which was issue/latency bound (26-38 % of HBM peak)
        const int lanes_per_pair = nci >> 3;
        const int n_items = nrows * taps * lanes_per_pair;
        for (int it = threadIdx.x; it < n_items; it += kFwdThreads) {
          const int rt = it / lanes_per_pair, cl = (it - rt * lanes_per_pair) * 8;
          const int r = rt / taps, tap = rt - r * taps;
          const float* trow = tile + r * kTileStride + tap + cl * taps;
          uint4 o;
          o.x = pack_bf16(trow[0], trow[taps]);
          o.y = pack_bf16(trow[2 * taps], trow[3 * taps]);
          o.z = pack_bf16(trow[4 * taps], trow[5 * taps]);
          o.w = pack_bf16(trow[6 * taps], trow[7 * taps]);
          *reinterpret_cast<uint4*>(obase + (size_t)r * d.kpad + tap * cin + cl) = o;
        }
      } else {
        // one warp per (row, tap): lanes run along ci, two channels per lane (4-byte stores, 128 contiguous bytes per warp)
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
    }
    if (out_dgrad != nullptr) {
      if (nrows == kGroupRows && (d.rows & 7) == 0 && (reinterpret_cast<uintptr_t>(out_dgrad) & 15) == 0) {
        // (ci, tap) pairs x 16 rows = 32 bytes: two lanes with one 16-byte store each (8 rows per lane), 16 pairs per warp
        // store instruction instead of 4
        const int half = threadIdx.x & 1;              // rows 0-7 / 8-15 of the group
        for (int j = threadIdx.x >> 1; j < nel; j += kFwdThreads / 2) {
          const int cl = j / taps, tap = j - cl * taps;
          const float* tcol = tile + (half * 8) * kTileStride + j;
          uint4 o;
          o.x = pack_bf16(tcol[0], tcol[kTileStride]);
          o.y = pack_bf16(tcol[2 * kTileStride], tcol[3 * kTileStride]);
          o.z = pack_bf16(tcol[4 * kTileStride], tcol[5 * kTileStride]);
          o.w = pack_bf16(tcol[6 * kTileStride], tcol[7 * kTileStride]);
          *reinterpret_cast<uint4*>(out_dgrad + ((size_t)(c0 + cl) * taps + (taps - 1 - tap)) * d.rows + o0 + half * 8) = o;
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
  if (out_fwd != nullptr && d.kpad > fan_in && blockIdx.y == 0) {
    const int pad = d.kpad - fan_in;
    for (int idx = threadIdx.x; idx < nrows * pad; idx += kFwdThreads) {
      const int r = idx / pad, j = idx - r * pad;
      out_fwd[(size_t)(o0 + r) * d.kpad + fan_in + j] = __float2bfloat16_rn(0.f);
    }
  }
}

// Backward: one CTA per parameter row. The dL/dw_hat row ([tap][cin], fp32) is staged in shared memory with four floats
// of padding per tap (16-byte stores stay aligned, the gather g[tap*cin + ci] for j = ci*taps + tap, four consecutive j per
// lane, is at worst 2-way conflicted); every global
// access is a coalesced 16-byte one when the row allows it (fan-in and cin multiples of 4), four of them in flight per
// thread (keeping the row in registers between the two passes bought nothing: the re-read hits L1 / L2).

template <int kBwdThreads>
__device__ __forceinline__ float block_sum_bwd(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kBwdThreads / 32; ++i) t += red[i];
  return t;
}

// TAPS = 9 / 1: compile-time division; 0: any tap count (d.taps)
template <int TAPS, int kBwdThreads>
__device__ __forceinline__ void weight_bwd_row(const float* __restrict__ w, const float* __restrict__ g, float* __restrict__ out,
                                               float* grow, float* red, int taps_rt, int cin, float inv_s, float norm) {
  const int taps = TAPS > 0 ? TAPS : taps_rt;
  const int fan_in = cin * taps;
  const int gs = cin + 4;
  const bool vec = ((cin & 3) == 0) && (((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) |
                                          reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  if (vec) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const int n4g = fan_in / 4;
    for (int k0 = threadIdx.x; k0 < n4g; k0 += 4 * kBwdThreads) {   // four 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k0 + u * kBwdThreads < n4g) v[u] = g4[k0 + u * kBwdThreads];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k4 = k0 + u * kBwdThreads;
        if (k4 < n4g) {
          const int k = 4 * k4, tap = k / cin, ci = k - tap * cin;   // cin % 4 == 0: the four elements share a tap
          *reinterpret_cast<float4*>(grow + tap * gs + ci) = v[u];
        }
      }
    }
  } else {
    for (int k = threadIdx.x; k < fan_in; k += kBwdThreads) {
      const int tap = k / cin, ci = k - tap * cin;
      grow[tap * gs + ci] = g[k];
    }
  }
  __syncthreads();
  float dot = 0.f;
  if (vec) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
    float4* o4 = reinterpret_cast<float4*>(out);
    const int n4 = fan_in / 4;
    auto gather = [&](int j4, float (&gv)[4]) {
      const int j = 4 * j4;
      int ci = j / taps, tap = j - ci * taps;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        gv[e] = grow[tap * gs + ci];
        if (++tap == taps) { tap = 0; ++ci; }
      }
    };
    // four 16-byte loads of w in flight per thread in both passes; the second pass re-reads the row from L1 / L2
    for (int j0 = threadIdx.x; j0 < n4; j0 += 4 * kBwdThreads) {
      float4 wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u * kBwdThreads < n4) wv[u] = w4[j0 + u * kBwdThreads];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u * kBwdThreads < n4) {
          float gv[4];
          gather(j0 + u * kBwdThreads, gv);
          dot += wv[u].x * gv[0] + wv[u].y * gv[1] + wv[u].z * gv[2] + wv[u].w * gv[3];
        }
      }
    }
    dot = block_sum_bwd<kBwdThreads>(dot, red);
    const float c2 = dot * inv_s * inv_s / fmaxf(norm, 1e-30f);
    for (int j0 = threadIdx.x; j0 < n4; j0 += 4 * kBwdThreads) {
      float4 wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u * kBwdThreads < n4) wv[u] = w4[j0 + u * kBwdThreads];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j4 = j0 + u * kBwdThreads;
        if (j4 < n4) {
          float gv[4];
          gather(j4, gv);
          o4[j4] = make_float4(gv[0] * inv_s - wv[u].x * c2, gv[1] * inv_s - wv[u].y * c2, gv[2] * inv_s - wv[u].z * c2,
                               gv[3] * inv_s - wv[u].w * c2);
        }
      }
    }
  } else {
    for (int j = threadIdx.x; j < fan_in; j += kBwdThreads) {
      const int ci = j / taps, tap = j - ci * taps;
      dot += w[j] * grow[tap * gs + ci];
    }
    dot = block_sum_bwd<kBwdThreads>(dot, red);
    const float c2 = dot * inv_s * inv_s / fmaxf(norm, 1e-30f);
    for (int j = threadIdx.x; j < fan_in; j += kBwdThreads) {
      const int ci = j / taps, tap = j - ci * taps;
      out[j] = grow[tap * gs + ci] * inv_s - w[j] * c2;
    }
  }
}

template <int kBwdThreads>
__global__ void __launch_bounds__(kBwdThreads)
weight_prep_bwd_kernel(const WeightDesc* __restrict__ table, int n_tensors) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  extern __shared__ __align__(16) float grow[];   // [taps][cin + 4]
  __shared__ float red[kBwdThreads / 32];
  const int ti = find_tensor(table, n_tensors, blockIdx.x);
  const WeightDesc d = table[ti];
  if (d.g_hat == nullptr || d.grad == nullptr) return;
  const int row = blockIdx.x - d.row_start;
  const int fan_in = d.cin * d.taps;
  const float* w = static_cast<const float*>(d.w) + (size_t)row * fan_in;
  const float* g = static_cast<const float*>(d.g_hat) + (size_t)out_row(d, row) * d.kpad;
  float* out = static_cast<float*>(d.grad) + (size_t)row * fan_in;
  const float* stats = static_cast<const float*>(d.stats);
  const float inv_s = stats[4 * row + 0];
  const float norm = stats[4 * row + 1];
  if (d.taps == 9) weight_bwd_row<9, kBwdThreads>(w, g, out, grow, red, 9, d.cin, inv_s, norm);
  else if (d.taps == 1) weight_bwd_row<1, kBwdThreads>(w, g, out, grow, red, 1, d.cin, inv_s, norm);
  else weight_bwd_row<0, kBwdThreads>(w, g, out, grow, red, d.taps, d.cin, inv_s, norm);
}

}  // namespace

int weight_prep_forward(const WeightDesc* table_dev, int n_tensors, int total_groups, int training, cudaStream_t stream) {
  if (total_groups <= 0) return 0;
  launch_pdl(weight_norm_kernel, 2 * total_groups, kFwdThreads, 0, stream, table_dev, n_tensors, training);
  TEDM_LAUNCH_CHECK();
  launch_pdl(weight_prep_fwd_kernel, dim3(total_groups, kTileSplit), kFwdThreads, 0, stream, table_dev, n_tensors, training);
  TEDM_LAUNCH_CHECK();
  return 0;
}

template <int T>
static int launch_weight_prep_bwd(const WeightDesc* table_dev, int n_tensors, int total_rows, size_t smem, cudaStream_t stream) {
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    TEDM_CUDA(cudaFuncSetAttribute(weight_prep_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_pdl(weight_prep_bwd_kernel<T>, total_rows, T, smem, stream, table_dev, n_tensors);
  TEDM_LAUNCH_CHECK();
  return 0;
}

int weight_prep_backward(const WeightDesc* table_dev, int n_tensors, int total_rows, int max_row_floats, cudaStream_t stream) {
  if (total_rows <= 0) return 0;
  const size_t smem = (size_t)max_row_floats * sizeof(float);
  TEDM_CHECK(smem <= 200 * 1024, "weight_prep_bwd: fan-in too large (%d floats)", max_row_floats);
  // short rows: 128 threads (more CTAs per SM to overlap the load / reduce / store phases of different rows);
  // long rows (the staging buffer limits the CTAs per SM): 256 threads. Measured: profiles/r1p_weight_prep.txt
  if (smem > 24 * 1024) return launch_weight_prep_bwd<256>(table_dev, n_tensors, total_rows, smem, stream);
  return launch_weight_prep_bwd<128>(table_dev, n_tensors, total_rows, smem, stream);
}

}  // namespace tedm
