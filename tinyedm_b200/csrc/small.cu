// Small fp32 layers and image-sized kernels of the EDM2 hot path (launch-latency / HBM bound).
//
//   sgemm                 generic fp32 GEMM on CUDA cores for the autocast-off islands (Linear, networks.py:46-64)
//   embedding_fwd/bwd     c_noise -> Fourier features -> sigma_embed -> class embed -> mp_add -> mp_silu
//                         (networks.py:121-178)
//   mod_finish fwd/bwd    m = embed(emb) * gain + 1 for all blocks at once (networks.py:255-258, :319-322)
//   scalelong fwd/bwd     learned skip gain (networks.py:106-118)
//   uncertainty fwd/bwd   log-variance head (networks.py:91-103)
//   conv_in_im2col        c_in * x, ones channel and the 3x3 patch gather feeding the tensor-core GEMM
//                         (networks.py:578-587)
//   conv_out fwd/bwd      1x1 conv to image channels fused with D = c_skip x + c_out gain_out F (:602-603)
//   wmse fwd/bwd          (uncertainty-)weighted MSE, metric.py:8-18 + edm.py:212-219
//   heun_step, diffuse    solvers.py:45-57, edm.py:84-93
#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

// ------------------------------------------------------------------------------------------------
// sgemm: C[M,N] = alpha * op(A) op(B) + beta * C   (row-major, explicit leading dimensions)
// ------------------------------------------------------------------------------------------------
constexpr int TS = 64, TK = 16;

__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, int M, int N, int K,
             int lda, int ldb, int ldc, int transA, int transB, float alpha, float beta, int k_chunk) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  __shared__ __align__(16) float As[TK][TS + 4];   // +4: rows stay 16-byte aligned for the float4 reads below
  __shared__ __align__(16) float Bs[TK][TS + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  // split-K: blockIdx.z owns K range [k_begin, k_end); partial results are combined with atomics into a zeroed C
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  const bool split = gridDim.z > 1;
  float acc[4][4] = {};
  // register-staged global loads: the tile of step k+1 is in flight while step k is multiplied out of shared memory
  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = threadIdx.x + t * 256;
      int mm, kk;
      if (transA) { mm = i % TS; kk = i / TS; } else { kk = i % TK; mm = i / TK; }
      const int gm = m0 + mm, gk = k0 + kk;
      ra[t] = 0.f;
      if (gm < M && gk < k_end) ra[t] = transA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      int nn, kb;
      if (transB) { kb = i % TK; nn = i / TK; } else { nn = i % TS; kb = i / TS; }
      const int gn = n0 + nn, gkb = k0 + kb;
      rb[t] = 0.f;
      if (gn < N && gkb < k_end) rb[t] = transB ? Bm[(size_t)gn * ldb + gkb] : Bm[(size_t)gkb * ldb + gn];
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = threadIdx.x + t * 256;
      int mm, kk;
      if (transA) { mm = i % TS; kk = i / TS; } else { kk = i % TK; mm = i / TK; }
      As[kk][mm] = ra[t];
      int nn, kb;
      if (transB) { kb = i % TK; nn = i / TK; } else { nn = i % TS; kb = i / TS; }
      Bs[kb][nn] = rb[t];
    }
  };
  if (k_begin < k_end) load_tile(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
    store_tile();
    __syncthreads();
    if (k0 + TK < k_end) load_tile(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) {
        float* c = C + (size_t)gm * ldc + gn;
        if (split) atomicAdd(c, alpha * acc[i][j]);
        else *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
      }
    }
}

// ------------------------------------------------------------------------------------------------
// embedding
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embedding_fwd_kernel(const EmbeddingArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  extern __shared__ float four_s[];  // [F]
  const int b = blockIdx.x;
  const float sigma = a.sigma[a.sigma_stride * b];
  const float c_noise = logf(sigma) * 0.25f;
  for (int f = threadIdx.x; f < a.F; f += blockDim.x) {
    const float v = cosf(c_noise * a.freqs[f] + a.phases[f]) * 1.41421356237f;
    four_s[f] = v;
    a.fourier[(size_t)b * a.F + f] = v;
  }
  __syncthreads();
  const float t = a.add_factor;
  const float inv_c = rsqrtf((1.f - t) * (1.f - t) + t * t);
  for (int e = threadIdx.x; e < a.E; e += blockDim.x) {
    const float* w = a.w_sigma + (size_t)e * a.F;
    float acc = 0.f;
    for (int f = 0; f < a.F; ++f) acc += w[f] * four_s[f];
    if (a.labels != nullptr) {
      const long long lab = a.labels[b];
      const float cls = a.w_class[(size_t)e * a.n_classes + lab] * sqrtf((float)a.n_classes);
      acc = ((1.f - t) * acc + t * cls) * inv_c;
    }
    a.pre[(size_t)b * a.E + e] = acc;
    a.emb[(size_t)b * a.E + e] = mp_silu_f(acc);
  }
}

// g_pre = g_emb * mp_silu'(pre); g_sig = g_pre * (1-t)/c (or g_pre); scatter class-weight gradient
__global__ void __launch_bounds__(256)
embedding_bwd_kernel(const EmbeddingBwdArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const int b = blockIdx.x;
  const float t = a.add_factor;
  const float inv_c = rsqrtf((1.f - t) * (1.f - t) + t * t);
  for (int e = threadIdx.x; e < a.E; e += blockDim.x) {
    const float gp = a.g_emb[(size_t)b * a.E + e] * mp_silu_grad_f(a.pre[(size_t)b * a.E + e]);
    if (a.labels != nullptr) {
      a.g_sig[(size_t)b * a.E + e] = gp * (1.f - t) * inv_c;
      const long long lab = a.labels[b];
      atomicAdd(a.g_w_class + (size_t)e * a.n_classes + lab, gp * t * inv_c * sqrtf((float)a.n_classes));
    } else {
      a.g_sig[(size_t)b * a.E + e] = gp;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// modulation finish: m[b, col] = lin[b, col] * gain[blk(col)] + 1
// ------------------------------------------------------------------------------------------------
__global__ void mod_finish_fwd_kernel(const float* __restrict__ lin, const float* const* __restrict__ gains,
                                      const int* __restrict__ col_block, float* __restrict__ m, int B, int N) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int col = (int)(i % N);
  m[i] = lin[i] * *gains[col_block[col]] + 1.0f;
}

// grid (block id, batch chunk): d_gain[blk] += sum dm*lin over the block's columns (atomic; zeroed by the caller);
// d_lin = dm * gain
__global__ void __launch_bounds__(256)
mod_finish_bwd_kernel(const float* __restrict__ lin, const float* __restrict__ dm, const float* const* __restrict__ gains,
                      const int* __restrict__ blk_start, float* __restrict__ d_lin, float* __restrict__ d_gain, int B,
                      int N) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  __shared__ float red[8];
  const int blk = blockIdx.x;
  const int c0 = blk_start[blk], c1 = blk_start[blk + 1];
  const int w = c1 - c0;
  const float g = *gains[blk];
  float acc = 0.f;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
      const size_t o = (size_t)b * N + c;
      const float d = dm[o];
      acc += d * lin[o];
      d_lin[o] = d * g;
    }
  }
  (void)w;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(d_gain + blk, t);
  }
}

// ------------------------------------------------------------------------------------------------
// ScaleLong MLP (networks.py:106-118): gain = sigmoid(W2 mp_silu(W1 [mean, 1])), R = C/16 hidden units
// ------------------------------------------------------------------------------------------------
// The arithmetic is ~10 MFLOP; what costs time is reading the two weight matrices (up to 2 x 148 KB, from L2) with
// enough loads in flight. A CTA therefore owns kSlG batch rows (every weight element it loads is used kSlG times), and
// each phase walks the weights in the direction that is contiguous in memory: W1 [R][C+1] row by row with the lanes
// along c, W2 [C][R] one 16-byte-vectorised row per thread (forward) or one row per warp iteration with the lanes along
// j (backward). The first version (one CTA per batch row, strided W2 reads) took 73 / 145 us forward / backward on the
// ImageNet-latent shapes (C = 768, B = 64).
constexpr int kSlG = 4;

__global__ void __launch_bounds__(256)
scalelong_fwd_kernel(const ScaleLongArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  extern __shared__ float sm[];  // aug[G][C+1], h[G][R]
  const int C = a.C, R = a.R, C1 = a.C + 1;
  float* aug = sm;
  float* h = sm + kSlG * C1;
  const int b0 = blockIdx.x * kSlG;
  for (int i = threadIdx.x; i < kSlG * C1; i += blockDim.x) {
    const int g = i / C1, c = i - g * C1;
    const int b = b0 + g;
    float v = 0.f;
    if (b < a.B) {
      v = c < C ? a.mean[(size_t)b * C + c] : 1.0f;
      a.aug_out[(size_t)b * C1 + c] = v;
    }
    aug[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < R; j += 8) {
    const float* w = a.w1 + (size_t)j * C1;
    float acc[kSlG] = {};
#pragma unroll 4
    for (int c = lane; c < C1; c += 32) {
      const float wv = w[c];
#pragma unroll
      for (int g = 0; g < kSlG; ++g) acc[g] = fmaf(wv, aug[g * C1 + c], acc[g]);
    }
#pragma unroll
    for (int g = 0; g < kSlG; ++g) {
      const float t = warp_sum(acc[g]);
      if (lane == 0 && b0 + g < a.B) {
        a.h_pre[(size_t)(b0 + g) * R + j] = t;
        const float hv = mp_silu_f(t);
        h[g * R + j] = hv;
        a.h_out[(size_t)(b0 + g) * R + j] = hv;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* w = a.w2 + (size_t)c * R;
    float acc[kSlG] = {};
    if ((R & 3) == 0) {
      for (int j = 0; j < R; j += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(w + j);
#pragma unroll
        for (int g = 0; g < kSlG; ++g) {
          const float* hg = h + g * R + j;
          acc[g] += wv.x * hg[0] + wv.y * hg[1] + wv.z * hg[2] + wv.w * hg[3];
        }
      }
    } else {
      for (int j = 0; j < R; ++j) {
        const float wv = w[j];
#pragma unroll
        for (int g = 0; g < kSlG; ++g) acc[g] = fmaf(wv, h[g * R + j], acc[g]);
      }
    }
#pragma unroll
    for (int g = 0; g < kSlG; ++g)
      if (b0 + g < a.B) a.gain[(size_t)(b0 + g) * C + c] = 1.0f / (1.0f + __expf(-acc[g]));
  }
}

__global__ void __launch_bounds__(256)
scalelong_bwd_kernel(const ScaleLongBwdArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  extern __shared__ float sm[];  // dp2[G][C], part[8][G][R], dhp[G][R]
  const int C = a.C, R = a.R, C1 = a.C + 1;
  float* dp2 = sm;
  float* part = sm + kSlG * C;
  float* dhp = part + 8 * kSlG * R;
  const int b0 = blockIdx.x * kSlG;
  for (int i = threadIdx.x; i < kSlG * C; i += blockDim.x) {
    const int g = i / C, c = i - g * C;
    const int b = b0 + g;
    float v = 0.f;
    if (b < a.B) {
      const float gn = a.gain[(size_t)b * C + c];
      // gain = sigmoid(pre2): d pre2 = d gain * g (1 - g); the split conv epilogue already delivers (d gain) * g
      v = a.d_gain[(size_t)b * C + c] * (a.d_gain_times_gain ? 1.0f : gn) * (1.0f - gn);
      a.d_pre2[(size_t)b * C + c] = v;
    }
    dp2[i] = v;
  }
  __syncthreads();
  // d h[g][j] = sum_c dp2[g][c] W2[c][j]: warp w takes c = w, w+8, ...; the lanes run along j (one contiguous W2 row)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j0 = 0; j0 < R; j0 += 64) {
    float acc[kSlG][2] = {};
    const int ja = j0 + lane, jb = j0 + 32 + lane;
#pragma unroll 4
    for (int c = warp; c < C; c += 8) {
      const float* w = a.w2 + (size_t)c * R;
      const float wa = ja < R ? w[ja] : 0.f;
      const float wb = jb < R ? w[jb] : 0.f;
#pragma unroll
      for (int g = 0; g < kSlG; ++g) {
        const float d = dp2[g * C + c];
        acc[g][0] = fmaf(d, wa, acc[g][0]);
        acc[g][1] = fmaf(d, wb, acc[g][1]);
      }
    }
#pragma unroll
    for (int g = 0; g < kSlG; ++g) {
      if (ja < R) part[(warp * kSlG + g) * R + ja] = acc[g][0];
      if (jb < R) part[(warp * kSlG + g) * R + jb] = acc[g][1];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSlG * R; i += blockDim.x) {
    const int g = i / R, j = i - g * R;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[(w * kSlG + g) * R + j];
    float v = 0.f;
    if (b0 + g < a.B) {
      v = t * mp_silu_grad_f(a.h_pre[(size_t)(b0 + g) * R + j]);
      a.d_hpre[(size_t)(b0 + g) * R + j] = v;
    }
    dhp[i] = v;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc[kSlG] = {};
#pragma unroll 4
    for (int j = 0; j < R; ++j) {
      const float wv = a.w1[(size_t)j * C1 + c];
#pragma unroll
      for (int g = 0; g < kSlG; ++g) acc[g] = fmaf(dhp[g * R + j], wv, acc[g]);
    }
#pragma unroll
    for (int g = 0; g < kSlG; ++g)
      if (b0 + g < a.B) a.d_mean[(size_t)(b0 + g) * C + c] = acc[g];
  }
}

// Weight gradients of both ScaleLong layers in ONE launch: dW2[c][j] += sum_b d_pre2[b][c] h[b][j]  (C x R),
// dW1[j][c] += sum_b d_hpre[b][j] aug[b][c]  (R x (C+1)). A CTA owns a 32-wide c tile of one of the two matrices (all R
// hidden units) and walks the batch in chunks of 32 rows staged through shared memory (coalesced loads, each operand
// element read once per CTA); blockIdx.y splits the batch. Results are added atomically (the flat g_hat buffer is
// zeroed once per step and may already hold earlier micro-batches).
constexpr int kSlWgJ = 8;   // hidden units per thread: R <= 64
__global__ void __launch_bounds__(256)
scalelong_wgrad_kernel(const float* __restrict__ d_pre2, const float* __restrict__ h, const float* __restrict__ d_hpre,
                       const float* __restrict__ aug, float* __restrict__ dw2, float* __restrict__ dw1, int B, int C, int R) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  extern __shared__ float sm[];   // A[32][33] (batch row, c), Bm[32][R] (batch row, j)
  float* As = sm;
  float* Bs = sm + 32 * 33;
  const int C1 = C + 1;
  const int tiles2 = (C + 31) / 32;
  const bool second = (int)blockIdx.x < tiles2;        // this CTA: a tile of dW2, else a tile of dW1
  const int c0 = (second ? (int)blockIdx.x : (int)blockIdx.x - tiles2) * 32;
  const int width = second ? C : C1;                     // row length of the batch-major operand holding the c axis
  const float* Ag = second ? d_pre2 : aug;
  const float* Bg = second ? h : d_hpre;
  const int chunk = ((B + gridDim.y - 1) / gridDim.y + 31) / 32 * 32;
  const int b_begin = blockIdx.y * chunk;
  const int b_end = min(B, b_begin + chunk);
  const int cl = threadIdx.x & 31, jg = threadIdx.x >> 5;
  float acc[kSlWgJ] = {};
  for (int bb = b_begin; bb < b_end; bb += 32) {
    __syncthreads();
#pragma unroll
    for (int r = jg; r < 32; r += 8) {
      const int b = bb + r, c = c0 + cl;
      As[r * 33 + cl] = (b < b_end && c < width) ? Ag[(size_t)b * width + c] : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * R; i += 256) {
      const int r = i / R, j = i - r * R;
      Bs[i] = (bb + r < b_end) ? Bg[(size_t)(bb + r) * R + j] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float av = As[r * 33 + cl];
#pragma unroll
      for (int t = 0; t < kSlWgJ; ++t) {
        const int j = jg + 8 * t;
        if (j < R) acc[t] = fmaf(av, Bs[r * R + j], acc[t]);
      }
    }
  }
  const int c = c0 + cl;
  if (c >= width) return;
#pragma unroll
  for (int t = 0; t < kSlWgJ; ++t) {
    const int j = jg + 8 * t;
    if (j < R) atomicAdd(second ? dw2 + (size_t)c * R + j : dw1 + (size_t)j * C1 + c, acc[t]);
  }
}

// ------------------------------------------------------------------------------------------------
// UncertaintyNet (per batch row); in = hidden = F
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
uncertainty_fwd_kernel(const UncertaintyArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  extern __shared__ float sm[];  // aug[F+1], h[F]
  float* aug = sm;
  float* h = sm + a.F + 1;
  const int b = blockIdx.x;
  for (int f = threadIdx.x; f < a.F; f += blockDim.x) {
    const float v = a.fourier[(size_t)b * a.F + f];
    aug[f] = v;
    a.aug_out[(size_t)b * (a.F + 1) + f] = v;
  }
  if (threadIdx.x == 0) {
    aug[a.F] = 1.0f;
    a.aug_out[(size_t)b * (a.F + 1) + a.F] = 1.0f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < a.F; j += blockDim.x) {
    const float* w = a.w1 + (size_t)j * (a.F + 1);
    float acc = 0.f;
    for (int f = 0; f <= a.F; ++f) acc += w[f] * aug[f];
    a.h_pre[(size_t)b * a.F + j] = acc;
    const float hv = mp_silu_f(acc);
    h[j] = hv;
    a.h_out[(size_t)b * a.F + j] = hv;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float acc = 0.f;
    for (int j = threadIdx.x; j < a.F; j += 32) acc += a.w2[j] * h[j];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) {
      a.u_raw[b] = acc;
      a.u[b] = acc * *a.gain;
    }
  }
}

// g_uraw[b] = g_u[b]*gain ; g_hpre[b,j] = g_uraw*w2[j]*silu'(h_pre)
__global__ void __launch_bounds__(256)
uncertainty_bwd_kernel(const UncertaintyBwdArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const int b = blockIdx.x;
  const float gr = a.g_u[b] * *a.gain;
  if (threadIdx.x == 0) a.g_uraw[b] = gr;
  for (int j = threadIdx.x; j < a.F; j += blockDim.x)
    a.g_hpre[(size_t)b * a.F + j] = gr * a.w2[j] * mp_silu_grad_f(a.h_pre[(size_t)b * a.F + j]);
}

// ------------------------------------------------------------------------------------------------
// conv_in im2col: (B,Ci,H,W) fp32 NCHW -> (B,H,W,64) bf16 patch matrix, k = tap*(Ci+1) + ci
// ------------------------------------------------------------------------------------------------
// CI > 0: image channel count known at compile time (1 MNIST, 3 CIFAR, 4 latents), so k -> (tap, ci) and the tap offsets
// fold into constants — the runtime-division version spent 57 us on a 33 MB tensor (instruction-bound); CI = 0: any count.
template <int CI>
__global__ void __launch_bounds__(256)
conv_in_im2col_kernel(const float* __restrict__ noisy, const float* __restrict__ sigma, int sigma_stride,
                      float sigma_data, __nv_bfloat16* __restrict__ out, int B, int Ci_rt, int H, int W) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (pixel, group of 8 k)
  const long long total = (long long)B * H * W * 8;
  if (idx >= total) return;
  const int Ci = CI > 0 ? CI : Ci_rt;
  const int g = (int)(idx & 7);
  const int pix = (int)(idx >> 3);
  const int hw = H * W;
  const int b = pix / hw, p = pix - b * hw;
  const int h = p / W, w = p - h * W;
  const float s = sigma[b * sigma_stride];
  const float c_in = rsqrtf(sigma_data * sigma_data + s * s);
  const int cin = Ci + 1;
  const float* img = noisy + (size_t)b * Ci * hw;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = g * 8 + i;
    float val = 0.f;
    if (k < 9 * cin) {
      const int tap = k / cin, ci = k - tap * cin;
      const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W)
        val = ci < Ci ? c_in * img[(ci * H + hh) * W + ww] : 1.0f;
    }
    v[i] = val;
  }
  uint4 o;
  o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(out + (size_t)pix * 64 + g * 8) = o;
}

// ------------------------------------------------------------------------------------------------
// conv_out (+ output preconditioning). One warp per pixel.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxCo = 4;

__global__ void __launch_bounds__(256)
conv_out_fwd_generic_kernel(const ConvOutArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long npix = (long long)a.B * a.HW;
  if (warp >= npix) return;
  const int b = (int)(warp / a.HW);
  const int p = (int)(warp - (long long)b * a.HW);
  float acc[kMaxCo] = {0.f, 0.f, 0.f, 0.f};
  const __nv_bfloat16* xr = a.x + warp * a.C;
  for (int c = lane * 2; c < a.C; c += 64) {
    const float2 xv = unpack_bf16(*reinterpret_cast<const uint32_t*>(xr + c));
#pragma unroll
    for (int o = 0; o < kMaxCo; ++o)
      if (o < a.Co) {
        const float2 wv = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.w + (size_t)o * a.C + c));
        acc[o] += xv.x * wv.x + xv.y * wv.y;
      }
  }
#pragma unroll
  for (int o = 0; o < kMaxCo; ++o) acc[o] = warp_sum(acc[o]);
  if (lane < a.Co) {
    float f = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
    const float s = a.sigma[b * a.sigma_stride];
    const float sd = a.sigma_data;
    const float c_skip = sd * sd / (s * s + sd * sd);
    const float c_out = s * sd * rsqrtf(s * s + sd * sd);
    const size_t o = ((size_t)b * a.Co + lane) * a.HW + p;
    if (a.f_raw != nullptr) a.f_raw[o] = f;
    a.D[o] = f * (*a.gain_out) * c_out + a.noisy[o] * c_skip;
  }
}

__global__ void __launch_bounds__(256)
conv_out_bwd_generic_kernel(const ConvOutBwdArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  // Each lane owns groups of 8 consecutive channels (one 16-byte access per pixel); a warp walks pixels.
  //   g_f = g_D * c_out(sigma_b);  d gain_out += g_f * f_raw;  g_x = gain_out * sum_o g_f[o] w[o];  dW[o] += gain_out g_f[o] x
  extern __shared__ float sm[];  // [warps][Co*C] partial dW
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long npix = (long long)a.B * a.HW;
  const long long warp0 = (long long)blockIdx.x * nw + wib;
  const long long nwarps = (long long)gridDim.x * nw;
  const int ngroups = a.C / 8;
  float dgain = 0.f;
  const float gain_out = *a.gain_out;
  const float sd = a.sigma_data;
  for (int i = threadIdx.x; i < nw * a.Co * a.C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  for (int cg0 = 0; cg0 < ngroups; cg0 += 32) {
    const int cg = cg0 + lane;
    const bool on = cg < ngroups;
    float wv[kMaxCo][8], dw[kMaxCo][8];
#pragma unroll
    for (int o = 0; o < kMaxCo; ++o) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { wv[o][i] = 0.f; dw[o][i] = 0.f; }
      if (on && o < a.Co) {
        const uint4 u = *reinterpret_cast<const uint4*>(a.w + (size_t)o * a.C + cg * 8);
        const float2 p0 = unpack_bf16(u.x), p1 = unpack_bf16(u.y), p2 = unpack_bf16(u.z), p3 = unpack_bf16(u.w);
        wv[o][0] = p0.x; wv[o][1] = p0.y; wv[o][2] = p1.x; wv[o][3] = p1.y;
        wv[o][4] = p2.x; wv[o][5] = p2.y; wv[o][6] = p3.x; wv[o][7] = p3.y;
      }
    }
    constexpr int PU = 4;   // pixels in flight per warp (all loads issued before the first use)
    for (long long pix0 = warp0 * PU; pix0 < npix; pix0 += nwarps * PU) {
      uint4 xr[PU];
      float gd[PU][kMaxCo], fr[PU][kMaxCo], cout_s[PU];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const long long pix = pix0 + u;
        cout_s[u] = 0.f;
#pragma unroll
        for (int o = 0; o < kMaxCo; ++o) { gd[u][o] = 0.f; fr[u][o] = 0.f; }
        if (pix < npix) {
          const int b = (int)(pix / a.HW);
          const int p = (int)(pix - (long long)b * a.HW);
          const float s = a.sigma[b * a.sigma_stride];
          cout_s[u] = s * sd * rsqrtf(s * s + sd * sd);
#pragma unroll
          for (int o = 0; o < kMaxCo; ++o)
            if (o < a.Co) {
              const size_t oi = ((size_t)b * a.Co + o) * a.HW + p;
              gd[u][o] = a.g_D[oi];
              if (cg0 == 0 && lane == 0) fr[u][o] = a.f_raw[oi];
            }
          if (on) xr[u] = *reinterpret_cast<const uint4*>(a.x + pix * a.C + cg * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const long long pix = pix0 + u;
        if (pix < npix) {
          float gf[kMaxCo];
#pragma unroll
          for (int o = 0; o < kMaxCo; ++o) {
            const float g1 = gd[u][o] * cout_s[u];
            dgain += g1 * fr[u][o];
            gf[o] = g1 * gain_out;
          }
          if (on) {
            const uint4 uu = xr[u];
            const float2 p0 = unpack_bf16(uu.x), p1 = unpack_bf16(uu.y), p2 = unpack_bf16(uu.z), p3 = unpack_bf16(uu.w);
            const float xv[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
            float gx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float t = 0.f;
#pragma unroll
              for (int o = 0; o < kMaxCo; ++o) {
                t += gf[o] * wv[o][i];
                dw[o][i] += gf[o] * xv[i];
              }
              gx[i] = t;
            }
            uint4 o4;
            o4.x = pack_bf16(gx[0], gx[1]); o4.y = pack_bf16(gx[2], gx[3]); o4.z = pack_bf16(gx[4], gx[5]); o4.w = pack_bf16(gx[6], gx[7]);
            *reinterpret_cast<uint4*>(a.g_x + pix * a.C + cg * 8) = o4;
          }
        }
      }
    }
    if (on) {
      for (int o = 0; o < a.Co; ++o)
#pragma unroll
        for (int i = 0; i < 8; ++i) sm[(size_t)wib * a.Co * a.C + (size_t)o * a.C + cg * 8 + i] = dw[o][i];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < a.Co * a.C; idx += blockDim.x) {
    float t = 0.f;
    for (int w2 = 0; w2 < nw; ++w2) t += sm[(size_t)w2 * a.Co * a.C + idx];
    atomicAdd(a.g_w + idx, t);
  }
  dgain = warp_sum(dgain);
  if (lane == 0 && dgain != 0.f) atomicAdd(a.g_gain_out, dgain);
}


// ---- C <= 256 (every shipped config: 256 / 128 / 192 channels into conv_out): a lane owns 8 consecutive channels of a
// pixel (one 16-byte access), w_hat lives in registers, 8 pixels are in flight per warp and 2 CTAs are resident per SM.
// Sum over lanes of v[j] for every j at once: after the call lane l holds the total of v[l % N] (N a power of two <= 32).
// N-1 exchanges for the transposing part instead of 5 N for N separate butterflies.
template <int N>
__device__ __forceinline__ float warp_sum_transposed(float (&v)[N], int lane) {
#pragma unroll
  for (int off = N / 2; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float send = hi ? v[j] : v[j + off];
      const float keep = hi ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  float t = v[0];
#pragma unroll
  for (int off = N; off < 32; off <<= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
  return t;
}

template <int CO>
__global__ void __launch_bounds__(256, 2)
conv_out_fwd_kernel(const ConvOutArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  constexpr int PU = 8;                                   // pixels in flight per warp
  constexpr int NV = CO == 1 ? 8 : (CO == 2 ? 16 : 32);   // reduction slots: index = o * PU + u
  const int lane = threadIdx.x & 31;
  const long long npix = (long long)a.B * a.HW;
  const long long ngrp = (npix + PU - 1) / PU;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const bool on = lane * 8 < a.C;
  float wv[CO][8];
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (on) u = *reinterpret_cast<const uint4*>(a.w + (size_t)o * a.C + lane * 8);
    const float2 p0 = unpack_bf16(u.x), p1 = unpack_bf16(u.y), p2 = unpack_bf16(u.z), p3 = unpack_bf16(u.w);
    wv[o][0] = p0.x; wv[o][1] = p0.y; wv[o][2] = p1.x; wv[o][3] = p1.y;
    wv[o][4] = p2.x; wv[o][5] = p2.y; wv[o][6] = p3.x; wv[o][7] = p3.y;
  }
  const float gain_out = *a.gain_out;
  const float sd = a.sigma_data;
  for (long long g = warp0; g < ngrp; g += nwarps) {
    const long long pix0 = g * PU;
    uint4 xr[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) {
      xr[u] = make_uint4(0, 0, 0, 0);
      if (on && pix0 + u < npix) xr[u] = *reinterpret_cast<const uint4*>(a.x + (pix0 + u) * a.C + lane * 8);
    }
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.f;
#pragma unroll
    for (int u = 0; u < PU; ++u) {
      const float2 p0 = unpack_bf16(xr[u].x), p1 = unpack_bf16(xr[u].y), p2 = unpack_bf16(xr[u].z), p3 = unpack_bf16(xr[u].w);
      const float xv[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t = fmaf(xv[i], wv[o][i], t);
        v[o * PU + u] = t;
      }
    }
    const float f = warp_sum_transposed<NV>(v, lane);
    const int slot = lane % NV;
    const int o = slot / PU, u = slot % PU;
    const long long pix = pix0 + u;
    if (lane < NV && o < CO && pix < npix) {
      const int b = (int)(pix / a.HW);
      const int p = (int)(pix - (long long)b * a.HW);
      const float s = a.sigma[b * a.sigma_stride];
      const float c_skip = sd * sd / (s * s + sd * sd);
      const float c_out = s * sd * rsqrtf(s * s + sd * sd);
      const size_t oi = ((size_t)b * CO + o) * a.HW + p;
      if (a.f_raw != nullptr) a.f_raw[oi] = f;
      a.D[oi] = f * gain_out * c_out + a.noisy[oi] * c_skip;
    }
  }
}

template <int CO>
__global__ void __launch_bounds__(256, 2)
conv_out_bwd_kernel(const ConvOutBwdArgs a) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  //   g_f = g_D * c_out(sigma_b);  d gain_out += g_f * f_raw;  g_x = gain_out * sum_o g_f[o] w[o];  dW[o] += gain_out g_f[o] x
  // A warp takes 32 consecutive pixels: lane l prepares g_f of pixel l (coalesced reads of g_D / f_raw), then the warp
  // walks the 32 pixels PU at a time with the per-pixel g_f broadcast by shuffles.
  constexpr int PU = CO == 4 ? 4 : 8;
  extern __shared__ float sm[];  // [warps][CO*C] partial dW
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long npix = (long long)a.B * a.HW;
  const long long nchunk = (npix + 31) / 32;
  const long long warp0 = (long long)blockIdx.x * nw + wib;
  const long long nwarps = (long long)gridDim.x * nw;
  const bool on = lane * 8 < a.C;
  const float gain_out = *a.gain_out;
  const float sd = a.sigma_data;
  uint4 wq[CO];          // w_hat stays packed (bf16 pairs) to leave registers for the pixels in flight
  float dw[CO][8];
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    wq[o] = make_uint4(0, 0, 0, 0);
    if (on) wq[o] = *reinterpret_cast<const uint4*>(a.w + (size_t)o * a.C + lane * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) dw[o][i] = 0.f;
  }
  float dgain = 0.f;
  for (long long ch = warp0; ch < nchunk; ch += nwarps) {
    const long long cpix0 = ch * 32;
    float gf[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) gf[o] = 0.f;
    {
      const long long pix = cpix0 + lane;
      if (pix < npix) {
        const int b = (int)(pix / a.HW);
        const int p = (int)(pix - (long long)b * a.HW);
        const float s = a.sigma[b * a.sigma_stride];
        const float cs = s * sd * rsqrtf(s * s + sd * sd);
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          const size_t oi = ((size_t)b * CO + o) * a.HW + p;
          const float g1 = a.g_D[oi] * cs;
          dgain = fmaf(g1, a.f_raw[oi], dgain);
          gf[o] = g1 * gain_out;
        }
      }
    }
#pragma unroll 1
    for (int sub = 0; sub < 32; sub += PU) {
      const long long pix0 = cpix0 + sub;
      if (pix0 >= npix) break;
      uint4 xr[PU];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        xr[u] = make_uint4(0, 0, 0, 0);
        if (on && pix0 + u < npix) xr[u] = *reinterpret_cast<const uint4*>(a.x + (pix0 + u) * a.C + lane * 8);
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        float g[CO];
#pragma unroll
        for (int o = 0; o < CO; ++o) g[o] = __shfl_sync(0xffffffffu, gf[o], sub + u);
        const float2 p0 = unpack_bf16(xr[u].x), p1 = unpack_bf16(xr[u].y), p2 = unpack_bf16(xr[u].z), p3 = unpack_bf16(xr[u].w);
        const float xv[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
        float gx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) gx[i] = 0.f;
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          const float2 w0 = unpack_bf16(wq[o].x), w1 = unpack_bf16(wq[o].y), w2 = unpack_bf16(wq[o].z), w3 = unpack_bf16(wq[o].w);
          const float wv[8] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y, w3.x, w3.y};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            gx[i] = fmaf(g[o], wv[i], gx[i]);
            dw[o][i] = fmaf(g[o], xv[i], dw[o][i]);
          }
        }
        if (on && pix0 + u < npix) {
          uint4 o4;
          o4.x = pack_bf16(gx[0], gx[1]); o4.y = pack_bf16(gx[2], gx[3]); o4.z = pack_bf16(gx[4], gx[5]); o4.w = pack_bf16(gx[6], gx[7]);
          *reinterpret_cast<uint4*>(a.g_x + (pix0 + u) * a.C + lane * 8) = o4;
        }
      }
    }
  }
  if (on) {
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[(size_t)wib * CO * a.C + (size_t)o * a.C + lane * 8 + i] = dw[o][i];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < CO * a.C; idx += blockDim.x) {
    float t = 0.f;
    for (int w2 = 0; w2 < nw; ++w2) t += sm[(size_t)w2 * CO * a.C + idx];
    atomicAdd(a.g_w + idx, t);
  }
  dgain = warp_sum(dgain);
  if (lane == 0 && dgain != 0.f) atomicAdd(a.g_gain_out, dgain);
}

// ------------------------------------------------------------------------------------------------
// weighted MSE (+ uncertainty)
// ------------------------------------------------------------------------------------------------
// weight_b = explicit weight[b] if given, else lambda(sigma_b) (edm.py:212) [* exp(-u_b) (edm.py:216)]
__device__ __forceinline__ float wmse_weight(const float* weight, const float* sigma, const float* u, float sigma_data, int b) {
  if (weight != nullptr) return weight[b];
  const float s = sigma[b];
  float w = (s * s + sigma_data * sigma_data) / ((s * sigma_data) * (s * sigma_data));
  if (u != nullptr) w *= __expf(-u[b]);
  return w;
}

__global__ void __launch_bounds__(256)
wmse_fwd_kernel(const float* __restrict__ D, const float* __restrict__ y, const float* __restrict__ sigma,
                const float* __restrict__ u, const float* __restrict__ weight, float sigma_data, float* __restrict__ mse,
                float* __restrict__ wsum, float* __restrict__ loss, int B, int n) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* d = D + (size_t)b * n;
  const float* t = y + (size_t)b * n;
  float acc = 0.f;
  if ((n & 3) == 0) {
    const float4* d4 = reinterpret_cast<const float4*>(d);
    const float4* t4 = reinterpret_cast<const float4*>(t);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
      const float4 a = d4[i], c = t4[i];
      const float e0 = a.x - c.x, e1 = a.y - c.y, e2 = a.z - c.z, e3 = a.w - c.w;
      acc += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float e = d[i] - t[i];
      acc += e * e;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float m = tot / (float)n;
    mse[b] = m;
    const float w = wmse_weight(weight, sigma, u, sigma_data, b);
    const float contrib = (weight == nullptr && u != nullptr) ? u[b] : 0.f;
    if (wsum != nullptr) atomicAdd(wsum, w * m);               // metric.py:16-18 running state
    atomicAdd(loss, (w * m + contrib) / (float)B);
  }
}

__global__ void __launch_bounds__(256)
wmse_bwd_kernel(const float* __restrict__ D, const float* __restrict__ y, const float* __restrict__ sigma,
                const float* __restrict__ u, const float* __restrict__ weight, const float* __restrict__ mse,
                const float* __restrict__ g_loss, float sigma_data, float* __restrict__ g_D, float* __restrict__ g_u,
                float* __restrict__ g_weight, int B, int n) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const int b = blockIdx.y;
  const float w = wmse_weight(weight, sigma, u, sigma_data, b);
  const float gl = *g_loss;
  const float k = gl * 2.0f * w / ((float)B * (float)n);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * n + i;
    g_D[o] = k * (D[o] - y[o]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (weight == nullptr && u != nullptr && g_u != nullptr) g_u[b] = gl * (1.0f - w * mse[b]) / (float)B;
    if (g_weight != nullptr) g_weight[b] = gl * mse[b] / (float)B;
  }
}

// ------------------------------------------------------------------------------------------------
// Heun stages / initial scaling / diffusion
// ------------------------------------------------------------------------------------------------
__global__ void heun_step_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ D,
                                 const float* __restrict__ d_prev, float* __restrict__ x_out, float* __restrict__ d_out,
                                 const float* __restrict__ ts, int step, int mode, long long n) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t0 = ts[step], t1 = ts[step + 1];
  if (mode == 0) {          // Euler predictor: d = (x0 - D)/t0 ; x1 = x0 + (t1 - t0) d        solvers.py:49-50
    const float x = x0[i];
    const float d = (x - D[i]) / t0;
    d_out[i] = d;
    x_out[i] = x + (t1 - t0) * d;
  } else if (mode == 1) {   // trapezoidal corrector                                           solvers.py:56-57
    const float dp = (x1[i] - D[i]) / t1;
    x_out[i] = x0[i] + (t1 - t0) * (0.5f * d_prev[i] + 0.5f * dp);
  } else {                  // x = x0 * t_0                                                     solvers.py:45
    x_out[i] = x0[i] * t0;
  }
}

__global__ void diffuse_kernel(const float* __restrict__ clean, const float* __restrict__ eps,
                               const float* __restrict__ noise, float P_mean, float P_std, float* __restrict__ noisy,
                               float* __restrict__ sigma, int B, int n) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * n) return;
  const int b = (int)(i / n);
  const float s = __expf(P_mean + eps[b] * P_std);
  if (i % n == 0) sigma[b] = s;
  noisy[i] = clean[i] + noise[i] * s;
}

// ------------------------------------------------------------------------------------------------
// Diffuser.forward with in-kernel normal draws, fused with the Denoiser's input block
// ------------------------------------------------------------------------------------------------
// Reference: epsilon = randn(B); sigma = exp(P_mean + epsilon P_std); noisy = clean + randn_like(clean) sigma
// (src/tinyedm/edm.py:84-93), then x = c_in noisy, ones channel, conv_in's 3x3 patch gather (networks.py:578-587):
// two Philox launches, three elementwise launches and an image-sized round trip per tensor. Here ONE kernel draws both
// normals with a counter-based Philox4x32-10 (key = seed; counter = (index / 4, stream tag, step)), so the value of
// element e is a pure function of (seed, step, e): any thread may regenerate a neighbour's noise instead of waiting for
// it, which is what lets the patch gather ride in the same pass. (seed, step) live on the DEVICE (`state[0..1]`; the host side
// increments the step per call, inside a captured CUDA graph too), so graph replays draw fresh noise and a re-seed reaches them.
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// the e-th standard normal of stream `tag` (0: image noise, 1: epsilon) at (seed, step): Box-Muller on two of the four
// words of Philox call e / 4 (elements 4i..4i+3 share one call)
__device__ __forceinline__ float philox_normal(unsigned long long e, uint32_t tag, unsigned long long seed,
                                               unsigned long long step) {
  const unsigned long long call = e >> 2;
  const uint4 r = philox4x32_10((uint32_t)call, (uint32_t)(call >> 32) | (tag << 30), (uint32_t)step, (uint32_t)(step >> 32),
                                (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t a = (e & 2) ? r.z : r.x, b = (e & 2) ? r.w : r.y;
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);     // (0, 1)
  const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return rad * ((e & 1) ? sn : cs);
}

// the four normals of Philox call `call` (elements 4 call .. 4 call + 3), same values as philox_normal element by element
__device__ __forceinline__ void philox_normal4(unsigned long long call, uint32_t tag, unsigned long long seed,
                                               unsigned long long step, float out[4]) {
  const uint4 r = philox4x32_10((uint32_t)call, (uint32_t)(call >> 32) | (tag << 30), (uint32_t)step, (uint32_t)(step >> 32),
                                (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = ((float)(r.z >> 8) + 0.5f) * (1.0f / 16777216.0f), u4 = ((float)(r.w >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  out[0] = ra * cs; out[1] = ra * sn;
  sincospif(2.0f * u4, &sn, &cs);
  out[2] = rb * cs; out[3] = rb * sn;
}

// One CTA per (image, block of kDiffRows rows). Phase 1 draws the noise of those rows plus a one-row halo above and
// below (every Philox call yields four consecutive pixels of a row; the halo rows are drawn again by the neighbouring
// CTA — the value of an element depends on its index only), writes `noisy` for its own rows and parks c_in * noisy for
// all of them in a zero-bordered shared-memory tile. Phase 2 is conv_in_im2col_kernel's gather out of that tile.
constexpr int kDiffRows = 8;
template <int CI>
__global__ void __launch_bounds__(256)
diffuse_philox_kernel(const float* __restrict__ clean, const long long* __restrict__ state,
                      float P_mean, float P_std, float sigma_data, float* __restrict__ noisy, float* __restrict__ sigma,
                      __nv_bfloat16* __restrict__ xcol, int B, int Ci_rt, int H, int W) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  extern __shared__ float tile[];      // [Ci][kDiffRows + 2][W + 2], border = 0 (the conv's zero padding)
  const int Ci = CI > 0 ? CI : Ci_rt;
  const int row_blocks = (H + kDiffRows - 1) / kDiffRows;
  const int b = blockIdx.x / row_blocks;
  const int h0 = (blockIdx.x - b * row_blocks) * kDiffRows;
  const unsigned long long seed = (unsigned long long)state[0], step = (unsigned long long)state[1];
  const int hw = H * W, TW = W + 2, TR = kDiffRows + 2;
  const float s = __expf(P_mean + philox_normal((unsigned long long)b, 1u, seed, step) * P_std);
  if (h0 == 0 && threadIdx.x == 0) sigma[b] = s;
  const float c_in = rsqrtf(sigma_data * sigma_data + s * s);
  const bool want_tile = xcol != nullptr;
  if (want_tile)
    for (int i = threadIdx.x; i < Ci * TR * TW; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  const size_t img = (size_t)b * Ci * hw;
  const int r_lo = want_tile ? -1 : 0, r_hi = want_tile ? kDiffRows + 1 : kDiffRows;
  if ((W & 3) == 0) {
    const int W4 = W >> 2;
    const int n_items = Ci * (r_hi - r_lo) * W4;
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
      const int w4 = i % W4;
      const int t = i / W4;
      const int r = t % (r_hi - r_lo) + r_lo, ci = t / (r_hi - r_lo);
      const int hh = h0 + r;
      if (hh < 0 || hh >= H) continue;
      const size_t e = img + (size_t)(ci * H + hh) * W + w4 * 4;       // multiple of 4: one Philox call
      float n4[4];
      philox_normal4(e >> 2, 0u, seed, step, n4);
      const float4 c = *reinterpret_cast<const float4*>(clean + e);
      const float4 v = make_float4(c.x + n4[0] * s, c.y + n4[1] * s, c.z + n4[2] * s, c.w + n4[3] * s);
      if (r >= 0 && r < kDiffRows) *reinterpret_cast<float4*>(noisy + e) = v;
      if (want_tile) {
        float* t_row = tile + (ci * TR + r + 1) * TW + 1 + w4 * 4;
        t_row[0] = c_in * v.x; t_row[1] = c_in * v.y; t_row[2] = c_in * v.z; t_row[3] = c_in * v.w;
      }
    }
  } else {
    const int n_items = Ci * (r_hi - r_lo) * W;
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
      const int ww = i % W;
      const int t = i / W;
      const int r = t % (r_hi - r_lo) + r_lo, ci = t / (r_hi - r_lo);
      const int hh = h0 + r;
      if (hh < 0 || hh >= H) continue;
      const size_t e = img + (size_t)(ci * H + hh) * W + ww;
      const float v = clean[e] + philox_normal(e, 0u, seed, step) * s;
      if (r >= 0 && r < kDiffRows) noisy[e] = v;
      if (want_tile) tile[(ci * TR + r + 1) * TW + 1 + ww] = c_in * v;
    }
  }
  if (!want_tile) return;
  __syncthreads();
  const int cin = Ci + 1;
  const int rows_here = min(kDiffRows, H - h0);
  for (int i = threadIdx.x; i < rows_here * W * 8; i += blockDim.x) {
    const int g = i & 7;
    const int pl = i >> 3;
    const int r = pl / W, w = pl - r * W;
    const int h = h0 + r;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < 9 * cin) {
        const int tap = k / cin, ci = k - tap * cin;
        const int dr = tap / 3, dc = tap % 3;              // tile coordinates already include the +1 border shift
        if (ci < Ci) {
          val = tile[(ci * TR + r + dr) * TW + w + dc];
        } else {
          const int hh = h + dr - 1, ww = w + dc - 1;
          val = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? 1.0f : 0.f;
        }
      }
      v[j] = val;
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(xcol + ((size_t)b * hw + (size_t)h * W + w) * 64 + g * 8) = o;
  }
}

__global__ void __launch_bounds__(256)
philox_draws_kernel(const long long* __restrict__ state, float* __restrict__ eps, float* __restrict__ noise, int B, long long n) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long seed = (unsigned long long)state[0], step = (unsigned long long)state[1];
  if (i < B) eps[i] = philox_normal((unsigned long long)i, 1u, seed, step);
  if (i < (long long)B * n) noise[i] = philox_normal((unsigned long long)i, 0u, seed, step);
}

// ------------------------------------------------------------------------------------------------
// Sample post-processing (callbacks.py:152-154): images = clamp(x * std * 2 + mean, 0, 1) -> NHWC -> * 255 -> uint8.
// One thread per pixel; explicit round-to-nearest multiplies / adds in torch's operation order so the truncated byte
// is bit-identical to the reference's (no FMA contraction).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
to_uint8_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ std,
                uint8_t* __restrict__ out, int B, int C, int HW) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW) return;
  const int b = (int)(i / HW), p = (int)(i - (long long)b * HW);
  for (int c = 0; c < C; ++c) {
    float v = __fadd_rn(__fmul_rn(__fmul_rn(x[((long long)b * C + c) * HW + p], std[c]), 2.0f), mean[c]);
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    out[i * C + c] = (uint8_t)__fmul_rn(v, 255.0f);
  }
}

}  // namespace

int to_uint8_images(const float* x, const float* mean, const float* std, uint8_t* out, int B, int C, int HW,
                    cudaStream_t stream) {
  const long long n = (long long)B * HW;
  if (n <= 0) return 0;
  TEDM_CHECK(C >= 1 && C <= 16, "to_uint8_images: unsupported channel count %d", C);
  launch_pdl(to_uint8_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, x, mean, std, out, B, C, HW);
  TEDM_LAUNCH_CHECK();
  return 0;
}

int sgemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
          float alpha, float beta, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid((N + TS - 1) / TS, (M + TS - 1) / TS);
  // long-K, few-tile problems (e.g. g_emb = d_lin (B x 5376) W (5376 x 256)) would run on a handful of CTAs: split K
  int splits = 1;
  const int tiles = grid.x * grid.y;
  // (beta == 1: the partial sums are atomically added onto the existing C, no memset)
  if ((beta == 0.f || beta == 1.f) && ((K >= 1024 && tiles < num_sms()) || (K >= 128 && tiles <= 8))) {
    splits = 2 * num_sms() / tiles;
    const int max_splits = K >= 1024 ? K / 256 : K / 32;   // tiny problems are latency bound: 2 k-steps per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int k_chunk = (K + splits - 1) / splits;
  k_chunk = (k_chunk + TK - 1) / TK * TK;
  splits = (K + k_chunk - 1) / k_chunk;
  grid.z = splits;
  if (splits > 1) {
    TEDM_CHECK(ldc == N, "sgemm: split-K needs a dense C");
    if (beta == 0.f) TEDM_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, stream));
  }
  launch_pdl(sgemm_kernel, grid, 256, 0, stream, A, B, C, M, N, K, lda, ldb, ldc, transA, transB, alpha, beta, k_chunk);
  TEDM_LAUNCH_CHECK();
  return 0;
}

int embedding_forward(const EmbeddingArgs& a, cudaStream_t stream) {
  TEDM_CHECK(a.B > 0 && a.F > 0 && a.E > 0, "embedding: empty problem");
  launch_pdl(embedding_fwd_kernel, a.B, 256, a.F * sizeof(float), stream, a);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int embedding_backward(const EmbeddingBwdArgs& a, cudaStream_t stream) {
  launch_pdl(embedding_bwd_kernel, a.B, 256, 0, stream, a);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int mod_finish_forward(const float* lin, const float* const* gains, const int* col_block, float* m, int B, int N,
                       cudaStream_t stream) {
  const long long n = (long long)B * N;
  launch_pdl(mod_finish_fwd_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, lin, gains, col_block, m, B, N);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int mod_finish_backward(const float* lin, const float* dm, const float* const* gains, const int* blk_start, float* d_lin,
                        float* d_gain, int B, int N, int n_blocks, cudaStream_t stream) {
  dim3 grid(n_blocks, B < 32 ? (B < 1 ? 1 : B) : 32);
  launch_pdl(mod_finish_bwd_kernel, grid, 256, 0, stream, lin, dm, gains, blk_start, d_lin, d_gain, B, N);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int scalelong_forward(const ScaleLongArgs& a, cudaStream_t stream) {
  if (a.B <= 0) return 0;
  const size_t smem = (size_t)kSlG * (a.C + 1 + a.R) * sizeof(float);
  TEDM_CHECK(smem <= 48 * 1024, "scalelong: skip width %d too large", a.C);
  launch_pdl(scalelong_fwd_kernel, (a.B + kSlG - 1) / kSlG, 256, smem, stream, a);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int scalelong_backward(const ScaleLongBwdArgs& a, cudaStream_t stream) {
  if (a.B <= 0) return 0;
  const size_t smem = (size_t)kSlG * (a.C + 9 * a.R) * sizeof(float);
  TEDM_CHECK(smem <= 48 * 1024, "scalelong: skip width %d too large", a.C);
  launch_pdl(scalelong_bwd_kernel, (a.B + kSlG - 1) / kSlG, 256, smem, stream, a);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int scalelong_wgrad(const float* d_pre2, const float* h, const float* d_hpre, const float* aug, float* dw2, float* dw1, int B,
                    int C, int R, cudaStream_t stream) {
  if (B <= 0) return 0;
  TEDM_CHECK(R <= 8 * kSlWgJ, "scalelong: %d hidden units not supported (max %d)", R, 8 * kSlWgJ);
  dim3 grid((C + 31) / 32 + (C + 1 + 31) / 32, B >= 128 ? 4 : (B >= 64 ? 2 : 1));
  launch_pdl(scalelong_wgrad_kernel, grid, 256, (32 * 33 + 32 * R) * sizeof(float), stream, d_pre2, h, d_hpre, aug, dw2, dw1, B, C, R);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int uncertainty_forward(const UncertaintyArgs& a, cudaStream_t stream) {
  launch_pdl(uncertainty_fwd_kernel, a.B, 256, (2 * a.F + 1) * sizeof(float), stream, a);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int uncertainty_backward(const UncertaintyBwdArgs& a, cudaStream_t stream) {
  launch_pdl(uncertainty_bwd_kernel, a.B, 256, 0, stream, a);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int conv_in_im2col(const float* noisy, const float* sigma, int sigma_stride, float sigma_data, __nv_bfloat16* out, int B,
                   int Ci, int H, int W, cudaStream_t stream) {
  TEDM_CHECK(9 * (Ci + 1) <= 64, "conv_in: at most 6 image channels supported (got %d)", Ci);
  const long long total = (long long)B * H * W * 8;
  TEDM_CHECK((long long)B * H * W < (1ll << 28), "conv_in: too many pixels (%d x %d x %d)", B, H, W);
  const unsigned grid = (unsigned)((total + 255) / 256);
  switch (Ci) {
    case 1: launch_pdl(conv_in_im2col_kernel<1>, grid, 256, 0, stream, noisy, sigma, sigma_stride, sigma_data, out, B, Ci, H, W); break;
    case 3: launch_pdl(conv_in_im2col_kernel<3>, grid, 256, 0, stream, noisy, sigma, sigma_stride, sigma_data, out, B, Ci, H, W); break;
    case 4: launch_pdl(conv_in_im2col_kernel<4>, grid, 256, 0, stream, noisy, sigma, sigma_stride, sigma_data, out, B, Ci, H, W); break;
    default: launch_pdl(conv_in_im2col_kernel<0>, grid, 256, 0, stream, noisy, sigma, sigma_stride, sigma_data, out, B, Ci, H, W); break;
  }
  TEDM_LAUNCH_CHECK();
  return 0;
}
int conv_out_forward(const ConvOutArgs& a, cudaStream_t stream) {
  TEDM_CHECK(a.Co >= 1 && a.Co <= kMaxCo && a.C % 64 == 0, "conv_out: unsupported Co=%d C=%d", a.Co, a.C);
  const long long npix = (long long)a.B * a.HW;
  if (a.C <= 256) {
    long long blocks = (npix + 63) / 64;  // 8 pixels per warp and trip
    if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
    if (blocks < 1) blocks = 1;
    switch (a.Co) {
      case 1: launch_pdl(conv_out_fwd_kernel<1>, (unsigned)blocks, 256, 0, stream, a); break;
      case 2: launch_pdl(conv_out_fwd_kernel<2>, (unsigned)blocks, 256, 0, stream, a); break;
      case 3: launch_pdl(conv_out_fwd_kernel<3>, (unsigned)blocks, 256, 0, stream, a); break;
      default: launch_pdl(conv_out_fwd_kernel<4>, (unsigned)blocks, 256, 0, stream, a); break;
    }
  } else {
    launch_pdl(conv_out_fwd_generic_kernel, (unsigned)((npix + 7) / 8), 256, 0, stream, a);
  }
  TEDM_LAUNCH_CHECK();
  return 0;
}
int conv_out_backward(const ConvOutBwdArgs& a, cudaStream_t stream) {
  TEDM_CHECK(a.Co >= 1 && a.Co <= kMaxCo && a.C % 64 == 0, "conv_out_bwd: unsupported Co=%d C=%d", a.Co, a.C);
  const long long npix = (long long)a.B * a.HW;
  long long blocks = (npix + 63) / 64;  // >= 8 pixels per warp
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)8 * a.Co * a.C * sizeof(float);
  TEDM_CHECK(smem <= 48 * 1024, "conv_out_bwd: C too large");
  if (a.C <= 256) {
    blocks = (npix + 255) / 256;          // 32 pixels per warp and trip
    if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
    if (blocks < 1) blocks = 1;
    switch (a.Co) {
      case 1: launch_pdl(conv_out_bwd_kernel<1>, (unsigned)blocks, 256, smem, stream, a); break;
      case 2: launch_pdl(conv_out_bwd_kernel<2>, (unsigned)blocks, 256, smem, stream, a); break;
      case 3: launch_pdl(conv_out_bwd_kernel<3>, (unsigned)blocks, 256, smem, stream, a); break;
      default: launch_pdl(conv_out_bwd_kernel<4>, (unsigned)blocks, 256, smem, stream, a); break;
    }
  } else {
    launch_pdl(conv_out_bwd_generic_kernel, (unsigned)blocks, 256, smem, stream, a);
  }
  TEDM_LAUNCH_CHECK();
  return 0;
}
int wmse_forward(const float* D, const float* y, const float* sigma, const float* u, const float* weight, float sigma_data,
                 float* mse, float* wsum, float* loss, int B, int n, cudaStream_t stream) {
  TEDM_CHECK(weight != nullptr || sigma != nullptr, "wmse: need either explicit weights or sigma");
  TEDM_CHECK(B > 0 && n > 0, "wmse: empty batch");
  TEDM_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  launch_pdl(wmse_fwd_kernel, B, 256, 0, stream, D, y, sigma, u, weight, sigma_data, mse, wsum, loss, B, n);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int wmse_backward(const float* D, const float* y, const float* sigma, const float* u, const float* weight, const float* mse,
                  const float* g_loss, float sigma_data, float* g_D, float* g_u, float* g_weight, int B, int n,
                  cudaStream_t stream) {
  TEDM_CHECK(weight != nullptr || sigma != nullptr, "wmse: need either explicit weights or sigma");
  dim3 grid((n + 1023) / 1024 < 1 ? 1 : (n + 1023) / 1024, B);
  launch_pdl(wmse_bwd_kernel, grid, 256, 0, stream, D, y, sigma, u, weight, mse, g_loss, sigma_data, g_D, g_u, g_weight, B, n);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int heun_step(const float* x0, const float* x1, const float* D, const float* d_prev, float* x_out, float* d_out,
              const float* ts, int step, int mode, long long n, cudaStream_t stream) {
  launch_pdl(heun_step_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, x0, x1, D, d_prev, x_out, d_out, ts, step, mode, n);
  TEDM_LAUNCH_CHECK();
  return 0;
}
int diffuse(const float* clean, const float* eps, const float* noise, float P_mean, float P_std, float* noisy,
            float* sigma, int B, int n, cudaStream_t stream) {
  const long long tot = (long long)B * n;
  launch_pdl(diffuse_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, clean, eps, noise, P_mean, P_std, noisy, sigma, B, n);
  TEDM_LAUNCH_CHECK();
  return 0;
}

int diffuse_philox(const float* clean, const long long* state, float P_mean, float P_std,
                   float sigma_data, float* noisy, float* sigma, __nv_bfloat16* xcol, int B, int Ci, int H, int W,
                   cudaStream_t stream) {
  TEDM_CHECK(xcol == nullptr || 9 * (Ci + 1) <= 64, "diffuse: the fused patch gather supports at most 6 image channels (got %d)", Ci);
  TEDM_CHECK((long long)B * H * W < (1ll << 28), "diffuse: too many pixels (%d x %d x %d)", B, H, W);
  if (B <= 0) return 0;
  const size_t smem = xcol != nullptr ? (size_t)Ci * (kDiffRows + 2) * (W + 2) * sizeof(float) : 0;
  TEDM_CHECK(smem <= 48 * 1024, "diffuse: image rows of %d pixels x %d channels do not fit the staging tile", W, Ci);
  const unsigned grid = (unsigned)(B * ((H + kDiffRows - 1) / kDiffRows));
  switch (Ci) {
    case 1: launch_pdl(diffuse_philox_kernel<1>, grid, 256, smem, stream, clean, state, P_mean, P_std, sigma_data, noisy, sigma, xcol, B, Ci, H, W); break;
    case 3: launch_pdl(diffuse_philox_kernel<3>, grid, 256, smem, stream, clean, state, P_mean, P_std, sigma_data, noisy, sigma, xcol, B, Ci, H, W); break;
    case 4: launch_pdl(diffuse_philox_kernel<4>, grid, 256, smem, stream, clean, state, P_mean, P_std, sigma_data, noisy, sigma, xcol, B, Ci, H, W); break;
    default: launch_pdl(diffuse_philox_kernel<0>, grid, 256, smem, stream, clean, state, P_mean, P_std, sigma_data, noisy, sigma, xcol, B, Ci, H, W); break;
  }
  TEDM_LAUNCH_CHECK();
  return 0;
}
int philox_normal_draws(const long long* state, float* eps, float* noise, int B, long long n, cudaStream_t stream) {
  const long long tot = (long long)B * n > B ? (long long)B * n : B;
  launch_pdl(philox_draws_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, state, eps, noise, B, n);
  TEDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace tedm
