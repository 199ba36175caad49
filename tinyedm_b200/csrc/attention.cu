// Cosine-normalised self-attention (reference: CosineAttention.forward, src/tinyedm/networks.py:191-207).
//
//   qkv (B,S,3C) bf16, channel = head*3*hd + d*3 + {q,k,v}   -> pixel_norm over hd of q, k AND v (:195)
//   y = softmax(q k^T / sqrt(hd)) v                           (:201)  -> (B,S,C), channel = head*hd + d (:202)
//
// S <= 256 and hd <= 128 here, so one CTA keeps a whole head's K and V in shared memory (flash-style,
// the S x S matrix never reaches HBM). Matrix products run on tensor cores through warp-level MMA
// (HMMA); round 1 keeps this 1.3%-of-FLOPs kernel on the legacy tensor path, see DESIGN.md.
#include <mma.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

using namespace nvcuda;
constexpr float kEps = 1e-4f;
constexpr int kAttnThreads = 256;
constexpr int kWarps = kAttnThreads / 32;

// ------------------------------------------------------------------------------------------------
// qkv de-interleave + pixel norm (forward / backward)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
qkv_norm_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B, int S, int heads,
                    int hd) {
  // out: [3][B][heads][S][hd]
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)B * S * heads;
  if (warp >= total) return;
  const int head = (int)(warp % heads);
  const int s = (int)((warp / heads) % S);
  const int b = (int)(warp / ((long long)heads * S));
  const int C3 = 3 * hd * heads;
  const __nv_bfloat16* src = qkv + ((long long)b * S + s) * C3 + head * 3 * hd;
  const int n = 3 * hd;
  float ss[3] = {0.f, 0.f, 0.f};
  for (int i = lane; i < n; i += 32) {
    const float v = __bfloat162float(src[i]);
    const int j = i % 3;
    ss[0] += j == 0 ? v * v : 0.f;
    ss[1] += j == 1 ? v * v : 0.f;
    ss[2] += j == 2 ? v * v : 0.f;
  }
  float inv[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) inv[j] = 1.0f / (kEps + sqrtf(warp_sum(ss[j]) / (float)hd));
  const long long plane = (long long)B * heads * S * hd;
  const long long row = (((long long)b * heads + head) * S + s) * hd;
  for (int i = lane; i < n; i += 32) {
    const int d = i / 3, j = i - d * 3;
    const float v = __bfloat162float(src[i]);
    out[j * plane + row + d] = __float2bfloat16_rn(v * (j == 0 ? inv[0] : (j == 1 ? inv[1] : inv[2])));
  }
}

__global__ void __launch_bounds__(256)
qkv_norm_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ g_out,
                    __nv_bfloat16* __restrict__ g_qkv, int B, int S, int heads, int hd) {
  // g_out: [3][B][heads][S][hd] gradient w.r.t. the normalised q,k,v; g_qkv: (B,S,3C)
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)B * S * heads;
  if (warp >= total) return;
  const int head = (int)(warp % heads);
  const int s = (int)((warp / heads) % S);
  const int b = (int)(warp / ((long long)heads * S));
  const int C3 = 3 * hd * heads;
  const long long base = ((long long)b * S + s) * C3 + head * 3 * hd;
  const long long plane = (long long)B * heads * S * hd;
  const long long row = (((long long)b * heads + head) * S + s) * hd;
  const int n = 3 * hd;
  float ss[3] = {0.f, 0.f, 0.f}, dot[3] = {0.f, 0.f, 0.f};
  for (int i = lane; i < n; i += 32) {
    const int d = i / 3, j = i - d * 3;
    const float v = __bfloat162float(qkv[base + i]);
    const float g = __bfloat162float(g_out[j * plane + row + d]);
    ss[0] += j == 0 ? v * v : 0.f;  dot[0] += j == 0 ? v * g : 0.f;
    ss[1] += j == 1 ? v * v : 0.f;  dot[1] += j == 1 ? v * g : 0.f;
    ss[2] += j == 2 ? v * v : 0.f;  dot[2] += j == 2 ? v * g : 0.f;
  }
  float inv_n[3], k[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float r = sqrtf(warp_sum(ss[j]) / (float)hd);
    const float dt = warp_sum(dot[j]);
    const float nn = kEps + r;
    inv_n[j] = 1.0f / nn;
    // g_u = g/n - u * (sum g*u) / (n^2 * r * hd)
    k[j] = dt / (nn * nn * fmaxf(r, 1e-20f) * (float)hd);
  }
  for (int i = lane; i < n; i += 32) {
    const int d = i / 3, j = i - d * 3;
    const float v = __bfloat162float(qkv[base + i]);
    const float g = __bfloat162float(g_out[j * plane + row + d]);
    const float in = j == 0 ? inv_n[0] : (j == 1 ? inv_n[1] : inv_n[2]);
    const float kk = j == 0 ? k[0] : (j == 1 ? k[1] : k[2]);
    g_qkv[base + i] = __float2bfloat16_rn(g * in - v * kk);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-level shared-memory GEMM on warp MMA:  C[M x N] (fp32, row-major ldc) = A * B
//   A: M x K, row-major (lda) if !A_COL else stored K x M row-major (i.e. A^T given)
//   B: K x N, row-major (ldb) if !B_COL else stored N x K row-major (i.e. B^T given)
// ------------------------------------------------------------------------------------------------
template <bool A_COL, bool B_COL>
__device__ __forceinline__ void smem_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* Bm, int ldb, float* C,
                                          int ldc, int M, int N, int K) {
  const int warp = threadIdx.x >> 5;
  const int mt = M / 16, nt = N / 16;
  for (int t = warp; t < mt * nt; t += kWarps) {
    const int mi = t / nt, ni = t - mi * nt;
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
    wmma::fill_fragment(acc, 0.f);
    for (int k = 0; k < K; k += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16,
                     typename std::conditional<A_COL, wmma::col_major, wmma::row_major>::type> fa;
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16,
                     typename std::conditional<B_COL, wmma::col_major, wmma::row_major>::type> fb;
      const __nv_bfloat16* pa = A_COL ? A + (size_t)k * lda + mi * 16 : A + (size_t)mi * 16 * lda + k;
      const __nv_bfloat16* pb = B_COL ? Bm + (size_t)ni * 16 * ldb + k : Bm + (size_t)k * ldb + ni * 16;
      wmma::load_matrix_sync(fa, pa, lda);
      wmma::load_matrix_sync(fb, pb, ldb);
      wmma::mma_sync(acc, fa, fb, acc);
    }
    wmma::store_matrix_sync(C + (size_t)mi * 16 * ldc + ni * 16, acc, ldc, wmma::mem_row_major);
  }
}

// loads `rows` rows of `hd` bf16 (global row stride g_ld) into smem (row stride s_ld); rows >= valid are zeroed
__device__ __forceinline__ void load_rows(__nv_bfloat16* dst, int s_ld, const __nv_bfloat16* src, long long g_ld, int rows,
                                          int valid, int hd) {
  const int vec_per_row = hd / 8;
  for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, v = i - r * vec_per_row;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r < valid) u = *reinterpret_cast<const uint4*>(src + (long long)r * g_ld + v * 8);
    *reinterpret_cast<uint4*>(dst + (size_t)r * s_ld + v * 8) = u;
  }
}

struct AttnSmem {
  int Sp, ldh, lds, ldp;
};

// ------------------------------------------------------------------------------------------------
// forward: grid (q blocks of QB rows, B*heads)
// ------------------------------------------------------------------------------------------------
constexpr int QB_FWD = 64;

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkvn, __nv_bfloat16* __restrict__ y, float* __restrict__ lse, int B,
                int S, int heads, int hd, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Sp = (S + 15) / 16 * 16;
  const int ldh = hd + 8, lds = Sp + 4, ldp = Sp + 8;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + (size_t)Sp * ldh;
  __nv_bfloat16* Qs = Vs + (size_t)Sp * ldh;
  __nv_bfloat16* Ps = Qs + (size_t)QB_FWD * ldh;
  float* Sc = reinterpret_cast<float*>(Ps + (size_t)QB_FWD * ldp);

  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * QB_FWD;
  const long long plane = (long long)B * heads * S * hd;
  const __nv_bfloat16* q = qkvn + (long long)bh * S * hd;
  const __nv_bfloat16* k = q + plane;
  const __nv_bfloat16* v = k + plane;
  int qvalid = S - q0;
  if (qvalid > QB_FWD) qvalid = QB_FWD;

  load_rows(Ks, ldh, k, hd, Sp, S, hd);
  load_rows(Vs, ldh, v, hd, Sp, S, hd);
  load_rows(Qs, ldh, q + (long long)q0 * hd, hd, QB_FWD, qvalid, hd);
  __syncthreads();
  smem_gemm<false, true>(Qs, ldh, Ks, ldh, Sc, lds, QB_FWD, Sp, hd);
  __syncthreads();
  // softmax, one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < QB_FWD; r += kWarps) {
    float* row = Sc + (size_t)r * lds;
    float mx = -INFINITY;
    for (int c = lane; c < S; c += 32) mx = fmaxf(mx, row[c] * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < S; c += 32) {
      const float e = __expf(row[c] * scale - mx);
      row[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int c = lane; c < Sp; c += 32) Ps[(size_t)r * ldp + c] = __float2bfloat16_rn(c < S ? row[c] * inv : 0.f);
    if (lane == 0 && r < qvalid && lse != nullptr) lse[(long long)bh * S + q0 + r] = mx + __logf(sum);
  }
  __syncthreads();
  // O = P V  -> reuse Sc as [QB][hd] fp32 (ld = hd + 4)
  const int ldo = hd + 4;
  smem_gemm<false, false>(Ps, ldp, Vs, ldh, Sc, ldo, QB_FWD, hd, Sp);
  __syncthreads();
  const int b = bh / heads, head = bh - b * heads;
  const int C = heads * hd;
  for (int i = threadIdx.x; i < qvalid * (hd / 2); i += blockDim.x) {
    const int r = i / (hd / 2), d = (i - r * (hd / 2)) * 2;
    const uint32_t pk = pack_bf16(Sc[(size_t)r * ldo + d], Sc[(size_t)r * ldo + d + 1]);
    *reinterpret_cast<uint32_t*>(y + ((long long)b * S + q0 + r) * C + head * hd + d) = pk;
  }
}

// ------------------------------------------------------------------------------------------------
// backward pass 1 (per q block): delta, dQ
// ------------------------------------------------------------------------------------------------
constexpr int QB_BWD = 32;

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkvn, const __nv_bfloat16* __restrict__ y,
                   const __nv_bfloat16* __restrict__ g_y, const float* __restrict__ lse, float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ g_qkvn, int B, int S, int heads, int hd, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Sp = (S + 15) / 16 * 16;
  const int ldh = hd + 8, lds = Sp + 4, ldp = Sp + 8;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + (size_t)Sp * ldh;
  __nv_bfloat16* Qs = Vs + (size_t)Sp * ldh;
  __nv_bfloat16* dOs = Qs + (size_t)QB_BWD * ldh;
  __nv_bfloat16* dSs = dOs + (size_t)QB_BWD * ldh;
  float* Sc = reinterpret_cast<float*>(dSs + (size_t)QB_BWD * ldp);
  float* dP = Sc + (size_t)QB_BWD * lds;
  float* dl = dP + (size_t)QB_BWD * lds;  // [QB]

  const int bh = blockIdx.y;
  const int b = bh / heads, head = bh - b * heads;
  const int C = heads * hd;
  const int q0 = blockIdx.x * QB_BWD;
  const long long plane = (long long)B * heads * S * hd;
  const __nv_bfloat16* q = qkvn + (long long)bh * S * hd;
  const __nv_bfloat16* k = q + plane;
  const __nv_bfloat16* v = k + plane;
  int qvalid = S - q0;
  if (qvalid > QB_BWD) qvalid = QB_BWD;

  load_rows(Ks, ldh, k, hd, Sp, S, hd);
  load_rows(Vs, ldh, v, hd, Sp, S, hd);
  load_rows(Qs, ldh, q + (long long)q0 * hd, hd, QB_BWD, qvalid, hd);
  load_rows(dOs, ldh, g_y + ((long long)b * S + q0) * C + head * hd, C, QB_BWD, qvalid, hd);
  // delta_r = sum_d dO * O
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < QB_BWD; r += kWarps) {
    float acc = 0.f;
    if (r < qvalid) {
      const long long o = ((long long)b * S + q0 + r) * C + head * hd;
      for (int d = lane; d < hd; d += 32) acc += __bfloat162float(g_y[o + d]) * __bfloat162float(y[o + d]);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      dl[r] = acc;
      if (r < qvalid) delta[(long long)bh * S + q0 + r] = acc;
    }
  }
  __syncthreads();
  smem_gemm<false, true>(Qs, ldh, Ks, ldh, Sc, lds, QB_BWD, Sp, hd);
  smem_gemm<false, true>(dOs, ldh, Vs, ldh, dP, lds, QB_BWD, Sp, hd);
  __syncthreads();
  for (int i = threadIdx.x; i < QB_BWD * Sp; i += blockDim.x) {
    const int r = i / Sp, c = i - r * Sp;
    float ds = 0.f;
    if (r < qvalid && c < S) {
      const float p = __expf(Sc[(size_t)r * lds + c] * scale - lse[(long long)bh * S + q0 + r]);
      ds = p * (dP[(size_t)r * lds + c] - dl[r]) * scale;
    }
    dSs[(size_t)r * ldp + c] = __float2bfloat16_rn(ds);
  }
  __syncthreads();
  const int ldo = hd + 4;
  smem_gemm<false, false>(dSs, ldp, Ks, ldh, Sc, ldo, QB_BWD, hd, Sp);
  __syncthreads();
  __nv_bfloat16* gq = g_qkvn + (long long)bh * S * hd + (long long)q0 * hd;
  for (int i = threadIdx.x; i < qvalid * (hd / 2); i += blockDim.x) {
    const int r = i / (hd / 2), d = (i - r * (hd / 2)) * 2;
    *reinterpret_cast<uint32_t*>(gq + (long long)r * hd + d) = pack_bf16(Sc[(size_t)r * ldo + d], Sc[(size_t)r * ldo + d + 1]);
  }
}

// ------------------------------------------------------------------------------------------------
// backward pass 2 (per key block): dK, dV
// ------------------------------------------------------------------------------------------------
constexpr int KB_BWD = 32;

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkvn, const __nv_bfloat16* __restrict__ g_y,
                    const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ g_qkvn,
                    int B, int S, int heads, int hd, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int Sp = (S + 15) / 16 * 16;
  const int ldh = hd + 8, lds = Sp + 4, ldp = Sp + 8;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* dOs = Qs + (size_t)Sp * ldh;
  __nv_bfloat16* Kb = dOs + (size_t)Sp * ldh;
  __nv_bfloat16* Vb = Kb + (size_t)KB_BWD * ldh;
  __nv_bfloat16* PT = Vb + (size_t)KB_BWD * ldh;
  __nv_bfloat16* dST = PT + (size_t)KB_BWD * ldp;
  float* ScT = reinterpret_cast<float*>(dST + (size_t)KB_BWD * ldp);
  float* dPT = ScT + (size_t)KB_BWD * lds;

  const int bh = blockIdx.y;
  const int b = bh / heads, head = bh - b * heads;
  const int C = heads * hd;
  const int k0 = blockIdx.x * KB_BWD;
  const long long plane = (long long)B * heads * S * hd;
  const __nv_bfloat16* q = qkvn + (long long)bh * S * hd;
  const __nv_bfloat16* k = q + plane;
  const __nv_bfloat16* v = k + plane;
  int kvalid = S - k0;
  if (kvalid > KB_BWD) kvalid = KB_BWD;

  load_rows(Qs, ldh, q, hd, Sp, S, hd);
  load_rows(dOs, ldh, g_y + (long long)b * S * C + head * hd, C, Sp, S, hd);
  load_rows(Kb, ldh, k + (long long)k0 * hd, hd, KB_BWD, kvalid, hd);
  load_rows(Vb, ldh, v + (long long)k0 * hd, hd, KB_BWD, kvalid, hd);
  __syncthreads();
  smem_gemm<false, true>(Kb, ldh, Qs, ldh, ScT, lds, KB_BWD, Sp, hd);   // [key][query]
  smem_gemm<false, true>(Vb, ldh, dOs, ldh, dPT, lds, KB_BWD, Sp, hd);
  __syncthreads();
  for (int i = threadIdx.x; i < KB_BWD * Sp; i += blockDim.x) {
    const int r = i / Sp, c = i - r * Sp;  // r: key, c: query
    float p = 0.f, ds = 0.f;
    if (r < kvalid && c < S) {
      p = __expf(ScT[(size_t)r * lds + c] * scale - lse[(long long)bh * S + c]);
      ds = p * (dPT[(size_t)r * lds + c] - delta[(long long)bh * S + c]) * scale;
    }
    PT[(size_t)r * ldp + c] = __float2bfloat16_rn(p);
    dST[(size_t)r * ldp + c] = __float2bfloat16_rn(ds);
  }
  __syncthreads();
  const int ldo = hd + 4;
  float* dV = ScT;  // reuse
  float* dK = dPT;
  smem_gemm<false, false>(PT, ldp, dOs, ldh, dV, ldo, KB_BWD, hd, Sp);
  smem_gemm<false, false>(dST, ldp, Qs, ldh, dK, ldo, KB_BWD, hd, Sp);
  __syncthreads();
  __nv_bfloat16* gk = g_qkvn + plane + (long long)bh * S * hd + (long long)k0 * hd;
  __nv_bfloat16* gv = gk + plane;
  for (int i = threadIdx.x; i < kvalid * (hd / 2); i += blockDim.x) {
    const int r = i / (hd / 2), d = (i - r * (hd / 2)) * 2;
    *reinterpret_cast<uint32_t*>(gk + (long long)r * hd + d) = pack_bf16(dK[(size_t)r * ldo + d], dK[(size_t)r * ldo + d + 1]);
    *reinterpret_cast<uint32_t*>(gv + (long long)r * hd + d) = pack_bf16(dV[(size_t)r * ldo + d], dV[(size_t)r * ldo + d + 1]);
  }
}

size_t fwd_smem(int S, int hd) {
  const int Sp = (S + 15) / 16 * 16;
  const size_t ldh = hd + 8, lds = Sp + 4, ldp = Sp + 8;
  size_t sc = (size_t)QB_FWD * lds * 4;
  size_t so = (size_t)QB_FWD * (hd + 4) * 4;
  return (2 * Sp + QB_FWD) * ldh * 2 + QB_FWD * ldp * 2 + (sc > so ? sc : so);
}
size_t bwd_dq_smem(int S, int hd) {
  const int Sp = (S + 15) / 16 * 16;
  const size_t ldh = hd + 8, lds = Sp + 4, ldp = Sp + 8;
  size_t sc = (size_t)QB_BWD * lds * 4;
  size_t so = (size_t)QB_BWD * (hd + 4) * 4;
  return (2 * Sp + 2 * QB_BWD) * ldh * 2 + QB_BWD * ldp * 2 + 2 * (sc > so ? sc : so) + QB_BWD * 4 + 64;
}
size_t bwd_dkv_smem(int S, int hd) {
  const int Sp = (S + 15) / 16 * 16;
  const size_t ldh = hd + 8, lds = Sp + 4, ldp = Sp + 8;
  size_t sc = (size_t)KB_BWD * lds * 4;
  size_t so = (size_t)KB_BWD * (hd + 4) * 4;
  return (2 * Sp + 2 * KB_BWD) * ldh * 2 + 2 * KB_BWD * ldp * 2 + 2 * (sc > so ? sc : so);
}

int check_dims(int S, int hd, int heads) {
  TEDM_CHECK(hd % 16 == 0 && hd >= 16, "attention: head_dim must be a multiple of 16 (got %d)", hd);
  TEDM_CHECK(S >= 1 && heads >= 1, "attention: empty problem");
  const size_t lim = 227 * 1024;
  TEDM_CHECK(fwd_smem(S, hd) <= lim && bwd_dq_smem(S, hd) <= lim && bwd_dkv_smem(S, hd) <= lim,
             "attention: S=%d, head_dim=%d does not fit the shared-memory resident kernel (round-1 limit)", S, hd);
  return 0;
}

}  // namespace

int attention_forward(const __nv_bfloat16* qkv, __nv_bfloat16* qkvn, __nv_bfloat16* y, float* lse, int B, int S,
                      int heads, int hd, cudaStream_t stream) {
  if (check_dims(S, hd, heads) != 0) return -1;
  const long long warps = (long long)B * S * heads;
  qkv_norm_fwd_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(qkv, qkvn, B, S, heads, hd);
  TEDM_LAUNCH_CHECK();
  const size_t smem = fwd_smem(S, hd);
  TEDM_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((S + QB_FWD - 1) / QB_FWD, B * heads);
  attn_fwd_kernel<<<grid, kAttnThreads, smem, stream>>>(qkvn, y, lse, B, S, heads, hd, 1.0f / sqrtf((float)hd));
  TEDM_LAUNCH_CHECK();
  return 0;
}

int attention_backward(const __nv_bfloat16* qkv, const __nv_bfloat16* qkvn, const __nv_bfloat16* y,
                       const __nv_bfloat16* g_y, const float* lse, float* delta, __nv_bfloat16* g_qkvn,
                       __nv_bfloat16* g_qkv, int B, int S, int heads, int hd, cudaStream_t stream) {
  if (check_dims(S, hd, heads) != 0) return -1;
  const float scale = 1.0f / sqrtf((float)hd);
  {
    const size_t smem = bwd_dq_smem(S, hd);
    TEDM_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((S + QB_BWD - 1) / QB_BWD, B * heads);
    attn_bwd_dq_kernel<<<grid, kAttnThreads, smem, stream>>>(qkvn, y, g_y, lse, delta, g_qkvn, B, S, heads, hd, scale);
    TEDM_LAUNCH_CHECK();
  }
  {
    const size_t smem = bwd_dkv_smem(S, hd);
    TEDM_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((S + KB_BWD - 1) / KB_BWD, B * heads);
    attn_bwd_dkv_kernel<<<grid, kAttnThreads, smem, stream>>>(qkvn, g_y, lse, delta, g_qkvn, B, S, heads, hd, scale);
    TEDM_LAUNCH_CHECK();
  }
  const long long warps = (long long)B * S * heads;
  qkv_norm_bwd_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(qkv, g_qkvn, g_qkv, B, S, heads, hd);
  TEDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace tedm
