// Cosine-normalised self-attention (reference: CosineAttention.forward, src/tinyedm/networks.py:191-207).
//
//   qkv (B,S,3C) bf16                                         -> pixel_norm over hd of q, k AND v (:195)
//   y = softmax(q k^T / sqrt(hd)) v                           (:201)  -> (B,S,C), channel = head*hd + d (:202)
// The reference's qkv channel order is head*3*hd + d*3 + {q,k,v} (:194); here the weight bank permutes the rows of
// the qkv conv weight so the convolution emits {q,k,v}*C + head*hd + d, i.e. contiguous hd-vectors per head.
// The q/k/v pixel norm and its backward are fused into these kernels (rows are normalised in shared memory right
// after the load; the adjoint is applied to dQ/dK/dV rows in registers before the store), so no normalised copy of
// q,k,v ever exists in HBM.
//
// Flash-style, register resident: S <= 256 and hd <= 128 here, so a CTA keeps the K and V of one (image, head)
// in shared memory and each warp owns 16 query rows whose score / probability tiles live only in registers
// (online softmax over 64-key chunks; the S x S matrix exists neither in HBM nor in shared memory).
// Matrix products are warp-level tensor-core MMAs (mma.sync.m16n8k16 bf16 -> fp32) fed by ldmatrix; the
// accumulator -> A-operand hand-off between the two GEMMs of each chain is register-local.
// Backward = two kernels without atomics: per query block (delta, dQ) and per key block (dK, dV); both
// recompute P from the saved log-sum-exp.
//
// This 1.3%-of-FLOPs op stays on the warp-level tensor path in round 1 (problem per head is 256x256x64: too
// small to amortise a tcgen05/TMEM pipeline without batching heads per CTA); see DESIGN.md.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr float kEps = 1e-4f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kThreads = 128;   // 4 warps x 16 rows = 64-row blocks
constexpr int kBlk = 64;

// ------------------------------------------------------------------------------------------------
// warp-level MMA plumbing
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}
// c (16x8 fp32) += a (16x16 bf16, row) * b (16x8 bf16, col)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// `rows` rows of HD bf16 (global row stride g_ld elements) -> smem rows of LD elements; rows >= valid are zero-filled.
template <int HD>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long g_ld, int rows, int valid) {
  constexpr int LD = HD + 8, VPR = HD / 8;
  for (int i = threadIdx.x; i < rows * VPR; i += kThreads) {
    const int r = i / VPR, v = i - r * VPR;
    const int rs = r < valid ? r : 0;
    cp_async16(dst + r * LD + v * 8, src + (long long)rs * g_ld + v * 8, r < valid ? 16 : 0);
  }
}

// Wide heads (> 64) do not keep their A fragments in registers: they are re-read from shared memory per k-step.
template <int HD>
struct KeepFrags {
  static constexpr bool value = HD <= 64;
};

// A-operand fragments (16 rows x HD) of this warp's rows from a [rows][LD] smem tile.
template <int HD>
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[HD / 16][4], const __nv_bfloat16* tile, int row0, int lane) {
  constexpr int LD = HD + 8;
  if constexpr (KeepFrags<HD>::value) {
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk)
      ldsm_x4(f[kk], tile + (row0 + (lane & 15)) * LD + kk * 16 + 8 * (lane >> 4));
  }
}

// c[nt] (16 x 8 each, NT n-tiles starting at smem row n0) += A(16 x HD) * Bt^T where Bt is stored [n][HD] (row = n index).
// A = this warp's 16 rows starting at a_row0 of a_tile: register fragments `a` when kept, else loaded per k-step.
template <int HD, int NT>
__device__ __forceinline__ void gemm_a_bt(float (&c)[NT][4], const uint32_t (&a)[HD / 16][4], const __nv_bfloat16* a_tile,
                                          int a_row0, const __nv_bfloat16* bt, int n0, int lane) {
  constexpr int LD = HD + 8;
  constexpr bool KEEP = KeepFrags<HD>::value;
#pragma unroll
  for (int k2 = 0; k2 < HD / 32; ++k2) {
    uint32_t a0[4], a1[4];
    if constexpr (KEEP) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { a0[i] = a[2 * k2][i]; a1[i] = a[2 * k2 + 1][i]; }
    } else {
      ldsm_x4(a0, a_tile + (a_row0 + (lane & 15)) * LD + (2 * k2) * 16 + 8 * (lane >> 4));
      ldsm_x4(a1, a_tile + (a_row0 + (lane & 15)) * LD + (2 * k2 + 1) * 16 + 8 * (lane >> 4));
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t b[4];
      ldsm_x4(b, bt + (n0 + nt * 8 + (lane & 7)) * LD + k2 * 32 + 8 * (lane >> 3));
      mma16816(c[nt], a0, b[0], b[1]);
      mma16816(c[nt], a1, b[2], b[3]);
    }
  }
  if constexpr (HD % 32 != 0) {   // one trailing 16-wide k-step (head_dim 144)
    constexpr int ks = HD / 16 - 1;
    uint32_t a0[4];
    if constexpr (KEEP) {
#pragma unroll
      for (int i = 0; i < 4; ++i) a0[i] = a[ks][i];
    } else {
      ldsm_x4(a0, a_tile + (a_row0 + (lane & 15)) * LD + ks * 16 + 8 * (lane >> 4));
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t b[2];
      ldsm_x2(b, bt + (n0 + nt * 8 + (lane & 7)) * LD + ks * 16 + 8 * ((lane >> 3) & 1));
      mma16816(c[nt], a0, b[0], b[1]);
    }
  }
}

// o[nd] (16 x 8 each over HD columns) += P(16 x 16*KS, fragments p[ks]) * Bm where Bm is stored [k][HD] starting at row k0.
template <int HD, int KS>
__device__ __forceinline__ void gemm_p_b(float (&o)[HD / 8][4], const uint32_t (&p)[KS][4], const __nv_bfloat16* bm, int k0,
                                         int lane) {
  constexpr int LD = HD + 8;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int n2 = 0; n2 < HD / 16; ++n2) {
      uint32_t b[4];
      ldsm_x4_t(b, bm + (k0 + ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + n2 * 16 + 8 * (lane >> 4));
      mma16816(o[2 * n2], p[ks], b[0], b[1]);
      mma16816(o[2 * n2 + 1], p[ks], b[2], b[3]);
    }
  }
}

// accumulator tiles (16 x 8*NT fp32) -> A fragments (16 x 16*(NT/2) bf16); register-local
template <int NT>
__device__ __forceinline__ void acc_to_a(uint32_t (&p)[NT / 2][4], const float (&c)[NT][4]) {
#pragma unroll
  for (int ks = 0; ks < NT / 2; ++ks) {
    p[ks][0] = pack_bf16(c[2 * ks][0], c[2 * ks][1]);
    p[ks][1] = pack_bf16(c[2 * ks][2], c[2 * ks][3]);
    p[ks][2] = pack_bf16(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    p[ks][3] = pack_bf16(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// 16 x HD fp32 accumulator of this warp -> bf16 rows staged in the warp's own smem rows -> coalesced 16-byte stores
template <int HD>
__device__ __forceinline__ void store_rows(const float (&o)[HD / 8][4], __nv_bfloat16* stage /* [16][LD] of this warp */,
                                           __nv_bfloat16* dst, long long g_ld, int valid_rows, int lane) {
  constexpr int LD = HD + 8, VPR = HD / 8;
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) {
    *reinterpret_cast<uint32_t*>(stage + g * LD + nd * 8 + 2 * t) = pack_bf16(o[nd][0], o[nd][1]);
    *reinterpret_cast<uint32_t*>(stage + (g + 8) * LD + nd * 8 + 2 * t) = pack_bf16(o[nd][2], o[nd][3]);
  }
  __syncwarp();
  for (int i = lane; i < 16 * VPR; i += 32) {
    const int r = i / VPR, v = i - r * VPR;
    if (r < valid_rows)
      *reinterpret_cast<uint4*>(dst + (long long)r * g_ld + v * 8) = *reinterpret_cast<const uint4*>(stage + r * LD + v * 8);
  }
}

// pixel_norm (networks.py:9-14) of every row of a [rows][LD] smem tile, in place: u -> u / (eps + rms(u)).
// One thread per row with 16-byte accesses (row stride LD*2 = HD*2+16 bytes: conflict-free per quarter warp).
template <int HD>
__device__ __forceinline__ void normalize_rows(__nv_bfloat16* tile, float* nrm_out, int rows) {
  constexpr int LD = HD + 8;
  for (int r = threadIdx.x; r < rows; r += kThreads) {
    uint4* row = reinterpret_cast<uint4*>(tile + r * LD);
    uint4 v[HD / 8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      v[i] = row[i];
      const float2 a = unpack_bf16(v[i].x), b = unpack_bf16(v[i].y), c = unpack_bf16(v[i].z), d = unpack_bf16(v[i].w);
      ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
    const float n = kEps + sqrtf(ss * (1.0f / HD));
    const float inv = 1.0f / n;
    if (nrm_out != nullptr) nrm_out[r] = n;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const float2 a = unpack_bf16(v[i].x), b = unpack_bf16(v[i].y), c = unpack_bf16(v[i].z), d = unpack_bf16(v[i].w);
      uint4 o;
      o.x = pack_bf16(a.x * inv, a.y * inv); o.y = pack_bf16(b.x * inv, b.y * inv);
      o.z = pack_bf16(c.x * inv, c.y * inv); o.w = pack_bf16(d.x * inv, d.y * inv);
      row[i] = o;
    }
  }
}

// Adjoint of the row normalisation applied to this warp's 16 x HD gradient tile, then stored:
//   g_u = g / n - y * (sum_d g*y) / ((n - eps) * HD)        (y = normalised row, n = eps + rms(u))
// ytile = the warp's own normalised rows in smem (also reused as the staging buffer), nrm = their n.
template <int HD>
__device__ __forceinline__ void norm_bwd_store_rows(float (&gr)[HD / 8][4], __nv_bfloat16* ytile, const float* nrm,
                                                    __nv_bfloat16* dst, long long g_ld, int valid_rows, int lane) {
  constexpr int LD = HD + 8;
  const int g = lane >> 2, t = lane & 3;
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) {
    const float2 y0 = unpack_bf16(*reinterpret_cast<const uint32_t*>(ytile + g * LD + nd * 8 + 2 * t));
    const float2 y1 = unpack_bf16(*reinterpret_cast<const uint32_t*>(ytile + (g + 8) * LD + nd * 8 + 2 * t));
    d0 += gr[nd][0] * y0.x + gr[nd][1] * y0.y;
    d1 += gr[nd][2] * y1.x + gr[nd][3] * y1.y;
  }
  d0 = quad_sum(d0);
  d1 = quad_sum(d1);
  const float n0 = nrm[g], n1 = nrm[g + 8];
  const float in0 = 1.0f / n0, in1 = 1.0f / n1;
  const float k0 = d0 / (fmaxf(n0 - kEps, 1e-20f) * HD), k1 = d1 / (fmaxf(n1 - kEps, 1e-20f) * HD);
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) {
    const float2 y0 = unpack_bf16(*reinterpret_cast<const uint32_t*>(ytile + g * LD + nd * 8 + 2 * t));
    const float2 y1 = unpack_bf16(*reinterpret_cast<const uint32_t*>(ytile + (g + 8) * LD + nd * 8 + 2 * t));
    gr[nd][0] = gr[nd][0] * in0 - y0.x * k0; gr[nd][1] = gr[nd][1] * in0 - y0.y * k0;
    gr[nd][2] = gr[nd][2] * in1 - y1.x * k1; gr[nd][3] = gr[nd][3] * in1 - y1.y * k1;
  }
  store_rows<HD>(gr, ytile, dst, g_ld, valid_rows, lane);
}

// ------------------------------------------------------------------------------------------------
// forward: grid (ceil(S/64), B*heads)
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ y, float* __restrict__ lse, int B, int S,
                int heads, float scale) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int Sp = (S + kBlk - 1) / kBlk * kBlk;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Vs = Ks + (size_t)Sp * LD;
  __nv_bfloat16* Qs = Vs + (size_t)Sp * LD;
  const int bh = blockIdx.y, q0 = blockIdx.x * kBlk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = bh / heads, head = bh - b * heads;
  const int C = heads * HD;
  const long long ld3 = 3LL * C;
  const __nv_bfloat16* q = qkv + (long long)b * S * ld3 + head * HD;   // + C: k, + 2C: v
  int qvalid = S - q0;
  if (qvalid > kBlk) qvalid = kBlk;
  load_tile<HD>(Qs, q + (long long)q0 * ld3, ld3, kBlk, qvalid);
  load_tile<HD>(Ks, q + C, ld3, Sp, S);
  load_tile<HD>(Vs, q + 2 * C, ld3, Sp, S);
  cp_async_wait_all();
  __syncthreads();
  normalize_rows<HD>(Qs, nullptr, kBlk);
  normalize_rows<HD>(Ks, nullptr, Sp);
  normalize_rows<HD>(Vs, nullptr, Sp);
  __syncthreads();

  uint32_t qf[HD / 16][4];
  load_a_frags<HD>(qf, Qs, warp * 16, lane);
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sc = scale * kLog2e;
  for (int kc = 0; kc < Sp; kc += kBlk) {
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    gemm_a_bt<HD, 8>(s, qf, Qs, warp * 16, Ks, kc, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kc + nt * 8 + 2 * t + (e & 1);
        const float v = key < S ? s[nt][e] * sc : -INFINITY;
        s[nt][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
    const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mn0); s[nt][1] = exp2f(s[nt][1] - mn0);
      s[nt][2] = exp2f(s[nt][2] - mn1); s[nt][3] = exp2f(s[nt][3] - mn1);
      ps0 += s[nt][0] + s[nt][1];
      ps1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + ps0;
    l1 = l1 * c1 + ps1;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
    uint32_t p[4][4];
    acc_to_a<8>(p, s);
    gemm_p_b<HD, 4>(o, p, Vs, kc, lane);
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { o[i][0] *= i0; o[i][1] *= i0; o[i][2] *= i1; o[i][3] *= i1; }
  const int r0 = q0 + warp * 16;
  if (lse != nullptr && t == 0) {
    if (r0 + g < S) lse[(long long)bh * S + r0 + g] = (m0 + log2f(l0)) * kLn2;
    if (r0 + g + 8 < S) lse[(long long)bh * S + r0 + g + 8] = (m1 + log2f(l1)) * kLn2;
  }
  store_rows<HD>(o, Qs + warp * 16 * LD, y + ((long long)b * S + r0) * C + head * HD, C, S - r0, lane);
}

// ------------------------------------------------------------------------------------------------
// backward 1: per 64-query block -> delta, dQ
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ y,
                   const __nv_bfloat16* __restrict__ g_y, const float* __restrict__ lse, float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ g_qkv, int B, int S, int heads, float scale) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int Sp = (S + kBlk - 1) / kBlk * kBlk;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Vs = Ks + (size_t)Sp * LD;
  __nv_bfloat16* Qs = Vs + (size_t)Sp * LD;
  __nv_bfloat16* dOs = Qs + (size_t)kBlk * LD;
  float* qn_s = reinterpret_cast<float*>(dOs + (size_t)kBlk * LD);   // [64] n = eps + rms of the raw q rows
  const int bh = blockIdx.y, q0 = blockIdx.x * kBlk;
  const int b = bh / heads, head = bh - b * heads;
  const int C = heads * HD;
  const long long ld3 = 3LL * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* q = qkv + (long long)b * S * ld3 + head * HD;
  int qvalid = S - q0;
  if (qvalid > kBlk) qvalid = kBlk;
  const __nv_bfloat16* go = g_y + ((long long)b * S + q0) * C + head * HD;
  const __nv_bfloat16* oo = y + ((long long)b * S + q0) * C + head * HD;
  load_tile<HD>(Qs, q + (long long)q0 * ld3, ld3, kBlk, qvalid);
  load_tile<HD>(dOs, go, C, kBlk, qvalid);
  load_tile<HD>(Ks, q + C, ld3, Sp, S);
  load_tile<HD>(Vs, q + 2 * C, ld3, Sp, S);
  // delta_r = sum_d dO * O for this warp's 16 rows: two lanes per row, 16-byte loads, one shuffle
  const int r0 = q0 + warp * 16;
  float dl0, dl1;
  {
    const int r = lane >> 1, half = lane & 1;
    float acc = 0.f;
    if (r0 + r < S) {
      const uint4* a = reinterpret_cast<const uint4*>(go + (long long)(warp * 16 + r) * C + half * (HD / 2));
      const uint4* c = reinterpret_cast<const uint4*>(oo + (long long)(warp * 16 + r) * C + half * (HD / 2));
#pragma unroll
      for (int i = 0; i < HD / 16; ++i) {
        const uint4 x = a[i], z = c[i];
        const float2 x0 = unpack_bf16(x.x), x1 = unpack_bf16(x.y), x2 = unpack_bf16(x.z), x3 = unpack_bf16(x.w);
        const float2 z0 = unpack_bf16(z.x), z1 = unpack_bf16(z.y), z2 = unpack_bf16(z.z), z3 = unpack_bf16(z.w);
        acc += x0.x * z0.x + x0.y * z0.y + x1.x * z1.x + x1.y * z1.y + x2.x * z2.x + x2.y * z2.y + x3.x * z3.x + x3.y * z3.y;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0 && r0 + r < S) delta[(long long)bh * S + r0 + r] = acc;
    dl0 = __shfl_sync(0xffffffffu, acc, 2 * g);
    dl1 = __shfl_sync(0xffffffffu, acc, 2 * (g + 8));
  }
  const float ls0 = (r0 + g < S ? lse[(long long)bh * S + r0 + g] : 0.f) * kLog2e;
  const float ls1 = (r0 + g + 8 < S ? lse[(long long)bh * S + r0 + g + 8] : 0.f) * kLog2e;
  cp_async_wait_all();
  __syncthreads();
  normalize_rows<HD>(Qs, qn_s, kBlk);
  normalize_rows<HD>(Ks, nullptr, Sp);
  normalize_rows<HD>(Vs, nullptr, Sp);
  __syncthreads();

  uint32_t qf[HD / 16][4], gf[HD / 16][4];
  load_a_frags<HD>(qf, Qs, warp * 16, lane);
  load_a_frags<HD>(gf, dOs, warp * 16, lane);
  float dq[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const float sc = scale * kLog2e;
  for (int kc = 0; kc < Sp; kc += kBlk) {
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
    gemm_a_bt<HD, 8>(s, qf, Qs, warp * 16, Ks, kc, lane);     // S  = Q K^T
    gemm_a_bt<HD, 8>(dp, gf, dOs, warp * 16, Vs, kc, lane);   // dP = dO V^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kc + nt * 8 + 2 * t + (e & 1);
        const float pr = key < S ? exp2f(s[nt][e] * sc - (e < 2 ? ls0 : ls1)) : 0.f;
        s[nt][e] = pr * (dp[nt][e] - (e < 2 ? dl0 : dl1)) * scale;   // dS
      }
    }
    uint32_t p[4][4];
    acc_to_a<8>(p, s);
    gemm_p_b<HD, 4>(dq, p, Ks, kc, lane);     // dQ += dS K
  }
  norm_bwd_store_rows<HD>(dq, Qs + warp * 16 * LD, qn_s + warp * 16,
                          g_qkv + ((long long)b * S + r0) * ld3 + head * HD, ld3, S - r0, lane);
}

// ------------------------------------------------------------------------------------------------
// backward 2: per 64-key block -> dK, dV   (query chunks of QC)
// ------------------------------------------------------------------------------------------------
template <int HD, int QC>
__global__ void __launch_bounds__(kThreads)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ g_y,
                    const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ g_qkv,
                    int B, int S, int heads, float scale) {
  pdl_trigger();   // the next kernel may be scheduled during this one's tail ...
  pdl_wait();      // ... and this one was: everything below needs its predecessors complete
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int Sp = (S + kBlk - 1) / kBlk * kBlk;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* dOs = Qs + (size_t)Sp * LD;
  __nv_bfloat16* Kb = dOs + (size_t)Sp * LD;
  __nv_bfloat16* Vb = Kb + (size_t)kBlk * LD;
  float* lse_s = reinterpret_cast<float*>(Vb + (size_t)kBlk * LD);
  float* dl_s = lse_s + Sp;
  float* kn_s = dl_s + Sp;     // [64] n of the raw k rows of this block
  float* vn_s = kn_s + kBlk;   // [64] n of the raw v rows
  const int bh = blockIdx.y, k0 = blockIdx.x * kBlk;
  const int b = bh / heads, head = bh - b * heads;
  const int C = heads * HD;
  const long long ld3 = 3LL * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
  const __nv_bfloat16* q = qkv + (long long)b * S * ld3 + head * HD;
  int kvalid = S - k0;
  if (kvalid > kBlk) kvalid = kBlk;
  load_tile<HD>(Qs, q, ld3, Sp, S);
  load_tile<HD>(dOs, g_y + (long long)b * S * C + head * HD, C, Sp, S);
  load_tile<HD>(Kb, q + C + (long long)k0 * ld3, ld3, kBlk, kvalid);
  load_tile<HD>(Vb, q + 2 * C + (long long)k0 * ld3, ld3, kBlk, kvalid);
  for (int i = threadIdx.x; i < Sp; i += kThreads) {
    lse_s[i] = i < S ? lse[(long long)bh * S + i] * kLog2e : INFINITY;   // padded queries: P = 0
    dl_s[i] = i < S ? delta[(long long)bh * S + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  normalize_rows<HD>(Qs, nullptr, Sp);
  normalize_rows<HD>(Kb, kn_s, kBlk);
  normalize_rows<HD>(Vb, vn_s, kBlk);
  __syncthreads();

  uint32_t kf[HD / 16][4], vf[HD / 16][4];
  load_a_frags<HD>(kf, Kb, warp * 16, lane);
  load_a_frags<HD>(vf, Vb, warp * 16, lane);
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  const float sc = scale * kLog2e;
  constexpr int NT = QC / 8;
  for (int qc = 0; qc < Sp; qc += QC) {
    float st[NT][4], dpt[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
    gemm_a_bt<HD, NT>(st, kf, Kb, warp * 16, Qs, qc, lane);     // S^T  = K Q^T   [key][query]
    gemm_a_bt<HD, NT>(dpt, vf, Vb, warp * 16, dOs, qc, lane);   // dP^T = V dO^T
    uint32_t pt[NT / 2][4], dst[NT / 2][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int qa = qc + nt * 8 + 2 * t;
      const float la = lse_s[qa], lb = lse_s[qa + 1], da = dl_s[qa], db = dl_s[qa + 1];
      const float p0 = exp2f(st[nt][0] * sc - la), p1 = exp2f(st[nt][1] * sc - lb);
      const float p2 = exp2f(st[nt][2] * sc - la), p3 = exp2f(st[nt][3] * sc - lb);
      st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
      dpt[nt][0] = p0 * (dpt[nt][0] - da) * scale; dpt[nt][1] = p1 * (dpt[nt][1] - db) * scale;
      dpt[nt][2] = p2 * (dpt[nt][2] - da) * scale; dpt[nt][3] = p3 * (dpt[nt][3] - db) * scale;
    }
    acc_to_a<NT>(pt, st);
    acc_to_a<NT>(dst, dpt);
    gemm_p_b<HD, NT / 2>(dv, pt, dOs, qc, lane);   // dV += P^T dO
    gemm_p_b<HD, NT / 2>(dk, dst, Qs, qc, lane);   // dK += dS^T Q
  }
  const int r0 = k0 + warp * 16;
  __nv_bfloat16* gk = g_qkv + ((long long)b * S + r0) * ld3 + C + head * HD;
  norm_bwd_store_rows<HD>(dk, Kb + warp * 16 * LD, kn_s + warp * 16, gk, ld3, S - r0, lane);
  norm_bwd_store_rows<HD>(dv, Vb + warp * 16 * LD, vn_s + warp * 16, gk + C, ld3, S - r0, lane);
}

size_t smem_fwd(int S, int hd) {
  const size_t Sp = (S + kBlk - 1) / kBlk * kBlk, LD = hd + 8;
  return (2 * Sp + kBlk) * LD * 2;
}
size_t smem_dq(int S, int hd) {
  const size_t Sp = (S + kBlk - 1) / kBlk * kBlk, LD = hd + 8;
  return (2 * Sp + 2 * kBlk) * LD * 2 + kBlk * 4;
}
size_t smem_dkv(int S, int hd) {
  const size_t Sp = (S + kBlk - 1) / kBlk * kBlk, LD = hd + 8;
  return (2 * Sp + 2 * kBlk) * LD * 2 + 2 * Sp * 4 + 2 * kBlk * 4;
}

int check_dims(int S, int hd, int heads) {
  TEDM_CHECK(hd == 32 || hd == 64 || hd == 128 || hd == 144 || hd == 192,
             "attention: head_dim must be one of 32, 64, 128, 144, 192 (got %d)", hd);
  TEDM_CHECK(S >= 1 && heads >= 1, "attention: empty problem");
  TEDM_CHECK(smem_dkv(S, hd) <= 227 * 1024, "attention: S=%d, head_dim=%d does not fit the shared-memory resident kernel", S, hd);
  return 0;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  TEDM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int HD>
int launch_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* y, float* lse, int B, int S, int heads, cudaStream_t stream) {
  const size_t smem = smem_fwd(S, HD);
  if (set_smem(attn_fwd_kernel<HD>, smem) != 0) return -1;
  dim3 grid((S + kBlk - 1) / kBlk, B * heads);
  launch_pdl(attn_fwd_kernel<HD>, grid, kThreads, smem, stream, qkv, y, lse, B, S, heads, 1.0f / sqrtf((float)HD));
  TEDM_LAUNCH_CHECK();
  return 0;
}

template <int HD, int QC>
int launch_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* g_y, const float* lse, float* delta,
               __nv_bfloat16* g_qkv, int B, int S, int heads, cudaStream_t stream) {
  const float scale = 1.0f / sqrtf((float)HD);
  dim3 grid((S + kBlk - 1) / kBlk, B * heads);
  {
    const size_t smem = smem_dq(S, HD);
    if (set_smem(attn_bwd_dq_kernel<HD>, smem) != 0) return -1;
    launch_pdl(attn_bwd_dq_kernel<HD>, grid, kThreads, smem, stream, qkv, y, g_y, lse, delta, g_qkv, B, S, heads, scale);
    TEDM_LAUNCH_CHECK();
  }
  {
    const size_t smem = smem_dkv(S, HD);
    if (set_smem(attn_bwd_dkv_kernel<HD, QC>, smem) != 0) return -1;
    launch_pdl(attn_bwd_dkv_kernel<HD, QC>, grid, kThreads, smem, stream, qkv, g_y, lse, delta, g_qkv, B, S, heads, scale);
    TEDM_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace

int attention_forward(const __nv_bfloat16* qkv, __nv_bfloat16* y, float* lse, int B, int S, int heads, int hd,
                      cudaStream_t stream) {
  if (check_dims(S, hd, heads) != 0) return -1;
  static const bool use_tc = [] { const char* e = getenv("TEDM_ATTN_TC"); return !(e != nullptr && e[0] == '0'); }();
  if (use_tc && attention_tc_supported(S, hd)) return attention_forward_tc(qkv, y, lse, B, S, heads, stream);
  switch (hd) {
    case 32: return launch_fwd<32>(qkv, y, lse, B, S, heads, stream);
    case 64: return launch_fwd<64>(qkv, y, lse, B, S, heads, stream);
    case 128: return launch_fwd<128>(qkv, y, lse, B, S, heads, stream);
    case 144: return launch_fwd<144>(qkv, y, lse, B, S, heads, stream);
    default: return launch_fwd<192>(qkv, y, lse, B, S, heads, stream);
  }
}

int attention_backward(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* g_y, const float* lse,
                       float* delta, __nv_bfloat16* g_qkv, int B, int S, int heads, int hd, cudaStream_t stream) {
  if (check_dims(S, hd, heads) != 0) return -1;
  static const bool use_tc = [] { const char* e = getenv("TEDM_ATTN_TC"); return !(e != nullptr && e[0] == '0'); }();
  // tcgen05 backward: the fused one-CTA-per-head kernel at S = 256; at S = 64 the two-kernel tcgen05 version (measured
  // 62 us vs 45 us for the warp-MMA kernels at B = 256: one CTA per SM, phases not overlapped — TEDM_ATTN_BWD64_TC=0
  // switches back, ~0.1 ms per CIFAR step)
  static const bool tc64 = [] { const char* e = getenv("TEDM_ATTN_BWD64_TC"); return !(e != nullptr && e[0] == '0'); }();
  if (use_tc && attention_tc_supported(S, hd) && (S == 256 || tc64))
    return attention_backward_tc(qkv, y, g_y, lse, delta, g_qkv, B, S, heads, stream);
  switch (hd) {
    case 32: return launch_bwd<32, 64>(qkv, y, g_y, lse, delta, g_qkv, B, S, heads, stream);
    case 64: return launch_bwd<64, 64>(qkv, y, g_y, lse, delta, g_qkv, B, S, heads, stream);
    case 128: return launch_bwd<128, 32>(qkv, y, g_y, lse, delta, g_qkv, B, S, heads, stream);
    case 144: return launch_bwd<144, 32>(qkv, y, g_y, lse, delta, g_qkv, B, S, heads, stream);
    default: return launch_bwd<192, 32>(qkv, y, g_y, lse, delta, g_qkv, B, S, heads, stream);
  }
}

}  // namespace tedm
