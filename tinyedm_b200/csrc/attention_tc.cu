// Cosine-normalised self-attention on tcgen05/TMEM (sm_100a), head_dim 64: forward for S = 256 or 64 keys per head,
// backward for S = 256 (fused one-CTA-per-head kernel; the older two-kernel version behind TEDM_ATTN_BWD_FUSED=0).
//
// Reference: CosineAttention.forward, src/tinyedm/networks.py:191-207 — pixel_norm over hd of q, k and v (:195), then
// F.scaled_dot_product_attention(q, k, v) with scale 1/sqrt(hd) (:201), output channel = head*hd + d (:202).
// qkv is (B,S,3C) bf16 with channel = {q,k,v}*C + head*hd + d (the weight bank permutes the qkv conv's rows).
//
// One CTA owns ONE M-tile of 128 query rows; two CTAs are resident per SM (112 KB of shared memory, 256 TMEM columns
// each) so that the TMA load / normalisation of one overlaps the MMA / softmax of the other:
//     S = 256 : half of one (image, head): 128 queries, 256 keys;
//     S =  64 : two consecutive (image, head) pairs: 128 queries, 128 keys and a block-diagonal mask (a query only sees
//               the 64 keys of its own pair).
// Pipeline per CTA (5 warps):
//     warp 4      TMA: Q tile, K, V (rows of 64 bf16, 128B-swizzled)                                 -> bar_load
//     warps 0..3  pixel_norm of every row in shared memory (thread-per-row)                           -> fence, sync
//     warp 4      tcgen05.mma  S = Q K^T   (M=128, N=keys, K=64)  fp32 in TMEM                        -> bar_s
//     warps 0..3  softmax straight out of TMEM (thread = query row): max, then per 64-key chunk exp2 / sum and the
//                 unnormalised P chunk (bf16) -> a 2-slot ring in swizzled shared memory              -> p_ready[slot]
//     warp 4      tcgen05.mma  O += P_c V_c (M=128, N=64, K=64; V is the MN-major B operand) into the TMEM columns
//                 the softmax has already consumed; commit frees the ring slot                        -> p_free[slot], bar_o
//     warps 0..3  O / rowsum -> bf16 -> y, log-sum-exp -> lse
// The S x S score matrix exists only in TMEM and, 64 keys at a time as bf16 P, in shared memory.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr int kHD = 64;
constexpr int kTileQ = 128;
constexpr int kQBytes = kTileQ * 128;                       // 16 KB
constexpr int kKVBytesMax = 256 * 128;                      // 32 KB each
constexpr int kPChunkBytes = 128 * 128;                     // [128 queries][64 keys] bf16
constexpr int kOffQ = 0, kOffK = kQBytes, kOffV = kQBytes + kKVBytesMax, kOffP = kQBytes + 2 * kKVBytesMax;
constexpr int kOffBarsA = kOffP + 2 * kPChunkBytes;
constexpr int kSmemBytesA = kOffBarsA + 128;                // 114 816 B: two CTAs per SM
constexpr int kThreadsA = 160;
constexpr float kEpsA = 1e-4f;
constexpr float kLog2eA = 1.4426950408889634f;
constexpr float kLn2A = 0.6931471805599453f;
static_assert(2 * (kSmemBytesA + 1024) <= 233472, "two CTAs per SM");

__device__ __forceinline__ uint4* prow(uint8_t* buf, int m, int j) {
  return reinterpret_cast<uint4*>(buf + m * 128 + ((j ^ (m & 7)) << 4));
}

// pixel_norm of one 64-element row held in a 128B-swizzled slab, in place (bf16 result, like the reference's cast)
__device__ __forceinline__ void normalize_row(uint8_t* slab, int r) {
  uint4 v[8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = *prow(slab, r, j);
    const float2 a = unpack_bf16(v[j].x), b = unpack_bf16(v[j].y), c = unpack_bf16(v[j].z), d = unpack_bf16(v[j].w);
    ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
  }
  const float inv = 1.0f / (kEpsA + sqrtf(ss * (1.0f / kHD)));
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 a = unpack_bf16(v[j].x), b = unpack_bf16(v[j].y), c = unpack_bf16(v[j].z), d = unpack_bf16(v[j].w);
    uint4 o;
    o.x = pack_bf16(a.x * inv, a.y * inv); o.y = pack_bf16(b.x * inv, b.y * inv);
    o.z = pack_bf16(c.x * inv, c.y * inv); o.w = pack_bf16(d.x * inv, d.y * inv);
    *prow(slab, r, j) = o;
  }
}

// NK = keys of the tile (256 for S = 256, 128 for S = 64)
template <int S>
__global__ void __launch_bounds__(kThreadsA, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                   __nv_bfloat16* __restrict__ y, float* __restrict__ lse, int n_pairs, int heads, float scale) {
  constexpr int NK = S == 256 ? 256 : 128;
  constexpr int NCH = NK / 64;                      // 64-key chunks
  constexpr int TILES_PER_PAIR = S == 256 ? 2 : 1;  // S = 64: one tile holds 2 pairs
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBarsA);
  uint64_t* bar_load = bars;          // TMA bytes
  uint64_t* bar_s = bars + 1;         // scores in TMEM
  uint64_t* p_ready = bars + 2;       // [2] 128 softmax threads wrote ring slot (and are done with those S columns)
  uint64_t* p_free = bars + 4;        // [2] the MMAs reading ring slot have completed
  uint64_t* bar_o = bars + 6;         // O in TMEM
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform warp index
  const int C = heads * kHD;
  if ((smem_u32(smem) & 1023u) != 0) __trap();      // the swizzled tiles need a 1024-byte aligned base
  if (threadIdx.x == 0) pdl_trigger();
  // tile -> (first pair, query offset)
  const int tile = blockIdx.x;
  const int pair0 = S == 256 ? tile / TILES_PER_PAIR : tile * 2;
  const int q_off = S == 256 ? (tile % TILES_PER_PAIR) * kTileQ : 0;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_kv);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&p_ready[i], 128);
        mbar_init(&p_free[i], 1);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // the prologue above overlapped the previous kernel's tail; its results are needed from here on

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, kQBytes + 2 * NK * 128);
      if (S == 256) {
        const int b = pair0 / heads, head = pair0 - b * heads;
        tma_load_2d(smem + kOffQ, &tmap_q, bar_load, head * kHD, b * S + q_off);
        tma_load_2d(smem + kOffK, &tmap_kv, bar_load, C + head * kHD, b * S);
        tma_load_2d(smem + kOffV, &tmap_kv, bar_load, 2 * C + head * kHD, b * S);
      } else {
        for (int pp = 0; pp < 2; ++pp) {
          int pair = pair0 + pp;
          if (pair >= n_pairs) pair = n_pairs - 1;      // tail tile: duplicate the last pair (results discarded)
          const int b = pair / heads, head = pair - b * heads;
          tma_load_2d(smem + kOffQ + pp * 64 * 128, &tmap_kv, bar_load, head * kHD, b * S);
          tma_load_2d(smem + kOffK + pp * 64 * 128, &tmap_kv, bar_load, C + head * kHD, b * S);
          tma_load_2d(smem + kOffV + pp * 64 * 128, &tmap_kv, bar_load, 2 * C + head * kHD, b * S);
        }
      }
    }
  } else {
    // ---- pixel_norm of the q, k, v rows (thread-per-row) ----
    mbar_wait_bounded(bar_load, 0);
    normalize_row(smem + kOffQ, threadIdx.x);
    for (int r = threadIdx.x; r < NK; r += 128) {
      normalize_row(smem + kOffK, r);
      normalize_row(smem + kOffV, r);
    }
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (warp == 4) {
    if (lane == 0) {
      tc_fence_after();
      {  // ---- S = Q K^T ----
        const uint32_t idesc_s = make_idesc_bf16(128, NK, 0, 0);
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kOffQ), 0, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kOffK), 0, 1024);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      // ---- O += P_c V_c, chunk by chunk through the 2-slot ring ----
      const uint32_t idesc_o = make_idesc_bf16(128, kHD, 0, 1);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int slot = c & 1;
        mbar_wait_bounded(&p_ready[slot], (c >> 1) & 1);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kOffP + slot * kPChunkBytes), 0, 1024);
        // V rows (keys) 64c .. 64c+63: MN-major B operand, 16 keys (K) = 16 rows of 128 B per MMA
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kOffV + c * 64 * 128), 64 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 128), idesc_o, (c | k) != 0 ? 1u : 0u);
        umma_commit(&p_free[slot]);
      }
      umma_commit(bar_o);
    }
  } else {
    // ---- softmax + output (thread = query row) ----
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    // keys this row may see: all NK for S = 256; the 64 of its own pair for S = 64
    const int kb = S == 256 ? 0 : (m >> 6) * 64;
    const int kn = S == 256 ? NK : 64;
    const float sc = scale * kLog2eA;
    mbar_wait_bounded(bar_s, 0);
    tc_fence_after();
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent chains (a single one is latency bound)
#pragma unroll 1
    for (int c0 = 0; c0 < kn; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(t_row + kb + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
    }
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    const float mxs = mx * sc;
    float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const int slot = c & 1;
      uint8_t* pbuf = smem + kOffP + slot * kPChunkBytes;
      if (c >= 2) mbar_wait_bounded(&p_free[slot], ((c >> 1) - 1) & 1);   // the MMAs of chunk c-2 released the slot
      const bool live = (c * 64 >= kb) && (c * 64 < kb + kn);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (live) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 64 + h * 32, r);
          tmem_ld_wait();
          float pv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            pv[i] = exp2f(fmaf(__uint_as_float(r[i]), sc, -mxs));
            sum4[i & 3] += pv[i];     // fp32 row sum of the unrounded probabilities (as the warp-MMA kernel does)
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16(pv[g * 8 + 0], pv[g * 8 + 1]); o.y = pack_bf16(pv[g * 8 + 2], pv[g * 8 + 3]);
            o.z = pack_bf16(pv[g * 8 + 4], pv[g * 8 + 5]); o.w = pack_bf16(pv[g * 8 + 6], pv[g * 8 + 7]);
            *prow(pbuf, m, h * 4 + g) = o;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) *prow(pbuf, m, h * 4 + g) = make_uint4(0, 0, 0, 0);   // keys of the other pair
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[slot]);
    }
    // ---- epilogue ----
    mbar_wait_bounded(bar_o, 0);
    tc_fence_after();
    const int pair = S == 256 ? pair0 : pair0 + (m >> 6);
    const int s_idx = S == 256 ? q_off + m : (m & 63);
    const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
    const float inv = 1.0f / sum;
    uint32_t o0[32], o1[32];
    tmem_ld32(t_row, o0);
    tmem_ld32(t_row + 32, o1);
    tmem_ld_wait();
    if (pair < n_pairs) {
      const int b = pair / heads, head = pair - b * heads;
      __nv_bfloat16* dst = y + ((long long)b * S + s_idx) * C + head * kHD;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(o0[g * 8 + 0]) * inv, __uint_as_float(o0[g * 8 + 1]) * inv);
        o.y = pack_bf16(__uint_as_float(o0[g * 8 + 2]) * inv, __uint_as_float(o0[g * 8 + 3]) * inv);
        o.z = pack_bf16(__uint_as_float(o0[g * 8 + 4]) * inv, __uint_as_float(o0[g * 8 + 5]) * inv);
        o.w = pack_bf16(__uint_as_float(o0[g * 8 + 6]) * inv, __uint_as_float(o0[g * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[g] = o;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(o1[g * 8 + 0]) * inv, __uint_as_float(o1[g * 8 + 1]) * inv);
        o.y = pack_bf16(__uint_as_float(o1[g * 8 + 2]) * inv, __uint_as_float(o1[g * 8 + 3]) * inv);
        o.z = pack_bf16(__uint_as_float(o1[g * 8 + 4]) * inv, __uint_as_float(o1[g * 8 + 5]) * inv);
        o.w = pack_bf16(__uint_as_float(o1[g * 8 + 6]) * inv, __uint_as_float(o1[g * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[4 + g] = o;
      }
      if (lse != nullptr) lse[(long long)pair * S + s_idx] = (mxs + log2f(sum)) * kLn2A;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =====================================================================================================================
// Backward (autograd of networks.py:195-201). Same tiling as the forward, two kernels without atomics:
//   dq  kernel: tile = 128 queries;  S = Qn Kn^T and dP = dO Vn^T into TMEM (2 x 256 columns), per 64-key chunk
//               dS = exp2(S c - lse) (dP - delta) / sqrt(hd) -> bf16 ring -> dQn += dS Kn; also writes delta = rowsum(dO o O)
//   dkv kernel: tile = 128 keys;     S^T = Kn Qn^T and dP^T = Vn dO^T into TMEM, per 64-query chunk P^T and dS^T -> two
//               bf16 rings -> dVn += P^T dO, dKn += dS^T Qn
// Each result row then goes through the adjoint of its pixel norm  g_u = g/n - y (g.y) / ((n - eps) hd)  in registers.
// =====================================================================================================================
// 8 compute warps (two per TMEM lane quarter: "sub" 0/1 split the 64-key chunks) + 1 TMA/MMA warp. One compute warp per
// scheduler (the first version) was latency bound: ~3 700 dependent instructions per thread and nothing to hide them.
constexpr int kThreadsB = 288;
// dq: Q tile 16K | dO tile 16K | K 32K | V 32K | dS ring 4 x 16K (one slot per key chunk: no slot recycling)
constexpr int kDqOffQ = 0, kDqOffDO = 16384, kDqOffK = 32768, kDqOffV = 65536, kDqOffRing = 98304, kDqOffBars = 163840;
constexpr int kDqSmem = kDqOffBars + 128;
// dkv: K tile 16K | V tile 16K | Q 32K | dO 32K | P^T ring 4 x 16K | dS^T ring 4 x 16K | lse2[256] | delta[256]
constexpr int kKvOffK = 0, kKvOffV = 16384, kKvOffQ = 32768, kKvOffDO = 65536, kKvOffRingP = 98304, kKvOffRingS = 163840,
              kKvOffVec = 229376, kKvOffBars = 229376 + 2048;
constexpr int kKvSmem = kKvOffBars + 128;
static_assert(kKvSmem <= 232448, "shared memory budget");

// pixel-norm adjoint of one 64-element gradient row held as two 32-float halves; y = the normalised row in a swizzled slab
__device__ __forceinline__ void norm_adjoint_store(const uint32_t (&g0)[32], const uint32_t (&g1)[32], uint8_t* slab, int r,
                                                   float n, __nv_bfloat16* dst) {
  float yv[64];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 u = *prow(slab, r, j);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    yv[j * 8 + 0] = a.x; yv[j * 8 + 1] = a.y; yv[j * 8 + 2] = b.x; yv[j * 8 + 3] = b.y;
    yv[j * 8 + 4] = c.x; yv[j * 8 + 5] = c.y; yv[j * 8 + 6] = d.x; yv[j * 8 + 7] = d.y;
  }
  float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    d4[i & 3] += __uint_as_float(g0[i]) * yv[i];
    d4[i & 3] += __uint_as_float(g1[i]) * yv[32 + i];
  }
  const float dot = (d4[0] + d4[1]) + (d4[2] + d4[3]);
  const float inv_n = 1.0f / n;
  const float k = dot / (fmaxf(n - kEpsA, 1e-20f) * kHD);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = j * 8 + e;
      const float g = i < 32 ? __uint_as_float(g0[i]) : __uint_as_float(g1[i - 32]);
      o[e] = g * inv_n - yv[i] * k;
    }
    uint4 u;
    u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
    reinterpret_cast<uint4*>(dst)[j] = u;
  }
}

// like normalize_row, returns n = eps + rms of the raw row
__device__ __forceinline__ float normalize_row_n(uint8_t* slab, int r) {
  uint4 v[8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = *prow(slab, r, j);
    const float2 a = unpack_bf16(v[j].x), b = unpack_bf16(v[j].y), c = unpack_bf16(v[j].z), d = unpack_bf16(v[j].w);
    ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
  }
  const float n = kEpsA + sqrtf(ss * (1.0f / kHD));
  const float inv = 1.0f / n;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 a = unpack_bf16(v[j].x), b = unpack_bf16(v[j].y), c = unpack_bf16(v[j].z), d = unpack_bf16(v[j].w);
    uint4 o;
    o.x = pack_bf16(a.x * inv, a.y * inv); o.y = pack_bf16(b.x * inv, b.y * inv);
    o.z = pack_bf16(c.x * inv, c.y * inv); o.w = pack_bf16(d.x * inv, d.y * inv);
    *prow(slab, r, j) = o;
  }
  return n;
}

template <int S>
__global__ void __launch_bounds__(kThreadsB, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                      const __grid_constant__ CUtensorMap tmap_do, const __nv_bfloat16* __restrict__ y,
                      const float* __restrict__ lse, float* __restrict__ delta, __nv_bfloat16* __restrict__ g_qkv,
                      int n_pairs, int heads, float scale) {
  constexpr int NK = S == 256 ? 256 : 128;
  constexpr int NCH = NK / 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDqOffBars);
  uint64_t* bar_load = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* p_ready = bars + 2;   // [4] one per key chunk
  uint64_t* bar_o = bars + 6;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform warp index
  const int C = heads * kHD;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) pdl_trigger();
  const int tile = blockIdx.x;
  const int pair0 = S == 256 ? tile / 2 : tile * 2;
  const int q_off = S == 256 ? (tile & 1) * kTileQ : 0;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_kv);
      tma_prefetch_desc(&tmap_do);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      for (int i = 0; i < 4; ++i) mbar_init(&p_ready[i], 128);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // the prologue above overlapped the previous kernel's tail; its results are needed from here on

  float n_q = 1.f, dl = 0.f, ls2 = 0.f;
  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, 2 * kQBytes + 2 * NK * 128);
      if (S == 256) {
        const int b = pair0 / heads, head = pair0 - b * heads;
        tma_load_2d(smem + kDqOffQ, &tmap_q, bar_load, head * kHD, b * S + q_off);
        tma_load_2d(smem + kDqOffDO, &tmap_do, bar_load, head * kHD, b * S + q_off);
        tma_load_2d(smem + kDqOffK, &tmap_kv, bar_load, C + head * kHD, b * S);
        tma_load_2d(smem + kDqOffV, &tmap_kv, bar_load, 2 * C + head * kHD, b * S);
      } else {
        for (int pp = 0; pp < 2; ++pp) {
          int pair = pair0 + pp;
          if (pair >= n_pairs) pair = n_pairs - 1;
          const int b = pair / heads, head = pair - b * heads;
          tma_load_2d(smem + kDqOffQ + pp * 8192, &tmap_kv, bar_load, head * kHD, b * S);
          tma_load_2d(smem + kDqOffDO + pp * 8192, &tmap_do, bar_load, head * kHD, b * S);
          tma_load_2d(smem + kDqOffK + pp * 8192, &tmap_kv, bar_load, C + head * kHD, b * S);
          tma_load_2d(smem + kDqOffV + pp * 8192, &tmap_kv, bar_load, 2 * C + head * kHD, b * S);
        }
      }
    }
  } else {
    const int m = threadIdx.x & 127;            // query row of this thread (both subs own the same rows)
    const int sub = threadIdx.x >> 7;
    const int pair = S == 256 ? pair0 : pair0 + (m >> 6);
    const int s_idx = S == 256 ? q_off + m : (m & 63);
    const bool live = pair < n_pairs;
    const int pr = live ? pair : n_pairs - 1;
    const int b = pr / heads, head = pr - b * heads;
    // delta = sum_d dO * O of this thread's query row (O straight from global, dO from the TMA tile)
    const uint4* orow = reinterpret_cast<const uint4*>(y + ((long long)b * S + s_idx) * C + head * kHD);
    uint4 ov[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ov[j] = orow[j];
    ls2 = lse[(long long)pr * S + s_idx] * kLog2eA;
    mbar_wait_bounded(bar_load, 0);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 g = *prow(smem + kDqOffDO, m, j);
      const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
      const float2 o0 = unpack_bf16(ov[j].x), o1 = unpack_bf16(ov[j].y), o2 = unpack_bf16(ov[j].z), o3 = unpack_bf16(ov[j].w);
      acc += g0.x * o0.x + g0.y * o0.y + g1.x * o1.x + g1.y * o1.y + g2.x * o2.x + g2.y * o2.y + g3.x * o3.x + g3.y * o3.y;
    }
    dl = acc;
    if (live && sub == 0) delta[(long long)pair * S + s_idx] = acc;
    if (sub == 0) n_q = normalize_row_n(smem + kDqOffQ, m);
    for (int r = threadIdx.x; r < NK; r += 256) {
      normalize_row(smem + kDqOffK, r);
      normalize_row(smem + kDqOffV, r);
    }
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (warp == 8) {
    if (lane == 0) {
      tc_fence_after();
      const uint32_t idesc_s = make_idesc_bf16(128, NK, 0, 0);
      {
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kDqOffQ), 0, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kDqOffK), 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
      }
      {
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kDqOffDO), 0, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kDqOffV), 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + 256, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
      const uint32_t idesc_o = make_idesc_bf16(128, kHD, 0, 1);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        mbar_wait_bounded(&p_ready[c], 0);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kDqOffRing + c * kPChunkBytes), 0, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kDqOffK + c * 64 * 128), 64 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 128), idesc_o, (c | k) != 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
  } else {
    const int q = warp & 3;
    const int sub = warp >> 2;                     // which half of the key chunks this warp handles
    const int m = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int kb = S == 256 ? 0 : (m >> 6) * 64;
    const int kn = S == 256 ? NK : 64;
    const float sc = scale * kLog2eA;
    mbar_wait_bounded(bar_s, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = sub * (NCH / 2); c < (sub + 1) * (NCH / 2); ++c) {
      uint8_t* pbuf = smem + kDqOffRing + c * kPChunkBytes;
      const bool on = (c * 64 >= kb) && (c * 64 < kb + kn);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (on) {
          uint32_t rs[32], rp[32];
          tmem_ld32(t_row + c * 64 + h * 32, rs);
          tmem_ld32(t_row + 256 + c * 64 + h * 32, rp);
          tmem_ld_wait();
          float ds[32];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            ds[i] = exp2f(fmaf(__uint_as_float(rs[i]), sc, -ls2)) * (__uint_as_float(rp[i]) - dl) * scale;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16(ds[g * 8 + 0], ds[g * 8 + 1]); o.y = pack_bf16(ds[g * 8 + 2], ds[g * 8 + 3]);
            o.z = pack_bf16(ds[g * 8 + 4], ds[g * 8 + 5]); o.w = pack_bf16(ds[g * 8 + 6], ds[g * 8 + 7]);
            *prow(pbuf, m, h * 4 + g) = o;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) *prow(pbuf, m, h * 4 + g) = make_uint4(0, 0, 0, 0);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[c]);
    }
    mbar_wait_bounded(bar_o, 0);
    tc_fence_after();
    const int pair = S == 256 ? pair0 : pair0 + (m >> 6);
    const int s_idx = S == 256 ? q_off + m : (m & 63);
    uint32_t g0[32], g1[32];
    tmem_ld32(t_row, g0);
    tmem_ld32(t_row + 32, g1);
    tmem_ld_wait();
    if (pair < n_pairs && sub == 0) {
      const int b = pair / heads, head = pair - b * heads;
      norm_adjoint_store(g0, g1, smem + kDqOffQ, m, n_q, g_qkv + ((long long)b * S + s_idx) * 3 * C + head * kHD);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int S>
__global__ void __launch_bounds__(kThreadsB, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmap_t, const __grid_constant__ CUtensorMap tmap_all,
                       const __grid_constant__ CUtensorMap tmap_do_all, const float* __restrict__ lse,
                       const float* __restrict__ delta, __nv_bfloat16* __restrict__ g_qkv, int n_pairs, int heads,
                       float scale) {
  constexpr int NQ = S == 256 ? 256 : 128;   // queries seen by the tile's keys
  constexpr int NCH = NQ / 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kKvOffBars);
  uint64_t* bar_load = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* p_ready = bars + 2;   // [4] one per query chunk
  uint64_t* bar_o = bars + 6;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);
  float* lse_s = reinterpret_cast<float*>(smem + kKvOffVec);
  float* dl_s = lse_s + 256;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform warp index
  const int C = heads * kHD;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) pdl_trigger();
  const int tile = blockIdx.x;
  const int pair0 = S == 256 ? tile / 2 : tile * 2;
  const int k_off = S == 256 ? (tile & 1) * 128 : 0;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_t);
      tma_prefetch_desc(&tmap_all);
      tma_prefetch_desc(&tmap_do_all);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      for (int i = 0; i < 4; ++i) mbar_init(&p_ready[i], 128);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // the prologue above overlapped the previous kernel's tail; its results are needed from here on

  float n_k = 1.f, n_v = 1.f;
  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, 2 * kQBytes + 2 * NQ * 128);
      if (S == 256) {
        const int b = pair0 / heads, head = pair0 - b * heads;
        tma_load_2d(smem + kKvOffK, &tmap_t, bar_load, C + head * kHD, b * S + k_off);
        tma_load_2d(smem + kKvOffV, &tmap_t, bar_load, 2 * C + head * kHD, b * S + k_off);
        tma_load_2d(smem + kKvOffQ, &tmap_all, bar_load, head * kHD, b * S);
        tma_load_2d(smem + kKvOffDO, &tmap_do_all, bar_load, head * kHD, b * S);
      } else {
        for (int pp = 0; pp < 2; ++pp) {
          int pair = pair0 + pp;
          if (pair >= n_pairs) pair = n_pairs - 1;
          const int b = pair / heads, head = pair - b * heads;
          tma_load_2d(smem + kKvOffK + pp * 8192, &tmap_all, bar_load, C + head * kHD, b * S);
          tma_load_2d(smem + kKvOffV + pp * 8192, &tmap_all, bar_load, 2 * C + head * kHD, b * S);
          tma_load_2d(smem + kKvOffQ + pp * 8192, &tmap_all, bar_load, head * kHD, b * S);
          tma_load_2d(smem + kKvOffDO + pp * 8192, &tmap_do_all, bar_load, head * kHD, b * S);
        }
      }
    }
  } else {
    const int m = threadIdx.x & 127;            // key row of this thread (both subs own the same rows)
    const int sub = threadIdx.x >> 7;
    // per-query log-sum-exp (log2 domain) and delta of every query this tile sees
    for (int i = threadIdx.x; i < NQ; i += 256) {
      const int pair = S == 256 ? pair0 : pair0 + (i >> 6);
      const int pr = pair < n_pairs ? pair : n_pairs - 1;
      const int qi = S == 256 ? i : (i & 63);
      lse_s[i] = lse[(long long)pr * S + qi] * kLog2eA;
      dl_s[i] = delta[(long long)pr * S + qi];
    }
    mbar_wait_bounded(bar_load, 0);
    if (sub == 0) n_k = normalize_row_n(smem + kKvOffK, m);   // sub 0 owns the dK epilogue, sub 1 the dV epilogue
    else n_v = normalize_row_n(smem + kKvOffV, m);
    for (int r = threadIdx.x; r < NQ; r += 256) normalize_row(smem + kKvOffQ, r);
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (warp == 8) {
    if (lane == 0) {
      tc_fence_after();
      const uint32_t idesc_s = make_idesc_bf16(128, NQ, 0, 0);
      {  // S^T = Kn Qn^T
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kKvOffK), 0, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kKvOffQ), 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
      }
      {  // dP^T = Vn dO^T
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kKvOffV), 0, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kKvOffDO), 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + 256, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
      const uint32_t idesc_o = make_idesc_bf16(128, kHD, 0, 1);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        mbar_wait_bounded(&p_ready[c], 0);
        tc_fence_after();
        const uint64_t ap_desc = make_smem_desc_sw128(smem_u32(smem + kKvOffRingP + c * kPChunkBytes), 0, 1024);
        const uint64_t as_desc = make_smem_desc_sw128(smem_u32(smem + kKvOffRingS + c * kPChunkBytes), 0, 1024);
        const uint64_t bdo_desc = make_smem_desc_sw128(smem_u32(smem + kKvOffDO + c * 64 * 128), 64 * 128, 1024);
        const uint64_t bq_desc = make_smem_desc_sw128(smem_u32(smem + kKvOffQ + c * 64 * 128), 64 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dVn += P^T_c dO_c   -> columns [0, 64)
          umma_bf16(tmem_base, ap_desc + (uint64_t)(k * 2), bdo_desc + (uint64_t)(k * 128), idesc_o, (c | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dKn += dS^T_c Qn_c  -> columns [256, 320)
          umma_bf16(tmem_base + 256, as_desc + (uint64_t)(k * 2), bq_desc + (uint64_t)(k * 128), idesc_o, (c | k) != 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
  } else {
    const int q = warp & 3;
    const int sub = warp >> 2;                        // which half of the query chunks this warp handles
    const int m = q * 32 + lane;                      // key row of the tile
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int qb = S == 256 ? 0 : (m >> 6) * 64;      // queries this key is seen by
    const int qn = S == 256 ? NQ : 64;
    const float sc = scale * kLog2eA;
    mbar_wait_bounded(bar_s, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = sub * (NCH / 2); c < (sub + 1) * (NCH / 2); ++c) {
      uint8_t* pbuf = smem + kKvOffRingP + c * kPChunkBytes;
      uint8_t* sbuf = smem + kKvOffRingS + c * kPChunkBytes;
      const bool on = (c * 64 >= qb) && (c * 64 < qb + qn);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (on) {
          uint32_t rs[32], rp[32];
          tmem_ld32(t_row + c * 64 + h * 32, rs);
          tmem_ld32(t_row + 256 + c * 64 + h * 32, rp);
          tmem_ld_wait();
          const float* lq = lse_s + c * 64 + h * 32;
          const float* dq_ = dl_s + c * 64 + h * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pv[8], dv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int i = g * 8 + e;
              pv[e] = exp2f(fmaf(__uint_as_float(rs[i]), sc, -lq[i]));
              dv[e] = pv[e] * (__uint_as_float(rp[i]) - dq_[i]) * scale;
            }
            uint4 o;
            o.x = pack_bf16(pv[0], pv[1]); o.y = pack_bf16(pv[2], pv[3]); o.z = pack_bf16(pv[4], pv[5]); o.w = pack_bf16(pv[6], pv[7]);
            *prow(pbuf, m, h * 4 + g) = o;
            o.x = pack_bf16(dv[0], dv[1]); o.y = pack_bf16(dv[2], dv[3]); o.z = pack_bf16(dv[4], dv[5]); o.w = pack_bf16(dv[6], dv[7]);
            *prow(sbuf, m, h * 4 + g) = o;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            *prow(pbuf, m, h * 4 + g) = make_uint4(0, 0, 0, 0);
            *prow(sbuf, m, h * 4 + g) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[c]);
    }
    mbar_wait_bounded(bar_o, 0);
    tc_fence_after();
    const int pair = S == 256 ? pair0 : pair0 + (m >> 6);
    const int s_idx = S == 256 ? k_off + m : (m & 63);
    uint32_t g0[32], g1[32];
    if (pair < n_pairs) {
      const int b = pair / heads, head = pair - b * heads;
      __nv_bfloat16* base = g_qkv + ((long long)b * S + s_idx) * 3 * C + head * kHD;
      if (sub == 0) {
        tmem_ld32(t_row + 256, g0);
        tmem_ld32(t_row + 256 + 32, g1);
        tmem_ld_wait();
        norm_adjoint_store(g0, g1, smem + kKvOffK, m, n_k, base + C);
      } else {
        tmem_ld32(t_row, g0);
        tmem_ld32(t_row + 32, g1);
        tmem_ld_wait();
        norm_adjoint_store(g0, g1, smem + kKvOffV, m, n_v, base + 2 * C);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// Fused backward for S = 256: ONE CTA per (image, head) computes dQ, dK and dV. The transposed scores are computed once
// per (key block, query chunk) and feed all three gradient GEMMs, so q/k/v/dO are loaded and normalised once and every
// exponential is evaluated once (the two-kernel version above does each of those twice and launches 4x the CTAs).
//   unit u = (j, i, h): key block j (128 keys = TMEM lanes), query chunk (i, h) = 64 queries 128 i + 64 h ..
//     MMA      S^T_u = Kn_j Qn_u^T, dP^T_u = Vn_j dO_u^T        (M=128, N=64, K=64) -> 2-slot TMEM rings
//     threads  P^T = exp2(S^T c - lse_q), dS^T = P^T (dP^T - delta_q)/sqrt(hd) -> bf16, swizzled [128 keys][64 queries]
//     MMA      dVn_j += P^T_u dO_u,  dKn_j += dS^T_u Qn_u        (M=128 keys, N=64, K=64 queries)
//              dQn_i += dS_(j,i) Kn_j after both chunks of the block: A = the two dS^T chunks read MN-major
//                                                                 (M=128 queries, N=64, K=128 keys)
// TMEM (512 columns): S^T ring 2x64 | dP^T ring 2x64 | dVn_j 64 | dKn_j 64 | dQn_0 64 | dQn_1 64. dVn_0 / dKn_0 are
// drained (norm adjoint + store) after the 4 units of key block 0, before block 1 reuses their columns.
// Shared memory: Q, K, V, dO 4 x 32 KB | P^T ring 2 x 16 KB | dS^T ring 4 x 16 KB (a block's two chunks stay until
// its dQ MMA has read them) | lse, delta vectors.
// =====================================================================================================================
constexpr int kFuOffQ = 0, kFuOffK = 32768, kFuOffV = 65536, kFuOffDO = 98304, kFuOffPT = 131072, kFuOffDST = 163840,
              kFuOffVec = 229376, kFuOffBars = 229376 + 2048;
constexpr int kFuSmem = kFuOffBars + 128;
static_assert(kFuSmem <= 232448, "shared memory budget");
constexpr uint32_t kFuColS = 0, kFuColDP = 128, kFuColDV = 256, kFuColDK = 320, kFuColDQ = 384;

__global__ void __launch_bounds__(kThreadsB, 1)
attn_bwd_fused_tc_kernel(const __grid_constant__ CUtensorMap tmap_all, const __grid_constant__ CUtensorMap tmap_do_all,
                         const __nv_bfloat16* __restrict__ y, const float* __restrict__ lse,
                         __nv_bfloat16* __restrict__ g_qkv, int heads, float scale) {
  constexpr int S = 256, NU = 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFuOffBars);
  uint64_t* bar_load = bars;
  uint64_t* bar_s = bars + 1;        // [2] S^T / dP^T of a unit are in TMEM ring slot u & 1
  uint64_t* p_ready = bars + 3;      // [2] 256 threads wrote P^T / dS^T of a unit (and are done with its TMEM slot)
  uint64_t* mma_done = bars + 5;     // [2] every MMA issued up to and including the unit's gradient MMAs completed
  uint64_t* kv_drained = bars + 7;   // 256 threads read dVn_0 / dKn_0 out of TMEM
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  float* lse_s = reinterpret_cast<float*>(smem + kFuOffVec);
  float* dl_s = lse_s + 256;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform warp index
  const int C = heads * kHD;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) pdl_trigger();
  const int pair = blockIdx.x;
  const int b = pair / heads, head = pair - b * heads;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_all);
      tma_prefetch_desc(&tmap_do_all);
      mbar_init(bar_load, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_s[i], 1);
        mbar_init(&p_ready[i], 256);
        mbar_init(&mma_done[i], 1);
      }
      mbar_init(kv_drained, 256);
      mbar_fence_init();
      // the 128 KB of operands are requested before anything else (the issuing thread initialised the barrier itself)
      pdl_wait();
      mbar_expect_tx(bar_load, 4 * 32768);
      tma_load_2d(smem + kFuOffQ, &tmap_all, bar_load, head * kHD, b * S);
      tma_load_2d(smem + kFuOffK, &tmap_all, bar_load, C + head * kHD, b * S);
      tma_load_2d(smem + kFuOffV, &tmap_all, bar_load, 2 * C + head * kHD, b * S);
      tma_load_2d(smem + kFuOffDO, &tmap_do_all, bar_load, head * kHD, b * S);
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // O row and log-sum-exp of this thread's query: in flight across the TMEM allocation / CTA barrier
  uint4 ov[8];
  float lse_r = 0.f;
  if (warp != 8) {
    pdl_wait();
    const int r = threadIdx.x;
    const uint4* orow = reinterpret_cast<const uint4*>(y + ((long long)b * S + r) * C + head * kHD);
#pragma unroll
    for (int j = 0; j < 8; ++j) ov[j] = orow[j];
    lse_r = lse[(long long)pair * S + r];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // norms of the rows whose gradients this thread finishes: two rows of K (sub 0) or V (sub 1), one row of Q
  float n_kv[2] = {1.f, 1.f}, n_q = 1.f;
  if (warp != 8) {
    const int r = threadIdx.x;                    // query row 0..255 owned by this thread for lse / delta / Q norm
    const int sub = threadIdx.x >> 7, m = threadIdx.x & 127;
    lse_s[r] = lse_r * kLog2eA;
    mbar_wait_bounded(bar_load, 0);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 g = *prow(smem + kFuOffDO, r, j);
      const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
      const float2 o0 = unpack_bf16(ov[j].x), o1 = unpack_bf16(ov[j].y), o2 = unpack_bf16(ov[j].z), o3 = unpack_bf16(ov[j].w);
      acc += g0.x * o0.x + g0.y * o0.y + g1.x * o1.x + g1.y * o1.y + g2.x * o2.x + g2.y * o2.y + g3.x * o3.x + g3.y * o3.y;
    }
    dl_s[r] = acc;                                // delta = rowsum(dO o O)
    n_q = normalize_row_n(smem + kFuOffQ, r);
    uint8_t* kv = smem + (sub == 0 ? kFuOffK : kFuOffV);
    n_kv[0] = normalize_row_n(kv, m);
    n_kv[1] = normalize_row_n(kv, 128 + m);
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (warp == 8) {
    if (lane == 0) {
      tc_fence_after();
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);     // K-major A and B
      const uint32_t idesc_g = make_idesc_bf16(128, kHD, 0, 1);    // K-major A (P^T / dS^T chunk), MN-major B (dO / Q rows)
      const uint32_t idesc_q = make_idesc_bf16(128, kHD, 1, 1);    // MN-major A (dS^T chunks read as dS), MN-major B (K rows)
      auto issue_scores = [&](int u) {
        const int j = u >> 2, qc = u & 3;       // query chunk (i, h) = 64 qc ..
        const uint32_t slot = (uint32_t)(u & 1) * 64;
        {
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffK + j * 16384), 0, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffQ + qc * 8192), 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + kFuColS + slot, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
        }
        {
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffV + j * 16384), 0, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffDO + qc * 8192), 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + kFuColDP + slot, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&bar_s[u & 1]);
      };
      issue_scores(0);
      issue_scores(1);
#pragma unroll 1
      for (int u = 0; u < NU; ++u) {
        const int j = u >> 2, qc = u & 3, i = qc >> 1, h = qc & 1;
        mbar_wait_bounded(&p_ready[u & 1], (u >> 1) & 1);
        tc_fence_after();
        if (u + 2 < NU) issue_scores(u + 2);
        if (u == 4) {                                // dVn_0 / dKn_0 must have left TMEM before block 1 overwrites them
          mbar_wait_bounded(kv_drained, 0);
          tc_fence_after();
        }
        const uint64_t ap_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffPT + (u & 1) * kPChunkBytes), 0, 1024);
        const uint64_t as_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffDST + (u & 3) * kPChunkBytes), 0, 1024);
        const uint64_t bdo_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffDO + qc * 8192), 64 * 128, 1024);
        const uint64_t bq_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffQ + qc * 8192), 64 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dVn_j += P^T_u dO_u
          umma_bf16(tmem_base + kFuColDV, ap_desc + (uint64_t)(k * 2), bdo_desc + (uint64_t)(k * 128), idesc_g, (qc | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dKn_j += dS^T_u Qn_u
          umma_bf16(tmem_base + kFuColDK, as_desc + (uint64_t)(k * 2), bq_desc + (uint64_t)(k * 128), idesc_g, (qc | k) != 0 ? 1u : 0u);
        if (h == 1) {                 // dQn_i += dS_(j,i) Kn_j : A = chunks (i,0), (i,1) of this block, 16 KB apart
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffDST + ((u & 3) - 1) * kPChunkBytes), kPChunkBytes, 1024);
          const uint64_t bk_desc = make_smem_desc_sw128(smem_u32(smem + kFuOffK + j * 16384), 64 * 128, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + kFuColDQ + (uint32_t)i * 64, a_desc + (uint64_t)(k * 128), bk_desc + (uint64_t)(k * 128), idesc_q,
                      (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&mma_done[u & 1]);
      }
    }
  } else {
    const int q = warp & 3;
    const int sub = warp >> 2;                        // which 32 queries of each 64-query chunk this thread converts
    const int m = q * 32 + lane;                      // key row within the block = TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sc = scale * kLog2eA;
    __nv_bfloat16* gq_base = g_qkv + (long long)b * S * 3 * C + head * kHD;
    // sub 0 finishes dK rows, sub 1 dV rows (block j: key 128 j + m)
    auto drain_kv = [&](int j) {
      uint32_t g0[32], g1[32];
      const uint32_t col = sub == 0 ? kFuColDK : kFuColDV;
      tmem_ld32(t_row + col, g0);
      tmem_ld32(t_row + col + 32, g1);
      tmem_ld_wait();
      if (j == 0) {
        tc_fence_before();
        mbar_arrive(kv_drained);
      }
      const int key = 128 * j + m;
      norm_adjoint_store(g0, g1, smem + (sub == 0 ? kFuOffK : kFuOffV), key, n_kv[j],
                         gq_base + (long long)key * 3 * C + (sub == 0 ? C : 2 * C));
    };
#pragma unroll 1
    for (int u = 0; u < NU; ++u) {
      const int qc = u & 3;
      uint8_t* pbuf = smem + kFuOffPT + (u & 1) * kPChunkBytes;
      uint8_t* sbuf = smem + kFuOffDST + (u & 3) * kPChunkBytes;
      mbar_wait_bounded(&bar_s[u & 1], (u >> 1) & 1);
      tc_fence_after();
      uint32_t rs[32], rp[32];
      tmem_ld32(t_row + kFuColS + (uint32_t)(u & 1) * 64 + sub * 32, rs);
      tmem_ld32(t_row + kFuColDP + (uint32_t)(u & 1) * 64 + sub * 32, rp);
      tmem_ld_wait();
      // the ring slots written below were last read by the MMAs of unit u - 2 (P^T) and of the block two back (dS^T)
      if (u >= 2) mbar_wait_bounded(&mma_done[u & 1], ((u - 2) >> 1) & 1);
      const float* lq = lse_s + qc * 64 + sub * 32;
      const float* dq_ = dl_s + qc * 64 + sub * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float pv[8], dv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = g * 8 + e;
          pv[e] = exp2f(fmaf(__uint_as_float(rs[c]), sc, -lq[c]));
          dv[e] = pv[e] * (__uint_as_float(rp[c]) - dq_[c]) * scale;
        }
        uint4 o;
        o.x = pack_bf16(pv[0], pv[1]); o.y = pack_bf16(pv[2], pv[3]); o.z = pack_bf16(pv[4], pv[5]); o.w = pack_bf16(pv[6], pv[7]);
        *prow(pbuf, m, sub * 4 + g) = o;
        o.x = pack_bf16(dv[0], dv[1]); o.y = pack_bf16(dv[2], dv[3]); o.z = pack_bf16(dv[4], dv[5]); o.w = pack_bf16(dv[6], dv[7]);
        *prow(sbuf, m, sub * 4 + g) = o;
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[u & 1]);
      if (u == 4) {   // key block 0 is complete once the MMAs of unit 3 are; by now (one unit later) they have retired
        mbar_wait_bounded(&mma_done[1], 1);
        tc_fence_after();
        drain_kv(0);
      }
    }
    mbar_wait_bounded(&mma_done[1], 1);               // unit 7: parity (7 >> 1) & 1
    tc_fence_after();
    drain_kv(1);
    {  // dQ: sub 0 finishes queries 0..127 (accumulator 0), sub 1 queries 128..255 (accumulator 1)
      uint32_t g0[32], g1[32];
      tmem_ld32(t_row + kFuColDQ + (uint32_t)sub * 64, g0);
      tmem_ld32(t_row + kFuColDQ + (uint32_t)sub * 64 + 32, g1);
      tmem_ld_wait();
      const int qr = 128 * sub + m;
      norm_adjoint_store(g0, g1, smem + kFuOffQ, qr, n_q, gq_base + (long long)qr * 3 * C);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_bwd_fused(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* g_y, const float* lse,
                     __nv_bfloat16* g_qkv, int B, int heads, cudaStream_t stream) {
  constexpr int S = 256;
  const int C = heads * kHD;
  CUtensorMap t_all, t_do_all;
  {
    uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)B * S};
    uint64_t strides[1] = {(uint64_t)3 * C * 2};
    uint32_t box[2] = {64, (uint32_t)S};
    if (encode_tmap(&t_all, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)B * S};
    uint64_t strides[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {64, (uint32_t)S};
    if (encode_tmap(&t_do_all, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g_y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  }
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(attn_bwd_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmem));
  }
  TEDM_CUDA(launch_pdl(attn_bwd_fused_tc_kernel, B * heads, kThreadsB, kFuSmem, stream, t_all, t_do_all, y, lse, g_qkv, heads,
                       1.0f / sqrtf((float)kHD)));
  return 0;
}

template <int S>
int launch_bwd_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* g_y, const float* lse, float* delta,
                  __nv_bfloat16* g_qkv, int B, int heads, cudaStream_t stream) {
  const int C = heads * kHD;
  CUtensorMap t_tile, t_all, t_do_tile, t_do_all;
  {
    uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)B * S};
    uint64_t strides[1] = {(uint64_t)3 * C * 2};
    uint32_t box_t[2] = {64, (uint32_t)(S == 256 ? 128 : 64)};
    uint32_t box_a[2] = {64, (uint32_t)S};
    if (encode_tmap(&t_tile, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box_t, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
    if (encode_tmap(&t_all, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box_a, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)B * S};
    uint64_t strides[1] = {(uint64_t)C * 2};
    uint32_t box_t[2] = {64, (uint32_t)(S == 256 ? 128 : 64)};
    uint32_t box_a[2] = {64, (uint32_t)S};
    if (encode_tmap(&t_do_tile, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g_y, dims, strides, box_t, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
    if (encode_tmap(&t_do_all, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g_y, dims, strides, box_a, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  }
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem));
    TEDM_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kKvSmem));
  }
  const int n_pairs = B * heads;
  const int grid = S == 256 ? n_pairs * 2 : (n_pairs + 1) / 2;
  const float scale = 1.0f / sqrtf((float)kHD);
  launch_pdl(attn_bwd_dq_tc_kernel<S>, grid, kThreadsB, kDqSmem, stream, t_tile, t_all, t_do_tile, y, lse, delta, g_qkv, n_pairs, heads, scale);
  TEDM_LAUNCH_CHECK();
  launch_pdl(attn_bwd_dkv_tc_kernel<S>, grid, kThreadsB, kKvSmem, stream, t_tile, t_all, t_do_all, lse, delta, g_qkv, n_pairs, heads, scale);
  TEDM_LAUNCH_CHECK();
  return 0;
}

template <int S>
int launch_tc(const __nv_bfloat16* qkv, __nv_bfloat16* y, float* lse, int B, int heads, cudaStream_t stream) {
  const int C = heads * kHD;
  CUtensorMap tq, tkv;
  uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)B * S};
  uint64_t strides[1] = {(uint64_t)3 * C * 2};
  uint32_t box_q[2] = {64, (uint32_t)(S == 256 ? kTileQ : 64)};   // 128 queries of a head (S = 256) / one pair (S = 64)
  uint32_t box_kv[2] = {64, (uint32_t)S};                          // all keys of a head
  if (encode_tmap(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  if (encode_tmap(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesA));
  }
  const int n_pairs = B * heads;
  const int grid = S == 256 ? n_pairs * 2 : (n_pairs + 1) / 2;
  launch_pdl(attn_fwd_tc_kernel<S>, grid, kThreadsA, kSmemBytesA, stream, tq, tkv, y, lse, n_pairs, heads, 1.0f / sqrtf((float)kHD));
  TEDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool attention_tc_supported(int S, int hd) { return hd == kHD && (S == 256 || S == 64); }

int attention_backward_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* g_y, const float* lse,
                          float* delta, __nv_bfloat16* g_qkv, int B, int S, int heads, cudaStream_t stream) {
  // TEDM_ATTN_BWD_FUSED=0 falls back to the two-kernel version (A/B switch)
  static const bool fused = [] { const char* e = getenv("TEDM_ATTN_BWD_FUSED"); return !(e != nullptr && e[0] == '0'); }();
  if (S == 256 && fused) return launch_bwd_fused(qkv, y, g_y, lse, g_qkv, B, heads, stream);
  if (S == 256) return launch_bwd_tc<256>(qkv, y, g_y, lse, delta, g_qkv, B, heads, stream);
  return launch_bwd_tc<64>(qkv, y, g_y, lse, delta, g_qkv, B, heads, stream);
}

int attention_forward_tc(const __nv_bfloat16* qkv, __nv_bfloat16* y, float* lse, int B, int S, int heads, cudaStream_t stream) {
  TEDM_CHECK((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
             "attention: pointers must be 16-byte aligned");
  if (S == 256) return launch_tc<256>(qkv, y, lse, B, heads, stream);
  return launch_tc<64>(qkv, y, lse, B, heads, stream);
}

}  // namespace tedm
