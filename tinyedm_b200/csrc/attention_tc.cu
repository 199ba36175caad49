// Cosine-normalised self-attention forward on tcgen05/TMEM (sm_100a), head_dim 64, S = 256 or 64 keys per head.
//
// Reference: CosineAttention.forward, src/tinyedm/networks.py:191-207 — pixel_norm over hd of q, k and v (:195), then
// F.scaled_dot_product_attention(q, k, v) with scale 1/sqrt(hd) (:201), output channel = head*hd + d (:202).
// qkv is (B,S,3C) bf16 with channel = {q,k,v}*C + head*hd + d (the weight bank permutes the qkv conv's rows).
//
// One CTA owns a SLAB of 256 query rows:
//     S = 256 : one (image, head): 2 M-tiles of 128 queries, both attend to the same 256 keys;
//     S =  64 : four consecutive (image, head) pairs: 2 M-tiles of 2 pairs each, each tile has its own 128 keys and a
//               block-diagonal mask (a query only sees the 64 keys of its own pair).
// Pipeline per CTA (9 warps):
//     warp 8      TMA: Q, K, V slabs (3 x 32 KB, 128B-swizzled rows of 64 bf16)                      -> bar_load
//     warps 0..7  pixel_norm of the 768 rows in shared memory (one row per thread and tensor)         -> fence, sync
//     warp 8      tcgen05.mma  S_t = Q_t K_t^T   (M=128, N=keys, K=64)  fp32 in TMEM                  -> bar_s[t]
//     warps 4t..  softmax of tile t straight out of TMEM (thread = query row): max, exp2, sum;
//                 unnormalised P (bf16) -> swizzled shared memory (the A operand of the next MMA)     -> p_ready[t]
//     warp 8      tcgen05.mma  O_t = P_t V_t     (M=128, N=64, K=keys; V is the MN-major B operand),
//                 accumulating into the TMEM columns S_t no longer needs                              -> bar_o[t]
//     warps 4t..  O / rowsum -> bf16 -> y, log-sum-exp -> lse
// The S x S score matrix exists only in TMEM (512 columns = 2 tiles x 256 keys) and, as bf16 P, in shared memory.
#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr int kHD = 64;
constexpr int kSlabRows = 256;
constexpr int kSlabBytes = kSlabRows * 128;                 // 32 KB: 256 rows of 64 bf16
constexpr int kPChunkBytes = 128 * 128;                     // [128 queries][64 keys] bf16
constexpr int kOffQ = 0, kOffK = kSlabBytes, kOffV = 2 * kSlabBytes, kOffP = 3 * kSlabBytes;
constexpr int kOffBarsA = kOffP + 2 * 4 * kPChunkBytes;     // two tiles x up to 4 key chunks
constexpr int kSmemBytesA = kOffBarsA + 128 + 1024;
constexpr int kThreadsA = 288;
constexpr float kEpsA = 1e-4f;
constexpr float kLog2eA = 1.4426950408889634f;
constexpr float kLn2A = 0.6931471805599453f;
static_assert(kSmemBytesA <= 232448, "shared memory budget");

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

// NK = keys per M-tile (256 for S = 256, 128 for S = 64), PAIR = rows per (image, head) pair (= S)
template <int S>
__global__ void __launch_bounds__(kThreadsA, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ y, float* __restrict__ lse,
                   int n_pairs, int heads, float scale) {
  constexpr int NK = S == 256 ? 256 : 128;
  constexpr int PAIRS = kSlabRows / S;              // (image, head) pairs per CTA
  constexpr int NCH = NK / 64;                      // 64-key chunks of P per tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBarsA);
  uint64_t* bar_load = bars;          // TMA bytes
  uint64_t* bar_s = bars + 1;         // [2] scores of tile t in TMEM
  uint64_t* p_ready = bars + 3;       // [2] 128 softmax threads of tile t wrote P (and are done with S_t)
  uint64_t* bar_o = bars + 5;         // [2] O_t in TMEM
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = heads * kHD;
  const int pair0 = blockIdx.x * PAIRS;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      mbar_init(bar_load, 1);
      for (int t = 0; t < 2; ++t) {
        mbar_init(&bar_s[t], 1);
        mbar_init(&p_ready[t], 128);
        mbar_init(&bar_o[t], 1);
      }
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

  if (warp == 8) {
    if (lane == 0) {
      // ---- loads: per pair, S rows x 64 channels of q, k and v (rows of image b, channel block of the head) ----
      mbar_expect_tx(bar_load, 3 * kSlabBytes);
      for (int pp = 0; pp < PAIRS; ++pp) {
        int pair = pair0 + pp;
        if (pair >= n_pairs) pair = n_pairs - 1;          // tail CTA: duplicate the last pair (results discarded)
        const int b = pair / heads, head = pair - b * heads;
        const int row = b * S;
        for (int part = 0; part < 3; ++part)
          tma_load_2d(smem + part * kSlabBytes + pp * S * 128, &tmap_qkv, bar_load, part * C + head * kHD, row);
      }
    }
  } else {
    // ---- pixel_norm of q, k, v rows (thread i: row i of each slab) ----
    mbar_wait_bounded(bar_load, 0);
    const int r = threadIdx.x;
    normalize_row(smem + kOffQ, r);
    normalize_row(smem + kOffK, r);
    normalize_row(smem + kOffV, r);
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (warp == 8) {
    if (lane == 0) {
      tc_fence_after();
      // ---- S_t = Q_t K_t^T ----
      const uint32_t idesc_s = make_idesc_bf16(128, NK, 0, 0);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t q_addr = smem_u32(smem + kOffQ + t * 128 * 128);
        const uint32_t k_addr = smem_u32(smem + kOffK + (S == 256 ? 0 : t * 128 * 128));
        const uint64_t a_desc = make_smem_desc_sw128(q_addr, 0, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(k_addr, 0, 1024);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          umma_bf16(tmem_base + t * 256, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar_s[t]);
      }
      // ---- O_t = P_t V_t ----
      const uint32_t idesc_o = make_idesc_bf16(128, kHD, 0, 1);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait_bounded(&p_ready[t], 0);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + kOffV + (S == 256 ? 0 : t * 128 * 128));
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const uint32_t p_addr = smem_u32(smem + kOffP + (t * 4 + c) * kPChunkBytes);
          const uint64_t a_desc = make_smem_desc_sw128(p_addr, 0, 1024);
          // V rows (keys) 64c .. 64c+63: MN-major B operand, 16 keys (K) = 16 rows of 128 B per MMA
          const uint64_t b_desc = make_smem_desc_sw128(v_addr + c * 64 * 128, 64 * 128, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + t * 256, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 128), idesc_o,
                      (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(&bar_o[t]);
      }
    }
  } else {
    // ---- softmax + output of tile t (warps 4t .. 4t+3; thread = query row) ----
    const int t = warp >> 2, q = warp & 3;
    const int m = q * 32 + lane;                       // row within the tile
    const int slab_row = t * 128 + m;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256;
    // keys this row may see: all NK for S = 256; the 64 of its own pair for S = 64
    const int kb = S == 256 ? 0 : (m >> 6) * 64;
    const int kn = S == 256 ? NK : 64;
    const float sc = scale * kLog2eA;
    mbar_wait_bounded(&bar_s[t], 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < kn; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(t_row + kb + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
    const float mxs = mx * sc;
    float sum = 0.f;
    uint8_t* pbuf = smem + kOffP + t * 4 * kPChunkBytes;
#pragma unroll 1
    for (int c0 = 0; c0 < NK; c0 += 32) {
      uint4* dst[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) dst[g] = prow(pbuf + (c0 >> 6) * kPChunkBytes, m, ((c0 & 63) >> 3) + g);
      if (c0 >= kb && c0 < kb + kn) {
        uint32_t r[32];
        tmem_ld32(t_row + c0, r);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pv[i] = exp2f(fmaf(__uint_as_float(r[i]), sc, -mxs));
          // the MMA consumes bf16 probabilities: the row sum uses the same rounded values (as SDPA's bf16 P does)
          pv[i] = bf16_round(pv[i]);
          sum += pv[i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16(pv[g * 8 + 0], pv[g * 8 + 1]); o.y = pack_bf16(pv[g * 8 + 2], pv[g * 8 + 3]);
          o.z = pack_bf16(pv[g * 8 + 4], pv[g * 8 + 5]); o.w = pack_bf16(pv[g * 8 + 6], pv[g * 8 + 7]);
          *dst[g] = o;
        }
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) *dst[g] = make_uint4(0, 0, 0, 0);   // masked keys of the other pair
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    mbar_arrive(&p_ready[t]);
    // ---- epilogue ----
    mbar_wait_bounded(&bar_o[t], 0);
    tc_fence_after();
    const int pair = pair0 + slab_row / S;
    const int s_idx = slab_row % S;
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
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int S>
int launch_tc(const __nv_bfloat16* qkv, __nv_bfloat16* y, float* lse, int B, int heads, cudaStream_t stream) {
  const int C = heads * kHD;
  CUtensorMap tm;
  uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)B * S};
  uint64_t strides[1] = {(uint64_t)3 * C * 2};
  uint32_t box[2] = {64, (uint32_t)S};
  if (encode_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesA));
    configured = true;
  }
  const int n_pairs = B * heads;
  const int per_cta = kSlabRows / S;
  const int grid = (n_pairs + per_cta - 1) / per_cta;
  attn_fwd_tc_kernel<S><<<grid, kThreadsA, kSmemBytesA, stream>>>(tm, y, lse, n_pairs, heads, 1.0f / sqrtf((float)kHD));
  TEDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool attention_tc_supported(int S, int hd) { return hd == kHD && (S == 256 || S == 64); }

int attention_forward_tc(const __nv_bfloat16* qkv, __nv_bfloat16* y, float* lse, int B, int S, int heads, cudaStream_t stream) {
  TEDM_CHECK((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
             "attention: pointers must be 16-byte aligned");
  if (S == 256) return launch_tc<256>(qkv, y, lse, B, heads, stream);
  return launch_tc<64>(qkv, y, lse, B, heads, stream);
}

}  // namespace tedm
