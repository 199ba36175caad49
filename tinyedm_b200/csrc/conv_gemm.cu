// MPConv as an implicit-GEMM convolution on tcgen05/TMEM, fed by TMA (sm_100a).
//
// Replaces F.conv2d(x, w_hat, padding="same") of the reference's Conv2d (src/tinyedm/networks.py:31-38)
// for both the forward pass and the data gradient (dgrad == the same stride-1 "same" convolution run
// with the spatially flipped, channel-transposed weight that weight_prep.cu emits).
//
// Layout: activations NHWC bf16 (B,H,W,C), C % 64 == 0; weights bf16 [Cout][tap][Cin] (K-major rows).
// GEMM view: D[M = pixels, N = Cout] = sum_{tap, ci} A[pixel + off(tap), ci] * Wt[Cout][tap*Cin + ci].
//
// One persistent CTA per SM. Warp 0 = TMA producer, warp 1 = MMA issuer (single thread), warp 2 = TMEM
// allocator, warps 4..7 = epilogue (TMEM -> registers -> fused epilogue -> global). The M tile is
// NB images x RH rows x W columns (<= 128 pixels) fetched by ONE 4-D TMA box per (tap, 64-channel
// slice): the halo / zero padding comes for free from TMA out-of-bounds zero fill (negative or
// too-large W/H coordinates). Accumulators are double buffered in TMEM (2 x BN columns) so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 bf16 = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kNumThreads = 384;     // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue (2 per TMEM lane quarter)
constexpr int kEpiWarps = 8;

template <int BN>
struct GemmCfg {
  static constexpr int kBTileBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*row dots*/;
};

struct Tile {
  int b0, h0, n0;
  long long p_base;  // first output pixel (global pixel index) of the tile
  long long p_limit; // pixels >= p_limit are outside the tile's images
};

__device__ __forceinline__ Tile decode_tile(const ConvGemmParams& p, int tile) {
  Tile t;
  int mt = tile / p.n_tiles;
  int nt = tile - mt * p.n_tiles;
  t.n0 = nt * p.block_n;
  if (p.NB == 1) {
    t.b0 = mt / p.tiles_h;
    t.h0 = (mt - t.b0 * p.tiles_h) * p.RH;
    t.p_limit = (long long)(t.b0 + 1) * p.H * p.W;
  } else {
    t.b0 = mt * p.NB;
    t.h0 = 0;
    t.p_limit = (long long)p.B * p.H * p.W;
  }
  t.p_base = ((long long)t.b0 * p.H + t.h0) * p.W;
  return t;
}

// 32 values per lane -> lane j ends up with the sum over all 32 lanes of value j (31 shuffles, no shared memory).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// pulls `bytes` of a row into L2 (issued one tile ahead of its use)
__device__ __forceinline__ void prefetch_row_l2(const void* ptr, int bytes) {
  const char* c = static_cast<const char*>(ptr);
  for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + o));
}

template <int CW>
__device__ __forceinline__ void row_load(const __nv_bfloat16* src, float (&f)[CW]) {
#pragma unroll
  for (int g = 0; g < CW / 8; ++g) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + g * 8);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[g * 8 + 0] = a.x; f[g * 8 + 1] = a.y; f[g * 8 + 2] = b.x; f[g * 8 + 3] = b.y;
    f[g * 8 + 4] = c.x; f[g * 8 + 5] = c.y; f[g * 8 + 6] = d.x; f[g * 8 + 7] = d.y;
  }
}

template <int CW>
__device__ __forceinline__ void row_store(__nv_bfloat16* dst, const float (&v)[CW], int n_valid) {
#pragma unroll
  for (int g = 0; g < CW / 8; ++g) {
    if (g * 8 < n_valid) {
      uint4 o;
      o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
      o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
      o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
      o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
      *reinterpret_cast<uint4*>(dst + g * 8) = o;
    }
  }
}

template <int CW>
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&r)[CW]);
template <>
__device__ __forceinline__ void tmem_ld_cw<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cw<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }

// v[i] *= keep(e0 + i) / (1 - p) for the CW consecutive elements starting at the even element index e0
template <int CW>
__device__ __forceinline__ void dropout_apply(float (&v)[CW], unsigned long long e0, float drop_p, uint32_t dseed) {
  const float keep_scale = 1.0f / (1.0f - drop_p);
  const uint32_t thresh = (uint32_t)(drop_p * 65536.0f);
#pragma unroll
  for (int i = 0; i < CW; i += 2) {
    const uint32_t bits = dropout_bits2(e0 + i, dseed);
    v[i] = (bits & 0xFFFFu) >= thresh ? v[i] * keep_scale : 0.f;
    v[i + 1] = (bits >> 16) >= thresh ? v[i + 1] * keep_scale : 0.f;
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const ConvGemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform: role loops stay in uniform registers
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_wait();                       // programmatic dependent launch: the prologue above overlaps the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) pdl_trigger();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t a_bytes = (uint32_t)p.NB * p.RH * p.W * kBlockK * 2;
    const uint32_t tx_bytes = a_bytes + Cfg::kBTileBytes;
    const int kc_per_tap = p.Cin / kBlockK;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      Tile t = decode_tile(p, tile);
      for (int tap = 0; tap < p.taps; ++tap) {
        const int dr = (p.taps == 9) ? tap / 3 - 1 : 0;
        const int ds = (p.taps == 9) ? tap % 3 - 1 : 0;
        for (int kc = 0; kc < kc_per_tap; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader_lane) {
            uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
            uint8_t* b_dst = a_dst + kATileBytes;
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_4d(a_dst, &tmap_a, &full_bar[stage], kc * kBlockK, ds, t.h0 + dr, t.b0);
            tma_load_2d(b_dst, &tmap_b, &full_bar[stage], tap * p.Cin + kc * kBlockK, t.n0);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t smem_base = smem_u32(smem);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      Tile t = decode_tile(p, tile);
      int n_this = p.Cout - t.n0;
      if (n_this > BN) n_this = BN;
      const uint32_t idesc = make_idesc_bf16(kBlockM, n_this, 0, 0);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader_lane) {
          const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr, 0, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + kATileBytes, 0, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (leader_lane) umma_commit(&tmem_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =====================
    constexpr int CW = (EPI == EPI_SILU_BWD) ? 16 : 32;   // columns per step (register budget of the heavy adjoint)
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2; // which half of the tile's columns this warp owns
    const int m = q * 32 + lane;      // accumulator row == pixel within the tile
    const int rows_in_tile = p.NB * p.RH * p.W;
    const int HW = p.H * p.W;
    int acc = 0;
    uint32_t acc_phase = 0;
    // dropout seed = per-launch salt + a device-resident step counter (fresh masks under CUDA-graph replay)
    unsigned long long seed64 = ((unsigned long long)p.seed_hi << 32) | p.seed_lo;
    if (p.seed_ptr != nullptr) seed64 += *p.seed_ptr * 0x9E3779B97F4A7C15ull;
    const uint32_t dseed = dropout_seed((uint32_t)seed64, (uint32_t)(seed64 >> 32));
    // a warp's 32 rows belong to ONE image when tiles never mix images or images are a multiple of 32 pixels
    const bool warp_one_image = (p.NB == 1) || (HW % 32 == 0);
    constexpr bool kNeedAux = (EPI == EPI_MODSILU_BWD || EPI == EPI_SILU_BWD);
    constexpr bool kMayRes = (EPI == EPI_AXPBY || EPI == EPI_SILU_BWD);
    const bool use_res = kMayRes && p.res != nullptr;
    const bool use_old = (EPI == EPI_SILU_BWD) && p.accumulate_out;
    float* row_dots = reinterpret_cast<float*>(tmem_ptr + 4);   // [2 parities][2 halves][128 rows]
    int parity = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      Tile t = decode_tile(p, tile);
      int n_this = p.Cout - t.n0;
      if (n_this > BN) n_this = BN;
      // column range of this warp: [cb, ce)
      int cs = ((n_this / 2 + 31) / 32) * 32;
      if (cs > n_this) cs = n_this;
      const int cb = half == 0 ? 0 : cs, ce = half == 0 ? cs : n_this;
      const long long pix = t.p_base + m;
      const bool valid = (m < rows_in_tile) && (pix < t.p_limit);
      const int b = valid ? (int)(pix / HW) : 0;
      const long long row_off = pix * p.Cout + t.n0;
      __nv_bfloat16* out_row = p.out + row_off;
      // epilogue operands live in HBM: start the NEXT tile's rows towards L2 now (a whole tile of lead time)
      if ((kNeedAux || use_res || use_old) && tile + (int)gridDim.x < total_tiles) {
        const Tile tn = decode_tile(p, tile + gridDim.x);
        const long long pn = tn.p_base + m;
        if (m < rows_in_tile && pn < tn.p_limit) {
          int nn = p.Cout - tn.n0;
          if (nn > BN) nn = BN;
          int csn = ((nn / 2 + 31) / 32) * 32;
          if (csn > nn) csn = nn;
          const int cbn = half == 0 ? 0 : csn, cen = half == 0 ? csn : nn;
          const long long off = pn * p.Cout + tn.n0 + cbn;
          if (cen > cbn) {
            if (kNeedAux) prefetch_row_l2(p.aux + off, (cen - cbn) * 2);
            if (use_res) prefetch_row_l2(p.res + off, (cen - cbn) * 2);
            if (use_old) prefetch_row_l2(p.out + off, (cen - cbn) * 2);
          }
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;

      float inv_n = 1.0f, kk = 0.f;
      if constexpr (EPI == EPI_SILU_BWD) if (p.nrm != nullptr) {
        // sweep 1 of the fused pixel-norm adjoint: dot = sum_c v*x with v = alpha*acc*silu'(x) + beta*res; the two
        // column halves of a row live in two warps: partial dots are exchanged through shared memory
        float dot = 0.f;
        for (int c0 = cb; c0 < ce; c0 += CW) {
          uint32_t r[CW];
          tmem_ld_cw<CW>(t_row + c0, r);
          tmem_ld_wait();
          if (valid) {
            float xv[CW], rv[CW];
            row_load<CW>(p.aux + row_off + c0, xv);
            if (use_res) row_load<CW>(p.res + row_off + c0, rv);
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              float v = __uint_as_float(r[i]) * p.alpha * mp_silu_grad_f(xv[i]);
              if (use_res) v += p.beta * rv[i];
              dot += v * xv[i];
            }
          }
        }
        float* slot = row_dots + parity * 256;
        slot[half * 128 + m] = dot;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // the two warps sharing this lane quarter
        dot = slot[m] + slot[128 + m];
        const float n = valid ? p.nrm[pix] : 1.0f;
        inv_n = 1.0f / n;
        kk = dot / (fmaxf(n - 1e-4f, 1e-20f) * (float)p.Cout);
      }
      for (int c0 = cb; c0 < ce; c0 += CW) {
        uint32_t r[CW];
        tmem_ld_cw<CW>(t_row + c0, r);
        tmem_ld_wait();
        float v[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
        if constexpr (EPI == EPI_MODSILU_BWD) {
          // backward of h = drop(mp_silu(raw * m)) fused into the data gradient of the next conv:
          //   gz = g_h * keep/(1-p) * mp_silu'(raw*m);  g_raw = gz * m;  d_mod[b,c] += sum_pixels gz * raw
          float dm[CW];
          if (valid) {
            float rw[CW];
            row_load<CW>(p.aux + row_off + c0, rw);
            const float* mrow = p.mod + (long long)b * p.mod_stride + t.n0 + c0;
            if (p.drop_p > 0.f) dropout_apply<CW>(v, (unsigned long long)pix * p.Cout + t.n0 + c0, p.drop_p, dseed);
#pragma unroll
            for (int g = 0; g < CW / 4; ++g) {
              const float4 mm = *reinterpret_cast<const float4*>(mrow + g * 4);
              const float mv[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i = g * 4 + j;
                const float gz = v[i] * mp_silu_grad_f(rw[i] * mv[j]);
                v[i] = gz * mv[j];
                dm[i] = gz * rw[i];
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < CW; ++i) dm[i] = 0.f;
          }
          if (warp_one_image) {
            const long long pw = t.p_base + q * 32;          // first pixel of this warp's rows
            const float tot = warp_transpose_reduce(dm, lane);
            if (q * 32 < rows_in_tile && pw < t.p_limit && c0 + lane < n_this)
              atomicAdd(p.d_mod + (long long)(pw / HW) * p.mod_stride + t.n0 + c0 + lane, tot);
          } else {
            // rows of several images in one warp: one masked transpose-reduce per image instead of CW atomics per row
            const int b_lo = __reduce_min_sync(0xffffffffu, valid ? b : 0x7fffffff);
            const int b_hi = __reduce_max_sync(0xffffffffu, valid ? b : -1);
            for (int bb = b_lo; bb <= b_hi; ++bb) {
              float part[CW];
#pragma unroll
              for (int i = 0; i < CW; ++i) part[i] = (valid && b == bb) ? dm[i] : 0.f;
              const float tot = warp_transpose_reduce(part, lane);
              if (c0 + lane < n_this) atomicAdd(p.d_mod + (long long)bb * p.mod_stride + t.n0 + c0 + lane, tot);
            }
          }
        }
        if (valid) {
          if constexpr (EPI == EPI_MODSILU) {
            const float* mrow = p.mod + (long long)b * p.mod_stride + t.n0 + c0;
            // the reference's conv output is bf16 before the fp32 modulation island (networks.py:253-258);
            // the stored (bf16) pre-modulation value is also what backward sees
#pragma unroll
            for (int i = 0; i < CW; ++i) v[i] = bf16_round(v[i]);
            if (p.out2 != nullptr) row_store<CW>(p.out2 + row_off + c0, v, n_this - c0);
#pragma unroll
            for (int g = 0; g < CW / 4; ++g) {
              if (c0 + g * 4 < n_this) {
                const float4 mm = *reinterpret_cast<const float4*>(mrow + g * 4);
                v[g * 4 + 0] = mp_silu_f(v[g * 4 + 0] * mm.x);
                v[g * 4 + 1] = mp_silu_f(v[g * 4 + 1] * mm.y);
                v[g * 4 + 2] = mp_silu_f(v[g * 4 + 2] * mm.z);
                v[g * 4 + 3] = mp_silu_f(v[g * 4 + 3] * mm.w);
              }
            }
            if (p.drop_p > 0.f) dropout_apply<CW>(v, (unsigned long long)pix * p.Cout + t.n0 + c0, p.drop_p, dseed);
          } else if constexpr (EPI == EPI_AXPBY) {
            float rv[CW];
            row_load<CW>(p.res + row_off + c0, rv);
            const float rs = p.nrm != nullptr ? p.beta / p.nrm[pix] : p.beta;   // nrm: res is the un-normalised tensor
#pragma unroll
            for (int i = 0; i < CW; ++i) v[i] += rs * rv[i];
          } else if constexpr (EPI == EPI_SILU_BWD) {
            // g_x = alpha*acc * mp_silu'(x) + beta*res, [pixel-norm adjoint], (+ what is already in `out`)
            float xv[CW];
            row_load<CW>(p.aux + row_off + c0, xv);
#pragma unroll
            for (int i = 0; i < CW; ++i) v[i] *= mp_silu_grad_f(xv[i]);
            if (use_res) {
              float rv[CW];
              row_load<CW>(p.res + row_off + c0, rv);
#pragma unroll
              for (int i = 0; i < CW; ++i) v[i] += p.beta * rv[i];
            }
            if (p.nrm != nullptr) {
#pragma unroll
              for (int i = 0; i < CW; ++i) v[i] = v[i] * inv_n - xv[i] * kk;
            }
            if (use_old) {
              float ov[CW];
              row_load<CW>(out_row + c0, ov);
#pragma unroll
              for (int i = 0; i < CW; ++i) v[i] += ov[i];
            }
            if (p.out_bias != nullptr) {
              const float* brow = p.out_bias + (long long)b * p.Cout + t.n0 + c0;
#pragma unroll
              for (int i = 0; i < CW; ++i)
                if (c0 + i < n_this) v[i] = fmaf(p.out_bias_scale, brow[i], v[i]);
            }
          }
          row_store<CW>(out_row + c0, v, n_this - c0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      parity ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int EPI>
int launch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const ConvGemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::kSmemBytes));
  }
  int tiles = p.m_tiles * p.n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  TEDM_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, EPI>, ta, tb, p));
  return 0;
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const ConvGemmParams& p, cudaStream_t stream) {
  switch (p.epi) {
    case EPI_PLAIN: return launch_epi<BN, EPI_PLAIN>(ta, tb, p, stream);
    case EPI_MODSILU: return launch_epi<BN, EPI_MODSILU>(ta, tb, p, stream);
    case EPI_AXPBY: return launch_epi<BN, EPI_AXPBY>(ta, tb, p, stream);
    case EPI_MODSILU_BWD: return launch_epi<BN, EPI_MODSILU_BWD>(ta, tb, p, stream);
    case EPI_SILU_BWD: return launch_epi<BN, EPI_SILU_BWD>(ta, tb, p, stream);
    default: return fail("conv_gemm: unknown epilogue %d", p.epi);
  }
}

}  // namespace

// Picks the M-tile geometry for an (H, W) feature map: NB images x RH rows x W columns <= 128 pixels.
int conv_tile_geometry(int H, int W, int* RH, int* NB) {
  if (W > kBlockM || W < 1 || H < 1) return -1;
  if (H * W >= kBlockM) {
    *NB = 1;
    int rh = kBlockM / W;
    if (rh > H) rh = H;
    *RH = rh;
  } else {
    *RH = H;
    *NB = kBlockM / (H * W);
  }
  return 0;
}

// 0 = single-CTA kernel only, 1 = CTA-pair kernel wherever it applies (default), set by TEDM_CONV_PAIR
static int conv_pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("TEDM_CONV_PAIR");
    mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return mode;
}

bool conv_split_supported(const ConvGemmArgs& a) {
  return conv_pair_mode() == 1 && a.B > 0 && a.nrm == nullptr && a.epi == EPI_SILU_BWD && a.split_c > 0 && a.split_c % 64 == 0 && a.split_c < a.Cout &&
         conv_pair_supported(a);
}

static bool conv_takes_pair(const ConvGemmArgs& a) {
  return a.block_n_override == 0 && conv_pair_mode() == 1 && a.B > 0 && conv_pair_supported(a) &&
         ((a.epi == EPI_SILU_BWD && a.nrm != nullptr) || 4 * conv_pair_tiles(a) > num_sms());   // at least half of the CTA pairs get a tile
}

int conv_colsum_slots(const ConvGemmArgs& a) {
  if ((a.epi != EPI_PLAIN && a.epi != EPI_AXPBY) || !conv_takes_pair(a)) return 0;
  return conv_pair_colsum_slots(a);
}

int conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t stream) {
  TEDM_CHECK(a.ksize == 1 || a.ksize == 3, "conv_gemm: kernel size must be 1 or 3 (got %d)", a.ksize);
  if (a.split_c > 0) {   // the split epilogue exists in the CTA-pair kernel only (callers ask conv_split_supported first)
    TEDM_CHECK(conv_split_supported(a), "conv_gemm: split epilogue not available for this problem");
    return conv_pair_launch(a, stream);
  }
  if (a.col_partial != nullptr)
    TEDM_CHECK(conv_colsum_slots(a) > 0, "conv_gemm: column sums not available for this problem (ask conv_colsum_slots first)");
  if (conv_takes_pair(a)) return conv_pair_launch(a, stream);
  TEDM_CHECK(a.Cin % 64 == 0, "conv_gemm: Cin must be a multiple of 64 (got %d)", a.Cin);
  TEDM_CHECK(a.Cout % 16 == 0 && a.Cout >= 16, "conv_gemm: Cout must be a multiple of 16 (got %d)", a.Cout);
  TEDM_CHECK(a.B > 0 && a.H > 0 && a.W > 0, "conv_gemm: empty input");
  TEDM_CHECK((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
             "conv_gemm: pointers must be 16-byte aligned");
  ConvGemmParams p{};
  p.B = a.B; p.H = a.H; p.W = a.W; p.Cin = a.Cin; p.Cout = a.Cout;
  p.taps = a.ksize * a.ksize;
  TEDM_CHECK(conv_tile_geometry(a.H, a.W, &p.RH, &p.NB) == 0, "conv_gemm: unsupported spatial size %dx%d", a.H, a.W);
  p.tiles_h = (a.H + p.RH - 1) / p.RH;
  p.m_tiles = (p.NB == 1) ? a.B * p.tiles_h : (a.B + p.NB - 1) / p.NB;
  // N-tile width: the widest tile has the best operand reuse, but a grid that leaves most SMs idle (e.g. 64 tiles at
  // 8x8, B=128) is better served by narrower tiles. Cost model: waves x (MMA time ~ BN, + fixed per-tile overhead).
  int bn = 64;
  {
    long long best = -1;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      const int c = cands[i];
      if (c > 64 && a.Cout <= c / 2) continue;                 // would waste more than half of the tile
      if (a.epi == EPI_SILU_BWD && a.nrm != nullptr && a.Cout > c) continue;   // fused pixel-norm adjoint needs one N tile
      const long long tiles = (long long)p.m_tiles * ((a.Cout + c - 1) / c);
      const long long waves = (tiles + num_sms() - 1) / num_sms();
      const long long cost = waves * (c + 48);
      if (best < 0 || cost < best) { best = cost; bn = c; }
    }
  }
  if (a.block_n_override) bn = a.block_n_override;
  p.block_n = bn;
  p.n_tiles = (a.Cout + bn - 1) / bn;
  p.k_blocks = p.taps * (a.Cin / 64);
  p.epi = a.epi; p.alpha = a.alpha; p.out = a.out; p.out2 = a.out2; p.res = a.res;
  p.beta = a.beta; p.mod = a.mod; p.mod_stride = a.mod_stride;
  p.drop_p = a.drop_p; p.seed_lo = (uint32_t)a.seed; p.seed_hi = (uint32_t)(a.seed >> 32); p.seed_ptr = a.seed_ptr;
  p.aux = a.aux; p.d_mod = a.d_mod; p.nrm = a.nrm; p.accumulate_out = a.accumulate_out;
  p.out_bias = a.epi == EPI_SILU_BWD ? a.out_bias : nullptr; p.out_bias_scale = a.out_bias_scale;
  if (a.epi == EPI_MODSILU) TEDM_CHECK(a.mod != nullptr, "conv_gemm: MODSILU epilogue needs mod");
  if (a.epi == EPI_AXPBY) TEDM_CHECK(a.res != nullptr, "conv_gemm: AXPBY epilogue needs res");
  if (a.epi == EPI_MODSILU_BWD)
    TEDM_CHECK(a.mod != nullptr && a.aux != nullptr && a.d_mod != nullptr, "conv_gemm: MODSILU_BWD epilogue needs mod, raw and d_mod");
  if (a.epi == EPI_SILU_BWD) {
    TEDM_CHECK(a.aux != nullptr, "conv_gemm: SILU_BWD epilogue needs x");
    TEDM_CHECK(a.nrm == nullptr || a.epi != EPI_SILU_BWD || a.Cout <= 256, "conv_gemm: fused pixel-norm adjoint needs Cout <= 256 (one N tile)");
  }
  TEDM_CHECK(a.Cout % 32 == 0 || a.epi == EPI_PLAIN || a.epi == EPI_AXPBY || a.epi == EPI_MODSILU,
             "conv_gemm: backward epilogues need Cout %% 32 == 0");

  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t strides[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2};
    uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)p.RH, (uint32_t)p.NB};
    if (encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.x, dims, strides, box,
                    CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  {
    uint64_t K = (uint64_t)p.taps * a.Cin;
    uint64_t dims[2] = {K, (uint64_t)a.Cout};
    uint64_t strides[1] = {K * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    if (encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.w, dims, strides, box,
                    CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  switch (bn) {
    case 256: return launch<256>(ta, tb, p, stream);
    case 128: return launch<128>(ta, tb, p, stream);
    case 64: return launch<64>(ta, tb, p, stream);
    default: return fail("conv_gemm: unsupported block_n %d", bn);
  }
}

}  // namespace tedm
