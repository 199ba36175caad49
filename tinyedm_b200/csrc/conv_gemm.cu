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
#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 bf16 = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kNumThreads = 256;

template <int BN>
struct GemmCfg {
  static constexpr int kBTileBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
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

template <int BN>
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

  const int warp = threadIdx.x >> 5;
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
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
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
            uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
            uint8_t* b_dst = a_dst + kATileBytes;
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_4d(a_dst, &tmap_a, &full_bar[stage], kc * kBlockK, ds, t.h0 + dr, t.b0);
            tma_load_2d(b_dst, &tmap_b, &full_bar[stage], tap * p.Cin + kc * kBlockK, t.n0);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
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
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + kATileBytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr, 0, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(b_addr, 0, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 bytes (encoded >>4) per 16-element K step inside the 128-byte swizzle row
            umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;      // accumulator row == pixel within the tile
    const int rows_in_tile = p.NB * p.RH * p.W;
    const int HW = p.H * p.W;
    int acc = 0;
    uint32_t acc_phase = 0;
    // dropout seed = per-launch salt + a device-resident step counter (fresh masks under CUDA-graph replay)
    unsigned long long seed64 = ((unsigned long long)p.seed_hi << 32) | p.seed_lo;
    if (p.seed_ptr != nullptr) seed64 += *p.seed_ptr * 0x9E3779B97F4A7C15ull;
    const uint32_t seed_lo = (uint32_t)seed64, seed_hi = (uint32_t)(seed64 >> 32);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      Tile t = decode_tile(p, tile);
      int n_this = p.Cout - t.n0;
      if (n_this > BN) n_this = BN;
      const long long pix = t.p_base + m;
      const bool valid = (m < rows_in_tile) && (pix < t.p_limit);
      const int b = valid ? (int)(pix / HW) : 0;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      __nv_bfloat16* out_row = p.out + pix * p.Cout + t.n0;
      for (int c0 = 0; c0 < n_this; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + c0, r);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
          if (p.epi == EPI_MODSILU) {
            const float* mrow = p.mod + (long long)b * p.mod_stride + t.n0 + c0;
            if (p.out2 != nullptr) {
              __nv_bfloat16* raw_row = p.out2 + pix * p.Cout + t.n0 + c0;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (c0 + g * 8 < n_this) {
                  uint4 o;
                  o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
                  o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                  o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
                  o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                  *reinterpret_cast<uint4*>(raw_row + g * 8) = o;
                }
              }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (c0 + g * 4 < n_this) {
                float4 mm = *reinterpret_cast<const float4*>(mrow + g * 4);
                // the reference's conv output is bf16 before the fp32 modulation island (networks.py:253-258);
                // the stored (bf16) pre-modulation value is also what backward sees
                float r0 = bf16_round(v[g * 4 + 0]);
                float r1 = bf16_round(v[g * 4 + 1]);
                float r2 = bf16_round(v[g * 4 + 2]);
                float r3 = bf16_round(v[g * 4 + 3]);
                v[g * 4 + 0] = mp_silu_f(r0 * mm.x);
                v[g * 4 + 1] = mp_silu_f(r1 * mm.y);
                v[g * 4 + 2] = mp_silu_f(r2 * mm.z);
                v[g * 4 + 3] = mp_silu_f(r3 * mm.w);
              }
            }
            if (p.drop_p > 0.f) {
              const float keep_scale = 1.0f / (1.0f - p.drop_p);
              const uint32_t thresh = (uint32_t)(p.drop_p * 4294967296.0);
              const unsigned long long e0 = (unsigned long long)pix * p.Cout + t.n0 + c0;
              const uint32_t s_lo = seed_lo, s_hi = seed_hi;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                unsigned long long ctr = (e0 >> 2) + g;  // one Philox call per 4 consecutive channels
                uint4 rnd = philox4x32((uint32_t)ctr, (uint32_t)(ctr >> 32), s_lo, s_hi);
                v[g * 4 + 0] = rnd.x >= thresh ? v[g * 4 + 0] * keep_scale : 0.f;
                v[g * 4 + 1] = rnd.y >= thresh ? v[g * 4 + 1] * keep_scale : 0.f;
                v[g * 4 + 2] = rnd.z >= thresh ? v[g * 4 + 2] * keep_scale : 0.f;
                v[g * 4 + 3] = rnd.w >= thresh ? v[g * 4 + 3] * keep_scale : 0.f;
              }
            }
          } else if (p.epi == EPI_AXPBY) {
            const __nv_bfloat16* res_row = p.res + pix * p.Cout + t.n0 + c0;
            const float wa = p.beta, wb = 1.0f;  // alpha already applied to v[]
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (c0 + g * 8 < n_this) {
                uint4 rr = *reinterpret_cast<const uint4*>(res_row + g * 8);
                float2 f0 = unpack_bf16(rr.x), f1 = unpack_bf16(rr.y), f2 = unpack_bf16(rr.z),
                       f3 = unpack_bf16(rr.w);
                v[g * 8 + 0] = wa * f0.x + wb * v[g * 8 + 0];
                v[g * 8 + 1] = wa * f0.y + wb * v[g * 8 + 1];
                v[g * 8 + 2] = wa * f1.x + wb * v[g * 8 + 2];
                v[g * 8 + 3] = wa * f1.y + wb * v[g * 8 + 3];
                v[g * 8 + 4] = wa * f2.x + wb * v[g * 8 + 4];
                v[g * 8 + 5] = wa * f2.y + wb * v[g * 8 + 5];
                v[g * 8 + 6] = wa * f3.x + wb * v[g * 8 + 6];
                v[g * 8 + 7] = wa * f3.y + wb * v[g * 8 + 7];
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c0 + g * 8 < n_this) {
              uint4 o;
              o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
              o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
              o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
              o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
              *reinterpret_cast<uint4*>(out_row + c0 + g * 8) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const ConvGemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::kSmemBytes));
    configured = true;
  }
  int tiles = p.m_tiles * p.n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  conv_gemm_kernel<BN><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  TEDM_LAUNCH_CHECK();
  return 0;
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

int conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t stream) {
  TEDM_CHECK(a.ksize == 1 || a.ksize == 3, "conv_gemm: kernel size must be 1 or 3 (got %d)", a.ksize);
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
  int bn = a.Cout > 128 ? 256 : (a.Cout > 64 ? 128 : 64);
  if (a.block_n_override) bn = a.block_n_override;
  p.block_n = bn;
  p.n_tiles = (a.Cout + bn - 1) / bn;
  p.k_blocks = p.taps * (a.Cin / 64);
  p.epi = a.epi; p.alpha = a.alpha; p.out = a.out; p.out2 = a.out2; p.res = a.res;
  p.beta = a.beta; p.mod = a.mod; p.mod_stride = a.mod_stride;
  p.drop_p = a.drop_p; p.seed_lo = (uint32_t)a.seed; p.seed_hi = (uint32_t)(a.seed >> 32); p.seed_ptr = a.seed_ptr;
  if (a.epi == EPI_MODSILU) TEDM_CHECK(a.mod != nullptr, "conv_gemm: MODSILU epilogue needs mod");
  if (a.epi == EPI_AXPBY) TEDM_CHECK(a.res != nullptr, "conv_gemm: AXPBY epilogue needs res");

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
