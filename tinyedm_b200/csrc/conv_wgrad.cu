// Weight gradient of the MPConv implicit GEMM on tcgen05/TMEM (sm_100a).
//
// autograd of F.conv2d w.r.t. its weight (reference: src/tinyedm/networks.py:37, no custom backward in
// the reference) restated as a GEMM whose contraction runs over PIXELS:
//     dW[co][tap][ci] = alpha * sum_p G[p][co] * X[p + off(tap)][ci]
// Both operands are NHWC bf16, i.e. the contraction index (pixel) is the slow one: they are fed to the
// tensor core as MN-major tiles (TMA box = rows of 64 channels, 128B swizzle; UMMA descriptors with
// LBO = distance between 64-channel boxes, SBO = 1024 B between 8-pixel groups).
// The zero padding of the shifted X tile comes from TMA out-of-bounds fill, the G tile uses the same
// 4-D box so both operands hold exactly the same pixel set (rows beyond the image are zero in both).
//
// Three kernels, chosen by conv_wgrad_launch:
//   conv_wgrad_pair_kernel    CTA pair, M = 256 output channels, N <= 256 input channels of one tap (Cout % 256 == 0,
//                             Cin % 128 == 0): the CIFAR net
//   conv_wgrad_pair_t_kernel  CTA pair, dW held transposed: M = 4 slabs of the flattened (tap, ci) axis, N = output
//                             channels in blocks of 256 / 192 / 128: every other multiple of 64 (ImageNet-latent, MNIST)
//   conv_wgrad_kernel         single CTA, work item = (co block of 128, tap, ci block of <= 256, K split), accumulator
//                             128 x 256 fp32 in TMEM, split-K partial results combined with fp32 vector atomics: the rest
//                             (Cout < 128, fewer than 3 k slabs) and the A/B baseline (TEDM_CONV_PAIR=0, splits = -1)
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr int kThreads = 256;
constexpr int kBM = 128;      // Cout rows per CTA
constexpr int kBN = 256;      // Cin columns per CTA (per tap)
constexpr int kPrefRows = 64;  // pixels per pipeline stage the shared-memory budget is sized for (4 stages of 48 KB)
constexpr int kMaxRows = 128;  // largest pixel tile (then 2 stages in the single-CTA kernel); see wgrad_geometry
constexpr int kMaxStages = 4;
constexpr int kTileBytes = 4 * (kBM / 64 + kBN / 64) * kPrefRows * 128;  // 192 KB of operand stages
constexpr int kSmemBytes = kTileBytes + 1024 /*align*/ + 1024 /*barriers*/;

struct WgradParams {
  int B, H, W, Cin, Cout, taps;
  int RH, NB, rows;        // pixel tile = NB images x RH rows x W cols; rows % 16 == 0, rows <= 128
  int stages, stage_bytes; // pipeline depth and bytes per stage (6 boxes of rows*128 B)
  int tiles_h, p_tiles;    // pixel tiles
  int co_blks, ci_blks, items, splits, tiles_per_split;
  float alpha;
  int use_atomics;
  float* dw;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* acc_full = bars + 2 * kMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  smem += 1024;
  const int kStages = p.stages;
  const int kStageBytes = p.stage_bytes;
  const int kBoxBytesMax = p.rows * 128;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;

  // work item decode
  const int item = blockIdx.x % p.items;
  const int split = blockIdx.x / p.items;
  const int ci_blk = item % p.ci_blks;
  const int tap = (item / p.ci_blks) % p.taps;
  const int co_blk = item / (p.ci_blks * p.taps);
  const int co0 = co_blk * kBM;
  const int ci0 = ci_blk * kBN;
  int n_this = p.Cin - ci0;
  if (n_this > kBN) n_this = kBN;
  const int n_boxes = n_this / 64;
  const int dr = (p.taps == 9) ? tap / 3 - 1 : 0;
  const int ds = (p.taps == 9) ? tap % 3 - 1 : 0;
  const int t_begin = split * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.p_tiles) t_end = p.p_tiles;
  const int box_bytes = p.rows * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_g);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, kBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                      // programmatic dependent launch: the prologue above overlaps the previous kernel's tail
  if (threadIdx.x == 0) pdl_trigger();

  if (warp == 0) {
    const bool leader_lane = elect_one();   // warp-uniform loop, one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = (uint32_t)(2 + n_boxes) * box_bytes;
    for (int t = t_begin; t < t_end; ++t) {
      const int bi = t / p.tiles_h;
      const int b0 = bi * p.NB;
      const int h0 = (t - bi * p.tiles_h) * p.RH;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (leader_lane) {
        uint8_t* s = smem + stage * kStageBytes;
        mbar_expect_tx(&full_bar[stage], tx);
        tma_load_4d(s, &tmap_g, &full_bar[stage], co0, 0, h0, b0);
        tma_load_4d(s + box_bytes, &tmap_g, &full_bar[stage], co0 + 64, 0, h0, b0);
        uint8_t* xs = s + 2 * kBoxBytesMax;
        for (int j = 0; j < n_boxes; ++j)
          tma_load_4d(xs + j * box_bytes, &tmap_x, &full_bar[stage], ci0 + j * 64, ds, h0 + dr, b0);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t idesc = make_idesc_bf16(kBM, n_this, 1, 1);
    const int ksteps = p.rows / 16;
    const uint32_t smem_base = smem_u32(smem);
    uint32_t accumulate = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (leader_lane) {
        const uint32_t a_addr = smem_base + stage * kStageBytes;
        const uint32_t b_addr = a_addr + 2 * kBoxBytesMax;
        const uint64_t a_desc = make_smem_desc_sw128(a_addr, box_bytes, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(b_addr, box_bytes, 1024);
        for (int k = 0; k < ksteps; ++k) {
          umma_bf16(tmem_base, a_desc + (uint64_t)(k * 128), b_desc + (uint64_t)(k * 128), idesc, accumulate);
          accumulate = 1;
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (leader_lane) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    if (t_end > t_begin) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      float* dst = p.dw + ((long long)co * p.taps + tap) * p.Cin + ci0;
      for (int c0 = 0; c0 < n_this; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + c0, r);
        tmem_ld_wait();
        if (co < p.Cout) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 v = make_float4(__uint_as_float(r[g * 4 + 0]) * p.alpha, __uint_as_float(r[g * 4 + 1]) * p.alpha,
                                   __uint_as_float(r[g * 4 + 2]) * p.alpha, __uint_as_float(r[g * 4 + 3]) * p.alpha);
            float4* d4 = reinterpret_cast<float4*>(dst + c0 + g * 4);
            if (p.use_atomics) {
              atomicAdd(d4, v);
            } else {
              *d4 = v;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kBN);
  }
}

// Pixel-tile geometry for the contraction: rows = NB*RH*W (whole image rows, NB images when one image is smaller than
// the tile) must be a multiple of 16 — one MMA contracts 16 pixels — and at most kMaxRows; the largest such tile wins
// (fewer, longer pipeline stages measured faster than kPrefRows-sized ones on every shape of the three configs).
// Widths with lcm(16, W) > kMaxRows (9, 13, 17, 18, ...) have no tile: the launchers refuse them, and the stand-alone
// Conv2d module pads such maps with zero columns first (networks.py `_wgrad_width`).
int wgrad_geometry(int H, int W, int* RH, int* NB) {
  int best = 0;
  for (int rh = 1; rh <= H; ++rh) {
    if (rh * W > kMaxRows) break;
    const int nb_max = kMaxRows / (rh * W);
    for (int nb = 1; nb <= nb_max; ++nb) {
      const int rows = rh * W * nb;
      if (rows % 16 != 0) continue;
      if (rows > best) { best = rows; *RH = rh; *NB = nb; }   // among equal row counts: fewer images per box
    }
  }
  return best > 0 ? 0 : -1;
}


// ------------------------------------------------------------------------------------------------------------------
// CTA-pair version (tcgen05 cta_group::2): one MMA of M = 256 output channels x N <= 256 input channels per 16 pixels,
// issued by the leader CTA of a 2-CTA cluster. Each CTA stages its own 128 rows of G but only HALF of the X tile, so the
// shared-memory data pipe carries 8 KB of operand reads per MMA and SM instead of 12 and 32 KB of TMA fill per 64
// pixels instead of 48 (see profiles/r1c_conv_gemm_ncu_full.md). The split-K partial sums leave through swizzled
// shared memory and TMA reduce-add (fp32) instead of 64 scattered 16-byte atomics per thread.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kPairThreads = 384;
constexpr int kPairStagesMax = 6;
constexpr int kPairTileBytes = 6 * 4 * kPrefRows * 128;   // 192 KB of operand stages (4 boxes per stage and CTA)
constexpr int kPairSmemBytes = kPairTileBytes + 1024 /*align*/ + 1024 /*barriers*/;

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                       const __grid_constant__ CUtensorMap tmap_dw, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* full_bar = bars;                       // leader: its producer's arrival + the bytes of both CTAs' loads
  uint64_t* empty_bar = bars + kPairStagesMax;     // MMA commit arrives in both CTAs
  uint64_t* acc_full = bars + 2 * kPairStagesMax;  // MMA commit arrives in both CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  smem += 1024;
  const int kStages = p.stages;
  const int kStageBytes = p.stage_bytes;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (see conv_pair.cu)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cl = (int)cluster_id_x();

  // work item decode: (co block of 256, tap, ci block of <= 256, K split)
  const int item = cl % p.items;
  const int split = cl / p.items;
  const int ci_blk = item % p.ci_blks;
  const int tap = (item / p.ci_blks) % p.taps;
  const int co_blk = item / (p.ci_blks * p.taps);
  const int co0 = co_blk * 256 + (int)rank * 128;   // this CTA's 128 rows of the accumulator
  const int ci0 = ci_blk * kBN;
  int n_this = p.Cin - ci0;
  if (n_this > kBN) n_this = kBN;
  const int nb_half = (n_this >> 1) / 64;           // 64-channel boxes of X staged by this CTA
  const int dr = (p.taps == 9) ? tap / 3 - 1 : 0;
  const int ds = (p.taps == 9) ? tap % 3 - 1 : 0;
  const int t_begin = split * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.p_tiles) t_end = p.p_tiles;
  const int box_bytes = p.rows * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_g);
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dw);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_ptr, kBN);
    tmem_relinquish_pair();
  }
  pdl_wait();                      // programmatic dependent launch: the prologue above overlaps the previous kernel's tail
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (threadIdx.x == 0) pdl_trigger();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // warp-uniform loop, one elected lane issues (keeps the loop state in uniform registers)
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_pair = 2u * (uint32_t)(2 + nb_half) * box_bytes;
    for (int t = t_begin; t < t_end; ++t) {
      const int bi = t / p.tiles_h;
      const int b0 = bi * p.NB;
      const int h0 = (t - bi * p.tiles_h) * p.RH;
      mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
      if (leader_lane) {
        uint8_t* s = smem + stage * kStageBytes;
        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
        if (rank == 0) mbar_expect_tx(&full_bar[stage], tx_pair);
        tma_load_4d_pair(s, &tmap_g, full_leader, co0, 0, h0, b0);
        tma_load_4d_pair(s + box_bytes, &tmap_g, full_leader, co0 + 64, 0, h0, b0);
        uint8_t* xs = s + 2 * box_bytes;
        for (int j = 0; j < nb_half; ++j)
          tma_load_4d_pair(xs + j * box_bytes, &tmap_x, full_leader, ci0 + (int)rank * (n_this >> 1) + j * 64, ds, h0 + dr, b0);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const bool leader_lane = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = make_idesc_bf16(256, n_this, 1, 1);
      const int ksteps = p.rows / 16;
      const uint32_t smem_base = smem_u32(smem);
      uint32_t accumulate = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait_bounded(&full_bar[stage], phase);
        tc_fence_after();
        if (leader_lane) {
          const uint32_t a_addr = smem_base + stage * kStageBytes;
          const uint32_t b_addr = a_addr + 2 * box_bytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr, box_bytes, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(b_addr, box_bytes, 1024);
          for (int k = 0; k < ksteps; ++k) {
            umma_bf16_pair(tmem_base, a_desc + (uint64_t)(k * 128), b_desc + (uint64_t)(k * 128), idesc, accumulate);
            accumulate = 1;
          }
          umma_commit_pair(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (leader_lane) umma_commit_pair(acc_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // epilogue: 128 rows (co) x n_this columns (ci) of fp32 -> swizzled [128][32] chunks -> TMA reduce-add / store
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int m = q * 32 + lane;
    const bool elected = (warp == 4 + 4 * half) && lane == 0;
    if (t_end > t_begin) {
      mbar_wait_bounded(acc_full, 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      const int nch = n_this / 32;
      const int cb = half == 0 ? 0 : nch / 2, ce = half == 0 ? nch / 2 : nch;
      for (int c = cb; c < ce; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        uint8_t* buf = smem + c * (128 * 128);   // all operand stages are dead once acc_full has fired
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 v = make_float4(__uint_as_float(r[g * 4 + 0]) * p.alpha, __uint_as_float(r[g * 4 + 1]) * p.alpha,
                                 __uint_as_float(r[g * 4 + 2]) * p.alpha, __uint_as_float(r[g * 4 + 3]) * p.alpha);
          *reinterpret_cast<float4*>(buf + m * 128 + ((g ^ (m & 7)) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + half, 128);
      if (elected) {
        for (int c = cb; c < ce; ++c) {
          const uint8_t* buf = smem + c * (128 * 128);
          if (p.use_atomics) tma_reduce_add_2d(&tmap_dw, buf, tap * p.Cin + ci0 + c * 32, co0);
          else tma_store_2d(&tmap_dw, buf, tap * p.Cin + ci0 + c * 32, co0);
        }
        bulk_commit();
        bulk_wait_read0();   // shared memory must outlive the reads; the global updates complete with the grid
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kBN);
  }
}

bool wgrad_pair_supported(const ConvWgradArgs& a) {
  if (a.Cout % 256 != 0) return false;
  if (a.Cin % 128 != 0) return false;                    // each CTA stages whole 64-channel boxes of its half
  if (a.Cin > 256 && a.Cin % 256 != 0) return false;
  return true;
}

int conv_wgrad_pair_launch(const ConvWgradArgs& a, cudaStream_t stream) {
  WgradParams p{};
  p.B = a.B; p.H = a.H; p.W = a.W; p.Cin = a.Cin; p.Cout = a.Cout; p.taps = a.ksize * a.ksize;
  TEDM_CHECK(wgrad_geometry(a.H, a.W, &p.RH, &p.NB) == 0, "conv_wgrad: unsupported spatial size %dx%d", a.H, a.W);
  p.rows = p.RH * p.NB * a.W;
  const int n_blk = a.Cin < kBN ? a.Cin : kBN;
  p.stage_bytes = (2 + (n_blk / 2) / 64) * p.rows * 128;
  p.stages = kPairTileBytes / p.stage_bytes;
  if (p.stages > kPairStagesMax) p.stages = kPairStagesMax;
  TEDM_CHECK(p.stages >= 2, "conv_wgrad: pixel tile of %d rows does not fit two pipeline stages", p.rows);
  TEDM_CHECK(p.stages * p.stage_bytes >= (n_blk / 32) * 128 * 128, "conv_wgrad: epilogue staging does not fit");
  p.tiles_h = (a.H + p.RH - 1) / p.RH;
  p.p_tiles = ((a.B + p.NB - 1) / p.NB) * p.tiles_h;
  p.co_blks = a.Cout / 256;
  p.ci_blks = (a.Cin + kBN - 1) / kBN;
  p.items = p.co_blks * p.ci_blks * p.taps;
  int splits = a.splits_override;
  const int max_clusters = num_sms() / 2;
  if (splits <= 0) splits = max_clusters / p.items;
  if (splits > p.p_tiles) splits = p.p_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.p_tiles + splits - 1) / splits;
  splits = (p.p_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.splits = splits;
  p.alpha = a.alpha;
  p.use_atomics = (splits > 1 || a.accumulate) ? 1 : 0;
  p.dw = a.dw;
  if (splits > 1 && !a.accumulate)
    TEDM_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.Cout * p.taps * a.Cin, stream));

  CUtensorMap tg, tx, tdw;
  {
    uint64_t dims[4] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t strides[3] = {(uint64_t)a.Cout * 2, (uint64_t)a.W * a.Cout * 2, (uint64_t)a.H * a.W * a.Cout * 2};
    uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)p.RH, (uint32_t)p.NB};
    if (encode_tmap(&tg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.g, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  {
    uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t strides[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2};
    uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)p.RH, (uint32_t)p.NB};
    if (encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  {
    const uint64_t K = (uint64_t)p.taps * a.Cin;
    uint64_t dims[2] = {K, (uint64_t)a.Cout};
    uint64_t strides[1] = {K * 4};
    uint32_t box[2] = {32, 128};
    if (encode_tmap(&tdw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.dw, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * p.items * splits);
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = kPairSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  TEDM_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_pair_kernel, tg, tx, tdw, p));
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Transposed CTA-pair version for the channel counts the kernel above cannot tile (multiples of 64 that are not
// multiples of 256 / 128: the 192-, 384- and 576-channel levels of the ImageNet-latent net, the 128-channel level of
// the MNIST net). The accumulator holds dW TRANSPOSED:
//     M (256 TMEM lanes over the CTA pair) = 4 consecutive 64-wide slabs of the flattened (tap, ci) axis of dW —
//       k = tap*Cin + ci IS that axis, so a slab is one X box with its own tap shift, and any Cin % 64 == 0 tiles
//       with at most one padding slab per launch (27 slabs -> 28 for 3x3 x 192 channels);
//     N (<= 256 columns) = output channels, G boxes, half of them staged by each CTA. Cout % 192 == 0 uses N = 192:
//       each CTA stages the two boxes that cover its 96 columns (the MMA reads 1.5 swizzle atoms), so the 192-, 384-
//       and 576-channel layers waste no tensor work; other widths round the last block up to a multiple of 128
//       (out-of-bounds G columns are zero-filled by TMA, their results clipped by the TMA reduce).
// Epilogue: a TMEM lane is one k, a register one co; for a fixed register the 32 lanes of a warp write 32 consecutive
// floats of one 128-byte row of a [co][32 k] staging block (conflict-free), which leaves by TMA reduce-add into
// dW[co][k] — the transposition costs nothing.
// ------------------------------------------------------------------------------------------------------------------
struct WgradTParams {
  int Cin, Cout, taps;
  int RH, NB, rows, stages, stage_bytes, tiles_h, p_tiles;
  int slabs_per_tap, total_slabs, m_blks, n_blks, n_tile, items, tiles_per_split;
  int n192;          // 1: N = 192 blocks, each CTA stages the boxes at rank*96 and rank*96 + 64
  float alpha;
  int use_atomics;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_wgrad_pair_t_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                         const __grid_constant__ CUtensorMap tmap_dw, const WgradTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kPairStagesMax;
  uint64_t* acc_full = bars + 2 * kPairStagesMax;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  smem += 1024;
  const int kStages = p.stages;
  const int kStageBytes = p.stage_bytes;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cl = (int)cluster_id_x();

  // work item: (block of 4 k slabs, block of output channels, K split)
  const int item = cl % p.items;
  const int split = cl / p.items;
  const int n_blk = item % p.n_blks;
  const int m_blk = item / p.n_blks;
  const int slab0 = m_blk * 4 + (int)rank * 2;      // this CTA's two slabs = its 128 accumulator lanes
  const int n0 = n_blk * p.n_tile;
  int n_this = p.n_tile;
  if (!p.n192) {
    const int left = (p.Cout - n0 + 127) & ~127;
    if (n_this > left) n_this = left;
  }
  const int nb_g = p.n192 ? 2 : (n_this >> 1) / 64;  // G boxes staged by this CTA
  const int g_c0 = n0 + (int)rank * (n_this >> 1);
  int tap_j[2], ci_j[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    int s = slab0 + j;
    if (s >= p.total_slabs) s = p.total_slabs - 1;   // padding slab: any valid box, its results are clipped
    tap_j[j] = s / p.slabs_per_tap;
    ci_j[j] = (s - tap_j[j] * p.slabs_per_tap) * 64;
  }
  const int t_begin = split * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.p_tiles) t_end = p.p_tiles;
  const int box_bytes = p.rows * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_g);
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dw);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_ptr, kBN);
    tmem_relinquish_pair();
  }
  pdl_wait();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (threadIdx.x == 0) pdl_trigger();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_pair = 2u * (uint32_t)(2 + nb_g) * box_bytes;
    const int dr0 = (p.taps == 9) ? tap_j[0] / 3 - 1 : 0, ds0 = (p.taps == 9) ? tap_j[0] % 3 - 1 : 0;
    const int dr1 = (p.taps == 9) ? tap_j[1] / 3 - 1 : 0, ds1 = (p.taps == 9) ? tap_j[1] % 3 - 1 : 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int bi = t / p.tiles_h;
      const int b0 = bi * p.NB;
      const int h0 = (t - bi * p.tiles_h) * p.RH;
      mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
      if (leader_lane) {
        uint8_t* s = smem + stage * kStageBytes;
        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
        if (rank == 0) mbar_expect_tx(&full_bar[stage], tx_pair);
        tma_load_4d_pair(s, &tmap_x, full_leader, ci_j[0], ds0, h0 + dr0, b0);
        tma_load_4d_pair(s + box_bytes, &tmap_x, full_leader, ci_j[1], ds1, h0 + dr1, b0);
        uint8_t* gs = s + 2 * box_bytes;
        for (int j = 0; j < nb_g; ++j)
          tma_load_4d_pair(gs + j * box_bytes, &tmap_g, full_leader, g_c0 + j * 64, 0, h0, b0);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const bool leader_lane = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = make_idesc_bf16(256, n_this, 1, 1);
      const int ksteps = p.rows / 16;
      const uint32_t smem_base = smem_u32(smem);
      uint32_t accumulate = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait_bounded(&full_bar[stage], phase);
        tc_fence_after();
        if (leader_lane) {
          const uint32_t a_addr = smem_base + stage * kStageBytes;
          const uint32_t b_addr = a_addr + 2 * box_bytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr, box_bytes, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(b_addr, box_bytes, 1024);
          for (int k = 0; k < ksteps; ++k) {
            umma_bf16_pair(tmem_base, a_desc + (uint64_t)(k * 128), b_desc + (uint64_t)(k * 128), idesc, accumulate);
            accumulate = 1;
          }
          umma_commit_pair(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (leader_lane) umma_commit_pair(acc_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // epilogue: lane m of quarter q holds k = slab0*64 + q*32 + lane; registers run over co. Staging block of chunk c
    // (32 output channels) and quarter q: [32 co rows][32 k floats], 128B-swizzled, 4 KB.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    if (t_end > t_begin) {
      mbar_wait_bounded(acc_full, 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      const int nch = n_this / 32;
      const int cb = half == 0 ? 0 : nch / 2, ce = half == 0 ? nch / 2 : nch;
      const int k0 = slab0 * 64 + q * 32;
      const uint32_t col_off = (uint32_t)(lane & 3) * 4u;
      for (int c = cb; c < ce; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        uint8_t* buf = smem + (c * 4 + q) * 4096;   // all operand stages are dead once acc_full has fired
#pragma unroll
        for (int i = 0; i < 32; ++i)
          *reinterpret_cast<float*>(buf + i * 128 + ((((uint32_t)lane >> 2) ^ (uint32_t)(i & 7)) << 4) + col_off) =
              __uint_as_float(r[i]) * p.alpha;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        for (int c = cb; c < ce; ++c) {
          const uint8_t* buf = smem + (c * 4 + q) * 4096;
          if (p.use_atomics) tma_reduce_add_2d(&tmap_dw, buf, k0, n0 + c * 32);
          else tma_store_2d(&tmap_dw, buf, k0, n0 + c * 32);
        }
        bulk_commit();
        bulk_wait_read0();
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kBN);
  }
}

static int wgrad_t_mode() {   // TEDM_WGRAD_T=0: A/B switch back to the single-CTA kernel; TEDM_WGRAD_N192=0: no N = 192 blocks
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("TEDM_WGRAD_T");
    const char* n = getenv("TEDM_WGRAD_N192");
    mode = (e != nullptr && e[0] == '0') ? 0 : 1;
    if (mode && !(n != nullptr && n[0] == '0')) mode = 2;
  }
  return mode;
}

// K splits of one tiling: minimise (waves of clusters) x (MMA time of a cluster's pixel tiles + an epilogue's worth of time,
// in units of one 256-column pixel tile). Returns the cost, the split count through *splits.
static long long wgrad_t_best_splits(int items, int p_tiles, int n_tile, int* splits) {
  const int max_clusters = num_sms() / 2;
  const int kEpiUnits = 12 * 256;
  long long best = -1;
  for (int s = 1; s <= p_tiles && s <= 4 * max_clusters; ++s) {
    const int tps = (p_tiles + s - 1) / s;
    const int s_eff = (p_tiles + tps - 1) / tps;
    const long long waves = ((long long)items * s_eff + max_clusters - 1) / max_clusters;
    const long long cost = waves * ((long long)tps * n_tile + kEpiUnits);
    if (best < 0 || cost < best) { best = cost; *splits = s_eff; }
  }
  return best;
}

// Fills the geometry; returns false when the launch should stay on the single-CTA kernel.
bool wgrad_pair_t_plan(const ConvWgradArgs& a, WgradTParams* p, int* splits) {
  if (wgrad_t_mode() == 0) return false;
  if (a.Cin % 64 != 0 || a.Cout % 64 != 0 || a.Cout < 128) return false;
  p->Cin = a.Cin; p->Cout = a.Cout; p->taps = a.ksize * a.ksize;
  if (wgrad_geometry(a.H, a.W, &p->RH, &p->NB) != 0) return false;
  p->rows = p->RH * p->NB * a.W;
  p->stage_bytes = 4 * p->rows * 128;
  p->stages = kPairTileBytes / p->stage_bytes;
  if (p->stages > kPairStagesMax) p->stages = kPairStagesMax;
  if (p->stages < 2) return false;
  if (p->stages * p->stage_bytes < (256 / 32) * 4 * 4096) return false;   // epilogue staging
  p->slabs_per_tap = a.Cin / 64;
  p->total_slabs = p->taps * p->slabs_per_tap;
  if (p->total_slabs < 3) return false;   // a 1x1 conv over <= 128 channels would fill less than 3/4 of the 256 accumulator lanes
  p->m_blks = (p->total_slabs + 3) / 4;
  p->tiles_h = (a.H + p->RH - 1) / p->RH;
  p->p_tiles = ((a.B + p->NB - 1) / p->NB) * p->tiles_h;
  // N = 256 blocks (last one rounded up to a multiple of 128) or, when Cout is a multiple of 192, exact N = 192 blocks
  int s256 = 1, s192 = 1;
  const int nb256 = (a.Cout + 255) / 256;
  const long long c256 = wgrad_t_best_splits(p->m_blks * nb256, p->p_tiles, 256, &s256);
  long long c192 = -1;
  if (wgrad_t_mode() == 2 && a.Cout % 192 == 0 && a.Cout % 256 != 0)
    c192 = wgrad_t_best_splits(p->m_blks * (a.Cout / 192), p->p_tiles, 192, &s192);
  p->n192 = (c192 >= 0 && c192 < c256) ? 1 : 0;
  p->n_tile = p->n192 ? 192 : 256;
  p->n_blks = (a.Cout + p->n_tile - 1) / p->n_tile;
  p->items = p->m_blks * p->n_blks;
  *splits = p->n192 ? s192 : s256;
  return true;
}

int conv_wgrad_pair_t_launch(const ConvWgradArgs& a, WgradTParams p, int splits, cudaStream_t stream) {
  if (a.splits_override > 0) splits = a.splits_override;
  if (splits > p.p_tiles) splits = p.p_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.p_tiles + splits - 1) / splits;
  splits = (p.p_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.alpha = a.alpha;
  p.use_atomics = (splits > 1 || a.accumulate) ? 1 : 0;
  if (splits > 1 && !a.accumulate)
    TEDM_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.Cout * p.taps * a.Cin, stream));

  CUtensorMap tg, tx, tdw;
  {
    uint64_t dims[4] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t strides[3] = {(uint64_t)a.Cout * 2, (uint64_t)a.W * a.Cout * 2, (uint64_t)a.H * a.W * a.Cout * 2};
    uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)p.RH, (uint32_t)p.NB};
    if (encode_tmap(&tg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.g, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  {
    uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t strides[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2};
    uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)p.RH, (uint32_t)p.NB};
    if (encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  {
    const uint64_t K = (uint64_t)p.taps * a.Cin;
    uint64_t dims[2] = {K, (uint64_t)a.Cout};
    uint64_t strides[1] = {K * 4};
    uint32_t box[2] = {32, 32};
    if (encode_tmap(&tdw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.dw, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_wgrad_pair_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * p.items * splits);
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = kPairSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  TEDM_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_pair_t_kernel, tg, tx, tdw, p));
  return 0;
}

static int wgrad_pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("TEDM_CONV_PAIR");
    mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return mode;
}

}  // namespace

int conv_wgrad_launch(const ConvWgradArgs& a, cudaStream_t stream) {
  TEDM_CHECK(a.ksize == 1 || a.ksize == 3, "conv_wgrad: kernel size must be 1 or 3 (got %d)", a.ksize);
  TEDM_CHECK(a.Cin % 64 == 0 && a.Cout % 64 == 0, "conv_wgrad: Cin/Cout must be multiples of 64 (got %d/%d)",
             a.Cin, a.Cout);
  TEDM_CHECK(a.B > 0 && a.H > 0 && a.W > 0, "conv_wgrad: empty input");
  if (a.splits_override >= 0 && wgrad_pair_mode() == 1) {
    if (wgrad_pair_supported(a)) return conv_wgrad_pair_launch(a, stream);
    WgradTParams tp{};
    int t_splits = 1;
    if (wgrad_pair_t_plan(a, &tp, &t_splits)) return conv_wgrad_pair_t_launch(a, tp, t_splits, stream);
  }
  WgradParams p{};
  p.B = a.B; p.H = a.H; p.W = a.W; p.Cin = a.Cin; p.Cout = a.Cout; p.taps = a.ksize * a.ksize;
  TEDM_CHECK(wgrad_geometry(a.H, a.W, &p.RH, &p.NB) == 0, "conv_wgrad: unsupported spatial size %dx%d", a.H, a.W);
  p.rows = p.RH * p.NB * a.W;
  p.stage_bytes = (kBM / 64 + kBN / 64) * p.rows * 128;
  p.stages = kTileBytes / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  TEDM_CHECK(p.stages >= 2, "conv_wgrad: pixel tile of %d rows does not fit two pipeline stages", p.rows);
  p.tiles_h = (a.H + p.RH - 1) / p.RH;
  p.p_tiles = ((a.B + p.NB - 1) / p.NB) * p.tiles_h;
  p.co_blks = (a.Cout + kBM - 1) / kBM;
  p.ci_blks = (a.Cin + kBN - 1) / kBN;
  p.items = p.co_blks * p.ci_blks * p.taps;
  int splits = a.splits_override;
  if (splits <= 0) {
    // one CTA per SM (192 KB of operand stages each): aim for exactly one full wave, never a ragged second one
    splits = num_sms() / p.items;
  }
  if (splits > p.p_tiles) splits = p.p_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.p_tiles + splits - 1) / splits;
  splits = (p.p_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.splits = splits;
  p.alpha = a.alpha;
  p.use_atomics = (splits > 1 || a.accumulate) ? 1 : 0;
  p.dw = a.dw;
  if (splits > 1 && !a.accumulate)
    TEDM_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.Cout * p.taps * a.Cin, stream));

  CUtensorMap tg, tx;
  {
    uint64_t dims[4] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t strides[3] = {(uint64_t)a.Cout * 2, (uint64_t)a.W * a.Cout * 2, (uint64_t)a.H * a.W * a.Cout * 2};
    uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)p.RH, (uint32_t)p.NB};
    if (encode_tmap(&tg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.g, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  {
    uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t strides[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2};
    uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)p.RH, (uint32_t)p.NB};
    if (encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
      return -1;
  }
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  launch_pdl(conv_wgrad_kernel, p.items * splits, kThreads, kSmemBytes, stream, tg, tx, p);
  TEDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace tedm
