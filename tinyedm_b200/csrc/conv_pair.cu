// MPConv implicit GEMM on a CTA PAIR: tcgen05.mma.cta_group::2 (M = 256 pixels x N = 256 channels per pair tile),
// TMA-fed, with the whole epilogue I/O staged through 128B-swizzled shared memory and moved by TMA.
//
// Same operation as conv_gemm.cu (F.conv2d(x, w_hat, padding="same") of src/tinyedm/networks.py:31-38, forward and
// data gradient, with the fused epilogues of kernels.h), restructured after the r1c ncu capture
// (profiles/r1c_conv_gemm_ncu_full.md): the single-CTA kernel is bound by the L1/shared-memory data pipe, which
// carries the tensor core's operand reads (12 KB per 128x256x16 MMA), the TMA fills (48 KB per k block) and the
// epilogue's row-per-thread global accesses (32 wavefronts per instruction). Here
//   * the two CTAs of a cluster (one TPC) issue ONE MMA of M=256: each CTA stages its own 128 pixels of A but only
//     HALF of the weight tile (128 of the 256 output channels) -> 8 KB of operand reads per MMA and SM instead of 12,
//     32 KB of TMA fill per k block instead of 48;
//   * epilogue operands (residual, saved activations) arrive by TMA one 64-channel chunk ahead, results leave by TMA
//     store; each thread touches only its own 128-byte row of a swizzled buffer (conflict-free 16-byte accesses).
//
// Roles per CTA (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only), warp 2 = TMEM
// allocator, warps 4..11 = epilogue (4 TMEM lane quarters x 2 column halves [x kSub warps per 64-column chunk]). Accumulators are double buffered in
// TMEM (2 x 256 columns in each CTA).
#include <mutex>

#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr int kBM = 128;                              // pixels per CTA (256 per pair tile)
constexpr int kBK = 64;                               // 64 bf16 = one 128-byte swizzle row
constexpr int kBN = 256;                              // output channels per pair tile
constexpr int kATileBytes = kBM * kBK * 2;            // 16 KB
constexpr int kBHalfBytes = (kBN / 2) * kBK * 2;      // 16 KB: this CTA's half of the weight tile
constexpr int kStageBytes = kATileBytes + kBHalfBytes;
constexpr int kStages = 4;
constexpr int kEpiBufBytes = kBM * 128;               // 128 pixels x 64 channels bf16
constexpr int kEpiBufsPerHalf = 3;
// Epilogue warps sharing a (TMEM lane quarter, column half); each owns 64/kSub columns of a chunk. kSub = 2 (16 epilogue
// warps, 16-column register steps) was measured on B200: no faster than kSub = 1 — under the power cap the extra warps buy
// nothing because the epilogue already overlaps the next tile's main loop — and slower for the two-sweep variant.
constexpr int kSub = 1;
constexpr int kEpiWarps = 8 * kSub;
constexpr int kThreads = 128 + 32 * kEpiWarps;        // 384 (640 with kSub = 2)
constexpr int kHalfThreads = 128 * kSub;              // threads behind one named barrier / one set of staging buffers
constexpr int CW = kSub == 1 ? 32 : 16;               // columns per register step (640 threads would leave <= 102 registers each)
constexpr int kStepsPerWarp = 64 / kSub / CW;
constexpr int kTmemCols = 512;
constexpr int kOffEpi = kStages * kStageBytes;
constexpr int kOffBars = kOffEpi + 2 * kEpiBufsPerHalf * kEpiBufBytes;
constexpr int kSmemBytes = kOffBars + 256 /*barriers*/ + 1024 /*row dots*/ + 1024 /*alignment slack*/;
static_assert(kSmemBytes <= 232448, "shared memory budget");

struct MTile {
  int b0, h0;
  long long p_base;   // global pixel index of the tile's first pixel
  long long p_limit;  // pixels >= p_limit are not part of this tile's images
};

__device__ __forceinline__ MTile decode_m(const ConvGemmParams& p, int mt) {
  MTile t;
  const long long total = (long long)p.B * p.H * p.W;
  if (p.NB == 1) {
    t.b0 = mt / p.tiles_h;
    t.h0 = (mt - t.b0 * p.tiles_h) * p.RH;
    t.p_limit = (long long)(t.b0 + 1) * p.H * p.W;
  } else {
    t.b0 = mt * p.NB;
    t.h0 = 0;
    t.p_limit = total;
  }
  if (t.p_limit > total) t.p_limit = total;
  t.p_base = ((long long)t.b0 * p.H + t.h0) * p.W;
  return t;
}

// 16-byte piece j (0..7) of row m in a [128 rows][64 bf16] buffer written/read by TMA with SWIZZLE_128B
__device__ __forceinline__ uint4* srow(uint8_t* buf, int m, int j) {
  return reinterpret_cast<uint4*>(buf + m * 128 + ((j ^ (m & 7)) << 4));
}
template <int N>
__device__ __forceinline__ void srow_load(uint8_t* buf, int m, int j0, float (&f)[N]) {
#pragma unroll
  for (int g = 0; g < N / 8; ++g) {
    const uint4 u = *srow(buf, m, j0 + g);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[g * 8 + 0] = a.x; f[g * 8 + 1] = a.y; f[g * 8 + 2] = b.x; f[g * 8 + 3] = b.y;
    f[g * 8 + 4] = c.x; f[g * 8 + 5] = c.y; f[g * 8 + 6] = d.x; f[g * 8 + 7] = d.y;
  }
}
template <int N>
__device__ __forceinline__ void srow_store(uint8_t* buf, int m, int j0, const float (&v)[N]) {
#pragma unroll
  for (int g = 0; g < N / 8; ++g) {
    uint4 o;
    o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
    o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
    o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
    o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
    *srow(buf, m, j0 + g) = o;
  }
}

// N values per lane (N = 16 or 32) -> lane j < N ends up with the sum over all 32 lanes of value j (N - 1 exchange
// shuffles, + 1 for N = 16 where lanes j and j + 16 hold the two halves of column j)
template <int N>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[N], int lane) {
#pragma unroll
  for (int s = N / 2; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  if (N == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
  return v[0];
}

template <int N>
__device__ __forceinline__ void dropout_apply(float (&v)[N], unsigned long long e0, float drop_p, uint32_t dseed) {
  const float keep_scale = 1.0f / (1.0f - drop_p);
  const uint32_t thresh = (uint32_t)(drop_p * 65536.0f);
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    const uint32_t bits = dropout_bits2(e0 + i, dseed);
    v[i] = (bits & 0xFFFFu) >= thresh ? v[i] * keep_scale : 0.f;
    v[i + 1] = (bits >> 16) >= thresh ? v[i + 1] * keep_scale : 0.f;
  }
}

// the 16-column overloads serve kSub = 2 (16 epilogue warps per CTA), kept selectable at compile time
[[maybe_unused]] __device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
[[maybe_unused]] __device__ __forceinline__ void tmem_st_cw(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st16(taddr, r); }
__device__ __forceinline__ void tmem_st_cw(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st32(taddr, r); }

// Work item -> (pair M-tile, first output channel, channels). The tail of a launch whose whole tiles do not fill the
// last wave of CTA pairs is cut into half-N items so that the remainder occupies (almost) every pair for half a tile
// time instead of fewer than half of the pairs for a whole one (16x16 maps at B = 256: 3.5 tile times instead of 4).
struct NTile { int pmt, n0, n_this; };
__device__ __forceinline__ NTile decode_item(const ConvGemmParams& p, int item) {
  int tile = item, half = -1;
  if (item >= p.split_from) {
    const int h = item - p.split_from;
    tile = p.split_from + (h >> 1);
    half = h & 1;
  }
  NTile t;
  t.pmt = tile / p.n_tiles;
  t.n0 = (tile - t.pmt * p.n_tiles) * kBN;
  t.n_this = p.Cout - t.n0;
  if (t.n_this > kBN) t.n_this = kBN;
  if (half >= 0) {           // only whole 256-channel tiles are split (host side guarantees Cout % 256 == 0)
    t.n_this >>= 1;
    t.n0 += half * t.n_this;
  }
  return t;
}

// Position of one epilogue half (4 warps) in its stream of 64-channel chunks.
struct ChunkPos {
  int ptile;   // pair tile
  int pass;    // 0, or 1 = second sweep of the fused pixel-norm adjoint
  int c;       // chunk within the tile
  int cb, ce;  // this half's chunk range in the tile (may be empty)
  bool valid;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_o2,
                 const __grid_constant__ CUtensorMap tmap_i0, const __grid_constant__ CUtensorMap tmap_i1,
                 const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* full_bar = bars;                    // [kStages] leader: its producer's arrival + the bytes of BOTH CTAs' loads
  uint64_t* empty_bar = bars + kStages;         // [kStages] own producer waits; MMA commit arrives in both CTAs
  uint64_t* tmem_full = bars + 2 * kStages;     // [2] MMA commit arrives in both CTAs
  uint64_t* tmem_empty = tmem_full + 2;         // [2] leader: 16 arrivals (8 epilogue warps of each CTA)
  uint64_t* in_full = tmem_empty + 2;           // [2] one per epilogue half: chunk operands have landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(in_full + 2);
  float* row_dots = reinterpret_cast<float*>(smem + kOffBars + 256);   // [2 halves][128 rows]

  // warp index through a shuffle: provably warp-uniform, so role branches and the producer / MMA loops below stay in
  // uniform registers (the r1j profile showed the single MMA-issuing thread spending ~120 instructions per k block on
  // R2UR / ELECT sequences and losing issue slots to the epilogue warps of its scheduler)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)cluster_id_x();
  const int ncl = (int)num_clusters_x();
  const int pair_tiles = p.work_items;   // whole pair tiles, then (optionally) half-N items: see decode_item

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_o);
    if ((EPI == EPI_MODSILU || EPI == EPI_SILU_BWD) && p.out2 != nullptr) tma_prefetch_desc(&tmap_o2);
    if (EPI == EPI_AXPBY || EPI == EPI_MODSILU_BWD || EPI == EPI_SILU_BWD) tma_prefetch_desc(&tmap_i0);
    if (EPI == EPI_SILU_BWD && p.res != nullptr) tma_prefetch_desc(&tmap_i1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);
      mbar_init(&in_full[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_ptr, kTmemCols);
    tmem_relinquish_pair();
  }
  // Everything above (descriptor prefetch, barrier init, TMEM allocation) may overlap the tail of the previous kernel in
  // the stream (programmatic dependent launch); from here on its results are visible. Our own dependents may be
  // scheduled as soon as every CTA of this grid is running: they park in their prologue until this grid has finished.
  pdl_wait();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (threadIdx.x == 0) pdl_trigger();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own 128 pixels of A, own half of the weight tile ==============
    // The whole warp runs the (uniform) loop; one elected lane issues the barrier arrival and the TMA loads.
    {
      const bool leader_lane = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = (uint32_t)p.NB * p.RH * p.W * kBK * 2;
      const uint32_t tx_pair = 2 * (a_bytes + kBHalfBytes);
      const int kc_per_tap = p.Cin / kBK;
      for (int ptile = cid; ptile < pair_tiles; ptile += ncl) {
        const NTile nt = decode_item(p, ptile);
        const MTile t = decode_m(p, 2 * nt.pmt + (int)rank);
        const int b_row = nt.n0 + (int)rank * (nt.n_this >> 1);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dr = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int ds = (p.taps == 9) ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < kc_per_tap; ++kc) {
            mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
            if (leader_lane) {
              uint8_t* a_dst = smem + stage * kStageBytes;
              uint8_t* b_dst = a_dst + kATileBytes;
              const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
              // one arrival (the leader's) per phase; the peer's bytes only count towards the transaction total. They can
              // never reach a stale phase: the peer refills a stage only after the MMAs that consumed it have completed.
              if (rank == 0) mbar_expect_tx(&full_bar[stage], tx_pair);
              tma_load_4d_pair(a_dst, &tmap_a, full_leader, kc * kBK, ds, t.h0 + dr, t.b0);
              tma_load_2d_pair(b_dst, &tmap_b, full_leader, tap * p.Cin + kc * kBK, b_row);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores ==================
    // Warp-uniform loop (all lanes wait on the barriers), one elected lane issues the MMAs and commits.
    if (rank == 0) {
      const bool leader_lane = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      for (int ptile = cid; ptile < pair_tiles; ptile += ncl) {
        const uint32_t idesc = make_idesc_bf16(2 * kBM, decode_item(p, ptile).n_this, 0, 0);
        mbar_wait_bounded(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait_bounded(&full_bar[stage], phase);
          tc_fence_after();
          if (leader_lane) {
            const uint32_t a_addr = smem_base + stage * kStageBytes;
            const uint64_t a_desc = make_smem_desc_sw128(a_addr, 0, 1024);
            const uint64_t b_desc = make_smem_desc_sw128(a_addr + kATileBytes, 0, 1024);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_pair(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (leader_lane) umma_commit_pair(&tmem_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int half = (warp - 4) / (4 * kSub);       // which half of the tile's 64-channel chunks this warp works on
    const int sub = ((warp - 4) >> 2) % kSub;       // which 64/kSub columns of a chunk this warp owns
    const int m = q * 32 + lane;          // accumulator row == pixel within this CTA's tile
    const bool elected = (warp == 4 + 4 * kSub * half) && lane == 0;
    const int bar_half = 1 + half;        // named barrier of this half (128 threads)
    const int rows_in_tile = p.NB * p.RH * p.W;
    const uint32_t box_bytes = (uint32_t)rows_in_tile * 128;
    const bool row_ok = m < rows_in_tile;
    const int HW = p.H * p.W;
    uint8_t* buf0 = smem + kOffEpi + (half * kEpiBufsPerHalf + 0) * kEpiBufBytes;
    uint8_t* buf1 = buf0 + kEpiBufBytes;
    uint8_t* buf2 = buf1 + kEpiBufBytes;
    uint64_t* my_in_full = &in_full[half];
    const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);

    unsigned long long seed64 = ((unsigned long long)p.seed_hi << 32) | p.seed_lo;
    if (p.seed_ptr != nullptr) seed64 += *p.seed_ptr * 0x9E3779B97F4A7C15ull;
    const uint32_t dseed = dropout_seed((uint32_t)seed64, (uint32_t)(seed64 >> 32));
    const bool warp_one_image = (p.NB == 1) || (HW % 32 == 0);

    constexpr bool kHasIn = (EPI == EPI_AXPBY || EPI == EPI_MODSILU_BWD || EPI == EPI_SILU_BWD);
    constexpr bool kPingPong = (EPI == EPI_PLAIN || EPI == EPI_AXPBY || EPI == EPI_MODSILU_BWD);
    const bool use_res = (EPI == EPI_SILU_BWD) && p.res != nullptr;
    const bool use_old = (EPI == EPI_SILU_BWD) && p.accumulate_out;
    const bool two_sweep = (EPI == EPI_SILU_BWD) && p.nrm != nullptr;
    // split epilogue (ConvGemmArgs::split_c): chunks of the skip half never load an old value
    const int split_c = (EPI == EPI_SILU_BWD && p.split_c > 0) ? p.split_c : (1 << 30);
    auto chunk_old = [&](int c0) { return use_old && c0 < split_c; };
    // MODSILU: h ping-pongs between buf1 and buf2 like the outputs of the kPingPong epilogues; the raw copy (training only)
    // goes through buf0 and is written after h, behind a mid-chunk "previous stores have been read" hand-shake that has
    // had a whole chunk of work to complete
    const bool modsilu_raw = (EPI == EPI_MODSILU) && p.out2 != nullptr;

    auto tile_range = [&](int ptile, int& cb, int& ce) {
      const int nch = decode_item(p, ptile).n_this >> 6;
      const int h0n = (nch + 1) >> 1;
      cb = half == 0 ? 0 : h0n;
      ce = half == 0 ? h0n : nch;
    };
    auto advance = [&](ChunkPos& s) {
      if (s.cb < s.ce) {
        if (s.c + 1 < s.ce) { ++s.c; return; }
        if (two_sweep && s.pass == 0) { s.pass = 1; s.c = s.cb; return; }
      }
      s.ptile += ncl;
      s.pass = 0;
      if (s.ptile >= pair_tiles) { s.valid = false; return; }
      tile_range(s.ptile, s.cb, s.ce);
      s.c = s.cb;
    };
    // operands of chunk `s` -> shared memory (issued by the elected thread of the half)
    auto issue_loads = [&](const ChunkPos& s) {
      const NTile nt = decode_item(p, s.ptile);
      const MTile t = decode_m(p, 2 * nt.pmt + (int)rank);
      const int c0 = nt.n0 + s.c * 64;
      if constexpr (EPI == EPI_AXPBY || EPI == EPI_MODSILU_BWD) {
        mbar_expect_tx(my_in_full, box_bytes);
        tma_load_4d(buf0, &tmap_i0, my_in_full, c0, 0, t.h0, t.b0);
      } else if constexpr (EPI == EPI_SILU_BWD) {
        const bool ld_res = use_res && !(two_sweep && s.pass == 1);
        const bool ld_old = chunk_old(c0) && !(two_sweep && s.pass == 0);
        if (ld_old) bulk_wait_read0();   // the store that last read buf2 must be done before TMA overwrites it
        mbar_expect_tx(my_in_full, box_bytes * (1u + (ld_res ? 1u : 0u) + (ld_old ? 1u : 0u)));
        tma_load_4d(buf0, &tmap_i0, my_in_full, c0, 0, t.h0, t.b0);
        if (ld_res) tma_load_4d(buf1, &tmap_i1, my_in_full, c0, 0, t.h0, t.b0);
        if (ld_old) tma_load_4d(buf2, &tmap_o, my_in_full, c0, 0, t.h0, t.b0);
      }
    };

    ChunkPos pos;
    pos.ptile = cid;
    pos.pass = 0;
    pos.valid = pos.ptile < pair_tiles;
    pos.cb = pos.ce = pos.c = 0;
    if (pos.valid) {
      tile_range(pos.ptile, pos.cb, pos.ce);
      pos.c = pos.cb;
    }
    if (kHasIn && elected) {
      ChunkPos f = pos;
      while (f.valid && f.cb >= f.ce) advance(f);
      if (f.valid) issue_loads(f);
    }

    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t in_phase = 0;
    int qi = 0;             // chunks processed by this half (ping-pong index)
    int cur_tile = -1;
    // per-tile state
    MTile t{};
    int n0 = 0, n_this = 0;
    long long pix = 0;
    bool valid = false;
    int b = 0;
    uint32_t t_row = 0;
    float dot = 0.f, inv_n = 1.f, kk = 0.f;
    float res_scale = 1.f;  // AXPBY with nrm: the residual operand is the UN-normalised tensor, scaled by 1/nrm of its pixel

    int cur_mt = 0;         // M tile of this CTA within the current pair tile
    // column sums over this warp's 32 rows of the values as stored (bf16), fixed order: see ConvGemmArgs::col_partial
    auto col_sums = [&](const float (&v)[CW], int cc) {
      float cs[CW];
#pragma unroll
      for (int i = 0; i < CW; ++i) cs[i] = (row_ok && valid) ? bf16_round(v[i]) : 0.f;
      const float tot = warp_transpose_reduce<CW>(cs, lane);
      if (cur_mt < p.m_tiles && lane < CW && cc + lane < n_this)
        p.col_partial[((long long)cur_mt * 4 + q) * p.Cout + n0 + cc + lane] = tot;
    };

    while (pos.valid) {
      if (pos.ptile != cur_tile) {
        cur_tile = pos.ptile;
        const NTile nt = decode_item(p, pos.ptile);
        n0 = nt.n0;
        n_this = nt.n_this;
        cur_mt = 2 * nt.pmt + (int)rank;
        t = decode_m(p, cur_mt);
        pix = t.p_base + m;
        valid = row_ok && pix < t.p_limit;
        b = valid ? (int)(pix / HW) : 0;
        mbar_wait_bounded(&tmem_full[acc], acc_phase);
        tc_fence_after();
        t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBN;
        dot = 0.f;
        if (EPI == EPI_AXPBY && p.nrm != nullptr) res_scale = valid ? p.beta / p.nrm[pix] : 0.f;
        else res_scale = p.beta;
      }
      if (pos.cb < pos.ce) {
        const bool sweep1 = two_sweep && pos.pass == 0;   // accumulates the row dot, parks v in TMEM, writes nothing
        const bool writes_out = !sweep1;
        // a single output buffer that is not refilled by TMA needs an explicit "previous store has been read" hand-shake
        const bool cur_old = chunk_old(n0 + pos.c * 64);
        if (EPI == EPI_SILU_BWD && !cur_old && writes_out) {
          if (elected) bulk_wait_read0();
          named_bar_sync(bar_half, kHalfThreads);
        }
        if (kHasIn) {
          mbar_wait_bounded(my_in_full, in_phase);
          in_phase ^= 1;
        }
        uint8_t* obuf = (EPI == EPI_PLAIN) ? ((qi & 1) ? buf1 : buf0)
                        : ((kPingPong || EPI == EPI_MODSILU) ? ((qi & 1) ? buf2 : buf1) : buf2);
#pragma unroll 1
        for (int s = 0; s < kStepsPerWarp; ++s) {
          const int col_in_chunk = sub * (64 / kSub) + s * CW;
          const int cc = pos.c * 64 + col_in_chunk;   // first column of this step within the tile
          const int j0 = col_in_chunk / 8;
          uint32_t r[CW];
          tmem_ld_cw(t_row + cc, r);
          tmem_ld_wait();
          float v[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
          if constexpr (EPI == EPI_PLAIN) {
            if (row_ok) srow_store<CW>(obuf, m, j0, v);
            if (p.col_partial != nullptr) col_sums(v, cc);
          } else if constexpr (EPI == EPI_MODSILU) {
            // the reference's conv output is bf16 before the fp32 modulation island (networks.py:253-258)
#pragma unroll
            for (int i = 0; i < CW; ++i) v[i] = bf16_round(v[i]);
            if (row_ok) {
              const float* mrow = p.mod + (long long)b * p.mod_stride + n0 + cc;
#pragma unroll
              for (int g = 0; g < CW / 4; ++g) {
                const float4 mm = *reinterpret_cast<const float4*>(mrow + g * 4);
                v[g * 4 + 0] = mp_silu_f(v[g * 4 + 0] * mm.x);
                v[g * 4 + 1] = mp_silu_f(v[g * 4 + 1] * mm.y);
                v[g * 4 + 2] = mp_silu_f(v[g * 4 + 2] * mm.z);
                v[g * 4 + 3] = mp_silu_f(v[g * 4 + 3] * mm.w);
              }
              if (p.drop_p > 0.f) dropout_apply<CW>(v, (unsigned long long)pix * p.Cout + n0 + cc, p.drop_p, dseed);
              srow_store<CW>(obuf, m, j0, v);
            }
          } else if constexpr (EPI == EPI_AXPBY) {
            if (row_ok) {
              float rv[CW];
              srow_load<CW>(buf0, m, j0, rv);
#pragma unroll
              for (int i = 0; i < CW; ++i) v[i] += res_scale * rv[i];
              srow_store<CW>(obuf, m, j0, v);
            }
            if (p.col_partial != nullptr) col_sums(v, cc);
          } else if constexpr (EPI == EPI_MODSILU_BWD) {
            // backward of h = drop(mp_silu(raw * m)):  gz = g_h * keep/(1-p) * mp_silu'(raw*m);  g_raw = gz * m;
            // d_mod[b,c] += sum_pixels gz * raw
            float dm[CW];
            if (valid) {
              float rw[CW];
              srow_load<CW>(buf0, m, j0, rw);
              const float* mrow = p.mod + (long long)b * p.mod_stride + n0 + cc;
              if (p.drop_p > 0.f) dropout_apply<CW>(v, (unsigned long long)pix * p.Cout + n0 + cc, p.drop_p, dseed);
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
              for (int i = 0; i < CW; ++i) { dm[i] = 0.f; v[i] = 0.f; }
            }
            if (row_ok) srow_store<CW>(obuf, m, j0, v);
            if (warp_one_image) {
              const long long pw = t.p_base + q * 32;   // first pixel of this warp's rows
              const float tot = warp_transpose_reduce<CW>(dm, lane);
              if (q * 32 < rows_in_tile && pw < t.p_limit && lane < CW && cc + lane < n_this)
                atomicAdd(p.d_mod + (long long)(pw / HW) * p.mod_stride + n0 + cc + lane, tot);
            } else {
              // the warp's 32 rows straddle images (HW = 49, 16, ...): one masked transpose-reduce per image instead of
              // CW atomics per row (7x7 maps at B = 128 ran this epilogue at 393 TFLOP/s against 890 for its siblings)
              const int b_lo = __reduce_min_sync(0xffffffffu, valid ? b : 0x7fffffff);
              const int b_hi = __reduce_max_sync(0xffffffffu, valid ? b : -1);
              for (int bb = b_lo; bb <= b_hi; ++bb) {
                float part[CW];
#pragma unroll
                for (int i = 0; i < CW; ++i) part[i] = (valid && b == bb) ? dm[i] : 0.f;
                const float tot = warp_transpose_reduce<CW>(part, lane);
                if (lane < CW && cc + lane < n_this) atomicAdd(p.d_mod + (long long)bb * p.mod_stride + n0 + cc + lane, tot);
              }
            }
          } else if constexpr (EPI == EPI_SILU_BWD) {
            // g_x = alpha*acc * mp_silu'(x) + beta*res, [pixel-norm adjoint: g/n - x*dot/((n-eps) n C)], (+ old out)
            float xv[CW];
            if (row_ok) srow_load<CW>(buf0, m, j0, xv);
            else {
#pragma unroll
              for (int i = 0; i < CW; ++i) xv[i] = 0.f;
            }
            if (two_sweep && pos.pass == 1) {
              // v was parked in TMEM by the first sweep
#pragma unroll
              for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]) * inv_n - xv[i] * kk;
            } else {
#pragma unroll
              for (int i = 0; i < CW; ++i) v[i] *= mp_silu_grad_f(xv[i]);
              if (use_res && row_ok) {
                float rv[CW];
                srow_load<CW>(buf1, m, j0, rv);
#pragma unroll
                for (int i = 0; i < CW; ++i) v[i] += p.beta * rv[i];
              }
            }
            if (sweep1) {
#pragma unroll
              for (int i = 0; i < CW; ++i) {
                dot += v[i] * xv[i];
                r[i] = __float_as_uint(v[i]);
              }
              tmem_st_cw(t_row + cc, r);
            } else if (n0 + cc >= split_c) {
              // skip half of a decoder block's concatenated gradient: d(gain)*gain reduction over pixels, then * gain
              float dg[CW];
#pragma unroll
              for (int i = 0; i < CW; ++i) dg[i] = valid ? v[i] * xv[i] : 0.f;
              const int cs = n0 + cc - split_c;                // first channel within the skip tensor
              if (valid) {
                const float* grow_ = p.mod + (long long)b * p.mod_stride + cs;
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) {
                  const float4 gg = *reinterpret_cast<const float4*>(grow_ + g * 4);
                  v[g * 4 + 0] *= gg.x; v[g * 4 + 1] *= gg.y; v[g * 4 + 2] *= gg.z; v[g * 4 + 3] *= gg.w;
                }
              }
              if (row_ok) srow_store<CW>(buf2, m, j0, v);
              if (warp_one_image) {
                const long long pw = t.p_base + q * 32;   // first pixel of this warp's rows
                const float tot = warp_transpose_reduce<CW>(dg, lane);
                if (q * 32 < rows_in_tile && pw < t.p_limit && lane < CW && cc + lane < n_this)
                  atomicAdd(p.d_mod + (long long)(pw / HW) * p.mod_stride + cs + lane, tot);
              } else {
                const int b_lo = __reduce_min_sync(0xffffffffu, valid ? b : 0x7fffffff);
                const int b_hi = __reduce_max_sync(0xffffffffu, valid ? b : -1);
                for (int bb = b_lo; bb <= b_hi; ++bb) {
                  float part[CW];
#pragma unroll
                  for (int i = 0; i < CW; ++i) part[i] = (valid && b == bb) ? dg[i] : 0.f;
                  const float tot = warp_transpose_reduce<CW>(part, lane);
                  if (lane < CW && cc + lane < n_this) atomicAdd(p.d_mod + (long long)bb * p.mod_stride + cs + lane, tot);
                }
              }
            } else if (row_ok) {
              if (cur_old) {
                float ov[CW];
                srow_load<CW>(buf2, m, j0, ov);
#pragma unroll
                for (int i = 0; i < CW; ++i) v[i] += ov[i];
              }
              if (p.out_bias != nullptr && valid) {
                const int cb_ = p.split_c > 0 ? p.split_c : p.Cout;     // channels of `out`
                const float* brow = p.out_bias + (long long)b * cb_ + n0 + cc;
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) {
                  const float4 bb = *reinterpret_cast<const float4*>(brow + g * 4);
                  v[g * 4 + 0] = fmaf(p.out_bias_scale, bb.x, v[g * 4 + 0]);
                  v[g * 4 + 1] = fmaf(p.out_bias_scale, bb.y, v[g * 4 + 1]);
                  v[g * 4 + 2] = fmaf(p.out_bias_scale, bb.z, v[g * 4 + 2]);
                  v[g * 4 + 3] = fmaf(p.out_bias_scale, bb.w, v[g * 4 + 3]);
                }
              }
              srow_store<CW>(buf2, m, j0, v);
            }
          }
        }
        if (sweep1) tmem_st_wait();
        if (modsilu_raw) {
          // un-modulated conv output (bf16) for the backward pass: second, cheap read of the accumulator columns
          if (elected) bulk_wait_read0();
          named_bar_sync(bar_half, kHalfThreads);
#pragma unroll 1
          for (int s = 0; s < kStepsPerWarp; ++s) {
            const int col_in_chunk = sub * (64 / kSub) + s * CW;
            uint32_t r[CW];
            tmem_ld_cw(t_row + pos.c * 64 + col_in_chunk, r);
            tmem_ld_wait();
            float v[CW];
#pragma unroll
            for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
            if (row_ok) srow_store<CW>(buf0, m, col_in_chunk / 8, v);
          }
        }
        if (writes_out) fence_proxy_async_smem();
        // the store issued one chunk ago has released the other buffer
        if ((kPingPong || (EPI == EPI_MODSILU && !modsilu_raw)) && elected) bulk_wait_read0();
        named_bar_sync(bar_half, kHalfThreads);
        ChunkPos nx = pos;
        advance(nx);
        if (elected) {
          if (writes_out) {
            const int c0 = n0 + pos.c * 64;
            if (c0 >= split_c) tma_store_4d(&tmap_o2, obuf, c0 - split_c, 0, t.h0, t.b0);
            else tma_store_4d(&tmap_o, obuf, c0, 0, t.h0, t.b0);
            if (EPI == EPI_MODSILU && p.out2 != nullptr) tma_store_4d(&tmap_o2, buf0, c0, 0, t.h0, t.b0);
            bulk_commit();
          }
          if (kHasIn) {
            ChunkPos f = nx;
            while (f.valid && f.cb >= f.ce) advance(f);
            if (f.valid) issue_loads(f);
          }
        }
        ++qi;
        // end of the first sweep over this half's chunks: combine the two halves' partial dots of each row
        if (sweep1 && pos.c + 1 == pos.ce) {
          // 2 * kSub warps hold partial dots of the same 32 rows: first the sub-warps of a half combine, then the halves
          float tot = dot;
          for (int w2 = kSub - 1; w2 >= 1; --w2) {
            if (sub == w2) row_dots[half * 128 + m] = tot;
            named_bar_sync(3 + q, 64 * kSub);
            if (sub == w2 - 1) tot += row_dots[half * 128 + m];
            named_bar_sync(3 + q, 64 * kSub);
          }
          if (sub == 0) row_dots[half * 128 + m] = tot;
          named_bar_sync(3 + q, 64 * kSub);
          const float d2 = row_dots[m] + row_dots[128 + m];
          named_bar_sync(3 + q, 64 * kSub);
          const float n = valid ? p.nrm[pix] : 1.0f;
          inv_n = 1.0f / n;
          kk = d2 / (fmaxf(n - 1e-4f, 1e-20f) * (float)p.Cout);
        }
        pos = nx;
      } else {
        // this half has no chunk in this tile (64-channel tile): with two_sweep its dot share is zero
        if (two_sweep) {
          for (int w2 = kSub - 1; w2 >= 1; --w2) {
            named_bar_sync(3 + q, 64 * kSub);
            named_bar_sync(3 + q, 64 * kSub);
          }
          if (sub == 0) row_dots[half * 128 + m] = 0.f;
          named_bar_sync(3 + q, 64 * kSub);
          named_bar_sync(3 + q, 64 * kSub);
        }
        advance(pos);
      }
      if (!pos.valid || pos.ptile != cur_tile) {
        // tile drained by this warp: hand the accumulator buffer back to the MMA issuer (leader CTA)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc == 0 ? tmem_empty_leader0 : tmem_empty_leader1);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    if (elected) bulk_wait_read0();   // shared memory must outlive the stores' reads; their global writes complete with the grid
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled costs ~1 us and a step issues ~900 of them: cache by (pointer, geometry).
struct TmapKey {
  const void* base;
  int d0, d1, d2, d3, b0, b1, b2, b3, rank;
  bool operator==(const TmapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && b0 == o.b0 && b1 == o.b1 &&
           b2 == o.b2 && b3 == o.b3 && rank == o.rank;
  }
};
struct TmapEntry {
  TmapKey key;
  CUtensorMap map;
  bool used;
};
constexpr int kTmapCacheSize = 4096;
TmapEntry g_tmap_cache[kTmapCacheSize];
std::mutex g_tmap_mutex;

size_t tmap_hash(const TmapKey& k) {
  size_t h = reinterpret_cast<size_t>(k.base) * 0x9E3779B97F4A7C15ull;
  h ^= ((size_t)k.d0 * 31 + k.d1) * 0x85EBCA6Bull + ((size_t)k.d2 * 131 + k.d3) * 0xC2B2AE35ull;
  h ^= ((size_t)k.b0 * 7 + k.b1 * 13 + k.b2 * 17 + k.b3 * 19 + k.rank) * 0x27D4EB2Full;
  return (h >> 17) % kTmapCacheSize;
}

// NHWC activation map, box = {64 channels, W, RH, NB}
int nhwc_tmap(CUtensorMap* out, const void* base, int C, int W, int H, int B, int RH, int NB) {
  TmapKey k{base, C, W, H, B, 64, W, RH, NB, 4};
  const size_t slot = tmap_hash(k);
  {
    std::lock_guard<std::mutex> g(g_tmap_mutex);
    if (g_tmap_cache[slot].used && g_tmap_cache[slot].key == k) {
      *out = g_tmap_cache[slot].map;
      return 0;
    }
  }
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)W, (uint32_t)RH, (uint32_t)NB};
  if (encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
    return -1;
  std::lock_guard<std::mutex> g(g_tmap_mutex);
  g_tmap_cache[slot].key = k;
  g_tmap_cache[slot].map = *out;
  g_tmap_cache[slot].used = true;
  return 0;
}

int weight_tmap(CUtensorMap* out, const void* base, int K, int Cout, int box_rows) {
  TmapKey k{base, K, Cout, 0, 0, 64, box_rows, 0, 0, 2};
  const size_t slot = tmap_hash(k);
  {
    std::lock_guard<std::mutex> g(g_tmap_mutex);
    if (g_tmap_cache[slot].used && g_tmap_cache[slot].key == k) {
      *out = g_tmap_cache[slot].map;
      return 0;
    }
  }
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)Cout};
  uint64_t strides[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {64, (uint32_t)box_rows};
  if (encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0)
    return -1;
  std::lock_guard<std::mutex> g(g_tmap_mutex);
  g_tmap_cache[slot].key = k;
  g_tmap_cache[slot].map = *out;
  g_tmap_cache[slot].used = true;
  return 0;
}

template <int EPI>
int launch_pair(const CUtensorMap* maps, const ConvGemmParams& p, int clusters, cudaStream_t stream) {
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_pair_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  TEDM_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<EPI>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p));
  return 0;
}

}  // namespace

bool conv_pair_supported(const ConvGemmArgs& a) {
  if (a.ksize != 1 && a.ksize != 3) return false;
  if (a.Cin % 64 != 0 || a.Cout % 64 != 0 || a.Cout < 128) return false;
  if (a.epi == EPI_SILU_BWD && a.nrm != nullptr && a.Cout > kBN) return false;   // fused pixel-norm adjoint: one N tile
  int RH, NB;
  if (conv_tile_geometry(a.H, a.W, &RH, &NB) != 0) return false;
  return true;
}

// Rows of ConvGemmArgs::col_partial per image (0: the tile geometry does not keep a warp's 32 rows inside one image)
int conv_pair_colsum_slots(const ConvGemmArgs& a) {
  int RH, NB;
  if (conv_tile_geometry(a.H, a.W, &RH, &NB) != 0) return 0;
  const int HW = a.H * a.W;
  if (NB == 1) return ((a.H + RH - 1) / RH) * 4;
  if (HW % 32 == 0 && NB * HW == 128) return HW / 32;
  return 0;
}

// Pair tiles (256 pixels x <=256 channels) of the problem: a launch that cannot occupy even half of the 74 CTA pairs is
// better served by the single-CTA kernel's narrower tiles (shorter critical path per tile).
int conv_pair_tiles(const ConvGemmArgs& a) {
  int RH, NB;
  if (conv_tile_geometry(a.H, a.W, &RH, &NB) != 0) return 0;
  const int tiles_h = (a.H + RH - 1) / RH;
  const int m_tiles = (NB == 1) ? a.B * tiles_h : (a.B + NB - 1) / NB;
  return ((m_tiles + 1) / 2) * ((a.Cout + kBN - 1) / kBN);
}

int conv_pair_launch(const ConvGemmArgs& a, cudaStream_t stream) {
  TEDM_CHECK(conv_pair_supported(a), "conv_pair: unsupported problem (Cin %d, Cout %d, %dx%d, k%d)", a.Cin, a.Cout, a.H,
             a.W, a.ksize);
  TEDM_CHECK(a.B > 0, "conv_pair: empty input");
  ConvGemmParams p{};
  p.B = a.B; p.H = a.H; p.W = a.W; p.Cin = a.Cin; p.Cout = a.Cout;
  p.taps = a.ksize * a.ksize;
  conv_tile_geometry(a.H, a.W, &p.RH, &p.NB);
  p.tiles_h = (a.H + p.RH - 1) / p.RH;
  p.m_tiles = (p.NB == 1) ? a.B * p.tiles_h : (a.B + p.NB - 1) / p.NB;
  p.block_n = kBN;
  p.n_tiles = (a.Cout + kBN - 1) / kBN;
  p.k_blocks = p.taps * (a.Cin / 64);
  p.epi = a.epi; p.alpha = a.alpha; p.out = a.out; p.out2 = a.out2; p.res = a.res;
  p.beta = a.beta; p.mod = a.mod; p.mod_stride = a.mod_stride;
  p.drop_p = a.drop_p; p.seed_lo = (uint32_t)a.seed; p.seed_hi = (uint32_t)(a.seed >> 32); p.seed_ptr = a.seed_ptr;
  p.aux = a.aux; p.d_mod = a.d_mod; p.nrm = a.nrm; p.accumulate_out = a.accumulate_out;
  p.col_partial = (a.epi == EPI_PLAIN || a.epi == EPI_AXPBY) ? a.col_partial : nullptr;
  p.out_bias = a.epi == EPI_SILU_BWD ? a.out_bias : nullptr; p.out_bias_scale = a.out_bias_scale;
  p.split_c = a.epi == EPI_SILU_BWD ? a.split_c : 0;
  if (p.split_c > 0)
    TEDM_CHECK(p.split_c % 64 == 0 && p.split_c < a.Cout && a.out2 != nullptr && a.mod != nullptr && a.d_mod != nullptr &&
                   a.nrm == nullptr && a.mod_stride >= a.Cout - p.split_c,
               "conv_pair: split epilogue needs split_c %% 64 == 0, out2, gain, d_gx and no fused pixel norm");
  if (a.epi == EPI_MODSILU) TEDM_CHECK(a.mod != nullptr, "conv_pair: MODSILU epilogue needs mod");
  if (a.epi == EPI_AXPBY) TEDM_CHECK(a.res != nullptr, "conv_pair: AXPBY epilogue needs res");
  if (a.epi == EPI_MODSILU_BWD)
    TEDM_CHECK(a.mod != nullptr && a.aux != nullptr && a.d_mod != nullptr, "conv_pair: MODSILU_BWD epilogue needs mod, raw and d_mod");
  if (a.epi == EPI_SILU_BWD) TEDM_CHECK(a.aux != nullptr, "conv_pair: SILU_BWD epilogue needs x");

  CUtensorMap maps[6];
  if (nhwc_tmap(&maps[0], a.x, a.Cin, a.W, a.H, a.B, p.RH, p.NB) != 0) return -1;
  if (weight_tmap(&maps[1], a.w, p.taps * a.Cin, a.Cout, kBN / 2) != 0) return -1;
  if (nhwc_tmap(&maps[2], a.out, p.split_c > 0 ? p.split_c : a.Cout, a.W, a.H, a.B, p.RH, p.NB) != 0) return -1;
  maps[3] = maps[2]; maps[4] = maps[2]; maps[5] = maps[2];
  if (a.epi == EPI_MODSILU && a.out2 != nullptr && nhwc_tmap(&maps[3], a.out2, a.Cout, a.W, a.H, a.B, p.RH, p.NB) != 0) return -1;
  if (p.split_c > 0 && nhwc_tmap(&maps[3], a.out2, a.Cout - p.split_c, a.W, a.H, a.B, p.RH, p.NB) != 0) return -1;
  const void* in0 = a.epi == EPI_AXPBY ? (const void*)a.res : (const void*)a.aux;
  if ((a.epi == EPI_AXPBY || a.epi == EPI_MODSILU_BWD || a.epi == EPI_SILU_BWD) &&
      nhwc_tmap(&maps[4], in0, a.Cout, a.W, a.H, a.B, p.RH, p.NB) != 0) return -1;
  if (a.epi == EPI_SILU_BWD && a.res != nullptr && nhwc_tmap(&maps[5], a.res, a.Cout, a.W, a.H, a.B, p.RH, p.NB) != 0) return -1;

  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  int clusters = num_sms() / 2;
  if (clusters > pair_tiles) clusters = pair_tiles;
  // tail split (see decode_item): the remainder tiles of the last wave become half-N items when that fills the wave
  // better. Not with the fused pixel-norm adjoint, whose row dot needs all channels of a pixel in one CTA.
  p.split_from = p.work_items = pair_tiles;
  const int rem = pair_tiles % clusters;
  static const bool tail_split = [] { const char* e = getenv("TEDM_CONV_TAIL_SPLIT"); return !(e != nullptr && e[0] == '0'); }();
  if (tail_split && rem > 0 && 2 * rem <= clusters && pair_tiles > clusters && a.Cout % kBN == 0 &&
      !(a.epi == EPI_SILU_BWD && a.nrm != nullptr)) {
    p.split_from = pair_tiles - rem;
    p.work_items = pair_tiles + rem;
  }
  switch (a.epi) {
    case EPI_PLAIN: return launch_pair<EPI_PLAIN>(maps, p, clusters, stream);
    case EPI_MODSILU: return launch_pair<EPI_MODSILU>(maps, p, clusters, stream);
    case EPI_AXPBY: return launch_pair<EPI_AXPBY>(maps, p, clusters, stream);
    case EPI_MODSILU_BWD: return launch_pair<EPI_MODSILU_BWD>(maps, p, clusters, stream);
    case EPI_SILU_BWD: return launch_pair<EPI_SILU_BWD>(maps, p, clusters, stream);
    default: return fail("conv_pair: unknown epilogue %d", a.epi);
  }
}

}  // namespace tedm
