// Cosine-normalised self-attention on tcgen05/TMEM for EVERY (head_dim, S) the reference's configs use — the generic
// companion of attention_tc.cu (which is specialised for head_dim 64 and S in {64, 256}):
//     MNIST            head_dim  64, S = 196     head_dim 128, S = 49
//     ImageNet latent  head_dim 144, S = 256     head_dim 192, S = 64
// and, with the same code, any head_dim that is a multiple of 16 up to 192 and any S <= 256.
//
// Reference: CosineAttention.forward, src/tinyedm/networks.py:191-207 — pixel_norm over hd of q, k and v (:195), then
// F.scaled_dot_product_attention(q, k, v) with scale 1/sqrt(hd) (:201), output channel = head*hd + d (:202); backward =
// torch autograd of the same. qkv is (B,S,3C) bf16 with channel = {q,k,v}*C + head*hd + d.
//
// Three kernels:
//   qkv_norm_kernel     pixel_norm of every (pixel, {q,k,v}, head) row, ONCE per layer: qn (bf16, the reference casts the
//                       normalised tensor to bf16 too) and n = eps + rms (fp32, for the adjoint). Pure HBM streaming.
//   attn_fwd_gen_kernel one CTA per (image, head, tile of 128 queries). Q tile resident; K then V stream through a 4-stage
//                       TMA ring in 64-row chunks; S = Q K^T for all keys in TMEM (<= 256 columns), softmax straight out of
//                       TMEM (thread = query row), bf16 P through a 2-slot ring, O += P V into further TMEM columns.
//   attn_bwd_gen_kernel<MODE>  MODE 0: one CTA per 128-QUERY tile, streams (K, V) chunks, produces dQ and delta;
//                       MODE 1: one CTA per 128-KEY tile, streams (Q, dO) chunks, produces dK and dV. Per chunk two
//                       "score" MMAs (S = A X^T, dP = B Y^T, N = 64) into TMEM, 256 threads turn them into bf16 P / dS
//                       chunks, which feed the accumulating MMAs with the streamed chunk as the MN-major B operand.
//                       The pixel-norm adjoint g_u = g/n - y (g.y)/((n - eps) hd) of each result row runs in the epilogue.
// Shared-memory tiles are 64-channel "slabs": [rows][64 bf16] with 128-byte rows, 128B-swizzled exactly as TMA writes
// them, so every MMA descriptor is the proven single-atom form of attention_tc.cu; head_dim 144 = slabs of 64 + 64 + 16
// channels (the MMAs of the last slab run K = 16 resp. N = 16; the rest of that slab is never read).
// Rows beyond S inside a 64-row chunk (S = 196 / 49) belong to the next image or are TMA zero fill: keys beyond S are
// masked to probability 0, query rows beyond S are computed and dropped.
#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr float kEpsG = 1e-4f;
constexpr float kLog2eG = 1.4426950408889634f;
constexpr float kLn2G = 0.6931471805599453f;
constexpr int kSlab128 = 128 * 128;   // bytes of one slab of a 128-row tile
constexpr int kSlab64 = 64 * 128;     // bytes of one slab of a 64-row chunk
constexpr int kRingTile = 128 * 128;  // [128 rows][64 bf16]

__device__ __forceinline__ uint4* srow(uint8_t* slab, int r, int j) {
  return reinterpret_cast<uint4*>(slab + r * 128 + ((j ^ (r & 7)) << 4));
}

// ---------------------------------------------------------------------------------------------------------------------
// pixel_norm of the q, k, v rows (networks.py:195): one warp per (pixel, plane, head) group of hd contiguous channels
// ---------------------------------------------------------------------------------------------------------------------
// A group of hd channels is hd / 8 16-byte chunks (8 for hd = 64, 16 for 128, 18 for 144, 24 for 192): a warp serves
// 32 / SEG groups at once, SEG = 8, 16 or 32 lanes per group, so that (almost) every lane has a load in flight.
template <int SEG>
__global__ void __launch_bounds__(256)
qkv_norm_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ qn, float* __restrict__ norms,
                long long n_groups, int hd) {
  pdl_trigger();
  pdl_wait();
  constexpr int GPW = 32 / SEG;          // groups per warp
  const int lane = threadIdx.x & 31;
  const int li = lane % SEG;
  const long long g = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * GPW + lane / SEG;
  const int chunks = hd >> 3;            // 16-byte chunks per row (hd % 8 == 0)
  const bool on = g < n_groups && li < chunks;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (on) v = reinterpret_cast<const uint4*>(qkv + g * hd)[li];
  const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
  float ss = a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
#pragma unroll
  for (int o = SEG / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float n = kEpsG + sqrtf(ss / (float)hd);
  const float inv = 1.0f / n;
  if (on) {
    uint4 o;
    o.x = pack_bf16(a.x * inv, a.y * inv); o.y = pack_bf16(b.x * inv, b.y * inv);
    o.z = pack_bf16(c.x * inv, c.y * inv); o.w = pack_bf16(d.x * inv, d.y * inv);
    reinterpret_cast<uint4*>(qn + g * hd)[li] = o;
    if (li == 0) norms[g] = n;
  }
}

// Tile geometry. Normal mode: a tile = 128 consecutive rows (queries or keys) of ONE (image, head) pair, streamed chunk c =
// rows 64c.. of the same pair. Packed mode (S <= 64: MNIST 7x7, ImageNet-latent 8x8, CIFAR 8x8): a tile = TWO pairs, rows
// 0-63 the first and 64-127 the second (each padded to 64 rows), chunk c = the rows of pair c, and a row only sees the
// columns of its own pair (block-diagonal mask) — half the CTAs and twice the useful rows per MMA.
struct Geom {
  int S, heads, n_pairs, pair0, t_off;
  bool packed;
  __device__ __forceinline__ int pair_clamped(int p) const { return p < n_pairs ? p : n_pairs - 1; }
  // (pair, row inside the pair, valid) of tile row m
  __device__ __forceinline__ int row_pair(int m) const { return packed ? pair_clamped(pair0 + (m >> 6)) : pair0; }
  __device__ __forceinline__ int row_idx(int m) const { return packed ? (m & 63) : t_off + m; }
  __device__ __forceinline__ bool row_live(int m) const {
    return packed ? (pair0 + (m >> 6) < n_pairs && (m & 63) < S) : (t_off + m < S);
  }
  // is column j of chunk c a real key/query for tile row m
  __device__ __forceinline__ bool col_live(int m, int c, int j) const {
    return packed ? (c == (m >> 6) && j < S) : (c * 64 + j < S);
  }
  // pair and first tensor row of 64-row box `hh` of the resident tile / of streamed chunk c
  __device__ __forceinline__ int box_pair(int hh) const { return packed ? pair_clamped(pair0 + hh) : pair0; }
  __device__ __forceinline__ int res_row(int hh) const {
    const int p = box_pair(hh);
    return (p / heads) * S + (packed ? 0 : t_off + 64 * hh);
  }
  __device__ __forceinline__ int chunk_row(int c) const {
    const int p = box_pair(c);
    return (p / heads) * S + (packed ? 0 : 64 * c);
  }
  __device__ __forceinline__ int box_head(int hh) const { const int p = box_pair(hh); return p - (p / heads) * heads; }
};

__device__ __forceinline__ Geom make_geom(int S, int heads, int n_pairs) {
  Geom g;
  g.S = S; g.heads = heads; g.n_pairs = n_pairs;
  g.packed = S <= 64;
  if (g.packed) {
    g.pair0 = blockIdx.x * 2;
    g.t_off = 0;
  } else {
    const int tiles = (S + 127) >> 7;
    g.pair0 = blockIdx.x / tiles;
    g.t_off = (blockIdx.x - g.pair0 * tiles) * 128;
  }
  return g;
}

// number of 16-wide k steps / valid columns of slab s for head dim hd
__device__ __forceinline__ int slab_cols(int hd, int s) { return min(64, hd - 64 * s); }

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kFwdStages = 4;
constexpr int kFwdThreads = 192;   // 4 softmax warps + TMA warp + MMA warp

template <int NSLAB>
struct FwdLayout {
  static constexpr int kOffQ = 0;
  static constexpr int kOffRing = NSLAB * kSlab128;
  static constexpr int kStageBytes = NSLAB * kSlab64;
  static constexpr int kOffP = kOffRing + kFwdStages * kStageBytes;
  static constexpr int kOffBars = kOffP + 2 * kRingTile;
  static constexpr int kSmem = kOffBars + 256;
};

// head_dim <= 64 (NSLAB == 1): O (<= 64 columns) lands in the score columns of key chunk 0, which the softmax has consumed
// before the first P V MMA is issued, so 256 TMEM columns and 80 KB of shared memory suffice and TWO CTAs share an SM —
// the load / softmax / epilogue phases of one overlap the MMAs of the other (these tiles are far too small to fill an SM).
template <int NSLAB>
__global__ void __launch_bounds__(kFwdThreads, NSLAB == 1 ? 2 : 1)
attn_fwd_gen_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ y, float* __restrict__ lse,
                    int S, int hd, int heads, int n_pairs, float scale) {
  using L = FwdLayout<NSLAB>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBars);
  uint64_t* bar_q = bars;                 // Q tile landed
  uint64_t* full = bars + 1;              // [kFwdStages] chunk landed
  uint64_t* empty = full + kFwdStages;    // [kFwdStages] the MMAs reading the stage completed
  uint64_t* bar_s = empty + kFwdStages;   // all scores in TMEM
  uint64_t* p_ready = bar_s + 1;          // [2] 128 softmax threads wrote the P slot
  uint64_t* p_free = p_ready + 2;         // [2] the MMAs reading the P slot completed
  uint64_t* bar_o = p_free + 2;           // O complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_o + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) pdl_trigger();
  const int C = heads * hd;
  const Geom geo = make_geom(S, heads, n_pairs);
  const int NC = geo.packed ? 2 : (S + 63) >> 6;           // 64-key chunks

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      mbar_init(bar_q, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      for (int i = 0; i < kFwdStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&p_ready[i], 128); mbar_init(&p_free[i], 1); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, NSLAB == 1 ? 256 : 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_o = NSLAB == 1 ? tmem_base : tmem_base + 256;
  pdl_wait();

  if (warp == 4) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      mbar_expect_tx(bar_q, NSLAB * kSlab128);
      for (int s = 0; s < NSLAB; ++s)
        for (int hh = 0; hh < 2; ++hh)
          tma_load_2d(smem + L::kOffQ + s * kSlab128 + hh * kSlab64, &tmap_qkv, bar_q, geo.box_head(hh) * hd + 64 * s, geo.res_row(hh));
      for (int i = 0; i < 2 * NC; ++i) {
        const int st = i % kFwdStages;
        if (i >= kFwdStages) mbar_wait_bounded(&empty[st], ((i / kFwdStages) - 1) & 1);
        const int plane = i < NC ? 1 : 2, c = i < NC ? i : i - NC;
        mbar_expect_tx(&full[st], L::kStageBytes);
        for (int s = 0; s < NSLAB; ++s)
          tma_load_2d(smem + L::kOffRing + st * L::kStageBytes + s * kSlab64, &tmap_qkv, &full[st],
                      plane * C + geo.box_head(c) * hd + 64 * s, geo.chunk_row(c));
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      mbar_wait_bounded(bar_q, 0);
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      for (int c = 0; c < NC; ++c) {
        const int st = c % kFwdStages;
        mbar_wait_bounded(&full[st], (c / kFwdStages) & 1);
        tc_fence_after();
        bool first = true;
        for (int s = 0; s < NSLAB; ++s) {
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + L::kOffQ + s * kSlab128), 0, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + L::kOffRing + st * L::kStageBytes + s * kSlab64), 0, 1024);
          const int ks = slab_cols(hd, s) >> 4;
          for (int k = 0; k < ks; ++k) {
            umma_bf16(tmem_base + 64 * c, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s, first ? 0u : 1u);
            first = false;
          }
        }
        umma_commit(&empty[st]);
      }
      umma_commit(bar_s);
      for (int c = 0; c < NC; ++c) {
        const int i = NC + c, st = i % kFwdStages, slot = c & 1;
        mbar_wait_bounded(&full[st], (i / kFwdStages) & 1);
        mbar_wait_bounded(&p_ready[slot], (c >> 1) & 1);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + L::kOffP + slot * kRingTile), 0, 1024);
        for (int s = 0; s < NSLAB; ++s) {
          // V chunk slab: MN-major B operand (N = channels contiguous, K = 16 keys = 16 rows of 128 B per MMA)
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + L::kOffRing + st * L::kStageBytes + s * kSlab64), kSlab64, 1024);
          const uint32_t idesc_o = make_idesc_bf16(128, slab_cols(hd, s), 0, 1);
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_o + 64 * s, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 128), idesc_o, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[st]);
        umma_commit(&p_free[slot]);
      }
      umma_commit(bar_o);
    }
  } else {
    // ---------------- softmax + output (thread = query row) ----------------
    const int m = threadIdx.x;                       // 0..127
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sc = scale * kLog2eG;
    mbar_wait_bounded(bar_s, 0);
    tc_fence_after();
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
    for (int c0 = 0; c0 < NC * 64; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(t_row + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (geo.col_live(m, c0 >> 6, (c0 & 63) + i)) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
    }
    const float mxs = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sc;
    float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = 0; c < NC; ++c) {
      const int slot = c & 1;
      uint8_t* pbuf = smem + L::kOffP + slot * kRingTile;
      if (c >= 2) mbar_wait_bounded(&p_free[slot], ((c >> 1) - 1) & 1);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 64 + h * 32, r);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pv[i] = geo.col_live(m, c, h * 32 + i) ? exp2f(fmaf(__uint_as_float(r[i]), sc, -mxs)) : 0.f;
          sum4[i & 3] += pv[i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16(pv[g * 8 + 0], pv[g * 8 + 1]); o.y = pack_bf16(pv[g * 8 + 2], pv[g * 8 + 3]);
          o.z = pack_bf16(pv[g * 8 + 4], pv[g * 8 + 5]); o.w = pack_bf16(pv[g * 8 + 6], pv[g * 8 + 7]);
          *srow(pbuf, m, h * 4 + g) = o;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[slot]);
    }
    mbar_wait_bounded(bar_o, 0);
    tc_fence_after();
    const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
    const float inv = 1.0f / sum;
    const bool live = geo.row_live(m);
    const int pair = geo.row_pair(m);
    const int q = live ? geo.row_idx(m) : 0;
    const int b = pair / heads, head = pair - b * heads;
    __nv_bfloat16* dst = y + ((long long)(b * S + q)) * C + head * hd;
    const uint32_t t_o = tmem_o + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int col = 0; col < hd; col += 32) {
      const int s = col >> 6;
      const uint32_t taddr = t_o + 64 * s + (col & 63);
      if (hd - col >= 32) {
        uint32_t r[32];
        tmem_ld32(taddr, r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[g * 8 + 0]) * inv, __uint_as_float(r[g * 8 + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(r[g * 8 + 2]) * inv, __uint_as_float(r[g * 8 + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(r[g * 8 + 4]) * inv, __uint_as_float(r[g * 8 + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(r[g * 8 + 6]) * inv, __uint_as_float(r[g * 8 + 7]) * inv);
            reinterpret_cast<uint4*>(dst + col)[g] = o;
          }
        }
      } else {
        uint32_t r[16];
        tmem_ld16(taddr, r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[g * 8 + 0]) * inv, __uint_as_float(r[g * 8 + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(r[g * 8 + 2]) * inv, __uint_as_float(r[g * 8 + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(r[g * 8 + 4]) * inv, __uint_as_float(r[g * 8 + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(r[g * 8 + 6]) * inv, __uint_as_float(r[g * 8 + 7]) * inv);
            reinterpret_cast<uint4*>(dst + col)[g] = o;
          }
        }
      }
    }
    if (live && lse != nullptr) lse[(long long)pair * S + q] = (mxs + log2f(sum)) * kLn2G;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NSLAB == 1 ? 256 : 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 320;   // 8 compute warps (two per TMEM lane quarter: columns 0-31 / 32-63 of a chunk) + TMA + MMA
constexpr int kBwdStages = 2;

template <int NSLAB, int MODE>
struct BwdLayout {
  // score buffers in TMEM == ring slots in shared memory. One when TMEM is short (MODE 1 with three slabs: 128 + 2 x 192
  // columns) and for head_dim <= 64, where 256 columns and <= 100 KB per CTA let two CTAs share an SM instead
  static constexpr int kNB = (NSLAB == 1 || (MODE == 1 && NSLAB == 3)) ? 1 : 2;
  static constexpr int kTmemCols = NSLAB == 1 ? 256 : 512;
  static constexpr int kRings = MODE == 1 ? 2 : 1;                  // MODE 1 keeps P^T and dS^T
  static constexpr int kOffA = 0;                                   // resident tile A (MODE 0: Qn, MODE 1: Kn)
  static constexpr int kOffB = NSLAB * kSlab128;                    // resident tile B (MODE 0: dO, MODE 1: Vn)
  static constexpr int kOffStage = 2 * NSLAB * kSlab128;
  static constexpr int kStageBytes = 2 * NSLAB * kSlab64;           // X chunk then Y chunk
  static constexpr int kOffRing = kOffStage + kBwdStages * kStageBytes;
  static constexpr int kOffVec = kOffRing + kRings * kNB * kRingTile;   // lse2[256], delta[256] (MODE 1)
  static constexpr int kOffBars = kOffVec + 2048;
  static constexpr int kSmem = kOffBars + 256;
  // TMEM columns: kNB x (S 64 | dP 64), then the accumulators (NSLAB x 64 each)
  static constexpr int kTmemAcc = kNB * 128;
};

// pixel-norm adjoint of one gradient row of hd columns sitting in TMEM at t_acc (this thread's lane): two sweeps
// (dot product, then the result), y = the normalised row in the resident slabs
template <int NSLAB>
__device__ __forceinline__ void norm_adjoint_row(uint32_t t_acc, uint8_t* res, int m, int hd, float n, bool live,
                                                 __nv_bfloat16* dst) {
  float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int col = 0; col < hd; col += 16) {
    uint32_t r[16];
    tmem_ld16(t_acc + 64 * (col >> 6) + (col & 63), r);
    tmem_ld_wait();
    uint8_t* slab = res + (col >> 6) * kSlab128;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const uint4 u = *srow(slab, m, ((col & 63) >> 3) + g);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      d4[0] += __uint_as_float(r[g * 8 + 0]) * a.x + __uint_as_float(r[g * 8 + 4]) * c.x;
      d4[1] += __uint_as_float(r[g * 8 + 1]) * a.y + __uint_as_float(r[g * 8 + 5]) * c.y;
      d4[2] += __uint_as_float(r[g * 8 + 2]) * b.x + __uint_as_float(r[g * 8 + 6]) * d.x;
      d4[3] += __uint_as_float(r[g * 8 + 3]) * b.y + __uint_as_float(r[g * 8 + 7]) * d.y;
    }
  }
  const float dot = (d4[0] + d4[1]) + (d4[2] + d4[3]);
  const float inv_n = 1.0f / n;
  const float kk = dot / (fmaxf(n - kEpsG, 1e-20f) * (float)hd);
#pragma unroll 1
  for (int col = 0; col < hd; col += 16) {
    uint32_t r[16];
    tmem_ld16(t_acc + 64 * (col >> 6) + (col & 63), r);
    tmem_ld_wait();
    uint8_t* slab = res + (col >> 6) * kSlab128;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const uint4 u = *srow(slab, m, ((col & 63) >> 3) + g);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      uint4 o;
      o.x = pack_bf16(__uint_as_float(r[g * 8 + 0]) * inv_n - a.x * kk, __uint_as_float(r[g * 8 + 1]) * inv_n - a.y * kk);
      o.y = pack_bf16(__uint_as_float(r[g * 8 + 2]) * inv_n - b.x * kk, __uint_as_float(r[g * 8 + 3]) * inv_n - b.y * kk);
      o.z = pack_bf16(__uint_as_float(r[g * 8 + 4]) * inv_n - c.x * kk, __uint_as_float(r[g * 8 + 5]) * inv_n - c.y * kk);
      o.w = pack_bf16(__uint_as_float(r[g * 8 + 6]) * inv_n - d.x * kk, __uint_as_float(r[g * 8 + 7]) * inv_n - d.y * kk);
      if (live) reinterpret_cast<uint4*>(dst + col)[g] = o;
    }
  }
}

template <int NSLAB, int MODE>
__global__ void __launch_bounds__(kBwdThreads, NSLAB == 1 ? 2 : 1)
attn_bwd_gen_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                    const __nv_bfloat16* __restrict__ y, const float* __restrict__ norms, const float* __restrict__ lse,
                    float* __restrict__ delta, __nv_bfloat16* __restrict__ g_qkv, int S, int hd, int heads, int n_pairs,
                    float scale) {
  using L = BwdLayout<NSLAB, MODE>;
  constexpr int NB = L::kNB;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBars);
  uint64_t* bar_res = bars;                 // resident tiles landed
  uint64_t* full = bars + 1;                // [2] stage landed
  uint64_t* empty = full + kBwdStages;      // [2] the accumulating MMAs reading the stage completed
  uint64_t* sp_ready = empty + kBwdStages;  // [NB] scores of a chunk in TMEM
  uint64_t* sp_free = sp_ready + 2;         // [NB] 256 threads finished reading the score buffer
  uint64_t* ring_ready = sp_free + 2;       // [NB] 256 threads wrote the ring slot(s)
  uint64_t* ring_free = ring_ready + 2;     // [NB] the MMAs reading the ring slot completed
  uint64_t* bar_o = ring_free + 2;          // accumulators complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_o + 1);
  float* lse_s = reinterpret_cast<float*>(smem + L::kOffVec);
  float* dl_s = lse_s + 256;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) pdl_trigger();
  const int C = heads * hd;
  const Geom geo = make_geom(S, heads, n_pairs);            // tile rows: queries (MODE 0) / keys (MODE 1)
  const int NC = geo.packed ? 2 : (S + 63) >> 6;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      tma_prefetch_desc(&tmap_do);
      mbar_init(bar_res, 1);
      mbar_init(bar_o, 1);
      for (int i = 0; i < kBwdStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sp_ready[i], 1); mbar_init(&sp_free[i], 256);
        mbar_init(&ring_ready[i], 256); mbar_init(&ring_free[i], 1);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, L::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_acc = tmem_base + L::kTmemAcc;
  pdl_wait();

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      mbar_expect_tx(bar_res, 2 * NSLAB * kSlab128);
      for (int s = 0; s < NSLAB; ++s)
        for (int hh = 0; hh < 2; ++hh) {
          const int r = geo.res_row(hh);
          const int head = geo.box_head(hh);
          if (MODE == 0) {
            tma_load_2d(smem + L::kOffA + s * kSlab128 + hh * kSlab64, &tmap_qkv, bar_res, head * hd + 64 * s, r);
            tma_load_2d(smem + L::kOffB + s * kSlab128 + hh * kSlab64, &tmap_do, bar_res, head * hd + 64 * s, r);
          } else {
            tma_load_2d(smem + L::kOffA + s * kSlab128 + hh * kSlab64, &tmap_qkv, bar_res, C + head * hd + 64 * s, r);
            tma_load_2d(smem + L::kOffB + s * kSlab128 + hh * kSlab64, &tmap_qkv, bar_res, 2 * C + head * hd + 64 * s, r);
          }
        }
      for (int c = 0; c < NC; ++c) {
        const int st = c % kBwdStages;
        if (c >= kBwdStages) mbar_wait_bounded(&empty[st], ((c / kBwdStages) - 1) & 1);
        mbar_expect_tx(&full[st], L::kStageBytes);
        uint8_t* xs = smem + L::kOffStage + st * L::kStageBytes;
        uint8_t* ys = xs + NSLAB * kSlab64;
        const int head = geo.box_head(c), crow = geo.chunk_row(c);
        for (int s = 0; s < NSLAB; ++s) {
          if (MODE == 0) {   // X = K chunk, Y = V chunk
            tma_load_2d(xs + s * kSlab64, &tmap_qkv, &full[st], C + head * hd + 64 * s, crow);
            tma_load_2d(ys + s * kSlab64, &tmap_qkv, &full[st], 2 * C + head * hd + 64 * s, crow);
          } else {           // X = Q chunk, Y = dO chunk
            tma_load_2d(xs + s * kSlab64, &tmap_qkv, &full[st], head * hd + 64 * s, crow);
            tma_load_2d(ys + s * kSlab64, &tmap_do, &full[st], head * hd + 64 * s, crow);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      mbar_wait_bounded(bar_res, 0);
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      for (int c = 0; c <= NC; ++c) {
        if (c < NC) {
          const int st = c % kBwdStages, buf = c % NB;
          mbar_wait_bounded(&full[st], (c / kBwdStages) & 1);
          if (c >= NB) mbar_wait_bounded(&sp_free[buf], ((c / NB) - 1) & 1);
          tc_fence_after();
          uint8_t* xs = smem + L::kOffStage + st * L::kStageBytes;
          uint8_t* ys = xs + NSLAB * kSlab64;
          for (int which = 0; which < 2; ++which) {      // 0: S = A X^T, 1: dP = B Y^T
            bool first = true;
            for (int s = 0; s < NSLAB; ++s) {
              const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + (which ? L::kOffB : L::kOffA) + s * kSlab128), 0, 1024);
              const uint64_t b_desc = make_smem_desc_sw128(smem_u32((which ? ys : xs) + s * kSlab64), 0, 1024);
              const int ks = slab_cols(hd, s) >> 4;
              for (int k = 0; k < ks; ++k) {
                umma_bf16(tmem_base + buf * 128 + which * 64, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc_s,
                          first ? 0u : 1u);
                first = false;
              }
            }
          }
          umma_commit(&sp_ready[buf]);
        }
        if (c >= 1) {
          const int cc = c - 1, st = cc % kBwdStages, slot = cc % NB;
          mbar_wait_bounded(&ring_ready[slot], (cc / NB) & 1);
          tc_fence_after();
          uint8_t* xs = smem + L::kOffStage + st * L::kStageBytes;
          uint8_t* ys = xs + NSLAB * kSlab64;
          uint8_t* ring0 = smem + L::kOffRing + slot * kRingTile;                       // MODE 0: dS; MODE 1: P^T
          uint8_t* ring1 = smem + L::kOffRing + (NB + slot) * kRingTile;                // MODE 1: dS^T
          for (int s = 0; s < NSLAB; ++s) {
            const uint32_t idesc_o = make_idesc_bf16(128, slab_cols(hd, s), 0, 1);
            if (MODE == 0) {   // dQ += dS K_c
              const uint64_t a_desc = make_smem_desc_sw128(smem_u32(ring0), 0, 1024);
              const uint64_t b_desc = make_smem_desc_sw128(smem_u32(xs + s * kSlab64), kSlab64, 1024);
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_acc + 64 * s, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 128), idesc_o, (cc | k) != 0 ? 1u : 0u);
            } else {           // dV += P^T dO_c ; dK += dS^T Q_c
              const uint64_t a_p = make_smem_desc_sw128(smem_u32(ring0), 0, 1024);
              const uint64_t a_s = make_smem_desc_sw128(smem_u32(ring1), 0, 1024);
              const uint64_t b_do = make_smem_desc_sw128(smem_u32(ys + s * kSlab64), kSlab64, 1024);
              const uint64_t b_q = make_smem_desc_sw128(smem_u32(xs + s * kSlab64), kSlab64, 1024);
              for (int k = 0; k < 4; ++k) {
                umma_bf16(tmem_acc + 64 * (NSLAB + s), a_p + (uint64_t)(k * 2), b_do + (uint64_t)(k * 128), idesc_o, (cc | k) != 0 ? 1u : 0u);
                umma_bf16(tmem_acc + 64 * s, a_s + (uint64_t)(k * 2), b_q + (uint64_t)(k * 128), idesc_o, (cc | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty[st]);
          umma_commit(&ring_free[slot]);
        }
      }
      umma_commit(bar_o);
    }
  } else {
    // ---------------- 8 compute warps ----------------
    const int m = threadIdx.x & 127;              // row of the tile (query in MODE 0, key in MODE 1)
    const int half = threadIdx.x >> 7;            // columns [32 half, 32 half + 32) of every 64-wide chunk
    const int quarter = warp & 3;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = scale * kLog2eG;
    const bool live = geo.row_live(m);
    const int pair = geo.row_pair(m);
    const int b = pair / heads, head = pair - b * heads;
    const int row0 = b * S;
    const int r_idx = geo.row_idx(m);             // row index inside the (image, head)
    const int r_cl = r_idx < S ? r_idx : S - 1;
    float ls2 = 0.f, dl = 0.f;
    if (MODE == 0) {
      // delta = sum_d dO * O of this query row (O straight from global, dO from the resident tile)
      ls2 = lse[(long long)pair * S + r_cl] * kLog2eG;
      const __nv_bfloat16* orow = y + ((long long)(row0 + r_cl)) * C + head * hd;
      mbar_wait_bounded(bar_res, 0);
      float acc = 0.f;
      for (int col = 0; col < hd; col += 8) {
        const uint4 o = *reinterpret_cast<const uint4*>(orow + col);
        const uint4 g = *srow(smem + L::kOffB + (col >> 6) * kSlab128, m, (col & 63) >> 3);
        const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
        const float2 o0 = unpack_bf16(o.x), o1 = unpack_bf16(o.y), o2 = unpack_bf16(o.z), o3 = unpack_bf16(o.w);
        acc += g0.x * o0.x + g0.y * o0.y + g1.x * o1.x + g1.y * o1.y + g2.x * o2.x + g2.y * o2.y + g3.x * o3.x + g3.y * o3.y;
      }
      dl = acc;
      if (live && half == 0) delta[(long long)pair * S + r_idx] = acc;
    } else {
      for (int i = threadIdx.x; i < NC * 64; i += 256) {     // statistics of the streamed queries, chunk by chunk
        const int cp = geo.box_pair(i >> 6);
        const int qraw = geo.packed ? (i & 63) : i;
        const int qi = qraw < S ? qraw : S - 1;
        lse_s[i] = lse[(long long)cp * S + qi] * kLog2eG;
        dl_s[i] = delta[(long long)cp * S + qi];
      }
      named_bar_sync(1, 256);
    }
#pragma unroll 1
    for (int c = 0; c < NC; ++c) {
      const int buf = c % NB;
      mbar_wait_bounded(&sp_ready[buf], (c / NB) & 1);
      tc_fence_after();
      if (c >= NB) mbar_wait_bounded(&ring_free[buf], ((c / NB) - 1) & 1);
      uint8_t* ring0 = smem + L::kOffRing + buf * kRingTile;
      uint8_t* ring1 = smem + L::kOffRing + (NB + buf) * kRingTile;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {           // 16 columns at a time: ~100 registers, two CTAs per SM when NSLAB == 1
        uint32_t rs[16], rp[16];
        tmem_ld16(t_row + buf * 128 + half * 32 + sub * 16, rs);
        tmem_ld16(t_row + buf * 128 + 64 + half * 32 + sub * 16, rp);
        tmem_ld_wait();
        float pv[16], ds[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int cj = half * 32 + sub * 16 + i;         // column inside the chunk
          const int col = c * 64 + cj;                     // key (MODE 0) / query (MODE 1) slot of this column
          const float l2 = MODE == 0 ? ls2 : lse_s[col];
          const float dd = MODE == 0 ? dl : dl_s[col];
          const float p = geo.col_live(m, c, cj) ? exp2f(fmaf(__uint_as_float(rs[i]), sc, -l2)) : 0.f;
          pv[i] = p;
          ds[i] = p * (__uint_as_float(rp[i]) - dd) * scale;
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 o;
          o.x = pack_bf16(ds[g * 8 + 0], ds[g * 8 + 1]); o.y = pack_bf16(ds[g * 8 + 2], ds[g * 8 + 3]);
          o.z = pack_bf16(ds[g * 8 + 4], ds[g * 8 + 5]); o.w = pack_bf16(ds[g * 8 + 6], ds[g * 8 + 7]);
          *srow(MODE == 0 ? ring0 : ring1, m, half * 4 + sub * 2 + g) = o;
          if (MODE == 1) {
            uint4 p4;
            p4.x = pack_bf16(pv[g * 8 + 0], pv[g * 8 + 1]); p4.y = pack_bf16(pv[g * 8 + 2], pv[g * 8 + 3]);
            p4.z = pack_bf16(pv[g * 8 + 4], pv[g * 8 + 5]); p4.w = pack_bf16(pv[g * 8 + 6], pv[g * 8 + 7]);
            *srow(ring0, m, half * 4 + sub * 2 + g) = p4;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&sp_free[buf]);                     // both halves of the score buffer are in registers / written out
      fence_proxy_async_smem();
      mbar_arrive(&ring_ready[buf]);
    }
    // ---------------- epilogue: pixel-norm adjoint of the result rows ----------------
    mbar_wait_bounded(bar_o, 0);
    tc_fence_after();
    const uint32_t t_acc_row = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    const long long grow = (long long)(row0 + r_cl) * 3 * C;
    const long long nrow = (long long)(row0 + r_cl) * 3 * heads;
    if (MODE == 0) {
      if (half == 0)
        norm_adjoint_row<NSLAB>(t_acc_row, smem + L::kOffA, m, hd, norms[nrow + head], live, g_qkv + grow + head * hd);
    } else {
      if (half == 0)    // dK (accumulator 0), y = Kn
        norm_adjoint_row<NSLAB>(t_acc_row, smem + L::kOffA, m, hd, norms[nrow + heads + head], live, g_qkv + grow + C + head * hd);
      else              // dV (accumulator 1), y = Vn
        norm_adjoint_row<NSLAB>(t_acc_row + 64 * NSLAB, smem + L::kOffB, m, hd, norms[nrow + 2 * heads + head], live,
                                g_qkv + grow + 2 * C + head * hd);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::kTmemCols);
  }
}

int make_maps(CUtensorMap* t_qkv, CUtensorMap* t_do, const __nv_bfloat16* qn, const __nv_bfloat16* g_y, int B, int S, int C) {
  uint32_t box[2] = {64, 64};
  {
    uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)B * S};
    uint64_t strides[1] = {(uint64_t)3 * C * 2};
    if (encode_tmap(t_qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qn, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  }
  if (t_do != nullptr) {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)B * S};
    uint64_t strides[1] = {(uint64_t)C * 2};
    if (encode_tmap(t_do, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g_y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0) return -1;
  }
  return 0;
}

template <int NSLAB>
int launch_fwd_gen(const __nv_bfloat16* qn, __nv_bfloat16* y, float* lse, int B, int S, int heads, int hd, cudaStream_t stream) {
  using L = FwdLayout<NSLAB>;
  CUtensorMap t_qkv;
  if (make_maps(&t_qkv, nullptr, qn, nullptr, B, S, heads * hd) != 0) return -1;
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured))
    TEDM_CUDA(cudaFuncSetAttribute(attn_fwd_gen_kernel<NSLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmem));
  const int n_pairs = B * heads;
  const int grid = S <= 64 ? (n_pairs + 1) / 2 : n_pairs * ((S + 127) / 128);
  launch_pdl(attn_fwd_gen_kernel<NSLAB>, grid, kFwdThreads, L::kSmem, stream, t_qkv, y, lse, S, hd, heads, n_pairs,
             1.0f / sqrtf((float)hd));
  TEDM_LAUNCH_CHECK();
  return 0;
}

template <int NSLAB>
int launch_bwd_gen(const __nv_bfloat16* qn, const float* norms, const __nv_bfloat16* y, const __nv_bfloat16* g_y, const float* lse,
                   float* delta, __nv_bfloat16* g_qkv, int B, int S, int heads, int hd, cudaStream_t stream) {
  CUtensorMap t_qkv, t_do;
  if (make_maps(&t_qkv, &t_do, qn, g_y, B, S, heads * hd) != 0) return -1;
  static unsigned long long configured = 0;
  if (first_use_on_device(&configured)) {
    TEDM_CUDA(cudaFuncSetAttribute(attn_bwd_gen_kernel<NSLAB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdLayout<NSLAB, 0>::kSmem));
    TEDM_CUDA(cudaFuncSetAttribute(attn_bwd_gen_kernel<NSLAB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdLayout<NSLAB, 1>::kSmem));
  }
  const int n_pairs = B * heads;
  const int grid = S <= 64 ? (n_pairs + 1) / 2 : n_pairs * ((S + 127) / 128);
  const float scale = 1.0f / sqrtf((float)hd);
  launch_pdl(attn_bwd_gen_kernel<NSLAB, 0>, grid, kBwdThreads, BwdLayout<NSLAB, 0>::kSmem, stream, t_qkv, t_do, y, norms, lse, delta,
             g_qkv, S, hd, heads, n_pairs, scale);
  TEDM_LAUNCH_CHECK();
  launch_pdl(attn_bwd_gen_kernel<NSLAB, 1>, grid, kBwdThreads, BwdLayout<NSLAB, 1>::kSmem, stream, t_qkv, t_do, y, norms, lse, delta,
             g_qkv, S, hd, heads, n_pairs, scale);
  TEDM_LAUNCH_CHECK();
  return 0;
}

static_assert(BwdLayout<3, 0>::kSmem <= 232448 && BwdLayout<3, 1>::kSmem <= 232448, "shared memory budget");
static_assert(FwdLayout<3>::kSmem <= 232448, "shared memory budget");

int check_gen(int S, int hd, const void* a, const void* b) {
  TEDM_CHECK(S >= 1 && S <= 256, "attention (normalised): S = %d not supported (1..256)", S);
  TEDM_CHECK(hd >= 16 && hd <= 192 && hd % 16 == 0, "attention (normalised): head_dim %d not supported (multiples of 16 up to 192)", hd);
  TEDM_CHECK((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0,
             "attention: pointers must be 16-byte aligned");
  return 0;
}

}  // namespace

int qkv_normalize(const __nv_bfloat16* qkv, __nv_bfloat16* qn, float* norms, long long rows, int heads, int hd, cudaStream_t stream) {
  TEDM_CHECK(hd % 8 == 0 && hd <= 256, "qkv_normalize: head_dim %d not supported", hd);
  const long long groups = rows * 3 * heads;
  if (groups <= 0) return 0;
  const int chunks = hd / 8;
  if (chunks <= 8) launch_pdl(qkv_norm_kernel<8>, (unsigned)((groups + 31) / 32), 256, 0, stream, qkv, qn, norms, groups, hd);
  else if (chunks <= 16) launch_pdl(qkv_norm_kernel<16>, (unsigned)((groups + 15) / 16), 256, 0, stream, qkv, qn, norms, groups, hd);
  else launch_pdl(qkv_norm_kernel<32>, (unsigned)((groups + 7) / 8), 256, 0, stream, qkv, qn, norms, groups, hd);
  TEDM_LAUNCH_CHECK();
  return 0;
}

int attention_forward_normalized(const __nv_bfloat16* qn, __nv_bfloat16* y, float* lse, int B, int S, int heads, int hd,
                                 cudaStream_t stream) {
  if (check_gen(S, hd, qn, y) != 0) return -1;
  if (B <= 0) return 0;
  const int nslab = (hd + 63) / 64;
  if (nslab == 1) return launch_fwd_gen<1>(qn, y, lse, B, S, heads, hd, stream);
  if (nslab == 2) return launch_fwd_gen<2>(qn, y, lse, B, S, heads, hd, stream);
  return launch_fwd_gen<3>(qn, y, lse, B, S, heads, hd, stream);
}

int attention_backward_normalized(const __nv_bfloat16* qn, const float* norms, const __nv_bfloat16* y, const __nv_bfloat16* g_y,
                                  const float* lse, float* delta, __nv_bfloat16* g_qkv, int B, int S, int heads, int hd,
                                  cudaStream_t stream) {
  if (check_gen(S, hd, qn, g_qkv) != 0) return -1;
  if (B <= 0) return 0;
  const int nslab = (hd + 63) / 64;
  if (nslab == 1) return launch_bwd_gen<1>(qn, norms, y, g_y, lse, delta, g_qkv, B, S, heads, hd, stream);
  if (nslab == 2) return launch_bwd_gen<2>(qn, norms, y, g_y, lse, delta, g_qkv, B, S, heads, hd, stream);
  return launch_bwd_gen<3>(qn, norms, y, g_y, lse, delta, g_qkv, B, S, heads, hd, stream);
}

}  // namespace tedm
