// HBM-bandwidth-bound activation kernels of the EDM2 blocks (NHWC bf16, 16-byte vector accesses).
//
//   block_prep_{fwd,bwd}   resample (avg-pool / nearest-exact), skip concat with the ScaleLong gain,
//                          pixel_norm and mp_silu in ONE pass           networks.py:9-14, :67-88, :246-252, :306-316
//   modsilu_bwd            backward of dropout(mp_silu(r * m[b,c])) incl. the per-(b,c) modulation
//                          gradient reduction                           networks.py:255-261, :319-325
//   channel_dot            out[b,c] = scale * sum_hw A[b,hw,c0+c] * B[b,hw,c] (B optional): ScaleLong mean
//                          (networks.py:115) and its gain gradient
// One warp owns one pixel (all channels): reductions over C are warp shuffles, every global access is a
// 16-byte vector and consecutive lanes touch consecutive 16-byte chunks (fully coalesced).
#include "common.cuh"
#include "kernels.h"

namespace tedm {

namespace {

constexpr int kMaxVec = 6;  // 6 * 32 lanes * 8 channels = 1536 channels max (kernels are templated on the actual count)
constexpr float kEps = 1e-4f;

struct Vec8 {
  float v[8];
};

__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  Vec8 r;
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
  return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  u.x = pack_bf16(r.v[0], r.v[1]); u.y = pack_bf16(r.v[2], r.v[3]);
  u.z = pack_bf16(r.v[4], r.v[5]); u.w = pack_bf16(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ Vec8 load8f(const float* p) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  Vec8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// ------------------------------------------------------------------------------------------------
// block_prep forward
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
block_prep_fwd_kernel(const PrepArgs a) {
  const int H = a.resample == 1 ? a.Hin / 2 : (a.resample == 2 ? a.Hin * 2 : a.Hin);
  const int W = a.resample == 1 ? a.Win / 2 : (a.resample == 2 ? a.Win * 2 : a.Win);
  const int C = a.C1 + a.C2;
  const int nvec = C / 8;
  const long long npix = (long long)a.B * H * W;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long pix = warp0; pix < npix; pix += nwarps) {
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    Vec8 val[NV];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int v = lane + it * 32;
      if (v < nvec) {
        const int c0 = v * 8;
        const bool from_skip = c0 >= a.C1;
        const __nv_bfloat16* src = from_skip ? a.skip : a.in;
        const int cs = from_skip ? a.C2 : a.C1;
        const int cc = from_skip ? c0 - a.C1 : c0;
        Vec8 x;
        if (a.resample == 1) {
          const long long base = (((long long)b * a.Hin + 2 * h) * a.Win + 2 * w) * cs + cc;
          Vec8 p0 = load8(src + base), p1 = load8(src + base + cs);
          Vec8 p2 = load8(src + base + (long long)a.Win * cs), p3 = load8(src + base + (long long)a.Win * cs + cs);
#pragma unroll
          for (int i = 0; i < 8; ++i) x.v[i] = 0.25f * (p0.v[i] + p1.v[i] + p2.v[i] + p3.v[i]);
        } else {
          const int hs = a.resample == 2 ? h >> 1 : h, ws = a.resample == 2 ? w >> 1 : w;
          x = load8(src + (((long long)b * a.Hin + hs) * a.Win + ws) * cs + cc);
        }
        if (from_skip && a.gain != nullptr) {
          Vec8 g = load8f(a.gain + (long long)b * a.C2 + cc);
#pragma unroll
          for (int i = 0; i < 8; ++i) x.v[i] *= g.v[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += x.v[i] * x.v[i];
        val[it] = x;
      }
    }
    float inv_n = 1.0f;
    if (a.pixelnorm) {
      ss = warp_sum(ss);
      const float n = kEps + sqrtf(ss / (float)C);
      inv_n = 1.0f / n;
      if (a.nrm_out != nullptr && lane == 0) a.nrm_out[pix] = n;
    }
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int v = lane + it * 32;
      if (v < nvec) {
        Vec8 x = val[it];
#pragma unroll
        for (int i = 0; i < 8; ++i) x.v[i] = bf16_round(x.v[i] * inv_n);
        const long long o = pix * C + v * 8;
        if (a.x_out != nullptr) store8(a.x_out + o, x);
        if (a.a_out != nullptr) {
          Vec8 s;
#pragma unroll
          for (int i = 0; i < 8; ++i) s.v[i] = mp_silu_f(x.v[i]);
          store8(a.a_out + o, s);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// block_prep backward. One warp per pixel of the PRE-resample grid when upsampling was applied
// (it sums its 2x2 children), otherwise per post-resample pixel.
//   g_x_total = beta * g_res + g_a * mp_silu'(x)
//   pixel_norm: g_u = g/n - x * sum_c(g*x) / ((n - eps) * C)
//   resample adjoint, concat split: g_in = g_u[:C1] ; g_skip = g_u[C1:] * gain (+ d_mean/(Hin*Win))
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void prep_bwd_pixel_grad(const PrepBwdArgs& a, long long pix, int C, int lane, int nvec,
                                                    Vec8 (&g)[NV]) {
  float dot = 0.f;
  Vec8 xs[NV];
#pragma unroll
  for (int it = 0; it < NV; ++it) {
    const int v = lane + it * 32;
    if (v < nvec) {
      const long long o = pix * C + v * 8;
      Vec8 acc;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
      if (a.g_res != nullptr) {
        Vec8 r = load8(a.g_res + o);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] = a.beta * r.v[i];
      }
      if (a.g_a != nullptr || a.pixelnorm) xs[it] = load8(a.x + o);
      if (a.g_a != nullptr) {
        Vec8 ga = load8(a.g_a + o);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] += ga.v[i] * mp_silu_grad_f(xs[it].v[i]);
      }
      if (a.pixelnorm) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dot += acc.v[i] * xs[it].v[i];
      }
      g[it] = acc;
    }
  }
  if (a.pixelnorm) {
    dot = warp_sum(dot);
    const float n = a.nrm[pix];
    const float inv_n = 1.0f / n;
    const float k = dot / (fmaxf(n - kEps, 1e-20f) * (float)C);
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int v = lane + it * 32;
      if (v < nvec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[it].v[i] = g[it].v[i] * inv_n - xs[it].v[i] * k;
      }
    }
  }
}

__device__ __forceinline__ void prep_bwd_store(const PrepBwdArgs& a, int b, long long src_pix, int v, const Vec8& gin,
                                               float scale) {
  // src_pix: pixel index on the pre-resample grid (B,Hin,Win)
  const int c0 = v * 8;
  Vec8 o;
  if (c0 < a.C1) {
    __nv_bfloat16* dst = a.g_in + src_pix * a.C1 + c0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = gin.v[i] * scale;
    if (a.accumulate_in) {
      Vec8 old = load8(dst);
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] += old.v[i];
    }
    store8(dst, o);
  } else if (a.g_skip != nullptr) {
    const int cc = c0 - a.C1;
    __nv_bfloat16* dst = a.g_skip + src_pix * a.C2 + cc;
    Vec8 gn;
    if (a.gain != nullptr) gn = load8f(a.gain + (long long)b * a.C2 + cc);
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = gin.v[i] * scale * (a.gain != nullptr ? gn.v[i] : 1.0f);
    if (a.d_mean != nullptr) {
      Vec8 dm = load8f(a.d_mean + (long long)b * a.C2 + cc);
      const float inv_hw = 1.0f / (float)(a.Hin * a.Win);
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] += dm.v[i] * inv_hw;
    }
    if (a.accumulate_skip) {
      Vec8 old = load8(dst);
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] += old.v[i];
    }
    store8(dst, o);
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
block_prep_bwd_kernel(const PrepBwdArgs a) {
  const int H = a.resample == 1 ? a.Hin / 2 : (a.resample == 2 ? a.Hin * 2 : a.Hin);
  const int W = a.resample == 1 ? a.Win / 2 : (a.resample == 2 ? a.Win * 2 : a.Win);
  const int C = a.C1 + a.C2;
  const int nvec = C / 8;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  if (a.resample == 2) {
    // iterate over low-res (source) pixels; sum the four high-res children
    const long long npix = (long long)a.B * a.Hin * a.Win;
    for (long long sp = warp0; sp < npix; sp += nwarps) {
      const int ws = (int)(sp % a.Win);
      const int hs = (int)((sp / a.Win) % a.Hin);
      const int b = (int)(sp / ((long long)a.Win * a.Hin));
      Vec8 tot[NV];
#pragma unroll
      for (int it = 0; it < NV; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) tot[it].v[i] = 0.f;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          const long long pix = ((long long)b * H + 2 * hs + dy) * W + 2 * ws + dx;
          Vec8 g[NV];
          prep_bwd_pixel_grad<NV>(a, pix, C, lane, nvec, g);
#pragma unroll
          for (int it = 0; it < NV; ++it)
            if (lane + it * 32 < nvec)
#pragma unroll
              for (int i = 0; i < 8; ++i) tot[it].v[i] += g[it].v[i];
        }
#pragma unroll
      for (int it = 0; it < NV; ++it) {
        const int v = lane + it * 32;
        if (v < nvec) prep_bwd_store(a, b, sp, v, tot[it], 1.0f);
      }
    }
  } else {
    const long long npix = (long long)a.B * H * W;
    for (long long pix = warp0; pix < npix; pix += nwarps) {
      const int w = (int)(pix % W);
      const int h = (int)((pix / W) % H);
      const int b = (int)(pix / ((long long)W * H));
      Vec8 g[NV];
      prep_bwd_pixel_grad<NV>(a, pix, C, lane, nvec, g);
#pragma unroll
      for (int it = 0; it < NV; ++it) {
        const int v = lane + it * 32;
        if (v < nvec) {
          if (a.resample == 1) {
            for (int dy = 0; dy < 2; ++dy)
              for (int dx = 0; dx < 2; ++dx)
                prep_bwd_store(a, b, ((long long)b * a.Hin + 2 * h + dy) * a.Win + 2 * w + dx, v, g[it], 0.25f);
          } else {
            prep_bwd_store(a, b, pix, v, g[it], 1.0f);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// modsilu backward:  h = drop(mp_silu(r * m));  given g_h:
//   g_z = g_h * keep/(1-p) * mp_silu'(r*m);  g_r = g_z * m;  dm[b,c] += sum_hw g_z * r
// Block = 256 threads = 8 pixel lanes x (C/8 <= 32.. vectors); loops over a chunk of pixels of ONE image,
// accumulates dm in registers, then one shared-memory reduction and one atomicAdd per (b, c).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
modsilu_bwd_kernel(const ModSiluBwdArgs a) {
  extern __shared__ float sred[];  // [rows][C]
  const int nvec = a.C / 8;
  const int rows = blockDim.x / nvec;  // pixel rows processed in parallel
  const int v = threadIdx.x % nvec;
  const int rr = threadIdx.x / nvec;
  const int b = blockIdx.y;
  const int chunk = (a.HW + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * chunk;
  int p_end = p_begin + chunk;
  if (p_end > a.HW) p_end = a.HW;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (rr < rows) {
    const Vec8 m = load8f(a.mod + (long long)b * a.mod_stride + v * 8);
    const float keep_scale = a.drop_p > 0.f ? 1.0f / (1.0f - a.drop_p) : 1.0f;
    const uint32_t thresh = (uint32_t)(a.drop_p * 65536.0f);
    unsigned long long seed64 = ((unsigned long long)a.seed_hi << 32) | a.seed_lo;
    if (a.seed_ptr != nullptr) seed64 += *a.seed_ptr * 0x9E3779B97F4A7C15ull;
    const uint32_t dseed = dropout_seed((uint32_t)seed64, (uint32_t)(seed64 >> 32));
    for (int p = p_begin + rr; p < p_end; p += rows) {
      const long long o = ((long long)b * a.HW + p) * a.C + v * 8;
      Vec8 g = load8(a.g_h + o);
      const Vec8 r = load8(a.raw + o);
      if (a.drop_p > 0.f) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const uint32_t bits = dropout_bits2((unsigned long long)o + i, dseed);
          g.v[i] = (bits & 0xFFFFu) >= thresh ? g.v[i] * keep_scale : 0.f;
          g.v[i + 1] = (bits >> 16) >= thresh ? g.v[i + 1] * keep_scale : 0.f;
        }
      }
      Vec8 out;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gz = g.v[i] * mp_silu_grad_f(r.v[i] * m.v[i]);
        out.v[i] = gz * m.v[i];
        acc[i] += gz * r.v[i];
      }
      store8(a.g_raw + o, out);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sred[rr * a.C + v * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float t = 0.f;
    for (int r2 = 0; r2 < rows; ++r2) t += sred[r2 * a.C + c];
    atomicAdd(a.d_mod + (long long)b * a.mod_stride + c, t);
  }
}

// ------------------------------------------------------------------------------------------------
// channel_dot: out[b, c] (+)= scale * sum_p A[b,p,a_off + c] * (Bm ? Bm[b,p,c] : 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
channel_dot_kernel(const ChannelDotArgs a) {
  extern __shared__ float sred[];
  const int nvec = a.C / 8;
  const int rows = blockDim.x / nvec;
  const int v = threadIdx.x % nvec;
  const int rr = threadIdx.x / nvec;
  const int b = blockIdx.y;
  const int chunk = (a.HW + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * chunk;
  int p_end = p_begin + chunk;
  if (p_end > a.HW) p_end = a.HW;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (rr < rows) {
    for (int p = p_begin + rr; p < p_end; p += rows) {
      const long long pa = ((long long)b * a.HW + p) * a.CA + a.a_off + v * 8;
      Vec8 x = load8(a.A + pa);
      if (a.Bm != nullptr) {
        Vec8 y = load8(a.Bm + ((long long)b * a.HW + p) * a.C + v * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += x.v[i] * y.v[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += x.v[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sred[rr * a.C + v * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float t = 0.f;
    for (int r2 = 0; r2 < rows; ++r2) t += sred[r2 * a.C + c];
    atomicAdd(a.out + (long long)b * a.C + c, t * a.scale);
  }
}


// ------------------------------------------------------------------------------------------------
// Flat fast paths (no resample, no pixel norm): pure elementwise over 8-channel vectors, 4 independent vectors per
// thread in flight so that enough bytes are outstanding to saturate HBM.
//   concat_silu:  x = [in | skip * gain[b]],  a = mp_silu(x)                         (networks.py:309-316)
//   split_grad:   g_in (+)= g_cat[:, :C1];  g_skip (+)= g_cat[:, C1:] * gain[b] + d_mean[b]/HW
// ------------------------------------------------------------------------------------------------
constexpr int kUnroll = 4;

__global__ void __launch_bounds__(256)
concat_silu_kernel(const PrepArgs a, long long nvec_total, int vec_per_pix, int hw) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int C = a.C1 + a.C2;
  for (long long i0 = base; i0 < nvec_total; i0 += stride * kUnroll) {
    Vec8 x[kUnroll];
    long long idx[kUnroll];
    bool on[kUnroll], sk[kUnroll];
    int cc[kUnroll];
    long long pix[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      idx[u] = i0 + u * stride;
      on[u] = idx[u] < nvec_total;
      if (on[u]) {
        pix[u] = idx[u] / vec_per_pix;
        const int c0 = (int)(idx[u] - pix[u] * vec_per_pix) * 8;
        sk[u] = c0 >= a.C1;
        cc[u] = sk[u] ? c0 - a.C1 : c0;
        x[u] = load8(sk[u] ? a.skip + pix[u] * a.C2 + cc[u] : a.in + pix[u] * a.C1 + cc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (!on[u]) continue;
      if (sk[u] && a.gain != nullptr) {
        const Vec8 g = load8f(a.gain + (pix[u] / hw) * a.C2 + cc[u]);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[u].v[i] = bf16_round(x[u].v[i] * g.v[i]);
      }
      const long long o = pix[u] * C + (sk[u] ? cc[u] + a.C1 : cc[u]);
      if (a.x_out != nullptr) store8(a.x_out + o, x[u]);
      if (a.a_out != nullptr) {
        Vec8 sv;
#pragma unroll
        for (int i = 0; i < 8; ++i) sv.v[i] = mp_silu_f(x[u].v[i]);
        store8(a.a_out + o, sv);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
split_grad_kernel(const PrepBwdArgs a, long long nvec_total, int vec_per_pix, int hw) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int C = a.C1 + a.C2;
  const float inv_hw = 1.0f / (float)hw;
  for (long long i0 = base; i0 < nvec_total; i0 += stride * kUnroll) {
    Vec8 g[kUnroll], old[kUnroll];
    bool on[kUnroll], sk[kUnroll], has_old[kUnroll];
    int cc[kUnroll];
    long long pix[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long idx = i0 + u * stride;
      on[u] = idx < nvec_total;
      has_old[u] = false;
      if (on[u]) {
        pix[u] = idx / vec_per_pix;
        const int c0 = (int)(idx - pix[u] * vec_per_pix) * 8;
        sk[u] = c0 >= a.C1;
        cc[u] = sk[u] ? c0 - a.C1 : c0;
        g[u] = load8(a.g_res + pix[u] * C + c0);
        if (!sk[u] && a.accumulate_in) { old[u] = load8(a.g_in + pix[u] * a.C1 + cc[u]); has_old[u] = true; }
        if (sk[u] && a.g_skip != nullptr && a.accumulate_skip) { old[u] = load8(a.g_skip + pix[u] * a.C2 + cc[u]); has_old[u] = true; }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (!on[u]) continue;
      Vec8 o;
      if (!sk[u]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = g[u].v[i] * a.beta + (has_old[u] ? old[u].v[i] : 0.f);
        store8(a.g_in + pix[u] * a.C1 + cc[u], o);
      } else if (a.g_skip != nullptr) {
        const long long brow = (pix[u] / hw) * a.C2 + cc[u];
        Vec8 gn, dm;
        if (a.gain != nullptr) gn = load8f(a.gain + brow);
        if (a.d_mean != nullptr) dm = load8f(a.d_mean + brow);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = g[u].v[i] * a.beta * (a.gain != nullptr ? gn.v[i] : 1.0f);
          if (a.d_mean != nullptr) v += dm.v[i] * inv_hw;
          o.v[i] = v + (has_old[u] ? old[u].v[i] : 0.f);
        }
        store8(a.g_skip + pix[u] * a.C2 + cc[u], o);
      }
    }
  }
}

int flat_grid(long long nvec) {
  long long blocks = (nvec + 256LL * kUnroll - 1) / (256LL * kUnroll);
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

int grid_for_warps(long long nwarps_needed, int warps_per_block) {
  long long blocks = (nwarps_needed + warps_per_block - 1) / warps_per_block;
  long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

int block_prep_forward(const PrepArgs& a, cudaStream_t stream) {
  const int C = a.C1 + a.C2;
  TEDM_CHECK(a.C1 % 8 == 0 && a.C2 % 8 == 0 && C / 8 <= kMaxVec * 32 && C > 0, "block_prep: unsupported channels %d+%d", a.C1, a.C2);
  TEDM_CHECK(a.resample != 1 || (a.Hin % 2 == 0 && a.Win % 2 == 0), "block_prep: downsample needs even H, W");
  TEDM_CHECK((a.C2 == 0) == (a.skip == nullptr), "block_prep: skip pointer / C2 mismatch");
  const int H = a.resample == 1 ? a.Hin / 2 : (a.resample == 2 ? a.Hin * 2 : a.Hin);
  const int W = a.resample == 1 ? a.Win / 2 : (a.resample == 2 ? a.Win * 2 : a.Win);
  const long long npix = (long long)a.B * H * W;
  if (npix == 0) return 0;
  if (a.resample == 0 && !a.pixelnorm) {   // flat elementwise fast path
    const long long nvec = npix * (C / 8);
    concat_silu_kernel<<<flat_grid(nvec), 256, 0, stream>>>(a, nvec, C / 8, H * W);
    TEDM_LAUNCH_CHECK();
    return 0;
  }
  const int nv = (C / 8 + 31) / 32;
  const int grid = grid_for_warps(npix, 8);
  switch (nv) {
    case 1: block_prep_fwd_kernel<1><<<grid, 256, 0, stream>>>(a); break;
    case 2: block_prep_fwd_kernel<2><<<grid, 256, 0, stream>>>(a); break;
    case 3: block_prep_fwd_kernel<3><<<grid, 256, 0, stream>>>(a); break;
    case 4: block_prep_fwd_kernel<4><<<grid, 256, 0, stream>>>(a); break;
    default: block_prep_fwd_kernel<6><<<grid, 256, 0, stream>>>(a); break;
  }
  TEDM_LAUNCH_CHECK();
  return 0;
}

int block_prep_backward(const PrepBwdArgs& a, cudaStream_t stream) {
  const int C = a.C1 + a.C2;
  TEDM_CHECK(a.C1 % 8 == 0 && a.C2 % 8 == 0 && C / 8 <= kMaxVec * 32 && C > 0, "block_prep_bwd: unsupported channels %d+%d", a.C1, a.C2);
  TEDM_CHECK(a.g_res != nullptr || a.g_a != nullptr, "block_prep_bwd: no incoming gradient");
  TEDM_CHECK(!(a.pixelnorm && (a.nrm == nullptr || a.x == nullptr)), "block_prep_bwd: pixelnorm needs x and nrm");
  TEDM_CHECK(!(a.g_a != nullptr && a.x == nullptr), "block_prep_bwd: g_a needs x");
  const long long npix = a.resample == 1 ? (long long)a.B * (a.Hin / 2) * (a.Win / 2) : (long long)a.B * a.Hin * a.Win;
  if (npix == 0) return 0;
  if (a.resample == 0 && !a.pixelnorm && a.g_a == nullptr) {   // flat elementwise fast path (gradient split / copy)
    const long long nvec = npix * (C / 8);
    split_grad_kernel<<<flat_grid(nvec), 256, 0, stream>>>(a, nvec, C / 8, a.Hin * a.Win);
    TEDM_LAUNCH_CHECK();
    return 0;
  }
  const int nv = (C / 8 + 31) / 32;
  const int grid = grid_for_warps(npix, 8);
  switch (nv) {
    case 1: block_prep_bwd_kernel<1><<<grid, 256, 0, stream>>>(a); break;
    case 2: block_prep_bwd_kernel<2><<<grid, 256, 0, stream>>>(a); break;
    case 3: block_prep_bwd_kernel<3><<<grid, 256, 0, stream>>>(a); break;
    case 4: block_prep_bwd_kernel<4><<<grid, 256, 0, stream>>>(a); break;
    default: block_prep_bwd_kernel<6><<<grid, 256, 0, stream>>>(a); break;
  }
  TEDM_LAUNCH_CHECK();
  return 0;
}

static int pick_chunks(int B, int HW, int rows) {
  // enough CTAs to fill the machine, but keep >= 4 pixel iterations per CTA row group
  int want = (4 * num_sms() + B - 1) / B;
  int maxc = (HW + 4 * rows - 1) / (4 * rows);
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  return want;
}

int modsilu_backward(const ModSiluBwdArgs& a, cudaStream_t stream) {
  TEDM_CHECK(a.C % 8 == 0 && a.C / 8 <= 256 && a.C > 0, "modsilu_bwd: unsupported C=%d", a.C);
  const int nvec = a.C / 8;
  const int rows = 256 / nvec;
  TEDM_CHECK(rows >= 1, "modsilu_bwd: C too large");
  dim3 grid(pick_chunks(a.B, a.HW, rows), a.B);
  size_t smem = (size_t)rows * a.C * sizeof(float);
  modsilu_bwd_kernel<<<grid, 256, smem, stream>>>(a);
  TEDM_LAUNCH_CHECK();
  return 0;
}

int channel_dot(const ChannelDotArgs& a, cudaStream_t stream) {
  TEDM_CHECK(a.C % 8 == 0 && a.C / 8 <= 256 && a.C > 0 && a.CA % 8 == 0 && a.a_off % 8 == 0, "channel_dot: unsupported C=%d", a.C);
  const int nvec = a.C / 8;
  const int rows = 256 / nvec;
  dim3 grid(pick_chunks(a.B, a.HW, rows), a.B);
  size_t smem = (size_t)rows * a.C * sizeof(float);
  channel_dot_kernel<<<grid, 256, smem, stream>>>(a);
  TEDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace tedm
