// Internal (C++) launch interfaces shared between the .cu translation units and api.cu.
// The public C ABI is include/tinyedm_b200.h.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tedm {

enum ConvEpilogue : int {
  EPI_PLAIN = 0,    // out = alpha * acc
  EPI_MODSILU = 1,  // out = dropout(mp_silu(acc * mod[b, c])); optional raw copy of acc in out2
  EPI_MPADD = 2,    // out = ((1 - t) * res + t * acc) * inv_c            (mp_add, networks.py:87-88)
};

struct ConvGemmArgs {
  const __nv_bfloat16* x;  // (B,H,W,Cin) NHWC
  const __nv_bfloat16* w;  // [Cout][ksize*ksize][Cin]
  __nv_bfloat16* out;      // (B,H,W,Cout)
  int B, H, W, Cin, Cout, ksize;
  int epi;
  float alpha;
  __nv_bfloat16* out2;       // MODSILU: raw conv output (may be null)
  const __nv_bfloat16* res;  // MPADD: residual (B,H,W,Cout)
  float t, inv_c;
  const float* mod;          // MODSILU: (B, mod_stride) fp32, column offset already applied
  int mod_stride;
  float drop_p;
  uint64_t seed;
  int block_n_override;      // 0 = heuristic
};

// Device-side parameter block of the implicit-GEMM kernel.
struct ConvGemmParams {
  int B, H, W, Cin, Cout, taps;
  int RH, NB, tiles_h, m_tiles, n_tiles, block_n, k_blocks;
  int epi;
  float alpha;
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
  const __nv_bfloat16* res;
  float t, inv_c;
  const float* mod;
  int mod_stride;
  float drop_p;
  uint32_t seed_lo, seed_hi;
};

int conv_tile_geometry(int H, int W, int* RH, int* NB);
int conv_gemm_launch(const ConvGemmArgs& a, cudaStream_t stream);

struct ConvWgradArgs {
  const __nv_bfloat16* g;  // (B,H,W,Cout) gradient w.r.t. the conv output
  const __nv_bfloat16* x;  // (B,H,W,Cin) conv input
  float* dw;               // [Cout][ksize*ksize][Cin] fp32
  int B, H, W, Cin, Cout, ksize;
  float alpha;
  int accumulate;          // 0: dw = result, 1: dw += result
  int splits_override;     // 0 = heuristic
};
int conv_wgrad_launch(const ConvWgradArgs& a, cudaStream_t stream);

}  // namespace tedm
