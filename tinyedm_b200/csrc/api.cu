// extern "C" entry points of libtinyedm_b200.so (declared in include/tinyedm_b200.h).
#include "../../include/tinyedm_b200.h"

#include "common.cuh"
#include "kernels.h"

using namespace tedm;

extern "C" {

int tedm_version(void) { return 100; }

const char* tedm_last_error(void) { return last_error(); }

int tedm_init(int device) {
  TEDM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TEDM_CUDA(cudaGetDeviceProperties(&prop, device));
  TEDM_CHECK(prop.major == 10, "tinyedm_b200 requires an sm_100-class GPU (found sm_%d%d, %s)", prop.major, prop.minor,
             prop.name);
  return 0;
}

int tedm_conv2d_forward(const void* x, const void* w, void* out, int B, int H, int W, int Cin, int Cout, int ksize,
                        int epilogue, float alpha, void* raw, const void* res, float t, const float* mod,
                        int mod_stride, float drop_p, uint64_t seed, int block_n, tedm_stream_t stream) {
  ConvGemmArgs a{};
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.w = static_cast<const __nv_bfloat16*>(w);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize;
  a.epi = epilogue; a.alpha = alpha;
  a.out2 = static_cast<__nv_bfloat16*>(raw);
  a.res = static_cast<const __nv_bfloat16*>(res);
  a.t = t;
  a.inv_c = 1.0f / sqrtf((1.0f - t) * (1.0f - t) + t * t);
  a.mod = mod; a.mod_stride = mod_stride; a.drop_p = drop_p; a.seed = seed;
  a.block_n_override = block_n;
  return conv_gemm_launch(a, static_cast<cudaStream_t>(stream));
}

int tedm_conv2d_wgrad(const void* g, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, int ksize,
                      float alpha, int accumulate, int splits, tedm_stream_t stream) {
  ConvWgradArgs a{};
  a.g = static_cast<const __nv_bfloat16*>(g);
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.dw = dw;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize;
  a.alpha = alpha; a.accumulate = accumulate; a.splits_override = splits;
  return conv_wgrad_launch(a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
