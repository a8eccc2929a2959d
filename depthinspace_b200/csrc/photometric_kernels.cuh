// Argument blocks and launcher declarations shared by the k x k window kernels:
//   ext_cuda.photometric_loss_{forward,backward}   (reference: model/ext_functions.py:115-140, 156-183)
//   RectifiedPatternSimilarityLoss.tforward fused  (reference: model/networks.py:354-377): pattern warp ->
//     window loss -> sigma-weighted partial sums (+ un-normalised d/d disp)
// The kernels live in census_kernels.cuh (soft census), box_kernels.cuh (mse / sad) and pattern_multi.cuh
// (all scales in one pass); one translation unit per window radius (see photometric_inst.cu).
#pragma once
#include "window.cuh"

namespace dis {

struct PhotoArgs {
  const float* es; const float* ta; const float* grad_out;
  float* out; float* grad_es;
  int N, C, H, W;
  float eps, inv_k2;
  int vec_ok;  // W % 4 == 0 and all pointers 16-byte aligned
};

struct PatternLossArgs {
  const float* disp; const float* im; const float* std_in; const float* pattern;
  float* proj; float* diff; float* grad_num; float* partials;
  const float* grad_scale;   // optional device scalar: grad_num is multiplied by it (final gradient, no scaling pass)
  int N, H, W;
  float eps, inv_k2, inv_w, inv_h;
  int vec_ok;
};

// store 4 horizontally adjacent values of tile row (gy) starting at gx
__device__ __forceinline__ void store_quad(float* __restrict__ plane, int gy, int gx, int H, int W, int vec_ok,
                                           float v0, float v1, float v2, float v3) {
  if (gy >= H || gx >= W) return;
  float* p = plane + (size_t)gy * W + gx;
  if (vec_ok && gx + 3 < W) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v0, v1, v2, v3));
  } else {
    p[0] = v0;
    if (gx + 1 < W) p[1] = v1;
    if (gx + 2 < W) p[2] = v2;
    if (gx + 3 < W) p[3] = v3;
  }
}

// Host launchers, explicitly instantiated once per radius by photometric_inst.cu (-DDIS_R=<R>).
template <int R> int launch_photometric(const PhotoArgs& a, int type, bool backward, cudaStream_t s);
template <int R> int launch_pattern_loss(const PatternLossArgs& a, int type, cudaStream_t s);

}  // namespace dis
