// Shared device helpers for the sm_100a kernels of libdis_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include "../../include/dis_b200.h"

namespace dis {

// ---- launch bookkeeping -------------------------------------------------------------
void set_last_cuda_error(cudaError_t e);
inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return DIS_ERR_CUDA_LAUNCH;
  }
  return DIS_OK;
}
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- exact (non-contracted) fp32 arithmetic -------------------------------------------
// The reference builds its sampling coordinates with one torch kernel per op, so no op is
// ever fused with its neighbour.  These wrappers keep nvcc from contracting mul+add.
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

__device__ __forceinline__ float rsqrt_fast(float x) {  // x >= eps > 0: one MUFU.RSQ, no denormal fix-up
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- packed fp32x2 arithmetic (sm_100+: FADD2 / FMUL2 / FFMA2, one issue slot for two lanes) ----------------
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 bc2(float v) { return pk2(v, v); }  // ptxas folds this into a broadcast operand

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// streaming 128-bit / 32-bit loads that do not pollute L1 (inputs are read once per CTA)
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

// ---- torch-CUDA grid_sample arithmetic (align_corners=True, bilinear) -----------------
// reference: 2 * (pix / (size-1) - 0.5)   (model/networks.py:363-364,
// model/multi_frame_networks.py:95-96); on CUDA, tensor / python-scalar is tensor * (1/scalar).
__device__ __forceinline__ float normalize_coord(float pix, float inv_size_m1) {
  return fmul(2.0f, fsub(fmul(pix, inv_size_m1), 0.5f));
}
// ATen/native/cuda/GridSampler.cuh: ((coord + 1) / 2) * (size - 1)
__device__ __forceinline__ float unnormalize_coord(float g, int size) {
  return fmul(fmul(fadd(g, 1.0f), 0.5f), (float)(size - 1));
}
__device__ __forceinline__ float safe_int_range(float x) {
  if (x > (float)(INT_MAX - 1) || x < (float)INT_MIN || !isfinite(x)) return -100.0f;
  return x;
}

struct Bilinear {
  int x0, y0;                 // nw corner
  float wnw, wne, wsw, wse;   // corner weights, ATen order
  float fx, fy;               // clipped source coordinates
  float gx_mult, gy_mult;     // d(source)/d(normalised) including the border-clip mask
};

template <bool BORDER>
__device__ __forceinline__ float source_index(float g, int size, float& mult) {
  float c = unnormalize_coord(g, size);
  mult = (float)(size - 1) * 0.5f;
  if (BORDER) {  // clip_coordinates_set_grad: borders count as outside for the gradient
    if (c <= 0.0f) { c = 0.0f; mult = 0.0f; }
    else if (c >= (float)(size - 1)) { c = (float)(size - 1); mult = 0.0f; }
  }
  return safe_int_range(c);
}

template <bool BORDER>
__device__ __forceinline__ void bilinear_setup(float gx, float gy, int H, int W, Bilinear& b) {
  const float ix = source_index<BORDER>(gx, W, b.gx_mult);
  const float iy = source_index<BORDER>(gy, H, b.gy_mult);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  b.x0 = (int)fx0;
  b.y0 = (int)fy0;
  const float x1 = (float)(b.x0 + 1), y1 = (float)(b.y0 + 1), x0 = (float)b.x0, y0 = (float)b.y0;
  b.wnw = fmul(fsub(x1, ix), fsub(y1, iy));
  b.wne = fmul(fsub(ix, x0), fsub(y1, iy));
  b.wsw = fmul(fsub(x1, ix), fsub(iy, y0));
  b.wse = fmul(fsub(ix, x0), fsub(iy, y0));
  b.fx = ix;
  b.fy = iy;
}

__device__ __forceinline__ bool in_bounds(int y, int x, int H, int W) {
  return y >= 0 && y < H && x >= 0 && x < W;
}

// corner fetch with the within_bounds_2d() predicate; OOB corners contribute nothing
struct Corners { float nw, ne, sw, se; bool bnw, bne, bsw, bse; };

__device__ __forceinline__ Corners fetch_corners(const float* __restrict__ img, int H, int W, const Bilinear& b) {
  Corners c;
  c.bnw = in_bounds(b.y0, b.x0, H, W);
  c.bne = in_bounds(b.y0, b.x0 + 1, H, W);
  c.bsw = in_bounds(b.y0 + 1, b.x0, H, W);
  c.bse = in_bounds(b.y0 + 1, b.x0 + 1, H, W);
  const float* p = img + (ptrdiff_t)b.y0 * W + b.x0;
  c.nw = c.bnw ? __ldg(p) : 0.0f;
  c.ne = c.bne ? __ldg(p + 1) : 0.0f;
  c.sw = c.bsw ? __ldg(p + W) : 0.0f;
  c.se = c.bse ? __ldg(p + W + 1) : 0.0f;
  return c;
}
// out_acc += v * w in ATen compiles to a chain of FMAs in nw, ne, sw, se order; a skipped
// (out-of-bounds) corner leaves the accumulator untouched, which fma(0, w, acc) reproduces
// for finite weights.
__device__ __forceinline__ float blend(const Corners& c, const Bilinear& b) {
  float acc = 0.0f;
  if (c.bnw) acc = __fmaf_rn(c.nw, b.wnw, acc);
  if (c.bne) acc = __fmaf_rn(c.ne, b.wne, acc);
  if (c.bsw) acc = __fmaf_rn(c.sw, b.wsw, acc);
  if (c.bse) acc = __fmaf_rn(c.se, b.wse, acc);
  return acc;
}
// d out / d source-x and source-y per unit upstream gradient (grid_sampler_2d_backward)
__device__ __forceinline__ float blend_dx(const Corners& c, const Bilinear& b) {
  const float wy0 = (float)(b.y0 + 1) - b.fy, wy1 = b.fy - (float)b.y0;
  float g = 0.0f;
  if (c.bnw) g -= c.nw * wy0;
  if (c.bne) g += c.ne * wy0;
  if (c.bsw) g -= c.sw * wy1;
  if (c.bse) g += c.se * wy1;
  return g;
}
__device__ __forceinline__ float blend_dy(const Corners& c, const Bilinear& b) {
  const float wx0 = (float)(b.x0 + 1) - b.fx, wx1 = b.fx - (float)b.x0;
  float g = 0.0f;
  if (c.bnw) g -= c.nw * wx0;
  if (c.bne) g -= c.ne * wx1;
  if (c.bsw) g += c.sw * wx0;
  if (c.bse) g += c.se * wx1;
  return g;
}

// pattern warp of one pixel: proj value and d proj / d disp (model/networks.py:358-367)
__device__ __forceinline__ float pattern_warp_pixel(const float* __restrict__ pattern, float disp, int h, int w,
                                                    int H, int W, float inv_w, float inv_h, float* dproj,
                                                    int* x0 = nullptr, int* y0 = nullptr) {
  const float gx = normalize_coord(fsub((float)w, disp), inv_w);
  const float gy = normalize_coord((float)h, inv_h);
  Bilinear b;
  bilinear_setup<true>(gx, gy, H, W, b);
  const Corners c = fetch_corners(pattern, H, W, b);
  if (dproj) *dproj = -(((b.gx_mult * blend_dx(c, b)) * 2.0f) * inv_w);
  if (x0) *x0 = b.x0;
  if (y0) *y0 = b.y0;
  return blend(c, b);
}

// ---- pattern warp split into a per-position y half and a per-disparity x half ---------------------
// Same arithmetic as pattern_warp_pixel (bit-identical results); the y half is shared by every
// disparity map sampled at the same pixel, corner fetches are branch-free (clamped address, value
// forced to 0 when ATen's within_bounds_2d() fails: fma(0, w, acc) == acc, i.e. "corner skipped").
struct WarpRow {
  int off0, off1;      // element offsets of pattern rows y0 and y0+1 (clamped so that the address is always valid)
  float wy0, wy1;      // (y0 + 1) - iy, iy - y0
  int by;              // bit 0: row y0 inside the image, bit 1: row y0+1 inside
};

// Branch-free border clip.  Forward semantics of ATen's clip_coordinates (min/max, so NaN clips to 0) and the
// gradient mask of clip_coordinates_set_grad (zero on and outside the border).
__device__ __forceinline__ float border_source_index(float g, int size, float& mult) {
  const float c = unnormalize_coord(g, size);
  const float hi = (float)(size - 1);
  mult = (c > 0.0f && c < hi) ? hi * 0.5f : 0.0f;
  return fminf(fmaxf(c, 0.0f), hi);
}

__device__ __forceinline__ WarpRow warp_row_setup(int h, int H, int W, float inv_h) {
  float unused;
  const float iy = border_source_index(normalize_coord((float)h, inv_h), H, unused);
  const int y0 = (int)floorf(iy);
  WarpRow r;
  r.wy0 = fsub((float)(y0 + 1), iy);
  r.wy1 = fsub(iy, (float)y0);
  r.by = ((unsigned)y0 < (unsigned)H ? 1 : 0) | ((unsigned)(y0 + 1) < (unsigned)H ? 2 : 0);
  r.off0 = clampi(y0, 0, H - 1) * W;
  r.off1 = clampi(y0 + 1, 0, H - 1) * W;
  return r;
}

// value of the warped pattern; *dproj (optional) = d value / d disp
__device__ __forceinline__ float warp_col_sample(const float* __restrict__ pattern, const WarpRow& r, float disp, int w,
                                                 int W, float inv_w, float* dproj) {
  float mult;
  const float ix = border_source_index(normalize_coord(fsub((float)w, disp), inv_w), W, mult);
  const int x0 = (int)floorf(ix);
  const float wx0 = fsub((float)(x0 + 1), ix), wx1 = fsub(ix, (float)x0);
  const bool bx0 = (unsigned)x0 < (unsigned)W, bx1 = (unsigned)(x0 + 1) < (unsigned)W;
  const int c0 = clampi(x0, 0, W - 1), c1 = clampi(x0 + 1, 0, W - 1);
  const float l00 = __ldg(pattern + r.off0 + c0), l01 = __ldg(pattern + r.off0 + c1);
  const float l10 = __ldg(pattern + r.off1 + c0), l11 = __ldg(pattern + r.off1 + c1);
  const bool by0 = r.by & 1, by1 = r.by & 2;
  const float vnw = (by0 && bx0) ? l00 : 0.0f, vne = (by0 && bx1) ? l01 : 0.0f;
  const float vsw = (by1 && bx0) ? l10 : 0.0f, vse = (by1 && bx1) ? l11 : 0.0f;
  float acc = __fmaf_rn(vnw, fmul(wx0, r.wy0), 0.0f);
  acc = __fmaf_rn(vne, fmul(wx1, r.wy0), acc);
  acc = __fmaf_rn(vsw, fmul(wx0, r.wy1), acc);
  acc = __fmaf_rn(vse, fmul(wx1, r.wy1), acc);
  if (dproj) {
    float g = 0.0f;   // grid_sampler_2d_backward: gix
    g -= vnw * r.wy0; g += vne * r.wy0; g -= vsw * r.wy1; g += vse * r.wy1;
    *dproj = -(((mult * g) * 2.0f) * inv_w);
  }
  return acc;
}

// Same value as warp_col_sample() without the per-corner within_bounds_2d() predicates: after the border clip the
// source coordinates lie in [0, W-1] x [0, H-1], so the only corners that can fall outside the image are x0+1 == W
// (ix == W-1, weight wx1 == 0) and y0+1 == H (weight wy1 == 0).  Reading the clamped neighbour instead of skipping
// the corner adds v * 0 for a finite pattern value v: the same sum (the sign of a zero result may differ; the
// values compare equal), and d proj / d disp is masked by mult == 0 there.  Needs a finite pattern.
__device__ __forceinline__ float warp_col_sample_clamped(const float* __restrict__ row0, const float* __restrict__ row1,
                                                         float wy0, float wy1, float disp, int w, int W, float inv_w,
                                                         float* dproj) {
  float mult;
  const float ix = border_source_index(normalize_coord(fsub((float)w, disp), inv_w), W, mult);
  const int x0 = (int)floorf(ix);
  const float wx0 = fsub((float)(x0 + 1), ix), wx1 = fsub(ix, (float)x0);
  const int c1 = min(x0 + 1, W - 1);
  const float vnw = __ldg(row0 + x0), vne = __ldg(row0 + c1);
  const float vsw = __ldg(row1 + x0), vse = __ldg(row1 + c1);
  float acc = __fmaf_rn(vnw, fmul(wx0, wy0), 0.0f);
  acc = __fmaf_rn(vne, fmul(wx1, wy0), acc);
  acc = __fmaf_rn(vsw, fmul(wx0, wy1), acc);
  acc = __fmaf_rn(vse, fmul(wx1, wy1), acc);
  if (dproj) {
    float g = 0.0f;   // grid_sampler_2d_backward: gix
    g -= vnw * wy0; g += vne * wy0; g -= vsw * wy1; g += vse * wy1;
    *dproj = -(((mult * g) * 2.0f) * inv_w);
  }
  return acc;
}

// ---- reductions -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum of two values; result valid in thread 0.  scratch: >= 2 * (threads/32) floats
template <int THREADS>
__device__ __forceinline__ void block_sum2(float& a, float& b, float* scratch) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  if (lane == 0) { scratch[2 * wid] = a; scratch[2 * wid + 1] = b; }
  __syncthreads();
  if (tid == 0) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) { sa += scratch[2 * i]; sb += scratch[2 * i + 1]; }
    a = sa; b = sb;
  }
}

__device__ __forceinline__ float sign0(float v) { return (float)((v > 0.0f) - (v < 0.0f)); }

}  // namespace dis
