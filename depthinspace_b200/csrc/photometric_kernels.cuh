// Kernels built on window.cuh:
//   photometric_fwd_kernel / photometric_bwd_kernel  -- ext_cuda.photometric_loss_{forward,backward}
//                                                       (reference: model/ext_functions.py:115-140, 156-183)
//   pattern_loss_kernel -- RectifiedPatternSimilarityLoss.tforward fused in one pass
//                          (reference: model/networks.py:354-377): pattern warp -> window loss ->
//                          sigma-weighted partial sums (+ un-normalised d/d disp)
// One translation unit per window radius (see photometric_inst.cu).
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
  int N, H, W;
  float eps, inv_k2, inv_w, inv_h;
  int vec_ok;
};

template <int R>
constexpr size_t photo_smem_bytes(bool with_w) {
  return sizeof(float) * (TileGeom<R>::SIZE * (with_w ? 3 : 2) + NFIX + 2 * (NTHREADS / 32));
}
template <int R>
constexpr size_t pattern_smem_bytes() {
  return sizeof(float) * (TileGeom<R>::SIZE * 3 + TH * TW + NFIX + 2 * (NTHREADS / 32));
}

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

// ---------------------------------------------------------------------------------------
template <int TYPE, int R>
__global__ void __launch_bounds__(NTHREADS, 2) photometric_fwd_kernel(PhotoArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float* se = smem;
  float* st = smem + G::SIZE;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;
  const size_t hw = (size_t)a.H * a.W;
  float total[2][4] = {};
  for (int c = 0; c < a.C; ++c) {
    const float* e = a.es + ((size_t)n * a.C + c) * hw;
    const float* t = a.ta + ((size_t)n * a.C + c) * hw;
    if (c) __syncthreads();
    for (int idx = tid; idx < G::ROWS * G::COLS; idx += NTHREADS) {
      const int j = idx / G::COLS, i = idx - j * G::COLS;
      const size_t g = (size_t)clampi(y0 - R + j, 0, a.H - 1) * a.W + clampi(x0 - R + i, 0, a.W - 1);
      se[j * G::PITCH + i] = __ldg(e + g);
      st[j * G::PITCH + i] = __ldg(t + g);
    }
    __syncthreads();
    float acc[2][4], gacc[2][4], ec[2][4], tc[2][4], wc[2][4];
    window_patch<TYPE, R, true, false>(se, st, nullptr, tx, ty, a.eps, acc, gacc, ec, tc, wc);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) total[r][i] += acc[r][i];
  }
  const float s = fwd_scale<TYPE>() * a.inv_k2;
  float* out = a.out + (size_t)n * hw;
#pragma unroll
  for (int r = 0; r < 2; ++r)
    store_quad(out, y0 + 2 * ty + r, x0 + 4 * tx, a.H, a.W, a.vec_ok, total[r][0] * s, total[r][1] * s,
               total[r][2] * s, total[r][3] * s);
}

// ---------------------------------------------------------------------------------------
template <int TYPE, int R>
__global__ void __launch_bounds__(NTHREADS, 2) photometric_bwd_kernel(PhotoArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float* se = smem;
  float* st = smem + G::SIZE;
  float* sw = smem + 2 * G::SIZE;
  float* fix = smem + 3 * G::SIZE;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int n = blockIdx.z / a.C;
  const size_t hw = (size_t)a.H * a.W;
  const float* e = a.es + (size_t)blockIdx.z * hw;
  const float* t = a.ta + (size_t)blockIdx.z * hw;
  const float* go = a.grad_out + (size_t)n * hw;
  for (int idx = tid; idx < G::ROWS * G::COLS; idx += NTHREADS) {
    const int j = idx / G::COLS, i = idx - j * G::COLS;
    const int gy = y0 - R + j, gx = x0 - R + i;
    const bool inside = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
    const size_t g = (size_t)clampi(gy, 0, a.H - 1) * a.W + clampi(gx, 0, a.W - 1);
    se[j * G::PITCH + i] = __ldg(e + g);
    st[j * G::PITCH + i] = __ldg(t + g);
    sw[j * G::PITCH + i] = inside ? __ldg(go + g) : 0.0f;
  }
  __syncthreads();
  float acc[2][4], gacc[2][4], ec[2][4], tc[2][4], wc[2][4];
  window_patch<TYPE, R, false, true>(se, st, sw, tx, ty, a.eps, acc, gacc, ec, tc, wc);
  const bool edge_tile = (x0 == 0) || (y0 == 0) || (x0 + TW >= a.W) || (y0 + TH >= a.H);
  if (edge_tile) {  // block-uniform
    border_fixup<TYPE, R>(se, st, sw, fix, x0, y0, a.H, a.W, a.eps, tid);
    __syncthreads();
  }
  float* ge = a.grad_es + (size_t)blockIdx.z * hw;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float v[4];
    const int gy = y0 + 2 * ty + r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gx = x0 + 4 * tx + i;
      float g = gacc[r][i];
      if (edge_tile && gy < a.H && gx < a.W) {
        const int slot = border_slot(2 * ty + r, 4 * tx + i, gy, gx, a.H, a.W);
        if (slot >= 0) g = fix[slot];
      }
      v[i] = finish_grad<TYPE>(g, ec[r][i], tc[r][i], a.eps) * a.inv_k2;
    }
    store_quad(ge, gy, x0 + 4 * tx, a.H, a.W, a.vec_ok, v[0], v[1], v[2], v[3]);
  }
}

// ---------------------------------------------------------------------------------------
template <int TYPE, int R, bool GRAD>
__global__ void __launch_bounds__(NTHREADS, 2) pattern_loss_kernel(PatternLossArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float* se = smem;                      // warped pattern, replicate-clamped halo
  float* st = smem + G::SIZE;            // LCN image,      replicate-clamped halo
  float* sw = smem + 2 * G::SIZE;        // sigma (or 1),   zero outside the image
  float* sdd = smem + 3 * G::SIZE;       // d proj / d disp of the tile's own pixels
  float* fix = sdd + TH * TW;
  float* red = fix + NFIX;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;
  const size_t hw = (size_t)a.H * a.W;
  const float* disp = a.disp + (size_t)n * hw;
  const float* im = a.im + (size_t)n * hw;
  const float* sd = a.std_in ? a.std_in + (size_t)n * hw : nullptr;

#pragma unroll 2
  for (int idx = tid; idx < G::ROWS * G::COLS; idx += NTHREADS) {
    const int j = idx / G::COLS, i = idx - j * G::COLS;
    const int gy = y0 - R + j, gx = x0 - R + i;
    const bool inside = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
    const int cy = clampi(gy, 0, a.H - 1), cx = clampi(gx, 0, a.W - 1);
    const size_t g = (size_t)cy * a.W + cx;
    const bool own = GRAD && j >= R && j < R + TH && i >= R && i < R + TW;
    float dd = 0.0f;
    const WarpRow row = warp_row_setup(cy, a.H, a.W, a.inv_h);
    const float e = warp_col_sample(a.pattern, row, __ldg(disp + g), cx, a.W, a.inv_w, own ? &dd : nullptr);
    se[j * G::PITCH + i] = e;
    st[j * G::PITCH + i] = __ldg(im + g);
    sw[j * G::PITCH + i] = inside ? (sd ? __ldg(sd + g) : 1.0f) : 0.0f;
    if (own) sdd[(j - R) * TW + (i - R)] = dd;
  }
  __syncthreads();

  float acc[2][4], gacc[2][4], ec[2][4], tc[2][4], wc[2][4];
  window_patch<TYPE, R, true, GRAD>(se, st, sw, tx, ty, a.eps, acc, gacc, ec, tc, wc);
  if (!GRAD) {  // window_patch only loads the centre weights for the backward half
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) wc[r][i] = sw[(2 * ty + r + R) * G::PITCH + 4 * tx + i + R];
  }
  const bool edge_tile = (x0 == 0) || (y0 == 0) || (x0 + TW >= a.W) || (y0 + TH >= a.H);
  if (GRAD && edge_tile) {
    border_fixup<TYPE, R>(se, st, sw, fix, x0, y0, a.H, a.W, a.eps, tid);
    __syncthreads();
  }

  const float fs = fwd_scale<TYPE>() * a.inv_k2;
  float num = 0.0f, den = 0.0f;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int gy = y0 + 2 * ty + r;
    float d[4], gv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gx = x0 + 4 * tx + i;
      const bool valid = gy < a.H && gx < a.W;
      d[i] = acc[r][i] * fs;
      if (valid) { num = fmaf(wc[r][i], d[i], num); den += wc[r][i]; }
      if (GRAD) {
        float g = gacc[r][i];
        if (edge_tile && valid) {
          const int slot = border_slot(2 * ty + r, 4 * tx + i, gy, gx, a.H, a.W);
          if (slot >= 0) g = fix[slot];
        }
        gv[i] = finish_grad<TYPE>(g, ec[r][i], tc[r][i], a.eps) * a.inv_k2 * sdd[(2 * ty + r) * TW + 4 * tx + i];
      }
    }
    if (a.diff) store_quad(a.diff + (size_t)n * hw, gy, x0 + 4 * tx, a.H, a.W, a.vec_ok, d[0], d[1], d[2], d[3]);
    if (a.proj) store_quad(a.proj + (size_t)n * hw, gy, x0 + 4 * tx, a.H, a.W, a.vec_ok, ec[r][0], ec[r][1], ec[r][2], ec[r][3]);
    if (GRAD) store_quad(a.grad_num + (size_t)n * hw, gy, x0 + 4 * tx, a.H, a.W, a.vec_ok, gv[0], gv[1], gv[2], gv[3]);
  }
  block_sum2<NTHREADS>(num, den, red);
  if (tid == 0) {
    const size_t b = ((size_t)n * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    a.partials[2 * b] = num;
    a.partials[2 * b + 1] = den;
  }
}

// Host launchers, explicitly instantiated once per radius by photometric_inst.cu (-DDIS_R=<R>).
template <int R> int launch_photometric(const PhotoArgs& a, int type, bool backward, cudaStream_t s);
template <int R> int launch_pattern_loss(const PatternLossArgs& a, int type, cudaStream_t s);

}  // namespace dis
