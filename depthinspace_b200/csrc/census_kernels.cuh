// Soft-census window loss, single scale, with the estimate and the target travelling together as fp32x2.
//
// The estimate plane e and the target plane t are staged INTERLEAVED in shared memory as float2 (e, t), so the two
// halves of every tap  de = e(q) - e(p), dt = t(q) - t(p),  x = d^2 + eps,  g = d * rsqrt(x)  are one FADD2, one
// FFMA2 and one FMUL2 (14 issue slots per tap forward+backward instead of 17 for the scalar form in window.cuh).
// The two products are rounded separately by the packed multiply and subtracted across the halves, so e == t gives
// exactly 0 (the reference's |.| has subgradient 0 there) without any extra work.
// Same tile / thread layout and the same gather-form backward + border-line fix-up as window.cuh.
//   census_fwd_kernel / census_bwd_kernel : ext_cuda.photometric_loss_{forward,backward}, types 2 and 3
//   census_pattern_loss_kernel            : RectifiedPatternSimilarityLoss.tforward fused (model/networks.py:354-377)
#pragma once
#include "photometric_kernels.cuh"

namespace dis {

#ifndef DIS_CENSUS_FILL_UNROLL
#define DIS_CENSUS_FILL_UNROLL 3
#endif
constexpr int kCensusFillUnroll = DIS_CENSUS_FILL_UNROLL;   // pattern-warp staging positions in flight per thread

template <int TYPE, bool FWD, bool BWD>
__device__ __forceinline__ void census_tap2(u64 etc, u64 etq, float wq, u64 eps2, float& acc, float& ga, float& gb) {
  const u64 d2 = sub2(etq, etc);                 // (de, dt)
  const u64 x2 = fma2(d2, d2, eps2);
  float xe, xt;
  upk2(x2, xe, xt);
  const float re = rsqrt_fast(xe), rt = rsqrt_fast(xt);
  float ge, gt;
  upk2(mul2(d2, pk2(re, rt)), ge, gt);           // both products rounded once, like the scalar __fmul_rn pair
  const float diff2 = ge - gt;                   // = 2 (h(de) - h(dt))
  if (FWD) acc = (TYPE == CENSUS_MSE) ? fmaf(diff2, diff2, acc) : acc + fabsf(diff2);
  if (BWD) {
    const float r3 = re * re * re;               // h'(de) = 0.5 eps r^3
    const float u = (TYPE == CENSUS_MSE) ? diff2 * r3 : signed_mag(r3, diff2);
    ga = fmaf(u, wq, ga);
    gb += u;
  }
}

// Main loop over the thread's 4 x 2 patch; set: interleaved (e, t) plane, sw: weight plane (zero outside the image).
template <int TYPE, int R, bool FWD, bool BWD>
__device__ __forceinline__ void census_patch(const float2* __restrict__ set, const float* __restrict__ sw, int tx, int ty,
                                             float eps, float (&acc)[2][4], float (&gacc)[2][4], float (&ec)[2][4],
                                             float (&tc)[2][4], float (&wc)[2][4]) {
  using G = TileGeom<R>;
  constexpr int NW = 4 + 2 * R;           // window values per thread-row
  constexpr int NWE = (NW + 1) & ~1;      // float2 pairs are fetched two at a time (128-bit)
  u64 etc[2][4];
  float gb[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = (2 * ty + r + R) * G::PITCH + 4 * tx + i + R;
      const float2 v = set[o];
      ec[r][i] = v.x;
      tc[r][i] = v.y;
      etc[r][i] = pk2(v.x, v.y);
      wc[r][i] = BWD ? sw[o] : 0.0f;
      acc[r][i] = gacc[r][i] = gb[r][i] = 0.0f;
    }
  const u64 eps2 = bc2(eps);
#pragma unroll 1
  for (int j = 0; j < 2 * R + 2; ++j) {
    u64 er[NWE];
    float wr[4 * G::NV];
    const int base = (2 * ty + j) * G::PITCH + 4 * tx;
#pragma unroll
    for (int v = 0; v < NWE / 2; ++v) {
      const float4 q = *reinterpret_cast<const float4*>(set + base + 2 * v);
      er[2 * v] = pk2(q.x, q.y);
      er[2 * v + 1] = pk2(q.z, q.w);
    }
    if (BWD) {
#pragma unroll
      for (int v = 0; v < G::NV; ++v) {
        const float4 c = *reinterpret_cast<const float4*>(sw + base + 4 * v);
        wr[4 * v] = c.x; wr[4 * v + 1] = c.y; wr[4 * v + 2] = c.z; wr[4 * v + 3] = c.w;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (j - r < 0 || j - r > 2 * R) continue;  // warp-uniform
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int dx = 0; dx <= 2 * R; ++dx)
          census_tap2<TYPE, FWD, BWD>(etc[r][i], er[i + dx], BWD ? wr[i + dx] : 0.0f, eps2, acc[r][i], gacc[r][i], gb[r][i]);
    }
  }
  if (BWD) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) gacc[r][i] = fmaf(wc[r][i], gb[r][i], gacc[r][i]);
  }
}

// exact backward accumulator of one border-line pixel (clamp multiplicities), interleaved plane
template <int TYPE, int R>
__device__ float census_border_gacc(const float2* __restrict__ set, const float* __restrict__ sw, int ly, int lx, int gy,
                                    int gx, int H, int W, float eps) {
  using G = TileGeom<R>;
  const int c = (ly + R) * G::PITCH + lx + R;
  const float2 ct = set[c];
  const float wc = sw[c];
  float acc = 0.0f, ga = 0.0f, gb = 0.0f;
  for (int dy = -R; dy <= R; ++dy) {
    const int py = gy + dy;
    const int my = (py >= 0 && py < H) ? clamp_multiplicity(py, gy, H, R) : 0;
    for (int dx = -R; dx <= R; ++dx) {
      const int px = gx + dx;
      const int mx = (px >= 0 && px < W) ? clamp_multiplicity(px, gx, W, R) : 0;
      const int o = c + dy * G::PITCH + dx;
      const float2 q = set[o];
      tap<TYPE, false, true>(ct.x, ct.y, q.x, q.y, (float)(my * mx) * sw[o], eps, acc, ga, gb);
    }
  }
  return fmaf(wc, gb, ga);
}

template <int TYPE, int R>
__device__ __forceinline__ void census_border_fixup(const float2* __restrict__ set, const float* __restrict__ sw,
                                                    float* __restrict__ fix, int x0, int y0, int H, int W, float eps, int tid) {
  if (tid < NFIX) {
    int ly, lx;
    if (tid < TH) { ly = tid; lx = 0 - x0; }
    else if (tid < 2 * TH) { ly = tid - TH; lx = W - 1 - x0; }
    else if (tid < 2 * TH + TW) { ly = 0 - y0; lx = tid - 2 * TH; }
    else { ly = H - 1 - y0; lx = tid - 2 * TH - TW; }
    const int gy = y0 + ly, gx = x0 + lx;
    if (ly >= 0 && ly < TH && lx >= 0 && lx < TW && gy < H && gx < W)
      fix[tid] = census_border_gacc<TYPE, R>(set, sw, ly, lx, gy, gx, H, W, eps);
  }
}

template <int R>
constexpr size_t census_smem_bytes(bool with_w, bool with_dd) {
  return sizeof(float) * ((size_t)TileGeom<R>::SIZE * (with_w ? 3 : 2) + (with_dd ? TH * TW : 0) + NFIX + 2 * (NTHREADS / 32));
}

// ---------------------------------------------------------------------------------------------------------------
template <int TYPE, int R>
__global__ void __launch_bounds__(NTHREADS, 2) census_fwd_kernel(PhotoArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float2* set = reinterpret_cast<float2*>(smem);
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;
  const size_t hw = (size_t)a.H * a.W;
  float total[2][4] = {};
  for (int c = 0; c < a.C; ++c) {
    const float* e = a.es + ((size_t)n * a.C + c) * hw;
    const float* t = a.ta + ((size_t)n * a.C + c) * hw;
    if (c) __syncthreads();
#pragma unroll 4
    for (int idx = tid; idx < G::ROWS * G::PITCH; idx += NTHREADS) {
      const int j = idx / G::PITCH, i = idx - j * G::PITCH;
      const int g = clampi(y0 - R + j, 0, a.H - 1) * a.W + clampi(x0 - R + i, 0, a.W - 1);
      set[idx] = make_float2(__ldg(e + g), __ldg(t + g));
    }
    __syncthreads();
    float acc[2][4], gacc[2][4], ec[2][4], tc[2][4], wc[2][4];
    census_patch<TYPE, R, true, false>(set, nullptr, tx, ty, a.eps, acc, gacc, ec, tc, wc);
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

template <int TYPE, int R>
__global__ void __launch_bounds__(NTHREADS, 2) census_bwd_kernel(PhotoArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float2* set = reinterpret_cast<float2*>(smem);
  float* sw = smem + 2 * G::SIZE;
  float* fix = smem + 3 * G::SIZE;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int n = blockIdx.z / a.C;
  const size_t hw = (size_t)a.H * a.W;
  const float* e = a.es + (size_t)blockIdx.z * hw;
  const float* t = a.ta + (size_t)blockIdx.z * hw;
  const float* go = a.grad_out + (size_t)n * hw;
#pragma unroll 4
  for (int idx = tid; idx < G::ROWS * G::PITCH; idx += NTHREADS) {
    const int j = idx / G::PITCH, i = idx - j * G::PITCH;
    const int gy = y0 - R + j, gx = x0 - R + i;
    const bool inside = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
    const int g = clampi(gy, 0, a.H - 1) * a.W + clampi(gx, 0, a.W - 1);
    set[idx] = make_float2(__ldg(e + g), __ldg(t + g));
    sw[idx] = inside ? __ldg(go + g) : 0.0f;
  }
  __syncthreads();
  float acc[2][4], gacc[2][4], ec[2][4], tc[2][4], wc[2][4];
  census_patch<TYPE, R, false, true>(set, sw, tx, ty, a.eps, acc, gacc, ec, tc, wc);
  const bool edge_tile = (x0 == 0) || (y0 == 0) || (x0 + TW >= a.W) || (y0 + TH >= a.H);
  if (edge_tile) {  // block-uniform
    census_border_fixup<TYPE, R>(set, sw, fix, x0, y0, a.H, a.W, a.eps, tid);
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

template <int TYPE, int R, bool GRAD>
__global__ void __launch_bounds__(NTHREADS, 2) census_pattern_loss_kernel(PatternLossArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float2* set = reinterpret_cast<float2*>(smem);   // (warped pattern, LCN image), replicate-clamped halo
  float* sw = smem + 2 * G::SIZE;                  // sigma (or 1), zero outside the image
  float* sdd = smem + 3 * G::SIZE;                 // d proj / d disp of the tile's own pixels
  float* fix = sdd + TH * TW;
  float* red = fix + NFIX;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;
  const size_t hw = (size_t)a.H * a.W;
  const float* disp = a.disp + (size_t)n * hw;
  const float* im = a.im + (size_t)n * hw;
  const float* sd = a.std_in ? a.std_in + (size_t)n * hw : nullptr;
#pragma unroll kCensusFillUnroll
  for (int idx = tid; idx < G::ROWS * G::COLS; idx += NTHREADS) {
    const int j = idx / G::COLS, i = idx - j * G::COLS;
    const int gy = y0 - R + j, gx = x0 - R + i;
    const bool inside = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
    const int cy = clampi(gy, 0, a.H - 1), cx = clampi(gx, 0, a.W - 1);
    const int g = cy * a.W + cx;
    const bool own = GRAD && j >= R && j < R + TH && i >= R && i < R + TW;
    float dd = 0.0f;
    const WarpRow row = warp_row_setup(cy, a.H, a.W, a.inv_h);
    const float e = warp_col_sample(a.pattern, row, __ldg(disp + g), cx, a.W, a.inv_w, own ? &dd : nullptr);
    set[j * G::PITCH + i] = make_float2(e, __ldg(im + g));
    sw[j * G::PITCH + i] = inside ? (sd ? __ldg(sd + g) : 1.0f) : 0.0f;
    if (own) sdd[(j - R) * TW + (i - R)] = dd;
  }
  __syncthreads();

  float acc[2][4], gacc[2][4], ec[2][4], tc[2][4], wc[2][4];
  census_patch<TYPE, R, true, GRAD>(set, sw, tx, ty, a.eps, acc, gacc, ec, tc, wc);
  if (!GRAD) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) wc[r][i] = sw[(2 * ty + r + R) * G::PITCH + 4 * tx + i + R];
  }
  const bool edge_tile = (x0 == 0) || (y0 == 0) || (x0 + TW >= a.W) || (y0 + TH >= a.H);
  if (GRAD && edge_tile) {
    census_border_fixup<TYPE, R>(set, sw, fix, x0, y0, a.H, a.W, a.eps, tid);
    __syncthreads();
  }
  const float fs = fwd_scale<TYPE>() * a.inv_k2;
  const float gk = (GRAD && a.grad_scale) ? a.inv_k2 * __ldg(a.grad_scale) : a.inv_k2;
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
        gv[i] = finish_grad<TYPE>(g, ec[r][i], tc[r][i], a.eps) * gk * sdd[(2 * ty + r) * TW + 4 * tx + i];
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

}  // namespace dis
