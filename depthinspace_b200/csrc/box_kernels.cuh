// mse / sad window losses as separable box filters (O(1) work per pixel instead of k^2 taps).
//
// reference: photometric_loss_pytorch types 'mse' / 'sad' (model/ext_functions.py:165-168): the k x k window sum
// of a POINT-WISE term s(q) = (e-t)^2 or |e-t| over the replicate-padded planes is a box filter of s, and the
// gradient is f'(e-t)(x) times the (multiplicity-aware) box filter of the zero-extended upstream weights:
//   out(p)  = (1/k^2) sum_o s(clamp(p+o))
//   grad(x) = f'(e(x)-t(x)) (1/k^2) sum_o m(x,o) Wz(x+o)          (see window.cuh for m)
// Same 64x32 tile / 16x16 threads / 4x2 patch per thread as the census kernels, so the epilogues match.
// Pass 1 (horizontal): sliding (2R+1)-sums of 4 adjacent columns from one 128-bit-aligned register row.
// Pass 2 (vertical):   each thread sums the 2R+2 rows above/below its 4x2 patch.
// These kernels are HBM-bound (the census kernels are not): ~25 instructions per pixel on top of the tile load.
#pragma once
#include "photometric_kernels.cuh"

namespace dis {

template <int R>
struct BoxGeom {
  using G = TileGeom<R>;
  static constexpr int HP = TW;                       // pitch of the horizontally summed planes
  static constexpr int HSIZE = G::ROWS * HP;
};

template <int TYPE>
__device__ __forceinline__ float point_term(float e, float t) {
  const float d = e - t;
  return TYPE == MSE ? d * d : fabsf(d);
}
template <int TYPE>
__device__ __forceinline__ float point_deriv(float e, float t) {
  return TYPE == MSE ? 2.0f * (e - t) : sign0(e - t);
}

// horizontal pass over one plane: src [ROWS][G::PITCH] (tile origin -R) -> dst [ROWS][TW]
template <int R>
__device__ __forceinline__ void box_rows(const float* __restrict__ src, float* __restrict__ dst, int tid) {
  using G = TileGeom<R>;
  for (int item = tid; item < G::ROWS * (TW / 4); item += NTHREADS) {
    const int row = item / (TW / 4), q = item - row * (TW / 4);
    float v[4 * G::NV];
#pragma unroll
    for (int k = 0; k < G::NV; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(src + row * G::PITCH + 4 * q + 4 * k);
      v[4 * k] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = a.z; v[4 * k + 3] = a.w;
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) s += v[k];
    float o[4];
    o[0] = s;
#pragma unroll
    for (int c = 1; c < 4; ++c) { s += v[c + 2 * R] - v[c - 1]; o[c] = s; }
    *reinterpret_cast<float4*>(dst + row * BoxGeom<R>::HP + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// vertical pass for the thread's 4x2 patch: rows (2ty .. 2ty+1), cols (4tx .. 4tx+3) of the tile
template <int R>
__device__ __forceinline__ void box_cols(const float* __restrict__ hs, int tx, int ty, float (&out)[2][4]) {
  constexpr int HP = BoxGeom<R>::HP;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = hs + (2 * ty) * HP + 4 * tx;   // tile row 2ty - R  == plane row 2ty
#pragma unroll
  for (int j = 0; j <= 2 * R; ++j) {
    const float4 a = *reinterpret_cast<const float4*>(p + j * HP);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  out[0][0] = acc.x; out[0][1] = acc.y; out[0][2] = acc.z; out[0][3] = acc.w;
  const float4 top = *reinterpret_cast<const float4*>(p);
  const float4 bot = *reinterpret_cast<const float4*>(p + (2 * R + 1) * HP);
  out[1][0] = acc.x + (bot.x - top.x); out[1][1] = acc.y + (bot.y - top.y);
  out[1][2] = acc.z + (bot.z - top.z); out[1][3] = acc.w + (bot.w - top.w);
}

template <int R>
constexpr size_t box_smem_bytes(int planes, int hplanes, int own_planes) {
  return sizeof(float) * ((size_t)planes * TileGeom<R>::SIZE + (size_t)hplanes * BoxGeom<R>::HSIZE +
                          (size_t)own_planes * TH * TW + NFIX + 2 * (NTHREADS / 32));
}

// ---------------------------------------------------------------------------------------------------------------
template <int TYPE, int R>
__global__ void __launch_bounds__(NTHREADS, 3) box_photometric_fwd_kernel(PhotoArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float* sS = smem;
  float* hS = smem + G::SIZE;
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
      sS[idx] = point_term<TYPE>(__ldg(e + g), __ldg(t + g));
    }
    __syncthreads();
    box_rows<R>(sS, hS, tid);
    __syncthreads();
    float o[2][4];
    box_cols<R>(hS, tx, ty, o);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) total[r][i] += o[r][i];
  }
  float* out = a.out + (size_t)n * hw;
#pragma unroll
  for (int r = 0; r < 2; ++r)
    store_quad(out, y0 + 2 * ty + r, x0 + 4 * tx, a.H, a.W, a.vec_ok, total[r][0] * a.inv_k2, total[r][1] * a.inv_k2,
               total[r][2] * a.inv_k2, total[r][3] * a.inv_k2);
}

// gsum of the thread's patch with the border-line fix-up applied; sW: zero-extended weights [ROWS][PITCH]
template <int TYPE, int R>
__device__ __forceinline__ void box_weight_sums(const float* __restrict__ sW, float* __restrict__ hW, float* __restrict__ fix,
                                                int tx, int ty, int tid, int x0, int y0, int H, int W, float (&gsum)[2][4]) {
  box_rows<R>(sW, hW, tid);
  const bool edge_tile = (x0 == 0) || (y0 == 0) || (x0 + TW >= W) || (y0 + TH >= H);
  if (edge_tile) border_fixup<TYPE, R>(sW, sW, sW, fix, x0, y0, H, W, 0.0f, tid);   // mse/sad: only the weights matter
  __syncthreads();
  box_cols<R>(hW, tx, ty, gsum);
  if (edge_tile) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gy = y0 + 2 * ty + r, gx = x0 + 4 * tx + i;
        if (gy < H && gx < W) {
          const int slot = border_slot(2 * ty + r, 4 * tx + i, gy, gx, H, W);
          if (slot >= 0) gsum[r][i] = fix[slot];
        }
      }
  }
}

template <int TYPE, int R>
__global__ void __launch_bounds__(NTHREADS, 3) box_photometric_bwd_kernel(PhotoArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* hW = smem + G::SIZE;
  float* fix = hW + BoxGeom<R>::HSIZE;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;
  const size_t hw = (size_t)a.H * a.W;
  const float* go = a.grad_out + (size_t)n * hw;
#pragma unroll 4
  for (int idx = tid; idx < G::ROWS * G::PITCH; idx += NTHREADS) {
    const int j = idx / G::PITCH, i = idx - j * G::PITCH;
    const int gy = y0 - R + j, gx = x0 - R + i;
    sW[idx] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) ? __ldg(go + (size_t)gy * a.W + gx) : 0.0f;
  }
  __syncthreads();
  float gsum[2][4];
  box_weight_sums<TYPE, R>(sW, hW, fix, tx, ty, tid, x0, y0, a.H, a.W, gsum);
  // the window weights are shared by every channel: only the point-wise derivative changes
  for (int c = 0; c < a.C; ++c) {
    const float* e = a.es + ((size_t)n * a.C + c) * hw;
    const float* t = a.ta + ((size_t)n * a.C + c) * hw;
    float* ge = a.grad_es + ((size_t)n * a.C + c) * hw;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int gy = y0 + 2 * ty + r, gx = x0 + 4 * tx;
      if (gy >= a.H || gx >= a.W) continue;
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const size_t o = (size_t)gy * a.W + min(gx + i, a.W - 1);
        v[i] = point_deriv<TYPE>(__ldg(e + o), __ldg(t + o)) * gsum[r][i] * a.inv_k2;
      }
      store_quad(ge, gy, gx, a.H, a.W, a.vec_ok, v[0], v[1], v[2], v[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <int TYPE, int R, bool GRAD>
__global__ void __launch_bounds__(NTHREADS, 3) box_pattern_loss_kernel(PatternLossArgs a) {
  using G = TileGeom<R>;
  extern __shared__ __align__(16) float smem[];
  float* sS = smem;                                   // point-wise term, replicate-clamped halo
  float* sW = smem + G::SIZE;                         // sigma (or 1), zero outside the image
  float* hS = smem + 2 * G::SIZE;
  float* hW = hS + BoxGeom<R>::HSIZE;
  float* sE = hW + BoxGeom<R>::HSIZE;                 // own pixels: warped pattern (for proj output)
  float* sF = sE + TH * TW;                           // own pixels: f'(e - t) * d proj / d disp
  float* fix = sF + TH * TW;
  float* red = fix + NFIX;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;
  const size_t hw = (size_t)a.H * a.W;
  const float* disp = a.disp + (size_t)n * hw;
  const float* im = a.im + (size_t)n * hw;
  const float* sd = a.std_in ? a.std_in + (size_t)n * hw : nullptr;
  // latency-bound staging (disp load -> 4 dependent pattern gathers): keep four positions in flight
#pragma unroll 4
  for (int idx = tid; idx < G::ROWS * G::PITCH; idx += NTHREADS) {
    const int j = idx / G::PITCH, i = idx - j * G::PITCH;
    const int gy = y0 - R + j, gx = x0 - R + i;
    const bool inside = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
    const int cy = clampi(gy, 0, a.H - 1), cx = clampi(gx, 0, a.W - 1);
    const int g = cy * a.W + cx;
    const bool own = j >= R && j < R + TH && i >= R && i < R + TW;
    float dd = 0.0f;
    const WarpRow row = warp_row_setup(cy, a.H, a.W, a.inv_h);
    const float e = warp_col_sample(a.pattern, row, __ldg(disp + g), cx, a.W, a.inv_w, (GRAD && own) ? &dd : nullptr);
    const float t = __ldg(im + g);
    sS[idx] = point_term<TYPE>(e, t);
    sW[idx] = inside ? (sd ? __ldg(sd + g) : 1.0f) : 0.0f;
    if (own) {
      sE[(j - R) * TW + (i - R)] = e;
      if (GRAD) sF[(j - R) * TW + (i - R)] = point_deriv<TYPE>(e, t) * dd;
    }
  }
  __syncthreads();
  box_rows<R>(sS, hS, tid);
  float gsum[2][4];
  if (GRAD) box_weight_sums<TYPE, R>(sW, hW, fix, tx, ty, tid, x0, y0, a.H, a.W, gsum);   // contains the barrier
  else __syncthreads();
  float acc[2][4];
  box_cols<R>(hS, tx, ty, acc);

  const float gk = (GRAD && a.grad_scale) ? a.inv_k2 * __ldg(a.grad_scale) : a.inv_k2;
  float num = 0.0f, den = 0.0f;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int gy = y0 + 2 * ty + r;
    float d[4], gv[4], ev[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gx = x0 + 4 * tx + i;
      const bool valid = gy < a.H && gx < a.W;
      const float wc = sW[(2 * ty + r + R) * G::PITCH + 4 * tx + i + R];
      d[i] = acc[r][i] * a.inv_k2;
      if (valid) { num = fmaf(wc, d[i], num); den += wc; }
      ev[i] = sE[(2 * ty + r) * TW + 4 * tx + i];
      if (GRAD) gv[i] = sF[(2 * ty + r) * TW + 4 * tx + i] * gsum[r][i] * gk;
    }
    if (a.diff) store_quad(a.diff + (size_t)n * hw, gy, x0 + 4 * tx, a.H, a.W, a.vec_ok, d[0], d[1], d[2], d[3]);
    if (a.proj) store_quad(a.proj + (size_t)n * hw, gy, x0 + 4 * tx, a.H, a.W, a.vec_ok, ev[0], ev[1], ev[2], ev[3]);
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
