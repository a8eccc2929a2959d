// Sobel filter and edge-aware disparity smoothness.
// reference: SobelFilter (model/networks.py:697-730), DisparitySmoothLoss.tforward (:419-431)
//   g  = sobel5x5(replicate_pad2(disp)) -> (gx, gy);  gi = sobel5x5(replicate_pad2(im))
//   val = mean | g * exp(-|255 gi|) |   over [N,2,H,W]
//
// smooth_loss_kernel does the whole forward in one pass over a 64x32 tile and, on request, the exact
// gradient w.r.t. disp in the same pass (gather form of the adjoint: conv-transpose of
// u = sign(g a) a, followed by the adjoint of the replicate padding, folded into the border pixels).
#include "common.cuh"

namespace dis {
namespace {

constexpr int STW = 64, STH = 32, SNT = 256;
#ifndef DIS_SMOOTH_FILL_UNROLL
#define DIS_SMOOTH_FILL_UNROLL 4
#endif
constexpr int kSmoothFillUnroll = DIS_SMOOTH_FILL_UNROLL;   // staging loads in flight per thread (the phase is latency-bound)
constexpr int IN_H = STH + 8, IN_W = STW + 8, IN_P = IN_W;      // inputs with halo 4 (replicate-clamped)
constexpr int U_H = STH + 4, U_W = STW + 4, U_P = U_W;          // u with halo 2 (zero outside the image)

// kx[i][j] / 240 rounded to float exactly like torch.from_numpy(kx).float()  (:701-705); ky = kx^T
#define SOB(v) ((float)((v) / 240.0))
__device__ __forceinline__ void sobel5_at(const float* __restrict__ t, int pitch, float& gx, float& gy) {
  // t points at the window's top-left element (row 0, col 0 of the 5x5 support)
  constexpr float k[5][5] = {{SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)},
                             {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                             {SOB(-10.0), SOB(-20.0), 0.f, SOB(20.0), SOB(10.0)},
                             {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                             {SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)}};
  float ax = 0.f, ay = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float v = t[i * pitch + j];
      if (j != 2) ax = fmaf(k[i][j], v, ax);
      if (i != 2) ay = fmaf(k[j][i], v, ay);
    }
  gx = ax;
  gy = ay;
}
// adjoint tap: sum_{i,j} kx[i][j] * ux(q - (i-2, j-2)) + ky[i][j] * uy(q - (i-2, j-2))
template <typename F>
__device__ __forceinline__ float sobel5_adjoint(F u_at, int qy, int qx) {
  constexpr float k[5][5] = {{SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)},
                             {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                             {SOB(-10.0), SOB(-20.0), 0.f, SOB(20.0), SOB(10.0)},
                             {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                             {SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)}};
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float2 u = u_at(qy - (i - 2), qx - (j - 2));
      if (j != 2) acc = fmaf(k[i][j], u.x, acc);
      if (i != 2) acc = fmaf(k[j][i], u.y, acc);
    }
  return acc;
}

__device__ __forceinline__ void load_row8(const float* __restrict__ p, float (&r)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}

// 4 horizontally adjacent Sobel responses from a 5 x 8 register window of (disp, ambient) pairs: every FFMA2
// advances both images.
// out: gx2[q] = (sobel_x(disp), sobel_x(amb)), gy2[q] likewise, for 4 adjacent responses.
__device__ __forceinline__ void sobel5_quad2(const u64 (&win)[5][8], u64 (&gx2)[4], u64 (&gy2)[4]) {
  constexpr float k[5][5] = {{SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)},
                             {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                             {SOB(-10.0), SOB(-20.0), 0.f, SOB(20.0), SOB(10.0)},
                             {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                             {SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)}};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    u64 ax = pk2(0.f, 0.f), ay = pk2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        if (j != 2) ax = fma2(bc2(k[i][j]), win[i][q + j], ax);
        if (i != 2) ay = fma2(bc2(k[j][i]), win[i][q + j], ay);
      }
    gx2[q] = ax;
    gy2[q] = ay;
  }
}

__device__ __forceinline__ void load_row8x2(const float2* __restrict__ p, u64 (&r)[8]) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 a = *reinterpret_cast<const float4*>(p + 2 * v);
    r[2 * v] = pk2(a.x, a.y);
    r[2 * v + 1] = pk2(a.z, a.w);
  }
}

// Tile 64 x 32.  Shared planes of float2, pitches chosen so that 128-bit row loads are 16-byte aligned and, for
// threads that walk down consecutive rows, bank-conflict free (pitch * 2 = 20 or 12 mod 32):
//   sin : (disp, ambient), replicate-clamped, origin (-4,-4): 40 rows x 72 cols, pitch 74
//   sux, suy : u = sign(g a) a, zero outside the image, origin (-2,-2): 36 rows x 68 cols (plain float planes: the
//              adjoint taps carry different weights for u_x and u_y, and only a scalar FFMA takes an immediate)
// Responses are produced in quads starting at tile-local column -2 + 4q, so their 8-wide input windows
// (origin -4 + 4q) and their stores (origin -2 + 4q) are both aligned; the adjoint quads start at 4q and read
// the u window starting at 4q - 2, aligned again.  Both images go through the Sobel taps together as fp32x2.
constexpr int PIN = 74, PU = U_P;

template <bool GRAD>
__global__ void __launch_bounds__(SNT) smooth_loss_kernel(const float* __restrict__ disp, const float* __restrict__ im,
                                                          float* __restrict__ grad_sum, float* __restrict__ partials,
                                                          int H, int W, int vec_ok, float grad_scale, int accumulate) {
  __shared__ __align__(16) float2 sin[IN_H * PIN];
  __shared__ __align__(16) float sux[GRAD ? U_H * PU : 4];
  __shared__ __align__(16) float suy[GRAD ? U_H * PU : 4];
  __shared__ float red[2 * (SNT / 32)];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * STW, y0 = blockIdx.y * STH, n = blockIdx.z;
  const size_t hw = (size_t)H * W;
  const float* d = disp + (size_t)n * hw;
  const float* a = im + (size_t)n * hw;
#pragma unroll kSmoothFillUnroll
  for (int idx = tid; idx < IN_H * IN_W; idx += SNT) {
    const int j = idx / IN_W, i = idx - j * IN_W;
    const int g = clampi(y0 - 4 + j, 0, H - 1) * W + clampi(x0 - 4 + i, 0, W - 1);
    sin[j * PIN + i] = make_float2(__ldg(d + g), __ldg(a + g));
  }
  __syncthreads();

  float lsum = 0.f, lcnt = 0.f;
  constexpr int UROWS = GRAD ? U_H : STH;   // rows -2..33 with the gradient, 0..31 without
  constexpr int UQ = U_W / 4;               // 17 quads: columns -2..65
  for (int item = tid; item < UROWS * UQ; item += SNT) {
    const int q = item / UROWS, jr = item - q * UROWS;   // consecutive threads walk down the rows of one quad
    const int ly = GRAD ? jr - 2 : jr;      // tile-local row of the responses
    const int lx = 4 * q - 2;               // tile-local column of the first response
    const int gy = y0 + ly;
    float ux[4] = {0.f, 0.f, 0.f, 0.f}, uy[4] = {0.f, 0.f, 0.f, 0.f};
    if (gy >= 0 && gy < H) {
      // response (ly, lx+c) reads tile-local rows ly-2..ly+2, cols lx+c-2..lx+c+2 = smem rows ly+2.., cols 4q+c..
      u64 win[5][8];
#pragma unroll
      for (int i = 0; i < 5; ++i) load_row8x2(sin + (ly + 2 + i) * PIN + 4 * q, win[i]);
      u64 gx2[4], gy2[4];
      sobel5_quad2(win, gx2, gy2);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int gx = x0 + lx + c;
        if (gx >= 0 && gx < W) {
          float gdx, gix, gdy, giy;
          upk2(gx2[c], gdx, gix);
          upk2(gy2[c], gdy, giy);
          const float ax = expf(-fabsf(255.0f * gix)), ay = expf(-fabsf(255.0f * giy));
          const float vx = gdx * ax, vy = gdy * ay;
          if (ly >= 0 && ly < STH && lx + c >= 0 && lx + c < STW) { lsum += fabsf(vx) + fabsf(vy); lcnt += 2.0f; }
          ux[c] = sign0(vx) * ax;
          uy[c] = sign0(vy) * ay;
        }
      }
    }
    if (GRAD) {  // response (ly, lx+c) lives at u-plane (ly+2, 4q+c)
      *reinterpret_cast<float4*>(sux + (ly + 2) * PU + 4 * q) = make_float4(ux[0], ux[1], ux[2], ux[3]);
      *reinterpret_cast<float4*>(suy + (ly + 2) * PU + 4 * q) = make_float4(uy[0], uy[1], uy[2], uy[3]);
    }
  }
  if (GRAD) {
    __syncthreads();
    float* go = grad_sum + (size_t)n * hw;
    constexpr float k[5][5] = {{SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)},
                               {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                               {SOB(-10.0), SOB(-20.0), 0.f, SOB(20.0), SOB(10.0)},
                               {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                               {SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)}};
    auto u_at = [&](int ly, int lx) -> float2 {  // tile-local; zero beyond the stored halo (= outside the image)
      if (ly < -2 || ly >= STH + 2 || lx < -2 || lx >= STW + 2) return make_float2(0.f, 0.f);
      return make_float2(sux[(ly + 2) * PU + lx + 2], suy[(ly + 2) * PU + lx + 2]);
    };
    for (int item = tid; item < STH * (STW / 4); item += SNT) {
      const int ly = item / (STW / 4), lx = 4 * (item - ly * (STW / 4));
      const int gy = y0 + ly, gx = x0 + lx;
      if (gy >= H || gx >= W) continue;
      // grad(q) = sum_{i,j} kx[i][j] ux(q - (i-2, j-2)) + ky[i][j] uy(q - (i-2, j-2))
      // u window: tile-local rows ly-2..ly+2, cols lx-2..lx+5 = u-plane rows ly.., cols lx..lx+7
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 5; ++r) {   // u row offset r-2  => i = 4 - r
        float vx[8], vy[8];
        load_row8(sux + (ly + r) * PU + lx, vx);
        load_row8(suy + (ly + r) * PU + lx, vy);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int t = 0; t < 5; ++t) {  // u col offset t-2 => j = 4 - t
            const int i = 4 - r, j = 4 - t;
            if (j != 2) acc[c] = fmaf(k[i][j], vx[c + t], acc[c]);
            if (i != 2) acc[c] = fmaf(k[j][i], vy[c + t], acc[c]);
          }
      }
      // border-line pixels also collect the pad-2 positions that replicate-clamp onto them
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int px_ = gx + c;
        if (px_ < W && (gy == 0 || gy == H - 1 || px_ == 0 || px_ == W - 1)) {
          const int ylo = (gy == 0) ? -2 : 0, yhi = (gy == H - 1) ? 2 : 0;
          const int xlo = (px_ == 0) ? -2 : 0, xhi = (px_ == W - 1) ? 2 : 0;
          float v = 0.f;
          for (int py2 = ylo; py2 <= yhi; ++py2)
            for (int px2 = xlo; px2 <= xhi; ++px2) v += sobel5_adjoint(u_at, ly + py2, lx + c + px2);
          acc[c] = v;
        }
      }
      float* o = go + (size_t)gy * W + gx;
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] *= grad_scale;      // 1 for the plain sum; weight / count for a final gradient
      if (accumulate) {                                       // add to a gradient another term already left there
        if (vec_ok && gx + 3 < W) {
          const float4 old = *reinterpret_cast<const float4*>(o);
          acc[0] += old.x; acc[1] += old.y; acc[2] += old.z; acc[3] += old.w;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) if (gx + c < W) acc[c] += o[c];
        }
      }
      if (vec_ok && gx + 3 < W) __stcs(reinterpret_cast<float4*>(o), make_float4(acc[0], acc[1], acc[2], acc[3]));
      else {
#pragma unroll
        for (int c = 0; c < 4; ++c) if (gx + c < W) o[c] = acc[c];
      }
    }
  }
  block_sum2<SNT>(lsum, lcnt, red);
  if (tid == 0) {
    const size_t b = ((size_t)n * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partials[2 * b] = lsum;
    partials[2 * b + 1] = lcnt;
  }
}

// Stand-alone Sobel (API completeness; direct clamped global reads, one thread per pixel)
__global__ void __launch_bounds__(256) sobel_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int H,
                                                        int W, int ksize, size_t total) {
  const size_t hw = (size_t)H * W;
  const int r = ksize / 2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    const float* p = x + n * hw;
    float gx = 0.f, gy = 0.f;
    if (ksize == 5) {
      float win[25];
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) win[i * 5 + j] = __ldg(p + (size_t)clampi(h + i - r, 0, H - 1) * W + clampi(w + j - r, 0, W - 1));
      sobel5_at(win, 5, gx, gy);
    } else {
      constexpr float k3[3][3] = {{(float)(-1.0 / 8.0), 0.f, (float)(1.0 / 8.0)},
                                  {(float)(-2.0 / 8.0), 0.f, (float)(2.0 / 8.0)},
                                  {(float)(-1.0 / 8.0), 0.f, (float)(1.0 / 8.0)}};
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float v = __ldg(p + (size_t)clampi(h + i - r, 0, H - 1) * W + clampi(w + j - r, 0, W - 1));
          gx = fmaf(k3[i][j], v, gx);
          gy = fmaf(k3[j][i], v, gy);
        }
    }
    out[(n * 2 + 0) * hw + pix] = gx;
    out[(n * 2 + 1) * hw + pix] = gy;
  }
}

// adjoint of sobel_fwd_kernel in gather form (atomics-free, deterministic)
__global__ void __launch_bounds__(256) sobel_bwd_kernel(const float* __restrict__ go, float* __restrict__ gx_out, int H,
                                                        int W, int ksize, size_t total) {
  const size_t hw = (size_t)H * W;
  const int r = ksize / 2;
  const double div = ksize == 5 ? 240.0 : 8.0;
  constexpr double k5[5][5] = {{-5, -4, 0, 4, 5}, {-8, -10, 0, 10, 8}, {-10, -20, 0, 20, 10}, {-8, -10, 0, 10, 8}, {-5, -4, 0, 4, 5}};
  constexpr double k3[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    const float* ux = go + (n * 2 + 0) * hw;
    const float* uy = go + (n * 2 + 1) * hw;
    const int ylo = (h == 0) ? -r : 0, yhi = (h == H - 1) ? r : 0;
    const int xlo = (w == 0) ? -r : 0, xhi = (w == W - 1) ? r : 0;
    float acc = 0.f;
    for (int py = ylo; py <= yhi; ++py)
      for (int px = xlo; px <= xhi; ++px)
        for (int i = 0; i < ksize; ++i)
          for (int j = 0; j < ksize; ++j) {
            const int sy = h + py - (i - r), sx = w + px - (j - r);
            if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
            const float kx = (float)((ksize == 5 ? k5[i][j] : k3[i][j]) / div);
            const float ky = (float)((ksize == 5 ? k5[j][i] : k3[j][i]) / div);
            acc = fmaf(kx, __ldg(ux + (size_t)sy * W + sx), acc);
            acc = fmaf(ky, __ldg(uy + (size_t)sy * W + sx), acc);
          }
    gx_out[idx] = acc;
  }
}

inline int flat_grid(size_t total) {
  const size_t want = (total + 255) / 256, cap = 148 * 16;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

int smooth_loss_num_partials(int N, int H, int W) { return N * ((H + STH - 1) / STH) * ((W + STW - 1) / STW); }

int smooth_loss_forward(const float* disp, const float* im, float* grad_sum, float* partials, int N, int H, int W,
                        float grad_scale, int accumulate, cudaStream_t s) {
  dim3 grid((W + STW - 1) / STW, (H + STH - 1) / STH, N);
  const int vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(grad_sum) & 15) == 0);
  if (grad_sum) smooth_loss_kernel<true><<<grid, SNT, 0, s>>>(disp, im, grad_sum, partials, H, W, vec_ok, grad_scale, accumulate);
  else smooth_loss_kernel<false><<<grid, SNT, 0, s>>>(disp, im, nullptr, partials, H, W, vec_ok, grad_scale, 0);
  return check_launch();
}

int sobel_forward(const float* x, float* out, int N, int H, int W, int ksize, cudaStream_t s) {
  const size_t total = (size_t)N * H * W;
  sobel_fwd_kernel<<<flat_grid(total), 256, 0, s>>>(x, out, H, W, ksize, total);
  return check_launch();
}

int sobel_backward(const float* go, float* gx, int N, int H, int W, int ksize, cudaStream_t s) {
  const size_t total = (size_t)N * H * W;
  sobel_bwd_kernel<<<flat_grid(total), 256, 0, s>>>(go, gx, H, W, ksize, total);
  return check_launch();
}

}  // namespace dis
