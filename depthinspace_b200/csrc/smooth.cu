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

template <bool GRAD>
__global__ void __launch_bounds__(SNT) smooth_loss_kernel(const float* __restrict__ disp, const float* __restrict__ im,
                                                          float* __restrict__ grad_sum, float* __restrict__ partials,
                                                          int H, int W) {
  __shared__ float sd[IN_H * IN_P];
  __shared__ float si[IN_H * IN_P];
  __shared__ float2 su[GRAD ? U_H * U_P : 1];
  __shared__ float red[2 * (SNT / 32)];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * STW, y0 = blockIdx.y * STH, n = blockIdx.z;
  const size_t hw = (size_t)H * W;
  const float* d = disp + (size_t)n * hw;
  const float* a = im + (size_t)n * hw;
  for (int idx = tid; idx < IN_H * IN_W; idx += SNT) {
    const int j = idx / IN_W, i = idx - j * IN_W;
    const size_t g = (size_t)clampi(y0 - 4 + j, 0, H - 1) * W + clampi(x0 - 4 + i, 0, W - 1);
    sd[j * IN_P + i] = __ldg(d + g);
    si[j * IN_P + i] = __ldg(a + g);
  }
  __syncthreads();

  // u over the tile plus a halo of 2; the loss itself only over the tile's own pixels
  float lsum = 0.f, lcnt = 0.f;
  for (int idx = tid; idx < U_H * U_W; idx += SNT) {
    const int j = idx / U_W, i = idx - j * U_W;
    const int gy = y0 - 2 + j, gx = x0 - 2 + i;
    const bool own = j >= 2 && j < 2 + STH && i >= 2 && i < 2 + STW;
    if (!GRAD && !own) continue;
    float2 u = make_float2(0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      float gdx, gdy, gix, giy;
      sobel5_at(sd + j * IN_P + i, IN_P, gdx, gdy);  // input tile origin is (-4,-4): window top-left = (j, i)
      sobel5_at(si + j * IN_P + i, IN_P, gix, giy);
      const float ax = expf(-fabsf(255.0f * gix)), ay = expf(-fabsf(255.0f * giy));
      const float vx = gdx * ax, vy = gdy * ay;
      if (own) { lsum += fabsf(vx) + fabsf(vy); lcnt += 2.0f; }
      u = make_float2(sign0(vx) * ax, sign0(vy) * ay);
    }
    if (GRAD) su[j * U_P + i] = u;
  }
  if (GRAD) {
    __syncthreads();
    float* go = grad_sum + (size_t)n * hw;
    auto u_at = [&](int ly, int lx) -> float2 {  // tile-local coords; zero beyond the stored halo (= outside image)
      if (ly < -2 || ly >= STH + 2 || lx < -2 || lx >= STW + 2) return make_float2(0.f, 0.f);
      return su[(ly + 2) * U_P + lx + 2];
    };
    for (int idx = tid; idx < STH * STW; idx += SNT) {
      const int ly = idx / STW, lx = idx - ly * STW;
      const int gy = y0 + ly, gx = x0 + lx;
      if (gy >= H || gx >= W) continue;
      // pre-image of (gy,gx) under replicate clamping of the pad-2 domain
      const int ylo = (gy == 0) ? -2 : 0, yhi = (gy == H - 1) ? 2 : 0;
      const int xlo = (gx == 0) ? -2 : 0, xhi = (gx == W - 1) ? 2 : 0;
      float acc = 0.f;
      for (int py = ylo; py <= yhi; ++py)
        for (int px = xlo; px <= xhi; ++px) acc += sobel5_adjoint(u_at, ly + py, lx + px);
      go[(size_t)gy * W + gx] = acc;
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
                        cudaStream_t s) {
  dim3 grid((W + STW - 1) / STW, (H + STH - 1) / STH, N);
  if (grad_sum) smooth_loss_kernel<true><<<grid, SNT, 0, s>>>(disp, im, grad_sum, partials, H, W);
  else smooth_loss_kernel<false><<<grid, SNT, 0, s>>>(disp, im, nullptr, partials, H, W);
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
