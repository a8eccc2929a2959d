// k x k block-window machinery shared by the photometric-loss kernels.
//
// A CTA of 256 threads (16 x 16) owns a 64 x 32 pixel output tile; every thread owns a
// 4 (x) by 2 (y) pixel patch.  The estimate e, target t and weight w planes are staged in
// shared memory WITH their halo:
//   * e and t are stored replicate-clamped, i.e. exactly the padded planes the reference
//     builds with F.pad(mode='replicate') (model/ext_functions.py:158-159), so window taps
//     never need index arithmetic;
//   * w (the per-pixel upstream weight of the backward pass) is stored ZERO outside the
//     image, which turns the scatter of the reference's autograd into an atomics-free
//     gather (see grad derivation below).
// Each thread walks the 2R+2 shared-memory rows that intersect its two windows, pulls the
// 4+2R contiguous values per plane with 128-bit loads into registers and evaluates its
// 4 x 2 x (2R+1) taps from registers: ~0.03 LDS per tap (main loops: census_kernels.cuh for
// the soft-census types, box_kernels.cuh for mse / sad).
//
// Gradient in gather form.  With W = grad_out / k^2 and phi(x->q) the derivative of the
// per-tap term w.r.t. the neighbour value (odd under swapping centre and neighbour),
//   census:  grad(x) = -sum_o ( m(x,o) * Wz(x+o) + W(x) ) * phi(x -> clamp(x+o))
//   mse/sad: grad(x) = f'(e(x) - t(x)) * sum_o m(x,o) * Wz(x+o)
// where Wz is W zero-extended and m(x,o) counts how many taps of window(x+o) clamp onto x:
// m = 1 unless x lies on the image border line.  The main loop uses m = 1; the
// O(H + W) border-line pixels are recomputed with the exact multiplicity afterwards.
#pragma once
#include "common.cuh"

namespace dis {

constexpr int TW = 64;         // tile width  (16 threads x 4 px)
constexpr int TH = 32;         // tile height (16 threads x 2 px)
constexpr int NTHREADS = 256;
constexpr int MAX_R = 7;       // block_size <= 15
constexpr int NFIX = 2 * TH + 2 * TW;  // border-line candidates per tile

template <int R>
struct TileGeom {
  static constexpr int NV = (4 + 2 * R + 3) / 4;  // float4 per thread-row window
  static constexpr int PITCH = 60 + 4 * NV;       // >= TW + 2R, multiple of 4
  static constexpr int ROWS = TH + 2 * R;
  static constexpr int SIZE = ROWS * PITCH;       // floats per plane
  static constexpr int COLS = TW + 2 * R;         // columns that carry data
};

enum LossType { MSE = 0, SAD = 1, CENSUS_MSE = 2, CENSUS_SAD = 3 };

// sign(v) * mag for mag > 0, and 0 when v == 0 (torch.abs has subgradient 0 at 0): LOP3 + FSETP + FSEL
__device__ __forceinline__ float signed_mag(float mag, float v) {
  const float m = __int_as_float(__float_as_int(mag) | (__float_as_int(v) & 0x80000000));
  return (v == 0.0f) ? 0.0f : m;
}

// One window tap.  acc: forward accumulator.  Backward accumulators: ga = sum u * Wz(q) and, for the
// census types, gb = sum u (the centre weight W(x) is applied once at the end: gacc = ga + W(x) * gb).
template <int TYPE, bool FWD, bool BWD>
__device__ __forceinline__ void tap(float ec, float tc, float eq, float tq, float wq, float eps,
                                    float& acc, float& ga, float& gb) {
  if (TYPE == MSE) {
    if (FWD) { const float d = eq - tq; acc = fmaf(d, d, acc); }
    if (BWD) ga += wq;
  } else if (TYPE == SAD) {
    if (FWD) acc += fabsf(eq - tq);
    if (BWD) ga += wq;
  } else {
    // soft census: h(d) = 0.5 (1 + d / sqrt(d^2 + eps));  diff2 = 2 (h(de) - h(dt))
    const float de = eq - ec, dt = tq - tc;
    const float re = rsqrt_fast(fmaf(de, de, eps));
    const float rt = rsqrt_fast(fmaf(dt, dt, eps));
    // both products rounded separately (no FMA contraction): e == t must give diff2 == 0 exactly,
    // because the reference's |.| has subgradient 0 there
    const float diff2 = __fmul_rn(de, re) - __fmul_rn(dt, rt);
    if (FWD) acc = (TYPE == CENSUS_MSE) ? fmaf(diff2, diff2, acc) : acc + fabsf(diff2);
    if (BWD) {
      const float r3 = re * re * re;  // h'(de) = 0.5 eps r^3
      const float u = (TYPE == CENSUS_MSE) ? diff2 * r3 : signed_mag(r3, diff2);
      ga = fmaf(u, wq, ga);
      gb += u;
    }
  }
}
// forward scale: out = acc * fwd_scale / k^2 ; backward: grad = gacc * bwd_scale(...) / k^2
template <int TYPE>
__device__ __forceinline__ float fwd_scale() {
  return TYPE == CENSUS_MSE ? 0.25f : (TYPE == CENSUS_SAD ? 0.5f : 1.0f);
}
// turns gacc into d(sum_p W(p) out(p)) / d e(x) given the centre values
template <int TYPE>
__device__ __forceinline__ float finish_grad(float gacc, float ec, float tc, float eps) {
  if (TYPE == MSE) return 2.0f * (ec - tc) * gacc;
  if (TYPE == SAD) return sign0(ec - tc) * gacc;
  // census_mse: f' = 2 (diff2 / 2), census_sad: f' = sign(diff2); both times h'(de) = 0.5 eps r^3
  return -0.5f * eps * gacc;
}

// How many offsets o in [-R,R] satisfy clamp(p + o, 0, size-1) == x  (p, x in [0,size)).
__device__ __forceinline__ int clamp_multiplicity(int p, int x, int size, int R) {
  const int lo = (x == 0) ? -(1 << 20) : x;
  const int hi = (x == size - 1) ? (1 << 20) : x;
  return max(0, min(R, hi - p) - max(-R, lo - p) + 1);
}

// Exact backward accumulator of one border-line pixel (tile-local coords ly, lx; global gy, gx).
template <int TYPE, int R>
__device__ float border_pixel_gacc(const float* __restrict__ se, const float* __restrict__ st,
                                   const float* __restrict__ sw, int ly, int lx, int gy, int gx, int H, int W,
                                   float eps) {
  using G = TileGeom<R>;
  constexpr bool CENSUS = (TYPE >= CENSUS_MSE);
  const int c = (ly + R) * G::PITCH + lx + R;
  const float ec = se[c], tc = st[c], wc = sw[c];
  float acc = 0.0f, ga = 0.0f, gb = 0.0f;
  for (int dy = -R; dy <= R; ++dy) {
    const int py = gy + dy;
    const int my = (py >= 0 && py < H) ? clamp_multiplicity(py, gy, H, R) : 0;
    for (int dx = -R; dx <= R; ++dx) {
      const int px = gx + dx;
      const int mx = (px >= 0 && px < W) ? clamp_multiplicity(px, gx, W, R) : 0;
      const int o = c + dy * G::PITCH + dx;
      const float wq = (float)(my * mx) * sw[o];  // sw is already 0 outside the image
      tap<TYPE, false, true>(ec, tc, se[o], st[o], wq, eps, acc, ga, gb);
    }
  }
  return CENSUS ? fmaf(wc, gb, ga) : ga;
}

// slot of an own pixel in the per-tile border-line table, or -1
__device__ __forceinline__ int border_slot(int ly, int lx, int gy, int gx, int H, int W) {
  if (gx == 0) return ly;
  if (gx == W - 1) return TH + ly;
  if (gy == 0) return 2 * TH + lx;
  if (gy == H - 1) return 2 * TH + TW + lx;
  return -1;
}

// Recompute gacc of every border-line pixel of this tile into fix[NFIX] (one pixel per thread).
template <int TYPE, int R>
__device__ __forceinline__ void border_fixup(const float* __restrict__ se, const float* __restrict__ st,
                                             const float* __restrict__ sw, float* __restrict__ fix, int x0, int y0,
                                             int H, int W, float eps, int tid) {
  if (tid < NFIX) {
    int ly, lx;
    if (tid < TH) { ly = tid; lx = 0 - x0; }
    else if (tid < 2 * TH) { ly = tid - TH; lx = W - 1 - x0; }
    else if (tid < 2 * TH + TW) { ly = 0 - y0; lx = tid - 2 * TH; }
    else { ly = H - 1 - y0; lx = tid - 2 * TH - TW; }
    const int gy = y0 + ly, gx = x0 + lx;
    if (ly >= 0 && ly < TH && lx >= 0 && lx < TW && gy < H && gx < W)
      fix[tid] = border_pixel_gacc<TYPE, R>(se, st, sw, ly, lx, gy, gx, H, W, eps);
  }
}

}  // namespace dis
