// Multi-scale fused pattern loss: S = 2*NPAIR disparity maps of the SAME frames against one LCN image.
//
// reference: the photometric loop of single_frame_worker.Worker.loss_forward
// (model/single_frame_worker.py:108-115) calls RectifiedPatternSimilarityLoss once per output scale with
// the same (im, std).  Everything that depends only on the image -- the target-side soft-census term
// g(t(q) - t(p)), the sigma weights, the image/std tile loads -- is scale independent, so one launch
// evaluates it once and reuses it for all scales (1.25 instead of 2 rsqrt per tap-scale).
//
// Blackwell specifics: the estimate planes of two scales are stored INTERLEAVED (float2) in shared
// memory and processed with packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2, sm_100+): one issue slot
// per two scales.  Scalar operands (eps, the target term, the weight) are broadcast by the instruction
// itself.  B200 issues 128 lane-instructions and 16 MUFU per clock per SM; with packing the issue slots,
// the MUFU pipe and the FP32 lanes are balanced at ~0.32 clk per tap (4 scales) per SM.
//
// Only the census types take this path (mse / sad have no scale-independent arithmetic worth sharing).
// CTA = 256 threads = 32 (x) x 8 (y); tile 64 x 32 pixels; a thread owns 2 adjacent pixels of rows ty + 8g.
#pragma once
#include "window.cuh"

namespace dis {

constexpr int MTW = 64, MTH = 32;             // tile (halo re-computation 1.41x at R = 4)
constexpr int MROWS_PER_PASS = 8;             // 256 threads = 32 (x, 2 px each) x 8 (y); 4 passes cover the tile
constexpr int MFIX = 2 * MTH + 2 * MTW;       // border-line candidates per tile
constexpr int MAX_SCALES = 4;
#ifndef DIS_FILL_UNROLL
#define DIS_FILL_UNROLL 3
#endif
constexpr int kFillUnroll = DIS_FILL_UNROLL;

template <int R>
struct MultiGeom {
  static constexpr int NW = 2 + 2 * R;                 // window values per thread-row (always even)
  static constexpr int PITCH = (MTW - 2) + NW;         // == MTW + 2R, even
  static constexpr int ROWS = MTH + 2 * R;
  static constexpr int COLS = MTW + 2 * R;
  static constexpr int PLANE = ROWS * PITCH;
};

struct PatternMultiArgs {
  const float* disp[MAX_SCALES];
  float* grad_num[MAX_SCALES];   // all NULL => forward only
  const float* im;
  const float* std_in;
  const float* pattern;
  const float* grad_scale;       // optional device float[S]: the stored gradient of scale s is multiplied by it
  float* partials;               // [S][num_blocks][2] = (num_s, den)
  int N, H, W;
  int num_blocks;                // CTAs of the whole call (stride between scales in `partials`)
  int block_offset;              // first CTA index of this launch (batch chunking)
  float eps, inv_k2, inv_w, inv_h;
  int vec_ok;                    // W even and grad pointers 8-byte aligned
};

// For the widest windows (R >= 6, 4 scales) the d proj / d disp plane would push a CTA past half of the SM's shared
// memory and halve the occupancy; there it is re-derived in the epilogue instead (a few % more arithmetic).
template <int R, int NPAIR>
__host__ __device__ constexpr bool multi_recompute_dd() { return R >= 6 && NPAIR == 2; }

template <int R, int NPAIR>
constexpr size_t pattern_multi_smem_bytes() {
  using G = MultiGeom<R>;
  return sizeof(float) * ((size_t)2 * NPAIR * G::PLANE + 2 * G::PLANE +
                          (multi_recompute_dd<R, NPAIR>() ? 0 : 2 * NPAIR * MTH * MTW) + 2 * NPAIR * MFIX + 64) +
         sizeof(WarpRow) * G::ROWS;
}

// exact backward accumulator of one border-line pixel for scale s (same maths as window.cuh border_pixel_gacc)
template <int TYPE, int R, int NPAIR>
__device__ float multi_border_gacc(const float2* __restrict__ se, const float* __restrict__ st,
                                   const float* __restrict__ sw, int s, int ly, int lx, int gy, int gx, int H, int W,
                                   float eps) {
  using G = MultiGeom<R>;
  const float2* plane = se + (s >> 1) * G::PLANE;
  const int c = (ly + R) * G::PITCH + lx + R;
  auto ev = [&](int o) { const float2 v = plane[o]; return (s & 1) ? v.y : v.x; };
  const float ec = ev(c), tc = st[c], wc = sw[c];
  float acc = 0.0f, ga = 0.0f, gb = 0.0f;
  for (int dy = -R; dy <= R; ++dy) {
    const int py = gy + dy;
    const int my = (py >= 0 && py < H) ? clamp_multiplicity(py, gy, H, R) : 0;
    for (int dx = -R; dx <= R; ++dx) {
      const int px = gx + dx;
      const int mx = (px >= 0 && px < W) ? clamp_multiplicity(px, gx, W, R) : 0;
      const int o = c + dy * G::PITCH + dx;
      tap<TYPE, false, true>(ec, tc, ev(o), st[o], (float)(my * mx) * sw[o], eps, acc, ga, gb);
    }
  }
  return fmaf(wc, gb, ga);
}

template <int TYPE, int R, int NPAIR, bool GRAD>
__global__ void __launch_bounds__(256, 2) pattern_multi_kernel(PatternMultiArgs a) {
  static_assert(TYPE == CENSUS_MSE || TYPE == CENSUS_SAD, "multi-scale path is for the census types");
  using G = MultiGeom<R>;
  constexpr int S = 2 * NPAIR;
  constexpr bool RDD = multi_recompute_dd<R, NPAIR>();
  extern __shared__ __align__(16) float smem[];
  float2* se = reinterpret_cast<float2*>(smem);              // [NPAIR][ROWS][PITCH] float2 = (scale 2p, scale 2p+1)
  float* st = smem + 2 * NPAIR * G::PLANE;                   // LCN image, replicate-clamped
  float* sw = st + G::PLANE;                                 // sigma (or 1), zero outside the image
  float* sdd = sw + G::PLANE;                                // [S][MTH][MTW] d proj / d disp of own pixels
  float* fix = sdd + (RDD ? 0 : S * MTH * MTW);              // [S][MFIX]
  float* red = fix + S * MFIX;                               // block-reduction scratch
  WarpRow* srow = reinterpret_cast<WarpRow*>(red + 64);      // y half of the pattern warp, one entry per tile row
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int x0 = blockIdx.x * MTW, y0 = blockIdx.y * MTH, n = blockIdx.z;
  const size_t hw = (size_t)a.H * a.W;
  const size_t fo = (size_t)n * hw;
  if (tid < G::ROWS) srow[tid] = warp_row_setup(clampi(y0 - R + tid, 0, a.H - 1), a.H, a.W, a.inv_h);
  __syncthreads();
  const float* dptr[S];
#pragma unroll
  for (int s = 0; s < S; ++s) dptr[s] = a.disp[s] + fo;
  const float* imp = a.im + fo;
  const float* sdp = a.std_in ? a.std_in + fo : nullptr;

  // ---- stage the tile: S pattern warps per position (shared y half), image and sigma once ---------------
  // (unrolled so that the loads of several positions are in flight together: the phase is latency-bound)
#pragma unroll kFillUnroll
  for (int idx = tid; idx < G::ROWS * G::COLS; idx += 256) {
    const int j = idx / G::COLS, i = idx - j * G::COLS;
    const int gy = y0 - R + j, gx = x0 - R + i;
    const bool inside = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
    const int cy = clampi(gy, 0, a.H - 1), cx = clampi(gx, 0, a.W - 1);
    const int g = cy * a.W + cx;
    float dv[S];
#pragma unroll
    for (int s = 0; s < S; ++s) dv[s] = __ldg(dptr[s] + g);            // S independent loads in flight
    const float tv = __ldg(imp + g);
    const float wv = inside ? (sdp ? __ldg(sdp + g) : 1.0f) : 0.0f;
    const WarpRow row = srow[j];
    const bool own = GRAD && !RDD && j >= R && j < R + MTH && i >= R && i < R + MTW;
    float ev[S], dd[S];
#pragma unroll
    for (int s = 0; s < S; ++s) ev[s] = warp_col_sample(a.pattern, row, dv[s], cx, a.W, a.inv_w, own ? &dd[s] : nullptr);
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) se[p * G::PLANE + j * G::PITCH + i] = make_float2(ev[2 * p], ev[2 * p + 1]);
    st[j * G::PITCH + i] = tv;
    sw[j * G::PITCH + i] = wv;
    if (own) {
#pragma unroll
      for (int s = 0; s < S; ++s) sdd[(s * MTH + (j - R)) * MTW + (i - R)] = dd[s];
    }
  }
  __syncthreads();

  const bool edge_tile = (x0 == 0) || (y0 == 0) || (x0 + MTW >= a.W) || (y0 + MTH >= a.H);
  // ---- border-line pixels: exact clamp multiplicities (one pixel-scale per thread) ---------------------
  if (GRAD && edge_tile) {
    for (int item = tid; item < S * MFIX; item += 256) {
      const int s = item / MFIX, c = item - s * MFIX;
      int ly, lx;
      if (c < MTH) { ly = c; lx = 0 - x0; }
      else if (c < 2 * MTH) { ly = c - MTH; lx = a.W - 1 - x0; }
      else if (c < 2 * MTH + MTW) { ly = 0 - y0; lx = c - 2 * MTH; }
      else { ly = a.H - 1 - y0; lx = c - 2 * MTH - MTW; }
      const int gy = y0 + ly, gx = x0 + lx;
      if (ly >= 0 && ly < MTH && lx >= 0 && lx < MTW && gy < a.H && gx < a.W)
        fix[item] = multi_border_gacc<TYPE, R, NPAIR>(se, st, sw, s, ly, lx, gy, gx, a.H, a.W, a.eps);
    }
    __syncthreads();
  }

  const float fs = fwd_scale<TYPE>() * a.inv_k2;
  const float gs = -0.5f * a.eps * a.inv_k2;
  float gss[S];                  // census chain-rule constant x the caller's per-scale factor (weight / sum of sigma)
#pragma unroll
  for (int s = 0; s < S; ++s) gss[s] = (GRAD && a.grad_scale) ? gs * __ldg(a.grad_scale + s) : gs;
  const u64 eps2 = bc2(a.eps);
  float num[S], den = 0.0f;
#pragma unroll
  for (int s = 0; s < S; ++s) num[s] = 0.0f;

  // ---- window loop: thread = pixels (2tx, 2tx+1) of rows ty, ty+8, ty+16, ty+24 ------------------------
#pragma unroll 1
  for (int pass = 0; pass < MTH / MROWS_PER_PASS; ++pass) {
    const int ly = ty + pass * MROWS_PER_PASS;
    u64 ec[NPAIR][2];
    float acc[NPAIR][2][2];
    u64 ga[NPAIR][2], gb[NPAIR][2];
    float tc[2], wc[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int o = (ly + R) * G::PITCH + 2 * tx + i + R;
      tc[i] = st[o];
      wc[i] = sw[o];
#pragma unroll
      for (int p = 0; p < NPAIR; ++p) {
        const float2 v = se[p * G::PLANE + o];
        ec[p][i] = pk2(v.x, v.y);
        acc[p][i][0] = acc[p][i][1] = 0.0f;
        ga[p][i] = gb[p][i] = pk2(0.0f, 0.0f);
      }
    }
#pragma unroll 1
    for (int j = 0; j <= 2 * R; ++j) {
      const int base = (ly + j) * G::PITCH + 2 * tx;
      u64 er[NPAIR][G::NW];
      float tr[G::NW], wr[G::NW];
#pragma unroll
      for (int v = 0; v < G::NW / 2; ++v) {
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) {
          const float4 q = *reinterpret_cast<const float4*>(se + p * G::PLANE + base + 2 * v);
          er[p][2 * v] = pk2(q.x, q.y);
          er[p][2 * v + 1] = pk2(q.z, q.w);
        }
        const float2 t2 = *reinterpret_cast<const float2*>(st + base + 2 * v);
        tr[2 * v] = t2.x; tr[2 * v + 1] = t2.y;
        const float2 w2 = *reinterpret_cast<const float2*>(sw + base + 2 * v);
        wr[2 * v] = w2.x; wr[2 * v + 1] = w2.y;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int dx = 0; dx <= 2 * R; ++dx) {
          const int k = i + dx;
          // target side, shared by all scales: gt = dt * rt (rounded) and its exact rounding residual
          const float dt = tr[k] - tc[i];
          const float rt = rsqrt_fast(fmaf(dt, dt, a.eps));
          const float gt = __fmul_rn(dt, rt);
          const float rest = __fmaf_rn(dt, rt, -gt);
#pragma unroll
          for (int p = 0; p < NPAIR; ++p) {
            const u64 de = sub2(er[p][k], ec[p][i]);
            const u64 xe = fma2(de, de, eps2);
            float x0f, x1f;
            upk2(xe, x0f, x1f);
            const float r0 = rsqrt_fast(x0f), r1 = rsqrt_fast(x1f);
            const u64 re = pk2(r0, r1);
            // diff2 = (de*re - gt) - (dt*rt - gt): both products enter through an exact FMA residual, so
            // e == t gives exactly 0 (the reference's |.| has subgradient 0 there) and e != t costs one rounding
            const u64 diff = sub2(fma2(de, re, bc2(-gt)), bc2(rest));
            float d0, d1;
            upk2(diff, d0, d1);
            if (TYPE == CENSUS_SAD) {
              acc[p][i][0] += fabsf(d0);
              acc[p][i][1] += fabsf(d1);
            } else {
              acc[p][i][0] = fmaf(d0, d0, acc[p][i][0]);
              acc[p][i][1] = fmaf(d1, d1, acc[p][i][1]);
            }
            if (GRAD) {
              const u64 r3 = mul2(mul2(re, re), re);
              u64 u;
              if (TYPE == CENSUS_SAD) {
                float q0, q1;
                upk2(r3, q0, q1);
                u = pk2(signed_mag(q0, d0), signed_mag(q1, d1));
              } else {
                u = mul2(diff, r3);
              }
              ga[p][i] = fma2(u, bc2(wr[k]), ga[p][i]);
              gb[p][i] = add2(gb[p][i], u);
            }
          }
        }
    }

    // ---- per-pass epilogue ---------------------------------------------------------------------------
    const int gy = y0 + ly;
    float gout[S][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int gx = x0 + 2 * tx + i;
      const bool valid = gy < a.H && gx < a.W;
      if (valid) den += wc[i];
      float ddv[S];
      if (GRAD && RDD) {
        const WarpRow row = srow[ly + R];
        const int cy = min(gy, a.H - 1), cx = min(gx, a.W - 1);
#pragma unroll
        for (int s = 0; s < S; ++s) warp_col_sample(a.pattern, row, __ldg(dptr[s] + cy * a.W + cx), cx, a.W, a.inv_w, &ddv[s]);
      }
      int slot = -1;
      if (GRAD && edge_tile && valid) {
        if (gx == 0) slot = ly;
        else if (gx == a.W - 1) slot = MTH + ly;
        else if (gy == 0) slot = 2 * MTH + 2 * tx + i;
        else if (gy == a.H - 1) slot = 2 * MTH + MTW + 2 * tx + i;
      }
#pragma unroll
      for (int p = 0; p < NPAIR; ++p) {
        float g0 = 0.f, g1 = 0.f;
        if (GRAD) {
          float a0, a1, b0, b1;
          upk2(ga[p][i], a0, a1);
          upk2(gb[p][i], b0, b1);
          g0 = fmaf(wc[i], b0, a0);
          g1 = fmaf(wc[i], b1, a1);
          if (slot >= 0) { g0 = fix[(2 * p) * MFIX + slot]; g1 = fix[(2 * p + 1) * MFIX + slot]; }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = 2 * p + h;
          if (valid) num[s] = fmaf(wc[i], acc[p][i][h] * fs, num[s]);
          if (GRAD) gout[s][i] = (h ? g1 : g0) * gss[s] * (RDD ? ddv[s] : sdd[(s * MTH + ly) * MTW + 2 * tx + i]);
        }
      }
    }
    if (GRAD && gy < a.H) {
      const int gx = x0 + 2 * tx;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float* p = a.grad_num[s] + fo + (size_t)gy * a.W + gx;
        if (a.vec_ok && gx + 1 < a.W) __stcs(reinterpret_cast<float2*>(p), make_float2(gout[s][0], gout[s][1]));
        else {
          if (gx < a.W) p[0] = gout[s][0];
          if (gx + 1 < a.W) p[1] = gout[s][1];
        }
      }
    }
  }

  // block reduction of S numerators + 1 denominator (fixed order)
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int s = 0; s < S; ++s) num[s] = warp_sum(num[s]);
  den = warp_sum(den);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) red[wid * (S + 1) + s] = num[s];
    red[wid * (S + 1) + S] = den;
  }
  __syncthreads();
  if (tid <= S) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w * (S + 1) + tid];
    red[56 + tid] = v;
  }
  __syncthreads();
  if (tid < S) {
    const size_t b = (size_t)a.block_offset + ((size_t)n * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    a.partials[((size_t)tid * a.num_blocks + b) * 2] = red[56 + tid];
    a.partials[((size_t)tid * a.num_blocks + b) * 2 + 1] = red[56 + S];
  }
}

template <int R> int launch_pattern_multi(const PatternMultiArgs& a, int S, int type, cudaStream_t s);

}  // namespace dis
