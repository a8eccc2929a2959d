// mse / sad pattern loss without a per-pixel loss map: a point-wise kernel behind ONE box filter of the weights.
//
// reference: RectifiedPatternSimilarityLoss.tforward (model/networks.py:354-377) with
// photometric_loss_pytorch types 'mse' / 'sad' (model/ext_functions.py:156-168):
//   out(p) = (1/k^2) sum_o s(clamp(p + o)),  s(q) = (e(q) - t(q))^2 or |e(q) - t(q)|,  val = sum_p w(p) out(p) / sum_p w(p)
// The window sum is linear in the point-wise term, so the numerator can be summed over the SOURCE pixel instead:
//   sum_p w(p) out(p) = sum_q s(q) M(q),   M(q) = (1/k^2) sum_{(p, o): clamp(p + o) = q} w(p)
// and M is also all the gradient needs:  d num / d e(q) = s'(q) M(q).  M depends on the weights (sigma of the LCN, or 1)
// only: it is computed once per batch of frames (two separable passes, padding multiplicity included) and shared by
// every disparity scale; the loss itself is then point-wise: per scale one coalesced read of the disparity and one
// write of the gradient, the image and M read once for all scales.  (The tile kernel in box_kernels.cuh box-filters s
// AND w per scale; it remains the path when the caller asks for the per-pixel map.)
#include "common.cuh"
#include "window.cuh"

namespace dis {
namespace {

constexpr int PL_MAX_S = 4;          // disparity scales per launch
constexpr int PL_MAX_T = 256;        // threads = columns of a strip
constexpr int PL_ROWS = 4;           // rows per CTA (one pixel per thread and row, all loads issued up front)
constexpr int BW_RUN = 16;           // rows one thread slides down in the vertical pass

// 1-D operator A(w)(q) = sum over padded positions c that clamp onto q of sum_{|d| <= R} wz(c + d), wz = w inside
// [0, n) and 0 outside.  Vertical pass: one thread per (column, run of BW_RUN rows), fp32 sliding sum inside the run.
// grid (columns / 256, runs, frames)
template <int R>
__global__ void __launch_bounds__(256) box_weight_v_kernel(const float* __restrict__ w, float* __restrict__ V, int H, int W) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= W) return;
  const size_t fo = (size_t)blockIdx.z * H * W + x;
  const float* wn = w ? w + fo : nullptr;
  float* vn = V + fo;
  auto wz = [&](int r) -> float { return (r >= 0 && r < H) ? (wn ? __ldg(wn + (size_t)r * W) : 1.0f) : 0.0f; };
  auto window = [&](int c) -> float {   // sum_{|d| <= R} wz(c + d)
    float s = 0.f;
#pragma unroll
    for (int d = -R; d <= R; ++d) s += wz(c + d);
    return s;
  };
  const int y0 = blockIdx.y * BW_RUN;
  float s = window(y0);
  // entering / leaving rows of the whole run first (independent loads), then the sliding sums
  float in[BW_RUN], out[BW_RUN];
#pragma unroll
  for (int k = 0; k < BW_RUN; ++k) { in[k] = wz(y0 + k + 1 + R); out[k] = wz(y0 + k - R); }
#pragma unroll
  for (int k = 0; k < BW_RUN; ++k) {
    const int y = y0 + k;
    if (y < H) {
      float v = s;
      if (y == 0) for (int c = -R; c < 0; ++c) v += window(c);
      if (y == H - 1) for (int c = H; c < H + R; ++c) v += window(c);
      vn[(size_t)y * W] = v;
    }
    s += in[k] - out[k];
  }
}

// horizontal pass over V, scaled by 1 / k^2: one thread per 4 consecutive pixels of a row (aligned 128-bit loads of the
// 4 + 2R values in reach, sliding sum, one 128-bit store), BW_RUN rows per CTA.  grid (quads / 128, rows of all frames / BW_RUN)
template <int R>
__global__ void __launch_bounds__(128) box_weight_h_kernel(const float* __restrict__ V, float* __restrict__ M, int W, float inv_k2,
                                                           size_t rows, int vec_ok) {
  constexpr int PAD = 4 * ((R + 3) / 4), NV = 1 + 2 * ((R + 3) / 4);   // floats before x0 / float4 loads per row
  const int x0 = 4 * (blockIdx.x * 128 + threadIdx.x);
  if (x0 >= W) return;
  const size_t r0 = (size_t)blockIdx.y * BW_RUN;
  // no padded column in reach and every load inside the row: plain window, no bounds checks
  const bool interior = vec_ok && x0 - R >= 1 && x0 + 3 + R <= W - 2 && x0 - PAD >= 0 && x0 + 4 + PAD <= W;
#pragma unroll 4
  for (int k = 0; k < BW_RUN; ++k) {
    if (r0 + k >= rows) break;
    const float* row = V + (r0 + k) * W;
    float* mrow = M + (r0 + k) * W;
    if (interior) {
      float v[4 * NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(row + x0 - PAD) + q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
      float o[4], s = 0.f;
#pragma unroll
      for (int d = 0; d <= 2 * R; ++d) s += v[PAD - R + d];
      o[0] = s;
#pragma unroll
      for (int c = 1; c < 4; ++c) { s += v[PAD + R + c] - v[PAD - R + c - 1]; o[c] = s; }
      __stcs(reinterpret_cast<float4*>(mrow + x0), make_float4(o[0] * inv_k2, o[1] * inv_k2, o[2] * inv_k2, o[3] * inv_k2));
      continue;
    }
    auto vz = [&](int c) -> float { return (c >= 0 && c < W) ? __ldg(row + c) : 0.0f; };
    auto window = [&](int c) -> float {
      float s = 0.f;
      for (int d = -R; d <= R; ++d) s += vz(c + d);
      return s;
    };
    for (int x = x0; x < min(x0 + 4, W); ++x) {
      float v = window(x);
      if (x == 0) for (int c = -R; c < 0; ++c) v += window(c);
      if (x == W - 1) for (int c = W; c < W + R; ++c) v += window(c);
      mrow[x] = v * inv_k2;
    }
  }
}

struct PointLossArgs {
  const float* disp[PL_MAX_S];
  float* grad[PL_MAX_S];
  float* proj[PL_MAX_S];
  const float* im; const float* std_in; const float* wbox; const float* pattern;   // wbox: M
  const float* grad_scale;   // optional, S device floats
  float* partials;           // [S][num_blocks] (num_s, den) pairs
  int H, W, strip;
  int num_blocks;            // partial slots per scale (all frames)
  int block_offset;          // slot of this launch's first CTA (frame chunking)
  float inv_w, inv_h;
};

// grid (strips, row groups, frames): a CTA owns PL_ROWS rows of a strip of columns, one pixel per thread and row.  The
// kernel is latency-bound (coalesced loads -> dependent pattern gathers -> store): all loads of the thread's pixels are
// issued before the first gather; the y half of the warp depends on the row only and is shared by every scale.
template <int TYPE, int S, bool GRAD>
__global__ void __launch_bounds__(PL_MAX_T) point_pattern_loss_kernel(const __grid_constant__ PointLossArgs a) {
  __shared__ float red[2 * (PL_MAX_T / 32)];
  const int tid = threadIdx.x;
  if (tid < 2 * (PL_MAX_T / 32)) red[tid] = 0.f;   // block_sum2 adds PL_MAX_T / 32 warp slots; the CTA may have fewer warps
  const int w = blockIdx.x * a.strip + tid;
  const bool col_ok = tid < a.strip && w < a.W;
  const int wc = min(w, a.W - 1);
  const int h0 = blockIdx.y * PL_ROWS;
  const size_t fo = (size_t)blockIdx.z * a.H * a.W;
  float gs[S];
#pragma unroll
  for (int s = 0; s < S; ++s) gs[s] = (GRAD && a.grad_scale) ? __ldg(a.grad_scale + s) : 1.0f;
  float num[S], den = 0.f;
#pragma unroll
  for (int s = 0; s < S; ++s) num[s] = 0.f;
  bool ok[PL_ROWS];
  float tv[PL_ROWS], mv[PL_ROWS], dv[PL_ROWS][S];
#pragma unroll
  for (int k = 0; k < PL_ROWS; ++k) {
    ok[k] = col_ok && h0 + k < a.H;
    const size_t i = fo + (size_t)min(h0 + k, a.H - 1) * a.W + wc;
    tv[k] = ld_stream(a.im + i);
    mv[k] = ok[k] ? ld_stream(a.wbox + i) : 0.0f;     // a pixel outside the strip / frame contributes nothing
    if (ok[k]) den += a.std_in ? ld_stream(a.std_in + i) : 1.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) dv[k][s] = ld_stream(a.disp[s] + i);
  }
#pragma unroll
  for (int k = 0; k < PL_ROWS; ++k) {
    const int h = min(h0 + k, a.H - 1);
    const size_t i = fo + (size_t)h * a.W + wc;
    const WarpRow row = warp_row_setup(h, a.H, a.W, a.inv_h);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float dd = 0.f;
      const float e = warp_col_sample_clamped(a.pattern + row.off0, a.pattern + row.off1, row.wy0, row.wy1, dv[k][s], wc, a.W,
                                              a.inv_w, GRAD ? &dd : nullptr);
      const float d = e - tv[k];
      num[s] = fmaf(TYPE == MSE ? d * d : fabsf(d), mv[k], num[s]);
      if (ok[k]) {
        if (a.proj[s]) __stcs(a.proj[s] + i, e);
        if (GRAD) __stcs(a.grad[s] + i, ((TYPE == MSE ? 2.0f * d : sign0(d)) * dd) * (mv[k] * gs[s]));
      }
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s) {
    float n2 = num[s], d2 = den;
    __syncthreads();
    block_sum2<PL_MAX_T>(n2, d2, red);
    if (tid == 0) {
      const size_t b = (size_t)s * a.num_blocks + a.block_offset + ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      a.partials[2 * b] = n2;
      a.partials[2 * b + 1] = d2;
    }
  }
}

template <int TYPE, int S>
int launch_point(const PointLossArgs& a, bool grad, dim3 grid, int threads, cudaStream_t s) {
  if (grad) point_pattern_loss_kernel<TYPE, S, true><<<grid, threads, 0, s>>>(a);
  else point_pattern_loss_kernel<TYPE, S, false><<<grid, threads, 0, s>>>(a);
  return check_launch();
}

constexpr int PL_MAX_FRAMES = 32768;   // frames per launch (grid.z limit)
struct PointPlan { int nstrips, strip, threads, groups; };
inline PointPlan point_plan(int H, int W) {
  PointPlan p;
  p.nstrips = (W + PL_MAX_T - 1) / PL_MAX_T;
  p.strip = (W + p.nstrips - 1) / p.nstrips;          // columns per strip: even split (432 -> 2 x 216)
  p.threads = ((p.strip + 31) / 32) * 32;
  p.groups = (H + PL_ROWS - 1) / PL_ROWS;
  return p;
}

}  // namespace

int point_loss_num_partials(int N, int H, int W) {
  const PointPlan p = point_plan(H, W);
  return N * p.nstrips * p.groups;
}

// workspace: N*H*W floats (the vertical pass); wbox: N*H*W floats
int box_weight(const float* std_in, float* workspace, float* wbox, int N, int H, int W, int block_size, cudaStream_t s) {
  const int R = block_size / 2, runs = (H + BW_RUN - 1) / BW_RUN;
  const size_t hw = (size_t)H * W;
  const float inv_k2 = 1.0f / (float)(block_size * block_size);
  for (int n0 = 0; n0 < N; n0 += PL_MAX_FRAMES) {
    const int nb = N - n0 < PL_MAX_FRAMES ? N - n0 : PL_MAX_FRAMES;
    const dim3 grid((W + 255) / 256, runs, nb);
    const float* w = std_in ? std_in + n0 * hw : nullptr;
    float* v = workspace + n0 * hw;
    switch (R) {
#define DIS_BW_CASE(R_) case R_: box_weight_v_kernel<R_><<<grid, 256, 0, s>>>(w, v, H, W); break;
      DIS_BW_CASE(0) DIS_BW_CASE(1) DIS_BW_CASE(2) DIS_BW_CASE(3) DIS_BW_CASE(4) DIS_BW_CASE(5) DIS_BW_CASE(6) DIS_BW_CASE(7)
#undef DIS_BW_CASE
      default: return DIS_ERR_UNSUPPORTED_BLOCK_SIZE;
    }
    if (int rc = check_launch()) return rc;
  }
  const size_t rows = (size_t)N * H, per_launch = (size_t)65535 * BW_RUN;
  for (size_t r0 = 0; r0 < rows; r0 += per_launch) {
    const size_t nr = rows - r0 < per_launch ? rows - r0 : per_launch;
    const dim3 grid(((W + 3) / 4 + 127) / 128, (unsigned)((nr + BW_RUN - 1) / BW_RUN));
    const float* v = workspace + r0 * W;
    float* m = wbox + r0 * W;
    const int vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(m)) & 15) == 0;
    switch (R) {
#define DIS_BW_CASE(R_) case R_: box_weight_h_kernel<R_><<<grid, 128, 0, s>>>(v, m, W, inv_k2, nr, vec_ok); break;
      DIS_BW_CASE(0) DIS_BW_CASE(1) DIS_BW_CASE(2) DIS_BW_CASE(3) DIS_BW_CASE(4) DIS_BW_CASE(5) DIS_BW_CASE(6) DIS_BW_CASE(7)
#undef DIS_BW_CASE
    }
    if (int rc = check_launch()) return rc;
  }
  return DIS_OK;
}

int point_pattern_loss(const float* const* disps, int S, const float* im, const float* std_in, const float* wbox,
                       const float* pattern, float* const* projs, float* const* grads, const float* grad_scale,
                       float* partials, int N, int H, int W, int type, cudaStream_t s) {
  const size_t hw = (size_t)H * W;
  const PointPlan p = point_plan(H, W);
  const int per_frame = p.nstrips * p.groups;
  for (int n0 = 0; n0 < N; n0 += PL_MAX_FRAMES) {
    const int nb = N - n0 < PL_MAX_FRAMES ? N - n0 : PL_MAX_FRAMES;
    PointLossArgs a{};
    bool grad = false;
    for (int i = 0; i < S; ++i) {
      a.disp[i] = disps[i] + n0 * hw;
      a.grad[i] = (grads && grads[i]) ? grads[i] + n0 * hw : nullptr;
      a.proj[i] = (projs && projs[i]) ? projs[i] + n0 * hw : nullptr;
      grad |= a.grad[i] != nullptr;
    }
    for (int i = 0; grad && i < S; ++i)
      if (!a.grad[i]) return DIS_ERR_NULL_POINTER;   // gradients for all scales or for none
    a.im = im + n0 * hw; a.std_in = std_in ? std_in + n0 * hw : nullptr; a.wbox = wbox + n0 * hw; a.pattern = pattern;
    a.grad_scale = grad_scale; a.partials = partials;
    a.H = H; a.W = W; a.strip = p.strip;
    a.num_blocks = N * per_frame; a.block_offset = n0 * per_frame;
    a.inv_w = 1.0f / (float)(W - 1); a.inv_h = 1.0f / (float)(H - 1);
    const dim3 grid(p.nstrips, p.groups, nb);
    int rc = DIS_ERR_UNSUPPORTED_COMBINATION;
#define DIS_PL_CASE(T_, S_) if (type == T_ && S == S_) rc = launch_point<T_, S_>(a, grad, grid, p.threads, s);
    DIS_PL_CASE(MSE, 1) DIS_PL_CASE(MSE, 2) DIS_PL_CASE(MSE, 3) DIS_PL_CASE(MSE, 4)
    DIS_PL_CASE(SAD, 1) DIS_PL_CASE(SAD, 2) DIS_PL_CASE(SAD, 3) DIS_PL_CASE(SAD, 4)
#undef DIS_PL_CASE
    if (rc) return rc;
  }
  return DIS_OK;
}

}  // namespace dis
