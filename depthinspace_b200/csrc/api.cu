// extern "C" surface of libdis_b200.so (declared in include/dis_b200.h): argument validation, batch
// chunking (grid.z <= 65535) and dispatch to the per-radius kernel instantiations.  No global mutable
// state: the only static is a thread_local error string.
#include <string.h>
#include "photometric_kernels.cuh"
#include "pattern_multi.cuh"
#include "pattern_march.cuh"
#include <stdlib.h>

namespace dis {

static thread_local char g_last_error[256] = "";
void set_last_cuda_error(cudaError_t e) {
  strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}

// defined in lcn.cu / misc.cu / smooth.cu / flow_warp.cu
int lcn_forward(const float*, float*, float*, int, int, int, int, float, int, cudaStream_t);
int lcn_backward(const float*, const float*, const float*, const float*, const float*, float*, float*, int, int, int,
                 int, float, cudaStream_t);
int lcn_forward_ex(const float*, float*, float*, float*, int, int, int, int, float, int, int, int, size_t, cudaStream_t);
int pattern_warp_forward(const float*, const float*, float*, float*, int32_t*, int32_t*, int, int, int, cudaStream_t);
int reduce_pairs(const float*, int, int, float*, cudaStream_t);
int scale_by_device_scalar(const float*, float*, size_t, const float*, const float*, cudaStream_t);
int mul(const float*, const float*, float*, size_t, cudaStream_t);
int l1_num_partials(size_t);
int l1_forward(const float*, const float*, float*, float*, size_t, cudaStream_t);
int masked_l1_forward(const float*, const float*, const float*, float, float*, float*, size_t, cudaStream_t);
int smooth_loss_num_partials(int, int, int);
int point_loss_num_partials(int, int, int);
int box_weight(const float*, float*, float*, int, int, int, int, cudaStream_t);
int point_pattern_loss(const float* const*, int, const float*, const float*, const float*, const float*, float* const*,
                       float* const*, const float*, float*, int, int, int, int, cudaStream_t);
int smooth_loss_forward(const float*, const float*, float*, float*, int, int, int, float, int, cudaStream_t);
int sobel_forward(const float*, float*, int, int, int, int, cudaStream_t);
int sobel_backward(const float*, float*, int, int, int, int, cudaStream_t);
int flow_warp_forward(const float*, const float*, float*, float*, int32_t*, int32_t*, int, int, int, int, cudaStream_t);
int flow_warp_backward(const float*, const float*, const float*, float*, float*, int, int, int, int, cudaStream_t);
int flow_warp_gather_forward(const float*, const float* const*, float*, int, int, int, int, int, int, cudaStream_t);
int flow_warp_gather_backward(const float* const*, const float*, float*, int, int, int, int, int, int, cudaStream_t);
int flow_warp_gather_all_forward(const float*, const float* const*, float*, int, int, int, int, int, cudaStream_t);
int flow_warp_gather_all_backward(const float* const*, const float*, float*, int, int, int, int, int, cudaStream_t);

int flow_consistency_blocks_per_frame(int, int);
int flow_consistency_forward(const float*, const float*, const float*, const float*, const float*, const float*,
                             const float*, const float*, const float*, const float*, int, const float*, const float*,
                             const float*, float, float, float*, float*, float*, float*, float*, int, int, int,
                             cudaStream_t);
int combine2(const float*, const float*, float*, size_t, const float*, const float*, const float*, float, cudaStream_t);
int geometric_grad_combine(const float* const*, const int*, int, const float*, const float*, float, float*, int, size_t,
                           cudaStream_t);

int resize_bilinear_forward(const float* const*, float* const*, int, int, int, int, int, int, int, int, cudaStream_t);
int resize_bilinear_backward(const float*, float*, int, int, int, int, int, int, cudaStream_t);

int ext_nn(const float*, const float*, long long*, long, long, int, cudaStream_t);
int ext_crosscheck(const long long*, const long long*, uint8_t*, long, long, cudaStream_t);
int ext_proj_nn(const float*, const float*, const float*, long long*, int, int, int, int, cudaStream_t);
int ext_xcorrvol(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);

int conv3d_out_size(int, int, int);
size_t conv3d_scratch_elems(int, int, int, int);
int conv3d_gather_forward(const float*, const float*, const float*, float*, float*, uint8_t*, float*, int, int, int, int,
                          int, int, int, int, cudaStream_t);
int conv3d_rank(const float*, const float*, float*, uint8_t*, float*, int, int, int, int, int, int, int, cudaStream_t);
int conv3d_gather_features(const float*, const uint8_t*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
int conv3d_gather_backward(const float*, const float*, const uint8_t*, float*, float*, int, int, int, int, int, int, int,
                           int, cudaStream_t);

namespace {

constexpr int MAX_GRID_Z = 65535;

template <typename... P>
bool aligned16(P... ptrs) {
  uintptr_t acc = 0;
  ((acc |= reinterpret_cast<uintptr_t>(ptrs)), ...);
  return (acc & 15) == 0;
}

int check_block(int block_size, int type) {
  if (type < 0 || type > 3) return DIS_ERR_INVALID_LOSS_TYPE;
  if (block_size < 1 || (block_size & 1) == 0 || block_size / 2 > MAX_R) return DIS_ERR_UNSUPPORTED_BLOCK_SIZE;
  return DIS_OK;
}

int dispatch_photometric(int R, const PhotoArgs& a, int type, bool backward, cudaStream_t s) {
  switch (R) {
    case 0: return launch_photometric<0>(a, type, backward, s);
    case 1: return launch_photometric<1>(a, type, backward, s);
    case 2: return launch_photometric<2>(a, type, backward, s);
    case 3: return launch_photometric<3>(a, type, backward, s);
    case 4: return launch_photometric<4>(a, type, backward, s);
    case 5: return launch_photometric<5>(a, type, backward, s);
    case 6: return launch_photometric<6>(a, type, backward, s);
    case 7: return launch_photometric<7>(a, type, backward, s);
  }
  return DIS_ERR_UNSUPPORTED_BLOCK_SIZE;
}

int dispatch_pattern_loss(int R, const PatternLossArgs& a, int type, cudaStream_t s) {
  switch (R) {
    case 0: return launch_pattern_loss<0>(a, type, s);
    case 1: return launch_pattern_loss<1>(a, type, s);
    case 2: return launch_pattern_loss<2>(a, type, s);
    case 3: return launch_pattern_loss<3>(a, type, s);
    case 4: return launch_pattern_loss<4>(a, type, s);
    case 5: return launch_pattern_loss<5>(a, type, s);
    case 6: return launch_pattern_loss<6>(a, type, s);
    case 7: return launch_pattern_loss<7>(a, type, s);
  }
  return DIS_ERR_UNSUPPORTED_BLOCK_SIZE;
}

int dispatch_pattern_multi(int R, const PatternMultiArgs& a, int S, int type, cudaStream_t s) {
  switch (R) {
    case 0: return launch_pattern_multi<0>(a, S, type, s);
    case 1: return launch_pattern_multi<1>(a, S, type, s);
    case 2: return launch_pattern_multi<2>(a, S, type, s);
    case 3: return launch_pattern_multi<3>(a, S, type, s);
    case 4: return launch_pattern_multi<4>(a, S, type, s);
    case 5: return launch_pattern_multi<5>(a, S, type, s);
    case 6: return launch_pattern_multi<6>(a, S, type, s);
    case 7: return launch_pattern_multi<7>(a, S, type, s);
  }
  return DIS_ERR_UNSUPPORTED_BLOCK_SIZE;
}

int dispatch_pattern_march(int R, const PatternMarchArgs& a, const MarchPlan& plan, int S, int type, cudaStream_t s) {
  switch (R) {
    case 1: return launch_pattern_march<1>(a, plan, S, type, s);
    case 2: return launch_pattern_march<2>(a, plan, S, type, s);
    case 3: return launch_pattern_march<3>(a, plan, S, type, s);
    case 4: return launch_pattern_march<4>(a, plan, S, type, s);
    case 5: return launch_pattern_march<5>(a, plan, S, type, s);
    case 6: return launch_pattern_march<6>(a, plan, S, type, s);
    case 7: return launch_pattern_march<7>(a, plan, S, type, s);
  }
  return DIS_ERR_UNSUPPORTED_BLOCK_SIZE;
}

int dispatch_photometric_march(int R, const PatternMarchArgs& a, const MarchPlan& plan, int mode, int type, cudaStream_t s) {
  switch (R) {
    case 1: return launch_photometric_march<1>(a, plan, mode, type, s);
    case 2: return launch_photometric_march<2>(a, plan, mode, type, s);
    case 3: return launch_photometric_march<3>(a, plan, mode, type, s);
    case 4: return launch_photometric_march<4>(a, plan, mode, type, s);
    case 5: return launch_photometric_march<5>(a, plan, mode, type, s);
    case 6: return launch_photometric_march<6>(a, plan, mode, type, s);
    case 7: return launch_photometric_march<7>(a, plan, mode, type, s);
  }
  return DIS_ERR_UNSUPPORTED_BLOCK_SIZE;
}

// Tuning / A-B knobs, read per call (no cached state): DIS_MULTI_IMPL=tile selects the round-1 tile kernel,
// DIS_MARCH_BAND_ROWS / DIS_MARCH_WARPS override the band plan.
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
bool use_march(int R) {
  if (R < 1) return false;
  const char* v = getenv("DIS_MULTI_IMPL");
  return !(v && strcmp(v, "tile") == 0);
}

}  // namespace

// Column bands: the CTA width (warps) that wastes the fewest lanes over the extended width W + 2R; row bands: the
// count that minimises (waves of CTAs) x (rows marched per CTA, including the R halo rows recomputed per band).
MarchPlan march_plan(int N, int H, int W, int R) {
  const int EW = W + 2 * R, EH = H + 2 * R;
  const int RH = ((R + MV - 1) / MV) * MV;
  MarchPlan best{};
  long best_lanes = -1;
  const int force_w = env_int("DIS_MARCH_WARPS", 0);
  for (int nw = 1; nw <= MARCH_MAX_WARPS; ++nw) {
    if (force_w && nw != force_w) continue;
    const int LW = 32 * nw, own = LW - 2 * R;
    if (own < 8) continue;
    const int ncb = EW <= LW ? 1 : 1 + (EW - LW + own - 1) / own;
    const long lanes = (long)ncb * LW;
    if (best_lanes < 0 || lanes <= best_lanes) {   // ties: the wider CTA
      best_lanes = lanes;
      best.nwarps = nw;
      best.ncb = ncb;
    }
  }
  const int slots = 148 * MARCH_CTAS_PER_SM;
  const int min_rows = 2 * RH + 2 * MV;
  int band_rows = env_int("DIS_MARCH_BAND_ROWS", 0);
  if (band_rows <= 0) {
    // cost of a plan = (waves of CTAs) x (rows marched per CTA); among the plans within 5 % of the cheapest take the one
    // with the longest bands (fewer halo rows and prologues: measured 6.82 ms at 130 rows vs 6.89 at 88, 256 frames)
    double best_cost = -1.0;
    for (int pass = 0; pass < 2; ++pass)
      for (int nrb = 1; nrb <= 64; ++nrb) {
        int rows = (EH + nrb - 1) / nrb;
        rows = ((rows + MV - 1) / MV) * MV;
        if (rows < min_rows) break;
        const long ctas = (long)N * best.ncb * nrb;
        const double cost = (double)((ctas + slots - 1) / slots) * (rows + RH);
        if (pass == 0) {
          if (best_cost < 0.0 || cost < best_cost) best_cost = cost;
        } else if (cost <= 1.05 * best_cost) {
          band_rows = rows;
          break;
        }
      }
    if (band_rows <= 0) band_rows = ((EH + MV - 1) / MV) * MV;
  }
  band_rows = ((band_rows + MV - 1) / MV) * MV;
  if (band_rows < min_rows) band_rows = min_rows;
  int nrb = (EH + band_rows - 1) / band_rows;
  // the last band must own image row H-1 together with the virtual rows below it
  while (nrb > 1 && EH - (nrb - 1) * band_rows < R + 1) --nrb;
  best.nrb = nrb;
  best.band_rows = band_rows;
  return best;
}

}  // namespace dis

using namespace dis;

extern "C" {

int dis_abi_version(void) { return 4; }

int dis_ext_nn(const float* in0, const float* in1, int64_t* out, int64_t n0, int64_t n1, int dim, void* stream) {
  if (n0 < 0 || n1 < 0 || (n0 + 255) / 256 > INT_MAX) return DIS_ERR_BAD_SHAPE;
  if (n0 == 0) return DIS_OK;
  if (!in0 || !out || (n1 > 0 && !in1)) return DIS_ERR_NULL_POINTER;
  return ext_nn(in0, in1, reinterpret_cast<long long*>(out), (long)n0, (long)n1, dim, as_stream(stream));
}

int dis_ext_crosscheck(const int64_t* in0, const int64_t* in1, uint8_t* out, int64_t n0, int64_t n1, void* stream) {
  if (n0 < 0 || n1 < 0 || (n0 + 255) / 256 > INT_MAX) return DIS_ERR_BAD_SHAPE;
  if (n0 == 0) return DIS_OK;
  if (!in0 || !out || (n1 > 0 && !in1)) return DIS_ERR_NULL_POINTER;
  return ext_crosscheck(reinterpret_cast<const long long*>(in0), reinterpret_cast<const long long*>(in1), out, (long)n0,
                        (long)n1, as_stream(stream));
}

int dis_ext_proj_nn(const float* xyz0, const float* xyz1, const float* K, int64_t* out, int bs, int H, int W,
                    int patch_size, void* stream) {
  if (bs < 0 || H < 1 || W < 1 || patch_size < 1 || patch_size > 255) return DIS_ERR_BAD_SHAPE;
  if (bs == 0) return DIS_OK;
  if (!xyz0 || !xyz1 || !K || !out) return DIS_ERR_NULL_POINTER;
  return ext_proj_nn(xyz0, xyz1, K, reinterpret_cast<long long*>(out), bs, H, W, patch_size, as_stream(stream));
}

int dis_ext_xcorrvol(const float* in0, const float* in1, float* out, int C, int H, int W, int n_disps, int block_size,
                     void* stream) {
  if (C < 1 || H < 1 || W < 1 || n_disps < 0 || block_size < 1 || block_size > 255) return DIS_ERR_BAD_SHAPE;
  if (n_disps == 0) return DIS_OK;
  if (!in0 || !in1 || !out) return DIS_ERR_NULL_POINTER;
  return ext_xcorrvol(in0, in1, out, C, H, W, n_disps, block_size, as_stream(stream));
}

int dis_resize_bilinear_forward(const float* const* ins, float* const* outs, int count, int N, int C, int H, int W, int oh,
                                int ow, int mode, void* stream) {
  if (!ins || !outs) return DIS_ERR_NULL_POINTER;
  if (count < 0 || N < 0 || C < 1 || H < 1 || W < 1 || oh < 1 || ow < 1 || (long)oh * ow > INT_MAX || (long)H * W > INT_MAX)
    return DIS_ERR_BAD_SHAPE;
  if (mode < 0 || mode > 2 || (mode == 1 && C != 2)) return DIS_ERR_UNSUPPORTED_COMBINATION;
  for (int i = 0; i < count; ++i)
    if (!ins[i] || !outs[i]) return DIS_ERR_NULL_POINTER;
  if (count == 0 || N == 0) return DIS_OK;
  return resize_bilinear_forward(ins, outs, count, N, C, H, W, oh, ow, mode, as_stream(stream));
}

int dis_resize_bilinear_backward(const float* grad_out, float* grad_in, int N, int C, int H, int W, int oh, int ow,
                                 void* stream) {
  if (!grad_out || !grad_in) return DIS_ERR_NULL_POINTER;
  if (N < 0 || C < 1 || H < 1 || W < 1 || oh < 1 || ow < 1 || (long)oh * ow > INT_MAX || (long)H * W > INT_MAX) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  return resize_bilinear_backward(grad_out, grad_in, N, C, H, W, oh, ow, as_stream(stream));
}

const char* dis_status_string(int status) {
  switch (status) {
    case DIS_OK: return "ok";
    case DIS_ERR_INVALID_LOSS_TYPE: return "invalid loss type";
    case DIS_ERR_BAD_SHAPE: return "bad shape";
    case DIS_ERR_UNSUPPORTED_BLOCK_SIZE: return "unsupported block_size (odd, 1..15)";
    case DIS_ERR_NULL_POINTER: return "null pointer";
    case DIS_ERR_CUDA_LAUNCH: return "CUDA launch failed";
    case DIS_ERR_UNSUPPORTED_KSIZE: return "unsupported Sobel ksize (3 or 5)";
    case DIS_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case DIS_ERR_UNSUPPORTED_COMBINATION: return "unsupported combination of arguments for this entry point";
  }
  return "unknown status";
}

const char* dis_last_cuda_error(void) { return g_last_error; }

int dis_lcn_forward(const float* x, float* lcn, float* std_out, int N, int H, int W, int radius, float eps,
                    void* stream) {
  if (!x || !lcn || !std_out) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 1 || W < 1 || radius < 1 || radius > 8 || radius >= H || radius >= W) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  const int vec_ok = (W % 4 == 0) && aligned16(x, lcn, std_out);
  const size_t hw = (size_t)H * W;
  for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
    const int nb = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z;
    if (int rc = lcn_forward(x + n0 * hw, lcn + n0 * hw, std_out + n0 * hw, nb, H, W, radius, eps, vec_ok,
                             as_stream(stream)))
      return rc;
  }
  return DIS_OK;
}

int dis_lcn_prepare_input(const float* x, float* im_cat, float* std_out, int bs, int tl, int H, int W, int radius,
                          float eps, void* stream) {
  if (!x || !im_cat || !std_out) return DIS_ERR_NULL_POINTER;
  if (bs < 0 || tl < 1 || H < 1 || W < 1 || radius < 1 || radius > 8 || radius >= H || radius >= W ||
      (long)bs * tl > MAX_GRID_Z)
    return DIS_ERR_BAD_SHAPE;
  if (bs == 0) return DIS_OK;
  const size_t hw = (size_t)H * W;
  const int vec_ok = (W % 4 == 0) && aligned16(x, im_cat, std_out);
  return lcn_forward_ex(x, im_cat, std_out, im_cat + hw, bs * tl, H, W, radius, eps, vec_ok, tl, bs, 2 * hw,
                        as_stream(stream));
}

int dis_lcn_backward(const float* x, const float* lcn, const float* std_in, const float* g_lcn, const float* g_std,
                     float* grad_x, float* workspace, int N, int H, int W, int radius, float eps, void* stream) {
  if (!x || !lcn || !std_in || !grad_x || !workspace || (!g_lcn && !g_std)) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 1 || W < 1 || radius < 1 || radius > 8 || radius >= H || radius >= W) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  return lcn_backward(x, lcn, std_in, g_lcn, g_std, grad_x, workspace, N, H, W, radius, eps, as_stream(stream));
}

int dis_photometric_loss_forward(const float* es, const float* ta, float* out, int N, int C, int H, int W,
                                 int block_size, int type, float eps, void* stream) {
  if (int rc = check_block(block_size, type)) return rc;
  if (!es || !ta || !out) return DIS_ERR_NULL_POINTER;
  if (N < 0 || C < 1 || H < 1 || W < 1) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  const size_t hw = (size_t)H * W;
  if (type >= CENSUS_MSE && C == 1 && H >= 2 && W >= 2 && use_march(block_size / 2)) {   // pair-symmetric marching kernel
    const MarchPlan plan = march_plan(N, H, W, block_size / 2);
    for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
      PatternMarchArgs a{};
      a.es = es + (size_t)n0 * hw; a.im = ta + (size_t)n0 * hw; a.out = out + (size_t)n0 * hw;
      a.grad[0] = a.out;        // (non-null: selects the strip path)
      a.N = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z; a.H = H; a.W = W;
      a.ncb = plan.ncb; a.nrb = plan.nrb; a.band_rows = plan.band_rows;
      a.eps = eps; a.inv_k2 = 1.0f / (float)(block_size * block_size);
      if (int rc = dispatch_photometric_march(block_size / 2, a, plan, MARCH_MAP, type, as_stream(stream))) return rc;
    }
    return DIS_OK;
  }
  for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
    PhotoArgs a{};
    a.es = es + (size_t)n0 * C * hw; a.ta = ta + (size_t)n0 * C * hw; a.out = out + (size_t)n0 * hw;
    a.N = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z; a.C = C; a.H = H; a.W = W;
    a.eps = eps; a.inv_k2 = 1.0f / (float)(block_size * block_size);
    a.vec_ok = (W % 4 == 0) && aligned16(out);
    if (int rc = dispatch_photometric(block_size / 2, a, type, false, as_stream(stream))) return rc;
  }
  return DIS_OK;
}

int dis_photometric_loss_backward(const float* es, const float* ta, const float* grad_out, float* grad_es, int N,
                                  int C, int H, int W, int block_size, int type, float eps, void* stream) {
  if (int rc = check_block(block_size, type)) return rc;
  if (!es || !ta || !grad_out || !grad_es) return DIS_ERR_NULL_POINTER;
  if (N < 0 || C < 1 || H < 1 || W < 1 || C > MAX_GRID_Z) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  const size_t hw = (size_t)H * W;
  if (type >= CENSUS_MSE && C == 1 && H >= 2 && W >= 2 && use_march(block_size / 2)) {
    const MarchPlan plan = march_plan(N, H, W, block_size / 2);
    for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
      PatternMarchArgs a{};
      a.es = es + (size_t)n0 * hw; a.im = ta + (size_t)n0 * hw; a.std_in = grad_out + (size_t)n0 * hw;
      a.grad[0] = grad_es + (size_t)n0 * hw;
      a.N = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z; a.H = H; a.W = W;
      a.ncb = plan.ncb; a.nrb = plan.nrb; a.band_rows = plan.band_rows;
      a.eps = eps; a.inv_k2 = 1.0f / (float)(block_size * block_size);
      if (int rc = dispatch_photometric_march(block_size / 2, a, plan, MARCH_GRAD_E, type, as_stream(stream))) return rc;
    }
    return DIS_OK;
  }
  const int step = MAX_GRID_Z / C;
  for (int n0 = 0; n0 < N; n0 += step) {
    PhotoArgs a{};
    a.es = es + (size_t)n0 * C * hw; a.ta = ta + (size_t)n0 * C * hw; a.grad_out = grad_out + (size_t)n0 * hw;
    a.grad_es = grad_es + (size_t)n0 * C * hw;
    a.N = N - n0 < step ? N - n0 : step; a.C = C; a.H = H; a.W = W;
    a.eps = eps; a.inv_k2 = 1.0f / (float)(block_size * block_size);
    a.vec_ok = (W % 4 == 0) && aligned16(grad_es);
    if (int rc = dispatch_photometric(block_size / 2, a, type, true, as_stream(stream))) return rc;
  }
  return DIS_OK;
}

int dis_pattern_warp_forward(const float* disp, const float* pattern, float* proj, float* dproj_ddisp,
                             int32_t* corner_x0, int32_t* corner_y0, int N, int H, int W, void* stream) {
  if (!disp || !pattern || !proj) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  return pattern_warp_forward(disp, pattern, proj, dproj_ddisp, corner_x0, corner_y0, N, H, W, as_stream(stream));
}

int dis_pattern_loss_num_partials(int N, int H, int W) {
  if (N < 0 || H < 1 || W < 1) return DIS_ERR_BAD_SHAPE;
  long most = (long)N * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);      // tile kernels (mse / sad, maps, 1 x 1 windows)
  for (int R = 1; R <= MAX_R; ++R) {                                      // marching kernel (census types)
    const MarchPlan p = march_plan(N, H, W, R);
    const long n = (long)N * p.ncb * p.nrb;
    if (n > most) most = n;
  }
  return most > INT_MAX ? DIS_ERR_BAD_SHAPE : (int)most;
}

int dis_pattern_loss_forward(const float* disp, const float* im, const float* std_in, const float* pattern,
                             float* proj, float* diff, float* grad_num, float* partials, int N, int H, int W,
                             int block_size, int type, float eps, void* stream) {
  return dis_pattern_loss_forward_scaled(disp, im, std_in, pattern, proj, diff, grad_num, nullptr, partials, N, H, W,
                                         block_size, type, eps, stream);
}

int dis_pattern_loss_forward_scaled(const float* disp, const float* im, const float* std_in, const float* pattern,
                                    float* proj, float* diff, float* grad_num, const float* grad_scale, float* partials,
                                    int N, int H, int W, int block_size, int type, float eps, void* stream) {
  if (int rc = check_block(block_size, type)) return rc;
  if (!disp || !im || !pattern || !partials) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  const size_t hw = (size_t)H * W;
  const int num_blocks = dis_pattern_loss_num_partials(N, H, W);
  if (num_blocks < 0) return num_blocks;
  // census types without the per-pixel loss map: the pair-symmetric marching kernel (one scale: the packed pair is
  // (estimate, target)); everything else stays on the tile kernels
  if (type >= CENSUS_MSE && !diff && use_march(block_size / 2)) {
    const int R = block_size / 2;
    const MarchPlan plan = march_plan(N, H, W, R);
    const int per_frame = plan.ncb * plan.nrb;
    for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
      PatternMarchArgs a{};
      a.disp[0] = disp + n0 * hw;
      a.grad[0] = grad_num ? grad_num + n0 * hw : nullptr;
      a.proj = proj ? proj + n0 * hw : nullptr;
      a.im = im + n0 * hw; a.std_in = std_in ? std_in + n0 * hw : nullptr; a.pattern = pattern;
      a.partials = partials;
      a.grad_scale = grad_scale;
      a.N = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z; a.H = H; a.W = W;
      a.ncb = plan.ncb; a.nrb = plan.nrb; a.band_rows = plan.band_rows;
      a.num_blocks = num_blocks; a.total_blocks = N * per_frame; a.block_offset = n0 * per_frame;
      a.eps = eps; a.inv_k2 = 1.0f / (float)(block_size * block_size);
      a.inv_w = 1.0f / (float)(W - 1); a.inv_h = 1.0f / (float)(H - 1);
      if (int rc = dispatch_pattern_march(R, a, plan, 1, type, as_stream(stream))) return rc;
    }
    return DIS_OK;
  }
  const int per_frame = ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
  if (num_blocks > N * per_frame) {   // slots the tile kernels do not write must read as zero
    const cudaError_t e = cudaMemsetAsync(partials, 0, sizeof(float) * 2 * (size_t)num_blocks, as_stream(stream));
    if (e != cudaSuccess) { set_last_cuda_error(e); return DIS_ERR_CUDA_LAUNCH; }
  }
  for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
    PatternLossArgs a{};
    a.disp = disp + n0 * hw; a.im = im + n0 * hw; a.std_in = std_in ? std_in + n0 * hw : nullptr;
    a.pattern = pattern;
    a.proj = proj ? proj + n0 * hw : nullptr; a.diff = diff ? diff + n0 * hw : nullptr;
    a.grad_num = grad_num ? grad_num + n0 * hw : nullptr;
    a.grad_scale = grad_scale;
    a.partials = partials + (size_t)2 * n0 * per_frame;
    a.N = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z; a.H = H; a.W = W;
    a.eps = eps; a.inv_k2 = 1.0f / (float)(block_size * block_size);
    a.inv_w = 1.0f / (float)(W - 1); a.inv_h = 1.0f / (float)(H - 1);
    a.vec_ok = (W % 4 == 0) && aligned16(proj, diff, grad_num);
    if (int rc = dispatch_pattern_loss(block_size / 2, a, type, as_stream(stream))) return rc;
  }
  return DIS_OK;
}

int dis_reduce_pairs(const float* partials, int n, float* out3, void* stream) {
  if (!partials || !out3) return DIS_ERR_NULL_POINTER;
  if (n < 0) return DIS_ERR_BAD_SHAPE;
  return reduce_pairs(partials, n, 1, out3, as_stream(stream));
}

int dis_reduce_pairs_batched(const float* partials, int n, int count, float* out3, void* stream) {
  if (!partials || !out3) return DIS_ERR_NULL_POINTER;
  if (n < 0 || count < 0 || count > MAX_GRID_Z) return DIS_ERR_BAD_SHAPE;
  if (count == 0) return DIS_OK;
  return reduce_pairs(partials, n, count, out3, as_stream(stream));
}

int dis_pattern_loss_multi_num_partials(int N, int H, int W) {
  if (N < 0 || H < 1 || W < 1) return DIS_ERR_BAD_SHAPE;
  long most = (long)N * ((H + MTH - 1) / MTH) * ((W + MTW - 1) / MTW);   // tile kernel (1 x 1 windows, A/B runs)
  for (int R = 1; R <= MAX_R; ++R) {                                        // marching kernel: depends on the window
    const MarchPlan p = march_plan(N, H, W, R);
    const long n = (long)N * p.ncb * p.nrb;
    if (n > most) most = n;
  }
  return most > INT_MAX ? DIS_ERR_BAD_SHAPE : (int)most;
}

int dis_pattern_loss_multi_forward(const float* const* disps, int S, const float* im, const float* std_in,
                                   const float* pattern, float* const* grad_nums, float* partials, int N, int H,
                                   int W, int block_size, int type, float eps, void* stream) {
  return dis_pattern_loss_multi_forward_scaled(disps, S, im, std_in, pattern, grad_nums, nullptr, partials, N, H, W,
                                               block_size, type, eps, stream);
}

int dis_pattern_loss_multi_forward_scaled(const float* const* disps, int S, const float* im, const float* std_in,
                                          const float* pattern, float* const* grad_nums, const float* grad_scale,
                                          float* partials, int N, int H, int W, int block_size, int type, float eps,
                                          void* stream) {
  if (int rc = check_block(block_size, type)) return rc;
  if (type != CENSUS_MSE && type != CENSUS_SAD) return DIS_ERR_UNSUPPORTED_COMBINATION;
  if (S != 2 && S != 4) return DIS_ERR_UNSUPPORTED_COMBINATION;
  if (!disps || !im || !pattern || !partials) return DIS_ERR_NULL_POINTER;
  bool any_grad = false, all_grad = true;
  for (int s = 0; s < S; ++s) {
    if (!disps[s]) return DIS_ERR_NULL_POINTER;
    const bool g = grad_nums && grad_nums[s];
    any_grad |= g;
    all_grad &= g;
  }
  if (any_grad && !all_grad) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  const size_t hw = (size_t)H * W;
  if (use_march(block_size / 2)) {
    const int R = block_size / 2;
    const MarchPlan plan = march_plan(N, H, W, R);
    const int num_blocks = dis_pattern_loss_multi_num_partials(N, H, W);
    if (num_blocks < 0) return num_blocks;
    const int per_frame = plan.ncb * plan.nrb;
    for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
      PatternMarchArgs a{};
      for (int s = 0; s < S; ++s) {
        a.disp[s] = disps[s] + n0 * hw;
        a.grad[s] = any_grad ? grad_nums[s] + n0 * hw : nullptr;
      }
      a.im = im + n0 * hw; a.std_in = std_in ? std_in + n0 * hw : nullptr; a.pattern = pattern;
      a.partials = partials;
      a.grad_scale = grad_scale;
      a.N = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z; a.H = H; a.W = W;
      a.ncb = plan.ncb; a.nrb = plan.nrb; a.band_rows = plan.band_rows;
      a.num_blocks = num_blocks; a.total_blocks = N * per_frame; a.block_offset = n0 * per_frame;
      a.eps = eps; a.inv_k2 = 1.0f / (float)(block_size * block_size);
      a.inv_w = 1.0f / (float)(W - 1); a.inv_h = 1.0f / (float)(H - 1);
      if (int rc = dispatch_pattern_march(R, a, plan, S, type, as_stream(stream))) return rc;
    }
    return DIS_OK;
  }
  const int per_frame = ((H + MTH - 1) / MTH) * ((W + MTW - 1) / MTW);
  const int num_blocks_all = dis_pattern_loss_multi_num_partials(N, H, W);
  if (num_blocks_all < 0) return num_blocks_all;
  if (num_blocks_all > N * per_frame) {   // slots this kernel does not write must read as zero
    const cudaError_t e = cudaMemsetAsync(partials, 0, sizeof(float) * 2 * S * (size_t)num_blocks_all, as_stream(stream));
    if (e != cudaSuccess) { set_last_cuda_error(e); return DIS_ERR_CUDA_LAUNCH; }
  }
  for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
    PatternMultiArgs a{};
    uintptr_t align = 0;
    for (int s = 0; s < S; ++s) {
      a.disp[s] = disps[s] + n0 * hw;
      a.grad_num[s] = any_grad ? grad_nums[s] + n0 * hw : nullptr;
      align |= reinterpret_cast<uintptr_t>(a.grad_num[s]);
    }
    a.im = im + n0 * hw; a.std_in = std_in ? std_in + n0 * hw : nullptr; a.pattern = pattern;
    a.partials = partials;
    a.grad_scale = grad_scale;
    a.N = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z; a.H = H; a.W = W;
    a.num_blocks = num_blocks_all; a.block_offset = n0 * per_frame;
    a.eps = eps; a.inv_k2 = 1.0f / (float)(block_size * block_size);
    a.inv_w = 1.0f / (float)(W - 1); a.inv_h = 1.0f / (float)(H - 1);
    a.vec_ok = (W % 2 == 0) && (align & 7) == 0;
    if (int rc = dispatch_pattern_multi(block_size / 2, a, S, type, as_stream(stream))) return rc;
  }
  return DIS_OK;
}

int dis_pattern_loss_point_num_partials(int N, int H, int W) {
  if (N < 0 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if ((size_t)H * W > (size_t)INT_MAX / 2 || H > 4 * 65535 || (size_t)N * ((H + 3) / 4) * ((W + 255) / 256) > (size_t)INT_MAX / 16) return DIS_ERR_BAD_SHAPE;
  return point_loss_num_partials(N, H, W);
}

int dis_pattern_loss_point_forward(const float* const* disps, int S, const float* im, const float* std_in,
                                   const float* pattern, float* const* projs, float* const* grad_nums,
                                   const float* grad_scale, float* workspace, int reuse_wbox, float* partials, int N, int H,
                                   int W, int block_size, int type, void* stream) {
  if (int rc = check_block(block_size, type)) return rc;
  if (type != MSE && type != SAD) return DIS_ERR_UNSUPPORTED_COMBINATION;
  if (S < 1 || S > 4) return DIS_ERR_UNSUPPORTED_COMBINATION;
  if (!disps || !im || !pattern || !workspace || !partials) return DIS_ERR_NULL_POINTER;
  for (int s = 0; s < S; ++s)
    if (!disps[s]) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (dis_pattern_loss_point_num_partials(N, H, W) < 0) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  float* wbox = workspace + (size_t)N * H * W;
  if (!reuse_wbox)
    if (int rc = box_weight(std_in, workspace, wbox, N, H, W, block_size, as_stream(stream))) return rc;
  return point_pattern_loss(disps, S, im, std_in, wbox, pattern, projs, grad_nums, grad_scale, partials, N, H, W, type,
                            as_stream(stream));
}

int dis_scale_by_device_scalar(const float* in, float* out, size_t n, const float* numer, const float* denom,
                               void* stream) {
  if (!in || !out || !numer) return DIS_ERR_NULL_POINTER;
  if (n == 0) return DIS_OK;
  return scale_by_device_scalar(in, out, n, numer, denom, as_stream(stream));
}

int dis_l1_num_partials(size_t n) { return l1_num_partials(n); }

int dis_l1_forward(const float* a, const float* b, float* sign_out, float* partials, size_t n, void* stream) {
  if (!a || !partials) return DIS_ERR_NULL_POINTER;   // b == NULL: sum |a|
  if (n == 0) return DIS_ERR_BAD_SHAPE;
  return l1_forward(a, b, sign_out, partials, n, as_stream(stream));
}

int dis_masked_l1_forward(const float* a, const float* b, const float* noise, float threshold, float* sign_out,
                          float* partials, size_t n, void* stream) {
  if (!a || !b || !partials) return DIS_ERR_NULL_POINTER;
  if (n == 0) return DIS_ERR_BAD_SHAPE;
  return masked_l1_forward(a, b, noise, threshold, sign_out, partials, n, as_stream(stream));
}

int dis_mul(const float* a, const float* b, float* out, size_t n, void* stream) {
  if (!a || !b || !out) return DIS_ERR_NULL_POINTER;
  if (n == 0) return DIS_OK;
  return mul(a, b, out, n, as_stream(stream));
}

int dis_sobel_forward(const float* x, float* out, int N, int H, int W, int ksize, void* stream) {
  if (ksize != 3 && ksize != 5) return DIS_ERR_UNSUPPORTED_KSIZE;
  if (!x || !out) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 1 || W < 1) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  return sobel_forward(x, out, N, H, W, ksize, as_stream(stream));
}

int dis_sobel_backward(const float* grad_out, float* grad_x, int N, int H, int W, int ksize, void* stream) {
  if (ksize != 3 && ksize != 5) return DIS_ERR_UNSUPPORTED_KSIZE;
  if (!grad_out || !grad_x) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 1 || W < 1) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  return sobel_backward(grad_out, grad_x, N, H, W, ksize, as_stream(stream));
}

int dis_smooth_loss_num_partials(int N, int H, int W) {
  if (N < 0 || H < 1 || W < 1) return DIS_ERR_BAD_SHAPE;
  return smooth_loss_num_partials(N, H, W);
}

int dis_smooth_loss_forward(const float* disp, const float* im, float* grad_sum, float* partials, int N, int H,
                            int W, void* stream) {
  return dis_smooth_loss_forward_scaled(disp, im, grad_sum, partials, N, H, W, 1.0f, 0, stream);
}

int dis_smooth_loss_forward_scaled(const float* disp, const float* im, float* grad_sum, float* partials, int N, int H,
                                   int W, float grad_scale, int accumulate, void* stream) {
  if (!disp || !im || !partials) return DIS_ERR_NULL_POINTER;
  if (N < 0 || H < 1 || W < 1) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  const size_t hw = (size_t)H * W;
  const int per_frame = smooth_loss_num_partials(1, H, W);
  for (int n0 = 0; n0 < N; n0 += MAX_GRID_Z) {
    const int nb = N - n0 < MAX_GRID_Z ? N - n0 : MAX_GRID_Z;
    if (int rc = smooth_loss_forward(disp + n0 * hw, im + n0 * hw, grad_sum ? grad_sum + n0 * hw : nullptr,
                                     partials + (size_t)2 * n0 * per_frame, nb, H, W, grad_scale, accumulate, as_stream(stream)))
      return rc;
  }
  return DIS_OK;
}

int dis_flow_warp_forward(const float* x, const float* flow, float* out, float* fb_mask_out, int32_t* corner_x0,
                          int32_t* corner_y0, int N, int C, int H, int W, void* stream) {
  if (!x || !flow || !out) return DIS_ERR_NULL_POINTER;
  if (N < 0 || C < 1 || H < 2 || W < 2 || (fb_mask_out && C != 2)) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  return flow_warp_forward(x, flow, out, fb_mask_out, corner_x0, corner_y0, N, C, H, W, as_stream(stream));
}

int dis_flow_warp_backward(const float* x, const float* flow, const float* grad_out, float* grad_x,
                           float* grad_flow, int N, int C, int H, int W, void* stream) {
  if (!flow || !grad_out || (!grad_x && !grad_flow) || (grad_flow && !x)) return DIS_ERR_NULL_POINTER;
  if (N < 0 || C < 1 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (N == 0) return DIS_OK;
  return flow_warp_backward(x, flow, grad_out, grad_x, grad_flow, N, C, H, W, as_stream(stream));
}

int dis_flow_warp_gather_forward(const float* x, const float* const* flows, float* out, int tl, int tidx, int bs, int C,
                                 int H, int W, void* stream) {
  if (!x || !out || (!flows && tl > 1)) return DIS_ERR_NULL_POINTER;
  if (tl < 1 || tl > 8 || tidx < 0 || tidx >= tl || bs < 0 || C < 1 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (bs == 0) return DIS_OK;
  return flow_warp_gather_forward(x, flows, out, tl, tidx, bs, C, H, W, as_stream(stream));
}

int dis_flow_warp_gather_backward(const float* const* flows, const float* grad_out, float* grad_x, int tl, int tidx,
                                  int bs, int C, int H, int W, void* stream) {
  if (!grad_out || !grad_x || (!flows && tl > 1)) return DIS_ERR_NULL_POINTER;
  if (tl < 1 || tl > 8 || tidx < 0 || tidx >= tl || bs < 0 || C < 1 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (bs == 0) return DIS_OK;
  return flow_warp_gather_backward(flows, grad_out, grad_x, tl, tidx, bs, C, H, W, as_stream(stream));
}

int dis_flow_warp_gather_all_forward(const float* x, const float* const* flows, float* out, int tl, int bs, int C, int H,
                                     int W, void* stream) {
  if (!x || !out || (!flows && tl > 1)) return DIS_ERR_NULL_POINTER;
  if (tl < 1 || tl > 8 || bs < 0 || C < 1 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (bs == 0) return DIS_OK;
  return flow_warp_gather_all_forward(x, flows, out, tl, bs, C, H, W, as_stream(stream));
}

int dis_flow_warp_gather_all_backward(const float* const* flows, const float* grad_out, float* grad_x, int tl, int bs,
                                      int C, int H, int W, void* stream) {
  if (!grad_out || !grad_x || (!flows && tl > 1)) return DIS_ERR_NULL_POINTER;
  if (tl < 1 || tl > 8 || bs < 0 || C < 1 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  if (bs == 0) return DIS_OK;
  return flow_warp_gather_all_backward(flows, grad_out, grad_x, tl, bs, C, H, W, as_stream(stream));
}

int dis_flow_consistency_num_partials(int bs, int H, int W) {
  if (bs < 0 || H < 2 || W < 2) return DIS_ERR_BAD_SHAPE;
  return bs * flow_consistency_blocks_per_frame(H, W);
}

int dis_flow_consistency_forward(const float* depth0, const float* depth1, const float* R0, const float* t0,
                                 const float* R1, const float* t1, const float* flow0, const float* flow1,
                                 const float* amb0, const float* amb1, int amb_channels, const float* primary_depth1,
                                 const float* K, const float* ray, float clamp, float fb_scale, float* loss_mask,
                                 float* orig_mask, float* grad_depth0, float* grad_depth1, float* partials, int bs,
                                 int H, int W, void* stream) {
  if (!depth0 || !depth1 || !R0 || !t0 || !R1 || !t1 || !flow0 || !flow1 || !amb0 || !amb1 || !K || !ray || !partials)
    return DIS_ERR_NULL_POINTER;
  if (bs < 0 || bs > MAX_GRID_Z || H < 2 || W < 2 || amb_channels < 1) return DIS_ERR_BAD_SHAPE;
  if (bs == 0) return DIS_OK;
  return flow_consistency_forward(depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, amb_channels,
                                  primary_depth1, K, ray, clamp, fb_scale, loss_mask, orig_mask, grad_depth0,
                                  grad_depth1, partials, bs, H, W, as_stream(stream));
}

int dis_combine2(const float* a, const float* b, float* out, size_t n, const float* numer, const float* den_a,
                 const float* den_b, float eps, void* stream) {
  if (!a || !out || !numer || !den_a || (b && !den_b)) return DIS_ERR_NULL_POINTER;
  if (n == 0) return DIS_OK;
  return combine2(a, b, out, n, numer, den_a, den_b, eps, as_stream(stream));
}

int dis_geometric_grad_combine(const float* const* planes, const int* frame_of, int n_terms, const float* scale,
                               const float* disp, float baseline_focal, float* grad_disp, int tl, int bs, int H, int W,
                               void* stream) {
  if (!disp || !grad_disp || (n_terms > 0 && (!planes || !frame_of || !scale))) return DIS_ERR_NULL_POINTER;
  if (tl < 1 || bs < 0 || H < 1 || W < 1 || n_terms < 0) return DIS_ERR_BAD_SHAPE;
  if (bs == 0) return DIS_OK;
  return geometric_grad_combine(planes, frame_of, n_terms, scale, disp, baseline_focal, grad_disp, tl,
                                (size_t)bs * H * W, as_stream(stream));
}

int dis_conv3d_out_size(int n, int ksize, int stride) {
  if (n < 1 || ksize < 1 || (ksize & 1) == 0 || stride < 1) return DIS_ERR_BAD_SHAPE;
  return conv3d_out_size(n, ksize, stride);
}

size_t dis_conv3d_scratch_elems(int tl, int bs, int h, int w) {
  if (tl < 1 || bs < 0 || h < 1 || w < 1) return 0;
  return conv3d_scratch_elems(tl, bs, h, w);
}

static int check_conv3d(int tl, int bs, int C, int h, int w, int ksize, int stride, int neighbors) {
  if (tl < 1 || bs < 0 || C < 1 || h < 1 || w < 1 || stride < 1) return DIS_ERR_BAD_SHAPE;
  if (ksize < 1 || (ksize & 1) == 0 || ksize * ksize * tl > 64 || neighbors < 1 || neighbors > ksize * ksize * tl ||
      neighbors > 16)
    return DIS_ERR_UNSUPPORTED_COMBINATION;
  return DIS_OK;
}

int dis_conv3d_gather_forward(const float* xyz, const float* feat, const float* mask, float* xyz_nb, float* feat_nb,
                              uint8_t* idx, float* scratch, int tl, int bs, int C, int h, int w, int ksize, int stride,
                              int neighbors, void* stream) {
  if (int rc = check_conv3d(tl, bs, C, h, w, ksize, stride, neighbors)) return rc;
  if (!xyz || !feat || !mask || !xyz_nb || !feat_nb || !idx || !scratch) return DIS_ERR_NULL_POINTER;
  if (bs == 0) return DIS_OK;
  return conv3d_gather_forward(xyz, feat, mask, xyz_nb, feat_nb, idx, scratch, tl, bs, C, h, w, ksize, stride,
                               neighbors, as_stream(stream));
}

int dis_conv3d_rank(const float* xyz, const float* mask, float* xyz_nb, uint8_t* idx, float* scratch, int tl, int bs, int h,
                    int w, int ksize, int stride, int neighbors, void* stream) {
  if (int rc = check_conv3d(tl, bs, 1, h, w, ksize, stride, neighbors)) return rc;
  if (!xyz || !mask || !xyz_nb || !idx || !scratch) return DIS_ERR_NULL_POINTER;
  if (bs == 0) return DIS_OK;
  return conv3d_rank(xyz, mask, xyz_nb, idx, scratch, tl, bs, h, w, ksize, stride, neighbors, as_stream(stream));
}

int dis_conv3d_gather_features(const float* feat, const uint8_t* idx, float* feat_nb, int tl, int bs, int C, int h, int w,
                               int ksize, int stride, int neighbors, void* stream) {
  if (int rc = check_conv3d(tl, bs, C, h, w, ksize, stride, neighbors)) return rc;
  if (!feat || !idx || !feat_nb) return DIS_ERR_NULL_POINTER;
  if (bs == 0) return DIS_OK;
  return conv3d_gather_features(feat, idx, feat_nb, tl, bs, C, h, w, ksize, stride, neighbors, as_stream(stream));
}

int dis_conv3d_gather_backward(const float* g_xyz_nb, const float* g_feat_nb, const uint8_t* idx, float* g_xyz,
                               float* g_feat, int tl, int bs, int C, int h, int w, int ksize, int stride,
                               int neighbors, void* stream) {
  if (int rc = check_conv3d(tl, bs, C, h, w, ksize, stride, neighbors)) return rc;
  if (!idx || (!g_xyz && !g_feat) || (g_xyz && !g_xyz_nb) || (g_feat && !g_feat_nb)) return DIS_ERR_NULL_POINTER;
  if (bs == 0) return DIS_OK;
  return conv3d_gather_backward(g_xyz_nb, g_feat_nb, idx, g_xyz, g_feat, tl, bs, C, h, w, ksize, stride, neighbors,
                                as_stream(stream));
}

}  // extern "C"
