// One translation unit per window radius R = DIS_R (block_size = 2R+1): keeps every k x k loop fully
// unrolled without a single multi-minute compile.  build.py compiles R = 0..7 in parallel.
#include <map>
#include <mutex>
#include "photometric_kernels.cuh"
#include "pattern_multi.cuh"
#include "pattern_march.cuh"
#include "box_kernels.cuh"
#include "census_kernels.cuh"

#ifndef DIS_R
#error "compile with -DDIS_R=<window radius>"
#endif

namespace dis {
namespace {

// Opt a kernel in to > 48 KB of dynamic shared memory.  Done once per kernel and process (thread-safe static
// initialisation), never on the launch path: keeps launches capturable into CUDA graphs.
template <typename K>
int prepare(K kernel, size_t smem) {
  if (smem <= 48 * 1024) return DIS_OK;
  struct Once {
    cudaError_t err;
    Once(K k, size_t bytes) : err(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)) {}
  };
  static std::map<std::pair<const void*, int>, cudaError_t> done;   // (kernel, device) pairs already opted in
  static std::mutex mu;
  int device = 0;
  cudaGetDevice(&device);
  std::lock_guard<std::mutex> lock(mu);
  const std::pair<const void*, int> key(reinterpret_cast<const void*>(kernel), device);
  auto it = done.find(key);
  if (it == done.end()) it = done.emplace(key, Once(kernel, smem).err).first;
  if (it->second != cudaSuccess) { set_last_cuda_error(it->second); return DIS_ERR_CUDA_LAUNCH; }
  return DIS_OK;
}

dim3 tile_grid(int H, int W, int z) { return dim3((W + TW - 1) / TW, (H + TH - 1) / TH, z); }

// mse / sad: separable box filters (HBM-bound)
template <int TYPE, int R>
int box_photometric_t(const PhotoArgs& a, bool backward, cudaStream_t s) {
  const dim3 block(16, 16);
  if (!backward) {
    const size_t smem = box_smem_bytes<R>(1, 1, 0);
    if (int rc = prepare(box_photometric_fwd_kernel<TYPE, R>, smem)) return rc;
    box_photometric_fwd_kernel<TYPE, R><<<tile_grid(a.H, a.W, a.N), block, smem, s>>>(a);
  } else {
    const size_t smem = box_smem_bytes<R>(1, 1, 0);
    if (int rc = prepare(box_photometric_bwd_kernel<TYPE, R>, smem)) return rc;
    box_photometric_bwd_kernel<TYPE, R><<<tile_grid(a.H, a.W, a.N), block, smem, s>>>(a);
  }
  return check_launch();
}

template <int TYPE, int R>
int box_pattern_loss_t(const PatternLossArgs& a, cudaStream_t s) {
  const dim3 block(16, 16);
  const size_t smem = box_smem_bytes<R>(2, 2, 2);
  if (a.grad_num) {
    if (int rc = prepare(box_pattern_loss_kernel<TYPE, R, true>, smem)) return rc;
    box_pattern_loss_kernel<TYPE, R, true><<<tile_grid(a.H, a.W, a.N), block, smem, s>>>(a);
  } else {
    if (int rc = prepare(box_pattern_loss_kernel<TYPE, R, false>, smem)) return rc;
    box_pattern_loss_kernel<TYPE, R, false><<<tile_grid(a.H, a.W, a.N), block, smem, s>>>(a);
  }
  return check_launch();
}

template <int TYPE, int R>
int photometric_t(const PhotoArgs& a, bool backward, cudaStream_t s) {
  const dim3 block(16, 16);
  if (!backward) {
    const size_t smem = census_smem_bytes<R>(false, false);
    if (int rc = prepare(census_fwd_kernel<TYPE, R>, smem)) return rc;
    census_fwd_kernel<TYPE, R><<<tile_grid(a.H, a.W, a.N), block, smem, s>>>(a);
  } else {
    const size_t smem = census_smem_bytes<R>(true, false);
    if (int rc = prepare(census_bwd_kernel<TYPE, R>, smem)) return rc;
    census_bwd_kernel<TYPE, R><<<tile_grid(a.H, a.W, a.N * a.C), block, smem, s>>>(a);
  }
  return check_launch();
}

template <int TYPE, int R>
int pattern_loss_t(const PatternLossArgs& a, cudaStream_t s) {
  const dim3 block(16, 16);
  const size_t smem = census_smem_bytes<R>(true, true);
  if (a.grad_num) {
    if (int rc = prepare(census_pattern_loss_kernel<TYPE, R, true>, smem)) return rc;
    census_pattern_loss_kernel<TYPE, R, true><<<tile_grid(a.H, a.W, a.N), block, smem, s>>>(a);
  } else {
    if (int rc = prepare(census_pattern_loss_kernel<TYPE, R, false>, smem)) return rc;
    census_pattern_loss_kernel<TYPE, R, false><<<tile_grid(a.H, a.W, a.N), block, smem, s>>>(a);
  }
  return check_launch();
}

template <int TYPE, int R, int NPAIR>
int pattern_multi_t(const PatternMultiArgs& a, cudaStream_t s) {
  const dim3 block(32, 8);
  const dim3 grid((a.W + MTW - 1) / MTW, (a.H + MTH - 1) / MTH, a.N);
  const size_t smem = pattern_multi_smem_bytes<R, NPAIR>();
  if (a.grad_num[0]) {
    if (int rc = prepare(pattern_multi_kernel<TYPE, R, NPAIR, true>, smem)) return rc;
    pattern_multi_kernel<TYPE, R, NPAIR, true><<<grid, block, smem, s>>>(a);
  } else {
    if (int rc = prepare(pattern_multi_kernel<TYPE, R, NPAIR, false>, smem)) return rc;
    pattern_multi_kernel<TYPE, R, NPAIR, false><<<grid, block, smem, s>>>(a);
  }
  return check_launch();
}

#if DIS_R >= 1
template <int TYPE, int R, int NS>
int pattern_march_t(const PatternMarchArgs& a, const MarchPlan& plan, cudaStream_t s) {
  const dim3 block(32 * plan.nwarps);
  const dim3 grid(plan.ncb, plan.nrb, a.N);
  const size_t smem = pattern_march_smem_bytes<R, NS>(plan.nwarps);
  const size_t smem_max = pattern_march_smem_bytes<R, NS>(MARCH_MAX_WARPS);   // the opt-in is made once per kernel
  if (a.grad[0]) {
    if (int rc = prepare(pattern_march_kernel<TYPE, R, NS, true>, smem_max)) return rc;
    pattern_march_kernel<TYPE, R, NS, true><<<grid, block, smem, s>>>(a);
  } else {
    if (int rc = prepare(pattern_march_kernel<TYPE, R, NS, false>, smem_max)) return rc;
    pattern_march_kernel<TYPE, R, NS, false><<<grid, block, smem, s>>>(a);
  }
  return check_launch();
}
template <int TYPE, int R, int MODE>
int photometric_march_t(const PatternMarchArgs& a, const MarchPlan& plan, cudaStream_t s) {
  const dim3 block(32 * plan.nwarps);
  const dim3 grid(plan.ncb, plan.nrb, a.N);
  const size_t smem = pattern_march_smem_bytes<R, 1>(plan.nwarps);
  if (int rc = prepare(pattern_march_kernel<TYPE, R, 1, true, MODE>, pattern_march_smem_bytes<R, 1>(MARCH_MAX_WARPS))) return rc;
  pattern_march_kernel<TYPE, R, 1, true, MODE><<<grid, block, smem, s>>>(a);
  return check_launch();
}
template <int TYPE, int R>
int pattern_march_s(const PatternMarchArgs& a, const MarchPlan& plan, int S, cudaStream_t s) {
  switch (S) {
    case 1: return pattern_march_t<TYPE, R, 1>(a, plan, s);
    case 2: return pattern_march_t<TYPE, R, 2>(a, plan, s);
    case 4: return pattern_march_t<TYPE, R, 4>(a, plan, s);
  }
  return DIS_ERR_UNSUPPORTED_COMBINATION;
}
#endif

}  // namespace

template <>
int launch_pattern_march<DIS_R>(const PatternMarchArgs& a, const MarchPlan& plan, int S, int type, cudaStream_t s) {
#if DIS_R >= 1
  if (type == CENSUS_SAD) return pattern_march_s<CENSUS_SAD, DIS_R>(a, plan, S, s);
  if (type == CENSUS_MSE) return pattern_march_s<CENSUS_MSE, DIS_R>(a, plan, S, s);
#endif
  return DIS_ERR_UNSUPPORTED_COMBINATION;
}

template <>
int launch_photometric_march<DIS_R>(const PatternMarchArgs& a, const MarchPlan& plan, int mode, int type, cudaStream_t s) {
#if DIS_R >= 1
  if (type == CENSUS_SAD)
    return mode == MARCH_MAP ? photometric_march_t<CENSUS_SAD, DIS_R, MARCH_MAP>(a, plan, s) : photometric_march_t<CENSUS_SAD, DIS_R, MARCH_GRAD_E>(a, plan, s);
  if (type == CENSUS_MSE)
    return mode == MARCH_MAP ? photometric_march_t<CENSUS_MSE, DIS_R, MARCH_MAP>(a, plan, s) : photometric_march_t<CENSUS_MSE, DIS_R, MARCH_GRAD_E>(a, plan, s);
#endif
  return DIS_ERR_UNSUPPORTED_COMBINATION;
}

template <>
int launch_pattern_multi<DIS_R>(const PatternMultiArgs& a, int S, int type, cudaStream_t s) {
  if (type == CENSUS_SAD) return S == 4 ? pattern_multi_t<CENSUS_SAD, DIS_R, 2>(a, s) : pattern_multi_t<CENSUS_SAD, DIS_R, 1>(a, s);
  if (type == CENSUS_MSE) return S == 4 ? pattern_multi_t<CENSUS_MSE, DIS_R, 2>(a, s) : pattern_multi_t<CENSUS_MSE, DIS_R, 1>(a, s);
  return DIS_ERR_UNSUPPORTED_COMBINATION;
}

template <>
int launch_photometric<DIS_R>(const PhotoArgs& a, int type, bool backward, cudaStream_t s) {
  switch (type) {
    case MSE: return box_photometric_t<MSE, DIS_R>(a, backward, s);
    case SAD: return box_photometric_t<SAD, DIS_R>(a, backward, s);
    case CENSUS_MSE: return photometric_t<CENSUS_MSE, DIS_R>(a, backward, s);
    case CENSUS_SAD: return photometric_t<CENSUS_SAD, DIS_R>(a, backward, s);
  }
  return DIS_ERR_INVALID_LOSS_TYPE;
}

template <>
int launch_pattern_loss<DIS_R>(const PatternLossArgs& a, int type, cudaStream_t s) {
  switch (type) {
    case MSE: return box_pattern_loss_t<MSE, DIS_R>(a, s);
    case SAD: return box_pattern_loss_t<SAD, DIS_R>(a, s);
    case CENSUS_MSE: return pattern_loss_t<CENSUS_MSE, DIS_R>(a, s);
    case CENSUS_SAD: return pattern_loss_t<CENSUS_SAD, DIS_R>(a, s);
  }
  return DIS_ERR_INVALID_LOSS_TYPE;
}

}  // namespace dis
