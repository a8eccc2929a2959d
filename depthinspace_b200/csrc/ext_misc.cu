// The four ext ops the reference wraps but never calls (model/ext_functions.py:41-110): nn, crosscheck, proj_nn, xcorrvol.
// Their definitions live in the un-vendored Connecting-the-Dots torchext (autonomousvision/connecting_the_dots, no pinned
// version): PARITY UNPINNED.  Semantics restated from that project's published ext_kernel.h functors:
//   nn(in0 [n0,D], in1 [n1,D])            -> int64 [n0]: index of the L2-nearest row of in1 (first minimum wins)
//   crosscheck(in0 int64 [n0], in1 int64 [n1]) -> uint8 [n0]: in1[in0[i]] == i
//   proj_nn(xyz0, xyz1 [bs,H,W,3], K [3,3], patch) -> int64 [bs,H,W]: project xyz0 with K, round, search the patch x patch
//                                             window of xyz1 around that pixel for the 3-D nearest point (flat index or -1)
//   xcorrvol(in0, in1 [C,H,W], n_disps, block) -> float [n_disps,H,W]: sum over channels of the zero-normalised cross
//                                             correlation of the block at (h,w) in in0 with the block at (h,w-d) in in1
//                                             (replicate-clamped taps, + 1e-8 on the norm)
// No callers in the reference, so these are plain one-thread-per-output kernels (nn stages in1 through shared memory).
#include "common.cuh"

namespace dis {
namespace {

constexpr int NN_TILE = 256, NN_MAX_DIM = 8;

__global__ void __launch_bounds__(NN_TILE) nn_kernel(const float* __restrict__ in0, const float* __restrict__ in1,
                                                     long long* __restrict__ out, long n0, long n1, int dim) {
  __shared__ float tile[NN_TILE * NN_MAX_DIM];
  const long i = (long)blockIdx.x * NN_TILE + threadIdx.x;
  float q[NN_MAX_DIM];
#pragma unroll
  for (int d = 0; d < NN_MAX_DIM; ++d) q[d] = (i < n0 && d < dim) ? in0[i * dim + d] : 0.0f;
  float best = 1e9f;
  long long arg = -1;
  for (long j0 = 0; j0 < n1; j0 += NN_TILE) {
    const int m = (int)min((long)NN_TILE, n1 - j0);
    __syncthreads();
    for (int k = threadIdx.x; k < m * dim; k += NN_TILE) tile[k] = in1[j0 * dim + k];
    __syncthreads();
    for (int j = 0; j < m; ++j) {
      float dist = 0.0f;
      for (int d = 0; d < dim; ++d) {
        const float diff = q[d] - tile[j * dim + d];
        dist += diff * diff;
      }
      if (dist < best) { best = dist; arg = j0 + j; }
    }
  }
  if (i < n0) out[i] = arg;
}

__global__ void __launch_bounds__(256) crosscheck_kernel(const long long* __restrict__ in0, const long long* __restrict__ in1,
                                                         uint8_t* __restrict__ out, long n0, long n1) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n0) return;
  const long long j = in0[i];
  out[i] = (j >= 0 && j < n1 && in1[j] == i) ? 1 : 0;
}

__global__ void __launch_bounds__(256) proj_nn_kernel(const float* __restrict__ xyz0, const float* __restrict__ xyz1,
                                                      const float* __restrict__ K, long long* __restrict__ out, long total,
                                                      int H, int W, int patch) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const long b = i / ((long)H * W);
  const float x = xyz0[i * 3], y = xyz0[i * 3 + 1], z = xyz0[i * 3 + 2];
  const float d = K[6] * x + K[7] * y + K[8] * z;
  const float u = (K[0] * x + K[1] * y + K[2] * z) / d;
  const float v = (K[3] * x + K[4] * y + K[5] * z) / d;
  long long arg = -1;
  float best = 1e9f;
  if (u > -1e8f && u < 1e8f && v > -1e8f && v < 1e8f) {     // (NaN / huge projections match nothing)
    const int u0 = (int)(u + 0.5f), v0 = (int)(v + 0.5f);
    for (int p = 0; p < patch * patch; ++p) {
      const int u1 = u0 + p % patch - patch / 2, v1 = v0 + p / patch - patch / 2;
      if (u1 < 0 || v1 < 0 || u1 >= W || v1 >= H) continue;
      const long j = (b * H + v1) * W + u1;
      const float dx = x - xyz1[j * 3], dy = y - xyz1[j * 3 + 1], dz = z - xyz1[j * 3 + 2];
      const float dist = dx * dx + dy * dy + dz * dz;
      if (dist < best) { best = dist; arg = j; }
    }
  }
  out[i] = arg;
}

__global__ void __launch_bounds__(256) xcorrvol_kernel(const float* __restrict__ in0, const float* __restrict__ in1,
                                                       float* __restrict__ out, int C, int H, int W, int n_disps, int block) {
  const long o = (long)blockIdx.x * 256 + threadIdx.x;
  if (o >= (long)n_disps * H * W) return;
  const int d = (int)(o / ((long)H * W)), h = (int)((o / W) % H), w = (int)(o % W);
  const float inv = 1.0f / (float)(block * block);
  float val = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float* a = in0 + (size_t)c * H * W;
    const float* b = in1 + (size_t)c * H * W;
    float mu0 = 0.0f, mu1 = 0.0f;
    for (int bh = 0; bh < block; ++bh) {
      const int h0 = clampi(h + bh - block / 2, 0, H - 1);
      for (int bw = 0; bw < block; ++bw) {
        const int w0 = w + bw - block / 2;
        mu0 += a[(size_t)h0 * W + clampi(w0, 0, W - 1)] * inv;
        mu1 += b[(size_t)h0 * W + clampi(w0 - d, 0, W - 1)] * inv;
      }
    }
    float s0 = 0.0f, s1 = 0.0f, dot = 0.0f;
    for (int bh = 0; bh < block; ++bh) {
      const int h0 = clampi(h + bh - block / 2, 0, H - 1);
      for (int bw = 0; bw < block; ++bw) {
        const int w0 = w + bw - block / 2;
        const float v0 = a[(size_t)h0 * W + clampi(w0, 0, W - 1)] - mu0;
        const float v1 = b[(size_t)h0 * W + clampi(w0 - d, 0, W - 1)] - mu1;
        dot += v0 * v1;
        s0 += v0 * v0;
        s1 += v1 * v1;
      }
    }
    val += dot / (sqrtf(s0 * s1) + 1e-8f);
  }
  out[o] = val;
}

inline unsigned blocks(long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

int ext_nn(const float* in0, const float* in1, long long* out, long n0, long n1, int dim, cudaStream_t s) {
  if (dim < 1 || dim > NN_MAX_DIM) return DIS_ERR_BAD_SHAPE;
  nn_kernel<<<(unsigned)((n0 + NN_TILE - 1) / NN_TILE), NN_TILE, 0, s>>>(in0, in1, out, n0, n1, dim);
  return check_launch();
}
int ext_crosscheck(const long long* in0, const long long* in1, uint8_t* out, long n0, long n1, cudaStream_t s) {
  crosscheck_kernel<<<blocks(n0), 256, 0, s>>>(in0, in1, out, n0, n1);
  return check_launch();
}
int ext_proj_nn(const float* xyz0, const float* xyz1, const float* K, long long* out, int bs, int H, int W, int patch,
                cudaStream_t s) {
  const long total = (long)bs * H * W;
  proj_nn_kernel<<<blocks(total), 256, 0, s>>>(xyz0, xyz1, K, out, total, H, W, patch);
  return check_launch();
}
int ext_xcorrvol(const float* in0, const float* in1, float* out, int C, int H, int W, int n_disps, int block, cudaStream_t s) {
  xcorrvol_kernel<<<blocks((long)n_disps * H * W), 256, 0, s>>>(in0, in1, out, C, H, W, n_disps, block);
  return check_launch();
}

}  // namespace dis
