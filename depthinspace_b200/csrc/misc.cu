// Stand-alone pattern warp and the small reduction / scaling utilities of the C-ABI.
#include "common.cuh"

namespace dis {
namespace {

// reference: model/networks.py:356-367 (pattern_proj) -- one thread per pixel, coalesced rows; the four
// pattern gathers hit L1/L2 (the pattern is batch-shared, 0.9 MB at 512x432).
__global__ void __launch_bounds__(256) pattern_warp_kernel(const float* __restrict__ disp,
                                                           const float* __restrict__ pattern,
                                                           float* __restrict__ proj, float* __restrict__ dproj,
                                                           int32_t* __restrict__ cx0, int32_t* __restrict__ cy0,
                                                           int H, int W, float inv_w, float inv_h, size_t total) {
  const size_t hw = (size_t)H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(idx % hw);
    const int h = pix / W, w = pix - h * W;
    float dd;
    int x0, y0;
    const float e = pattern_warp_pixel(pattern, ld_stream(disp + idx), h, w, H, W, inv_w, inv_h,
                                       dproj ? &dd : nullptr, &x0, &y0);
    proj[idx] = e;
    if (dproj) dproj[idx] = dd;
    if (cx0) cx0[idx] = x0;
    if (cy0) cy0[idx] = y0;
  }
}

// Fixed-order, single-CTA reduction of n (a,b) pairs in fp64: bitwise reproducible run to run.
__global__ void __launch_bounds__(1024) reduce_pairs_kernel(const float* __restrict__ p, int n, float* __restrict__ out3) {
  __shared__ double sa[1024], sb[1024];
  p += (size_t)2 * n * blockIdx.x;   // one CTA per independent segment
  out3 += 3 * blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) { a += (double)p[2 * i]; b += (double)p[2 * i + 1]; }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out3[0] = (float)sa[0];
    out3[1] = (float)sb[0];
    out3[2] = (float)(sa[0] / sb[0]);
  }
}

__global__ void __launch_bounds__(256) scale_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n4,
                                                    size_t n, const float* __restrict__ numer,
                                                    const float* __restrict__ denom) {
  const float s = denom ? __fdiv_rn(__ldg(numer), __ldg(denom)) : __ldg(numer);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t; i < n4; i += stride) {
    float4 v = __ldcs(reinterpret_cast<const float4*>(in) + i);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    __stcs(reinterpret_cast<float4*>(out) + i, v);
  }
  for (size_t i = 4 * n4 + t; i < n; i += stride) out[i] = in[i] * s;
}

__global__ void __launch_bounds__(256) mul_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                  float* __restrict__ out, size_t n4, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t; i < n4; i += stride) {
    const float4 u = __ldcs(reinterpret_cast<const float4*>(a) + i);
    const float4 v = b ? __ldcs(reinterpret_cast<const float4*>(b) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    __stcs(reinterpret_cast<float4*>(out) + i, make_float4(u.x * v.x, u.y * v.y, u.z * v.z, u.w * v.w));
  }
  for (size_t i = 4 * n4 + t; i < n; i += stride) out[i] = a[i] * b[i];
}

// mean |a - b| in one pass: per-CTA partial sums (sum |a-b|, count) and, optionally, sign(a - b) (the gradient of
// the sum w.r.t. a).  Auxiliary L1 terms of the workers: pseudo-GT (single_frame_worker.py:152-155), primary
// disparity (multi_frame_worker.py:160-165).
__global__ void __launch_bounds__(256) l1_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                 float* __restrict__ sgn, float* __restrict__ partials, size_t n4, size_t n) {
  __shared__ float red[16];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f, c = 0.f;
  for (size_t i = t; i < n4; i += stride) {
    const float4 u = __ldcs(reinterpret_cast<const float4*>(a) + i);
    const float4 v = b ? __ldcs(reinterpret_cast<const float4*>(b) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
    s += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
    c += 4.0f;
    if (sgn) __stcs(reinterpret_cast<float4*>(sgn) + i, make_float4(sign0(d0), sign0(d1), sign0(d2), sign0(d3)));
  }
  for (size_t i = 4 * n4 + t; i < n; i += stride) {
    const float d = a[i] - (b ? b[i] : 0.f);
    s += fabsf(d);
    c += 1.0f;
    if (sgn) sgn[i] = sign0(d);
  }
  block_sum2<256>(s, c, red);
  if (threadIdx.x == 0) { partials[2 * blockIdx.x] = s; partials[2 * blockIdx.x + 1] = c; }
}

// Masked L1 with additive noise, the SGM warm-up term of the single-frame worker (single_frame_worker.py:158-163):
//   valid = (b > threshold);  d = a - b + noise;  partials = (sum |d| * valid, sum valid);  sgn = sign(d) * valid
// (the reference evaluates a - b first and adds the noise to the difference: same order here)
__global__ void __launch_bounds__(256) masked_l1_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        const float* __restrict__ noise, float threshold,
                                                        float* __restrict__ sgn, float* __restrict__ partials, size_t n) {
  __shared__ float red[16];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float s = 0.f, c = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float bv = __ldcs(b + i);
    const float valid = bv > threshold ? 1.0f : 0.0f;
    float d = __ldcs(a + i) - bv;
    if (noise) d += __ldcs(noise + i);
    s = fmaf(fabsf(d), valid, s);
    c += valid;
    if (sgn) __stcs(sgn + i, sign0(d) * valid);
  }
  block_sum2<256>(s, c, red);
  if (threadIdx.x == 0) { partials[2 * blockIdx.x] = s; partials[2 * blockIdx.x + 1] = c; }
}

inline int stream_grid(size_t work_items, int threads) {
  const size_t want = (work_items + threads - 1) / threads;
  const size_t cap = 148 * 16;  // 16 resident 256-thread CTAs x 148 SMs is plenty for a streaming loop
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

int pattern_warp_forward(const float* disp, const float* pattern, float* proj, float* dproj, int32_t* cx0,
                         int32_t* cy0, int N, int H, int W, cudaStream_t s) {
  const size_t total = (size_t)N * H * W;
  const float inv_w = 1.0f / (float)(W - 1), inv_h = 1.0f / (float)(H - 1);
  pattern_warp_kernel<<<stream_grid(total, 256), 256, 0, s>>>(disp, pattern, proj, dproj, cx0, cy0, H, W, inv_w,
                                                              inv_h, total);
  return check_launch();
}

int reduce_pairs(const float* partials, int n, int count, float* out3, cudaStream_t s) {
  reduce_pairs_kernel<<<count, 1024, 0, s>>>(partials, n, out3);
  return check_launch();
}

int scale_by_device_scalar(const float* in, float* out, size_t n, const float* numer, const float* denom,
                           cudaStream_t s) {
  const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const size_t n4 = vec ? n / 4 : 0;
  scale_kernel<<<stream_grid(n4 ? n4 : n, 256), 256, 0, s>>>(in, out, n4, n, numer, denom);
  return check_launch();
}

int l1_num_partials(size_t n) { return stream_grid((n + 3) / 4, 256); }

int l1_forward(const float* a, const float* b, float* sgn, float* partials, size_t n, cudaStream_t s) {
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(sgn)) & 15) == 0;
  // the grid must match l1_num_partials(n) whatever the alignment: scalar fallback keeps the same grid
  l1_kernel<<<l1_num_partials(n), 256, 0, s>>>(a, b, sgn, partials, vec ? n / 4 : 0, n);
  return check_launch();
}

int masked_l1_forward(const float* a, const float* b, const float* noise, float threshold, float* sgn, float* partials,
                      size_t n, cudaStream_t s) {
  masked_l1_kernel<<<l1_num_partials(n), 256, 0, s>>>(a, b, noise, threshold, sgn, partials, n);
  return check_launch();
}

int mul(const float* a, const float* b, float* out, size_t n, cudaStream_t s) {
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const size_t n4 = vec ? n / 4 : 0;
  mul_kernel<<<stream_grid(n4 ? n4 : n, 256), 256, 0, s>>>(a, b, out, n4, n);
  return check_launch();
}

}  // namespace dis
