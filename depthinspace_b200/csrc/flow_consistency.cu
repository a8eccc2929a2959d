// Flow-consistency (geometric) loss, one direction (frame 0 -> frame 1), fused into a single pass.
// reference: Single_Frame_Flow_Consistency_Loss.fwd  (model/networks.py:619-655)
//            Multi_Frame_Flow_Consistency_Loss.fwd   (model/networks.py:564-601)
//            ProjectionBaseLoss.unproject/transform/project (model/networks.py:455-488)
//
//   d1      = z of  K (R1 ((depth0 * ray - t0) R0)^T + t1)            (depth of pixel p seen from view 1)
//   depth10 = bilinear(depth1; p + flow0(p))   zeros padding, align_corners=True, same fp32 coordinate round trip
//   diff    = |d1 - depth10|  (SF: clamped to [0, clamp])
//   fb_mask = |flow0 + flow10|^2 < 0.5 + fb_scale (|flow0|^2 + |flow10|^2),  flow10 = bilinear(flow1; ...)
//   vc_mask = mean_c |amb0 - amb10| < 0.01,                                 amb10  = bilinear(amb1; ...)
//   rf_mask = |bilinear(uv0; ...) - p|^2 < 1  (MF only), uv0 = projection into view 0 of view-1 pixels lifted
//             with primary_depth1
//   loss    = sum(diff * mask) / (sum(mask) + 1e-8)
//
// The reference runs 2 bmm + 3-4 grid_sample + ~40 elementwise kernels per direction and (SF) a blocking
// device->host copy.  Here one thread owns one pixel: the sampling coordinates and corner weights are
// computed once and shared by every sampled plane; the loss partial sums, the masks and both gradients
// (direct w.r.t. depth0; scatter w.r.t. depth1, RED.ADD like ATen's grid_sampler backward) come out of the
// same pass.  Gradients are un-normalised (d sum(diff*mask)); the caller scales by 1/(sum(mask)+1e-8).
#include "common.cuh"

namespace dis {
namespace {

struct FCArgs {
  const float* depth0; const float* depth1;
  const float* R0; const float* t0; const float* R1; const float* t1;   // [bs,3,3], [bs,3]
  const float* flow0; const float* flow1; const float* amb0; const float* amb1;
  const float* primary_depth1;   // NULL => no reprojection mask (single-frame variant)
  const float* K; const float* ray;   // [3,3] ; [H*W,3]
  float* loss_mask; float* orig_mask; float* grad_depth0; float* grad_depth1; float* partials;
  int bs, amb_c, H, W;
  float clamp, fb_scale, inv_w, inv_h;
};

// row-vector conventions of the reference: v' = v M  (bmm(xyz, M))
struct Pose { float R0[9], t0[3], R1[9], t1[3]; };

__device__ __forceinline__ void project_to_view1(const Pose& P, const float* __restrict__ K, float depth,
                                                 const float* __restrict__ ray, float out[3]) {
  // unproject (:455-461): xyz = depth * ray - t0 ; xyz = xyz R0
  const float x = depth * ray[0] - P.t0[0], y = depth * ray[1] - P.t0[1], z = depth * ray[2] - P.t0[2];
  float w[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) w[j] = fmaf(z, P.R0[6 + j], fmaf(y, P.R0[3 + j], x * P.R0[j]));
  // project (:467-476): xyz = xyz R1^T + t1 ; uv = xyz K^T
  float c[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) c[j] = fmaf(w[2], P.R1[3 * j + 2], fmaf(w[1], P.R1[3 * j + 1], w[0] * P.R1[3 * j])) + P.t1[j];
#pragma unroll
  for (int j = 0; j < 3; ++j) out[j] = fmaf(c[2], K[3 * j + 2], fmaf(c[1], K[3 * j + 1], c[0] * K[3 * j]));
}

// four-corner gather of one plane with ATen's in-bounds predicate: addresses come from the shared corner offsets
struct CornerIdx { int o[4]; bool in[4]; };
__device__ __forceinline__ void gather4(const float* __restrict__ plane, const CornerIdx& ci, float (&v)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = ci.in[k] ? __ldg(plane + ci.o[k]) : 0.0f;
}
__device__ __forceinline__ float blend4(const float (&v)[4], const CornerIdx& ci, const Bilinear& b) {
  float acc = 0.0f;
  if (ci.in[0]) acc = __fmaf_rn(v[0], b.wnw, acc);
  if (ci.in[1]) acc = __fmaf_rn(v[1], b.wne, acc);
  if (ci.in[2]) acc = __fmaf_rn(v[2], b.wsw, acc);
  if (ci.in[3]) acc = __fmaf_rn(v[3], b.wse, acc);
  return acc;
}

template <bool AMB1>
__global__ void __launch_bounds__(256, 4) flow_consistency_kernel(FCArgs a) {
  __shared__ Pose pose;      // 0 -> 1
  __shared__ Pose pose_inv;  // 1 -> 0 (reprojection mask)
  __shared__ float sK[9];
  __shared__ float red[2 * 8];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid < 9) {
    pose.R0[tid] = a.R0[n * 9 + tid]; pose.R1[tid] = a.R1[n * 9 + tid];
    pose_inv.R0[tid] = a.R1[n * 9 + tid]; pose_inv.R1[tid] = a.R0[n * 9 + tid];
    sK[tid] = a.K[tid];
  }
  if (tid < 3) {
    pose.t0[tid] = a.t0[n * 3 + tid]; pose.t1[tid] = a.t1[n * 3 + tid];
    pose_inv.t0[tid] = a.t1[n * 3 + tid]; pose_inv.t1[tid] = a.t0[n * 3 + tid];
  }
  __syncthreads();
  const int hw = a.H * a.W;
  const size_t fo = (size_t)n * hw;
  float num = 0.f, den = 0.f;
  for (int pix = blockIdx.x * 256 + tid; pix < hw; pix += gridDim.x * 256) {
    const int h = pix / a.W, w = pix - h * a.W;
    // sampling position p + flow0(p), reference op order (:626-631)
    const float fx = __ldg(a.flow0 + (size_t)n * 2 * hw + pix), fy = __ldg(a.flow0 + (size_t)n * 2 * hw + hw + pix);
    Bilinear b;
    bilinear_setup<false>(normalize_coord(fadd(fx, (float)w), a.inv_w), normalize_coord(fadd(fy, (float)h), a.inv_h), a.H,
                          a.W, b);
    // corner offsets / predicates once, then every gather of the pixel is issued before the first use
    CornerIdx ci;
    ci.in[0] = in_bounds(b.y0, b.x0, a.H, a.W); ci.in[1] = in_bounds(b.y0, b.x0 + 1, a.H, a.W);
    ci.in[2] = in_bounds(b.y0 + 1, b.x0, a.H, a.W); ci.in[3] = in_bounds(b.y0 + 1, b.x0 + 1, a.H, a.W);
    // (corner offsets from coordinates clamped to [-1, size]: identical for every corner that passes the predicate, and no
    //  int overflow for garbage flows whose y0 * W would not fit)
    ci.o[0] = clampi(b.y0, -1, a.H) * a.W + clampi(b.x0, -1, a.W); ci.o[1] = ci.o[0] + 1; ci.o[2] = ci.o[0] + a.W; ci.o[3] = ci.o[2] + 1;
    float vd[4], vfx[4], vfy[4], va[4];
    gather4(a.depth1 + fo, ci, vd);
    gather4(a.flow1 + (size_t)n * 2 * hw, ci, vfx);
    gather4(a.flow1 + (size_t)n * 2 * hw + hw, ci, vfy);
    if (AMB1) gather4(a.amb1 + fo, ci, va);
    const float dep0 = __ldg(a.depth0 + fo + pix);
    const float amb0v = AMB1 ? __ldg(a.amb0 + fo + pix) : 0.f;
    const float r0 = __ldg(a.ray + (size_t)pix * 3), r1 = __ldg(a.ray + (size_t)pix * 3 + 1), r2 = __ldg(a.ray + (size_t)pix * 3 + 2);
    // depth of p seen from view 1
    const float rayv[3] = {r0, r1, r2};
    float uvd[3];
    project_to_view1(pose, sK, dep0, rayv, uvd);
    const float d1 = uvd[2];
    const float depth10 = blend4(vd, ci, b);
    const float raw = d1 - depth10;
    float diff = fabsf(raw);
    const bool clamped = a.clamp > 0.f && diff > a.clamp;   // torch.clamp passes gradient on [0, clamp] inclusive
    if (a.clamp > 0.f) diff = fminf(diff, a.clamp);
    if (a.orig_mask) a.orig_mask[fo + pix] = (diff < a.clamp) ? 1.0f : 0.0f;
    // forward-backward flow mask (:644-646 / :589-591)
    const float f10x = blend4(vfx, ci, b);
    const float f10y = blend4(vfy, ci, b);
    const float sx = fadd(fx, f10x), sy = fadd(fy, f10y);
    const float lhs = fadd(fmul(sx, sx), fmul(sy, sy));
    const float mag = fadd(fadd(fmul(fx, fx), fmul(fy, fy)), fadd(fmul(f10x, f10x), fmul(f10y, f10y)));
    float mask = (lhs < fadd(0.5f, fmul(a.fb_scale, mag))) ? 1.0f : 0.0f;
    // visibility mask (:648-649): mean over channels of |amb0 - amb10| < 0.01
    float vc = 0.f;
    if (AMB1) {
      vc = fabsf(fsub(amb0v, blend4(va, ci, b)));
    } else {
      for (int c = 0; c < a.amb_c; ++c) {
        float vv[4];
        gather4(a.amb1 + ((size_t)n * a.amb_c + c) * hw, ci, vv);
        vc = fadd(vc, fabsf(fsub(__ldg(a.amb0 + ((size_t)n * a.amb_c + c) * hw + pix), blend4(vv, ci, b))));
      }
      vc = vc / (float)a.amb_c;
    }
    mask *= (vc < 0.01f) ? 1.0f : 0.0f;
    // reprojection mask (:591-595): view-1 pixels lifted with primary_depth1, projected into view 0, sampled at p+flow
    if (a.primary_depth1) {
      float wu = 0.f, wv = 0.f;
      const int cx[4] = {b.x0, b.x0 + 1, b.x0, b.x0 + 1}, cy[4] = {b.y0, b.y0, b.y0 + 1, b.y0 + 1};
      const float wgt[4] = {b.wnw, b.wne, b.wsw, b.wse};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!in_bounds(cy[k], cx[k], a.H, a.W)) continue;
        const int q = cy[k] * a.W + cx[k];
        float o[3];
        project_to_view1(pose_inv, sK, __ldg(a.primary_depth1 + fo + q), a.ray + (size_t)q * 3, o);
        const float dz = fmaxf(o[2], 0.f) + 1e-12f;    // relu(d) + 1e-12 (:484)
        wu = __fmaf_rn(__fdiv_rn(o[0], dz), wgt[k], wu);
        wv = __fmaf_rn(__fdiv_rn(o[1], dz), wgt[k], wv);
      }
      const float du = fsub(wu, (float)w), dv = fsub(wv, (float)h);
      mask *= (fadd(fmul(du, du), fmul(dv, dv)) < 1.0f) ? 1.0f : 0.0f;
    }
    if (a.loss_mask) a.loss_mask[fo + pix] = mask;
    num = fmaf(diff, mask, num);
    den += mask;
    if (a.grad_depth0 || a.grad_depth1) {
      // d(diff*mask)/d d1 = mask * sign(raw) inside the clamp range; d d1 / d depth0 = ((ray R0) R1^T K^T)[2]
      const float g = (clamped ? 0.f : sign0(raw)) * mask;
      if (a.grad_depth0) {
        const float* ray = rayv;
        float wv3[3], c3[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) wv3[j] = fmaf(ray[2], pose.R0[6 + j], fmaf(ray[1], pose.R0[3 + j], ray[0] * pose.R0[j]));
#pragma unroll
        for (int j = 0; j < 3; ++j) c3[j] = fmaf(wv3[2], pose.R1[3 * j + 2], fmaf(wv3[1], pose.R1[3 * j + 1], wv3[0] * pose.R1[3 * j]));
        const float dd1 = fmaf(c3[2], sK[8], fmaf(c3[1], sK[7], c3[0] * sK[6]));
        a.grad_depth0[fo + pix] = g * dd1;
      }
      if (a.grad_depth1 && g != 0.f) {
        float* gp = a.grad_depth1 + fo + (ptrdiff_t)b.y0 * a.W + b.x0;
        if (ci.in[0]) atomicAdd(gp, -g * b.wnw);
        if (ci.in[1]) atomicAdd(gp + 1, -g * b.wne);
        if (ci.in[2]) atomicAdd(gp + a.W, -g * b.wsw);
        if (ci.in[3]) atomicAdd(gp + a.W + 1, -g * b.wse);
      }
    }
  }
  block_sum2<256>(num, den, red);
  if (tid == 0) {
    a.partials[2 * ((size_t)n * gridDim.x + blockIdx.x)] = num;
    a.partials[2 * ((size_t)n * gridDim.x + blockIdx.x) + 1] = den;
  }
}

// out = numer * (a / (den_a + eps) + b / (den_b + eps))   (b may be NULL)
__global__ void __launch_bounds__(256) combine2_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ out, size_t n, const float* __restrict__ numer,
                                                       const float* __restrict__ den_a, const float* __restrict__ den_b,
                                                       float eps) {
  const float g = __ldg(numer);
  const float sa = g / (__ldg(den_a) + eps);
  const float sb = b ? g / (__ldg(den_b) + eps) : 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
    out[i] = b ? fmaf(__ldcs(b + i), sb, __ldcs(a + i) * sa) : __ldcs(a + i) * sa;
}

// Gradient of all geometric terms of the loss assembly w.r.t. the disparity maps in one pass: for frame f (blockIdx.y)
//   grad_disp[f] = d depth / d disp * sum_e scale[e] * plane_e        over the planes that belong to frame f
// with depth = bf / (max(disp, 0) + 1e-12) (DispToDepth, model/networks.py:311-319; relu has derivative 0 at 0).
// Replaces, per training step, 12 two-plane combines, the select-backward zero-fills / strided copies and the gradient
// accumulations autograd would run for the pair loop, and DispToDepth's backward.
constexpr int GC_MAX_FRAMES = 8, GC_MAX_PER_FRAME = 16;
struct GeoCombineArgs {
  const float* plane[GC_MAX_FRAMES][GC_MAX_PER_FRAME];
  short sidx[GC_MAX_FRAMES][GC_MAX_PER_FRAME];
  short count[GC_MAX_FRAMES];
};

__global__ void __launch_bounds__(256) geometric_grad_combine_kernel(const __grid_constant__ GeoCombineArgs a,
                                                                     const float* __restrict__ scale,
                                                                     const float* __restrict__ disp, float bf,
                                                                     float* __restrict__ grad_disp, size_t plane_elems) {
  const int f = blockIdx.y, cnt = a.count[f];
  float sc[GC_MAX_PER_FRAME];
#pragma unroll
  for (int e = 0; e < GC_MAX_PER_FRAME; ++e) sc[e] = e < cnt ? __ldg(scale + a.sidx[f][e]) : 0.f;
  const float* d = disp + (size_t)f * plane_elems;
  float* o = grad_disp + (size_t)f * plane_elems;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < plane_elems; i += (size_t)gridDim.x * 256) {
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < GC_MAX_PER_FRAME; ++e)
      if (e < cnt) acc = fmaf(__ldcs(a.plane[f][e] + i), sc[e], acc);
    const float dv = __ldcs(d + i);
    const float x = fmaxf(dv, 0.f) + 1e-12f;
    o[i] = dv > 0.f ? -(acc * bf) / (x * x) : 0.f;
  }
}

constexpr int FC_BLOCKS_PER_FRAME_MAX = 64;

}  // namespace

int flow_consistency_blocks_per_frame(int H, int W) {
  const int want = (H * W + 255) / 256;
  return want < FC_BLOCKS_PER_FRAME_MAX ? want : FC_BLOCKS_PER_FRAME_MAX;
}

int flow_consistency_forward(const float* depth0, const float* depth1, const float* R0, const float* t0, const float* R1,
                             const float* t1, const float* flow0, const float* flow1, const float* amb0,
                             const float* amb1, int amb_c, const float* primary_depth1, const float* K, const float* ray,
                             float clamp, float fb_scale, float* loss_mask, float* orig_mask, float* grad_depth0,
                             float* grad_depth1, float* partials, int bs, int H, int W, cudaStream_t s) {
  if (grad_depth1) {
    cudaError_t e = cudaMemsetAsync(grad_depth1, 0, sizeof(float) * (size_t)bs * H * W, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return DIS_ERR_CUDA_LAUNCH; }
  }
  FCArgs a{depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, primary_depth1, K, ray,
           loss_mask, orig_mask, grad_depth0, grad_depth1, partials, bs, amb_c, H, W,
           clamp, fb_scale, 1.0f / (float)(W - 1), 1.0f / (float)(H - 1)};
  dim3 grid(flow_consistency_blocks_per_frame(H, W), bs);
  if (amb_c == 1) flow_consistency_kernel<true><<<grid, 256, 0, s>>>(a);
  else flow_consistency_kernel<false><<<grid, 256, 0, s>>>(a);
  return check_launch();
}

int geometric_grad_combine(const float* const* planes, const int* frame_of, int n_terms, const float* scale,
                           const float* disp, float bf, float* grad_disp, int tl, size_t plane_elems, cudaStream_t s) {
  if (tl < 1 || tl > GC_MAX_FRAMES) return DIS_ERR_BAD_SHAPE;
  GeoCombineArgs a{};
  for (int k = 0; k < n_terms; ++k) {
    const int f = frame_of[k];
    if (f < 0 || f >= tl || !planes[k]) return DIS_ERR_BAD_SHAPE;
    if (a.count[f] >= GC_MAX_PER_FRAME) return DIS_ERR_UNSUPPORTED_COMBINATION;
    a.plane[f][a.count[f]] = planes[k];
    a.sidx[f][a.count[f]] = (short)k;
    ++a.count[f];
  }
  const size_t want = (plane_elems + 255) / 256, cap = 148 * 8;
  const dim3 grid((unsigned)(want < cap ? (want ? want : 1) : cap), tl);
  geometric_grad_combine_kernel<<<grid, 256, 0, s>>>(a, scale, disp, bf, grad_disp, plane_elems);
  return check_launch();
}

int combine2(const float* a, const float* b, float* out, size_t n, const float* numer, const float* den_a,
             const float* den_b, float eps, cudaStream_t s) {
  const size_t want = (n + 255) / 256, cap = 148 * 16;
  combine2_kernel<<<(int)(want < cap ? (want ? want : 1) : cap), 256, 0, s>>>(a, b, out, n, numer, den_a, den_b, eps);
  return check_launch();
}

}  // namespace dis
