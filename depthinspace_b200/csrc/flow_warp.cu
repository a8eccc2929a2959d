// Optical-flow warp used by DIS-MF fusion.
// reference: warp(x, flow), model/multi_frame_networks.py:83-99
//   (u + flow_x, v + flow_y) -> 2 * (c / (size-1) - 0.5) -> grid_sample(bilinear, zeros, align_corners=True)
// plus the forward-backward consistency mask of gather_warped_xyz (:205-207).
//
// The sampling coordinates and the four corner weights depend only on (n, h, w), so one thread owns
// one pixel and loops over the C channels: the coordinate math is paid once and every channel costs
// four (mostly L1-resident) gathers and one coalesced store.
// Backward w.r.t. x is a true scatter; like ATen (GridSampler.cuh safe_add_2d) it uses fp32 global
// reductions (RED.ADD), so it is reproducible up to summation order only.  Backward w.r.t. flow is a
// per-pixel gather and is exact/deterministic.
#include "common.cuh"

namespace dis {
namespace {

__device__ __forceinline__ void flow_bilinear(const float* __restrict__ flow, size_t n, int h, int w, int H, int W,
                                              float inv_w, float inv_h, Bilinear& b, float& fx, float& fy) {
  const size_t hw = (size_t)H * W;
  fx = ld_stream(flow + (n * 2 + 0) * hw + (size_t)h * W + w);
  fy = ld_stream(flow + (n * 2 + 1) * hw + (size_t)h * W + w);
  const float gx = normalize_coord(fadd(fx, (float)w), inv_w);
  const float gy = normalize_coord(fadd(fy, (float)h), inv_h);
  bilinear_setup<false>(gx, gy, H, W, b);
}

// one pixel of the forward warp: all C channels of xp (sample base) -> op (sample base + pix); returns channels 0, 1
__device__ __forceinline__ void warp_pixel_fwd(const float* __restrict__ xp, float* __restrict__ op, const Bilinear& b,
                                               int C, int H, int W, size_t hw, float& o0, float& o1) {
  // corner offsets and predicates once per pixel; channels in groups of 4 so that 16 gathers are in flight
  const bool bnw = in_bounds(b.y0, b.x0, H, W), bne = in_bounds(b.y0, b.x0 + 1, H, W);
  const bool bsw = in_bounds(b.y0 + 1, b.x0, H, W), bse = in_bounds(b.y0 + 1, b.x0 + 1, H, W);
  const int o_nw = b.y0 * W + b.x0;
  o0 = 0.f; o1 = 0.f;
  int c = 0;
  for (; c + 4 <= C; c += 4) {
    float v[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* p = xp + (size_t)(c + k) * hw + o_nw;
      v[k][0] = bnw ? __ldg(p) : 0.f; v[k][1] = bne ? __ldg(p + 1) : 0.f;
      v[k][2] = bsw ? __ldg(p + W) : 0.f; v[k][3] = bse ? __ldg(p + W + 1) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float acc = 0.f;
      if (bnw) acc = __fmaf_rn(v[k][0], b.wnw, acc);
      if (bne) acc = __fmaf_rn(v[k][1], b.wne, acc);
      if (bsw) acc = __fmaf_rn(v[k][2], b.wsw, acc);
      if (bse) acc = __fmaf_rn(v[k][3], b.wse, acc);
      __stcs(op + (size_t)(c + k) * hw, acc);
      if (c + k == 0) o0 = acc;
      if (c + k == 1) o1 = acc;
    }
  }
  for (; c < C; ++c) {
    const Corners cr = fetch_corners(xp + (size_t)c * hw, H, W, b);
    const float v = blend(cr, b);
    __stcs(op + (size_t)c * hw, v);
    if (c == 0) o0 = v;
    if (c == 1) o1 = v;
  }
}

// one pixel of the backward warp w.r.t. x: scatter go (sample base + pix) into gxp (sample base) with RED.ADD
__device__ __forceinline__ void warp_pixel_bwd_x(const float* __restrict__ gop, float* __restrict__ gxp, const Bilinear& b,
                                                 int C, int H, int W, size_t hw) {
  const bool bnw = in_bounds(b.y0, b.x0, H, W), bne = in_bounds(b.y0, b.x0 + 1, H, W);
  const bool bsw = in_bounds(b.y0 + 1, b.x0, H, W), bse = in_bounds(b.y0 + 1, b.x0 + 1, H, W);
  const ptrdiff_t o_nw = (ptrdiff_t)b.y0 * W + b.x0;
#pragma unroll 8
  for (int c = 0; c < C; ++c) {
    const float g = ld_stream(gop + (size_t)c * hw);
    float* p = gxp + (size_t)c * hw + o_nw;
    if (bnw) atomicAdd(p, b.wnw * g);
    if (bne) atomicAdd(p + 1, b.wne * g);
    if (bsw) atomicAdd(p + W, b.wsw * g);
    if (bse) atomicAdd(p + W + 1, b.wse * g);
  }
}

__global__ void __launch_bounds__(256) flow_warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ flow,
                                                            float* __restrict__ out, float* __restrict__ fb_mask,
                                                            int32_t* __restrict__ cx0, int32_t* __restrict__ cy0,
                                                            int C, int H, int W, float inv_w, float inv_h, size_t total) {
  const size_t hw = (size_t)H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    Bilinear b;
    float fx, fy;
    flow_bilinear(flow, n, h, w, H, W, inv_w, inv_h, b, fx, fy);
    if (cx0) cx0[idx] = b.x0;
    if (cy0) cy0[idx] = b.y0;
    float o0, o1;
    warp_pixel_fwd(x + n * C * hw, out + n * C * hw + pix, b, C, H, W, hw, o0, o1);
    if (fb_mask) {
      // (f + f_warped)^2 summed < 0.01 * (|f|^2 + |f_warped|^2) + 0.5   (multi_frame_networks.py:205-207)
      const float sx = fadd(fx, o0), sy = fadd(fy, o1);
      const float diff = fadd(fmul(sx, sx), fmul(sy, sy));
      const float mag = fadd(fadd(fmul(fx, fx), fmul(fy, fy)), fadd(fmul(o0, o0), fmul(o1, o1)));
      fb_mask[idx] = (diff < fadd(fmul(0.01f, mag), 0.5f)) ? 1.0f : 0.0f;
    }
  }
}

// Gather step of FuseNet (multi_frame_networks.py:187-214, 347-360) in one launch.  A slot table tells every
// blockIdx.x which frame it reads (src), which slice of the stacked output it owns (dst) and which flow warps it
// (flow == NULL: the own frame, a plain copy).  x is [tl,bs,C,H,W]; out is [n_dst,bs,C,H,W] with n_dst = tl for
// one target frame and tl*tl for all of them (out[tidx*tl + k]).  The slot is the FAST grid dimension and the tables
// list slots of the same source frame next to each other: the CTAs that read (forward) or reduce into (backward) the
// same pixels of a frame run at the same time and meet in L2 instead of each making its own trip to HBM.
constexpr int MAX_TL = 8, MAX_SLOTS = MAX_TL * MAX_TL;
struct GatherArgs {
  const float* flow[MAX_SLOTS];
  short src[MAX_SLOTS];
  short dst[MAX_SLOTS];
};

__global__ void __launch_bounds__(256) flow_warp_gather_fwd_kernel(const float* __restrict__ x, const __grid_constant__ GatherArgs g,
                                                                   float* __restrict__ out, int C, int H, int W,
                                                                   float inv_w, float inv_h, size_t total, size_t slot_stride) {
  const size_t hw = (size_t)H * W;
  const int slot = blockIdx.x;
  const float* flow = g.flow[slot];
  const float* xs = x + (size_t)g.src[slot] * slot_stride;
  float* os = out + (size_t)g.dst[slot] * slot_stride;
  for (size_t idx = (size_t)blockIdx.y * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.y * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    if (!flow) {
#pragma unroll 8
      for (int c = 0; c < C; ++c) __stcs(os + (n * C + c) * hw + pix, ld_stream(xs + (n * C + c) * hw + pix));
    } else {
      Bilinear b;
      float fx, fy, o0, o1;
      flow_bilinear(flow, n, h, w, H, W, inv_w, inv_h, b, fx, fy);
      warp_pixel_fwd(xs + n * C * hw, os + n * C * hw + pix, b, C, H, W, hw, o0, o1);
    }
  }
}

// adjoint: copy slots  gx[src] = go[dst]  (plain stores);  warp slots  gx[src] += scatter(go[dst])  (RED.ADD).
// A copy slot and a warp slot with the same src must not share a launch (the launcher orders them).
__global__ void __launch_bounds__(256) flow_warp_gather_bwd_kernel(const float* __restrict__ go, const __grid_constant__ GatherArgs g,
                                                                   float* __restrict__ gx, int C, int H, int W,
                                                                   float inv_w, float inv_h, size_t total, size_t slot_stride) {
  const size_t hw = (size_t)H * W;
  const int slot = blockIdx.x;
  const float* flow = g.flow[slot];
  const float* gs = go + (size_t)g.dst[slot] * slot_stride;
  float* xs = gx + (size_t)g.src[slot] * slot_stride;
  for (size_t idx = (size_t)blockIdx.y * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.y * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    if (!flow) {
#pragma unroll 8
      for (int c = 0; c < C; ++c) xs[(n * C + c) * hw + pix] = ld_stream(gs + (n * C + c) * hw + pix);
    } else {
      Bilinear b;
      float fx, fy;
      flow_bilinear(flow, n, h, w, H, W, inv_w, inv_h, b, fx, fy);
      warp_pixel_bwd_x(gs + n * C * hw + pix, xs + n * C * hw, b, C, H, W, hw);
    }
  }
}

__global__ void __launch_bounds__(256) flow_warp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ flow,
                                                            const float* __restrict__ go, float* __restrict__ gx,
                                                            float* __restrict__ gflow, int C, int H, int W,
                                                            float inv_w, float inv_h, size_t total) {
  const size_t hw = (size_t)H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    Bilinear b;
    float fx, fy;
    flow_bilinear(flow, n, h, w, H, W, inv_w, inv_h, b, fx, fy);
    const bool bnw = in_bounds(b.y0, b.x0, H, W), bne = in_bounds(b.y0, b.x0 + 1, H, W);
    const bool bsw = in_bounds(b.y0 + 1, b.x0, H, W), bse = in_bounds(b.y0 + 1, b.x0 + 1, H, W);
    const ptrdiff_t o_nw = (ptrdiff_t)b.y0 * W + b.x0;
    float gfx = 0.f, gfy = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float g = ld_stream(go + (n * C + c) * hw + pix);
      if (gx) {
        float* p = gx + (n * C + c) * hw + o_nw;
        if (bnw) atomicAdd(p, b.wnw * g);
        if (bne) atomicAdd(p + 1, b.wne * g);
        if (bsw) atomicAdd(p + W, b.wsw * g);
        if (bse) atomicAdd(p + W + 1, b.wse * g);
      }
      if (gflow) {
        const Corners cr = fetch_corners(x + (n * C + c) * hw, H, W, b);
        gfx = fmaf(blend_dx(cr, b), g, gfx);
        gfy = fmaf(blend_dy(cr, b), g, gfy);
      }
    }
    if (gflow) {
      gflow[(n * 2 + 0) * hw + pix] = ((b.gx_mult * gfx) * 2.0f) * inv_w;
      gflow[(n * 2 + 1) * hw + pix] = ((b.gy_mult * gfy) * 2.0f) * inv_h;
    }
  }
}

inline int flat_grid(size_t total) {
  const size_t want = (total + 255) / 256, cap = 148 * 32;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

// slot table of one target frame: slot 0 = own frame, slot k = k-th other frame (flows[k-1] = flow_{tidx,j_k})
int make_gather_args(const float* const* flows, int tl, int tidx, GatherArgs& g) {
  if (tl < 1 || tl > MAX_TL || tidx < 0 || tidx >= tl) return DIS_ERR_BAD_SHAPE;
  g.src[0] = (short)tidx;
  g.dst[0] = 0;
  g.flow[0] = nullptr;
  for (int j = 0, k = 1; j < tl; ++j) {
    if (j == tidx) continue;
    if (!flows[k - 1]) return DIS_ERR_NULL_POINTER;
    g.src[k] = (short)j;
    g.dst[k] = (short)k;
    g.flow[k] = flows[k - 1];
    ++k;
  }
  return DIS_OK;
}

// slot tables of ALL target frames: flows[i*tl + j] = flow_{ij} (diagonal ignored); copy slots and warp slots in
// separate tables (the backward pass must finish the copies before the reductions start)
int make_gather_all_args(const float* const* flows, int tl, GatherArgs& copies, GatherArgs& warps) {
  if (tl < 1 || tl > MAX_TL) return DIS_ERR_BAD_SHAPE;
  for (int i = 0; i < tl; ++i) {
    copies.src[i] = (short)i;
    copies.dst[i] = (short)(i * tl);
    copies.flow[i] = nullptr;
  }
  int nw = 0;
  for (int j = 0; j < tl; ++j)            // source frame-major: the tl-1 slots that touch frame j sit next to each other
    for (int i = 0; i < tl; ++i) {
      if (i == j) continue;
      if (!flows[i * tl + j]) return DIS_ERR_NULL_POINTER;
      warps.src[nw] = (short)j;
      warps.dst[nw] = (short)(i * tl + (j < i ? j + 1 : j));   // slot k of target i: own frame first, others ascending
      warps.flow[nw] = flows[i * tl + j];
      ++nw;
    }
  return DIS_OK;
}

}  // namespace

int flow_warp_gather_forward(const float* x, const float* const* flows, float* out, int tl, int tidx, int bs, int C,
                             int H, int W, cudaStream_t s) {
  GatherArgs g;
  if (int rc = make_gather_args(flows, tl, tidx, g)) return rc;
  const size_t total = (size_t)bs * H * W;
  const dim3 grid(tl, flat_grid(total));
  flow_warp_gather_fwd_kernel<<<grid, 256, 0, s>>>(x, g, out, C, H, W, 1.0f / (float)(W - 1), 1.0f / (float)(H - 1), total,
                                                   total * C);
  return check_launch();
}

int flow_warp_gather_backward(const float* const* flows, const float* go, float* gx, int tl, int tidx, int bs, int C,
                              int H, int W, cudaStream_t s) {
  GatherArgs g;
  if (int rc = make_gather_args(flows, tl, tidx, g)) return rc;
  const size_t total = (size_t)bs * H * W, stride = total * C;
  // zero the scatter targets (every slice but the own frame's, which is overwritten by a plain copy)
  cudaError_t e = cudaSuccess;
  if (tidx > 0) e = cudaMemsetAsync(gx, 0, sizeof(float) * stride * tidx, s);
  if (e == cudaSuccess && tidx + 1 < tl)
    e = cudaMemsetAsync(gx + stride * (tidx + 1), 0, sizeof(float) * stride * (tl - 1 - tidx), s);
  if (e != cudaSuccess) { set_last_cuda_error(e); return DIS_ERR_CUDA_LAUNCH; }
  const dim3 grid(tl, flat_grid(total));
  flow_warp_gather_bwd_kernel<<<grid, 256, 0, s>>>(go, g, gx, C, H, W, 1.0f / (float)(W - 1), 1.0f / (float)(H - 1), total,
                                                   stride);
  return check_launch();
}

int flow_warp_gather_all_forward(const float* x, const float* const* flows, float* out, int tl, int bs, int C, int H, int W,
                                 cudaStream_t s) {
  GatherArgs copies, warps;
  if (int rc = make_gather_all_args(flows, tl, copies, warps)) return rc;
  const size_t total = (size_t)bs * H * W;
  const float iw = 1.0f / (float)(W - 1), ih = 1.0f / (float)(H - 1);
  flow_warp_gather_fwd_kernel<<<dim3(tl, flat_grid(total)), 256, 0, s>>>(x, copies, out, C, H, W, iw, ih, total, total * C);
  if (int rc = check_launch()) return rc;
  if (tl > 1) {
    flow_warp_gather_fwd_kernel<<<dim3(tl * (tl - 1), flat_grid(total)), 256, 0, s>>>(x, warps, out, C, H, W, iw, ih, total,
                                                                                     total * C);
    return check_launch();
  }
  return DIS_OK;
}

int flow_warp_gather_all_backward(const float* const* flows, const float* go, float* gx, int tl, int bs, int C, int H, int W,
                                  cudaStream_t s) {
  GatherArgs copies, warps;
  if (int rc = make_gather_all_args(flows, tl, copies, warps)) return rc;
  const size_t total = (size_t)bs * H * W;
  const float iw = 1.0f / (float)(W - 1), ih = 1.0f / (float)(H - 1);
  // gx[i] = go[i][0] first (no zero-fill needed), then every warped slot reduces into its source frame
  flow_warp_gather_bwd_kernel<<<dim3(tl, flat_grid(total)), 256, 0, s>>>(go, copies, gx, C, H, W, iw, ih, total, total * C);
  if (int rc = check_launch()) return rc;
  if (tl > 1) {
    flow_warp_gather_bwd_kernel<<<dim3(tl * (tl - 1), flat_grid(total)), 256, 0, s>>>(go, warps, gx, C, H, W, iw, ih, total,
                                                                                     total * C);
    return check_launch();
  }
  return DIS_OK;
}

int flow_warp_forward(const float* x, const float* flow, float* out, float* fb_mask, int32_t* cx0, int32_t* cy0, int N,
                      int C, int H, int W, cudaStream_t s) {
  const size_t total = (size_t)N * H * W;
  flow_warp_fwd_kernel<<<flat_grid(total), 256, 0, s>>>(x, flow, out, fb_mask, cx0, cy0, C, H, W,
                                                        1.0f / (float)(W - 1), 1.0f / (float)(H - 1), total);
  return check_launch();
}

int flow_warp_backward(const float* x, const float* flow, const float* go, float* gx, float* gflow, int N, int C, int H,
                       int W, cudaStream_t s) {
  const size_t total = (size_t)N * H * W;
  if (gx) {
    cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * total * C, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return DIS_ERR_CUDA_LAUNCH; }
  }
  flow_warp_bwd_kernel<<<flat_grid(total), 256, 0, s>>>(x, flow, go, gx, gflow, C, H, W, 1.0f / (float)(W - 1),
                                                        1.0f / (float)(H - 1), total);
  return check_launch();
}

}  // namespace dis
