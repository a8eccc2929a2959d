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
#include <stdlib.h>
#include <string.h>
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
  const int o_nw = clampi(b.y0, -1, H) * W + clampi(b.x0, -1, W);   // (clamped: no overflow; unchanged where a corner is in bounds)
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

// ---- all-frames gather, backward, as a tile-local GATHER (deterministic, no global atomics) -------------------------------
// d/dx of gather_warped_all: gx[j] = go[j][own slot] + sum over the tl-1 frames i that sample frame j of
// warp^T(go[i][slot of j]).  The transposed warp is a scatter; here it is inverted per 32 x 8 tile of frame j:
//   1. every pixel p of the tile's neighbourhood (halo GT_F) evaluates its sampling position once (channel independent)
//      and appends (p, weight) to the shared-memory list of each of its <= 4 target cells q inside the tile
//      (integer shared-memory atomics only hand out list slots; the lists are then sorted by p: the order of the sum,
//      and with it the result, is fixed);
//   2. every thread owns one q and walks its lists for all C channels: plain coalesced store, no RED.
// The neighbourhood of a tile is displaced by the reverse flow at the tile centre (flow_ji ~ inverse of flow_ij); a (p, q)
// pair whose p falls outside it (inconsistent flows, occlusion edges) is left to a second, scatter-style launch
// (flow_gather_far_kernel), ordered after this one; a tile whose lists overflow (> GT_K pixels landing in one cell of one
// slot: folds of the flow field) switches to tile-local reductions for that tile only.
constexpr int GT_X = 32, GT_Y = 8, GT_F = 5, GT_K = 8, GT_S = 3;
constexpr int GT_RW = GT_X + 2 * GT_F, GT_RH = GT_Y + 2 * GT_F, GT_CH = 8;

struct GatherTileArgs {
  const float* flow[MAX_TL][MAX_TL - 1];   // [source frame j][incoming slot]: flow_{ij}, sampled at frame i's pixels
  const float* rflow[MAX_TL][MAX_TL - 1];  // flow_{ji}: where the cells of frame j come from in frame i (approximate inverse)
  short dst[MAX_TL][MAX_TL - 1];           // slice of go that holds warp(x[j], flow_ij)
  short copy_dst[MAX_TL];                  // slice of go that holds the own frame
  int tl;
};


// Where the pixels that land in tile (tx0, ty0) of frame j come from: the tile displaced by the reverse flow at its centre
// (rounded; clamped so that garbage flows cannot overflow the index arithmetic).  Both launches derive it the same way.
__device__ __forceinline__ void gt_shift(const float* __restrict__ rflow, size_t n, int tx0, int ty0, int H, int W, int& sx, int& sy) {
  const size_t hw = (size_t)H * W;
  const int cx = min(tx0 + GT_X / 2, W - 1), cy = min(ty0 + GT_Y / 2, H - 1);
  const float fx = __ldg(rflow + (n * 2 + 0) * hw + (size_t)cy * W + cx);
  const float fy = __ldg(rflow + (n * 2 + 1) * hw + (size_t)cy * W + cx);
  sx = (fx > -4096.0f && fx < 4096.0f) ? __float2int_rn(fx) : 0;     // (NaN compares false: shift 0)
  sy = (fy > -4096.0f && fy < 4096.0f) ? __float2int_rn(fy) : 0;
}
// is pixel (px, py) inside the candidate region of that tile?
__device__ __forceinline__ bool gt_in_region(int px, int py, int tx0, int ty0, int sx, int sy) {
  const int rx = px - (tx0 + sx - GT_F), ry = py - (ty0 + sy - GT_F);
  return rx >= 0 && rx < GT_RW && ry >= 0 && ry < GT_RH;
}

__global__ void __launch_bounds__(256) flow_gather_all_bwd_tile_kernel(const float* __restrict__ go, const __grid_constant__ GatherTileArgs g,
                                                                       float* __restrict__ gx, int bs, int C, int H, int W,
                                                                       float inv_w, float inv_h, size_t slot_stride) {
  __shared__ unsigned short lcand[GT_S][GT_K][GT_X * GT_Y];   // index of p in the tile's candidate region
  __shared__ float lw[GT_S][GT_K][GT_X * GT_Y];               // its bilinear weight
  __shared__ int cnt[GT_S][GT_X * GT_Y];
  __shared__ int overflow;
  __shared__ int shift[GT_S][2];
  const int tid = threadIdx.x;
  const int x0t = blockIdx.x * GT_X, y0t = blockIdx.y * GT_Y;
  const int j = blockIdx.z / bs, n = blockIdx.z - j * bs;
  const size_t hw = (size_t)H * W;
  const int qlx = tid & (GT_X - 1), qly = tid >> 5;
  const int qx = x0t + qlx, qy = y0t + qly;
  const bool q_ok = qx < W && qy < H;
  const size_t qoff = (size_t)qy * W + qx;
  float* gxs = gx + (size_t)j * slot_stride + (size_t)n * C * hw;
  const int nslots = g.tl - 1;

  for (int s0 = 0; s0 < nslots || s0 == 0; s0 += GT_S) {
    const int ns = min(GT_S, nslots - s0);
    // ---- 1. lists ---------------------------------------------------------------------------------------------------
    if (tid == 0) overflow = 0;
    if (tid < ns) gt_shift(g.rflow[j][s0 + tid], n, x0t, y0t, H, W, shift[tid][0], shift[tid][1]);
    for (int i = tid; i < GT_S * GT_X * GT_Y; i += 256) (&cnt[0][0])[i] = 0;
    __syncthreads();
    for (int s = 0; s < ns; ++s) {
      const float* flow = g.flow[j][s0 + s];
      const int rx0 = x0t + shift[s][0] - GT_F, ry0 = y0t + shift[s][1] - GT_F;    // origin of the candidate region
      for (int cand = tid; cand < GT_RW * GT_RH; cand += 256) {
        const int ly = cand / GT_RW, lx = cand - ly * GT_RW;
        const int px = rx0 + lx, py = ry0 + ly;
        if (px < 0 || px >= W || py < 0 || py >= H) continue;
        Bilinear b;
        float fx, fy;
        flow_bilinear(flow, n, py, px, H, W, inv_w, inv_h, b, fx, fy);
        const float wgt[4] = {b.wnw, b.wne, b.wsw, b.wse};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int cx = b.x0 + (k & 1), cy = b.y0 + (k >> 1);
          const int tx = cx - x0t, ty = cy - y0t;
          if (tx < 0 || tx >= GT_X || ty < 0 || ty >= GT_Y || cx >= W || cy >= H) continue;   // not a cell of this tile
          if (wgt[k] == 0.0f) continue;                                   // zero weights add nothing
          const int q = ty * GT_X + tx;
          const int slot = atomicAdd(&cnt[s][q], 1);
          if (slot < GT_K) {
            lcand[s][slot][q] = (unsigned short)cand;     // [slot][q]: a warp's 32 cells hit 32 different banks
            lw[s][slot][q] = wgt[k];
          } else {
            overflow = 1;
          }
        }
      }
    }
    __syncthreads();
    const bool first = (s0 == 0);
    if (!overflow) {
      // ---- 2a. gather: sort my lists by candidate index (fixed summation order), then all channels ---------------------
      if (q_ok) {
        for (int s = 0; s < ns; ++s) {
          const int m = cnt[s][tid];
          for (int a = 1; a < m; ++a) {
            const unsigned short ec = lcand[s][a][tid];
            const float ev = lw[s][a][tid];
            int b2 = a - 1;
            while (b2 >= 0 && lcand[s][b2][tid] > ec) {
              lcand[s][b2 + 1][tid] = lcand[s][b2][tid];
              lw[s][b2 + 1][tid] = lw[s][b2][tid];
              --b2;
            }
            lcand[s][b2 + 1][tid] = ec;
            lw[s][b2 + 1][tid] = ev;
          }
        }
        const float* own = go + (size_t)g.copy_dst[j] * slot_stride + (size_t)n * C * hw + qoff;
        for (int c0 = 0; c0 < C; c0 += GT_CH) {
          float acc[GT_CH];
#pragma unroll
          for (int c = 0; c < GT_CH; ++c)
            acc[c] = (c0 + c < C) ? (first ? ld_stream(own + (size_t)(c0 + c) * hw) : gxs[(size_t)(c0 + c) * hw + qoff]) : 0.0f;
          for (int s = 0; s < ns; ++s) {
            const float* gs = go + (size_t)g.dst[j][s0 + s] * slot_stride + ((size_t)n * C + c0) * hw;
            const int rx0 = x0t + shift[s][0] - GT_F, ry0 = y0t + shift[s][1] - GT_F;
            const int m = cnt[s][tid];
            for (int e = 0; e < m; ++e) {
              const int ecand = lcand[s][e][tid];
              const float ew = lw[s][e][tid];
              const int ly = ecand / GT_RW, lx = ecand - ly * GT_RW;
              const float* pp = gs + (size_t)(ry0 + ly) * W + (rx0 + lx);
#pragma unroll
              for (int c = 0; c < GT_CH; ++c)
                if (c0 + c < C) acc[c] = fmaf(ew, __ldg(pp + (size_t)c * hw), acc[c]);
            }
          }
#pragma unroll
          for (int c = 0; c < GT_CH; ++c)
            if (c0 + c < C) gxs[(size_t)(c0 + c) * hw + qoff] = acc[c];
        }
      }
    } else {
      // ---- 2b. a list overflowed: this tile only falls back to reductions into its own cells -----------------------------
      if (first && q_ok) {
        const float* own = go + (size_t)g.copy_dst[j] * slot_stride + (size_t)n * C * hw + qoff;
        for (int c = 0; c < C; ++c) gxs[(size_t)c * hw + qoff] = ld_stream(own + (size_t)c * hw);
      }
      __syncthreads();
      for (int s = 0; s < ns; ++s) {
        const float* flow = g.flow[j][s0 + s];
        const float* gs = go + (size_t)g.dst[j][s0 + s] * slot_stride + (size_t)n * C * hw;
        const int rx0 = x0t + shift[s][0] - GT_F, ry0 = y0t + shift[s][1] - GT_F;
        for (int cand = tid; cand < GT_RW * GT_RH; cand += 256) {
          const int ly = cand / GT_RW, lx = cand - ly * GT_RW;
          const int px = rx0 + lx, py = ry0 + ly;
          if (px < 0 || px >= W || py < 0 || py >= H) continue;
          Bilinear b;
          float fx, fy;
          flow_bilinear(flow, n, py, px, H, W, inv_w, inv_h, b, fx, fy);
          const float wgt[4] = {b.wnw, b.wne, b.wsw, b.wse};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cx = b.x0 + (k & 1), cy = b.y0 + (k >> 1);
            const int tx = cx - x0t, ty = cy - y0t;
            if (tx < 0 || tx >= GT_X || ty < 0 || ty >= GT_Y || cx >= W || cy >= H) continue;
            if (wgt[k] == 0.0f) continue;
            for (int c = 0; c < C; ++c)
              atomicAdd(gxs + (size_t)c * hw + (size_t)cy * W + cx, wgt[k] * __ldg(gs + (size_t)c * hw + (size_t)py * W + px));
          }
        }
      }
    }
    __syncthreads();
    if (nslots == 0) break;
  }
}

// pairs (p, q) whose p lies outside the candidate region of q's tile (the flow deviates from the reverse flow at the tile
// centre by more than GT_F pixels: occlusion edges, inconsistent flows): reductions, after the tile kernel has written every cell
struct FarArgs {
  const float* flow[MAX_SLOTS];
  const float* rflow[MAX_SLOTS];
  short src[MAX_SLOTS];
  short dst[MAX_SLOTS];
};
__global__ void __launch_bounds__(256) flow_gather_far_kernel(const float* __restrict__ go, const __grid_constant__ FarArgs g,
                                                              float* __restrict__ gx, int C, int H, int W, float inv_w,
                                                              float inv_h, size_t total, size_t slot_stride) {
  const size_t hw = (size_t)H * W;
  const int slot = blockIdx.x;
  const float* flow = g.flow[slot];
  const float* gs = go + (size_t)g.dst[slot] * slot_stride;
  float* xs = gx + (size_t)g.src[slot] * slot_stride;
  for (size_t idx = (size_t)blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.y * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    Bilinear b;
    float fx, fy;
    flow_bilinear(flow, n, h, w, H, W, inv_w, inv_h, b, fx, fy);
    const float wgt[4] = {b.wnw, b.wne, b.wsw, b.wse};
    int last_tx = INT_MIN, last_ty = INT_MIN, sx = 0, sy = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cx = b.x0 + (k & 1), cy = b.y0 + (k >> 1);
      if (!in_bounds(cy, cx, H, W) || wgt[k] == 0.0f) continue;
      const int tx0 = (cx / GT_X) * GT_X, ty0 = (cy / GT_Y) * GT_Y;
      if (tx0 != last_tx || ty0 != last_ty) {
        gt_shift(g.rflow[slot], n, tx0, ty0, H, W, sx, sy);
        last_tx = tx0;
        last_ty = ty0;
      }
      if (gt_in_region(w, h, tx0, ty0, sx, sy)) continue;            // the tile kernel has it
      for (int c = 0; c < C; ++c)
        atomicAdd(xs + (n * C + c) * hw + (size_t)cy * W + cx, wgt[k] * ld_stream(gs + (n * C + c) * hw + pix));
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
  // DIS_GATHER_BWD=tile selects the tile-local gather formulation (deterministic where the flows are consistent).  Measured
  // at tl 4, bs 32, C 32, 256x216 it LOSES to the reductions (4.2 vs 3.4 ms): 6 200 instructions per cell (list building
  // ~1 800, list walk ~2 000 per 8-channel chunk pass) make it issue-bound at 67 % -- kept as an experiment, not the default.
  const char* impl = getenv("DIS_GATHER_BWD");
  if (impl && strcmp(impl, "tile") == 0 && (long)tl * bs <= 65535) {
    GatherTileArgs t{};
    t.tl = tl;
    for (int j = 0; j < tl; ++j) {
      t.copy_dst[j] = (short)(j * tl);
      int k = 0;
      for (int i = 0; i < tl; ++i) {
        if (i == j) continue;
        t.flow[j][k] = flows[i * tl + j];
        t.rflow[j][k] = flows[j * tl + i];
        t.dst[j][k] = (short)(i * tl + (j < i ? j + 1 : j));
        ++k;
      }
    }
    FarArgs far{};
    for (int k = 0; k < tl * (tl - 1); ++k) {
      far.flow[k] = warps.flow[k];
      far.src[k] = warps.src[k];
      far.dst[k] = warps.dst[k];
    }
    for (int j = 0, k = 0; j < tl; ++j)       // same order as make_gather_all_args: source frame major
      for (int i = 0; i < tl; ++i)
        if (i != j) far.rflow[k++] = flows[j * tl + i];
    const dim3 grid((W + GT_X - 1) / GT_X, (H + GT_Y - 1) / GT_Y, tl * bs);
    flow_gather_all_bwd_tile_kernel<<<grid, 256, 0, s>>>(go, t, gx, bs, C, H, W, iw, ih, total * C);
    if (int rc = check_launch()) return rc;
    if (tl > 1) {
      flow_gather_far_kernel<<<dim3(tl * (tl - 1), flat_grid(total)), 256, 0, s>>>(go, far, gx, C, H, W, iw, ih, total, total * C);
      return check_launch();
    }
    return DIS_OK;
  }
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
