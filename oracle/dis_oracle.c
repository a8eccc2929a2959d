/*
 * dis_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by, or
 * executed from the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it).
 *
 * Plain-C CPU restatement of the DepthInSpace self-supervision hot path.
 * Every function cites the reference file:line it follows (paths relative to
 * the upstream checkout, idiap/DepthInSpace).
 *
 * Parity status: the block-window photometric loss lives in an un-vendored
 * third-party dependency (autonomousvision/connecting_the_dots, torchext/, no
 * pinned version; reference call sites model/ext_functions.py:124,137).  Its
 * only in-tree definition is photometric_loss_pytorch
 * (model/ext_functions.py:156-183); this file restates THAT function and is
 * pinned against it (tests/test_oracle_pinning.py imports the reference in the
 * build container; the .npz files under tests/golden hold vectors generated from it by
 * oracle/gen_golden.py).  At the ext_cuda boundary itself parity is unpinned
 * (no upstream tests / golden vectors exist).
 *
 * Built twice by oracle/Makefile: -DREAL=float (same op order / roundings as
 * the fp32 CUDA path: contraction is disabled with -ffp-contract=off and an
 * explicit fma is written wherever the ATen CUDA kernel contracts) and
 * -DREAL=double (higher-precision evaluation of the same formulas, used as a
 * tie-breaker when two fp32 evaluations legitimately differ).
 *
 * Coordinate arithmetic follows the torch *CUDA* op sequence (what the
 * reference runs on, model/worker.py:131 'cuda:0'):
 *   - tensor / python-scalar on CUDA is tensor * (1/scalar)
 *     (ATen/native/cuda/BinaryDivTrueKernel.cu, scalar fast path),
 *   - grid_sample un-normalises with ((g + 1) / 2) * (size - 1)
 *     (ATen/native/cuda/GridSampler.cuh:13-22), clips for 'border'
 *     (:41-45), floors, and blends corners nw,ne,sw,se in that order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

#define ORC_API __attribute__((visibility("default")))

static inline real r_fma(real a, real b, real c) {
  return sizeof(real) == 4 ? (real)fmaf((float)a, (float)b, (float)c) : (real)fma(a, b, c);
}
static inline real r_sqrt(real a) { return sizeof(real) == 4 ? (real)sqrtf((float)a) : (real)sqrt(a); }
static inline real r_abs(real a) { return a < 0 ? -a : a; }
static inline real r_floor(real a) { return sizeof(real) == 4 ? (real)floorf((float)a) : (real)floor(a); }
static inline real r_exp(real a) { return sizeof(real) == 4 ? (real)expf((float)a) : (real)exp(a); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline real sgn(real v) { return (real)((v > 0) - (v < 0)); }

ORC_API int orc_sizeof_real(void) { return (int)sizeof(real); }

/* ------------------------------------------------------------------------
 * Local contrast normalisation.  model/networks.py:667-689
 *   box = conv_ones(reflect_pad(x)); mu = box / n; s2 = conv_ones(reflect_pad(x*x))
 *   std = sqrt(max(s2 / n - mu*mu + 1e-6, 0)) + eps ; lcn = (x - mu) / std
 * ---------------------------------------------------------------------- */
static inline int reflect_idx(int i, int n) { /* torch ReflectionPad2d: no edge repeat */
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

ORC_API void orc_lcn_forward(const real* x, real* lcn, real* std_out, int N, int H, int W,
                             int radius, real eps) {
  const int k = 2 * radius + 1;
  const real n = (real)(k * k);
  for (int b = 0; b < N; ++b) {
    const real* xb = x + (size_t)b * H * W;
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        real s1 = 0, s2 = 0;
        for (int dy = -radius; dy <= radius; ++dy) {
          const int yy = reflect_idx(h + dy, H);
          for (int dx = -radius; dx <= radius; ++dx) {
            const real v = xb[(size_t)yy * W + reflect_idx(w + dx, W)];
            s1 += v;
            s2 += v * v;
          }
        }
        const real mu = s1 / n;
        real var = s2 / n - mu * mu + (real)1e-6;
        if (var < 0) var = 0;
        const real sd = r_sqrt(var) + eps;
        const size_t o = (size_t)b * H * W + (size_t)h * W + w;
        std_out[o] = sd;
        lcn[o] = (xb[(size_t)h * W + w] - mu) / sd;
      }
  }
}

/* ------------------------------------------------------------------------
 * Block-window photometric loss.  model/ext_functions.py:156-183
 *   out[n,0,h,w] = (1/k^2) sum_c sum_{dy,dx} f(es, ta), neighbour replicate-clamped
 *   type 0 mse, 1 sad, 2 census_mse, 3 census_sad (model/ext_functions.py:142-154)
 * ---------------------------------------------------------------------- */
static inline real soft_step(real d, real eps) { /* :174-175  0.5*(1 + d/sqrt(d*d+eps)) */
  return (real)0.5 * ((real)1 + d / r_sqrt(d * d + eps));
}

ORC_API int orc_photometric_forward(const real* es, const real* ta, real* out, int N, int C,
                                    int H, int W, int block_size, int type, real eps) {
  if (type < 0 || type > 3) return -1; /* 'invalid loss type' :153,:180 */
  const int p = block_size / 2;
  const real k2 = (real)(block_size * block_size);
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        real acc = 0;
        for (int c = 0; c < C; ++c) {
          const real* e = es + ((size_t)n * C + c) * H * W;
          const real* t = ta + ((size_t)n * C + c) * H * W;
          const real ec = e[(size_t)h * W + w], tc = t[(size_t)h * W + w];
          for (int dy = -p; dy <= p; ++dy) {
            const int yy = clampi(h + dy, 0, H - 1);
            for (int dx = -p; dx <= p; ++dx) {
              const int xx = clampi(w + dx, 0, W - 1);
              const real eq = e[(size_t)yy * W + xx], tq = t[(size_t)yy * W + xx];
              real f;
              if (type == 0) { const real d = eq - tq; f = d * d; }
              else if (type == 1) { f = r_abs(eq - tq); }
              else {
                const real diff = soft_step(eq - ec, eps) - soft_step(tq - tc, eps);
                f = (type == 2) ? diff * diff : r_abs(diff);
              }
              acc += f;
            }
          }
        }
        out[((size_t)n * H + h) * W + w] = acc / k2;
      }
  return 0;
}

/* Gradient w.r.t. es only (model/ext_functions.py:140 returns None for ta):
 * the autograd of photometric_loss_pytorch, written as an explicit scatter.
 * torch.abs has subgradient 0 at 0. */
ORC_API int orc_photometric_backward(const real* es, const real* ta, const real* grad_out,
                                     real* grad_es, int N, int C, int H, int W,
                                     int block_size, int type, real eps) {
  if (type < 0 || type > 3) return -1;
  const int p = block_size / 2;
  const real k2 = (real)(block_size * block_size);
  memset(grad_es, 0, sizeof(real) * (size_t)N * C * H * W);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      const real* e = es + ((size_t)n * C + c) * H * W;
      const real* t = ta + ((size_t)n * C + c) * H * W;
      real* g = grad_es + ((size_t)n * C + c) * H * W;
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          const real wgt = grad_out[((size_t)n * H + h) * W + w] / k2;
          const real ec = e[(size_t)h * W + w], tc = t[(size_t)h * W + w];
          for (int dy = -p; dy <= p; ++dy) {
            const int yy = clampi(h + dy, 0, H - 1);
            for (int dx = -p; dx <= p; ++dx) {
              const int xx = clampi(w + dx, 0, W - 1);
              const size_t q = (size_t)yy * W + xx;
              const real eq = e[q], tq = t[q];
              if (type == 0) { g[q] += wgt * (real)2 * (eq - tq); }
              else if (type == 1) { g[q] += wgt * sgn(eq - tq); }
              else {
                const real de = eq - ec;
                const real xe = de * de + eps;
                const real diff = soft_step(de, eps) - soft_step(tq - tc, eps);
                const real dh = (real)0.5 * eps / (xe * r_sqrt(xe)); /* d soft_step / d de */
                const real df = (type == 2) ? (real)2 * diff : sgn(diff);
                const real v = wgt * df * dh;
                g[q] += v;
                g[(size_t)h * W + w] -= v;
              }
            }
          }
        }
    }
  return 0;
}

/* ------------------------------------------------------------------------
 * Bilinear sampling, torch CUDA grid_sample semantics (align_corners=True).
 * ---------------------------------------------------------------------- */
typedef struct {
  int x0, y0;          /* nw corner index */
  real wnw, wne, wsw, wse;
  real fx, fy;         /* source coordinates after clipping */
  real gx_mult, gy_mult; /* d(source coord)/d(normalised coord) incl. clip gradient */
} bilin_t;

static inline real unnormalize(real g, int size) { /* GridSampler.cuh:13-17 */
  return ((g + (real)1) / (real)2) * (real)(size - 1);
}
static inline real normalize_cuda(real pix, real inv) {
  /* reference: 2 * (pix / (size-1) - 0.5)  (networks.py:363-364, multi_frame_networks.py:95-96)
   * with the CUDA scalar-division fast path pix * (1/(size-1)). */
  return (real)2 * (pix * inv - (real)0.5);
}
static inline real safe_int_range(real x) { /* GridSampler.cuh:137-143 */
  if (x > (real)(INT_MAX - 1) || x < (real)INT_MIN || !isfinite((double)x)) return (real)-100.0;
  return x;
}
/* border: clip + gradient mask (GridSampler.cuh:41-63); zeros: pass through */
static inline real source_index(real g, int size, int border, real* mult) {
  real c = unnormalize(g, size);
  real m = (real)(size - 1) / (real)2;
  if (border) {
    if (c <= (real)0) { c = 0; m = 0; }
    else if (c >= (real)(size - 1)) { c = (real)(size - 1); m = 0; }
  }
  *mult = m;
  return safe_int_range(c);
}
static inline void bilin_setup(real gx, real gy, int H, int W, int border, bilin_t* b) {
  const real ix = source_index(gx, W, border, &b->gx_mult);
  const real iy = source_index(gy, H, border, &b->gy_mult);
  const real fx0 = r_floor(ix), fy0 = r_floor(iy);
  b->x0 = (int)fx0; b->y0 = (int)fy0;
  const real x1 = (real)(b->x0 + 1), y1 = (real)(b->y0 + 1), x0 = (real)b->x0, y0 = (real)b->y0;
  b->wnw = (x1 - ix) * (y1 - iy);
  b->wne = (ix - x0) * (y1 - iy);
  b->wsw = (x1 - ix) * (iy - y0);
  b->wse = (ix - x0) * (iy - y0);
  b->fx = ix; b->fy = iy;
}
static inline int inb(int y, int x, int H, int W) { return y >= 0 && y < H && x >= 0 && x < W; }
static inline real bilin_sample(const real* img, int H, int W, const bilin_t* b) {
  real acc = 0; /* out_acc += v * w, contracted to fma by nvcc in the ATen kernel */
  if (inb(b->y0, b->x0, H, W))         acc = r_fma(img[(size_t)b->y0 * W + b->x0], b->wnw, acc);
  if (inb(b->y0, b->x0 + 1, H, W))     acc = r_fma(img[(size_t)b->y0 * W + b->x0 + 1], b->wne, acc);
  if (inb(b->y0 + 1, b->x0, H, W))     acc = r_fma(img[(size_t)(b->y0 + 1) * W + b->x0], b->wsw, acc);
  if (inb(b->y0 + 1, b->x0 + 1, H, W)) acc = r_fma(img[(size_t)(b->y0 + 1) * W + b->x0 + 1], b->wse, acc);
  return acc;
}
/* d out / d (source x), per unit upstream gradient (grid_sampler_2d_backward) */
static inline real bilin_dx(const real* img, int H, int W, const bilin_t* b) {
  const real y1 = (real)(b->y0 + 1), y0 = (real)b->y0;
  real g = 0;
  if (inb(b->y0, b->x0, H, W))         g -= img[(size_t)b->y0 * W + b->x0] * (y1 - b->fy);
  if (inb(b->y0, b->x0 + 1, H, W))     g += img[(size_t)b->y0 * W + b->x0 + 1] * (y1 - b->fy);
  if (inb(b->y0 + 1, b->x0, H, W))     g -= img[(size_t)(b->y0 + 1) * W + b->x0] * (b->fy - y0);
  if (inb(b->y0 + 1, b->x0 + 1, H, W)) g += img[(size_t)(b->y0 + 1) * W + b->x0 + 1] * (b->fy - y0);
  return g;
}
static inline real bilin_dy(const real* img, int H, int W, const bilin_t* b) {
  const real x1 = (real)(b->x0 + 1), x0 = (real)b->x0;
  real g = 0;
  if (inb(b->y0, b->x0, H, W))         g -= img[(size_t)b->y0 * W + b->x0] * (x1 - b->fx);
  if (inb(b->y0, b->x0 + 1, H, W))     g -= img[(size_t)b->y0 * W + b->x0 + 1] * (b->fx - x0);
  if (inb(b->y0 + 1, b->x0, H, W))     g += img[(size_t)(b->y0 + 1) * W + b->x0] * (x1 - b->fx);
  if (inb(b->y0 + 1, b->x0 + 1, H, W)) g += img[(size_t)(b->y0 + 1) * W + b->x0 + 1] * (b->fx - x0);
  return g;
}

/* ------------------------------------------------------------------------
 * Disparity-driven horizontal pattern warp.  model/networks.py:358-367
 *   x = u - disp ; y = v ; normalise ; grid_sample(pattern, border, align_corners)
 * Also returns corner index x0,y0 (int32) for bit-exact index checks and
 * d pattern_proj / d disp (the autograd chain of :358-367).
 * ---------------------------------------------------------------------- */
ORC_API void orc_pattern_warp(const real* disp, const real* pattern, real* proj, real* dproj_ddisp,
                              int32_t* ix0, int32_t* iy0, int N, int H, int W) {
  const real inv_w = (real)((float)1.0f / (float)(W - 1));
  const real inv_h = (real)((float)1.0f / (float)(H - 1));
  const real invw = sizeof(real) == 4 ? inv_w : (real)1 / (real)(W - 1);
  const real invh = sizeof(real) == 4 ? inv_h : (real)1 / (real)(H - 1);
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const size_t o = ((size_t)n * H + h) * W + w;
        const real px = (real)w - disp[o];
        const real gx = normalize_cuda(px, invw);
        const real gy = normalize_cuda((real)h, invh);
        bilin_t b;
        bilin_setup(gx, gy, H, W, 1, &b);
        proj[o] = bilin_sample(pattern, H, W, &b);
        if (ix0) ix0[o] = b.x0;
        if (iy0) iy0[o] = b.y0;
        if (dproj_ddisp) {
          /* grid grad = gx_mult * gix; then *2 (networks.py:363), *inv_w, then d(u-disp)/ddisp = -1 */
          const real gix = bilin_dx(pattern, H, W, &b);
          dproj_ddisp[o] = -(((b.gx_mult * gix) * (real)2) * invw);
        }
      }
}

/* ------------------------------------------------------------------------
 * Optical-flow warp.  model/multi_frame_networks.py:83-99
 *   (u + fx, v + fy) -> normalise -> grid_sample(x, zeros, align_corners)
 * backward: grad w.r.t. x (scatter, GridSampler.cuh:250-264 safe_add_2d) and,
 * optionally, w.r.t. flow.
 * ---------------------------------------------------------------------- */
static inline void flow_setup(const real* flow, int n, int h, int w, int H, int W, real invw,
                              real invh, bilin_t* b) {
  const size_t hw = (size_t)H * W;
  const real fx = flow[((size_t)n * 2 + 0) * hw + (size_t)h * W + w];
  const real fy = flow[((size_t)n * 2 + 1) * hw + (size_t)h * W + w];
  const real gx = normalize_cuda(fx + (real)w, invw);
  const real gy = normalize_cuda(fy + (real)h, invh);
  bilin_setup(gx, gy, H, W, 0, b);
}

ORC_API void orc_flow_warp_forward(const real* x, const real* flow, real* out, int32_t* ix0,
                                   int32_t* iy0, int N, int C, int H, int W) {
  const real invw = sizeof(real) == 4 ? (real)((float)1.0f / (float)(W - 1)) : (real)1 / (real)(W - 1);
  const real invh = sizeof(real) == 4 ? (real)((float)1.0f / (float)(H - 1)) : (real)1 / (real)(H - 1);
  const size_t hw = (size_t)H * W;
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        bilin_t b;
        flow_setup(flow, n, h, w, H, W, invw, invh, &b);
        if (ix0) ix0[(size_t)n * hw + (size_t)h * W + w] = b.x0;
        if (iy0) iy0[(size_t)n * hw + (size_t)h * W + w] = b.y0;
        for (int c = 0; c < C; ++c)
          out[((size_t)n * C + c) * hw + (size_t)h * W + w] =
              bilin_sample(x + ((size_t)n * C + c) * hw, H, W, &b);
      }
}

ORC_API void orc_flow_warp_backward(const real* x, const real* flow, const real* grad_out,
                                    real* grad_x, real* grad_flow, int N, int C, int H, int W) {
  const real invw = sizeof(real) == 4 ? (real)((float)1.0f / (float)(W - 1)) : (real)1 / (real)(W - 1);
  const real invh = sizeof(real) == 4 ? (real)((float)1.0f / (float)(H - 1)) : (real)1 / (real)(H - 1);
  const size_t hw = (size_t)H * W;
  if (grad_x) memset(grad_x, 0, sizeof(real) * (size_t)N * C * hw);
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        bilin_t b;
        flow_setup(flow, n, h, w, H, W, invw, invh, &b);
        real gfx = 0, gfy = 0;
        for (int c = 0; c < C; ++c) {
          const real go = grad_out[((size_t)n * C + c) * hw + (size_t)h * W + w];
          if (grad_x) {
            real* gx = grad_x + ((size_t)n * C + c) * hw;
            if (inb(b.y0, b.x0, H, W))         gx[(size_t)b.y0 * W + b.x0] += b.wnw * go;
            if (inb(b.y0, b.x0 + 1, H, W))     gx[(size_t)b.y0 * W + b.x0 + 1] += b.wne * go;
            if (inb(b.y0 + 1, b.x0, H, W))     gx[(size_t)(b.y0 + 1) * W + b.x0] += b.wsw * go;
            if (inb(b.y0 + 1, b.x0 + 1, H, W)) gx[(size_t)(b.y0 + 1) * W + b.x0 + 1] += b.wse * go;
          }
          if (grad_flow) {
            const real* xc = x + ((size_t)n * C + c) * hw;
            gfx += bilin_dx(xc, H, W, &b) * go;
            gfy += bilin_dy(xc, H, W, &b) * go;
          }
        }
        if (grad_flow) {
          grad_flow[((size_t)n * 2 + 0) * hw + (size_t)h * W + w] = ((b.gx_mult * gfx) * (real)2) * invw;
          grad_flow[((size_t)n * 2 + 1) * hw + (size_t)h * W + w] = ((b.gy_mult * gfy) * (real)2) * invh;
        }
      }
}

/* ------------------------------------------------------------------------
 * Sobel 5x5 / 3x3 on a replicate-padded image.  model/networks.py:697-730
 * out [N,2,H,W] = (gx, gy); weights are float32-rounded (torch .float()).
 * ---------------------------------------------------------------------- */
static void sobel_weights(int ksize, real* kx /* ksize*ksize, row-major [i][j] */) {
  static const double k5[25] = {-5, -4, 0, 4, 5, -8, -10, 0, 10, 8, -10, -20, 0, 20, 10,
                                -8, -10, 0, 10, 8, -5, -4, 0, 4, 5};
  static const double k3[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};
  if (ksize == 5) for (int i = 0; i < 25; ++i) kx[i] = (real)(float)(k5[i] / 240.0);
  else            for (int i = 0; i < 9; ++i)  kx[i] = (real)(float)(k3[i] / 8.0);
}

ORC_API int orc_sobel_forward(const real* x, real* out, int N, int H, int W, int ksize) {
  if (ksize != 5 && ksize != 3) return -1;
  real kx[25];
  sobel_weights(ksize, kx);
  const int r = ksize / 2;
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const real* xb = x + (size_t)n * H * W;
        real gx = 0, gy = 0;
        for (int i = 0; i < ksize; ++i)
          for (int j = 0; j < ksize; ++j) {
            const real v = xb[(size_t)clampi(h + i - r, 0, H - 1) * W + clampi(w + j - r, 0, W - 1)];
            gx += kx[i * ksize + j] * v;
            gy += kx[j * ksize + i] * v; /* ky = kx^T  (:707) */
          }
        out[(((size_t)n * 2 + 0) * H + h) * W + w] = gx;
        out[(((size_t)n * 2 + 1) * H + h) * W + w] = gy;
      }
  return 0;
}

/* adjoint of orc_sobel_forward: grad_out [N,2,H,W] -> grad_x [N,1,H,W] */
ORC_API int orc_sobel_backward(const real* grad_out, real* grad_x, int N, int H, int W, int ksize) {
  if (ksize != 5 && ksize != 3) return -1;
  real kx[25];
  sobel_weights(ksize, kx);
  const int r = ksize / 2;
  memset(grad_x, 0, sizeof(real) * (size_t)N * H * W);
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const real ux = grad_out[(((size_t)n * 2 + 0) * H + h) * W + w];
        const real uy = grad_out[(((size_t)n * 2 + 1) * H + h) * W + w];
        real* g = grad_x + (size_t)n * H * W;
        for (int i = 0; i < ksize; ++i)
          for (int j = 0; j < ksize; ++j)
            g[(size_t)clampi(h + i - r, 0, H - 1) * W + clampi(w + j - r, 0, W - 1)] +=
                kx[i * ksize + j] * ux + kx[j * ksize + i] * uy;
      }
  return 0;
}

/* ------------------------------------------------------------------------
 * Edge-aware disparity smoothness.  model/networks.py:419-431
 *   val = mean | sobel(disp) * exp(-|255 * sobel(im)|) |   over [N,2,H,W]
 * Returns val; if grad_disp != NULL also d val / d disp.
 * ---------------------------------------------------------------------- */
ORC_API double orc_smooth_loss(const real* disp, const real* im, real* grad_disp, int N, int H, int W) {
  const size_t M = (size_t)N * 2 * H * W;
  real* gd = (real*)malloc(sizeof(real) * M);
  real* gi = (real*)malloc(sizeof(real) * M);
  orc_sobel_forward(disp, gd, N, H, W, 5);
  orc_sobel_forward(im, gi, N, H, W, 5);
  double total = 0;
  for (size_t i = 0; i < M; ++i) {
    const real a = r_exp(-r_abs((real)255 * gi[i]));
    const real v = gd[i] * a;
    total += (double)r_abs(v);
    gi[i] = sgn(v) * a / (real)M; /* d val / d gd */
  }
  if (grad_disp) orc_sobel_backward(gi, grad_disp, N, H, W, 5);
  free(gd);
  free(gi);
  return total / (double)M;
}

/* ------------------------------------------------------------------------
 * RectifiedPatternSimilarityLoss.tforward.  model/networks.py:354-377
 *   val = sum(mask * diff) / sum(mask), mask = std (or ones when std==NULL)
 * Outputs: *num, *den (double), optional per-pixel diff, pattern_proj, and
 * d val / d disp (for upstream gradient 1).
 * ---------------------------------------------------------------------- */
ORC_API int orc_pattern_loss(const real* disp, const real* im, const real* std_in, const real* pattern,
                             real* proj_out, real* diff_out, real* grad_disp, double* num, double* den,
                             int N, int H, int W, int block_size, int type, real eps) {
  const size_t M = (size_t)N * H * W;
  real* proj = proj_out ? proj_out : (real*)malloc(sizeof(real) * M);
  real* dpd = (real*)malloc(sizeof(real) * M);
  real* diff = diff_out ? diff_out : (real*)malloc(sizeof(real) * M);
  orc_pattern_warp(disp, pattern, proj, dpd, NULL, NULL, N, H, W);
  int rc = orc_photometric_forward(proj, im, diff, N, 1, H, W, block_size, type, eps);
  if (rc == 0) {
    double sn = 0, sd = 0;
    for (size_t i = 0; i < M; ++i) {
      const real m = std_in ? std_in[i] : (real)1;
      sn += (double)(m * diff[i]);
      sd += (double)m;
    }
    *num = sn; *den = sd;
    if (grad_disp) {
      real* go = (real*)malloc(sizeof(real) * M);
      real* ge = (real*)malloc(sizeof(real) * M);
      for (size_t i = 0; i < M; ++i) go[i] = (std_in ? std_in[i] : (real)1) / (real)sd;
      orc_photometric_backward(proj, im, go, ge, N, 1, H, W, block_size, type, eps);
      for (size_t i = 0; i < M; ++i) grad_disp[i] = ge[i] * dpd[i];
      free(go); free(ge);
    }
  }
  if (!proj_out) free(proj);
  if (!diff_out) free(diff);
  free(dpd);
  return rc;
}

/* ------------------------------------------------------------------------
 * Flow-consistency (geometric) loss, one direction.  model/networks.py:619-655 (single frame),
 * :564-601 (multi frame, adds the reprojection mask), ProjectionBaseLoss :455-488.
 * Row-vector conventions of the reference: bmm(xyz, R0), bmm(xyz, R1^T), bmm(xyz, K^T).
 * Outputs: *num = sum(diff*mask), *den = sum(mask); optional mask, orig_mask, and the gradients of
 * num/(den+1e-8) w.r.t. depth0 (direct) and depth1 (scatter through the bilinear sampling).
 * ---------------------------------------------------------------------- */
static void project_view(const real* R0, const real* t0, const real* R1, const real* t1, const real* K,
                         real depth, const real* ray, real out[3]) {
  real c0[3], w[3], c1[3];
  for (int j = 0; j < 3; ++j) c0[j] = depth * ray[j] - t0[j];
  for (int j = 0; j < 3; ++j) w[j] = c0[0] * R0[j] + c0[1] * R0[3 + j] + c0[2] * R0[6 + j];
  for (int j = 0; j < 3; ++j) c1[j] = w[0] * R1[3 * j] + w[1] * R1[3 * j + 1] + w[2] * R1[3 * j + 2] + t1[j];
  for (int j = 0; j < 3; ++j) out[j] = c1[0] * K[3 * j] + c1[1] * K[3 * j + 1] + c1[2] * K[3 * j + 2];
}

ORC_API void orc_flow_consistency(const real* depth0, const real* depth1, const real* R0, const real* t0,
                                  const real* R1, const real* t1, const real* flow0, const real* flow1,
                                  const real* amb0, const real* amb1, int amb_c, const real* primary_depth1,
                                  const real* K, const real* ray, real clamp, real fb_scale, real* mask_out,
                                  real* orig_mask, real* grad_depth0, real* grad_depth1, double* num, double* den,
                                  int bs, int H, int W) {
  const real invw = sizeof(real) == 4 ? (real)((float)1.0f / (float)(W - 1)) : (real)1 / (real)(W - 1);
  const real invh = sizeof(real) == 4 ? (real)((float)1.0f / (float)(H - 1)) : (real)1 / (real)(H - 1);
  const size_t hw = (size_t)H * W;
  double sn = 0, sd = 0;
  real* g = (real*)malloc(sizeof(real) * bs * hw);   /* d(diff*mask)/d d1 */
  for (int n = 0; n < bs; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const size_t pix = (size_t)h * W + w, o = n * hw + pix;
        bilin_t b;
        flow_setup(flow0, n, h, w, H, W, invw, invh, &b);
        real uvd[3];
        project_view(R0 + 9 * n, t0 + 3 * n, R1 + 9 * n, t1 + 3 * n, K, depth0[o], ray + 3 * pix, uvd);
        const real depth10 = bilin_sample(depth1 + n * hw, H, W, &b);
        const real raw = uvd[2] - depth10;
        real diff = r_abs(raw);
        const int clamped = clamp > 0 && diff > clamp;
        if (clamp > 0 && diff > clamp) diff = clamp;
        if (orig_mask) orig_mask[o] = diff < clamp ? 1 : 0;
        const real fx = flow0[(n * 2 + 0) * hw + pix], fy = flow0[(n * 2 + 1) * hw + pix];
        const real f10x = bilin_sample(flow1 + (n * 2 + 0) * hw, H, W, &b);
        const real f10y = bilin_sample(flow1 + (n * 2 + 1) * hw, H, W, &b);
        const real sx = fx + f10x, sy = fy + f10y;
        real m = (sx * sx + sy * sy) < (real)0.5 + fb_scale * ((fx * fx + fy * fy) + (f10x * f10x + f10y * f10y)) ? 1 : 0;
        real vc = 0;
        for (int c = 0; c < amb_c; ++c)
          vc += r_abs(amb0[((size_t)n * amb_c + c) * hw + pix] - bilin_sample(amb1 + ((size_t)n * amb_c + c) * hw, H, W, &b));
        if (amb_c > 1) vc = vc / (real)amb_c;
        m *= vc < (real)0.01 ? 1 : 0;
        if (primary_depth1) {
          real wu = 0, wv = 0;
          const int cx[4] = {b.x0, b.x0 + 1, b.x0, b.x0 + 1}, cy[4] = {b.y0, b.y0, b.y0 + 1, b.y0 + 1};
          const real wg[4] = {b.wnw, b.wne, b.wsw, b.wse};
          for (int k = 0; k < 4; ++k) {
            if (!inb(cy[k], cx[k], H, W)) continue;
            const size_t q = (size_t)cy[k] * W + cx[k];
            real p3[3];
            project_view(R1 + 9 * n, t1 + 3 * n, R0 + 9 * n, t0 + 3 * n, K, primary_depth1[n * hw + q], ray + 3 * q, p3);
            const real dz = (p3[2] > 0 ? p3[2] : 0) + (real)1e-12;
            wu = r_fma(p3[0] / dz, wg[k], wu);
            wv = r_fma(p3[1] / dz, wg[k], wv);
          }
          const real du = wu - (real)w, dv = wv - (real)h;
          m *= (du * du + dv * dv) < (real)1 ? 1 : 0;
        }
        if (mask_out) mask_out[o] = m;
        sn += (double)(diff * m);
        sd += (double)m;
        g[o] = (clamped ? 0 : sgn(raw)) * m;
      }
  *num = sn;
  *den = sd;
  if (grad_depth0 || grad_depth1) {
    const real scale = (real)(1.0 / (sd + 1e-8));
    if (grad_depth1) memset(grad_depth1, 0, sizeof(real) * bs * hw);
    for (int n = 0; n < bs; ++n)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          const size_t pix = (size_t)h * W + w, o = n * hw + pix;
          if (grad_depth0) {
            real a[3], z[3] = {0, 0, 0}, b3[3];
            project_view(R0 + 9 * n, z, R1 + 9 * n, z, K, (real)1, ray + 3 * pix, a);   /* linear part: d d1/d depth0 */
            (void)b3;
            grad_depth0[o] = g[o] * a[2] * scale;
          }
          if (grad_depth1 && g[o] != 0) {
            bilin_t b;
            flow_setup(flow0, n, h, w, H, W, invw, invh, &b);
            real* gd = grad_depth1 + n * hw;
            const real v = -g[o] * scale;
            if (inb(b.y0, b.x0, H, W))         gd[(size_t)b.y0 * W + b.x0] += b.wnw * v;
            if (inb(b.y0, b.x0 + 1, H, W))     gd[(size_t)b.y0 * W + b.x0 + 1] += b.wne * v;
            if (inb(b.y0 + 1, b.x0, H, W))     gd[(size_t)(b.y0 + 1) * W + b.x0] += b.wsw * v;
            if (inb(b.y0 + 1, b.x0 + 1, H, W)) gd[(size_t)(b.y0 + 1) * W + b.x0 + 1] += b.wse * v;
          }
        }
  }
  free(g);
}
