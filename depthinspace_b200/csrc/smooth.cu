// Sobel filter and edge-aware disparity smoothness.
// reference: SobelFilter (model/networks.py:697-730), DisparitySmoothLoss.tforward (:419-431)
//   g  = sobel5x5(replicate_pad2(disp)) -> (gx, gy);  gi = sobel5x5(replicate_pad2(im))
//   val = mean | g * exp(-|255 gi|) |   over [N,2,H,W]
//
// smooth_march_kernel does the whole forward and, on request, the exact gradient w.r.t. disp in one pass.  The 5x5 Sobel
// pair is rank 2:  kx = a (x) [-1 0 0 0 1] + b (x) [0 -1 0 1 0]  with the vertical smoothers a = (5 8 10 8 5) / 240 and
// b = (4 10 20 10 4) / 240, ky = kx^T, so
//   gx(y, x) = VA(y, x+2) - VA(y, x-2) + VB(y, x+1) - VB(y, x-1),   VA = a *_vertical v,  VB = b *_vertical v
//   gy(y, x) = HA(y+2, x) - HA(y-2, x) + HB(y+1, x) - HB(y-1, x),   HA = a *_horizontal v, HB = b *_horizontal v
// (22 instead of 40 multiply-adds per pixel and image), and the adjoint is the same filter pair applied to
// u = sign(g a) a with the sign flipped (kx is odd in x and even in y), followed by the adjoint of the replicate
// padding (the padded positions fold onto the border pixels).
#include "common.cuh"
#include <map>
#include <mutex>

namespace dis {
namespace {

#ifndef DIS_SMOOTH_MAX_T
#define DIS_SMOOTH_MAX_T 224
#endif
constexpr int SM_MAX_T = DIS_SMOOTH_MAX_T;    // threads = columns of a strip with a halo of 4 on either side (432 = 2 x 216; 216 + 8 = 224)
constexpr int SM_HALO = 4;       // v at +-4 -> g, u at +-2 -> gradient
constexpr int SM_MIN_BAND = 16;  // shortest row band (the partials buffer is sized for it)
#ifndef DIS_SMOOTH_MIN_CTAS
#define DIS_SMOOTH_MIN_CTAS 4
#endif

// kx[i][j] / 240 rounded to float exactly like torch.from_numpy(kx).float()  (:701-705); ky = kx^T
#define SOB(v) ((float)((v) / 240.0))
__device__ __forceinline__ void sobel5_at(const float* __restrict__ t, int pitch, float& gx, float& gy) {
  // t points at the window's top-left element (row 0, col 0 of the 5x5 support)
  constexpr float k[5][5] = {{SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)},
                             {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                             {SOB(-10.0), SOB(-20.0), 0.f, SOB(20.0), SOB(10.0)},
                             {SOB(-8.0), SOB(-10.0), 0.f, SOB(10.0), SOB(8.0)},
                             {SOB(-5.0), SOB(-4.0), 0.f, SOB(4.0), SOB(5.0)}};
  float ax = 0.f, ay = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float v = t[i * pitch + j];
      if (j != 2) ax = fmaf(k[i][j], v, ax);
      if (i != 2) ay = fmaf(k[j][i], v, ay);
    }
  gx = ax;
  gy = ay;
}

// A CTA owns a strip of columns (one thread per column, halo included) and marches down a band of rows, TWO rows per
// step and ONE barrier per step.  Step (s, s+1), for each of its two rows r:
//   A  loads row r of (disp, ambient) as an fp32x2 pair (replicate-clamped), pushes it into the thread's 5-row register
//      ring, forms the vertical smoothers VA, VB of row y = r - 2 and posts v(r), VA(y), VB(y) in shared memory;
//   -- barrier --
//   C  (gradient) takes the u-side quantities the neighbours posted in the previous step, finishes the padded-domain
//      gradient of row r - 6, and emits row r - 8 (held one more step so that the border columns can collect the two
//      padded columns that fold onto them; the padded rows fold through a register);
//   B  reads the neighbours' v(r), VA(y), VB(y): HA(r), HB(r) update the pending gy sums of rows r-2..r+2, gx(y) is a
//      difference of four neighbours; g(y) -> loss and u(y); the vertical smoothers of ux for row q = y - 2 and uy(y)
//      are posted for the next step.
// Everything a step posts goes to buffer (step & 1) and is read before the barrier of the next step; the next write to
// that buffer comes after that barrier.  Rings are indexed (PH + k) % 5 with the loop unrolled by 5 steps (10 rows): no
// register moves.  The posted rows carry two pad columns on either side, so a neighbour is an immediate offset.
struct SmoothRings {
  u64 v[5];      // (disp, ambient) rows r-4..r
  u64 gy[5];     // pending gy sums of rows r-2..r+2
  float ux[5];   // ux rows y-4..y
  float ga[5];   // pending uy-side gradient sums of rows y'-2..y'+2
};
template <int V> struct SmoothPhase { static constexpr int value = V; };
constexpr int SM_PAD = 2, SM_ROW = SM_MAX_T + 2 * SM_PAD;
struct SmoothPost {                       // what one step posts, per row r in {0, 1}
  ulonglong2 ab[2][SM_ROW];               // (VA, VB) of row y
  float4 u[2][SM_ROW];                    // (uy(y), UA(q), UB(q), padded-domain gradient of row r - 6)
  u64 v[2][SM_ROW];                       // (disp, ambient) of row r
};

// sign(v) * a for a > 0
__device__ __forceinline__ float signed_mag0(float a, float v) { return v == 0.0f ? 0.0f : copysignf(a, v); }

template <bool GRAD>
__global__ void __launch_bounds__(SM_MAX_T, DIS_SMOOTH_MIN_CTAS)
smooth_march_kernel(const float* __restrict__ disp, const float* __restrict__ im, float* __restrict__ grad_sum,
                    float* __restrict__ partials, int H, int W, int strip, int band_rows, int per_frame, float grad_scale,
                    int accumulate) {
  extern __shared__ __align__(16) unsigned char smooth_smem[];
  SmoothPost* post = reinterpret_cast<SmoothPost*>(smooth_smem);   // [2]
  __shared__ double red[SM_MAX_T / 32];
  constexpr float A0 = SOB(5.0), A1 = SOB(8.0), A2 = SOB(10.0), B0 = SOB(4.0), B1 = SOB(10.0), B2 = SOB(20.0);
  const int T = blockDim.x, t = threadIdx.x, n = blockIdx.z, tp = t + SM_PAD;
  const int X0 = blockIdx.x * strip, X1 = min(X0 + strip, W);
  const int xc = X0 - SM_HALO + t, cx = clampi(xc, 0, W - 1);
  const bool col_img = xc >= 0 && xc < W, col_out = xc >= X0 && xc < X1;
  const bool fold_l = GRAD && col_out && xc == 0, fold_r = GRAD && col_out && xc == W - 1;
  const int y_begin = blockIdx.y * band_rows, y_end = min(y_begin + band_rows, H);
  // padded-domain gradient rows [qb, qe) of this band: the first / last band also owns the two padded rows above / below
  const int qb = (GRAD && y_begin == 0) ? -2 : y_begin, qe = (GRAD && y_end == H) ? H + 2 : y_end;
  const int u_first = GRAD ? qb - 2 : y_begin;              // first row whose g is needed (rings are primed by then)
  const int s0 = u_first - 2, s1 = GRAD ? qe + 7 : y_end + 1;   // first / last row to load
  const size_t hw = (size_t)H * W;
  const float* dn = disp + (size_t)n * hw + cx;
  const float* an = im + (size_t)n * hw + cx;
  float* go = GRAD ? grad_sum + (size_t)n * hw + cx : nullptr;

  SmoothRings R;
#pragma unroll
  for (int k = 0; k < 5; ++k) { R.v[k] = 0ull; R.gy[k] = 0ull; R.ux[k] = 0.f; R.ga[k] = 0.f; }
  float lsum = 0.f, facc = 0.f, pre[4];
  double dsum = 0.0;
  int buf = 0;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int ro = clampi(s0 + r, 0, H - 1) * W;
    pre[2 * r] = __ldg(dn + ro);
    pre[2 * r + 1] = __ldg(an + ro);
  }

  auto step = [&](auto phase, const int s) __attribute__((always_inline)) {
    constexpr int P0 = decltype(phase)::value * 2;   // ring phase of the step's first row
#define RING(a, ph, k) a[((ph) + (k)) % 5]
    SmoothPost& cur = post[buf];
    const SmoothPost& prv = post[buf ^ 1];
    // ---- A (the rows were requested a step ago; the next two are requested now)
    const u64 vnew[2] = {pk2(pre[0], pre[1]), pk2(pre[2], pre[3])};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int ro = clampi(s + 2 + r, 0, H - 1) * W;
      pre[2 * r] = __ldg(dn + ro);
      pre[2 * r + 1] = __ldg(an + ro);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      RING(R.v, P0 + r, 4) = vnew[r];
      const u64 t04 = add2(RING(R.v, P0 + r, 0), vnew[r]), t13 = add2(RING(R.v, P0 + r, 1), RING(R.v, P0 + r, 3));
      const u64 va = fma2(bc2(A2), RING(R.v, P0 + r, 2), fma2(bc2(A1), t13, mul2(bc2(A0), t04)));
      const u64 vb = fma2(bc2(B2), RING(R.v, P0 + r, 2), fma2(bc2(B1), t13, mul2(bc2(B0), t04)));
      cur.v[r][tp] = vnew[r];
      cur.ab[r][tp] = make_ulonglong2(va, vb);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = s + r;
      float gnew = 0.f;
      // ---- C: the previous step posted uy(y' = row - 4), UA / UB(q' = row - 6) and the gradient of row - 8
      if (GRAD && s > s0) {
        const float4 m2 = prv.u[r][tp - 2], m1 = prv.u[r][tp - 1], own = prv.u[r][tp], p1 = prv.u[r][tp + 1], p2 = prv.u[r][tp + 2];
        const int qq = row - 8;   // row whose padded-domain gradient was posted
        if (qq >= qb && qq < qe) {
          float g = own.w;
          if (fold_l) g += m1.w + m2.w;
          if (fold_r) g += p1.w + p2.w;
          facc += g;
          if ((qq >= 0 && qq < H - 1) || qq == H + 1) {   // last padded row that folds onto image row clamp(qq)
            if (col_out) {
              float* o = go + clampi(qq, 0, H - 1) * W;
              float v = facc * grad_scale;      // 1 for the plain sum; weight / count for a final gradient
              if (accumulate) v += *o;          // add to a gradient another term already left there
              __stcs(o, v);
            }
            facc = 0.f;
          }
        }
        const float e2 = m2.x + p2.x, e1 = m1.x + p1.x;
        const float wa = fmaf(A2, own.x, fmaf(A1, e1, A0 * e2)), wb = fmaf(B2, own.x, fmaf(B1, e1, B0 * e2));
        const float gfin = RING(R.ga, P0 + r, 0) + wa;     // row y' - 2 = row - 6 complete
        RING(R.ga, P0 + r, 1) += wb;
        RING(R.ga, P0 + r, 3) -= wb;
        RING(R.ga, P0 + r, 4) = -wa;
        gnew = -(((p2.y - m2.y) + (p1.z - m1.z)) + gfin);
      }
      // ---- B
      const u64 vm2 = cur.v[r][tp - 2], vm1 = cur.v[r][tp - 1], vp1 = cur.v[r][tp + 1], vp2 = cur.v[r][tp + 2];
      const u64 e2 = add2(vm2, vp2), e1 = add2(vm1, vp1);
      const u64 ha = fma2(bc2(A2), vnew[r], fma2(bc2(A1), e1, mul2(bc2(A0), e2)));
      const u64 hb = fma2(bc2(B2), vnew[r], fma2(bc2(B1), e1, mul2(bc2(B0), e2)));
      const u64 gy2 = add2(RING(R.gy, P0 + r, 0), ha);   // row y = row - 2 complete
      RING(R.gy, P0 + r, 1) = add2(RING(R.gy, P0 + r, 1), hb);
      RING(R.gy, P0 + r, 3) = sub2(RING(R.gy, P0 + r, 3), hb);
      RING(R.gy, P0 + r, 4) = sub2(0ull, ha);
      const ulonglong2 am2 = cur.ab[r][tp - 2], am1 = cur.ab[r][tp - 1], ap1 = cur.ab[r][tp + 1], ap2 = cur.ab[r][tp + 2];
      const u64 gx2 = add2(sub2(ap2.x, am2.x), sub2(ap1.y, am1.y));
      float gdx, gix, gdy, giy;
      upk2(gx2, gdx, gix);
      upk2(gy2, gdy, giy);
      const int y = row - 2;
      const float ax = expf(-fabsf(255.0f * gix)), ay = expf(-fabsf(255.0f * giy));
      const float vx = gdx * ax, vy = gdy * ay;
      if (y >= y_begin && y < y_end && col_out) lsum += fabsf(vx) + fabsf(vy);
      if (GRAD) {
        const bool ok = y >= u_first && y >= 0 && y < H && col_img;   // u is zero outside the image
        const float ux = ok ? signed_mag0(ax, vx) : 0.f, uy = ok ? signed_mag0(ay, vy) : 0.f;
        RING(R.ux, P0 + r, 4) = ux;
        const float c04 = RING(R.ux, P0 + r, 0) + ux, c13 = RING(R.ux, P0 + r, 1) + RING(R.ux, P0 + r, 3);
        const float ua = fmaf(A2, RING(R.ux, P0 + r, 2), fmaf(A1, c13, A0 * c04));
        const float ub = fmaf(B2, RING(R.ux, P0 + r, 2), fmaf(B1, c13, B0 * c04));
        cur.u[r][tp] = make_float4(uy, ua, ub, gnew);
      }
    }
    buf ^= 1;
#undef RING
  };

  // groups of 5 steps (10 rows) restore the ring phase
  int base = s0;
  for (; base + 9 <= s1; base += 10) {
    step(SmoothPhase<0>{}, base);
    step(SmoothPhase<1>{}, base + 2);
    step(SmoothPhase<2>{}, base + 4);
    step(SmoothPhase<3>{}, base + 6);
    step(SmoothPhase<4>{}, base + 8);
    dsum += (double)lsum;   // fp32 only over 10 rows: the band sum keeps fp64 accuracy
    lsum = 0.f;
  }
  if (base <= s1) {          // the last 1..4 steps (a row past s1 is harmless: clamped load, nothing emitted)
    step(SmoothPhase<0>{}, base);
    if (base + 2 <= s1) step(SmoothPhase<1>{}, base + 2);
    if (base + 4 <= s1) step(SmoothPhase<2>{}, base + 4);
    if (base + 6 <= s1) step(SmoothPhase<3>{}, base + 6);
    if (base + 8 <= s1) step(SmoothPhase<4>{}, base + 8);
    dsum += (double)lsum;
  }

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
  if ((t & 31) == 0) red[t >> 5] = dsum;
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int i = 0; i < T / 32; ++i) tot += red[i];
    float* pf = partials + 2 * (size_t)n * per_frame;
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    pf[2 * b] = (float)tot;
    pf[2 * b + 1] = 2.0f * (float)(y_end - y_begin) * (float)(X1 - X0);
    if (b == 0)   // slots of the bands a finer plan would have used
      for (int i = gridDim.x * gridDim.y; i < per_frame; ++i) { pf[2 * i] = 0.f; pf[2 * i + 1] = 0.f; }
  }
}

// Stand-alone Sobel (API completeness; direct clamped global reads, one thread per pixel)
__global__ void __launch_bounds__(256) sobel_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int H,
                                                        int W, int ksize, size_t total) {
  const size_t hw = (size_t)H * W;
  const int r = ksize / 2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    const float* p = x + n * hw;
    float gx = 0.f, gy = 0.f;
    if (ksize == 5) {
      float win[25];
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) win[i * 5 + j] = __ldg(p + (size_t)clampi(h + i - r, 0, H - 1) * W + clampi(w + j - r, 0, W - 1));
      sobel5_at(win, 5, gx, gy);
    } else {
      constexpr float k3[3][3] = {{(float)(-1.0 / 8.0), 0.f, (float)(1.0 / 8.0)},
                                  {(float)(-2.0 / 8.0), 0.f, (float)(2.0 / 8.0)},
                                  {(float)(-1.0 / 8.0), 0.f, (float)(1.0 / 8.0)}};
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float v = __ldg(p + (size_t)clampi(h + i - r, 0, H - 1) * W + clampi(w + j - r, 0, W - 1));
          gx = fmaf(k3[i][j], v, gx);
          gy = fmaf(k3[j][i], v, gy);
        }
    }
    out[(n * 2 + 0) * hw + pix] = gx;
    out[(n * 2 + 1) * hw + pix] = gy;
  }
}

// adjoint of sobel_fwd_kernel in gather form (atomics-free, deterministic)
__global__ void __launch_bounds__(256) sobel_bwd_kernel(const float* __restrict__ go, float* __restrict__ gx_out, int H,
                                                        int W, int ksize, size_t total) {
  const size_t hw = (size_t)H * W;
  const int r = ksize / 2;
  const double div = ksize == 5 ? 240.0 : 8.0;
  constexpr double k5[5][5] = {{-5, -4, 0, 4, 5}, {-8, -10, 0, 10, 8}, {-10, -20, 0, 20, 10}, {-8, -10, 0, 10, 8}, {-5, -4, 0, 4, 5}};
  constexpr double k3[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    const float* ux = go + (n * 2 + 0) * hw;
    const float* uy = go + (n * 2 + 1) * hw;
    const int ylo = (h == 0) ? -r : 0, yhi = (h == H - 1) ? r : 0;
    const int xlo = (w == 0) ? -r : 0, xhi = (w == W - 1) ? r : 0;
    float acc = 0.f;
    for (int py = ylo; py <= yhi; ++py)
      for (int px = xlo; px <= xhi; ++px)
        for (int i = 0; i < ksize; ++i)
          for (int j = 0; j < ksize; ++j) {
            const int sy = h + py - (i - r), sx = w + px - (j - r);
            if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
            const float kx = (float)((ksize == 5 ? k5[i][j] : k3[i][j]) / div);
            const float ky = (float)((ksize == 5 ? k5[j][i] : k3[j][i]) / div);
            acc = fmaf(kx, __ldg(ux + (size_t)sy * W + sx), acc);
            acc = fmaf(ky, __ldg(uy + (size_t)sy * W + sx), acc);
          }
    gx_out[idx] = acc;
  }
}

inline int flat_grid(size_t total) {
  const size_t want = (total + 255) / 256, cap = 148 * 16;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

struct SmoothPlan {
  int threads, strip, nstrips, band_rows, nbands, per_frame;
};
SmoothPlan smooth_plan(int N, int H, int W) {
  SmoothPlan p;
  int t = ((W + 2 * SM_HALO + 31) / 32) * 32;
  if (t > SM_MAX_T) t = SM_MAX_T;
  p.threads = t;
  p.strip = t - 2 * SM_HALO;
  p.nstrips = (W + p.strip - 1) / p.strip;
  // every band re-primes 10 rows, so bands should be long; shorten them only when the batch cannot fill the SMs
  p.band_rows = 64;
  while (p.band_rows > SM_MIN_BAND && (long)N * p.nstrips * ((H + p.band_rows - 1) / p.band_rows) < 148L * DIS_SMOOTH_MIN_CTAS) p.band_rows >>= 1;
  p.nbands = (H + p.band_rows - 1) / p.band_rows;
  p.per_frame = p.nstrips * ((H + SM_MIN_BAND - 1) / SM_MIN_BAND);   // independent of N: callers chunk the batch
  return p;
}

int smooth_loss_num_partials(int N, int H, int W) { return N * smooth_plan(N, H, W).per_frame; }

// > 48 KB of dynamic shared memory needs an opt-in per kernel and device; done once, off the launch path
template <typename K>
int smooth_prepare(K kernel) {
  static std::map<std::pair<const void*, int>, cudaError_t> done;
  static std::mutex mu;
  int device = 0;
  cudaGetDevice(&device);
  std::lock_guard<std::mutex> lock(mu);
  const std::pair<const void*, int> key(reinterpret_cast<const void*>(kernel), device);
  auto it = done.find(key);
  if (it == done.end())
    it = done.emplace(key, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(SmoothPost)))).first;
  if (it->second != cudaSuccess) { set_last_cuda_error(it->second); return DIS_ERR_CUDA_LAUNCH; }
  return DIS_OK;
}

int smooth_loss_forward(const float* disp, const float* im, float* grad_sum, float* partials, int N, int H, int W,
                        float grad_scale, int accumulate, cudaStream_t s) {
  const SmoothPlan p = smooth_plan(N, H, W);
  dim3 grid(p.nstrips, p.nbands, N);
  const size_t smem = 2 * sizeof(SmoothPost);
  if (grad_sum) {
    if (int rc = smooth_prepare(smooth_march_kernel<true>)) return rc;
    smooth_march_kernel<true><<<grid, p.threads, smem, s>>>(disp, im, grad_sum, partials, H, W, p.strip, p.band_rows, p.per_frame,
                                                            grad_scale, accumulate);
  } else {
    if (int rc = smooth_prepare(smooth_march_kernel<false>)) return rc;
    smooth_march_kernel<false><<<grid, p.threads, smem, s>>>(disp, im, nullptr, partials, H, W, p.strip, p.band_rows, p.per_frame,
                                                             grad_scale, 0);
  }
  return check_launch();
}

int sobel_forward(const float* x, float* out, int N, int H, int W, int ksize, cudaStream_t s) {
  const size_t total = (size_t)N * H * W;
  sobel_fwd_kernel<<<flat_grid(total), 256, 0, s>>>(x, out, H, W, ksize, total);
  return check_launch();
}

int sobel_backward(const float* go, float* gx, int N, int H, int W, int ksize, cudaStream_t s) {
  const size_t total = (size_t)N * H * W;
  sobel_bwd_kernel<<<flat_grid(total), 256, 0, s>>>(go, gx, H, W, ksize, total);
  return check_launch();
}

}  // namespace dis
