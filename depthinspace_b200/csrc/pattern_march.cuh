// Pair-symmetric, column-marching fused pattern loss (soft census types): every unordered pixel pair of the
// k x k windows is evaluated ONCE and serves the forward value and the gradient of both of its pixels.
//
// reference: the photometric loop of single_frame_worker.Worker.loss_forward
// (model/single_frame_worker.py:108-115) -> RectifiedPatternSimilarityLoss.tforward (model/networks.py:354-377)
// -> ext_functions.photometric_loss(..., 'census_sad' / 'census_mse') (model/ext_functions.py:156-183).
//
// Maths.  With g(d) = d / sqrt(d^2 + eps) the per-tap term is f(p,o) = F(g(e(p+o) - e(p)) - g(t(p+o) - t(p))) with
// F even, so f(p,o) = f(p+o,-o), and its derivative u(p,o) w.r.t. the neighbour is odd: u(p,o) = -u(p+o,-o).
// Hence, over the "forward" half window (dy > 0, or dy == 0 and dx > 0), one evaluation per unordered pair {p,q}:
//     numerator  += f * (w(p) + w(q))
//     G(p)       += v,   G(q) -= v,        v = (w(p) + w(q)) * u(p,o)
// Replicate padding (model/ext_functions.py:158-159) is handled by making the padded plane THE image: the
// (H + 2R) x (W + 2R) extended image carries the clamped e / t values and weight 0 on its R virtual border
// rows / columns; the gradient collected by a virtual pixel is folded onto the border pixel it replicates
// (= the backward of F.pad).  Clamp multiplicities come out exactly, with no special border pass.
//
// Organisation.  A CTA owns a band of the extended image: up to 8 warps side by side (lane = column), marching
// down the band MV = 2 rows per step.  Per step a thread owns pixels p0 = (row, col), p1 = (row + 1, col) and
// visits the cells q of rows row .. row + R + 1: the values of q are loaded once for both pairs (p0,q), (p1,q).
// The p side of the gradient stays in registers; the q side goes into a WARP-PRIVATE strip of shared-memory
// accumulators (RING rows x (32 + 2R) cells, all scales of a cell in one word of up to 128 bits), one
// read-modify-write per cell.  Strips of neighbouring warps are merged in a fixed order when a row retires: no
// atomics, bitwise reproducible.  Staged planes (warped pattern of every scale, LCN image, sigma) live in a ring
// of RING rows; every pixel of the band is staged exactly once (halo: R rows at the top of a band, R columns
// between column bands).  d proj / d disp is parked in the gradient output buffer at staging time and read back
// (L2 hit) when the row retires, instead of occupying shared memory for R + 2 rows.
//
// Scales.  NS = 2 / 4: the estimates of two scales share a packed fp32x2 register (FADD2 / FFMA2 / FMUL2), the target
// side of the soft census is computed once per pair for all scales.  NS = 1: (estimate, target) form the packed pair.
// Cost per unordered pair: NS = 4: 5 MUFU.RSQ, ~41 issue slots (the tile kernel in pattern_multi.cuh pays 43 slots and
// 5 MUFU per ORDERED pair, i.e. twice); NS = 1: 2 MUFU.RSQ, ~16 slots (census_kernels.cuh: 14 slots, 2 MUFU per ordered pair).
#pragma once
#include <type_traits>
#include "window.cuh"

namespace dis {

constexpr int MV = 2;                  // rows per thread per step
constexpr int MARCH_MAX_WARPS = 8;     // CTA <= 256 threads
constexpr int MARCH_CTAS_PER_SM = 3;
#ifndef DIS_MARCH_SIGN_CLAMP
#define DIS_MARCH_SIGN_CLAMP 0
#endif
#ifndef DIS_MARCH_UNROLL
#define DIS_MARCH_UNROLL 9
#endif
#ifndef DIS_MARCH_SYNCWARP
#define DIS_MARCH_SYNCWARP 1
#endif
#ifndef DIS_MARCH_PREFETCH
#define DIS_MARCH_PREFETCH 0     // streaming loads of the next step's rows issued one step ahead (measured slower: registers)
#endif
#ifndef DIS_MARCH_STASH_EARLY
#define DIS_MARCH_STASH_EARLY 0  // read-back of the parked d proj / d disp issued before the barrier (measured slower)
#endif
constexpr int MARCH_UNROLL = DIS_MARCH_UNROLL;   // cells per iteration of the rolled cell loops
constexpr int MARCH_MAX_REGS = 80;      // 6 warps per SM sub-partition (16 K registers each): 3 CTAs of 7 warps

template <int R>
struct MarchGeom {
  static constexpr int RH = ((R + MV - 1) / MV) * MV;  // halo rows recomputed at the top of a band
  static constexpr int RING = MV + R;                   // live rows of the staged planes and of the strips
  static constexpr int SW = 32 + 2 * R;                 // strip cells per row
};

struct PatternMarchArgs {
  const float* disp[4];
  float* grad[4];          // all NULL => forward only
  float* proj;             // optional (NS == 1 only): the warped pattern [N,1,H,W] (model/networks.py:367)
  const float* es;         // MODE 1 / 2: the estimate plane itself (no pattern warp); im = target, std_in = upstream weights
  float* out;              // MODE 1: the per-pixel loss map [N,1,H,W]
  const float* im;
  const float* std_in;
  const float* pattern;
  const float* grad_scale; // optional device float[NS]
  float* partials;         // [NS][num_blocks][2] = (num_s, den)
  int N, H, W;
  int ncb, nrb;            // column / row bands per frame
  int band_rows;           // extended rows owned by a row band (multiple of MV); the last band takes the rest
  int num_blocks;          // partial slots per scale (>= total_blocks; the tail is zero-filled)
  int total_blocks;        // CTAs of the whole call
  int block_offset;        // first CTA index of this launch (batch chunking)
  float eps, inv_k2, inv_w, inv_h;
};

// ---- per-cell words ---------------------------------------------------------------------------------------------------
// MCell<NP>: the staged estimates of one pixel, NP packed fp32x2 values (NS = 4: two scale pairs; NS = 2: one;
// NS = 1: the pair (estimate, target)).
template <int NP>
struct MCell {
  u64 v[NP];
};
template <int NP>
__device__ __forceinline__ MCell<NP> mcell_load(const u64* p) {
  MCell<NP> c;
  if (NP == 2) {
    const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(p);
    c.v[0] = q.x;
    c.v[NP - 1] = q.y;
  } else {
    c.v[0] = *p;
  }
  return c;
}

// MAcc<NS>: the gradient accumulator of one pixel, one float per scale (packed in pairs for NS >= 2).  Accesses to the
// accumulator strips are volatile, so that the read-modify-write sequences of consecutive cells keep their program
// order (the next cell of a lane is the previous cell of its neighbour lane; the warp runs them in lockstep and the LSU
// executes a warp's shared-memory operations in order), while the plain loads of the staged planes stay free to move.
template <int NS>
struct MAcc {
  static constexpr int NP = NS / 2;
  static constexpr int BYTES = 4 * NS;
  u64 v[NP];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int p = 0; p < NP; ++p) v[p] = 0ull;
  }
  __device__ __forceinline__ void add(const MAcc& o) {
#pragma unroll
    for (int p = 0; p < NP; ++p) v[p] = add2(v[p], o.v[p]);
  }
  __device__ __forceinline__ void sub(const MAcc& o) {
#pragma unroll
    for (int p = 0; p < NP; ++p) v[p] = sub2(v[p], o.v[p]);
  }
  __device__ __forceinline__ void ld(unsigned addr) {
    if (NP == 2) asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[NP - 1]) : "r"(addr));
    else asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v[0]) : "r"(addr));
  }
  __device__ __forceinline__ void st(unsigned addr) const {
    if (NP == 2) asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %2};" ::"r"(addr), "l"(v[0]), "l"(v[NP - 1]));
    else asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(v[0]));
  }
  __device__ __forceinline__ float get(int s) const {
    float lo, hi;
    upk2(v[s >> 1], lo, hi);
    return (s & 1) ? hi : lo;
  }
};
template <>
struct MAcc<1> {
  static constexpr int BYTES = 4;
  float v;
  __device__ __forceinline__ void zero() { v = 0.0f; }
  __device__ __forceinline__ void add(const MAcc& o) { v += o.v; }
  __device__ __forceinline__ void sub(const MAcc& o) { v -= o.v; }
  __device__ __forceinline__ void ld(unsigned addr) { asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); }
  __device__ __forceinline__ void st(unsigned addr) const { asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }
  __device__ __forceinline__ float get(int) const { return v; }
};

// shared memory: ring of staged rows (estimate cells, (t, w)), per-warp accumulator strips, reduction scratch
template <int R, int NS>
__host__ __device__ constexpr size_t pattern_march_smem_bytes(int nwarps) {
  using G = MarchGeom<R>;
  const size_t np = (NS + 1) / 2;
  const size_t pitch = 32 * (size_t)nwarps + 2 * R;
  const size_t strips = ((size_t)nwarps * G::RING * G::SW * 4 * NS + 15) / 16 * 16;
  return G::RING * pitch * np * 8 + G::RING * pitch * 8 + strips + (size_t)nwarps * (NS + 1) * 8 +
         (size_t)(NS + 1) * 32 * nwarps * 8;
}

// ---- NS = 2 / 4: estimate side of one unordered pair {p (centre), q} for all scales, given the target side
// (ngt = -(dt * rt) rounded, rest = the exact rounding residual of dt * rt) and the pair weight ws = w(p) + w(q).
// acc: forward accumulators, gp: p-side gradient, cell: running sum for q's cell.
template <int TYPE, int NS, bool GRAD, bool FIRST>
__device__ __forceinline__ void march_scales(const MCell<NS / 2>& ec, const MCell<NS / 2>& eq, float ngt, float rest, float ws,
                                             u64 eps2, float (&acc)[NS], MAcc<NS>& gp, MAcc<NS>& cell) {
#pragma unroll
  for (int p = 0; p < NS / 2; ++p) {
    const u64 de = sub2(eq.v[p], ec.v[p]);
    const u64 xe = fma2(de, de, eps2);
    float x0f, x1f;
    upk2(xe, x0f, x1f);
    const u64 re = pk2(rsqrt_fast(x0f), rsqrt_fast(x1f));
    // diff = (de*re - gt) - (dt*rt - gt): both products enter through an exact FMA residual, so e == t gives
    // exactly 0 (the reference's |.| has subgradient 0 there)
    const u64 diff = sub2(fma2(de, re, bc2(ngt)), bc2(rest));
    float d0, d1;
    upk2(diff, d0, d1);
    if (TYPE == CENSUS_SAD) {
      acc[2 * p] = fmaf(fabsf(d0), ws, acc[2 * p]);
      acc[2 * p + 1] = fmaf(fabsf(d1), ws, acc[2 * p + 1]);
    } else {
      acc[2 * p] = fmaf(d0 * d0, ws, acc[2 * p]);
      acc[2 * p + 1] = fmaf(d1 * d1, ws, acc[2 * p + 1]);
    }
    if (GRAD) {
      const u64 r3 = mul2(mul2(re, re), re);
      u64 u;
      if (TYPE == CENSUS_SAD) {
        float q0, q1;
        upk2(r3, q0, q1);
#if DIS_MARCH_SIGN_CLAMP
        // sign(d) * q with sign(0) = 0 as clamp(d * 2^96, -q, q): one packed FMUL for two scales + two FMNMX per scale
        // instead of LOP3 + FSETP + FSEL (a non-zero d is >= 2^-60 here and q = rsqrt^3 <= eps^-1.5: exact for eps >= 1e-12)
        float t0, t1;
        upk2(mul2(diff, bc2(7.9228162514264338e28f)), t0, t1);
        u = pk2(fminf(fmaxf(t0, -q0), q0), fminf(fmaxf(t1, -q1), q1));
#else
        u = pk2(signed_mag(q0, d0), signed_mag(q1, d1));
#endif
      } else {
        u = mul2(diff, r3);
      }
      const u64 v = mul2(u, bc2(ws));
      gp.v[p] = add2(gp.v[p], v);
      cell.v[p] = FIRST ? v : add2(cell.v[p], v);
    }
  }
}

// one pair: scalar target side
template <int TYPE, int NS, bool GRAD>
__device__ __forceinline__ void march_pair1(const MCell<NS / 2>& ec, float tc, float wc, const MCell<NS / 2>& eq, float tq,
                                            float wq, float eps, u64 eps2, float (&acc)[NS], MAcc<NS>& gp, MAcc<NS>& cell) {
  const float dt = tq - tc;
  const float rt = rsqrt_fast(fmaf(dt, dt, eps));
  const float gt = __fmul_rn(dt, rt);
  const float rest = __fmaf_rn(dt, rt, -gt);
  march_scales<TYPE, NS, GRAD, true>(ec, eq, -gt, rest, wq + wc, eps2, acc, gp, cell);
}

// both pixels of the thread against the same q: the two target sides share packed instructions
template <int TYPE, int NS, bool GRAD>
__device__ __forceinline__ void march_pair2(const MCell<NS / 2>& ec0, const MCell<NS / 2>& ec1, u64 tc2, u64 wc2,
                                            const MCell<NS / 2>& eq, float tq, float wq, u64 eps2, float (&acc)[NS],
                                            MAcc<NS>& gp0, MAcc<NS>& gp1, MAcc<NS>& cell) {
  const u64 dt = sub2(bc2(tq), tc2);
  const u64 xt = fma2(dt, dt, eps2);
  float x0f, x1f;
  upk2(xt, x0f, x1f);
  const u64 rt = pk2(rsqrt_fast(x0f), rsqrt_fast(x1f));
  const u64 gt = mul2(dt, rt);
  const u64 ngt = mul2(gt, bc2(-1.0f));
  const u64 rest = fma2(dt, rt, ngt);
  const u64 ws = add2(bc2(wq), wc2);
  float n0, n1, r0, r1, w0, w1;
  upk2(ngt, n0, n1);
  upk2(rest, r0, r1);
  upk2(ws, w0, w1);
  march_scales<TYPE, NS, GRAD, true>(ec0, eq, n0, r0, w0, eps2, acc, gp0, cell);
  march_scales<TYPE, NS, GRAD, false>(ec1, eq, n1, r1, w1, eps2, acc, gp1, cell);
}

// ---- NS = 1: the cell is the packed pair (estimate, target); both sides of the soft census share every instruction.
// The two products d * r are rounded separately (one FMUL2) before they are subtracted, so e == t gives exactly 0.
// MAP: instead of the weighted loss sum and its gradient, the tap value itself goes to both pixels (per-pixel loss map).
template <int TYPE, bool GRAD, bool MAP = false>
__device__ __forceinline__ void march_et_pair1(u64 c_et, float wc, u64 q_et, float wq, u64 eps2, float& acc, MAcc<1>& gp,
                                               MAcc<1>& cell) {
  const u64 d = sub2(q_et, c_et);
  const u64 x = fma2(d, d, eps2);
  float xe, xt;
  upk2(x, xe, xt);
  const float re = rsqrt_fast(xe), rt = rsqrt_fast(xt);
  float pe, pt;
  upk2(mul2(d, pk2(re, rt)), pe, pt);
  const float diff = __fsub_rn(pe, pt);
  if (MAP) {
    const float f = (TYPE == CENSUS_SAD) ? fabsf(diff) : diff * diff;
    gp.v += f;
    cell.v = f;
    return;
  }
  const float ws = wq + wc;
  acc = (TYPE == CENSUS_SAD) ? fmaf(fabsf(diff), ws, acc) : fmaf(diff * diff, ws, acc);
  if (GRAD) {
    const float r3 = re * re * re;
    const float u = (TYPE == CENSUS_SAD) ? signed_mag(r3, diff) : diff * r3;
    const float v = u * ws;
    gp.v += v;
    cell.v = v;
  }
}
template <int TYPE, bool GRAD, bool MAP = false>
__device__ __forceinline__ void march_et_pair2(u64 c0_et, u64 c1_et, u64 wc2, u64 q_et, float wq, u64 eps2, float& acc,
                                               MAcc<1>& gp0, MAcc<1>& gp1, MAcc<1>& cell) {
  const u64 d0 = sub2(q_et, c0_et), d1 = sub2(q_et, c1_et);
  const u64 x0 = fma2(d0, d0, eps2), x1 = fma2(d1, d1, eps2);
  float xe0, xt0, xe1, xt1;
  upk2(x0, xe0, xt0);
  upk2(x1, xe1, xt1);
  const float re0 = rsqrt_fast(xe0), rt0 = rsqrt_fast(xt0), re1 = rsqrt_fast(xe1), rt1 = rsqrt_fast(xt1);
  float pe0, pt0, pe1, pt1;
  upk2(mul2(d0, pk2(re0, rt0)), pe0, pt0);
  upk2(mul2(d1, pk2(re1, rt1)), pe1, pt1);
  const float diff0 = __fsub_rn(pe0, pt0), diff1 = __fsub_rn(pe1, pt1);
  if (MAP) {
    const float f0 = (TYPE == CENSUS_SAD) ? fabsf(diff0) : diff0 * diff0;
    const float f1 = (TYPE == CENSUS_SAD) ? fabsf(diff1) : diff1 * diff1;
    gp0.v += f0;
    gp1.v += f1;
    cell.v = f0 + f1;
    return;
  }
  const u64 ws2 = add2(bc2(wq), wc2);
  float ws0, ws1;
  upk2(ws2, ws0, ws1);
  if (TYPE == CENSUS_SAD) {
    acc = fmaf(fabsf(diff0), ws0, acc);
    acc = fmaf(fabsf(diff1), ws1, acc);
  } else {
    acc = fmaf(diff0 * diff0, ws0, acc);
    acc = fmaf(diff1 * diff1, ws1, acc);
  }
  if (GRAD) {
    const u64 rp = pk2(re0, re1);
    const u64 r3 = mul2(mul2(rp, rp), rp);
    u64 u;
    if (TYPE == CENSUS_SAD) {
      float q0, q1;
      upk2(r3, q0, q1);
      u = pk2(signed_mag(q0, diff0), signed_mag(q1, diff1));
    } else {
      u = mul2(pk2(diff0, diff1), r3);
    }
    float v0, v1;
    upk2(mul2(u, ws2), v0, v1);
    gp0.v += v0;
    gp1.v += v1;
    cell.v = v0 + v1;
  }
}

// Sum of everything the strips hold for CTA-lane position P of strip row rq (own strip, then the halo cells of the
// left and of the right neighbour warp: fixed order); the cells are zeroed for their next use.
template <int R, int NS>
__device__ __forceinline__ MAcc<NS> strip_collect(unsigned strips_addr, int P, int rq, int nwarps) {
  using G = MarchGeom<R>;
  const int w = P >> 5, l = P & 31;
  MAcc<NS> zero, t, c;
  zero.zero();
  const unsigned own = strips_addr + (unsigned)(((w * G::RING + rq) * G::SW + l + R) * MAcc<NS>::BYTES);
  t.ld(own);
  zero.st(own);
  if (l < R && w > 0) {
    const unsigned h = strips_addr + (unsigned)((((w - 1) * G::RING + rq) * G::SW + 32 + R + l) * MAcc<NS>::BYTES);
    c.ld(h);
    zero.st(h);
    t.add(c);
  }
  if (l >= 32 - R && w < nwarps - 1) {
    const unsigned h = strips_addr + (unsigned)((((w + 1) * G::RING + rq) * G::SW + l - 32 + R) * MAcc<NS>::BYTES);
    c.ld(h);
    zero.st(h);
    t.add(c);
  }
  return t;
}

// MODE 0: fused pattern loss (pattern warp by `disp`, sigma-weighted sums, d/d disp).
// MODE 1 / 2 (NS == 1): the ext boundary ops on given planes es / ta (C == 1), reference model/ext_functions.py:115-140:
//   1  photometric_loss_forward:  the per-pixel loss map (no weights, nothing folded: virtual pixels have no output)
//   2  photometric_loss_backward: d/d es for upstream weights grad_out (in the `std_in` slot), written to grad[0]
constexpr int MARCH_FUSED = 0, MARCH_MAP = 1, MARCH_GRAD_E = 2;

template <int TYPE, int R, int NS, bool GRAD, int MODE = MARCH_FUSED>
__global__ void __maxnreg__(MARCH_MAX_REGS) pattern_march_kernel(PatternMarchArgs a) {
  static_assert(TYPE == CENSUS_MSE || TYPE == CENSUS_SAD, "pair symmetry is a property of the census types");
  static_assert(R >= 1, "a 1 x 1 window has no pairs");
  static_assert(NS == 1 || NS == 2 || NS == 4, "1, 2 or 4 scales");
  static_assert(MODE == MARCH_FUSED || (NS == 1 && GRAD), "the ext modes work on one plane and use the strips");
  constexpr bool MAP = (MODE == MARCH_MAP);
  using G = MarchGeom<R>;
  constexpr int S = NS;
  constexpr int NP = (NS + 1) / 2;          // packed words per staged cell
  constexpr bool ET = (NS == 1);            // the packed pair is (estimate, target)
  using Acc = MAcc<NS>;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int LW = blockDim.x, nwarps = LW >> 5;
  const int PITCH = LW + 2 * R;
  extern __shared__ __align__(16) unsigned char march_smem[];
  u64* ring_e = reinterpret_cast<u64*>(march_smem);                                  // [RING][PITCH][NP]
  float2* ring_tw = reinterpret_cast<float2*>(ring_e + (size_t)G::RING * PITCH * NP);     // [RING][PITCH] (t, w)
  unsigned char* strips = reinterpret_cast<unsigned char*>(ring_tw + (size_t)G::RING * PITCH);   // [nwarps][RING][SW] Acc
  const size_t strip_bytes = ((size_t)nwarps * G::RING * G::SW * Acc::BYTES + 15) / 16 * 16;
  double* red = reinterpret_cast<double*>(strips + strip_bytes);                     // [nwarps][S+1]
  double* sums = red + nwarps * (S + 1);                                             // [S+1][LW]

  const int H = a.H, W = a.W;
  const int EH = H + 2 * R, EW = W + 2 * R;
  const int cb = blockIdx.x, rb = blockIdx.y, n = blockIdx.z;
  // ---- my column ---------------------------------------------------------------------------------------
  const int c0 = cb * (LW - 2 * R);
  const int ec = c0 + tid;                                   // extended column of this thread
  const int own_c0 = c0 + (cb > 0 ? R : 0);
  const int own_c1 = (cb == a.ncb - 1) ? EW : c0 + LW - R;
  const bool col_own = ec >= own_c0 && ec < own_c1;          // (implies ec < EW)
  const bool col_in = ec >= R && ec < EW - R;                // a real image column
  const bool out_lane = col_own && col_in;
  const int x = clampi(ec - R, 0, W - 1);
  // ---- my rows -----------------------------------------------------------------------------------------
  const int own_r0 = rb * a.band_rows;
  const int own_r1 = (rb == a.nrb - 1) ? EH : own_r0 + a.band_rows;
  const int row_start = rb > 0 ? own_r0 - G::RH : 0;
  const int nsteps = (own_r1 - row_start + MV - 1) / MV;

  const size_t hw = (size_t)H * W;
  const size_t fo = (size_t)n * hw;
  // (plane pointers are re-derived from the parameter block at each use: they would cost 20 registers otherwise)

  const float gs = -0.5f * a.eps * a.inv_k2;
  float gss[S];            // census chain-rule constant x the caller's per-scale factor
#pragma unroll
  for (int s = 0; s < S; ++s) gss[s] = (GRAD && a.grad_scale) ? gs * __ldg(a.grad_scale + s) : gs;
  const u64 eps2 = bc2(a.eps);

  // ---- one-time initialisation: zero strips, finite zero-weight padding cells of the ring --------------------
  for (int i = tid; i < (int)(strip_bytes / 4); i += LW) reinterpret_cast<float*>(strips)[i] = 0.0f;
  for (int i = tid; i < G::RING * 2 * R; i += LW) {
    const int r = i / (2 * R), k = i - r * 2 * R;
    const int slot = r * PITCH + (k < R ? k : LW + k);
#pragma unroll
    for (int p = 0; p < NP; ++p) ring_e[(size_t)slot * NP + p] = 0ull;
    ring_tw[slot] = make_float2(0.0f, 0.0f);
  }

  // Staging of one extended row in two halves: the streaming loads, then the pattern warp of every scale and the
  // stores into the ring.
  struct Raw {
    float dv[S];
    float tv, wv;
  };
  auto load_raw = [&](int er) __attribute__((always_inline)) {
    const int y = clampi(er - R, 0, H - 1);
    const bool inside = col_in && er >= R && er < EH - R;
    const size_t g = (size_t)y * W + x;
    Raw r;
#pragma unroll
    for (int s = 0; s < S; ++s) r.dv[s] = __ldg((MODE == MARCH_FUSED ? a.disp[s] : a.es) + fo + g);
    r.tv = __ldg(a.im + fo + g);
    r.wv = (inside && !MAP) ? (a.std_in ? __ldg(a.std_in + fo + g) : 1.0f) : 0.0f;
    return r;
  };
  auto finish_row = [&](int er, const Raw& r) __attribute__((always_inline)) {
    const int y = clampi(er - R, 0, H - 1);
    const bool inside = col_in && er >= R && er < EH - R;
    const WarpRow row = warp_row_setup(y, H, W, a.inv_h);
    const float* prow0 = a.pattern + row.off0;
    const float* prow1 = a.pattern + row.off1;
    const bool own_px = inside && col_own && er >= own_r0 && er < own_r1;
    const bool want_dd = GRAD && own_px && MODE == MARCH_FUSED;
    float ev[S], dd[S];
#pragma unroll
    for (int s = 0; s < S; ++s)
      ev[s] = MODE == MARCH_FUSED ? warp_col_sample_clamped(prow0, prow1, row.wy0, row.wy1, r.dv[s], x, W, a.inv_w, want_dd ? &dd[s] : nullptr)
                                  : r.dv[s];
    const int slot = (er % G::RING) * PITCH + R + tid;
    if (ET) {
      ring_e[slot] = pk2(ev[0], r.tv);
    } else {
#pragma unroll
      for (int p = 0; p < NP; ++p) ring_e[(size_t)slot * NP + p] = pk2(ev[2 * p], ev[2 * p + (NS > 1 ? 1 : 0)]);
    }
    ring_tw[slot] = make_float2(r.tv, r.wv);
    const size_t g = (size_t)y * W + x;
    if (want_dd) {
#pragma unroll
      for (int s = 0; s < S; ++s) a.grad[s][fo + g] = dd[s] * gss[s];   // parked; multiplied by G(p) when the row retires
    }
    if (ET && MODE == MARCH_FUSED && a.proj && own_px) a.proj[fo + g] = ev[0];
  };

  for (int er = row_start; er < row_start + R; ++er) finish_row(er, load_raw(er));
  Raw raw[MV];
  if (DIS_MARCH_PREFETCH) {
#pragma unroll
    for (int i = 0; i < MV; ++i) raw[i] = load_raw(row_start + R + i);
  }

  // per-thread fp64 running sums (S numerators + the denominator) live in shared memory, not in 10 registers
  double* my_sums = sums + tid;
#pragma unroll
  for (int s = 0; s <= S; ++s) my_sums[s * LW] = 0.0;
  // vertical fold: first the top virtual rows (added to image row 0, then reset), later image row H-1 plus the
  // virtual rows below it -- never both at once (H >= 2), so one accumulator serves both
  Acc facc;
  facc.zero();
  const unsigned strips_addr = (unsigned)__cvta_generic_to_shared(strips);
  const unsigned my_strip = strips_addr + (unsigned)((wid * G::RING * G::SW + lane + R) * Acc::BYTES);
  constexpr unsigned STRIP_ROW_BYTES = G::SW * Acc::BYTES;

#pragma unroll 1
  for (int step = 0; step < nsteps; ++step) {
    const int ys = row_start + step * MV;
    const int r0 = ys % G::RING;
    // ---- the MV new rows of this step enter the ring ---------------------------------------------------------
    if (!DIS_MARCH_PREFETCH) {
#pragma unroll
      for (int i = 0; i < MV; ++i) raw[i] = load_raw(ys + R + i);
    }
#pragma unroll
    for (int i = 0; i < MV; ++i) finish_row(ys + R + i, raw[i]);
    __syncthreads();
    if (DIS_MARCH_PREFETCH && step + 1 < nsteps) {
#pragma unroll
      for (int i = 0; i < MV; ++i) raw[i] = load_raw(ys + MV + R + i);
    }

    MCell<NP> pe[MV];
    Acc gp[MV];
    float ptc[MV], pwc[MV];
    float acc[S];
#pragma unroll
    for (int s = 0; s < S; ++s) acc[s] = 0.0f;
#pragma unroll
    for (int i = 0; i < MV; ++i) {
      int rq = r0 + i;
      if (rq >= G::RING) rq -= G::RING;
      const int slot = rq * PITCH + R + tid;
      pe[i] = mcell_load<NP>(ring_e + (size_t)slot * NP);
      const float2 tw = ring_tw[slot];
      ptc[i] = tw.x;
      pwc[i] = tw.y;
      gp[i].zero();
    }
    const u64 tc2 = pk2(ptc[0], ptc[1]), wc2 = pk2(pwc[0], pwc[1]);
    float stash[MV][S];     // d proj / d disp x constants of my two pixels, parked in the gradient buffer at staging
    int stash_at[MV];

    // one cell q = (row ys + j, column + dx); P0 / P1: which of my two pixels pair with it.
    auto cell = [&](int rq, int dx, auto p0_tag, auto p1_tag) __attribute__((always_inline)) {
      constexpr bool P0 = decltype(p0_tag)::value, P1 = decltype(p1_tag)::value;
      const int slot = rq * PITCH + R + tid + dx;
      const unsigned c = my_strip + (unsigned)rq * STRIP_ROW_BYTES + dx * Acc::BYTES;
      Acc cur;
      if (GRAD) cur.ld(c);     // issued first: its latency hides behind the pair arithmetic
      const MCell<NP> eq = mcell_load<NP>(ring_e + (size_t)slot * NP);
      const float2 tw = ring_tw[slot];
      Acc cs;
      if constexpr (ET) {
        if (P0 && P1) march_et_pair2<TYPE, GRAD, MAP>(pe[0].v[0], pe[1].v[0], wc2, eq.v[0], tw.y, eps2, acc[0], gp[0], gp[1], cs);
        else if (P0) march_et_pair1<TYPE, GRAD, MAP>(pe[0].v[0], pwc[0], eq.v[0], tw.y, eps2, acc[0], gp[0], cs);
        else march_et_pair1<TYPE, GRAD, MAP>(pe[1].v[0], pwc[1], eq.v[0], tw.y, eps2, acc[0], gp[1], cs);
      } else {
        if (P0 && P1) march_pair2<TYPE, NS, GRAD>(pe[0], pe[1], tc2, wc2, eq, tw.x, tw.y, eps2, acc, gp[0], gp[1], cs);
        else if (P0) march_pair1<TYPE, NS, GRAD>(pe[0], ptc[0], pwc[0], eq, tw.x, tw.y, a.eps, eps2, acc, gp[0], cs);
        else march_pair1<TYPE, NS, GRAD>(pe[1], ptc[1], pwc[1], eq, tw.x, tw.y, a.eps, eps2, acc, gp[1], cs);
      }
      if (GRAD) {
        if (MAP) cur.add(cs);      // the tap value goes to both pixels; the gradient is antisymmetric
        else cur.sub(cs);
        cur.st(c);
#if DIS_MARCH_SYNCWARP
        __syncwarp();   // memory-model form of the hand-over to the neighbour lane (the LSU already keeps the order)
#endif
      }
    };
    using T_ = std::true_type;
    using F_ = std::false_type;
    auto ring_row = [&](int j) {
      const int rq = r0 + j;
      return rq >= G::RING ? rq - G::RING : rq;
    };
    // p0 alone: row ys (dy = 0: dx > 0) and the left half of row ys + 1 (dy = 1; p1 joins for dx > 0)
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      const int rq = ring_row(j);
      const int lo = j == 0 ? 1 : -R, hi = j == 0 ? R : 0;
#pragma unroll 1
      for (int dx = lo; dx <= hi; ++dx) cell(rq, dx, T_{}, F_{});
    }
    // both pixels: right half of row ys + 1, then rows ys + 2 .. ys + R in full
#pragma unroll 1
    for (int j = 1; j <= R; ++j) {
      const int rq = ring_row(j);
      if (MARCH_UNROLL >= 2 * R + 1 && j == 1) {   // (fully unrolled variant: keep immediate offsets)
#pragma unroll
        for (int dx = 1; dx <= R; ++dx) cell(rq, dx, T_{}, T_{});
      } else {
#pragma unroll MARCH_UNROLL
        for (int dx = (MARCH_UNROLL >= 2 * R + 1 || j != 1) ? -R : 1; dx <= R; ++dx) cell(rq, dx, T_{}, T_{});
      }
    }
    {  // p1 alone: row ys + R + 1 (dy = R)
      const int rq = ring_row(R + 1);
#pragma unroll MARCH_UNROLL
      for (int dx = -R; dx <= R; ++dx) cell(rq, dx, F_{}, T_{});
    }
    if (GRAD) {  // my own pixels' p-side sums join the strip
#pragma unroll
      for (int i = 0; i < MV; ++i) {
        const unsigned c = my_strip + (unsigned)ring_row(i) * STRIP_ROW_BYTES;
        Acc cur;
        cur.ld(c);
        cur.add(gp[i]);
        cur.st(c);
      }
    }
    auto load_stash = [&]() __attribute__((always_inline)) {
#pragma unroll
      for (int i = 0; i < MV; ++i) {
        const int er = ys + i;
        const int y = er >= EH - R - 1 ? H - 1 : max(er - R, 0);
        stash_at[i] = y * W + x;
#pragma unroll
        for (int s = 0; s < S; ++s)
          stash[i][s] = (out_lane && er >= own_r0 && er < own_r1 && er >= R) ? a.grad[s][fo + stash_at[i]] : 0.0f;
      }
    };
    if (GRAD && DIS_MARCH_STASH_EARLY && MODE == MARCH_FUSED) load_stash();   // L2 hits, in flight across the barrier
    if (MODE == MARCH_FUSED && ys >= own_r0) {  // pairs are counted by the band that owns p (halo steps only feed the strips)
#pragma unroll
      for (int s = 0; s < S; ++s) my_sums[s * LW] += (double)acc[s];
    }
    __syncthreads();

    // ---- rows ys, ys + 1 retire: merge strips, fold virtual pixels, write the gradient ------------------------
    if (GRAD && !DIS_MARCH_STASH_EARLY && MODE == MARCH_FUSED) load_stash();
    if (out_lane) {
#pragma unroll
      for (int i = 0; i < MV; ++i) {
        const int er = ys + i;
        const int rq = ring_row(i);
        const bool row_own = er >= own_r0 && er < own_r1;
        if (MODE == MARCH_FUSED && row_own) my_sums[S * LW] += (double)ring_tw[rq * PITCH + R + tid].y;
        if (GRAD) {
          Acc t = strip_collect<R, NS>(strips_addr, tid, rq, nwarps);
          if (MAP) {      // loss map: real pixels only, nothing to fold
            if (row_own && er >= R && er < EH - R)
              a.out[fo + (size_t)(er - R) * W + x] = t.get(0) * (fwd_scale<TYPE>() * a.inv_k2);
            continue;
          }
          if (ec == R) {  // left border column: virtual columns 0 .. R-1 fold onto it
            for (int k = 0; k < R; ++k) t.add(strip_collect<R, NS>(strips_addr, tid - R + k, rq, nwarps));
          }
          if (ec == EW - R - 1) {  // right border column
            for (int k = 1; k <= R; ++k) t.add(strip_collect<R, NS>(strips_addr, tid + k, rq, nwarps));
          }
          if (row_own) {
            bool emit = true;
            if (er < R) {                      // top virtual row
              facc.add(t);
              emit = false;
            } else if (er >= EH - R - 1) {     // image row H-1 and the virtual rows below it
              facc.add(t);
              emit = (er == EH - 1);
              t = facc;
            } else if (er == R) {              // image row 0
              t.add(facc);
              facc.zero();
            }
            if (emit) {
              if (MODE == MARCH_GRAD_E) {
                const int y = er >= EH - R - 1 ? H - 1 : er - R;
                a.grad[0][fo + (size_t)y * W + x] = t.get(0) * gs;
              } else {
#pragma unroll
                for (int s = 0; s < S; ++s) a.grad[s][fo + stash_at[i]] = t.get(s) * stash[i][s];
              }
            }
          }
        }
      }
    }
    // (no barrier here: the next step's new rows go into the ring rows of ys, ys+1, which every warp finished reading
    //  before the barrier above; strip rows ys, ys+1 are written again only after the barrier that follows the next staging)
  }

  if (MODE != MARCH_FUSED) return;
  // ---- CTA reduction of S numerators + the denominator (fixed order) ------------------------------------------
  const double fs = (double)(fwd_scale<TYPE>() * a.inv_k2);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    double v = col_own ? my_sums[s * LW] * fs : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid * (S + 1) + s] = v;
  }
  {
    double v = my_sums[S * LW];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid * (S + 1) + S] = v;
  }
  __syncthreads();
  if (tid < S) {
    double num = 0.0, den = 0.0;
    for (int w = 0; w < nwarps; ++w) {
      num += red[w * (S + 1) + tid];
      den += red[w * (S + 1) + S];
    }
    const size_t blocks_per_frame = (size_t)a.ncb * a.nrb;
    const size_t b = (size_t)a.block_offset + (size_t)n * blocks_per_frame + (size_t)rb * a.ncb + cb;
    float* out = a.partials + (size_t)tid * a.num_blocks * 2;
    out[b * 2] = (float)num;
    out[b * 2 + 1] = (float)den;
    for (size_t z = (size_t)a.total_blocks + b; z < (size_t)a.num_blocks; z += a.total_blocks) {
      out[z * 2] = 0.0f;
      out[z * 2 + 1] = 0.0f;
    }
  }
}

// Band plan for one call (host side): warps per CTA / column bands minimising idle lanes, row bands balancing the
// halo re-computation against the number of CTA waves.
struct MarchPlan {
  int nwarps, ncb, nrb, band_rows;
};
MarchPlan march_plan(int N, int H, int W, int R);

template <int R> int launch_pattern_march(const PatternMarchArgs& a, const MarchPlan& plan, int S, int type, cudaStream_t s);
// ext boundary ops (model/ext_functions.py:115-140) on planes es / ta, C == 1: mode MARCH_MAP or MARCH_GRAD_E
template <int R> int launch_photometric_march(const PatternMarchArgs& a, const MarchPlan& plan, int mode, int type, cudaStream_t s);

}  // namespace dis
