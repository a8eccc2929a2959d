// Neighbour selection + gather of FuseNet's Conv3D (the step right after the flow warps in DIS-MF).
// reference: Conv3D.tforward, model/multi_frame_networks.py:469-501
//   unfold xyz / feat / mask [tl,bs,C,h,w] into k x k x tl candidates per output pixel (zero padding, stride s),
//   rank the candidates by squared distance to the centre ray in the normalised image plane
//   (masked-out candidates get global_max + 1), keep the `neighbors` smallest, gather xyz_local and feat.
// The reference materialises three unfolded tensors of k*k*tl (=36) times the input and runs topk + two gathers;
// here one thread ranks the 36 candidates of its output pixel in registers and only the 9 winners are written.
// Ties (equal keys) are broken by the lowest candidate index; torch.topk(sorted=False) leaves them unspecified.
// Backward is a deterministic gather (no atomics): every source element looks up the <= k*k output pixels whose
// window contains it and adds the gradients of the slots that selected it.
#include <cstdint>
#include "common.cuh"

namespace dis {
namespace {

constexpr int MAX_CAND = 64;
constexpr int MAX_KK = 64;   // window positions k*k (api.cu admits k*k*tl <= 64)

struct C3Args {
  const float* xyz; const float* feat; const float* mask;
  float* xyz_nb; float* feat_nb; uint8_t* idx; float* gmax; float* plane;
  int tl, bs, C, h, w, k, stride, nb, oh, ow;
};

// squared plane distance of candidate (t, ky, kx) of output pixel (b, oy, ox) to the centre candidate, and its mask
__device__ __forceinline__ void cand_pos(const C3Args& a, int oy, int ox, int ky, int kx, int& y, int& x, bool& inside) {
  const int pad = (a.k - 1) / 2;
  y = oy * a.stride + ky - pad;
  x = ox * a.stride + kx - pad;
  inside = y >= 0 && y < a.h && x >= 0 && x < a.w;
}

__device__ __forceinline__ void load_xyz(const C3Args& a, int t, int b, int y, int x, bool inside, float v[3]) {
  if (!inside) { v[0] = v[1] = v[2] = 0.f; return; }   // F.pad(constant 0), :472
  const size_t hw = (size_t)a.h * a.w;
  const float* p = a.xyz + ((size_t)(t * a.bs + b) * 3) * hw + (size_t)y * a.w + x;
  v[0] = __ldg(p); v[1] = __ldg(p + hw); v[2] = __ldg(p + 2 * hw);
}

// plane = xyz / (z + 1e-12)  (:489), once per source element instead of once per (output pixel, candidate): every
// element is a candidate of up to k*k output pixels and both ranking passes need it (IEEE divisions: ~10 instructions
// each).  Zero-padded candidates have xyz = 0, i.e. plane = 0 / 1e-12 = 0 exactly, and need no storage.
__global__ void __launch_bounds__(256) conv3d_plane_kernel(const float* __restrict__ xyz, float* __restrict__ plane,
                                                           size_t hw, size_t total) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const size_t tb = i / hw, pix = i - tb * hw;
    const float* p = xyz + tb * 3 * hw + pix;
    float* o = plane + tb * 3 * hw + pix;
    const float x = __ldg(p), y = __ldg(p + hw), z = __ldg(p + 2 * hw);
    const float den = z + 1e-12f;
    o[0] = __fdiv_rn(x, den); o[hw] = __fdiv_rn(y, den); o[2 * hw] = __fdiv_rn(z, den);
  }
}

__device__ __forceinline__ void load_plane(const C3Args& a, int t, int b, int y, int x, bool inside, float v[3]) {
  if (!inside) { v[0] = v[1] = v[2] = 0.f; return; }
  const size_t hw = (size_t)a.h * a.w;
  const float* p = a.plane + ((size_t)(t * a.bs + b) * 3) * hw + (size_t)y * a.w + x;
  v[0] = __ldg(p); v[1] = __ldg(p + hw); v[2] = __ldg(p + 2 * hw);
}

// sum_c (plane_c - centre_plane_c)^2  (:493-495)
__device__ __forceinline__ float plane_sq(const float v[3], const float cpl[3]) {
  const float d0 = v[0] - cpl[0], d1 = v[1] - cpl[1], d2 = v[2] - cpl[2];
  return __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0)));
}

// FIXED: the reference configuration (3 x 3 window, 4 frames = 36 candidates, 9 neighbours), every loop unrolled.
// The winners are kept as a sorted list in registers: a candidate is inserted behind every entry whose key is <= its
// own (compare-and-select per slot, no branches) -- "ascending key, ties -> lowest candidate index", because candidates
// arrive in index order.
template <bool SELECT, bool FIXED>
__global__ void __launch_bounds__(128) conv3d_rank_kernel(C3Args a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = a.bs * a.oh * a.ow;
  float local_max = 0.f;
  if (m < M) {
    const int b = m / (a.oh * a.ow), r = m - b * a.oh * a.ow, oy = r / a.ow, ox = r - oy * a.ow;
    const int kw = FIXED ? 3 : a.k, tl = FIXED ? 4 : a.tl;
    const int ncand = FIXED ? 36 : a.k * a.k * a.tl;
    const int nb = FIXED ? 9 : a.nb;
    const int pad = (kw - 1) / 2;
    // centre candidate: (ky, kx) = (pad, pad), t = 0   (tidx = (k*k // 2) * tl, :491); always inside the image
    int cy, cx; bool cin;
    cand_pos(a, oy, ox, pad, pad, cy, cx, cin);
    float cpl[3];
    load_plane(a, 0, b, cy, cx, cin, cpl);
    constexpr int NBMAX = FIXED ? 9 : 16;
    constexpr float SENTINEL = 3.402823466e38f;       // FLT_MAX: above every key (keys are <= global max + 1)
    float bestk[NBMAX];
    int besti[NBMAX];
#pragma unroll
    for (int j = 0; j < NBMAX; ++j) { bestk[j] = SENTINEL; besti[j] = 0; }
    const float big = SELECT ? __ldg(a.gmax) + 1.0f : 0.f;
    const size_t hw = (size_t)a.h * a.w;
#pragma unroll (FIXED ? 36 : 1)
    for (int c = 0; c < ncand; ++c) {
      const int t = c % tl, kk = c / tl, ky = kk / kw, kx = kk - ky * kw;   // index = (ky*k + kx)*tl + t, :485
      int y, x; bool in;
      cand_pos(a, oy, ox, ky, kx, y, x, in);
      float v[3];
      load_plane(a, t, b, y, x, in, v);
      const float sq = plane_sq(v, cpl);
      if (SELECT) {
        const float mk = in ? __ldg(a.mask + (size_t)(t * a.bs + b) * hw + (size_t)y * a.w + x) : 0.f;
        const float key = __fmaf_rn(mk, sq, __fmul_rn(1.0f - mk, big));     // mask*sq + (1-mask)*(max+1), :497
        {   // branch-free insertion (an early "not better than the current worst" exit saved < 10 %)
          bool lt[NBMAX];
#pragma unroll
          for (int j = 0; j < NBMAX; ++j) lt[j] = key < bestk[j];           // monotone: the list is sorted
#pragma unroll
          for (int j = NBMAX - 1; j >= 1; --j) {                           // descending: reads the old neighbours
            bestk[j] = lt[j - 1] ? bestk[j - 1] : (lt[j] ? key : bestk[j]);
            besti[j] = lt[j - 1] ? besti[j - 1] : (lt[j] ? c : besti[j]);
          }
          bestk[0] = lt[0] ? key : bestk[0];
          besti[0] = lt[0] ? c : besti[0];
        }
      } else {
        local_max = fmaxf(local_max, sq);
      }
    }
    if (SELECT) {
      float cxyz[3];
      load_xyz(a, 0, b, cy, cx, cin, cxyz);
#pragma unroll
      for (int j = 0; j < NBMAX; ++j) {
        if (j >= nb) break;
        const int best = besti[j];
        a.idx[(size_t)m * nb + j] = (uint8_t)best;
        const int t = best % tl, kk = best / tl, ky = kk / kw, kx = kk - ky * kw;
        int y, x; bool in;
        cand_pos(a, oy, ox, ky, kx, y, x, in);
        float v[3];
        load_xyz(a, t, b, y, x, in, v);
        float* o = a.xyz_nb + ((size_t)m * nb + j) * 3;
        o[0] = v[0] - cxyz[0]; o[1] = v[1] - cxyz[1]; o[2] = v[2] - cxyz[2];     // xyz_local, :492
      }
    }
  }
  if (!SELECT) {
    // global max of the squared distances (:497 xyz_sq.max()): order-independent, so an integer atomicMax on the
    // bit pattern of the non-negative floats is deterministic
    local_max = fmaxf(local_max, 0.f);
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(a.gmax), __float_as_int(local_max));
  }
}

// feat_nb[m, j, :] = feat[t, b, :, y, x] of the selected candidate (zeros when it lies in the padding).
// The source is channel-planar (stride h*w between channels), the destination channel-contiguous, so the copy is a
// transpose.  One WARP owns a (32 consecutive output pixels) x (one slot) item: it reads with lanes = pixels
// (neighbouring pixels pick neighbouring candidates: mostly the same 128-byte lines) with all 32 channel loads in
// flight, turns the 32 x 32 tile around in its private slice of shared memory (__syncwarp only, no CTA barrier) and
// writes with lanes = channels, one full 128-byte line per pixel.
constexpr int FT = 32, FW = 8;   // pixels per item, warps per CTA
__global__ void __launch_bounds__(32 * FW) conv3d_feat_gather_kernel(C3Args a) {
  __shared__ float tiles[FW][32][FT + 1];
  const int M = a.bs * a.oh * a.ow;
  const size_t hw = (size_t)a.h * a.w;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  float (*tile)[FT + 1] = tiles[wv];
  const size_t items = (size_t)((M + FT - 1) / FT) * a.nb;
  for (size_t it = (size_t)blockIdx.x * FW + wv; it < items; it += (size_t)gridDim.x * FW) {
    const int j = (int)(it % a.nb), m0 = (int)(it / a.nb) * FT, m = m0 + lane;
    long long off = -1;                              // element offset of channel 0 of my pixel's candidate
    if (m < M) {
      const int b = m / (a.oh * a.ow), r = m - b * a.oh * a.ow, oy = r / a.ow, ox = r - oy * a.ow;
      const int c = a.idx[(size_t)m * a.nb + j];
      const int t = c % a.tl, kk = c / a.tl, ky = kk / a.k, kx = kk - ky * a.k;
      int y, x; bool in;
      cand_pos(a, oy, ox, ky, kx, y, x, in);
      if (in) off = (long long)((size_t)(t * a.bs + b) * a.C * hw + (size_t)y * a.w + x);
    }
    for (int c0 = 0; c0 < a.C; c0 += 32) {
      const int nc = min(32, a.C - c0);
      float v[32];
#pragma unroll
      for (int ch = 0; ch < 32; ++ch) v[ch] = (off >= 0 && ch < nc) ? __ldg(a.feat + off + (size_t)(c0 + ch) * hw) : 0.f;
#pragma unroll
      for (int ch = 0; ch < 32; ++ch) tile[ch][lane] = v[ch];
      __syncwarp();
      const int np = min(FT, M - m0);
      if (lane < nc) {
        float* o = a.feat_nb + ((size_t)m0 * a.nb + j) * a.C + c0 + lane;
#pragma unroll 8
        for (int p = 0; p < np; ++p) __stcs(o + (size_t)p * a.nb * a.C, tile[lane][p]);
      }
      __syncwarp();
    }
  }
}

struct C3BwdArgs {
  const float* g_xyz_nb; const float* g_feat_nb; const uint8_t* idx;
  float* g_xyz; float* g_feat;
  int tl, bs, C, h, w, k, stride, nb, oh, ow;
};

// one thread per source element (t, b, y, x): deterministic gather of the gradients of every slot that selected it
__global__ void __launch_bounds__(256) conv3d_gather_bwd_kernel(C3BwdArgs a) {
  const size_t hw = (size_t)a.h * a.w;
  const size_t total = (size_t)a.tl * a.bs * hw;
  const int pad = (a.k - 1) / 2;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int tb = (int)(i / hw), pix = (int)(i - (size_t)tb * hw), y = pix / a.w, x = pix - y * a.w;
    const int t = tb / a.bs, b = tb - t * a.bs;
    float gx[3] = {0.f, 0.f, 0.f};
    int match[MAX_KK];  // (m * nb + j) of the slots that picked this element; <= k*k of them
    int nmatch = 0;
    for (int ky = 0; ky < a.k; ++ky)
      for (int kx = 0; kx < a.k; ++kx) {
        // output pixel (oy, ox) sees (y, x) as candidate (ky, kx) iff oy*stride + ky - pad == y
        const int ny = y - ky + pad, nx = x - kx + pad;
        if (ny < 0 || nx < 0 || ny % a.stride || nx % a.stride) continue;
        const int oy = ny / a.stride, ox = nx / a.stride;
        if (oy >= a.oh || ox >= a.ow) continue;
        const int m = (b * a.oh + oy) * a.ow + ox;
        const int cand = (ky * a.k + kx) * a.tl + t;
        for (int j = 0; j < a.nb; ++j)
          if (a.idx[(size_t)m * a.nb + j] == cand) { match[nmatch++] = m * a.nb + j; break; }
      }
    if (a.g_xyz) {
      for (int q = 0; q < nmatch; ++q)
        for (int c = 0; c < 3; ++c) gx[c] += __ldg(a.g_xyz_nb + (size_t)match[q] * 3 + c);
      if (t == 0 && y % a.stride == 0 && x % a.stride == 0 && y / a.stride < a.oh && x / a.stride < a.ow) {
        const int m = (b * a.oh + y / a.stride) * a.ow + x / a.stride;      // this element is the centre of pixel m
        for (int j = 0; j < a.nb; ++j)
          for (int c = 0; c < 3; ++c) gx[c] -= __ldg(a.g_xyz_nb + ((size_t)m * a.nb + j) * 3 + c);
      }
      for (int c = 0; c < 3; ++c) a.g_xyz[((size_t)tb * 3 + c) * hw + pix] = gx[c];
    }
    if (a.g_feat) {
      for (int ch = 0; ch < a.C; ++ch) {
        float s = 0.f;
        for (int q = 0; q < nmatch; ++q) s += __ldg(a.g_feat_nb + (size_t)match[q] * a.C + ch);
        a.g_feat[((size_t)tb * a.C + ch) * hw + pix] = s;
      }
    }
  }
}

// g_feat[t, b, :, y, x] = sum of g_feat_nb over the slots that selected (t, b, y, x): the transpose of the forward
// copy, with the same warp-private scheme.  A warp takes 32 consecutive pixels of one (t, b) plane: with lanes = pixels
// it finds, per window position, the slot that picked the pixel; then a quarter-warp per pixel sums the matching rows
// of g_feat_nb with 128-bit loads (8 lanes x float4 = one 32-channel row, 4 pixels per instruction, 9 in flight); the
// sums go through the warp's shared-memory tile and are written with lanes = pixels.  Deterministic: fixed summation
// order, no atomics.  Needs C % 4 == 0 and a window of at most 5 x 5 (else conv3d_gather_bwd_kernel does the work).
constexpr int BW = 4, BKK = 25;
template <int KK, bool S1>   // KK = 9: the reference's 3 x 3 window, fully unrolled (0: generic); S1: stride == 1
__global__ void __launch_bounds__(32 * BW) conv3d_feat_bwd_kernel(C3BwdArgs a, int z0) {
  __shared__ float tiles[BW][32][FT + 1];
  __shared__ int smatch[BW][BKK][FT];
  const size_t hw = (size_t)a.h * a.w;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int tb = z0 + blockIdx.y, t = tb / a.bs, b = tb - t * a.bs;
  const int pix0 = (blockIdx.x * BW + wv) * FT;
  if (pix0 >= (int)hw) return;                       // whole warp
  float (*tile)[FT + 1] = tiles[wv];
  int (*match)[FT] = smatch[wv];
  const int pad = (a.k - 1) / 2, kk2 = KK > 0 ? KK : a.k * a.k, kw = KK == 9 ? 3 : a.k;
  const int pix = pix0 + lane;
  const int y = pix / a.w, x = pix - y * a.w;
#pragma unroll
  for (int kk = 0; kk < kk2; ++kk) {
    const int ky = kk / kw, kx = kk - ky * kw;
    int found = -1;
    // output pixel (oy, ox) sees (y, x) as candidate (ky, kx) iff oy*stride + ky - pad == y
    const int ny = y - ky + pad, nx = x - kx + pad;
    const bool on_grid = S1 || (ny % a.stride == 0 && nx % a.stride == 0);
    if (pix < (int)hw && ny >= 0 && nx >= 0 && on_grid) {
      const int oy = S1 ? ny : ny / a.stride, ox = S1 ? nx : nx / a.stride;
      if (oy < a.oh && ox < a.ow) {
        const int m = (b * a.oh + oy) * a.ow + ox;
        const int cand = kk * a.tl + t;
        const uint8_t* row = a.idx + (size_t)m * a.nb;
        // no early exit: the loads of one row are independent (a candidate occupies at most one slot)
#pragma unroll 9
        for (int j = 0; j < a.nb; ++j)
          if (__ldg(row + j) == cand) found = m * a.nb + j;
      }
    }
    match[kk][lane] = found;
  }
  __syncwarp();
  const int pp = lane >> 3, cq = lane & 7;           // pixel within the group of 4, channel quad
  for (int c0 = 0; c0 < a.C; c0 += 32) {
    const int nc = min(32, a.C - c0);
    const bool live = 4 * cq < nc;
    const float* src = a.g_feat_nb + c0 + 4 * cq;
    for (int p0 = 0; p0 < FT; p0 += 4) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        if (KK > 0) {
          float4 v[KK > 0 ? KK : 1];
#pragma unroll
          for (int q = 0; q < KK; ++q) {
            const int mt = match[q][p0 + pp];
            v[q] = mt >= 0 ? __ldg(reinterpret_cast<const float4*>(src + (size_t)mt * a.C)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int q = 0; q < KK; ++q) { sum.x += v[q].x; sum.y += v[q].y; sum.z += v[q].z; sum.w += v[q].w; }
        } else {
          for (int kk = 0; kk < kk2; ++kk) {
            const int mt = match[kk][p0 + pp];
            if (mt >= 0) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)mt * a.C));
              sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
          }
        }
      }
      tile[4 * cq][p0 + pp] = sum.x; tile[4 * cq + 1][p0 + pp] = sum.y;
      tile[4 * cq + 2][p0 + pp] = sum.z; tile[4 * cq + 3][p0 + pp] = sum.w;
    }
    __syncwarp();
    if (pix < (int)hw) {
      float* o = a.g_feat + ((size_t)tb * a.C + c0) * hw + pix;
#pragma unroll 8
      for (int ch = 0; ch < nc; ++ch) o[(size_t)ch * hw] = tile[ch][lane];
    }
    __syncwarp();
  }
}

inline int flat_grid(size_t total, int threads) {
  const size_t want = (total + threads - 1) / threads, cap = 148 * 32;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

int conv3d_out_size(int n, int k, int stride) { return (n + 2 * ((k - 1) / 2) - k) / stride + 1; }

size_t conv3d_scratch_elems(int tl, int bs, int h, int w) { return 4 + (size_t)tl * bs * 3 * h * w; }

// scratch: [0] = global max of the squared plane distances, [4 ...] = plane coordinates [tl,bs,3,h,w]
// neighbour selection only: xyz_nb and idx (depends on xyz and mask, not on the features)
int conv3d_rank(const float* xyz, const float* mask, float* xyz_nb, uint8_t* idx, float* scratch, int tl, int bs, int h, int w,
                int k, int stride, int nb, cudaStream_t s) {
  float* gmax = scratch;
  C3Args a{xyz, nullptr, mask, xyz_nb, nullptr, idx, gmax, scratch + 4, tl, bs, 0, h, w, k, stride, nb,
           conv3d_out_size(h, k, stride), conv3d_out_size(w, k, stride)};
  const int M = bs * a.oh * a.ow;
  cudaError_t e = cudaMemsetAsync(gmax, 0, sizeof(float), s);
  if (e != cudaSuccess) { set_last_cuda_error(e); return DIS_ERR_CUDA_LAUNCH; }
  const size_t n_src = (size_t)tl * bs * h * w;
  conv3d_plane_kernel<<<flat_grid(n_src, 256), 256, 0, s>>>(xyz, a.plane, (size_t)h * w, n_src);
  if (k == 3 && tl == 4 && nb == 9) {   // the FIXED instantiation hard-codes 36 candidates AND 9 neighbours
    conv3d_rank_kernel<false, true><<<(M + 127) / 128, 128, 0, s>>>(a);
    conv3d_rank_kernel<true, true><<<(M + 127) / 128, 128, 0, s>>>(a);
  } else {
    conv3d_rank_kernel<false, false><<<(M + 127) / 128, 128, 0, s>>>(a);
    conv3d_rank_kernel<true, false><<<(M + 127) / 128, 128, 0, s>>>(a);
  }
  return check_launch();
}

// feature gather for a given selection (the selection of a FuseNet level serves every Conv3D layer of that level and the
// checkpoint recompute: the ranking is paid once)
int conv3d_gather_features(const float* feat, const uint8_t* idx, float* feat_nb, int tl, int bs, int C, int h, int w, int k,
                           int stride, int nb, cudaStream_t s) {
  C3Args a{nullptr, feat, nullptr, nullptr, feat_nb, const_cast<uint8_t*>(idx), nullptr, nullptr, tl, bs, C, h, w, k, stride, nb,
           conv3d_out_size(h, k, stride), conv3d_out_size(w, k, stride)};
  const int M = bs * a.oh * a.ow;
  conv3d_feat_gather_kernel<<<flat_grid((size_t)((M + FT - 1) / FT) * nb, FW), 32 * FW, 0, s>>>(a);
  return check_launch();
}

int conv3d_gather_forward(const float* xyz, const float* feat, const float* mask, float* xyz_nb, float* feat_nb,
                          uint8_t* idx, float* scratch, int tl, int bs, int C, int h, int w, int k, int stride, int nb,
                          cudaStream_t s) {
  if (int rc = conv3d_rank(xyz, mask, xyz_nb, idx, scratch, tl, bs, h, w, k, stride, nb, s)) return rc;
  return conv3d_gather_features(feat, idx, feat_nb, tl, bs, C, h, w, k, stride, nb, s);
}

int conv3d_gather_backward(const float* g_xyz_nb, const float* g_feat_nb, const uint8_t* idx, float* g_xyz, float* g_feat,
                           int tl, int bs, int C, int h, int w, int k, int stride, int nb, cudaStream_t s) {
  C3BwdArgs a{g_xyz_nb, g_feat_nb, idx, g_xyz, g_feat, tl, bs, C, h, w, k, stride, nb,
              conv3d_out_size(h, k, stride), conv3d_out_size(w, k, stride)};
  const bool fast_feat = g_feat && k * k <= BKK && C % 4 == 0 && (reinterpret_cast<uintptr_t>(g_feat_nb) & 15) == 0;
  if (g_xyz || (g_feat && !fast_feat)) {
    C3BwdArgs ax = a;
    if (fast_feat) ax.g_feat = nullptr;
    if (!g_xyz) ax.g_xyz = nullptr;
    conv3d_gather_bwd_kernel<<<flat_grid((size_t)tl * bs * h * w, 256), 256, 0, s>>>(ax);
    if (int rc = check_launch()) return rc;
  }
  if (fast_feat) {
    const int tiles = (h * w + FT - 1) / FT;
    for (int z0 = 0; z0 < tl * bs; z0 += 65535) {
      const int nz = tl * bs - z0 < 65535 ? tl * bs - z0 : 65535;
      const dim3 grid((tiles + BW - 1) / BW, nz);
      if (k == 3 && stride == 1) conv3d_feat_bwd_kernel<9, true><<<grid, 32 * BW, 0, s>>>(a, z0);
      else if (k == 3) conv3d_feat_bwd_kernel<9, false><<<grid, 32 * BW, 0, s>>>(a, z0);
      else conv3d_feat_bwd_kernel<0, false><<<grid, 32 * BW, 0, s>>>(a, z0);
      if (int rc = check_launch()) return rc;
    }
  }
  return DIS_OK;
}

}  // namespace dis
