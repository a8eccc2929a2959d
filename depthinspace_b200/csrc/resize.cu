// Bilinear resize with align_corners = True, fused with what FuseNet does next to it:
//   resize_like            model/multi_frame_networks.py:42-52   (plain)
//   resize_flow_like       :54-68   (every flow of the dict in ONE launch; x / y channel rescaled by the size ratio)
//   resize_flow_masks_like :70-81 and (resize_like(mask) > 0.5).float() at :394   (threshold fused)
// The reference issues F.interpolate + two in-place multiplies per dict entry (36 launches for the 12 flows of a
// 4-frame track, per resolution).  Arithmetic follows ATen's upsample_bilinear2d_out_frame (UpSampleBilinear2d.cu):
// source index = scale * dst with scale = (in - 1) / (out - 1) in fp32, lambda weights from the truncated index, and the
// blend in the contraction order nvcc gives ATen's expression (fma(w0, a, w1 * b) per row, fma(h0, top, h1 * bottom)).
// HBM-bound: one thread per output pixel, coordinates computed once and reused for every channel / dict entry plane.
#include "common.cuh"

namespace dis {
namespace {

constexpr int RESIZE_MAX_PLANES = 56;   // dict entries per launch (8 frames: 8 * 7 ordered pairs)

struct ResizeArgs {
  const float* in[RESIZE_MAX_PLANES];
  float* out[RESIZE_MAX_PLANES];
  int N, C, H, W, oh, ow;
  float rh, rw;        // (H - 1) / (oh - 1), (W - 1) / (ow - 1)  (0 when the output has one row / column)
  float sx, sy;        // flow mode: channel 0 *= sx, channel 1 *= sy
  int mode;            // 0 plain, 1 flow, 2 mask (> 0.5 -> 1.0 / 0.0)
};

struct Src {
  int i0, ip;          // first source index, 1 if a second one exists
  float l0, l1;        // weights of i0 and i0 + ip
};
__device__ __forceinline__ Src source(float scale, int dst, int size) {
  const float r = __fmul_rn(scale, (float)dst);
  Src s;
  s.i0 = (int)r;
  s.ip = (s.i0 < size - 1) ? 1 : 0;
  s.l1 = __fsub_rn(r, (float)s.i0);
  s.l0 = __fsub_rn(1.0f, s.l1);
  return s;
}

__global__ void __launch_bounds__(256) resize_bilinear_kernel(ResizeArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.oh * a.ow) return;
  const int oy = p / a.ow, ox = p - oy * a.ow;
  const Src sy = source(a.rh, oy, a.H), sx = source(a.rw, ox, a.W);
  const int n = blockIdx.y;
  const float* __restrict__ in = a.in[blockIdx.z] + (size_t)n * a.C * a.H * a.W;
  float* __restrict__ out = a.out[blockIdx.z] + (size_t)n * a.C * a.oh * a.ow;
  const int o00 = sy.i0 * a.W + sx.i0;
  const int dxo = sx.ip, dyo = sy.ip * a.W;
  for (int c = 0; c < a.C; ++c) {
    const float* pl = in + (size_t)c * a.H * a.W + o00;
    const float v00 = __ldg(pl), v01 = __ldg(pl + dxo), v10 = __ldg(pl + dyo), v11 = __ldg(pl + dyo + dxo);
    const float top = __fmaf_rn(sx.l0, v00, __fmul_rn(sx.l1, v01));
    const float bot = __fmaf_rn(sx.l0, v10, __fmul_rn(sx.l1, v11));
    float v = __fmaf_rn(sy.l0, top, __fmul_rn(sy.l1, bot));
    if (a.mode == 1) v = __fmul_rn(v, c == 0 ? a.sx : a.sy);
    else if (a.mode == 2) v = v > 0.5f ? 1.0f : 0.0f;
    out[(size_t)c * a.oh * a.ow + p] = v;
  }
}

// adjoint of the plain resize (gradient w.r.t. the input): scatter with RED.ADD, as ATen's backward does
__global__ void __launch_bounds__(256) resize_bilinear_bwd_kernel(const float* __restrict__ g_out, float* __restrict__ g_in,
                                                                  int C, int H, int W, int oh, int ow, float rh, float rw) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= oh * ow) return;
  const int oy = p / ow, ox = p - oy * ow;
  const Src sy = source(rh, oy, H), sx = source(rw, ox, W);
  const int n = blockIdx.y;
  const float* go = g_out + (size_t)n * C * oh * ow + p;
  float* gi = g_in + (size_t)n * C * H * W + sy.i0 * W + sx.i0;
  for (int c = 0; c < C; ++c) {
    const float g = __ldg(go + (size_t)c * oh * ow);
    float* pl = gi + (size_t)c * H * W;
    atomicAdd(pl, sy.l0 * sx.l0 * g);
    atomicAdd(pl + sx.ip, sy.l0 * sx.l1 * g);
    atomicAdd(pl + sy.ip * W, sy.l1 * sx.l0 * g);
    atomicAdd(pl + sy.ip * W + sx.ip, sy.l1 * sx.l1 * g);
  }
}

float area_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f; }

}  // namespace

int resize_bilinear_forward(const float* const* ins, float* const* outs, int count, int N, int C, int H, int W, int oh, int ow,
                            int mode, cudaStream_t s) {
  for (int z0 = 0; z0 < count; z0 += RESIZE_MAX_PLANES) {
    ResizeArgs a{};
    const int nz = count - z0 < RESIZE_MAX_PLANES ? count - z0 : RESIZE_MAX_PLANES;
    for (int i = 0; i < nz; ++i) {
      a.in[i] = ins[z0 + i];
      a.out[i] = outs[z0 + i];
    }
    a.N = N; a.C = C; a.H = H; a.W = W; a.oh = oh; a.ow = ow;
    a.rh = area_scale(H, oh); a.rw = area_scale(W, ow);
    a.sx = (float)((double)ow / (double)W); a.sy = (float)((double)oh / (double)H);
    a.mode = mode;
    for (int n0 = 0; n0 < N; n0 += 65535) {
      ResizeArgs b = a;
      for (int i = 0; i < nz; ++i) {
        b.in[i] += (size_t)n0 * C * H * W;
        b.out[i] += (size_t)n0 * C * oh * ow;
      }
      const int nb = N - n0 < 65535 ? N - n0 : 65535;
      resize_bilinear_kernel<<<dim3((oh * ow + 255) / 256, nb, nz), 256, 0, s>>>(b);
      if (int rc = check_launch()) return rc;
    }
  }
  return DIS_OK;
}

int resize_bilinear_backward(const float* g_out, float* g_in, int N, int C, int H, int W, int oh, int ow, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(g_in, 0, sizeof(float) * (size_t)N * C * H * W, s);
  if (e != cudaSuccess) { set_last_cuda_error(e); return DIS_ERR_CUDA_LAUNCH; }
  for (int n0 = 0; n0 < N; n0 += 65535) {
    const int nb = N - n0 < 65535 ? N - n0 : 65535;
    resize_bilinear_bwd_kernel<<<dim3((oh * ow + 255) / 256, nb), 256, 0, s>>>(
        g_out + (size_t)n0 * C * oh * ow, g_in + (size_t)n0 * C * H * W, C, H, W, oh, ow, area_scale(H, oh), area_scale(W, ow));
    if (int rc = check_launch()) return rc;
  }
  return DIS_OK;
}

}  // namespace dis
