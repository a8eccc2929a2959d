// Micro-benchmark: cost of the feature-gradient scatter as 4 scalar RED.ADD per value vs one red.global.add.v4.f32 per
// 4 target cells (sm_90+), same addresses touched.  nvcc -arch=sm_100a -O3 tools/ubench_red.cu -o tools/bin/ubench_red
#include <cstdio>
#include <cuda_runtime.h>

constexpr int W = 216, H = 256, C = 32;

__global__ void scalar4(float* __restrict__ gx, const float* __restrict__ go, int bs) {
  const size_t hw = (size_t)H * W, total = (size_t)bs * hw;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t n = idx / hw;
    const int pix = (int)(idx - n * hw), h = pix / W, w = pix - h * W;
    const int y0 = min(h + 1, H - 2), x0 = min(w + 2, W - 2);
    for (int c = 0; c < C; ++c) {
      const float g = go[(n * C + c) * hw + pix];
      float* p = gx + (n * C + c) * hw + (size_t)y0 * W + x0;
      atomicAdd(p, 0.25f * g); atomicAdd(p + 1, 0.25f * g); atomicAdd(p + W, 0.25f * g); atomicAdd(p + W + 1, 0.25f * g);
    }
  }
}
__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// one warp = 32 consecutive pixels of a row; lanes 0..9 flush row y0, lanes 16..25 row y0+1 (aligned float4 groups)
__global__ void vector4(float* __restrict__ gx, const float* __restrict__ go, int bs) {
  const size_t hw = (size_t)H * W, total = (size_t)bs * hw;
  const int lane = threadIdx.x & 31;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total + 31; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t base = idx - lane;                 // first pixel of the warp
    if (base >= total) break;
    const size_t n = base / hw;
    const int pix0 = (int)(base - n * hw), h = pix0 / W, w0 = pix0 - h * W;
    const int y0 = min(h + 1, H - 2);
    const int xg = ((min(w0 + 2, W - 40)) / 4) * 4;  // aligned start of the 40-wide target span
    const int row = lane >> 4, grp = lane & 15;
    const bool act = grp < 10;
    for (int c = 0; c < C; ++c) {
      const float g = idx < total ? go[(n * C + c) * hw + pix0 + lane] : 0.f;
      const float v = __shfl_sync(0xffffffffu, g, (grp * 3) & 31) * 0.25f;   // stand-in for the staged sums
      if (act) red_v4(gx + (n * C + c) * hw + (size_t)(y0 + row) * W + xg + 4 * grp, v, v, v, v);
    }
  }
}

int main() {
  const int bs = 32 * 3;   // 3 warp slots of a frame at once
  const size_t n = (size_t)bs * C * H * W;
  float *gx, *go;
  cudaMalloc(&gx, n * 4); cudaMalloc(&go, n * 4);
  cudaMemset(gx, 0, n * 4); cudaMemset(go, 0, n * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int k = 0; k < 2; ++k) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (k == 0) scalar4<<<148 * 32, 256>>>(gx, go, bs); else vector4<<<148 * 32, 256>>>(gx, go, bs);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%s: %.3f ms (%s)\n", k == 0 ? "4 scalar RED per value" : "1 v4 RED per 4 cells", ms, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
