// Pipe-throughput microbenchmarks used to size the census kernel (run on the B200 box):
//   scalar FFMA, packed FFMA2 (fma.rn.f32x2), MUFU.RSQ, and the census tap mix.
// Prints lane-ops per clock per SM from clock64() deltas.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ITER = 2048, CH = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* clk, float seed) {
  float a[CH], b[CH];
  unsigned long long p[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { a[i] = seed + i + threadIdx.x; b[i] = 0.5f + i; p[i] = (unsigned long long)__float_as_uint(a[i]) << 32 | __float_as_uint(b[i]); }
  const unsigned long long m = (unsigned long long)__float_as_uint(1.0001f) << 32 | __float_as_uint(0.9999f);
  const unsigned long long c = (unsigned long long)__float_as_uint(0.001f) << 32 | __float_as_uint(0.002f);
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) { a[i] = fmaf(a[i], 1.0001f, b[i]); b[i] = fmaf(b[i], 0.9999f, a[i]); }               // 2 FFMA (reg operands)
      if (MODE == 1) { p[i] = fma2(p[i], m, c); p[i] = fma2(p[i], m, p[(i + 1) % CH]); }                       // 2 FFMA2 = 4 lane-FMA
      if (MODE == 2) { a[i] = rsq(a[i] + 1.5f); b[i] = rsq(b[i] + 1.5f); }                                   // 2 MUFU + 2 FADD
      if (MODE == 3) {                                                                                       // census tap: 2 MUFU + ~15 FP32
        float de = a[i] - b[i], dt = b[i] - seed;
        float re = rsq(fmaf(de, de, 0.5f)), rt = rsq(fmaf(dt, dt, 0.5f));
        float d2 = de * re - dt * rt;
        float r3 = re * re * re;
        float u = __int_as_float(__float_as_int(r3) | (__float_as_int(d2) & 0x80000000));
        u = d2 == 0.f ? 0.f : u;
        a[i] += fabsf(d2); b[i] = fmaf(u, rt, b[i]) + u;
      }
      if (MODE == 4) { a[i] = a[i] * 1.0001f + 0.5f; b[i] = b[i] + a[i]; }                                    // FFMA(imm) + FADD
    }
  }
  long long t1 = clock64();
  float s = 0; 
#pragma unroll
  for (int i = 0; i < CH; ++i) s += a[i] + b[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter_chain, int ctas_per_sm) {
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int grid = nsm * ctas_per_sm;
  float* out; long long* clk; cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&clk, grid * 8);
  k<MODE><<<grid, 256>>>(out, clk, 1.0f); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, clk, 1.0f); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[grid]; cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  double lane_ops_per_cta = 256.0 * ITER * CH * lane_ops_per_iter_chain;
  printf("%-28s ctas/SM=%d  %.1f lane-ops/clk/SM  (%.3f ms, %.0f clk/CTA, eff clock %.0f MHz)\n", name, ctas_per_sm,
         lane_ops_per_cta * ctas_per_sm / avg, ms, avg, avg / (ms * 1e3));
  cudaFree(out); cudaFree(clk); delete[] h;
}

int main() {
  for (int c : {2, 4, 8}) {
    run<0>("FFMA scalar (3-reg)", 2, c);
    run<4>("FFMA imm + FADD", 2, c);
    run<1>("FFMA2 packed (lane-FMAs)", 4, c);
    run<2>("MUFU.RSQ (+FADD)", 2, c);
    run<3>("census tap (taps)", 1, c);
  }
  return 0;
}
