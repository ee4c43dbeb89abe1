// Probe (GPU box): issue/throughput of packed fp32x2 FMA (sm_100 FFMA2) vs scalar FFMA, register-resident chains.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 1.0001f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { x[i].x = fmaf(x[i].x, A.x, B.x); x[i].y = fmaf(x[i].y, A.y, B.y); }
        else if (MODE == 1) x[i] = ffma2(x[i], A, B);
        else x[i] = ffma2(x[i], x[(i + 1) & 7], B);       // three distinct register operands
      }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  if (s == 123456789.f) out[0] = s;
}
template <int MODE> double run(float* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 8, threads = 256, iters = 2048;
  double best = 0;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 2 * 8 * 16 * (double)iters * blocks * threads;
    if (r && fl / (ms * 1e-3) > best) best = fl / (ms * 1e-3);
  }
  return best * 1e-12;
}
int main() {
  float* d; cudaMalloc(&d, 4);
  printf("scalar FFMA x2       : %.1f TFLOP/s\n", run<0>(d));
  printf("FFMA2 (const operands): %.1f TFLOP/s\n", run<1>(d));
  printf("FFMA2 (3 reg operands): %.1f TFLOP/s\n", run<2>(d));
  return cudaGetLastError() != cudaSuccess;
}
