// Accuracy probe (run on the GPU box): MUFU-seeded fp64 reciprocal / rsqrt with Newton steps vs IEEE results.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double rcp_n(double a, int steps) {
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  for (int i = 0; i < steps; ++i) { double e = fma(-a, y, 1.0); y = fma(y, e, y); }
  return y;
}
__device__ double rsqrt_n(double a, int steps) {
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = 0.5 * a;
  for (int i = 0; i < steps; ++i) { double e = fma(-h * y, y, 0.5); y = fma(y, e, y); }
  return y;
}
__device__ double rcp_3(double a) {   // one third-order step
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a, y, 1.0);
  return fma(fma(y, e, y), e, y);
}
__device__ double rsqrt_3(double a) {
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-(a * y), y, 1.0);
  return fma(y * e, fma(e, 0.375, 0.5), y);
}
__global__ void probe(double* out) {
  // out[2*s]: max rel err rcp with s steps; out[2*s+1]: rsqrt
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long st = 0x9E3779B97F4A7C15ull * (tid + 1);
  double mr[4] = {0, 0, 0, 0}, ms[4] = {0, 0, 0, 0}, m3r = 0, m3s = 0;
  for (int k = 0; k < 4096; ++k) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    const double u = (double)(st >> 11) * (1.0 / 9007199254740992.0);
    const double a = exp2(-40.0 + 80.0 * u) * (1.0 + u);   // 1e-12 .. 1e12
    for (int s = 0; s < 4; ++s) {
      mr[s] = fmax(mr[s], fabs(rcp_n(a, s) * a - 1.0));
      const double r = rsqrt_n(a, s);
      ms[s] = fmax(ms[s], fabs(r * r * a - 1.0) * 0.5);
    }
    m3r = fmax(m3r, fabs(fma(rcp_3(a), a, -1.0)));
    const double r3 = rsqrt_3(a);
    m3s = fmax(m3s, fabs(fma(r3 * r3, a, -1.0)) * 0.5);
  }
  for (int s = 0; s < 4; ++s) {
    atomicMax((unsigned long long*)&out[2 * s], (unsigned long long)__double_as_longlong(mr[s]));
    atomicMax((unsigned long long*)&out[2 * s + 1], (unsigned long long)__double_as_longlong(ms[s]));
  }
  atomicMax((unsigned long long*)&out[8], (unsigned long long)__double_as_longlong(m3r));
  atomicMax((unsigned long long*)&out[9], (unsigned long long)__double_as_longlong(m3s));
}
int main() {
  double* d; cudaMalloc(&d, 10 * sizeof(double)); cudaMemset(d, 0, 10 * sizeof(double));
  probe<<<64, 128>>>(d);
  double h[10]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  for (int s = 0; s < 4; ++s) printf("newton steps %d: rcp max rel err %.3e   rsqrt max rel err %.3e\n", s, h[2 * s], h[2 * s + 1]);
  printf("one third-order step: rcp max rel err %.3e   rsqrt max rel err %.3e\n", h[8], h[9]);
  return cudaGetLastError() != cudaSuccess;
}
