// Accuracy probe (run on the GPU box): the fp64 sincos / atan2 of csrc/gik_core.cuh against the CUDA library forms.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -I motion-planning-and-control-for-dual-manipulator-robot_b200/csrc \
//        -o build/f64_trig tools/probes/f64_trig.cu && build/f64_trig
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "gik_core.cuh"

__global__ void probe(double* out, double range) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long st = 0x9E3779B97F4A7C15ull * (tid + 1);
  double es = 0, ec = 0, ea = 0, er = 0;
  for (int k = 0; k < 4096; ++k) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    const double u = (double)(st >> 11) * (1.0 / 9007199254740992.0);
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    const double v = (double)(st >> 11) * (1.0 / 9007199254740992.0);
    const double x = range * (2.0 * u - 1.0);
    double s, c, s0, c0;
    gik::sincos_<true>(x, s, c);
    sincos(x, &s0, &c0);
    es = fmax(es, fabs(s - s0)); ec = fmax(ec, fabs(c - c0));
    // atan2 with y >= 0 over all magnitudes of the ratio, including tiny angles (relative error matters there)
    const double y = exp2(-60.0 * v) * u, xx = (2.0 * v - 1.0);
    const double a = gik::atan2_pos(y, xx), a0 = atan2(y, xx);
    ea = fmax(ea, fabs(a - a0));
    er = fmax(er, fabs(a - a0) / fmax(a0, 1e-300));
  }
  atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(es));
  atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(ec));
  atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(ea));
  atomicMax((unsigned long long*)&out[3], (unsigned long long)__double_as_longlong(er));
}

int main() {
  double* d; cudaMalloc(&d, 4 * sizeof(double));
  const double ranges[3] = {3.6, 100.0, 1e6};
  for (double range : ranges) {
    cudaMemset(d, 0, 4 * sizeof(double));
    probe<<<256, 128>>>(d, range);
    double h[4]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("|x| <= %-8g sincos max abs err: sin %.3e cos %.3e | atan2_pos: max abs err %.3e, max rel err %.3e\n", range,
           h[0], h[1], h[2], h[3]);
  }
  return cudaGetLastError() != cudaSuccess;
}
