// FP64 throughput probes: the denominators of the roofline for the FP64-bound kernels.
// vcgpu_fp64_peak times (a) independent DFMA chains and (b) independent mma.sync.m8n8k4.f64 chains
// on every SM and reports the best of a few repetitions, in TFLOP/s (FMA = 2 flop).
#pragma once
#include <cuda_runtime.h>

namespace vc {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = seed + threadIdx.x * 1e-9 + k;
  const double m = 1.0 + 1e-12, c = 1e-12;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fma(x[k], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  double d[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k][0] = d[k][1] = 0.0;
  const double a = seed + 1e-9 * threadIdx.x, b = seed - 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(d[k][0]), "+d"(d[k][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += d[k][0] + d[k][1];
  if (s == 123.456) out[0] = s;
}

}  // namespace vc
