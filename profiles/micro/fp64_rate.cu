// fp64 instruction throughput per SM on sm_100a (B200): DFMA, DADD, DMUL, DSETP+SEL (min), fmin, mixed with FP32/INT.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, double a, double b, int iters) {
  double x0 = a + threadIdx.x, x1 = a * 2 + threadIdx.x, x2 = a * 3, x3 = a * 4 + threadIdx.x, x4 = a * 5, x5 = a * 6, x6 = a * 7, x7 = a * 8;
  for (int i = 0; i < iters; i++) {
#define STEP(x)                                             \
  if (OP == 0) x = fma(x, b, a);                            \
  else if (OP == 1) x = x + b;                              \
  else if (OP == 2) x = x * b;                              \
  else if (OP == 3) x = (x < b) ? x + 1.0 : b;              \
  else if (OP == 4) x = fmin(x, b) + a;                     \
  else if (OP == 5) x = (x < a ? x : a) + b;                \
  else if (OP == 6) x = copysign(fabs(x) + a, b);           \
  else if (OP == 7) x = a / (x + b);
    STEP(x0) STEP(x1) STEP(x2) STEP(x3) STEP(x4) STEP(x5) STEP(x6) STEP(x7)
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <int OP>
void run(const char* name, int ops_per_step) {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  const int iters = 4096, blocks = 148 * 2, threads = 1024;
  k<OP><<<blocks, threads>>>(out, 1.0000001, 0.9999999, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(out, 1.0000001, 0.9999999, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)blocks * threads * iters * 8;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %8.3f ms  %7.2f Gstep/s  = %6.2f steps/clk/SM (at %d MHz)  [%d fp64 op(s) per step]\n", name, ms, n / ms * 1e-6,
         n / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000, ops_per_step);
  cudaFree(out);
}
int main() {
  run<0>("DFMA", 1); run<1>("DADD", 1); run<2>("DMUL", 1); run<3>("DSETP + DADD + SEL", 2); run<4>("fmin + DADD", 2);
  run<5>("cmp-select min + DADD", 2); run<6>("fabs/copysign + DADD", 1); run<7>("DIV + DADD", 2);
  return 0;
}
