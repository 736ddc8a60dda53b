// Microbenchmark: latency / single-warp throughput of DFMA, DADD, FFMA, SHFL, LDS on this GPU (one warp, one CTA).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, int n) {
  double a = out[0], b = out[1], c = out[2];
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x0 = fma(x0, b, c); }                       // dependent DFMA
  long long t1 = clock64();
  for (int i = 0; i < n; ++i) { x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c);
                                x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c); }
  long long t2 = clock64();
  float f0 = (float)a, fb = (float)b, fc = (float)c;
  for (int i = 0; i < n; ++i) { f0 = fmaf(f0, fb, fc); }
  long long t3 = clock64();
  double s = x0;
  for (int i = 0; i < n; ++i) { s = __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(s), 1), __shfl_up_sync(0xffffffffu, __double2loint(s), 1)); }
  long long t4 = clock64();
  for (int i = 0; i < n; ++i) { x1 = x1 + x0; }                                // dependent DADD
  long long t5 = clock64();
  float e = f0;
  for (int i = 0; i < n; ++i) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e)); }
  long long t6 = clock64();
  out[3] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + f0 + s + e;
  if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; clk[4] = t5 - t4; clk[5] = t6 - t5; }
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 64);
  double h[4] = {1.0, 0.999, 0.001, 0}; cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
  int n = 4096;
  for (int r = 0; r < 2; ++r) k<<<1, 32>>>(d, c, n);
  long long hc[6]; cudaMemcpy(hc, c, 48, cudaMemcpyDeviceToHost);
  printf("dep DFMA %.1f clk | 8 indep DFMA %.1f clk per DFMA | dep FFMA %.1f | dep 64-bit SHFL %.1f | dep DADD %.1f | dep EX2 %.1f\n",
         hc[0] / (double)n, hc[1] / (8.0 * n), hc[2] / (double)n, hc[3] / (double)n, hc[4] / (double)n, hc[5] / (double)n);
  return 0;
}
