// Latency micro-benchmarks for the fp64 building blocks of the Jacobi kernels (B200).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.000000001, acc = 0;
  long long t0, t1;
  // dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  acc += x;
  // dependent DADD chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = x + y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  acc += x;
  // dependent shuffle(double)+add chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 << (i % 5));
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  acc += x;
  // dependent rsqrt chain
  x = fabs(x) + 2.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) x = rsqrt(x) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  acc += x;
  // dependent division chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) x = 3.0 / x + 1.25;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  acc += x;
  // dependent sqrt chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) x = sqrt(x) + 2.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  acc += x;
  // dependent FFMA chain
  float f = (float)x, gq = 1.0001f;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) f = fmaf(f, gq, 1e-6f);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  acc += f;
  // LDS dependent chain
  __shared__ int idx[256];
  idx[threadIdx.x] = (threadIdx.x + 1) & 255;
  __syncthreads();
  int p = threadIdx.x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) p = idx[p];
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[threadIdx.x] = acc + p;
}
__global__ void tput(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  {
    double* o; cudaMalloc(&o, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps : {2, 4, 8}) {
      int thr = warps * 32 * 4 / 4;  // threads per CTA
      tput<<<148 * 4, thr>>>(o, 1000); cudaDeviceSynchronize();
      cudaEventRecord(e0);
      tput<<<148 * 4, thr>>>(o, 20000);
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fl = 2.0 * 64 * 20000.0 * 148 * 4 * thr;
      printf("FP64 FMA throughput, 4 CTAs/SM x %d threads: %.2f TFLOP/s (%.3f ms)\n", thr, fl / ms / 1e9, ms);
    }
  }
  double* o; long long* c;
  cudaMalloc(&o, 1024 * 8); cudaMalloc(&c, 64);
  for (int nt : {32, 128, 512}) {
    k<<<1, nt>>>(o, c, 1.0); cudaDeviceSynchronize();
    k<<<1, nt>>>(o, c, 1.0); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    printf("threads=%d  DFMA %.1f  DADD %.1f  shfl64+add %.1f  rsqrt64(+add) %.1f  div64(+add) %.1f  sqrt64(+add) %.1f  FFMA %.1f  LDS %.1f cycles/op\n",
           nt, h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 256.0, h[7] / 64.0);
  }
  return 0;
}
