// Is FP64 mma.sync (DMMA) faster than / concurrent with DFMA on B200?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>  // 0 dmma only, 1 dfma only, 2 both
__global__ void k(double* out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x + i;
  double f[8];
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 0.5 + i;
  double a = 1.0000001, b = 0.9999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE != 1) dmma(c[2 * u], c[2 * u + 1], a, b);
      if (MODE != 0) { f[u] = fma(f[u], a, b); }
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  for (int i = 0; i < 8; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double* o, int bps = 4, int thr = 256) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * bps, iters = 20000;
  k<MODE><<<blocks, thr>>>(o, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<blocks, thr>>>(o, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warps = (double)blocks * thr / 32;
  double fl_mma = (MODE != 1) ? warps * iters * 8.0 * 512.0 : 0;
  double fl_fma = (MODE != 0) ? warps * iters * 8.0 * 64.0 : 0;
  printf("%-12s (%d CTA/SM x %d thr) %.3f ms  DMMA %.2f TF/s  DFMA %.2f TF/s  total %.2f TF/s\n", name, ms, fl_mma / ms / 1e9, fl_fma / ms / 1e9,
         (fl_mma + fl_fma) / ms / 1e9);
}
int main() {
  double* o; cudaMalloc(&o, 148 * 4 * 256 * 8);
  run<0>("dmma only", o); run<1>("dfma only", o); run<2>("dmma+dfma", o);
  run<0>("dmma only", o, 2, 256); run<0>("dmma only", o, 1, 256); run<0>("dmma only", o, 1, 128); run<0>("dmma only", o, 2, 128);
  return 0;
}
