// dmma_lds_bench.cu -- ceiling of an LDS-fed DMMA inner loop on B200: the krgemm2 warp tile
// (16 rows x 16 cols x 4 weight indices: 2 A + 8 B fragment loads per 16 DMMAs) with operands
// resident in shared memory and nothing else (no global traffic, no pipeline, no barriers).
// tools/dmma_bench.cu (same registers for every DMMA) reaches 37.1 TF/s; this is the number a
// real FP64 tensor-core contraction can approach.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
constexpr int ALD = 20, BNP = 20, MA = 128;
template <int NB>   // B loads per k4: 8 = as in krgemm2; 0 = B fragments kept in registers
__global__ void __launch_bounds__(512, 1) k(double* out, int iters) {
  extern __shared__ double sm[];
  double* Bs = sm;                         // [4][MA][BNP]
  double* As = sm + 4 * MA * BNP;          // [16 warps][16][ALD]
  for (int i = threadIdx.x; i < 4 * MA * BNP + 16 * 16 * ALD; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
  const double* Asb = As + wid * 16 * ALD + g * ALD + tq;
  double acc[4][2][2][2] = {};
  double bfr[4][2];
  for (int p = 0; p < 4; ++p) { bfr[p][0] = Bs[p * 7 + lane]; bfr[p][1] = Bs[p * 11 + lane]; }
  for (int it = 0; it < iters; ++it) {
    const double* Bsb = Bs + ((it & 7) * 16 + tq) * BNP + g;
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      double af[2], bf[4][2];
      af[0] = Asb[k4 * 4];
      af[1] = Asb[8 * ALD + k4 * 4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (NB) {
          const double* bp = Bsb + (p * MA + k4 * 4) * BNP;
          bf[p][0] = bp[0];
          bf[p][1] = bp[8];
        } else {
          bf[p][0] = bfr[p][0];
          bf[p][1] = bfr[p][1];
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) dmma(acc[p][mi][ni][0], acc[p][mi][ni][1], af[mi], bf[p][ni]);
    }
  }
  double s = 0;
  for (int p = 0; p < 4; ++p)
    for (int mi = 0; mi < 2; ++mi)
      for (int ni = 0; ni < 2; ++ni) s += acc[p][mi][ni][0] + acc[p][mi][ni][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NB>
void run(const char* name, double* o, int threads) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148, iters = 4000;
  const size_t sh = (4 * MA * BNP + 16 * 16 * ALD) * sizeof(double);
  cudaFuncSetAttribute(k<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
  k<NB><<<blocks, threads, sh>>>(o, 10); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NB><<<blocks, threads, sh>>>(o, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = (double)blocks * (threads / 32) * iters * 64.0 * 512.0;
  printf("%-34s %2d warps/SM: %.3f ms  %.2f TF/s  (%s)\n", name, threads / 32, ms, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double* o; cudaMalloc(&o, 148 * 512 * 8);
  run<8>("A+B fragments from LDS", o, 512); run<8>("A+B fragments from LDS", o, 256); run<8>("A+B fragments from LDS", o, 128);
  run<0>("A from LDS, B in registers", o, 512); run<0>("A from LDS, B in registers", o, 256);
  return 0;
}
