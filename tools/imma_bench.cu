// imma_bench.cu -- throughput of the legacy tensor path for INT8 (mma.sync m16n8k32 s8*s8+s32) and
// FP16 (m16n8k16 f16*f16+f32) on B200, next to DMMA (tools/dmma_bench.cu: 37 TF/s).  Decides whether an
// Ozaki-style error-free splitting of the FP64 contractions onto integer MMAs is worth building on
// mma.sync before a tcgen05 (kind::i8) version exists.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imma_bench tools/imma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void imma_kernel(int* out, int iters) {
  int acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  unsigned a0 = threadIdx.x * 0x01010101u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 ^ 0x55u, b1 = b0 + 7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(acc[i][0]), "+r"(acc[i][1]), "+r"(acc[i][2]), "+r"(acc[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void hmma_kernel(float* out, int iters) {
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  unsigned a0 = 0x3c003c00u, a1 = a0, a2 = a0, a3 = a0, b0 = 0x38003800u, b1 = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 20000;
  int* d;
  cudaMalloc(&d, blocks * threads * sizeof(int));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    imma_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 2.0 * 16 * 8 * 32 * 8.0 * iters * (threads / 32) * blocks;
    printf("IMMA m16n8k32 s8: %.1f TOPS (%.2f ms)\n", ops / ms / 1e9, ms);
  }
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    hmma_kernel<<<blocks, threads>>>((float*)d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 2.0 * 16 * 8 * 16 * 8.0 * iters * (threads / 32) * blocks;
    printf("HMMA m16n8k16 f16->f32: %.1f TFLOPS (%.2f ms)\n", ops / ms / 1e9, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
