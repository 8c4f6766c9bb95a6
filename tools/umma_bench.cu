// umma_bench.cu -- cost model of tcgen05.mma kind::i8 (M=128, K=32) on B200: cycles per instruction as
// a function of N and of whether consecutive instructions accumulate into the same TMEM columns.
// One CTA per SM (all SMs busy, like the real kernel), operands = whatever is in shared memory
// (timing only).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_bench tools/umma_bench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void umma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_i8(int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

// mode 0: every MMA accumulates into the same columns; mode 1: alternate between two disjoint
// column ranges; mode 2: rotate over 512/N disjoint ranges; mode 3: same D, but A alternates
// between two smem tiles (operand reuse check)
// fpw > 0: warps 1..fpw run `fpiters` x 16 independent FP64 ops each (fpop 0: DFMA reg operands, 1: DADD with a
// constant operand, 2: FFMA fp32) while warp 0 issues the MMAs; their cycles go to out[148 + block]
__global__ void __launch_bounds__(288, 1) umma_bench_kernel(int N, int mode, int reps, long long* out, int fpw, int fpiters,
                                                            int fpop, double* sink) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_addr(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 2 * 16384, bar = sB + 32768, slot = bar + 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(raw + (slot - smem_addr(raw)));
  for (int i = threadIdx.x; i < (2 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_addr(raw)))[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x < 32) {
    const uint64_t ad = desc_sw128(sA), bd = desc_sw128(sB);
    const uint32_t id = idesc_i8(N);
    const int nreg = (mode == 1) ? 2 : (mode == 2 ? 512 / N : 1);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
      t0 = clock64();
      if (elect_one()) {
        int r = 0;
        for (int k = 0; k < reps; ++k) {
          const uint32_t d = tmem + (uint32_t)(r * N);
          const uint64_t a = ad + ((mode == 3 && (k & 1)) ? (16384 >> 4) : 0) + (uint64_t)(2 * (k & 3));
          umma_i8(d, a, bd + (uint64_t)(2 * (k & 3)), id, 1u);
          if (++r == nreg) r = 0;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      }
      __syncwarp();
      asm volatile(
          "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar),
          "r"((uint32_t)pass)
          : "memory");
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  else if ((int)(threadIdx.x >> 5) <= fpw) {
    double v[16];
    float f[16];
    for (int i = 0; i < 16; ++i) { v[i] = threadIdx.x + i; f[i] = threadIdx.x + i; }
    const double a = 1.0000001, b = 0.999;
    const long long t0 = clock64();
    for (int it = 0; it < fpiters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (fpop == 0) v[i] = fma(v[i], a, b);
        else if (fpop == 1) v[i] = v[i] - 4503601774854144.0;
        else f[i] = fmaf(f[i], 1.0000001f, 0.999f);
      }
    }
    const long long t1 = clock64();
    double s2 = 0;
    for (int i = 0; i < 16; ++i) s2 += v[i] + f[i];
    if (s2 == 1.2345) sink[0] = s2;
    if ((threadIdx.x & 31) == 0 && threadIdx.x == 32) out[148 + blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* d;
  double* sink;
  cudaMalloc(&d, 296 * sizeof(long long));
  cudaMalloc(&sink, 8);
  const size_t sh = 1024 + 2 * 16384 + 32768 + 64;
  cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
  const int reps = 512;
  printf("tcgen05.mma kind::i8 M=128 K=32, %d instructions per CTA, 148 CTAs; ideal = N/2 cycles per instruction\n", reps);
  const char* names[] = {"same D", "2 D ranges", "rotate D", "same D, 2 A tiles"};
  for (int N : {32, 64, 96, 128, 192, 256}) {
    for (int mode = 0; mode < 2; ++mode) {
      umma_bench_kernel<<<148, 288, sh>>>(N, mode, reps, d, 0, 0, 0, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("N %d mode %d: %s\n", N, mode, cudaGetErrorString(e));
        return 1;
      }
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      double s = 0;
      for (auto v : h) s += (double)v;
      s /= 148;
      printf("N %3d  %-18s %8.1f cycles/instr (ideal %5.1f)  -> %5.1f %% of peak\n", N, names[mode], s / reps, N / 2.0,
             100.0 * (N / 2.0) / (s / reps));
    }
  }
  // FP64 / FP32 work in other warps of the same SM while the tensor pipe runs N=128 MMAs back to back
  const char* ops[] = {"DFMA reg", "DADD const", "FFMA fp32"};
  for (int mma_reps : {0, 2048}) {
    for (int fpw : {4, 8}) {
      for (int fpop = 0; fpop < 3; ++fpop) {
        const int fpiters = 400;
        cudaMemset(d, 0, 296 * sizeof(long long));
        umma_bench_kernel<<<148, 288, sh>>>(128, 0, mma_reps, d, fpw, fpiters, fpop, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("fp test: %s\n", cudaGetErrorString(e));
          return 1;
        }
        std::vector<long long> h(296);
        cudaMemcpy(h.data(), d, 296 * sizeof(long long), cudaMemcpyDeviceToHost);
        double s = 0, f = 0;
        for (int i = 0; i < 148; ++i) { s += (double)h[i]; f += (double)h[148 + i]; }
        s /= 148; f /= 148;
        printf("MMA N=128 x %4d: %7.1f cyc/instr | %d warps of %s: %6.2f cycles per warp-instruction (%d x 16 per warp)\n", mma_reps,
               mma_reps ? s / mma_reps : 0.0, fpw, ops[fpop], f / (fpiters * 16.0), fpiters);
      }
    }
  }
  // sustained int8 tensor peak at whatever clock the power cap allows (wall clock, CUDA events)
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int big = 16384;
    umma_bench_kernel<<<148, 288, sh>>>(256, 0, big, d, 0, 0, 0, sink);
    cudaEventRecord(e0);
    umma_bench_kernel<<<148, 288, sh>>>(256, 0, big, d, 0, 0, 0, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = 0;
    for (auto v : h) cyc += (double)v;
    cyc /= 148;
    const double ops = 148.0 * 2.0 * big * 2.0 * 128.0 * 256.0 * 32.0;   // 2 passes inside the kernel
    printf("INT8_PEAK_TOPS %.1f  (tcgen05.mma kind::i8 M=128 N=256 K=32 back to back on 148 SMs, %.3f ms, SM clock %.0f MHz under load)\n",
           ops / (ms * 1e-3) / 1e12, ms, cyc / (ms * 0.5 * 1e-3) / 1e6);
  }
  cudaFree(d);
  return 0;
}
