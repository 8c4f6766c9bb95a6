// tnml_ozaki.cu -- the Khatri-Rao projection GEMM of the fixedL bond update on the 5th-generation
// tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, operands staged by TMA), at float64
// accuracy through an error-free splitting ("Ozaki scheme").
//
//   Out[row][j] = sum_p w_p(row) * ( sum_a In[row][a] * Bm[(a*S+p)*ldb + j] )        (= tnml::krgemm)
//
// replaces the first half of P = B * t.v (fixedL.cc:318,377,399,416).  tcgen05 has no f64 kind, and
// the reference's CG does not tolerate fp32-level noise, so both operands are cut into
// signed 7-bit slices,
//   In[row][a]  = ea[row] * sum_i 2^(-7(i+1)) A_i[row][a],   |A_i| <= 64   (one power of two per row)
//   Bm[..][c]   = eb[c]   * sum_j 2^(-7(j+1)) B_j[c][a],     |B_j| <= 64   (one power of two per column)
// the slice products A_i * B_j^T with i + j < NS are EXACT in int32 (K <= 128: |sum| <= 128*64*64*8),
// and only the final combination sum_L 2^(-7(L+2)) acc_L rounds, in float64, like an FMA chain would.
// With NS = 8 the dropped tail is 2^-57 of (row max) x (column max): float64 GEMM accuracy.
//
// Mapping (one persistent CTA per SM, 19 warps, warp-specialised; see the comments at each role):
//   * tile = 128 rows x 32 columns (32/S output columns x S weight indices) x NS accumulator levels;
//     level L of a tile lives in TMEM columns [32L, 32L+32): the MMA of A-slice i multiplies it with
//     the B slices 0..NS-1-i at once -- they are contiguous in shared memory, so this is ONE
//     tcgen05.mma with N = 32(NS-i) whose 32-column groups land on levels i..NS-1.  36 slice products
//     = 8 instructions per 32-deep k-step.
//   * warp 0: TMA producer.  A (all NS slices of 128 rows, 16 KB each, 128B-swizzled K-major) stays
//     resident in shared memory for every column tile of the row tile and is refilled slice by slice
//     (one mbarrier per slice) as soon as the last column tile's MMAs of that slice have retired;
//     B column tiles (NS x 32 rows x 128 B = 32 KB) stream through a 2-stage ring from L2.
//   * warps 1-2: MMA issuers (even / odd A slices; accumulators are handed over zeroed, so the
//     order of the two warps' instructions in the tensor pipe does not matter).
//   * warps 3-18: epilogue (tcgen05.ld of all levels, exact integer merge, float64 combination with
//     the column scale and the output-side Khatri-Rao weights, store), overlapping the next tile's
//     MMAs through two TMEM accumulator buffers (2 x 256 columns).
// Measured cost model (tools/umma_bench.cu, B200): a kind::i8 M=128 K=32 instruction takes
// max(N/2, 32 + N/4) cycles (shared-memory operand bandwidth below N = 128); FP64 instructions of
// other warps are throttled ~20x while tcgen05.mma is in flight (47 vs 2.3 cycles per warp
// instruction; FP32 and integer are not) -- hence the integer-heavy epilogue.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "tnml_kernels.cuh"

namespace tnml {

namespace {

constexpr int OZ_KB = 128;                    // bytes (= int8 elements) per operand row: one 128B swizzle span
constexpr int OZ_TM = 128;                    // rows per tile (UMMA M)
constexpr int OZ_TN = 32;                     // columns per accumulator level per tile
constexpr int OZ_ASLICE = OZ_TM * OZ_KB;      // 16 KB
constexpr int OZ_BSLICE = OZ_TN * OZ_KB;      // 4 KB
constexpr int OZ_THREADS = 608;               // warp 0: TMA producer, warps 1-2: MMA issuers, warps 3-18: epilogue

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "OZ_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra OZ_DONE;\n\t"
      "bra OZ_WAIT;\n\t"
      "OZ_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// one lane of a converged warp (the compiler keeps the operands of single-thread tcgen05/TMA
// instructions in uniform registers when the issuing branch is guarded by elect.sync)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, issued by one thread for the CTA
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they have retired
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, 128B swizzle (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;    // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// instruction descriptor: S32 accumulate, signed 8-bit A and B, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_i8(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OZ_TM >> 4) << 24);
}
// exact int64 -> double for |x| < 2^51: the bits of 1.5 * 2^52 + x, minus the bias
__device__ __forceinline__ double ll2d(long long x) {
  return __longlong_as_double(x + 0x4338000000000000LL) - 6755399441055744.0;
}
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// Khatri-Rao weights of one row: S=2 -> (f1_0, f1_1); S=4 -> f1_s * f2_q at p = s*2+q
template <int S>
__device__ __forceinline__ void kr_weights_oz(const double* __restrict__ f1, const double* __restrict__ f2, long img,
                                              double (&w)[S]) {
  const double a0 = f1[img * 2], a1 = f1[img * 2 + 1];
  if (S == 2) {
    w[0] = a0;
    w[1] = a1;
  } else {
    const double b0 = f2[img * 2], b1 = f2[img * 2 + 1];
    w[0] = a0 * b0;
    w[1] = a0 * b1;
    w[S - 2] = a1 * b0;
    w[S - 1] = a1 * b1;
  }
}

// next signed 7-bit digit of x (|x| <= 0.5): q = rint(128 x) as a two's-complement byte, x <- 128 x - q.
// Round-to-nearest-even through the 1.5 * 2^52 shift (128 x is exact, so the fused add rounds exactly like
// rint, and the digit sits in the low word of the sum): three FP64 pipe instructions instead of a
// FRND + F2I pair on the quarter-rate conversion unit, which was what bounded the slicing kernels.
__device__ __forceinline__ uint32_t oz_digit(double& x) {
  constexpr double SHIFT = 6755399441055744.0;   // 1.5 * 2^52
  const double t = fma(x, 128.0, SHIFT);
  x = fma(x, 128.0, SHIFT - t);                  // SHIFT - t = -q exactly
  return (uint32_t)__double2loint(t) & 0xFFu;
}

// ---- slicing ------------------------------------------------------------------------------------
// one warp per row of In [rows][ldin] (ma <= 128): ea[row] and NS int8 planes, written as
// A8[rowtile][slice][128 rows][128 bytes]; rows beyond `rows` and columns beyond ma are zero.
// A warp takes OZ_RPW consecutive rows per iteration and issues all their loads (two 16-byte loads per
// lane and row when the rows are 16-byte aligned) before the first row is cut: the kernel streams
// 8 ma bytes in and 8 x 128 bytes out per row and needs the loads in flight to approach HBM speed.
constexpr int OZ_RPW = 4;
__global__ void __launch_bounds__(256) oz_slice_rows_kernel(const double* __restrict__ In, long ldin, int ma, long rows,
                                                            long rows_pad, int ns, int8_t* __restrict__ A8,
                                                            double* __restrict__ ea) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarp = ((long)gridDim.x * blockDim.x) >> 5;
  const bool vec = ((reinterpret_cast<uintptr_t>(In) & 15) == 0) && ((ldin & 1) == 0);
  const int a0 = lane * 4;
  for (long r0 = warp * OZ_RPW; r0 < rows_pad; r0 += nwarp * OZ_RPW) {
    double x[OZ_RPW][4];
#pragma unroll
    for (int u = 0; u < OZ_RPW; ++u) {
      const long r = r0 + u;
      const double* src = In + r * ldin + a0;
      if (r < rows && vec && a0 + 3 < ma) {
        const double2 v0 = __ldg(reinterpret_cast<const double2*>(src));
        const double2 v1 = __ldg(reinterpret_cast<const double2*>(src) + 1);
        x[u][0] = v0.x;
        x[u][1] = v0.y;
        x[u][2] = v1.x;
        x[u][3] = v1.y;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) x[u][k] = (r < rows && a0 + k < ma) ? __ldg(src + k) : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < OZ_RPW; ++u) {
      const long r = r0 + u;   // < rows_pad: rows_pad is a multiple of 128, hence of OZ_RPW
      double mx = fmax(fmax(fabs(x[u][0]), fabs(x[u][1])), fmax(fabs(x[u][2]), fabs(x[u][3])));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      int e = 0;
      if (mx > 0.0) {
        frexp(mx, &e);   // mx = f * 2^e, f in [0.5, 1)
        e += 1;          // |x| * 2^-e <= 0.5
      }
      const double sc = pow2(-e);
      if (lane == 0) ea[r] = pow2(e);
      const long rt = r >> 7, rl = r & 127;
      uint32_t* dst = reinterpret_cast<uint32_t*>(A8 + ((rt * ns) * OZ_TM + rl) * OZ_KB) + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k) x[u][k] *= sc;
      for (int i = 0; i < ns; ++i) {
        uint32_t pk = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) pk |= (uint32_t)oz_digit(x[u][k]) << (8 * k);
        dst[(long)i * (OZ_ASLICE / 4)] = pk;
      }
    }
  }
}

// one warp per column c = (coltile, jl, p) of the bond-tensor operand: B8[coltile][slice][32][128], eb[coltile*32 + jl*S + p]
template <int S>
__global__ void __launch_bounds__(256) oz_slice_cols_kernel(const double* __restrict__ Bm, long ldb, int ma, int J, int coltiles,
                                                            int ns, int8_t* __restrict__ B8, double* __restrict__ eb) {
  constexpr int JT = OZ_TN / S;
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long ncol = (long)coltiles * OZ_TN;
  if (warp >= ncol) return;
  const int ct = (int)(warp / OZ_TN), c = (int)(warp % OZ_TN);
  const int jl = c / S, p = c % S;   // column order inside a tile: output column major, weight index fastest
  const int j = ct * JT + jl;
  double x[4];
  double mx = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int a = lane * 4 + k;
    x[k] = (j < J && a < ma) ? Bm[((long)a * S + p) * ldb + j] : 0.0;
    mx = fmax(mx, fabs(x[k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.0) {
    frexp(mx, &e);
    e += 1;
  }
  const double sc = pow2(-e);
  if (lane == 0) eb[warp] = pow2(e);
  uint32_t* dst = reinterpret_cast<uint32_t*>(B8 + (((long)ct * ns) * OZ_TN + c) * OZ_KB) + lane;
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] *= sc;
  for (int i = 0; i < ns; ++i) {
    uint32_t pk = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) pk |= (uint32_t)oz_digit(x[k]) << (8 * k);
    dst[(long)i * (OZ_BSLICE / 4)] = pk;
  }
}

// ---- the GEMM -----------------------------------------------------------------------------------
template <int NS, int S>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const double* __restrict__ ea, const double* __restrict__ eb, const double* __restrict__ f1,
               const double* __restrict__ f2, int div, double* __restrict__ Out, long ldout, long rows, int J,
               int coltiles, long ntiles, long long* __restrict__ dbg) {
  constexpr int JT = OZ_TN / S;
  constexpr uint32_t BSTAGE = NS * OZ_BSLICE;
#ifdef OZ_PROFILE
  const long long k_t0 = clock64();
#define OZ_T(var) const long long var = clock64()
#define OZ_ACC(slot, a, b) prof[slot] += (b) - (a)
#else
#define OZ_T(var)
#define OZ_ACC(slot, a, b)
#endif
  extern __shared__ uint8_t oz_smem_raw[];
  const uint32_t base = (smem_addr(oz_smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                               // [NS][128][128]
  const uint32_t sB = sA + NS * OZ_ASLICE;                // [2][NS][32][128]
  const uint32_t sBar = sB + 2 * BSTAGE;
  const uint32_t a_full = sBar, a_empty = sBar + 8 * NS;
  const uint32_t b_full = sBar + 16 * NS, b_empty = b_full + 16;
  const uint32_t acc_full = b_empty + 16, acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(oz_smem_raw + (tmem_slot - smem_addr(oz_smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous tile range of this CTA; tile t -> (row tile t / coltiles, column tile t % coltiles)
  const long t0 = (ntiles * blockIdx.x) / gridDim.x, t1 = (ntiles * (blockIdx.x + 1)) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(a_full + 8 * i, 1);
      mbar_init(a_empty + 8 * i, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 2);      // one tcgen05.commit per issuer warp
      mbar_init(acc_full + 8 * s, 2);
      mbar_init(acc_empty + 8 * s, 512);  // every epilogue thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: all 512 columns (two accumulator buffers of NS*32 <= 256 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  // first tile of the chunk; (rt, ct) then advance without divisions
  const long rt0 = t0 / coltiles;
  const int ct0 = (int)(t0 - rt0 * coltiles);
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t bphase = 0, aphase = 0;
      long rt = rt0;
      int ct = ct0;
      bool new_rt = true;
      for (long t = t0; t < t1; ++t) {
        mbar_wait(b_empty + 8 * stage, bphase ^ 1);
        mbar_expect_tx(b_full + 8 * stage, BSTAGE);
        tma_load_2d(sB + stage * BSTAGE, &tmB, b_full + 8 * stage, 0, ct * NS * OZ_TN);
        if (new_rt) {
          for (int i = 0; i < NS; ++i) {
            mbar_wait(a_empty + 8 * i, aphase ^ 1);
            mbar_expect_tx(a_full + 8 * i, OZ_ASLICE);
            tma_load_2d(sA + i * OZ_ASLICE, &tmA, a_full + 8 * i, 0, (int)((rt * NS + i) * OZ_TM));
          }
          aphase ^= 1;
        }
        new_rt = false;
        if (++ct == coltiles) {
          ct = 0;
          ++rt;
          new_rt = true;
        }
        if (++stage == 2) {
          stage = 0;
          bphase ^= 1;
        }
      }
    }
  } else if (warp <= 2) {
    // ===== MMA issuers: warp 1 issues the even A slices, warp 2 the odd ones =====
    // Every accumulator buffer is handed over ZEROED by the epilogue, so every MMA accumulates and
    // the order in which the two warps' instructions reach the tensor pipe does not matter (integer
    // adds).  Why two warps: one thread needs ~13 instructions (R2UR of the descriptors, ...) per
    // tcgen05.mma and our MMAs are short (N = 32..256: 40..128 cycles); with a single issuer the
    // issue thread, not the tensor pipe, set the pace (tools/umma_bench: the pipe itself runs at
    // max(N/2, 32 + N/4) cycles per instruction, 99 % of peak from N = 128 on).
    // The whole warp runs the loop converged; one elected lane issues (uniform-register operands).
    const int par = warp - 1;
    int stage = 0;   // B stage == accumulator buffer: both rings advance once per tile
    uint32_t ph = 0, aphase = 0;
    int ct = ct0;
    bool new_rt = true;
#ifdef OZ_PROFILE
    long long prof[4] = {0, 0, 0, 0};
#endif
    const uint64_t adesc0 = umma_desc_sw128(sA);
    const uint64_t bdesc0 = umma_desc_sw128(sB);
    for (long t = t0; t < t1; ++t) {
      const bool last_of_rt = (t + 1 == t1) || (ct + 1 == coltiles);
      OZ_T(c0);
      mbar_wait(b_full + 8 * stage, ph);
      OZ_T(c1);
      mbar_wait(acc_empty + 8 * stage, ph);   // completed once by the epilogue's initial hand-over
      OZ_T(c2);
      OZ_ACC(0, c0, c1);
      OZ_ACC(1, c1, c2);
      tc_fence_after();
      const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (BSTAGE >> 4));
      const uint32_t dbase = tmem + (uint32_t)(stage * 256);
      if (!new_rt && !last_of_rt) {
        if (elect_one()) {
#pragma unroll
          for (int i2 = 0; i2 < (NS + 1) / 2; ++i2) {
            const int i = 2 * i2 + par;
            if (i < NS) {
              const uint64_t adesc = adesc0 + (uint64_t)(i * (OZ_ASLICE >> 4));
              const uint32_t idesc = umma_idesc_i8((NS - i) * OZ_TN);
#pragma unroll
              for (int kb = 0; kb < OZ_KB / 32; ++kb)   // 32 int8 per k-step: +32 B inside the swizzle span
                umma_i8(dbase + i * OZ_TN, adesc + (uint64_t)(2 * kb), bdesc + (uint64_t)(2 * kb), idesc, 1u);
            }
          }
          umma_commit(b_empty + 8 * stage);
          umma_commit(acc_full + 8 * stage);
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int i2 = 0; i2 < (NS + 1) / 2; ++i2) {
          const int i = 2 * i2 + par;
          if (i < NS) {
            if (new_rt) {
              OZ_T(c3);
              mbar_wait(a_full + 8 * i, aphase);
              OZ_T(c4);
              OZ_ACC(2, c3, c4);
              tc_fence_after();
            }
            if (elect_one()) {
              const uint64_t adesc = adesc0 + (uint64_t)(i * (OZ_ASLICE >> 4));
              const uint32_t idesc = umma_idesc_i8((NS - i) * OZ_TN);
#pragma unroll
              for (int kb = 0; kb < OZ_KB / 32; ++kb)
                umma_i8(dbase + i * OZ_TN, adesc + (uint64_t)(2 * kb), bdesc + (uint64_t)(2 * kb), idesc, 1u);
              if (last_of_rt) umma_commit(a_empty + 8 * i);   // slice i may be refilled for the next row tile
            }
            __syncwarp();
          }
        }
        if (elect_one()) {
          umma_commit(b_empty + 8 * stage);
          umma_commit(acc_full + 8 * stage);
        }
        __syncwarp();
        if (new_rt) aphase ^= 1;
      }
      new_rt = false;
      if (++ct == coltiles) {
        ct = 0;
        new_rt = true;
      }
      if (++stage == 2) {
        stage = 0;
        ph ^= 1;
      }
    }
#ifdef OZ_PROFILE
    if (dbg && warp == 1 && lane == 0) {
      dbg[blockIdx.x * 8 + 0] = prof[0];
      dbg[blockIdx.x * 8 + 1] = prof[1];
      dbg[blockIdx.x * 8 + 2] = prof[2];
      dbg[blockIdx.x * 8 + 3] = clock64() - k_t0;
    }
#endif
  } else {
    // ===== epilogue: warps 3..18; warp w owns TMEM lanes 32*(w%4) .. +31 (hardware rule) and the
    // column quarter (w-3)/4 of every level: 8 columns = JT/4 output columns x S weight indices.
    //
    // Measured on B200 (tools/umma_bench): while tcgen05.mma instructions are in flight an FP64
    // instruction costs a warp ~47 cycles (2.3 when the tensor pipe is idle; FP32/integer are not
    // affected), whatever the number of independent operations in flight.  So the epilogue keeps
    // its FP64 instruction count per warp low: (a) the NS int32 levels of a column are merged
    // EXACTLY in integer arithmetic into two int64 (levels 0-3 and 4-7), (b) those are converted
    // with the 2^52 trick, (c) one FMA pair applies the column scale, one FMA the row weight --
    // 5 FP64 instructions per column, 8 columns per warp, 16 epilogue warps.
    // All level loads are in flight before the single wait; the columns are zeroed right after
    // they have been read and the buffer goes back to the MMA warps before the arithmetic.
    constexpr int HC = OZ_TN / 4;   // columns per warp
    constexpr int HJ = JT / 4;      // output columns per warp
    const int q = warp & 3;
    const int quarter = (warp - 3) >> 2;
    const int rl = q * 32 + lane;   // row inside the tile = TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(quarter * HC);
    // initial hand-over: both accumulator buffers zeroed
#pragma unroll
    for (int b2 = 0; b2 < 2; ++b2) {
#pragma unroll
      for (int L = 0; L < NS; ++L) tmem_st8_zero(tlane + b2 * 256 + L * OZ_TN);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(acc_empty);
    mbar_arrive(acc_empty + 8);
    int buf = 0;
    uint32_t accphase = 0;
#ifdef OZ_PROFILE
    long long prof[4] = {0, 0, 0, 0};
#endif
    long rt = rt0;
    int ct = ct0;
    bool new_rt = true;
    long row = 0;
    bool rok = false;
    double w[S];
    const bool vec2 = ((ldout & 1) == 0) && ((reinterpret_cast<size_t>(Out) & 15) == 0) && (HJ % 2 == 0);
    for (long t = t0; t < t1; ++t) {
      if (new_rt) {   // the row (weights, scale) changes once per row tile only
        row = rt * OZ_TM + rl;
        rok = row < rows;
#pragma unroll
        for (int p = 0; p < S; ++p) w[p] = 0.0;
        if (rok) {
          kr_weights_oz<S>(f1, f2, row / div, w);
          const double earow = ea[row];
#pragma unroll
          for (int p = 0; p < S; ++p) w[p] *= earow;   // exact: earow is a power of two
        }
      }
      // column scales (powers of two): lane c < 8 holds the exponent word of eb[c] of this warp's
      // quarter, broadcast by shuffle below; 2^-35 / 2^-63 are folded in by integer exponent arithmetic
      const int ebw = __double2hiint(__ldg(eb + (long)ct * OZ_TN + quarter * HC + (lane & (HC - 1))));
      OZ_T(e0);
      mbar_wait(acc_full + 8 * buf, accphase);
      OZ_T(e1);
      OZ_ACC(0, e0, e1);
      tc_fence_after();
      const uint32_t taddr = tlane + (uint32_t)(buf * 256);
      uint32_t acc[8][HC];
#pragma unroll
      for (int L = 0; L < NS; ++L) tmem_ld8(taddr + L * OZ_TN, acc[L]);
      tmem_ld_wait();
#pragma unroll
      for (int L = 0; L < NS; ++L) tmem_st8_zero(taddr + L * OZ_TN);
      tmem_st_wait();
      tc_fence_before();
#ifndef OZ_LATE_RELEASE
      mbar_arrive(acc_empty + 8 * buf);   // the (zeroed) accumulator buffer goes back to the MMA warps
#endif
      OZ_T(e2);
      OZ_ACC(1, e1, e2);
#pragma unroll
      for (int L = NS; L < 8; ++L)
#pragma unroll
        for (int c = 0; c < HC; ++c) acc[L][c] = 0u;
      double o[HJ];
#pragma unroll
      for (int j = 0; j < HJ; ++j) o[j] = 0.0;
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        // sum_L acc_L 2^(-7(L+2)) = hi * 2^-35 + lo * 2^-63, hi/lo exact (|acc_L| <= (L+1) 2^19)
        const int p01 = (int)acc[0][c] * 128 + (int)acc[1][c], p23 = (int)acc[2][c] * 128 + (int)acc[3][c];
        const int p45 = (int)acc[4][c] * 128 + (int)acc[5][c], p67 = (int)acc[6][c] * 128 + (int)acc[7][c];
        const long long hi = (long long)p01 * 16384 + p23, lo = (long long)p45 * 16384 + p67;
        const int e = __shfl_sync(0xffffffffu, ebw, c);
        const double k1 = __hiloint2double(e - (35 << 20), 0), k2 = __hiloint2double(e - (63 << 20), 0);
#ifdef OZ_NO_MATH   // experiment: no FP64 in the epilogue (wrong results, timing only)
        o[c / S] = __longlong_as_double((hi ^ lo) + e);
#else
        const double v = fma(ll2d(lo), k2, ll2d(hi) * k1);
        o[c / S] = fma(w[c % S], v, o[c / S]);
#endif
      }
#ifdef OZ_LATE_RELEASE
      // hand the buffer back only after the FP64 part: the tile after next must not start (and
      // throttle the FP64 pipe again) before this tile's arithmetic is done
      if (o[0] == 1.2345e300) w[0] = 0.0;   // keeps the arithmetic above the release
      mbar_arrive(acc_empty + 8 * buf);
#endif
      if (rok) {
        const int jb = ct * JT + quarter * HJ;
        double* orow = Out + row * ldout + jb;
        const int jn = (J - jb < HJ) ? (J - jb) : HJ;
        if (vec2 && jn == HJ) {
#pragma unroll
          for (int j = 0; j + 1 < HJ; j += 2) *reinterpret_cast<double2*>(orow + j) = make_double2(o[j], o[j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < HJ; ++j)
            if (j < jn) orow[j] = o[j];
        }
      }
      OZ_T(e3);
      OZ_ACC(2, e2, e3);
      new_rt = false;
      if (++ct == coltiles) {
        ct = 0;
        ++rt;
        new_rt = true;
      }
      if (++buf == 2) {
        buf = 0;
        accphase ^= 1;
      }
    }
#ifdef OZ_PROFILE
    if (dbg && threadIdx.x == 96) {   // first epilogue warp
      dbg[blockIdx.x * 8 + 4] = prof[0];
      dbg[blockIdx.x * 8 + 5] = prof[1];
      dbg[blockIdx.x * 8 + 6] = prof[2];
      dbg[blockIdx.x * 8 + 7] = t1 - t0;
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

long long* g_oz_dbg = nullptr;   // -DOZ_PROFILE builds: [CTA][8] cycle counters (tools/oz_test)

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tmap_encode_fn get_tmap_encode() {
  static tmap_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (tmap_encode_fn)p;
  }
  return fn;
}
// 2-D byte tensor [nrows][128], box [boxrows][128], 128B swizzle
bool make_map(CUtensorMap* m, const void* ptr, uint64_t nrows, uint32_t boxrows) {
  tmap_encode_fn enc = get_tmap_encode();
  if (!enc) return false;
  const cuuint64_t dims[2] = {OZ_KB, nrows};
  const cuuint64_t strides[1] = {OZ_KB};
  const cuuint32_t box[2] = {OZ_KB, boxrows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NS, int S>
bool oz_launch(cudaStream_t st, const int8_t* A8, const double* ea, long rows, const double* f1, const double* f2,
               int div, const int8_t* B8, const double* eb, int J, double* Out, long ldout, int num_sm) {
  constexpr int JT = OZ_TN / S;
  const long rowtiles = (rows + OZ_TM - 1) / OZ_TM;
  const int coltiles = (J + JT - 1) / JT;
  const long ntiles = rowtiles * coltiles;
  CUtensorMap tmA, tmB;
  if (!make_map(&tmA, A8, (uint64_t)rowtiles * NS * OZ_TM, OZ_TM)) return false;
  if (!make_map(&tmB, B8, (uint64_t)coltiles * NS * OZ_TN, NS * OZ_TN)) return false;
  const size_t sh = 1024 + (size_t)NS * OZ_ASLICE + 2 * (size_t)NS * OZ_BSLICE + 16 * NS + 64 + 16;
  static unsigned long long attr = 0;
  if (first_on_device(attr))
    cudaFuncSetAttribute(oz_gemm_kernel<NS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
  const int grid = (int)((ntiles < num_sm) ? ntiles : num_sm);
  oz_gemm_kernel<NS, S><<<grid, OZ_THREADS, sh, st>>>(tmA, tmB, ea, eb, f1, f2, div, Out, ldout, rows, J, coltiles, ntiles,
                                                      g_oz_dbg);
  return true;
}

}  // namespace

// ---- host interface -------------------------------------------------------------------------------
void oz_set_debug_buffer(long long* p) { g_oz_dbg = p; }
size_t oz_a8_bytes(long rows, int ns) { return (size_t)((rows + OZ_TM - 1) / OZ_TM) * ns * OZ_ASLICE; }
long oz_rows_pad(long rows) { return ((rows + OZ_TM - 1) / OZ_TM) * OZ_TM; }
size_t oz_b8_bytes(int S, int J, int ns) {
  const int JT = OZ_TN / S;
  return (size_t)((J + JT - 1) / JT) * ns * OZ_BSLICE;
}
long oz_cols_pad(int S, int J) {
  const int JT = OZ_TN / S;
  return (long)((J + JT - 1) / JT) * OZ_TN;
}
bool oz_supported(int S, int ma, int ns) { return (S == 2 || S == 4) && ma >= 1 && ma <= OZ_KB && ns >= 6 && ns <= 8; }

void oz_slice_rows(cudaStream_t st, const double* In, long ldin, int ma, long rows, int ns, int8_t* A8, double* ea) {
  const long rp = oz_rows_pad(rows);
  long blocks = (rp / OZ_RPW + 7) / 8;   // 8 warps per block, OZ_RPW rows per warp and iteration
  if (blocks > 148 * 16) blocks = 148 * 16;
  oz_slice_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(In, ldin, ma, rows, rp, ns, A8, ea);
}

void oz_slice_cols(cudaStream_t st, int S, const double* Bm, long ldb, int ma, int J, int ns, int8_t* B8, double* eb) {
  const int JT = OZ_TN / S;
  const int coltiles = (J + JT - 1) / JT;
  const long ncol = (long)coltiles * OZ_TN;
  const unsigned blocks = (unsigned)((ncol + 7) / 8);
  if (S == 2)
    oz_slice_cols_kernel<2><<<blocks, 256, 0, st>>>(Bm, ldb, ma, J, coltiles, ns, B8, eb);
  else
    oz_slice_cols_kernel<4><<<blocks, 256, 0, st>>>(Bm, ldb, ma, J, coltiles, ns, B8, eb);
}

bool oz_krgemm(cudaStream_t st, int S, int ns, const int8_t* A8, const double* ea, long rows, const double* f1,
               const double* f2, int div, const int8_t* B8, const double* eb, int J, double* Out, long ldout,
               int num_sm) {
  if (rows <= 0 || J <= 0) return true;
#define OZ_CASE(NSV)                                                                                          \
  case NSV:                                                                                                   \
    return (S == 2) ? oz_launch<NSV, 2>(st, A8, ea, rows, f1, f2, div, B8, eb, J, Out, ldout, num_sm)        \
                    : oz_launch<NSV, 4>(st, A8, ea, rows, f1, f2, div, B8, eb, J, Out, ldout, num_sm);
  switch (ns) {
    OZ_CASE(6)
    OZ_CASE(7)
    OZ_CASE(8)
    default:
      return false;
  }
#undef OZ_CASE
}

}  // namespace tnml
