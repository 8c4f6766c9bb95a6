// tnml_kernels.cu -- float64 contraction kernels of the fixedL bond update for
// sm_100a.  See tnml_kernels.cuh for what each replaces in the reference.
//
// Shared inner product core: a 128 x 64 (x NB) output tile per CTA, 256
// threads, 8 x 4 (x NB) register tile per thread, BK = 16, double-buffered
// shared memory with register prefetch.  Both operands are generated while
// they are staged (Khatri-Rao factors l_n (x) phi_n), so the dense projected
// input t.v of fixedL.cc:183-185 never exists.
#include "tnml_kernels.cuh"

#include <cstdio>
#include <cstdlib>

namespace tnml {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int BK = 16;
constexpr int NTHR = 256;

// FP64 tensor-core inner product: mma.sync.aligned.m8n8k4 (SASS DMMA.8x8x4).  On B200 DMMA
// and DFMA share one FP64 pipe (37.1 vs 36.3 TF/s measured, 35.7 together), but one DMMA
// retires 256 FMAs per issue slot, so the pipe can be kept busy with far fewer instructions
// and shared-memory reads than the 8x4 register-tiled DFMA loop it replaces (measured 16 TF/s).
// Fragments (PTX ISA, .f64 m8n8k4): g = lane>>2, t = lane&3;  A[g][t], B[t][g], C[g][2t..2t+1].
// CTA tile 128 x 64, 8 warps stacked along M, warp tile 16 x 64 = 2 x 8 mma tiles (16-row
// granularity lets the host pick the rows per tile that balances the waves).
constexpr int LDA = BM + 4;   // k-row strides = 8 banks mod 32: conflict-free fragment loads
constexpr int LDB = BN + 4;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void tile_mma(const double* __restrict__ A, const double* __restrict__ B,
                                         double (&acc)[2][8][2], int g, int t) {
  // A -> As[buf] + wm*16 ; B -> Bs[buf].  Warp tile 16 (M) x 64 (N) = 2 x 8 mma tiles.
#pragma unroll
  for (int k4 = 0; k4 < BK / 4; ++k4) {
    double af[2], bf[8];
    const double* ap = A + (k4 * 4 + t) * LDA + g;
    const double* bp = B + (k4 * 4 + t) * LDB + g;
    af[0] = ap[0];
    af[1] = ap[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bf[i] = bp[i * 8];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
  }
}

// Khatri-Rao weights of one row: S=2 -> (f1_0, f1_1); S=4 -> f1_s * f2_q at p = s*2+q
template <int S>
__device__ __forceinline__ void kr_weights(const double* __restrict__ f1, const double* __restrict__ f2, long img,
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

// ---------------------------------------------------------------------------
// Out[row][j] = sum_{a,p} In[row][a] * w_p(row) * Bm[(a*S+p)*ldb + j].  K = S*ma.
// 128(bm) x 64 tile per CTA, 8x4 per thread, two CTAs per SM; `bm` (rows per tile, <= 128)
// is chosen by the host so that the tile count is a whole number of waves.
template <int S>
__global__ void __launch_bounds__(NTHR, 2)
krgemm_kernel(const double* __restrict__ In, long ldin, int ma, const double* __restrict__ f1,
              const double* __restrict__ f2, int div, const double* __restrict__ Bm, long ldb, int J,
              double* __restrict__ Out, long ldout, long rows, int bm) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                  // [2][BK][LDA]
  double* Bs = smem + 2 * BK * LDA;   // [2][BK][LDB]
  constexpr int AK = BK / S;         // a-values per k-tile
  constexpr int APT = AK / 2;        // a-values per loader thread (2 threads per row)
  const int t = threadIdx.x;
  const int lane = t & 31, wm = t >> 5, g = lane >> 2, tq = lane & 3;
  const long row0 = (long)blockIdx.x * bm;
  const int j0 = blockIdx.y * BN;

  const int lrow = t >> 1, lah = (t & 1) * APT;
  const long grow = row0 + lrow;
  const bool rok = (grow < rows) && (lrow < bm);
  double w[S];
#pragma unroll
  for (int p = 0; p < S; ++p) w[p] = 0.0;
  if (rok) kr_weights<S>(f1, f2, grow / div, w);
  const int bk = t >> 4, bc = (t & 15) * 4;

  double acc[2][8][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double ra[APT];
  double rb[4];
  const int nk = (ma + AK - 1) / AK;
  const int K2 = ma * S;

  auto gload = [&](int kt) {
    const int a0 = kt * AK + lah;
#pragma unroll
    for (int i = 0; i < APT; ++i) {
      int a = a0 + i;
      ra[i] = (rok && a < ma) ? In[grow * ldin + a] : 0.0;
    }
    const int k2 = kt * BK + bk;
    const bool kok = k2 < K2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int j = j0 + bc + i;
      rb[i] = (kok && j < J) ? Bm[(long)k2 * ldb + j] : 0.0;
    }
  };
  auto sstore = [&](int buf) {
    double* A = As + buf * BK * LDA;
#pragma unroll
    for (int i = 0; i < APT; ++i)
#pragma unroll
      for (int p = 0; p < S; ++p) A[((lah + i) * S + p) * LDA + lrow] = ra[i] * w[p];
    double2* B = reinterpret_cast<double2*>(Bs + (buf * BK + bk) * LDB + bc);
    B[0] = make_double2(rb[0], rb[1]);
    B[1] = make_double2(rb[2], rb[3]);
  };

  gload(0);
  sstore(0);
  __syncthreads();
  const bool active = (wm * 16 < bm);
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
    if (active) tile_mma(As + buf * BK * LDA + wm * 16, Bs + buf * BK * LDB, acc, g, tq);
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }
  if (!active) return;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int lr = wm * 16 + mi * 8 + g;
    const long r = row0 + lr;
    if (r >= rows || lr >= bm) continue;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      const int j = j0 + ni * 8 + 2 * tq;
      if (j < J) Out[r * ldout + j] = acc[mi][ni][0];
      if (j + 1 < J) Out[r * ldout + j + 1] = acc[mi][ni][1];
    }
  }
}

// ---------------------------------------------------------------------------
// krgemm2: the same contraction with the Khatri-Rao weights moved to the OUTPUT side,
//   Out[row][j] = sum_p w_p(row) * ( sum_a In[row][a] * Bm[(a*S+p)*ldb + j] ),
// i.e. S plain GEMMs with K = ma sharing the A operand, combined in the epilogue.  Motivation
// (ncu source view of krgemm_kernel<4>, profiles/r01): 19 % of the stall samples sit on the
// per-k-tile barrier and 14 % on the DMULs that generate In*w_p (they queue behind the DMMAs on
// the one FP64 pipe).  Here the A operand is the raw environment slice, so
//   * it is staged by cp.async (no register staging, no multiplies);
//   * every WARP runs its own 3-stage cp.async pipeline over its own 16-row tiles -- no block or
//     group barrier in the main loop, the warps drift freely and keep the DMMA pipe fed (the
//     first version with 128-row tiles shared by 8 warps still lost 9 % to the barrier);
//   * the pipeline runs across tile boundaries (the next tile's stages are in flight during the
//     epilogue of the current one);
//   * the B panel (all K rows x 16 columns) stays resident in shared memory for the whole CTA.
// One persistent 512-thread CTA per SM; every warp owns a contiguous range of rows cut in 16-row
// tiles, a last tile of <= 8 rows issues only half of the MMAs, so all SMs get the same work.
constexpr int G2_BN = 16;     // columns per CTA column tile
constexpr int G2_BNP = 20;    // padded B row: t-stride of 20 doubles is conflict-free per half warp
constexpr int G2_AK = 16;     // a-values per A stage
constexpr int G2_ALD = 20;    // padded A row (row-major tile [16][16+4])
constexpr int G2_WR = 16;     // rows per warp tile
constexpr int G2_WARPS = 16;

__device__ __forceinline__ void cp_async8(double* dst, const double* src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(double* dst, const double* src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// MMAs of k-steps [K0, K1) of one pipeline stage: MI x NI tiles of 8 x 8 per weight index p; only
// the first k4n k-steps of the stage exist (ma need not be a multiple of 16)
template <int S, int MI, int NI, int K0, int K1>
__device__ __forceinline__ void krgemm2_mma(double (&acc)[S][2][2][2], const double* __restrict__ Asb,
                                            const double* __restrict__ Bsb, int pstride, int k4n) {
#pragma unroll
  for (int k4 = K0; k4 < K1; ++k4) {
    if (k4n >= G2_AK / 4 || k4 < k4n) {
      double af[MI], bf[S][NI];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) af[mi] = Asb[mi * 8 * G2_ALD + k4 * 4];
#pragma unroll
      for (int p = 0; p < S; ++p) {
        const double* bp = Bsb + p * pstride + k4 * 4 * G2_BNP;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bf[p][ni] = bp[ni * 8];
      }
#pragma unroll
      for (int p = 0; p < S; ++p)
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma884(acc[p][mi][ni][0], acc[p][mi][ni][1], af[mi], bf[p][ni]);
    }
  }
}
template <int S, int K0, int K1>
__device__ __forceinline__ void krgemm2_mma_sel(double (&acc)[S][2][2][2], const double* __restrict__ Asb,
                                                const double* __restrict__ Bsb, int pstride, int k4n, int bmt,
                                                int nin) {
  // only the MMAs that produce existing rows / columns are issued: a last row tile of <= 8 rows
  // drops rows 8..15, a last column tile of <= 8 columns drops columns 8..15
  if (bmt > 8) {
    if (nin == 2) krgemm2_mma<S, 2, 2, K0, K1>(acc, Asb, Bsb, pstride, k4n);
    else krgemm2_mma<S, 2, 1, K0, K1>(acc, Asb, Bsb, pstride, k4n);
  } else {
    if (nin == 2) krgemm2_mma<S, 1, 2, K0, K1>(acc, Asb, Bsb, pstride, k4n);
    else krgemm2_mma<S, 1, 1, K0, K1>(acc, Asb, Bsb, pstride, k4n);
  }
}

template <int S, int STAGES>
__global__ void __launch_bounds__(32 * G2_WARPS, 1)
krgemm2_kernel(const double* __restrict__ In, long ldin, int ma, const double* __restrict__ f1,
               const double* __restrict__ f2, int div, const double* __restrict__ Bm, long ldb, int J,
               double* __restrict__ Out, long ldout, long rows, int X, int P, int ma_pad, int vec16) {
  extern __shared__ __align__(16) double smem[];
  double* Bs = smem;                                            // [S][ma_pad][G2_BNP]
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  double* Aw = smem + (long)S * ma_pad * G2_BNP + (long)wid * STAGES * G2_WR * G2_ALD;   // [STAGES][16][G2_ALD]
  const int cg = blockIdx.x % X, y = blockIdx.x / X;
  const int coltiles = (J + G2_BN - 1) / G2_BN;
  const int nk = ma_pad / G2_AK;
  // contiguous row range of this warp (worker w of W)
  const long W = (long)G2_WARPS * P, w = (long)G2_WARPS * y + wid;
  const long rbeg = (rows * w) / W, rend = (rows * (w + 1)) / W;
  const long ntile = (rend - rbeg + G2_WR - 1) / G2_WR;

  // Prefetch cursor: the A chunk (16 rows x 16 a) of step (pf_tile, pf_kt) goes to stage pf_stage.
  // All per-lane offsets are loop invariants; the cursor advances without divisions.  The ncu
  // source view of the first version showed 2 IMADs per DMMA and all four warps of a scheduler
  // doing this bookkeeping at the same time (DMMA pipe idle 25 %), so (a) it is cheap now and
  // (b) it is issued in the MIDDLE of the MMA block of the current step, where the DMMA queue
  // hides it.
  // lane -> (row r0 + c*rstep, a-offset q0) of its c-th copy: 4 copies of 16 B or 8 copies of 8 B
  const int r0 = vec16 ? (lane >> 3) : (lane >> 4);
  const int q0 = vec16 ? 2 * (lane & 7) : (lane & 15);
  const int rstep = vec16 ? 4 : 2;
  const int dst_lane = r0 * G2_ALD + q0;
  const long src_lane = (long)r0 * ldin + q0;
  const long src_step = (long)rstep * ldin;
  const int pstride = ma_pad * G2_BNP;
  long pf_tile = 0;
  int pf_kt = 0, pf_stage = 0;
  auto issue_next = [&]() {
    if (pf_tile < ntile) {
      const long row0 = rbeg + pf_tile * G2_WR;
      const int bmt = (int)((rend - row0 < G2_WR) ? (rend - row0) : G2_WR);
      const int a = pf_kt * G2_AK + q0;
      double* dst = Aw + pf_stage * (G2_WR * G2_ALD) + dst_lane;
      const double* src = In + row0 * ldin + pf_kt * G2_AK + src_lane;
#ifdef KR2_NO_CPASYNC
      if (false) {
#else
      if (vec16) {
#endif
        const int abytes = (a < ma) ? ((ma - a >= 2) ? 16 : 8) : 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool ok = (r0 + 4 * c < bmt) && (abytes > 0);
          cp_async16(dst + c * 4 * G2_ALD, ok ? (src + c * src_step) : In, ok ? abytes : 0);
        }
#ifdef KR2_NO_CPASYNC
      } else if (false) {
#else
      } else {
#endif
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const bool ok = (r0 + 2 * c < bmt) && (a < ma);
          cp_async8(dst + c * 2 * G2_ALD, ok ? (src + c * src_step) : In, ok ? 8 : 0);
        }
      }
      if (++pf_kt == nk) {
        pf_kt = 0;
        ++pf_tile;
      }
      if (++pf_stage == STAGES) pf_stage = 0;
    }
    cp_async_commit();
  };
  (void)rstep;

  for (int ct = cg; ct < coltiles; ct += X) {
    const int j0 = ct * G2_BN;
    const int nin = (J - j0 > 8) ? 2 : 1;
    __syncthreads();   // every warp is done with the previous B panel
    for (int idx = tid; idx < S * ma_pad * G2_BN; idx += 32 * G2_WARPS) {
      const int n = idx & (G2_BN - 1);
      const int k2 = idx >> 4;            // = a*S + p
      const int a = k2 / S, p = k2 - a * S;
      const double v = (a < ma && j0 + n < J) ? Bm[(long)k2 * ldb + j0 + n] : 0.0;
      Bs[(p * ma_pad + a) * G2_BNP + n] = v;
    }
    __syncthreads();
    double acc[S][2][2][2];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) acc[p][mi][ni][0] = acc[p][mi][ni][1] = 0.0;
    pf_tile = 0;
    pf_kt = 0;
    pf_stage = 0;
#pragma unroll
    for (int s0 = 0; s0 < STAGES - 1; ++s0) issue_next();
    int stage = 0;
    for (long tile = 0; tile < ntile; ++tile) {
      const long row0 = rbeg + tile * G2_WR;
      const int bmt = (int)((rend - row0 < G2_WR) ? (rend - row0) : G2_WR);
      for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncwarp();   // this stage visible to all lanes; all lanes finished reading the previous one
        const double* Asb = Aw + stage * (G2_WR * G2_ALD) + g * G2_ALD + tq;
        const double* Bsb = Bs + (kt * G2_AK + tq) * G2_BNP + g;
        const int rem = ma - kt * G2_AK;
        const int k4n = (rem >= G2_AK) ? (G2_AK / 4) : ((rem + 3) >> 2);
        if (k4n == G2_AK / 4 && bmt > 8 && nin == 2) {
          // common case, no conditions inside: the compiler software-pipelines the fragment loads
          // of k-step k+1 under the MMAs of k-step k
          krgemm2_mma<S, 2, 2, 0, 1>(acc, Asb, Bsb, pstride, 4);
          issue_next();   // refills the stage read in the previous step
          krgemm2_mma<S, 2, 2, 1, G2_AK / 4>(acc, Asb, Bsb, pstride, 4);
        } else {
          krgemm2_mma_sel<S, 0, 1>(acc, Asb, Bsb, pstride, k4n, bmt, nin);
          issue_next();
          krgemm2_mma_sel<S, 1, G2_AK / 4>(acc, Asb, Bsb, pstride, k4n, bmt, nin);
        }
        if (++stage == STAGES) stage = 0;
      }
      // tile finished: combine the S partial products with the row's weights
#ifdef KR2_NO_EPILOGUE
      if (acc[0][0][0][0] == 1.2345e300)
#endif
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int lr = mi * 8 + g;
        if (lr < bmt) {
          const long r = row0 + lr;
          double wgt[S];
          kr_weights<S>(f1, f2, r / div, wgt);
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) {
            double o0 = 0.0, o1 = 0.0;
#pragma unroll
            for (int p = 0; p < S; ++p) {
              o0 = fma(wgt[p], acc[p][mi][ni][0], o0);
              o1 = fma(wgt[p], acc[p][mi][ni][1], o1);
            }
            const int j = j0 + ni * 8 + 2 * tq;
            if (j < J) Out[r * ldout + j] = o0;
            if (j + 1 < J) Out[r * ldout + j + 1] = o1;
          }
        }
      }
#pragma unroll
      for (int p = 0; p < S; ++p)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) acc[p][mi][ni][0] = acc[p][mi][ni][1] = 0.0;
    }
    cp_async_wait<0>();
    __syncwarp();
  }
}

// returns false when the resident B panel does not fit (large ma): caller falls back to krgemm_kernel
template <int S>
static bool krgemm2_launch(cudaStream_t st, const double* In, long ldin, int ma, const double* f1, const double* f2,
                           int div, const double* Bm, long ldb, int J, double* Out, long ldout, long rows,
                           int num_sm) {
  const int ma_pad = ((ma + G2_AK - 1) / G2_AK) * G2_AK;
  const size_t bbytes = (size_t)S * ma_pad * G2_BNP * sizeof(double);
  const size_t a3 = (size_t)G2_WARPS * 3 * G2_WR * G2_ALD * sizeof(double);
  const size_t a2 = (size_t)G2_WARPS * 2 * G2_WR * G2_ALD * sizeof(double);
  const size_t lim = 227 * 1024;
  int stages = 0;
  if (bbytes + a3 <= lim) stages = 3;
  else if (bbytes + a2 <= lim) stages = 2;
  if (!stages) return false;
  const int coltiles = (J + G2_BN - 1) / G2_BN;
  // X column groups x P CTAs each: maximise (SMs used) x (column-tile balance)
  int X = 1, P = num_sm;
  double best = -1.0;
  for (int x = 1; x <= coltiles && x <= num_sm; ++x) {
    const int pp = num_sm / x;
    const double util = (double)(x * pp) / num_sm;
    const double bal = (double)coltiles / ((double)x * ((coltiles + x - 1) / x));
    // rows are split over 16*pp warps: very short ranges waste the pipeline fill
    const double rowsper = (double)rows / ((double)G2_WARPS * pp);
    const double fill = rowsper / (rowsper + 4.0);
    const double score = util * bal * fill;
    if (score > best + 1e-12) {
      best = score;
      X = x;
      P = pp;
    }
  }
  const int vec16 = ((ldin & 1) == 0 && (((size_t)In) & 15) == 0) ? 1 : 0;
  static unsigned long long attr = 0;   // one bit per device: function attributes are per device
  if (first_on_device(attr)) {
    cudaFuncSetAttribute(krgemm2_kernel<S, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim);
    cudaFuncSetAttribute(krgemm2_kernel<S, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim);
  }
  if (stages == 3)
    krgemm2_kernel<S, 3><<<X * P, 32 * G2_WARPS, bbytes + a3, st>>>(In, ldin, ma, f1, f2, div, Bm, ldb, J, Out, ldout,
                                                                  rows, X, P, ma_pad, vec16);
  else
    krgemm2_kernel<S, 2><<<X * P, 32 * G2_WARPS, bbytes + a2, st>>>(In, ldin, ma, f1, f2, div, Bm, ldb, J, Out, ldout,
                                                                  rows, X, P, ma_pad, vec16);
  return true;
}

static int g_krgemm_variant = -1;
void krgemm_set_variant(int v) { g_krgemm_variant = v; }

void krgemm(cudaStream_t st, int S, const double* In, long ldin, int ma, const double* f1, const double* f2,
            int div, const double* Bm, long ldb, int J, double* Out, long ldout, long rows, int num_sm) {
  if (rows <= 0 || J <= 0) return;
  if (g_krgemm_variant < 0) {   // TNML_KRGEMM=1: register-staged kernel only (A/B comparisons)
    const char* e = getenv("TNML_KRGEMM");
    g_krgemm_variant = e ? atoi(e) : 2;
  }
  if (g_krgemm_variant >= 2 && rows >= 512) {   // (3 = tcgen05 projection, chosen by the caller; here it means 2)
    const bool ok = (S == 2) ? krgemm2_launch<2>(st, In, ldin, ma, f1, f2, div, Bm, ldb, J, Out, ldout, rows, num_sm)
                             : krgemm2_launch<4>(st, In, ldin, ma, f1, f2, div, Bm, ldb, J, Out, ldout, rows, num_sm);
    if (ok) return;
  }
  const int coltiles = (J + BN - 1) / BN;
  // rows per tile (multiple of the 16-row warp tile): minimise the makespan
  // ceil(tiles / SMs) * (bm + fixed per-tile overhead) -- the FP64 pipe is the bottleneck, so an
  // SM's time is the sum of the rows it is handed
  long bm = BM, best = -1;
  for (long cand = BM; cand >= 32; cand -= 16) {
    long tiles = ((rows + cand - 1) / cand) * coltiles;
    long cost = ((tiles + num_sm - 1) / num_sm) * (cand + 6);
    if (best < 0 || cost < best) {
      best = cost;
      bm = cand;
    }
  }
  dim3 grid((unsigned)((rows + bm - 1) / bm), (unsigned)coltiles);
  size_t sh = (size_t)(2 * BK * LDA + 2 * BK * LDB) * sizeof(double);
  static unsigned long long attr = 0;   // one bit per device: function attributes are per device
  if (first_on_device(attr)) {
    cudaFuncSetAttribute(krgemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(krgemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  }
  if (S == 2)
    krgemm_kernel<2><<<grid, NTHR, sh, st>>>(In, ldin, ma, f1, f2, div, Bm, ldb, J, Out, ldout, rows, (int)bm);
  else
    krgemm_kernel<4><<<grid, NTHR, sh, st>>>(In, ldin, ma, f1, f2, div, Bm, ldb, J, Out, ldout, rows, (int)bm);
}

// ---------------------------------------------------------------------------
// Gpart[split][(a*S+p)][j] = sum_{row in split} In[row][a] * w_p(row) * Z[row][j].  M = S*ma, K = rows.
template <int S>
__global__ void __launch_bounds__(NTHR, 2)
krgram_kernel(const double* __restrict__ In, long ldin, int ma, const double* __restrict__ f1,
              const double* __restrict__ f2, const double* __restrict__ Z, long ldz, int J,
              double* __restrict__ Gpart, long rows, long rows_per_split) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                  // [2][BK][LDA]   column index = (a_local*S+p)
  double* Bs = smem + 2 * BK * LDA;   // [2][BK][LDB]
  constexpr int AM = BM / S;         // a-values per m-tile
  constexpr int APT = AM / 16;       // a-values per loader thread (16 threads per row)
  const int t = threadIdx.x;
  const int lane = t & 31, wm = t >> 5, g = lane >> 2, tq = lane & 3;
  const int a0 = blockIdx.x * AM;
  const int j0 = blockIdx.y * BN;
  const long rbeg = (long)blockIdx.z * rows_per_split;
  const long rend = (rbeg + rows_per_split < rows) ? (rbeg + rows_per_split) : rows;
  const int lrow = t >> 4, ca = (t & 15) * APT, c4 = (t & 15) * 4;

  double acc[2][8][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double ra[APT], rz[4], w[S];
  const long nrow = (rend > rbeg) ? (rend - rbeg) : 0;
  const int nk = (int)((nrow + BK - 1) / BK);

  auto gload = [&](int kt) {
    const long r = rbeg + (long)kt * BK + lrow;
    const bool ok = r < rend;
#pragma unroll
    for (int i = 0; i < APT; ++i) {
      int a = a0 + ca + i;
      ra[i] = (ok && a < ma) ? In[r * ldin + a] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int j = j0 + c4 + i;
      rz[i] = (ok && j < J) ? Z[r * ldz + j] : 0.0;
    }
#pragma unroll
    for (int p = 0; p < S; ++p) w[p] = 0.0;
    if (ok) kr_weights<S>(f1, f2, r, w);
  };
  auto sstore = [&](int buf) {
    // shared-memory column index = p*AM + a_local (p-major): consecutive lanes write
    // consecutive 16-byte words (the a-major order gave 8-way bank conflicts)
    double* A = As + (buf * BK + lrow) * LDA + ca;
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int i = 0; i < APT; i += 2)
        *reinterpret_cast<double2*>(A + p * AM + i) = make_double2(ra[i] * w[p], ra[i + 1] * w[p]);
    double2* B = reinterpret_cast<double2*>(Bs + (buf * BK + lrow) * LDB + c4);
    B[0] = make_double2(rz[0], rz[1]);
    B[1] = make_double2(rz[2], rz[3]);
  };

  if (nk > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
    tile_mma(As + buf * BK * LDA + wm * 16, Bs + buf * BK * LDB, acc, g, tq);
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  const long M2 = (long)S * ma;
  double* Gp = Gpart + (long)blockIdx.z * (M2 * J);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int ml = wm * 16 + mi * 8 + g;          // p-major local column
    const int pp = ml / AM, al = ml - pp * AM;
    if (a0 + al >= ma) continue;
    const long m2 = (long)(a0 + al) * S + pp;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      const int j = j0 + ni * 8 + 2 * tq;
      if (j < J) Gp[m2 * J + j] = acc[mi][ni][0];
      if (j + 1 < J) Gp[m2 * J + j + 1] = acc[mi][ni][1];
    }
  }
}

// ---------------------------------------------------------------------------
// krgram2: the same split-K contraction with RAW operand tiles.  The environment slice In, the
// back-propagated Z and the two feature pairs of 16 images are staged by cp.async, three stages
// deep; the Khatri-Rao weight w_p(n) = f1[n][p>>1] * f2[n][p&1] is applied when the A fragment is
// loaded (one DMUL per fragment element: the warp's 16 rows of the tile belong to one p).
// Removes from krgram_kernel the register-staged loads (18 % long-scoreboard stalls in the ncu
// capture), the DMUL -> STS chain in front of every barrier and 4x of the shared-memory stores.
constexpr int R2_BK = 16;      // images per stage
constexpr int R2_AM = 32;      // a-values per M tile (x 4 weights = 128 rows)
constexpr int R2_ALD = 36;     // padded In row: 8t + 2g bank pattern, conflict-free per half warp
constexpr int R2_ZLD = 68;     // padded Z row
constexpr int R2_STAGES = 3;
constexpr int R2_STAGE_DOUBLES = R2_BK * (R2_ALD + R2_ZLD + 4);

__global__ void __launch_bounds__(NTHR, 2)
krgram2_kernel(const double* __restrict__ In, long ldin, int ma, const double* __restrict__ f1,
               const double* __restrict__ f2, const double* __restrict__ Z, long ldz, int J,
               double* __restrict__ Gpart, long rows, long rows_per_split, int vecA, int vecZ) {
  extern __shared__ __align__(16) double smem[];
  constexpr int S = 4;
  const int t = threadIdx.x;
  const int lane = t & 31, wm = t >> 5, g = lane >> 2, tq = lane & 3;
  const int a0 = blockIdx.x * R2_AM;
  const int j0 = blockIdx.y * BN;
  const long rbeg = (long)blockIdx.z * rows_per_split;
  const long rend = (rbeg + rows_per_split < rows) ? (rbeg + rows_per_split) : rows;
  const long nrow = (rend > rbeg) ? (rend - rbeg) : 0;
  const int nk = (int)((nrow + R2_BK - 1) / R2_BK);

  auto issue = [&](int kt) {
    double* As = smem + (long)(kt % R2_STAGES) * R2_STAGE_DOUBLES;   // [16][R2_ALD]
    double* Zs = As + R2_BK * R2_ALD;                                // [16][R2_ZLD]
    double* Fs = Zs + R2_BK * R2_ZLD;                                // [16][4] = f1[n][0..1], f2[n][0..1]
    const long r0 = rbeg + (long)kt * R2_BK;
    if (vecA) {
      const int lr = t >> 4, q = 2 * (t & 15), a = a0 + q;
      const bool ok = (r0 + lr < rend) && (a < ma);
      cp_async16(As + lr * R2_ALD + q, ok ? (In + (r0 + lr) * ldin + a) : In, ok ? ((ma - a >= 2) ? 16 : 8) : 0);
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = t + i * NTHR, lr = c >> 5, q = c & 31, a = a0 + q;
        const bool ok = (r0 + lr < rend) && (a < ma);
        cp_async8(As + lr * R2_ALD + q, ok ? (In + (r0 + lr) * ldin + a) : In, ok ? 8 : 0);
      }
    }
    if (vecZ) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = t + i * NTHR, lr = c >> 5, q = 2 * (c & 31), j = j0 + q;
        const bool ok = (r0 + lr < rend) && (j < J);
        cp_async16(Zs + lr * R2_ZLD + q, ok ? (Z + (r0 + lr) * ldz + j) : Z, ok ? ((J - j >= 2) ? 16 : 8) : 0);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = t + i * NTHR, lr = c >> 6, q = c & 63, j = j0 + q;
        const bool ok = (r0 + lr < rend) && (j < J);
        cp_async8(Zs + lr * R2_ZLD + q, ok ? (Z + (r0 + lr) * ldz + j) : Z, ok ? 8 : 0);
      }
    }
    if (t < 2 * R2_BK) {
      const int lr = t & (R2_BK - 1);
      const double* f = (t < R2_BK) ? f1 : f2;
      const bool ok = (r0 + lr < rend);
      cp_async16(Fs + lr * 4 + ((t < R2_BK) ? 0 : 2), ok ? (f + (r0 + lr) * 2) : f, ok ? 16 : 0);
    }
  };

  double acc[2][8][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int s0 = 0; s0 < R2_STAGES - 1; ++s0) {
    if (s0 < nk) issue(s0);
    cp_async_commit();
  }
  const int pp = wm >> 1;                          // weight index of this warp's 16 rows
  const int fs_off = pp >> 1, fq_off = 2 + (pp & 1);
  const int arow = (wm & 1) * 16 + g;              // a-value (within the tile) of fragment row g
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<R2_STAGES - 2>();
    __syncthreads();
    if (kt + R2_STAGES - 1 < nk) issue(kt + R2_STAGES - 1);
    cp_async_commit();
    const double* As = smem + (long)(kt % R2_STAGES) * R2_STAGE_DOUBLES;
    const double* Zs = As + R2_BK * R2_ALD;
    const double* Fs = Zs + R2_BK * R2_ZLD;
#pragma unroll
    for (int k4 = 0; k4 < R2_BK / 4; ++k4) {
      const int n = k4 * 4 + tq;
      const double w = Fs[n * 4 + fs_off] * Fs[n * 4 + fq_off];
      double af[2], bf[8];
      af[0] = As[n * R2_ALD + arow] * w;
      af[1] = As[n * R2_ALD + arow + 8] * w;
      const double* bp = Zs + n * R2_ZLD + g;
#pragma unroll
      for (int i = 0; i < 8; ++i) bf[i] = bp[i * 8];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();
  const long M2 = (long)S * ma;
  double* Gp = Gpart + (long)blockIdx.z * (M2 * J);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int al = (wm & 1) * 16 + mi * 8 + g;
    if (a0 + al >= ma) continue;
    const long m2 = (long)(a0 + al) * S + pp;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      const int j = j0 + ni * 8 + 2 * tq;
      if (j < J) Gp[m2 * J + j] = acc[mi][ni][0];
      if (j + 1 < J) Gp[m2 * J + j + 1] = acc[mi][ni][1];
    }
  }
}

static int g_krgram_variant = -1;
void krgram_set_variant(int v) { g_krgram_variant = v; }

int krgram_splits(int ma, int S, int J, long rows, int num_sm) {
  int mt = (S * ma + BM - 1) / BM, nt = (J + BN - 1) / BN;
  long tiles = (long)mt * nt;
  int s = (int)((2L * num_sm) / tiles);           // one full wave of two CTAs per SM
  if (s < 1) s = 1;
  long maxs = (rows + 8 * BK - 1) / (8 * BK);     // at least 128 rows per split
  if (s > maxs) s = (int)maxs;
  if (s < 1) s = 1;
  if (s > 65535) s = 65535;
  return s;
}

void krgram(cudaStream_t st, int S, const double* In, long ldin, int ma, const double* f1, const double* f2,
            const double* Z, long ldz, int J, double* Gpart, long rows, int nsplit) {
  dim3 grid((unsigned)((S * ma + BM - 1) / BM), (unsigned)((J + BN - 1) / BN), (unsigned)nsplit);
  long rps = (rows + nsplit - 1) / nsplit;
  rps = ((rps + BK - 1) / BK) * BK;
  size_t sh = (size_t)(2 * BK * LDA + 2 * BK * LDB) * sizeof(double);
  static unsigned long long attr = 0;   // one bit per device: function attributes are per device
  if (first_on_device(attr)) {
    cudaFuncSetAttribute(krgram_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(krgram_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  }
  if (g_krgram_variant < 0) {   // TNML_KRGRAM=1: register-staged kernel only
    const char* e = getenv("TNML_KRGRAM");
    g_krgram_variant = e ? atoi(e) : 2;
  }
  if (S == 4 && g_krgram_variant == 2) {
    static unsigned long long attr2 = 0;   // one bit per device: function attributes are per device
    const size_t sh2 = (size_t)R2_STAGES * R2_STAGE_DOUBLES * sizeof(double);
    if (first_on_device(attr2)) {
      cudaFuncSetAttribute(krgram2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    }
    const int vecA = ((ldin & 1) == 0 && (((size_t)In) & 15) == 0) ? 1 : 0;
    const int vecZ = ((ldz & 1) == 0 && (((size_t)Z) & 15) == 0) ? 1 : 0;
    krgram2_kernel<<<grid, NTHR, sh2, st>>>(In, ldin, ma, f1, f2, Z, ldz, J, Gpart, rows, rps, vecA, vecZ);
    return;
  }
  if (S == 2)
    krgram_kernel<2><<<grid, NTHR, sh, st>>>(In, ldin, ma, f1, f2, Z, ldz, J, Gpart, rows, rps);
  else
    krgram_kernel<4><<<grid, NTHR, sh, st>>>(In, ldin, ma, f1, f2, Z, ldz, J, Gpart, rows, rps);
}

__global__ void reduce_partials_kernel(const double* __restrict__ Gpart, int nsplit, long n,
                                       double* __restrict__ G) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < nsplit; ++k) s += Gpart[(long)k * n + i];
  G[i] = s;
}
void reduce_partials(cudaStream_t st, const double* Gpart, int nsplit, long n, double* G) {
  if (n <= 0) return;
  reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Gpart, nsplit, n, G);
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Statistics of one pass (cost per label, #correct, sum |P|^2): every CTA leaves 16 partial sums,
// the CTA that finishes last adds them up in a fixed order (deterministic, independent of which CTA
// happens to be last) and resets the ticket.  Replaces the separate <<<1,512>>> reduce launch that
// followed each of the 9 label-environment passes of a bond update (17 us each, cold).
__device__ __forceinline__ void stats_finalize(const double* __restrict__ sp, int nblocks, double* __restrict__ stats,
                                               unsigned* __restrict__ ticket) {
  __shared__ int is_last;
  __shared__ double part[8][16];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < 128) {
    const int i = threadIdx.x & 15, g = threadIdx.x >> 4;
    double s = 0.0;
    for (int b = g; b < nblocks; b += 8) s += __ldcg(sp + (long)b * 16 + i);
    part[g][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += part[g][threadIdx.x];
    stats[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// MCH = ceil(m/32) <= 4: the image's whole fat environment (10 x m doubles) is loaded with
// all loads in flight at once and stays in registers for the backward contraction (one HBM
// read per image; the loop version below issued 11 loads per round trip and re-read F).
// MCH = 0: generic loop version for m > 128.
// the label-carrying environment is read once per launch (9.6 KB per image, 576 MB per launch at config 3):
// streaming (evict-first) loads keep it from flushing the operands the neighbouring launches re-read
// (int8 planes of the thin environment, Q, Z) out of the 126 MB L2
#ifndef FAT_NO_STREAM
#define FAT_LDF(p) __ldcs(p)
#else
#define FAT_LDF(p) (*(p))
#endif
template <int MODE, int MCH>
__global__ void __launch_bounds__((MCH > 0) ? 384 : 512, 1)
fat_kernel_t(const double* __restrict__ Q, const double* __restrict__ F, int m,
             const int32_t* __restrict__ labels, double* __restrict__ P, double* __restrict__ Z,
             int32_t* __restrict__ pred, double* __restrict__ stats_partial, long NT,
             double* __restrict__ stats_out, unsigned* __restrict__ ticket) {
  constexpr int WPB = (MCH > 0) ? 12 : 16;  // warps per block: one persistent CTA per SM
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long gw = (long)blockIdx.x * WPB + warp;
  const long nw = (long)gridDim.x * WPB;
  double cst[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) cst[l] = 0.0;
  double ncor = 0.0, pap = 0.0;

  for (long n = gw; n < NT; n += nw) {
    const double* q = Q + n * m;
    const double* Fn = F + n * (long)NL * m;
    double pl[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) pl[l] = 0.0;
    constexpr int C = (MCH > 0) ? MCH : 1;
    constexpr bool GIVEN_P = (MODE == FAT_BWD || MODE == FAT_BWD_OUTER);
    constexpr bool DO_Z = (MODE == FAT_GRAD || MODE == FAT_BWD);
    constexpr bool DO_ZOUTER = (MODE == FAT_GRAD_OUTER || MODE == FAT_BWD_OUTER);
    double qv[C];
    double fv[NL][C];
    if (MCH > 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int f = lane + 32 * c;
        const bool ok = f < m;
        qv[c] = (ok && !(GIVEN_P && DO_Z)) ? FAT_LDF(q + f) : 0.0;     // FAT_BWD needs no Q; Q is consumed here once
#pragma unroll
        for (int l = 0; l < NL; ++l) fv[l][c] = (ok && !(GIVEN_P && DO_ZOUTER)) ? FAT_LDF(Fn + (long)l * m + f) : 0.0;
      }
      if (!GIVEN_P) {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int l = 0; l < NL; ++l) pl[l] = fma(qv[c], fv[l][c], pl[l]);
      }
    } else if (!GIVEN_P) {
      for (int f = lane; f < m; f += 32) {
        double x = q[f];
#pragma unroll
        for (int l = 0; l < NL; ++l) pl[l] = fma(x, Fn[(long)l * m + f], pl[l]);
      }
    }
    if (GIVEN_P) {
#pragma unroll
      for (int l = 0; l < NL; ++l) pl[l] = P[n * NL + l];
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int l = 0; l < NL; ++l) pl[l] += __shfl_xor_sync(0xffffffffu, pl[l], o);
      }
    }
    if (P != nullptr && !GIVEN_P) {
      double mine = 0.0;
#pragma unroll
      for (int l = 0; l < NL; ++l) mine = (lane == l) ? pl[l] : mine;
      if (lane < NL) P[n * NL + lane] = mine;
    }
    if (MODE == FAT_PAP) {
#pragma unroll
      for (int l = 0; l < NL; ++l) pap = fma(pl[l], pl[l], pap);
    } else {
      const int lab = labels[n];
      int am = 0;
      double mx = fabs(pl[0]);
#pragma unroll
      for (int l = 1; l < NL; ++l) {
        double w = fabs(pl[l]);
        if (w > mx) {  // first strict maximum, util.h:42-57
          mx = w;
          am = l;
        }
      }
      ncor += (am == lab) ? 1.0 : 0.0;
      if (pred != nullptr && lane == 0) pred[n] = am;
      double dp[NL];
      double e = 0.0;
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        dp[l] = ((l == lab) ? 1.0 : 0.0) - pl[l];
        e = fma(dp[l], dp[l], e);
      }
#pragma unroll
      for (int l = 0; l < NL; ++l) cst[l] += (l == lab) ? e : 0.0;
      if (DO_Z) {
        double* z = Z + n * m;
        if (MCH > 0) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int f = lane + 32 * c;
            double sacc = 0.0;
#pragma unroll
            for (int l = 0; l < NL; ++l) sacc = fma(dp[l], fv[l][c], sacc);
            if (f < m) z[f] = sacc;
          }
        } else {
          for (int f = lane; f < m; f += 32) {
            double sacc = 0.0;
#pragma unroll
            for (int l = 0; l < NL; ++l) sacc = fma(dp[l], Fn[(long)l * m + f], sacc);
            z[f] = sacc;
          }
        }
      } else if (DO_ZOUTER) {
        double* z = Z + n * (long)NL * m;
        if (MCH > 0) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int f = lane + 32 * c;
            if (f < m) {
#pragma unroll
              for (int l = 0; l < NL; ++l) z[(long)l * m + f] = dp[l] * qv[c];
            }
          }
        } else {
          for (int f = lane; f < m; f += 32) {
            double x = q[f];
#pragma unroll
            for (int l = 0; l < NL; ++l) z[(long)l * m + f] = dp[l] * x;
          }
        }
      }
    }
  }
  __shared__ double red[WPB][12];
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < NL; ++l) red[warp][l] = cst[l];
    red[warp][10] = ncor;
    red[warp][11] = pap;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double s = 0.0;
    if (threadIdx.x < 12)
      for (int w = 0; w < WPB; ++w) s += red[w][threadIdx.x];
    stats_partial[(long)blockIdx.x * 16 + threadIdx.x] = s;
  }
  stats_finalize(stats_partial, gridDim.x, stats_out, ticket);
}

// ---------------------------------------------------------------------------
// fat_bulk_kernel: the same per-image work with the label-carrying environment of an image
// (NL x m doubles = 9.6 KB at m = 120, contiguous in HBM) brought in by ONE bulk asynchronous copy
// (TMA, cp.async.bulk ... mbarrier::complete_tx) into a warp-private double buffer in shared
// memory: the next image's 9.6 KB are in flight while the warp works on the current one, no
// registers are tied up by the loads, and the number of resident warps is set by shared memory
// (11 warps x 2 x 9.6 KB at m = 120) instead of by the 168 registers of fat_kernel_t.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1)
fat_bulk_kernel(const double* __restrict__ Q, const double* __restrict__ F, int m,
                const int32_t* __restrict__ labels, double* __restrict__ P, double* __restrict__ Z,
                int32_t* __restrict__ pred, double* __restrict__ stats_partial, int nblocks_stats, long NT,
                double* __restrict__ stats_out, unsigned* __restrict__ ticket) {
  constexpr bool GIVEN_P = (MODE == FAT_BWD);
  constexpr bool DO_Z = (MODE == FAT_GRAD || MODE == FAT_BWD);
  constexpr int C = 4;                       // m <= 128
  constexpr int NBUF = 2;                    // ring of NBUF buffers per warp: NBUF-1 images in flight
  extern __shared__ __align__(128) unsigned char smraw[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smraw);     // [16 warps][NBUF] (<= 512 B)
  double* red = reinterpret_cast<double*>(smraw + 512);                        // [16][12]
  double* bufs = reinterpret_cast<double*>(smraw + 512 + 16 * 12 * sizeof(double));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int bufd = NL * m;
  const unsigned bytes = (unsigned)bufd * sizeof(double);
  double* mybuf = bufs + (long)warp * NBUF * bufd;
  const unsigned bar0 = smem_u32(bars + warp * NBUF);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NBUF; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long gw = (long)blockIdx.x * wpb + warp;
  const long nw = (long)gridDim.x * wpb;
  auto issue = [&](long n, int b) {
    if (lane == 0) {
      mbar_expect_tx(bar0 + 8 * b, bytes);
      bulk_g2s(mybuf + (long)b * bufd, F + n * (long)bufd, bytes, bar0 + 8 * b);
    }
  };
  double cst[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) cst[l] = 0.0;
  double ncor = 0.0, pap = 0.0;
#pragma unroll
  for (int i = 0; i < NBUF - 1; ++i)
    if (gw + i * nw < NT) issue(gw + i * nw, i);
  int it = 0, b = 0;
  unsigned par = 0;
  for (long n = gw; n < NT; n += nw, ++it) {
    __syncwarp();                               // every lane is done reading the buffer of iteration it-1
    {
      const int bn = (b == 0) ? (NBUF - 1) : (b - 1);   // = (it + NBUF - 1) % NBUF: freed by iteration it-1
      if (n + (NBUF - 1) * nw < NT) issue(n + (NBUF - 1) * nw, bn);
    }
    double qv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int f = lane + 32 * c;
      qv[c] = (!GIVEN_P && f < m) ? Q[n * m + f] : 0.0;
    }
    double pl[NL];
    if (GIVEN_P) {
#pragma unroll
      for (int l = 0; l < NL; ++l) pl[l] = P[n * NL + l];
    }
    const int lab = (MODE == FAT_PAP) ? 0 : labels[n];
    mbar_wait(bar0 + 8 * b, par);
    const double* Fn = mybuf + (long)b * bufd;
    if (!GIVEN_P) {
#pragma unroll
      for (int l = 0; l < NL; ++l) pl[l] = 0.0;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int f = lane + 32 * c;
        if (f < m) {
#pragma unroll
          for (int l = 0; l < NL; ++l) pl[l] = fma(qv[c], Fn[l * m + f], pl[l]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int l = 0; l < NL; ++l) pl[l] += __shfl_xor_sync(0xffffffffu, pl[l], o);
      }
      if (P != nullptr) {
        double mine = 0.0;
#pragma unroll
        for (int l = 0; l < NL; ++l) mine = (lane == l) ? pl[l] : mine;
        if (lane < NL) P[n * NL + lane] = mine;
      }
    }
    if (MODE == FAT_PAP) {
#pragma unroll
      for (int l = 0; l < NL; ++l) pap = fma(pl[l], pl[l], pap);
    } else {
      int am = 0;
      double mx = fabs(pl[0]);
#pragma unroll
      for (int l = 1; l < NL; ++l) {
        double w = fabs(pl[l]);
        if (w > mx) {  // first strict maximum, util.h:42-57
          mx = w;
          am = l;
        }
      }
      ncor += (am == lab) ? 1.0 : 0.0;
      if (pred != nullptr && lane == 0) pred[n] = am;
      double dp[NL];
      double e = 0.0;
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        dp[l] = ((l == lab) ? 1.0 : 0.0) - pl[l];
        e = fma(dp[l], dp[l], e);
      }
#pragma unroll
      for (int l = 0; l < NL; ++l) cst[l] += (l == lab) ? e : 0.0;
      if (DO_Z) {
        double* z = Z + n * m;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int f = lane + 32 * c;
          if (f < m) {
            double sacc = 0.0;
#pragma unroll
            for (int l = 0; l < NL; ++l) sacc = fma(dp[l], Fn[l * m + f], sacc);
            z[f] = sacc;
          }
        }
      }
    }
    if (++b == NBUF) {
      b = 0;
      par ^= 1;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < NL; ++l) red[warp * 12 + l] = cst[l];
    red[warp * 12 + 10] = ncor;
    red[warp * 12 + 11] = pap;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double s = 0.0;
    if (threadIdx.x < 12)
      for (int w = 0; w < wpb; ++w) s += red[w * 12 + threadIdx.x];
    stats_partial[(long)blockIdx.x * 16 + threadIdx.x] = s;
  }
  (void)nblocks_stats;
  stats_finalize(stats_partial, gridDim.x, stats_out, ticket);
}

static int g_fat_variant = -1;
void fat_set_variant(int v) { g_fat_variant = v; }

template <int MODE>
static bool fat_bulk_launch(cudaStream_t st, const double* Q, const double* F, int m, const int32_t* labels, double* P,
                            double* Z, int32_t* pred, double* stats_partial, int nblocks, long NT, double* stats_out,
                            unsigned* ticket) {
  if (m > 128 || m < 4) return false;
  const size_t per_warp = (size_t)2 * NL * m * sizeof(double);   // NBUF buffers
  int wpb = (int)((220 * 1024 - 512 - 16 * 12 * sizeof(double)) / per_warp);
  if (wpb > 16) wpb = 16;
  if (wpb < 4) return false;
  const int grid = nblocks;        // = number of SMs (fat_blocks): one persistent CTA per SM
  if (grid < 1) return false;
  const size_t sh = 512 + 16 * 12 * sizeof(double) + (size_t)wpb * per_warp;
  static unsigned long long attr = 0;   // one bit per device: function attributes are per device
  if (first_on_device(attr)) {
    cudaFuncSetAttribute(fat_bulk_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  }
  fat_bulk_kernel<MODE><<<grid, 32 * wpb, sh, st>>>(Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out,
                                                    ticket);
  return true;
}

int fat_blocks(int num_sm) { return num_sm; }

template <int MODE>
static void fat_launch(cudaStream_t st, const double* Q, const double* F, int m, const int32_t* labels, double* P,
                       double* Z, int32_t* pred, double* sp, int nb, long NT, double* so, unsigned* tk) {
  const int mch = (m + 31) / 32;
  if (mch == 1)
    fat_kernel_t<MODE, 1><<<nb, 384, 0, st>>>(Q, F, m, labels, P, Z, pred, sp, NT, so, tk);
  else if (mch == 2)
    fat_kernel_t<MODE, 2><<<nb, 384, 0, st>>>(Q, F, m, labels, P, Z, pred, sp, NT, so, tk);
  else if (mch == 3)
    fat_kernel_t<MODE, 3><<<nb, 384, 0, st>>>(Q, F, m, labels, P, Z, pred, sp, NT, so, tk);
  else if (mch == 4)
    fat_kernel_t<MODE, 4><<<nb, 384, 0, st>>>(Q, F, m, labels, P, Z, pred, sp, NT, so, tk);
  else
    fat_kernel_t<MODE, 0><<<nb, 512, 0, st>>>(Q, F, m, labels, P, Z, pred, sp, NT, so, tk);
}

void fat_kernel(cudaStream_t st, int mode, const double* Q, const double* F, int m, const int32_t* labels,
                double* P, double* Z, int32_t* pred, double* stats_partial, int nblocks, long NT, double* stats_out,
                unsigned* ticket) {
  // Measured on B200 (bench.py, m = 120): the register-resident kernel streams the label environment
  // at 5.06 TB/s, the bulk-copy kernel at 4.77 TB/s with 2 buffers per warp (11 warps/SM) and
  // 4.70 TB/s with 3 (7 warps/SM): shared memory caps the bytes in flight at about the level the
  // 12 x 168-register warps reach anyway.  So the register kernel stays the default (variant 1);
  // TNML_FAT=2 / tnml_set_option("fat_variant", 2) selects the bulk-copy kernel.
  if (g_fat_variant < 0) {
    const char* e = getenv("TNML_FAT");
    g_fat_variant = e ? atoi(e) : 1;
  }
  if (g_fat_variant == 2 && NT >= 1024) {
    bool done = false;
    if (mode == FAT_GRAD) done = fat_bulk_launch<FAT_GRAD>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
    else if (mode == FAT_PAP) done = fat_bulk_launch<FAT_PAP>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
    else if (mode == FAT_COST) done = fat_bulk_launch<FAT_COST>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
    else if (mode == FAT_BWD) done = fat_bulk_launch<FAT_BWD>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
    if (done) return;
  }
  switch (mode) {
    case FAT_GRAD:
      fat_launch<FAT_GRAD>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
      break;
    case FAT_PAP:
      fat_launch<FAT_PAP>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
      break;
    case FAT_COST:
      fat_launch<FAT_COST>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
      break;
    case FAT_BWD:
      fat_launch<FAT_BWD>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
      break;
    case FAT_BWD_OUTER:
      fat_launch<FAT_BWD_OUTER>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
      break;
    default:
      fat_launch<FAT_GRAD_OUTER>(st, Q, F, m, labels, P, Z, pred, stats_partial, nblocks, NT, stats_out, ticket);
      break;
  }
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ long geom_off(const BondGeom& g, int a, int s, int t, int b, int l) {
  return a * g.sa + s * g.ss + t * g.st + b * g.sb + l * g.sl;
}

__global__ void form_bond_kernel(const double* __restrict__ Wb, const double* __restrict__ Wb1, int m,
                                 BondGeom g, double* __restrict__ B) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = g.size();
  if (idx >= n) return;
  int l = (int)(idx % g.nl);
  long r = idx / g.nl;
  int b = (int)(r % g.mr);
  r /= g.mr;
  int t = (int)(r % 2);
  r /= 2;
  int s = (int)(r % 2);
  int a = (int)(r / 2);
  double acc = 0.0;
  const long as = (long)a * 2 + s;
  for (int k = 0; k < m; ++k) {
    double u = g.lab_b ? Wb[(as * m + k) * NL + l] : Wb[as * m + k];
    long i1 = ((long)k * 2 + t) * g.mr + b;
    double v = g.lab_b1 ? Wb1[i1 * NL + l] : Wb1[i1];
    acc = fma(u, v, acc);
  }
  B[geom_off(g, a, s, t, b, l)] = acc;
}
void form_bond(cudaStream_t st, const double* Wb, const double* Wb1, int m, BondGeom g, double* B) {
  long n = g.size();
  form_bond_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Wb, Wb1, m, g, B);
}

__global__ void bond_layout_kernel(const double* __restrict__ src, BondGeom g, double* __restrict__ dst,
                                   int to_host) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = g.size();
  if (idx >= n) return;
  int l = (int)(idx % g.nl);
  long r = idx / g.nl;
  int b = (int)(r % g.mr);
  r /= g.mr;
  int t = (int)(r % 2);
  r /= 2;
  int s = (int)(r % 2);
  int a = (int)(r / 2);
  long off = geom_off(g, a, s, t, b, l);
  if (to_host)
    dst[idx] = src[off];
  else
    dst[off] = src[idx];
}
void bond_to_host_layout(cudaStream_t st, const double* Bc, BondGeom g, double* Bh) {
  long n = g.size();
  bond_layout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Bc, g, Bh, 1);
}
void bond_from_host_layout(cudaStream_t st, const double* Bh, BondGeom g, double* Bc) {
  long n = g.size();
  bond_layout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Bh, g, Bc, 0);
}

__global__ void permute_site_kernel(const double* __restrict__ W, int ma, int mb, int nl, int right,
                                    double* __restrict__ Bm) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = (long)ma * 2 * mb * nl;
  if (idx >= n) return;
  int l = (int)(idx % nl);
  long r = idx / nl;
  int b = (int)(r % mb);
  r /= mb;
  int s = (int)(r % 2);
  int a = (int)(r / 2);
  double v = W[idx];
  if (!right)
    Bm[((long)a * 2 + s) * ((long)nl * mb) + (long)l * mb + b] = v;
  else
    Bm[((long)b * 2 + s) * ((long)nl * ma) + (long)l * ma + a] = v;
}
void permute_site(cudaStream_t st, const double* W, int ma, int mb, int nl, int right, double* Bm) {
  long n = (long)ma * 2 * mb * nl;
  permute_site_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, ma, mb, nl, right, Bm);
}

__global__ void fill_kernel(double* x, long n, double v) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = v;
}
void fill(cudaStream_t st, double* x, long n, double v) {
  if (n <= 0) return;
  long blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, v);
}

__global__ void axpby_kernel(long n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}
void axpby(cudaStream_t st, long n, double a, const double* x, double b, double* y) {
  if (n <= 0) return;
  axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a, x, b, y);
}

// y = (*a)*x + b*y with the scalar read from device memory: the CG step size a = |r|^2 / pAp
// (fixedL.cc:405) never visits the host
__global__ void axpby_dev_kernel(long n, const double* __restrict__ a_ptr, const double* __restrict__ x, double b,
                                 double* __restrict__ y) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const double a = *a_ptr;
  if (i < n) y[i] = fma(a, x[i], b * y[i]);
}
void axpby_dev(cudaStream_t st, long n, const double* a_ptr, const double* x, double b, double* y) {
  if (n <= 0) return;
  axpby_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a_ptr, x, b, y);
}
// ---- device-side scalars of the CG recurrence (fixedL.cc:388-443): cgs = [0] |r|^2, [1] beta, [2] converged,
// [3] passes recorded, [4] step a, [8..15] C/NT per pass, [16..23] |r| per pass; `tail` = the 16 statistics behind G
// (cost per label in 0..9, |G|^2 in 12)
__global__ void cg_begin_kernel(const double* __restrict__ tail, double* __restrict__ cgs) {
  if (threadIdx.x == 0) {
    cgs[0] = tail[12];
    cgs[1] = 0.0;
    cgs[2] = 0.0;
    cgs[3] = 0.0;
    cgs[4] = 0.0;
  }
}
void cg_begin(cudaStream_t st, const double* tail, double* cgs) { cg_begin_kernel<<<1, 32, 0, st>>>(tail, cgs); }

__global__ void cg_step_kernel(double* __restrict__ cgs, const double* __restrict__ pAp, double lambda,
                               const double* __restrict__ pp) {
  if (threadIdx.x == 0) {
    double d = pAp[0];
    if (lambda != 0.0) d += lambda * pp[0];
    cgs[4] = (cgs[2] != 0.0) ? 0.0 : cgs[0] / d;      // a = |r|^2 / pAp (405); frozen once |r| < cconv
  }
}
void cg_step(cudaStream_t st, double* cgs, const double* pAp, double lambda, const double* pp) {
  cg_step_kernel<<<1, 32, 0, st>>>(cgs, pAp, lambda, pp);
}

__global__ void cg_after_grad_kernel(const double* __restrict__ tail, double* __restrict__ cgs, double lambda,
                                     const double* __restrict__ bb, double NTg, double cconv) {
  if (threadIdx.x == 0 && cgs[2] == 0.0) {
    const double nrr = tail[12];
    cgs[1] = nrr / cgs[0];                            // beta = (|nr|/|r|)^2 (423)
    cgs[0] = nrr;                                     // r = nr (424)
    double C = 0.0;
    for (int l = 0; l < NL; ++l) C += tail[l];        // 427
    if (lambda != 0.0) C += lambda * bb[0];           // 428
    const int nd = (int)cgs[3];
    if (nd < 8) {
      cgs[8 + nd] = C / NTg;                          // 429
      cgs[16 + nd] = sqrt(nrr);
    }
    cgs[3] = (double)(nd + 1);
    if (sqrt(nrr) < cconv) cgs[2] = 1.0;              // 432-436: no further update of p or B
  }
}
void cg_after_grad(cudaStream_t st, const double* tail, double* cgs, double lambda, const double* bb, double NTg,
                   double cconv) {
  cg_after_grad_kernel<<<1, 32, 0, st>>>(tail, cgs, lambda, bb, NTg, cconv);
}

// y = x + (*b) * y   (p = r + beta p, fixedL.cc:442)
__global__ void xpby_dev_kernel(long n, const double* __restrict__ x, const double* __restrict__ b_ptr,
                                double* __restrict__ y) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const double b = *b_ptr;
  if (i < n) y[i] = fma(b, y[i], x[i]);
}
void xpby_dev(cudaStream_t st, long n, const double* x, const double* b_ptr, double* y) {
  if (n <= 0) return;
  xpby_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, x, b_ptr, y);
}

constexpr int DOT_BLOCKS = 128;
__global__ void dot_kernel(long n, const double* __restrict__ x, const double* __restrict__ y,
                           double* __restrict__ scratch) {
  __shared__ double red[256];
  double s = 0.0;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)DOT_BLOCKS * 256) s = fma(x[i], y[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) scratch[blockIdx.x] = red[0];
}
__global__ void dot_final_kernel(const double* __restrict__ scratch, double* __restrict__ out) {
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < DOT_BLOCKS; ++b) s += scratch[b];
    out[0] = s;
  }
}
void dot(cudaStream_t st, long n, const double* x, const double* y, double* scratch, double* out) {
  dot_kernel<<<DOT_BLOCKS, 256, 0, st>>>(n, x, y, scratch);
  dot_final_kernel<<<1, 32, 0, st>>>(scratch, out);
}

}  // namespace tnml
