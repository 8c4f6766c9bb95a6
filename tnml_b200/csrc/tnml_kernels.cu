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

namespace tnml {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int BK = 16;
constexpr int NTHR = 256;

template <int NB>
__device__ __forceinline__ void tile_fma(const double* __restrict__ A, const double* __restrict__ B,
                                         double (&acc)[NB][8][4], int tx) {
  // A -> As[buf] + ty*8 ; B -> Bs[buf][0]
#pragma unroll
  for (int k = 0; k < BK; ++k) {
    double a[8];
    const double2* ap = reinterpret_cast<const double2*>(A + k * BM);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double2 v = ap[i];
      a[2 * i] = v.x;
      a[2 * i + 1] = v.y;
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const double* bp = B + (q * BK + k) * BN;
      double2 b01 = *reinterpret_cast<const double2*>(bp + tx * 2);
      double2 b23 = *reinterpret_cast<const double2*>(bp + 32 + tx * 2);
      double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[q][i][j] = fma(a[i], b[j], acc[q][i][j]);
    }
  }
}

__device__ __forceinline__ int tile_col(int tx, int jj) { return (jj < 2) ? (tx * 2 + jj) : (32 + tx * 2 + (jj - 2)); }

// ---------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(NTHR)
krgemm_kernel(const double* __restrict__ In, long ldin, int ma, const double* __restrict__ f1, int div,
              const double* __restrict__ Bm0, const double* __restrict__ Bm1, long ldb, int J,
              const double* __restrict__ f2, double* __restrict__ Out, long ldout, long rows) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                 // [2][BK][BM]
  double* Bs = smem + 2 * BK * BM;   // [2][NB][BK][BN]
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const long row0 = (long)blockIdx.x * BM;
  const int j0 = blockIdx.y * BN;

  const int lrow = t >> 1, lah = (t & 1) * 4;
  const long grow = row0 + lrow;
  const bool rok = grow < rows;
  double fa0 = 0.0, fa1 = 0.0;
  if (rok) {
    long img = grow / div;
    fa0 = f1[img * 2];
    fa1 = f1[img * 2 + 1];
  }
  const int bk = t >> 4, bc = (t & 15) * 4;

  double acc[NB][8][4];
#pragma unroll
  for (int q = 0; q < NB; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][i][j] = 0.0;

  double ra[4];
  double rb[NB][4];
  const int nk = (ma + 7) / 8;

  auto gload = [&](int kt) {
    const int a0 = kt * 8 + lah;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int a = a0 + i;
      ra[i] = (rok && a < ma) ? In[grow * ldin + a] : 0.0;
    }
    const int k2 = kt * 16 + bk;
    const bool kok = k2 < 2 * ma;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const double* Bq = (q == 0) ? Bm0 : Bm1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int j = j0 + bc + i;
        rb[q][i] = (kok && j < J) ? Bq[(long)k2 * ldb + j] : 0.0;
      }
    }
  };
  auto sstore = [&](int buf) {
    double* A = As + buf * BK * BM;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = (lah + i) * 2;
      A[k * BM + lrow] = ra[i] * fa0;
      A[(k + 1) * BM + lrow] = ra[i] * fa1;
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      double* B = Bs + ((buf * NB + q) * BK + bk) * BN + bc;
#pragma unroll
      for (int i = 0; i < 4; ++i) B[i] = rb[q][i];
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
    tile_fma<NB>(As + buf * BK * BM + ty * 8, Bs + buf * NB * BK * BN, acc, tx);
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long r = row0 + ty * 8 + i;
    if (r >= rows) continue;
    double w0 = 1.0, w1 = 0.0;
    if (NB == 2) {
      long img = r / div;
      w0 = f2[img * 2];
      w1 = f2[img * 2 + 1];
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int j = j0 + tile_col(tx, jj);
      if (j < J) {
        double v = (NB == 2) ? fma(w1, acc[NB - 1][i][jj], w0 * acc[0][i][jj]) : acc[0][i][jj];
        Out[r * ldout + j] = v;
      }
    }
  }
}

void krgemm(cudaStream_t st, int NB, const double* In, long ldin, int ma, const double* f1, int div,
            const double* Bm0, const double* Bm1, long ldb, int J, const double* f2, double* Out,
            long ldout, long rows) {
  if (rows <= 0 || J <= 0) return;
  dim3 grid((unsigned)((rows + BM - 1) / BM), (unsigned)((J + BN - 1) / BN));
  size_t sh = (size_t)(2 * BK * BM + 2 * NB * BK * BN) * sizeof(double);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(krgemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(krgemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    attr = true;
  }
  if (NB == 1)
    krgemm_kernel<1><<<grid, NTHR, sh, st>>>(In, ldin, ma, f1, div, Bm0, Bm1, ldb, J, f2, Out, ldout, rows);
  else
    krgemm_kernel<2><<<grid, NTHR, sh, st>>>(In, ldin, ma, f1, div, Bm0, Bm1, ldb, J, f2, Out, ldout, rows);
}

// ---------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(NTHR)
krgram_kernel(const double* __restrict__ In, long ldin, int ma, const double* __restrict__ f1,
              const double* __restrict__ f2, const double* __restrict__ Z, long ldz, int J,
              double* __restrict__ Gpart, long rows, long rows_per_split) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                 // [2][BK][BM]   BM index = (a_local*2+s)
  double* Bs = smem + 2 * BK * BM;   // [2][NB][BK][BN]
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int a0 = blockIdx.x * (BM / 2);
  const int j0 = blockIdx.y * BN;
  const long rbeg = (long)blockIdx.z * rows_per_split;
  const long rend = (rbeg + rows_per_split < rows) ? (rbeg + rows_per_split) : rows;
  const int lrow = t >> 4, c4 = (t & 15) * 4;

  double acc[NB][8][4];
#pragma unroll
  for (int q = 0; q < NB; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][i][j] = 0.0;

  double ra[4], rz[4], rf[4];  // rf: f1_0, f1_1, f2_0, f2_1
  const long nrow = (rend > rbeg) ? (rend - rbeg) : 0;
  const int nk = (int)((nrow + BK - 1) / BK);

  auto gload = [&](int kt) {
    const long r = rbeg + (long)kt * BK + lrow;
    const bool ok = r < rend;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int a = a0 + c4 + i;
      ra[i] = (ok && a < ma) ? In[r * ldin + a] : 0.0;
      int j = j0 + c4 + i;
      rz[i] = (ok && j < J) ? Z[r * ldz + j] : 0.0;
    }
    rf[0] = ok ? f1[r * 2] : 0.0;
    rf[1] = ok ? f1[r * 2 + 1] : 0.0;
    if (NB == 2) {
      rf[2] = ok ? f2[r * 2] : 0.0;
      rf[3] = ok ? f2[r * 2 + 1] : 0.0;
    } else {
      rf[2] = 1.0;
      rf[3] = 0.0;
    }
  };
  auto sstore = [&](int buf) {
    double2* A = reinterpret_cast<double2*>(As + (buf * BK + lrow) * BM + c4 * 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) A[i] = make_double2(ra[i] * rf[0], ra[i] * rf[1]);
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      double2* B = reinterpret_cast<double2*>(Bs + ((buf * NB + q) * BK + lrow) * BN + c4);
      B[0] = make_double2(rz[0] * rf[2 + q], rz[1] * rf[2 + q]);
      B[1] = make_double2(rz[2] * rf[2 + q], rz[3] * rf[2 + q]);
    }
  };

  if (nk > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
    tile_fma<NB>(As + buf * BK * BM + ty * 8, Bs + buf * NB * BK * BN, acc, tx);
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  const long M2 = 2L * ma;
  double* Gp = Gpart + (long)blockIdx.z * (M2 * NB * J);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long m2 = (long)a0 * 2 + ty * 8 + i;
    if (m2 >= M2) continue;
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        int j = j0 + tile_col(tx, jj);
        if (j < J) Gp[(m2 * NB + q) * J + j] = acc[q][i][jj];
      }
  }
}

int krgram_splits(int ma, int J, int NB, long rows, int num_sm) {
  int mt = (2 * ma + BM - 1) / BM, nt = (J + BN - 1) / BN;
  long tiles = (long)mt * nt;
  int s = (int)((2L * num_sm + tiles - 1) / tiles);
  long maxs = (rows + 8 * BK - 1) / (8 * BK);  // at least 128 rows per split
  if (s > maxs) s = (int)maxs;
  if (s < 1) s = 1;
  if (s > 65535) s = 65535;
  return s;
}

void krgram(cudaStream_t st, int NB, const double* In, long ldin, int ma, const double* f1,
            const double* f2, const double* Z, long ldz, int J, double* Gpart, long rows, int nsplit) {
  dim3 grid((unsigned)((2 * ma + BM - 1) / BM), (unsigned)((J + BN - 1) / BN), (unsigned)nsplit);
  long rps = (rows + nsplit - 1) / nsplit;
  rps = ((rps + BK - 1) / BK) * BK;
  size_t sh = (size_t)(2 * BK * BM + 2 * NB * BK * BN) * sizeof(double);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(krgram_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(krgram_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    attr = true;
  }
  if (NB == 1)
    krgram_kernel<1><<<grid, NTHR, sh, st>>>(In, ldin, ma, f1, f2, Z, ldz, J, Gpart, rows, rps);
  else
    krgram_kernel<2><<<grid, NTHR, sh, st>>>(In, ldin, ma, f1, f2, Z, ldz, J, Gpart, rows, rps);
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

template <int MODE>
__global__ void __launch_bounds__(256)
fat_kernel_t(const double* __restrict__ Q, const double* __restrict__ F, int m,
             const int32_t* __restrict__ labels, double* __restrict__ P, double* __restrict__ Z,
             int32_t* __restrict__ pred, double* __restrict__ stats_partial, long NT) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long gw = (long)blockIdx.x * 8 + warp;
  const long nw = (long)gridDim.x * 8;
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
    for (int f = lane; f < m; f += 32) {
      double qv = q[f];
#pragma unroll
      for (int l = 0; l < NL; ++l) pl[l] = fma(qv, Fn[(long)l * m + f], pl[l]);
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) pl[l] = warp_allsum(pl[l]);
    if (P != nullptr) {
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
      if (MODE == FAT_GRAD) {
        double* z = Z + n * m;
        for (int f = lane; f < m; f += 32) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < NL; ++l) s = fma(dp[l], Fn[(long)l * m + f], s);
          z[f] = s;
        }
      } else if (MODE == FAT_GRAD_OUTER) {
        double* z = Z + n * (long)NL * m;
        for (int f = lane; f < m; f += 32) {
          double qv = q[f];
#pragma unroll
          for (int l = 0; l < NL; ++l) z[(long)l * m + f] = dp[l] * qv;
        }
      }
    }
  }
  __shared__ double red[8][12];
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
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    stats_partial[(long)blockIdx.x * 16 + threadIdx.x] = s;
  }
}

int fat_blocks(int num_sm) { return num_sm * 8; }

void fat_kernel(cudaStream_t st, int mode, const double* Q, const double* F, int m, const int32_t* labels,
                double* P, double* Z, int32_t* pred, double* stats_partial, int nblocks, long NT) {
  switch (mode) {
    case FAT_GRAD:
      fat_kernel_t<FAT_GRAD><<<nblocks, 256, 0, st>>>(Q, F, m, labels, P, Z, pred, stats_partial, NT);
      break;
    case FAT_PAP:
      fat_kernel_t<FAT_PAP><<<nblocks, 256, 0, st>>>(Q, F, m, labels, P, Z, pred, stats_partial, NT);
      break;
    case FAT_COST:
      fat_kernel_t<FAT_COST><<<nblocks, 256, 0, st>>>(Q, F, m, labels, P, Z, pred, stats_partial, NT);
      break;
    default:
      fat_kernel_t<FAT_GRAD_OUTER><<<nblocks, 256, 0, st>>>(Q, F, m, labels, P, Z, pred, stats_partial, NT);
      break;
  }
}

__global__ void reduce_stats_kernel(const double* __restrict__ sp, int nblocks, double* __restrict__ stats) {
  // 16 warps, one per statistic: lanes stride over the per-block partials in a fixed order
  const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += sp[(long)b * 16 + i];
  s = warp_allsum(s);
  if (lane == 0) stats[i] = s;
}
void reduce_stats(cudaStream_t st, const double* sp, int nblocks, double* stats) {
  reduce_stats_kernel<<<1, 512, 0, st>>>(sp, nblocks, stats);
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
