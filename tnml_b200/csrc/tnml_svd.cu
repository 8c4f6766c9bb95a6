// tnml_svd.cu -- truncated SVD of the bond tensor on the device, float64.
// Replaces ITensor's svd(B,U,S,V,{Cutoff,Maxm,Minm}) + `W.Aref(c+dc) *= S`
// (fixedL.cc:519-521).
//
// Method: block one-sided (Hestenes) Jacobi.  The bond matrix is held in its
// tall orientation X[small][big] (column-major, columns = the smaller index
// set, at most 2*maxm of them); pairs of columns are rotated until mutually
// orthogonal, the same rotations accumulate in J.  Columns are grouped in
// blocks of 16: one "diagonal" launch orthogonalises the pairs inside every
// block, then a round-robin tournament over block pairs (one CTA per block
// pair, one warp per column pair, 16 inner rounds) covers the cross pairs.
// Dot products are warp-shuffle reductions.  One-sided Jacobi is accurate for
// small singular values in the relative sense, which matters because Minm
// (fixedL.cc:593) forces the trailing vectors to be kept.
#include "tnml_kernels.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace tnml {

constexpr int JW = 16;  // block width (columns)

// process-wide variant switches (tnml_set_option "svd_cluster" / "svd_cross" / "svd_precond"; -1 = take
// the TNML_SVD_* environment variable or the default).  Testing and A/B timing only.
static int g_svd_cluster = -1, g_svd_cross = -1, g_svd_precond = -1;
void svd_set_variant(const char* what, int v) {
  if (what[4] == 'c' && what[5] == 'l') g_svd_cluster = v;        // svd_cluster
  else if (what[4] == 'c' && what[5] == 'r') g_svd_cross = v;     // svd_cross
  else g_svd_precond = v;                                          // svd_precond: 1 one QR, 3 sort + two QRs
}

__device__ long long* g_qr_dbg = nullptr;   // optional [ns][4] clock64 stamps (TNML_QR_DEBUG)


__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// rotate columns ci, cj of X (length nb) and of J (length ns); returns |cos angle|
__device__ __forceinline__ double rotate_pair(double* __restrict__ X, double* __restrict__ J, int nb, int ns,
                                              int ci, int cj, double tol, int lane) {
  if (ci >= ns || cj >= ns) return 0.0;
  double* xi = X + (long)ci * nb;
  double* xj = X + (long)cj * nb;
  double a = 0.0, b = 0.0, g = 0.0;
  for (int r = lane; r < nb; r += 32) {
    double u = xi[r], v = xj[r];
    a = fma(u, u, a);
    b = fma(v, v, b);
    g = fma(u, v, g);
  }
  a = wsum(a);
  b = wsum(b);
  g = wsum(g);
  double den = sqrt(a) * sqrt(b);
  if (!(den > 0.0)) return 0.0;
  double off = fabs(g) / den;
  if (off <= tol) return off;
  double zeta = (b - a) / (2.0 * g);
  double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  double c = 1.0 / sqrt(1.0 + t * t);
  double s = c * t;
  for (int r = lane; r < nb; r += 32) {
    double u = xi[r], v = xj[r];
    xi[r] = c * u - s * v;
    xj[r] = s * u + c * v;
  }
  double* ji = J + (long)ci * ns;
  double* jj = J + (long)cj * ns;
  for (int r = lane; r < ns; r += 32) {
    double u = ji[r], v = jj[r];
    ji[r] = c * u - s * v;
    jj[r] = s * u + c * v;
  }
  return off;
}

__device__ __forceinline__ void atomic_max_pos(double* addr, double v) {
  // v >= 0: IEEE order == integer order
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// circle-method round robin for n (even) players: pair k of round r
__device__ __forceinline__ void rr_pair(int n, int r, int k, int& p, int& q) {
  if (k == 0) {
    p = n - 1;
    q = r;
  } else {
    p = (r + k) % (n - 1);
    q = (r - k + (n - 1)) % (n - 1);
  }
}

// pairs inside each block of JW columns: 15 rounds x 8 pairs, 8 warps
__global__ void __launch_bounds__(256)
jacobi_diag_kernel(double* __restrict__ X, double* __restrict__ J, int nb, int ns, double tol,
                   double* __restrict__ info, const int* __restrict__ flags) {
  if (flags[0]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * JW;
  double mo = 0.0;
  for (int r = 0; r < JW - 1; ++r) {
    int p, q;
    rr_pair(JW, r, warp, p, q);
    double off = rotate_pair(X, J, nb, ns, c0 + p, c0 + q, tol, lane);
    mo = fmax(mo, off);
    __syncthreads();
  }
  if (lane == 0 && mo > 0.0) atomic_max_pos(info, mo);
}

// cross pairs of block pair (P,Q) chosen by outer round R: 16 rounds x 16 pairs, 16 warps
__global__ void __launch_bounds__(512)
jacobi_offdiag_kernel(double* __restrict__ X, double* __restrict__ J, int nb, int ns, int nblk_e, int R,
                      double tol, double* __restrict__ info, const int* __restrict__ flags) {
  if (flags[0]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int P, Q;
  rr_pair(nblk_e, R, blockIdx.x, P, Q);
  double mo = 0.0;
  for (int r = 0; r < JW; ++r) {
    int ci = P * JW + warp;
    int cj = Q * JW + ((warp + r) & (JW - 1));
    double off = rotate_pair(X, J, nb, ns, ci, cj, tol, lane);
    mo = fmax(mo, off);
    __syncthreads();
  }
  if (lane == 0 && mo > 0.0) atomic_max_pos(info, mo);
}


// ---- shared-memory resident variants (used when 2*W columns of X and J fit) ----
// rotate columns (already in shared memory) xi/xj (length nb) and ji/jj (length ns)
// xi/xj point at staged columns laid out [X part (nb) | J part (ns)], ld = nb+ns.
// The rotation is derived with two rsqrt (no divide / sqrt chain): with
// d = b-a, h = 2g:  cos2t = |d|/r, sin2t = sign(d) h/r, r = sqrt(d^2+h^2),
// c = sqrt((1+cos2t)/2), s = sin2t/(2c)  (|t| <= pi/4, same rotation as the
// classical tan formula).  Returns g^2/(a b) scaled test value (0 if converged).
__device__ __forceinline__ double rotate_pair_smem(double* __restrict__ xi, double* __restrict__ xj, int nb,
                                                   int ld, double tol2, int lane) {
  double a = 0.0, b = 0.0, g = 0.0;
  for (int r = lane; r < nb; r += 32) {
    double u = xi[r], v = xj[r];
    a = fma(u, u, a);
    b = fma(v, v, b);
    g = fma(u, v, g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    g += __shfl_xor_sync(0xffffffffu, g, o);
  }
  const double ab = a * b;
  const double gg = g * g;
  if (!(ab > 0.0) || gg <= tol2 * ab) return (ab > 0.0) ? gg / ab : 0.0;
  const double d = b - a, h = 2.0 * g;
  const double rinv = rsqrt(fma(d, d, h * h));
  const double c2 = fma(0.5 * fabs(d), rinv, 0.5);   // cos^2 t
  const double rc = rsqrt(c2);
  const double c = c2 * rc;
  const double s = copysign(0.5, d) * h * rinv * rc;
  for (int r = lane; r < ld; r += 32) {
    double u = xi[r], v = xj[r];
    xi[r] = fma(c, u, -s * v);
    xj[r] = fma(s, u, c * v);
  }
  return gg / ab;
}

// Fully unrolled variant for rows, ns <= 32*R: the two A columns stay in registers between
// the dot products and the update, every shared-memory access of a phase is issued
// back to back (the loop version above is latency bound: ~26 cycles per instruction).
template <int R>
__device__ __forceinline__ double rotate_pair_regs(double* __restrict__ xi, double* __restrict__ xj, int nb,
                                                   int ns, double tol2, int lane) {
  double u[R], v[R];
  double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int r = lane + 32 * i;
    u[i] = (r < nb) ? xi[r] : 0.0;
    v[i] = (r < nb) ? xj[r] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    a = fma(u[i], u[i], a);
    b = fma(v[i], v[i], b);
    g = fma(u[i], v[i], g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    g += __shfl_xor_sync(0xffffffffu, g, o);
  }
  const double ab = a * b;
  const double gg = g * g;
  if (!(ab > 0.0)) return 0.0;
  if (gg <= tol2 * ab) return tol2 * 0.5;   // converged pair: report "below threshold"
  const double d = b - a, h = 2.0 * g;
  const double rinv = rsqrt(fma(d, d, h * h));
  const double c2 = fma(0.5 * fabs(d), rinv, 0.5);
  const double rc = rsqrt(c2);
  const double c = c2 * rc;
  const double s = copysign(0.5, d) * h * rinv * rc;
  double* ji = xi + nb;
  double* jj = xj + nb;
  double p[R], q[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int r = lane + 32 * i;
    p[i] = (r < ns) ? ji[r] : 0.0;
    q[i] = (r < ns) ? jj[r] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int r = lane + 32 * i;
    if (r < nb) {
      xi[r] = fma(c, u[i], -s * v[i]);
      xj[r] = fma(s, u[i], c * v[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int r = lane + 32 * i;
    if (r < ns) {
      ji[r] = fma(c, p[i], -s * q[i]);
      jj[r] = fma(s, p[i], c * q[i]);
    }
  }
  return gg / ab;
}

// Stage columns [c0, c0+W) (slot 0..W-1) and [c1, c1+W) (slot W..2W-1) of X and J.
template <int W>
__device__ __forceinline__ void stage_cols(double* __restrict__ sm, double* __restrict__ X, double* __restrict__ J,
                                           int nb, int ns, int c0, int c1, bool store, int ncols) {
  // flat index over (slot k, row r) so that every thread has many independent
  // global accesses in flight (the staged set is re-read from L2 every launch)
  const int ld = nb + ns;
  if (((nb | ns) & 1) == 0) {   // 16-byte path: every column start is 16-byte aligned
    const int ld2 = ld >> 1, nb2 = nb >> 1, ns2 = ns >> 1;
    const int total = ncols * ld2;
    double2* sm2 = reinterpret_cast<double2*>(sm);
#pragma unroll 4
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int k = i / ld2, r = i - k * ld2;
      const int c = (k < W) ? (c0 + k) : (c1 + k - W);
      if (c >= ns) continue;
      double2* gp = (r < nb2) ? (reinterpret_cast<double2*>(X + (long)c * nb) + r)
                              : (reinterpret_cast<double2*>(J + (long)c * ns) + (r - nb2));
      if (!store)
        sm2[i] = __ldcg(gp);
      else
        __stcg(gp, sm2[i]);
    }
    (void)ns2;
    return;
  }
  const int total = ncols * ld;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int k = i / ld, r = i - k * ld;
    const int c = (k < W) ? (c0 + k) : (c1 + k - W);
    if (c >= ns) continue;
    double* gp = (r < nb) ? (X + (long)c * nb + r) : (J + (long)c * ns + (r - nb));
    if (!store)
      sm[i] = *gp;
    else
      *gp = sm[i];
  }
}

template <int W>
__global__ void __launch_bounds__(32 * W)
jacobi_offdiag_smem_kernel(double* __restrict__ X, double* __restrict__ J, int nb, int ns, int nblk_e, int R,
                           double tol2, double* __restrict__ info, const int* __restrict__ flags) {
  if (flags[0]) return;
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int P, Q;
  rr_pair(nblk_e, R, blockIdx.x, P, Q);
  if (P * W >= ns || Q * W >= ns) return;  // dummy block: nothing to do
  const int ld = nb + ns;
  stage_cols<W>(sm, X, J, nb, ns, P * W, Q * W, false, 2 * W);
  __syncthreads();
  double mo = 0.0;
  for (int r = 0; r < W; ++r) {
    const int ki = warp, kj = W + ((warp + r) & (W - 1));
    if (P * W + ki < ns && Q * W + (kj - W) < ns) {
      double* si = sm + (long)ki * ld;
      double* sj = sm + (long)kj * ld;
      mo = fmax(mo, (nb <= 256 && ns <= 256) ? rotate_pair_regs<8>(si, sj, nb, ns, tol2, lane)
                                             : rotate_pair_smem(si, sj, nb, ld, tol2, lane));
    }
    __syncthreads();
  }
  stage_cols<W>(sm, X, J, nb, ns, P * W, Q * W, true, 2 * W);
  if (lane == 0 && mo > 0.0) atomic_max_pos(info, mo);
}

template <int W>
__global__ void __launch_bounds__(16 * W)
jacobi_diag_smem_kernel(double* __restrict__ X, double* __restrict__ J, int nb, int ns, double tol2,
                        double* __restrict__ info, const int* __restrict__ flags) {
  if (flags[0]) return;
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * W;
  const int ld = nb + ns;
  stage_cols<W>(sm, X, J, nb, ns, c0, 0, false, W);
  __syncthreads();
  double mo = 0.0;
  for (int r = 0; r < W - 1; ++r) {
    int p, q;
    rr_pair(W, r, warp, p, q);
    if (c0 + p < ns && c0 + q < ns) {
      double* si = sm + (long)p * ld;
      double* sj = sm + (long)q * ld;
      mo = fmax(mo, (nb <= 256 && ns <= 256) ? rotate_pair_regs<8>(si, sj, nb, ns, tol2, lane)
                                             : rotate_pair_smem(si, sj, nb, ld, tol2, lane));
    }
    __syncthreads();
  }
  stage_cols<W>(sm, X, J, nb, ns, c0, 0, true, W);
  if (lane == 0 && mo > 0.0) atomic_max_pos(info, mo);
}


// ---------------------------------------------------------------------------
// Gram-based block Jacobi: one CTA per block pair (32 staged columns of [A | J]).
//   1. G = A_blk^T A_blk (32 x 32) with FP64 tensor-core MMAs (DMMA.8x8x4)
//   2. NR rounds of 16 disjoint plane rotations applied to G itself (G <- R^T G R, ping-pong
//      buffers, one barrier per round) while the product of the rotations accumulates in RA;
//      no dot products / warp shuffles inside the round chain
//   3. [A | J]_blk <- [A | J]_blk * RA, again with DMMA, written back
// Mathematically the same rotations as the column-wise kernel above (which is bound by
// shared-memory bandwidth: every round re-reads and re-writes all 32 staged columns).
__device__ __forceinline__ void dmma884s(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Branch-free (two of them are evaluated per thread and must interleave).  rot = cos^2 of the
// angle between the two columns when a rotation is applied, 0 when the pair is already
// orthogonal to tolerance.
__device__ __forceinline__ void plane_rot(double a, double b, double g, double tol2, double& c, double& s,
                                          double& rot) {
  const double ab = a * b, gg = g * g;
  const bool doit = (ab > 0.0) && (gg > tol2 * ab);
  const double d = b - a, h = 2.0 * g;
  const double r2 = doit ? fma(d, d, h * h) : 1.0;
  const double rinv = rsqrt(r2);
  const double c2 = fma(0.5 * fabs(d), rinv, 0.5);
  const double rc = rsqrt(c2);
  const double cc = c2 * rc;
  const double ss = copysign(0.5, d) * h * rinv * rc;
  c = doit ? cc : 1.0;
  s = doit ? ss : 0.0;
  rot = doit ? gg / ab : 0.0;   // cos^2 of the angle that was rotated away (0: no rotation)
}

constexpr int GLD = 36;   // leading dimension of the 32 x 32 Gram / rotation matrices in smem

template <int W>   // block width: 2W staged columns, all (2W choose 2) pairs in 2W-1 rounds
__global__ void __launch_bounds__(32 * W)
jacobi_gram_kernel(double* __restrict__ A, double* __restrict__ Jm, int rows, int ns, int ld, int nblk_e, int R,
                   double tol2, double* __restrict__ info, const int* __restrict__ flags) {
  constexpr int NC = 2 * W;         // staged columns
  constexpr int NR = NC - 1;        // rounds
  constexpr int NTH = 32 * W;       // threads (512 / 256)
  constexpr int CP = NTH / 256;     // column parities of the staging map
  constexpr int NWARP = NTH / 32;
  constexpr int TG = NC / 8;        // 8x8 tiles per side of the Gram matrix
  if (flags[0]) return;
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                       // [NC][ld]
  double* G0 = sm + (long)NC * ld;      // [NC][GLD]
  double* G1 = G0 + NC * GLD;
  double* RA = G1 + NC * GLD;
  double* CS = RA + NC * GLD;                                        // [2][W][2] (double buffered)
  unsigned short* PQ = reinterpret_cast<unsigned short*>(CS + 4 * W);   // [NR][W]  p | q << 8
  unsigned char* POS = reinterpret_cast<unsigned char*>(PQ + NR * W);    // [NR][NC] pair*2 + (0 first, 1 second)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  int P, Q;
  rr_pair(nblk_e, R, blockIdx.x, P, Q);
  const int c0 = P * W, c1 = Q * W;
  if (c0 >= ns && c1 >= ns) return;
  const int tot = rows + ns;
  long long tstamp[6];
  tstamp[0] = clock64();
  // ---- stage (invalid columns and the padding rows are zero).  Thread = (row pair rc2,
  //      column parity): 16 independent 16-byte loads in flight per thread, no index division.
  const int rc2 = tid & 255, cpar = tid >> 8;
  const int ld2 = ld >> 1, rows2 = rows >> 1, tot2 = tot >> 1;   // rows, ns are even (2m or 2m*NL)
  for (int rr2 = rc2; rr2 < ld2; rr2 += 256) {   // more than 512 staged rows (m > 128): several passes
    double2 v[NC / CP];
#pragma unroll
    for (int it = 0; it < NC / CP; ++it) {
      const int k = cpar + CP * it;
      const int c = (k < W) ? (c0 + k) : (c1 + k - W);
      v[it] = make_double2(0.0, 0.0);
      if (c < ns && rr2 < tot2)
        v[it] = (rr2 < rows2) ? __ldcg(reinterpret_cast<const double2*>(A + (long)c * rows) + rr2)
                              : __ldcg(reinterpret_cast<const double2*>(Jm + (long)c * ns) + (rr2 - rows2));
    }
#pragma unroll
    for (int it = 0; it < NC / CP; ++it) {
      const int k = cpar + CP * it;
      reinterpret_cast<double2*>(S + (long)k * ld)[rr2] = v[it];
    }
  }
  for (int i = tid; i < NC * GLD; i += NTH) RA[i] = ((i / GLD) == (i % GLD)) ? 1.0 : 0.0;
  for (int i = tid; i < NR * W; i += NTH) {
    const int rd = i / W, k = i - rd * W;
    int p, q;
    rr_pair(NC, rd, k, p, q);
    PQ[i] = (unsigned short)(p | (q << 8));
    POS[rd * NC + p] = (unsigned char)(2 * k);
    POS[rd * NC + q] = (unsigned char)(2 * k + 1);
  }
  __syncthreads();
  tstamp[1] = clock64();
  // ---- Gram of the A part with DMMA: 8 x 8 tiles.  The accumulation over the rows is a chain of
  //      dependent DMMAs (latency ~40 cycles each), so the rows are split over KS warps per tile
  //      and over 4 independent accumulators per warp; partial sums meet in G0 / G1.
  constexpr int KS = (NWARP >= 2 * TG * TG) ? 2 : 1;
  if (warp < KS * TG * TG) {
    const int tile = warp % (TG * TG), kh = warp / (TG * TG);
    const int ti = tile / TG, tj = tile - ti * TG;
    const int nk4 = (rows + 3) >> 2;                       // k-steps of 4 rows
    const int kbeg = (nk4 * kh) / KS, kend = (nk4 * (kh + 1)) / KS;
    const double* pa = S + (long)(ti * 8 + g) * ld + t;
    const double* pb = S + (long)(tj * 8 + g) * ld + t;
    double d[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) d[u][0] = d[u][1] = 0.0;
    int k = kbeg;
    for (; k + 4 <= kend; k += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r0 = (k + u) * 4;
        const bool ok = (r0 + t) < rows;
        dmma884s(d[u][0], d[u][1], ok ? pa[r0] : 0.0, ok ? pb[r0] : 0.0);
      }
    }
    for (; k < kend; ++k) {
      const int r0 = k * 4;
      const bool ok = (r0 + t) < rows;
      dmma884s(d[0][0], d[0][1], ok ? pa[r0] : 0.0, ok ? pb[r0] : 0.0);
    }
    double* Gd = kh ? G1 : G0;
    Gd[(ti * 8 + g) * GLD + tj * 8 + 2 * t] = (d[0][0] + d[1][0]) + (d[2][0] + d[3][0]);
    Gd[(ti * 8 + g) * GLD + tj * 8 + 2 * t + 1] = (d[0][1] + d[1][1]) + (d[2][1] + d[3][1]);
  }
  __syncthreads();
  if (KS == 2) {
    for (int i = tid; i < NC * NC; i += NTH) {
      const int r = i / NC, c = i - r * NC;
      G0[r * GLD + c] += G1[r * GLD + c];
    }
    __syncthreads();
  }
  tstamp[2] = clock64();
  // ---- rotation rounds on the Gram matrix.  Per round: W*W threads update one 2x2 block of
  //      G each (ping-pong), NC*W/2 threads update RA, and W "look-ahead" threads compute the
  //      rotations of the NEXT round from privately recomputed entries of the updated G, so that
  //      the serial chain (load -> two rsqrt -> store) overlaps the updates: one barrier per round.
  double* cur = G0;
  double* nxt = G1;
  double mo = 0.0;
  if (tid < W) {   // rotations of round 0
    const int pq = PQ[tid];
    const int p = pq & 0xff, q = pq >> 8;
    double c, sn, rot;
    plane_rot(cur[p * GLD + p], cur[q * GLD + q], cur[p * GLD + q], tol2, c, sn, rot);
    CS[2 * tid] = c;
    CS[2 * tid + 1] = sn;
    mo = fmax(mo, rot);
  }
  __syncthreads();
  constexpr int RE = (W == 16) ? 4 : 2;          // pairs per RA-update thread
  constexpr int T_BLK = W * W, T_RA = NC * W / RE;
  static_assert(T_BLK + T_RA + W <= NTH, "thread budget");
  for (int rd = 0; rd < NR; ++rd) {
    const double* cs = CS + (rd & 1) * 2 * W;
    double* csn = CS + ((rd + 1) & 1) * 2 * W;
    const unsigned short* pqr = PQ + rd * W;
    if (tid < T_BLK) {
      const int k = tid / W, l = tid - k * W;
      const int pqk = pqr[k], pql = pqr[l];
      const int pk = pqk & 0xff, qk = pqk >> 8, pl = pql & 0xff, ql = pql >> 8;
      const double ck = cs[2 * k], sk = cs[2 * k + 1], cl = cs[2 * l], sl = cs[2 * l + 1];
      const double g00 = cur[pk * GLD + pl], g01 = cur[pk * GLD + ql];
      const double g10 = cur[qk * GLD + pl], g11 = cur[qk * GLD + ql];
      // T = R_k^T * Gb ; Gb' = T * R_l   with R = [[c, s], [-s, c]]
      const double t00 = ck * g00 - sk * g10, t01 = ck * g01 - sk * g11;
      const double t10 = sk * g00 + ck * g10, t11 = sk * g01 + ck * g11;
      nxt[pk * GLD + pl] = t00 * cl - t01 * sl;
      nxt[pk * GLD + ql] = t00 * sl + t01 * cl;
      nxt[qk * GLD + pl] = t10 * cl - t11 * sl;
      nxt[qk * GLD + ql] = t10 * sl + t11 * cl;
    } else if (tid < T_BLK + T_RA) {   // RA <- RA * R : row i, two of the W pairs per thread
      const int u = tid - T_BLK, i = u / (W / RE), l0 = (u - i * (W / RE)) * RE;
#pragma unroll
      for (int e = 0; e < RE; ++e) {
        const int pql = pqr[l0 + e];
        const int pl = pql & 0xff, ql = pql >> 8;
        const double cl = cs[2 * (l0 + e)], sl = cs[2 * (l0 + e) + 1];
        const double a = RA[i * GLD + pl], b = RA[i * GLD + ql];
        RA[i * GLD + pl] = cl * a - sl * b;
        RA[i * GLD + ql] = sl * a + cl * b;
      }
    } else if (tid >= NTH - W && rd + 1 < NR) {   // look-ahead: pair j of round rd+1
      const int j = tid - (NTH - W);
      const int pqn = PQ[(rd + 1) * W + j];
      const int x = pqn & 0xff, y = pqn >> 8;
      // position of x, y in THIS round's pairing: partner and rotation column
      const int ix = POS[rd * NC + x], iy = POS[rd * NC + y];
      const int kx = ix >> 1, ky = iy >> 1;
      const int pqx = pqr[kx], pqy = pqr[ky];
      const int xp = pqx & 0xff, xq = pqx >> 8, yp = pqy & 0xff, yq = pqy >> 8;
      // column of R that produces the new vector: first of pair -> (c, -s), second -> (s, c)
      const double cx = cs[2 * kx], sx = cs[2 * kx + 1], cy = cs[2 * ky], sy = cs[2 * ky + 1];
      const double ux = (ix & 1) ? sx : cx, vx = (ix & 1) ? cx : -sx;   // new_x = ux*old[xp] + vx*old[xq]
      const double uy = (iy & 1) ? sy : cy, vy = (iy & 1) ? cy : -sy;
      auto quad = [&](int ap, int aq, double ua, double va, int bp, int bq, double ub, double vb) {
        const double g00 = cur[ap * GLD + bp], g01 = cur[ap * GLD + bq];
        const double g10 = cur[aq * GLD + bp], g11 = cur[aq * GLD + bq];
        return ua * (g00 * ub + g01 * vb) + va * (g10 * ub + g11 * vb);
      };
      const double gxx = quad(xp, xq, ux, vx, xp, xq, ux, vx);
      const double gyy = quad(yp, yq, uy, vy, yp, yq, uy, vy);
      const double gxy = quad(xp, xq, ux, vx, yp, yq, uy, vy);
      double c, sn, rot;
      plane_rot(gxx, gyy, gxy, tol2, c, sn, rot);
      csn[2 * j] = c;
      csn[2 * j + 1] = sn;
      mo = fmax(mo, rot);
    }
    __syncthreads();
    double* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  tstamp[3] = clock64();
  // ---- apply the accumulated rotation to the staged rows with DMMA, 8-row blocks per warp
  //      (two blocks per iteration for more independent chains was measured slower: 3248 vs 2659 cycles)
  const int nrb = (tot + 7) / 8;
  for (int rb = warp; rb < nrb; rb += NWARP) {
    const int r0 = rb * 8;
    double acc[TG][2];
#pragma unroll
    for (int n = 0; n < TG; ++n) acc[n][0] = acc[n][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < NC; k0 += 4) {
      const double a = S[(long)(k0 + t) * ld + r0 + g];
#pragma unroll
      for (int n = 0; n < TG; ++n) {
        const double b = RA[(k0 + t) * GLD + n * 8 + g];
        dmma884s(acc[n][0], acc[n][1], a, b);
      }
    }
    __syncwarp();
#pragma unroll
    for (int n = 0; n < TG; ++n) {
      S[(long)(n * 8 + 2 * t) * ld + r0 + g] = acc[n][0];
      S[(long)(n * 8 + 2 * t + 1) * ld + r0 + g] = acc[n][1];
    }
  }
  __syncthreads();
  tstamp[4] = clock64();
  // ---- write back
  for (int rr2 = rc2; rr2 < tot2; rr2 += 256) {
#pragma unroll
    for (int it = 0; it < NC / CP; ++it) {
      const int k = cpar + CP * it;
      const int c = (k < W) ? (c0 + k) : (c1 + k - W);
      if (c < ns) {
        const double2 v = reinterpret_cast<const double2*>(S + (long)k * ld)[rr2];
        if (rr2 < rows2)
          __stcg(reinterpret_cast<double2*>(A + (long)c * rows) + rr2, v);
        else
          __stcg(reinterpret_cast<double2*>(Jm + (long)c * ns) + (rr2 - rows2), v);
      }
    }
  }
  if ((tid < W || tid >= NTH - W) && mo > 0.0) atomic_max_pos(info, mo);
  if (g_qr_dbg != nullptr && tid == 0 && blockIdx.x == 0) {
    tstamp[5] = clock64();
    for (int i = 0; i < 6; ++i) g_qr_dbg[2048 + i] = tstamp[i];
  }
}

// ---------------------------------------------------------------------------
// Cluster-resident Jacobi: the whole iteration (all sweeps) in ONE launch of one 16-CTA thread-block
// cluster.  The working set ([A | J], <= 256 columns x 512 rows = 1 MB) lives in the distributed
// shared memory of the cluster for the whole SVD: every CTA holds the two 8-column blocks of its
// current block pair, runs the same Gram / rotation-round / apply phases as jacobi_gram_kernel,
// and the apply phase stores every updated column DIRECTLY into the shared memory of the CTA
// that needs it in the next round (st.shared::cluster through a generic pointer).  One
// cluster barrier per round replaces a kernel boundary; the per-round staging from / write-back
// to L2 (2.7 of the 11 us per round of the multi-launch version) and the launch gaps disappear.
// The round-robin tournament is the same circle method, so the rotations are the same.
__device__ __forceinline__ void rr_where(int n, int r, int x, int& k, int& slot) {
  // inverse of rr_pair: in round r block x sits in pair k as p (slot 0) or q (slot 1)
  if (x == n - 1) {
    k = 0;
    slot = 0;
    return;
  }
  if (x == r) {
    k = 0;
    slot = 1;
    return;
  }
  const int d = (x - r + (n - 1)) % (n - 1);      // x = (r + d) mod (n-1)
  if (d <= n / 2 - 1) {
    k = d;
    slot = 0;
  } else {
    k = (n - 1) - d;                               // x = (r - k) mod (n-1)
    slot = 1;
  }
}

// NTH threads per CTA: the rotation rounds are a latency chain run by the first NTR = 32 W threads
// (their mapping needs exactly 2W x 2W threads); the Gram and apply phases are throughput work and use
// all NTH / 32 warps (16 warps: Gram 1.7k -> 0.9k cycles, apply 3.7k -> 1.9k per block pairing).
template <int W, int NTH>
__global__ void __launch_bounds__(NTH, 1)
jacobi_cluster_kernel(double* __restrict__ A, double* __restrict__ Jm, int rows, int ns, int ld, int nblk_e,
                      double tol2, double conv, int max_sweeps, int cross_only, double* __restrict__ info,
                      double* __restrict__ sweepmax, int* __restrict__ flags) {
  namespace cg = cooperative_groups;
  constexpr int NC = 2 * W, NR = NC - 1, NTR = 32 * W, NWARP = NTH / 32, TG = NC / 8;
  constexpr int KH = NWARP / (TG * TG);   // k-splits of the Gram products (partial sums in G0..G[KH-1])
  static_assert(KH >= 1 && KH <= 4 && NWARP == KH * TG * TG, "Gram phase: warps = tiles x k-splits");
  static_assert(W == 8, "cluster kernel is written for 8-column blocks (256 threads)");
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int ncta = nblk_e / 2;
  const bool active = crank < ncta;
  extern __shared__ __align__(16) double sm[];
  double* buf0 = sm;                                 // [NC][ld]
  double* buf1 = sm + (long)NC * ld;                 // [NC][ld]
  double* G0 = buf1 + (long)NC * ld;                 // [NC][GLD]
  double* G1 = G0 + NC * GLD;
  double* GX = G1 + NC * GLD;                        // [2][NC][GLD] extra Gram partials (KH = 4)
  double* RA = GX + 2 * NC * GLD;
  double* CS = RA + NC * GLD;                        // [2][W][2]
  double** DST = reinterpret_cast<double**>(CS + 4 * W);   // [NC] destination column of the next round
  // inner tournaments: "full" = all pairs of the 16 staged columns (15 rounds), used for the first
  // block pairing of every sweep (it covers the pairs inside each block); "cross" = only the 8 x 8
  // pairs between the two blocks (8 rounds) for the other 28 pairings.  Every column pair is then
  // rotated exactly once per sweep instead of the within-block pairs 29 times; on bond matrices
  // this costs 0-1 extra sweeps for 45 % fewer rotation rounds (tools/jacobi_precond_study.py).
  unsigned short* PQf = reinterpret_cast<unsigned short*>(DST + NC);   // [NR][W]
  unsigned short* PQc = PQf + NR * W;                                   // [W][W]
  unsigned char* POSf = reinterpret_cast<unsigned char*>(PQc + W * W);  // [NR][NC]
  unsigned char* POSc = POSf + NR * NC;                                 // [W][NC]
  // look-ahead table: everything the thread that prepares pair j of inner round rd+1 needs
  // (positions and partners of its two columns in round rd), packed in one word -- one shared-memory
  // load on the serial chain instead of three dependent ones
  unsigned* LAf = reinterpret_cast<unsigned*>(POSc + W * NC);           // [NR-1][W]
  unsigned* LAc = LAf + (NR - 1) * W;                                   // [W-1][W]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int tot = rows + ns;
  const int rows2 = rows >> 1, tot2 = tot >> 1, ld2 = ld >> 1;
  const int nrd = nblk_e - 1;                        // rounds per sweep

  // ---- tables + initial load of the block pair of round 0 into buf0
  for (int i = tid; i < NR * W; i += NTH) {
    const int rd = i / W, k = i - rd * W;
    int p, q;
    rr_pair(NC, rd, k, p, q);
    PQf[i] = (unsigned short)(p | (q << 8));
    POSf[rd * NC + p] = (unsigned char)(2 * k);
    POSf[rd * NC + q] = (unsigned char)(2 * k + 1);
  }
  for (int i = tid; i < W * W; i += NTH) {
    const int rd = i / W, k = i - rd * W;
    const int p = k, q = W + ((k + rd) & (W - 1));
    PQc[i] = (unsigned short)(p | (q << 8));
    POSc[rd * NC + p] = (unsigned char)(2 * k);
    POSc[rd * NC + q] = (unsigned char)(2 * k + 1);
  }
  __syncthreads();
  for (int i = tid; i < (NR - 1) * W + (W - 1) * W; i += NTH) {
    const bool fullm = i < (NR - 1) * W;
    const int ii = fullm ? i : i - (NR - 1) * W;
    const int rd = ii / W, j = ii - rd * W;
    const unsigned short* PQm = fullm ? PQf : PQc;
    const unsigned char* POSm = fullm ? POSf : POSc;
    const int pqn = PQm[(rd + 1) * W + j];
    const int x = pqn & 0xff, y = pqn >> 8;
    const int ix = POSm[rd * NC + x], iy = POSm[rd * NC + y];
    const int kx = ix >> 1, ky = iy >> 1;
    const int pqx = PQm[rd * W + kx], pqy = PQm[rd * W + ky];
    const unsigned word = (unsigned)(pqx & 0xf) | ((unsigned)((pqx >> 8) & 0xf) << 4) | ((unsigned)(pqy & 0xf) << 8) |
                          ((unsigned)((pqy >> 8) & 0xf) << 12) | ((unsigned)kx << 16) | ((unsigned)ky << 20) |
                          ((unsigned)(ix & 1) << 24) | ((unsigned)(iy & 1) << 25);
    (fullm ? LAf : LAc)[ii] = word;
  }
  if (active) {
    int P, Q;
    rr_pair(nblk_e, 0, crank, P, Q);
    for (int i = tid; i < NC * ld2; i += NTH) {
      const int k = i / ld2, r2 = i - k * ld2;
      const int c = (k < W) ? (P * W + k) : (Q * W + k - W);
      double2 v = make_double2(0.0, 0.0);
      if (c < ns && r2 < tot2)
        v = (r2 < rows2) ? __ldcg(reinterpret_cast<const double2*>(A + (long)c * rows) + r2)
                         : __ldcg(reinterpret_cast<const double2*>(Jm + (long)c * ns) + (r2 - rows2));
      reinterpret_cast<double2*>(buf0 + (long)k * ld)[r2] = v;
    }
  }
  cluster.sync();

  int cur = 0, R = 0, sweeps_done = 0, converged = 0;
  // Per-pairing setup that does not depend on the incoming columns: identity in RA and the
  // destination (owner CTA, slot) of every local column in the following pairing.  It runs between
  // barrier.cluster.arrive and .wait of the previous pairing, hidden behind the barrier latency.
  auto setup_round = [&](int Rcur, int curbuf) {
    if (!active) return;
    const int Rnext = (Rcur + 1 == nrd) ? 0 : Rcur + 1;
    double* Snext = curbuf ? buf0 : buf1;
    if (tid < NC) {
      int P, Q;
      rr_pair(nblk_e, Rcur, crank, P, Q);
      const int blk = (tid < W) ? P : Q;
      int k2, slot;
      rr_where(nblk_e, Rnext, blk, k2, slot);
      double* remote = cluster.map_shared_rank(Snext, (unsigned)k2);
      DST[tid] = remote + (long)(slot * W + (tid & (W - 1))) * ld;
    }
    for (int i = tid; i < NC * GLD; i += NTH) RA[i] = ((i / GLD) == (i % GLD)) ? 1.0 : 0.0;
  };
  setup_round(0, 0);
  __syncthreads();
  long long ph[6] = {0, 0, 0, 0, 0, 0};   // TNML_QR_DEBUG: cycles in setup / gram / rounds / apply / cluster barrier
  const bool prof = (g_qr_dbg != nullptr) && crank == 0 && tid == 0;
  for (int sw = 0; sw < max_sweeps; ++sw) {
    double mo = 0.0;
    for (int rr = 0; rr < nrd; ++rr) {
      const int Rn = (R + 1 == nrd) ? 0 : R + 1;
      const bool full = (rr == 0) || !cross_only;
      const int nir = full ? NR : W;                       // inner rounds of this block pairing
      const unsigned short* PQ = full ? PQf : PQc;
      const unsigned* LA = full ? LAf : LAc;
      long long c0 = prof ? clock64() : 0, c1 = c0, c2 = c0, c3 = c0;
      if (active) {
        double* S = cur ? buf1 : buf0;
        // ---- Gram (rows split over two warps per tile, 4 accumulator chains)
        if (prof) c1 = clock64();
        {
          const int tile = warp % (TG * TG), kh = warp / (TG * TG);
          const int ti = tile / TG, tj = tile - ti * TG;
          const int nk4 = (rows + 3) >> 2;
          const int kbeg = (nk4 * kh) / KH, kend = (nk4 * (kh + 1)) / KH;
          const double* pa = S + (long)(ti * 8 + g) * ld + t;
          const double* pb = S + (long)(tj * 8 + g) * ld + t;
          double d[4][2];
#pragma unroll
          for (int u = 0; u < 4; ++u) d[u][0] = d[u][1] = 0.0;
          int k = kbeg;
          for (; k + 4 <= kend; k += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r0 = (k + u) * 4;
              const bool ok = (r0 + t) < rows;
              dmma884s(d[u][0], d[u][1], ok ? pa[r0] : 0.0, ok ? pb[r0] : 0.0);
            }
          }
          for (; k < kend; ++k) {
            const int r0 = k * 4;
            const bool ok = (r0 + t) < rows;
            dmma884s(d[0][0], d[0][1], ok ? pa[r0] : 0.0, ok ? pb[r0] : 0.0);
          }
          double* Gd = (kh == 0) ? G0 : (kh == 1 ? G1 : GX + (kh - 2) * NC * GLD);
          Gd[(ti * 8 + g) * GLD + tj * 8 + 2 * t] = (d[0][0] + d[1][0]) + (d[2][0] + d[3][0]);
          Gd[(ti * 8 + g) * GLD + tj * 8 + 2 * t + 1] = (d[0][1] + d[1][1]) + (d[2][1] + d[3][1]);
        }
        __syncthreads();
        if (tid < NC * NC) {
          const int r = tid / NC, c = tid - r * NC;
          double acc = G0[r * GLD + c];
          if (KH > 1) acc += G1[r * GLD + c];
          if (KH > 2) acc += GX[r * GLD + c] + GX[NC * GLD + r * GLD + c];
          G0[r * GLD + c] = acc;
        }
        __syncthreads();
        // ---- rotation rounds on the Gram matrix (identical to jacobi_gram_kernel)
        if (prof) c2 = clock64();
        double* curG = G0;
        double* nxtG = G1;
        if (tid < W) {
          const int pq = PQ[tid];
          const int p = pq & 0xff, q = pq >> 8;
          double c, sn, rot;
          plane_rot(curG[p * GLD + p], curG[q * GLD + q], curG[p * GLD + q], tol2, c, sn, rot);
          CS[2 * tid] = c;
          CS[2 * tid + 1] = sn;
          mo = fmax(mo, rot);
        }
        __syncthreads();
        constexpr int RE = 2;
        constexpr int T_BLK = W * W, T_RA = NC * W / RE;
        static_assert(T_BLK + T_RA + W <= NTR, "thread budget");
        for (int rd = 0; rd < nir; ++rd) {
          const double* cs = CS + (rd & 1) * 2 * W;
          double* csn = CS + ((rd + 1) & 1) * 2 * W;
          const unsigned short* pqr = PQ + rd * W;
          if (tid < T_BLK) {
            const int k = tid / W, l = tid - k * W;
            const int pqk = pqr[k], pql = pqr[l];
            const int pk = pqk & 0xff, qk = pqk >> 8, pl = pql & 0xff, ql = pql >> 8;
            const double ck = cs[2 * k], sk = cs[2 * k + 1], cl = cs[2 * l], sl = cs[2 * l + 1];
            const double g00 = curG[pk * GLD + pl], g01 = curG[pk * GLD + ql];
            const double g10 = curG[qk * GLD + pl], g11 = curG[qk * GLD + ql];
            const double t00 = ck * g00 - sk * g10, t01 = ck * g01 - sk * g11;
            const double t10 = sk * g00 + ck * g10, t11 = sk * g01 + ck * g11;
            nxtG[pk * GLD + pl] = t00 * cl - t01 * sl;
            nxtG[pk * GLD + ql] = t00 * sl + t01 * cl;
            nxtG[qk * GLD + pl] = t10 * cl - t11 * sl;
            nxtG[qk * GLD + ql] = t10 * sl + t11 * cl;
          } else if (tid < T_BLK + T_RA) {
            const int u = tid - T_BLK, i = u / (W / RE), l0 = (u - i * (W / RE)) * RE;
#pragma unroll
            for (int e = 0; e < RE; ++e) {
              const int pql = pqr[l0 + e];
              const int pl = pql & 0xff, ql = pql >> 8;
              const double cl = cs[2 * (l0 + e)], sl = cs[2 * (l0 + e) + 1];
              const double a = RA[i * GLD + pl], b = RA[i * GLD + ql];
              RA[i * GLD + pl] = cl * a - sl * b;
              RA[i * GLD + ql] = sl * a + cl * b;
            }
          } else if (tid >= NTR - W && tid < NTR && rd + 1 < nir) {
            const int j = tid - (NTR - W);
            const unsigned la = LA[rd * W + j];
            const int xp = la & 0xf, xq = (la >> 4) & 0xf, yp = (la >> 8) & 0xf, yq = (la >> 12) & 0xf;
            const int kx = (la >> 16) & 0xf, ky = (la >> 20) & 0xf;
            const bool px = (la >> 24) & 1, py = (la >> 25) & 1;
            const double cx = cs[2 * kx], sx = cs[2 * kx + 1], cy = cs[2 * ky], sy = cs[2 * ky + 1];
            const double ux = px ? sx : cx, vx = px ? cx : -sx;
            const double uy = py ? sy : cy, vy = py ? cy : -sy;
            auto quad = [&](int ap, int aq, double ua, double va, int bp, int bq, double ub, double vb) {
              const double g00 = curG[ap * GLD + bp], g01 = curG[ap * GLD + bq];
              const double g10 = curG[aq * GLD + bp], g11 = curG[aq * GLD + bq];
              return ua * (g00 * ub + g01 * vb) + va * (g10 * ub + g11 * vb);
            };
            const double gxx = quad(xp, xq, ux, vx, xp, xq, ux, vx);
            const double gyy = quad(yp, yq, uy, vy, yp, yq, uy, vy);
            const double gxy = quad(xp, xq, ux, vx, yp, yq, uy, vy);
            double c, sn, rot;
            plane_rot(gxx, gyy, gxy, tol2, c, sn, rot);
            csn[2 * j] = c;
            csn[2 * j + 1] = sn;
            mo = fmax(mo, rot);
          }
          __syncthreads();
          double* tmp = curG;
          curG = nxtG;
          nxtG = tmp;
        }
        // ---- apply the accumulated rotation; the result goes straight to the next round's owners
        if (prof) c3 = clock64();
        const int nrb = (tot + 7) / 8;
        for (int rb = warp; rb < nrb; rb += NWARP) {
          const int r0 = rb * 8;
          double acc[TG][2];
#pragma unroll
          for (int n = 0; n < TG; ++n) acc[n][0] = acc[n][1] = 0.0;
#pragma unroll
          for (int k0 = 0; k0 < NC; k0 += 4) {
            const double a = S[(long)(k0 + t) * ld + r0 + g];
#pragma unroll
            for (int n = 0; n < TG; ++n) {
              const double b = RA[(k0 + t) * GLD + n * 8 + g];
              dmma884s(acc[n][0], acc[n][1], a, b);
            }
          }
#pragma unroll
          for (int n = 0; n < TG; ++n) {
            DST[n * 8 + 2 * t][r0 + g] = acc[n][0];
            DST[n * 8 + 2 * t + 1][r0 + g] = acc[n][1];
          }
        }
      }
      const long long c4 = prof ? clock64() : 0;
      // every column of the next pairing must have arrived and this pairing's buffer must be free:
      // split barrier, the next pairing's setup runs in its shadow
      __syncthreads();                       // all warps of this CTA are done with RA / DST / S
      cluster.barrier_arrive();
      setup_round(Rn, cur ^ 1);
      cluster.barrier_wait();
      if (prof) {
        const long long c5 = clock64();
        ph[0] += c1 - c0;
        ph[1] += c2 - c1;
        ph[2] += c3 - c2;
        ph[3] += c4 - c3;
        ph[4] += c5 - c4;
        ph[5] += 1;
      }
      cur ^= 1;
      R = Rn;
    }
    // ---- end of sweep: largest rotated cos^2 over the cluster
    if (active && (tid < W || (tid >= NTR - W && tid < NTR)) && mo > 0.0) atomic_max_pos(sweepmax + sw, mo);
    __threadfence();
    cluster.sync();
    const double smax = __ldcg(sweepmax + sw);
    sweeps_done = sw + 1;
    if (smax <= conv) {
      converged = 1;
      break;
    }
  }
  // ---- write the blocks this CTA holds (pair of round R) back to global memory
  if (active) {
    const double* S = cur ? buf1 : buf0;
    int P, Q;
    rr_pair(nblk_e, R, crank, P, Q);
    for (int i = tid; i < NC * ld2; i += NTH) {
      const int k = i / ld2, r2 = i - k * ld2;
      const int c = (k < W) ? (P * W + k) : (Q * W + k - W);
      if (c < ns && r2 < tot2) {
        const double2 v = reinterpret_cast<const double2*>(S + (long)k * ld)[r2];
        if (r2 < rows2)
          __stcg(reinterpret_cast<double2*>(A + (long)c * rows) + r2, v);
        else
          __stcg(reinterpret_cast<double2*>(Jm + (long)c * ns) + (r2 - rows2), v);
      }
    }
  }
  if (prof)
    for (int i = 0; i < 6; ++i) g_qr_dbg[2064 + i] = ph[i];
  if (crank == 0 && tid == 0) {
    info[3] += (double)sweeps_done;
    info[6] = (double)converged;
    flags[0] = converged;
  }
}

__global__ void jacobi_sweep_end_kernel(double* __restrict__ info, int* __restrict__ flags, double tol) {
  if (threadIdx.x == 0 && !flags[0]) {
    info[3] += 1.0;
    info[5] = info[0];
    if (info[0] <= tol) flags[0] = 1;
    info[0] = 0.0;
  }
}

__global__ void svd_init_kernel(double* __restrict__ J, int ns, double* __restrict__ info, int* __restrict__ flags) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = (long)ns * ns;
  if (idx < n) J[idx] = ((idx / ns) == (idx % ns)) ? 1.0 : 0.0;
  if (idx < 8) info[idx] = 0.0;
  if (idx < 4) flags[idx] = 0;
}

struct SvdGeom {
  BondGeom g;
  int nlA, nlB;  // label multiplicity on side A (site b) / side B (site b+1)
  int nA, nB;    // 2*ml*nlA, 2*mr*nlB
  int bigIsA;    // X columns are indexed by the small side, rows by the big side
  int nb, ns;
};

__device__ __forceinline__ double bond_elem(const double* __restrict__ Bc, const SvdGeom& sg, int ia, int ib) {
  int la = ia % sg.nlA, as = ia / sg.nlA;
  int lb = ib % sg.nlB, tb = ib / sg.nlB;
  int s = as & 1, a = as >> 1;
  int b = tb % sg.g.mr, t = tb / sg.g.mr;
  int l = (sg.nlA > 1) ? la : lb;
  return Bc[a * sg.g.sa + s * sg.g.ss + t * sg.g.st + b * sg.g.sb + l * sg.g.sl];
}

__global__ void svd_gather_kernel(const double* __restrict__ Bc, SvdGeom sg, double* __restrict__ X) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = (long)sg.nb * sg.ns;
  if (idx >= n) return;
  int r = (int)(idx % sg.nb);
  int c = (int)(idx / sg.nb);
  int ia = sg.bigIsA ? r : c;
  int ib = sg.bigIsA ? c : r;
  X[idx] = bond_elem(Bc, sg, ia, ib);
}

// sigma^2, sort (descending, stable), ITensor truncation rule.  One CTA.
__global__ void __launch_bounds__(1024)
svd_finalize_kernel(const double* __restrict__ X, int nb, int ns, double* __restrict__ sig2,
                    int* __restrict__ perm, double cutoff, int maxm, int minm, int do_rel,
                    double* __restrict__ info) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < ns; c += 32) {
    const double* x = X + (long)c * nb;
    double a = 0.0;
    for (int r = lane; r < nb; r += 32) a = fma(x[r], x[r], a);
    a = wsum(a);
    if (lane == 0) sig2[c] = a;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < ns; c += blockDim.x) {
    double v = sig2[c];
    int rank = 0;
    for (int k = 0; k < ns; ++k) {
      double u = sig2[k];
      rank += (u > v || (u == v && k < c)) ? 1 : 0;
    }
    perm[rank] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // ASSUMED ITensor v2 `truncate`; mirrors oracle.truncate_spectrum
    int m = ns;
    double terr = 0.0;
    while (m > maxm) {
      terr += sig2[perm[m - 1]];
      --m;
    }
    double scale = 1.0;
    if (do_rel) {
      double s = 0.0;
      for (int k = 0; k < ns; ++k) s += sig2[perm[k]];
      scale = (s == 0.0) ? 1.0 : s;
    }
    while (m > minm && m > 1 && terr + sig2[perm[m - 1]] < cutoff * scale) {
      terr += sig2[perm[m - 1]];
      --m;
    }
    info[1] = (double)m;
    info[2] = terr / scale;
    int nz = 0;                                   // kept vectors whose singular value is exactly zero
    for (int k = 0; k < m; ++k) nz += (sig2[perm[k]] > 0.0) ? 0 : 1;
    info[7] = (double)nz;
  }
}

// ITensor's svd returns orthonormal U columns also for zero singular values (Minm can force them
// to be kept: real MNIST border sites have phi = [1, 0] for every image, so bond matrices there are
// rank deficient).  The scatter kernels leave such columns zero; this pass completes them: for each
// zero column, the unit vector e_r with the smallest weight in the span of the other columns,
// projected out of that span (one projection from the row of U, one re-orthogonalisation).
// iso(r, k) = W[(r / nl) * (m * nl) + k * nl + (r % nl)]  (dir 1)  |  W[k * nrows + r]  (dir 2).  One CTA.
__global__ void __launch_bounds__(1024)
svd_complete_iso_kernel(double* __restrict__ W, int nrows, int m, int nl, int dir1) {
  extern __shared__ double csm[];
  double* v = csm;                 // [nrows]
  double* coef = csm + nrows;      // [m]
  double* red = coef + m;          // [32]
  __shared__ int s_row, s_col;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  auto at = [&](int r, int k) -> double& {
    return dir1 ? W[(long)(r / nl) * ((long)m * nl) + (long)k * nl + (r % nl)] : W[(long)k * nrows + r];
  };
  for (int guard = 0; guard < m; ++guard) {
    // first zero column (column norms by warps)
    if (tid == 0) s_col = m;
    __syncthreads();
    for (int k = warp; k < m; k += nw) {
      double a = 0.0;
      for (int r = lane; r < nrows; r += 32) a = fma(at(r, k), at(r, k), a);
      a = wsum(a);
      if (lane == 0 && a < 0.25) atomicMin(&s_col, k);
    }
    __syncthreads();
    const int c = s_col;
    if (c >= m) return;
    // row with the smallest weight sum_k U[r][k]^2
    double best = 1e300;
    int brow = 0;
    for (int r = tid; r < nrows; r += blockDim.x) {
      double a = 0.0;
      for (int k = 0; k < m; ++k) a = fma(at(r, k), at(r, k), a);
      if (a < best) {
        best = a;
        brow = r;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
      if (ob < best || (ob == best && orow < brow)) {
        best = ob;
        brow = orow;
      }
    }
    if (lane == 0) {
      red[warp] = best;
      reinterpret_cast<int*>(red + 32)[warp] = brow;
    }
    __syncthreads();
    if (tid == 0) {
      double b = red[0];
      int br = reinterpret_cast<int*>(red + 32)[0];
      for (int w2 = 1; w2 < nw; ++w2) {
        const double ob = red[w2];
        const int orow = reinterpret_cast<int*>(red + 32)[w2];
        if (ob < b || (ob == b && orow < br)) {
          b = ob;
          br = orow;
        }
      }
      s_row = br;
    }
    __syncthreads();
    const int e = s_row;
    for (int k = tid; k < m; k += blockDim.x) coef[k] = at(e, k);
    __syncthreads();
    for (int r = tid; r < nrows; r += blockDim.x) {
      double a = (r == e) ? 1.0 : 0.0;
      for (int k = 0; k < m; ++k) a = fma(-at(r, k), coef[k], a);
      v[r] = a;
    }
    __syncthreads();
    // re-orthogonalise against every other column, then normalise
    for (int k = warp; k < m; k += nw) {
      double d = 0.0;
      for (int r = lane; r < nrows; r += 32) d = fma(at(r, k), v[r], d);
      d = wsum(d);
      if (lane == 0) coef[k] = (k == c) ? 0.0 : d;
    }
    __syncthreads();
    double nn = 0.0;
    for (int r = tid; r < nrows; r += blockDim.x) {
      double a = v[r];
      for (int k = 0; k < m; ++k) a = fma(-at(r, k), coef[k], a);
      v[r] = a;
      nn = fma(a, a, nn);
    }
    nn = wsum(nn);
    if (lane == 0) red[warp] = nn;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w2 = 0; w2 < nw; ++w2) t += red[w2];
      red[0] = 1.0 / sqrt(t);
    }
    __syncthreads();
    const double inv = red[0];
    for (int r = tid; r < nrows; r += blockDim.x) at(r, c) = v[r] * inv;
    __threadfence_block();
    __syncthreads();
  }
}

// iso side gets unit vectors, the other side sigma * unit vectors
__global__ void svd_scatter_kernel(const double* __restrict__ X, const double* __restrict__ J,
                                   const double* __restrict__ sig2, const int* __restrict__ perm, SvdGeom sg,
                                   int isoIsA, int m, double* __restrict__ Wb, double* __restrict__ Wb1) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long nAm = (long)sg.nA * m, nBm = (long)sg.nB * m;
  if (idx >= nAm + nBm) return;
  const bool sideA = idx < nAm;
  long e = sideA ? idx : idx - nAm;
  int k, i;  // kept vector index, index on this side
  if (sideA) {  // W_b[as][k][l] : e = (as*m + k)*nlA + l
    int l = (int)(e % sg.nlA);
    long r = e / sg.nlA;
    k = (int)(r % m);
    int as = (int)(r / m);
    i = as * sg.nlA + l;
  } else {  // W_{b+1}[k][tb][l] : e = k*nB + ib
    k = (int)(e / sg.nB);
    i = (int)(e % sg.nB);
  }
  const int c = perm[k];
  const bool thisIsBig = (sideA == (sg.bigIsA != 0));
  const bool thisIsIso = (sideA == (isoIsA != 0));
  double v = thisIsBig ? X[(long)c * sg.nb + i] : J[(long)c * sg.ns + i];
  // X-type carries sigma, J-type is a unit vector
  if (thisIsIso && thisIsBig) {
    double sg1 = sqrt(sig2[c]);
    v = (sg1 > 0.0) ? v / sg1 : 0.0;
  } else if (!thisIsIso && !thisIsBig) {
    v *= sqrt(sig2[c]);
  }
  if (sideA)
    Wb[e] = v;
  else
    Wb1[e] = v;
}


// ---------------------------------------------------------------------------
// QR preconditioning (Drmac-Veselic): X = Q R by Householder, then one-sided
// Jacobi on R^T.  On graded / nearly rank-deficient bond matrices plain Jacobi
// needs 25-40 sweeps, Jacobi on R^T about 10 (measured, DESIGN.md "SVD").
//
// Dataflow Householder QR: one warp owns one column (kept in registers); it
// applies reflector k as soon as column k has published it (flag in global
// memory), then forms its own reflector.  No grid-wide barrier.  All ns warps
// must be co-resident (ns/8 CTAs <= number of SMs).
template <int RPL>   // rows per lane: nb <= 32*RPL
__global__ void __launch_bounds__(256, 1)
qr_dataflow_kernel(double* __restrict__ X, int nb, int ns, double* __restrict__ tau, volatile int* ready) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= ns) return;
  double* aj = X + (long)j * nb;
  double a[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    a[i] = (r < nb) ? aj[r] : 0.0;
  }
  for (int k = 0; k < j; ++k) {
    if (lane == 0)
      while (ready[k] == 0) {
      }
    __syncwarp();
    __threadfence();
    const double* vk = X + (long)k * nb;
    const double tk = __ldcg(tau + k);
    double dot = 0.0;
    if (RPL <= 20) {
      double v[RPL <= 20 ? RPL : 1];
#pragma unroll
      for (int i = 0; i < (RPL <= 20 ? RPL : 1); ++i) {
        int r = lane + 32 * i;
        v[i] = (r > k && r < nb) ? __ldcg(vk + r) : ((r == k) ? 1.0 : 0.0);
        dot = fma(v[i], a[i], dot);
      }
      dot = wsum(dot);
      const double f = tk * dot;
#pragma unroll
      for (int i = 0; i < (RPL <= 20 ? RPL : 1); ++i) a[i] = fma(-f, v[i], a[i]);
    } else {
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        int r = lane + 32 * i;
        double vv = (r > k && r < nb) ? __ldcg(vk + r) : ((r == k) ? 1.0 : 0.0);
        dot = fma(vv, a[i], dot);
      }
      dot = wsum(dot);
      const double f = tk * dot;
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        int r = lane + 32 * i;
        double vv = (r > k && r < nb) ? __ldcg(vk + r) : ((r == k) ? 1.0 : 0.0);
        a[i] = fma(-f, vv, a[i]);
      }
    }
  }
  // own reflector (LAPACK dlarfg): H = I - tau v v^T, v(j) = 1
  double alpha = 0.0, xn2 = 0.0;
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    if (r == j) alpha = a[i];
    if (r > j && r < nb) xn2 = fma(a[i], a[i], xn2);
  }
  alpha = wsum(alpha);
  xn2 = wsum(xn2);
  double tj = 0.0, beta = alpha, scale = 0.0;
  if (xn2 > 0.0) {
    beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
    tj = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    if (r < nb) {
      double out = a[i];
      if (r == j) out = beta;
      if (r > j) out = a[i] * scale;
      __stcg(aj + r, out);
    }
  }
  if (lane == 0) __stcg(tau + j, tj);
  __threadfence();
  __syncwarp();
  if (lane == 0) ready[j] = 1;
}

// Tall columns (class-C bond matrices, nb = 20 m > 1280 rows): 75+ rows per lane do not fit the
// register file (the register-resident kernels above spilled 1-8 KB per thread and made the first
// QR of a 2400 x 240 matrix cost 9 ms), so the warp's column lives in shared memory and every
// loop is a plain strided loop.  Same dataflow: one warp per column, reflector k consumed as soon
// as its flag is published.
__global__ void __launch_bounds__(256, 1)
qr_dataflow_smem_kernel(double* __restrict__ X, int nb, int ns, double* __restrict__ tau, volatile int* ready) {
  extern __shared__ __align__(16) double qcol[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * (blockDim.x >> 5) + warp;   // 8 columns per CTA, fewer when 8 do not fit in shared memory
  if (j >= ns) return;
  double* a = qcol + (long)warp * nb;
  double* aj = X + (long)j * nb;
  for (int r = lane; r < nb; r += 32) a[r] = aj[r];
  __syncwarp();
  for (int k = 0; k < j; ++k) {
    if (lane == 0)
      while (ready[k] == 0) {
      }
    __syncwarp();
    __threadfence();
    const double* vk = X + (long)k * nb;
    const double tk = __ldcg(tau + k);
    double dot = (lane == 0) ? a[k] : 0.0;               // v(k) = 1
    for (int r = k + 1 + lane; r < nb; r += 32) dot = fma(__ldcg(vk + r), a[r], dot);
    dot = wsum(dot);
    const double f = tk * dot;
    if (lane == 0) a[k] -= f;
    for (int r = k + 1 + lane; r < nb; r += 32) a[r] = fma(-f, __ldcg(vk + r), a[r]);
    __syncwarp();
  }
  // own reflector (LAPACK dlarfg): H = I - tau v v^T, v(j) = 1
  const double alpha = a[j];
  double xn2 = 0.0;
  for (int r = j + 1 + lane; r < nb; r += 32) xn2 = fma(a[r], a[r], xn2);
  xn2 = wsum(xn2);
  double tj = 0.0, beta = alpha, scale = 0.0;
  if (xn2 > 0.0) {
    beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
    tj = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  for (int r = lane; r < nb; r += 32) {
    double out = a[r];
    if (r == j) out = beta;
    if (r > j) out = a[r] * scale;
    __stcg(aj + r, out);
  }
  if (lane == 0) __stcg(tau + j, tj);
  __threadfence();
  __syncwarp();
  if (lane == 0) ready[j] = 1;
}

// Y[i] = Q * [src[:, perm[i]] * sc; 0] for tall columns: the vector lives in shared memory.
__global__ void __launch_bounds__(256, 1)
apply_q_smem_kernel(const double* __restrict__ X, const double* __restrict__ tau, const double* __restrict__ Jm,
                    const int* __restrict__ perm, int nb, int ns, int m, double* __restrict__ Y,
                    const double* __restrict__ scale_sig2, const int* __restrict__ rowperm) {
  extern __shared__ __align__(16) double qcol[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i0 >= m) return;
  double* y = qcol + (long)warp * nb;
  const double* jc = Jm + (long)perm[i0] * ns;
  double sc = 1.0;
  if (scale_sig2 != nullptr) {
    const double sg = sqrt(scale_sig2[perm[i0]]);
    sc = (sg > 0.0) ? 1.0 / sg : 0.0;
  }
  for (int r = lane; r < nb; r += 32) y[r] = (r < ns) ? jc[r] * sc : 0.0;
  __syncwarp();
  for (int k = ns - 1; k >= 0; --k) {
    const double* vk = X + (long)k * nb;
    const double tk = tau[k];
    double dot = (lane == 0) ? y[k] : 0.0;
    for (int r = k + 1 + lane; r < nb; r += 32) dot = fma(__ldg(vk + r), y[r], dot);
    dot = wsum(dot);
    const double f = tk * dot;
    if (lane == 0) y[k] -= f;
    for (int r = k + 1 + lane; r < nb; r += 32) y[r] = fma(-f, __ldg(vk + r), y[r]);
    __syncwarp();
  }
  double* out = Y + (long)i0 * nb;
  for (int r = lane; r < nb; r += 32) out[rowperm ? rowperm[r] : r] = y[r];
}

// Same factorisation, 32 consecutive columns per CTA (32 warps): reflectors of the CTA's own
// columns travel through shared memory (the pivot chain k -> k+1 stays on one SM for 31 of 32
// steps), reflectors of earlier CTAs come from global memory with batched flag polls (the
// flags are monotone: ready[k] implies ready[k-1]) and one-step-ahead prefetch.
template <int RPL>
__global__ void __launch_bounds__(1024)
qr_block_kernel(double* __restrict__ X, int nb, int ns, double* __restrict__ tau, volatile int* ready) {
  extern __shared__ __align__(16) double qsm[];
  double* sv = qsm;                         // [32][nb]
  double* stau = qsm + 32 * (long)nb;       // [32]
  volatile int* sflag = reinterpret_cast<volatile int*>(stau + 32);  // [32]
  // The SM arbiter issues highest-warp-id first (B300_MICROARCH.md): map the FIRST column of
  // the block to the LAST warp so that the warp holding the next pivot always has priority
  // over the warps that merely apply the reflector (measured: 5.7k -> cycles per pivot).
  const int lane = threadIdx.x & 31, w = 31 - (threadIdx.x >> 5);
  const int c0 = blockIdx.x * 32;
  const int j = c0 + w;
  if (threadIdx.x < 32) sflag[threadIdx.x] = 0;
  __syncthreads();
  if (j >= ns) return;
  double* aj = X + (long)j * nb;
  double a[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    a[i] = (r < nb) ? aj[r] : 0.0;
  }
  // ---- phase A: reflectors published by earlier CTAs (global memory)
  int navail = 0;
  double vn[RPL];
  double tn = 0.0;
  auto gload = [&](int k) {
    const double* vk = X + (long)k * nb;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      int r = lane + 32 * i;
      vn[i] = (r > k && r < nb) ? __ldcg(vk + r) : ((r == k) ? 1.0 : 0.0);
    }
    tn = __ldcg(tau + k);
  };
  for (int k = 0; k < c0; ++k) {
    if (k >= navail) {
      while (true) {
        int idx = navail + lane;
        int f = (idx < c0) ? ready[idx] : 0;
        unsigned m = __ballot_sync(0xffffffffu, f != 0);
        int cnt = __ffs(~m) - 1;            // leading run of ready flags
        if (m == 0xffffffffu) cnt = 32;
        if (cnt > 0) {
          navail += cnt;
          break;
        }
        __nanosleep(400);   // do not hammer the L2 slice that holds the flags
      }
      __threadfence();
      gload(k);
    }
    double v[RPL];
    const double tk = tn;
#pragma unroll
    for (int i = 0; i < RPL; ++i) v[i] = vn[i];
    if (k + 1 < navail) gload(k + 1);       // prefetch the next reflector
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < RPL; ++i) dot = fma(v[i], a[i], dot);
    dot = wsum(dot);
    const double f = tk * dot;
#pragma unroll
    for (int i = 0; i < RPL; ++i) a[i] = fma(-f, v[i], a[i]);
    if (k + 1 < c0 && k + 1 >= navail) {
      // next one not yet known ready: it is (re)loaded after the poll at the loop top
    }
  }
  long long t_a = clock64();
  // ---- phase B: reflectors of this CTA (shared memory; stored with explicit 0 / 1 so that
  //      no per-element guards are needed).  Waiting warps back off with nanosleep so that
  //      the pivot warp owns the issue slots and the shared-memory pipe.
  for (int kk = 0; kk < w; ++kk) {
    // Warp-uniform wait (every lane reads the flag, the vote result is uniform): a lane-0-only
    // spin makes the warp divergent and the compiler then wraps every later __shfl_sync in
    // WARPSYNC.COLLECTIVE, which cost ~4k cycles per pivot.  Only the next three pivot
    // owners poll continuously, the others back off.
    while (true) {
      const int f = sflag[kk];
      if (__all_sync(0xffffffffu, f != 0)) break;
      if (kk < w - 3) __nanosleep(200);
    }
    __threadfence_block();
    if (kk == w - 1) t_a = clock64();   // the flag this warp's pivot was waiting for
    const double* vk = sv + (long)kk * nb;
    const double tk = stau[kk];
    double v[RPL];
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      int r = lane + 32 * i;
      v[i] = (r < nb) ? vk[r] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < RPL; ++i) dot = fma(v[i], a[i], dot);
    dot = wsum(dot);
    const double f = tk * dot;
#pragma unroll
    for (int i = 0; i < RPL; ++i) a[i] = fma(-f, v[i], a[i]);
  }
  // ---- own reflector (dlarfg)
  const long long t_b = clock64();
  double alpha = 0.0, xn2 = 0.0;
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    if (r == j) alpha = a[i];
    if (r > j && r < nb) xn2 = fma(a[i], a[i], xn2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
    xn2 += __shfl_xor_sync(0xffffffffu, xn2, o);
  }
  double tj = 0.0, beta = alpha, scale = 0.0;
  if (xn2 > 0.0) {
    beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
    tj = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  double* svw = sv + (long)w * nb;
  double outv[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    double out = a[i];
    if (r == j) out = beta;
    if (r > j) out = a[i] * scale;
    outv[i] = out;
    if (r < nb) svw[r] = (r < j) ? 0.0 : ((r == j) ? 1.0 : out);
  }
  if (lane == 0) stau[w] = tj;
  // global copies are ISSUED before the cta-scope publication (they are not waited for), so
  // that the batched gpu-scope fence of a later warp is cumulative over them
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    if (r < nb) __stcg(aj + r, outv[i]);
  }
  if (lane == 0) __stcg(tau + j, tj);
  __threadfence_block();
  __syncwarp();
  if (lane == 0) sflag[w] = 1;
  if (g_qr_dbg != nullptr && lane == 0) {
    g_qr_dbg[j * 4 + 0] = t_a;
    g_qr_dbg[j * 4 + 1] = t_b;
    g_qr_dbg[j * 4 + 2] = clock64();
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    g_qr_dbg[j * 4 + 3] = sm;
  }
  // Global publication in batches of 8 columns: a gpu-scope fence (MEMBAR.GL + L1
  // invalidate) stalls the memory pipe of the whole SM for ~2 us; issued by every pivot warp
  // it sat on the pivot chain (640 us per QR).  The last column of each batch fences once --
  // the earlier columns of the batch were observed through the cta-scope flag chain, so the
  // fence is cumulative over their stores (PTX memory model causality order).
  const bool last_in_cta = (j == ns - 1) || (w == 31);
  if ((w & 7) == 7 || last_in_cta) {
    __threadfence_block();
    __threadfence();
    __syncwarp();
    const int first = c0 + (w & ~7);
    if (lane <= (w & 7)) ready[first + lane] = 1;
  }
}

// M = R^T as a column-major ns x ns matrix: M[:, j] = row j of R (upper triangular)
__global__ void rt_form_kernel(const double* __restrict__ X, int nb, int ns, double* __restrict__ M) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)ns * ns) return;
  int c = (int)(idx % ns);   // row index inside column j of M
  int j = (int)(idx / ns);
  M[idx] = (c >= j) ? X[(long)c * nb + j] : 0.0;
}

// big-side unit vectors: Y[i] = Q * [J'[:, perm[i]]; 0].  One warp per kept vector.
// scale_sig2 != nullptr: the source column is sigma * unit vector, normalise it (zero if sigma = 0);
// rowperm != nullptr: out[rowperm[r]] = y[r] (undo the column sort of the first QR).
template <int RPL>
__global__ void __launch_bounds__(256, 1)
apply_q_kernel(const double* __restrict__ X, const double* __restrict__ tau, const double* __restrict__ Jm,
               const int* __restrict__ perm, int nb, int ns, int m, double* __restrict__ Y,
               const double* __restrict__ scale_sig2, const int* __restrict__ rowperm) {
  const int lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i0 >= m) return;
  const double* jc = Jm + (long)perm[i0] * ns;
  double sc = 1.0;
  if (scale_sig2 != nullptr) {
    const double sg = sqrt(scale_sig2[perm[i0]]);
    sc = (sg > 0.0) ? 1.0 / sg : 0.0;
  }
  double y[RPL];
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    y[i] = (r < ns) ? jc[r] * sc : 0.0;
  }
  double vn[RPL];
  double tn = 0.0;
  auto vload = [&](int k) {
    const double* vk = X + (long)k * nb;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      int r = lane + 32 * i;
      vn[i] = (r > k && r < nb) ? vk[r] : ((r == k) ? 1.0 : 0.0);
    }
    tn = tau[k];
  };
  vload(ns - 1);
  for (int k = ns - 1; k >= 0; --k) {
    double v[RPL];
    const double tk = tn;
#pragma unroll
    for (int i = 0; i < RPL; ++i) v[i] = vn[i];
    if (k > 0) vload(k - 1);   // prefetch: hides the L2 latency of the next reflector
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < RPL; ++i) dot = fma(v[i], y[i], dot);
    dot = wsum(dot);
    const double f = tk * dot;
#pragma unroll
    for (int i = 0; i < RPL; ++i) y[i] = fma(-f, v[i], y[i]);
  }
  double* out = Y + (long)i0 * nb;
#pragma unroll
  for (int i = 0; i < RPL; ++i) {
    int r = lane + 32 * i;
    if (r < nb) out[rowperm ? rowperm[r] : r] = y[i];
  }
}

// ---- column sort (cheap substitute for pivoting) + second QR --------------------------------
// perm0[rank] = column with the rank-th largest norm (stable).  One CTA.
__global__ void __launch_bounds__(1024)
svd_colsort_kernel(const double* __restrict__ X, int nb, int ns, double* __restrict__ nrm2, int* __restrict__ perm0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < ns; c += 32) {
    const double* x = X + (long)c * nb;
    double a = 0.0;
    for (int r = lane; r < nb; r += 32) a = fma(x[r], x[r], a);
    a = wsum(a);
    if (lane == 0) nrm2[c] = a;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < ns; c += blockDim.x) {
    const double v = nrm2[c];
    int rank = 0;
    for (int k = 0; k < ns; ++k) {
      const double u = nrm2[k];
      rank += (u > v || (u == v && k < c)) ? 1 : 0;
    }
    perm0[rank] = c;
  }
}

__global__ void svd_permute_cols_kernel(const double* __restrict__ src, int nb, int ns, const int* __restrict__ perm0,
                                        double* __restrict__ dst) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nb * ns) return;
  const int r = (int)(idx % nb), c = (int)(idx / nb);
  dst[idx] = src[(long)perm0[c] * nb + r];
}

// two-QR path scatter: both sides are unit vectors (Y: [k][nb] big side, Y2: [k][ns] small side)
__global__ void svd_scatter_qr2_kernel(const double* __restrict__ Y, const double* __restrict__ Y2,
                                       const double* __restrict__ sig2, const int* __restrict__ perm, SvdGeom sg,
                                       int isoIsA, int m, double* __restrict__ Wb, double* __restrict__ Wb1) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long nAm = (long)sg.nA * m, nBm = (long)sg.nB * m;
  if (idx >= nAm + nBm) return;
  const bool sideA = idx < nAm;
  long e = sideA ? idx : idx - nAm;
  int k, i;
  if (sideA) {
    int l = (int)(e % sg.nlA);
    long r = e / sg.nlA;
    k = (int)(r % m);
    int as = (int)(r / m);
    i = as * sg.nlA + l;
  } else {
    k = (int)(e / sg.nB);
    i = (int)(e % sg.nB);
  }
  const bool thisIsBig = (sideA == (sg.bigIsA != 0));
  const bool thisIsIso = (sideA == (isoIsA != 0));
  double v = thisIsBig ? Y[(long)k * sg.nb + i] : Y2[(long)k * sg.ns + i];
  if (!thisIsIso) v *= sqrt(sig2[perm[k]]);
  if (sideA)
    Wb[e] = v;
  else
    Wb1[e] = v;
}

// QR path scatter: big side = unit vectors Y[k][nb]; small side = columns of M (sigma-scaled)
__global__ void svd_scatter_qr_kernel(const double* __restrict__ Y, const double* __restrict__ M,
                                      const double* __restrict__ sig2, const int* __restrict__ perm, SvdGeom sg,
                                      int isoIsA, int m, double* __restrict__ Wb, double* __restrict__ Wb1) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long nAm = (long)sg.nA * m, nBm = (long)sg.nB * m;
  if (idx >= nAm + nBm) return;
  const bool sideA = idx < nAm;
  long e = sideA ? idx : idx - nAm;
  int k, i;
  if (sideA) {
    int l = (int)(e % sg.nlA);
    long r = e / sg.nlA;
    k = (int)(r % m);
    int as = (int)(r / m);
    i = as * sg.nlA + l;
  } else {
    k = (int)(e / sg.nB);
    i = (int)(e % sg.nB);
  }
  const int c = perm[k];
  const bool thisIsBig = (sideA == (sg.bigIsA != 0));
  const bool thisIsIso = (sideA == (isoIsA != 0));
  double v = thisIsBig ? Y[(long)k * sg.nb + i] : M[(long)c * sg.ns + i];
  const double sg1 = sqrt(sig2[c]);
  if (thisIsBig && !thisIsIso) v *= sg1;                       // unit -> sigma * unit
  if (!thisIsBig && thisIsIso) v = (sg1 > 0.0) ? v / sg1 : 0.0;  // sigma * unit -> unit
  if (sideA)
    Wb[e] = v;
  else
    Wb1[e] = v;
}

static int ensure(SvdWork& w, long nX, int ns) {
  // cudaMalloc / cudaFree synchronise the device and were measured at up to 1.4 s next to a large
  // stream-ordered pool: size every buffer for the largest bond matrix (2*maxm*NL x 2*maxm) at once
  if (w.hint_m > 0) {
    const long hs = 2L * w.hint_m;
    if (ns <= hs && nX <= hs * hs * NL) {
      ns = (int)std::max<long>(ns, hs);
      nX = std::max(nX, hs * hs * NL);
    }
  }
  if (nX > w.capX) {
    if (w.X) cudaFree(w.X);
    if (cudaMalloc(&w.X, nX * sizeof(double)) != cudaSuccess) return -1;
    w.capX = nX;
  }
  long nJ = (long)ns * ns;
  if (nJ > w.capJ) {
    if (w.J) cudaFree(w.J);
    if (cudaMalloc(&w.J, nJ * sizeof(double)) != cudaSuccess) return -1;
    w.capJ = nJ;
  }
  if (ns > w.capS) {
    if (w.sig2) cudaFree(w.sig2);
    if (w.perm) cudaFree(w.perm);
    if (cudaMalloc(&w.sig2, ns * sizeof(double)) != cudaSuccess) return -1;
    if (cudaMalloc(&w.perm, ns * sizeof(int)) != cudaSuccess) return -1;
    w.capS = ns;
  }
  if (!w.info) {
    if (cudaMalloc(&w.info, 8 * sizeof(double)) != cudaSuccess) return -1;
    if (cudaMalloc(&w.flags, 4 * sizeof(int)) != cudaSuccess) return -1;
    if (cudaMalloc(&w.sweepmax, 64 * sizeof(double)) != cudaSuccess) return -1;
  }
  if (nJ > w.capM) {   // QR path: R^T, tau, ready flags
    if (w.M) cudaFree(w.M);
    if (w.tau) cudaFree(w.tau);
    if (w.ready) cudaFree(w.ready);
    if (cudaMalloc(&w.M, nJ * sizeof(double)) != cudaSuccess) return -1;
    if (cudaMalloc(&w.tau, ns * sizeof(double)) != cudaSuccess) return -1;
    if (cudaMalloc(&w.ready, ns * sizeof(int)) != cudaSuccess) return -1;
    w.capM = nJ;
  }
  if (nX > w.capY) {
    if (w.Y) cudaFree(w.Y);
    if (cudaMalloc(&w.Y, nX * sizeof(double)) != cudaSuccess) return -1;
    w.capY = nX;
  }
  if (nJ > w.capM2) {   // two-QR path
    if (w.M2) cudaFree(w.M2);
    if (w.tau2) cudaFree(w.tau2);
    if (w.Y2) cudaFree(w.Y2);
    if (w.perm0) cudaFree(w.perm0);
    if (cudaMalloc(&w.M2, nJ * sizeof(double)) != cudaSuccess) return -1;
    if (cudaMalloc(&w.tau2, ns * sizeof(double)) != cudaSuccess) return -1;
    if (cudaMalloc(&w.Y2, nJ * sizeof(double)) != cudaSuccess) return -1;
    if (cudaMalloc(&w.perm0, ns * sizeof(int)) != cudaSuccess) return -1;
    w.capM2 = nJ;
  }
  return 0;
}

// Block one-sided Jacobi on the column-major matrix A (rows x ns), rotations accumulated
// in Jm (ns x ns, must hold the starting orthogonal matrix).  Returns 0 when converged.
static int jacobi_iterate(cudaStream_t st, SvdWork& w, double* A, double* Jm, int rows, int ns, long& nl) {
  // |cos angle| threshold.  LAPACK dgesvj uses sqrt(rows)*eps; the rounding noise of a
  // dot product sits right at that level and (with 28k pairs per sweep) stalls
  // convergence, so use 8x.
  const double tol = 8.0 * std::sqrt((double)rows) * 1.1102230246251565e-16;
  const double tol2 = tol * tol;
  // shared-memory resident path when 2*W staged columns (A and J parts) fit
  const size_t need16 = (size_t)2 * 16 * (rows + ns) * sizeof(double);
  const size_t need8 = (size_t)2 * 8 * (rows + ns) * sizeof(double);
  int Wd = 0;
  if (need16 <= 200 * 1024) Wd = 16;
  else if (need8 <= 200 * 1024) Wd = 8;
  static unsigned long long attr = 0;   // one bit per device: function attributes are per device
  if (first_on_device(attr)) {
    cudaFuncSetAttribute(jacobi_offdiag_smem_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(jacobi_offdiag_smem_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(jacobi_diag_smem_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(jacobi_diag_smem_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  // Gram-based kernel: 2W staged columns of length rows+ns plus three (2W) x 36 matrices.
  // W = 8 by default: the round chain costs ~ W^3 per launch while a launch covers ~ W^2 pairs,
  // and more (smaller) CTAs run concurrently.
  const int gld = ((rows + ns + 15) / 16) * 16 + 4;
  static int use_gram = -1, gram_w = 8;
  if (use_gram < 0) {
    const char* e = getenv("TNML_SVD_GRAM");
    use_gram = e ? atoi(e) : 1;
    const char* ew = getenv("TNML_SVD_GRAM_W");
    if (ew && atoi(ew) == 16) gram_w = 16;
  }
  static unsigned long long attr_gram = 0;
  if (first_on_device(attr_gram)) {
    cudaFuncSetAttribute(jacobi_gram_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(jacobi_gram_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  }
  const int GW = gram_w;
  const size_t need_gram = ((size_t)2 * GW * gld + 3 * 2 * GW * GLD + 4 * GW) * sizeof(double) +
                           (size_t)(2 * GW - 1) * GW * sizeof(unsigned short) + (size_t)(2 * GW - 1) * 2 * GW + 16;
  const bool gram = use_gram && need_gram <= 220 * 1024 && ns > GW && (rows % 2 == 0) && (ns % 2 == 0);
  const int bw = gram ? GW : (Wd ? Wd : JW);
  const int nblk = (ns + bw - 1) / bw;
  const int nblk_e = (nblk % 2) ? nblk + 1 : nblk;
  // gram: largest cos^2 rotated away during the sweep.  Cyclic Jacobi converges quadratically, so a
  // sweep whose largest angle had |cos| <= 1e-8 leaves every pair orthogonal to ~1e-16: no extra
  // "clean" sweep is needed to confirm (measured on bond matrices: ... 1e-5, 1e-10, 0).
  // smem kernels: off^2.
  static double gram_conv = -1.0;
  if (gram_conv < 0.0) {
    const char* e = getenv("TNML_SVD_STOP");
    gram_conv = e ? atof(e) : 1e-16;
  }
  const double conv = gram ? gram_conv : (Wd ? tol2 : tol);
  const int max_sweeps = 60;
  int hflag = 0;
  // ---- cluster-resident path: all sweeps in one launch of a 16-CTA cluster (<= 256 columns)
  // per handle (= per device): a device whose cluster launch is refused falls back alone
  int& use_cluster = w.cluster_ok;
  if (use_cluster < 0) {
    const char* e = getenv("TNML_SVD_CLUSTER");
    use_cluster = e ? atoi(e) : 1;
    if (use_cluster) {
      if (cudaFuncSetAttribute(jacobi_cluster_kernel<8, 512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
          cudaFuncSetAttribute(jacobi_cluster_kernel<8, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
        cudaGetLastError();
        use_cluster = 0;
      }
    }
  }
  const bool want_cluster = (g_svd_cluster >= 0) ? (g_svd_cluster != 0 && use_cluster >= 0 && use_cluster != 0) : (use_cluster != 0);
  if (gram && want_cluster && GW == 8 && nblk_e <= 32) {
    const size_t need_cl = ((size_t)2 * 16 * gld + 5 * 16 * GLD + 4 * 8) * sizeof(double) + 16 * sizeof(double*) +
                           (size_t)(15 + 8) * 8 * sizeof(unsigned short) + (size_t)(15 + 8) * 16 +
                           (size_t)(14 + 7) * 8 * sizeof(unsigned) + 32;
    if (need_cl <= 220 * 1024) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(16, 1, 1);
      cfg.blockDim = dim3(512, 1, 1);
      cfg.dynamicSmemBytes = need_cl;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 16;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (!w.cluster_checked) {
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, jacobi_cluster_kernel<8, 512>, &cfg) != cudaSuccess || ncl < 1) {
          cudaGetLastError();
          use_cluster = 0;
        }
        w.cluster_checked = 1;
      }
      if (use_cluster) {
        if (cudaMemsetAsync(w.sweepmax, 0, 64 * sizeof(double), st) != cudaSuccess) return -2;
        static int cross_only = -1;
        if (cross_only < 0) {
          const char* e = getenv("TNML_SVD_CROSS");
          cross_only = e ? atoi(e) : 1;
        }
        const int cross_now = (g_svd_cross >= 0) ? g_svd_cross : cross_only;
        if (cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel<8, 512>, A, Jm, rows, ns, gld, nblk_e, tol2, conv, max_sweeps,
                               cross_now, w.info, w.sweepmax, w.flags) == cudaSuccess) {
          nl += 2;
          return 1;   // convergence flag is in info[6]: the caller reads it with the truncation results (one sync)
        }
        cudaGetLastError();   // cluster launch refused (e.g. no GPC with 16 free SMs): multi-launch path from now on
        use_cluster = 0;
      }
    }
  }
  for (int sw = 0; sw < max_sweeps && !hflag; ++sw) {
    if (gram) {
      // one sweep = nblk_e-1 launches + bookkeeping; replayed as a CUDA graph (the launch gaps of
      // ~300 back-to-back 10 us kernels were ~0.8 ms per SVD)
      const long key[6] = {(long)(size_t)A, (long)(size_t)Jm, rows, ns, GW, (long)(size_t)w.info};
      int gi = -1, lru = 0;
      for (int k = 0; k < SvdWork::NGRAPH; ++k) {
        bool eq = w.gexec[k] != nullptr;
        for (int i = 0; i < 6 && eq; ++i) eq = (w.gkey[k][i] == key[i]);
        if (eq) {
          gi = k;
          break;
        }
        if (w.gstamp[k] < w.gstamp[lru]) lru = k;
      }
      bool same = gi >= 0;
      if (!same) gi = lru;
      w.gstamp[gi] = ++w.gclock;
      cudaGraphExec_t& gexec = w.gexec[gi];
      long* gkey = w.gkey[gi];
      static int use_graph = -1;
      if (use_graph < 0) use_graph = getenv("TNML_SVD_NOGRAPH") ? 0 : 1;
      auto enqueue = [&]() {
        for (int R = 0; R < nblk_e - 1; ++R) {
          if (GW == 16)
            jacobi_gram_kernel<16><<<nblk_e / 2, 512, need_gram, st>>>(A, Jm, rows, ns, gld, nblk_e, R, tol2, w.info,
                                                                       w.flags);
          else
            jacobi_gram_kernel<8><<<nblk_e / 2, 256, need_gram, st>>>(A, Jm, rows, ns, gld, nblk_e, R, tol2, w.info,
                                                                      w.flags);
        }
        jacobi_sweep_end_kernel<<<1, 32, 0, st>>>(w.info, w.flags, conv);
      };
      if (use_graph) {
        if (!same) {
          if (gexec) cudaGraphExecDestroy(gexec);
          gexec = nullptr;
          cudaGraph_t graph;
          if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            enqueue();
            if (cudaStreamEndCapture(st, &graph) == cudaSuccess && graph) {
              if (cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) gexec = nullptr;
              cudaGraphDestroy(graph);
            }
          }
          for (int i = 0; i < 6; ++i) gkey[i] = gexec ? key[i] : 0;
        }
        if (gexec && cudaGraphLaunch(gexec, st) != cudaSuccess) return -2;
        if (!gexec) enqueue();
      } else {
        enqueue();
      }
      nl += nblk_e;
      if (sw >= 3) {
        if (cudaMemcpyAsync(&hflag, w.flags, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -2;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
      }
      continue;
    }
    if (Wd == 16) {
      jacobi_diag_smem_kernel<16><<<nblk, 16 * 16, need16 / 2, st>>>(A, Jm, rows, ns, tol2, w.info, w.flags);
    } else if (Wd == 8) {
      jacobi_diag_smem_kernel<8><<<nblk, 16 * 8, need8 / 2, st>>>(A, Jm, rows, ns, tol2, w.info, w.flags);
    } else {
      jacobi_diag_kernel<<<nblk, 256, 0, st>>>(A, Jm, rows, ns, tol, w.info, w.flags);
    }
    nl += 1;
    if (nblk > 1) {
      for (int R = 0; R < nblk_e - 1; ++R) {
        if (Wd == 16)
          jacobi_offdiag_smem_kernel<16><<<nblk_e / 2, 32 * 16, need16, st>>>(A, Jm, rows, ns, nblk_e, R, tol2, w.info,
                                                                            w.flags);
        else if (Wd == 8)
          jacobi_offdiag_smem_kernel<8><<<nblk_e / 2, 32 * 8, need8, st>>>(A, Jm, rows, ns, nblk_e, R, tol2, w.info,
                                                                          w.flags);
        else
          jacobi_offdiag_kernel<<<nblk_e / 2, 512, 0, st>>>(A, Jm, rows, ns, nblk_e, R, tol, w.info, w.flags);
        nl += 1;
      }
    }
    jacobi_sweep_end_kernel<<<1, 32, 0, st>>>(w.info, w.flags, conv);
    nl += 1;
    if (sw >= 3 || nblk == 1) {
      if (cudaMemcpyAsync(&hflag, w.flags, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -2;
      if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
    }
  }
  return hflag ? 0 : -5;
}

template <int RPL>
static void launch_qr(cudaStream_t st, SvdWork& w, double* Xq, double* tau, int nb, int ns) {
  qr_dataflow_kernel<RPL><<<(ns + 7) / 8, 256, 0, st>>>(Xq, nb, ns, tau, w.ready);
}
static void launch_qr_block8(cudaStream_t st, SvdWork& w, double* Xq, double* tau, int nb, int ns) {
  static unsigned long long attr = 0;   // one bit per device: function attributes are per device
  if (first_on_device(attr)) {
    cudaFuncSetAttribute(qr_block_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  }
  size_t sh = (size_t)(32L * nb + 32) * sizeof(double) + 32 * sizeof(int);
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    dbg_on = getenv("TNML_QR_DEBUG") ? 1 : 0;
    if (dbg_on) {
      cudaMalloc(&dbg, 4096 * 4 * sizeof(long long));
      cudaMemcpyToSymbol(g_qr_dbg, &dbg, sizeof(dbg));
    }
  }
  qr_block_kernel<8><<<(ns + 31) / 32, 1024, sh, st>>>(Xq, nb, ns, tau, w.ready);
  if (dbg_on) {
    cudaStreamSynchronize(st);
    std::vector<long long> hd(ns * 4);
    cudaMemcpy(hd.data(), dbg, ns * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long gs[6], rs[5];
    cudaMemcpy(gs, dbg + 2048, sizeof(gs), cudaMemcpyDeviceToHost);
    cudaMemcpy(rs, dbg + 2056, sizeof(rs), cudaMemcpyDeviceToHost);
    {
      FILE* fg = fopen("gpurun_out/gram_debug.txt", "w");
      if (fg) {
        fprintf(fg, "gram kernel phases (cycles): stage %lld gram %lld rounds %lld apply %lld store %lld\n", gs[1] - gs[0],
                gs[2] - gs[1], gs[3] - gs[2], gs[4] - gs[3], gs[5] - gs[4]);
        long long cs[6];
        cudaMemcpy(cs, dbg + 2064, sizeof(cs), cudaMemcpyDeviceToHost);
        if (cs[5] > 0)
          fprintf(fg, "cluster kernel, CTA 0, cycles per block pairing over %lld pairings: setup %lld gram %lld rounds %lld "
                      "apply %lld cluster barrier %lld\n", cs[5], cs[0] / cs[5], cs[1] / cs[5], cs[2] / cs[5], cs[3] / cs[5],
                  cs[4] / cs[5]);
        fclose(fg);
      }
    }
    FILE* f = fopen("gpurun_out/qr_debug.txt", "w");
    if (f) {
      fprintf(f, "(gram stamps are written by the LAST gram launch of the previous svd)\n");
      for (int jj = 0; jj < ns; ++jj)
        fprintf(f, "%d sm=%lld wait_end=%lld apply_end=%lld publish=%lld\n", jj, hd[jj * 4 + 3], hd[jj * 4 + 0], hd[jj * 4 + 1],
                hd[jj * 4 + 2]);
      fclose(f);
    }
  }
}
template <int RPL>
static void launch_apply_q(cudaStream_t st, const double* Xq, const double* tau, const double* src, const int* perm,
                           int nb, int ns, int m, double* Yout, const double* scale_sig2, const int* rowperm) {
  apply_q_kernel<RPL><<<(m + 7) / 8, 256, 0, st>>>(Xq, tau, src, perm, nb, ns, m, Yout, scale_sig2, rowperm);
}
// Householder QR of the column-major nb x ns matrix Xq in place (reflectors below the diagonal)
static int run_qr(cudaStream_t st, SvdWork& w, double* Xq, double* tau, int nb, int ns) {
  if (cudaMemsetAsync(w.ready, 0, ns * sizeof(int), st) != cudaSuccess) return -2;
  if (nb <= 32 * 8) launch_qr_block8(st, w, Xq, tau, nb, ns);
  else if (nb <= 32 * 20) launch_qr<20>(st, w, Xq, tau, nb, ns);
  else if (nb <= 32 * 40) launch_qr<40>(st, w, Xq, tau, nb, ns);
  else {
    static unsigned long long attr = 0;   // one bit per device: function attributes are per device
    if (first_on_device(attr)) {
      cudaFuncSetAttribute(qr_dataflow_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    // columns per CTA: as many of 8 as fit in 200 KB (20 m rows at m = 300: 4 columns of 48 KB)
    int wpb = (int)((200 * 1024) / ((size_t)nb * sizeof(double)));
    wpb = wpb > 8 ? 8 : (wpb < 1 ? 1 : wpb);
    qr_dataflow_smem_kernel<<<(ns + wpb - 1) / wpb, 32 * wpb, (size_t)wpb * nb * sizeof(double), st>>>(Xq, nb, ns, tau,
                                                                                                    w.ready);
  }
  return 0;
}
static void run_apply_q(cudaStream_t st, const double* Xq, const double* tau, const double* src, const int* perm,
                        int nb, int ns, int m, double* Yout, const double* scale_sig2, const int* rowperm) {
  if (nb <= 32 * 8) launch_apply_q<8>(st, Xq, tau, src, perm, nb, ns, m, Yout, scale_sig2, rowperm);
  else if (nb <= 32 * 20) launch_apply_q<20>(st, Xq, tau, src, perm, nb, ns, m, Yout, scale_sig2, rowperm);
  else if (nb <= 32 * 40) launch_apply_q<40>(st, Xq, tau, src, perm, nb, ns, m, Yout, scale_sig2, rowperm);
  else {
    static unsigned long long attr = 0;   // one bit per device: function attributes are per device
    if (first_on_device(attr)) {
      cudaFuncSetAttribute(apply_q_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    int wpb = (int)((200 * 1024) / ((size_t)nb * sizeof(double)));
    wpb = wpb > 8 ? 8 : (wpb < 1 ? 1 : wpb);
    apply_q_smem_kernel<<<(m + wpb - 1) / wpb, 32 * wpb, (size_t)wpb * nb * sizeof(double), st>>>(Xq, tau, src, perm, nb, ns,
                                                                                              m, Yout, scale_sig2, rowperm);
  }
}

int svd_split(cudaStream_t st, SvdWork& w, const double* Bc, BondGeom g, int dir, double cutoff, int maxm,
              int minm, int do_rel_cutoff, double* Wb_out, double* Wb1_out, int* newm, double* truncerr,
              int* sweeps, long* launches) {
  SvdGeom sg;
  sg.g = g;
  sg.nlA = g.lab_b ? NL : 1;
  sg.nlB = g.lab_b1 ? NL : 1;
  sg.nA = 2 * g.ml * sg.nlA;
  sg.nB = 2 * g.mr * sg.nlB;
  sg.bigIsA = (sg.nA >= sg.nB) ? 1 : 0;
  sg.nb = sg.bigIsA ? sg.nA : sg.nB;
  sg.ns = sg.bigIsA ? sg.nB : sg.nA;
  const int ns = sg.ns, nb = sg.nb;
  if (ensure(w, (long)nb * ns, ns) != 0) return -2;
  long nl = 0;
  if (w.use_qr < 0) {
    const char* e = getenv("TNML_SVD_QR");
    w.use_qr = e ? atoi(e) : 2;   // 2 = auto (sort + two QRs from 32 columns on)
  }
  // QR preconditioning pays off from a few dozen columns on; it needs one resident warp per
  // column and the column in registers (nb <= 32*96)
  const int use_qr = (g_svd_precond >= 0) ? g_svd_precond : w.use_qr;
  // (the tall columns of a class-C bond live in shared memory: one column must fit in 200 KB)
  const bool qr = (use_qr == 1 || use_qr == 3 || (use_qr == 2 && ns >= 32)) && (size_t)nb * sizeof(double) <= 200 * 1024 &&
                  ns <= 8 * 140;

  long nJ = (long)ns * ns;
  long ninit = nJ > 8 ? nJ : 8;
  long nX = (long)nb * ns;
  // 3 = column sort + two QRs (X P0 = Q1 R1, R1^T = Q2 R2, Jacobi on R2^T): about half the Jacobi
  // sweeps of the one-QR path on freshly optimised bond matrices (tools/jacobi_precond_study.py)
  const bool qr2 = qr && (use_qr == 2 || use_qr == 3);
  svd_init_kernel<<<(unsigned)((ninit + 255) / 256), 256, 0, st>>>(w.J, ns, w.info, w.flags);
  nl += 1;
  int rc;
  if (qr2) {
    svd_gather_kernel<<<(unsigned)((nX + 255) / 256), 256, 0, st>>>(Bc, sg, w.Y);
    svd_colsort_kernel<<<1, 1024, 0, st>>>(w.Y, nb, ns, w.sig2, w.perm0);
    svd_permute_cols_kernel<<<(unsigned)((nX + 255) / 256), 256, 0, st>>>(w.Y, nb, ns, w.perm0, w.X);
    if (run_qr(st, w, w.X, w.tau, nb, ns) != 0) return -2;
    rt_form_kernel<<<(unsigned)((nJ + 255) / 256), 256, 0, st>>>(w.X, nb, ns, w.M);
    if (run_qr(st, w, w.M, w.tau2, ns, ns) != 0) return -2;
    rt_form_kernel<<<(unsigned)((nJ + 255) / 256), 256, 0, st>>>(w.M, ns, ns, w.M2);
    nl += 9;
    rc = jacobi_iterate(st, w, w.M2, w.J, ns, ns, nl);
    if (rc == -2) return rc;
    svd_finalize_kernel<<<1, 1024, 0, st>>>(w.M2, ns, ns, w.sig2, w.perm, cutoff, maxm, minm, do_rel_cutoff, w.info);
  } else if (qr) {
    svd_gather_kernel<<<(unsigned)((nX + 255) / 256), 256, 0, st>>>(Bc, sg, w.X);
    if (run_qr(st, w, w.X, w.tau, nb, ns) != 0) return -2;
    rt_form_kernel<<<(unsigned)((nJ + 255) / 256), 256, 0, st>>>(w.X, nb, ns, w.M);
    nl += 4;
    rc = jacobi_iterate(st, w, w.M, w.J, ns, ns, nl);
    if (rc == -2) return rc;
    svd_finalize_kernel<<<1, 1024, 0, st>>>(w.M, ns, ns, w.sig2, w.perm, cutoff, maxm, minm, do_rel_cutoff, w.info);
  } else {
    svd_gather_kernel<<<(unsigned)((nX + 255) / 256), 256, 0, st>>>(Bc, sg, w.X);
    nl += 1;
    rc = jacobi_iterate(st, w, w.X, w.J, nb, ns, nl);
    if (rc == -2) return rc;
    svd_finalize_kernel<<<1, 1024, 0, st>>>(w.X, nb, ns, w.sig2, w.perm, cutoff, maxm, minm, do_rel_cutoff, w.info);
  }
  nl += 1;
  double hinfo[8];
  if (cudaMemcpyAsync(hinfo, w.info, sizeof(hinfo), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -2;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
  if (rc == 1) rc = (hinfo[6] != 0.0) ? 0 : -5;   // cluster-resident Jacobi: deferred convergence check
  const int m = (int)hinfo[1];
  *newm = m;
  *truncerr = hinfo[2];
  *sweeps = (int)hinfo[3];
  const int isoIsA = (dir == 1) ? 1 : 0;
  long nout = (long)(sg.nA + sg.nB) * m;
  if (qr2) {
    // X P0 = Q1 W Sigma (Q2 J)^T with W Sigma = the rotated columns of R2^T:
    // big side = Q1 [W; 0], small side = P0 (Q2 J)
    // the two applications are independent (15 CTAs each at m = 120): run them side by side
    if (!w.st2) {
      if (cudaStreamCreateWithFlags(&w.st2, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&w.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&w.ev_join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        w.st2 = nullptr;
      }
    }
    if (w.st2) {
      cudaEventRecord(w.ev_fork, st);
      cudaStreamWaitEvent(w.st2, w.ev_fork, 0);
      run_apply_q(w.st2, w.M, w.tau2, w.J, w.perm, ns, ns, m, w.Y2, nullptr, w.perm0);
      cudaEventRecord(w.ev_join, w.st2);
      run_apply_q(st, w.X, w.tau, w.M2, w.perm, nb, ns, m, w.Y, w.sig2, nullptr);
      cudaStreamWaitEvent(st, w.ev_join, 0);
    } else {
      run_apply_q(st, w.X, w.tau, w.M2, w.perm, nb, ns, m, w.Y, w.sig2, nullptr);
      run_apply_q(st, w.M, w.tau2, w.J, w.perm, ns, ns, m, w.Y2, nullptr, w.perm0);
    }
    svd_scatter_qr2_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(w.Y, w.Y2, w.sig2, w.perm, sg, isoIsA, m,
                                                                        Wb_out, Wb1_out);
    nl += 3;
  } else if (qr) {
    run_apply_q(st, w.X, w.tau, w.J, w.perm, nb, ns, m, w.Y, nullptr, nullptr);
    svd_scatter_qr_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(w.Y, w.M, w.sig2, w.perm, sg, isoIsA, m,
                                                                       Wb_out, Wb1_out);
    nl += 2;
  } else {
    svd_scatter_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(w.X, w.J, w.sig2, w.perm, sg, isoIsA, m,
                                                                      Wb_out, Wb1_out);
    nl += 1;
  }
  if (hinfo[7] > 0.0) {   // zero singular values among the kept ones: complete the isometry like ITensor does
    const int nrows = isoIsA ? sg.nA : sg.nB;
    const size_t sh = (size_t)(nrows + m + 64 + 32) * sizeof(double);
    if (sh <= 200 * 1024) {
      static unsigned long long attr = 0;
      if (first_on_device(attr))
        cudaFuncSetAttribute(svd_complete_iso_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      svd_complete_iso_kernel<<<1, 1024, sh, st>>>(isoIsA ? Wb_out : Wb1_out, nrows, m, isoIsA ? sg.nlA : 1, isoIsA);
      nl += 1;
    }
  }
  if (launches) *launches += nl;
  if (cudaGetLastError() != cudaSuccess) return -2;
  return rc;
}

}  // namespace tnml
