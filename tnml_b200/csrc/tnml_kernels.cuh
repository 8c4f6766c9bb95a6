// tnml_kernels.cuh -- launch wrappers of the sm_100a kernels behind the C-ABI.
// All arithmetic is float64: the reference computes in ITensor `Real` = double
// and the CG of fixedL.cc:349-445 is too ill-conditioned for less (DESIGN.md
// "Precision").
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tnml {

constexpr int NL = 10;

// cudaFuncSetAttribute is per device: opt-ins are done once per device of the process (one handle
// per GPU may live in the same process), tracked in a 64-bit mask owned by the call site.
inline bool first_on_device(unsigned long long& mask) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// ---- Khatri-Rao GEMM family -------------------------------------------------
// krgemm: Out[row][j] = sum_{a,p} In[row][a] * w_p(img(row)) * Bm[(a*S+p)*ldb + j],  img(row) = row/div
//   S = 2: w = (f1_0, f1_1) ; S = 4: w_{s*2+q} = f1_s * f2_q.  K = S*ma; the Khatri-Rao operand
//   l_n (x) phi_n (x) phi'_n is generated while it is staged.
// Used for: projected input Q (fixedL.cc:318,377,399,416 restructured) and
// environment advance (fixedL.cc:144-150, 221-229).
void krgemm(cudaStream_t st, int S, const double* In, long ldin, int ma, const double* f1, const double* f2,
            int div, const double* Bm, long ldb, int J, double* Out, long ldout, long rows, int num_sm);
// 2 (default): persistent kernel with output-side Khatri-Rao weights, cp.async-staged A tiles and a
// resident B panel (from 512 rows on, while the panel fits); 1: register-staged kernel only
void krgemm_set_variant(int v);

// ---- the same projection on tcgen05 (tnml_ozaki.cu): int8 error-free splitting, TMA, TMEM -------------
// Operands are cut into `ns` signed 7-bit planes once (rows of In: per environment, i.e. once per
// bond; columns of Bm: per CG pass) and multiplied exactly by tcgen05.mma kind::i8.
bool oz_supported(int S, int ma, int ns);                 // S in {2,4}, ma <= 128, ns in 6..8
size_t oz_a8_bytes(long rows, int ns);                    // plane buffer of the row operand
long oz_rows_pad(long rows);                              // entries of ea[]
size_t oz_b8_bytes(int S, int J, int ns);                 // plane buffer of the column operand
long oz_cols_pad(int S, int J);                           // entries of eb[]
void oz_slice_rows(cudaStream_t st, const double* In, long ldin, int ma, long rows, int ns, int8_t* A8, double* ea);
void oz_slice_cols(cudaStream_t st, int S, const double* Bm, long ldb, int ma, int J, int ns, int8_t* B8, double* eb);
// Out[row][j] = sum_p w_p(row/div) sum_a In[row][a] Bm[(a*S+p)*ldb + j]; false: not launched
bool oz_krgemm(cudaStream_t st, int S, int ns, const int8_t* A8, const double* ea, long rows, const double* f1,
               const double* f2, int div, const int8_t* B8, const double* eb, int J, double* Out, long ldout,
               int num_sm);

void oz_set_debug_buffer(long long* dev_buf);               // -DOZ_PROFILE builds only: [CTA][8] cycle counters

// krgram: Gpart[split][(a*S+p)][j] = sum_{row in split} In[row][a] * w_p(row) * Z[row][j]
// The rank-1 gradient accumulation of fixedL.cc:379,418 as one K=NT contraction, split-K over
// CTAs, partials reduced in a fixed order (deterministic).
int krgram_splits(int ma, int S, int J, long rows, int num_sm);
void krgram(cudaStream_t st, int S, const double* In, long ldin, int ma, const double* f1, const double* f2,
            const double* Z, long ldz, int J, double* Gpart, long rows, int nsplit);
// 2 (default): cp.async-staged raw tiles, weights applied at fragment load; 1: register-staged kernel
void krgram_set_variant(int v);
// G[i] = sum_split Gpart[split][i]   (+ optional: G[i] -= lambda*B[i])
void reduce_partials(cudaStream_t st, const double* Gpart, int nsplit, long n, double* G);

// ---- fat-environment kernel -------------------------------------------------
// One warp per image.  P[n][l] = sum_f Q[n][f]*F[n][l][f].
enum FatMode : int {
  FAT_GRAD = 0,   // dP = delta_{label} - P ; stats ; Z[n][f] = sum_l dP[l]*F[n][l][f]
  FAT_PAP = 1,    // stats[11] += sum_l P[l]^2
  FAT_COST = 2,   // stats (cost per label, ncorrect), pred
  FAT_GRAD_OUTER = 3,  // class C: Zfat[n][l][f] = dP[l]*Q[n][f]
  FAT_BWD = 4,         // P[n][l] is given (linear update of the forward outputs): stats + Z only
  FAT_BWD_OUTER = 5,   // class C variant of FAT_BWD
};
// stats layout (double[16]): [0..9] cost per label, [10] ncorrect, [11] sum |P|^2
// The statistics of the pass are reduced inside the launch (last CTA, fixed order) into stats_out[16];
// `ticket` is a zero-initialised device counter owned by the caller.
void fat_kernel(cudaStream_t st, int mode, const double* Q, const double* F, int m, const int32_t* labels,
                double* P, double* Z, int32_t* pred, double* stats_partial, int nblocks, long NT, double* stats_out,
                unsigned* ticket);
int fat_blocks(int num_sm);
// 1 (default): register-resident kernel; 2: bulk-async-copy (TMA, cp.async.bulk + mbarrier) double-buffered
// kernel for m <= 128 from 1024 images on (measured slower: 4.77 vs 5.06 TB/s)
void fat_set_variant(int v);

// ---- small dense tensor helpers ---------------------------------------------
// geometry of a bond-shaped tensor in its device ("canonical") layout
struct BondGeom {
  int ml, mr, nl;          // nl = 1 or NL
  int lab_b, lab_b1;       // label on site b / on site b+1
  long sa, ss, st, sb, sl; // strides of (alpha, s, t, beta, label)
  __host__ __device__ long size() const { return (long)ml * 4 * mr * nl; }
};
// B[alpha,s,t,beta(,l)] = sum_m Wb[alpha,s,m(,l)] * Wb1[m,t,beta(,l)]   (fixedL.cc:494)
void form_bond(cudaStream_t st, const double* Wb, const double* Wb1, int m, BondGeom g, double* B);
// host layout [ml][2][2][mr][nl] <-> canonical
void bond_to_host_layout(cudaStream_t st, const double* Bc, BondGeom g, double* Bh);
void bond_from_host_layout(cudaStream_t st, const double* Bh, BondGeom g, double* Bc);

// site tensor W[a][s][b](,[l]) -> GEMM operand for an environment advance:
//  left : Bm[(a*2+s)][(l*mb + b)]  = W[a][s][b][l]
//  right: Bm[(b*2+s)][(l*ma + a)]  = W[a][s][b][l]
void permute_site(cudaStream_t st, const double* W, int ma, int mb, int nl, int right, double* Bm);

void fill(cudaStream_t st, double* x, long n, double v);
// y = a*x + b*y
void axpby(cudaStream_t st, long n, double a, const double* x, double b, double* y);
// y = (*a_ptr)*x + b*y with the scalar read from device memory
void axpby_dev(cudaStream_t st, long n, const double* a_ptr, const double* x, double b, double* y);
// y = x + (*b_ptr)*y
void xpby_dev(cudaStream_t st, long n, const double* x, const double* b_ptr, double* y);
// scalars of the CG recurrence on the device (fixedL.cc:388-443); layout of cgs[32] in tnml_kernels.cu
void cg_begin(cudaStream_t st, const double* tail, double* cgs);
void cg_step(cudaStream_t st, double* cgs, const double* pAp, double lambda, const double* pp);
void cg_after_grad(cudaStream_t st, const double* tail, double* cgs, double lambda, const double* bb, double NTg,
                   double cconv);
// out[0] = sum x*y (deterministic two-stage)
void dot(cudaStream_t st, long n, const double* x, const double* y, double* scratch, double* out);

// ---- truncated SVD (tnml_svd.cu) ----------------------------------------------
struct SvdWork {
  double* X = nullptr;     // [small][big]  column-major working matrix (columns = "small" index)
  double* J = nullptr;     // [small][small] accumulated rotations
  double* sig2 = nullptr;  // [small]
  int* perm = nullptr;     // [small]
  double* info = nullptr;  // [8]: 0 maxoff of last sweep, 1 newm, 2 truncerr, 3 sweeps, 4 sum sig2
  int* flags = nullptr;    // [4]: 0 converged
  double* sweepmax = nullptr;  // [64] largest rotated cos^2 per sweep (cluster-resident Jacobi)
  long capX = 0, capJ = 0;
  int capS = 0;
  // QR-preconditioned path
  double* M = nullptr;     // [ns][ns] R^T, then its rotated columns
  double* tau = nullptr;   // [ns]
  int* ready = nullptr;    // [ns] dataflow flags
  double* Y = nullptr;     // [kept][nb] big-side unit vectors
  long capM = 0, capY = 0;
  // second QR (R1^T = Q2 R2, Jacobi on R2^T) + column sort
  double* M2 = nullptr;    // [ns][ns] R2^T, then its rotated columns
  double* tau2 = nullptr;  // [ns]
  double* Y2 = nullptr;    // [kept][ns] small-side unit vectors (before the column permutation)
  int* perm0 = nullptr;    // [ns] columns of X sorted by decreasing norm
  long capM2 = 0;
  int use_qr = -1;         // -1 auto, 0 never, 1 one QR, 3 sort + two QRs (TNML_SVD_QR)
  int hint_m = 0;          // largest link dimension expected (maxm): buffers are sized for it at once
  int cluster_ok = -1;     // cluster-resident Jacobi usable on this handle's device (-1: not probed yet)
  int cluster_checked = 0;
  cudaStream_t st2 = nullptr;              // side stream: the two apply-Q launches run concurrently
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // one Jacobi sweep captured as a CUDA graph, one executable per (buffers, dims) seen -- in a real
  // sweep the bond matrix has a different size at almost every bond, and re-instantiating the
  // graph every time costs more than it saves
  static constexpr int NGRAPH = 96;
  cudaGraphExec_t gexec[NGRAPH] = {};
  long gkey[NGRAPH][6] = {};
  long gstamp[NGRAPH] = {};
  long gclock = 0;
};
// variant switches for tests / A-B timing: "svd_cluster" (0: multi-launch Jacobi), "svd_cross" (0: full
// inner tournaments), "svd_precond" (1: one QR, 3: column sort + two QRs, 0: none)
void svd_set_variant(const char* what, int v);
// Gather canonical B into X (tall orientation), run block one-sided Jacobi,
// sort, apply ITensor's truncation rule, scatter U -> W(c), S*V -> W(c+dc).
// dir = 1: rows are (alpha,s[,l]) ; dir = 2: rows are (t,beta[,l]).
// Wb_out / Wb1_out must hold 2*ml*maxkeep*nl and maxkeep*2*mr*nl doubles.
// Returns (via host sync) newm, truncerr, sweeps.
int svd_split(cudaStream_t st, SvdWork& w, const double* Bc, BondGeom g, int dir, double cutoff, int maxm,
              int minm, int do_rel_cutoff, double* Wb_out, double* Wb1_out, int* newm, double* truncerr,
              int* sweeps, long* launches);

}  // namespace tnml
