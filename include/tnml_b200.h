/*
 * tnml_b200.h -- C-ABI of the B200-native replacement for TNML's `fixedL`
 * per-bond hot path (reference: /root/reference/fixedL.cc, paralleldo.h).
 *
 * The reference has no FFI layer: fixedL.cc calls ITensor directly.  The seam
 * is placed where `mldmrg` (fixedL.cc:451-570) calls TrainStates / cgrad /
 * quadcost / svd, one entry point per phase; each declaration cites the
 * reference code it replaces.  Plain pointers and sizes only; no exceptions
 * cross the boundary.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; the message is
 *     available from tnml_last_error(h) (or tnml_last_error(NULL) for
 *     create-time failures).  The reference's `Error(...)` throws
 *     (fixedL.cc:88,295,362,505,734) map to negative codes.
 *   - sites and bonds are 1-indexed like the reference (sites 1..N, bond b
 *     joins sites b and b+1); the label index lives on site jc = N/2
 *     (fixedL.cc:616) and has dimension TNML_NL = 10 (fixedL.cc:15).
 *   - host tensors are row-major float64:
 *       site tensor   [ml][d][mr]            (label site: [ml][d][mr][NL])
 *       bond tensor   [ml][d][d][mr]         (touching jc: [ml][d][d][mr][NL])
 *   - the caller owns host buffers; the library owns all device memory.
 *   - one driver thread per handle; one handle drives one GPU (one process
 *     per GPU, like the shards of ParallelDo, paralleldo.h:32-43); calls are
 *     stream-ordered, getters synchronise.
 *   - there is NO CPU fallback: without a CUDA device tnml_create fails.
 */
#ifndef TNML_B200_H
#define TNML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNML_NL 10          /* fixedL.cc:15  const size_t NL = 10 */
#define TNML_D 2            /* fixedL.cc:586 int d = 2           */

#define TNML_OK 0
#define TNML_ERR_INVALID (-1)   /* bad argument / bad state (reference: Error()) */
#define TNML_ERR_CUDA (-2)      /* CUDA runtime error                           */
#define TNML_ERR_NODEVICE (-3)  /* no CUDA device: the product has no CPU path  */
#define TNML_ERR_NCCL (-4)      /* NCCL error / library not loadable            */
#define TNML_ERR_NOCONV (-5)    /* Jacobi SVD did not converge                  */

#define TNML_FROMLEFT 1         /* ITensor Direction Fromleft  (ha==1) */
#define TNML_FROMRIGHT 2        /* ITensor Direction Fromright (ha==2) */

typedef struct tnml_handle_s* tnml_handle;

/* Library / build identification ("tnml_b200 <ver> sm_100a ..."). */
const char* tnml_version(void);
const char* tnml_last_error(tnml_handle h);

/* Create the per-GPU state on CUDA device `device`.  Replaces the TrainStates
 * constructor (fixedL.cc:76-96); `flags` is reserved (0). */
int tnml_create(int device, int flags, tnml_handle* out);
int tnml_destroy(tnml_handle h);

/* Training images of THIS shard.  `feat` is TState::data for each image,
 * feat[(n*N + (j-1))*d + (k-1)] = phi(img_n(j),k)  (fixedL.cc:39-46), `labels`
 * in 0..9 (fixedL.cc:650).  NT_global / first are the global image count and
 * the offset of this shard (ParallelDo bounds, paralleldo.h:35-42); they only
 * matter for multi-rank runs (pass NT, 0 otherwise). */
int tnml_set_images(tnml_handle h, int64_t NT, int N, const double* feat,
                    const int32_t* labels, int64_t NT_global, int64_t first);

/* MPS site tensors W.A(j) (fixedL.cc:670-728 builds them; 493-521 uses them). */
int tnml_set_site(tnml_handle h, int j, int ml, int mr, int has_label, const double* data);
int tnml_get_site_dims(tnml_handle h, int j, int* ml, int* mr, int* has_label);
int tnml_get_site(tnml_handle h, int j, double* data, size_t capacity_elems);

/* TrainStates::init (fixedL.cc:122-157): right environments E_N..E_3 for every
 * image, kept resident in HBM (the reference spills them to proj_images/),
 * then setBond(1). */
int tnml_init_envs(tnml_handle h);

/* TrainStates::setBond (fixedL.cc:159-190).  Selects LE=slot[b-1], RE=slot[b+2];
 * the dense t.v of fixedL.cc:183-185 is never materialised. */
int tnml_set_bond(tnml_handle h, int b);

/* oB = W.A(c)*W.A(c+dc); B = oB (fixedL.cc:493-498): forms the bond tensor of
 * the current bond on the device. */
int tnml_bond_form(tnml_handle h);
/* Host access to the device bond tensor (layout above).  `has_label` != 0 iff
 * the bond touches site jc. */
int tnml_bond_dims(tnml_handle h, int* ml, int* mr, int* has_label);
int tnml_bond_load(tnml_handle h, const double* B, size_t n_elems);
int tnml_bond_store(tnml_handle h, double* B, size_t capacity_elems);

/* cgrad (fixedL.cc:349-445), including the cross-rank all-reduce of the
 * gradient and of the scalars when a communicator is attached.
 * cost_per_pass / rnorm_per_pass (may be NULL) receive, for every pass that
 * evaluates a new gradient, C/NT (the "Cost =" line, fixedL.cc:429) and |r|
 * (fixedL.cc:439); *npass_done = number of entries written (<= Npass-1). */
int tnml_cgrad(tnml_handle h, int Npass, double lambda, double cconv,
               double* cost_per_pass, double* rnorm_per_pass, int* npass_done);

/* svd(B, W.Aref(c), S, W.Aref(c+dc), {Cutoff,Maxm,Minm}); W.Aref(c+dc) *= S
 * (fixedL.cc:519-521).  dir = TNML_FROMLEFT (ha==1, c=b) or TNML_FROMRIGHT
 * (ha==2, c=b+1).  The new site tensors stay on the device (tnml_get_site).
 * do_rel_cutoff mirrors ITensor's DoRelCutoff (release dependent). */
int tnml_svd_split(tnml_handle h, int dir, double cutoff, int maxm, int minm,
                   int do_rel_cutoff, int* newm, double* truncerr);

/* quadcost (fixedL.cc:280-344) of newB = W.A(c)*W.A(c+dc) (fixedL.cc:527,532)
 * when use_sites != 0, else of the current device bond tensor.  Returns the
 * UN-normalised C, the per-label costs CL (fixedL.cc:333) and the number of
 * images with l == argmax_l |P_l| (fixedL.cc:321-326). */
int tnml_quadcost(tnml_handle h, int use_sites, double lambda, double* C,
                  double* C_label /*[10]*/, int64_t* ncorrect);

/* TrainStates::shiftE (fixedL.cc:192-233). */
int tnml_shift_env(tnml_handle h, int b, int dir);

/* One iteration of the mldmrg loop body (fixedL.cc:478-563):
 * setBond, form B, cgrad, svd, quadcost(newB), shiftE.  All outputs optional. */
typedef struct tnml_bond_params {
  int Npass;            /* fixedL.cc:607 */
  double lambda;        /* fixedL.cc:601 */
  double cconv;         /* fixedL.cc:608 */
  double cutoff;        /* fixedL.cc:591 */
  int maxm, minm;       /* fixedL.cc:592-593 */
  int do_rel_cutoff;
} tnml_bond_params;

typedef struct tnml_bond_result {
  int origm, newm;              /* fixedL.cc:493,522 */
  double truncerr;              /* fixedL.cc:523 */
  double cost;                  /* newC (un-normalised), fixedL.cc:532 */
  double cost_label[TNML_NL];
  int64_t ncorrect;
  double normB, dB;             /* |B|, |B-newB| (fixedL.cc:528-530) */
  int npass_done;
  double cg_cost[8];            /* C/NT per CG pass */
  double cg_rnorm[8];
  int svd_sweeps;
} tnml_bond_result;

int tnml_bond_update(tnml_handle h, int b, int ha, const tnml_bond_params* p,
                     tnml_bond_result* out);

/* Predicted label argmax_l |P_l| for every image of this shard, and the raw
 * outputs P (either may be NULL), from the last quadcost. */
int tnml_predict(tnml_handle h, int32_t* labels_out, double* P_out /*[NT][10]*/);

/* fullTest (util.h:123-200, driven by fulltest.cc:7-99): classify every image currently held
 * by the handle with the MPS currently held (tnml_set_site): per image the full contraction
 * toverlap(psi,img,cent) (util.h:19-40) and argmax_l |W_l| (util.h:160-169).  Implemented as
 * right-environment build + one forward pass at bond 1, i.e. the same kernels as quadcost.
 * pred_out (may be NULL) receives the predicted label of every image of this shard;
 * *ncorrect the number of correct ones (all-reduced over ranks). */
int tnml_fulltest(tnml_handle h, int32_t* pred_out, int64_t* ncorrect);

/* Read an environment slot back (testing / checkpointing):
 * thin -> [NT][m], fat -> [NT][NL][m]. */
int tnml_get_env(tnml_handle h, int slot, int* m, int* is_fat, double* data, size_t capacity_elems);

/* ---- multi-GPU: one process per GPU, images sharded like ParallelDo --------
 * The only exchange is sum-all-reduce of the gradient (+ packed scalars),
 * replacing stdx::accumulate over per-thread partials (fixedL.cc:385,402,421).
 * Rank 0 obtains an id, distributes it out of band (torch.distributed / MPI /
 * a file), every rank calls tnml_comm_init_rank. */
#define TNML_UNIQUE_ID_BYTES 128
int tnml_comm_get_unique_id(uint8_t* id /*[128]*/);
int tnml_comm_init_rank(tnml_handle h, int nranks, int rank, const uint8_t* id);
/* Host-side control values that every rank must agree on (the reference is one process: the LAMBDA /
 * WRITE_WF sentinel files of fixedL.cc:542-559 and the time-seeded RNG of :702-728 are read once there):
 * the `n` doubles of rank `root` replace vals[] on every rank.  No-op without a communicator. */
int tnml_comm_broadcast(tnml_handle h, double* vals, int n, int root);

/* Options (name, value).  "cg_reuse_forward" (default 0): 0 = every CG pass recomputes the
 * residual from scratch exactly like fixedL.cc:412-421; 1 = the forward outputs are updated
 * linearly, P(B + a p) = P(B) + a P(p) with P(p) taken from the pAp pass (fixedL.cc:393-402), and
 * only the backward half is re-evaluated -- same mathematics, 3 of 13 projection passes fewer.
 * "reserve_m": environment slots are allocated for this link dimension so that they never have to
 * grow during the sweeps (tnml_bond_update raises it to maxm by itself).
 * "env_budget_gb" (default 0 = keep everything in HBM): environment tiering.  The reference keeps
 * every per-image environment on disk (proj_images/, fixedL.cc:153,231) and reads the two a bond
 * needs (fixedL.cc:177-178); here at most this many GiB of environment slots stay in HBM, the rest
 * lives in pinned host memory.  Slots are evicted farthest-next-use first and the slot the next
 * bond needs is fetched on a copy stream while the current bond computes.  Results are bit-identical
 * to the all-resident run.
 * "krgemm_variant" (per handle; -1 = automatic): kernel of the projection / environment advance.
 * 3 = tcgen05 int8 error-free splitting (default where it applies: >= 1024 images, 48 <= link
 * dimension <= 128), 2 = persistent FP64 mma.sync kernel, 1 = register-staged FP64 kernel.
 * "oz_slices" (per handle, 6..8, default 8): 7-bit planes per operand of variant 3 (8: element
 * error 1e-15, 7: 1e-13, 6: 1e-11 relative to row scale x column scale).
 * Process-wide variant switches for tests and A/B timing (-1 restores the default):
 * "krgram_variant", "fat_variant", "svd_cluster" (0: multi-launch Jacobi instead of the
 * cluster-resident kernel), "svd_cross" (0: full inner tournaments), "svd_precond"
 * (0: plain Jacobi, 1: one QR, 3: column sort + two QRs).  DESIGN.md section 11 lists them all. */
int tnml_set_option(tnml_handle h, const char* name, double value);

/* Counters for the roofline report: kernel launches issued by this library,
 * algorithmic bytes / flops accumulated since the last reset. */
typedef struct tnml_stats {
  int64_t launches;
  double alg_bytes;
  double alg_flops;
  double ms_proj, ms_grad, ms_fat, ms_svd, ms_shift, ms_other; /* if timing on */
  int64_t tier_evictions, tier_fetches;  /* environment slots moved HBM -> host / host -> HBM */
  double tier_bytes;                     /* bytes that crossed PCIe for the environment tier */
} tnml_stats;
int tnml_get_stats(tnml_handle h, tnml_stats* out, int reset);
int tnml_set_timing(tnml_handle h, int on); /* CUDA-event timing of each phase */
int tnml_synchronize(tnml_handle h);
/* Raw CUDA stream the library launches on (cudaStream_t as void*), so callers
 * can bracket regions with their own events. */
void* tnml_stream(tnml_handle h);

#ifdef __cplusplus
}
#endif
#endif /* TNML_B200_H */
