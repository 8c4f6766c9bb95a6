// tnml_capi.cu -- the C-ABI of include/tnml_b200.h: per-GPU state ("TrainStates"
// on the device), the CG driver of fixedL.cc:349-445, quadcost, svd, shiftE.
// No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX: per-phase ranges for Nsight timelines (no-ops without a tool attached)

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tnml_b200.h"
#include "tnml_kernels.cuh"

using namespace tnml;

namespace {

std::string g_err;  // create-time errors

struct Slot {
  double* p = nullptr;   // device copy (nullptr: not resident)
  size_t bytes = 0;      // capacity of the device buffer
  int m = 0;
  int fat = 0;
  int kind = 0;  // 0 empty, 1 left env, 2 right env
  // host tier (the reference keeps every environment in proj_images/ and reads two
  // per bond, fixedL.cc:177-178,231): pinned copy, valid iff host_valid
  double* host = nullptr;
  size_t host_bytes = 0;
  bool host_valid = false;
  cudaEvent_t ready = nullptr;   // recorded on the fetch stream after an H2D fetch
  bool pending = false;          // the main stream has not waited for `ready` yet
  cudaEvent_t saved = nullptr;   // recorded on the eviction stream after the D2H write-back
};
struct Site {
  double* d = nullptr;
  size_t cap = 0;  // elements
  int ml = 0, mr = 0, lab = 0;
  long size() const { return (long)ml * 2 * mr * (lab ? NL : 1); }
};

struct DBuf {
  double* p = nullptr;
  size_t cap = 0;  // elements
};

// minimal NCCL surface, resolved with dlopen so the library loads without NCCL
struct NcclId {
  char internal[128];
};
typedef int (*nccl_get_uid_t)(NcclId*);
typedef int (*nccl_init_rank_t)(void**, int, NcclId, int);
typedef int (*nccl_allreduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_destroy_t)(void*);
typedef const char* (*nccl_errstr_t)(int);
struct NcclApi {
  void* lib = nullptr;
  nccl_get_uid_t get_uid = nullptr;
  nccl_init_rank_t init_rank = nullptr;
  nccl_allreduce_t allreduce = nullptr;
  nccl_destroy_t destroy = nullptr;
  nccl_errstr_t errstr = nullptr;
};
NcclApi g_nccl;
bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
    return false;
  }
  g_nccl.get_uid = (nccl_get_uid_t)dlsym(lib, "ncclGetUniqueId");
  g_nccl.init_rank = (nccl_init_rank_t)dlsym(lib, "ncclCommInitRank");
  g_nccl.allreduce = (nccl_allreduce_t)dlsym(lib, "ncclAllReduce");
  g_nccl.destroy = (nccl_destroy_t)dlsym(lib, "ncclCommDestroy");
  g_nccl.errstr = (nccl_errstr_t)dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy) {
    err = "libnccl lacks a required symbol";
    dlclose(lib);
    return false;
  }
  g_nccl.lib = lib;
  return true;
}

enum Phase { PH_PROJ = 0, PH_GRAD, PH_FAT, PH_SVD, PH_SHIFT, PH_OTHER, PH_COUNT };

}  // namespace

struct tnml_handle_s {
  int device = 0;
  cudaStream_t st = nullptr;
  int num_sm = 148;
  std::string err;

  long NT = 0, NTg = 0, first = 0;
  int N = 0, jc = 0;
  double* feat = nullptr;  // [N+2][NT][2]
  int32_t* labels = nullptr;
  double* ones = nullptr;  // [NT]
  std::vector<Slot> slot;
  std::vector<Site> W;
  int currb = -1;

  // current bond
  BondGeom g{};
  int cls = 0;  // 0 L, 1 C, 2 R
  bool bond_valid = false;
  DBuf B, r, p, G, T, Gpart, Q, Z, Bm;
  double* P = nullptr;  // [NT][NL]
  double* PV = nullptr; // [NT][NL]  p*v_n of the current CG pass (cg_reuse_forward)
  int cg_reuse_forward = 0;
  // tcgen05 projection (tnml_ozaki.cu): int8 planes of the thin environment of the current bond
  // (cut once per bond) and of the bond-shaped operand (cut per pass)
  int8_t* oz_A8 = nullptr;
  double* oz_ea = nullptr;
  size_t oz_capA = 0, oz_capea = 0;
  int8_t* oz_B8 = nullptr;
  double* oz_eb = nullptr;
  size_t oz_capB = 0, oz_capeb = 0;
  long oz_tag_bond = -1, oz_tag_gen = -1;   // (bond, env generation) the planes in oz_A8 belong to
  int oz_tag_ns = 0;
  const double* oz_tag_ptr = nullptr;       // ... and the environment they were cut from
  long env_gen = 0;                          // bumped whenever an environment slot is (re)written
  int oz_slices = 8;                         // 8: float64-class accuracy (2^-57 of row max x column max)
  int krgemm_variant = -1;                   // -1 default (3 where supported), 3 tcgen05, 2 DMMA persistent, 1 register-staged
  int reserve_m = 0;       // env slots are sized for this link dim (set from maxm by tnml_bond_update)
  // environment tiering: at most env_budget bytes of slots resident in HBM (0 = everything)
  cudaStream_t cp = nullptr;     // eviction stream (D2H) -- copies overlap the kernels and each other
  cudaStream_t cpin = nullptr;   // fetch stream (H2D)
  cudaEvent_t ev_main = nullptr;
  size_t env_budget = 0, env_resident = 0;
  int moving = 1;                // 1: sweeping right (Fromleft), 2: sweeping left
  int32_t* pred = nullptr;
  double* stats_partial = nullptr;
  unsigned* ticket = nullptr;   // last-CTA ticket of the statistics reduction inside fat_kernel
  int nfat_blocks = 0;
  double* dscal = nullptr;  // [64] device scalars
  double* cgs = nullptr;    // [32] device state of the CG recurrence (see cgrad_enqueue)
  int last_svd_sweeps = 0;
  double* dot_scratch = nullptr;
  double* hpin = nullptr;  // pinned host [64]
  SvdWork svd;

  void* comm = nullptr;
  int nranks = 1, rank = 0;

  tnml_stats stats{};
  bool timing = false;
  struct Ev {
    int ph;
    cudaEvent_t a, b;
  };
  std::vector<Ev> evs;
  std::vector<cudaEvent_t> evpool;
};

namespace {

int fail(tnml_handle h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_err = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(h, TNML_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKL()                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                           \
    if (e_ != cudaSuccess)                                                                         \
      return fail(h, TNML_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define TRY(expr)            \
  do {                       \
    int rc_ = (expr);        \
    if (rc_ != 0) return rc_; \
  } while (0)

// Work buffers grow with the link dimensions during the first sweeps.  Growing the stream-ordered
// pool is expensive (measured: 0.3 - 1.4 s stalls on single bonds of an otherwise 5 ms/bond sweep),
// so a buffer that has to grow is sized for `want` (what it will need at maxm, when known) or at
// least 1.5x its old capacity.
int ensure(tnml_handle h, DBuf& b, size_t n, size_t want = 0) {
  if (n <= b.cap) return 0;
  size_t target = std::max(n, std::max(want, b.cap + b.cap / 2));
  if (b.p) CK(cudaFreeAsync(b.p, h->st));
  b.p = nullptr;
  b.cap = 0;
  if (target > n) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || fr < target * sizeof(double) + (tot >> 3)) target = n;
  }
  CK(cudaMallocAsync(&b.p, target * sizeof(double), h->st));
  b.cap = target;
  return 0;
}

struct PhaseTimer {
  tnml_handle h;
  cudaEvent_t a = nullptr, b = nullptr;
  int ph;
  PhaseTimer(tnml_handle h_, int ph_) : h(h_), ph(ph_) {
    static const char* const names[PH_COUNT] = {"tnml:proj", "tnml:grad", "tnml:fat", "tnml:svd", "tnml:shift", "tnml:other"};
    nvtxRangePushA(names[ph]);
    if (!h->timing) return;
    a = get();
    b = get();
    cudaEventRecord(a, h->st);
  }
  ~PhaseTimer() {
    nvtxRangePop();
    if (!h->timing) return;
    cudaEventRecord(b, h->st);
    h->evs.push_back({ph, a, b});
  }
  cudaEvent_t get() {
    if (!h->evpool.empty()) {
      cudaEvent_t e = h->evpool.back();
      h->evpool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

void drain_events(tnml_handle h) {
  if (h->evs.empty()) return;
  cudaStreamSynchronize(h->st);
  for (auto& e : h->evs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.a, e.b);
    double* dst = &h->stats.ms_proj;
    dst[e.ph] += ms;
    h->evpool.push_back(e.a);
    h->evpool.push_back(e.b);
  }
  h->evs.clear();
}

const double* featp(tnml_handle h, int j) { return h->feat + (long)j * h->NT * 2; }

// ---- environment tier ---------------------------------------------------------------------------
size_t slot_used(tnml_handle h, const Slot& s) { return (size_t)h->NT * s.m * (s.fat ? NL : 1) * sizeof(double); }

// make the main stream wait for an in-flight fetch of this slot
int use_slot(tnml_handle h, Slot& s) {
  if (s.pending) {
    CK(cudaStreamWaitEvent(h->st, s.ready, 0));
    s.pending = false;
  }
  return 0;
}

int free_slot_dev(tnml_handle h, Slot& s, cudaStream_t on) {
  if (!s.p) return 0;
  CK(cudaFreeAsync(s.p, on));
  h->env_resident -= s.bytes;
  s.p = nullptr;
  s.bytes = 0;
  return 0;
}

// write the slot back to pinned host memory (if the host copy is stale) and release the device copy
int evict_slot(tnml_handle h, Slot& s) {
  if (!s.p) return 0;
  TRY(use_slot(h, s));
  CK(cudaEventRecord(h->ev_main, h->st));          // everything queued so far may still read it
  CK(cudaStreamWaitEvent(h->cp, h->ev_main, 0));
  const size_t used = slot_used(h, s);
  if (!s.host_valid) {
    if (s.host_bytes < used) {
      // pinning host memory runs at only 1-2 GB/s (measured), so the buffer is sized for the
      // largest link dimension the slot can reach and allocated once
      const size_t want = std::max(used, (size_t)h->NT * std::max(s.m, h->reserve_m) * (s.fat ? NL : 1) * sizeof(double));
      if (s.host) CK(cudaFreeHost(s.host));
      s.host = nullptr;
      s.host_bytes = 0;
      if (cudaMallocHost(&s.host, want) != cudaSuccess) {
        cudaGetLastError();
        CK(cudaMallocHost(&s.host, used));
        s.host_bytes = used;
      } else {
        s.host_bytes = want;
      }
    }
    CK(cudaMemcpyAsync(s.host, s.p, used, cudaMemcpyDeviceToHost, h->cp));
    if (!s.saved) CK(cudaEventCreateWithFlags(&s.saved, cudaEventDisableTiming));
    CK(cudaEventRecord(s.saved, h->cp));
    s.host_valid = true;
    h->stats.tier_bytes += (double)used;
  }
  h->stats.tier_evictions += 1;
  return free_slot_dev(h, s, h->cp);
}

// The slot whose next use lies farthest in the future: moving right at bond b the left
// environments behind us (small j) are needed again only on the way back, lowest j last; the right
// environments ahead are needed soon, highest j last.  Mirror image when moving left.
int pick_victim(tnml_handle h, int lo_keep, int hi_keep) {
  const int N = h->N;
  if (h->moving == 1) {
    for (int j = 1; j < lo_keep; ++j)
      if (h->slot[j].p) return j;
    for (int j = N; j > hi_keep; --j)
      if (h->slot[j].p) return j;
  } else {
    for (int j = N; j > hi_keep; --j)
      if (h->slot[j].p) return j;
    for (int j = 1; j < lo_keep; ++j)
      if (h->slot[j].p) return j;
  }
  return -1;
}

// evict until `need` more bytes fit under the budget; slots lo_keep..hi_keep are protected
int make_room(tnml_handle h, size_t need, int lo_keep, int hi_keep) {
  if (h->env_budget == 0) return 0;
  while (h->env_resident + need > h->env_budget) {
    const int v = pick_victim(h, lo_keep, hi_keep);
    if (v < 0) break;   // only protected slots left: exceed the budget rather than fail
    TRY(evict_slot(h, h->slot[v]));
  }
  return 0;
}

// bring slot j back to HBM on the copy stream (no-op if resident or empty)
int fetch_slot(tnml_handle h, int j, int lo_keep, int hi_keep) {
  if (j < 1 || j > h->N) return 0;
  Slot& s = h->slot[j];
  if (s.kind == 0 || s.p) return 0;
  if (!s.host_valid) return fail(h, TNML_ERR_INVALID, "environment slot %d lost (neither in HBM nor on the host)", j);
  const size_t used = slot_used(h, s);
  TRY(make_room(h, used, lo_keep, hi_keep));
  CK(cudaMallocAsync(&s.p, used, h->cpin));
  s.bytes = used;
  h->env_resident += used;
  if (s.saved) CK(cudaStreamWaitEvent(h->cpin, s.saved, 0));   // the write-back of this slot has landed
  CK(cudaMemcpyAsync(s.p, s.host, used, cudaMemcpyHostToDevice, h->cpin));
  if (!s.ready) CK(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
  CK(cudaEventRecord(s.ready, h->cpin));
  s.pending = true;
  h->stats.tier_fetches += 1;
  h->stats.tier_bytes += (double)used;
  return 0;
}

// env accessors for bond b: returns pointer, dim, fat flag
struct EnvRef {
  const double* p;
  int m;
  int fat;
};
int left_env(tnml_handle h, int b, EnvRef& e) {
  if (b - 1 < 1) {
    e = {h->ones, 1, 0};
    return 0;
  }
  Slot& s = h->slot[b - 1];
  if (s.kind != 1) return fail(h, TNML_ERR_INVALID, "left environment slot %d not built (kind=%d)", b - 1, s.kind);
  TRY(fetch_slot(h, b - 1, b - 1, b + 2));
  TRY(use_slot(h, s));
  e = {s.p, s.m, s.fat};
  return 0;
}
int right_env(tnml_handle h, int b, EnvRef& e) {
  if (b + 2 > h->N) {
    e = {h->ones, 1, 0};
    return 0;
  }
  Slot& s = h->slot[b + 2];
  if (s.kind != 2) return fail(h, TNML_ERR_INVALID, "right environment slot %d not built (kind=%d)", b + 2, s.kind);
  TRY(fetch_slot(h, b + 2, b - 1, b + 2));
  TRY(use_slot(h, s));
  e = {s.p, s.m, s.fat};
  return 0;
}

int bond_class(tnml_handle h, int b) {
  if (b + 1 < h->jc) return 0;
  if (b > h->jc) return 2;
  return 1;
}

int setup_geom(tnml_handle h, int b) {
  Site& wb = h->W[b];
  Site& wb1 = h->W[b + 1];
  if (!wb.d || !wb1.d) return fail(h, TNML_ERR_INVALID, "site tensors %d/%d not set", b, b + 1);
  if (wb.mr != wb1.ml) return fail(h, TNML_ERR_INVALID, "link mismatch between sites %d and %d", b, b + 1);
  BondGeom g{};
  g.ml = wb.ml;
  g.mr = wb1.mr;
  g.lab_b = wb.lab;
  g.lab_b1 = wb1.lab;
  g.nl = (wb.lab || wb1.lab) ? NL : 1;
  h->cls = bond_class(h, b);
  if ((h->cls == 1) != (g.nl == NL)) return fail(h, TNML_ERR_INVALID, "Label Index not on site %d", h->jc);
  const long ml = g.ml, mr = g.mr;
  if (h->cls == 0) {  // Bc[alpha][s][t][beta]
    g.sa = 4 * mr, g.ss = 2 * mr, g.st = mr, g.sb = 1, g.sl = 0;
  } else if (h->cls == 2) {  // Bc[beta][t][s][alpha]
    g.sb = 4 * ml, g.st = 2 * ml, g.ss = ml, g.sa = 1, g.sl = 0;
  } else {  // Bc[alpha][s][t][l][beta]
    g.sa = 4 * NL * mr, g.ss = 2 * NL * mr, g.st = NL * mr, g.sl = mr, g.sb = 1;
  }
  h->g = g;
  EnvRef le, re;
  TRY(left_env(h, b, le));
  TRY(right_env(h, b, re));
  if (le.m != g.ml || re.m != g.mr)
    return fail(h, TNML_ERR_INVALID, "environment dims (%d,%d) do not match bond dims (%d,%d) at bond %d", le.m,
                re.m, g.ml, g.mr, b);
  const bool wantLfat = (h->cls == 2), wantRfat = (h->cls == 0);
  if ((le.fat != 0) != wantLfat || (re.fat != 0) != wantRfat)
    return fail(h, TNML_ERR_INVALID, "environment label placement inconsistent at bond %d", b);
  return 0;
}

int ensure_bytes(tnml_handle h, void** p, size_t& cap, size_t bytes) {
  if (bytes <= cap) return 0;
  if (*p) CK(cudaFreeAsync(*p, h->st));
  *p = nullptr;
  cap = 0;
  CK(cudaMallocAsync(p, bytes, h->st));
  cap = bytes;
  return 0;
}

// plane buffers of the tcgen05 kernels for `rows` operand rows and S*J operand columns; a re-allocation
// of the row planes drops the tag that says which environment they hold
int oz_buffers(tnml_handle h, long rows, int S, int J) {
  const int8_t* before = h->oz_A8;
  TRY(ensure_bytes(h, (void**)&h->oz_A8, h->oz_capA, oz_a8_bytes(rows, 8)));
  TRY(ensure_bytes(h, (void**)&h->oz_ea, h->oz_capea, (size_t)oz_rows_pad(rows) * sizeof(double)));
  TRY(ensure_bytes(h, (void**)&h->oz_B8, h->oz_capB, oz_b8_bytes(S, J, 8)));
  TRY(ensure_bytes(h, (void**)&h->oz_eb, h->oz_capeb, (size_t)oz_cols_pad(S, J) * sizeof(double)));
  if (h->oz_A8 != before) h->oz_tag_bond = -1;
  return 0;
}
// true if oz_A8 already holds the planes of environment `env` (cut for this bond, no advance since)
bool oz_planes_valid(tnml_handle h, const double* env, int ns) {
  return h->oz_tag_bond == h->currb && h->oz_tag_gen == h->env_gen && h->oz_tag_ns == ns && h->oz_tag_ptr == env;
}

// Q[n][j] = sum_p w_p(n) sum_a thin[n][a] X[(a*4+p)*J + j]: the projection half of P = B * t.v
// (fixedL.cc:318,377,399,416).  Default: tcgen05 int8 kernel on error-free 7-bit planes -- the planes
// of the thin environment are cut once per bond (it does not change during the bond update), those
// of X every call.  Fallback / variants 1,2: FP64 mma.sync kernels.
int project(tnml_handle h, const double* thin, int mt, const double* f1, const double* f2, const double* X, int J) {
  const long NT = h->NT;
  int variant = h->krgemm_variant;
  if (variant < 0) {
    const char* e = getenv("TNML_KRGEMM");
    variant = e ? atoi(e) : 3;
  }
  const int ns = h->oz_slices;
  if (variant == 3 && NT >= 1024 && mt >= 48 && oz_supported(4, mt, ns)) {
    TRY(oz_buffers(h, NT, 4, J));
    if (!oz_planes_valid(h, thin, ns)) {
      oz_slice_rows(h->st, thin, mt, mt, NT, ns, h->oz_A8, h->oz_ea);
      CKL();
      h->oz_tag_bond = h->currb;
      h->oz_tag_gen = h->env_gen;
      h->oz_tag_ns = ns;
      h->oz_tag_ptr = thin;
      h->stats.launches += 1;
    }
    oz_slice_cols(h->st, 4, X, J, mt, J, ns, h->oz_B8, h->oz_eb);
    CKL();
    if (oz_krgemm(h->st, 4, ns, h->oz_A8, h->oz_ea, NT, f1, f2, 1, h->oz_B8, h->oz_eb, J, h->Q.p, J, h->num_sm)) {
      CKL();
      h->stats.launches += 1;
      return 0;
    }
    cudaGetLastError();   // tensor-map encoding unavailable: FP64 kernels below
  }
  krgemm(h->st, 4, thin, mt, mt, f1, f2, 1, X, J, J, h->Q.p, J, NT, h->num_sm);
  CKL();
  return 0;
}

// forward pass: P[n][l] for bond-shaped tensor X (canonical layout).
// mode: FAT_GRAD (also produces Z), FAT_PAP, FAT_COST.  Leaves stats in dstats.
int forward(tnml_handle h, const double* X, int mode, double* dstats, double* Pout = nullptr) {
  if (!Pout) Pout = h->P;
  const int b = h->currb;
  const BondGeom& g = h->g;
  EnvRef le, re;
  TRY(left_env(h, b, le));
  TRY(right_env(h, b, re));
  const long NT = h->NT;
  if (h->cls != 1) {
    const bool L = (h->cls == 0);
    const EnvRef& thin = L ? le : re;
    const EnvRef& fat = L ? re : le;
    const double* fth = featp(h, L ? b : b + 1);
    const double* ffa = featp(h, L ? b + 1 : b);
    const int mt = thin.m, mf = fat.m;
    const size_t wantq = (size_t)NT * std::max(mf, h->reserve_m);
    TRY(ensure(h, h->Q, (size_t)NT * mf, wantq));
    if (mode == FAT_GRAD) TRY(ensure(h, h->Z, (size_t)NT * mf, wantq));
    {
      PhaseTimer t(h, PH_PROJ);
      TRY(project(h, thin.p, mt, fth, ffa, X, mf));
    }
    {
      PhaseTimer t(h, PH_FAT);
      fat_kernel(h->st, mode, h->Q.p, fat.p, mf, h->labels, Pout, h->Z.p, h->pred, h->stats_partial,
                 h->nfat_blocks, NT, dstats, h->ticket);
      CKL();
    }
    h->stats.launches += 2;
    h->stats.alg_flops += (double)NT * (8.0 * mt * mf + 2.0 * NL * mf * (mode == FAT_GRAD ? 2 : 1));
    h->stats.alg_bytes += 8.0 * NT * ((double)mt + (double)NL * mf + 4);
  } else {
    const long J = (long)NL * g.mr;
    const size_t wantq = (size_t)NT * NL * std::max(g.mr, h->reserve_m);
    TRY(ensure(h, h->Q, (size_t)NT * J, wantq));
    if (mode == FAT_GRAD) TRY(ensure(h, h->Z, (size_t)NT * J, wantq));
    {
      PhaseTimer t(h, PH_PROJ);
      TRY(project(h, le.p, g.ml, featp(h, b), featp(h, b + 1), X, (int)J));
    }
    {
      PhaseTimer t(h, PH_FAT);
      fat_kernel(h->st, mode == FAT_GRAD ? FAT_GRAD_OUTER : mode, re.p, h->Q.p, g.mr, h->labels, Pout, h->Z.p,
                 h->pred, h->stats_partial, h->nfat_blocks, NT, dstats, h->ticket);
      CKL();
    }
    h->stats.launches += 2;
    h->stats.alg_flops += (double)NT * (8.0 * NL * g.ml * g.mr + 2.0 * NL * g.mr);
    h->stats.alg_bytes += 8.0 * NT * ((double)g.ml + g.mr + 4);
  }
  return 0;
}

// backward: G = sum_n dP_n (x) v_n from Z (left by forward(FAT_GRAD)); G gets a 16-double tail
int backward(tnml_handle h) {
  const int b = h->currb;
  const BondGeom& g = h->g;
  EnvRef le, re;
  TRY(left_env(h, b, le));
  TRY(right_env(h, b, re));
  const long NT = h->NT;
  const long n = g.size();
  TRY(ensure(h, h->G, (size_t)n + 16));
  const double* thin;
  int mt;
  const double *f1, *f2;
  long J;
  if (h->cls == 0) {
    thin = le.p, mt = le.m, f1 = featp(h, b), f2 = featp(h, b + 1), J = re.m;
  } else if (h->cls == 2) {
    thin = re.p, mt = re.m, f1 = featp(h, b + 1), f2 = featp(h, b), J = le.m;
  } else {
    thin = le.p, mt = le.m, f1 = featp(h, b), f2 = featp(h, b + 1), J = (long)NL * g.mr;
  }
  const int ns = krgram_splits(mt, 4, (int)J, NT, h->num_sm);
  TRY(ensure(h, h->Gpart, (size_t)ns * n));
  PhaseTimer t(h, PH_GRAD);
  krgram(h->st, 4, thin, mt, mt, f1, f2, h->Z.p, J, (int)J, h->Gpart.p, NT, ns);
  CKL();
  reduce_partials(h->st, h->Gpart.p, ns, n, h->G.p);
  CKL();
  h->stats.launches += 2;
  h->stats.alg_flops += (double)NT * 8.0 * mt * J;
  h->stats.alg_bytes += 8.0 * NT * ((double)mt + J + 4);
  return 0;
}

int allreduce(tnml_handle h, double* buf, size_t count) {
  if (!h->comm) return 0;
  int rc = g_nccl.allreduce(buf, buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, h->comm, h->st);
  if (rc != 0)
    return fail(h, TNML_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
  h->stats.launches += 1;
  return 0;
}

int fetch(tnml_handle h, const double* dsrc, int n, double* hdst) {
  CK(cudaMemcpyAsync(h->hpin, dsrc, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  memcpy(hdst, h->hpin, n * sizeof(double));
  return 0;
}

// gradient evaluation at X: G (+16 stats tail, all-reduced).  After the all-reduce: G -= lambda*X
// (fixedL.cc:386,422) and |G|^2 into slot 12 of the 16-double tail.  Nothing is read back: the CG
// scalars (|r|^2, beta, step size, costs per pass, convergence flag) live in h->cgs on the device.
int finish_grad(tnml_handle h, const double* X, double lambda) {
  const long n = h->g.size();
  if (lambda != 0.0) {
    axpby(h->st, n, -lambda, X, 1.0, h->G.p);
    CKL();
    h->stats.launches += 1;
  }
  dot(h->st, n, h->G.p, h->G.p, h->dot_scratch, h->G.p + n + 12);
  CKL();
  h->stats.launches += 2;
  return 0;
}

int grad_eval(tnml_handle h, const double* X, double lambda) {
  const long n = h->g.size();
  TRY(ensure(h, h->G, (size_t)n + 16));
  TRY(forward(h, X, FAT_GRAD, h->dscal));
  TRY(backward(h));
  CK(cudaMemcpyAsync(h->G.p + n, h->dscal, 16 * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  TRY(allreduce(h, h->G.p, (size_t)n + 16));
  return finish_grad(h, X, lambda);
}

// cg_reuse_forward: the forward outputs are linear in the bond tensor, P(B + a p) = P(B) + a P(p),
// and P(p) was just computed for pAp -- so the next residual needs only the backward half:
// dP = delta - P, cost statistics, Z (one pass over the fat environment) and the krgram contraction.
int grad_from_P(tnml_handle h, const double* X, double lambda) {
  const int b = h->currb;
  const BondGeom& g = h->g;
  EnvRef le, re;
  TRY(left_env(h, b, le));
  TRY(right_env(h, b, re));
  const long NT = h->NT;
  const long n = g.size();
  TRY(ensure(h, h->G, (size_t)n + 16));
  {
    PhaseTimer t(h, PH_FAT);
    if (h->cls != 1) {
      const EnvRef& fat = (h->cls == 0) ? re : le;
      TRY(ensure(h, h->Z, (size_t)NT * fat.m, (size_t)NT * std::max(fat.m, h->reserve_m)));
      fat_kernel(h->st, FAT_BWD, h->Q.p, fat.p, fat.m, h->labels, h->P, h->Z.p, h->pred, h->stats_partial,
                 h->nfat_blocks, NT, h->dscal, h->ticket);
      h->stats.alg_bytes += 8.0 * NT * ((double)NL * fat.m + fat.m + NL);
      h->stats.alg_flops += (double)NT * 2.0 * NL * fat.m;
    } else {
      TRY(ensure(h, h->Z, (size_t)NT * NL * g.mr, (size_t)NT * NL * std::max(g.mr, h->reserve_m)));
      fat_kernel(h->st, FAT_BWD_OUTER, re.p, h->Q.p, g.mr, h->labels, h->P, h->Z.p, h->pred, h->stats_partial,
                 h->nfat_blocks, NT, h->dscal, h->ticket);
    }
    CKL();
    h->stats.launches += 1;
  }
  TRY(backward(h));
  CK(cudaMemcpyAsync(h->G.p + n, h->dscal, 16 * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  TRY(allreduce(h, h->G.p, (size_t)n + 16));
  return finish_grad(h, X, lambda);
}


// `reserve` >= bytes: size to allocate when the slot has to grow.  During sweeps the link dims
// grow towards maxm; growing a slot means a fresh cudaMallocAsync of up to 0.6 GB (measured:
// 20 ms stalls on ~1 bond in 10 during the first sweeps), so slots are sized for maxm at once
// while that fits comfortably in HBM.
int alloc_slot(tnml_handle h, Slot& s, size_t bytes, size_t reserve, int lo_keep, int hi_keep) {
  TRY(use_slot(h, s));
  if (s.p && s.bytes >= bytes) return 0;
  TRY(free_slot_dev(h, s, h->st));
  if (h->env_budget) reserve = bytes;   // tiering: exact sizes, the budget is what matters
  TRY(make_room(h, bytes, lo_keep, hi_keep));
  if (reserve > bytes) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr > reserve + (tot >> 3) &&
        cudaMallocAsync(&s.p, reserve, h->st) == cudaSuccess) {
      s.bytes = reserve;
      h->env_resident += reserve;
      return 0;
    }
    cudaGetLastError();
    s.p = nullptr;
  }
  CK(cudaMallocAsync(&s.p, bytes, h->st));
  s.bytes = bytes;
  h->env_resident += bytes;
  return 0;
}

// env advance through site c.  right=0: new left env slot[c] from slot[c-1];
// right=1: new right env slot[c] from slot[c+1].  (fixedL.cc:144-150, 221-229)
int advance_env(tnml_handle h, int c, int right) {
  Site& w = h->W[c];
  if (!w.d) return fail(h, TNML_ERR_INVALID, "site tensor %d not set", c);
  const long NT = h->NT;
  const int prevc = right ? c + 1 : c - 1;
  const bool hasPrev = (prevc >= 1 && prevc <= h->N);
  EnvRef pe{h->ones, 1, 0};
  const int klo = std::min(c, prevc), khi = std::max(c, prevc);   // slots this call touches
  if (hasPrev) {
    Slot& ps = h->slot[prevc];
    if (ps.kind != (right ? 2 : 1))
      return fail(h, TNML_ERR_INVALID, "cannot advance env to site %d: slot %d not a %s env", c, prevc,
                  right ? "right" : "left");
    TRY(fetch_slot(h, prevc, klo, khi));
    TRY(use_slot(h, ps));
    pe = {ps.p, ps.m, ps.fat};
  }
  const int kin = right ? w.mr : w.ml;    // contracted link
  const int kout = right ? w.ml : w.mr;   // new env dim
  if (pe.m != kin) return fail(h, TNML_ERR_INVALID, "env dim %d != link dim %d at site %d", pe.m, kin, c);
  if (pe.fat && w.lab) return fail(h, TNML_ERR_INVALID, "two label indices at site %d", c);
  const int outfat = (pe.fat || w.lab) ? 1 : 0;
  Slot& ns = h->slot[c];
  const size_t bytes = (size_t)NT * kout * (outfat ? NL : 1) * sizeof(double);
  const size_t reserve = (size_t)NT * std::max(kout, h->reserve_m) * (outfat ? NL : 1) * sizeof(double);
  TRY(alloc_slot(h, ns, bytes, reserve, klo, khi));
  ns.host_valid = false;   // about to be overwritten
  const double* Bm = w.d;
  PhaseTimer t(h, PH_SHIFT);
  if (right || w.lab) {
    TRY(ensure(h, h->Bm, (size_t)w.size()));
    permute_site(h->st, w.d, w.ml, w.mr, w.lab ? NL : 1, right, h->Bm.p);
    CKL();
    h->stats.launches += 1;
    Bm = h->Bm.p;
  }
  const long rows = pe.fat ? NT * NL : NT;
  const int div = pe.fat ? NL : 1;
  const int J = kout * (w.lab ? NL : 1);
  // Same tcgen05 int8 kernel as the projection, S = 2 weights.  Label-carrying previous environment (10 NT
  // rows, one image per NL rows: the expensive advance of a leftward class-L / rightward class-R bond): its
  // planes are cut here, into the projection's buffers (the bond update that used them is over).  Thin
  // previous environment: it is the operand the finished bond update projected with, so its planes are
  // normally still there (tags) and the advance costs one kernel over NT rows.
  int variant = h->krgemm_variant;
  if (variant < 0) {
    const char* e = getenv("TNML_KRGEMM");
    variant = e ? atoi(e) : 3;
  }
  bool done = false;
  // (a thin environment whose planes would have to be cut first only from m = 96 on: below that the FP64
  // kernel, whose cost falls with m^2 where the int8 kernel's K stays padded to 128, is as fast)
  if (variant == 3 && hasPrev && NT >= 1024 && kin >= 48 && oz_supported(2, kin, h->oz_slices) &&
      (pe.fat || kin >= 96 || oz_planes_valid(h, pe.p, h->oz_slices))) {
    const int nsl = h->oz_slices;
    TRY(oz_buffers(h, rows, 2, J));
    int nl = 2;
    if (pe.fat || !oz_planes_valid(h, pe.p, nsl)) {
      h->oz_tag_bond = -1;   // the planes in the buffer are replaced by this environment's ...
      oz_slice_rows(h->st, pe.p, kin, kin, rows, nsl, h->oz_A8, h->oz_ea);
      CKL();
      ++nl;
    }
    h->oz_tag_bond = -1;     // ... and no later projection uses them: the environments change below
    oz_slice_cols(h->st, 2, Bm, J, kin, J, nsl, h->oz_B8, h->oz_eb);
    CKL();
    if (oz_krgemm(h->st, 2, nsl, h->oz_A8, h->oz_ea, rows, featp(h, c), nullptr, div, h->oz_B8, h->oz_eb, J, ns.p, J,
                  h->num_sm)) {
      CKL();
      h->stats.launches += nl;
      done = true;
    } else {
      cudaGetLastError();
    }
  }
  if (!done) {
    krgemm(h->st, 2, pe.p, kin, kin, featp(h, c), nullptr, div, Bm, J, J, ns.p, J, rows, h->num_sm);
    CKL();
    h->stats.launches += 1;
  }
  h->stats.alg_flops += (double)rows * 4.0 * kin * J;
  h->stats.alg_bytes += 8.0 * rows * ((double)kin + J);
  ns.m = kout;
  ns.fat = outfat;
  ns.kind = right ? 2 : 1;
  ++h->env_gen;
  return 0;
}

int set_site_dev(tnml_handle h, int j, int ml, int mr, int lab) {
  Site& s = h->W[j];
  size_t n = (size_t)ml * 2 * mr * (lab ? NL : 1);
  if (n > s.cap) {
    if (s.d) CK(cudaFreeAsync(s.d, h->st));
    s.d = nullptr;
    const size_t rm = (size_t)std::min(h->reserve_m, 1024);   // site tensors are small: size them for maxm at once
    const size_t want = std::max(n, std::max(rm * 2 * rm * (lab ? NL : 1), s.cap + s.cap / 2));
    s.cap = 0;
    CK(cudaMallocAsync(&s.d, want * sizeof(double), h->st));
    s.cap = want;
  }
  s.ml = ml;
  s.mr = mr;
  s.lab = lab;
  return 0;
}

}  // namespace

extern "C" {

const char* tnml_version(void) { return "tnml_b200 0.2 (sm_100a, float64 results: tcgen05 int8 error-free projection + DMMA gradient, cluster Jacobi SVD)"; }

const char* tnml_last_error(tnml_handle h) { return h ? h->err.c_str() : g_err.c_str(); }

int tnml_create(int device, int flags, tnml_handle* out) {
  (void)flags;
  if (!out) return fail(nullptr, TNML_ERR_INVALID, "out == NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, TNML_ERR_NODEVICE, "no CUDA device (%s); tnml_b200 has no CPU path",
                e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, TNML_ERR_INVALID, "device %d out of range (%d)", device, ndev);
  tnml_handle h = new tnml_handle_s();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete h;
    return fail(nullptr, TNML_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  h->num_sm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->cp, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->cpin, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess) {
    delete h;
    return fail(nullptr, TNML_ERR_CUDA, "cudaStreamCreate failed");
  }
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  h->nfat_blocks = fat_blocks(h->num_sm);
  if (cudaMalloc(&h->stats_partial, (size_t)h->nfat_blocks * 16 * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->ticket, 64) != cudaSuccess || cudaMemset(h->ticket, 0, 64) != cudaSuccess ||
      cudaMalloc(&h->dscal, 64 * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->cgs, 32 * sizeof(double)) != cudaSuccess || cudaMemset(h->cgs, 0, 32 * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->dot_scratch, 1024 * sizeof(double)) != cudaSuccess ||
      cudaMallocHost(&h->hpin, 64 * sizeof(double)) != cudaSuccess) {
    delete h;
    return fail(nullptr, TNML_ERR_CUDA, "workspace allocation failed");
  }
  *out = h;
  return TNML_OK;
}

int tnml_destroy(tnml_handle h) {
  if (!h) return TNML_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->st);
  if (h->comm && g_nccl.destroy) g_nccl.destroy(h->comm);
  if (h->cp) cudaStreamSynchronize(h->cp);
  if (h->cpin) cudaStreamSynchronize(h->cpin);
  for (auto& s : h->slot) {
    if (s.p) cudaFree(s.p);
    if (s.host) cudaFreeHost(s.host);
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.saved) cudaEventDestroy(s.saved);
  }
  for (auto& s : h->W)
    if (s.d) cudaFree(s.d);
  DBuf* bufs[] = {&h->B, &h->r, &h->p, &h->G, &h->T, &h->Gpart, &h->Q, &h->Z, &h->Bm};
  for (DBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  void* ptrs[] = {h->feat, h->labels, h->ones, h->P, h->PV, h->pred, h->stats_partial, h->ticket, h->dscal, h->dot_scratch,
                  h->oz_A8, h->oz_ea, h->oz_B8, h->oz_eb, h->cgs,
                  h->svd.X, h->svd.J, h->svd.sig2, h->svd.perm, h->svd.info, h->svd.flags,
                  h->svd.M, h->svd.tau, h->svd.ready, h->svd.Y, h->svd.M2, h->svd.tau2, h->svd.Y2, h->svd.perm0, h->svd.sweepmax};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& ge : h->svd.gexec)
    if (ge) cudaGraphExecDestroy(ge);
  if (h->svd.st2) {
    cudaStreamSynchronize(h->svd.st2);
    cudaStreamDestroy(h->svd.st2);
    cudaEventDestroy(h->svd.ev_fork);
    cudaEventDestroy(h->svd.ev_join);
  }
  if (h->hpin) cudaFreeHost(h->hpin);
  for (auto& e : h->evs) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  for (auto e : h->evpool) cudaEventDestroy(e);
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->cp) cudaStreamDestroy(h->cp);
  if (h->cpin) cudaStreamDestroy(h->cpin);
  cudaStreamDestroy(h->st);
  delete h;
  return TNML_OK;
}

int tnml_set_images(tnml_handle h, int64_t NT, int N, const double* feat, const int32_t* labels,
                    int64_t NT_global, int64_t first) {
  if (!h) return TNML_ERR_INVALID;
  if (NT <= 0 || N < 4 || !feat || !labels) return fail(h, TNML_ERR_INVALID, "bad image set (NT=%ld N=%d)", (long)NT, N);
  for (int64_t n = 0; n < NT; ++n)
    if (labels[n] < 0 || labels[n] >= NL) return fail(h, TNML_ERR_INVALID, "label %d of image %ld out of range", labels[n], (long)n);
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaStreamSynchronize(h->cp));
  CK(cudaStreamSynchronize(h->cpin));
  for (auto& s : h->slot) {
    if (s.p) cudaFree(s.p);
    if (s.host) cudaFreeHost(s.host);
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.saved) cudaEventDestroy(s.saved);
  }
  h->env_resident = 0;
  for (auto& s : h->W)
    if (s.d) cudaFree(s.d);
  void* ptrs[] = {h->feat, h->labels, h->ones, h->P, h->PV, h->pred};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  h->feat = nullptr, h->labels = nullptr, h->ones = nullptr, h->P = nullptr, h->PV = nullptr, h->pred = nullptr;
  h->NT = NT;
  h->NTg = NT_global > 0 ? NT_global : NT;
  h->first = first;
  h->N = N;
  h->jc = N / 2;  // fixedL.cc:616
  h->slot.assign(N + 2, Slot());
  h->W.assign(N + 2, Site());
  h->currb = -1;
  h->bond_valid = false;
  ++h->env_gen;
  const size_t nf = (size_t)(N + 2) * NT * 2;
  CK(cudaMalloc(&h->feat, nf * sizeof(double)));
  CK(cudaMalloc(&h->labels, NT * sizeof(int32_t)));
  CK(cudaMalloc(&h->ones, NT * sizeof(double)));
  CK(cudaMalloc(&h->P, (size_t)NT * NL * sizeof(double)));
  CK(cudaMalloc(&h->PV, (size_t)NT * NL * sizeof(double)));
  CK(cudaMalloc(&h->pred, NT * sizeof(int32_t)));
  // [NT][N][2] -> [N+2][NT][2] (site-major: one bond's features are contiguous)
  std::vector<double> stage(nf, 0.0);
  for (int64_t n = 0; n < NT; ++n)
    for (int j = 1; j <= N; ++j) {
      const double* src = feat + ((size_t)n * N + (j - 1)) * 2;
      double* dst = stage.data() + ((size_t)j * NT + n) * 2;
      dst[0] = src[0];
      dst[1] = src[1];
    }
  CK(cudaMemcpyAsync(h->feat, stage.data(), nf * sizeof(double), cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(h->labels, labels, NT * sizeof(int32_t), cudaMemcpyHostToDevice, h->st));
  fill(h->st, h->ones, NT, 1.0);
  CKL();
  CK(cudaStreamSynchronize(h->st));
  return TNML_OK;
}

int tnml_set_site(tnml_handle h, int j, int ml, int mr, int has_label, const double* data) {
  if (!h) return TNML_ERR_INVALID;
  if (h->N == 0) return fail(h, TNML_ERR_INVALID, "tnml_set_images must be called first");
  if (j < 1 || j > h->N || ml < 1 || mr < 1 || !data) return fail(h, TNML_ERR_INVALID, "bad site %d (%d,%d)", j, ml, mr);
  if ((has_label != 0) != (j == h->jc)) return fail(h, TNML_ERR_INVALID, "Label Index not on site %d", h->jc);
  CK(cudaSetDevice(h->device));
  TRY(set_site_dev(h, j, ml, mr, has_label ? 1 : 0));
  CK(cudaMemcpyAsync(h->W[j].d, data, (size_t)h->W[j].size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (j == h->currb || j == h->currb + 1) h->bond_valid = false;
  return TNML_OK;
}

int tnml_get_site_dims(tnml_handle h, int j, int* ml, int* mr, int* has_label) {
  if (!h || j < 1 || j > h->N || !h->W[j].d) return h ? fail(h, TNML_ERR_INVALID, "site %d not set", j) : TNML_ERR_INVALID;
  if (ml) *ml = h->W[j].ml;
  if (mr) *mr = h->W[j].mr;
  if (has_label) *has_label = h->W[j].lab;
  return TNML_OK;
}

int tnml_get_site(tnml_handle h, int j, double* data, size_t capacity_elems) {
  if (!h || j < 1 || j > h->N || !h->W[j].d) return h ? fail(h, TNML_ERR_INVALID, "site %d not set", j) : TNML_ERR_INVALID;
  size_t n = (size_t)h->W[j].size();
  if (capacity_elems < n) return fail(h, TNML_ERR_INVALID, "buffer too small for site %d (%zu < %zu)", j, capacity_elems, n);
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(data, h->W[j].d, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return TNML_OK;
}

int tnml_init_envs(tnml_handle h) {
  if (!h || h->N == 0) return h ? fail(h, TNML_ERR_INVALID, "no images") : TNML_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  for (int j = 1; j <= h->N; ++j)
    if (!h->W[j].d) return fail(h, TNML_ERR_INVALID, "site tensor %d not set", j);
  if (!h->W[h->jc].lab) return fail(h, TNML_ERR_INVALID, "Label Index not on site %d", h->jc);
  for (auto& s : h->slot) {
    s.kind = 0;
    s.host_valid = false;
  }
  // size the int8 plane buffers once, for their largest user (the advance of a label-carrying environment,
  // NL NT rows; class-C operand columns), so that the sweeps never stop for a 0.6 GB allocation
  {
    int variant = h->krgemm_variant;
    if (variant < 0) {
      const char* e = getenv("TNML_KRGEMM");
      variant = e ? atoi(e) : 3;
    }
    if (variant == 3 && h->NT >= 1024 && h->reserve_m >= 48 && h->reserve_m <= 128)
      TRY(oz_buffers(h, h->NT * NL, 4, NL * h->reserve_m));
  }
  h->moving = 2;   // building right environments N..3: the high slots are needed last
  for (int n = h->N; n >= 3; --n) TRY(advance_env(h, n, 1));
  h->moving = 1;
  h->currb = -1;
  return tnml_set_bond(h, 1);
}

int tnml_set_bond(tnml_handle h, int b) {
  if (!h) return TNML_ERR_INVALID;
  if (b < 1 || b >= h->N) return fail(h, TNML_ERR_INVALID, "bond %d out of range", b);
  if (h->currb == b) return TNML_OK;  // fixedL.cc:162
  h->currb = b;
  h->bond_valid = false;
  if (h->env_budget) {
    // tiering: the two environments of this bond (fixedL.cc:177-178 reads them from disk), then
    // the one the next bond will need, fetched on the copy stream while this bond computes
    CK(cudaSetDevice(h->device));
    TRY(fetch_slot(h, b - 1, b - 1, b + 2));
    TRY(fetch_slot(h, b + 2, b - 1, b + 2));
    if (h->moving == 1)
      TRY(fetch_slot(h, b + 3, b - 1, b + 3));
    else
      TRY(fetch_slot(h, b - 2, b - 2, b + 2));
  }
  return TNML_OK;
}

int tnml_bond_form(tnml_handle h) {
  if (!h || h->currb < 1) return h ? fail(h, TNML_ERR_INVALID, "no current bond") : TNML_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const int b = h->currb;
  TRY(setup_geom(h, b));
  const long n = h->g.size();
  TRY(ensure(h, h->B, (size_t)n));
  PhaseTimer t(h, PH_OTHER);
  form_bond(h->st, h->W[b].d, h->W[b + 1].d, h->W[b].mr, h->g, h->B.p);
  CKL();
  h->stats.launches += 1;
  h->bond_valid = true;
  return TNML_OK;
}

int tnml_bond_dims(tnml_handle h, int* ml, int* mr, int* has_label) {
  if (!h || !h->bond_valid) return h ? fail(h, TNML_ERR_INVALID, "no bond tensor formed") : TNML_ERR_INVALID;
  if (ml) *ml = h->g.ml;
  if (mr) *mr = h->g.mr;
  if (has_label) *has_label = (h->g.nl == NL);
  return TNML_OK;
}

int tnml_bond_load(tnml_handle h, const double* B, size_t n_elems) {
  if (!h || h->currb < 1 || !B) return h ? fail(h, TNML_ERR_INVALID, "no current bond") : TNML_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  TRY(setup_geom(h, h->currb));
  const long n = h->g.size();
  if ((long)n_elems != n) return fail(h, TNML_ERR_INVALID, "bond tensor has %ld elements, got %zu", n, n_elems);
  TRY(ensure(h, h->B, (size_t)n));
  TRY(ensure(h, h->T, (size_t)n));
  CK(cudaMemcpyAsync(h->T.p, B, n * sizeof(double), cudaMemcpyHostToDevice, h->st));
  bond_from_host_layout(h->st, h->T.p, h->g, h->B.p);
  CKL();
  CK(cudaStreamSynchronize(h->st));
  h->bond_valid = true;
  return TNML_OK;
}

int tnml_bond_store(tnml_handle h, double* B, size_t capacity_elems) {
  if (!h || !h->bond_valid || !B) return h ? fail(h, TNML_ERR_INVALID, "no bond tensor formed") : TNML_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const long n = h->g.size();
  if ((long)capacity_elems < n) return fail(h, TNML_ERR_INVALID, "buffer too small for bond tensor");
  TRY(ensure(h, h->T, (size_t)n));
  bond_to_host_layout(h->st, h->B.p, h->g, h->T.p);
  CKL();
  CK(cudaMemcpyAsync(B, h->T.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return TNML_OK;
}

// cgrad (fixedL.cc:349-445), enqueue half: every kernel of the Npass passes is queued without a
// single host read-back.  The scalars of the recurrence stay in h->cgs (device):
//   [0] |r|^2  [1] beta  [2] converged flag  [3] passes recorded  [4] step a  [8..15] C/NT per pass  [16..23] |r| per pass
// `|r| < cconv` (fixedL.cc:432-436) cannot break a queue that is already enqueued; instead the flag
// freezes the state: every later step size is 0 and no later cost is recorded, so B, the costs and
// the pass count are exactly what the reference's `break` leaves behind (the remaining passes are
// wasted work in that -- in practice never taken -- case).
static int cgrad_enqueue(tnml_handle h, int Npass, double lambda, double cconv) {
  const long n = h->g.size();
  TRY(ensure(h, h->p, (size_t)n));
  double* cgs = h->cgs;
  // r = sum_n (delta - B v_n) v_n - lambda B     (fixedL.cc:373-386); p = r (388)
  TRY(grad_eval(h, h->B.p, lambda));
  cg_begin(h->st, h->G.p + n, cgs);
  CKL();
  CK(cudaMemcpyAsync(h->p.p, h->G.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  h->stats.launches += 1;
  for (int pass = 1; pass <= Npass; ++pass) {
    // pAp = sum_n |p v_n|^2 + lambda |p|^2          (393-403)
    TRY(forward(h, h->p.p, FAT_PAP, h->dscal, h->cg_reuse_forward ? h->PV : nullptr));
    TRY(allreduce(h, h->dscal, 16));
    if (lambda != 0.0) {
      dot(h->st, n, h->p.p, h->p.p, h->dot_scratch, h->dscal + 21);
      CKL();
      h->stats.launches += 2;
    }
    // a = |r|^2 / pAp (405) and B += a p (406)
    cg_step(h->st, cgs, h->dscal + 11, lambda, h->dscal + 21);
    CKL();
    axpby_dev(h->st, n, cgs + 4, h->p.p, 1.0, h->B.p);
    CKL();
    h->stats.launches += 2;
    if (pass == Npass) break;                         // 409
    if (h->cg_reuse_forward) {
      axpby_dev(h->st, (long)h->NT * NL, cgs + 4, h->PV, 1.0, h->P);   // P(B + a p) = P(B) + a P(p)
      CKL();
      h->stats.launches += 1;
      TRY(grad_from_P(h, h->B.p, lambda));
    } else {
      TRY(grad_eval(h, h->B.p, lambda));              // 412-422 (incl. nr -= lambda*B)
    }
    if (lambda != 0.0) {
      dot(h->st, n, h->B.p, h->B.p, h->dot_scratch, h->dscal + 23);   // 428
      CKL();
      h->stats.launches += 2;
    }
    // beta = (|nr|/|r|)^2 (423), r = nr (424), C (427-428), cost line (429), |r| < cconv (432)
    cg_after_grad(h->st, h->G.p + n, cgs, lambda, h->dscal + 23, (double)h->NTg, cconv);
    CKL();
    xpby_dev(h->st, n, h->G.p, cgs + 1, h->p.p);      // p = r + beta p (442)
    CKL();
    h->stats.launches += 2;
  }
  return 0;
}

static void cgrad_collect(const double* hc, double* cost_per_pass, double* rnorm_per_pass, int* npass_done) {
  int nd = (int)hc[3];
  if (nd > 8) nd = 8;
  for (int i = 0; i < nd; ++i) {
    if (cost_per_pass) cost_per_pass[i] = hc[8 + i];
    if (rnorm_per_pass) rnorm_per_pass[i] = hc[16 + i];
  }
  if (npass_done) *npass_done = nd;
}

int tnml_cgrad(tnml_handle h, int Npass, double lambda, double cconv, double* cost_per_pass,
               double* rnorm_per_pass, int* npass_done) {
  if (!h || !h->bond_valid) return h ? fail(h, TNML_ERR_INVALID, "no bond tensor formed") : TNML_ERR_INVALID;
  if (Npass < 1) return fail(h, TNML_ERR_INVALID, "Npass must be >= 1");
  CK(cudaSetDevice(h->device));
  TRY(cgrad_enqueue(h, Npass, lambda, cconv));
  double hc[24];
  TRY(fetch(h, h->cgs, 24, hc));                      // the one read-back of the whole CG
  cgrad_collect(hc, cost_per_pass, rnorm_per_pass, npass_done);
  return TNML_OK;
}

int tnml_svd_split(tnml_handle h, int dir, double cutoff, int maxm, int minm, int do_rel_cutoff, int* newm,
                   double* truncerr) {
  if (!h || !h->bond_valid) return h ? fail(h, TNML_ERR_INVALID, "no bond tensor formed") : TNML_ERR_INVALID;
  if (dir != TNML_FROMLEFT && dir != TNML_FROMRIGHT) return fail(h, TNML_ERR_INVALID, "bad direction %d", dir);
  if (maxm < 1) return fail(h, TNML_ERR_INVALID, "maxm must be >= 1");
  CK(cudaSetDevice(h->device));
  const int b = h->currb;
  const BondGeom g = h->g;
  const int nlA = g.lab_b ? NL : 1, nlB = g.lab_b1 ? NL : 1;
  const int nA = 2 * g.ml * nlA, nB = 2 * g.mr * nlB;
  const int keep = std::min(std::min(nA, nB), maxm);
  // new buffers (the old site tensors are not inputs of the SVD: B is)
  TRY(set_site_dev(h, b, g.ml, keep, g.lab_b));
  TRY(set_site_dev(h, b + 1, keep, g.mr, g.lab_b1));
  int m = 0, sweeps = 0;
  double terr = 0.0;
  int rc;
  h->svd.hint_m = std::max(h->reserve_m, maxm < 4096 ? maxm : 0);
  {
    PhaseTimer t(h, PH_SVD);
    rc = svd_split(h->st, h->svd, h->B.p, g, dir, cutoff, maxm, minm, do_rel_cutoff, h->W[b].d, h->W[b + 1].d, &m,
                   &terr, &sweeps, (long*)&h->stats.launches);
  }
  if (rc == -2) return fail(h, TNML_ERR_CUDA, "svd kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc == -5) return fail(h, TNML_ERR_NOCONV, "Jacobi SVD did not converge in 60 sweeps");
  h->W[b].mr = m;
  h->W[b + 1].ml = m;
  h->stats.alg_flops += 4.0 * std::max(nA, nB) * (double)std::min(nA, nB) * std::min(nA, nB);
  h->last_svd_sweeps = sweeps;
  if (newm) *newm = m;
  if (truncerr) *truncerr = terr;
  return TNML_OK;
}

// quadcost, enqueue half: forward pass + all-reduce of the statistics into dscal[0..15], lambda*|X|^2
// term into dscal[18].  The caller reads dscal back (one read-back for everything it needs).
static int quadcost_enqueue(tnml_handle h, int use_sites, double lambda) {
  const double* X;
  if (use_sites) {
    const int b = h->currb;
    TRY(setup_geom(h, b));
    TRY(ensure(h, h->T, (size_t)h->g.size()));
    form_bond(h->st, h->W[b].d, h->W[b + 1].d, h->W[b].mr, h->g, h->T.p);
    CKL();
    h->stats.launches += 1;
    X = h->T.p;
  } else {
    if (!h->bond_valid) return fail(h, TNML_ERR_INVALID, "no bond tensor formed");
    X = h->B.p;
  }
  TRY(forward(h, X, FAT_COST, h->dscal));
  TRY(allreduce(h, h->dscal, 16));
  if (lambda != 0.0) {
    dot(h->st, h->g.size(), X, X, h->dot_scratch, h->dscal + 18);
    CKL();
    h->stats.launches += 2;
  }
  return 0;
}
static void quadcost_collect(const double* hs, double lambda, double* C, double* C_label, int64_t* ncorrect) {
  double c = 0.0;
  for (int l = 0; l < NL; ++l) {
    c += hs[l];
    if (C_label) C_label[l] = hs[l];
  }
  if (lambda != 0.0) c += lambda * hs[18];  // fixedL.cc:329
  if (C) *C = c;
  if (ncorrect) *ncorrect = (int64_t)llround(hs[10]);
}

int tnml_quadcost(tnml_handle h, int use_sites, double lambda, double* C, double* C_label, int64_t* ncorrect) {
  if (!h || h->currb < 1) return h ? fail(h, TNML_ERR_INVALID, "no current bond") : TNML_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  TRY(quadcost_enqueue(h, use_sites, lambda));
  double hs[19];
  TRY(fetch(h, h->dscal, 19, hs));
  quadcost_collect(hs, lambda, C, C_label, ncorrect);
  return TNML_OK;
}

int tnml_shift_env(tnml_handle h, int b, int dir) {
  if (!h) return TNML_ERR_INVALID;
  if (b < 1 || b >= h->N) return fail(h, TNML_ERR_INVALID, "bond %d out of range", b);
  if (dir != TNML_FROMLEFT && dir != TNML_FROMRIGHT) return fail(h, TNML_ERR_INVALID, "bad direction %d", dir);
  CK(cudaSetDevice(h->device));
  const int c = (dir == TNML_FROMLEFT) ? b : b + 1;  // fixedL.cc:196
  h->moving = (dir == TNML_FROMLEFT) ? 1 : 2;
  return advance_env(h, c, dir == TNML_FROMLEFT ? 0 : 1);
}

int tnml_bond_update(tnml_handle h, int b, int ha, const tnml_bond_params* p, tnml_bond_result* out) {
  if (!h || !p) return TNML_ERR_INVALID;
  if (ha != 1 && ha != 2) return fail(h, TNML_ERR_INVALID, "ha must be 1 or 2");
  tnml_bond_result res;
  memset(&res, 0, sizeof(res));
  if (p->maxm > h->reserve_m) h->reserve_m = p->maxm;
  h->moving = ha;
  TRY(tnml_set_bond(h, b));                                            // 488
  TRY(tnml_bond_form(h));                                              // 493-498
  res.origm = h->W[b].mr;
  if (!h->bond_valid) return fail(h, TNML_ERR_INVALID, "no bond tensor formed");
  if (p->Npass < 1) return fail(h, TNML_ERR_INVALID, "Npass must be >= 1");
  TRY(cgrad_enqueue(h, p->Npass, p->lambda, p->cconv));                   // 504 (scalars read back below)
  const long n = h->g.size();
  TRY(tnml_svd_split(h, ha == 1 ? TNML_FROMLEFT : TNML_FROMRIGHT, p->cutoff, p->maxm, p->minm, p->do_rel_cutoff,
                     &res.newm, &res.truncerr));                        // 519-521
  res.svd_sweeps = h->last_svd_sweeps;
  // quadcost(newB) (527-532, newB in T), |B| and |B - newB| (528-530) and shiftE (540) are all
  // enqueued before the single read-back of their scalars, so the GPU stays busy while the host waits
  TRY(quadcost_enqueue(h, 1, p->lambda));
  dot(h->st, n, h->B.p, h->B.p, h->dot_scratch, h->dscal + 16);
  CKL();
  axpby(h->st, n, -1.0, h->T.p, 1.0, h->B.p);
  CKL();
  dot(h->st, n, h->B.p, h->B.p, h->dot_scratch, h->dscal + 17);
  CKL();
  h->stats.launches += 5;
  h->bond_valid = false;
  TRY(tnml_shift_env(h, b, ha == 1 ? TNML_FROMLEFT : TNML_FROMRIGHT));  // 540
  double hs[19 + 24];
  CK(cudaMemcpyAsync(h->hpin + 19, h->cgs, 24 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  TRY(fetch(h, h->dscal, 19, hs));                                       // one synchronisation for both
  memcpy(hs + 19, h->hpin + 19, 24 * sizeof(double));
  quadcost_collect(hs, p->lambda, &res.cost, res.cost_label, &res.ncorrect);
  cgrad_collect(hs + 19, res.cg_cost, res.cg_rnorm, &res.npass_done);
  res.normB = std::sqrt(hs[16]);
  res.dB = std::sqrt(hs[17]);
  if (out) *out = res;
  return TNML_OK;
}

int tnml_predict(tnml_handle h, int32_t* labels_out, double* P_out) {
  if (!h || h->NT == 0) return h ? fail(h, TNML_ERR_INVALID, "no images") : TNML_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  if (labels_out) CK(cudaMemcpyAsync(labels_out, h->pred, h->NT * sizeof(int32_t), cudaMemcpyDeviceToHost, h->st));
  if (P_out) CK(cudaMemcpyAsync(P_out, h->P, (size_t)h->NT * NL * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return TNML_OK;
}

int tnml_fulltest(tnml_handle h, int32_t* pred_out, int64_t* ncorrect) {
  if (!h || h->N == 0) return h ? fail(h, TNML_ERR_INVALID, "no images") : TNML_ERR_INVALID;
  TRY(tnml_init_envs(h));           // right environments of every image, then setBond(1)
  double C = 0.0;
  int64_t nc = 0;
  TRY(tnml_quadcost(h, 1, 0.0, &C, nullptr, &nc));   // P = full contraction, argmax |P_l|
  if (ncorrect) *ncorrect = nc;
  if (pred_out) TRY(tnml_predict(h, pred_out, nullptr));
  return TNML_OK;
}

int tnml_get_env(tnml_handle h, int slot, int* m, int* is_fat, double* data, size_t capacity_elems) {
  if (!h || slot < 1 || slot > h->N) return h ? fail(h, TNML_ERR_INVALID, "bad slot %d", slot) : TNML_ERR_INVALID;
  Slot& s = h->slot[slot];
  if (s.kind == 0) return fail(h, TNML_ERR_INVALID, "slot %d empty", slot);
  if (m) *m = s.m;
  if (is_fat) *is_fat = s.fat;
  if (data) {
    size_t n = (size_t)h->NT * s.m * (s.fat ? NL : 1);
    if (capacity_elems < n) return fail(h, TNML_ERR_INVALID, "buffer too small for slot %d", slot);
    CK(cudaSetDevice(h->device));
    TRY(fetch_slot(h, slot, slot, slot));
    TRY(use_slot(h, s));
    CK(cudaMemcpyAsync(data, s.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
  }
  return TNML_OK;
}

int tnml_comm_get_unique_id(uint8_t* id) {
  if (!id) return TNML_ERR_INVALID;
  std::string err;
  if (!load_nccl(err)) return fail(nullptr, TNML_ERR_NCCL, "%s", err.c_str());
  NcclId uid;
  int rc = g_nccl.get_uid(&uid);
  if (rc != 0) return fail(nullptr, TNML_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
  memcpy(id, uid.internal, TNML_UNIQUE_ID_BYTES);
  return TNML_OK;
}

int tnml_comm_init_rank(tnml_handle h, int nranks, int rank, const uint8_t* id) {
  if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return h ? fail(h, TNML_ERR_INVALID, "bad comm args") : TNML_ERR_INVALID;
  std::string err;
  if (!load_nccl(err)) return fail(h, TNML_ERR_NCCL, "%s", err.c_str());
  CK(cudaSetDevice(h->device));
  NcclId uid;
  memcpy(uid.internal, id, TNML_UNIQUE_ID_BYTES);
  void* comm = nullptr;
  int rc = g_nccl.init_rank(&comm, nranks, uid, rank);
  if (rc != 0) return fail(h, TNML_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
  h->comm = comm;
  h->nranks = nranks;
  h->rank = rank;
  return TNML_OK;
}

int tnml_comm_broadcast(tnml_handle h, double* vals, int n, int root) {
  if (!h || !vals || n < 1 || n > 16) return h ? fail(h, TNML_ERR_INVALID, "bad broadcast args (1 <= n <= 16)") : TNML_ERR_INVALID;
  if (!h->comm) return TNML_OK;
  if (root < 0 || root >= h->nranks) return fail(h, TNML_ERR_INVALID, "bad root %d", root);
  CK(cudaSetDevice(h->device));
  // sum-all-reduce with zeros from every rank but the root (the library's only collective)
  for (int i = 0; i < n; ++i) h->hpin[i] = (h->rank == root) ? vals[i] : 0.0;
  CK(cudaMemcpyAsync(h->dscal + 40, h->hpin, n * sizeof(double), cudaMemcpyHostToDevice, h->st));
  TRY(allreduce(h, h->dscal + 40, (size_t)n));
  TRY(fetch(h, h->dscal + 40, n, vals));
  return TNML_OK;
}

int tnml_set_option(tnml_handle h, const char* name, double value) {
  if (!h || !name) return TNML_ERR_INVALID;
  if (strcmp(name, "cg_reuse_forward") == 0) {
    h->cg_reuse_forward = (value != 0.0);
    return TNML_OK;
  }
  if (strcmp(name, "svd_cluster") == 0 || strcmp(name, "svd_cross") == 0 || strcmp(name, "svd_precond") == 0) {
    svd_set_variant(name, (int)value);          // process-wide (testing / A-B timing)
    return TNML_OK;
  }
  if (strcmp(name, "fat_variant") == 0) {
    fat_set_variant((int)value);
    return TNML_OK;
  }
  if (strcmp(name, "krgram_variant") == 0) {
    krgram_set_variant((int)value);
    return TNML_OK;
  }
  if (strcmp(name, "krgemm_variant") == 0) {   // per handle: 3 tcgen05 int8 (default), 2 DMMA persistent, 1 register-staged
    h->krgemm_variant = (int)value;
    krgemm_set_variant((int)value == 1 ? 1 : -1);   // the environment advance follows variant 1 (process-wide switch)
    return TNML_OK;
  }
  if (strcmp(name, "oz_slices") == 0) {
    if (value < 6 || value > 8) return fail(h, TNML_ERR_INVALID, "oz_slices must be 6, 7 or 8");
    h->oz_slices = (int)value;
    return TNML_OK;
  }
  if (strcmp(name, "env_budget_gb") == 0) {
    if (value < 0) return fail(h, TNML_ERR_INVALID, "env_budget_gb must be >= 0");
    h->env_budget = (size_t)(value * 1073741824.0);
    return TNML_OK;
  }
  if (strcmp(name, "reserve_m") == 0) {
    h->reserve_m = (int)value;
    return TNML_OK;
  }
  return fail(h, TNML_ERR_INVALID, "unknown option %s", name);
}

int tnml_get_stats(tnml_handle h, tnml_stats* out, int reset) {
  if (!h) return TNML_ERR_INVALID;
  cudaSetDevice(h->device);
  drain_events(h);
  if (out) *out = h->stats;
  if (reset) memset(&h->stats, 0, sizeof(h->stats));
  return TNML_OK;
}

int tnml_set_timing(tnml_handle h, int on) {
  if (!h) return TNML_ERR_INVALID;
  cudaSetDevice(h->device);
  drain_events(h);
  h->timing = (on != 0);
  return TNML_OK;
}

int tnml_synchronize(tnml_handle h) {
  if (!h) return TNML_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaStreamSynchronize(h->cp));
  CK(cudaStreamSynchronize(h->cpin));
  return TNML_OK;
}

void* tnml_stream(tnml_handle h) { return h ? (void*)h->st : nullptr; }

}  // extern "C"
