// fixedL.cc -- drop-in for the reference program `fixedL <inputfile>`
// (/root/reference/fixedL.cc): same input-file keys, same files in the working
// directory (W, sites, WRITE_WF, LAMBDA), same log lines, same function names
// (TState, TrainStates::{init,setBond,shiftE}, quadcost, cgrad, mldmrg).  The
// per-bond arithmetic the reference does with ITensor on host threads is done
// by libtnml_b200.so on the GPU through the C-ABI (include/tnml_b200.h).
//
// Deliberate differences (all logged at start-up):
//   * `imglen` is honoured (2x2 block mean, image.h:316-346); the reference parses
//     nothing of the sort and always runs 28x28.  Default 28.
//   * the initial W is built like the reference (sum of `ninitial` random product states per
//     label, fixedL.cc:702-728, or the W0..W9 merge, 682-701) but the random picks come from a
//     SEEDED generator (`seed` key; seed < 0 = time-seeded like ITensor's Global::random(),
//     SURVEY F7).  `initial = random` selects a seeded random MPS of bond dimension `minitial`
//     instead; `init_only = yes` stops after writing `W` (no GPU needed).
//   * `W` / `sites` files use this program's own binary format.
//   * `trace = <file>` appends one JSON line per bond update (sweep, bond, link dimensions, truncation
//     error, cost, #correct, wall time, kernel launches, algorithmic bytes / flops from tnml_get_stats;
//     `trace_phases = yes` adds the CUDA-event phase times) next to the reference's log lines.
//   * nthread/Nbatch are accepted and validated like the reference but the
//     parallelism is GPUs: one process per GPU (TNML_RANK / TNML_WORLD_SIZE),
//     images sharded with ParallelDo's bounds (paralleldo.h:32-43).
#include <unistd.h>

#include <array>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "../../include/tnml_b200.h"
#include "itensor_lite.h"
#include "initial_w.h"
#include "mnist.h"

using namespace itensor;
using std::array;
using std::max;
using std::min;
using std::string;
using std::vector;

const size_t NL = 10;
const auto Label = IndexType("Label");

#define TN(call)                                                                   \
  do {                                                                             \
    int rc_ = (call);                                                              \
    if (rc_ != 0) Error(format("%s failed (%d): %s", #call, rc_, tnml_last_error(h_))); \
  } while (0)

// Represents a range of integers (paralleldo.h:8-18)
struct Bound {
  size_t n = 0, begin = 0, end = 0;
  size_t size() const { return end - begin; }
};
vector<Bound> ParallelDoBounds(int Nthread, size_t Ntask) {  // paralleldo.h:32-43
  vector<Bound> b(Nthread);
  size_t th = Ntask / Nthread, c = 0;
  for (int n = 0; n < Nthread; ++n) {
    b[n] = Bound{(size_t)n, c, c + th};
    c += th;
  }
  b.back().end = Ntask;
  return b;
}

// Struct holding info about training "states" (fixedL.cc:18-62)
struct TState {
  long n = -1;
  int l = -1;
  int d = 0;
  vector<Real> data;
  template <typename Func, typename ImgType>
  TState(int n_, int l_, int d_, ImgType const& img, Func const& phi) : n(n_), l(l_), d(d_) {
    data.resize(img.size() * d);
    size_t i = 0;
    for (size_t j = 1; j <= img.size(); ++j)
      for (int k = 1; k <= d; ++k) data[i++] = phi(img(j), k);
  }
  Real operator()(int i, int k) const { return data.at(d * i + k - d - 1); }  // 1-indexed
};

class TrainStates {
 public:
  tnml_handle h_ = nullptr;
  int N = 0;
  long NTlocal = 0, NTglobal = 0;
  int currb_ = -1;
  int rank_ = 0, world_ = 1;
  vector<Index> links;  // links[j] joins sites j and j+1
  Index L;
  SiteSet sites_;

  TrainStates(vector<TState>&& ts, int N_, SiteSet const& sites, int Nthread, int Nbatch, int device, int rank,
              int world)
      : N(N_), rank_(rank), world_(world), sites_(sites) {
    const int totNtrain = (int)ts.size();
    if (totNtrain % Nbatch != 0) {  // fixedL.cc:84-89
      printfln("totNtrain=%d, Nbatch=%d, totNtrain%%Nbatch=%d", totNtrain, Nbatch, totNtrain % Nbatch);
      Error("totNtrain not commensurate with Nbatch");
    }
    if (Nthread > 16) Error("Need to increase size of futs");  // paralleldo.h:55-56
    NTglobal = totNtrain;
    auto bs = ParallelDoBounds(world, totNtrain);
    for (auto& b : bs) printfln("GPU shard %d %d -> %d (%d)", (int)b.n, (int)b.begin, (int)b.end, (int)b.size());
    const Bound mine = bs.at(rank);
    NTlocal = (long)mine.size();
    vector<double> feat((size_t)NTlocal * N * 2);
    vector<int32_t> labels(NTlocal);
    for (long i = 0; i < NTlocal; ++i) {
      auto& t = ts.at(mine.begin + i);
      std::memcpy(&feat[(size_t)i * N * 2], t.data.data(), sizeof(double) * N * 2);
      labels[i] = t.l;
    }
    if (tnml_create(device, 0, &h_) != 0) Error(string("tnml_create: ") + tnml_last_error(nullptr));
    TN(tnml_set_images(h_, NTlocal, N, feat.data(), labels.data(), NTglobal, (long)mine.begin));
    if (world > 1) initComm();
  }
  ~TrainStates() {
    if (h_) tnml_destroy(h_);
  }
  int size() const { return (int)NTglobal; }

  // MPS site tensors travel as [ml][d][mr]([NL]) == index order (left, site, right [, L])
  void upload(MPS const& W, int j) {
    auto const& A = W.A(j);
    const bool lab = (bool)findtype(A, Label);
    const long ml = links.at(j - 1).m(), mr = links.at(j).m();
    if ((long)A.data().size() != ml * 2 * mr * (lab ? (long)NL : 1)) Error(format("site %d has unexpected size", j));
    TN(tnml_set_site(h_, j, (int)ml, (int)mr, lab ? 1 : 0, A.data().data()));
  }
  void download(MPS& W, int j) {
    int ml, mr, lab;
    TN(tnml_get_site_dims(h_, j, &ml, &mr, &lab));
    if (links.at(j - 1).m() != ml) links.at(j - 1) = Index(format("l%d", j - 1), ml, Link);
    if (links.at(j).m() != mr) links.at(j) = Index(format("l%d", j), mr, Link);
    vector<Index> is{links[j - 1], sites_(j), links[j]};
    if (lab) is.push_back(L);
    vector<Real> d((size_t)ml * 2 * mr * (lab ? NL : 1));
    TN(tnml_get_site(h_, j, d.data(), d.size()));
    W.setA(j, ITensor(is, std::move(d)));
  }

  // size the environment slots for link dimension m so that they never grow during the sweeps
  void reserve(int m) { TN(tnml_set_option(h_, "reserve_m", (double)m)); }
  void init(MPS const& W) {  // fixedL.cc:122-157
    for (int j = 1; j <= N; ++j) upload(W, j);
    TN(tnml_init_envs(h_));
    currb_ = 1;
  }
  void setBond(int b) {  // fixedL.cc:159-190
    if (currb_ == b) return;
    currb_ = b;
    TN(tnml_set_bond(h_, b));
  }
  void shiftE(MPS const&, int b, Direction dir) {  // fixedL.cc:192-233
    const int c = (dir == Fromleft) ? b : b + 1;
    const int prevc = (dir == Fromleft) ? b - 1 : b + 2;
    if (prevc >= 1 && prevc <= N)
      printfln("## Advancing E from %d to %d", prevc, c);
    else
      printfln("## Making new E at %d", c);
    TN(tnml_shift_env(h_, b, dir == Fromleft ? TNML_FROMLEFT : TNML_FROMRIGHT));
  }

 private:
  void initComm() {
    // rank 0 publishes the NCCL id through a file in the working directory
    const char* f = "tnml_nccl_uid";
    uint8_t id[TNML_UNIQUE_ID_BYTES];
    if (rank_ == 0) {
      if (tnml_comm_get_unique_id(id) != 0) Error(string("tnml_comm_get_unique_id: ") + tnml_last_error(nullptr));
      std::ofstream o(string(f) + ".tmp", std::ios::binary);
      o.write((char*)id, sizeof(id));
      o.close();
      std::rename((string(f) + ".tmp").c_str(), f);
    } else {
      for (int t = 0; t < 600 && !fileExists(f); ++t) std::this_thread::sleep_for(std::chrono::milliseconds(100));
      std::ifstream i(f, std::ios::binary);
      if (!i.read((char*)id, sizeof(id))) Error("could not read tnml_nccl_uid");
    }
    TN(tnml_comm_init_rank(h_, world_, rank_, id));
    if (rank_ == 0) {
      std::this_thread::sleep_for(std::chrono::milliseconds(500));
      std::remove(f);
    }
  }
};

// Compute squared distance of the actual output of the model from the ideal
// output (fixedL.cc:280-344).  B lives on the device: use_sites selects
// newB = W.A(c)*W.A(c+dc) (fixedL.cc:527,532) or the current bond tensor.
Real quadcost(bool use_sites, TrainStates const& ts, Args const& args = Args::global(), int64_t* ncor_out = nullptr) {
  auto h_ = ts.h_;
  auto NT = ts.size();
  auto lambda = args.getReal("lambda", 0.);
  auto showlabels = args.getBool("ShowLabels", false);
  double C = 0, CL[NL];
  int64_t ncor = 0;
  TN(tnml_quadcost(h_, use_sites ? 1 : 0, lambda, &C, CL, &ncor));
  if (showlabels)
    for (size_t l = 0; l < NL; ++l) printfln("  Label l=%d C%d = %.10f", (int)l, (int)l, CL[l] / NT);
  if (ncor_out) *ncor_out = ncor;
  long ninc = NT - ncor;
  printfln("Percent correct = %.4f%%, # incorrect = %d/%d", ncor * 100. / NT, (int)ninc, (int)(ncor + ninc));
  return C;
}

// Conjugate gradient (fixedL.cc:349-445) on the device bond tensor
void cgrad(TrainStates& ts, Args const& args) {
  auto h_ = ts.h_;
  auto Npass = args.getInt("Npass");
  auto lambda = args.getReal("lambda", 0.);
  auto cconv = args.getReal("cconv", 1E-10);
  printfln("In cgrad, lambda = %.3E", lambda);
  double costs[8], rn[8];
  int nd = 0;
  TN(tnml_cgrad(h_, Npass, lambda, cconv, costs, rn, &nd));
  for (int pass = 1; pass <= Npass; ++pass) {
    println("  Conj grad pass ", pass);
    if (pass > nd) break;
    printfln("  Cost = %.10f", costs[pass - 1]);
    if (rn[pass - 1] < cconv) {
      printfln("  |r| = %.1E < %.1E, breaking", rn[pass - 1], cconv);
      break;
    }
    printfln("  |r| = %.1E", rn[pass - 1]);
  }
}

// M.L. DMRG (fixedL.cc:451-570)
void mldmrg(MPS& W, TrainStates& ts, Sweeps const& sweeps, Args args) {
  auto h_ = ts.h_;
  auto N = W.N();
  auto NT = ts.size();
  auto method = args.getString("Method");
  auto pause_step = args.getBool("PauseStep", false);
  auto do_rel = args.getBool("DoRelCutoff", false);
  auto cargs = Args{args, "Normalize", false};
  // per-bond JSONL trace (rank 0): off unless the input file names a `trace` file
  FILE* trace = nullptr;
  auto tracefile = args.getString("Trace", "");
  auto trace_phases = args.getBool("TracePhases", false);
  if (!tracefile.empty() && ts.rank_ == 0) {
    trace = std::fopen(tracefile.c_str(), "a");
    if (!trace) Error(format("cannot open trace file \"%s\"", tracefile.c_str()));
  }
  if (trace && trace_phases) TN(tnml_set_timing(h_, 1));
  tnml_stats st;

  for (int sw = 1; sw <= sweeps.nsweep(); ++sw) {
    printfln("\nSweep %d maxm=%d minm=%d", sw, sweeps.maxm(sw), sweeps.minm(sw));
    for (int b = 1, ha = 1; ha <= 2; sweepnext(b, ha, N)) {
      auto c = (ha == 1) ? b : b + 1;
      auto dc = (ha == 1) ? +1 : -1;
      (void)dc;
      auto tb0 = std::chrono::steady_clock::now();
      if (trace) TN(tnml_get_stats(h_, &st, 1));
      ts.setBond(b);
      printfln("Sweep %d Half %d Bond %d", sw, ha, c);
      int origm = 0;
      TN(tnml_get_site_dims(h_, b, nullptr, &origm, nullptr));
      TN(tnml_bond_form(h_));  // oB = W.A(c)*W.A(c+dc); B = oB

      if (method == "conj")
        cgrad(ts, args);
      else
        Error(format("method type \"%s\" not recognized", method.c_str()));
      printfln("Sweep %d Half %d Bond %d", sw, ha, c);

      // SVD B back apart into MPS tensors
      int newm = 0;
      double truncerr = 0;
      TN(tnml_svd_split(h_, ha == 1 ? TNML_FROMLEFT : TNML_FROMRIGHT, sweeps.cutoff(sw), sweeps.maxm(sw),
                        sweeps.minm(sw), do_rel ? 1 : 0, &newm, &truncerr));
      printfln("SVD trunc err = %.2E", truncerr);
      printfln("Original m=%d, New m=%d", origm, newm);

      auto largs = Args{cargs, "ShowLabels", true, "lambda", args.getReal("lambda", 0.)};
      int64_t ncor = 0;
      auto newC = quadcost(true, ts, largs, &ncor);
      printfln("--> After SVD, Cost = %.10f", newC / NT);

      // Update E's (MPS environment tensors)
      ts.shiftE(W, b, ha == 1 ? Fromleft : Fromright);
      if (trace) {
        TN(tnml_synchronize(h_));
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count();
        TN(tnml_get_stats(h_, &st, 0));
        std::fprintf(trace,
                     "{\"sweep\": %d, \"half\": %d, \"bond\": %d, \"b\": %d, \"origm\": %d, \"newm\": %d, \"truncerr\": %.6e, "
                     "\"cost\": %.12e, \"ncorrect\": %lld, \"NT\": %d, \"wall_ms\": %.4f, \"launches\": %lld, "
                     "\"alg_bytes\": %.6e, \"alg_flops\": %.6e",
                     sw, ha, c, b, origm, newm, truncerr, newC / NT, (long long)ncor, (int)NT, ms, (long long)st.launches,
                     st.alg_bytes, st.alg_flops);
        if (trace_phases)
          std::fprintf(trace,
                       ", \"phase_ms\": {\"proj\": %.4f, \"grad\": %.4f, \"fat\": %.4f, \"svd\": %.4f, \"shift\": %.4f, "
                       "\"other\": %.4f}",
                       st.ms_proj, st.ms_grad, st.ms_fat, st.ms_svd, st.ms_shift, st.ms_other);
        std::fprintf(trace, "}\n");
        std::fflush(trace);
      }

      // Sentinel files (fixedL.cc:542-559).  The reference is ONE process; here rank 0 alone looks at
      // (and removes) the files and its findings are broadcast, so that every rank writes W at the
      // same bond and continues with the same lambda (r = G - lambda*B and the replicated SVD must
      // stay bit-identical across ranks).
      double ctl[3] = {0., 0., 0.};   // WRITE_WF seen, LAMBDA seen, new lambda
      if (ts.rank_ == 0) {
        if (fileExists("WRITE_WF")) {
          std::remove("WRITE_WF");
          ctl[0] = 1.;
        }
        if (fileExists("LAMBDA")) {
          std::ifstream lf("LAMBDA");
          Real lambda = 0.;
          lf >> lambda;
          lf.close();
          std::remove("LAMBDA");
          ctl[1] = 1.;
          ctl[2] = lambda;
        }
      }
      if (ts.world_ > 1) TN(tnml_comm_broadcast(h_, ctl, 3, 0));
      if (ctl[0] != 0.) {
        println("File WRITE_WF found");
        println("Writing W to disk");
        for (int j = 1; j <= N; ++j) ts.download(W, j);
        if (ts.rank_ == 0) writeToFile("W", W);
      }
      if (ctl[1] != 0.) {
        args.add("lambda", ctl[2]);
        println("new lambda = ", ctl[2]);
      }
      if (pause_step) {
        println("(Paused, press enter to continue)");
        getchar();
      }
    }  // loop over c,dc
    println("Writing W to disk");
    for (int j = 1; j <= N; ++j) ts.download(W, j);
    if (ts.rank_ == 0) writeToFile("W", W);
  }  // loop over sweeps
  if (trace) std::fclose(trace);
}

// ---- deterministic initial W (replaces fixedL.cc:702-728) ---------------------
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return ((next() >> 11) + 0.5) / 9007199254740992.0; }
  double normal() {
    double u = uniform(), v = uniform();
    return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v);
  }
};

// rows of M (r x c, r <= c) -> orthonormal rows; returns R with M_old = R * M_new
static vector<double> orthonormalize_rows(vector<double>& M, long r, long c) {
  vector<double> R(r * r, 0.0);
  for (long i = 0; i < r; ++i) {
    for (int rep = 0; rep < 2; ++rep)  // MGS with re-orthogonalisation
      for (long k = 0; k < i; ++k) {
        double dot = 0;
        for (long j = 0; j < c; ++j) dot += M[i * c + j] * M[k * c + j];
        for (long j = 0; j < c; ++j) M[i * c + j] -= dot * M[k * c + j];
        R[i * r + k] += dot;
      }
    double nrm = 0;
    for (long j = 0; j < c; ++j) nrm += M[i * c + j] * M[i * c + j];
    nrm = std::sqrt(nrm);
    if (nrm == 0) Error("rank-deficient site tensor in initial W");
    for (long j = 0; j < c; ++j) M[i * c + j] /= nrm;
    R[i * r + i] = nrm;
  }
  return R;
}

MPS randomMPS(SiteSet const& sites, vector<Index>& links, Index const& L, int c, int m, uint64_t seed) {
  const int N = sites.N(), d = (int)sites(1).m();
  links.assign(N + 1, Index());
  for (int j = 0; j <= N; ++j) {
    long cap = std::min(j, N - j);
    long dim = (j == 0 || j == N) ? 1 : (cap >= 40 ? (long)m : std::min<long>(m, 1L << cap));
    links[j] = Index(format("l%d", j), dim, Link);
  }
  Rng rng(seed);
  vector<vector<double>> A(N + 1);
  const double noise = 0.3;
  for (int j = 1; j <= N; ++j) {
    const long ml = links[j - 1].m(), mr = links[j].m(), nl = (j == c) ? (long)NL : 1;
    A[j].assign(ml * d * mr * nl, 0.0);
    vector<double> lw(nl, 1.0);
    if (j == c)
      for (auto& w : lw) w = 1.0 + 0.5 * rng.normal();
    for (long a = 0; a < ml; ++a)
      for (int s = 0; s < d; ++s)
        for (long b = 0; b < mr; ++b)
          for (long l = 0; l < nl; ++l) {
            double v = noise * rng.normal() / std::sqrt((double)std::max(ml, mr));
            if (s == 0 && a == b) v += lw[l];
            A[j][((a * d + s) * mr + b) * nl + l] = v;
          }
  }
  for (int j = N; j >= 2; --j) {  // right-canonicalise, centre ends on site 1
    const long ml = links[j - 1].m(), mr = links[j].m(), nl = (j == c) ? (long)NL : 1;
    const long cols = d * mr * nl;
    if (ml > cols) Error("initial link dimension too large");
    auto R = orthonormalize_rows(A[j], ml, cols);
    double rn = 0;
    for (double v : R) rn += v * v;
    rn = std::sqrt(rn / ml);
    // A[j-1][..., b', ...] = sum_b A[j-1][..., b, ...] * R[b][b'] / rn  (R lower-triangular: old = R * new)
    const long pml = links[j - 2].m(), pnl = (j - 1 == c) ? (long)NL : 1;
    vector<double> out(A[j - 1].size(), 0.0);
    for (long a = 0; a < pml; ++a)
      for (int s = 0; s < d; ++s)
        for (long l = 0; l < pnl; ++l)
          for (long bp = 0; bp < ml; ++bp) {
            double acc = 0;
            for (long b = 0; b < ml; ++b) acc += A[j - 1][((a * d + s) * ml + b) * pnl + l] * R[b * ml + bp];
            out[((a * d + s) * ml + bp) * pnl + l] = acc / rn;
          }
    A[j - 1].swap(out);
  }
  auto normalise = [](vector<double>& v) {
    double n = 0;
    for (double x : v) n += x * x;
    n = std::sqrt(n);
    for (double& x : v) x /= n;
  };
  normalise(A[1]);
  normalise(A[c]);  // W.Aref(c) /= norm(W.A(c))  (fixedL.cc:725)
  MPS W(sites);
  for (int j = 1; j <= N; ++j) {
    vector<Index> is{links[j - 1], sites(j), links[j]};
    if (j == c) is.push_back(L);
    W.setA(j, ITensor(is, std::move(A[j])));
  }
  return W;
}

int main(int argc, const char* argv[]) {
  if (argc != 2) {
    printfln("Usage: %s inputfile", argv[0]);
    return 0;  // fixedL.cc:579-583
  }
  try {
    auto input = InputGroup(argv[1], "input");
    int d = 2;
    auto datadir = input.getString("datadir", "/Users/mstoudenmire/software/tnml/mllib/MNIST");
    auto Ntrain = input.getInt("Ntrain", 60000);
    auto Nbatch = input.getInt("Nbatch", 10);
    auto Nsweep = input.getInt("Nsweep", 50);
    auto cutoff = input.getReal("cutoff", 1E-10);
    auto maxm = input.getInt("maxm", 5000);
    auto minm = input.getInt("minm", max(10, maxm / 2));
    auto ninitial = input.getInt("ninitial", 100);
    auto Nthread = input.getInt("nthread", 1);
    auto replace = input.getYesNo("replace", false);
    auto pause_step = input.getYesNo("pause_step", false);
    auto lambda = input.getReal("lambda", 0.);
    auto method = input.getString("method", "conj");
    auto alpha = input.getReal("alpha", 0.01);
    auto clip = input.getReal("clip", 1.0);
    auto Npass = input.getInt("Npass", 4);
    auto cconv = input.getReal("cconv", 1E-10);
    // keys this build adds
    auto imglen = input.getInt("imglen", 28);
    auto seed = input.getInt("seed", 1);
    auto minitial = input.getInt("minitial", 10);
    auto device = input.getInt("device", -1);
    auto dorel = input.getYesNo("dorelcutoff", false);
    auto initial = input.getString("initial", "sum");
    auto init_only = input.getYesNo("init_only", false);
    auto tracefile = input.getString("trace", "");
    auto trace_phases = input.getYesNo("trace_phases", false);

    int rank = std::getenv("TNML_RANK") ? atoi(std::getenv("TNML_RANK")) : 0;
    int world = std::getenv("TNML_WORLD_SIZE") ? atoi(std::getenv("TNML_WORLD_SIZE")) : 1;
    if (device < 0) device = rank;
    // every rank builds the same initial W from the same seed; a time seed would give each rank its own
    if (world > 1 && seed < 0) Error("seed < 0 (time-seeded initial W) needs an explicit seed when TNML_WORLD_SIZE > 1");

    auto train = mllib::readMNIST(datadir, mllib::Train, Ntrain);
    if (imglen != 28) {
      printfln("imglen = %d: reducing images by block mean (not in the reference, SURVEY F4)", imglen);
      mllib::reduce(train, imglen);
    }
    auto N = (int)train.front().size();
    auto c = N / 2;
    printfln("%d sites of dimension %d", N, d);
    SiteSet sites;
    if (fileExists("sites")) {
      sites = readFromFile<SiteSet>("sites");
      if (sites(1).m() != (long)d) {
        printfln("Error: d=%d but dimension of first site is %d", d, (int)sites(1).m());
        return 1;
      }
    } else {
      sites = SiteSet(N, d);
      if (rank == 0) writeToFile("sites", sites);
    }

    // Local feature map (fixedL.cc:637-642)
    auto phi = [](Real g, int n) -> Real {
      if (g < 0 || g > 255.) Error(format("Expected g=%f to be in [0,255]", g));
      auto x = g / 255.;
      return std::pow(x / 4., n - 1);
    };

    println("Converting training set to MPS");
    auto states = vector<TState>();
    auto n = 1;
    for (auto& img : train) states.emplace_back(n++, img.label, d, img, phi);
    int totNtrain = (int)states.size();
    printfln("Total of %d training images", totNtrain);

    Index L;
    MPS W;
    vector<Index> wlinks;
    bool have_links = false;
    if (fileExists("W")) {
      println("Reading W from disk");
      W = readFromFile<MPS>("W", sites);
      L = findtype(W.A(c), Label);
      if (!L) {
        printfln("Expected W to have Label type Index at site %d", c);
        return 1;
      }
    } else if (fileExists("W0")) {
      // fixedL.cc:682-701
      println("Found separate W0,W1,...,W9 MPS: summing");
      L = Index("L", 10, Label);
      vector<initw::RMPS> ipsis;
      for (int n = 0; n < (int)NL; ++n) {
        auto in = initw::from_mps(readFromFile<MPS>(format("W%d", n), sites));
        initw::tag_label(in, c, n, (int)NL, 1.0);
        ipsis.push_back(std::move(in));
      }
      printfln("Summing all %d label states together", (int)ipsis.size());
      auto R = initw::sum(ipsis, 1E-10, 1000000, dorel);
      W = initw::to_mps(R, sites, wlinks, L, c);
      have_links = true;
      println("Done making initial W");
      if (rank == 0) writeToFile("W", W);
    } else if (initial == "random") {
      L = Index("L", 10, Label);
      printfln("Making initial W: seeded random MPS (seed=%d, m=%d), label on site %d", seed, minitial, c);
      W = randomMPS(sites, wlinks, L, c, minitial, (uint64_t)seed);
      have_links = true;
      println("Done making initial W");
      if (rank == 0) writeToFile("W", W);
    } else {
      // fixedL.cc:702-728: make initial W by summing training states together
      L = Index("L", 10, Label);
      Rng rng(seed < 0 ? (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count() : (uint64_t)seed);
      vector<vector<long>> picks;
      auto R = initw::initial_sum(N, d, c, (int)NL, train, phi, ninitial, [&rng]() { return rng.uniform(); }, dorel,
                                  &picks);
      for (int n = 0; n < (int)NL; ++n) {   // which images were drawn (positions in the training set)
        string line = format("  label %d picks:", n);
        for (long w : picks[n]) line += format(" %ld", w);
        println(line);
      }
      W = initw::to_mps(R, sites, wlinks, L, c);
      have_links = true;
      println("Done making initial W");
      if (rank == 0) writeToFile("W", W);
      printfln("overlap(W,W) = %.12f", initw::overlap(R, R));
    }
    if (!have_links) {
      wlinks.assign(N + 1, Index());
      for (int j = 1; j <= N; ++j) {
        auto const& is = W.A(j).inds();
        wlinks[j - 1] = is.at(0);
        wlinks[j] = is.at(2);
      }
    }
    println("Done making initial W");
    if (!findtype(W.A(c), Label)) Error(format("Label Index not on site %d", c));
    if (init_only) return 0;

    auto ts = TrainStates(std::move(states), N, sites, Nthread, Nbatch, device, rank, world);
    printfln("%s", tnml_version());
    ts.links = wlinks;
    ts.L = L;
    train.clear();  // to save memory
    if (!findtype(W.A(c), Label)) Error(format("Label Index not on site %d", c));

    // Project training states (product states) into environment of W MPS
    printf("Projecting training states...");
    ts.reserve(maxm);
    ts.init(W);
    println("done");

    println("Calling quadcost...");
    auto h_ = ts.h_;
    TN(tnml_bond_form(h_));
    auto C = quadcost(false, ts, {"lambda", lambda});
    printfln("Before starting DMRG Cost = %.10f", C / totNtrain);

    auto sweeps = Sweeps(Nsweep, minm, maxm, cutoff);
    auto args = Args{"lambda", lambda, "Method",  method,  "Npass",     Npass,      "alpha",       alpha, "clip",
                     clip,     "cconv", cconv, "Replace", replace, "PauseStep", pause_step, "DoRelCutoff", dorel,
                     "Trace", tracefile, "TracePhases", trace_phases};
    auto t0 = std::chrono::steady_clock::now();
    mldmrg(W, ts, sweeps, args);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printfln("mldmrg: %d sweeps, %d bond updates in %.3f s = %.2f bond-updates/sec", Nsweep, Nsweep * 2 * (N - 1), secs,
             Nsweep * 2 * (N - 1) / secs);

    println("Writing W to disk");
    if (rank == 0) writeToFile("W", W);
  } catch (ITError const& e) {
    fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  }
  return 0;
}
