// fixedl_ref_cpu.cpp -- CPU restatement (C++17, float64) of the reference's per-bond
// hot loops, in the reference's own LITERAL formulation.  TEST / BASELINE
// INFRASTRUCTURE ONLY (same status as oracle/fixedl_oracle.py): it is the
// `cpu_baseline` / `--impl reference` leg of bench.py and a second, independent
// pin of the numpy oracle.  PARITY UNPINNED: ITensor is absent (see the oracle
// header); this file follows /root/reference/fixedL.cc line by line instead:
//
//   setBond   fixedL.cc:159-190  dense t.v = A(b) x A(b+1) x LE x RE per image
//   cgrad     fixedL.cc:349-445  per-image P = B*t.v, dP = delta - P,
//                                tensors[nt] += dP*dag(t.v), per-thread partials
//                                reduced serially in thread order (385,402,421)
//   quadcost  fixedL.cc:280-344  per-label costs, argmax |P_l| (util.h:42-57)
//   threads   paralleldo.h:32-67 contiguous bounds, one std::async per bound
//
// Environments are kept in RAM (the reference spills them to proj_images/ on
// disk), ITensor's per-operation allocation/index matching is not reproduced:
// this is a LOWER bound on the reference's time.
//
// I/O: reads a little-endian binary problem file written by
// oracle/cpu_ref.py::write_problem, writes B after cgrad + costs.
//   usage: fixedl_ref_cpu <problem.bin> <result.bin> <nthread> [reps]
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <future>
#include <vector>

static const int NL = 10;  // fixedL.cc:15

struct Bound {
  size_t n, begin, end;
};
// ParallelDo(Nthread, Ntask), paralleldo.h:32-43
static std::vector<Bound> make_bounds(int nthread, size_t ntask) {
  std::vector<Bound> b(nthread);
  size_t th = ntask / nthread, c = 0;
  for (int n = 0; n < nthread; ++n) {
    b[n] = {(size_t)n, c, c + th};
    c += th;
  }
  b.back().end = ntask;
  return b;
}
template <class F>
static void parallel_do(const std::vector<Bound>& bs, F f) {  // paralleldo.h:51-67
  std::vector<std::future<void>> futs;
  for (auto& b : bs) futs.push_back(std::async(std::launch::async, f, b));
  for (auto& x : futs) x.wait();
}

struct Problem {
  int64_t NT, ml, mr, cls, Npass;  // cls 0: label on right env, 1: label on B, 2: label on left env
  double lambda, cconv;
  std::vector<double> x, y;    // [NT][2]
  std::vector<double> l, r;    // [NT][ml] or [NT][NL][ml]; [NT][mr] or [NT][NL][mr]
  std::vector<int32_t> labels;
  std::vector<double> B;       // [ml][2][2][mr] (cls 0,2) or [ml][2][2][mr][NL] (cls 1)
};

static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 4) {
    printf("Usage: %s problem.bin result.bin nthread [reps]\n", argv[0]);
    return 0;
  }
  const int nthread = std::max(1, atoi(argv[3]));
  const int reps = argc > 4 ? atoi(argv[4]) : 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", argv[1]);
    return 1;
  }
  Problem P;
  int64_t hdr[5];
  double dh[2];
  if (!rd(f, hdr, sizeof(hdr)) || !rd(f, dh, sizeof(dh))) return 1;
  P.NT = hdr[0], P.ml = hdr[1], P.mr = hdr[2], P.cls = hdr[3], P.Npass = hdr[4];
  P.lambda = dh[0], P.cconv = dh[1];
  const size_t NT = P.NT, ml = P.ml, mr = P.mr;
  const size_t nll = (P.cls == 2) ? NL : 1, nlr = (P.cls == 0) ? NL : 1, nlb = (P.cls == 1) ? NL : 1;
  P.x.resize(NT * 2), P.y.resize(NT * 2), P.l.resize(NT * nll * ml), P.r.resize(NT * nlr * mr);
  P.labels.resize(NT), P.B.resize(ml * 4 * mr * nlb);
  if (!rd(f, P.x.data(), P.x.size() * 8) || !rd(f, P.y.data(), P.y.size() * 8) ||
      !rd(f, P.l.data(), P.l.size() * 8) || !rd(f, P.r.data(), P.r.size() * 8) ||
      !rd(f, P.labels.data(), NT * 4) || !rd(f, P.B.data(), P.B.size() * 8)) {
    fprintf(stderr, "short problem file\n");
    return 1;
  }
  fclose(f);

  // dense t.v, index order (alpha, s, t, beta, L) like B (+L): vsz per image
  const size_t bsz = ml * 4 * mr;           // B-shaped part without label
  const size_t vsz = bsz * ((P.cls == 1) ? 1 : NL);
  const size_t Bsz = P.B.size();
  const auto bounds = make_bounds(nthread, NT);
  std::vector<double> v;
  double t_setbond = 1e300, t_cgrad = 1e300, t_quad = 1e300;  // best of `reps` (kind to the CPU)
  std::vector<double> B, costs;
  double Cfinal = 0;
  long ncor = 0;

  for (int rep = 0; rep < reps; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    // ---- setBond: fixedL.cc:183-185 -------------------------------------------------
    v.assign(NT * vsz, 0.0);
    parallel_do(bounds, [&](Bound b) {
      for (size_t n = b.begin; n < b.end; ++n) {
        double* vn = &v[n * vsz];
        const double* xn = &P.x[n * 2];
        const double* yn = &P.y[n * 2];
        for (size_t a = 0; a < ml; ++a)
          for (int s = 0; s < 2; ++s)
            for (int t = 0; t < 2; ++t)
              for (size_t be = 0; be < mr; ++be) {
                const size_t i = ((a * 2 + s) * 2 + t) * mr + be;
                const double xy = xn[s] * yn[t];
                if (P.cls == 1) {
                  vn[i] = xy * P.l[n * ml + a] * P.r[n * mr + be];
                } else if (P.cls == 0) {
                  for (int L = 0; L < NL; ++L) vn[i * NL + L] = xy * P.l[n * ml + a] * P.r[(n * NL + L) * mr + be];
                } else {
                  for (int L = 0; L < NL; ++L) vn[i * NL + L] = xy * P.l[(n * NL + L) * ml + a] * P.r[n * mr + be];
                }
              }
      }
    });
    auto t1 = std::chrono::steady_clock::now();

    // P_n = B * t.v  (fixedL.cc:318,377,399,416)
    auto project = [&](const std::vector<double>& T, size_t n, double* Pn) {
      const double* vn = &v[n * vsz];
      for (int L = 0; L < NL; ++L) Pn[L] = 0.0;
      if (P.cls == 1) {
        for (size_t i = 0; i < bsz; ++i) {
          const double vi = vn[i];
          const double* Ti = &T[i * NL];
          for (int L = 0; L < NL; ++L) Pn[L] += Ti[L] * vi;
        }
      } else {
        for (size_t i = 0; i < bsz; ++i) {
          const double Ti = T[i];
          const double* vi = &vn[i * NL];
          for (int L = 0; L < NL; ++L) Pn[L] += Ti * vi[L];
        }
      }
    };
    // tensors[nt] += dP * dag(t.v)  (fixedL.cc:379,418)
    auto accumulate = [&](std::vector<double>& G, size_t n, const double* dP) {
      const double* vn = &v[n * vsz];
      if (P.cls == 1) {
        for (size_t i = 0; i < bsz; ++i) {
          const double vi = vn[i];
          double* Gi = &G[i * NL];
          for (int L = 0; L < NL; ++L) Gi[L] += dP[L] * vi;
        }
      } else {
        for (size_t i = 0; i < bsz; ++i) {
          const double* vi = &vn[i * NL];
          double g = 0.0;
          for (int L = 0; L < NL; ++L) g += dP[L] * vi[L];
          G[i] += g;
        }
      }
    };
    auto norm2 = [&](const std::vector<double>& T) {
      double s = 0;
      for (double z : T) s += z * z;
      return s;
    };
    // gradient at T: returns sum over thread partials in thread order; C = sum |dP|^2
    auto gradient = [&](const std::vector<double>& T, std::vector<double>& G, double& C) {
      std::vector<std::vector<double>> tensors(nthread, std::vector<double>(Bsz, 0.0));
      std::vector<double> reals(nthread, 0.0);
      parallel_do(bounds, [&](Bound b) {
        double Pn[NL], dP[NL];
        for (size_t n = b.begin; n < b.end; ++n) {
          project(T, n, Pn);
          double e = 0;
          for (int L = 0; L < NL; ++L) {
            dP[L] = ((L == P.labels[n]) ? 1.0 : 0.0) - Pn[L];
            e += dP[L] * dP[L];
          }
          accumulate(tensors[b.n], n, dP);
          reals[b.n] += e;
        }
      });
      G.assign(Bsz, 0.0);
      C = 0;
      for (int nt = 0; nt < nthread; ++nt) {  // stdx::accumulate in thread order
        for (size_t i = 0; i < Bsz; ++i) G[i] += tensors[nt][i];
        C += reals[nt];
      }
    };

    // ---- cgrad: fixedL.cc:349-445 ----------------------------------------------------
    B = P.B;
    costs.clear();
    std::vector<double> r, p, nr;
    double C;
    gradient(B, r, C);
    if (P.lambda != 0.0)
      for (size_t i = 0; i < Bsz; ++i) r[i] -= P.lambda * B[i];
    p = r;
    for (int pass = 1; pass <= P.Npass; ++pass) {
      std::vector<double> reals(nthread, 0.0);
      parallel_do(bounds, [&](Bound b) {
        double Pn[NL];
        for (size_t n = b.begin; n < b.end; ++n) {
          project(p, n, Pn);
          for (int L = 0; L < NL; ++L) reals[b.n] += Pn[L] * Pn[L];
        }
      });
      double pAp = 0;
      for (double z : reals) pAp += z;
      pAp += P.lambda * norm2(p);
      const double a = norm2(r) / pAp;
      for (size_t i = 0; i < Bsz; ++i) B[i] += a * p[i];
      if (pass == P.Npass) break;
      gradient(B, nr, C);
      if (P.lambda != 0.0)
        for (size_t i = 0; i < Bsz; ++i) nr[i] -= P.lambda * B[i];
      const double beta = norm2(nr) / norm2(r);
      r = nr;
      C += P.lambda * norm2(B);
      costs.push_back(C / (double)NT);
      if (std::sqrt(norm2(r)) < P.cconv) break;
      for (size_t i = 0; i < Bsz; ++i) p[i] = r[i] + beta * p[i];
    }
    auto t2 = std::chrono::steady_clock::now();

    // ---- quadcost: fixedL.cc:280-344 -------------------------------------------------
    {
      std::vector<std::array<double, NL>> reals(nthread);
      std::vector<long> ints(nthread, 0);
      for (auto& a : reals) a.fill(0.0);
      parallel_do(bounds, [&](Bound b) {
        double Pn[NL];
        for (size_t n = b.begin; n < b.end; ++n) {
          project(B, n, Pn);
          double e = 0;
          int am = 0;
          double mx = std::fabs(Pn[0]);
          for (int L = 0; L < NL; ++L) {
            double d = ((L == P.labels[n]) ? 1.0 : 0.0) - Pn[L];
            e += d * d;
            if (std::fabs(Pn[L]) > mx) {
              mx = std::fabs(Pn[L]);
              am = L;
            }
          }
          reals[b.n][P.labels[n]] += e;
          if (am == P.labels[n]) ints[b.n] += 1;
        }
      });
      Cfinal = P.lambda * norm2(B);
      for (int L = 0; L < NL; ++L)
        for (int nt = 0; nt < nthread; ++nt) Cfinal += reals[nt][L];
      ncor = 0;
      for (long z : ints) ncor += z;
    }
    auto t3 = std::chrono::steady_clock::now();
    t_setbond = std::min(t_setbond, std::chrono::duration<double>(t1 - t0).count());
    t_cgrad = std::min(t_cgrad, std::chrono::duration<double>(t2 - t1).count());
    t_quad = std::min(t_quad, std::chrono::duration<double>(t3 - t2).count());
  }

  FILE* g = fopen(argv[2], "wb");
  if (!g) return 1;
  int64_t nc = costs.size();
  double tt[3] = {t_setbond, t_cgrad, t_quad};
  int64_t nco = ncor;
  fwrite(&nc, 8, 1, g);
  fwrite(costs.data(), 8, costs.size(), g);
  fwrite(&Cfinal, 8, 1, g);
  fwrite(&nco, 8, 1, g);
  fwrite(tt, 8, 3, g);
  fwrite(B.data(), 8, B.size(), g);
  fclose(g);
  printf("NT=%zu ml=%zu mr=%zu cls=%ld nthread=%d  setBond %.4f s  cgrad %.4f s  quadcost %.4f s  total %.4f s\n", NT,
         ml, mr, (long)P.cls, nthread, tt[0], tt[1], tt[2], tt[0] + tt[1] + tt[2]);
  return 0;
}
