// hosttest.cc -- self-test of the host-side pieces that need no GPU: the ITensor-shaped surface
// (itensor_lite.h) and the MPS algebra of the initial-W construction (initial_w.h).
// Prints "hosttest: PASS" and returns 0, or the first failed check and returns 1.
#include <cstdio>
#include <cstdlib>
#include <sstream>

#include "initial_w.h"
#include "itensor_lite.h"

using namespace itensor;

static int fails = 0;
#define CHECK(cond, what)                                     \
  do {                                                        \
    if (!(cond)) {                                            \
      printf("hosttest: FAIL %s (%s:%d)\n", what, __FILE__, __LINE__); \
      ++fails;                                                \
    }                                                         \
  } while (0)

static double rnd(unsigned long long& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return ((s >> 11) + 0.5) / 9007199254740992.0 - 0.5;
}

static MPS random_mps(SiteSet const& sites, int m, unsigned long long seed) {
  const int N = sites.N();
  std::vector<Index> links(N + 1);
  for (int j = 0; j <= N; ++j) links[j] = Index(format("l%d", j), (j == 0 || j == N) ? 1 : m, Link);
  MPS W(sites);
  for (int j = 1; j <= N; ++j) {
    ITensor A(links[j - 1], sites(j), links[j]);
    std::vector<Real> d(links[j - 1].m() * 2 * links[j].m());
    for (auto& v : d) v = rnd(seed);
    W.setA(j, ITensor(std::vector<Index>{links[j - 1], sites(j), links[j]}, std::move(d)));
  }
  return W;
}

int main() {
  // ---- small SVD: reconstruction, orthonormality, ordering (tall, wide, rank deficient)
  unsigned long long seed = 12345;
  for (auto shape : {std::pair<long, long>{12, 5}, {5, 12}, {9, 9}}) {
    const long r = shape.first, c = shape.second, k = std::min(r, c);
    std::vector<Real> M(r * c);
    for (auto& v : M) v = rnd(seed);
    if (r == 9)   // make it rank deficient: last row = first row
      for (long j = 0; j < c; ++j) M[(r - 1) * c + j] = M[j];
    std::vector<Real> U, s, Vt;
    initw::svd_small(M, r, c, U, s, Vt);
    double err = 0, orth = 0;
    for (long i = 0; i < r; ++i)
      for (long j = 0; j < c; ++j) {
        double acc = 0;
        for (long q = 0; q < k; ++q) acc += U[i * k + q] * s[q] * Vt[q * c + j];
        err = std::max(err, std::fabs(acc - M[i * c + j]));
      }
    for (long a = 0; a < k; ++a)
      for (long b = 0; b < k; ++b) {
        if (s[a] < 1e-12 || s[b] < 1e-12) continue;
        double acc = 0;
        for (long i = 0; i < r; ++i) acc += U[i * k + a] * U[i * k + b];
        orth = std::max(orth, std::fabs(acc - (a == b ? 1.0 : 0.0)));
      }
    bool sorted = true;
    for (long q = 1; q < k; ++q) sorted = sorted && (s[q] <= s[q - 1] + 1e-15);
    CHECK(err < 1e-13, "svd_small reconstruction");
    CHECK(orth < 1e-12, "svd_small orthonormal U");
    CHECK(sorted, "svd_small descending");
    if (r == 9) CHECK(s[k - 1] < 1e-13, "svd_small rank deficiency");
  }
  // ---- ITensor-lite contraction / addition / norm
  {
    Index i("i", 3), j("j", 4), k2("k", 2);
    ITensor A(i, j), B(j, k2);
    for (long a = 1; a <= 3; ++a)
      for (long b = 1; b <= 4; ++b) A.set(i(a), j(b), a + 0.1 * b);
    for (long b = 1; b <= 4; ++b)
      for (long c = 1; c <= 2; ++c) B.set(j(b), k2(c), b - 0.5 * c);
    ITensor C = A * B;
    double ref = 0;
    for (long b = 1; b <= 4; ++b) ref += (2 + 0.1 * b) * (b - 0.5 * 2);
    CHECK(std::fabs(C.real(i(2), k2(2)) - ref) < 1e-13, "ITensor contraction");
    ITensor D = A + A;
    CHECK(std::fabs(norm(D) - 2 * norm(A)) < 1e-13, "ITensor addition / norm");
    CHECK((bool)commonIndex(A, B) && commonIndex(A, B) == j, "commonIndex");
  }
  // ---- MPS algebra: overlap is bilinear over sum; compression respects Maxm; truncation error small
  {
    SiteSet sites(10, 2);
    MPS a = random_mps(sites, 3, 1), b = random_mps(sites, 2, 2), c = random_mps(sites, 4, 3);
    const double ac = overlap(a, c), bc = overlap(b, c);
    MPS ab = sum(std::vector<MPS>{a, b}, Args("Cutoff", 1E-14));
    CHECK(std::fabs(overlap(ab, c) - (ac + bc)) < 1e-12 * (std::fabs(ac) + std::fabs(bc) + 1e-300), "overlap(sum(a,b),c)");
    const double nab = overlap(ab, ab), ref = overlap(a, a) + 2 * overlap(a, b) + overlap(b, b);
    CHECK(std::fabs(nab - ref) < 1e-12 * ref, "norm of sum");
    long mx = 0;
    for (int jj = 1; jj < 10; ++jj) mx = std::max<long>(mx, ab.A(jj).inds().at(2).m());
    CHECK(mx <= 5, "direct sum + exact compression keeps m <= m_a + m_b");
    MPS abc = sum(std::vector<MPS>{a, b, c}, Args("Cutoff", 1E-14, "Maxm", 4));
    long mx2 = 0;
    for (int jj = 1; jj < 10; ++jj) mx2 = std::max<long>(mx2, abc.A(jj).inds().at(2).m());
    CHECK(mx2 <= 4, "Maxm respected");
    // product states: sum of identical states = 2 x state, bond dimension 1 after compression
    auto img = std::vector<double>{0.1, 0.9, 0.3, 0.0, 0.5, 0.7, 0.2, 0.4, 0.6, 0.8};
    struct Img { std::vector<double> d; size_t size() const { return d.size(); } double operator()(size_t i) const { return d[i - 1]; } };
    Img im{img};
    auto phi = [](double g, int n) { return n == 1 ? 1.0 : g; };
    initw::RMPS p1 = initw::makeMPS(10, 2, im, phi);
    initw::RMPS p2 = initw::sum(std::vector<initw::RMPS>{p1, p1}, 1E-10, 10, false);
    long mx3 = 0;
    for (int jj = 1; jj < 10; ++jj) mx3 = std::max<long>(mx3, p2[jj].mr);
    CHECK(mx3 == 1, "sum of two identical product states has bond dimension 1");
    CHECK(std::fabs(initw::overlap(p2, p1) - 2 * initw::overlap(p1, p1)) < 1e-12 * initw::overlap(p1, p1), "2 x state");
  }
  // ---- sweepnext / Sweeps / Args
  {
    int b = 1, ha = 1, cnt = 0, N = 5;
    std::vector<int> seq;
    for (; ha != 3 && cnt < 100; sweepnext(b, ha, N)) {
      seq.push_back(b * 10 + ha);
      ++cnt;
    }
    CHECK(cnt == 2 * (N - 1) && seq.front() == 11 && seq[N - 2] == 41 && seq[N - 1] == 42 && seq.back() == 12, "sweepnext");
    Sweeps sw(3, 10, 40, 1E-9);
    CHECK(sw.maxm(2) == 40 && sw.minm(3) == 10 && sw.cutoff(1) == 1E-9, "Sweeps");
    Args a1("lambda", 0.5, "Npass", 4);
    Args a2{a1, "Maxm", 7};
    CHECK(a2.getInt("Npass") == 4 && a2.getInt("Maxm") == 7 && a2.getReal("lambda") == 0.5 && a2.getInt("none", 3) == 3, "Args");
  }
  // ---- svd(T, U, S, V, Args) -> Spectrum (fixedL.cc:519-523 call shape), W(c+dc) *= S, Print/PrintData
  {
    Index l("l", 3, Link), s1("s1", 2, Site), s2("s2", 2, Site), r("r", 4, Link);
    std::vector<Real> d(3 * 2 * 2 * 4);
    for (auto& v : d) v = rnd(seed);
    ITensor B(std::vector<Index>{l, s1, s2, r}, std::move(d));
    ITensor U(l, s1), S, V;                 // U shares (l, s1) with B: those are the rows
    auto spec = svd(B, U, S, V, {"Cutoff", 0.0, "Maxm", 100, "Minm", 1});
    CHECK(spec.numEigsKept() == 6 && spec.truncerr() == 0.0, "svd keeps min(6,8) values without truncation");
    ITensor R = U * S * V;
    CHECK(norm(R - B) < 1e-12 * norm(B), "U*S*V == B");
    ITensor UU = U * U;                     // contracts every index: sum of squares = number of columns
    CHECK(std::fabs(UU.data()[0] - 6.0) < 1e-12, "U has orthonormal columns");
    ITensor Vn = V;
    Vn *= S;                                // W.Aref(c+dc) *= S (fixedL.cc:521)
    CHECK(norm(U * Vn - B) < 1e-12 * norm(B), "U*(S*V) == B");
    ITensor U2(l, s1), S2, V2;
    auto spec2 = svd(B, U2, S2, V2, {"Cutoff", 0.0, "Maxm", 3, "Minm", 1});
    double tail = 0;
    for (int i = 3; i < 6; ++i) tail += spec.eigsKept()[i];
    CHECK(spec2.numEigsKept() == 3 && std::fabs(spec2.truncerr() - tail) < 1e-12, "truncerr = discarded weight");
    // rank-deficient: zero singular value kept by Minm still gives an isometry (ITensor semantics)
    ITensor Z(std::vector<Index>{l, s1, s2, r});
    Z.set(l(1), s1(1), s2(1), r(1), 2.0);
    ITensor U3(l, s1), S3, V3;
    svd(Z, U3, S3, V3, {"Cutoff", 1E-10, "Maxm", 4, "Minm", 4});
    ITensor UU3 = U3 * U3;
    CHECK(std::fabs(UU3.data()[0] - 4.0) < 1e-12, "zero singular values: U completed to an isometry");
    std::ostringstream os;
    os << l << " " << B;
    printData(os, Z);
    CHECK(os.str().find("(l,3,Link)") != std::string::npos && os.str().find("(1,1,1,1) 2.0000000000") != std::string::npos,
          "Print / PrintData formatting");
  }
  if (fails == 0) printf("hosttest: PASS\n");
  return fails ? 1 : 0;
}
