// initial_w.h -- construction of the initial weight MPS exactly the way fixedL.cc does it when no
// `W` file exists (fixedL.cc:682-728, util.h:76-121):
//   * per label: `ninitial` random training images of that label as product-state MPS (makeMPS),
//     summed and compressed with {"Cutoff",1E-10,"Maxm",10};  site c gets 0.1 * setElt(L(label))
//   * the ten label states summed with {"Cutoff",1E-8,"Maxm",10};  W.Aref(c) /= norm(W.A(c))
//   * or, when W0..W9 exist: those ten MPS tagged with setElt(L(label)) and summed, Cutoff 1E-10.
// Host-side, float64, bond dimensions <= 20: nothing here needs the GPU.
//
// ITensor v2 semantics restated (ASSUMED, unverifiable offline -- SURVEY 8c(5)):
//   sum(vector<MPS>,args): terms are added in pairs, recursively;  sum(A,B,args) = direct sum of
//   the link spaces followed by orthogonalize(args): one half sweep to the left without
//   truncation, then one half sweep to the right truncating every bond with the SVD rule
//   (sigma^2 spectrum, Maxm, Cutoff; DoRelCutoff as for the bond SVD).  The gauge differs from
//   ITensor's at most by where the orthogonality centre ends (here: site N); the state is the same.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "itensor_lite.h"

namespace initw {

using itensor::Real;

struct RSite {             // a[(l*p + s)*mr + r]
  long ml = 1, p = 2, mr = 1;
  std::vector<Real> a;
};
using RMPS = std::vector<RSite>;   // sites 1..N (element 0 unused)

// Thin SVD of the row-major rows x cols matrix M by one-sided (Hestenes) Jacobi on the smaller
// side: M = U diag(s) Vt, s descending, k = min(rows, cols).  Vectors of zero singular values are
// zero vectors (they are always truncated).
inline void svd_small(std::vector<Real> const& M, long rows, long cols, std::vector<Real>& U, std::vector<Real>& s,
                      std::vector<Real>& Vt) {
  const bool tr = rows < cols;                  // work on A = M (rows >= cols) or A = M^T
  const long nr = tr ? cols : rows, nc = tr ? rows : cols;
  std::vector<Real> A(nr * nc), V(nc * nc, 0.0);  // column-major: column j at A[j*nr ..]
  for (long i = 0; i < rows; ++i)
    for (long j = 0; j < cols; ++j) {
      const Real v = M[i * cols + j];
      if (tr) A[i * nr + j] = v; else A[j * nr + i] = v;
    }
  for (long j = 0; j < nc; ++j) V[j * nc + j] = 1.0;
  const Real tol = 4.0 * std::sqrt((Real)nr) * 1.1102230246251565e-16;
  for (int sweep = 0; sweep < 80; ++sweep) {
    bool rotated = false;
    for (long p = 0; p + 1 < nc; ++p)
      for (long q = p + 1; q < nc; ++q) {
        Real* xp = &A[p * nr];
        Real* xq = &A[q * nr];
        Real a = 0, b = 0, g = 0;
        for (long r = 0; r < nr; ++r) {
          a += xp[r] * xp[r];
          b += xq[r] * xq[r];
          g += xp[r] * xq[r];
        }
        if (!(a * b > 0.0) || g * g <= tol * tol * a * b) continue;
        rotated = true;
        const Real zeta = (b - a) / (2.0 * g);
        const Real t = std::copysign(1.0, zeta) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const Real c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        for (long r = 0; r < nr; ++r) {
          const Real u = xp[r], v = xq[r];
          xp[r] = c * u - sn * v;
          xq[r] = sn * u + c * v;
        }
        Real* vp = &V[p * nc];
        Real* vq = &V[q * nc];
        for (long r = 0; r < nc; ++r) {
          const Real u = vp[r], v = vq[r];
          vp[r] = c * u - sn * v;
          vq[r] = sn * u + c * v;
        }
      }
    if (!rotated) break;
  }
  std::vector<Real> sg(nc);
  std::vector<long> ord(nc);
  for (long j = 0; j < nc; ++j) {
    Real n2 = 0;
    for (long r = 0; r < nr; ++r) n2 += A[j * nr + r] * A[j * nr + r];
    sg[j] = std::sqrt(n2);
    ord[j] = j;
  }
  std::stable_sort(ord.begin(), ord.end(), [&](long x, long y) { return sg[x] > sg[y]; });
  const long k = nc;
  s.assign(k, 0.0);
  U.assign(rows * k, 0.0);
  Vt.assign(k * cols, 0.0);
  for (long i = 0; i < k; ++i) {
    const long j = ord[i];
    s[i] = sg[j];
    const Real inv = sg[j] > 0.0 ? 1.0 / sg[j] : 0.0;
    // A[:,j]/sigma is the singular vector on the long side, V[:,j] on the short side
    if (!tr) {
      for (long r = 0; r < rows; ++r) U[r * k + i] = A[j * nr + r] * inv;
      for (long r = 0; r < cols; ++r) Vt[i * cols + r] = (sg[j] > 0.0) ? V[j * nc + r] : 0.0;
    } else {
      for (long r = 0; r < cols; ++r) Vt[i * cols + r] = A[j * nr + r] * inv;
      for (long r = 0; r < rows; ++r) U[r * k + i] = (sg[j] > 0.0) ? V[j * nc + r] : 0.0;
    }
  }
}

// ASSUMED ITensor v2 truncation; same rule as the device SVD and the oracle.
inline long truncate_spectrum(std::vector<Real> const& P, long maxm, long minm, Real cutoff, bool do_rel,
                              Real* truncerr = nullptr) {
  long m = (long)P.size();
  Real terr = 0;
  while (m > maxm) terr += P[--m];
  Real scale = 1.0;
  if (do_rel) {
    Real sum = 0;
    for (Real v : P) sum += v;
    scale = (sum == 0.0) ? 1.0 : sum;
  }
  while (m > minm && m > 1 && terr + P[m - 1] < cutoff * scale) terr += P[--m];
  if (truncerr) *truncerr = terr / scale;
  return m;
}

// util.h:76-102 makeMPS: bond dimension 1, A_j[s] = phi(img(j), s)
template <class Img, class Phi>
RMPS makeMPS(int N, int d, Img const& img, Phi const& phi) {
  if ((size_t)N != img.size()) itensor::Error("Mismatched sizes");
  RMPS psi(N + 1);
  for (int j = 1; j <= N; ++j) {
    psi[j].ml = psi[j].mr = 1;
    psi[j].p = d;
    psi[j].a.resize(d);
    for (int n = 1; n <= d; ++n) psi[j].a[n - 1] = phi(img(j), n);
  }
  return psi;
}

inline RMPS direct_sum(RMPS const& A, RMPS const& B) {
  const int N = (int)A.size() - 1;
  if ((int)B.size() - 1 != N) itensor::Error("sum of MPS of different length");
  RMPS C(N + 1);
  for (int j = 1; j <= N; ++j) {
    RSite const &x = A[j], &y = B[j];
    if (x.p != y.p) itensor::Error("sum of MPS with different site dimensions");
    RSite& z = C[j];
    z.p = x.p;
    z.ml = (j == 1) ? 1 : x.ml + y.ml;
    z.mr = (j == N) ? 1 : x.mr + y.mr;
    z.a.assign(z.ml * z.p * z.mr, 0.0);
    const long lo = (j == 1) ? 0 : x.ml, ro = (j == N) ? 0 : x.mr;
    for (long l = 0; l < x.ml; ++l)
      for (long s = 0; s < x.p; ++s)
        for (long r = 0; r < x.mr; ++r) z.a[(l * z.p + s) * z.mr + r] += x.a[(l * x.p + s) * x.mr + r];
    for (long l = 0; l < y.ml; ++l)
      for (long s = 0; s < y.p; ++s)
        for (long r = 0; r < y.mr; ++r) z.a[((lo + l) * z.p + s) * z.mr + ro + r] += y.a[(l * y.p + s) * y.mr + r];
  }
  return C;
}

// orthogonalize(args): left half sweep without truncation (only exactly vanishing directions are
// dropped), then right half sweep truncating with {Cutoff, Maxm}.
inline void orthogonalize(RMPS& W, Real cutoff, long maxm, bool do_rel) {
  const int N = (int)W.size() - 1;
  std::vector<Real> U, s, Vt;
  for (int j = N; j >= 2; --j) {
    RSite& A = W[j];
    const long rows = A.ml, cols = A.p * A.mr;
    svd_small(A.a, rows, cols, U, s, Vt);
    long k = 0;
    while (k < (long)s.size() && s[k] > 1e-14 * s[0]) ++k;
    if (k == 0) k = 1;
    const long kk = (long)s.size();
    std::vector<Real> na(k * cols);
    for (long i = 0; i < k; ++i)
      for (long c = 0; c < cols; ++c) na[i * cols + c] = Vt[i * cols + c];
    // A_{j-1}[.., r'] = sum_r A_{j-1}[.., r] U[r][r'] s[r']
    RSite& L = W[j - 1];
    std::vector<Real> nl(L.ml * L.p * k, 0.0);
    for (long q = 0; q < L.ml * L.p; ++q)
      for (long r = 0; r < rows; ++r) {
        const Real v = L.a[q * rows + r];
        if (v == 0.0) continue;
        for (long i = 0; i < k; ++i) nl[q * k + i] += v * U[r * kk + i] * s[i];
      }
    A.a.swap(na);
    A.ml = k;
    L.a.swap(nl);
    L.mr = k;
  }
  for (int j = 1; j <= N - 1; ++j) {
    RSite& A = W[j];
    const long rows = A.ml * A.p, cols = A.mr;
    svd_small(A.a, rows, cols, U, s, Vt);
    std::vector<Real> P(s.size());
    for (size_t i = 0; i < s.size(); ++i) P[i] = s[i] * s[i];
    const long m = truncate_spectrum(P, maxm, 1, cutoff, do_rel);
    const long kk = (long)s.size();
    std::vector<Real> na(rows * m);
    for (long q = 0; q < rows; ++q)
      for (long i = 0; i < m; ++i) na[q * m + i] = U[q * kk + i];
    // A_{j+1}[i, ..] = sum_r s[i] Vt[i][r] A_{j+1}[r, ..]
    RSite& R = W[j + 1];
    const long rc = R.p * R.mr;
    std::vector<Real> nr(m * rc, 0.0);
    for (long i = 0; i < m; ++i)
      for (long r = 0; r < cols; ++r) {
        const Real v = s[i] * Vt[i * cols + r];
        if (v == 0.0) continue;
        for (long c = 0; c < rc; ++c) nr[i * rc + c] += v * R.a[r * rc + c];
      }
    A.a.swap(na);
    A.mr = m;
    R.a.swap(nr);
    R.ml = m;
  }
}

inline RMPS sum2(RMPS const& A, RMPS const& B, Real cutoff, long maxm, bool do_rel) {
  RMPS C = direct_sum(A, B);
  orthogonalize(C, cutoff, maxm, do_rel);
  return C;
}

// sum(vector<MPS>,args): add all MPS in pairs, recursively
inline RMPS sum(std::vector<RMPS> const& terms, Real cutoff, long maxm, bool do_rel) {
  const size_t Nt = terms.size();
  if (Nt == 0) itensor::Error("sum of zero MPS");
  if (Nt == 1) return terms[0];
  if (Nt == 2) return sum2(terms[0], terms[1], cutoff, maxm, do_rel);
  std::vector<RMPS> nt;
  for (size_t n = 0; n + 1 < Nt; n += 2) nt.push_back(sum2(terms[n], terms[n + 1], cutoff, maxm, do_rel));
  if (Nt % 2 == 1) nt.push_back(terms.back());
  return sum(nt, cutoff, maxm, do_rel);
}

// <A|B>: contraction over every site index (site c of a label-tagged MPS has p = d*NL)
inline Real overlap(RMPS const& A, RMPS const& B) {
  const int N = (int)A.size() - 1;
  std::vector<Real> E{1.0};   // [la][lb]
  long ea = 1, eb = 1;
  for (int j = 1; j <= N; ++j) {
    RSite const &x = A[j], &y = B[j];
    std::vector<Real> T(ea * y.p * y.mr, 0.0);  // T[la][s][rb] = sum_lb E[la][lb] y[lb][s][rb]
    for (long la = 0; la < ea; ++la)
      for (long lb = 0; lb < eb; ++lb) {
        const Real e = E[la * eb + lb];
        if (e == 0.0) continue;
        for (long q = 0; q < y.p * y.mr; ++q) T[la * y.p * y.mr + q] += e * y.a[lb * y.p * y.mr + q];
      }
    std::vector<Real> F(x.mr * y.mr, 0.0);      // F[ra][rb] = sum_{la,s} x[la][s][ra] T[la][s][rb]
    for (long la = 0; la < ea; ++la)
      for (long s = 0; s < x.p; ++s)
        for (long ra = 0; ra < x.mr; ++ra) {
          const Real v = x.a[(la * x.p + s) * x.mr + ra];
          if (v == 0.0) continue;
          for (long rb = 0; rb < y.mr; ++rb) F[ra * y.mr + rb] += v * T[(la * y.p + s) * y.mr + rb];
        }
    E.swap(F);
    ea = x.mr;
    eb = y.mr;
  }
  return E[0];
}

// in.Aref(c) *= f * setElt(L(1+label)) : the site index of site c becomes (s, l), p = d*NL
inline void tag_label(RMPS& W, int c, int label, int NL, Real f) {
  RSite& A = W[c];
  std::vector<Real> na(A.ml * A.p * NL * A.mr, 0.0);
  for (long l = 0; l < A.ml; ++l)
    for (long s = 0; s < A.p; ++s)
      for (long r = 0; r < A.mr; ++r) na[(l * A.p * NL + s * NL + label) * A.mr + r] = f * A.a[(l * A.p + s) * A.mr + r];
  A.a.swap(na);
  A.p *= NL;
}

// RMPS (site c with p = d*NL) -> ITensor MPS with index order (left, site, right [, L])
inline itensor::MPS to_mps(RMPS const& R, itensor::SiteSet const& sites, std::vector<itensor::Index>& links,
                           itensor::Index const& L, int c) {
  using namespace itensor;
  const int N = sites.N();
  const long d = sites(1).m(), NL = L.m();
  links.assign(N + 1, Index());
  links[0] = Index("l0", 1, Link);
  for (int j = 1; j <= N; ++j) links[j] = Index(format("l%d", j), R[j].mr, Link);
  MPS W(sites);
  for (int j = 1; j <= N; ++j) {
    RSite const& A = R[j];
    std::vector<Index> is{links[j - 1], sites(j), links[j]};
    if (j != c) {
      if (A.p != d) Error("unexpected site dimension");
      W.setA(j, ITensor(is, std::vector<Real>(A.a)));
    } else {
      if (A.p != d * NL) Error(format("Label Index not on site %d", c));
      is.push_back(L);
      std::vector<Real> t(A.a.size());
      for (long l = 0; l < A.ml; ++l)
        for (long s = 0; s < d; ++s)
          for (long q = 0; q < NL; ++q)
            for (long r = 0; r < A.mr; ++r) t[((l * d + s) * A.mr + r) * NL + q] = A.a[(l * A.p + s * NL + q) * A.mr + r];
      W.setA(j, ITensor(is, std::move(t)));
    }
  }
  return W;
}

// ITensor MPS without label (index order left, site, right) -> RMPS
inline RMPS from_mps(itensor::MPS const& W) {
  const int N = W.N();
  RMPS R(N + 1);
  for (int j = 1; j <= N; ++j) {
    auto const& is = W.A(j).inds();
    if (is.size() != 3) itensor::Error("expected a label-free MPS (W0..W9)");
    R[j].ml = is[0].m();
    R[j].p = is[1].m();
    R[j].mr = is[2].m();
    R[j].a = W.A(j).data();
  }
  return R;
}

// util.h:104-121 randImg with an explicit generator (the reference draws from ITensor's
// time-seeded Global::random()): uniform index, retry until the label matches.
template <class ImgVec, class Uniform>
long randImg(ImgVec const& imgs, long label, Uniform&& uniform) {
  const int max_tries = 1000;
  for (int t = 0; t < max_tries; ++t) {
    long w = (long)((Real)imgs.size() * uniform());
    if (w < 0) w = 0;
    if (w >= (long)imgs.size()) w = (long)imgs.size() - 1;
    if (imgs[w].label == label) return w;
  }
  itensor::Error(itensor::format("Did not find image with requested label after %d tries", max_tries));
  return 0;
}

// fixedL.cc:702-728.  `picks_out` (optional) receives the chosen image positions per label.
template <class ImgVec, class Phi, class Uniform>
RMPS initial_sum(int N, int d, int c, int NL, ImgVec const& train, Phi const& phi, int ninitial, Uniform&& uniform,
                 bool do_rel, std::vector<std::vector<long>>* picks_out = nullptr) {
  std::vector<RMPS> ipsis;
  for (int n = 0; n < NL; ++n) {
    std::vector<RMPS> psis;
    std::vector<long> picks;
    for (int m = 0; m < ninitial; ++m) {
      const long w = randImg(train, n, uniform);
      picks.push_back(w);
      psis.push_back(makeMPS(N, d, train[w], phi));
    }
    itensor::printfln("Summing %d random label %d states", ninitial, n);
    RMPS s = sum(psis, 1E-10, 10, do_rel);
    tag_label(s, c, n, NL, 0.1);
    ipsis.push_back(std::move(s));
    if (picks_out) picks_out->push_back(picks);
  }
  itensor::printfln("Summing all %d label states together", (int)ipsis.size());
  RMPS W = sum(ipsis, 1E-8, 10, do_rel);
  Real nrm = 0;
  for (Real v : W[c].a) nrm += v * v;
  nrm = std::sqrt(nrm);
  for (Real& v : W[c].a) v /= nrm;   // W.Aref(c) /= norm(W.A(c))
  return W;
}

// any MPS (label-free, or with the label index on one site) -> RMPS; the label index is merged
// into that site's physical index ((s, l), p = d*NL) so that contractions run over it too
inline RMPS to_raw(itensor::MPS const& W) {
  const int N = W.N();
  RMPS R(N + 1);
  for (int j = 1; j <= N; ++j) {
    auto const& is = W.A(j).inds();
    if (is.size() == 3) {
      R[j].ml = is[0].m();
      R[j].p = is[1].m();
      R[j].mr = is[2].m();
      R[j].a = W.A(j).data();
    } else if (is.size() == 4) {   // (left, site, right, L), row-major
      const long ml = is[0].m(), d = is[1].m(), mr = is[2].m(), nl = is[3].m();
      R[j].ml = ml;
      R[j].p = d * nl;
      R[j].mr = mr;
      R[j].a.resize(ml * d * nl * mr);
      auto const& t = W.A(j).data();
      for (long l = 0; l < ml; ++l)
        for (long s = 0; s < d; ++s)
          for (long r = 0; r < mr; ++r)
            for (long q = 0; q < nl; ++q) R[j].a[(l * d * nl + s * nl + q) * mr + r] = t[((l * d + s) * mr + r) * nl + q];
    } else {
      itensor::Error("MPS site tensor of unexpected rank");
    }
  }
  return R;
}

}  // namespace initw

// ---- the ITensor-shaped entry points fixedL.cc uses on whole MPS (fixedL.cc:697,710,723,729) ----
namespace itensor {

// svd(T, U, S, V, {"Cutoff", c, "Maxm", m, "Minm", n [, "DoRelCutoff", b]}) -> Spectrum   (fixedL.cc:519-523)
// ITensor v2 semantics as assumed in SURVEY 8c(1-2): the indices T shares with the incoming U are the
// rows; U gets orthonormal columns over a new Link index, S is diagonal >= 0 descending, T ~ U*S*V;
// truncation by truncate_spectrum on sigma^2; Spectrum::truncerr() = discarded weight.  This host
// version (one-sided Jacobi, float64) serves the small tensors of the host-side code (initial W,
// tests); the bond update of `fixedL` itself goes through tnml_svd_split on the device.
class Spectrum {
  Real truncerr_ = 0;
  std::vector<Real> eigs_;

 public:
  Spectrum() {}
  Spectrum(std::vector<Real> const& eigs, Real terr) : truncerr_(terr), eigs_(eigs) {}
  Real truncerr() const { return truncerr_; }
  std::vector<Real> const& eigsKept() const { return eigs_; }
  int numEigsKept() const { return (int)eigs_.size(); }
};

inline Spectrum svd(ITensor const& T, ITensor& U, ITensor& S, ITensor& V, Args const& args = Args()) {
  if (!T) Error("svd of default ITensor");
  std::vector<Index> rowI, colI;
  for (auto const& I : T.inds()) {
    bool inU = false;
    if (U)
      for (auto const& J : U.inds()) inU = inU || (J == I);
    (inU ? rowI : colI).push_back(I);
  }
  if (rowI.empty() || colI.empty()) Error("svd: U must share some but not all indices with T");
  long nr = 1, nc = 1;
  for (auto const& I : rowI) nr *= I.m();
  for (auto const& I : colI) nc *= I.m();
  // M[(rows),(cols)] = T permuted: contract with nothing, just reorder through an identity trick:
  // walk all elements of T and scatter
  std::vector<Index> const& ti = T.inds();
  std::vector<long> stride_out(ti.size());
  {
    std::vector<Index> order = rowI;
    order.insert(order.end(), colI.begin(), colI.end());
    std::vector<long> ostr(order.size());
    long sacc = 1;
    for (int q = (int)order.size() - 1; q >= 0; --q) {
      ostr[q] = sacc;
      sacc *= order[q].m();
    }
    for (size_t i = 0; i < ti.size(); ++i)
      for (size_t q = 0; q < order.size(); ++q)
        if (order[q] == ti[i]) stride_out[i] = ostr[q];
  }
  std::vector<Real> M(nr * nc);
  {
    std::vector<long> cnt(ti.size(), 0);
    long dst = 0;
    for (size_t k = 0; k < T.data().size(); ++k) {
      M[dst] = T.data()[k];
      for (int q = (int)ti.size() - 1; q >= 0; --q) {
        dst += stride_out[q];
        if (++cnt[q] < ti[q].m()) break;
        dst -= stride_out[q] * ti[q].m();
        cnt[q] = 0;
      }
    }
  }
  std::vector<Real> Um, sv, Vt;
  initw::svd_small(M, nr, nc, Um, sv, Vt);
  const long k = (long)sv.size();
  std::vector<Real> P(k);
  for (long i = 0; i < k; ++i) P[i] = sv[i] * sv[i];
  Real terr = 0;
  const long m = initw::truncate_spectrum(P, args.getInt("Maxm", 1000000), args.getInt("Minm", 1),
                                          args.getReal("Cutoff", 0.0), args.getBool("DoRelCutoff", false), &terr);
  // ITensor returns orthonormal U columns also for zero singular values: complete them (Gram-Schmidt
  // of unit vectors against the columns kept so far)
  for (long c = 0; c < m; ++c) {
    Real nn = 0;
    for (long r = 0; r < nr; ++r) nn += Um[r * k + c] * Um[r * k + c];
    if (nn > 0.25) continue;
    for (long e = 0; e < nr; ++e) {
      std::vector<Real> v(nr, 0.0);
      v[e] = 1.0;
      for (int pass = 0; pass < 2; ++pass)
        for (long c2 = 0; c2 < m; ++c2) {
          if (c2 == c) continue;
          Real d = 0;
          for (long r = 0; r < nr; ++r) d += Um[r * k + c2] * v[r];
          for (long r = 0; r < nr; ++r) v[r] -= d * Um[r * k + c2];
        }
      Real vn = 0;
      for (Real x : v) vn += x * x;
      if (vn > 0.25) {
        vn = std::sqrt(vn);
        for (long r = 0; r < nr; ++r) Um[r * k + c] = v[r] / vn;
        break;
      }
    }
  }
  Index ul("ul", m, Link), vl("vl", m, Link);
  std::vector<Index> ui = rowI, vi{vl};
  ui.push_back(ul);
  vi.insert(vi.end(), colI.begin(), colI.end());
  std::vector<Real> ud(nr * m), sd(m * m, 0.0), vd(m * nc);
  for (long r = 0; r < nr; ++r)
    for (long c = 0; c < m; ++c) ud[r * m + c] = Um[r * k + c];
  for (long c = 0; c < m; ++c) sd[c * m + c] = sv[c];
  for (long c = 0; c < m; ++c)
    for (long j = 0; j < nc; ++j) vd[c * nc + j] = Vt[c * nc + j];
  U = ITensor(ui, std::move(ud));
  S = ITensor(std::vector<Index>{ul, vl}, std::move(sd));
  V = ITensor(vi, std::move(vd));
  return Spectrum(std::vector<Real>(P.begin(), P.begin() + m), terr);
}

// overlap(psi, phi) = <psi|phi>, every site (and label) index contracted
inline Real overlap(MPS const& A, MPS const& B) { return initw::overlap(initw::to_raw(A), initw::to_raw(B)); }

// sum(vector<MPS>, {"Cutoff", c, "Maxm", m}): label-free terms over the same site indices
inline MPS sum(std::vector<MPS> const& terms, Args const& args = Args()) {
  if (terms.empty()) Error("sum of zero MPS");
  std::vector<initw::RMPS> raw;
  for (auto const& t : terms) raw.push_back(initw::from_mps(t));
  const Real cutoff = args.getReal("Cutoff", 1E-13);
  const long maxm = args.getInt("Maxm", 1000000);
  const bool dorel = args.getBool("DoRelCutoff", false);
  initw::RMPS R = initw::sum(raw, cutoff, maxm, dorel);
  const int N = terms[0].N();
  MPS W(N);
  std::vector<Index> links(N + 1);
  links[0] = Index("l0", 1, Link);
  for (int j = 1; j <= N; ++j) links[j] = Index(format("l%d", j), R[j].mr, Link);
  for (int j = 1; j <= N; ++j) {
    Index site = terms[0].A(j).inds().at(1);
    W.setA(j, ITensor(std::vector<Index>{links[j - 1], site, links[j]}, std::vector<Real>(R[j].a)));
  }
  return W;
}

}  // namespace itensor
