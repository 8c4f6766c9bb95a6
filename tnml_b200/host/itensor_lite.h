// itensor_lite.h -- the ITensor-shaped surface that fixedL.cc / util.h touch
// (SURVEY.md 8b), as a small self-contained header.  NOT ITensor: only the
// members the reference program uses, dense real storage, value semantics.
// Heavy arithmetic on the hot path does not go through this header -- it runs
// in libtnml_b200.so (include/tnml_b200.h); these types carry the MPS, the
// configuration and the file I/O on the host, and provide small-tensor
// contraction (`*`) for initialisation / checks.
//
// Reference uses covered: Index/IndexVal/IndexType("Label") (util.h:17,
// fixedL.cc:669,685), ITensor * + - scalar norm sqr rank setElt findtype
// commonIndex real scaleTo (fixedL.cc:57-59,289,306,318-323,385,493-498,
// 519-530), MPS N/A/Aref/Anc/setA (util.h:76-102), SiteSet (618-632),
// Sweeps (749), Args (467,473-476,751-759), InputGroup (584-608),
// readFromFile/writeToFile/fileExists (619-631,671-674,700,727,764),
// printfln/println/Error.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <iostream>
#include <string>
#include <vector>

namespace itensor {

using Real = double;

struct ITError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
[[noreturn]] inline void Error(std::string const& msg) { throw ITError(msg); }

inline std::string format(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return buf;
}
template <typename... T>
void printfln(const char* fmt, T... t) {
  printf(fmt, t...);
  printf("\n");
}
inline void printfln(const char* s) { printf("%s\n", s); }
template <typename... T>
void println(T const&... t) {
  std::ostringstream o;
  (void)std::initializer_list<int>{(o << t, 0)...};
  printf("%s\n", o.str().c_str());
}

// ---- Index ---------------------------------------------------------------
struct IndexType {
  std::string name;
  explicit IndexType(const char* n = "Link") : name(n) {}
  bool operator==(IndexType const& o) const { return name == o.name; }
};
static const IndexType Link("Link"), Site("Site");

struct IndexVal;
class Index {
  long id_ = 0;
  long m_ = 0;
  std::string name_;
  IndexType type_;
  static long next_id() {
    static long c = 0;
    return ++c;
  }

 public:
  Index() {}
  Index(std::string const& name, long m, IndexType t = Link) : id_(next_id()), m_(m), name_(name), type_(t) {}
  long m() const { return m_; }
  long id() const { return id_; }
  std::string const& name() const { return name_; }
  IndexType const& type() const { return type_; }
  explicit operator bool() const { return id_ != 0; }
  bool operator==(Index const& o) const { return id_ == o.id_; }
  bool operator!=(Index const& o) const { return id_ != o.id_; }
  IndexVal operator()(long i) const;
  // serialisation keeps ids so that a reloaded MPS still matches its SiteSet
  void write(std::ostream& s) const {
    long n = name_.size(), tn = type_.name.size();
    s.write((char*)&id_, 8).write((char*)&m_, 8).write((char*)&n, 8).write(name_.data(), n);
    s.write((char*)&tn, 8).write(type_.name.data(), tn);
  }
  void read(std::istream& s) {
    long n = 0, tn = 0;
    s.read((char*)&id_, 8).read((char*)&m_, 8).read((char*)&n, 8);
    name_.resize(n);
    s.read(&name_[0], n);
    s.read((char*)&tn, 8);
    type_.name.resize(tn);
    s.read(&type_.name[0], tn);
  }
};
struct IndexVal {
  Index index;
  long val = 0;  // 1-indexed
};
inline IndexVal Index::operator()(long i) const {
  if (i < 1 || i > m_) Error(format("IndexVal %ld out of range 1..%ld", i, m_));
  return IndexVal{*this, i};
}

// ---- ITensor -------------------------------------------------------------
class ITensor {
  std::vector<Index> is_;
  std::vector<Real> d_;  // row-major in the order of is_
  bool valid_ = false;

 public:
  ITensor() {}
  explicit ITensor(std::vector<Index> const& is) : is_(is), valid_(true) {
    long n = 1;
    for (auto& i : is_) n *= i.m();
    d_.assign(n, 0.0);
  }
  ITensor(std::vector<Index> const& is, std::vector<Real>&& data) : is_(is), d_(std::move(data)), valid_(true) {}
  template <typename... I>
  explicit ITensor(Index const& i0, I const&... rest) : ITensor(std::vector<Index>{i0, rest...}) {}
  explicit operator bool() const { return valid_; }
  std::vector<Index> const& inds() const { return is_; }
  std::vector<Real> const& data() const { return d_; }
  std::vector<Real>& data() { return d_; }
  long r() const { return (long)is_.size(); }

  long offset(std::vector<IndexVal> const& ivs) const {
    if ((long)ivs.size() != r()) Error("wrong number of IndexVals");
    long off = 0;
    for (auto& I : is_) {
      bool found = false;
      for (auto& iv : ivs)
        if (iv.index == I) {
          off = off * I.m() + (iv.val - 1);
          found = true;
          break;
        }
      if (!found) Error("IndexVal for index " + I.name() + " missing");
    }
    return off;
  }
  template <typename... IV>
  Real real(IV const&... ivs) const {
    if (!valid_) Error("real() of default ITensor");
    return d_[offset(std::vector<IndexVal>{ivs...})];
  }
  template <typename... IV>
  void set(IV const&... args) {  // set(iv1, iv2, ..., value)
    setImpl(std::vector<IndexVal>{}, args...);
  }
  void scaleTo(Real) {}  // ITensor's lazy scale factor does not exist here

  ITensor& operator*=(Real x) {
    for (auto& v : d_) v *= x;
    return *this;
  }
  ITensor& operator/=(Real x) { return operator*=(1.0 / x); }
  ITensor& operator+=(ITensor const& o) { return addAssign(o, 1.0); }
  ITensor& operator-=(ITensor const& o) { return addAssign(o, -1.0); }
  ITensor& operator*=(ITensor const& o) {
    *this = contract(*this, o);
    return *this;
  }

  static ITensor contract(ITensor const& A, ITensor const& B) {
    if (!A.valid_ || !B.valid_) Error("contraction with default ITensor");
    // classify indices
    std::vector<int> aU, aC, bU, bC;
    for (int i = 0; i < (int)A.is_.size(); ++i) {
      int j = B.find(A.is_[i]);
      if (j >= 0) {
        aC.push_back(i);
        bC.push_back(j);
      } else
        aU.push_back(i);
    }
    for (int j = 0; j < (int)B.is_.size(); ++j)
      if (A.find(B.is_[j]) < 0) bU.push_back(j);
    auto perm = [](ITensor const& T, std::vector<int> const& first, std::vector<int> const& second, long& n1,
                   long& n2) {
      std::vector<int> order = first;
      order.insert(order.end(), second.begin(), second.end());
      n1 = 1;
      for (int i : first) n1 *= T.is_[i].m();
      n2 = 1;
      for (int i : second) n2 *= T.is_[i].m();
      std::vector<long> stride(T.is_.size());
      long s = 1;
      for (int i = (int)T.is_.size() - 1; i >= 0; --i) {
        stride[i] = s;
        s *= T.is_[i].m();
      }
      std::vector<Real> out(T.d_.size());
      std::vector<long> dims, st;
      for (int i : order) {
        dims.push_back(T.is_[i].m());
        st.push_back(stride[i]);
      }
      std::vector<long> cnt(order.size(), 0);
      long src = 0;
      for (size_t k = 0; k < out.size(); ++k) {
        out[k] = T.d_[src];
        for (int q = (int)order.size() - 1; q >= 0; --q) {
          src += st[q];
          if (++cnt[q] < dims[q]) break;
          src -= st[q] * dims[q];
          cnt[q] = 0;
        }
      }
      return out;
    };
    long am, ak, bk, bn;
    auto Am = perm(A, aU, aC, am, ak);
    auto Bm = perm(B, bC, bU, bk, bn);
    std::vector<Index> ris;
    for (int i : aU) ris.push_back(A.is_[i]);
    for (int j : bU) ris.push_back(B.is_[j]);
    std::vector<Real> out(am * bn, 0.0);
    for (long i = 0; i < am; ++i)
      for (long k = 0; k < ak; ++k) {
        Real a = Am[i * ak + k];
        if (a == 0.0) continue;
        const Real* br = &Bm[k * bn];
        Real* orow = &out[i * bn];
        for (long j = 0; j < bn; ++j) orow[j] += a * br[j];
      }
    return ITensor(ris, std::move(out));
  }

 private:
  int find(Index const& I) const {
    for (int i = 0; i < (int)is_.size(); ++i)
      if (is_[i] == I) return i;
    return -1;
  }
  void setImpl(std::vector<IndexVal> ivs, Real v) { d_[offset(ivs)] = v; }
  template <typename... R>
  void setImpl(std::vector<IndexVal> ivs, IndexVal const& iv, R const&... rest) {
    ivs.push_back(iv);
    setImpl(ivs, rest...);
  }
  ITensor& addAssign(ITensor const& o, Real f) {
    if (!valid_) {  // ITensor{} + T == T  (fixedL.cc:374-385 relies on it)
      *this = o;
      for (auto& v : d_) v *= f;
      return *this;
    }
    if (!o.valid_) return *this;
    if (o.is_.size() != is_.size()) Error("adding ITensors of different rank");
    // permute o into this order
    std::vector<long> ostride(o.is_.size());
    long s = 1;
    for (int i = (int)o.is_.size() - 1; i >= 0; --i) {
      ostride[i] = s;
      s *= o.is_[i].m();
    }
    std::vector<long> st, dims;
    for (auto& I : is_) {
      int j = o.find(I);
      if (j < 0) Error("adding ITensors with different indices");
      st.push_back(ostride[j]);
      dims.push_back(I.m());
    }
    std::vector<long> cnt(is_.size(), 0);
    long src = 0;
    for (size_t k = 0; k < d_.size(); ++k) {
      d_[k] += f * o.d_[src];
      for (int q = (int)is_.size() - 1; q >= 0; --q) {
        src += st[q];
        if (++cnt[q] < dims[q]) break;
        src -= st[q] * dims[q];
        cnt[q] = 0;
      }
    }
    return *this;
  }
};
// ---- printing (ITensor's Print / PrintData / PAUSE / EXIT macros, util.h and fixedL.cc debugging aids) ----
inline std::ostream& operator<<(std::ostream& s, Index const& I) {
  return s << "(" << I.name() << "," << I.m() << "," << I.type().name << ")";
}
inline std::ostream& operator<<(std::ostream& s, IndexVal const& iv) { return s << iv.index << "=" << iv.val; }
inline Real norm(ITensor const& T);
inline std::ostream& operator<<(std::ostream& s, ITensor const& T) {
  if (!T) return s << "ITensor r=0: (default constructed)";
  s << "ITensor r=" << T.r() << ":";
  for (auto const& I : T.inds()) s << " " << I;
  return s << "  {norm=" << format("%.2f", norm(T)) << "}";
}
// PrintData: every element above 1E-10 with its index values, like ITensor's printData
inline void printData(std::ostream& s, ITensor const& T) {
  s << T << "\n";
  if (!T) return;
  std::vector<long> cnt(T.inds().size(), 1);
  for (size_t k = 0; k < T.data().size(); ++k) {
    if (std::fabs(T.data()[k]) > 1E-10) {
      s << "  (";
      for (size_t q = 0; q < cnt.size(); ++q) s << (q ? "," : "") << cnt[q];
      s << ") " << format("%.10f", T.data()[k]) << "\n";
    }
    for (int q = (int)cnt.size() - 1; q >= 0; --q) {
      if (++cnt[q] <= T.inds()[q].m()) break;
      cnt[q] = 1;
    }
  }
}
inline void pause_() {
  std::cout << "(Paused, press enter to continue)" << std::endl;
  std::cin.get();
}
#define Print(X) (std::cout << #X << " = " << (X) << std::endl)
#define PrintData(X) (std::cout << #X << " = ", itensor::printData(std::cout, (X)))
#define PAUSE itensor::pause_();
#define EXIT exit(0);

inline ITensor operator*(ITensor const& A, ITensor const& B) { return ITensor::contract(A, B); }
inline ITensor operator*(ITensor A, Real x) { return A *= x; }
inline ITensor operator*(Real x, ITensor A) { return A *= x; }
inline ITensor operator+(ITensor A, ITensor const& B) { return A += B; }
inline ITensor operator-(ITensor A, ITensor const& B) { return A -= B; }
inline Real norm(ITensor const& T) {
  Real s = 0;
  for (Real v : T.data()) s += v * v;
  return std::sqrt(s);
}
inline Real sqr(Real x) { return x * x; }
inline long rank(ITensor const& T) { return T.r(); }
inline ITensor dag(ITensor const& T) { return T; }  // real tensors
inline ITensor setElt(IndexVal const& iv) {
  ITensor T(iv.index);
  T.set(iv, 1.0);
  return T;
}
inline Index findtype(ITensor const& T, IndexType const& t) {
  for (auto& I : T.inds())
    if (I.type() == t) return I;
  return Index();
}
inline Index commonIndex(ITensor const& A, ITensor const& B, IndexType const* t = nullptr) {
  for (auto& I : A.inds())
    for (auto& J : B.inds())
      if (I == J && (!t || I.type() == *t)) return I;
  return Index();
}

// ---- SiteSet / MPS / Sweeps -----------------------------------------------
class SiteSet {
  std::vector<Index> s_;

 public:
  SiteSet() {}
  SiteSet(int N, int d) {
    s_.resize(N + 1);
    for (int j = 1; j <= N; ++j) s_[j] = Index(format("S%d", j), d, Site);
  }
  int N() const { return (int)s_.size() - 1; }
  Index const& operator()(int i) const { return s_.at(i); }
  void write(std::ostream& s) const {
    long n = N();
    s.write((char*)&n, 8);
    for (int j = 1; j <= n; ++j) s_[j].write(s);
  }
  void read(std::istream& s) {
    long n = 0;
    s.read((char*)&n, 8);
    s_.assign(n + 1, Index());
    for (int j = 1; j <= n; ++j) s_[j].read(s);
  }
};

class MPS {
  int N_ = 0;
  std::vector<ITensor> A_;

 public:
  MPS() {}
  explicit MPS(SiteSet const& sites) : N_(sites.N()), A_(sites.N() + 2) {}
  explicit MPS(int N) : N_(N), A_(N + 2) {}
  int N() const { return N_; }
  ITensor const& A(int i) const { return A_.at(i); }
  ITensor& Aref(int i) { return A_.at(i); }
  ITensor& Anc(int i) { return A_.at(i); }
  void setA(int i, ITensor const& T) { A_.at(i) = T; }
};

class Sweeps {
  int n_;
  int minm_, maxm_;
  Real cutoff_;

 public:
  Sweeps(int n, int minm, int maxm, Real cutoff) : n_(n), minm_(minm), maxm_(maxm), cutoff_(cutoff) {}
  int nsweep() const { return n_; }
  int maxm(int) const { return maxm_; }
  int minm(int) const { return minm_; }
  Real cutoff(int) const { return cutoff_; }
};
enum Direction { Fromleft = 1, Fromright = 2 };
// b=1..N-1 with ha=1, then b=N-1..1 with ha=2
inline void sweepnext(int& b, int& ha, int N) {
  const int inc = (ha == 1 ? +1 : -1);
  b += inc;
  if (b == (ha == 1 ? N : 0)) {
    b -= inc;
    ++ha;
  }
}

// ---- Args / InputGroup ------------------------------------------------------
class Args {
  std::map<std::string, std::string> kv_;

 public:
  Args() {}
  template <typename V, typename... R>
  Args(const char* k, V const& v, R const&... rest) {
    addAll(k, v, rest...);
  }
  template <typename... R>
  Args(Args const& o, R const&... rest) : kv_(o.kv_) {
    addAll(rest...);
  }
  void add(std::string const& k, Real v) { kv_[k] = format("%.17g", v); }
  void add(std::string const& k, int v) { kv_[k] = format("%d", v); }
  void add(std::string const& k, long v) { kv_[k] = format("%ld", v); }
  void add(std::string const& k, bool v) { kv_[k] = v ? "1" : "0"; }
  void add(std::string const& k, const char* v) { kv_[k] = v; }
  void add(std::string const& k, std::string const& v) { kv_[k] = v; }
  bool defined(std::string const& k) const { return kv_.count(k) > 0; }
  int getInt(std::string const& k) const { return std::stoi(get(k)); }
  int getInt(std::string const& k, int d) const { return defined(k) ? getInt(k) : d; }
  Real getReal(std::string const& k) const { return std::stod(get(k)); }
  Real getReal(std::string const& k, Real d) const { return defined(k) ? getReal(k) : d; }
  bool getBool(std::string const& k) const { return get(k) != "0"; }
  bool getBool(std::string const& k, bool d) const { return defined(k) ? getBool(k) : d; }
  std::string getString(std::string const& k) const { return get(k); }
  std::string getString(std::string const& k, std::string const& d) const { return defined(k) ? get(k) : d; }
  static Args& global() {
    static Args g;
    return g;
  }

 private:
  std::string const& get(std::string const& k) const {
    auto it = kv_.find(k);
    if (it == kv_.end()) Error("Args: missing key " + k);
    return it->second;
  }
  void addAll() {}
  template <typename V, typename... R>
  void addAll(const char* k, V const& v, R const&... rest) {
    add(k, v);
    addAll(rest...);
  }
};

// `input { key = value ... }`; unknown keys are ignored
class InputGroup {
  std::map<std::string, std::string> kv_;

 public:
  InputGroup(std::string const& file, std::string const& name) {
    std::ifstream f(file);
    if (!f) Error("Couldn't open input file " + file);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string txt = ss.str();
    auto p = txt.find(name);
    while (p != std::string::npos) {
      auto q = txt.find_first_not_of(" \t\r\n", p + name.size());
      if (q != std::string::npos && txt[q] == '{') {
        p = q;
        break;
      }
      p = txt.find(name, p + 1);
    }
    if (p == std::string::npos) Error("Couldn't find group " + name);
    auto e = txt.find('}', p);
    std::stringstream body(txt.substr(p + 1, e - p - 1));
    std::string line;
    while (std::getline(body, line)) {
      auto h = line.find('#');
      if (h != std::string::npos) line = line.substr(0, h);
      auto eq = line.find('=');
      if (eq == std::string::npos) continue;
      auto trim = [](std::string s) {
        auto a = s.find_first_not_of(" \t\r");
        auto b = s.find_last_not_of(" \t\r");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
      };
      kv_[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
  }
  bool has(std::string const& k) const { return kv_.count(k) > 0; }
  std::string getString(std::string const& k, std::string const& d) const { return has(k) ? kv_.at(k) : d; }
  int getInt(std::string const& k, int d) const { return has(k) ? (int)std::stod(kv_.at(k)) : d; }
  Real getReal(std::string const& k, Real d) const { return has(k) ? std::stod(kv_.at(k)) : d; }
  bool getYesNo(std::string const& k, bool d) const {
    if (!has(k)) return d;
    std::string v = kv_.at(k);
    std::transform(v.begin(), v.end(), v.begin(), ::tolower);
    return v == "yes" || v == "y" || v == "true" || v == "1";
  }
};

// ---- file I/O (own documented format; ITensor's binary format cannot be
//      verified offline) -------------------------------------------
inline bool fileExists(std::string const& f) { return (bool)std::ifstream(f); }

inline void write(std::ostream& s, ITensor const& T) {
  long r = T.r(), n = T.data().size();
  s.write((char*)&r, 8);
  for (auto& I : T.inds()) I.write(s);
  s.write((char*)&n, 8);
  s.write((char*)T.data().data(), n * 8);
}
inline void read(std::istream& s, ITensor& T) {
  long r = 0, n = 0;
  s.read((char*)&r, 8);
  std::vector<Index> is(r);
  for (auto& I : is) I.read(s);
  s.read((char*)&n, 8);
  std::vector<Real> d(n);
  s.read((char*)d.data(), n * 8);
  T = ITensor(is, std::move(d));
}
inline void writeToFile(std::string const& f, SiteSet const& s) {
  std::ofstream o(f, std::ios::binary);
  o.write("TNMLS1\0\0", 8);
  s.write(o);
}
inline void writeToFile(std::string const& f, MPS const& W) {
  std::ofstream o(f, std::ios::binary);
  o.write("TNMLW1\0\0", 8);
  long N = W.N();
  o.write((char*)&N, 8);
  for (int j = 1; j <= N; ++j) write(o, W.A(j));
}
template <class T>
T readFromFile(std::string const& f);
template <>
inline SiteSet readFromFile<SiteSet>(std::string const& f) {
  std::ifstream i(f, std::ios::binary);
  char magic[8];
  i.read(magic, 8);
  if (std::string(magic, 6) != "TNMLS1") Error("not a tnml_b200 sites file: " + f);
  SiteSet s;
  s.read(i);
  return s;
}
template <class T>
T readFromFile(std::string const& f, SiteSet const& sites);
template <>
inline MPS readFromFile<MPS>(std::string const& f, SiteSet const& sites) {
  std::ifstream i(f, std::ios::binary);
  char magic[8];
  i.read(magic, 8);
  if (std::string(magic, 6) != "TNMLW1") Error("not a tnml_b200 MPS file: " + f);
  long N = 0;
  i.read((char*)&N, 8);
  if (N != sites.N()) Error("MPS file has a different number of sites");
  MPS W(sites);
  for (int j = 1; j <= N; ++j) read(i, W.Aref(j));
  return W;
}

}  // namespace itensor
