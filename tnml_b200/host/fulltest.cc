// fulltest.cc -- drop-in for the reference program `fulltest <inputfile>`
// (/root/reference/fulltest.cc:7-99 + util.h:123-200 fullTest): load the MPS `fname`
// (default "W") and the `sites` file, classify the MNIST test set, print the per-digit report.
// The per-image contraction toverlap (util.h:19-40) runs on the GPU (tnml_fulltest).
// Added key: `imglen` (must match the training run; default 28 like the reference).
#include <array>
#include <cstring>

#include "../../include/tnml_b200.h"
#include "itensor_lite.h"
#include "mnist.h"

using namespace itensor;
using std::string;
using std::vector;

const auto Label = IndexType("Label");

int main(int argc, const char* argv[]) {
  if (argc != 2) {
    printfln("Usage: %s inputfile", argv[0]);
    return 0;
  }
  try {
    auto input = InputGroup(argv[1], "input");
    int d = 2;
    auto datadir = input.getString("datadir", "/Users/mstoudenmire/software/tnml/mllib/MNIST");
    auto fname = input.getString("fname", "W");
    auto feature = input.getString("feature", "series");
    auto imglen = input.getInt("imglen", 28);
    auto device = input.getInt("device", 0);

    std::array<long, 10> labels{{0, 1, 2, 3, 4, 5, 6, 7, 8, 9}};
    printf("Labels:");
    for (auto l : labels) printf(" %ld", l);
    printf("\n");

    auto test = mllib::readMNIST(datadir, mllib::Test, 50000);   // readMNIST default NT (mnist.h:453)
    if (imglen != 28) mllib::reduce(test, imglen);
    auto N = (int)test.front().size();
    SiteSet sites;
    if (fileExists("sites"))
      sites = readFromFile<SiteSet>("sites");
    else
      Error("Couldn't find file 'sites'");
    if (sites.N() != N) Error("sites file does not match the image size");

    enum Feature { Normal, Series };
    auto ftype = Series;
    if (feature == "norm" || feature == "normal")
      ftype = Normal;
    else if (feature == "series")
      ftype = Series;
    else
      Error(format("feature type \"%s\" not recognized", feature.c_str()));
    auto phi = [ftype](Real g, int n) -> Real {   // fulltest.cc:58-71
      if (g < 0 || g > 255.) Error(format("Expected g=%f to be in [0,255]", g));
      auto x = g / 255.;
      if (ftype == Normal) return n == 1 ? std::cos(M_PI / 2. * x) : std::sin(M_PI / 2. * x);
      return n == 1 ? 1. : x / 4.;
    };

    println("Converting test set to MPS");
    const long totNtest = (long)test.size();
    vector<double> feat((size_t)totNtest * N * d);
    vector<int32_t> lab(totNtest);
    for (long n = 0; n < totNtest; ++n) {
      lab[n] = test[n].label;
      for (int j = 1; j <= N; ++j)
        for (int k = 1; k <= d; ++k) feat[((size_t)n * N + (j - 1)) * d + (k - 1)] = phi(test[n](j), k);
    }
    printfln("Total of %d testing images", (int)totNtest);

    MPS psi;
    if (fileExists(fname))
      psi = readFromFile<MPS>(fname, sites);
    else
      Error(format("Couldn't find file '%s'", fname.c_str()));
    printfln("Running full test of %s", fname.c_str());

    tnml_handle h = nullptr;
    if (tnml_create(device, 0, &h) != 0) Error(string("tnml_create: ") + tnml_last_error(nullptr));
    auto ck = [&](int rc, const char* what) {
      if (rc != 0) Error(format("%s failed (%d): %s", what, rc, tnml_last_error(h)));
    };
    ck(tnml_set_images(h, totNtest, N, feat.data(), lab.data(), totNtest, 0), "tnml_set_images");
    long cent = 0;
    for (int j = 1; j <= N; ++j) {
      auto const& A = psi.A(j);
      const bool haslab = (bool)findtype(A, Label);
      if (haslab && cent == 0) cent = j;
      auto const& is = A.inds();
      ck(tnml_set_site(h, j, (int)is.at(0).m(), (int)is.at(2).m(), haslab ? 1 : 0, A.data().data()), "tnml_set_site");
    }
    if (cent == 0) Error("expected Label index at some site of psi MPS");   // util.h:140
    vector<int32_t> pred(totNtest);
    int64_t tncor = 0;
    ck(tnml_fulltest(h, pred.data(), &tncor), "tnml_fulltest");
    tnml_destroy(h);

    // report (util.h:185-199)
    std::array<long, 10> counts{}, ninc{};
    for (long n = 0; n < totNtest; ++n) {
      counts[lab[n]]++;
      if (pred[n] != lab[n]) ninc[lab[n]]++;
    }
    const long nte = totNtest, tninc = nte - tncor;
    printfln("%d/%d correct (%.2f%%), %d/%d incorrect (%.2f%%)", (int)tncor, (int)nte, tncor * 100. / nte, (int)tninc,
             (int)nte, tninc * 100. / nte);
    long tot_counts = 0;
    for (int l = 0; l < 10; ++l) {
      auto nt = counts[l];
      tot_counts += nt;
      if (nt == 0) continue;
      auto ni = ninc[l];
      auto nc = nt - ni;
      printfln("  Digit %d %d/%d correct (%.2f%%), %d/%d incorrect (%.2f%%)", l, (int)nc, (int)nt, nc * 100. / nt,
               (int)ni, (int)nt, ni * 100. / nt);
    }
    printfln("Total # test images = %d", (int)tot_counts);
  } catch (ITError const& e) {
    fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  }
  return 0;
}
