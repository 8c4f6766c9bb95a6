// mnist.h -- idx reader with readMNIST's selection semantics (mllib/mnist.h:443-530),
// the Data record (mllib/data.h:11-58) and the 2x2 block mean `reduce`
// (image.h:316-346; fixedL.cc never downsizes, `imglen` is our add-on).
#pragma once
#include <array>
#include <cstdint>
#include <fstream>
#include <string>
#include <vector>

#include "itensor_lite.h"

namespace mllib {

using itensor::Real;

enum DataType { Train, Test };

struct MNISTData {            // Data<Real,10>: n, label, data; operator() is 1-indexed
  static const int NL = 10;
  long n = -1;
  int label = -1;
  std::vector<Real> data;
  size_t size() const { return data.size(); }
  Real operator()(size_t i) const { return data.at(i - 1); }
};

inline uint32_t be32(std::ifstream& f) {
  unsigned char b[4];
  f.read((char*)b, 4);
  return (uint32_t(b[0]) << 24) | (uint32_t(b[1]) << 16) | (uint32_t(b[2]) << 8) | uint32_t(b[3]);
}

inline std::vector<MNISTData> readMNIST(std::string const& datadir, DataType type, long NT) {
  const std::string pre = (type == Train) ? "train" : "t10k";
  std::ifstream fi(datadir + "/" + pre + "-images-idx3-ubyte", std::ios::binary);
  std::ifstream fl(datadir + "/" + pre + "-labels-idx1-ubyte", std::ios::binary);
  if (!fi || !fl) itensor::Error("Error opening MNIST files in " + datadir);
  if (be32(fi) != 0x803) itensor::Error("Invalid magic number in image file");
  const uint32_t count = be32(fi), rows = be32(fi), cols = be32(fi);
  if (be32(fl) != 0x801) itensor::Error("Invalid magic number in label file");
  if (be32(fl) != count) itensor::Error("image / label count mismatch");
  std::vector<uint8_t> img((size_t)count * rows * cols), lab(count);
  fi.read((char*)img.data(), img.size());
  fl.read((char*)lab.data(), lab.size());
  std::array<long, 10> counts{};
  std::vector<MNISTData> tset;
  for (uint32_t i = 0; i < count; ++i) {   // per-label cap, file order (mnist.h:472-496)
    const int l = lab[i];
    if (counts[l] >= NT) continue;
    counts[l] += 1;
    MNISTData t;
    t.n = i;
    t.label = l;
    t.data.resize((size_t)rows * cols);
    const uint8_t* p = &img[(size_t)i * rows * cols];
    for (size_t j = 0; j < t.data.size(); ++j) t.data[j] = p[j] / 255.;   // mnist.h:495
    tset.push_back(std::move(t));
  }
  itensor::printfln("%sing set consists of %d images:", type == Train ? "Train" : "Test", (int)tset.size());
  for (int l = 0; l < 10; ++l) itensor::printfln("  %d of label %d", (int)counts[l], l);
  return tset;
}

// image.h:316-346
inline void reduce(std::vector<MNISTData>& set, long newlen) {
  if (set.empty()) return;
  long L = 1;
  while (L * L < (long)set.front().size()) ++L;
  if (newlen == L) return;
  const long bs = L / newlen, rem = L % bs;
  for (auto& t : set) {
    std::vector<Real> out(newlen * newlen);
    for (long ny = 0; ny < newlen; ++ny)
      for (long nx = 0; nx < newlen; ++nx) {
        Real avg = 0;
        long cnt = 0;
        for (long oy = rem + bs * ny; oy < rem + bs * ny + bs; ++oy)
          for (long ox = rem + bs * nx; ox < rem + bs * nx + bs; ++ox) {
            avg += t.data[oy * L + ox];
            ++cnt;
          }
        out[ny * newlen + nx] = avg / cnt;
      }
    t.data.swap(out);
  }
}

}  // namespace mllib
