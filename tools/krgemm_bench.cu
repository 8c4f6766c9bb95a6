// krgemm_bench.cu -- standalone timing of the projection kernel (tnml::krgemm) on the bench shape
// (rows 60000, ma 120, S 4, J 120) and a few others, without the rest of the library.  Builds the
// kernels' translation unit directly so that experiment switches (-DKR2_NO_EPILOGUE, ...) apply:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DKR2_...] -o tools/krgemm_bench tools/krgemm_bench.cu
#include "../tnml_b200/csrc/tnml_kernels.cu"
#include <vector>
int main(int argc, char** argv) {
  using namespace tnml;
  struct Shape { long rows; int ma, S, J, div; };
  Shape shapes[] = {{60000, 120, 4, 120, 1}, {60000, 120, 2, 120, 1}, {600000, 120, 2, 120, 10}, {60000, 77, 4, 77, 1},
                    {60000, 120, 4, 1200, 1}, {7500, 120, 4, 120, 1}};
  for (auto sh : shapes) {
    std::vector<double> hin((size_t)sh.rows * sh.ma), hb((size_t)sh.S * sh.ma * sh.J), hf((size_t)sh.rows * 2 / sh.div + 2);
    for (size_t i = 0; i < hin.size(); ++i) hin[i] = 1.0 + 1e-3 * (i % 97);
    for (size_t i = 0; i < hb.size(); ++i) hb[i] = 1.0 - 1e-3 * (i % 89);
    for (size_t i = 0; i < hf.size(); ++i) hf[i] = (i & 1) ? 1e-3 * (i % 13) : 1.0;
    double *in, *b, *f, *out;
    cudaMalloc(&in, hin.size() * 8); cudaMalloc(&b, hb.size() * 8); cudaMalloc(&f, hf.size() * 8);
    cudaMalloc(&out, (size_t)sh.rows * sh.J * 8);
    cudaMemcpy(in, hin.data(), hin.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(b, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(f, hf.data(), hf.size() * 8, cudaMemcpyHostToDevice);
    for (int variant = 1; variant <= 2; ++variant) {
      krgemm_set_variant(variant);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 3; ++i) krgemm(0, sh.S, in, sh.ma, sh.ma, f, f, sh.div, b, sh.J, sh.J, out, sh.J, sh.rows, 148);
      cudaEventRecord(e0);
      const int reps = 20;
      for (int i = 0; i < reps; ++i) krgemm(0, sh.S, in, sh.ma, sh.ma, f, f, sh.div, b, sh.J, sh.J, out, sh.J, sh.rows, 148);
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
      double fl = 2.0 * sh.rows * sh.S * sh.ma * sh.J;
      std::vector<double> ho(8);
      cudaMemcpy(ho.data(), out + (size_t)(sh.rows - 1) * sh.J, 8 * 8 > sh.J * 8 ? sh.J * 8 : 64, cudaMemcpyDeviceToHost);
      printf("rows %7ld ma %3d S %d J %4d variant %d: %.4f ms  %.2f TF/s  out[last][0]=%.6e (%s)\n", sh.rows, sh.ma, sh.S,
             sh.J, variant, ms, fl / ms / 1e9, ho[0], cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(in); cudaFree(b); cudaFree(f); cudaFree(out);
  }
  return 0;
}
