// oz_test.cu -- standalone check + timing of the tcgen05 int8 (Ozaki) projection kernel against a
// long-double CPU evaluation and against the FP64 DMMA kernel (tnml::krgemm) on the same operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/oz_test tools/oz_test.cu \
//        tnml_b200/csrc/tnml_ozaki.cu tnml_b200/csrc/tnml_kernels.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tnml_b200/csrc/tnml_kernels.cuh"

using namespace tnml;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0;
}

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

int main(int argc, char** argv) {
  struct Shape {
    long rows;
    int ma, S, J, div;
  };
  std::vector<Shape> shapes = {{300, 120, 4, 120, 1},   {1000, 77, 4, 50, 1},   {4096, 128, 4, 120, 1},
                               {60000, 120, 4, 120, 1}, {60000, 120, 2, 120, 1}, {8192, 120, 4, 1200, 1},
                               {20000, 100, 2, 95, 10}, {7500, 120, 4, 120, 1}};
  int nslices[] = {8, 7, 6};
  int only = (argc > 1) ? atoi(argv[1]) : -1;
  int fails = 0;
  for (size_t si = 0; si < shapes.size(); ++si) {
    if (only >= 0 && (int)si != only) continue;
    const Shape sh = shapes[si];
    const long nimg = (sh.rows + sh.div - 1) / sh.div;
    std::vector<double> hin((size_t)sh.rows * sh.ma), hb((size_t)sh.S * sh.ma * sh.J), hf1((size_t)nimg * 2), hf2((size_t)nimg * 2);
    for (auto& v : hin) v = (urand() - 0.5) * std::exp(6.0 * (urand() - 0.5));          // wide dynamic range
    for (long r = 0; r < sh.rows; r += 37) for (int a = 0; a < sh.ma; ++a) hin[(size_t)r * sh.ma + a] *= 1e-7;  // small rows
    for (auto& v : hb) v = (urand() - 0.5) * std::exp(4.0 * (urand() - 0.5));
    for (long n = 0; n < nimg; ++n) {
      hf1[2 * n] = 1.0;
      hf1[2 * n + 1] = 1e-3 * urand();
      hf2[2 * n] = 1.0;
      hf2[2 * n + 1] = 1e-3 * urand();
    }
    double *in, *b, *f1, *f2, *out, *out2, *ea, *eb;
    int8_t *A8, *B8;
    CK(cudaMalloc(&in, hin.size() * 8));
    CK(cudaMalloc(&b, hb.size() * 8));
    CK(cudaMalloc(&f1, hf1.size() * 8));
    CK(cudaMalloc(&f2, hf2.size() * 8));
    CK(cudaMalloc(&out, (size_t)sh.rows * sh.J * 8));
    CK(cudaMalloc(&out2, (size_t)sh.rows * sh.J * 8));
    CK(cudaMemcpy(in, hin.data(), hin.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(f1, hf1.data(), hf1.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(f2, hf2.data(), hf2.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&A8, oz_a8_bytes(sh.rows, 8)));
    CK(cudaMalloc(&B8, oz_b8_bytes(sh.S, sh.J, 8)));
    CK(cudaMalloc(&ea, oz_rows_pad(sh.rows) * 8));
    CK(cudaMalloc(&eb, oz_cols_pad(sh.S, sh.J) * 8));
    // FP64 DMMA kernel on the same operands
    krgemm(0, sh.S, in, sh.ma, sh.ma, f1, f2, sh.div, b, sh.J, sh.J, out2, sh.J, sh.rows, 148);
    CK(cudaDeviceSynchronize());
    std::vector<double> ho2((size_t)sh.rows * sh.J);
    CK(cudaMemcpy(ho2.data(), out2, ho2.size() * 8, cudaMemcpyDeviceToHost));
    // long-double reference on a sample of rows
    std::vector<long> srows;
    for (long r = 0; r < sh.rows; r += (sh.rows > 4000 ? sh.rows / 997 : 1)) srows.push_back(r);
    srows.push_back(sh.rows - 1);
    std::vector<long double> ref(srows.size() * sh.J);
    double refmax = 0.0;
    for (size_t k = 0; k < srows.size(); ++k) {
      const long r = srows[k], img = r / sh.div;
      double w[4];
      if (sh.S == 2) {
        w[0] = hf1[2 * img];
        w[1] = hf1[2 * img + 1];
      } else {
        w[0] = hf1[2 * img] * hf2[2 * img];
        w[1] = hf1[2 * img] * hf2[2 * img + 1];
        w[2] = hf1[2 * img + 1] * hf2[2 * img];
        w[3] = hf1[2 * img + 1] * hf2[2 * img + 1];
      }
      for (int j = 0; j < sh.J; ++j) {
        long double acc = 0.0L;
        for (int p = 0; p < sh.S; ++p) {
          long double t = 0.0L;
          for (int a = 0; a < sh.ma; ++a)
            t += (long double)hin[(size_t)r * sh.ma + a] * (long double)hb[((size_t)a * sh.S + p) * sh.J + j];
          acc += (long double)w[p] * t;
        }
        ref[k * sh.J + j] = acc;
      }
    }
    for (int ns : nslices) {
      if (!oz_supported(sh.S, sh.ma, ns)) continue;
      CK(cudaMemset(out, 0xFF, (size_t)sh.rows * sh.J * 8));
      oz_slice_rows(0, in, sh.ma, sh.ma, sh.rows, ns, A8, ea);
      oz_slice_cols(0, sh.S, b, sh.J, sh.ma, sh.J, ns, B8, eb);
      bool ok = oz_krgemm(0, sh.S, ns, A8, ea, sh.rows, f1, f2, sh.div, B8, eb, sh.J, out, sh.J, 148);
      cudaError_t e = cudaDeviceSynchronize();
      if (!ok || e != cudaSuccess) {
        printf("shape %zu ns %d: launch ok=%d, %s\n", si, ns, (int)ok, cudaGetErrorString(e));
        return 3;
      }
      std::vector<double> ho((size_t)sh.rows * sh.J);
      CK(cudaMemcpy(ho.data(), out, ho.size() * 8, cudaMemcpyDeviceToHost));
      // per-row normalised error vs long double (sampled rows) and vs the DMMA kernel (all rows)
      double worst = 0.0, worst2 = 0.0, worstd = 0.0;
      for (size_t k = 0; k < srows.size(); ++k) {
        const long r = srows[k];
        long double rm = 0.0L;
        for (int j = 0; j < sh.J; ++j) rm = fmaxl(rm, fabsl(ref[k * sh.J + j]));
        if (rm == 0.0L) rm = 1.0L;
        for (int j = 0; j < sh.J; ++j) {
          worst = fmax(worst, (double)(fabsl((long double)ho[(size_t)r * sh.J + j] - ref[k * sh.J + j]) / rm));
          worstd = fmax(worstd, (double)(fabsl((long double)ho2[(size_t)r * sh.J + j] - ref[k * sh.J + j]) / rm));
        }
      }
      long nbad = 0;
      for (long r = 0; r < sh.rows; ++r) {
        double rm = 0.0;
        for (int j = 0; j < sh.J; ++j) rm = fmax(rm, fabs(ho2[(size_t)r * sh.J + j]));
        if (rm == 0.0) rm = 1.0;
        for (int j = 0; j < sh.J; ++j) {
          const double d = fabs(ho[(size_t)r * sh.J + j] - ho2[(size_t)r * sh.J + j]) / rm;
          if (!(d < 1e-6)) ++nbad;
          if (d == d) worst2 = fmax(worst2, d);
        }
      }
      // timing
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      const int reps = 20;
      for (int i = 0; i < 3; ++i) oz_krgemm(0, sh.S, ns, A8, ea, sh.rows, f1, f2, sh.div, B8, eb, sh.J, out, sh.J, 148);
      cudaEventRecord(e0);
      for (int i = 0; i < reps; ++i) oz_krgemm(0, sh.S, ns, A8, ea, sh.rows, f1, f2, sh.div, B8, eb, sh.J, out, sh.J, 148);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms, ms_sr, ms_sc, ms_d;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= reps;
      // OZ_TEST_FLUSH=1: the same launch with a cold L2 (512 MB memset before every launch), as inside a bond
      // update where the label-environment stream runs between two projections
      float ms_cold = 0.f;
      if (getenv("OZ_TEST_FLUSH")) {
        static char* flush = nullptr;
        if (!flush) CK(cudaMalloc(&flush, (size_t)512 << 20));
        for (int i = 0; i < 10; ++i) {
          cudaMemsetAsync(flush, i, (size_t)512 << 20);
          cudaEventRecord(e0);
          oz_krgemm(0, sh.S, ns, A8, ea, sh.rows, f1, f2, sh.div, B8, eb, sh.J, out, sh.J, 148);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float t;
          cudaEventElapsedTime(&t, e0, e1);
          ms_cold += t / 10;
        }
        printf("   cold-L2 oz launch: %.4f ms (warm %.4f)\n", ms_cold, ms);
      }
      cudaEventRecord(e0);
      for (int i = 0; i < reps; ++i) oz_slice_rows(0, in, sh.ma, sh.ma, sh.rows, ns, A8, ea);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms_sr, e0, e1);
      ms_sr /= reps;
      cudaEventRecord(e0);
      for (int i = 0; i < reps; ++i) oz_slice_cols(0, sh.S, b, sh.J, sh.ma, sh.J, ns, B8, eb);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms_sc, e0, e1);
      ms_sc /= reps;
      cudaEventRecord(e0);
      for (int i = 0; i < reps; ++i) krgemm(0, sh.S, in, sh.ma, sh.ma, f1, f2, sh.div, b, sh.J, sh.J, out2, sh.J, sh.rows, 148);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms_d, e0, e1);
      ms_d /= reps;
#ifdef OZ_PROFILE
      {
        long long* dbg;
        CK(cudaMalloc(&dbg, 148 * 8 * sizeof(long long)));
        CK(cudaMemset(dbg, 0, 148 * 8 * sizeof(long long)));
        oz_set_debug_buffer(dbg);
        oz_krgemm(0, sh.S, ns, A8, ea, sh.rows, f1, f2, sh.div, B8, eb, sh.J, out, sh.J, 148);
        CK(cudaDeviceSynchronize());
        oz_set_debug_buffer(nullptr);
        std::vector<long long> hd(148 * 8);
        CK(cudaMemcpy(hd.data(), dbg, hd.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double s[8] = {0};
        int nc = 0;
        for (int c = 0; c < 148; ++c) if (hd[c * 8 + 3] > 0) { ++nc; for (int k = 0; k < 8; ++k) s[k] += (double)hd[c * 8 + k]; }
        if (nc) printf("   profile (mean cycles per CTA over %d CTAs, %.1f tiles): mma waits b_full %.0f acc_empty %.0f a_full %.0f | kernel %.0f | "
                       "epilogue wait %.0f tmem+math %.0f store %.0f\n", nc, s[7] / nc, s[0] / nc, s[1] / nc, s[2] / nc, s[3] / nc, s[4] / nc,
                       s[5] / nc, s[6] / nc);
        cudaFree(dbg);
      }
#endif
      const double fl = 2.0 * sh.rows * sh.S * sh.ma * sh.J;
      const double iops = 2.0 * sh.rows * sh.S * 128.0 * sh.J * (ns * (ns + 1) / 2);
      const double tol = (ns == 8) ? fmax(1e-13, 50.0 * worstd) : (ns == 7 ? 1e-11 : 1e-9);
      const bool pass = worst < tol && nbad == 0;
      if (!pass) ++fails;
      printf("rows %6ld ma %3d S %d J %4d div %2d ns %d: err_vs_ld %.2e (dmma %.2e) max|oz-dmma| %.2e bad %ld | oz %.4f ms = %.1f TF/s-equiv, "
             "%.0f TOPS int8 | slice rows %.4f cols %.4f ms | dmma %.4f ms %.1f TF/s  %s\n",
             sh.rows, sh.ma, sh.S, sh.J, sh.div, ns, worst, worstd, worst2, nbad, ms, fl / ms / 1e9, iops / ms / 1e9, ms_sr, ms_sc,
             ms_d, fl / ms_d / 1e9, pass ? "PASS" : "FAIL");
      fflush(stdout);
    }
    cudaFree(in); cudaFree(b); cudaFree(f1); cudaFree(f2); cudaFree(out); cudaFree(out2);
    cudaFree(A8); cudaFree(B8); cudaFree(ea); cudaFree(eb);
  }
  printf(fails ? "OZ_TEST FAILED (%d)\n" : "OZ_TEST OK\n", fails);
  return fails ? 1 : 0;
}
