// Measured machine peaks that the roofline fractions of bench.py are quoted against but MEASURED_PEAKS.json does not hold
// (SURVEY.md §6 / BASELINE.md §2: "FP64 vector peak: not measured — measure with an FMA loop").
#include "isph_internal.h"

namespace isph {

// 16 independent DFMA chains per thread (enough to cover the FP64 pipe latency at full occupancy), no memory traffic
__global__ void __launch_bounds__(256) k_fp64_fma_loop(double *out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) x[q] = (double)(threadIdx.x + q) * 1.0e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) x[q] = fma(x[q], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 16; ++q) s += x[q];
  if (s == 123.456) out[0] = s;                                   // never true: keeps the loop alive
}

}  // namespace isph

using namespace isph;
extern "C" int isph_measure_fp64_peak(isph_ctx *ctx, double *tflops) {
  if (!ctx || !tflops) return ISPH_FAILURE; Ctx *c = reinterpret_cast<Ctx *>(ctx);
  try {
    CUDA_CHECK(cudaSetDevice(c->device));
    int sms = 0; CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    c->hbuf.ensure(8192);
    const int grid = sms * 8, iters = 1 << 15; cudaEvent_t e0, e1; CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {                          // first repetition = warm-up
      CUDA_CHECK(cudaEventRecord(e0, c->stream));
      k_fp64_fma_loop<<<grid, 256, 0, c->stream>>>(c->hbuf.p + 4200, iters, 0.999999, 1.0e-7); ++c->launches;
      CUDA_CHECK(cudaEventRecord(e1, c->stream)); CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      const double tf = 2.0 * 16.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
  } catch (const std::exception &e) { c->err = e.what(); cudaGetLastError(); return ISPH_FAILURE; }
  return ISPH_SUCCESS;
}
