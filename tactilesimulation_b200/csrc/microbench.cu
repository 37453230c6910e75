// Measurement aid of the C ABI (include/tactilesim_b200.h): the fp64 FMA issue peak of the device the library runs on.
// The hot path is fp64 vector arithmetic (DH/Common.h:24: dtype = double); MEASURED_PEAKS.json holds HBM and bf16
// tensor peaks only, so the denominator of the fp64 roofline is measured here.
//
// Kernel: every thread runs 8 independent chains of dependent DFMAs (enough to cover the pipe latency at full
// occupancy); 2 blocks of 1024 threads per SM.  flops = 2 per FMA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace {
thread_local std::string g_mb_err;

__global__ void __launch_bounds__(1024, 2) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
  double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // never true: keeps the chains alive
}
}  // namespace

extern "C" const char* tsim_debug_last_error(void) { return g_mb_err.c_str(); }

// out[0] = achieved GFLOP/s (fp64, FMA = 2 flops), out[1] = kernel ms, out[2] = SM count, out[3] = SM clock (MHz, max)
extern "C" int tsim_debug_fp64_peak(int device, double* out) {
  if (!out) { g_mb_err = "tsim_debug_fp64_peak: null argument"; return 1; }
#define MB_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_mb_err = std::string(#x) + ": " + cudaGetErrorString(e_); return 1; } } while (0)
  MB_CK(cudaSetDevice(device));
  int nsm = 0, khz = 0;
  MB_CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  MB_CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
  double* d = 0;
  const int grid = nsm * 2, block = 1024, iters = 4096;
  MB_CK(cudaMalloc(&d, sizeof(double) * (size_t)grid * block));
  cudaEvent_t e0, e1;
  MB_CK(cudaEventCreate(&e0));
  MB_CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {          // first repetition warms the clocks up
    MB_CK(cudaEventRecord(e0, 0));
    dfma_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);
    MB_CK(cudaEventRecord(e1, 0));
    MB_CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    MB_CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  MB_CK(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double fmas = (double)grid * block * (double)iters * 64.0;
  out[0] = 2.0 * fmas / (best * 1e-3) / 1e9;
  out[1] = best;
  out[2] = nsm;
  out[3] = khz / 1000.0;
  return 0;
#undef MB_CK
}
