// FP64 CUDA-core peak of the device: independent DFMA / DADD / DMUL chains, all SMs busy.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu
// Prints one JSON line; bench.py / profiles quote it as the FP64 roofline denominator.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void __launch_bounds__(256) chains(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) x[i] = fma(x[i], a, b);
      else if (OP == 1) x[i] = x[i] + a;
      else if (OP == 2) x[i] = x[i] * a;
      else { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i])); x[i] = y; }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int OP, int ILP>
double run(int sms, int blocks_per_sm, int threads, double *out) {
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  chains<OP, ILP><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    chains<OP, ILP><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double ops = (double)sms * blocks_per_sm * threads * (double)iters * ILP;
  return ops / (best * 1e-3);   // thread-instructions per second
}

int main() {
  int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double *out; cudaMalloc(&out, 8);
  double fma8 = run<0, 8>(sms, 4, 256, out);
  double add8 = run<1, 8>(sms, 4, 256, out);
  double mul8 = run<2, 8>(sms, 4, 256, out);
  double rcp8 = run<3, 8>(sms, 4, 256, out);
  double fma1_16w = run<0, 1>(sms, 4, 128, out);   // 16 warps/SM, one dependent chain each: latency-bound rate
  double fma2_16w = run<0, 2>(sms, 4, 128, out);
  double fma4_16w = run<0, 4>(sms, 4, 128, out);
  double hz = khz * 1e3;
  printf("{\"sms\": %d, \"clock_mhz\": %.0f, \"dfma_tflops\": %.2f, \"dfma_per_clk_sm\": %.2f, \"dadd_per_clk_sm\": %.2f, \"dmul_per_clk_sm\": %.2f, "
         "\"mufu_rcp64h_per_clk_sm\": %.2f, \"dfma_per_clk_sm_16warps_ilp1\": %.2f, \"ilp2\": %.2f, \"ilp4\": %.2f}\n",
         sms, hz / 1e6, 2 * fma8 / 1e12, fma8 / hz / sms, add8 / hz / sms, mul8 / hz / sms, rcp8 / hz / sms,
         fma1_16w / hz / sms, fma2_16w / hz / sms, fma4_16w / hz / sms);
  return 0;
}
