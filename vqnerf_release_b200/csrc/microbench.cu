// FP32-FMA peak micro-benchmark ("FP32-FMA peak is not in MEASURED_PEAKS.json -- the builder must
// measure it").  Not part of the reference interface; used by bench.py to state the denominator of the
// FFMA-bound kernels' roofline.  mode 0: scalar FFMA, mode 1: packed fma.rn.f32x2 (sm_100 FFMA2).
#include "common.cuh"

template <int MODE>
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          unsigned long long d, x, y, z;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(acc[i]), "f"(acc[i + 1]));
          asm volatile("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(a));
          asm volatile("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(b));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(d));
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;   // keep the chain alive
}

extern "C" int vqn_microbench_fma(vqn_ctx* ctx, int mode, int iters, double* tflops_out) {
  VQN_CHECK_ARG(ctx && tflops_out && iters > 0 && (mode == 0 || mode == 1), "microbench_fma args");
  float* d = nullptr;
  VQN_CUDA(cudaMalloc(&d, 4));
  int blocks = ctx->sm_count * 8;
  cudaEvent_t e0, e1;
  VQN_CUDA(cudaEventCreate(&e0));
  VQN_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    VQN_CUDA(cudaEventRecord(e0, 0));
    if (mode == 0) fma_peak_kernel<0><<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
    else fma_peak_kernel<1><<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
    VQN_LAUNCHED(ctx);
    VQN_CUDA(cudaEventRecord(e1, 0));
    VQN_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    VQN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 256.0 * (double)iters * (double)blocks * 256.0;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return VQN_OK;
}
