// Shared declarations of the VQ kernels (vq.cu: C ABI + EMA; vq_mma.cu: the assignment kernel).
#pragma once
#include "common.cuh"

#define VQ_Z 256

struct VqParams {
  const float* x;
  long long n;
  const float* cb;       // [256,K]
  int K;
  const float* sel_mask; // [K] or null
  unsigned* maxdist;     // ordered-uint global max of distances (two-pass thres path)
  int normalize;
  long long* idx_out;
  float* quant_out;
  float* dist_out;
  float* znorm_out;
  double* stats;         // [K + 2 + 256*K]
  int want_dw;
};

// order-preserving float <-> uint maps (atomicMax / redux.min on distances)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// vq_mma.cu: launches the (optional) max-distance pass and the assignment pass for any K
int vq_assign_mma_launch(vqn_ctx* ctx, VqParams p, cudaStream_t s);
// vq_tc.cu: K > 32, indices only, on tcgen05/TMEM (3xTF32 split, fp64 near-tie re-score)
int vq_tc_assign_launch(vqn_ctx* ctx, const VqParams& p, cudaStream_t s);
int vq_tc_min_k();   // smallest K routed to vq_tc.cu (default 33; VQN_VQ_TC_MIN_K overrides, for measurements)
