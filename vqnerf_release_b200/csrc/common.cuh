// Shared device/host helpers for libvqnerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/vqnerf_b200.h"

#define VQN_SM_COUNT_FALLBACK 148

struct vqn_ctx {
  int device;
  int sm_count;
  int max_smem_optin;
  int* nonfinite_flag;  // device int, sticky check_numerics flag
  int* scratch;         // small persistent device scratch (block counts of the mask compaction, VQ max distance)
  size_t scratch_ints;
  std::atomic<long long> launches;
  void* pool;           // per-(kind, stream) device buffers of the tensor-core kernels (abi.cu: vqn_stream_scratch)
};

// Library-owned work buffers that must not be shared between streams (latent scratch of mlp_main, act' stash of the SDF
// gradient, codebook images of vq_tc): one buffer per (kind, stream), allocated on first use, freed with the context.
enum { VQN_SCRATCH_Z = 0, VQN_SCRATCH_STASH = 1, VQN_SCRATCH_VQ_W = 2, VQN_SCRATCH_VQ_C2 = 3, VQN_SCRATCH_SHADE = 4 };
void* vqn_stream_scratch(vqn_ctx* ctx, int kind, cudaStream_t stream, size_t bytes);   // nullptr: allocation failed
#define VQN_SCRATCH_INTS (1 << 16)

void vqn_set_error(const char* fmt, ...);

#define VQN_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      vqn_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
      return VQN_ERR_INVALID_ARG;                                  \
    }                                                              \
  } while (0)

#define VQN_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      vqn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return VQN_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// after a kernel launch
#define VQN_LAUNCHED(ctx)                         \
  do {                                            \
    (ctx)->launches.fetch_add(1);                 \
    VQN_CUDA(cudaGetLastError());                 \
  } while (0)

static inline cudaStream_t vqn_cs(vqn_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float vqn_sigmoid(float x) {
  // Keras sigmoid in fp32; expf (not __expf) keeps <= 2 ulp
  return 1.0f / (1.0f + expf(-x));
}

__device__ __forceinline__ float vqn_apply_act(float x, int act) {
  if (act == VQN_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == VQN_ACT_SIGMOID) return vqn_sigmoid(x);
  return x;
}

__device__ __forceinline__ float vqn_linear2srgb(float v) {
  // util/img.py:142-165
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  float lin = v * 12.92f;
  float nl = 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
  return v <= 0.0031308f ? lin : nl;
}

__device__ __forceinline__ float vqn_srgb2linear(float v) {
  // util/img.py:167-186
  float lin = v / 12.92f;
  float nl = powf((v + 0.055f) / 1.055f, 2.4f);
  return v <= 0.04045f ? lin : nl;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  // streaming 128-bit load, no L1 allocation (read-once data: lvis rows, latents)
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
