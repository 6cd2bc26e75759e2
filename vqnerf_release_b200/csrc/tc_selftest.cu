// Single-tile tcgen05 GEMM used to validate the tensor-core primitives (descriptors, 128-B swizzle, TMEM
// alloc/ld, tcgen05.commit -> mbarrier) on real hardware:  D[128,N] = A[128,K] . B[N,K]^T, fp32 accumulate.
//   mode 0: kind::tf32 single pass (inputs rounded to tf32 by the kernel)
//   mode 1: kind::f16 with bf16 operands (round-to-nearest)
//   mode 2: 3xTF32 split  A_hi B_hi + A_lo B_hi + A_hi B_lo  (fp32-parity arithmetic of VQN_PREC_TF32X3)
//   mode 3: the SHIPPED tf32x3 scheme with the A operand in TENSOR MEMORY (tc_selftest_ts_kernel): per 32-K chunk one TMEM
//           slot of 64 columns -- [tf32 hi x 32 | bf16 pairs of a_lo x 16 | bf16 pairs of a_hi x 16], written with tcgen05.st
//           -- against the weight chunk image of mlp_tc.cu (plane H tf32, plane C = [bf16 w_hi x 32 | bf16 w_lo x 32])
#include "common.cuh"
#include "tc_common.cuh"

__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ D, int N, int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  // manual 1024-B alignment of the dynamic region
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool bf16 = (mode == 1);
  const int E = bf16 ? 64 : 32;            // K elements per 128-byte chunk row
  const int nch = K / E;
  const int planes = (mode == 2) ? 2 : 1;  // hi (+ lo)
  const uint32_t a_tile = 128 * 128, b_tile = (uint32_t)N * 128;
  uint8_t* a_base = smem;                                   // [planes][nch][128 x 128 B]
  uint8_t* b_base = smem + (size_t)planes * nch * a_tile;   // [planes][nch][N x 128 B]

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }

  // fill operand tiles (generic proxy)
  for (int idx = tid; idx < 128 * K; idx += 128) {
    int r = idx / K, k = idx % K, ch = k / E, kk = k % E;
    float v = A[(size_t)r * K + k];
    if (bf16) {
      uint32_t off = tc::sw128_off(r, kk / 8) + (kk % 8) * 2;
      *reinterpret_cast<__nv_bfloat16*>(a_base + (size_t)ch * a_tile + off) = __float2bfloat16_rn(v);
    } else {
      uint32_t off = tc::sw128_off(r, kk / 4) + (kk % 4) * 4;
      float hi = tc::tf32_rna(v);
      *reinterpret_cast<float*>(a_base + (size_t)ch * a_tile + off) = hi;
      if (mode == 2)
        *reinterpret_cast<float*>(a_base + (size_t)(nch + ch) * a_tile + off) = tc::tf32_rna(v - hi);
    }
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    int r = idx / K, k = idx % K, ch = k / E, kk = k % E;
    float v = B[(size_t)r * K + k];
    if (bf16) {
      uint32_t off = tc::sw128_off(r, kk / 8) + (kk % 8) * 2;
      *reinterpret_cast<__nv_bfloat16*>(b_base + (size_t)ch * b_tile + off) = __float2bfloat16_rn(v);
    } else {
      uint32_t off = tc::sw128_off(r, kk / 4) + (kk % 4) * 4;
      float hi = tc::tf32_rna(v);
      *reinterpret_cast<float*>(b_base + (size_t)ch * b_tile + off) = hi;
      if (mode == 2)
        *reinterpret_cast<float*>(b_base + (size_t)(nch + ch) * b_tile + off) = tc::tf32_rna(v - hi);
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(bf16 ? tc::FMT_BF16 : tc::FMT_TF32, 128, N);
    uint32_t acc = 0;
    for (int ch = 0; ch < nch; ++ch) {
      for (int s = 0; s < 4; ++s) {   // four 32-byte K steps per chunk
        const uint32_t koff = 32 * s;
        uint64_t a_hi = tc::make_desc_sw128(tc::smem_u32(a_base + (size_t)ch * a_tile) + koff);
        uint64_t b_hi = tc::make_desc_sw128(tc::smem_u32(b_base + (size_t)ch * b_tile) + koff);
        if (bf16) {
          tc::mma_ss<false>(tmem_base, a_hi, b_hi, idesc, acc); acc = 1;
        } else {
          tc::mma_ss<true>(tmem_base, a_hi, b_hi, idesc, acc); acc = 1;
          if (mode == 2) {
            uint64_t a_lo = tc::make_desc_sw128(tc::smem_u32(a_base + (size_t)(nch + ch) * a_tile) + koff);
            uint64_t b_lo = tc::make_desc_sw128(tc::smem_u32(b_base + (size_t)(nch + ch) * b_tile) + koff);
            tc::mma_ss<true>(tmem_base, a_lo, b_hi, idesc, 1);
            tc::mma_ss<true>(tmem_base, a_hi, b_lo, idesc, 1);
          }
        }
      }
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  // epilogue: warp w owns TMEM lanes [32w, 32w+32); thread -> row tid
  const int ncb = (N + 31) / 32;
  for (int cb = 0; cb < ncb; ++cb) {
    float v[32];
    tc::tmem_ld32(tmem_base + ((uint32_t)(32 * warp) << 16) + cb * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (cb * 32 + j < N) D[(size_t)tid * N + cb * 32 + j] = v[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

// mode 3 (see the header): K <= 128 so that accumulator (256 columns) + A slots (64 per chunk) fit the 512 columns
__global__ void __launch_bounds__(128, 1) tc_selftest_ts_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nch = K / 32;
  const uint32_t plane = (uint32_t)N * 128;                 // chunk image: plane H, then plane C
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  for (int idx = tid; idx < N * K; idx += 128) {
    const int r = idx / K, k = idx % K, ch = k / 32, kk = k % 32;
    const float v = B[(size_t)r * K + k], hi = tc::tf32_rna(v);
    uint8_t* base = smem + (size_t)ch * 2 * plane;
    *reinterpret_cast<float*>(base + tc::sw128_off(r, kk / 4) + (kk % 4) * 4) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + plane + tc::sw128_off(r, kk / 8) + (kk % 8) * 2) = __float2bfloat16_rn(hi);
    *reinterpret_cast<__nv_bfloat16*>(base + plane + tc::sw128_off(r, 4 + kk / 8) + (kk % 8) * 2) = __float2bfloat16_rn(v - hi);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * warp) << 16);
  // thread tid = row tid: its K values of every chunk -> TMEM slot at column 256 + 64 ch
  for (int ch = 0; ch < nch; ++ch) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t hi[16], lo2[8], hi2[8];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float v0 = A[(size_t)tid * K + ch * 32 + 16 * h + j], v1 = A[(size_t)tid * K + ch * 32 + 16 * h + j + 1];
        const float h0 = tc::tf32_rna(v0), h1 = tc::tf32_rna(v1);
        hi[j] = __float_as_uint(h0); hi[j + 1] = __float_as_uint(h1);
        __nv_bfloat162 pl = __floats2bfloat162_rn(v0 - h0, v1 - h1), ph = __floats2bfloat162_rn(h0, h1);
        lo2[j / 2] = *reinterpret_cast<uint32_t*>(&pl); hi2[j / 2] = *reinterpret_cast<uint32_t*>(&ph);
      }
      const uint32_t slot = lane_addr + 256u + 64u * (uint32_t)ch;
      tc::tmem_st16(slot + 16u * h, hi);
      tc::tmem_st8(slot + 32u + 8u * h, lo2);
      tc::tmem_st8(slot + 48u + 8u * h, hi2);
    }
  }
  tc::tmem_st_wait();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, N), idesc_c = tc::make_idesc(tc::FMT_BF16, 128, N);
    uint32_t acc = 0;
    for (int ch = 0; ch < nch; ++ch) {
      const uint32_t w_addr = tc::smem_u32(smem + (size_t)ch * 2 * plane);
      const uint32_t a_slot = tmem_base + 256u + 64u * (uint32_t)ch;
      for (int s = 0; s < 4; ++s) {
        tc::mma_ts<true>(tmem_base, a_slot + 8u * s, tc::make_desc_sw128(w_addr + 32 * s), idesc, acc); acc = 1;
        tc::mma_ts<false>(tmem_base, a_slot + 32u + 8u * s, tc::make_desc_sw128(w_addr + plane + 32 * s), idesc_c, 1);
      }
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int cb = 0; cb < (N + 31) / 32; ++cb) {
    float v[32];
    tc::tmem_ld32(lane_addr + cb * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (cb * 32 + j < N) D[(size_t)tid * N + cb * 32 + j] = v[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

extern "C" int vqn_tc_selftest(vqn_ctx* ctx, int mode, int n, int k, const float* a, const float* b, float* d,
                               vqn_stream stream) {
  VQN_CHECK_ARG(ctx && a && b && d, "tc_selftest: null");
  VQN_CHECK_ARG(mode >= 0 && mode <= 3, "tc_selftest: mode");
  VQN_CHECK_ARG(n >= 16 && n <= 256 && n % 16 == 0, "tc_selftest: N multiple of 16 in [16,256]");
  if (mode == 3) {
    VQN_CHECK_ARG(k >= 32 && k <= 128 && k % 32 == 0, "tc_selftest: mode 3 needs K in {32, 64, 96, 128}");
    size_t smem3 = (size_t)(k / 32) * 2 * n * 128 + 1024;
    VQN_CUDA(cudaFuncSetAttribute(tc_selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    tc_selftest_ts_kernel<<<1, 128, smem3, vqn_cs(stream)>>>(a, b, d, n, k);
    VQN_LAUNCHED(ctx);
    return VQN_OK;
  }
  const int e = mode == 1 ? 64 : 32;
  VQN_CHECK_ARG(k >= e && k % e == 0, "tc_selftest: K multiple of the chunk");
  const int planes = mode == 2 ? 2 : 1;
  size_t smem = (size_t)planes * (k / e) * (128 * 128 + (size_t)n * 128) + 1024;
  VQN_CHECK_ARG((int)smem <= ctx->max_smem_optin, "tc_selftest: tiles exceed shared memory");
  VQN_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, vqn_cs(stream)>>>(a, b, d, n, k, mode);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
