// Dense-layer forward and backward-data of the training step on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Reference: networks/mlp.py:44-48 under tf.GradientTape (train_nfr.py:562-576).  Same contract as the warp-level
// kernels of train_dense.cu (vqn_dense_forward / vqn_dense_backward_data: arbitrary leading dimensions so layers read
// and write slices of the skip-connection concat buffers, activations kept / act' taken from the stored activation),
// same 3-term fp32-parity split (kind::tf32 leading term + ONE kind::f16 bf16 MMA of twice the K for both correction
// terms), but the products run as M = 128 UMMAs with the accumulator in TMEM:
//
//   C[M, N] = epilogue( A[M, K] . B[K, N] )      forward:  A = X,  B = W          epilogue: act(. + b) * scale + bias
//                                                 bwd data: A = dZ, B = W^T        epilogue: . * act'(Y_prev) (+= dX)
//
// One CTA per (128-row block, 128-column block).  Neither operand is pre-packed (the weights change every step and the
// C ABI takes plain pointers): of each 256-thread producer group, 128 threads own an A row and 128 threads a B row of
// the current 32-wide K chunk; they load their 32 values, split them (tf32 hi plane + bf16 correction plane) into the
// 128-B-swizzled K-major slot and hand it to the MMA thread through mbarriers; two groups alternate chunks.
// Measured at 8192 rows (benchmarks/dense_micro.py): 12-18 us forward, 15-32 us backward-data per layer against 8-29 /
// 11-32 us on the mma.sync path: a win only on the 256 x 256 layers (18 vs 29 us), which is where vqn_dense_forward
// routes to it.  The batch of a training step is too small for per-layer tensor-core kernels to pay off; the next step
// is the fused multi-layer form of mlp_tc.cu with activation saves.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

#define DT_M 128
#define DT_N 128
#define DT_PLANE (128u * 128u)                 // 16 KB: 128 rows x 128 B
#define DT_OPERAND (2u * DT_PLANE)             // tf32 plane + bf16 correction plane
#define DT_SLOT (2u * DT_OPERAND)              // A + B
#define DT_STAGES 2
#define DT_THREADS (32 * 17)                   // 16 producer/epilogue warps + 1 MMA warp
#define DT_SMEM (DT_STAGES * DT_SLOT + 1024)

namespace {

struct DtParams {
  const float* A; long long lda;
  const float* W; long long ldw;
  float* C; long long ldc;
  int M, N, K;                                 // GEMM sizes (bwd data: N = layer inputs, K = layer outputs)
  int bwd;                                     // 0: B(n,k) = W[k*ldw + n];  1: B(n,k) = W[n*ldw + k]
  const float* bias; int act; float out_scale, out_bias;
  const float* yprev; long long ldy; int act_prev; int accumulate;
};

// 16 consecutive K values of row r -> tf32 hi plane + bf16 correction plane (layout of mlp_tc.cu).  The correction MMA
// evaluates a_lo.b_hi + a_hi.b_lo in ONE product of K = 64, so the A operand stores [bf16(lo) x 32 | bf16(hi) x 32] and
// the B operand the opposite order [bf16(hi) x 32 | bf16(lo) x 32].
__device__ __forceinline__ void dt_store16(uint32_t op, int r, int j0, const float* v, bool b_operand) {
  const uint32_t row = op + (uint32_t)r * 128u;             // shared-space address: STS, not generic stores
  const uint32_t rx = (uint32_t)(r & 7);
#pragma unroll
  for (int qq = 0; qq < 2; ++qq) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { h[i] = tc::tf32_rna_fast(v[8 * qq + i]); l[i] = v[8 * qq + i] - h[i]; }
    const uint32_t c0 = (uint32_t)(j0 / 4 + 2 * qq);
    tc::sts128(row + (((c0) ^ rx) << 4), make_float4(h[0], h[1], h[2], h[3]));
    tc::sts128(row + (((c0 + 1) ^ rx) << 4), make_float4(h[4], h[5], h[6], h[7]));
    uint4 ul, uh;
    __nv_bfloat162 t0 = __floats2bfloat162_rn(l[0], l[1]), t1 = __floats2bfloat162_rn(l[2], l[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(l[4], l[5]), t3 = __floats2bfloat162_rn(l[6], l[7]);
    ul.x = *reinterpret_cast<uint32_t*>(&t0); ul.y = *reinterpret_cast<uint32_t*>(&t1);
    ul.z = *reinterpret_cast<uint32_t*>(&t2); ul.w = *reinterpret_cast<uint32_t*>(&t3);
    t0 = __floats2bfloat162_rn(h[0], h[1]); t1 = __floats2bfloat162_rn(h[2], h[3]);
    t2 = __floats2bfloat162_rn(h[4], h[5]); t3 = __floats2bfloat162_rn(h[6], h[7]);
    uh.x = *reinterpret_cast<uint32_t*>(&t0); uh.y = *reinterpret_cast<uint32_t*>(&t1);
    uh.z = *reinterpret_cast<uint32_t*>(&t2); uh.w = *reinterpret_cast<uint32_t*>(&t3);
    const uint32_t cc = (uint32_t)(j0 / 8 + qq);
    tc::sts128(row + DT_PLANE + ((cc ^ rx) << 4), b_operand ? uh : ul);
    tc::sts128(row + DT_PLANE + (((cc + 4) ^ rx) << 4), b_operand ? ul : uh);
  }
}

// 32 consecutive floats starting at p (elements >= n_valid read as zero); vectorised when the address allows
__device__ __forceinline__ void dt_load_row32(const float* p, int n_valid, float (&v)[32]) {
  if (n_valid >= 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = j < n_valid ? __ldg(p + j) : 0.f;
  }
}

__device__ __forceinline__ void dense_tc_body(const DtParams& p, const int bx, const int by) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[DT_STAGES], empty[DT_STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MMA_WARP = 16;
  const int m0 = bx * DT_M, n0 = by * DT_N;
  const int nvalid = min(DT_N, p.N - n0), npad = (nvalid + 15) / 16 * 16;
  const int nch = (p.K + 31) / 32;

  if (warp == MMA_WARP) tc::tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) {
    for (int i = 0; i < DT_STAGES; ++i) { tc::mbar_init(&full[i], 256); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(&acc_full, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 16) {
    const int grp = warp >> 3, i = tid & 255;
    const bool a_role = i < 128;
    const int r = a_role ? i : i - 128;                       // A row / B row (= output column) of this thread
    for (int c = grp; c < nch; c += 2) {
      const int slot = c & 1, k0 = c * 32, kv = min(32, p.K - k0);
      float v[32];
      if (a_role) {
        const int m = m0 + r;
        if (m < p.M) dt_load_row32(p.A + (long long)m * p.lda + k0, kv, v);
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
      } else if (r < nvalid) {
        if (p.bwd) dt_load_row32(p.W + (long long)(n0 + r) * p.ldw + k0, kv, v);
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = j < kv ? __ldg(p.W + (long long)(k0 + j) * p.ldw + n0 + r) : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      tc::mbar_wait(&empty[slot], (((uint32_t)c / DT_STAGES) & 1u) ^ 1u);
      const uint32_t op = tc::smem_u32(smem) + (uint32_t)slot * DT_SLOT + (a_role ? 0u : DT_OPERAND);
      dt_store16(op, r, 0, v, !a_role);
      dt_store16(op, r, 16, v + 16, !a_role);
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[slot]);
    }
    // ---- epilogue: 4 warps share a TMEM lane quarter; each thread handles two 16-column pieces of its row ----
    tc::mbar_wait(&acc_full, 0);
    tc::fence_after_sync();
    const int row = 32 * (warp & 3) + lane, m = m0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const int piece0 = warp >> 2;                             // 0..3
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      const int c16 = 16 * (piece0 + 4 * t);
      if (c16 >= npad) continue;                              // warp-uniform
      float v[16];
      tc::tmem_ld16(lane_addr + (uint32_t)c16, v);
      if (m < p.M) {                                          // (no early exit: the next tcgen05.ld is warp-collective)
        float* dst = p.C + (long long)m * p.ldc + n0 + c16;
        const int cnt = min(16, nvalid - c16);
        if (!p.bwd) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < cnt) {
              float x = v[j] + (p.bias ? __ldg(p.bias + n0 + c16 + j) : 0.f);
              v[j] = p.out_scale * vqn_apply_act(x, p.act) + p.out_bias;
            }
          }
        } else {
          if (p.act_prev != VQN_ACT_NONE) {
            const float* yp = p.yprev + (long long)m * p.ldy + n0 + c16;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (j < cnt) {
                const float y = yp[j];
                v[j] *= (p.act_prev == VQN_ACT_RELU) ? (y > 0.f ? 1.f : 0.f) : y * (1.f - y);
              }
            }
          }
          if (p.accumulate == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < cnt) v[j] += dst[j];
          }
        }
        if (p.bwd && p.accumulate == 2) {                       // several problems of a batch add into the same buffer
#pragma unroll
          for (int j = 0; j < 16; ++j) if (j < cnt) atomicAdd(dst + j, v[j]);
        } else if (cnt == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) if (j < cnt) dst[j] = v[j];
        }
      }
      __syncwarp();
    }
    tc::fence_before_sync();
  } else {
    {   // whole warp walks the loop, one elected lane issues (uniform operands: see tc::mma_ss_e)
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, DT_M, npad);
      const uint32_t idesc_c = tc::make_idesc(tc::FMT_BF16, DT_M, npad);
      uint32_t acc = 0;
      for (int c = 0; c < nch; ++c) {
        const int slot = c & 1;
        tc::mbar_wait_u(&full[slot], ((uint32_t)c / DT_STAGES) & 1u);
        tc::fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(smem + (size_t)slot * DT_SLOT);
        const uint32_t b_addr = a_addr + DT_OPERAND;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          tc::mma_ss_e<true>(tmem_u, tc::make_desc_sw128(a_addr + 32 * s), tc::make_desc_sw128(b_addr + 32 * s), idesc, acc);
          acc = 1;
          tc::mma_ss_e<false>(tmem_u, tc::make_desc_sw128(a_addr + DT_PLANE + 32 * s),
                            tc::make_desc_sw128(b_addr + DT_PLANE + 32 * s), idesc_c, 1);
        }
        tc::mma_commit_e(&empty[slot]);
      }
      tc::mma_commit_e(&acc_full);
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tc::tmem_dealloc(tmem_base, 128);
}

__global__ void __launch_bounds__(DT_THREADS, 1) dense_tc_kernel(const __grid_constant__ DtParams p) {
  dense_tc_body(p, (int)blockIdx.x, (int)blockIdx.y);
}

// up to DTB_MAX independent problems (the same-level backward-data GEMMs of the six head networks) in ONE launch
#define DTB_MAX 24
struct DtBatch { int count; int cta_start[DTB_MAX + 1]; int tiles_x[DTB_MAX]; DtParams p[DTB_MAX]; };
__global__ void __launch_bounds__(DT_THREADS, 1) dense_tc_batched_kernel(const __grid_constant__ DtBatch b) {
  int i = 0;
  while (i + 1 < b.count && (int)blockIdx.x >= b.cta_start[i + 1]) ++i;
  const int t = (int)blockIdx.x - b.cta_start[i];
  dense_tc_body(b.p[i], t % b.tiles_x[i], t / b.tiles_x[i]);
}

int dt_launch(vqn_ctx* ctx, const DtParams& p, cudaStream_t s) {
  static bool attr_set[16] = {false};
  const int dev = ctx->device & 15;
  if (!attr_set[dev]) {
    VQN_CUDA(cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DT_SMEM));
    attr_set[dev] = true;
  }
  dim3 grid((unsigned)((p.M + DT_M - 1) / DT_M), (unsigned)((p.N + DT_N - 1) / DT_N), 1);
  dense_tc_kernel<<<grid, DT_THREADS, DT_SMEM, s>>>(p);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// Weight gradients on tcgen05, batched:  dW[k_in, n_out] += X[rows, k_in]^T . dZ[rows, n_out],  db[n_out] += colsum(dZ)
// for up to WG_MAX layers in ONE launch.  The reduction runs over the ROWS, so both operands are read transposed: the
// producer thread that owns operand row m (a column of X, resp. of dZ) loads X[k0 + j][m] for the 32 rows of the chunk --
// for a fixed j the 32 threads of a warp read 128 contiguous bytes -- splits them and stores its row of the K-major slot,
// exactly as dense_tc_kernel does.  A CTA owns (layer, 128 x 128 output tile, row range); its accumulator is added into
// dW with fp32 atomics (as the warp-level kernel does); the B-role threads of the first row-tile also sum their column
// of dZ (the bias gradient).
// ---------------------------------------------------------------------------------------------
#define WG_MAX 32
struct WgProblem { const float* X; long long ldx; const float* dZ; long long lddz; float* dW; float* db; int M, N; long long rows; };
struct WgBatch { int count; int cta_start[WG_MAX + 1]; int tiles_m[WG_MAX], tiles_n[WG_MAX], per[WG_MAX]; WgProblem p[WG_MAX]; };

__global__ void __launch_bounds__(DT_THREADS, 1) dense_tc_wgrad_kernel(const __grid_constant__ WgBatch b) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[DT_STAGES], empty[DT_STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MMA_WARP = 16;
  int pi = 0;
  while (pi + 1 < b.count && (int)blockIdx.x >= b.cta_start[pi + 1]) ++pi;
  const WgProblem& p = b.p[pi];
  const int t = (int)blockIdx.x - b.cta_start[pi];
  const int mt = t % b.tiles_m[pi], nt = (t / b.tiles_m[pi]) % b.tiles_n[pi], sp = t / (b.tiles_m[pi] * b.tiles_n[pi]);
  const int m0 = mt * DT_M, n0 = nt * DT_N;
  const int nvalid = min(DT_N, p.N - n0), npad = (nvalid + 15) / 16 * 16;
  const long long k_begin = (long long)sp * b.per[pi];
  const long long k_end = min(p.rows, k_begin + (long long)b.per[pi]);
  const int nch = (int)((k_end - k_begin + 31) / 32);

  if (warp == MMA_WARP) tc::tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) {
    for (int i = 0; i < DT_STAGES; ++i) { tc::mbar_init(&full[i], 256); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(&acc_full, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 16) {
    const int grp = warp >> 3, i = tid & 255;
    const bool a_role = i < 128;
    const int r = a_role ? i : i - 128;                       // operand row: column m0 + r of X / column n0 + r of dZ
    const bool live = a_role ? (m0 + r < p.M) : (r < nvalid);
    const float* src = a_role ? p.X + m0 + r : p.dZ + n0 + r;
    const long long ld = a_role ? p.ldx : p.lddz;
    float csum = 0.f;
    for (int c = grp; c < nch; c += 2) {
      const int slot = c & 1;
      const long long k0 = k_begin + 32LL * c;
      float v[32];
      if (live) {
        if (k0 + 32 <= k_end) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __ldg(src + (k0 + j) * ld);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = k0 + j < k_end ? __ldg(src + (k0 + j) * ld) : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (!a_role) {
#pragma unroll
        for (int j = 0; j < 32; ++j) csum += v[j];
      }
      tc::mbar_wait(&empty[slot], (((uint32_t)c / DT_STAGES) & 1u) ^ 1u);
      const uint32_t op = tc::smem_u32(smem) + (uint32_t)slot * DT_SLOT + (a_role ? 0u : DT_OPERAND);
      dt_store16(op, r, 0, v, !a_role);
      dt_store16(op, r, 16, v + 16, !a_role);
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[slot]);
    }
    if (!a_role && live && mt == 0 && p.db && csum != 0.f) atomicAdd(p.db + n0 + r, csum);
    if (nch > 0) {
      // ---- epilogue: the accumulator tile is ADDED into dW (other row ranges add theirs) ----
      tc::mbar_wait(&acc_full, 0);
      tc::fence_after_sync();
      const int row = 32 * (warp & 3) + lane, m = m0 + row;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
      const int piece0 = warp >> 2;                             // 0..3
#pragma unroll 1
      for (int tt = 0; tt < 2; ++tt) {
        const int c16 = 16 * (piece0 + 4 * tt);
        if (c16 >= npad) continue;                              // warp-uniform
        float v[16];
        tc::tmem_ld16(lane_addr + (uint32_t)c16, v);
        if (m < p.M) {
          float* dst = p.dW + (long long)m * p.N + n0 + c16;
          const int cnt = min(16, nvalid - c16);
          if (cnt == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            // a thread owns a row of the tile: 4 vector reductions instead of 16 scalar ones (a quarter of the L1 wavefronts)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                           "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < cnt) atomicAdd(dst + j, v[j]);
          }
        }
        __syncwarp();
      }
      tc::fence_before_sync();
    }
  } else {
    {   // whole warp walks the loop, one elected lane issues (uniform operands: see tc::mma_ss_e)
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, DT_M, npad);
      const uint32_t idesc_c = tc::make_idesc(tc::FMT_BF16, DT_M, npad);
      uint32_t acc = 0;
      for (int c = 0; c < nch; ++c) {
        const int slot = c & 1;
        tc::mbar_wait_u(&full[slot], ((uint32_t)c / DT_STAGES) & 1u);
        tc::fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(smem + (size_t)slot * DT_SLOT);
        const uint32_t b_addr = a_addr + DT_OPERAND;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          tc::mma_ss_e<true>(tmem_u, tc::make_desc_sw128(a_addr + 32 * s), tc::make_desc_sw128(b_addr + 32 * s), idesc, acc);
          acc = 1;
          tc::mma_ss_e<false>(tmem_u, tc::make_desc_sw128(a_addr + DT_PLANE + 32 * s),
                            tc::make_desc_sw128(b_addr + DT_PLANE + 32 * s), idesc_c, 1);
        }
        tc::mma_commit_e(&empty[slot]);
      }
      if (nch > 0) tc::mma_commit_e(&acc_full);
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tc::tmem_dealloc(tmem_base, 128);
}

}  // namespace

// the backward-data GEMMs of one level of the head networks on tcgen05, one launch
int vqn_dense_tc_backward_data_batched(vqn_ctx* ctx, const vqn_dense_problem* pr, int count, cudaStream_t s) {
  static bool attr_set[16] = {false};
  const int dev = ctx->device & 15;
  if (!attr_set[dev]) {
    VQN_CUDA(cudaFuncSetAttribute(dense_tc_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DT_SMEM));
    attr_set[dev] = true;
  }
  for (int i0 = 0; i0 < count; i0 += DTB_MAX) {
    DtBatch b = {};
    int ctas = 0;
    for (int i = i0; i < count && i < i0 + DTB_MAX; ++i) {
      const vqn_dense_problem& q = pr[i];
      if (q.m == 0) continue;
      DtParams& p = b.p[b.count];
      p.A = q.a; p.lda = q.lda; p.W = q.w; p.ldw = q.n; p.C = q.out; p.ldc = q.ldo; p.M = (int)q.m; p.N = q.k; p.K = q.n; p.bwd = 1;
      p.yprev = q.yprev; p.ldy = q.ldy; p.act_prev = q.act_prev; p.accumulate = q.accumulate;
      b.tiles_x[b.count] = (int)((q.m + DT_M - 1) / DT_M);
      b.cta_start[b.count] = ctas;
      ctas += b.tiles_x[b.count] * ((q.k + DT_N - 1) / DT_N);
      ++b.count;
    }
    if (b.count == 0) continue;
    b.cta_start[b.count] = ctas;
    dense_tc_batched_kernel<<<ctas, DT_THREADS, DT_SMEM, s>>>(b);
    VQN_LAUNCHED(ctx);
  }
  return VQN_OK;
}

// every weight-gradient GEMM of a training step on tcgen05 (called by vqn_dense_backward_weights_batched)
int vqn_dense_tc_wgrad_batched(vqn_ctx* ctx, const vqn_dense_problem* pr, int count, cudaStream_t s) {
  static bool attr_set[16] = {false};
  const int dev = ctx->device & 15;
  if (!attr_set[dev]) {
    VQN_CUDA(cudaFuncSetAttribute(dense_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DT_SMEM));
    attr_set[dev] = true;
  }
  WgBatch b = {};
  long long base_tiles = 0;
  for (int i = 0; i < count; ++i)
    if (pr[i].m > 0) base_tiles += (long long)((pr[i].k + DT_M - 1) / DT_M) * ((pr[i].n + DT_N - 1) / DT_N);
  if (base_tiles == 0) return VQN_OK;
  // one CTA per SM (129 KB of shared memory): the row split fills whole waves of sm_count CTAs
  int splits_target = (int)((long long)ctx->sm_count / base_tiles);
  if (splits_target < 1) splits_target = 1;
  int ctas = 0;
  for (int i = 0; i < count; ++i) {
    const vqn_dense_problem& q = pr[i];
    if (q.m == 0) continue;
    WgProblem& p = b.p[b.count];
    p.X = q.a; p.ldx = q.lda; p.dZ = q.w; p.lddz = q.ldw; p.dW = q.out; p.db = q.colsum; p.M = q.k; p.N = q.n; p.rows = q.m;
    long long per = (q.m + splits_target - 1) / splits_target;
    per = (per + 31) / 32 * 32;
    if (per < 128) per = 128;
    const int splits = (int)((q.m + per - 1) / per);
    b.per[b.count] = (int)per;
    b.tiles_m[b.count] = (q.k + DT_M - 1) / DT_M; b.tiles_n[b.count] = (q.n + DT_N - 1) / DT_N;
    b.cta_start[b.count] = ctas;
    ctas += b.tiles_m[b.count] * b.tiles_n[b.count] * splits;
    ++b.count;
  }
  b.cta_start[b.count] = ctas;
  dense_tc_wgrad_kernel<<<ctas, DT_THREADS, DT_SMEM, s>>>(b);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

namespace {
}  // namespace

// rows from which the tcgen05 forward is taken for the WIDE layers (k, n >= 192); VQN_DENSE_TC_MIN_M overrides: 1 = every
// forward and backward-data call (parity tests, measurements), a negative value = never
long long vqn_dense_tc_min_m() {
  static bool init = false;
  static long long v = 1024;
  if (!init) { const char* e = getenv("VQN_DENSE_TC_MIN_M"); if (e) v = atoll(e); init = true; }
  return v;
}

int vqn_dense_tc_forward(vqn_ctx* ctx, const float* x, long long ldx, const float* w, const float* b, float* y,
                         long long ldy, long long m, int k, int n, int act, float out_scale, float out_bias,
                         cudaStream_t s) {
  DtParams p = {};
  p.A = x; p.lda = ldx; p.W = w; p.ldw = n; p.C = y; p.ldc = ldy; p.M = (int)m; p.N = n; p.K = k; p.bwd = 0;
  p.bias = b; p.act = act; p.out_scale = out_scale; p.out_bias = out_bias;
  return dt_launch(ctx, p, s);
}

int vqn_dense_tc_backward_data(vqn_ctx* ctx, const float* dz, long long lddz, const float* w, float* dx, long long lddx,
                               const float* yprev, long long ldyp, int act_prev, int accumulate, long long m, int k,
                               int n, cudaStream_t s) {
  DtParams p = {};
  p.A = dz; p.lda = lddz; p.W = w; p.ldw = n; p.C = dx; p.ldc = lddx; p.M = (int)m; p.N = k; p.K = n; p.bwd = 1;
  p.yprev = yprev; p.ldy = ldyp; p.act_prev = act_prev; p.accumulate = accumulate;
  return dt_launch(ctx, p, s);
}
