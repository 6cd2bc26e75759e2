// Dense-layer kernels of the training step (config #4): Keras Dense forward with saved activations,
// backward-data and backward-weights (networks/mlp.py:24-50 under tf.GradientTape, train_nfr.py:562-576).
//
// One GEMM core, C[M,N] = A[M,K] . B[K,N], 3-term split on the warp-level tensor cores (fp32 accumulate;
// hi = cvt.rna.tf32(v), lo = v - hi, C = Ah.Bh [mma.sync m16n8k8 tf32] + (Al.Bh + Ah.Bl) [ONE m16n8k16 bf16 MMA])
// so the gradients keep fp32 parity (1e-4 rel) with the reference.  The training batch is small (8192 rays per
// GPU, widths <= 384), so the layers are launch- and latency-bound rather than tensor-bound: the kernels are
// built for generality (arbitrary leading dimensions so layers read and write slices of the concat buffers
// of the skip connections, transposed operands read in place) and are meant to be replayed from a CUDA
// graph.  Block tile 64x64x16, 4 warps (2x2) of 32x32, operands staged in padded shared memory whose
// strides (20 for k-contiguous, 72 for m/n-contiguous tiles) make every fragment read conflict-free.
//
//   forward   (mlp.py:44-46)   Y  = act(X . W + b)                 A = X,    B = W
//   bwd data                   dX = (dZ . W^T) * act'(Y_prev)      A = dZ,   B = W^T (read in place)
//   bwd weight                 dW += X^T . dZ  (split over rows, fp32 atomics),  db += colsum(dZ) (fused)
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 128;
constexpr int LDK = BK + 4;    // k-contiguous tile  [rows][k]   stride 20 floats
constexpr int LDR = BM + 8;    // row-contiguous tile [k][rows]  stride 72 floats

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo with hi = rna_tf32(x) (returned as tf32 bits) and lo = x - hi (exact, returned as a float)
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, float& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = x - __uint_as_float(hi);
}
// {low half: bf16(a), high half: bf16(b)}
__device__ __forceinline__ unsigned pack_bf16(float a, float b) {
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// The two correction products a_lo.b_hi + a_hi.b_lo are 2^-11 of the leading term, so bf16 operands (2^-9) keep
// them to 2^-20: ONE m16n8k16 bf16 MMA evaluates both, with its 16 k-slots holding [lo(k=t), lo(k=t+4)] pairs in
// slots 0-7 and [hi(k=t), hi(k=t+4)] pairs in slots 8-15 (any k-permutation shared by A and B is valid).
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct GemmParams {
  const float* A; long long lda;   // A_TRANS ? element (m,k) at A[k*lda + m] : A[m*lda + k]
  const float* B; long long ldb;   // B_TRANS ? element (k,n) at B[n*ldb + k] : B[k*ldb + n]
  float* C; long long ldc;
  int M, N, K;
  int k_split;                      // rows of K per blockIdx.z (EPI_ATOMIC only)
  // epilogues
  const float* bias; int act;                      // EPI_FWD
  const float* yprev; long long ldy; int act_prev; // EPI_BWD: multiply by act'(yprev[m*ldy + n])
  int accumulate;                                  // EPI_BWD: C += result
  float out_scale, out_bias;                       // EPI_FWD: y = out_scale * act(.) + out_bias (albedo_slope/bias)
  float* colsum;                                   // EPI_ATOMIC: colsum[n] += sum_k B[k][n] (the Dense bias gradient)
};

enum { EPI_FWD = 0, EPI_BWD = 1, EPI_ATOMIC = 2 };

// load a 4-wide strip of a [rows x cols] operand tile; `contig` runs along the contiguous global axis
__device__ __forceinline__ float4 load4(const float* base, long long ld, int outer, int inner, int outer_lim,
                                        int inner_lim) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (outer >= outer_lim) return v;
  const float* p = base + (long long)outer * ld + inner;
  if (inner + 3 < inner_lim && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    v = *reinterpret_cast<const float4*>(p);
  } else {
    if (inner < inner_lim) v.x = p[0];
    if (inner + 1 < inner_lim) v.y = p[1];
    if (inner + 2 < inner_lim) v.z = p[2];
    if (inner + 3 < inner_lim) v.w = p[3];
  }
  return v;
}

template <bool A_TRANS, bool B_TRANS, int EPI>
__device__ __forceinline__ void gemm_tile(const GemmParams& p, const int bx, const int by, const int bz) {
  // A tile: !A_TRANS -> As[m][k] (stride LDK), A_TRANS -> As[k][m] (stride LDR); same for B with n
  __shared__ __align__(16) float As[A_TRANS ? BK * LDR : BM * LDK];
  __shared__ __align__(16) float Bs[B_TRANS ? BN * LDK : BK * LDR];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int m0 = bx * BM, n0 = by * BN;
  int k_begin = 0, k_end = p.K;
  if (EPI == EPI_ATOMIC) {
    k_begin = bz * p.k_split;
    k_end = min(p.K, k_begin + p.k_split);
  }
  float acc_m[2][4][4], acc_c[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc_m[i][j][e] = 0.f; acc_c[i][j][e] = 0.f; }

  // global -> register staging: each operand tile is 64 x 16 = 256 float4 strips, 2 per thread
  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int s = tid + i * THREADS;
      if (!A_TRANS) {           // strips along k: row = s / 4, kq = s % 4
        const int r = s >> 2, kq = (s & 3) * 4;
        ra[i] = load4(p.A, p.lda, m0 + r, k0 + kq, p.M, k_end);
      } else {                  // strips along m: k = s / 16, mq = s % 16
        const int kk = s >> 4, mq = (s & 15) * 4;
        ra[i] = load4(p.A, p.lda, k0 + kk, m0 + mq, k_end, p.M);
      }
      if (B_TRANS) {            // strips along k: col = s / 4
        const int c = s >> 2, kq = (s & 3) * 4;
        rb[i] = load4(p.B, p.ldb, n0 + c, k0 + kq, p.N, k_end);
      } else {                  // strips along n
        const int kk = s >> 4, nq = (s & 15) * 4;
        rb[i] = load4(p.B, p.ldb, k0 + kk, n0 + nq, k_end, p.N);
      }
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int s = tid + i * THREADS;
      if (!A_TRANS) *reinterpret_cast<float4*>(&As[(s >> 2) * LDK + (s & 3) * 4]) = ra[i];
      else *reinterpret_cast<float4*>(&As[(s >> 4) * LDR + (s & 15) * 4]) = ra[i];
      if (B_TRANS) *reinterpret_cast<float4*>(&Bs[(s >> 2) * LDK + (s & 3) * 4]) = rb[i];
      else *reinterpret_cast<float4*>(&Bs[(s >> 4) * LDR + (s & 15) * 4]) = rb[i];
    }
  };
  auto a_at = [&](int m, int k) { return A_TRANS ? As[k * LDR + m] : As[m * LDK + k]; };
  auto b_at = [&](int k, int n) { return B_TRANS ? Bs[n * LDK + k] : Bs[k * LDR + n]; };

  // bias gradient fused into the weight-gradient GEMM: the CTAs of the first row-tile also sum the columns of
  // every dZ tile they stage (thread t < 64 owns column n0 + t)
  const bool do_colsum = EPI == EPI_ATOMIC && p.colsum != nullptr && bx == 0 && tid < BN;
  float csum = 0.f;
  if (k_begin < k_end) gload(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    __syncthreads();
    sstore();
    __syncthreads();
    if (k0 + BK < k_end) gload(k0 + BK);
    if (EPI == EPI_ATOMIC && do_colsum) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) csum += Bs[B_TRANS ? tid * LDK + kk : kk * LDR + tid];
    }
#pragma unroll
    for (int ks = 0; ks < BK; ks += 8) {
      unsigned ah[2][4], ac[2][4];          // tf32 hi fragment, bf16 correction fragment
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = wm + i * 16 + g;
        float l0, l1, l2, l3;
        split_tf32(a_at(m, ks + t), ah[i][0], l0);
        split_tf32(a_at(m + 8, ks + t), ah[i][1], l1);
        split_tf32(a_at(m, ks + t + 4), ah[i][2], l2);
        split_tf32(a_at(m + 8, ks + t + 4), ah[i][3], l3);
        ac[i][0] = pack_bf16(l0, l2);                                                   // row g:   lo(k=t), lo(k=t+4)
        ac[i][1] = pack_bf16(l1, l3);                                                   // row g+8
        ac[i][2] = pack_bf16(__uint_as_float(ah[i][0]), __uint_as_float(ah[i][2]));      // row g:   hi(k=t), hi(k=t+4)
        ac[i][3] = pack_bf16(__uint_as_float(ah[i][1]), __uint_as_float(ah[i][3]));      // row g+8
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = wn + j * 8 + g;
        unsigned bh0, bh1;
        float bl0, bl1;
        split_tf32(b_at(ks + t, n), bh0, bl0);
        split_tf32(b_at(ks + t + 4, n), bh1, bl1);
        const unsigned bc0 = pack_bf16(__uint_as_float(bh0), __uint_as_float(bh1));     // pairs with the lo slots of A
        const unsigned bc1 = pack_bf16(bl0, bl1);                                       // pairs with the hi slots of A
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mma_tf32(acc_m[i][j], ah[i], bh0, bh1);
          mma_bf16(acc_c[i][j], ac[i], bc0, bc1);
        }
      }
    }
  }

  if (EPI == EPI_ATOMIC && do_colsum && n0 + tid < p.N) atomicAdd(&p.colsum[n0 + tid], csum);

  // epilogue: accumulator (i, j, e): row = wm + 16i + g + 8*(e>>1), col = wn + 8j + 2t + (e&1)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = m0 + wm + i * 16 + g + 8 * (e >> 1);
        const int n = n0 + wn + j * 8 + 2 * t + (e & 1);
        if (m >= p.M || n >= p.N) continue;
        float v = acc_m[i][j][e] + acc_c[i][j][e];
        float* dst = p.C + (long long)m * p.ldc + n;
        if (EPI == EPI_FWD) {
          if (p.bias) v += p.bias[n];
          v = vqn_apply_act(v, p.act);
          *dst = p.out_scale * v + p.out_bias;
        } else if (EPI == EPI_BWD) {
          if (p.act_prev != VQN_ACT_NONE) {
            const float y = p.yprev[(long long)m * p.ldy + n];
            v *= (p.act_prev == VQN_ACT_RELU) ? (y > 0.f ? 1.f : 0.f) : y * (1.f - y);
          }
          if (p.accumulate == 2) atomicAdd(dst, v);        // several problems of a batch add into the same buffer
          else { if (p.accumulate) v += *dst; *dst = v; }
        } else {
          atomicAdd(dst, v);
        }
      }
}

template <bool A_TRANS, bool B_TRANS, int EPI>
__global__ void __launch_bounds__(THREADS) dense_gemm_kernel(GemmParams p) {
  gemm_tile<A_TRANS, B_TRANS, EPI>(p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Batched form: up to GB_MAX independent GEMMs of one kind in ONE launch (the same-level layers of the six head networks,
// or every weight-gradient GEMM of the step).  The training batch is 8192 rows, a single layer is 20 us of latency on a
// fraction of the SMs, and a step had ~57 of them in a row; a CTA finds its problem in the prefix table of tile counts.
constexpr int GB_MAX = 32;
struct GemmBatch {
  int count;
  int tile_start[GB_MAX + 1];         // first linear tile of problem i; tile_start[count] = grid size
  int tiles_x[GB_MAX], tiles_y[GB_MAX];
  GemmParams p[GB_MAX];
};

template <bool A_TRANS, bool B_TRANS, int EPI>
__global__ void __launch_bounds__(THREADS) dense_gemm_batched_kernel(const __grid_constant__ GemmBatch b) {
  int i = 0;
  while (i + 1 < b.count && (int)blockIdx.x >= b.tile_start[i + 1]) ++i;
  const int t = (int)blockIdx.x - b.tile_start[i];
  const int bx = t % b.tiles_x[i], by = (t / b.tiles_x[i]) % b.tiles_y[i], bz = t / (b.tiles_x[i] * b.tiles_y[i]);
  gemm_tile<A_TRANS, B_TRANS, EPI>(b.p[i], bx, by, bz);
}

// dZ = dY * act'(Y) for the LAST layer of a net (no following GEMM epilogue to fuse into);
// scale folds d(out_scale * act + out_bias) (albedo_slope)
__global__ void act_grad_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ y,
                                long long ldy, int M, int N, int act, float scale, float out_scale,
                                float out_bias, float* __restrict__ dz, long long lddz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float v = dy[(long long)m * lddy + n] * scale;
  if (act != VQN_ACT_NONE) {
    // stored y = out_scale * a + out_bias  ->  a = (y - out_bias) / out_scale
    float a = (y[(long long)m * ldy + n] - out_bias) / out_scale;
    v *= (act == VQN_ACT_RELU) ? (a > 0.f ? 1.f : 0.f) : a * (1.f - a);
  }
  dz[(long long)m * lddz + n] = v;
}

// the last-layer activation gradients of several networks (the six heads of a training step) in ONE launch: blockIdx.y = job
constexpr int ACT_JOBS_MAX = 8;
struct ActJobs { vqn_act_job j[ACT_JOBS_MAX]; };
__global__ void act_grad_batched_kernel(const __grid_constant__ ActJobs jobs) {
  const vqn_act_job& q = jobs.j[blockIdx.y];
  const long long total = (long long)q.m * q.n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / q.n;
    const int n = (int)(i % q.n);
    float v = q.dy[m * q.lddy + n] * q.scale;
    if (q.act != VQN_ACT_NONE) {
      const float a = (q.y[m * q.ldy + n] - q.out_bias) / q.out_scale;
      v *= (q.act == VQN_ACT_RELU) ? (a > 0.f ? 1.f : 0.f) : a * (1.f - a);
    }
    q.dz[m * q.lddz + n] = v;
  }
}

}  // namespace

// train_tc.cu: the same two contracts on tcgen05/TMEM (taken from vqn_dense_tc_min_m() rows upwards)
long long vqn_dense_tc_min_m();
int vqn_dense_tc_forward(vqn_ctx* ctx, const float* x, long long ldx, const float* w, const float* b, float* y,
                         long long ldy, long long m, int k, int n, int act, float out_scale, float out_bias,
                         cudaStream_t s);
int vqn_dense_tc_backward_data(vqn_ctx* ctx, const float* dz, long long lddz, const float* w, float* dx, long long lddx,
                               const float* yprev, long long ldyp, int act_prev, int accumulate, long long m, int k,
                               int n, cudaStream_t s);

/* mlp.py:44-46: Y[M,N] (ld ldy) = out_scale * act(X[M,K] (ld ldx) . W[K,N] + b[N]) + out_bias */
extern "C" int vqn_dense_forward(vqn_ctx* ctx, const float* x, int64_t ldx, const float* w, const float* b, float* y,
                                 int64_t ldy, int64_t m, int k, int n, int act, float out_scale, float out_bias,
                                 vqn_stream stream) {
  VQN_CHECK_ARG(ctx && x && w && y, "dense_forward: null pointer");
  VQN_CHECK_ARG(m >= 0 && k > 0 && n > 0 && ldx >= k && ldy >= n, "dense_forward: bad shape");
  if (m == 0) return VQN_OK;
  // measured per layer at 8192 rows (benchmarks/dense_micro.py, CUDA-graph replay): the tcgen05 kernel wins only on the
  // wide layers (k = n = 256: 18 us vs 29 us); its fixed costs (132 KB smem carve-out, TMEM allocation, un-coalesced
  // row-per-thread operand loads) lose on the narrow ones (k = 63: 12 us vs 8 us).  VQN_DENSE_TC_MIN_M=1 forces it.
  const long long tc_min = vqn_dense_tc_min_m();
  if (tc_min >= 0 && m >= tc_min && (tc_min <= 1 || (k >= 192 && n >= 192)))
    return vqn_dense_tc_forward(ctx, x, ldx, w, b, y, ldy, m, k, n, act, out_scale, out_bias, vqn_cs(stream));
  GemmParams p = {};
  p.A = x; p.lda = ldx; p.B = w; p.ldb = n; p.C = y; p.ldc = ldy; p.M = (int)m; p.N = n; p.K = k;
  p.bias = b; p.act = act; p.out_scale = out_scale; p.out_bias = out_bias;
  dim3 grid((unsigned)((m + BM - 1) / BM), (unsigned)((n + BN - 1) / BN), 1);
  dense_gemm_kernel<false, false, EPI_FWD><<<grid, THREADS, 0, vqn_cs(stream)>>>(p);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

/* dX[M,K] (ld lddx) (+)= (dZ[M,N] (ld lddz) . W[K,N]^T) * act_prev'(Yprev[M,K] (ld ldyp)) */
extern "C" int vqn_dense_backward_data(vqn_ctx* ctx, const float* dz, int64_t lddz, const float* w, float* dx,
                                       int64_t lddx, const float* yprev, int64_t ldyp, int act_prev, int accumulate,
                                       int64_t m, int k, int n, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && dz && w && dx, "dense_backward_data: null pointer");
  VQN_CHECK_ARG(m >= 0 && k > 0 && n > 0 && lddz >= n && lddx >= k, "dense_backward_data: bad shape");
  VQN_CHECK_ARG(act_prev == VQN_ACT_NONE || yprev, "dense_backward_data: act_prev needs yprev");
  if (m == 0) return VQN_OK;
  if (vqn_dense_tc_min_m() == 1)      // never faster than the warp-level kernel at the training sizes (27 vs 22-32 us)
    return vqn_dense_tc_backward_data(ctx, dz, lddz, w, dx, lddx, yprev, ldyp, act_prev, accumulate, m, k, n,
                                      vqn_cs(stream));
  GemmParams p = {};
  p.A = dz; p.lda = lddz; p.B = w; p.ldb = n; p.C = dx; p.ldc = lddx; p.M = (int)m; p.N = k; p.K = n;
  p.yprev = yprev; p.ldy = ldyp; p.act_prev = act_prev; p.accumulate = accumulate;
  dim3 grid((unsigned)((m + BM - 1) / BM), (unsigned)((k + BN - 1) / BN), 1);
  dense_gemm_kernel<false, true, EPI_BWD><<<grid, THREADS, 0, vqn_cs(stream)>>>(p);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

/* dW[K,N] += X[M,K]^T . dZ[M,N];  db[N] += colsum(dZ)  (db may be NULL) */
extern "C" int vqn_dense_backward_weights(vqn_ctx* ctx, const float* x, int64_t ldx, const float* dz, int64_t lddz,
                                          float* dw, float* db, int64_t m, int k, int n, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && x && dz && dw, "dense_backward_weights: null pointer");
  VQN_CHECK_ARG(m >= 0 && k > 0 && n > 0 && ldx >= k && lddz >= n, "dense_backward_weights: bad shape");
  if (m == 0) return VQN_OK;
  GemmParams p = {};
  p.A = x; p.lda = ldx; p.B = dz; p.ldb = lddz; p.C = dw; p.ldc = n; p.M = k; p.N = n; p.K = (int)m;
  const int tiles = ((k + BM - 1) / BM) * ((n + BN - 1) / BN);
  int splits = (2 * ctx->sm_count + tiles - 1) / tiles;           // ~2 waves of CTAs
  int per = (int)((m + splits - 1) / splits);
  per = ((per + BK - 1) / BK) * BK;
  if (per < 4 * BK) per = 4 * BK;
  splits = (int)((m + per - 1) / per);
  p.k_split = per;
  p.colsum = db;
  dim3 grid((unsigned)((k + BM - 1) / BM), (unsigned)((n + BN - 1) / BN), (unsigned)splits);
  dense_gemm_kernel<true, false, EPI_ATOMIC><<<grid, THREADS, 0, vqn_cs(stream)>>>(p);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

/* dZ[M,N] = scale * dY * act'(Y): gradient through the last layer's activation (and albedo_slope) */
extern "C" int vqn_act_backward(vqn_ctx* ctx, const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t m,
                                int n, int act, float scale, float out_scale, float out_bias, float* dz, int64_t lddz,
                                vqn_stream stream) {
  VQN_CHECK_ARG(ctx && dy && dz && (act == VQN_ACT_NONE || y), "act_backward: null pointer");
  VQN_CHECK_ARG(out_scale != 0.f, "act_backward: out_scale == 0");
  if (m == 0) return VQN_OK;
  const long long total = (long long)m * n;
  act_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, vqn_cs(stream)>>>(dy, lddy, y, ldy, (int)m, n, act, scale,
                                                                              out_scale, out_bias, dz, lddz);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

/* vqn_act_backward for up to 8 networks in ONE launch */
extern "C" int vqn_act_backward_batched(vqn_ctx* ctx, const vqn_act_job* jobs, int count, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && jobs && count >= 0 && count <= ACT_JOBS_MAX, "act_backward_batched: 0..8 jobs");
  if (count == 0) return VQN_OK;
  ActJobs a;
  long long most = 0;
  for (int i = 0; i < count; ++i) {
    const vqn_act_job& q = jobs[i];
    VQN_CHECK_ARG(q.dy && q.dz && (q.act == VQN_ACT_NONE || q.y) && q.m >= 0 && q.n > 0, "act_backward_batched: null pointer");
    VQN_CHECK_ARG(q.out_scale != 0.f, "act_backward_batched: out_scale == 0");
    a.j[i] = q;
    most = std::max(most, (long long)q.m * q.n);
  }
  if (most == 0) return VQN_OK;
  dim3 grid((unsigned)std::min<long long>((most + 255) / 256, 1024), (unsigned)count);
  act_grad_batched_kernel<<<grid, 256, 0, vqn_cs(stream)>>>(a);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

/* Batched forms of vqn_dense_backward_data / vqn_dense_backward_weights: `count` <= 32 independent problems in ONE launch.
 * backward-data problems with accumulate == 2 ADD atomically (several problems may target the same dx). */
int vqn_dense_tc_backward_data_batched(vqn_ctx* ctx, const vqn_dense_problem* pr, int count, cudaStream_t s);   // train_tc.cu

extern "C" int vqn_dense_backward_data_batched(vqn_ctx* ctx, const vqn_dense_problem* pr, int count, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && pr && count >= 1 && count <= GB_MAX, "dense_backward_data_batched: 1 <= count <= 32");
  {
    // tcgen05 form (train_tc.cu): opt-in (VQN_BWD_TC=1 from 1024 rows upwards, 2 always -- the parity tests).  Measured at
    // 8192 rows: 1.315 ms per step against 1.272 ms with the warp-level kernel (one CTA per SM, ~8 waves of CTAs that each
    // pay TMEM allocation + barrier set-up for 1 - 8 chunks of work; routing only the long reductions to it: 1.265 vs 1.269)
    static int env = -1;
    if (env < 0) { const char* e = getenv("VQN_BWD_TC"); env = e ? atoi(e) : 0; }
    long long rows_min = 1LL << 62;
    for (int i = 0; i < count; ++i) {
      const vqn_dense_problem& q = pr[i];
      VQN_CHECK_ARG(q.a && q.w && q.out && q.m >= 0 && q.k > 0 && q.n > 0 && q.lda >= q.n && q.ldo >= q.k,
                    "dense_backward_data_batched: bad problem");
      VQN_CHECK_ARG(q.act_prev == VQN_ACT_NONE || q.yprev, "dense_backward_data_batched: act_prev needs yprev");
      if (q.m > 0 && q.m < rows_min) rows_min = q.m;
    }
    if (env && (rows_min >= 1024 || env == 2) && rows_min < (1LL << 62))
      return vqn_dense_tc_backward_data_batched(ctx, pr, count, vqn_cs(stream));
  }
  GemmBatch b = {};
  int tiles = 0;
  for (int i = 0; i < count; ++i) {
    const vqn_dense_problem& q = pr[i];
    VQN_CHECK_ARG(q.a && q.w && q.out && q.m >= 0 && q.k > 0 && q.n > 0 && q.lda >= q.n && q.ldo >= q.k,
                  "dense_backward_data_batched: bad problem");
    VQN_CHECK_ARG(q.act_prev == VQN_ACT_NONE || q.yprev, "dense_backward_data_batched: act_prev needs yprev");
    GemmParams& p = b.p[b.count];
    if (q.m == 0) continue;
    p.A = q.a; p.lda = q.lda; p.B = q.w; p.ldb = q.n; p.C = q.out; p.ldc = q.ldo; p.M = (int)q.m; p.N = q.k; p.K = q.n;
    p.yprev = q.yprev; p.ldy = q.ldy; p.act_prev = q.act_prev; p.accumulate = q.accumulate;
    b.tiles_x[b.count] = (int)((q.m + BM - 1) / BM); b.tiles_y[b.count] = (q.k + BN - 1) / BN;
    b.tile_start[b.count] = tiles;
    tiles += b.tiles_x[b.count] * b.tiles_y[b.count];
    ++b.count;
  }
  if (b.count == 0) return VQN_OK;
  b.tile_start[b.count] = tiles;
  dense_gemm_batched_kernel<false, true, EPI_BWD><<<tiles, THREADS, 0, vqn_cs(stream)>>>(b);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

int vqn_dense_tc_wgrad_batched(vqn_ctx* ctx, const vqn_dense_problem* pr, int count, cudaStream_t s);   // train_tc.cu

extern "C" int vqn_dense_backward_weights_batched(vqn_ctx* ctx, const vqn_dense_problem* pr, int count, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && pr && count >= 1 && count <= GB_MAX, "dense_backward_weights_batched: 1 <= count <= 32");
  {
    // tcgen05 form (train_tc.cu) from 1024 rows upwards; VQN_WGRAD_TC=0 keeps the warp-level kernel
    static int env = -1;
    if (env < 0) { const char* e = getenv("VQN_WGRAD_TC"); env = e ? atoi(e) : 1; }
    long long rows_min = 1LL << 62;
    for (int i = 0; i < count; ++i) {
      const vqn_dense_problem& q = pr[i];
      VQN_CHECK_ARG(q.a && q.w && q.out && q.m >= 0 && q.k > 0 && q.n > 0 && q.lda >= q.k && q.ldw >= q.n,
                    "dense_backward_weights_batched: bad problem");
      if (q.m > 0 && q.m < rows_min) rows_min = q.m;
    }
    if (env && (rows_min >= 1024 || env == 2) && rows_min < (1LL << 62))       // (2: any row count -- the parity tests)
      return vqn_dense_tc_wgrad_batched(ctx, pr, count, vqn_cs(stream));
  }
  GemmBatch b = {};
  // row splits: ~4 waves of CTAs over the whole batch, shared out in proportion to the tiles of each problem
  long long base_tiles = 0;
  for (int i = 0; i < count; ++i) base_tiles += (long long)((pr[i].k + BM - 1) / BM) * ((pr[i].n + BN - 1) / BN);
  if (base_tiles == 0) return VQN_OK;
  // every CTA of the batch does the same work (`per` rows of one 64 x 64 tile), so the launch costs whole waves: pick the
  // split that fills ONE wave of resident CTAs (672 CTAs on 592 slots were two waves: 384 us; 588 on 592: one)
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dense_gemm_batched_kernel<true, false, EPI_ATOMIC>, THREADS, 0) !=
            cudaSuccess || occ < 1)
      occ = 4;
    ctas_per_sm = occ;
  }
  const long long slots = (long long)ctx->sm_count * ctas_per_sm;
  int splits_target = (int)(slots / base_tiles);
  if (splits_target < 1) splits_target = 1;
  int tiles = 0;
  for (int i = 0; i < count; ++i) {
    const vqn_dense_problem& q = pr[i];
    VQN_CHECK_ARG(q.a && q.w && q.out && q.m >= 0 && q.k > 0 && q.n > 0 && q.lda >= q.k && q.ldw >= q.n,
                  "dense_backward_weights_batched: bad problem");
    if (q.m == 0) continue;
    GemmParams& p = b.p[b.count];
    // a = x [m, k] (ld lda), w = dz [m, n] (ld ldw), out = dW [k, n], colsum = db
    p.A = q.a; p.lda = q.lda; p.B = q.w; p.ldb = q.ldw; p.C = q.out; p.ldc = q.n; p.M = q.k; p.N = q.n; p.K = (int)q.m;
    int per = (int)((q.m + splits_target - 1) / splits_target);
    per = ((per + BK - 1) / BK) * BK;
    if (per < 4 * BK) per = 4 * BK;
    const int splits = (int)((q.m + per - 1) / per);
    p.k_split = per; p.colsum = q.colsum;
    b.tiles_x[b.count] = (q.k + BM - 1) / BM; b.tiles_y[b.count] = (q.n + BN - 1) / BN;
    b.tile_start[b.count] = tiles;
    tiles += b.tiles_x[b.count] * b.tiles_y[b.count] * splits;
    ++b.count;
  }
  if (b.count == 0) return VQN_OK;
  b.tile_start[b.count] = tiles;
  dense_gemm_batched_kernel<true, false, EPI_ATOMIC><<<tiles, THREADS, 0, vqn_cs(stream)>>>(b);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
