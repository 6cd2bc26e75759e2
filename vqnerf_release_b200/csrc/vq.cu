// VQ nearest-codeword assignment + EMA statistics (networks/vq_layers.py:257-344).
//
// Data layout in HBM: latents x[n,256] fp32 row-major (1 KB rows), codebook C[256,K] fp32 (column =
// codeword), indices int64[n].  Algorithmic bytes per latent: 4*256 in + 8 out = 1032 B.
//
// Kernel vq_assign_kernel<KC,R>: one warp handles R rows at a time.  Lane l owns the z-slice
// {4l..4l+3, 128+4l..128+4l+3} of every row, so a row is two fully coalesced 512 B LDG.128 requests
// (streamed, L1 no-allocate).  The matching codebook slice (8 z x KC codewords) stays in REGISTERS for
// the whole kernel when K <= KC (the shipped K = 15), so the inner loop is pure FFMA with no shared-memory
// traffic; larger K loops over KC-wide chunks staged in shared memory.  Partial dot products are summed
// across the 32 lanes with a halving butterfly (R*KC values -> 2 per lane, R*KC-2 shuffles instead of
// 5*R*KC), distances are formed in the reference's algebraic form  ||x||^2 - 2 x.c + ||c||^2
// (vq_layers.py:279-282) and the arg-min keeps the FIRST minimum (tf.argmax(-d), :292).  Rows whose best
// two candidates are closer than 4e-6 relative are re-scored in fp64 so that indices are exact whenever
// the true top-2 gap exceeds the 1e-6 tolerance of BASELINE.json.
#include "common.cuh"
#include "tc_common.cuh"

#define VQ_Z 256
#define VQ_THREADS 256

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

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

// butterfly step: R*KC values spread over lane bit `BIT`; after the step each lane holds half of them.
template <int NV>
__device__ __forceinline__ void halve(float* v, int lane, int bit) {
  const bool hi = (lane >> bit) & 1;
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) {
    float send = hi ? v[i] : v[i + NV / 2];
    float keep = hi ? v[i + NV / 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << bit);
  }
}

// MODE 0: full assignment; MODE 1: only the global max distance (first pass of the `thres` path)
template <int MODE, bool CHUNKED>
__global__ void __launch_bounds__(VQ_THREADS, 1) vq_assign_kernel(VqParams p) {
  constexpr int KC = 16, R = 2;
  extern __shared__ float smem[];
  // smem: cnorm[Kpad] | (stats) dw_s[K*256] + cnt_s[K] + elat_s | chunk staging [32 lanes][8 z][KC] (K > KC)
  const int K = p.K;
  const int nchunks = (K + KC - 1) / KC;
  const int Kpad = nchunks * KC;
  float* cnorm = smem;
  float* dw_s = cnorm + Kpad;
  const bool smem_dw = p.stats && p.want_dw && K <= 32;
  float* cnt_s = dw_s + (smem_dw ? K * VQ_Z : 0);
  float* elat_s = cnt_s + (p.stats ? Kpad : 0);
  float* stage = elat_s + (p.stats ? 4 : 0);   // only used when nchunks > 1

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ||c||^2 per codeword (vq_layers.py:282); padded codewords get +inf so they never win
  for (int k = tid; k < Kpad; k += VQ_THREADS) {
    float s = 0.f;
    if (k < K) {
      for (int z = 0; z < VQ_Z; ++z) { float c = p.cb[(size_t)z * K + k]; s = fmaf(c, c, s); }
    } else {
      s = __int_as_float(0x7f800000);
    }
    cnorm[k] = s;
  }
  if (p.stats) {
    for (int i = tid; i < (smem_dw ? K * VQ_Z : 0); i += VQ_THREADS) dw_s[i] = 0.f;
    for (int i = tid; i < Kpad; i += VQ_THREADS) cnt_s[i] = 0.f;
    if (tid < 4) elat_s[tid] = 0.f;
  }
  // register-resident codebook slice for chunk 0: creg[zi][k], zi -> z = (zi<4 ? 4*lane+zi : 128+4*lane+zi-4)
  float creg[CHUNKED ? 1 : 8][CHUNKED ? 1 : KC];
  if (!CHUNKED) {
#pragma unroll
    for (int zi = 0; zi < 8; ++zi) {
      int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
#pragma unroll
      for (int k = 0; k < KC; ++k) creg[CHUNKED ? 0 : zi][CHUNKED ? 0 : k] = k < K ? p.cb[(size_t)z * K + k] : 0.f;
    }
  }
  __syncthreads();

  const long long n_groups = (p.n + R - 1) / R;
  const long long warps_total = (long long)gridDim.x * (VQ_THREADS / 32);
  const long long iters = (n_groups + warps_total - 1) / warps_total;   // block-uniform trip count
  float local_max = -__int_as_float(0x7f800000);

  for (long long it = 0; it < iters; ++it) {
    const long long g = it * warps_total + (long long)blockIdx.x * (VQ_THREADS / 32) + warp;
    float x[R][8];
    float xs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      long long row = g * R + r;
      if (row < p.n) {
        const float* px = p.x + row * VQ_Z;
        float4 a = ldg_stream_f4(px + 4 * lane);
        float4 b = ldg_stream_f4(px + 128 + 4 * lane);
        x[r][0] = a.x; x[r][1] = a.y; x[r][2] = a.z; x[r][3] = a.w;
        x[r][4] = b.x; x[r][5] = b.y; x[r][6] = b.z; x[r][7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[r][i] = 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(x[r][i], x[r][i], s);
      s = warp_sum(s);
      if (p.normalize) {
        // caller's safe_l2_normalize(z_enc, axis=1) (vq_nfr.py:575) fused: x * rsqrt(max(sum x^2, 1e-6))
        float inv = rsqrtf(fmaxf(s, 1e-6f));
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[r][i] *= inv; s2 = fmaf(x[r][i], x[r][i], s2); }
        s = warp_sum(s2);
      }
      xs[r] = s;  // reduce_sum(flat_inputs**2, 1)
    }
    if (MODE == 0 && p.znorm_out) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        long long row = g * R + r;
        if (row < p.n) {
          float* po = p.znorm_out + row * VQ_Z;
          *reinterpret_cast<float4*>(po + 4 * lane) = make_float4(x[r][0], x[r][1], x[r][2], x[r][3]);
          *reinterpret_cast<float4*>(po + 128 + 4 * lane) = make_float4(x[r][4], x[r][5], x[r][6], x[r][7]);
        }
      }
    }

    // running best / second best per row, replicated over the 8-lane group that owns the row after the
    // butterfly: group id = (lane >> 3) & (R-1)... with R == 2 rows split on lane bit 4.
    float best = __int_as_float(0x7f800000), second = __int_as_float(0x7f800000);
    int best_i = 0, second_i = 0;
    const int my_r = (lane >> 4) & 1;

    for (int c = 0; c < nchunks; ++c) {
      float acc[R * KC];
#pragma unroll
      for (int i = 0; i < R * KC; ++i) acc[i] = 0.f;
      if (!CHUNKED) {
#pragma unroll
        for (int zi = 0; zi < 8; ++zi)
#pragma unroll
          for (int k = 0; k < KC; ++k)
#pragma unroll
            for (int r = 0; r < R; ++r)
              acc[r * KC + k] = fmaf(x[r][zi], creg[CHUNKED ? 0 : zi][CHUNKED ? 0 : k], acc[r * KC + k]);
      } else {
        // stage chunk c block-wide (trip counts are block-uniform, so the barriers are safe)
        __syncthreads();
        for (int i = tid; i < 32 * 8 * KC; i += VQ_THREADS) {
          int l = i / (8 * KC), rem = i % (8 * KC), zi = rem / KC, k = rem % KC;
          int z = (zi < 4 ? 0 : 128) + 4 * l + (zi & 3);
          int kk = c * KC + k;
          stage[l * (8 * KC + 4) + zi * KC + k] = kk < K ? p.cb[(size_t)z * K + kk] : 0.f;
        }
        __syncthreads();
        const float* st = stage + lane * (8 * KC + 4);
#pragma unroll
        for (int zi = 0; zi < 8; ++zi) {
          float cv[KC];
#pragma unroll
          for (int q = 0; q < KC / 4; ++q) {
            float4 t = *reinterpret_cast<const float4*>(st + zi * KC + 4 * q);
            cv[4 * q] = t.x; cv[4 * q + 1] = t.y; cv[4 * q + 2] = t.z; cv[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int k = 0; k < KC; ++k)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r * KC + k] = fmaf(x[r][zi], cv[k], acc[r * KC + k]);
        }
      }
      // butterfly: bit4 splits rows, bits 3..0 split codewords 16 -> 1 ... we stop at 2 values per lane
      // after using bits 4,3,2,1 and finish bit 0 with a plain exchange so each lane pair shares 2 values.
      halve<32>(acc, lane, 4);          // 32 -> 16 values: row my_r, k 0..15
      halve<16>(acc, lane, 3);          // 8 values
      halve<8>(acc, lane, 2);           // 4 values
      halve<4>(acc, lane, 1);           // 2 values
      halve<2>(acc, lane, 0);           // 1 value
      // lane now owns codeword k = 8*b3 + 4*b2 + 2*b1 + b0 of row my_r in this chunk
      const int k_local = ((lane >> 3) & 1) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + (lane & 1);
      const int k_glob = c * KC + k_local;
      float xsr = my_r ? xs[1] : xs[0];
      float d = xsr - 2.0f * acc[0] + cnorm[k_glob];   // +inf for padded codewords
      if (MODE == 1) {
        if (k_glob < K && g * R + my_r < p.n) local_max = fmaxf(local_max, d);   // padded rows must not vote
        continue;
      }
      if (p.sel_mask && k_glob < K) {
        float sel = p.sel_mask[k_glob];
        float mv = ord2f(*p.maxdist);
        d = d * sel + mv * (1.0f - sel);               // vq_layers.py:290
      }
      if (p.dist_out && k_glob < K) {
        long long row = g * R + my_r;
        if (row < p.n) p.dist_out[row * K + k_glob] = d;
      }
      // arg-min over the 16 lanes of this row (first minimum wins), tracking the runner-up
      float b = d, s = __int_as_float(0x7f800000);
      int bi = k_glob, si = k_glob;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, b, o);
        int obi = __shfl_xor_sync(0xffffffffu, bi, o);
        float os = __shfl_xor_sync(0xffffffffu, s, o);
        int osi = __shfl_xor_sync(0xffffffffu, si, o);
        bool other_wins = (ob < b) || (ob == b && obi < bi);
        float lose = other_wins ? b : ob;
        int losei = other_wins ? bi : obi;
        if (other_wins) { b = ob; bi = obi; }
        // runner-up = min(lose, s, os)
        if (os < s || (os == s && osi < si)) { s = os; si = osi; }
        if (lose < s || (lose == s && losei < si)) { s = lose; si = losei; }
      }
      // merge with the running best across chunks (earlier chunk = lower index wins ties)
      if (b < best) {
        if (best < s) { second = best; second_i = best_i; } else { second = s; second_i = si; }
        best = b; best_i = bi;
      } else {
        if (b < second) { second = b; second_i = bi; }
      }
    }
    if (MODE == 1) continue;

    // fp64 re-score of near-ties (top-2 gap below 4e-6 relative): exact ordering of the two candidates
    {
      bool near = (second - best) <= 4e-6f * fmaxf(fabsf(best), 1e-3f) && K > 1 &&
                  !(p.sel_mask && (p.sel_mask[best_i] == 0.f || p.sel_mask[second_i] == 0.f));
      unsigned need = __ballot_sync(0xffffffffu, near);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!((need >> (16 * r)) & 1u)) continue;       // warp-uniform
        int i1 = __shfl_sync(0xffffffffu, best_i, 16 * r);
        int i2 = __shfl_sync(0xffffffffu, second_i, 16 * r);
        double d1 = 0.0, d2 = 0.0;
#pragma unroll
        for (int zi = 0; zi < 8; ++zi) {
          int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
          double xv = (double)x[r][zi];
          double c1 = (double)p.cb[(size_t)z * K + i1], c2 = (double)p.cb[(size_t)z * K + i2];
          d1 += c1 * c1 - 2.0 * xv * c1;
          d2 += c2 * c2 - 2.0 * xv * c2;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          d1 += __shfl_xor_sync(0xffffffffu, d1, o);
          d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        bool swap = (d2 < d1) || (d2 == d1 && i2 < i1);
        if (my_r == r && swap) { best_i = i2; }
      }
    }

    // outputs
#pragma unroll
    for (int r = 0; r < R; ++r) {
      long long row = g * R + r;
      if (row >= p.n) continue;                          // warp-uniform
      int idx = __shfl_sync(0xffffffffu, best_i, 16 * r);
      if (lane == 0 && p.idx_out) p.idx_out[row] = (long long)idx;
      if (p.quant_out || p.stats) {
        float q[8];
        float e = 0.f;
#pragma unroll
        for (int zi = 0; zi < 8; ++zi) {
          int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
          float c = p.cb[(size_t)z * K + idx];           // quantize(): embedding_lookup(codebook^T, idx)
          float diff = c - x[r][zi];
          e = fmaf(diff, diff, e);                        // (sg(quantized) - inputs)^2, :302
          q[zi] = x[r][zi] + diff;                        // inputs + sg(quantized - inputs), :327
        }
        if (p.quant_out) {
          float* po = p.quant_out + row * VQ_Z;
          *reinterpret_cast<float4*>(po + 4 * lane) = make_float4(q[0], q[1], q[2], q[3]);
          *reinterpret_cast<float4*>(po + 128 + 4 * lane) = make_float4(q[4], q[5], q[6], q[7]);
        }
        if (p.stats) {
          e = warp_sum(e);
          if (lane == 0) { atomicAdd(&cnt_s[idx], 1.0f); atomicAdd(&elat_s[0], e); atomicAdd(&elat_s[1], 1.0f); }
          if (p.want_dw) {
#pragma unroll
            for (int zi = 0; zi < 8; ++zi) {
              int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
              if (smem_dw) atomicAdd(&dw_s[idx * VQ_Z + z], x[r][zi]);
              else atomicAdd(&p.stats[K + 2 + (size_t)z * K + idx], (double)x[r][zi]);
            }
          }
        }
      }
    }
  }

  if (MODE == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if (lane == 0) atomicMax(p.maxdist, f2ord(local_max));
    return;
  }
  if (p.stats) {
    __syncthreads();
    for (int k = tid; k < K; k += VQ_THREADS)
      if (cnt_s[k] != 0.f) atomicAdd(&p.stats[k], (double)cnt_s[k]);
    if (tid == 0) { atomicAdd(&p.stats[K], (double)elat_s[0]); atomicAdd(&p.stats[K + 1], (double)elat_s[1]); }
    if (smem_dw)
      for (int i = tid; i < K * VQ_Z; i += VQ_THREADS) {
        int k = i / VQ_Z, z = i % VQ_Z;
        float v = dw_s[i];
        if (v != 0.f) atomicAdd(&p.stats[K + 2 + (size_t)z * K + k], (double)v);
      }
  }
}


// ---------------------------------------------------------------------------------------------
// K <= 16 (the shipped K = 15): HBM-roofline kernel.
//
// Each warp owns a private ring of VQS_DEPTH x 2 KB shared-memory stages that one elected lane fills with
// cp.async.bulk (TMA engine, mbarrier complete_tx): 2 consecutive latent rows = one contiguous 2 KB copy, so
// 8 warps x 8 stages = 128 KB of reads are in flight per SM with no registers spent on prefetch.  Lane l
// owns z in {4l..4l+3, 128+4l..128+4l+3}; its 8 x 16 codebook slice lives in REGISTERS for the whole kernel.
// Issue-slot economy (the kernel is HBM-bound only if it needs < ~180 issue slots per row):
//  * packed fma.rn.f32x2 (FFMA2): one issue slot per two MACs (accumulator pairs = adjacent codewords),
//  * SEL-free butterfly: lane l keeps codeword (s ^ (l & 15)) in accumulator slot s and row (r ^ bit4(l)) in row
//    slot r, so every halving step is  v[i] += shfl_xor(v[i + half])  with no lane-dependent selects,
//  * arg-min by redux.sync.min on the order-preserving uint image of the distance + ballot/ffs (first index
//    wins), twice (best and runner-up) instead of a 4-round 4-register shuffle tournament.
// ---------------------------------------------------------------------------------------------
#define VQS_DEPTH 8
#define VQS_WARPS 8

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b,
                                                    unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int NV>
__device__ __forceinline__ void halve_nosel(float* v, int bit) {
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i + NV / 2], 1 << bit);
}

template <int MODE>
__global__ void __launch_bounds__(VQS_WARPS * 32, 1) vq_assign_small_kernel(VqParams p) {
  constexpr int KC = 16;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // smem: ring [warps][DEPTH][2 KB] | mbarriers [warps][DEPTH] | cnorm[16] | cnt_s[16] | elat_s[4] | dw_s[K*256]
  float* ring = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + VQS_WARPS * VQS_DEPTH * 2048);
  float* cnorm = reinterpret_cast<float*>(bars + VQS_WARPS * VQS_DEPTH);
  float* cnt_s = cnorm + KC;
  float* elat_s = cnt_s + KC;
  float* dw_s = elat_s + 4;
  const int K = p.K;
  const bool smem_dw = p.stats && p.want_dw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b4 = lane >> 4, kperm = lane & 15;

  // ||c||^2 (vq_layers.py:282): warp w reduces codewords w and w+8; padded codewords get +inf
  for (int k = warp; k < KC; k += VQS_WARPS) {
    float s = 0.f;
    if (k < K)
      for (int z = lane; z < VQ_Z; z += 32) { float c = p.cb[(size_t)z * K + k]; s = fmaf(c, c, s); }
    s = warp_sum(s);
    if (lane == 0) cnorm[k] = k < K ? s : __int_as_float(0x7f800000);
  }
  if (p.stats) {
    for (int i = tid; i < (smem_dw ? K * VQ_Z : 0); i += blockDim.x) dw_s[i] = 0.f;
    if (tid < KC) cnt_s[tid] = 0.f;
    if (tid < 4) elat_s[tid] = 0.f;
  }
  // register-resident codebook slice, permuted: slot s of this lane holds codeword s ^ (lane & 15);
  // packed as pairs of adjacent slots (2q, 2q+1) -> codewords ((2q) ^ kperm, (2q+1) ^ kperm)
  unsigned long long creg[8][KC / 2];
#pragma unroll
  for (int zi = 0; zi < 8; ++zi) {
    const int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
#pragma unroll
    for (int q = 0; q < KC / 2; ++q) {
      const int k0 = (2 * q) ^ kperm, k1 = (2 * q + 1) ^ kperm;
      float c0 = k0 < K ? p.cb[(size_t)z * K + k0] : 0.f;
      float c1 = k1 < K ? p.cb[(size_t)z * K + k1] : 0.f;
      creg[zi][q] = pack2(c0, c1);
    }
  }
  uint64_t* mybar = bars + warp * VQS_DEPTH;
  float* myring = ring + (size_t)warp * VQS_DEPTH * 512;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < VQS_DEPTH; ++s) tc::mbar_init(mybar + s, 1);
    tc::mbar_fence_init();
  }
  __syncthreads();

  const long long n_groups = (p.n + 1) / 2;
  const long long warps_total = (long long)gridDim.x * VQS_WARPS;
  const long long gw = (long long)blockIdx.x * VQS_WARPS + warp;
  // groups of this warp: g = it * warps_total + gw, it = 0 .. my_iters-1
  const int my_iters = n_groups > gw ? (int)((n_groups - gw + warps_total - 1) / warps_total) : 0;
  auto issue = [&](int it) {
    const long long g = (long long)it * warps_total + gw;
    const int s = it % VQS_DEPTH;
    const unsigned bytes = (g * 2 + 1 < p.n) ? 2048u : 1024u;
    tc::mbar_expect_tx(mybar + s, bytes);
    tc::bulk_g2s(myring + s * 512, p.x + g * 2 * VQ_Z, bytes, mybar + s);
  };
  if (lane == 0)
    for (int it = 0; it < my_iters && it < VQS_DEPTH; ++it) issue(it);

  float local_max = -__int_as_float(0x7f800000);
  const float cn_mine = cnorm[kperm];
  float sel_mine = 1.f;
  if (MODE == 0 && p.sel_mask) sel_mine = kperm < K ? p.sel_mask[kperm] : 1.f;

  for (int it = 0; it < my_iters; ++it) {
    const long long g = (long long)it * warps_total + gw;
    const int s = it % VQS_DEPTH;
    tc::mbar_wait(mybar + s, (unsigned)((it / VQS_DEPTH) & 1));
    // row slot r of this lane holds row g*2 + (r ^ b4)
    float x[2][8];
    const bool full = g * 2 + 1 < p.n;                       // warp-uniform
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float* px = myring + s * 512 + ((r ^ b4) * VQ_Z);
      float4 a = *reinterpret_cast<const float4*>(px + 4 * lane);
      float4 b = *reinterpret_cast<const float4*>(px + 128 + 4 * lane);
      x[r][0] = a.x; x[r][1] = a.y; x[r][2] = a.z; x[r][3] = a.w;
      x[r][4] = b.x; x[r][5] = b.y; x[r][6] = b.z; x[r][7] = b.w;
    }
    if (!full) {                                             // odd tail: the second row is stale smem
#pragma unroll
      for (int r = 0; r < 2; ++r)
        if ((r ^ b4) == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[r][i] = 0.f;
        }
    }
    __syncwarp();
    if (lane == 0 && it + VQS_DEPTH < my_iters) issue(it + VQS_DEPTH);

    // row norms (and the caller's l2_normalize, vq_nfr.py:575, when fused); xs[r] belongs to row r ^ b4
    float xs[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sq = fmaf(x[r][i], x[r][i], sq);
      xs[r] = sq;
    }
    // slot r of this lane holds row r ^ b4: fold the partner half-warp's other slot in first (1 shuffle), then
    // reduce over the 16 lanes of the half; afterwards xs_mine = ||row g*2 + b4||^2 and xs_other the other row's
    float xs_mine, xs_other;
    {
      float a = xs[0] + __shfl_xor_sync(0xffffffffu, xs[1], 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      xs_mine = a;
    }
    if (p.normalize) {
      xs_other = __shfl_xor_sync(0xffffffffu, xs_mine, 16);
      const float inv0 = rsqrtf(fmaxf(xs_mine, 1e-6f)), inv1 = rsqrtf(fmaxf(xs_other, 1e-6f));
      float sq0 = 0.f, sq1 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[0][i] *= inv0; sq0 = fmaf(x[0][i], x[0][i], sq0);
        x[1][i] *= inv1; sq1 = fmaf(x[1][i], x[1][i], sq1);
      }
      float a = sq0 + __shfl_xor_sync(0xffffffffu, sq1, 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      xs_mine = a;
    }
    if (MODE == 0 && p.znorm_out) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        long long row = g * 2 + (r ^ b4);
        if (row < p.n) {
          float* po = p.znorm_out + row * VQ_Z;
          *reinterpret_cast<float4*>(po + 4 * lane) = make_float4(x[r][0], x[r][1], x[r][2], x[r][3]);
          *reinterpret_cast<float4*>(po + 128 + 4 * lane) = make_float4(x[r][4], x[r][5], x[r][6], x[r][7]);
        }
      }
    }

    // partial dots: acc2[r][q] = (slot 2q, slot 2q+1) of row slot r
    unsigned long long acc2[2][KC / 2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      unsigned long long xx = pack2(x[r][0], x[r][0]);
#pragma unroll
      for (int q = 0; q < KC / 2; ++q) acc2[r][q] = fmul2(xx, creg[0][q]);
    }
#pragma unroll
    for (int zi = 1; zi < 8; ++zi)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        unsigned long long xx = pack2(x[r][zi], x[r][zi]);
#pragma unroll
        for (int q = 0; q < KC / 2; ++q) acc2[r][q] = ffma2(xx, creg[zi][q], acc2[r][q]);
      }
    float acc[2 * KC];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int q = 0; q < KC / 2; ++q) unpack2(acc2[r][q], acc[r * KC + 2 * q], acc[r * KC + 2 * q + 1]);
    halve_nosel<32>(acc, 4);   // row slots: keep slot 0 = row b4
    halve_nosel<16>(acc, 3);
    halve_nosel<8>(acc, 2);
    halve_nosel<4>(acc, 1);
    halve_nosel<2>(acc, 0);
    // lane now holds x.c for row (g*2 + b4), codeword kperm = lane & 15 (slot 0 ^ kperm)
    const long long my_row = g * 2 + b4;
    const bool row_ok = my_row < p.n;
    float d = xs_mine - 2.0f * acc[0] + cn_mine;               // +inf for padded codewords
    if (MODE == 1) {
      if (kperm < K && row_ok) local_max = fmaxf(local_max, d);
      continue;
    }
    if (p.sel_mask && kperm < K) {
      float mv = ord2f(*p.maxdist);
      d = d * sel_mine + mv * (1.0f - sel_mine);             // vq_layers.py:290
    }
    if (p.dist_out && kperm < K && row_ok) p.dist_out[my_row * K + kperm] = d;

    // arg-min per half-warp: first minimum wins (tf.argmax(-d), vq_layers.py:292), plus the runner-up
    const unsigned key = f2ord(d);
    const unsigned inf_key = 0xffffffffu;
    unsigned m0 = __reduce_min_sync(0xffffffffu, b4 ? inf_key : key);
    unsigned m1 = __reduce_min_sync(0xffffffffu, b4 ? key : inf_key);
    const unsigned mn = b4 ? m1 : m0;
    unsigned eq = __ballot_sync(0xffffffffu, key == mn);
    int best_i = __ffs((eq >> (16 * b4)) & 0xffffu) - 1;
    const unsigned key2 = (kperm == best_i) ? inf_key : key;
    unsigned n0 = __reduce_min_sync(0xffffffffu, b4 ? inf_key : key2);
    unsigned n1 = __reduce_min_sync(0xffffffffu, b4 ? key2 : inf_key);
    const unsigned mn2 = b4 ? n1 : n0;
    unsigned eq2 = __ballot_sync(0xffffffffu, key2 == mn2 && kperm != best_i);
    int second_i = __ffs((eq2 >> (16 * b4)) & 0xffffu) - 1;
    if (second_i < 0) second_i = best_i;
    const float best = ord2f(mn), second = ord2f(mn2);

    // fp64 re-score of near-ties (top-2 gap below 4e-6 relative): exact ordering of the two candidates
    {
      bool near = (second - best) <= 4e-6f * fmaxf(fabsf(best), 1e-3f) && K > 1 && row_ok &&
                  !(p.sel_mask && (p.sel_mask[best_i] == 0.f || p.sel_mask[second_i] == 0.f));
      unsigned need = __ballot_sync(0xffffffffu, near);
      if (need) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {                     // rr = actual row within the pair
          if (!((need >> (16 * rr)) & 1u)) continue;         // warp-uniform
          int i1 = __shfl_sync(0xffffffffu, best_i, 16 * rr);
          int i2 = __shfl_sync(0xffffffffu, second_i, 16 * rr);
          double d1 = 0.0, d2 = 0.0;
#pragma unroll
          for (int zi = 0; zi < 8; ++zi) {
            int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
            double xv = (double)(b4 == rr ? x[0][zi] : x[1][zi]);   // slot holding row rr on this lane
            double c1 = (double)p.cb[(size_t)z * K + i1], c2 = (double)p.cb[(size_t)z * K + i2];
            d1 += c1 * c1 - 2.0 * xv * c1;
            d2 += c2 * c2 - 2.0 * xv * c2;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
          }
          bool swap = (d2 < d1) || (d2 == d1 && i2 < i1);
          if (b4 == rr && swap) best_i = i2;
        }
      }
    }

    // outputs
    if (p.idx_out && kperm == 0 && row_ok) p.idx_out[my_row] = (long long)best_i;
    if (p.quant_out || p.stats) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        long long row = g * 2 + rr;
        if (row >= p.n) continue;                            // warp-uniform
        const int idx = __shfl_sync(0xffffffffu, best_i, 16 * rr);
        float q[8], xv[8];
        float e = 0.f;
#pragma unroll
        for (int zi = 0; zi < 8; ++zi) {
          int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
          xv[zi] = b4 == rr ? x[0][zi] : x[1][zi];
          float c = p.cb[(size_t)z * K + idx];               // quantize(): embedding_lookup(codebook^T, idx)
          float diff = c - xv[zi];
          e = fmaf(diff, diff, e);                            // (sg(quantized) - inputs)^2, :302
          q[zi] = xv[zi] + diff;                              // inputs + sg(quantized - inputs), :327
        }
        if (p.quant_out) {
          float* po = p.quant_out + row * VQ_Z;
          *reinterpret_cast<float4*>(po + 4 * lane) = make_float4(q[0], q[1], q[2], q[3]);
          *reinterpret_cast<float4*>(po + 128 + 4 * lane) = make_float4(q[4], q[5], q[6], q[7]);
        }
        if (p.stats) {
          e = warp_sum(e);
          if (lane == 0) { atomicAdd(&cnt_s[idx], 1.0f); atomicAdd(&elat_s[0], e); atomicAdd(&elat_s[1], 1.0f); }
          if (p.want_dw) {
#pragma unroll
            for (int zi = 0; zi < 8; ++zi) {
              int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
              atomicAdd(&dw_s[idx * VQ_Z + z], xv[zi]);
            }
          }
        }
      }
    }
  }

  if (MODE == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if (lane == 0) atomicMax(p.maxdist, f2ord(local_max));
    return;
  }
  if (p.stats) {
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x)
      if (cnt_s[k] != 0.f) atomicAdd(&p.stats[k], (double)cnt_s[k]);
    if (tid == 0) { atomicAdd(&p.stats[K], (double)elat_s[0]); atomicAdd(&p.stats[K + 1], (double)elat_s[1]); }
    if (smem_dw)
      for (int i = tid; i < K * VQ_Z; i += blockDim.x) {
        int k = i / VQ_Z, z = i % VQ_Z;
        float v = dw_s[i];
        if (v != 0.f) atomicAdd(&p.stats[K + 2 + (size_t)z * K + k], (double)v);
      }
  }
}

extern "C" int vqn_vq_assign(vqn_ctx* ctx, const float* inputs, int64_t n, int z_dim, const float* codebook,
                             int k, const float* sel_mask, int normalize_inputs, int64_t* indices,
                             float* quantize, float* distances, float* z_norm_out, double* stats,
                             int want_dw, vqn_stream stream) {
  VQN_CHECK_ARG(ctx, "vq_assign: null ctx");
  VQN_CHECK_ARG(z_dim == VQ_Z, "vq_assign: embedding_dim must be 256 (conv_width)");
  VQN_CHECK_ARG(k >= 1 && k <= 1024, "vq_assign: 1 <= K <= 1024");
  VQN_CHECK_ARG(n >= 0, "vq_assign: n < 0");
  if (n == 0) return VQN_OK;
  VQN_CHECK_ARG(inputs && codebook, "vq_assign: null input");
  VqParams p;
  p.x = inputs; p.n = n; p.cb = codebook; p.K = k; p.sel_mask = sel_mask; p.maxdist = nullptr;
  p.normalize = normalize_inputs; p.idx_out = (long long*)indices; p.quant_out = quantize;
  p.dist_out = distances; p.znorm_out = z_norm_out; p.stats = stats; p.want_dw = want_dw;
  cudaStream_t s = vqn_cs(stream);
  const int KC = 16;
  if (k <= KC) {
    size_t smem_s = (size_t)VQS_WARPS * VQS_DEPTH * 2048 + VQS_WARPS * VQS_DEPTH * 8 +
                    sizeof(float) * (KC + KC + 4 + ((stats && want_dw) ? (size_t)k * VQ_Z : 0));
    long long groups_s = (n + 1) / 2;
    long long want_s = (groups_s + VQS_WARPS - 1) / VQS_WARPS;
    int blocks_s = (int)(want_s < (long long)ctx->sm_count ? want_s : (long long)ctx->sm_count);
    VQN_CUDA(cudaFuncSetAttribute(vq_assign_small_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    if (sel_mask) {
      unsigned* md = reinterpret_cast<unsigned*>(ctx->scratch);      // persistent scratch slot 0
      VQN_CUDA(cudaMemsetAsync(md, 0, sizeof(unsigned), s));         // ordered 0 == most negative
      p.maxdist = md;
      VQN_CUDA(cudaFuncSetAttribute(vq_assign_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
      vq_assign_small_kernel<1><<<blocks_s, VQS_WARPS * 32, smem_s, s>>>(p);
      VQN_LAUNCHED(ctx);
    }
    vq_assign_small_kernel<0><<<blocks_s, VQS_WARPS * 32, smem_s, s>>>(p);
    VQN_LAUNCHED(ctx);
    return VQN_OK;
  }
  int nchunks = (k + KC - 1) / KC, kpad = nchunks * KC;
  size_t smem = sizeof(float) * (kpad + ((stats && want_dw && k <= 32) ? (size_t)k * VQ_Z : 0) +
                                 (stats ? kpad + 4 : 0) + (nchunks > 1 ? 32 * (8 * KC + 4) : 0));
  long long groups = (n + 1) / 2;
  long long want_blocks = (groups + (VQ_THREADS / 32) - 1) / (VQ_THREADS / 32);
  int blocks = (int)(want_blocks < (long long)ctx->sm_count ? want_blocks : (long long)ctx->sm_count);
  // occupancy: registers allow one 256-thread block per SM; use 2 x SM count when rows are plentiful so the
  // second block of a pair can start as soon as registers free up
  const bool chunked = nchunks > 1;
  auto k0 = chunked ? vq_assign_kernel<0, true> : vq_assign_kernel<0, false>;
  auto k1 = chunked ? vq_assign_kernel<1, true> : vq_assign_kernel<1, false>;
  VQN_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VQN_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned* maxdist = nullptr;
  if (sel_mask) {
    maxdist = reinterpret_cast<unsigned*>(ctx->scratch);          // persistent scratch slot 0
    VQN_CUDA(cudaMemsetAsync(maxdist, 0, sizeof(unsigned), s));   // ordered 0 == most negative
    p.maxdist = maxdist;
    k1<<<blocks, VQ_THREADS, smem, s>>>(p);
    VQN_LAUNCHED(ctx);
  }
  k0<<<blocks, VQ_THREADS, smem, s>>>(p);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// EMA update: vq_layers.py:304-321 + sonnet ExponentialMovingAverage; one block, K*256 elements.
// ---------------------------------------------------------------------------------------------
__global__ void vq_ema_kernel(const double* __restrict__ stats, int Z, int K, const float* __restrict__ cb,
                              float decay, float eps, float commit, int training, float* cs_hidden,
                              float* cs_average, float* dw_hidden, float* dw_average, long long* counters,
                              float* update, float* loss, float* perplexity) {
  extern __shared__ float sh[];       // cs_norm[K]
  __shared__ float red[32];
  __shared__ float nsum_s, perp_s;
  const int tid = threadIdx.x;
  const double rows = stats[K + 1];
  // perplexity (:328-330) and loss (:302,321)
  float part = 0.f;
  for (int k = tid; k < K; k += blockDim.x) {
    float pk = (float)(stats[k] / (rows > 0 ? rows : 1.0));
    part += pk * logf(pk + 1e-10f);
  }
  part = warp_sum(part);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid < 32) {
    float t = tid < (blockDim.x >> 5) ? red[tid] : 0.f;
    t = warp_sum(t);
    if (tid == 0) perp_s = expf(-t);
  }
  __syncthreads();
  if (tid == 0) {
    if (perplexity) *perplexity = perp_s;
    if (loss) *loss = commit * (float)(stats[K] / (rows > 0 ? rows * Z : 1.0));
  }
  if (!training) return;
  // Sonnet EMA on cluster sizes
  long long c_cs = counters[0] + 1, c_dw = counters[1] + 1;
  float deb_cs = 1.0f - powf(decay, (float)c_cs);
  float deb_dw = 1.0f - powf(decay, (float)c_dw);
  float part_n = 0.f;
  for (int k = tid; k < K; k += blockDim.x) {
    float v = (float)stats[k];
    float h = cs_hidden[k];
    h = h - (h - v) * (1.0f - decay);
    cs_hidden[k] = h;
    float a = h / deb_cs;
    cs_average[k] = a;
    sh[k] = a;
    part_n += a;
  }
  part_n = warp_sum(part_n);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = part_n;
  __syncthreads();
  if (tid < 32) {
    float t = tid < (blockDim.x >> 5) ? red[tid] : 0.f;
    t = warp_sum(t);
    if (tid == 0) nsum_s = t;
  }
  __syncthreads();
  const float nsum = nsum_s;
  for (int k = tid; k < K; k += blockDim.x)
    sh[k] = (sh[k] + eps) / (nsum + K * eps) * nsum;             // :311-313
  __syncthreads();
  for (int i = tid; i < Z * K; i += blockDim.x) {
    int k = i % K;
    float v = (float)stats[K + 2 + i];
    float h = dw_hidden[i];
    h = h - (h - v) * (1.0f - decay);
    dw_hidden[i] = h;
    float a = h / deb_dw;
    dw_average[i] = a;
    float w = a / sh[k];                                         // :315-316
    float used = stats[k] > 0.0 ? 1.0f : 0.0f;                   // :318
    update[i] = w * used + cb[i] * (1.0f - used);                // :319
  }
  __syncthreads();
  if (tid == 0) { counters[0] = c_cs; counters[1] = c_dw; }
}

extern "C" int vqn_vq_ema_update(vqn_ctx* ctx, const double* stats, int z_dim, int k, const float* codebook,
                                 float decay, float epsilon, float commitment_cost, int is_training,
                                 float* cs_hidden, float* cs_average, float* dw_hidden, float* dw_average,
                                 int64_t* counters, float* update, float* loss, float* perplexity,
                                 vqn_stream stream) {
  VQN_CHECK_ARG(ctx && stats && z_dim > 0 && k > 0 && k <= 4096, "vq_ema_update args");
  if (is_training)
    VQN_CHECK_ARG(codebook && cs_hidden && cs_average && dw_hidden && dw_average && counters && update,
                  "vq_ema_update: training needs EMA state + codebook + update");
  VQN_CHECK_ARG(decay >= 0.f && decay <= 1.f, "decay must be in range [0, 1]");   // vq_layers.py:236-237
  vq_ema_kernel<<<1, 1024, sizeof(float) * k, vqn_cs(stream)>>>(
      stats, z_dim, k, codebook, decay, epsilon, commitment_cost, is_training, cs_hidden, cs_average,
      dw_hidden, dw_average, (long long*)counters, update, loss, perplexity);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
