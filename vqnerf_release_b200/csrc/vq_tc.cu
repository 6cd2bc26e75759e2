// VQ nearest-codeword assignment for codebooks of K > 32 on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Reference: networks/vq_layers.py:279-292 (distances ||x||^2 - 2 x.C + ||c||^2, first-minimum arg-min).
// BASELINE.json configs[2] sweeps K up to 1024, where the assignment is a [n,256] x [256,K] GEMM with an arg-min
// epilogue and no longer HBM-bound (2 K FLOP against 4 B per latent value).  The warp-level kernel of vq_mma.cu
// (operands from registers, right for the HBM-bound K <= 64) tops out at ~46 TFLOP/s there; this kernel keeps the
// 3-term fp32-parity split of mlp_tc.cu (kind::tf32 leading term + ONE kind::f16 bf16 MMA of twice the K for the two
// correction terms) and the same warp roles:
//
//   * a persistent CTA per SM walks tiles of 128 latents; a "layer" is (tile, block of <= 256 codewords);
//   * 16 producer warps (2 groups) load the tile's K-chunks (128 rows x 32 values) COALESCED, split them into the
//     tf32 plane + bf16 correction plane of the 128-B-swizzled A slot, and accumulate ||x||^2 on the way;
//   * one thread streams the pre-split codebook chunk images (cp.async.bulk + mbarrier), one thread issues the MMAs;
//   * accumulators ping-pong between two 256-column TMEM regions: while the tensor cores fill one region with the
//     dot products of the next codeword block, the producers drain the other one -- c2 - 2 x.c per codeword, running
//     (best, runner-up) per row in registers, first-index tie-break;
//   * after a tile's last block the four partial (best, runner-up) pairs of a row are merged through shared memory;
//     rows whose two best distances are closer than 4e-5 relative are re-scored in fp64 (warp-cooperative), so the
//     index is exact whenever the true top-2 gap exceeds the 1e-6 tolerance of BASELINE.json.
//
// K <= 128 (one codeword block, vq_tc_kernel): the latent chunks come through the TMA engine -- a tensor map with the
// 128-byte swizzle, 4..7 [128 x 32] boxes in flight per SM -- and four dedicated drain warps (a thread owns a row) take the
// arg-min off the producers.  K > 128: vq_tc_big_kernel (producers drain, four partial results per row merged in smem).
// Measured (4 M latents): K = 64 1.36 ms (warp-level kernel: 2.13 ms, it pads K = 33..64 to 8 n-tiles), K = 128 1.64 ms
// (register loads, producers draining: 2.49 ms), K = 32 1.32 ms (warp-level kernel: 1.24 ms -> routing starts at K = 33).
//
// Indices only (no thres mask, no l2-normalise, no statistics): every other variant stays on vq_mma.cu.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"
#include "vq.cuh"

#define VT_M 128
#define VT_G 2                               // producer groups of 8 warps
#define VT_THREADS_BIG (32 * (8 * VT_G + 3))  // K > 128 kernel: + MMA warp, codebook-stream warp, (idle) TMA warp
#define VT_THREADS (32 * (8 * VT_G + 7))      // + MMA warp, codebook-stream warp, latent-TMA warp, 4 drain warps
#define VT_A_PLANE (VT_M * 128u)             // 16 KB
#define VT_A_SLOT (2u * VT_A_PLANE)          // tf32 plane + bf16 correction plane
#define VT_W_SLOT (256u * 128u * 2u)         // 64 KB
#define VT_STAGES 2
#define VT_KCHUNKS (VQ_Z / 32)               // 8 K-chunks of 32 values
#define VT_MAXK 1024
#define VT_SMEM (VT_STAGES * VT_A_SLOT + VT_STAGES * VT_W_SLOT + 1024)
// vq_tc_kernel (K <= 128, one codeword block): the codebook chunk is <= 32 KB, so half of the W ring holds a 4-stage staging
// ring that the TMA engine fills with [128 rows x 128 B] boxes of the latents (tensor map, SWIZZLE_128B)
#define VT_W_SLOT_SMALL (128u * 128u * 2u)   // 32 KB
#define VT_XSTAGES 4
#define VT_XSTAGES_MAX 7
#define VT_WSTAGES_MAX 8                     // barrier array size of the codebook ring (two slots are used)
#define VT_X_STAGE (VT_M * 128u)             // 16 KB

namespace {

struct VtParams {
  const float* x;          // [n,256]
  long long n;
  const float* cb;         // [256,K] fp32 (for the fp64 re-score)
  int K, nb;               // nb = ceil(K / 256) codeword blocks
  const uint8_t* wpack;    // block b at wpack + b * (VT_KCHUNKS * VT_W_SLOT): chunk images [Npad_b x 128 B] x 2 planes
  const float* c2;         // [nb * 256] ||c||^2, +huge for the padding columns
  long long* idx_out;
};

// Running (best, runner-up, third) of a row as PACKED keys: the distance with its low `ibits` mantissa bits replaced by
// the codeword index (ibits = log2 of the padded codeword count, <= 10).  The arg-min scan is then five FMNMX per
// codeword and no index bookkeeping (round 1 tracked two indices with predicated selects: 8 ALU-pipe operations per
// codeword, and the accumulator drain is what bounds K >= 128); the truncation (2^(ibits-23) relative) is far above the
// arithmetic error, so it is covered by the same exactness net: every near-tie is re-scored in fp64 below.
struct Top3 { float b, s, t; };

__device__ __forceinline__ void top3_push(Top3& m, float dp) {
  m.t = fminf(m.t, fmaxf(dp, m.s));
  m.s = fminf(m.s, fmaxf(dp, m.b));
  m.b = fminf(m.b, dp);
}
// running result across codeword blocks: the keys keep their BLOCK-LOCAL 8-bit index (so the truncation stays at
// 2^-15 for any K); the blocks of the best and of the runner-up are tracked here, once per block instead of per codeword
struct Run3 { float b, s, t; int bb, sb; };
__device__ __forceinline__ void run3_push(Run3& m, float k, int blk) {
  const bool p1 = k < m.b, p2 = k < m.s;
  m.t = fminf(m.t, fmaxf(k, m.s));
  m.s = fminf(m.s, fmaxf(k, m.b));
  m.sb = p1 ? m.bb : (p2 ? blk : m.sb);
  m.b = fminf(m.b, k);
  m.bb = p1 ? blk : m.bb;
}
__device__ __forceinline__ float vt_pack(float d, uint32_t kmask, int idx) {
  return __int_as_float((int)((__float_as_uint(d) & kmask) + (uint32_t)idx));
}

// Tolerances of the exactness net.  The tensor-core value of c^2 - 2 x.c carries the 3-term split error (~2^-20 of |x||c|)
// and the accumulator's truncation (48 accumulation steps of < 1 ulp each; measured on unit vectors: a common bias of
// ~3.5e-6 and a spread of +-5e-7), all proportional to SCALE = ||x||^2 + ||c||^2 >= 2 |x.c|; on top of that the packed
// keys are truncated by up to TRUNC = 2^(ibits-23) |d| each:
//   runner-up within VT_TOL2 * SCALE + 2 TRUNC of the best  -> the two candidates are re-scored in fp64;
//   THIRD-best within VT_TOL3 * SCALE + 2 TRUNC of the best -> three-way near-tie, the approximate values cannot even name
//   the candidates (round 2 sweep: 1 row of 4 M at K = 128 had the true winner ranked third): the row is re-scored in
//   fp64 against EVERY codeword.  Exact duplicates land in one of the two and resolve to the first index.
#define VT_TOL2 1.1e-5f
#define VT_TOL3 4.0e-6f

// fp64 distances (minus the row constant ||x||^2) of one latent row to every codeword, warp-cooperative: lane owns the
// codewords k = lane + 32 i, four at a time for ILP; returns the first arg-min (networks/vq_layers.py:279-292 in
// exact arithmetic)
__device__ __forceinline__ int vt_full_rescore(const float* __restrict__ xrow, const float* __restrict__ cb, int K, int lane) {
  double bestd = 1.0e300;
  int besti = 0x7fffffff;
  for (int k0 = lane; k0 < K; k0 += 128) {
    const int k1 = min(k0 + 32, K - 1), k2 = min(k0 + 64, K - 1), k3 = min(k0 + 96, K - 1);
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll 2
    for (int z = 0; z < VQ_Z; ++z) {
      const double x2 = 2.0 * (double)xrow[z];
      const float* cr = cb + (size_t)z * K;
      const double c0 = (double)cr[k0], c1 = (double)cr[k1], c2 = (double)cr[k2], c3 = (double)cr[k3];
      d0 += c0 * (c0 - x2); d1 += c1 * (c1 - x2); d2 += c2 * (c2 - x2); d3 += c3 * (c3 - x2);
    }
    if (d0 < bestd) { bestd = d0; besti = k0; }                  // increasing k: strict < keeps the first minimum
    if (k0 + 32 < K && d1 < bestd) { bestd = d1; besti = k0 + 32; }
    if (k0 + 64 < K && d2 < bestd) { bestd = d2; besti = k0 + 64; }
    if (k0 + 96 < K && d3 < bestd) { bestd = d3; besti = k0 + 96; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bestd, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (od < bestd || (od == bestd && oi < besti)) { bestd = od; besti = oi; }
  }
  return besti;
}

// The tile's final decision for the 32 rows a warp owns (one row per lane): m = the row's packed (best, runner-up,
// third), xsum = ||x||^2.  Near-ties are resolved in fp64, warp-cooperatively; returns the row's index.
__device__ __forceinline__ int vt_resolve(float mb, float ms, float mt, int bblk, int sblk, uint32_t kmask, float xsum,
                                          const float* c2_s, bool ok, const float* __restrict__ x, long long row0,
                                          const float* __restrict__ cb, int K, int lane) {
  const int bi = 256 * bblk + (int)(__float_as_uint(mb) & ~kmask), si = 256 * sblk + (int)(__float_as_uint(ms) & ~kmask);
  const float bv = __uint_as_float(__float_as_uint(mb) & kmask), sv = __uint_as_float(__float_as_uint(ms) & kmask);
  const float tv = __uint_as_float(__float_as_uint(mt) & kmask);
  int best = bi;
  const float scale = fmaxf(xsum + c2_s[min(bi, VT_MAXK - 1)], 1e-3f);
  const float trunc2 = 2.0f * __uint_as_float(0x34000000u) * (float)(~kmask + 1u);   // 2 * 2^-23 * 2^ibits
  const bool near3 = ok && K > 2 && (tv - bv) <= VT_TOL3 * scale + trunc2 * fmaxf(fabsf(bv), fabsf(tv));
  const bool near2 = ok && K > 1 && !near3 && (sv - bv) <= VT_TOL2 * scale + trunc2 * fmaxf(fabsf(bv), fabsf(sv));
  unsigned need = __ballot_sync(0xffffffffu, near2);
  while (need) {                                                // warp-uniform: fp64 re-score of the two candidates
    const int src = __ffs(need) - 1;
    need &= need - 1;
    const float* xr = x + (row0 + src) * VQ_Z;
    const int i1 = __shfl_sync(0xffffffffu, bi, src), i2 = __shfl_sync(0xffffffffu, si, src);
    double d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int mm = 0; mm < 8; ++mm) {
      const int z = lane + 32 * mm;
      const double xv = (double)xr[z];
      const double c1 = (double)cb[(size_t)z * K + i1], c2v = (double)cb[(size_t)z * K + i2];
      d1 += c1 * c1 - 2.0 * xv * c1;
      d2 += c2v * c2v - 2.0 * xv * c2v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    const bool first2 = (d2 < d1) || (d2 == d1 && i2 < i1);
    if (lane == src) best = first2 ? i2 : i1;
  }
  need = __ballot_sync(0xffffffffu, near3);
  while (need) {                                                // three-way near-tie: every codeword, in fp64
    const int src = __ffs(need) - 1;
    need &= need - 1;
    const int fi = vt_full_rescore(x + (row0 + src) * VQ_Z, cb, K, lane);
    if (lane == src) best = fi;
  }
  return best;
}

// codebook block -> per K-chunk swizzled images: plane H = tf32 hi [Npad x 32 fp32], plane C = [bf16(hi) x 32 | bf16(lo) x 32]
// (the layout of mlp_tc.cu's tc_pack_kernel; the codeword is the N row, z the K index), and c2 = ||c||^2
__global__ void vt_pack_kernel(const float* __restrict__ cb, int K, int nb, uint8_t* __restrict__ wpack,
                               float* __restrict__ c2) {
  const long long total = (long long)nb * VT_KCHUNKS * 256 * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % 32);
    const int nrow = (int)((i / 32) % 256);
    const int c = (int)((i / (32 * 256)) % VT_KCHUNKS);
    const int b = (int)(i / (32 * 256 * VT_KCHUNKS));
    const int nblk = min(256, K - 256 * b), npad = nblk <= 0 ? 16 : (nblk + 15) / 16 * 16;   // nblk <= 0: padding block
    if (nrow >= npad) continue;
    const int z = c * 32 + kk, col = 256 * b + nrow;
    const float v = nrow < nblk ? cb[(size_t)z * K + col] : 0.f;
    const size_t plane = (size_t)npad * 128;
    uint8_t* base = wpack + (size_t)b * VT_KCHUNKS * VT_W_SLOT + (size_t)c * 2 * plane;
    const float hi = tc::tf32_rna(v);
    *reinterpret_cast<float*>(base + tc::sw128_off(nrow, kk / 4) + (kk % 4) * 4) = hi;
    uint8_t* pc = base + plane;
    *reinterpret_cast<__nv_bfloat16*>(pc + tc::sw128_off(nrow, kk / 8) + (kk % 8) * 2) = __float2bfloat16_rn(hi);
    *reinterpret_cast<__nv_bfloat16*>(pc + tc::sw128_off(nrow, 4 + kk / 8) + (kk % 8) * 2) = __float2bfloat16_rn(v - hi);
  }
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nb * 256; j += gridDim.x * blockDim.x) {
    float s = 3.0e38f;                                         // padding columns never win
    if (j < K) {
      s = 0.f;
      for (int z = 0; z < VQ_Z; ++z) { const float v = cb[(size_t)z * K + j]; s = fmaf(v, v, s); }
    }
    c2[j] = s;
  }
}

// 16 consecutive K values of row r into the A slot (tf32 hi plane + [bf16(lo) | bf16(hi)] plane), as mlp_tc.cu
__device__ __forceinline__ void vt_store16(uint32_t slot, int r, int j0, const float (&v)[16]) {
  const uint32_t row = slot + (uint32_t)r * 128u;           // shared-space address: STS, not generic stores
  const uint32_t rx = (uint32_t)(r & 7);
#pragma unroll
  for (int qq = 0; qq < 2; ++qq) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { h[i] = tc::tf32_rna(v[8 * qq + i]); l[i] = v[8 * qq + i] - h[i]; }
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
    tc::sts128(row + VT_A_PLANE + ((cc ^ rx) << 4), ul);
    tc::sts128(row + VT_A_PLANE + (((cc + 4) ^ rx) << 4), uh);
  }
}

__device__ __forceinline__ void vt_group_bar(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

// 2-D tiled TMA load (tensor map in kernel parameter space) -> shared memory, completion counted on `bar` in bytes
__device__ __forceinline__ void vt_tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int crd0, int crd1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(tc::smem_u32(bar)), "r"(crd0),
               "r"(crd1) : "memory");
}

// K <= 128 (one codeword block).  There the kernel is bound by the latency of the latent loads, so they go
// through the TMA engine: one thread keeps VT_XSTAGES boxes of [128 latents x 32 values] in flight (tensor map with the
// 128-byte swizzle = the layout the producers read conflict-free), the producer groups only split and store.
__global__ void __launch_bounds__(VT_THREADS, 1) vq_tc_kernel(const __grid_constant__ VtParams p,
                                                              const __grid_constant__ CUtensorMap xmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[VT_STAGES], a_empty[VT_STAGES], w_full[VT_WSTAGES_MAX], w_empty[VT_WSTAGES_MAX];
  __shared__ __align__(8) uint64_t acc_full[2], drain_done[2];
  __shared__ __align__(8) uint64_t x_full[VT_XSTAGES_MAX], x_empty[VT_XSTAGES_MAX];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float c2_s[VT_MAXK];
  __shared__ float xs_s[2][VT_M * 4];                        // partial ||x||^2 of the 4 producer threads of a row, per tile parity
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + (size_t)VT_STAGES * VT_A_SLOT;
  // codebook ring: two chunk-sized slots (more were measured: no gain)
  const uint32_t npad0 = (uint32_t)((min(256, p.K) + 15) / 16 * 16);
  const uint32_t W_SLOT = npad0 * 256u;
  const uint32_t NWS = (uint32_t)VT_STAGES;
  // whatever the two codebook slots leave of the 128 KB goes to the TMA staging ring of the latents (4 boxes of
  // 16 KB at K = 128, 6 at K <= 64, 7 at K <= 32): the depth of that ring is what bounds the small-K case
  uint8_t* x_ring = w_ring + (size_t)NWS * W_SLOT;
  const uint32_t NXS = min((uint32_t)VT_XSTAGES_MAX, (2u * VT_W_SLOT - NWS * W_SLOT) / VT_X_STAGE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MMA_WARP = 8 * VT_G, W_WARP = 8 * VT_G + 1, X_WARP = 8 * VT_G + 2, D_WARP0 = 8 * VT_G + 3;

  if (warp == MMA_WARP) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < VT_STAGES; ++i) { tc::mbar_init(&a_full[i], 256); tc::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < VT_WSTAGES_MAX; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&drain_done[i], 128); }
    for (int i = 0; i < VT_XSTAGES_MAX; ++i) { tc::mbar_init(&x_full[i], 1); tc::mbar_init(&x_empty[i], 256); }
    tc::mbar_fence_init();
  }
  for (int i = tid; i < p.nb * 256; i += VT_THREADS) c2_s[i] = p.c2[i];
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const long long n_tiles = (p.n + VT_M - 1) / VT_M;
  const int nb = p.nb;
  const uint32_t kmask = ~((p.K <= 64 ? 64u : 128u) - 1u);         // packed-key index bits (one codeword block, K <= 128)

  if (warp < 8 * VT_G) {
    // ===================== producers: A chunks (split into the tf32 + bf16-correction planes), ||x||^2 =====================
    const int grp = warp >> 3, half = (warp >> 2) & 1;
    const int r = 32 * (warp & 3) + lane;                          // latent row of the tile
    uint32_t ga = 0;                                               // global chunk counter
    uint32_t tile_i = 0;
    const uint32_t rt_zero = (uint32_t)p.K & 0x40000000u;          // 0 (K <= 1024), not foldable at compile time
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_i) {
      float xs = 0.f;
      for (int b = 0; b < nb; ++b) {
        for (int c = 0; c < VT_KCHUNKS; ++c, ++ga) {
          if ((int)(ga % VT_G) != grp) continue;
          const int slot = (int)(ga % VT_STAGES);
          const uint32_t dst = tc::smem_u32(a_ring) + (uint32_t)slot * VT_A_SLOT;
          float v[16];
          // the chunk was put into staging buffer ga % VT_XSTAGES by the TMA engine (swizzled like the A planes)
          const int xs_i = (int)(ga % NXS);
          tc::mbar_wait(&x_full[xs_i], (ga / NXS) & 1u);
          const uint32_t stage = tc::smem_u32(x_ring) + (uint32_t)xs_i * VT_X_STAGE;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = tc::lds128(stage + r * 128 + (((4 * half + q) ^ (r & 7)) << 4));
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
          }
          // The staging buffer may be refilled -- but only once the LDS results above have RETURNED: an
          // mbarrier.arrive issued right behind the loads is not ordered behind them by the hardware (it overtakes
          // them when the shared-memory pipe is busy with the UMMA operand reads, and the TMA engine then overwrites
          // rows that have not been read yet: ~1e-3 of the rows got another tile's values, non-deterministically).
          // The arrive therefore takes a register dependency on the last word of every load (an address offset that
          // is zero at run time but unknown to the compiler).
          {
            uint32_t dep;
            asm volatile("{\n\t.reg .b32 t0, t1;\n\tor.b32 t0, %1, %2;\n\tor.b32 t1, %3, %4;\n\tor.b32 t0, t0, t1;\n\t"
                         "and.b32 %0, t0, %5;\n\t}"
                         : "=r"(dep) : "f"(v[3]), "f"(v[7]), "f"(v[11]), "f"(v[15]), "r"(rt_zero));
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&x_empty[xs_i]) + dep) : "memory");
          }
          tc::mbar_wait(&a_empty[slot], ((ga / VT_STAGES) & 1u) ^ 1u);
          if (b == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xs = fmaf(v[j], v[j], xs);
            // the group's last chunk of the tile's first block: publish this thread's share of ||x||^2 for the drain
            // warps BEFORE the chunk is handed over (they read it after the block's accumulator is complete)
            if (c + VT_G >= VT_KCHUNKS) xs_s[tile_i & 1u][r * 4 + grp * 2 + half] = xs;
          }
          vt_store16(dst, r, 16 * half, v);
          tc::fence_proxy_async();
          tc::fence_before_sync();
          tc::mbar_arrive(&a_full[slot]);
        }
      }
    }
  } else if (warp >= D_WARP0) {
    // ===================== drain warps: a thread owns a latent row, arg-min over every codeword block =====================
    const int q4 = warp & 3;                                       // a warp reaches the TMEM lanes 32 (warp id % 4) ..; the four
                                                                   // drain warps cover the four quarters
    const int r = 32 * q4 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q4) << 16);
    uint32_t gk = 0, tile_i = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_i) {
      Top3 m = {3.0e38f, 3.0e38f, 3.0e38f};
      float xsum = 0.f;
      for (int b = 0; b < nb; ++b, ++gk) {
        const int region = (int)(gk & 1u);
        tc::mbar_wait(&acc_full[region], (gk >> 1) & 1u);
        tc::fence_after_sync();
        const int nblk = min(256, p.K - 256 * b), npad = (nblk + 15) / 16 * 16;
        const float* c2b = c2_s + 256 * b;
        for (int c16 = 0; c16 < npad; c16 += 16) {
          float v[16];
          tc::tmem_ld16(lane_addr + (uint32_t)(region * 256 + c16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) top3_push(m, vt_pack(fmaf(-2.0f, v[j], c2b[c16 + j]), kmask, 256 * b + c16 + j));
        }
        if (b == nb - 1) {                                        // before the hand-over: the producers reuse xs_s two tiles on
#pragma unroll
          for (int q = 0; q < 4; ++q) xsum += xs_s[tile_i & 1u][r * 4 + q];
        }
        tc::fence_before_sync();
        tc::mbar_arrive(&drain_done[region]);                     // the tensor cores may refill this TMEM region
      }
      // ---- the tile's arg-min: near-ties re-scored in fp64, index written ----
      const long long row = tile * VT_M + r;
      const bool ok = row < p.n;
      const int best = vt_resolve(m.b, m.s, m.t, 0, 0, kmask, xsum, c2_s, ok, p.x, tile * VT_M + 32 * q4, p.cb, p.K, lane);
      if (ok) p.idx_out[row] = (long long)best;
    }
  } else if (warp == MMA_WARP) {
    {   // whole warp walks the loop, one elected lane issues (uniform operands: see tc::mma_ss_e)
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t ga = 0, gk = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int b = 0; b < nb; ++b, ++gk) {
          const int region = (int)(gk & 1u);
          const int nblk = min(256, p.K - 256 * b), npad = (nblk + 15) / 16 * 16;
          const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, VT_M, npad);
          const uint32_t idesc_c = tc::make_idesc(tc::FMT_BF16, VT_M, npad);
          const uint32_t d_tmem = tmem_u + (uint32_t)(region * 256);
          const uint32_t w_plane = (uint32_t)npad * 128;
          if (gk >= 2) {                                          // the drain of the layer that used this region
            tc::mbar_wait_u(&drain_done[region], ((gk - 2) >> 1) & 1u);
            tc::fence_after_sync();
          }
          uint32_t acc = 0;
          for (int c = 0; c < VT_KCHUNKS; ++c, ++ga) {
            const int s_ = (int)(ga % VT_STAGES);
            const int ws = (int)(ga % NWS);
            tc::mbar_wait_u(&a_full[s_], (ga / VT_STAGES) & 1u);
            tc::mbar_wait_u(&w_full[ws], (ga / NWS) & 1u);
            tc::fence_after_sync();
            const uint32_t a_addr = tc::smem_u32(a_ring + (size_t)s_ * VT_A_SLOT);
            const uint32_t w_addr = tc::smem_u32(w_ring + (size_t)ws * W_SLOT);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              tc::mma_ss_e<true>(d_tmem, tc::make_desc_sw128(a_addr + 32 * s), tc::make_desc_sw128(w_addr + 32 * s), idesc, acc);
              acc = 1;
              tc::mma_ss_e<false>(d_tmem, tc::make_desc_sw128(a_addr + VT_A_PLANE + 32 * s),
                                tc::make_desc_sw128(w_addr + w_plane + 32 * s), idesc_c, 1);
            }
            tc::mma_commit_e(&a_empty[s_]);
            tc::mma_commit_e(&w_empty[ws]);
          }
          tc::mma_commit_e(&acc_full[region]);
        }
      }
    }
    __syncwarp();
  } else if (warp == W_WARP) {
    if (lane == 0) {
      uint32_t gw = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int b = 0; b < nb; ++b) {
          const int nblk = min(256, p.K - 256 * b), npad = (nblk + 15) / 16 * 16;
          const uint32_t bytes = (uint32_t)npad * 128 * 2;
          const uint8_t* wb = p.wpack + (size_t)b * VT_KCHUNKS * VT_W_SLOT;
          for (int c = 0; c < VT_KCHUNKS; ++c, ++gw) {
            const int s_ = (int)(gw % NWS);
            tc::mbar_wait(&w_empty[s_], ((gw / NWS) & 1u) ^ 1u);
            tc::mbar_expect_tx(&w_full[s_], bytes);
            tc::bulk_g2s(w_ring + (size_t)s_ * W_SLOT, wb + (size_t)c * bytes, bytes, &w_full[s_]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == X_WARP) {
    // =========================== latent chunks through the TMA engine (one thread) ===========================
    if (lane == 0) {
      uint32_t gx = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int c = 0; c < VT_KCHUNKS; ++c, ++gx) {              // nb == 1
          const int xs_i = (int)(gx % NXS);
          tc::mbar_wait(&x_empty[xs_i], ((gx / NXS) & 1u) ^ 1u);
          tc::mbar_expect_tx(&x_full[xs_i], VT_X_STAGE);
          vt_tma_load_2d(x_ring + (size_t)xs_i * VT_X_STAGE, &xmap, &x_full[xs_i], c * 32, (int)(tile * VT_M));
        }
      }
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tc::tmem_dealloc(tmem_base, 512);
}

// K > 128: the 16 producer warps also drain the accumulators (at these sizes the A-chunk production has slack and the
// extra drain throughput matters: four dedicated drain warps were measured 10 % slower at K = 1024, 9.4 -> 10.4 ms).
//   PAIR = false (one codeword block, 128 < K <= 256): a "layer" is one block; TMEM regions ping-pong between tiles.
//   PAIR = true  (K > 256): a layer is a PAIR of blocks -- every A chunk feeds two MMAs, block 2j into TMEM region 0 and
//   block 2j+1 into region 1 -- because re-producing the A chunks per block (load, split, 32 KB of smem stores) was what
//   bounded the large-K case; the block count is padded to an even number (an all-padding block never wins).  Both
//   regions are busy during a layer, so its drain runs between layers instead of under the next layer's MMAs.
template <bool PAIR>
__global__ void __launch_bounds__(VT_THREADS_BIG, 1) vq_tc_big_kernel(const __grid_constant__ VtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[VT_STAGES], a_empty[VT_STAGES], w_full[VT_STAGES], w_empty[VT_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], drain_done[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float c2_s[VT_MAXK];
  __shared__ __align__(16) float4 merge_s[2][VT_M * 4];      // packed Top3 + ||x||^2 share of the 4 threads of a row, double-buffered by tile
  __shared__ int mergeb_s[2][VT_M * 4];                      // codeword block of their best | runner-up << 8
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + (size_t)VT_STAGES * VT_A_SLOT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MMA_WARP = 8 * VT_G, W_WARP = 8 * VT_G + 1;
  constexpr int BPL = PAIR ? 2 : 1;                                // codeword blocks per layer

  if (warp == MMA_WARP) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < VT_STAGES; ++i) {
      tc::mbar_init(&a_full[i], 256); tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&drain_done[i], 256 * VT_G); }
    tc::mbar_fence_init();
  }
  for (int i = tid; i < p.nb * 256; i += VT_THREADS_BIG) c2_s[i] = p.c2[i];
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const long long n_tiles = (p.n + VT_M - 1) / VT_M;
  const int nl = p.nb / BPL;                                       // layers per tile (p.nb is even in PAIR mode)
  const uint32_t kmask = ~255u;                                    // packed keys carry the index inside the 256-codeword block
  // columns of codeword block b (a padding block of PAIR mode has 16 all-padding columns)
  auto npad_of = [&](int b) { const int nblk = min(256, p.K - 256 * b); return nblk <= 0 ? 16 : (nblk + 15) / 16 * 16; };

  if (warp < 8 * VT_G) {
    // ===================== producers: A chunks, accumulator drain, arg-min =====================
    const int grp = warp >> 3, half = (warp >> 2) & 1, tg = tid & 255;
    const int r = 32 * (warp & 3) + lane;                          // TMEM lane == latent row of the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t ga = 0, gk = 0;                                       // global chunk / layer counters
    Run3 t2 = {3.0e38f, 3.0e38f, 3.0e38f, 0, 0};
    float xs = 0.f, xs_fin = 0.f;
    long long pend_tile = -1; int pend_l = 0; uint32_t pend_k = 0; uint32_t tiles_done = 0;

    auto drain = [&](long long tile, int l, uint32_t k) {
#pragma unroll
      for (int h2 = 0; h2 < BPL; ++h2) {
        const int region = PAIR ? h2 : (int)(k & 1u);
        const int b = PAIR ? 2 * l + h2 : l;
        tc::mbar_wait(&acc_full[region], PAIR ? (k & 1u) : ((k >> 1) & 1u));
        tc::fence_after_sync();
        const int npad = npad_of(b);
        const float* c2b = c2_s + 256 * b;
        Top3 blk = {3.0e38f, 3.0e38f, 3.0e38f};
        for (int cbk = grp; cbk * 32 < npad; cbk += VT_G) {
          const int c16 = cbk * 32 + 16 * half;
          if (c16 < npad) {
            float v[16];
            tc::tmem_ld16(lane_addr + (uint32_t)(region * 256 + c16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) top3_push(blk, vt_pack(fmaf(-2.0f, v[j], c2b[c16 + j]), kmask, c16 + j));
          }
        }
        tc::fence_before_sync();
        tc::mbar_arrive(&drain_done[region]);
        run3_push(t2, blk.b, b); run3_push(t2, blk.s, b); run3_push(t2, blk.t, b);
      }
      if (l + 1 < nl) return;
      // ---- last layer of the tile: merge the 4 partial results of each row, re-score near-ties, write the index ----
      const int buf = (int)(tiles_done & 1u);
      ++tiles_done;
      merge_s[buf][r * 4 + grp * 2 + half] = make_float4(t2.b, t2.s, t2.t, xs_fin);
      mergeb_s[buf][r * 4 + grp * 2 + half] = t2.bb | (t2.sb << 8);
      t2.b = 3.0e38f; t2.s = 3.0e38f; t2.t = 3.0e38f; t2.bb = 0; t2.sb = 0;
      asm volatile("bar.sync 3, %0;" ::"r"(256 * VT_G) : "memory");
      if (grp == 0 && half == 0) {
        Run3 m = {3.0e38f, 3.0e38f, 3.0e38f, 0, 0};
        float xsum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 e = merge_s[buf][r * 4 + q];
          const int eb = mergeb_s[buf][r * 4 + q];
          // a partial's third can only enter the merged top two through a tie, and ties go to the full re-score
          run3_push(m, e.x, eb & 255); run3_push(m, e.y, eb >> 8); run3_push(m, e.z, eb >> 8);
          xsum += e.w;
        }
        const long long row = tile * VT_M + r;
        const bool ok = row < p.n;
        const int best = vt_resolve(m.b, m.s, m.t, m.bb, m.sb, kmask, xsum, c2_s, ok, p.x, tile * VT_M + 32 * (warp & 3),
                                    p.cb, p.K, lane);
        if (ok) p.idx_out[row] = (long long)best;
      }
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 0; l < nl; ++l, ++gk) {
        for (int c = 0; c < VT_KCHUNKS; ++c, ++ga) {
          if ((int)(ga % VT_G) != grp) continue;
          const int slot = (int)(ga % VT_STAGES);
          const uint32_t dst = tc::smem_u32(a_ring) + (uint32_t)slot * VT_A_SLOT;
          // coalesced load of the chunk (a warp reads 4 rows x 128 B per instruction), issued BEFORE the slot is claimed.
          // (Keeping the group's next chunk in registers across the drain was measured slower: 9.6 -> 11.4 ms at
          // K = 1024, the extra live registers spill inside the arg-min loop.)
          float4 ldv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int f = tg + 256 * i, rr = f >> 3, ch = f & 7;
            const long long prow = tile * VT_M + rr;
            ldv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (prow < p.n) ldv[i] = __ldg(reinterpret_cast<const float4*>(p.x + prow * VQ_Z + c * 32 + 4 * ch));
          }
          tc::mbar_wait(&a_empty[slot], ((ga / VT_STAGES) & 1u) ^ 1u);
          const uint32_t stage = dst + VT_A_PLANE;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int f = tg + 256 * i, rr = f >> 3, ch = f & 7;
            tc::sts128(stage + rr * 128 + ((ch ^ (rr & 7)) << 4), ldv[i]);
          }
          vt_group_bar(grp);
          float v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = tc::lds128(stage + r * 128 + (((4 * half + q) ^ (r & 7)) << 4));
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
          }
          vt_group_bar(grp);                                     // every row has been read before plane C is overwritten
          if (l == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xs = fmaf(v[j], v[j], xs);
          }
          vt_store16(dst, r, 16 * half, v);
          tc::fence_proxy_async();
          tc::fence_before_sync();
          tc::mbar_arrive(&a_full[slot]);
        }
        if (PAIR) {
          if (l == nl - 1) { xs_fin = xs; xs = 0.f; }
          drain(tile, l, gk);                                    // both regions are busy: the next layer waits for this drain
        } else {
          // the previous layer (of the previous tile: xs_fin still holds ITS ||x||^2) is drained while the tensor cores
          // work on this one
          if (pend_tile >= 0) drain(pend_tile, pend_l, pend_k);
          pend_tile = tile; pend_l = l; pend_k = gk;
          if (l == nl - 1) { xs_fin = xs; xs = 0.f; }
        }
      }
    }
    if (!PAIR && pend_tile >= 0) drain(pend_tile, pend_l, pend_k);
  } else if (warp == MMA_WARP) {
    {   // whole warp walks the loop, one elected lane issues (uniform operands: see tc::mma_ss_e)
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t ga = 0, gk = 0, gw = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < nl; ++l, ++gk) {
          if (PAIR) {
            if (gk >= 1) {                                        // both regions: the drain of the previous layer
              tc::mbar_wait_u(&drain_done[0], (gk - 1) & 1u);
              tc::mbar_wait_u(&drain_done[1], (gk - 1) & 1u);
              tc::fence_after_sync();
            }
          } else if (gk >= 2) {                                   // the drain of the layer that used this region
            tc::mbar_wait_u(&drain_done[gk & 1u], ((gk - 2) >> 1) & 1u);
            tc::fence_after_sync();
          }
          uint32_t acc = 0;
          for (int c = 0; c < VT_KCHUNKS; ++c, ++ga) {
            const int s_ = (int)(ga % VT_STAGES);
            tc::mbar_wait_u(&a_full[s_], (ga / VT_STAGES) & 1u);
            const uint32_t a_addr = tc::smem_u32(a_ring + (size_t)s_ * VT_A_SLOT);
#pragma unroll
            for (int h2 = 0; h2 < BPL; ++h2, ++gw) {
              const int region = PAIR ? h2 : (int)(gk & 1u);
              const int npad = npad_of(PAIR ? 2 * l + h2 : l);
              const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, VT_M, npad);
              const uint32_t idesc_c = tc::make_idesc(tc::FMT_BF16, VT_M, npad);
              const uint32_t d_tmem = tmem_u + (uint32_t)(region * 256);
              const uint32_t w_plane = (uint32_t)npad * 128;
              const int ws = (int)(gw % VT_STAGES);
              tc::mbar_wait_u(&w_full[ws], (gw / VT_STAGES) & 1u);
              tc::fence_after_sync();
              const uint32_t w_addr = tc::smem_u32(w_ring + (size_t)ws * VT_W_SLOT);
              uint32_t first = acc;                               // 0 only for the first K-step of the layer's first chunk
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                tc::mma_ss_e<true>(d_tmem, tc::make_desc_sw128(a_addr + 32 * s), tc::make_desc_sw128(w_addr + 32 * s), idesc, first);
                first = 1;
                tc::mma_ss_e<false>(d_tmem, tc::make_desc_sw128(a_addr + VT_A_PLANE + 32 * s),
                                  tc::make_desc_sw128(w_addr + w_plane + 32 * s), idesc_c, 1);
              }
              tc::mma_commit_e(&w_empty[ws]);
            }
            acc = 1;
            tc::mma_commit_e(&a_empty[s_]);
          }
          if (PAIR) { tc::mma_commit_e(&acc_full[0]); tc::mma_commit_e(&acc_full[1]); }
          else tc::mma_commit_e(&acc_full[gk & 1u]);
        }
      }
    }
    __syncwarp();
  } else if (warp == W_WARP) {
    if (lane == 0) {
      uint32_t gw = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < nl; ++l) {
          for (int c = 0; c < VT_KCHUNKS; ++c) {
#pragma unroll
            for (int h2 = 0; h2 < BPL; ++h2, ++gw) {
              const int b = PAIR ? 2 * l + h2 : l;
              const uint32_t bytes = (uint32_t)npad_of(b) * 128 * 2;
              const uint8_t* wb = p.wpack + (size_t)b * VT_KCHUNKS * VT_W_SLOT;
              const int s_ = (int)(gw % VT_STAGES);
              tc::mbar_wait(&w_empty[s_], ((gw / VT_STAGES) & 1u) ^ 1u);
              tc::mbar_expect_tx(&w_full[s_], bytes);
              tc::bulk_g2s(w_ring + (size_t)s_ * VT_W_SLOT, wb + (size_t)c * bytes, bytes, &w_full[s_]);
            }
          }
        }
      }
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace

int vq_tc_min_k() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VQN_VQ_TC_MIN_K"); v = e ? atoi(e) : 33; if (v < 16) v = 16; }
  return v;
}

int vq_tc_assign_launch(vqn_ctx* ctx, const VqParams& q, cudaStream_t s) {
  // codebook images are rebuilt on every call (the codebook is a caller tensor), stream-ordered, in per-stream buffers
  uint8_t* wpack = static_cast<uint8_t*>(vqn_stream_scratch(ctx, VQN_SCRATCH_VQ_W, s, (size_t)4 * VT_KCHUNKS * VT_W_SLOT));
  float* c2 = static_cast<float*>(vqn_stream_scratch(ctx, VQN_SCRATCH_VQ_C2, s, sizeof(float) * VT_MAXK));
  if (!wpack || !c2) return VQN_ERR_CUDA;
  VtParams p;
  p.x = q.x; p.n = q.n; p.cb = q.cb; p.K = q.K; p.nb = (q.K + 255) / 256;
  if (p.nb > 1 && (p.nb & 1)) p.nb += 1;        // K > 256: blocks are processed in pairs (an all-padding block never wins)
  p.wpack = wpack; p.c2 = c2; p.idx_out = q.idx_out;
  vt_pack_kernel<<<256, 256, 0, s>>>(q.cb, q.K, p.nb, wpack, c2);
  VQN_LAUNCHED(ctx);
  const long long tiles = (q.n + VT_M - 1) / VT_M;
  const int blocks = (int)(tiles < (long long)ctx->sm_count ? tiles : (long long)ctx->sm_count);
  CUtensorMap xmap;
  memset(&xmap, 0, sizeof(xmap));
  const bool smallk = q.K <= 128 && (reinterpret_cast<uintptr_t>(q.x) & 15) == 0 && q.n < (1ll << 31);
  if (smallk) {
    // latents [n, 256] fp32 row-major; box = 32 values x 128 rows = one K-chunk of a tile; rows past n read as zero
    const cuuint64_t gdim[2] = {(cuuint64_t)VQ_Z, (cuuint64_t)q.n};
    const cuuint64_t gstr[1] = {(cuuint64_t)VQ_Z * sizeof(float)};
    const cuuint32_t box[2] = {32, VT_M};
    const cuuint32_t estr[2] = {1, 1};
    // the driver entry point is resolved through the runtime, so the library has no link-time dependency on libcuda.so
    // (it must load on machines without a driver: build checks, symbol tests)
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      VQN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess) { vqn_set_error("cuTensorMapEncodeTiled is not available"); return VQN_ERR_UNSUPPORTED; }
      encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const CUresult cr = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(q.x), gdim, gstr,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { vqn_set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return VQN_ERR_CUDA; }
    VQN_CUDA(cudaFuncSetAttribute(vq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VT_SMEM));
    vq_tc_kernel<<<blocks, VT_THREADS, VT_SMEM, s>>>(p, xmap);
  } else if (p.nb > 1) {
    VQN_CUDA(cudaFuncSetAttribute(vq_tc_big_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VT_SMEM));
    vq_tc_big_kernel<true><<<blocks, VT_THREADS_BIG, VT_SMEM, s>>>(p);
  } else {
    VQN_CUDA(cudaFuncSetAttribute(vq_tc_big_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VT_SMEM));
    vq_tc_big_kernel<false><<<blocks, VT_THREADS_BIG, VT_SMEM, s>>>(p);
  }
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
