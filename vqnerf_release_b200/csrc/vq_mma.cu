// VQ nearest-codeword assignment on the warp-level tensor-core path (networks/vq_layers.py:257-302).
//
// Why mma.sync here and not tcgen05: the assignment is HBM-bound for the shipped K = 15 (1032 B against
// 2*256*16 MACs per latent), so the kernel must spend as few issue slots and as little shared-memory
// bandwidth per byte as possible.  mma.sync takes its A operand from REGISTERS, so a latent row goes
// HBM -> registers -> tensor core without ever touching shared memory (a tcgen05 SS-MMA would need each byte
// written to and re-read from smem 3x for the hi/lo split, more than the 128 B/clk/SM smem port affords at
// 22 B/clk/SM of HBM).  The FFMA form needed ~245 issue slots per row (budget at the HBM roofline: 183);
// this one needs ~50.
//
// Warp tile = 16 latent rows.  Lane (g = lane>>2, t = lane&3) loads 16-byte pieces of rows g and g+8
// (z = 16c + 4t .. +3, c = 0..15: four lanes cover 64 contiguous bytes, two consecutive c's a full 128 B line)
// straight into the m16n8k8 A-fragment positions: k-step s = 2c+u uses z0 = 16c+4t+2u at column t and
// z0+1 at column t+4; the codebook B fragments are pre-split and stored in exactly that order in shared
// memory ({b0.hi, b1.hi, b0.lo, b1.lo} per lane -> one conflict-free LDS.128 per (k-step, n-tile)).
// fp32 parity comes from the 3-term split  x.c = xh.ch + (xl.ch + xh.cl),  hi = cvt.rna.tf32(v), lo = v - hi:
// the leading term is a tf32 MMA, the two corrections (2^-11 of it) ONE bf16 m16n8k16 MMA (4 instead of 6 MMAs
// per k-step at K <= 16), with the large and the small products in separate fp32 accumulators; rows
// whose best two distances are closer than 4e-5 relative are re-scored in fp64, so indices are exact
// whenever the true top-2 gap exceeds the 1e-6 tolerance of BASELINE.json.  Loads are software-pipelined
// in four phases per tile through two register buffers (one phase = 4 KB per warp in flight behind the
// one being consumed; 16 warps/SM).
//
// K > 8*NT (NT <= 8 n-tiles, i.e. K > 64): codeword chunks of 64 are staged block-wide, every warp sweeps a
// batch of 4 tiles per chunk (x re-read from L2) and the running (best, runner-up) pairs live in smem.
//
// Distances use the reference's algebraic form ||x||^2 - 2 x.c + ||c||^2 (vq_layers.py:279-282); arg-min keeps
// the FIRST minimum (tf.argmax(-d), :292).  The optional second pass re-reads the tile row-wise (L2 hits) for
// quantize / z_norm / EMA statistics.
#include "vq.cuh"

#define VQM_TPW 4            // tiles per warp per batch in chunked mode
// NT <= 4 n-tiles: 4-warp blocks, three per SM (<= 170 registers/thread, no spills, 12 warps/SM);
// NT == 8: the 128 KB of B fragments allow one block per SM, so it gets 8 warps.
template <int NT> struct VqmCfg { static constexpr int WARPS = NT <= 4 ? 4 : 8, BLOCKS = NT <= 4 ? 3 : 1; };

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// {low half: bf16(a), high half: bf16(b)}
__device__ __forceinline__ unsigned pack_bf16(float a, float b) {
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ unsigned cvt_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = cvt_tf32(x);
  lo = cvt_tf32(x - __uint_as_float(hi));
}
template <bool STREAM>
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  if (STREAM) return ldg_stream_f4(p);
  return __ldg(reinterpret_cast<const float4*>(p));
}

struct Top2 {
  float b, s, t;       // best, runner-up, THIRD-best value (a three-way near-tie sends the row to the full fp64 re-score)
  int bi, si;
};
__device__ __forceinline__ void top2_push(Top2& t, float d, int k) {   // k arrives in increasing order
  if (d < t.b) { t.t = t.s; t.s = t.b; t.si = t.bi; t.b = d; t.bi = k; }
  else if (d < t.s) { t.t = t.s; t.s = d; t.si = k; }
  else if (d < t.t) t.t = d;
}
// merge with another (best, runner-up, third) triple; equal distances -> lower index first
__device__ __forceinline__ void top2_merge(Top2& t, float ob, int obi, float os, int osi, float ot) {
  // third smallest of the union = min(third of each side, third smallest of the four indexed values)
  const float t3 = fminf(fminf(t.t, ot), fmaxf(fminf(t.s, os), fmaxf(t.b, ob)));
  const bool ow = (ob < t.b) || (ob == t.b && obi < t.bi);
  const float lose = ow ? t.b : ob;  const int losei = ow ? t.bi : obi;
  const float w2 = ow ? os : t.s;    const int w2i = ow ? osi : t.si;
  if (ow) { t.b = ob; t.bi = obi; }
  const bool l = (lose < w2) || (lose == w2 && losei < w2i);
  t.s = l ? lose : w2;  t.si = l ? losei : w2i;
  t.t = t3;
}

// Hot-loop form of the same triple: PACKED keys, the distance with its low 6 mantissa bits replaced by the chunk-local
// codeword index (< 64): five FMNMX per value, three shuffles per merge step, no index bookkeeping.  The truncation
// (2^-17 relative) is covered by the fp64 re-score thresholds below.
struct Top3 { float b, s, t; };
__device__ __forceinline__ void top3_push(Top3& m, float dp) {
  m.t = fminf(m.t, fmaxf(dp, m.s));
  m.s = fminf(m.s, fmaxf(dp, m.b));
  m.b = fminf(m.b, dp);
}
#define VQM_KMASK 0xffffffc0u
__device__ __forceinline__ float vqm_pack(float d, int kk) { return __uint_as_float((__float_as_uint(d) & VQM_KMASK) + (unsigned)kk); }
__device__ __forceinline__ float vqm_val(float p) { return __uint_as_float(__float_as_uint(p) & VQM_KMASK); }
__device__ __forceinline__ int vqm_idx(float p) { return (int)(__float_as_uint(p) & ~VQM_KMASK); }

// one pipeline phase: 4 column groups c0..c0+3 of rows g (vg) and g+8 (vh) = 8 k-steps
template <int NT>
__device__ __forceinline__ void compute_phase(const float4 (&vg)[4], const float4 (&vh)[4], int c0,
                                              const float4* __restrict__ bl, float (&accm)[NT][4],
                                              float (&accc)[NT][4], float& xs_g, float& xs_h) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int s = 2 * (c0 + cc) + u;
      const float a0 = u ? vg[cc].z : vg[cc].x, a2 = u ? vg[cc].w : vg[cc].y;
      const float a1 = u ? vh[cc].z : vh[cc].x, a3 = u ? vh[cc].w : vh[cc].y;
      xs_g = fmaf(a0, a0, xs_g); xs_g = fmaf(a2, a2, xs_g);
      xs_h = fmaf(a1, a1, xs_h); xs_h = fmaf(a3, a3, xs_h);
      // leading term: tf32 MMA on the hi parts; both correction terms (x_lo.c_hi + x_hi.c_lo, 2^-11 of it) in ONE
      // bf16 m16n8k16 MMA whose k-slots 0-7 hold the lo pairs [k=t, k=t+4] and slots 8-15 the hi pairs
      unsigned ah[4], ac[4];
      ah[0] = cvt_tf32(a0); ah[1] = cvt_tf32(a1); ah[2] = cvt_tf32(a2); ah[3] = cvt_tf32(a3);
      const float h0 = __uint_as_float(ah[0]), h1 = __uint_as_float(ah[1]);
      const float h2 = __uint_as_float(ah[2]), h3 = __uint_as_float(ah[3]);
      ac[0] = pack_bf16(a0 - h0, a2 - h2);
      ac[1] = pack_bf16(a1 - h1, a3 - h3);
      ac[2] = pack_bf16(h0, h2);
      ac[3] = pack_bf16(h1, h3);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float4 b = bl[(s * NT + j) * 32];
        mma_tf32(accm[j], ah, __float_as_uint(b.x), __float_as_uint(b.y));
        mma_bf16(accc[j], ac, __float_as_uint(b.z), __float_as_uint(b.w));
      }
    }
  }
}

// One phase = 64 latent columns of the two rows of a lane, as two 32-BYTE loads per row (LDG.256, sm_100): lane (g, t) owns
// columns 32 C + 8 t + [0, 8) of every 32-column block C -- the k order inside the MMAs is free as long as the B fragments
// (stage_chunk) use the same one.  A warp-level load touches 16 rows x 128 contiguous bytes whatever its width, so half the
// load instructions are half the L1 wavefronts of the 16-byte form.  vg[2m], vg[2m+1] = the two halves of block c0/2 + m.
template <bool STREAM>
__device__ __forceinline__ void ldg_f8(const float* p, float4& lo, float4& hi) {
  if (STREAM)
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(p));
  else
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(p));
}
template <bool STREAM>
__device__ __forceinline__ void load_phase(float4 (&vg)[4], float4 (&vh)[4], const float* pg, const float* ph,
                                           int c0) {
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    ldg_f8<STREAM>(pg + 32 * (c0 / 2 + m), vg[2 * m], vg[2 * m + 1]);
    ldg_f8<STREAM>(ph + 32 * (c0 / 2 + m), vh[2 * m], vh[2 * m + 1]);
  }
}

// MODE 0: assignment; MODE 1: global max distance only (first pass of the `thres` path)
template <int NT, int MODE, bool CHUNKED>
__global__ void __launch_bounds__(VqmCfg<NT>::WARPS * 32, VqmCfg<NT>::BLOCKS) vq_assign_mma_kernel(VqParams p) {
  constexpr int KCH = 8 * NT;                    // codewords per chunk
  constexpr int VQM_WARPS = VqmCfg<NT>::WARPS, VQM_THREADS = VQM_WARPS * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // smem: B fragments [32 k-steps][NT][32 lanes] float4 | cnorm[KCH] | top-2 state (chunked) | stats
  float4* bsm = reinterpret_cast<float4*>(smem_raw);
  float* cnorm = reinterpret_cast<float*>(bsm + 32 * NT * 32);
  float* st_b = cnorm + KCH;                     // chunked: [8 warps * TPW * 16 rows] x {b, s, t, bi, si}
  constexpr int ST_ROWS = CHUNKED ? VQM_WARPS * VQM_TPW * 16 : 0;
  float* st_s = st_b + ST_ROWS;
  float* st_t = st_s + ST_ROWS;
  int* st_bi = reinterpret_cast<int*>(st_t + ST_ROWS);
  int* st_si = st_bi + ST_ROWS;
  float* cnt_s = reinterpret_cast<float*>(st_si + ST_ROWS);   // stats: cnt_s[K] | elat_s[4] | dw_s[K*256]
  const int K = p.K;
  const bool smem_dw = p.stats && p.want_dw && K <= 32;
  float* elat_s = cnt_s + (p.stats ? K : 0);
  float* dw_s = elat_s + (p.stats ? 4 : 0);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int nchunks = (K + KCH - 1) / KCH;
  const float INF = 3.0e38f;                     // finite: a packed key must not become a NaN pattern

  if (p.stats) {
    for (int i = tid; i < K + 4 + (smem_dw ? K * VQ_Z : 0); i += VQM_THREADS) cnt_s[i] = 0.f;
  }
  auto stage_chunk = [&](int ch) {
    // B fragments of chunk ch: lane (g,t) of (k-step s = 2c+u, n-tile j) holds cb[z0][k], cb[z0+1][k],
    // z0 = 32 (c / 2) + 8t + 4 (c % 2) + 2u (the column order of load_phase), k = ch*KCH + 8j + g, each split into tf32 hi / lo
    for (int i = tid; i < 32 * NT * 32; i += VQM_THREADS) {
      const int ln = i & 31, j = (i >> 5) % NT, s = i / (32 * NT);
      const int gg = ln >> 2, tt = ln & 3, c = s >> 1, u = s & 1;
      const int z0 = 32 * (c >> 1) + 8 * tt + 4 * (c & 1) + 2 * u, k = ch * KCH + 8 * j + gg;
      float b0 = 0.f, b1 = 0.f;
      if (k < K) { b0 = p.cb[(size_t)z0 * K + k]; b1 = p.cb[(size_t)(z0 + 1) * K + k]; }
      // {b0.hi, b1.hi} as tf32 and the bf16 pairs {bf16(b0.hi), bf16(b1.hi)} (meets the x_lo slots) and
      // {bf16(b0.lo), bf16(b1.lo)} (meets the x_hi slots) of the correction MMA
      const unsigned h0 = cvt_tf32(b0), h1 = cvt_tf32(b1);
      const float fh0 = __uint_as_float(h0), fh1 = __uint_as_float(h1);
      bsm[i] = make_float4(fh0, fh1, __uint_as_float(pack_bf16(fh0, fh1)), __uint_as_float(pack_bf16(b0 - fh0, b1 - fh1)));
    }
    // ||c||^2 (vq_layers.py:282); padded codewords get +inf so they never win
    for (int kk = warp; kk < KCH; kk += VQM_WARPS) {
      const int k = ch * KCH + kk;
      float sq = 0.f;
      if (k < K)
        for (int z = lane; z < VQ_Z; z += 32) { float c = p.cb[(size_t)z * K + k]; sq = fmaf(c, c, sq); }
      sq = warp_sum(sq);
      if (lane == 0) cnorm[kk] = k < K ? sq : INF;
    }
  };

  const long long ntiles = (p.n + 15) / 16;
  float local_max = -INF;
  const float4* bl = bsm + lane;

  // ---- per-tile epilogue --------------------------------------------------------------------------------
  auto epilogue = [&](long long tile, int ch, int slot, float (&accm)[NT][4], float (&accc)[NT][4], float xs_g,
                      float xs_h) {
    const long long row_g = tile * 16 + g, row_h = row_g + 8;
    const bool ok_g = row_g < p.n, ok_h = row_h < p.n;
    xs_g += __shfl_xor_sync(0xffffffffu, xs_g, 1); xs_g += __shfl_xor_sync(0xffffffffu, xs_g, 2);
    xs_h += __shfl_xor_sync(0xffffffffu, xs_h, 1); xs_h += __shfl_xor_sync(0xffffffffu, xs_h, 2);
    float inv_g = 1.f, inv_h = 1.f;
    if (p.normalize) {
      // caller's safe_l2_normalize(z_enc, axis=1) (vq_nfr.py:575) folded in: x.c / |x|, |x|^2 / |x|^2
      inv_g = rsqrtf(fmaxf(xs_g, 1e-6f)); inv_h = rsqrtf(fmaxf(xs_h, 1e-6f));
      xs_g = xs_g * inv_g * inv_g; xs_h = xs_h * inv_h * inv_h;
    }
    Top3 pg = {INF, INF, INF}, ph = {INF, INF, INF};
    float mv = 0.f;
    if (MODE == 0 && p.sel_mask) mv = ord2f(*p.maxdist);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kk = 8 * j + 2 * t + e, k = ch * KCH + kk;
        const float cn = cnorm[kk];                         // +inf for padded codewords
        float dg = xs_g - 2.0f * ((accm[j][e] + accc[j][e]) * inv_g) + cn;
        float dh = xs_h - 2.0f * ((accm[j][2 + e] + accc[j][2 + e]) * inv_h) + cn;
        if (MODE == 1) {
          if (k < K) {
            if (ok_g) local_max = fmaxf(local_max, dg);
            if (ok_h) local_max = fmaxf(local_max, dh);
          }
          continue;
        }
        if (p.sel_mask && k < K) {
          const float sel = p.sel_mask[k];
          dg = dg * sel + mv * (1.0f - sel);                // vq_layers.py:290
          dh = dh * sel + mv * (1.0f - sel);
        }
        if (p.dist_out && k < K) {
          if (ok_g) p.dist_out[row_g * K + k] = dg;
          if (ok_h) p.dist_out[row_h * K + k] = dh;
        }
        top3_push(pg, vqm_pack(dg, kk));
        top3_push(ph, vqm_pack(dh, kk));
      }
    }
    if (MODE == 1) return;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, pg.b, o), os = __shfl_xor_sync(0xffffffffu, pg.s, o);
      float ot = __shfl_xor_sync(0xffffffffu, pg.t, o);
      top3_push(pg, ob); top3_push(pg, os); top3_push(pg, ot);
      ob = __shfl_xor_sync(0xffffffffu, ph.b, o); os = __shfl_xor_sync(0xffffffffu, ph.s, o);
      ot = __shfl_xor_sync(0xffffffffu, ph.t, o);
      top3_push(ph, ob); top3_push(ph, os); top3_push(ph, ot);
    }
    Top2 tg = {vqm_val(pg.b), vqm_val(pg.s), vqm_val(pg.t), ch * KCH + vqm_idx(pg.b), ch * KCH + vqm_idx(pg.s)};
    Top2 th = {vqm_val(ph.b), vqm_val(ph.s), vqm_val(ph.t), ch * KCH + vqm_idx(ph.b), ch * KCH + vqm_idx(ph.s)};
    if (CHUNKED) {
      // running state across codeword chunks (earlier chunk = lower indices, merged first)
      const int sg = slot * 16 + g, sh = sg + 8;
      if (ch > 0) {
        Top2 rg = {st_b[sg], st_s[sg], st_t[sg], st_bi[sg], st_si[sg]};
        Top2 rh = {st_b[sh], st_s[sh], st_t[sh], st_bi[sh], st_si[sh]};
        top2_merge(rg, tg.b, tg.bi, tg.s, tg.si, tg.t); tg = rg;
        top2_merge(rh, th.b, th.bi, th.s, th.si, th.t); th = rh;
      }
      if (ch + 1 < nchunks) {
        __syncwarp();
        if (t == 0) {
          st_b[sg] = tg.b; st_s[sg] = tg.s; st_t[sg] = tg.t; st_bi[sg] = tg.bi; st_si[sg] = tg.si;
          st_b[sh] = th.b; st_s[sh] = th.s; st_t[sh] = th.t; st_bi[sh] = th.bi; st_si[sh] = th.si;
        }
        __syncwarp();
        return;
      }
    }
    // fp64 re-score of near-ties.  The tensor-core distance carries the split error (~2^-20 |x||c|) and the fp32
    // accumulation error, both proportional to |x||c| <= SCALE = |d| + 2 ||x||^2: runner-up within 1.1e-5 SCALE of the best ->
    // exact ordering of the two candidates; THIRD-best within 4e-6 SCALE -> the approximate values cannot name the
    // candidates (three-way near-tie, a few rows per million): every codeword is re-scored in fp64.
    int best_g = tg.bi, best_h = th.bi;
    {
      // |d| + 2 ||x||^2 bounds 2 |x||c| from above up to a small factor without keeping ||c_best||^2 across chunks
      const float sc_g = fmaxf(fabsf(tg.b) + 2.f * xs_g, 1e-3f), sc_h = fmaxf(fabsf(th.b) + 2.f * xs_h, 1e-3f);
      // + 2 * 2^-17 |d|: the packed keys of the hot loop are truncated to 17 mantissa bits
      bool near3_g = ok_g && K > 2 && (tg.t - tg.b) <= 4.0e-6f * sc_g + 1.6e-5f * fmaxf(fabsf(tg.b), fabsf(tg.t));
      bool near3_h = ok_h && K > 2 && (th.t - th.b) <= 4.0e-6f * sc_h + 1.6e-5f * fmaxf(fabsf(th.b), fabsf(th.t));
      bool near_g = ok_g && K > 1 && !near3_g && (tg.s - tg.b) <= 1.1e-5f * sc_g + 1.6e-5f * fmaxf(fabsf(tg.b), fabsf(tg.s));
      bool near_h = ok_h && K > 1 && !near3_h && (th.s - th.b) <= 1.1e-5f * sc_h + 1.6e-5f * fmaxf(fabsf(th.b), fabsf(th.s));
      if (p.sel_mask) {
        // dropped codewords all sit at the mask value (vq_layers.py:290): ties among them keep the first index, which
        // the scan order already delivers; a re-score in exact arithmetic would not reproduce the masked values
        if (near_g && (p.sel_mask[tg.bi] == 0.f || p.sel_mask[tg.si] == 0.f)) near_g = false;
        if (near_h && (p.sel_mask[th.bi] == 0.f || p.sel_mask[th.si] == 0.f)) near_h = false;
        if (near3_g && p.sel_mask[tg.bi] == 0.f) near3_g = false;
        if (near3_h && p.sel_mask[th.bi] == 0.f) near3_h = false;
      }
      unsigned need_g = __ballot_sync(0xffffffffu, near_g && t == 0);
      unsigned need_h = __ballot_sync(0xffffffffu, near_h && t == 0);
      unsigned need3_g = __ballot_sync(0xffffffffu, near3_g && t == 0);
      unsigned need3_h = __ballot_sync(0xffffffffu, near3_h && t == 0);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        unsigned need = half ? need_h : need_g;
        while (need) {                                           // warp-uniform
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const long long row = tile * 16 + (src >> 2) + 8 * half;
          const int i1 = __shfl_sync(0xffffffffu, half ? th.bi : tg.bi, src);
          const int i2 = __shfl_sync(0xffffffffu, half ? th.si : tg.si, src);
          const float inv = __shfl_sync(0xffffffffu, half ? inv_h : inv_g, src);
          double d1 = 0.0, d2 = 0.0;
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const int z = lane + 32 * m;
            const double xv = (double)(p.x[row * VQ_Z + z] * inv);   // the fp32 latent the reference sees
            const double c1 = (double)p.cb[(size_t)z * K + i1], c2 = (double)p.cb[(size_t)z * K + i2];
            d1 += c1 * c1 - 2.0 * xv * c1;
            d2 += c2 * c2 - 2.0 * xv * c2;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
          }
          const bool swap = (d2 < d1) || (d2 == d1 && i2 < i1);
          if (swap && (lane >> 2) == (src >> 2)) { if (half) best_h = i2; else best_g = i2; }
        }
        need = half ? need3_h : need3_g;
#pragma unroll 1
        while (need) {                                           // warp-uniform: every (not dropped) codeword in fp64
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const long long row = tile * 16 + (src >> 2) + 8 * half;
          const float inv = __shfl_sync(0xffffffffu, half ? inv_h : inv_g, src);
          double bestd = 1.0e300;
          int besti = 0x7fffffff;
#pragma unroll 1
          for (int k = lane; k < K; k += 32) {
            if (p.sel_mask && p.sel_mask[k] == 0.f) continue;    // sits at the mask value, cannot beat a kept codeword
            double d = 0.0;
#pragma unroll 4
            for (int z = 0; z < VQ_Z; ++z) {
              const double xv = (double)(p.x[row * VQ_Z + z] * inv), c = (double)p.cb[(size_t)z * K + k];
              d += c * c - 2.0 * xv * c;
            }
            if (d < bestd) { bestd = d; besti = k; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bestd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (od < bestd || (od == bestd && oi < besti)) { bestd = od; besti = oi; }
          }
          if ((lane >> 2) == (src >> 2)) { if (half) best_h = besti; else best_g = besti; }
        }
      }
    }
    if (p.idx_out && t == 0) {
      if (ok_g) p.idx_out[row_g] = (long long)best_g;
      if (ok_h) p.idx_out[row_h] = (long long)best_h;
    }
    // second pass (training / quantize consumers): row-wise re-read (L2 hits), lane l owns z {4l.., 128+4l..}
    if (p.quant_out || p.znorm_out || p.stats) {
      // (the next row's latent is requested before the current row is processed: the loop is a serial chain of L2 round trips
      // otherwise -- 16 rows x ~1 us per warp at 8192 training rows)
      float4 a_nx = make_float4(0.f, 0.f, 0.f, 0.f), b_nx = a_nx;
      if (tile * 16 < p.n) {
        a_nx = __ldg(reinterpret_cast<const float4*>(p.x + tile * 16 * VQ_Z + 4 * lane));
        b_nx = __ldg(reinterpret_cast<const float4*>(p.x + tile * 16 * VQ_Z + 128 + 4 * lane));
      }
#pragma unroll 1
      for (int rr = 0; rr < 16; ++rr) {
        const long long row = tile * 16 + rr;
        if (row >= p.n) break;                                   // warp-uniform
        const int src = (rr & 7) * 4;
        const int idx = __shfl_sync(0xffffffffu, rr < 8 ? best_g : best_h, src);
        const float inv = __shfl_sync(0xffffffffu, rr < 8 ? inv_g : inv_h, src);
        const float4 a = a_nx, b = b_nx;
        if (rr + 1 < 16 && row + 1 < p.n) {
          a_nx = __ldg(reinterpret_cast<const float4*>(p.x + (row + 1) * VQ_Z + 4 * lane));
          b_nx = __ldg(reinterpret_cast<const float4*>(p.x + (row + 1) * VQ_Z + 128 + 4 * lane));
        }
        float xv[8] = {a.x * inv, a.y * inv, a.z * inv, a.w * inv, b.x * inv, b.y * inv, b.z * inv, b.w * inv};
        float q[8];
        float e = 0.f;
#pragma unroll
        for (int zi = 0; zi < 8; ++zi) {
          const int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
          const float c = p.cb[(size_t)z * K + idx];            // quantize(): embedding_lookup(codebook^T, idx)
          const float diff = c - xv[zi];
          e = fmaf(diff, diff, e);                               // (sg(quantized) - inputs)^2, :302
          q[zi] = xv[zi] + diff;                                 // inputs + sg(quantized - inputs), :327
        }
        if (p.znorm_out) {
          float* po = p.znorm_out + row * VQ_Z;
          *reinterpret_cast<float4*>(po + 4 * lane) = make_float4(xv[0], xv[1], xv[2], xv[3]);
          *reinterpret_cast<float4*>(po + 128 + 4 * lane) = make_float4(xv[4], xv[5], xv[6], xv[7]);
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
              const int z = (zi < 4 ? 0 : 128) + 4 * lane + (zi & 3);
              if (smem_dw) atomicAdd(&dw_s[idx * VQ_Z + z], xv[zi]);
              else atomicAdd(&p.stats[K + 2 + (size_t)z * K + idx], (double)xv[zi]);
            }
          }
        }
      }
    }
  };

  // ---- pipelined sweep over a sequence of tiles: first, first+step, ... (count tiles) --------------------
  auto sweep = [&](long long first, long long step, int count, int ch, int slot0) {
    if (count <= 0) return;
    float4 ag[4], ah[4], bg[4], bh[4];
    auto rowptr = [&](long long tile, int add) {
      long long r = tile * 16 + g + add;
      if (r >= p.n) r = p.n - 1;                                 // clamp: loaded, never used
      return p.x + r * VQ_Z + 8 * t;
    };
    const float* pg = rowptr(first, 0);
    const float* ph = rowptr(first, 8);
    load_phase<!CHUNKED>(ag, ah, pg, ph, 0);
    load_phase<!CHUNKED>(bg, bh, pg, ph, 4);
    long long tile = first;
#pragma unroll 1
    for (int i = 0; i < count; ++i) {
      float accm[NT][4], accc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { accm[j][e] = 0.f; accc[j][e] = 0.f; }
      float xs_g = 0.f, xs_h = 0.f;
      const long long next = tile + step;
      const bool more = i + 1 < count;
      const float* ng = more ? rowptr(next, 0) : pg;
      const float* nh = more ? rowptr(next, 8) : ph;
      compute_phase<NT>(ag, ah, 0, bl, accm, accc, xs_g, xs_h);
      load_phase<!CHUNKED>(ag, ah, pg, ph, 8);
      compute_phase<NT>(bg, bh, 4, bl, accm, accc, xs_g, xs_h);
      load_phase<!CHUNKED>(bg, bh, pg, ph, 12);
      compute_phase<NT>(ag, ah, 8, bl, accm, accc, xs_g, xs_h);
      if (more) load_phase<!CHUNKED>(ag, ah, ng, nh, 0);
      compute_phase<NT>(bg, bh, 12, bl, accm, accc, xs_g, xs_h);
      if (more) load_phase<!CHUNKED>(bg, bh, ng, nh, 4);
      epilogue(tile, ch, slot0 + i, accm, accc, xs_g, xs_h);
      tile = next; pg = ng; ph = nh;
    }
  };

  if (!CHUNKED) {
    stage_chunk(0);
    __syncthreads();
    const long long warps_total = (long long)gridDim.x * VQM_WARPS;
    const long long gw = (long long)blockIdx.x * VQM_WARPS + warp;
    const int count = ntiles > gw ? (int)((ntiles - gw + warps_total - 1) / warps_total) : 0;
    sweep(gw, warps_total, count, 0, 0);
  } else {
    const long long tiles_per_batch = VQM_WARPS * VQM_TPW;
    const long long nbatches = (ntiles + tiles_per_batch - 1) / tiles_per_batch;
    for (long long batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
      const long long first = batch * tiles_per_batch + warp * VQM_TPW;
      long long rem = ntiles - first;
      const int count = rem <= 0 ? 0 : (rem < VQM_TPW ? (int)rem : VQM_TPW);
      for (int ch = 0; ch < nchunks; ++ch) {
        __syncthreads();
        stage_chunk(ch);
        __syncthreads();
        sweep(first, 1, count, ch, warp * VQM_TPW);
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
    for (int k = tid; k < K; k += VQM_THREADS)
      if (cnt_s[k] != 0.f) atomicAdd(&p.stats[k], (double)cnt_s[k]);
    if (tid == 0) { atomicAdd(&p.stats[K], (double)elat_s[0]); atomicAdd(&p.stats[K + 1], (double)elat_s[1]); }
    if (smem_dw)
      for (int i = tid; i < K * VQ_Z; i += VQM_THREADS) {
        const int k = i / VQ_Z, z = i % VQ_Z;
        const float v = dw_s[i];
        if (v != 0.f) atomicAdd(&p.stats[K + 2 + (size_t)z * K + k], (double)v);
      }
  }
}

template <int NT, bool CHUNKED>
static int launch_nt(vqn_ctx* ctx, VqParams p, cudaStream_t s) {
  constexpr int VQM_WARPS = VqmCfg<NT>::WARPS, VQM_THREADS = VQM_WARPS * 32;
  const int K = p.K;
  if (reinterpret_cast<uintptr_t>(p.x) & 31) {
    vqn_set_error("vq_assign: inputs must be 32-byte aligned (32-byte row loads)");
    return VQN_ERR_INVALID_ARG;
  }
  const bool smem_dw = p.stats && p.want_dw && K <= 32;
  size_t smem = sizeof(float4) * 32 * NT * 32 + sizeof(float) * 8 * NT +
                (CHUNKED ? sizeof(float) * 5 * VQM_WARPS * VQM_TPW * 16 : 0) +
                sizeof(float) * (p.stats ? K + 4 + (smem_dw ? (size_t)K * VQ_Z : 0) : 0);
  const long long ntiles = (p.n + 15) / 16;
  long long want = CHUNKED ? (ntiles + VQM_WARPS * VQM_TPW - 1) / (VQM_WARPS * VQM_TPW)
                           : (ntiles + VQM_WARPS - 1) / VQM_WARPS;
  int per_sm = VqmCfg<NT>::BLOCKS;
  while (per_sm > 1 && (smem + 1024) * per_sm > 220 * 1024) --per_sm;
  long long cap = (long long)ctx->sm_count * per_sm;
  const int blocks = (int)(want < cap ? want : cap);
  auto k0 = vq_assign_mma_kernel<NT, 0, CHUNKED>;
  auto k1 = vq_assign_mma_kernel<NT, 1, CHUNKED>;
  VQN_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (p.sel_mask) {
    unsigned* md = reinterpret_cast<unsigned*>(ctx->scratch);        // persistent scratch slot 0
    VQN_CUDA(cudaMemsetAsync(md, 0, sizeof(unsigned), s));           // ordered 0 == most negative
    p.maxdist = md;
    VQN_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1<<<blocks, VQM_THREADS, smem, s>>>(p);
    VQN_LAUNCHED(ctx);
  }
  k0<<<blocks, VQM_THREADS, smem, s>>>(p);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

int vq_assign_mma_launch(vqn_ctx* ctx, VqParams p, cudaStream_t s) {
  if (p.K <= 8) return launch_nt<1, false>(ctx, p, s);
  if (p.K <= 16) return launch_nt<2, false>(ctx, p, s);
  if (p.K <= 32) return launch_nt<4, false>(ctx, p, s);
  if (p.K <= 64) return launch_nt<8, false>(ctx, p, s);
  return launch_nt<8, true>(ctx, p, s);
}
