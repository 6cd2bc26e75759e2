// Light-integral kernel, thread-per-point form (large batches: full-image relighting).
//
// Reference: models/shape.py:103-119, models/vq_nfr.py:694-733,830-874, util/microfacet.py:9-89 -- the same
// arithmetic as shade_kernel (shade.cu), with the roles of lanes swapped: a THREAD owns one surface point and
// walks the 512 lights, so the per-light data (light position, probe radiance x area) is warp-UNIFORM: every
// shared-memory read is a broadcast (one wavefront for 32 points).  Compared with the warp-per-point kernel
// this removes (i) its shared-memory bottleneck -- there 27 LDS.128 per 4 lights serve ONE point, i.e. 4x the
// 128 B/clk smem port at full FMA rate; here the same 27 reads serve 32 points --, (ii) the cross-lane
// reductions and (iii) the 48 registers of per-light weights: the 3 x (1+P) sums stay in the thread's registers.
// The point's visibility row is read as one 16-byte load per 4 lights (every 128-byte line is consumed by 8
// consecutive loads of the same thread and stays in L1 meanwhile).  (A first version kept the tables in the
// constant bank: 61 KB swept by 24 warps at different offsets thrashes the constant cache -- 6.7 ms vs 2.6 ms.)
//
// Light image (<= 63 488 B, built on the launching stream before every launch since the model light is trainable):
// per group of 4 lights, 3 float4 of light positions (x, y, z of the 4 lights), then per light NF4 float4 of
// radiance x area arranged as output PAIRS: pair m < NPC = (probe m: channel 0, channel 1), pair NPC + j = channel 2
// of probes (2j, 2j+1) (zero padded).  The 3 (1+P) sums are accumulated pairwise with the packed fma.rn.f32x2
// (FFMA2): the kernel is bound by issue slots, not by the FMA pipe, and the pairs halve the issue slots of the
// accumulation (27 -> 14 per light for 9 probes); the per-light factor is (e0, e1) or (e2, e2); an LDS.128 delivers
// two radiance pairs already in aligned 64-bit registers.  One persistent
// 512-thread block per SM copies the image into shared memory once and walks 32-point tiles, warp-strided.
#include <cuda_fp16.h>
#include "common.cuh"

#define SP_L 512
#define SP_GROUPS (SP_L / 4)
#define SP_MAXP 9
#define SP_NPAIR(npc) ((npc) + ((npc) + 1) / 2)         // output pairs: (c0, c1) per probe + channel 2 of two probes
#define SP_NF4(npc) ((SP_NPAIR(npc) + 1) / 2)           // float4 of radiance per light
#define SP_F4_PER_GROUP(npc) (3 + 4 * SP_NF4(npc))
#define SP_F4_MAX SP_F4_PER_GROUP(SP_MAXP)

#define SP_THREADS 512

namespace {

__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(u64& acc, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

// staging image in device memory, nf4 = SP_NF4(NPC) of the kernel instance that will read it
__global__ void shade_pt_prep_kernel(const float* __restrict__ lxyz, const float* __restrict__ lareas,
                                     const float* __restrict__ lights, int n_probes, int clip_light0, int npc,
                                     int nf4, float4* __restrict__ img) {
  const int per_group = 3 + 4 * nf4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // one float4 of the image
  if (i >= SP_GROUPS * per_group) return;
  const int grp = i / per_group, k = i % per_group;
  float v[4];
  if (k < 3) {
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = lxyz[3 * (4 * grp + q) + k];
  } else {
    const int l = 4 * grp + (k - 3) / nf4, j = (k - 3) % nf4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int m = 2 * j + (c >> 1);                          // output pair (see the header comment)
      const int p = m < npc ? m : 2 * (m - npc) + (c & 1), ch = m < npc ? (c & 1) : 2;
      float r = 0.f;
      if (p < n_probes) {
        r = lights[((size_t)p * SP_L + l) * 3 + ch];
        if (p == 0 && clip_light0) r = fmaxf(r, 0.f);          // clip(_light, 0, inf), vq_nfr.py:759
        r *= lareas[l];
      }
      v[c] = r;
    }
  }
  img[i] = make_float4(v[0], v[1], v[2], v[3]);
}

// LV: light-visibility format -- 0 none, 1 float32 (the reference's lvis.npy), 2 float16, 3 uint8 (v = q / 255): the
// compact formats are an opt-in of the host-buffer path (the fp32 rows are 98 % of a view's H2D bytes)
template <int NPC, int LV, bool SPLIT>
__global__ void __launch_bounds__(SP_THREADS, 1) shade_pt_kernel(vqn_shade_args a, const float4* __restrict__ img,
                                                                 int* nonfinite, float* __restrict__ part_buf,
                                                                 unsigned* __restrict__ part_cnt) {
  extern __shared__ __align__(16) float4 s_img[];
  constexpr int NF4 = SP_NF4(NPC), PER_GROUP = SP_F4_PER_GROUP(NPC), NPAIR = 2 * NF4;   // NPAIR includes the padding pair
  for (int k = threadIdx.x; k < SP_GROUPS * PER_GROUP; k += SP_THREADS) s_img[k] = img[k];
  // per-warp staging of the 32 x (3 NPC) outputs of a tile for the coalesced peer stores (fused gather)
  float* s_out = reinterpret_cast<float*>(s_img + SP_GROUPS * PER_GROUP) + (threadIdx.x >> 5) * (32 * (3 * NPC + 1));
  __syncthreads();
  long long n = a.n_dev ? (long long)*a.n_dev : a.n;
  if (n > a.n) n = a.n;
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * (SP_THREADS / 32);
  const long long warp0 = (long long)blockIdx.x * (SP_THREADS / 32) + (threadIdx.x >> 5);
  // Work units.  T tiles of 32 points over W persistent warps: the first F = T / W rounds are whole tiles; the R = T - F W
  // tiles of the last, partial round are cut into S light ranges each (S = the largest power of two <= 16 with
  // R S <= W), so that the tail costs 1/S of a round instead of a whole one -- at 80 000 points per GPU (an 800x800 view
  // over 8 GPUs) that is 2500 tiles on 2368 warps: 1.06 rounds of work used to take 2.  The S partial sums of a tile meet
  // in global scratch; the LAST warp to arrive adds them in a fixed order (deterministic) and runs the epilogue.
  // (SPLIT is chosen by the host from the row count it knows: only when the last round would fill less than a quarter of
  // the warps -- a fuller tail round already runs at the issue rate, and the instantiation without the split keeps the
  // constant loop bounds.)
  const long long T = (n + 31) / 32;
  const long long whole = SPLIT ? (T / warps_total) * warps_total : T;      // tiles processed whole
  const long long R = T - whole;
  int S = 1;
  if (SPLIT && R > 0) { const long long q = warps_total / R; S = q >= 16 ? 16 : q >= 8 ? 8 : q >= 4 ? 4 : 1; }
  for (long long u = warp0; u < whole + R * S; u += warps_total) {
  long long tile = u;
  int g0 = 0, g1 = SP_GROUPS, part = 0;
  const bool split = SPLIT && u >= whole && S > 1;
  if (SPLIT && u >= whole) {
    const long long ru = u - whole;
    tile = whole + ru / S;
    part = (int)(ru % S);
    g0 = part * (SP_GROUPS / S); g1 = g0 + SP_GROUPS / S;
  }
  const long long i_raw = tile * 32 + lane;
  const bool live = i_raw < n;
  const long long i = live ? i_raw : n - 1;                     // dead lanes of the last tile recompute row n-1
  const long long row = a.row_idx ? (long long)a.row_idx[i] : i;
  const float INV_PI = 0.318309886183790671538f;
  const float px = a.xyz[row * 3], py = a.xyz[row * 3 + 1], pz = a.xyz[row * 3 + 2];
  float vx = a.rayo[row * 3] - px, vy = a.rayo[row * 3 + 1] - py, vz = a.rayo[row * 3 + 2] - pz;
  {  // _calc_vdir: safe_l2_normalize (eps on the squared norm)
    const float inv = rsqrtf(fmaxf(vx * vx + vy * vy + vz * vz, 1e-6f));
    vx *= inv; vy *= inv; vz *= inv;
  }
  float nx = a.normal[row * 3], ny = a.normal[row * 3 + 1], nz = a.normal[row * 3 + 2];
  {  // _normal_correct: where(n.v >= 0, n, -n)
    const float c = nx * vx + ny * vy + nz * vz;
    if (!(c >= 0.f)) { nx = -nx; ny = -ny; nz = -nz; }
  }
  if (a.normal_out && live) { a.normal_out[row * 3] = nx; a.normal_out[row * 3 + 1] = ny; a.normal_out[row * 3 + 2] = nz; }
  const float inv_n = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-6f));
  const float vn = (nx * vx + ny * vy + nz * vz) * inv_n;
  const float alb0 = a.albedo[i * 3] * INV_PI, alb1 = a.albedo[i * 3 + 1] * INV_PI, alb2 = a.albedo[i * 3 + 2] * INV_PI;
  const float f00 = a.spec[i * 3], f01 = a.spec[i * 3 + 1], f02 = a.spec[i * 3 + 2];
  const float rough = a.rough[i];
  const float alpha = rough * rough, a2 = alpha * alpha;        // microfacet.py:25; helpers use alpha ** 2
  const float oma2 = 1.0f - a2, a2m1 = a2 - 1.0f;
  const float cv = fminf(fmaxf(vn, 0.f), 1.f);
  const float den_v = cv + sqrtf(fabsf(a2 + oma2 * cv * cv));
  const float g_v = den_v == 0.f ? 0.f : 2.0f * cv / den_v;
  const float avn = fabsf(vn);
  const float a_pt = avn == 0.f ? 0.f : a2 * g_v * (0.5f * INV_PI) / avn;

  u64 acc[NPAIR];                                               // (out[2m], out[2m+1]) pairs, fp32 x 2
#pragma unroll
  for (int m = 0; m < NPAIR; ++m) acc[m] = 0ull;
  constexpr bool HAS_LVIS = LV != 0;
  // one group = 4 lights = 16 / 8 / 4 bytes of the point's visibility row
  const unsigned char* lvb = reinterpret_cast<const unsigned char*>(a.lvis) +
                             (HAS_LVIS ? (size_t)row * SP_L * (LV == 1 ? 4 : LV == 2 ? 2 : 1) : 0);
  auto lv_load = [&](int g) -> float4 {
    if (LV == 1) return __ldg(reinterpret_cast<const float4*>(lvb) + g);
    if (LV == 2) {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(lvb) + g);
      const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
      const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
      return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    // uint8: byte b -> the float 2^23 + b by a byte permute into the mantissa; the subtraction of 2^23 is exact, then one
    // multiply (an FMA with the pre-scaled offset -2^23/255 would round that offset to 0.002 -- half a quantisation step)
    const unsigned raw = __ldg(reinterpret_cast<const unsigned*>(lvb) + g);
    const float k = 1.0f / 255.0f, o = 8388608.0f;
    return make_float4((__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7650)) - o) * k,
                       (__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7651)) - o) * k,
                       (__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7652)) - o) * k,
                       (__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7653)) - o) * k);
  };
  float4 lv_cur = make_float4(1.f, 1.f, 1.f, 1.f);
  if (HAS_LVIS) lv_cur = lv_load(g0);

#pragma unroll 1
  for (int grp = g0; grp < g1; ++grp) {
    float4 lv_nxt = lv_cur;
    if (HAS_LVIS && grp + 1 < g1) lv_nxt = lv_load(grp + 1);
    const float4* cg = s_img + grp * PER_GROUP;                // warp-uniform address: broadcast reads
    const float4 X = cg[0], Y = cg[1], Z = cg[2];
    const float xs4[4] = {X.x, X.y, X.z, X.w}, ys4[4] = {Y.x, Y.y, Y.z, Y.w}, zs4[4] = {Z.x, Z.y, Z.z, Z.w};
    const float lvv[4] = {lv_cur.x, lv_cur.y, lv_cur.z, lv_cur.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float dx = xs4[q] - px, dy = ys4[q] - py, dz = zs4[q] - pz;
      const float inv = fast_rsqrt(fmaxf(dx * dx + dy * dy + dz * dz, 1e-6f));   // _calc_ldir: l = d * inv
      const float cos_r = (dx * nx + dy * ny + dz * nz) * inv;  // _render: cos = l . n
      const float ln = cos_r * inv_n;                           // get_brdf: l . normalize(n)
      // h = normalize(l + v), formed componentwise as the reference does (microfacet.py:21-22).  The shortcut
      // |l + v|^2 = 2 + 2 l.v is cheaper but its rounding error is amplified by 2/q in q = 1 - (h.n)^2 (1 - a^2)
      // near the highlight (h ~ n, small roughness): 2e-4 relative on the specular lobe, outside the parity budget.
      const float hx = fmaf(dx, inv, vx), hy = fmaf(dy, inv, vy), hz = fmaf(dz, inv, vz);
      const float hi_ = fast_rsqrt(fmaxf(hx * hx + hy * hy + hz * hz, 1e-6f));
      const float hvr = (hx * vx + hy * vy + hz * vz) * hi_;
      const float hnr = (hx * nx + hy * ny + hz * nz) * (inv_n * hi_);
      const float hv = __saturatef(hvr);                        // clip(h . v, 0, 1)
      const float hn = __saturatef(hnr);                        // clip(h . n, 0, 1)
      const float om = 1.0f - hv, om2 = om * om;
      const float p5 = om2 * om2 * om;                          // (1 - h.v)^5
      const float q_ = fmaf(hn * hn, a2m1, 1.0f);
      const float cl = __saturatef(ln);
      const float den_l = cl + fast_sqrt(fabsf(fmaf(oma2, cl * cl, a2)));
      // glossy = F G D / (4 |l.n| |v.n|) with g(l) = 2 cl / den_l: the factor cl / |l.n| is 1 for every front-lit
      // light (0 < l.n <= 1) and the term is multiplied by front_lit * cos below, so the division by |l.n| and the
      // divide_no_nan guard drop out; q_^2 den_l >= a2^2.5 > 0 (the floor only matters for rough == 0, where a_pt == 0)
      const float S = a_pt * fast_rcp(fmaxf(q_ * q_ * den_l, 1e-30f));
      float wv = fmaxf(cos_r, 0.f);                             // front_lit * cos
      if (HAS_LVIS) wv *= lvv[q];
      const float sw = S * wv, psw = p5 * sw, dsw = sw - psw;   // F S w = psw + f0 (sw - psw)
      const float e0 = fmaf(alb0, wv, fmaf(f00, dsw, psw));
      const float e1 = fmaf(alb1, wv, fmaf(f01, dsw, psw));
      const float e2 = fmaf(alb2, wv, fmaf(f02, dsw, psw));
      const u64 e01 = pack2(e0, e1), e22 = pack2(e2, e2);
      const ulonglong2* rq = reinterpret_cast<const ulonglong2*>(cg + 3 + q * NF4);
#pragma unroll
      for (int j = 0; j < NF4; ++j) {
        const ulonglong2 R = rq[j];                             // radiance x area of output pairs 2j, 2j+1 for this light
        ffma2(acc[2 * j], 2 * j < NPC ? e01 : e22, R.x);
        ffma2(acc[2 * j + 1], 2 * j + 1 < NPC ? e01 : e22, R.y);
      }
    }
    lv_cur = lv_nxt;
  }
  const int NP = a.n_probes;
  float out[3 * NPC];
#pragma unroll
  for (int p = 0; p < NPC; ++p) {
    unpack2(acc[p], out[3 * p], out[3 * p + 1]);
    float lo, hi;
    unpack2(acc[NPC + p / 2], lo, hi);
    out[3 * p + 2] = (p & 1) ? hi : lo;
  }
  if (split) {
    // this warp's light range of the tile -> scratch [remainder tile][part][k][lane]; the last of the S warps finishes the tile
    const long long rt = tile - whole;
    float* pb = part_buf + ((size_t)rt * 16 + part) * (32 * 3 * SP_MAXP);
#pragma unroll
    for (int k = 0; k < 3 * NPC; ++k) pb[k * 32 + lane] = out[k];
    __threadfence();
    unsigned old = 0;
    if (lane == 0) old = atomicAdd(&part_cnt[rt], 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != (unsigned)(S - 1)) continue;                            // warp-uniform: not the last one
    __threadfence();
    if (lane == 0) part_cnt[rt] = 0;                                   // ready for the next launch
    const float* p0 = part_buf + (size_t)rt * 16 * (32 * 3 * SP_MAXP);
    // q outer, k inner: the 3 NPC loads of a part are independent and in flight together (one L2 latency per part; with
    // k outer the 16 x 27 loads ran back to back -- 50 us for the tail of ONE tile); the order of the sum over q is fixed
#pragma unroll
    for (int k = 0; k < 3 * NPC; ++k) out[k] = 0.f;
    for (int q = 0; q < S; ++q) {
      const float* pq = p0 + (size_t)q * (32 * 3 * SP_MAXP) + lane;
#pragma unroll
      for (int k = 0; k < 3 * NPC; ++k) out[k] += __ldcg(pq + k * 32);
    }
  }
#pragma unroll
  for (int p = 0; p < NPC; ++p) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float v = out[p * 3 + ch];
      if (a.use_gamma) v = powf(v * a.gamma_bias, a.gamma_index);     // vq_nfr.py:715-716
      if (p < NP && live && !isfinite(v)) atomicOr(nonfinite, 2);      // check_numerics (:731)
      v = fminf(fmaxf(v, 0.f), 1.f);                                   // clip_by_value (:718)
      if (a.to_srgb) v = vqn_linear2srgb(v);
      out[p * 3 + ch] = v;
      if (a.rgb && p < NP && live) a.rgb[(row * NP + p) * 3 + ch] = v;
    }
  }
  if (a.n_peers > 0) {
    // fused gather: stage the tile's rows, then store each row (3 NP contiguous floats) coalesced into the image
    // buffer of EVERY rank (P2P stores over NVLink; the local rank is one of the peers)
    constexpr int ST = 3 * NPC + 1;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 3 * NPC; ++k) s_out[lane * ST + k] = out[k];
    __syncwarp();
    const int w = 3 * NP;
    for (int pt = 0; pt < 32; ++pt) {
      const long long r_pt = __shfl_sync(0xffffffffu, row, pt);
      const bool live_pt = __shfl_sync(0xffffffffu, (int)live, pt) != 0;
      if (!live_pt) break;                                             // warp-uniform: dead lanes are at the end
      if (lane < w) {
        const float v = s_out[pt * ST + lane];
        const long long off = (a.peer_row0 + r_pt) * w + lane;
        for (int q = 0; q < a.n_peers; ++q) a.peer_rgb[q][off] = v;
      }
    }
  }
  }   // work-unit loop
}

template <int NPC>
int launch_pt(vqn_ctx* ctx, const vqn_shade_args& a, const float4* img, cudaStream_t s) {
  const size_t smem = sizeof(float4) * SP_GROUPS * SP_F4_PER_GROUP(NPC) + sizeof(float) * (SP_THREADS / 32) * 32 * (3 * NPC + 1);
  long long want = (a.n + SP_THREADS - 1) / SP_THREADS;
  const unsigned blocks = (unsigned)(want < (long long)ctx->sm_count ? want : (long long)ctx->sm_count);
  const int lv = !a.lvis ? 0 : a.lvis_format == VQN_LVIS_F16 ? 2 : a.lvis_format == VQN_LVIS_U8 ? 3 : 1;
  // scratch of the split last round: <= W / S tiles x 16 parts x [27][32] partial sums, + one arrival counter per tile
  const size_t warps = (size_t)blocks * (SP_THREADS / 32);
  const size_t cnt_bytes = (warps * sizeof(unsigned) + 255) / 256 * 256;
  const size_t part_bytes = warps / 2 * 16 * (32 * 3 * SP_MAXP) * sizeof(float);
  unsigned char* sc = static_cast<unsigned char*>(vqn_stream_scratch(ctx, VQN_SCRATCH_SHADE, s, cnt_bytes + part_bytes));
  if (!sc) return VQN_ERR_CUDA;
  unsigned* part_cnt = reinterpret_cast<unsigned*>(sc);
  float* part_buf = reinterpret_cast<float*>(sc + cnt_bytes);
  // split the last round only when it would be poorly occupied (see the kernel)
  const long long T_host = (a.n + 31) / 32, R_host = T_host % (long long)warps;
  const bool do_split = R_host > 0 && (long long)warps / R_host >= 4;
  if (do_split) VQN_CUDA(cudaMemsetAsync(part_cnt, 0, cnt_bytes, s));
#define SP_LAUNCH2(LVV, SPL)                                                                                            \
  do {                                                                                                                  \
    VQN_CUDA(cudaFuncSetAttribute(shade_pt_kernel<NPC, LVV, SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    shade_pt_kernel<NPC, LVV, SPL><<<blocks, SP_THREADS, smem, s>>>(a, img, ctx->nonfinite_flag, part_buf, part_cnt);   \
  } while (0)
#define SP_LAUNCH(LVV) do { if (do_split) SP_LAUNCH2(LVV, true); else SP_LAUNCH2(LVV, false); } while (0)
  if (lv == 0) SP_LAUNCH(0);
  else if (lv == 1) SP_LAUNCH(1);
  else if (lv == 2) SP_LAUNCH(2);
  else SP_LAUNCH(3);
#undef SP_LAUNCH2
#undef SP_LAUNCH
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

}  // namespace

// called by vqn_shade (shade.cu) for large un-split batches with at most SP_MAXP probes
int vqn_shade_pt_launch(vqn_ctx* ctx, const vqn_shade_args& a, cudaStream_t s) {
  // staging image in the upper half of the context's persistent scratch (stream-ordered: prep, then the kernel)
  float4* img = reinterpret_cast<float4*>(ctx->scratch + 32768);
  const int npc = a.n_probes <= 1 ? 1 : a.n_probes <= 3 ? 3 : a.n_probes <= 5 ? 5 : SP_MAXP;
  const int total = SP_GROUPS * SP_F4_PER_GROUP(npc);
  shade_pt_prep_kernel<<<(total + 127) / 128, 128, 0, s>>>(a.lxyz, a.lareas, a.lights, a.n_probes, a.clip_light0,
                                                           npc, SP_NF4(npc), img);
  VQN_LAUNCHED(ctx);
  if (npc == 1) return launch_pt<1>(ctx, a, img, s);
  if (npc == 3) return launch_pt<3>(ctx, a, img, s);
  if (npc == 5) return launch_pt<5>(ctx, a, img, s);
  return launch_pt<SP_MAXP>(ctx, a, img, s);
}
