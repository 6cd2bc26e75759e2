// Light-integral ("shade") kernel: _calc_ldir + _calc_vdir + _normal_correct + get_brdf + _render fused.
//
// Reference: models/shape.py:103-119, models/vq_nfr.py:694-733,830-874, util/microfacet.py:9-89.
// The reference materialises >= 15 tensors of shape [N,512,3]; here nothing per-(point,light) ever
// leaves registers.
//
// Mapping: one warp per surface point, lane l owns lights {128 j + 4 l + i : j,i in 0..3} so that
//   - the point's lvis row (2 KB) is four fully coalesced 512 B LDG.128 requests (streamed, prefetched one
//     point ahead),
//   - light geometry and probe radiance are SoA in shared memory, read as conflict-free LDS.128 that each
//     deliver four lights.
// Phase 1 computes, per light, the probe-independent per-channel weight
//     e_ch = (F_ch S + albedo_ch / pi) * w,   w = front * lvis * cos   (area is folded into the radiance
//     tables), S = D G / (4 |l.n| |v.n|), F_ch = f0_ch + (1 - f0_ch)(1 - h.v)^5
// and keeps it in registers (48 per lane).  Phase 2 is then 3 FFMA per (light, probe):
//     rgb[p][ch] += e_ch * L[p][l][ch]
// followed by a halving-butterfly reduction of the 3 values per probe across the warp.  Algorithmic HBM bytes per point: 2048 (lvis) + 36 + 28 + 12 (1 + P).
struct vqn_peer_ptrs { float* p[8]; int n; };
#include "common.cuh"

#define SH_L 512
#define SH_THREADS 256
#define SH_PC 4  // probes per accumulation pass

struct ShadeParams {
  vqn_shade_args a;
  int* nonfinite;
};

// single-MUFU approximations (<= 2 ulp): the per-light chain tolerates them within the 1e-4 budget
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ float gsub_f(float c, float a2) {
  // util/microfacet.py:49-69: 2c / (c + sqrt(|a2 + (1-a2) c^2|)), divide_no_nan
  float den = c + sqrtf(fabsf(a2 + (1.0f - a2) * c * c));
  return den == 0.f ? 0.f : 2.0f * c / den;
}

// halving butterfly over lane bit `bit`: NV values -> NV/2 (see vq.cu)
template <int NV>
__device__ __forceinline__ void sh_halve(float* v, int lane, int bit) {
  const bool hi = (lane >> bit) & 1;
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) {
    float send = hi ? v[i] : v[i + NV / 2];
    float keep = hi ? v[i + NV / 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << bit);
  }
}

template <bool HAS_LVIS, bool SPLIT>
__global__ void __launch_bounds__(SH_THREADS, 2) shade_kernel(ShadeParams P) {
  extern __shared__ __align__(16) float sm[];
  const vqn_shade_args& a = P.a;
  const int NP = a.n_probes;
  float* lx = sm;               // [512]
  float* ly = lx + SH_L;
  float* lz = ly + SH_L;
  float* rad = lz + SH_L;       // [NP][3][512]  radiance * area
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < SH_L; i += SH_THREADS) {
    lx[i] = a.lxyz[3 * i]; ly[i] = a.lxyz[3 * i + 1]; lz[i] = a.lxyz[3 * i + 2];
  }
  for (int i = tid; i < NP * 3 * SH_L; i += SH_THREADS) {
    int p = i / (3 * SH_L), rem = i % (3 * SH_L), ch = rem / SH_L, l = rem % SH_L;
    float v = a.lights[((size_t)p * SH_L + l) * 3 + ch];
    if (p == 0 && a.clip_light0) v = fmaxf(v, 0.f);   // light property: clip(_light, 0, inf), vq_nfr.py:759
    rad[i] = v * a.lareas[l];
  }
  __syncthreads();

  long long n = a.n_dev ? (long long)*a.n_dev : a.n;
  if (n > a.n) n = a.n;
  const long long warps_total = (long long)gridDim.x * (SH_THREADS / 32);
  long long i = (long long)blockIdx.x * (SH_THREADS / 32) + warp;

  float4 lv_next[4];
  long long row_next = 0;
  if (i < n) {
    row_next = a.row_idx ? (long long)a.row_idx[i] : i;
    if (HAS_LVIS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) lv_next[j] = ldg_stream_f4(a.lvis + row_next * SH_L + 128 * j + 4 * lane);
    }
  }
  const float INV_PI = 0.318309886183790671538f;

  for (; i < n; i += warps_total) {
    const long long row = row_next;
    float lvv[16];
    if (HAS_LVIS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lvv[4 * j] = lv_next[j].x; lvv[4 * j + 1] = lv_next[j].y;
        lvv[4 * j + 2] = lv_next[j].z; lvv[4 * j + 3] = lv_next[j].w;
      }
    }
    // prefetch the next point's visibility row (2 KB, streamed)
    long long inext = i + warps_total;
    if (inext < n) {
      row_next = a.row_idx ? (long long)a.row_idx[inext] : inext;
      if (HAS_LVIS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) lv_next[j] = ldg_stream_f4(a.lvis + row_next * SH_L + 128 * j + 4 * lane);
      }
    }
    // ---- per-point terms (warp-uniform, computed redundantly by every lane) ----
    const float px = a.xyz[row * 3], py = a.xyz[row * 3 + 1], pz = a.xyz[row * 3 + 2];
    float vx = a.rayo[row * 3] - px, vy = a.rayo[row * 3 + 1] - py, vz = a.rayo[row * 3 + 2] - pz;
    {  // _calc_vdir: safe_l2_normalize (eps on the squared norm)
      float inv = rsqrtf(fmaxf(vx * vx + vy * vy + vz * vz, 1e-6f));
      vx *= inv; vy *= inv; vz *= inv;
    }
    float nx = a.normal[row * 3], ny = a.normal[row * 3 + 1], nz = a.normal[row * 3 + 2];
    {  // _normal_correct: where(n.v >= 0, n, -n)
      float c = nx * vx + ny * vy + nz * vz;
      if (!(c >= 0.f)) { nx = -nx; ny = -ny; nz = -nz; }
    }
    if (a.normal_out && lane == 0) {
      a.normal_out[row * 3] = nx; a.normal_out[row * 3 + 1] = ny; a.normal_out[row * 3 + 2] = nz;
    }
    // get_brdf normalises the normal again (microfacet.py:20); _render uses it as passed (vq_nfr.py:701)
    const float inv_n = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-6f));
    const float vn = (nx * vx + ny * vy + nz * vz) * inv_n;
    const float alb0 = a.albedo[i * 3] * INV_PI, alb1 = a.albedo[i * 3 + 1] * INV_PI,
                alb2 = a.albedo[i * 3 + 2] * INV_PI;
    const float f00 = a.spec[i * 3], f01 = a.spec[i * 3 + 1], f02 = a.spec[i * 3 + 2];
    const float rough = a.rough[i];
    const float alpha = rough * rough;          // microfacet.py:25
    const float a2 = alpha * alpha;             // every helper uses alpha ** 2
    const float g_v = gsub_f(fminf(fmaxf(vn, 0.f), 1.f), a2);
    const float avn = fabsf(vn);
    // S = D g_l g_v / (4 |l.n| |v.n|) = a_pt * cl / (q^2 den_l |l.n|),  a_pt = a2 g_v / (2 pi |v.n|)
    const float a_pt = avn == 0.f ? 0.f : a2 * g_v * (0.5f * INV_PI) / avn;
    const float oma2 = 1.0f - a2, a2m1 = a2 - 1.0f;

    // ---- phase 1: per-light, per-channel effective weights e = F S w + albedo/pi w (probe independent) ----
    float e[16][3];
    float wd[SPLIT ? 16 : 1];   // SPLIT: e holds the glossy lobe only and wd the diffuse weight
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lb = 128 * j + 4 * lane;
      const float4 X = *reinterpret_cast<const float4*>(lx + lb);
      const float4 Y = *reinterpret_cast<const float4*>(ly + lb);
      const float4 Z = *reinterpret_cast<const float4*>(lz + lb);
      const float xs4[4] = {X.x, X.y, X.z, X.w}, ys4[4] = {Y.x, Y.y, Y.z, Y.w}, zs4[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int li = 4 * j + q;
        float dx = xs4[q] - px, dy = ys4[q] - py, dz = zs4[q] - pz;
        float inv = fast_rsqrt(fmaxf(dx * dx + dy * dy + dz * dz, 1e-6f));   // _calc_ldir
        dx *= inv; dy *= inv; dz *= inv;
        const float cos_r = dx * nx + dy * ny + dz * nz;                  // _render: cos = l . n
        const float ln = cos_r * inv_n;                                   // get_brdf: l . normalize(n)
        const float lv = dx * vx + dy * vy + dz * vz;
        // h = normalize(l + v): |l + v|^2 = 2 + 2 l.v for unit l, v
        // h = normalize(l + v), formed componentwise as the reference does (microfacet.py:21-22).  The shortcut
        // |l + v|^2 = 2 + 2 l.v is cheaper but its rounding error is amplified by 2/q in q = 1 - (h.n)^2 (1 - a^2)
        // near the highlight (h ~ n, small roughness): 2e-4 relative on the specular lobe, outside the parity budget.
        const float hx = dx + vx, hy = dy + vy, hz = dz + vz;
        const float hi_ = fast_rsqrt(fmaxf(hx * hx + hy * hy + hz * hz, 1e-6f));
        const float hvr = (hx * vx + hy * vy + hz * vz) * hi_;
        const float hnr = (hx * nx + hy * ny + hz * nz) * inv_n * hi_;
        const float hv = fminf(fmaxf(hvr, 0.f), 1.f);                     // h . v
        const float hn = fminf(fmaxf(hnr, 0.f), 1.f);                     // h . n
        const float om = 1.0f - hv;
        const float om2 = om * om;
        const float p5 = om2 * om2 * om;                                  // (1 - h.v)^5
        const float q_ = fmaf(hn * hn, a2m1, 1.0f);                       // _get_d denominator core
        const float cl = fminf(fmaxf(ln, 0.f), 1.f);
        const float den_l = cl + fast_sqrt(fabsf(fmaf(oma2, cl * cl, a2)));
        const float den = q_ * q_ * den_l * fabsf(ln);
        const float S = den == 0.f ? 0.f : __fdividef(a_pt * cl, den);
        float wv = cos_r > 0.f ? cos_r : 0.f;                              // front_lit * cos
        if (HAS_LVIS) wv *= lvv[li];
        const float sw = S * wv;
        const float psw = p5 * sw;
        // F_ch S w = f0 sw + (1 - f0) psw = psw + f0 (sw - psw)
        const float dsw = sw - psw;
        if (SPLIT) {
          e[li][0] = fmaf(f00, dsw, psw); e[li][1] = fmaf(f01, dsw, psw); e[li][2] = fmaf(f02, dsw, psw);
          wd[SPLIT ? li : 0] = wv;
        } else {
          e[li][0] = fmaf(alb0, wv, fmaf(f00, dsw, psw));
          e[li][1] = fmaf(alb1, wv, fmaf(f01, dsw, psw));
          e[li][2] = fmaf(alb2, wv, fmaf(f02, dsw, psw));
        }
      }
    }

    // ---- phase 2: probes, SH_PC at a time: 3 FFMA per (light, probe) ----
    for (int p0 = 0; p0 < NP; p0 += SH_PC) {
      float out[SH_PC * 3];
      float outd[SPLIT ? 3 : 1];
#pragma unroll
      for (int k = 0; k < SH_PC * 3; ++k) out[k] = 0.f;
      if (SPLIT) { outd[0] = 0.f; outd[SPLIT ? 1 : 0] = 0.f; outd[SPLIT ? 2 : 0] = 0.f; }
#pragma unroll
      for (int p = 0; p < SH_PC; ++p) {
        if (p0 + p < NP) {                       // warp-uniform
          const float* rp = rad + (size_t)(p0 + p) * 3 * SH_L;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int lb = 128 * j + 4 * lane;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              const float4 Lq = *reinterpret_cast<const float4*>(rp + ch * SH_L + lb);
              const float l4[4] = {Lq.x, Lq.y, Lq.z, Lq.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                out[p * 3 + ch] = fmaf(e[4 * j + q][ch], l4[q], out[p * 3 + ch]);
                if (SPLIT && p0 == 0 && p == 0)
                  outd[SPLIT ? ch : 0] = fmaf(wd[SPLIT ? 4 * j + q : 0], l4[q], outd[SPLIT ? ch : 0]);
              }
            }
          }
        }
      }
      if (SPLIT) {
        // diffuse lobe of probe 0 (mode != 'train', vq_nfr.py:605-610); every probe's total = glossy + diffuse,
        // so the diffuse weights are applied to all probes of this chunk through a second pass over wd
        float dsum[3] = {warp_sum(outd[0]) * alb0, warp_sum(outd[1]) * alb1, warp_sum(outd[2]) * alb2};
        float ssum[3] = {0.f, 0.f, 0.f};
        if (p0 == 0) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) ssum[ch] = warp_sum(out[ch]);
          if (lane == 0) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              float s_ = ssum[ch], d_ = dsum[ch];
              if (a.use_gamma) { s_ = powf(s_ * a.gamma_bias, a.gamma_index); d_ = powf(d_ * a.gamma_bias, a.gamma_index); }
              if (a.rgb_spec) a.rgb_spec[row * 3 + ch] = fminf(fmaxf(s_, 0.f), 1.f);
              if (a.rgb_diff) a.rgb_diff[row * 3 + ch] = fminf(fmaxf(d_, 0.f), 1.f);
            }
          }
        }
        // totals: add the diffuse part of every probe in this chunk
#pragma unroll
        for (int p = 0; p < SH_PC; ++p) {
          if (p0 + p < NP) {
            const float* rp = rad + (size_t)(p0 + p) * 3 * SH_L;
            float dd[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) {
                const float4 Lq = *reinterpret_cast<const float4*>(rp + ch * SH_L + 128 * j + 4 * lane);
                dd[ch] = fmaf(wd[SPLIT ? 4 * j : 0], Lq.x, dd[ch]);
                dd[ch] = fmaf(wd[SPLIT ? 4 * j + 1 : 0], Lq.y, dd[ch]);
                dd[ch] = fmaf(wd[SPLIT ? 4 * j + 2 : 0], Lq.z, dd[ch]);
                dd[ch] = fmaf(wd[SPLIT ? 4 * j + 3 : 0], Lq.w, dd[ch]);
              }
            out[p * 3 + 0] = fmaf(alb0, dd[0], out[p * 3 + 0]);
            out[p * 3 + 1] = fmaf(alb1, dd[1], out[p * 3 + 1]);
            out[p * 3 + 2] = fmaf(alb2, dd[2], out[p * 3 + 2]);
          }
        }
      }
      // cross-lane reduction: 12 -> 6 -> 3 values by halving on lane bits 4, 3, then 3 butterfly sums over 8 lanes
      sh_halve<12>(out, lane, 4);
      sh_halve<6>(out, lane, 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        out[k] += __shfl_xor_sync(0xffffffffu, out[k], 4);
        out[k] += __shfl_xor_sync(0xffffffffu, out[k], 2);
        out[k] += __shfl_xor_sync(0xffffffffu, out[k], 1);
      }
      if (a.rgb && (lane & 7) == 0) {
        const int base = 6 * ((lane >> 4) & 1) + 3 * ((lane >> 3) & 1);   // first of this lane's 3 values
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int vi = base + k, p = vi / 3, ch = vi % 3;
          if (p0 + p < NP) {
            float v = out[k];
            if (a.use_gamma) v = powf(v * a.gamma_bias, a.gamma_index);    // vq_nfr.py:715-716
            if (!isfinite(v)) atomicOr(P.nonfinite, 2);                    // check_numerics (:731)
            if (!a.no_clip) v = fminf(fmaxf(v, 0.f), 1.f);                 // clip_by_value (:718)
            if (a.to_srgb) v = vqn_linear2srgb(v);
            a.rgb[(row * NP + p0 + p) * 3 + ch] = v;
          }
        }
      }
    }
  }
}

int vqn_shade_pt_launch(vqn_ctx* ctx, const vqn_shade_args& a, cudaStream_t s);   // shade_pt.cu

extern "C" int vqn_shade(vqn_ctx* ctx, const vqn_shade_args* args, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && args, "shade: null args");
  const vqn_shade_args& a = *args;
  VQN_CHECK_ARG(a.xyz && a.rayo && a.normal && a.albedo && a.spec && a.rough, "shade: null per-point input");
  VQN_CHECK_ARG(a.lxyz && a.lareas && a.lights, "shade: null light tables");
  VQN_CHECK_ARG(a.n_probes >= 1 && a.n_probes <= 33, "shade: 1 <= n_probes <= 33");
  VQN_CHECK_ARG(a.n >= 0, "shade: n < 0");
  if (a.n == 0) return VQN_OK;
  ShadeParams P;
  P.a = a;
  P.nonfinite = ctx->nonfinite_flag;
  size_t smem = sizeof(float) * (3 * SH_L + (size_t)a.n_probes * 3 * SH_L);
  VQN_CHECK_ARG((int)smem <= ctx->max_smem_optin, "shade: probe tables exceed shared memory");
  const bool split = a.rgb_diff || a.rgb_spec;
  // large un-split batches: thread-per-point kernel with the light tables in the constant bank (shade_pt.cu);
  // small batches (training, 8192 rays) keep the warp-per-point kernel, which exposes 32x more parallelism
  VQN_CHECK_ARG(a.n_peers >= 0 && a.n_peers <= 8, "shade: 0 <= n_peers <= 8");
  VQN_CHECK_ARG(a.lvis_format >= VQN_LVIS_F32 && a.lvis_format <= VQN_LVIS_U8, "shade: unknown lvis_format");
  const bool compact_lvis = a.lvis && a.lvis_format != VQN_LVIS_F32;
  VQN_CHECK_ARG(!a.no_clip || (!a.use_gamma && !a.to_srgb && a.n_peers == 0 && !compact_lvis),
                "shade: no_clip returns the raw integral (no gamma / sRGB / gather / compact visibility)");
  if (!split && !a.no_clip && a.n_probes <= 9 && (a.n >= 32768 || a.n_peers > 0 || compact_lvis)) return vqn_shade_pt_launch(ctx, a, vqn_cs(stream));
  VQN_CHECK_ARG(a.n_peers == 0, "shade: the fused peer gather needs an un-split batch with at most 9 probes");
  if (compact_lvis) { vqn_set_error("shade: float16 / uint8 light visibility needs an un-split batch with at most 9 probes"); return VQN_ERR_UNSUPPORTED; }
  auto kern = a.lvis ? (split ? shade_kernel<true, true> : shade_kernel<true, false>)
                     : (split ? shade_kernel<false, true> : shade_kernel<false, false>);
  VQN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  VQN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SH_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  long long want = (a.n + (SH_THREADS / 32) - 1) / (SH_THREADS / 32);
  long long cap = (long long)ctx->sm_count * per_sm;   // persistent: whole multiples of the SM count
  int blocks = (int)(want < cap ? want : cap);
  kern<<<blocks, SH_THREADS, smem, vqn_cs(stream)>>>(P);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// Fine-grained materialising variants (API parity with _eval_brdf_at / _render; debug sizes only)
// ---------------------------------------------------------------------------------------------
__global__ void eval_brdf_kernel(const float* __restrict__ pts2l, const float* __restrict__ pts2c,
                                 const float* __restrict__ normal, const float* __restrict__ albedo,
                                 const float* __restrict__ spec, const float* __restrict__ rough, long long n,
                                 float* __restrict__ brdf, float* __restrict__ glossy, float* __restrict__ diffuse) {
  const float PI = 3.14159265358979323846f;
  long long total = n * SH_L;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long i = idx / SH_L;
    float lx = pts2l[idx * 3], ly = pts2l[idx * 3 + 1], lz = pts2l[idx * 3 + 2];
    float inv = rsqrtf(fmaxf(lx * lx + ly * ly + lz * lz, 1e-6f));
    lx *= inv; ly *= inv; lz *= inv;
    float vx = pts2c[i * 3], vy = pts2c[i * 3 + 1], vz = pts2c[i * 3 + 2];
    inv = rsqrtf(fmaxf(vx * vx + vy * vy + vz * vz, 1e-6f));
    vx *= inv; vy *= inv; vz *= inv;
    float nx = normal[i * 3], ny = normal[i * 3 + 1], nz = normal[i * 3 + 2];
    inv = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-6f));
    nx *= inv; ny *= inv; nz *= inv;
    float hx = lx + vx, hy = ly + vy, hz = lz + vz;
    inv = rsqrtf(fmaxf(hx * hx + hy * hy + hz * hz, 1e-6f));
    hx *= inv; hy *= inv; hz *= inv;
    float hv = fminf(fmaxf(hx * vx + hy * vy + hz * vz, 0.f), 1.f);
    float om = 1.f - hv;
    float p5 = om * om * om * om * om;
    float r = rough[i];
    float alpha = r * r, a2 = alpha * alpha;
    float hn = fminf(fmaxf(hx * nx + hy * ny + hz * nz, 0.f), 1.f);
    float dden = PI * (hn * hn * (a2 - 1.f) + 1.f) * (hn * hn * (a2 - 1.f) + 1.f);
    float d = dden == 0.f ? 0.f : a2 / dden;
    float ln = lx * nx + ly * ny + lz * nz, vn = vx * nx + vy * ny + vz * nz;
    float g = gsub_f(fminf(fmaxf(ln, 0.f), 1.f), a2) * gsub_f(fminf(fmaxf(vn, 0.f), 1.f), a2);
    float denom = 4.f * fabsf(ln) * fabsf(vn);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float f0 = spec[i * 3 + ch];
      float f = f0 + (1.f - f0) * p5;
      float gl = denom == 0.f ? 0.f : f * g * d / denom;
      float df = albedo[i * 3 + ch] / PI;
      if (glossy) glossy[idx * 3 + ch] = gl;
      if (diffuse) diffuse[idx * 3 + ch] = df;
      if (brdf) brdf[idx * 3 + ch] = gl + df;
    }
  }
}

extern "C" int vqn_eval_brdf(vqn_ctx* ctx, const float* pts2l, const float* pts2c, const float* normal,
                             const float* albedo, const float* spec, const float* rough, int64_t n, float* brdf,
                             float* glossy, float* diffuse, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && pts2l && pts2c && normal && albedo && spec && rough && n >= 0, "eval_brdf args");
  if (n == 0) return VQN_OK;
  long long want = (n * SH_L + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  eval_brdf_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(pts2l, pts2c, normal, albedo, spec, rough, n, brdf,
                                                       glossy, diffuse);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// _render for one probe from a materialised brdf: one warp per point
__global__ void render_kernel(const float* __restrict__ brdf, const float* __restrict__ l,
                              const float* __restrict__ normal, const float* __restrict__ lvis,
                              const float* __restrict__ lareas, const float* __restrict__ light, long long n,
                              int use_gamma, float gb, float gi, float* __restrict__ rgb, int* nonfinite) {
  int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < n; i += nw) {
    float nx = normal[i * 3], ny = normal[i * 3 + 1], nz = normal[i * 3 + 2];
    float acc[3] = {0.f, 0.f, 0.f};
    for (int li = lane; li < SH_L; li += 32) {
      long long idx = i * SH_L + li;
      float c = l[idx * 3] * nx + l[idx * 3 + 1] * ny + l[idx * 3 + 2] * nz;
      float vis = c > 0.f ? 1.f : 0.f;
      if (lvis) vis *= lvis[idx];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        acc[ch] += brdf[idx * 3 + ch] * (vis * light[li * 3 + ch]) * c * lareas[li];
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float v = warp_sum(acc[ch]);
      if (lane == 0) {
        if (use_gamma) v = powf(v * gb, gi);
        if (!isfinite(v)) atomicOr(nonfinite, 2);
        rgb[i * 3 + ch] = fminf(fmaxf(v, 0.f), 1.f);
      }
    }
  }
}

extern "C" int vqn_render(vqn_ctx* ctx, const float* brdf, const float* l, const float* normal,
                          const float* lvis, const float* lareas, const float* light, int64_t n, int use_gamma,
                          float gamma_bias, float gamma_index, float* rgb, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && brdf && l && normal && lareas && light && rgb && n >= 0, "render args");
  if (n == 0) return VQN_OK;
  long long want = (n + 7) / 8;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  render_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(brdf, l, normal, lvis, lareas, light, n, use_gamma,
                                                    gamma_bias, gamma_index, rgb, ctx->nonfinite_flag);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// Background rows of a fused image gather.  shade_pt_kernel stores only the live (alpha > 0) rows of a shard into the
// peers' image buffers; this writes zeros for the shard's OTHER rows (models/vq_nfr.py:347-370: scatter_nd leaves zeros
// at background rows), so that a gathered frame never shows pixels of the previous one.  Traffic: dead rows only.
__global__ void peer_clear_background_kernel(const float* __restrict__ alpha, int alpha_stride, long long n_local,
                                             long long peer_row0, int width, vqn_peer_ptrs pp) {
  const long long total = n_local * width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / width;
    if (alpha[row * alpha_stride] > 0.f) continue;
    const long long off = (peer_row0 + row) * width + (i - row * width);
    for (int q = 0; q < pp.n; ++q) pp.p[q][off] = 0.f;
  }
}

extern "C" int vqn_peer_clear_background(vqn_ctx* ctx, const float* alpha, int alpha_stride, int64_t n_local,
                                         int64_t peer_row0, int width, float* const* peer_ptrs, int n_peers,
                                         vqn_stream stream) {
  VQN_CHECK_ARG(ctx && alpha && alpha_stride >= 1 && n_local >= 0 && width >= 1, "peer_clear_background args");
  VQN_CHECK_ARG(peer_ptrs && n_peers >= 1 && n_peers <= 8, "peer_clear_background: 1 <= n_peers <= 8");
  if (n_local == 0) return VQN_OK;
  vqn_peer_ptrs pp;
  pp.n = n_peers;
  for (int q = 0; q < 8; ++q) pp.p[q] = q < n_peers ? peer_ptrs[q] : nullptr;
  const long long want = (n_local * width + 255) / 256;
  const int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  peer_clear_background_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(alpha, alpha_stride, n_local, peer_row0, width, pp);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// Frame hand-shake of the fused image gather on ONE destination rank (dist.PeerImage(dst=r)): a one-thread kernel per rank
// and frame instead of a symmetric-memory barrier over all ranks (~100 us at 8 GPUs against a 0.8 ms step).
//   sender:       frame f done (its P2P stores precede this kernel in stream order) -> flags_dst[rank] = f;
//                 then wait for ack >= f - 1: the destination has consumed frame f - 1, whose buffer frame f + 1 reuses
//   destination:  acknowledge frame f - 1 to every sender (its reads of that frame precede this kernel in stream order),
//                 then wait until flags[r] >= f for every sender r
// flags: world + 1 int32 per rank in symmetric memory ([r] = last frame finished by rank r, [world] = last frame
// acknowledged by the destination); counter: this rank's frame count (device memory, so a captured graph replays).
// ---------------------------------------------------------------------------------------------
struct PeerFlagPtrs { int* p[8]; };
__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) { int v; asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__global__ void peer_frame_sync_kernel(int* counter, const int* local_flags, PeerFlagPtrs peers, int world, int rank, int dst) {
  if (threadIdx.x != 0) return;
  __threadfence_system();
  const int f = ++(*counter);
  if (rank == dst) {
    for (int r = 0; r < world; ++r) if (r != rank) st_release_sys(peers.p[r] + world, f - 1);
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      const long long t0 = clock64();
      for (unsigned spins = 0; ld_acquire_sys(local_flags + r) < f; ++spins)   // a protocol bug traps (after ~60 s) instead of hanging
        if ((spins & 4095u) == 4095u && clock64() - t0 > 120000000000LL) { asm volatile("trap;"); }
    }
  } else {
    st_release_sys(peers.p[dst] + rank, f);
    const long long t0 = clock64();
    for (unsigned spins = 0; ld_acquire_sys(local_flags + world) < f - 1; ++spins)
      if ((spins & 4095u) == 4095u && clock64() - t0 > 120000000000LL) { asm volatile("trap;"); }
  }
  __threadfence_system();
}

extern "C" int vqn_peer_frame_sync(vqn_ctx* ctx, int32_t* counter, const int32_t* local_flags, int32_t* const* peer_flags,
                                   int world, int rank, int dst, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && counter && local_flags && peer_flags, "peer_frame_sync: null pointer");
  VQN_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world && dst >= 0 && dst < world,
                "peer_frame_sync: bad world / rank / dst");
  PeerFlagPtrs pp;
  for (int r = 0; r < 8; ++r) pp.p[r] = r < world ? peer_flags[r] : nullptr;
  peer_frame_sync_kernel<<<1, 32, 0, vqn_cs(stream)>>>(counter, local_flags, pp, world, rank, dst);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
