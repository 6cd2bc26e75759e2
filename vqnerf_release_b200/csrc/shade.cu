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
// Phase 1 computes, per light, the three probe-independent weights
//     w   = front * lvis * cos            (area is pre-multiplied into the radiance tables)
//     sw  = S * w,  S = D G / (4 |l.n| |v.n|)   (Fresnel-free glossy scalar)
//     psw = (1 - h.v)^5 * sw
// and keeps them in registers (48 per lane).  Phase 2 is 9 FFMA per (light, probe):
//     A += sw L, B += psw L, C += w L   (per channel)
// and  rgb = f0 A + (1 - f0) B + albedo/pi C,  because F = f0 + (1 - f0)(1 - h.v)^5 is affine in f0.
// Partial rgb sums are combined BEFORE the cross-lane reduction (linear), so only 3 values per probe are
// shuffled.  Algorithmic HBM bytes per point: 2048 (lvis) + 36 + 28 + 12 (1 + P).
#include "common.cuh"

#define SH_L 512
#define SH_THREADS 256
#define SH_PC 4  // probes per accumulation pass

struct ShadeParams {
  vqn_shade_args a;
  int* nonfinite;
};

__device__ __forceinline__ float gsub_f(float c, float a2) {
  // util/microfacet.py:49-69: 2c / (c + sqrt(|a2 + (1-a2) c^2|)), divide_no_nan
  float den = c + sqrtf(fabsf(a2 + (1.0f - a2) * c * c));
  return den == 0.f ? 0.f : 2.0f * c / den;
}

template <bool HAS_LVIS>
__global__ void __launch_bounds__(SH_THREADS) shade_kernel(ShadeParams P) {
  extern __shared__ __align__(16) float sm[];
  const vqn_shade_args& a = P.a;
  const int NP = a.n_probes;
  float* lx = sm;               // [512]
  float* ly = lx + SH_L;
  float* lz = ly + SH_L;
  float* rad = lz + SH_L;       // [NP][3][512]  radiance * area, clip(.,0,inf) applied to probe 0
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
    // prefetch next point's visibility row
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

    // ---- phase 1: per-light weights ----
    float w[16], sw[16], psw[16];
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
        float inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-6f));   // _calc_ldir
        dx *= inv; dy *= inv; dz *= inv;
        const float cos_r = dx * nx + dy * ny + dz * nz;                  // _render: cos = l . n
        const float ln = cos_r * inv_n;                                   // get_brdf: l . normalize(n)
        const float lv = dx * vx + dy * vy + dz * vz;
        // h = normalize(l + v): |l + v|^2 = 2 + 2 l.v for unit l, v
        const float t = 1.0f + lv;
        const float hinv = rsqrtf(fmaxf(2.0f * t, 1e-6f));
        const float hv = fminf(fmaxf(t * hinv, 0.f), 1.f);                // h . v
        const float hn = fminf(fmaxf((ln + vn) * hinv, 0.f), 1.f);        // h . n
        const float om = 1.0f - hv;
        const float om2 = om * om;
        const float p5 = om2 * om2 * om;                                  // (1 - h.v)^5
        const float q_ = hn * hn * (a2 - 1.0f) + 1.0f;                    // _get_d denominator core
        const float cl = fminf(fmaxf(ln, 0.f), 1.f);
        const float den_l = cl + sqrtf(fabsf(a2 + (1.0f - a2) * cl * cl));
        const float den = q_ * q_ * den_l * fabsf(ln);
        const float S = den == 0.f ? 0.f : a_pt * cl / den;
        const bool front = cos_r > 0.f;
        float wv = front ? cos_r : 0.f;
        if (HAS_LVIS) wv *= lvv[li];
        w[li] = wv;
        sw[li] = S * wv;
        psw[li] = p5 * sw[li];
      }
    }

    // ---- phase 2: probes, SH_PC at a time ----
    const bool want_split = (a.rgb_diff != nullptr) || (a.rgb_spec != nullptr);
    for (int p0 = 0; p0 < NP; p0 += SH_PC) {
      float A[SH_PC][3], B[SH_PC][3], C[SH_PC][3];
#pragma unroll
      for (int p = 0; p < SH_PC; ++p)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) { A[p][ch] = 0.f; B[p][ch] = 0.f; C[p][ch] = 0.f; }
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
                A[p][ch] = fmaf(sw[4 * j + q], l4[q], A[p][ch]);
                B[p][ch] = fmaf(psw[4 * j + q], l4[q], B[p][ch]);
                C[p][ch] = fmaf(w[4 * j + q], l4[q], C[p][ch]);
              }
            }
          }
        }
      }
      // combine (linear) then reduce 3 values per probe across the warp
      float out[SH_PC][3];
#pragma unroll
      for (int p = 0; p < SH_PC; ++p) {
        float s0 = f00 * A[p][0] + (1.f - f00) * B[p][0];
        float s1 = f01 * A[p][1] + (1.f - f01) * B[p][1];
        float s2 = f02 * A[p][2] + (1.f - f02) * B[p][2];
        float d0 = alb0 * C[p][0], d1 = alb1 * C[p][1], d2 = alb2 * C[p][2];
        if (want_split && p0 == 0 && p == 0) {   // probe 0 lobes (mode != 'train', vq_nfr.py:605-610)
          s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
          d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
          if (lane == 0) {
            float sv[3] = {s0, s1, s2}, dv[3] = {d0, d1, d2};
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              float s_ = sv[ch], d_ = dv[ch];
              if (a.use_gamma) { s_ = powf(s_ * a.gamma_bias, a.gamma_index); d_ = powf(d_ * a.gamma_bias, a.gamma_index); }
              s_ = fminf(fmaxf(s_, 0.f), 1.f); d_ = fminf(fmaxf(d_, 0.f), 1.f);
              if (a.rgb_spec) a.rgb_spec[row * 3 + ch] = s_;
              if (a.rgb_diff) a.rgb_diff[row * 3 + ch] = d_;
            }
          }
          out[p][0] = s0 + d0; out[p][1] = s1 + d1; out[p][2] = s2 + d2;   // already full sums
          // mark as reduced by zeroing on other lanes so the generic reduction below stays correct
          if (lane != 0) { out[p][0] = 0.f; out[p][1] = 0.f; out[p][2] = 0.f; }
        } else {
          out[p][0] = s0 + d0; out[p][1] = s1 + d1; out[p][2] = s2 + d2;
        }
      }
#pragma unroll
      for (int p = 0; p < SH_PC; ++p)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) out[p][ch] = warp_sum(out[p][ch]);
      if (a.rgb) {
        // lanes 0..3*SH_PC-1 each finish one (probe, channel)
#pragma unroll
        for (int p = 0; p < SH_PC; ++p)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            if (lane == p * 3 + ch && p0 + p < NP) {
              float v = out[p][ch];
              if (a.use_gamma) v = powf(v * a.gamma_bias, a.gamma_index);    // vq_nfr.py:715-716
              if (!isfinite(v)) atomicOr(P.nonfinite, 2);                    // check_numerics (:731)
              v = fminf(fmaxf(v, 0.f), 1.f);                                 // clip_by_value (:718)
              if (a.to_srgb) v = vqn_linear2srgb(v);
              a.rgb[(row * NP + p0 + p) * 3 + ch] = v;
            }
          }
      }
    }
  }
}

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
  auto kern = a.lvis ? shade_kernel<true> : shade_kernel<false>;
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
