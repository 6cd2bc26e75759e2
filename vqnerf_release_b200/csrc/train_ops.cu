// Backward / loss / optimizer kernels of the decomposition-stage training step (BASELINE config #4):
// Model.call(mode='train') + compute_loss under tf.GradientTape and the Adam(amsgrad) update
// (models/vq_nfr.py:534-692, 876-986; train_nfr.py:121-139, 562-576).
//
//  shade_bwd_kernel      d(rgb)/d(albedo, f0, rough, _light) of the fused light integral (analytic, nothing of shape
//                        [N,512,.] materialised; the forward terms are recomputed per light)
//  loss_train_kernel     per-example loss terms + their gradients w.r.t. rgb, vq_rgb, z_vq, spec
//  vq_bwd_kernel         straight-through estimator + commitment term + l2_normalize backward
//  combine_bwd_kernel    spec = ks*base, albedo = (1-ks)*base backward
//  sim_loss_kernel       codebook separation loss and its gradient through get_codebook()
//  adam_kernel           tf.keras.optimizers.Adam(amsgrad=True) dense update
#include "common.cuh"
#include <cstdlib>

#define TL 512

namespace {

__device__ __forceinline__ float fast_rsqrt_(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct ShadeBwdParams {
  const float *xyz, *rayo, *normal, *lvis;
  const int32_t* row_idx;
  long long n;
  const float *albedo, *spec, *rough;
  const float *lxyz, *lareas, *light;
  int clip_light0;
  const float* d_rgb;   // [n,3] compact, gradient w.r.t. the linear clipped rgb (clip has identity gradient, :718)
  float *d_albedo, *d_spec, *d_rough;   // compact
  float* d_light;       // [512,3], accumulated
};

// One warp per point, lane l owns lights {128 j + 4 l + q} exactly as shade_kernel.  With
//   rgb_ch = sum_l (F_ch S + alb_ch/pi) w_l R_l,ch,   F = p5 + f0 (1 - p5),   S = A g_v cl / (2 pi q^2 den_l |l.n| |v.n|),
//   A = rough^4, q = hn^2 (A - 1) + 1, den_l = cl + sqrt(A + (1 - A) cl^2), g_v = 2 cv / (cv + sqrt(A + (1 - A) cv^2)):
//   d alb_ch  = g_ch / pi * sum_l w R_ch
//   d f0_ch   = g_ch * sum_l (1 - p5) S w R_ch
//   d rough   = 4 rough^3 * sum_l S w (sum_ch g_ch F_ch R_ch) * [1/A + T_v - 2 hn^2/q - (1 - cl^2)/(2 sqrt(u_l) den_l)]
//   d light_l,ch += g_ch (F_ch S + alb_ch/pi) w area_l         (clip_by_value_preserve_gradient: identity, :759)
// Work split: SB_LW warps share a point, each owning 4 / SB_LW blocks of 128 lights (lane l: 4 consecutive lights of a
// block, the float4 the forward kernel reads).  With all 512 lights in one warp the 48 per-lane light gradients pushed the
// kernel to 161 registers = 8..12 warps per SM, and at 8192 points (55 per SM) it was latency-bound at 4x its issue bound.
// The partial sums of a point meet in shared memory in a fixed order (deterministic); the loop trip count is uniform per
// CTA.  Per-light divisions are the forward kernel's __fdividef (shade.cu); sqrt is sqrt.approx (1 ulp).
constexpr int SB_THREADS = 256, SB_WARPS = SB_THREADS / 32;
__device__ __forceinline__ float fast_sqrt_(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <int SB_LW, int MIN_CTAS>
__global__ void __launch_bounds__(SB_THREADS, MIN_CTAS) shade_bwd_kernel(ShadeBwdParams a) {
  constexpr int SB_PPI = SB_WARPS / SB_LW, SB_LB = 4 / SB_LW;
  __shared__ __align__(16) float lx[TL], ly[TL], lz[TL], rad[3 * TL], dls[3 * TL];
  __shared__ float part[2][SB_PPI][SB_LW][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ps = warp / SB_LW, lq = warp % SB_LW;
  // light tables permuted so that lane l of a warp finds light 4 l + q of a 128-block at [32 q + l]: the per-light scalar
  // reads in the loop are bank-conflict-free and nothing loop-invariant has to be held in registers
  for (int i = tid; i < TL; i += SB_THREADS) {
    const int pos = (i & ~127) + 32 * (i & 3) + ((i & 127) >> 2);
    lx[pos] = a.lxyz[3 * i]; ly[pos] = a.lxyz[3 * i + 1]; lz[pos] = a.lxyz[3 * i + 2];
  }
  for (int i = tid; i < 3 * TL; i += SB_THREADS) {
    const int ch = i / TL, l = i % TL;
    const int pos = (l & ~127) + 32 * (l & 3) + ((l & 127) >> 2);
    float v = a.light[l * 3 + ch];
    if (a.clip_light0) v = fmaxf(v, 0.f);
    rad[ch * TL + pos] = v * a.lareas[l];
    dls[i] = 0.f;
  }
  __syncthreads();
  const float INV_PI = 0.318309886183790671538f;
  float dl[4 * SB_LB][3];
#pragma unroll
  for (int k = 0; k < 4 * SB_LB; ++k) { dl[k][0] = 0.f; dl[k][1] = 0.f; dl[k][2] = 0.f; }
  int it = 0;
  for (long long base = (long long)blockIdx.x * SB_PPI; base < a.n; base += (long long)gridDim.x * SB_PPI, ++it) {
    const long long i = base + ps;
    const bool valid = i < a.n;                  // warp-uniform
    float g[3] = {0.f, 0.f, 0.f}, rough = 0.f;
    float s_alb[3] = {0.f, 0.f, 0.f}, s_f0[3] = {0.f, 0.f, 0.f}, s_rough = 0.f;
    if (valid) {
      const long long row = a.row_idx ? (long long)a.row_idx[i] : i;
      const float px = a.xyz[row * 3], py = a.xyz[row * 3 + 1], pz = a.xyz[row * 3 + 2];
      float vx = a.rayo[row * 3] - px, vy = a.rayo[row * 3 + 1] - py, vz = a.rayo[row * 3 + 2] - pz;
      {
        float inv = rsqrtf(fmaxf(vx * vx + vy * vy + vz * vz, 1e-6f));
        vx *= inv; vy *= inv; vz *= inv;
      }
      float nx = a.normal[row * 3], ny = a.normal[row * 3 + 1], nz = a.normal[row * 3 + 2];
      {
        float c = nx * vx + ny * vy + nz * vz;
        if (!(c >= 0.f)) { nx = -nx; ny = -ny; nz = -nz; }
      }
      const float inv_n = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-6f));
      const float vn = (nx * vx + ny * vy + nz * vz) * inv_n;
      const float alb[3] = {a.albedo[i * 3] * INV_PI, a.albedo[i * 3 + 1] * INV_PI, a.albedo[i * 3 + 2] * INV_PI};
      const float f0[3] = {a.spec[i * 3], a.spec[i * 3 + 1], a.spec[i * 3 + 2]};
      g[0] = a.d_rgb[i * 3]; g[1] = a.d_rgb[i * 3 + 1]; g[2] = a.d_rgb[i * 3 + 2];
      rough = a.rough[i];
      const float alpha = rough * rough, a2 = alpha * alpha;
      const float oma2 = 1.0f - a2, a2m1 = a2 - 1.0f;
      const float cv = fminf(fmaxf(vn, 0.f), 1.f);
      const float su_v = sqrtf(fabsf(a2 + oma2 * cv * cv));
      const float den_v = cv + su_v;
      const float g_v = den_v == 0.f ? 0.f : 2.0f * cv / den_v;
      const float avn = fabsf(vn);
      const float a_pt = avn == 0.f ? 0.f : a2 * g_v * (0.5f * INV_PI) / avn;
      // point-level part of d ln(S)/dA: 1/A + (d g_v/dA)/g_v
      const float t_pt = (a2 > 0.f ? 1.0f / a2 : 0.f) -
                         ((su_v * den_v) > 0.f ? (1.0f - cv * cv) / (2.0f * su_v * den_v) : 0.f);
#pragma unroll
      for (int j = 0; j < SB_LB; ++j) {
        const int blk = lq * SB_LB + j;
        const int pb = 128 * blk + lane;         // the lane's constants in the permuted tables: [pb + 32 q]
        float4 LV = make_float4(1.f, 1.f, 1.f, 1.f);
        if (a.lvis) LV = ldg_stream_f4(a.lvis + row * TL + 128 * blk + 4 * lane);
        const float lv4[4] = {LV.x, LV.y, LV.z, LV.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int li = 4 * j + q;
          float dx = lx[pb + 32 * q] - px, dy = ly[pb + 32 * q] - py, dz = lz[pb + 32 * q] - pz;
          const float inv = fast_rsqrt_(fmaxf(dx * dx + dy * dy + dz * dz, 1e-6f));
          dx *= inv; dy *= inv; dz *= inv;
          const float cos_r = dx * nx + dy * ny + dz * nz;
          const float ln = cos_r * inv_n;
          // h = normalize(l + v), formed componentwise as the reference does (microfacet.py:21-22).  The shortcut
          // |l + v|^2 = 2 + 2 l.v is cheaper but its rounding error is amplified by 2/q in q = 1 - (h.n)^2 (1 - a^2)
          // near the highlight (h ~ n, small roughness): 2e-4 relative on the specular lobe, outside the parity budget.
          const float hx = dx + vx, hy = dy + vy, hz = dz + vz;
          const float hi_ = fast_rsqrt_(fmaxf(hx * hx + hy * hy + hz * hz, 1e-6f));
          const float hvr = (hx * vx + hy * vy + hz * vz) * hi_;
          const float hnr = (hx * nx + hy * ny + hz * nz) * inv_n * hi_;
          const float hv = fminf(fmaxf(hvr, 0.f), 1.f);                     // h . v
          const float hn = fminf(fmaxf(hnr, 0.f), 1.f);                     // h . n
          const float om = 1.0f - hv, om2 = om * om;
          const float p5 = om2 * om2 * om;
          const float q_ = fmaf(hn * hn, a2m1, 1.0f);
          const float cl = fminf(fmaxf(ln, 0.f), 1.f);
          const float su_l = fast_sqrt_(fabsf(fmaf(oma2, cl * cl, a2)));
          const float den_l = cl + su_l;
          const float den = q_ * q_ * den_l * fabsf(ln);
          const float S = den == 0.f ? 0.f : __fdividef(a_pt * cl, den);
          const float wv = (cos_r > 0.f ? cos_r : 0.f) * lv4[q];
          const float sw = S * wv;
          const float R[3] = {rad[pb + 32 * q], rad[TL + pb + 32 * q], rad[2 * TL + pb + 32 * q]};
          float gFR = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float F = fmaf(f0[ch], 1.0f - p5, p5);
            s_alb[ch] = fmaf(wv, R[ch], s_alb[ch]);
            s_f0[ch] = fmaf((1.0f - p5) * sw, R[ch], s_f0[ch]);
            gFR = fmaf(g[ch] * F, R[ch], gFR);
            dl[li][ch] = fmaf(g[ch], fmaf(F, sw, alb[ch] * wv), dl[li][ch]);
          }
          const float sd = su_l * den_l;
          const float bk = t_pt - (q_ != 0.f ? __fdividef(2.0f * hn * hn, q_) : 0.f) -
                           (sd > 0.f ? __fdividef(1.0f - cl * cl, 2.0f * sd) : 0.f);
          s_rough = fmaf(sw * bk, gFR, s_rough);
        }
      }
    }
    // (outside the branch: `valid` is warp-uniform, but the compiler cannot know, and shuffles under a divergent-looking
    // branch cost a WARPSYNC + collective prologue each)
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) { s_alb[ch] = warp_sum(s_alb[ch]); s_f0[ch] = warp_sum(s_f0[ch]); }
    s_rough = warp_sum(s_rough);
    if (lane == 0) {
      float* pp = part[it & 1][ps][lq];
      pp[0] = s_alb[0]; pp[1] = s_alb[1]; pp[2] = s_alb[2];
      pp[3] = s_f0[0]; pp[4] = s_f0[1]; pp[5] = s_f0[2]; pp[6] = s_rough;
    }
    __syncthreads();                   // buffer (it & 1) is rewritten two iterations later, after the next barrier
    if (lq == 0 && valid && lane < 7) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < SB_LW; ++w) sum += part[it & 1][ps][w][lane];
      const int ch = lane < 3 ? lane : lane - 3;
      const float gc = ch == 0 ? g[0] : ch == 1 ? g[1] : g[2];
      if (lane < 3) a.d_albedo[i * 3 + lane] = gc * INV_PI * sum;
      else if (lane < 6) a.d_spec[i * 3 + lane - 3] = gc * sum;
      else a.d_rough[i] = 4.0f * rough * rough * rough * sum;
    }
  }
  // flush the per-lane light gradients: smem per block, then one global atomic per (light, channel)
#pragma unroll
  for (int j = 0; j < SB_LB; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        atomicAdd(&dls[ch * TL + 128 * (lq * SB_LB + j) + 4 * lane + q], dl[4 * j + q][ch]);
  __syncthreads();
  for (int i = tid; i < 3 * TL; i += SB_THREADS) {
    const int ch = i / TL, l = i % TL;
    const float v = dls[i] * a.lareas[l];
    if (v != 0.f) atomicAdd(&a.d_light[l * 3 + ch], v);
  }
}

// ---------------------------------------------------------------------------------------------------------
struct LossParams {
  const float *gtc, *rgb, *vqrgb, *z, *spec, *rough;
  long long n;           // rows (even: consecutive rows are (pixel, neighbour) pairs, train_nfr.py:447-448)
  int zdim, nerf;
  float combine_w, chroma_w, smooth_w, lambert_w, chr_alpha, chr_thres, inv_gbs;
  float *loss_rows;      // [n] per-example loss without the broadcast scalars (vqloss, sim_smooth)
  float *d_rgb, *d_vqrgb, *d_z, *d_spec;   // gradients of sum(loss)/global_bs
  float* sums;           // [6] accumulated: rgb, vqrgb, chromaticity, chr_smooth, lambert, total (sums over rows)
};

__device__ __forceinline__ void chroma3(const float (&v)[3], float (&c)[3], float& nrm) {
  nrm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);     // _rgb2chromaticity (:1135-1137), divide_no_nan
  const float inv = nrm == 0.f ? 0.f : 1.0f / nrm;
  c[0] = v[0] * inv; c[1] = v[1] * inv; c[2] = v[2] * inv;
}

// one warp per pair of rows (2i, 2i+1); lanes 0/1 own the scalar terms of the two rows, all lanes share z
__global__ void __launch_bounds__(256) loss_train_kernel(LossParams p) {
  const int lane = threadIdx.x & 31;
  const long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long npairs = p.n >> 1;
  __shared__ float bsum[6];
  if (threadIdx.x < 6) bsum[threadIdx.x] = 0.f;
  __syncthreads();
  if (pair < npairs) {
  float terms[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float schr[3] = {0.f, 0.f, 0.f};
  if (lane < 2) {
    const long long r = pair * 2 + lane;
    float gt[3], lin[3], pr[3], vq[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      gt[c] = p.gtc[r * 3 + c];
      lin[c] = p.nerf ? vqn_srgb2linear(gt[c]) : gt[c];               // :894-899
      pr[c] = p.rgb[r * 3 + c];
      vq[c] = p.vqrgb[r * 3 + c];
    }
    // rgb / vqrgb MSE (:924-929): mean over channels
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float e1 = pr[c] - lin[c], e2 = vq[c] - lin[c];
      terms[0] += p.combine_w * e1 * e1 * (1.0f / 3.0f);
      terms[1] += e2 * e2 * (1.0f / 3.0f);
      p.d_rgb[r * 3 + c] = p.combine_w * 2.0f * e1 * (1.0f / 3.0f) * p.inv_gbs;
    }
    // chromaticity (:936-939)
    float cpd[3], cgt[3], nv, ng;
    chroma3(vq, cpd, nv);
    chroma3(lin, cgt, ng);
    float dc[3], dot = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float e = cpd[c] - cgt[c];
      if (p.chroma_w > 0.f) terms[2] += p.chroma_w * e * e * (1.0f / 3.0f);
      dc[c] = p.chroma_w > 0.f ? p.chroma_w * 2.0f * e * (1.0f / 3.0f) : 0.f;
      dot += dc[c] * cpd[c];
    }
    const float inv_nv = nv == 0.f ? 0.f : 1.0f / nv;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float e2 = vq[c] - lin[c];
      p.d_vqrgb[r * 3 + c] = (2.0f * e2 * (1.0f / 3.0f) + (dc[c] - cpd[c] * dot) * inv_nv) * p.inv_gbs;
    }
    // lambert (:974-982): max_ch(spec) * where(sg(rough) < .5, 0, 2 sg(rough) - 1)
    if (p.lambert_w > 0.f) {
      const float s0 = p.spec[r * 3], s1 = p.spec[r * 3 + 1], s2 = p.spec[r * 3 + 2];
      int am = 0; float mx = s0;
      if (s1 > mx) { mx = s1; am = 1; }
      if (s2 > mx) { mx = s2; am = 2; }
      const float rg = p.rough[r];
      const float sg = rg < 0.5f ? 0.f : 2.0f * rg - 1.0f;
      terms[4] = p.lambert_w * mx * sg;
#pragma unroll
      for (int c = 0; c < 3; ++c) p.d_spec[r * 3 + c] = (c == am) ? p.lambert_w * sg * p.inv_gbs : 0.f;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) p.d_spec[r * 3 + c] = 0.f;
    }
    float ns;
    chroma3(gt, schr, ns);                                            // schr_gt on the sRGB ground truth (:935)
  }
  // pair smoothness (:942-956)
  float w = 0.f;
  {
    float d0 = schr[0] - __shfl_sync(0xffffffffu, schr[0], 1);
    float d1 = schr[1] - __shfl_sync(0xffffffffu, schr[1], 1);
    float d2 = schr[2] - __shfl_sync(0xffffffffu, schr[2], 1);
    float ce = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    ce = ce > p.chr_thres ? ce : 0.f;
    w = __shfl_sync(0xffffffffu, expf(-p.chr_alpha * ce), 0);
  }
  if (p.smooth_w > 0.f) {
    const float* za = p.z + (pair * 2) * p.zdim;
    const float* zb = za + p.zdim;
    float dotz = 0.f;
    for (int k = lane; k < p.zdim; k += 32) dotz = fmaf(za[k], zb[k], dotz);
    dotz = warp_sum(dotz);
    const float sl = p.smooth_w * w * (1.0f - dotz);
    terms[3] = sl;                                                    // both rows carry it
    const float coef = -2.0f * p.smooth_w * w * p.inv_gbs;            // d(2 sl)/d z_a = -2 w_s w z_b
    float* da = p.d_z + (pair * 2) * p.zdim;
    float* db = da + p.zdim;
    for (int k = lane; k < p.zdim; k += 32) { da[k] = coef * zb[k]; db[k] = coef * za[k]; }
  } else {
    float* da = p.d_z + (pair * 2) * p.zdim;
    for (int k = lane; k < 2 * p.zdim; k += 32) da[k] = 0.f;
  }
  if (lane < 2) {
    const float tot = terms[0] + terms[1] + terms[2] + terms[3] + terms[4];
    p.loss_rows[pair * 2 + lane] = tot;
  }
  // loss_dict sums (diagnostics; nondeterministic order is irrelevant at 1e-6)
  float t0 = terms[0], t1 = terms[1], t2 = terms[2], t3 = lane < 2 ? terms[3] : 0.f, t4 = terms[4];
  t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
  t2 += __shfl_xor_sync(0xffffffffu, t2, 1); t3 += __shfl_xor_sync(0xffffffffu, t3, 1);
  t4 += __shfl_xor_sync(0xffffffffu, t4, 1);
  // per block in shared memory first: 4096 warps adding to the same six addresses serialise in L2 (35 us of an 8192-row step)
  if (lane == 0 && p.sums) {
    atomicAdd(&bsum[0], t0); atomicAdd(&bsum[1], t1); atomicAdd(&bsum[2], t2); atomicAdd(&bsum[3], t3);
    atomicAdd(&bsum[4], t4); atomicAdd(&bsum[5], t0 + t1 + t2 + t3 + t4);
  }
  }
  __syncthreads();
  if (threadIdx.x < 6 && p.sums) atomicAdd(&p.sums[threadIdx.x], bsum[threadIdx.x]);
}

// z_vq = z_norm + sg(q - z_norm) (vq_layers.py:327) => d z_norm = d z_vq; commitment term
// vq_w * commitment * mean((sg(q) - z_norm)^2) (:302,321) => + coef (z_norm - q), coef = vq_w commit 2/(gbs Z);
// z_norm = z * rsqrt(max(sum z^2, 1e-6)) (util/math.py:63-64) => d z = inv (g - z_norm (g . z_norm)) (or inv g below eps)
__global__ void __launch_bounds__(256) vq_bwd_kernel(const float* __restrict__ z_enc, const long long* __restrict__ idx,
                                                     const float* __restrict__ cb, int K, const float* __restrict__ d_zvq,
                                                     float coef, long long n, int accumulate, float* __restrict__ d_zenc,
                                                     int act, float* __restrict__ dz_out, long long lddz) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* pz = z_enc + row * 256;
  const int k = (int)idx[row];
  float z[8], gq[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { z[i] = pz[lane + 32 * i]; s = fmaf(z[i], z[i], s); }
  s = warp_sum(s);
  const float inv = rsqrtf(fmaxf(s, 1e-6f));
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int zi = lane + 32 * i;
    const float zn = z[i] * inv;
    const float q = cb[(size_t)zi * K + k];
    gq[i] = d_zvq[row * 256 + zi] + coef * (zn - q);
    dot = fmaf(gq[i], zn, dot);
    z[i] = zn;
  }
  dot = warp_sum(dot);
  if (!(s > 1e-6f)) dot = 0.f;       // clamped norm: inv is a constant
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int zi = lane + 32 * i;
    const float v = inv * (gq[i] - z[i] * dot);
    float* dst = d_zenc + row * 256 + zi;
    const float tot = accumulate ? *dst + v : v;
    *dst = tot;
    if (dz_out) {                    // z_enc is the producing layer's activated output: also form that layer's dz
      const float y = pz[zi];
      dz_out[row * lddz + zi] = tot * (act == VQN_ACT_RELU ? (y > 0.f ? 1.f : 0.f) : act == VQN_ACT_SIGMOID ? y * (1.f - y) : 1.f);
    }
  }
}

__global__ void combine_bwd_kernel(const float* __restrict__ base, const float* __restrict__ ks,
                                   const float* __restrict__ d_albedo, const float* __restrict__ d_spec,
                                   const float* __restrict__ d_spec_extra, long long n, float* __restrict__ d_base,
                                   float* __restrict__ d_ks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float k = ks[i];
  float dk = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float ds = d_spec[i * 3 + c] + (d_spec_extra ? d_spec_extra[i * 3 + c] : 0.f);
    const float da = d_albedo[i * 3 + c];
    d_base[i * 3 + c] = k * ds + (1.0f - k) * da;        // spec = ks*base, albedo = (1-ks)*base (:590-591)
    dk = fmaf(base[i * 3 + c], ds - da, dk);
  }
  d_ks[i] = dk;
}

// sim_smooth (:958-972): -log(min_{i != j} || cn_i - cn_j ||), cn = get_codebook(raw) columns.  One block.
// The gradient reaches only the closest pair; the K diagonal zeros get weight 0 (TF's 0*0.5/0 there is defined
// as 0 here).
__global__ void __launch_bounds__(256) sim_loss_kernel(const float* raw, int Z, int K, float scale,
                                                       float* __restrict__ loss_out, float* __restrict__ d_raw,
                                                       int accumulate, int staged) {
  extern __shared__ float sm[];      // inv[K] | best_d[8] | best_pair[8] | (small codebooks) clip(raw) [Z*K]
  float* inv = sm;
  float* bd = inv + K;
  int* bp = reinterpret_cast<int*>(bd + 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // a small codebook (the shipped K = 15: 15 KB) is staged once with coalesced loads: the column reads below have a stride of
  // K floats, and from global memory every one of the ~14 dependent pair iterations per warp was an L2 round trip
  if (staged) {
    float* cs = reinterpret_cast<float*>(bp + 8);
    for (int i = tid; i < Z * K; i += 256) cs[i] = fminf(fmaxf(raw[i], 0.f), 1.f);
    __syncthreads();
    raw = cs;
  }
  for (int k = warp; k < K; k += 8) {
    float s = 0.f;
    for (int z = lane; z < Z; z += 32) { float c = fminf(fmaxf(raw[(size_t)z * K + k], 0.f), 1.f); s = fmaf(c, c, s); }
    s = warp_sum(s);
    if (lane == 0) inv[k] = rsqrtf(fmaxf(s, 1e-6f));
  }
  __syncthreads();
  const int npairs = K * (K - 1) / 2;
  float best = __int_as_float(0x7f800000);
  int best_pair = 0x7fffffff;
  for (int pr = warp; pr < npairs; pr += 8) {
    // pair index -> (i < j) in row-major order of the upper triangle
    int i = 0, rem = pr;
    while (rem >= K - 1 - i) { rem -= K - 1 - i; ++i; }
    const int j = i + 1 + rem;
    float s = 0.f;
    for (int z = lane; z < Z; z += 32) {
      const float ci = fminf(fmaxf(raw[(size_t)z * K + i], 0.f), 1.f) * inv[i];
      const float cj = fminf(fmaxf(raw[(size_t)z * K + j], 0.f), 1.f) * inv[j];
      const float d = ci - cj;
      s = fmaf(d, d, s);
    }
    s = warp_sum(s);
    if (s < best || (s == best && pr < best_pair)) { best = s; best_pair = pr; }
  }
  if (lane == 0) { bd[warp] = best; bp[warp] = best_pair; }
  __syncthreads();
  for (int w2 = 0; w2 < 8; ++w2)
    if (bd[w2] < best || (bd[w2] == best && bp[w2] < best_pair)) { best = bd[w2]; best_pair = bp[w2]; }
  if (!accumulate)
    for (int i = tid; i < Z * K; i += 256) d_raw[i] = 0.f;
  __syncthreads();
  if (npairs == 0) { if (tid == 0 && loss_out) *loss_out = 0.f; return; }
  int bi = 0, rem = best_pair;
  while (rem >= K - 1 - bi) { rem -= K - 1 - bi; ++bi; }
  const int bj = bi + 1 + rem;
  if (tid == 0 && loss_out) *loss_out = -0.5f * logf(best);           // -log(sqrt(D))
  // d(-log d)/d cn_i = -(cn_i - cn_j)/D ; then through the column normalisation (clip has identity gradient)
  if (warp < 2) {
    const int col = warp == 0 ? bi : bj;
    const float sign = warp == 0 ? -1.f : 1.f;
    float gv[8], cn[8];
    float dot = 0.f, ssq = 0.f;
    for (int q = 0; q < (Z + 31) / 32 && q < 8; ++q) {
      const int z = lane + 32 * q;
      gv[q] = 0.f; cn[q] = 0.f;
      if (z < Z) {
        const float ci_raw = fminf(fmaxf(raw[(size_t)z * K + bi], 0.f), 1.f);
        const float cj_raw = fminf(fmaxf(raw[(size_t)z * K + bj], 0.f), 1.f);
        const float ci = ci_raw * inv[bi], cj = cj_raw * inv[bj];
        gv[q] = sign * scale * (ci - cj) / best;
        cn[q] = warp == 0 ? ci : cj;
        const float cr = warp == 0 ? ci_raw : cj_raw;
        ssq = fmaf(cr, cr, ssq);
        dot = fmaf(gv[q], cn[q], dot);
      }
    }
    dot = warp_sum(dot);
    ssq = warp_sum(ssq);
    if (!(ssq > 1e-6f)) dot = 0.f;
    for (int q = 0; q < (Z + 31) / 32 && q < 8; ++q) {
      const int z = lane + 32 * q;
      if (z < Z) d_raw[(size_t)z * K + col] += inv[col] * (gv[q] - cn[q] * dot);
    }
  }
}

// tf.keras.optimizers.Adam(amsgrad=True) (optimizer_v2/adam.py): lr_t = lr sqrt(1-b2^t)/(1-b1^t) is computed by
// the caller; m += (g-m)(1-b1); v += (g^2-v)(1-b2); vhat = max(vhat, v); p -= lr_t m / (sqrt(vhat) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float* __restrict__ vhat, long long count, float lr_t,
                            const float* __restrict__ lr_t_dev, float b1, float b2, float eps) {
  if (lr_t_dev) lr_t = *lr_t_dev;      // CUDA-graph replay: the step-dependent rate lives in device memory
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.0f - b2);
    const float vh = fmaxf(vhat[i], vi);
    m[i] = mi; v[i] = vi; vhat[i] = vh;
    p[i] -= lr_t * mi / (sqrtf(vh) + eps);
  }
}

__global__ void copy_cols_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd,
                                 long long m, int w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * w) return;
  const long long r = i / w; const int c = (int)(i % w);
  dst[r * ldd + c] = src[r * lds + c];
}
// several column-block copies in ONE launch (the skip-concat halves and output compactions of the training step)
#define COPY_JOBS 16
struct CopyJob { const float* src; float* dst; long long lds, ldd, m; int w, pad_; };
struct CopyBatch { int count; int block_start[COPY_JOBS + 1]; CopyJob j[COPY_JOBS]; };
__global__ void copy_cols_batched_kernel(const __grid_constant__ CopyBatch b) {
  int q = 0;
  while (q + 1 < b.count && (int)blockIdx.x >= b.block_start[q + 1]) ++q;
  const CopyJob& j = b.j[q];
  const long long i = (long long)((int)blockIdx.x - b.block_start[q]) * blockDim.x + threadIdx.x;
  if (j.pad_) {                       // 16-byte form: width, leading dimensions and both bases are multiples of 4 floats
    const int w4 = j.w >> 2;
    if (i >= j.m * w4) return;
    const long long r = i / w4; const int c = (int)(i % w4) * 4;
    *reinterpret_cast<float4*>(j.dst + r * j.ldd + c) = *reinterpret_cast<const float4*>(j.src + r * j.lds + c);
    return;
  }
  if (i >= j.m * j.w) return;
  const long long r = i / j.w; const int c = (int)(i % j.w);
  j.dst[r * j.ldd + c] = j.src[r * j.lds + c];
}
// learnable tone scaling of non-'nerf' data (vq_nfr.py:715-718, 736-745); gpar = [_gamma_bias, _gamma_index] on the device
__global__ void gamma_fwd_kernel(const float* __restrict__ lin, const float* __restrict__ gpar, float* __restrict__ out,
                                 long long n) {
  const float g0 = gpar[0], g1 = fminf(fmaxf(gpar[1], 0.f), 5.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = fminf(fmaxf(powf(lin[i] * g0, g1), 0.f), 1.f);
}
__global__ void gamma_bwd_kernel(const float* __restrict__ lin, const float* __restrict__ gpar,
                                 const float* __restrict__ d_out, float* __restrict__ d_lin, float* __restrict__ d_gpar,
                                 long long n) {
  const float g0 = gpar[0], g1 = fminf(fmaxf(gpar[1], 0.f), 5.f);
  float s0 = 0.f, s1 = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = lin[i], u = x * g0, g = d_out[i];
    const float um1 = powf(u, g1 - 1.0f);            // d u^g1 / d u = g1 u^(g1 - 1)
    d_lin[i] = g * g1 * g0 * um1;
    s0 += g * g1 * x * um1;
    s1 += u > 0.f ? g * powf(u, g1) * logf(u) : 0.f; // TF's pow gradient uses log(x) where x > 0, else 0
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&d_gpar[0], s0); atomicAdd(&d_gpar[1], s1); }
}
__global__ void cast_f64_f32_kernel(const double* __restrict__ s, float* __restrict__ d, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (float)s[i];
}
// the buffers a training step accumulates into (gradients, VQ statistics, d_z) cleared in ONE launch; sizes in 4-byte words
#define ZERO_JOBS 8
struct ZeroBatch { int count; uint32_t* p[ZERO_JOBS]; long long words[ZERO_JOBS]; };
__global__ void zero_batched_kernel(const __grid_constant__ ZeroBatch b) {
  for (int q = 0; q < b.count; ++q) {
    uint32_t* p = b.p[q];
    const long long n = b.words[q];
    const long long stride = (long long)gridDim.x * blockDim.x, t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (((uintptr_t)p & 15) == 0) {
      const long long n4 = n >> 2;
      for (long long i = t; i < n4; i += stride) reinterpret_cast<uint4*>(p)[i] = make_uint4(0, 0, 0, 0);
      for (long long i = (n4 << 2) + t; i < n; i += stride) p[i] = 0;
    } else {
      for (long long i = t; i < n; i += stride) p[i] = 0;
    }
  }
}
__global__ void pack_stats_kernel(const double* __restrict__ s, float* __restrict__ d, long long n,
                                  float* __restrict__ rows_slot, float rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (float)s[i];
  if (i == 0 && rows_slot) rows_slot[0] += rows;
}
__global__ void cast_f32_f64_kernel(const float* __restrict__ s, double* __restrict__ d, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (double)s[i];
}

// the scalar tail of a training step (train_nfr.py:571, vq_nfr.py:971-986) in one thread instead of ~10 elementwise launches
__global__ void train_scalars_kernel(const float* __restrict__ sums, const float* __restrict__ vq_loss,
                                     const float* __restrict__ sim_loss, float inv_gbs, float vq_w, float sim_w,
                                     float* __restrict__ out) {
  if (blockIdx.x || threadIdx.x) return;
  const float rows_scale = sums[6] * inv_gbs;
  const float vq = vq_w * vq_loss[0];
  const float sim = sim_loss ? sim_w * sim_loss[0] : 0.f;
  out[0] = sums[5] * inv_gbs + rows_scale * (vq + sim);
  out[1] = vq;
  out[2] = sim;
  out[3] = rows_scale;
}

}  // namespace

extern "C" int vqn_shade_backward(vqn_ctx* ctx, const float* xyz, const float* rayo, const float* normal,
                                  const float* lvis, const int32_t* row_idx, int64_t n, const float* albedo,
                                  const float* spec, const float* rough, const float* lxyz, const float* lareas,
                                  const float* light, int clip_light0, const float* d_rgb, float* d_albedo,
                                  float* d_spec, float* d_rough, float* d_light, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && xyz && rayo && normal && albedo && spec && rough && lxyz && lareas && light && d_rgb &&
                    d_albedo && d_spec && d_rough && d_light, "shade_backward: null pointer");
  VQN_CHECK_ARG(n >= 0, "shade_backward: n < 0");
  if (n == 0) return VQN_OK;
  ShadeBwdParams a = {xyz, rayo, normal, lvis, row_idx, (long long)n, albedo, spec, rough, lxyz, lareas, light,
                      clip_light0, d_rgb, d_albedo, d_spec, d_rough, d_light};
  // two warps per point, two CTAs (16 warps, 128 registers) per SM: 40 us per 8192 points; four warps per point at three
  // CTAs per SM measured 58 us (the per-point set-up and reduction are repeated by every warp of a point), one warp 68-82 us
  const long long want = (n + 3) / 4, cap = (long long)ctx->sm_count * 2;
  shade_bwd_kernel<2, 2><<<(unsigned)(want < cap ? want : cap), SB_THREADS, 0, vqn_cs(stream)>>>(a);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_loss_train(vqn_ctx* ctx, const float* gtc, const float* rgb, const float* vqrgb, const float* z_vq,
                              const float* spec, const float* rough, int64_t n, int z_dim, int data_is_nerf,
                              float combine_weight, float chromaticity_weight, float mat_sloss_weight,
                              float lambert_weight, float chr_alpha, float chr_thres, float inv_global_bs,
                              float* loss_rows, float* d_rgb, float* d_vqrgb, float* d_z, float* d_spec, float* sums,
                              vqn_stream stream) {
  VQN_CHECK_ARG(ctx && gtc && rgb && vqrgb && z_vq && spec && rough && loss_rows && d_rgb && d_vqrgb && d_z && d_spec,
                "loss_train: null pointer");
  VQN_CHECK_ARG(n >= 0 && (n % 2) == 0, "loss_train: rows must come in (pixel, neighbour) pairs");
  if (n == 0) return VQN_OK;
  LossParams p = {gtc, rgb, vqrgb, z_vq, spec, rough, (long long)n, z_dim, data_is_nerf, combine_weight,
                  chromaticity_weight, mat_sloss_weight, lambert_weight, chr_alpha, chr_thres, inv_global_bs,
                  loss_rows, d_rgb, d_vqrgb, d_z, d_spec, sums};
  const long long warps = n / 2;
  loss_train_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, vqn_cs(stream)>>>(p);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_vq_backward_act(vqn_ctx* ctx, const float* z_enc, const int64_t* indices, const float* codebook, int k,
                                   const float* d_zvq, float commit_coef, int64_t n, int accumulate, float* d_zenc,
                                   int act, float* dz_out, int64_t lddz, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && z_enc && indices && codebook && d_zvq && d_zenc && k > 0 && n >= 0, "vq_backward args");
  VQN_CHECK_ARG(!dz_out || lddz >= 256, "vq_backward_act: lddz");
  if (n == 0) return VQN_OK;
  vq_bwd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, vqn_cs(stream)>>>(z_enc, (const long long*)indices, codebook, k,
                                                                    d_zvq, commit_coef, (long long)n, accumulate, d_zenc,
                                                                    act, dz_out, (long long)lddz);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
extern "C" int vqn_vq_backward(vqn_ctx* ctx, const float* z_enc, const int64_t* indices, const float* codebook, int k,
                               const float* d_zvq, float commit_coef, int64_t n, int accumulate, float* d_zenc,
                               vqn_stream stream) {
  return vqn_vq_backward_act(ctx, z_enc, indices, codebook, k, d_zvq, commit_coef, n, accumulate, d_zenc, VQN_ACT_NONE,
                             nullptr, 0, stream);
}

/* memset(ptrs[i], 0, bytes[i]) for up to 8 device buffers (4-byte aligned, sizes multiples of 4) in ONE launch */
extern "C" int vqn_zero_batched(vqn_ctx* ctx, void* const* ptrs, const int64_t* bytes, int count, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && ptrs && bytes && count >= 0 && count <= ZERO_JOBS, "zero_batched: 0..8 buffers");
  ZeroBatch b;
  b.count = 0;
  long long most = 0;
  for (int i = 0; i < count; ++i) {
    VQN_CHECK_ARG(bytes[i] >= 0 && bytes[i] % 4 == 0 && (bytes[i] == 0 || (ptrs[i] && (uintptr_t)ptrs[i] % 4 == 0)),
                  "zero_batched: buffers must be 4-byte aligned with sizes in multiples of 4");
    if (bytes[i] == 0) continue;
    b.p[b.count] = (uint32_t*)ptrs[i]; b.words[b.count] = bytes[i] / 4; ++b.count;
    most = most > bytes[i] / 16 ? most : bytes[i] / 16;
  }
  if (b.count == 0) return VQN_OK;
  long long blocks = (most + 255) / 256;
  blocks = blocks < 1 ? 1 : blocks > 4 * (long long)ctx->sm_count ? 4 * (long long)ctx->sm_count : blocks;
  zero_batched_kernel<<<(unsigned)blocks, 256, 0, vqn_cs(stream)>>>(b);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

/* stats32 = (float) stats64 for the gradient all-reduce buffer, and rows_slot[0] += rows (this rank's active rows) */
extern "C" int vqn_train_pack_stats(vqn_ctx* ctx, const double* stats64, float* stats32, int64_t count, float* rows_slot,
                                    float rows, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && stats64 && stats32 && count > 0, "train_pack_stats args");
  pack_stats_kernel<<<(unsigned)((count + 255) / 256), 256, 0, vqn_cs(stream)>>>(stats64, stats32, (long long)count,
                                                                                rows_slot, rows);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_material_combine_backward(vqn_ctx* ctx, const float* basecolor, const float* ks,
                                             const float* d_albedo, const float* d_spec, const float* d_spec_extra,
                                             int64_t n, float* d_basecolor, float* d_ks, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && basecolor && ks && d_albedo && d_spec && d_basecolor && d_ks && n >= 0, "combine_backward args");
  if (n == 0) return VQN_OK;
  combine_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, vqn_cs(stream)>>>(basecolor, ks, d_albedo, d_spec,
                                                                             d_spec_extra, (long long)n, d_basecolor, d_ks);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_codebook_sim_loss(vqn_ctx* ctx, const float* raw_codebook, int z_dim, int k, float grad_scale,
                                     float* loss_out, float* d_raw, int accumulate, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && raw_codebook && d_raw && z_dim > 0 && z_dim <= 256 && k >= 1 && k <= 1024, "sim_loss args");
  const int staged = (size_t)z_dim * k * sizeof(float) <= 40 * 1024 ? 1 : 0;
  sim_loss_kernel<<<1, 256, sizeof(float) * (k + 16 + (staged ? (size_t)z_dim * k : 0)), vqn_cs(stream)>>>(
      raw_codebook, z_dim, k, grad_scale, loss_out, d_raw, accumulate, staged);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_adam_amsgrad(vqn_ctx* ctx, float* param, const float* grad, float* m, float* v, float* vhat,
                                int64_t count, float lr_t, const float* lr_t_dev, float beta1, float beta2,
                                float epsilon, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && param && grad && m && v && vhat && count >= 0, "adam args");
  if (count == 0) return VQN_OK;
  long long want = (count + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  adam_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(param, grad, m, v, vhat, (long long)count, lr_t, lr_t_dev, beta1, beta2,
                                                  epsilon);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_copy_cols(vqn_ctx* ctx, const float* src, int64_t lds, float* dst, int64_t ldd, int64_t m, int w,
                             vqn_stream stream) {
  VQN_CHECK_ARG(ctx && src && dst && m >= 0 && w > 0 && lds >= w && ldd >= w, "copy_cols args");
  if (m == 0) return VQN_OK;
  const long long total = (long long)m * w;
  copy_cols_kernel<<<(unsigned)((total + 255) / 256), 256, 0, vqn_cs(stream)>>>(src, lds, dst, ldd, (long long)m, w);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_gamma_forward(vqn_ctx* ctx, const float* lin, const float* gpar, float* out, int64_t count,
                                 vqn_stream stream) {
  VQN_CHECK_ARG(ctx && lin && gpar && out && count >= 0, "gamma_forward args");
  if (count == 0) return VQN_OK;
  long long want = (count + 255) / 256;
  gamma_fwd_kernel<<<(unsigned)(want < 1184 ? want : 1184), 256, 0, vqn_cs(stream)>>>(lin, gpar, out, (long long)count);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
extern "C" int vqn_gamma_backward(vqn_ctx* ctx, const float* lin, const float* gpar, const float* d_out, float* d_lin,
                                  float* d_gpar, int64_t count, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && lin && gpar && d_out && d_lin && d_gpar && count >= 0, "gamma_backward args");
  if (count == 0) return VQN_OK;
  long long want = (count + 255) / 256;
  gamma_bwd_kernel<<<(unsigned)(want < 592 ? want : 592), 256, 0, vqn_cs(stream)>>>(lin, gpar, d_out, d_lin, d_gpar,
                                                                                   (long long)count);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_copy_cols_batched(vqn_ctx* ctx, const vqn_copy_job* jobs, int count, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && jobs && count >= 1 && count <= COPY_JOBS, "copy_cols_batched: 1 <= count <= 16");
  CopyBatch b;
  b.count = 0;
  int blocks = 0;
  for (int q = 0; q < count; ++q) {
    const vqn_copy_job& c = jobs[q];
    VQN_CHECK_ARG(c.src && c.dst && c.m >= 0 && c.w > 0 && c.lds >= c.w && c.ldd >= c.w, "copy_cols_batched: bad job");
    if (c.m == 0) continue;
    const int vec = (c.w % 4 == 0 && c.lds % 4 == 0 && c.ldd % 4 == 0 &&
                     ((uintptr_t)c.src | (uintptr_t)c.dst) % 16 == 0) ? 1 : 0;
    b.j[b.count] = {c.src, c.dst, (long long)c.lds, (long long)c.ldd, (long long)c.m, c.w, vec};
    b.block_start[b.count] = blocks;
    blocks += (int)(((long long)c.m * (vec ? c.w / 4 : c.w) + 255) / 256);
    ++b.count;
  }
  if (b.count == 0) return VQN_OK;
  b.block_start[b.count] = blocks;
  copy_cols_batched_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(b);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

extern "C" int vqn_cast_f64_f32(vqn_ctx* ctx, const double* src, float* dst, int64_t count, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && src && dst && count >= 0, "cast args");
  if (count == 0) return VQN_OK;
  cast_f64_f32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, vqn_cs(stream)>>>(src, dst, (long long)count);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
/* out[4] = { weighted loss, vq_w * vq_loss, sim_w * sim_loss, rows / global_bs } from the reduced sums
 * (sums[5] = total of the per-example terms, sums[6] = active rows); sim_loss may be NULL (weight 0). */
extern "C" int vqn_train_scalars(vqn_ctx* ctx, const float* sums, const float* vq_loss, const float* sim_loss,
                                 float inv_gbs, float vq_w, float sim_w, float* out, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && sums && vq_loss && out, "train_scalars: null pointer");
  train_scalars_kernel<<<1, 32, 0, vqn_cs(stream)>>>(sums, vq_loss, sim_loss, inv_gbs, vq_w, sim_w, out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
extern "C" int vqn_cast_f32_f64(vqn_ctx* ctx, const float* src, double* dst, int64_t count, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && src && dst && count >= 0, "cast args");
  if (count == 0) return VQN_OK;
  cast_f32_f64_kernel<<<(unsigned)((count + 255) / 256), 256, 0, vqn_cs(stream)>>>(src, dst, (long long)count);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
