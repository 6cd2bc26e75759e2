// NeuS geo-stage per-ray warp-scan kernels (secondary path).
//
// Reference: geo/NeuS-ours2/models/renderer.py  sample_pdf :39-69, up_sample :131-175,
// cat_z_vals :177-191, render_core :193-297.  One warp owns one ray; lane l owns samples
// [4l, 4l+4) (n_samples <= 128), so every cumulative product / sum is a 4-element serial scan followed
// by a 5-step warp shuffle scan, and neighbour samples come from lane l+1 by shuffle.  All per-ray state
// stays in registers; HBM traffic is exactly the rows read and written once (memory-bound kernels).
#include "common.cuh"

#define NS_MAXS 128
#define NS_PER 4

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// exclusive product scan over 128 blocked elements; v[j] in: factors, out: exclusive prefix products
__device__ __forceinline__ void warp_excl_cumprod4(float (&v)[NS_PER], int lane) {
  float loc[NS_PER];
  float run = 1.f;
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) { loc[j] = run; run *= v[j]; }
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= y;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.f;
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) v[j] = excl * loc[j];
}

// inclusive sum scan over 128 blocked elements
__device__ __forceinline__ void warp_incl_cumsum4(float (&v)[NS_PER], int lane) {
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) { run += v[j]; v[j] = run; }
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0.f;
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) v[j] += excl;
}

// ---------------------------------------------------------------------------------------------
// up_sample + sample_pdf(det=True)
// ---------------------------------------------------------------------------------------------
// One ray's up_sample (renderer.py:131-175 + sample_pdf det=True :39-69) by one warp.  zsrc / sdsrc: the ray's S sorted depths
// and SDF values (global or shared memory); s_cdf / s_z: the warp's scratch rows.  Lane q writes sample q + 32 k to
// z_samples (global, may be NULL), s_new (shared, may be NULL) and its position o + d z to pts_out (global, may be NULL).
__device__ __forceinline__ void up_sample_ray(const float* zsrc, const float* sdsrc, int S, float ox, float oy, float oz,
                                              float dx, float dy, float dz, float r_limit, int n_imp, float inv_s,
                                              float* s_cdf, float* s_z, float* z_samples, float* s_new, float* pts_out,
                                              int lane) {
  float z[NS_PER + 1], sd[NS_PER + 1], rad[NS_PER + 1];
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) {
    int i = NS_PER * lane + j;
    z[j] = i < S ? zsrc[i] : 0.f;
    sd[j] = i < S ? sdsrc[i] : 0.f;
    float px = ox + dx * z[j], py = oy + dy * z[j], pz = oz + dz * z[j];
    rad[j] = sqrtf(px * px + py * py + pz * pz);     // torch.linalg.norm(pts, ord=2)
  }
  // element NS_PER = first element of the next lane
  z[NS_PER] = __shfl_down_sync(0xffffffffu, z[0], 1);
  sd[NS_PER] = __shfl_down_sync(0xffffffffu, sd[0], 1);
  rad[NS_PER] = __shfl_down_sync(0xffffffffu, rad[0], 1);
  // cos_val per interval i (valid for i < S-1)
  float cosv[NS_PER];
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) cosv[j] = (sd[j + 1] - sd[j]) / (z[j + 1] - z[j] + 1e-5f);
  float prev_last = __shfl_up_sync(0xffffffffu, cosv[NS_PER - 1], 1);
  if (lane == 0) prev_last = 0.f;                     // prev_cos_val[:,0] = 0
  float om[NS_PER], alpha[NS_PER];
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) {
    int i = NS_PER * lane + j;
    float pc = j == 0 ? prev_last : cosv[j - 1];
    float c = fminf(pc, cosv[j]);
    c = fminf(fmaxf(c, -1e3f), 0.f);
    bool inside = (rad[j] < r_limit) || (rad[j + 1] < r_limit);
    c = inside ? c : 0.f;
    float mid = (sd[j] + sd[j + 1]) * 0.5f;
    float dist = z[j + 1] - z[j];
    float pe = mid - c * dist * 0.5f, ne = mid + c * dist * 0.5f;
    float pcdf = sigmoidf_(pe * inv_s), ncdf = sigmoidf_(ne * inv_s);
    float a = (pcdf - ncdf + 1e-5f) / (pcdf + 1e-5f);
    bool valid = i < S - 1;
    alpha[j] = valid ? a : 0.f;
    om[j] = valid ? (1.f - a + 1e-7f) : 1.f;
  }
  warp_excl_cumprod4(om, lane);                        // transmittance
  float w[NS_PER];
  float wsum = 0.f;
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) {
    int i = NS_PER * lane + j;
    w[j] = i < S - 1 ? alpha[j] * om[j] + 1e-5f : 0.f;  // sample_pdf: weights + 1e-5
    wsum += w[j];
  }
  wsum = warp_sum(wsum);
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) w[j] = w[j] / wsum;  // pdf
  warp_incl_cumsum4(w, lane);                          // cdf[1..S-1]
  __syncwarp();
  if (lane == 0) s_cdf[0] = 0.f;
#pragma unroll
  for (int j = 0; j < NS_PER; ++j) {
    int i = NS_PER * lane + j;
    if (i < S - 1) s_cdf[i + 1] = w[j];
    if (i < S) s_z[i] = z[j];
  }
  __syncwarp();
  // inverse CDF at u = linspace(0.5/n, 1-0.5/n, n)
  for (int q = lane; q < n_imp; q += 32) {
    float lo = 0.5f / n_imp, hi = 1.0f - 0.5f / n_imp;
    float u = n_imp > 1 ? lo + (hi - lo) * ((float)q / (float)(n_imp - 1)) : lo;
    if (q == n_imp - 1 && n_imp > 1) u = hi;
    // searchsorted(cdf, u, right=True): first idx with cdf[idx] > u, in [0, S]
    int a0 = 0, b0 = S;
    while (a0 < b0) { int m = (a0 + b0) >> 1; if (s_cdf[m] > u) b0 = m; else a0 = m + 1; }
    int below = max(0, a0 - 1), above = min(S - 1, a0);
    float cb = s_cdf[below], ca = s_cdf[above];
    float bb = s_z[below], ba = s_z[above];
    float den = ca - cb;
    if (den < 1e-5f) den = 1.f;
    float t = (u - cb) / den;
    const float zs = bb + t * (ba - bb);
    if (z_samples) z_samples[q] = zs;
    if (s_new) s_new[q] = zs;
    if (pts_out) { pts_out[q * 3] = ox + dx * zs; pts_out[q * 3 + 1] = oy + dy * zs; pts_out[q * 3 + 2] = oz + dz * zs; }
  }
  __syncwarp();
}

__global__ void neus_up_sample_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                      const float* __restrict__ z_vals, const float* __restrict__ sdf,
                                      long long n_rays, int S, float r_limit, int n_imp, float inv_s,
                                      float* __restrict__ z_samples, float* __restrict__ pts_out) {
  __shared__ float s_cdf[8][NS_MAXS + 1];
  __shared__ float s_z[8][NS_MAXS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ray = warp; ray < n_rays; ray += nw) {
    const float ox = rays_o[ray * 3], oy = rays_o[ray * 3 + 1], oz = rays_o[ray * 3 + 2];
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    up_sample_ray(z_vals + ray * S, sdf + ray * S, S, ox, oy, oz, dx, dy, dz, r_limit, n_imp, inv_s, s_cdf[wib], s_z[wib],
                  z_samples + ray * n_imp, nullptr, pts_out ? pts_out + ray * n_imp * 3 : nullptr, lane);
  }
}

extern "C" int vqn_neus_up_sample(vqn_ctx* ctx, const float* rays_o, const float* rays_d, const float* z_vals,
                                  const float* sdf, int64_t n_rays, int n_samples, float r_limit, int n_importance,
                                  float inv_s, float* z_samples, vqn_stream stream) {
  return vqn_neus_up_sample_pts(ctx, rays_o, rays_d, z_vals, sdf, n_rays, n_samples, r_limit, n_importance, inv_s, z_samples,
                                nullptr, stream);
}

extern "C" int vqn_neus_up_sample_pts(vqn_ctx* ctx, const float* rays_o, const float* rays_d, const float* z_vals,
                                      const float* sdf, int64_t n_rays, int n_samples, float r_limit, int n_importance,
                                      float inv_s, float* z_samples, float* pts_out, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && rays_o && rays_d && z_vals && sdf && z_samples, "up_sample: null");
  VQN_CHECK_ARG(n_samples >= 2 && n_samples <= NS_MAXS, "up_sample: 2 <= n_samples <= 128");
  VQN_CHECK_ARG(n_importance >= 1 && n_rays >= 0, "up_sample: n_importance >= 1");
  if (n_rays == 0) return VQN_OK;
  long long want = (n_rays + 7) / 8;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  neus_up_sample_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(rays_o, rays_d, z_vals, sdf, n_rays, n_samples,
                                                            r_limit, n_importance, inv_s, z_samples, pts_out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// cat_z_vals: stable rank sort of [z_vals ; new_z] carrying sdf
// ---------------------------------------------------------------------------------------------
__global__ void neus_cat_kernel(const float* __restrict__ z_vals, const float* __restrict__ new_z,
                                const float* __restrict__ sdf, const float* __restrict__ new_sdf, long long n_rays,
                                int S, int I, float* __restrict__ z_out, float* __restrict__ sdf_out);   // (below rank_merge)

extern "C" int vqn_neus_cat_z_vals(vqn_ctx* ctx, const float* z_vals, const float* new_z, const float* sdf,
                                   const float* new_sdf, int64_t n_rays, int n_samples, int n_importance,
                                   float* z_out, float* sdf_out, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && z_vals && new_z && z_out, "cat_z_vals: null");
  VQN_CHECK_ARG(n_samples >= 1 && n_importance >= 1 && n_samples + n_importance <= 2 * NS_MAXS,
                "cat_z_vals: n_samples + n_importance <= 256");
  VQN_CHECK_ARG(!sdf_out || sdf, "cat_z_vals: sdf_out needs sdf");
  if (n_rays == 0) return VQN_OK;
  long long want = (n_rays + 7) / 8;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  neus_cat_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(z_vals, new_z, sdf, new_sdf, n_rays, n_samples,
                                                      n_importance, z_out, sdf_out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// One hierarchical-sampling step in ONE launch (NeuSRenderer.render, renderer.py:343-366): cat_z_vals of step i (the SDF
// of the new samples has just been evaluated), up_sample of step i + 1 on the merged row while it is still in shared
// memory -- including the positions o + d z of the new samples, the next SDF call's input -- and, for the last step
// (whose cat_z_vals needs no SDF), the final merge and the mid-point positions / directions render_core asks for.
// 4 + 4 + 1 launches per render become 4; the same arithmetic as the separate kernels, bit for bit.
// ---------------------------------------------------------------------------------------------
// Stable sort of s_in[0..T) = [list A (S elements) ; list B (T - S elements)] -- the order torch.sort gives the reference
// (cat_z_vals, renderer.py:181-182): element i goes to position #{k : in[k] < in[i] or (in[k] == in[i] and k < i)}.
// Both lists are normally sorted already (A is a previous merge, B an inverse CDF at increasing u), and then the position is
// i + #{b < A_i} resp. j + #{a <= B_j}: two binary searches instead of T comparisons per element (T^2 = 16 k per ray was a
// quarter of the scan kernels' time).  An unsorted input (possible through the public cat_z_vals) takes the rank sort.
__device__ __forceinline__ void rank_merge(const float* s_in, int S, int T, float* z_dst, const float* sd_in, float* sd_dst,
                                           int lane) {
  bool sorted = true;
  for (int i = lane; i + 1 < T; i += 32) sorted = sorted && (i + 1 == S || s_in[i] <= s_in[i + 1]);
  sorted = __all_sync(0xffffffffu, sorted);
  for (int i = lane; i < T; i += 32) {
    const float v = s_in[i];
    int rank;
    if (sorted) {
      int lo, hi;
      if (i < S) {                    // #{b in B : b < v}
        lo = S; hi = T;
        while (lo < hi) { const int m = (lo + hi) >> 1; if (s_in[m] < v) lo = m + 1; else hi = m; }
        rank = i + (lo - S);
      } else {                        // #{a in A : a <= v}
        lo = 0; hi = S;
        while (lo < hi) { const int m = (lo + hi) >> 1; if (s_in[m] <= v) lo = m + 1; else hi = m; }
        rank = (i - S) + lo;
      }
    } else {
      rank = 0;
      for (int k = 0; k < T; ++k) {
        const float u = s_in[k];
        rank += (u < v) || (u == v && k < i);
      }
    }
    z_dst[rank] = v;
    if (sd_dst) sd_dst[rank] = sd_in[i];
  }
}

__global__ void neus_cat_kernel(const float* __restrict__ z_vals, const float* __restrict__ new_z,
                                const float* __restrict__ sdf, const float* __restrict__ new_sdf, long long n_rays,
                                int S, int I, float* __restrict__ z_out, float* __restrict__ sdf_out) {
  __shared__ float s_z[8][2 * NS_MAXS], s_sd[8][2 * NS_MAXS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int T = S + I;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ray = warp; ray < n_rays; ray += nw) {
    for (int i = lane; i < T; i += 32) {
      s_z[wib][i] = i < S ? z_vals[ray * S + i] : new_z[ray * I + (i - S)];
      if (sdf_out) s_sd[wib][i] = i < S ? sdf[ray * S + i] : (new_sdf ? new_sdf[ray * I + (i - S)] : 0.f);
    }
    __syncwarp();
    rank_merge(s_z[wib], S, T, z_out + ray * T, s_sd[wib], sdf_out ? sdf_out + ray * T : nullptr, lane);
    __syncwarp();
  }
}

__global__ void neus_scan_step_kernel(vqn_neus_step_args a) {
  __shared__ float s_in[8][2 * NS_MAXS], s_sdin[8][2 * NS_MAXS];      // unsorted [z_vals ; new_z] and their SDF values
  __shared__ float s_zm[8][NS_MAXS], s_sdm[8][NS_MAXS];               // merged row (T <= 128 when it is up-sampled)
  __shared__ float s_cdf[8][NS_MAXS + 1], s_z[8][NS_MAXS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int S = a.n_samples, I = a.n_new, T = S + I, I2 = a.n_importance, F = T + I2;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ray = warp; ray < a.n_rays; ray += nw) {
    const float ox = a.rays_o[ray * 3], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
    const float dx = a.rays_d[ray * 3], dy = a.rays_d[ray * 3 + 1], dz = a.rays_d[ray * 3 + 2];
    for (int i = lane; i < T; i += 32) {
      s_in[wib][i] = i < S ? a.z_vals[ray * S + i] : a.new_z[ray * I + (i - S)];
      s_sdin[wib][i] = i < S ? a.sdf[ray * S + i] : a.new_sdf[ray * I + (i - S)];
    }
    __syncwarp();
    rank_merge(s_in[wib], S, T, s_zm[wib], s_sdin[wib], s_sdm[wib], lane);
    __syncwarp();
    if (a.z_out) for (int i = lane; i < T; i += 32) { a.z_out[ray * T + i] = s_zm[wib][i]; a.sdf_out[ray * T + i] = s_sdm[wib][i]; }
    // up_sample of the next step on the merged row; the new samples land behind it in s_in for the final merge
    if (a.final_merge) for (int i = lane; i < T; i += 32) s_in[wib][i] = s_zm[wib][i];
    up_sample_ray(s_zm[wib], s_sdm[wib], T, ox, oy, oz, dx, dy, dz, a.r_limit, I2, a.inv_s, s_cdf[wib], s_z[wib],
                  a.new_z_out ? a.new_z_out + ray * I2 : nullptr, a.final_merge ? s_in[wib] + T : nullptr,
                  a.pts_out ? a.pts_out + ray * I2 * 3 : nullptr, lane);
    if (a.final_merge) {
      // last step: cat_z_vals(last=True) needs no SDF; then render_core's mid-point positions (:203-209)
      __syncwarp();
      rank_merge(s_in[wib], T, F, s_sdin[wib], nullptr, nullptr, lane);    // sorted depths -> s_sdin (free by now)
      __syncwarp();
      for (int i = lane; i < F; i += 32) {
        const float z = s_sdin[wib][i];
        const float dist = i + 1 < F ? s_sdin[wib][i + 1] - z : a.sample_dist;
        const float mz = z + dist * 0.5f;
        const long long idx = ray * F + i;
        a.z_final[idx] = z;
        a.mid_pts[idx * 3] = ox + dx * mz; a.mid_pts[idx * 3 + 1] = oy + dy * mz; a.mid_pts[idx * 3 + 2] = oz + dz * mz;
        if (a.mid_dirs) { a.mid_dirs[idx * 3] = dx; a.mid_dirs[idx * 3 + 1] = dy; a.mid_dirs[idx * 3 + 2] = dz; }
      }
    }
    __syncwarp();
  }
}

extern "C" int vqn_neus_scan_step(vqn_ctx* ctx, const vqn_neus_step_args* args, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && args, "scan_step: null");
  const vqn_neus_step_args& a = *args;
  VQN_CHECK_ARG(a.rays_o && a.rays_d && a.z_vals && a.new_z && a.sdf && a.new_sdf, "scan_step: null input");
  VQN_CHECK_ARG(a.n_samples >= 1 && a.n_new >= 1 && a.n_samples + a.n_new >= 2 && a.n_samples + a.n_new <= NS_MAXS,
                "scan_step: 2 <= n_samples + n_new <= 128");
  VQN_CHECK_ARG(a.n_importance >= 1 && a.n_samples + a.n_new + a.n_importance <= 2 * NS_MAXS, "scan_step: n_importance");
  VQN_CHECK_ARG((a.z_out == nullptr) == (a.sdf_out == nullptr), "scan_step: z_out and sdf_out go together");
  VQN_CHECK_ARG(!a.final_merge || (a.z_final && a.mid_pts), "scan_step: final_merge needs z_final and mid_pts");
  VQN_CHECK_ARG(a.final_merge || a.new_z_out, "scan_step: new_z_out missing");
  if (a.n_rays == 0) return VQN_OK;
  long long want = (a.n_rays + 7) / 8;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  neus_scan_step_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(a);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// mid-point sample positions for the SDF / colour networks
// ---------------------------------------------------------------------------------------------
__global__ void neus_mid_points_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                       const float* __restrict__ z_vals, long long n_rays, int S, float sample_dist,
                                       float* __restrict__ pts, float* __restrict__ dirs) {
  long long total = n_rays * S;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long ray = idx / S;
    int i = (int)(idx - ray * S);
    float z = z_vals[idx];
    float dist = i + 1 < S ? z_vals[idx + 1] - z : sample_dist;
    float mz = z + dist * 0.5f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float d = rays_d[ray * 3 + c];
      pts[idx * 3 + c] = rays_o[ray * 3 + c] + d * mz;
      if (dirs) dirs[idx * 3 + c] = d;
    }
  }
}

extern "C" int vqn_neus_mid_points(vqn_ctx* ctx, const float* rays_o, const float* rays_d, const float* z_vals,
                                   int64_t n_rays, int n_samples, float sample_dist, float* pts, float* dirs,
                                   vqn_stream stream) {
  VQN_CHECK_ARG(ctx && rays_o && rays_d && z_vals && pts && n_samples >= 1 && n_rays >= 0, "mid_points args");
  if (n_rays == 0) return VQN_OK;
  long long want = (n_rays * n_samples + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  neus_mid_points_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(rays_o, rays_d, z_vals, n_rays, n_samples, sample_dist,
                                                             pts, dirs);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// render_core compositing
// ---------------------------------------------------------------------------------------------
__global__ void neus_composite_kernel(vqn_neus_composite_args a) {
  const int lane = threadIdx.x & 31;
  const int S = a.n_samples;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  double ge_num = 0.0, ge_den = 0.0;
  for (long long ray = warp; ray < a.n_rays; ray += nw) {
    const float ox = a.rays_o[ray * 3], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
    const float dx = a.rays_d[ray * 3], dy = a.rays_d[ray * 3 + 1], dz = a.rays_d[ray * 3 + 2];
    float z[NS_PER];
#pragma unroll
    for (int j = 0; j < NS_PER; ++j) { int i = NS_PER * lane + j; z[j] = i < S ? a.z_vals[ray * S + i] : 0.f; }
    float znext0 = __shfl_down_sync(0xffffffffu, z[0], 1);
    float alpha[NS_PER], om[NS_PER], px[NS_PER], py[NS_PER], pz[NS_PER];
#pragma unroll
    for (int j = 0; j < NS_PER; ++j) {
      const int i = NS_PER * lane + j;
      const bool valid = i < S;
      const long long si = ray * S + (valid ? i : 0);
      float zn = j + 1 < NS_PER ? z[j + 1] : znext0;
      float dist = (i + 1 < S) ? zn - z[j] : a.sample_dist;          // :203-204
      float mz = z[j] + dist * 0.5f;
      px[j] = ox + dx * mz; py[j] = oy + dy * mz; pz[j] = oz + dz * mz;
      float sdfv = a.sdf[si];
      float gx = a.gradients[si * 3], gy = a.gradients[si * 3 + 1], gz = a.gradients[si * 3 + 2];
      float true_cos = dx * gx + dy * gy + dz * gz;
      float iter_cos = -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.0f - a.cos_anneal_ratio) +
                         fmaxf(-true_cos, 0.f) * a.cos_anneal_ratio);  // :235-236
      float en = sdfv + iter_cos * dist * 0.5f, ep = sdfv - iter_cos * dist * 0.5f;
      float pcdf = sigmoidf_(ep * a.inv_s), ncdf = sigmoidf_(en * a.inv_s);
      float al = (pcdf - ncdf + 1e-5f) / (pcdf + 1e-5f);
      al = fminf(fmaxf(al, 0.f), 1.f);
      alpha[j] = valid ? al : 0.f;
      om[j] = valid ? 1.f - al + 1e-7f : 1.f;
      float radius = sqrtf(px[j] * px[j] + py[j] * py[j] + pz[j] * pz[j]);
      if (valid) {
        if (a.cdf) a.cdf[si] = pcdf;
        if (a.inside_sphere) a.inside_sphere[si] = radius < a.radius ? 1.f : 0.f;
        if (a.mid_z_vals) a.mid_z_vals[si] = mz;
        if (a.dists) a.dists[si] = dist;
        if (a.grad_err_sums) {
          float relax = radius < a.radius * 1.1f ? 1.f : 0.f;
          float gn = sqrtf(gx * gx + gy * gy + gz * gz) - 1.0f;
          ge_num += (double)(relax * gn * gn);
          ge_den += (double)relax;
        }
      }
    }
    warp_excl_cumprod4(om, lane);
    float wsum = 0.f, wmax = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NS_PER; ++j) {
      const int i = NS_PER * lane + j;
      if (i < S) {
        const long long si = ray * S + i;
        float w = alpha[j] * om[j];
        if (a.weights) a.weights[si] = w;
        wsum += w; wmax = fmaxf(wmax, w);
        c0 = fmaf(a.sampled_color[si * 3], w, c0);
        c1 = fmaf(a.sampled_color[si * 3 + 1], w, c1);
        c2 = fmaf(a.sampled_color[si * 3 + 2], w, c2);
        s0 = fmaf(px[j], w, s0); s1 = fmaf(py[j], w, s1); s2 = fmaf(pz[j], w, s2);
      }
    }
    wsum = warp_sum(wsum);
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) {
      if (a.background_rgb) {
        c0 += a.background_rgb[0] * (1.f - wsum); c1 += a.background_rgb[1] * (1.f - wsum);
        c2 += a.background_rgb[2] * (1.f - wsum);
      }
      if (a.color) { a.color[ray * 3] = c0; a.color[ray * 3 + 1] = c1; a.color[ray * 3 + 2] = c2; }
      if (a.surf) { a.surf[ray * 3] = s0; a.surf[ray * 3 + 1] = s1; a.surf[ray * 3 + 2] = s2; }
      if (a.depth) {
        float ex = s0 - ox, ey = s1 - oy, ez = s2 - oz;
        a.depth[ray] = sqrtf(ex * ex + ey * ey + ez * ez);
      }
      if (a.weight_sum) a.weight_sum[ray] = wsum;
      if (a.weight_max) a.weight_max[ray] = wmax;
    }
  }
  if (a.grad_err_sums) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ge_num += __shfl_xor_sync(0xffffffffu, ge_num, o);
      ge_den += __shfl_xor_sync(0xffffffffu, ge_den, o);
    }
    if (lane == 0 && (ge_num != 0.0 || ge_den != 0.0)) {
      atomicAdd(&a.grad_err_sums[0], ge_num);
      atomicAdd(&a.grad_err_sums[1], ge_den);
    }
  }
}

extern "C" int vqn_neus_composite(vqn_ctx* ctx, const vqn_neus_composite_args* args, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && args, "composite: null");
  const vqn_neus_composite_args& a = *args;
  VQN_CHECK_ARG(a.rays_o && a.rays_d && a.z_vals && a.sdf && a.gradients && a.sampled_color, "composite: null input");
  VQN_CHECK_ARG(a.n_samples >= 1 && a.n_samples <= NS_MAXS && a.n_rays >= 0, "composite: 1 <= n_samples <= 128");
  if (a.n_rays == 0) return VQN_OK;
  long long want = (a.n_rays + 7) / 8;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  neus_composite_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(a);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// RenderingNetwork input columns (fields.py:147-156, mode 'idr'; embedder.py: [x, sin(x f), cos(x f), ...], f = 2^k)
// one thread per (point, column): coalesced 4-byte stores into the strided row buffer
// ---------------------------------------------------------------------------------------------
__global__ void neus_color_input_kernel(const float* __restrict__ pts, const float* __restrict__ dirs,
                                        const float* __restrict__ normals, long long n, int multires_view,
                                        float* __restrict__ rows, long long row_stride, int col_off, int width) {
  const int e_dim = 3 + 6 * multires_view;
  const long long total = n * width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / width;
    const int c = (int)(i - p * width);
    float v = 0.f;
    if (c < 3) v = pts[p * 3 + c];
    else if (c < 3 + e_dim) {
      const int e = c - 3;
      if (e < 3) v = dirs[p * 3 + e];
      else {
        const int f = (e - 3) / 6, w = (e - 3) % 6;
        const float a = dirs[p * 3 + (w % 3)] * exp2f((float)f);
        v = w < 3 ? sinf(a) : cosf(a);
      }
    } else if (c < 6 + e_dim) v = normals[p * 3 + (c - 3 - e_dim)];
    rows[p * row_stride + col_off + c] = v;
  }
}

extern "C" int vqn_neus_color_input(vqn_ctx* ctx, const float* pts, const float* dirs, const float* normals, int64_t n,
                                    int multires_view, float* rows, int64_t row_stride, int col_off, int width,
                                    vqn_stream stream) {
  VQN_CHECK_ARG(ctx && pts && dirs && normals && rows && n >= 0, "color_input: null argument");
  VQN_CHECK_ARG(multires_view >= 0 && width >= 9 + 6 * multires_view && col_off >= 0 &&
                    row_stride >= (int64_t)col_off + width, "color_input: columns do not fit the row");
  if (n == 0) return VQN_OK;
  long long want = (n * width + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
  neus_color_input_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(pts, dirs, normals, n, multires_view, rows, row_stride,
                                                               col_off, width);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// Light-visibility extraction, ray set-up and result scatter (gen_geo.py:182-257 compute_vis, :346-357 intersect_circle)
//   pair m = p * n_chunk + j  <->  surface point p, light l0 + j
//   d = normalize(lxyz[l] - surf[p]); front = d . normal[p] > 0; far = larger root of |surf + t d| = r;
//   near = min(0.1, far / 2)
// ---------------------------------------------------------------------------------------------
__global__ void neus_light_rays_kernel(const float* __restrict__ surf, const float* __restrict__ normal,
                                       const float* __restrict__ lxyz, long long n_pts, int l0, int n_chunk, float radius,
                                       float* __restrict__ rays_o, float* __restrict__ rays_d, float* __restrict__ near_o,
                                       float* __restrict__ far_o, float* __restrict__ front) {
  const long long total = n_pts * n_chunk;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    const long long p = m / n_chunk;
    const int l = l0 + (int)(m - p * n_chunk);
    const float sx = surf[p * 3], sy = surf[p * 3 + 1], sz = surf[p * 3 + 2];
    float dx = lxyz[l * 3] - sx, dy = lxyz[l * 3 + 1] - sy, dz = lxyz[l * 3 + 2] - sz;
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);          // torch.linalg.norm; division as in :208
    dx = dx / nrm; dy = dy / nrm; dz = dz / nrm;
    const float lcos = dx * normal[p * 3] + dy * normal[p * 3 + 1] + dz * normal[p * 3 + 2];
    const float b = 2.0f * (sx * dx + sy * dy + sz * dz);
    const float a = dx * dx + dy * dy + dz * dz;
    const float c = sx * sx + sy * sy + sz * sz - radius * radius;
    const float denom = 2.0f * a > 1e-7f ? 2.0f * a : 1e-7f;
    const float sq = sqrtf(b * b - 4.0f * a * c);
    const float t1 = (-b + sq) / denom, t2 = (-b - sq) / denom;
    const float t = t1 > t2 ? t1 : t2;
    const float n_far = t * 0.5f;
    rays_o[m * 3] = sx; rays_o[m * 3 + 1] = sy; rays_o[m * 3 + 2] = sz;
    rays_d[m * 3] = dx; rays_d[m * 3 + 1] = dy; rays_d[m * 3 + 2] = dz;
    far_o[m] = t;
    near_o[m] = 0.1f < n_far ? 0.1f : n_far;
    front[m] = lcos > 0.0f ? 1.0f : 0.0f;
  }
}

extern "C" int vqn_neus_light_rays(vqn_ctx* ctx, const float* surf, const float* normal, const float* lxyz,
                                   int64_t n_pts, int l0, int n_chunk, float radius, float* rays_o, float* rays_d,
                                   float* near_out, float* far_out, float* front, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && surf && normal && lxyz && rays_o && rays_d && near_out && far_out && front,
                "light_rays: null argument");
  VQN_CHECK_ARG(n_pts >= 0 && l0 >= 0 && n_chunk >= 1 && radius > 0.f, "light_rays: bad sizes");
  if (n_pts == 0) return VQN_OK;
  long long want = (n_pts * n_chunk + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
  neus_light_rays_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(surf, normal, lxyz, n_pts, l0, n_chunk, radius, rays_o,
                                                              rays_d, near_out, far_out, front);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// lvis[p, l0 + j] = 1 - weight_sum of the compacted ray i whose pair index is row_idx[i] = p * n_chunk + j
// (gen_geo.py:241-244: lvis_hit[front_lit_full] = 1 - occu); back-lit pairs keep the buffer's zero
__global__ void neus_lvis_scatter_kernel(const float* __restrict__ weight_sum, const int* __restrict__ row_idx,
                                         const int* __restrict__ n_dev, long long n_max, int l0, int n_chunk,
                                         int n_lights, float* __restrict__ lvis) {
  long long n = n_dev ? (long long)*n_dev : n_max;
  if (n > n_max) n = n_max;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long m = row_idx[i];
    const long long p = m / n_chunk;
    const int j = (int)(m - p * n_chunk);
    lvis[p * n_lights + l0 + j] = 1.0f - weight_sum[i];
  }
}

extern "C" int vqn_neus_lvis_scatter(vqn_ctx* ctx, const float* weight_sum, const int32_t* row_idx,
                                     const int32_t* n_dev, int64_t n_max, int l0, int n_chunk, int n_lights,
                                     float* lvis, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && weight_sum && row_idx && lvis && n_max >= 0, "lvis_scatter: null argument");
  VQN_CHECK_ARG(l0 >= 0 && n_chunk >= 1 && l0 + n_chunk <= n_lights, "lvis_scatter: light chunk outside the probe");
  if (n_max == 0) return VQN_OK;
  long long want = (n_max + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
  neus_lvis_scatter_kernel<<<blocks, 256, 0, vqn_cs(stream)>>>(weight_sum, row_idx, n_dev, n_max, l0, n_chunk, n_lights,
                                                                lvis);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
