// Context, error reporting and the small element-wise entry points of libvqnerf_b200.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <new>

#include <mutex>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

void vqn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int vqn_abi_version(void) { return VQN_ABI_VERSION; }

extern "C" const char* vqn_last_error(void) { return g_err; }

extern "C" const char* vqn_status_str(int s) {
  switch (s) {
    case VQN_OK: return "ok";
    case VQN_ERR_INVALID_ARG: return "invalid argument";
    case VQN_ERR_CUDA: return "CUDA error";
    case VQN_ERR_NONFINITE: return "non-finite value (check_numerics)";
    case VQN_ERR_UNSUPPORTED: return "unsupported";
    case VQN_ERR_ZERO_NORM: return "zero-norm direction";
    default: return "unknown status";
  }
}

namespace {
struct ScratchEntry { int kind; cudaStream_t stream; void* ptr; size_t bytes; };
// `retired`: blocks replaced by a larger one.  They are NOT freed before the context is destroyed: a captured CUDA graph
// may have their address baked into its kernel nodes, and a replay after a free would be a use-after-free.
struct ScratchPool { std::mutex mu; std::vector<ScratchEntry> entries; std::vector<void*> retired; };

// cudaMalloc is not allowed while the calling thread captures a stream (thread-local / global capture modes): the first
// use of a scratch kind on a stream must happen in the warm-up run on that SAME stream (torch.cuda.graph(stream=...)).
bool stream_is_capturing(cudaStream_t stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st != cudaStreamCaptureStatusNone;
}
}  // namespace

void* vqn_stream_scratch(vqn_ctx* ctx, int kind, cudaStream_t stream, size_t bytes) {
  ScratchPool* pool = static_cast<ScratchPool*>(ctx->pool);
  std::lock_guard<std::mutex> lock(pool->mu);
  ScratchEntry* hit = nullptr;
  for (ScratchEntry& e : pool->entries)
    if (e.kind == kind && e.stream == stream) { hit = &e; break; }
  if (hit && hit->bytes >= bytes) return hit->ptr;
  if (stream_is_capturing(stream)) {
    vqn_set_error("work buffer %d (%zu bytes) would have to be allocated during stream capture: run the call once on "
                  "this stream before capturing it", kind, bytes);
    return nullptr;
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { vqn_set_error("scratch allocation of %zu bytes failed", bytes); return nullptr; }
  if (hit) {
    pool->retired.push_back(hit->ptr);                   // grown: the old block stays alive (see ScratchPool)
    hit->ptr = p; hit->bytes = bytes;
  } else {
    pool->entries.push_back({kind, stream, p, bytes});
  }
  return p;
}

extern "C" int vqn_ctx_create(int device, vqn_ctx** out) {
  VQN_CHECK_ARG(out != nullptr, "out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    vqn_set_error("no CUDA device visible (%s): libvqnerf_b200 has no CPU fallback",
                  e == cudaSuccess ? "count == 0" : cudaGetErrorString(e));
    return VQN_ERR_CUDA;
  }
  VQN_CHECK_ARG(device >= 0 && device < count, "device index out of range");
  cudaDeviceProp prop;
  VQN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    vqn_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
    return VQN_ERR_UNSUPPORTED;
  }
  VQN_CUDA(cudaSetDevice(device));
  vqn_ctx* c = new (std::nothrow) vqn_ctx();
  VQN_CHECK_ARG(c != nullptr, "out of host memory");
  c->device = device;
  c->pool = new ScratchPool();
  c->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : VQN_SM_COUNT_FALLBACK;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->launches.store(0);
  VQN_CUDA(cudaMalloc(&c->nonfinite_flag, sizeof(int)));
  VQN_CUDA(cudaMemset(c->nonfinite_flag, 0, sizeof(int)));
  c->scratch_ints = VQN_SCRATCH_INTS;
  VQN_CUDA(cudaMalloc(&c->scratch, sizeof(int) * c->scratch_ints));
  *out = c;
  return VQN_OK;
}

extern "C" int vqn_ctx_destroy(vqn_ctx* ctx) {
  if (!ctx) return VQN_OK;
  cudaFree(ctx->nonfinite_flag);
  cudaFree(ctx->scratch);
  if (ctx->pool) {
    ScratchPool* pool = static_cast<ScratchPool*>(ctx->pool);
    for (ScratchEntry& e : pool->entries) cudaFree(e.ptr);
    for (void* r : pool->retired) cudaFree(r);
    delete pool;
  }
  delete ctx;
  return VQN_OK;
}

extern "C" int64_t vqn_vq_stats_size(int z_dim, int k) { return (int64_t)k + 2 + (int64_t)z_dim * k; }
extern "C" int64_t vqn_compact_workspace_size(int64_t n) { return n / 1024 + 2; }
extern "C" int64_t vqn_sample_pairs_workspace_size(int h, int w) { return (int64_t)(h - 2) * (w - 2); }

extern "C" int64_t vqn_ctx_launch_count(const vqn_ctx* ctx) { return ctx ? (int64_t)ctx->launches.load() : 0; }

extern "C" int vqn_ctx_check_numerics(vqn_ctx* ctx, vqn_stream stream) {
  VQN_CHECK_ARG(ctx, "ctx is NULL");
  int h = 0;
  VQN_CUDA(cudaMemcpyAsync(&h, ctx->nonfinite_flag, sizeof(int), cudaMemcpyDeviceToHost, vqn_cs(stream)));
  VQN_CUDA(cudaStreamSynchronize(vqn_cs(stream)));
  if (h != 0) {
    VQN_CUDA(cudaMemsetAsync(ctx->nonfinite_flag, 0, sizeof(int), vqn_cs(stream)));
    vqn_set_error("check_numerics: NaN/Inf in %s", (h & 1) ? "Z / MLP output" : "shaded radiance");
    return VQN_ERR_NONFINITE;
  }
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// gen_light_xyz: brdf/renderer.py:184-219 (+ sph2cart, xiuminglib/geometry/sph.py:184-190), float64
// ---------------------------------------------------------------------------------------------
extern "C" int vqn_gen_light_xyz(int h, int w, double radius, double* xyz, double* areas) {
  VQN_CHECK_ARG(h > 0 && w > 0 && xyz && areas, "bad light grid");
  const double pi = 3.14159265358979323846;
  double lat_step = pi / (h + 2), lng_step = 2 * pi / (w + 2);
  double sum = 0.0;
  for (int i = 0; i < h; ++i) {
    // np.linspace(a, b, n)[i] = a + i*(b-a)/(n-1)
    double lat0 = pi / 2 - lat_step, lat1 = -pi / 2 + lat_step;
    double lat = h > 1 ? lat0 + i * ((lat1 - lat0) / (h - 1)) : lat0;
    if (h > 1 && i == h - 1) lat = lat1;
    for (int j = 0; j < w; ++j) {
      double lng0 = pi - lng_step, lng1 = -pi + lng_step;
      double lng = w > 1 ? lng0 + j * ((lng1 - lng0) / (w - 1)) : lng0;
      if (w > 1 && j == w - 1) lng = lng1;
      double* p = xyz + ((size_t)i * w + j) * 3;
      p[0] = radius * cos(lat) * cos(lng);
      p[1] = radius * cos(lat) * sin(lng);
      p[2] = radius * sin(lat);
      double sc = sin(pi / 2 - lat);
      areas[(size_t)i * w + j] = sc;
      sum += sc;
    }
  }
  for (int i = 0; i < h * w; ++i) areas[i] = 4 * pi * areas[i] / sum;
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// small element-wise kernels
// ---------------------------------------------------------------------------------------------
__global__ void embed_kernel(const float* __restrict__ x, long long n, int n_freqs, float* __restrict__ out, long long ld) {
  // Embedder.__call__: one thread per (row, output column); precise sinf/cosf on x * 2^k
  int d = 3 + 6 * n_freqs;
  long long total = n * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i / d;
    int c = (int)(i - r * d);
    float v;
    if (c < 3) {
      v = x[r * 3 + c];
    } else {
      int q = c - 3, f = q / 6, w = q % 6;
      float a = x[r * 3 + (w % 3)] * exp2f((float)f);
      v = w < 3 ? sinf(a) : cosf(a);
    }
    out[r * ld + c] = v;
  }
}

extern "C" int vqn_embed_ld(vqn_ctx* ctx, const float* x, int64_t n, int n_freqs, float* out, int64_t ld_out, vqn_stream s);
extern "C" int vqn_embed(vqn_ctx* ctx, const float* x, int64_t n, int n_freqs, float* out, vqn_stream s) {
  return vqn_embed_ld(ctx, x, n, n_freqs, out, 3 + 6 * (int64_t)n_freqs, s);
}
extern "C" int vqn_embed_ld(vqn_ctx* ctx, const float* x, int64_t n, int n_freqs, float* out, int64_t ld_out, vqn_stream s) {
  VQN_CHECK_ARG(ctx && x && out && n >= 0 && n_freqs >= 0 && n_freqs <= 16, "embed args");
  VQN_CHECK_ARG(ld_out >= 3 + 6 * n_freqs, "embed: ld_out < 3 + 6 n_freqs");
  if (n == 0) return VQN_OK;
  long long total = n * (3 + 6 * n_freqs);
  int blocks = (int)((total + 255) / 256 < (long long)ctx->sm_count * 16 ? (total + 255) / 256
                                                                          : (long long)ctx->sm_count * 16);
  embed_kernel<<<blocks, 256, 0, vqn_cs(s)>>>(x, n, n_freqs, out, (long long)ld_out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// rows of x[n,d] -> x * rsqrt(max(sum x^2, 1e-6)); one warp per row
__global__ void l2norm_rows_kernel(const float* __restrict__ x, long long n, int d, float* __restrict__ out) {
  int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n; r += nwarps) {
    const float* p = x + r * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) { float v = p[c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    float inv = rsqrtf(fmaxf(s, 1e-6f));
    for (int c = lane; c < d; c += 32) out[r * d + c] = p[c] * inv;
  }
}

extern "C" int vqn_l2_normalize_rows(vqn_ctx* ctx, const float* x, int64_t n, int d, float* out, vqn_stream s) {
  VQN_CHECK_ARG(ctx && x && out && n >= 0 && d > 0, "l2_normalize_rows args");
  if (n == 0) return VQN_OK;
  long long want = (n + 7) / 8;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  l2norm_rows_kernel<<<blocks, 256, 0, vqn_cs(s)>>>(x, n, d, out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// get_codebook: clip(raw,0,1) then l2_normalize(axis=0) of [Z,K]; one block per column
__global__ void get_codebook_kernel(const float* __restrict__ raw, int z_dim, int k, float* __restrict__ out) {
  int col = blockIdx.x;
  __shared__ float red[32];
  float s = 0.f;
  for (int z = threadIdx.x; z < z_dim; z += blockDim.x) {
    float v = fminf(fmaxf(raw[(size_t)z * k + col], 0.f), 1.f);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = rsqrtf(fmaxf(t, 1e-6f));
  }
  __syncthreads();
  float inv = red[0];
  for (int z = threadIdx.x; z < z_dim; z += blockDim.x) {
    float v = fminf(fmaxf(raw[(size_t)z * k + col], 0.f), 1.f);
    out[(size_t)z * k + col] = v * inv;
  }
}

extern "C" int vqn_get_codebook(vqn_ctx* ctx, const float* raw, int z_dim, int k, float* out, vqn_stream s) {
  VQN_CHECK_ARG(ctx && raw && out && z_dim > 0 && k > 0, "get_codebook args");
  get_codebook_kernel<<<k, 128, 0, vqn_cs(s)>>>(raw, z_dim, k, out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

template <bool TO_SRGB>
__global__ void srgb_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = TO_SRGB ? vqn_linear2srgb(x[i]) : vqn_srgb2linear(x[i]);
}

static int launch_srgb(vqn_ctx* ctx, const float* x, int64_t n, float* out, vqn_stream s, bool to_srgb) {
  VQN_CHECK_ARG(ctx && x && out && n >= 0, "srgb args");
  if (n == 0) return VQN_OK;
  long long want = (n + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
  if (to_srgb) srgb_kernel<true><<<blocks, 256, 0, vqn_cs(s)>>>(x, n, out);
  else srgb_kernel<false><<<blocks, 256, 0, vqn_cs(s)>>>(x, n, out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
extern "C" int vqn_linear2srgb(vqn_ctx* ctx, const float* x, int64_t n, float* out, vqn_stream s) {
  return launch_srgb(ctx, x, n, out, s, true);
}
extern "C" int vqn_srgb2linear(vqn_ctx* ctx, const float* x, int64_t n, float* out, vqn_stream s) {
  return launch_srgb(ctx, x, n, out, s, false);
}

// ---------------------------------------------------------------------------------------------
// mask compaction: ind = where(alpha[:,0] > 0), order-preserving, single block-scan pass with a
// decoupled look-back replaced by a simple two-kernel scheme (counts per block, then scatter).
// ---------------------------------------------------------------------------------------------
#define CMP_BLOCK 1024
__global__ void compact_count_kernel(const float* __restrict__ alpha, long long n, int* __restrict__ block_counts) {
  long long i = blockIdx.x * (long long)CMP_BLOCK + threadIdx.x;
  int flag = (i < n && alpha[i] > 0.f) ? 1 : 0;
  int c = __syncthreads_count(flag);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}
// exclusive scan of block counts by one block (n_blocks <= a few thousand)
__global__ void compact_scan_kernel(int* __restrict__ block_counts, int n_blocks, int* __restrict__ n_active) {
  __shared__ int carry;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < n_blocks ? block_counts[i] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
      wsum[lane] = t;
    }
    __syncthreads();
    int excl = x - v + (w > 0 ? wsum[w - 1] : 0) + carry;
    if (i < n_blocks) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_active = carry;
}
__global__ void compact_scatter_kernel(const float* __restrict__ alpha, long long n,
                                       const int* __restrict__ block_offsets, int* __restrict__ row_idx) {
  __shared__ int wsum[32];
  long long i = blockIdx.x * (long long)CMP_BLOCK + threadIdx.x;
  int flag = (i < n && alpha[i] > 0.f) ? 1 : 0;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned bal = __ballot_sync(0xffffffffu, flag);
  int pre = __popc(bal & ((1u << lane) - 1));
  if (lane == 0) wsum[w] = __popc(bal);
  __syncthreads();
  if (w == 0) {
    int t = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
    wsum[lane] = t;
  }
  __syncthreads();
  int off = block_offsets[blockIdx.x] + (w > 0 ? wsum[w - 1] : 0) + pre;
  if (flag) row_idx[off] = (int)i;
}

extern "C" int vqn_compact_mask(vqn_ctx* ctx, const float* alpha, int64_t n, int32_t* row_idx,
                                int32_t* n_active, int32_t* workspace, vqn_stream s) {
  VQN_CHECK_ARG(ctx && n_active && n >= 0 && n < (1LL << 31), "compact_mask args");
  int n_blocks = (int)((n + CMP_BLOCK - 1) / CMP_BLOCK);
  if (n_blocks == 0) { VQN_CUDA(cudaMemsetAsync(n_active, 0, sizeof(int), vqn_cs(s))); return VQN_OK; }
  VQN_CHECK_ARG(alpha && row_idx, "compact_mask: null alpha / row_idx");
  VQN_CHECK_ARG(workspace || (size_t)n_blocks + 16 <= ctx->scratch_ints, "compact_mask: more than 64 M rows");
  int* counts = workspace ? workspace : ctx->scratch + 16;   // ctx scratch: one compaction in flight per ctx
  compact_count_kernel<<<n_blocks, CMP_BLOCK, 0, vqn_cs(s)>>>(alpha, n, counts);
  VQN_LAUNCHED(ctx);
  compact_scan_kernel<<<1, 1024, 0, vqn_cs(s)>>>(counts, n_blocks, n_active);
  VQN_LAUNCHED(ctx);
  compact_scatter_kernel<<<n_blocks, CMP_BLOCK, 0, vqn_cs(s)>>>(alpha, n, counts, row_idx);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

__global__ void scatter_rows_kernel(const float* __restrict__ compact, const int* __restrict__ row_idx,
                                    const int* __restrict__ n_dev, long long n_max, int c,
                                    float* __restrict__ out) {
  long long n = n_dev ? (long long)*n_dev : n_max;
  if (n > n_max) n = n_max;
  long long total = n * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i / c;
    int k = (int)(i - r * c);
    out[(long long)row_idx[r] * c + k] = compact[i];
  }
}

extern "C" int vqn_scatter_rows(vqn_ctx* ctx, const float* compact, const int32_t* row_idx,
                                const int32_t* n_dev, int64_t n_max, int c, float* out, vqn_stream s) {
  VQN_CHECK_ARG(ctx && n_max >= 0 && c > 0, "scatter_rows args");
  if (n_max == 0) return VQN_OK;
  VQN_CHECK_ARG(compact && row_idx && out, "scatter_rows: null pointer");
  long long want = (n_max * c + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
  scatter_rows_kernel<<<blocks, 256, 0, vqn_cs(s)>>>(compact, row_idx, n_dev, n_max, c, out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// up to 4 scatter_nd's that share the row list in ONE launch (fast_render scatters four material maps per view)
struct ScatterMulti { const float* src[4]; float* out[4]; int c[4]; int count; };
__global__ void scatter_rows_multi_kernel(ScatterMulti m, const int* __restrict__ row_idx, const int* __restrict__ n_dev,
                                          long long n_max) {
  long long n = n_dev ? (long long)*n_dev : n_max;
  if (n > n_max) n = n_max;
  int ctot = 0;
  for (int q = 0; q < m.count; ++q) ctot += m.c[q];
  const long long total = n * ctot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ctot;
    int k = (int)(i - r * ctot), q = 0;
    while (k >= m.c[q]) { k -= m.c[q]; ++q; }
    m.out[q][(long long)row_idx[r] * m.c[q] + k] = m.src[q][r * m.c[q] + k];
  }
}
extern "C" int vqn_scatter_rows_multi(vqn_ctx* ctx, const float* const* compact, const int32_t* widths, int count,
                                      const int32_t* row_idx, const int32_t* n_dev, int64_t n_max, float* const* outs,
                                      vqn_stream s) {
  VQN_CHECK_ARG(ctx && compact && widths && outs && count >= 1 && count <= 4 && n_max >= 0, "scatter_rows_multi args");
  if (n_max == 0) return VQN_OK;
  VQN_CHECK_ARG(row_idx, "scatter_rows_multi: null row_idx");
  ScatterMulti m = {};
  m.count = count;
  int ctot = 0;
  for (int q = 0; q < count; ++q) {
    VQN_CHECK_ARG(compact[q] && outs[q] && widths[q] > 0, "scatter_rows_multi: bad entry");
    m.src[q] = compact[q]; m.out[q] = outs[q]; m.c[q] = widths[q]; ctot += widths[q];
  }
  long long want = (n_max * ctot + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
  scatter_rows_multi_kernel<<<blocks, 256, 0, vqn_cs(s)>>>(m, row_idx, n_dev, n_max);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// ---------------------------------------------------------------------------------------------
// material combine: spec = ks * basecolor; albedo = (1 - ks) * basecolor (vq_nfr.py:330-331,590-591)
// and the optional opt_scale of fast_render (:333-336); compact [n,*] in and out.
// ---------------------------------------------------------------------------------------------
__global__ void material_combine_kernel(const float* __restrict__ base, const float* __restrict__ ks,
                                        const float* __restrict__ opt_scale, const int* __restrict__ n_dev,
                                        long long n_max, float* __restrict__ albedo, float* __restrict__ spec,
                                        float* __restrict__ albedo_s, float* __restrict__ spec_s) {
  long long n = n_dev ? (long long)*n_dev : n_max;
  if (n > n_max) n = n_max;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * 3;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i / 3;
    int c = (int)(i - r * 3);
    float b = base[i], k = ks[r];
    float sp = k * b, al = (1.0f - k) * b;
    if (albedo) albedo[i] = al;
    if (spec) spec[i] = sp;
    float sc = opt_scale ? opt_scale[c] : 1.0f;
    if (albedo_s) albedo_s[i] = opt_scale ? al * sc : al;
    if (spec_s) spec_s[i] = opt_scale ? sp * sc : sp;
  }
}

extern "C" int vqn_material_combine(vqn_ctx* ctx, const float* basecolor, const float* ks, const float* opt_scale,
                                    const int32_t* n_dev, int64_t n, float* albedo, float* spec, float* albedo_scaled,
                                    float* spec_scaled, vqn_stream s) {
  VQN_CHECK_ARG(ctx && basecolor && ks && n >= 0, "material_combine args");
  if (n == 0) return VQN_OK;
  long long want = (n * 3 + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  material_combine_kernel<<<blocks, 256, 0, vqn_cs(s)>>>(basecolor, ks, opt_scale, n_dev, n, albedo, spec,
                                                         albedo_scaled, spec_scaled);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

// fast_render's material edit (models/vq_nfr.py:258-260, 293-295, 324-330): rows whose edit_mask[..., 0] > 0 get
// src * (1 - m) + m * update with m = 1, i.e. the update; the scaled copies the shading kernel reads follow (:333-336)
__global__ void material_edit_kernel(const float* __restrict__ edit_mask, int mask_stride, const int* __restrict__ row_idx,
                                     const int* __restrict__ n_dev, long long n_max, float3 diff, int has_diff, float3 spc,
                                     int has_spec, float rgh, int has_rough, const float* __restrict__ opt_scale,
                                     float* __restrict__ albedo, float* __restrict__ spec, float* __restrict__ rough,
                                     float* __restrict__ albedo_s, float* __restrict__ spec_s) {
  long long n = n_dev ? (long long)*n_dev : n_max;
  if (n > n_max) n = n_max;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = row_idx ? (long long)row_idx[i] : i;
    const float m = edit_mask[row * mask_stride] > 0.f ? 1.f : 0.f;
    const float sc[3] = {opt_scale ? opt_scale[0] : 1.f, opt_scale ? opt_scale[1] : 1.f, opt_scale ? opt_scale[2] : 1.f};
    const float dv[3] = {diff.x, diff.y, diff.z}, sv[3] = {spc.x, spc.y, spc.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (has_diff) {
        const float v = albedo[i * 3 + c] * (1.f - m) + m * dv[c];
        albedo[i * 3 + c] = v;
        if (albedo_s) albedo_s[i * 3 + c] = opt_scale ? v * sc[c] : v;
      }
      if (has_spec) {
        const float v = spec[i * 3 + c] * (1.f - m) + m * sv[c];
        spec[i * 3 + c] = v;
        if (spec_s) spec_s[i * 3 + c] = opt_scale ? v * sc[c] : v;
      }
    }
    if (has_rough) rough[i] = rough[i] * (1.f - m) + m * rgh;
  }
}

extern "C" int vqn_material_edit(vqn_ctx* ctx, const float* edit_mask, int mask_stride, const int32_t* row_idx,
                                 const int32_t* n_dev, int64_t n, const float* diff3, const float* spec3,
                                 const float* rough1, const float* opt_scale, float* albedo, float* spec, float* rough,
                                 float* albedo_scaled, float* spec_scaled, vqn_stream s) {
  VQN_CHECK_ARG(ctx && edit_mask && mask_stride >= 1 && albedo && spec && rough && n >= 0, "material_edit args");
  if (n == 0 || (!diff3 && !spec3 && !rough1)) return VQN_OK;
  const float3 d = diff3 ? make_float3(diff3[0], diff3[1], diff3[2]) : make_float3(0.f, 0.f, 0.f);
  const float3 sp = spec3 ? make_float3(spec3[0], spec3[1], spec3[2]) : make_float3(0.f, 0.f, 0.f);
  long long want = (n + 255) / 256;
  int blocks = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  material_edit_kernel<<<blocks, 256, 0, vqn_cs(s)>>>(edit_mask, mask_stride, row_idx, n_dev, n, d, diff3 != nullptr, sp,
                                                      spec3 != nullptr, rough1 ? rough1[0] : 0.f, rough1 != nullptr,
                                                      opt_scale, albedo, spec, rough, albedo_scaled, spec_scaled);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
