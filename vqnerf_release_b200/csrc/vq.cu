// VQ layer C ABI: nearest-codeword assignment (kernel in vq_mma.cu) + Sonnet EMA update
// (networks/vq_layers.py:257-344).
//
// Data layout in HBM: latents x[n,256] fp32 row-major (1 KB rows), codebook C[256,K] fp32 (column =
// codeword), indices int64[n].  Algorithmic bytes per latent: 4*256 in + 8 out = 1032 B.
#include "vq.cuh"

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
  // codebooks of K > 32, indices only (BASELINE configs[2] sweep): GEMM-shaped -> tcgen05 kernel with an arg-min epilogue
  if (k >= vq_tc_min_k() && indices && !sel_mask && !normalize_inputs && !quantize && !distances && !z_norm_out && !stats)
    return vq_tc_assign_launch(ctx, p, s);
  return vq_assign_mma_launch(ctx, p, s);
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
