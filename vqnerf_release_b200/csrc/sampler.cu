// Training-batch assembler: outer_sample of nerfactor/train_nfr.py:380-467 on the device.
//
// The reference reshapes the 10 per-pixel tensors of one view to [H,W,.], draws ONE random 8-neighbour per interior
// pixel (tf.random.uniform, :414-415), keeps the (pixel, neighbour) pairs whose alphas both exceed alpha_thres
// (:428-431, boolean_mask = order-preserving compaction), draws n_rays_per_step of them with replacement
// (:438-439) and gathers [p1, p1_n, p2, p2_n, ...] rows of every tensor (:443-459) -- ~40 eager ops, a host sync
// (hw[0,:].numpy()) and two data-dependent shapes.  Here: one kernel flags the valid interior pixels and records
// their neighbour, the library's order-preserving compaction (abi.cu) lists them, one kernel draws the pairs, one
// gather kernel per tensor.  TensorFlow's RNG streams are not reproducible outside TF, so randomness comes from a
// counter-based hash (splitmix64 of (seed, stream, index)); the oracle restates the same generator.
#include "common.cuh"

namespace {

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned rnd_u32(unsigned long long seed, unsigned stream, unsigned long long idx) {
  return (unsigned)(splitmix64(splitmix64(seed ^ ((unsigned long long)stream << 56)) + idx) >> 32);
}

// interior pixel q = (i-1)*(W-2) + (j-1), i in [1,H-2], j in [1,W-2]: flag[q] = 1 when the pair is usable,
// nb[q] = linear index (i', j') of its randomly drawn neighbour
__global__ void pair_flag_kernel(const float* __restrict__ alpha, int H, int W, float thres, int use_thres,
                                 unsigned long long seed, float* __restrict__ flag, int* __restrict__ nb) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(H - 2) * (W - 2);
  if (q >= total) return;
  const int i = 1 + (int)(q / (W - 2)), j = 1 + (int)(q % (W - 2));
  // jitters (:401-402): [-1,-1],[-1,0],[-1,1],[0,-1],[0,1],[1,-1],[1,0],[1,1]
  const int jit = (int)(rnd_u32(seed, 1, (unsigned long long)q) % 8u);
  const int k = jit < 4 ? jit : jit + 1;               // skip the centre of the 3x3 neighbourhood
  const int di = k / 3 - 1, dj = k % 3 - 1;
  const long long n_lin = (long long)(i + di) * W + (j + dj);
  nb[q] = (int)n_lin;
  bool ok = true;
  if (use_thres) ok = alpha[(long long)i * W + j] > thres && alpha[n_lin] > thres;
  flag[q] = ok ? 1.0f : 0.0f;
}

// rows[2s] = pixel, rows[2s+1] = neighbour of a uniformly drawn valid pair (with replacement, :438-448)
__global__ void pair_select_kernel(const int* __restrict__ valid_q, const int* __restrict__ n_valid,
                                   const int* __restrict__ nb, int W, int bs, unsigned long long seed,
                                   int* __restrict__ rows) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= bs) return;
  const int cnt = *n_valid;
  if (cnt <= 0) { rows[2 * s] = -1; rows[2 * s + 1] = -1; return; }
  const int r = (int)(rnd_u32(seed, 2, (unsigned long long)s) % (unsigned)cnt);
  const int q = valid_q[r];
  const int i = 1 + q / (W - 2), j = 1 + q % (W - 2);
  rows[2 * s] = i * W + j;
  rows[2 * s + 1] = nb[q];
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ rows, long long n_out, int c,
                                   float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_out * c) return;
  const long long o = idx / c;
  const int k = (int)(idx % c);
  const int r = rows[o];
  out[idx] = r >= 0 ? src[(long long)r * c + k] : 0.0f;
}

}  // namespace

/* outer_sample index half (train_nfr.py:401-448): alpha [H*W] of one view -> rows int32 [2*bs] =
 * [p1, p1_n, p2, p2_n, ...] (linear pixel indices; -1 when the view has no valid pair) and n_valid[1].
 * Workspace: flag float [(H-2)(W-2)], nb / valid_q int32 [(H-2)(W-2)], compaction workspace as vqn_compact_mask. */
extern "C" int vqn_sample_pairs(vqn_ctx* ctx, const float* alpha, int h, int w, int use_alpha_thres, float alpha_thres,
                                int bs, uint64_t seed, float* flag_ws, int32_t* nb_ws, int32_t* valid_ws,
                                int32_t* compact_ws, int32_t* n_valid, int32_t* rows, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && alpha && flag_ws && nb_ws && valid_ws && n_valid && rows, "sample_pairs: null pointer");
  VQN_CHECK_ARG(h >= 3 && w >= 3 && bs >= 1, "sample_pairs: need H, W >= 3 and bs >= 1");
  const long long total = (long long)(h - 2) * (w - 2);
  cudaStream_t s = vqn_cs(stream);
  pair_flag_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(alpha, h, w, alpha_thres, use_alpha_thres,
                                                                  (unsigned long long)seed, flag_ws, nb_ws);
  VQN_LAUNCHED(ctx);
  int rc = vqn_compact_mask(ctx, flag_ws, total, valid_ws, n_valid, compact_ws, stream);
  if (rc != VQN_OK) return rc;
  pair_select_kernel<<<(bs + 255) / 256, 256, 0, s>>>(valid_ws, n_valid, nb_ws, w, bs, (unsigned long long)seed, rows);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

/* tf.gather_nd(tensor, select_ind) (:450-465): out[n_out, c] = src[rows[o], :] (zeros for rows[o] < 0) */
extern "C" int vqn_gather_rows(vqn_ctx* ctx, const float* src, const int32_t* rows, int64_t n_out, int c, float* out,
                               vqn_stream stream) {
  VQN_CHECK_ARG(ctx && src && rows && out && n_out >= 0 && c >= 1, "gather_rows args");
  if (n_out == 0) return VQN_OK;
  const long long total = (long long)n_out * c;
  gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, vqn_cs(stream)>>>(src, rows, (long long)n_out, c, out);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}
