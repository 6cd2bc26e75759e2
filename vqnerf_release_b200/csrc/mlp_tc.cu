// Fused small-MLP kernel on the 5th-gen tensor cores (tcgen05 + TMEM):  VQN_PREC_TF32X3 and VQN_PREC_BF16.
//
// Reference: networks/mlp.py:24-50, networks/embedder.py:23-47, models/vq_nfr.py:771-828.
//
// One persistent CTA per SM walks tiles of 128 surface points through a chain of Dense layers without the
// activations ever leaving the SM:
//
//   * accumulators live in TMEM (128 lanes = 128 points, up to 256 fp32 columns = layer width); two 256-column
//     regions ping-pong between consecutive layers (512 columns = the whole TMEM of the SM);
//   * the A operand of layer L+1 is produced CHUNK BY CHUNK (128 bytes of K per row: 32 tf32 / 64 bf16 values) by
//     the four epilogue warps, which drain layer L's accumulator with tcgen05.ld, add the bias, apply the
//     activation, convert (tf32 hi + lo planes, or bf16) and store the chunk in the 128-B-swizzled K-major layout
//     the UMMA descriptors expect; positional-encoding chunks (Embedder) and latent chunks (z rows from global)
//     enter the same ring, which is how the skip connections concat(y, x) are realised without copies;
//   * the B operand (weights) is streamed from L2 by one producer thread with 1-D bulk copies (TMA engine,
//     cp.async.bulk + mbarrier complete_tx) of host-pre-swizzled chunk images, through its own ring;
//   * one thread issues tcgen05.mma (M = 128, N = layer width, K = 8 tf32 / 16 bf16 per instruction) and
//     tcgen05.commit's the ring slots back to their producers and the accumulator to the epilogue warps.
//
// VQN_PREC_TF32X3 evaluates every product as a_hi w_hi + a_lo w_hi + a_hi w_lo with hi = rna_tf32(x), lo = x - hi,
// fp32 accumulation in TMEM: the leading term is a kind::tf32 MMA; the two correction terms are 2^-11 of it, so
// their operands only need ~9 bits and are fed as bf16 through ONE kind::f16 MMA of twice the K
// ([a_lo | a_hi] . [w_hi ; w_lo]): ~2^-20 relative per product, i.e. fp32-level parity (1e-4 budget of the north
// star), for 2/3 of the tensor time and 2/3 of the shared-memory operand traffic of three tf32 MMAs -- and the
// kernel is bound by exactly that traffic (the 128 B/clk smem port feeds UMMA operand reads, the A-chunk stores and
// the weight copies).  VQN_PREC_BF16 is a single kind::f16 MMA on bf16 operands (1e-2 budget).
#include <string.h>

#include "net.cuh"
#include "tc_common.cuh"

#define TC_M 128               // points per tile (= TMEM lanes)
// clock64 stamps for benchmarks/tc_trace.py: compiled in only with -DVQN_TC_TRACE (VQN_EXTRA_NVCC_FLAGS=-DVQN_TC_TRACE
// python -m vqnerf_release_b200.build --force); in the product build the hot loops carry no predicated CS2R / stores
#ifdef VQN_TC_TRACE
#define TC_STAMP(p, i) do { if (p) (p)[i] = clock64(); } while (0)
#else
#define TC_STAMP(p, i) do { } while (0)
#endif
#define TC_MAX_LAYERS 20
#define TC_NPAD_MAX 256

enum { SRC_DRAIN = 0, SRC_EMBED = 1, SRC_GLOBAL = 2,
       SRC_GRADINIT = 3,      // reverse mode: A = w_tail (x) act'(stash)            (d out / d pre of the last hidden layer)
       SRC_GRAD = 4,          // reverse mode: A = drain(previous accumulator) * act'(stash)
       SRC_BWD = 5,           // training backward (MODE 4): A = dz = drain(previous accumulator) * act'(saved forward output), dz also stored
       SRC_TAILGRAD = 6 };    // training backward: A = dz of the layer below a NARROW last layer, formed on the CUDA cores

struct TcLayer {
  const uint8_t* w;       // packed chunk images, chunk c at w + c * chunk_bytes
  const float* bias;      // [Npad]
  int N, Npad, act;
  int nseg, seg_type[2], seg_chunks[2], seg_first_chunk[2];   // seg_first_chunk: index inside the source
  int seg_ts[2];          // 1: the chunks of this segment are handed over in TENSOR MEMORY (TS-mode MMAs), see TC_TS_COL
  int out_slot;           // >= 0: accumulator is written to global outs[out_slot] (row-major [point][N])
  int bias_off;           // offset of this layer's bias inside the shared-memory bias table
  float post_scale, post_bias;
  int tmem_col;           // first TMEM column of this layer's accumulator
  // Folded skip connection of a narrow last layer: concat(y, x) . W = y . W_y + x . W_x with <= 4 outputs.  The x half
  // is accumulated on the CUDA cores by the producers of the network's FIRST layer while they hold the same x
  // values in registers (skip_w = W_x rows, skip_n outputs; 4 FMAs per value), the y half by the tail below -- the
  // narrow layer needs no MMA, no A chunks and no second pass over x.
  const float* skip_w;    // fp32 rows of the tail kernel that multiply x: [g_dim, skip_n]
  int skip_n, skip_woff;
  // Narrow tail: a last layer with <= 4 outputs that follows this layer (after the fold above its K is just this
  // layer's width) is evaluated on the CUDA cores while this layer's accumulator is drained -- 3 x width FMAs per
  // row instead of four more A chunks (32 KB of smem stores + fences each) for an MMA with N = 16.
  const float* tail_w;    // fp32 Keras kernel of the tail layer, rows [0, N) used, row stride tail_n
  const float* tail_b;    // its bias
  int tail_n, tail_act, tail_out_slot, tail_add_skip, tail_woff;
  float tail_scale, tail_bias;
  // Reverse-mode gradient (MODE 2, vqn_sdf_forward): stash_w >= 0: whoever applies this layer's activation (the next
  // layer's drain, or the tail) also stores act'(pre-activation) into stash slot stash_w of the CTA's scratch;
  // grad_stash: the slot a SRC_GRADINIT / SRC_GRAD layer multiplies by; grad_tail_woff: table of the tail weights;
  // egrad_col0 >= 0: columns [egrad_col0, egrad_col0 + 3 + 6 n_freqs) of this layer's accumulator are d out / d embedding
  // and are contracted with the embedding Jacobian (the skip half of the skip_in layer, and the whole last layer);
  // grad_final: this is the last backward layer -- combine the partial gradients and store d out / d x.
  int stash_w, grad_stash, grad_tail_woff, egrad_col0, grad_final;
  // Training forward (MODE 3, vqn_net_forward_train): whoever applies this layer's activation also stores the activated
  // values into save[row * save_ld + col] -- the layer outputs the backward pass needs (mlp.py under GradientTape)
  float* save;
  int save_ld;
  // Training backward (MODE 4, vqn_net_backward_train): this layer is the GEMM  d_in = dz_i . W_i^T  of forward layer i, its A
  // chunks are dz_i:  SRC_BWD: drain(previous accumulator) * act'(bw_y) (bw_y = the forward layer's saved output, bw_act its
  // activation); SRC_TAILGRAD: (sum_o tg_dz[row][o] W_last[c][o]) * act'(bw_y) -- the narrow last forward layer evaluated
  // backwards on the CUDA cores (tg_w = its Keras kernel [rows, tg_n], tg_n <= 3).  Either way dz_i is also STORED to bw_dz
  // (the weight gradients read it).  Final drain of the chain (out_slot >= 0): fin_mode 0 store / 1 add / 2 atomic add;
  // fin_y: multiply by act'(fin_y) first (the chain ends in a dz, not in an input gradient); fin_skip_n > 0: add the x half of
  // the folded skip concat, sum_o tg_dz[row][o] W_last[K_y + c][o].
  const float* bw_y; float* bw_dz; const float* tg_dz; const float* tg_w; const float* fin_y;
  int bw_y_ld, bw_act, bw_dz_ld, bw_n, tg_dz_ld, tg_n, tg_woff, tg_rows, fin_mode, fin_y_ld, fin_act, fin_skip_n, fin_skip_woff,
      fin_skip_row0;
};

struct TcProgram {
  int n_layers;
  int n_freqs;
  int g_dim;               // row length of the GLOBAL source (z_dim)
  const float* pts;        // [*,3] (EMBED source)
  const float* gsrc;       // [n,g_dim] (GLOBAL source)
  const int* row_idx;
  const int* n_dev;
  long long n;
  float* outs[4];
  int out_stride[4];
  // Per-CTA latent scratch (vqn_tc_mlp_main): the GLOBAL source / output slot `local_slot` is a [128, g_dim] tile PER CTA
  // (gsrc + blockIdx.x * 128 * g_dim), written by the final drain of the bottleneck and re-read by the heads of the
  // SAME tile -- 148 x 128 KB stay in L2, so the latent never travels to HBM and encoder + heads are one launch.
  // out_dup: optional [n, g_dim] global copy of that slot (callers that want z).
  int g_local, local_slot;
  float* out_dup;
  // Jet mode (vqn_sdf_forward with grad_out): a tile is 32 points x 4 rows -- row 4p is the value of point p, rows
  // 4p+1..4p+3 its derivatives with respect to x, y, z; every layer is linear in the tangent rows and the activation
  // multiplies them by act'(pre-activation of the value row), fetched from the neighbouring TMEM lane by a shuffle.
  int jet;
  float* jet_grad;         // [n,3]: tangent rows of the narrow tail (d sdf / d x); reverse mode: the gradient output
  // Reverse mode (MODE 2): forward pass on value rows (128 points per tile) with act' of every hidden layer stashed in a
  // per-CTA scratch ([slot][16-column piece][row][16] fp32, L2-resident), then the transposed layers back to the
  // embedding and a contraction with the embedding's Jacobian: 2 row-passes per point instead of the 4 of the jets.
  int reverse;
  float* stash;            // sm_count x TC_STASH_FLOATS
  int train;               // MODE 3
  int train_bwd;           // MODE 4
  int* nonfinite;
  long long* trace;        // diagnostic (vqn_debug_tc_trace): clock64 stamps of CTA 0's MMA thread, 4 per layer
  TcLayer layers[TC_MAX_LAYERS];
};

struct TcPack {
  int n_layers;
  int Npad[VQN_MAX_LAYERS];
  int n_chunks[VQN_MAX_LAYERS];
  uint8_t* w[VQN_MAX_LAYERS];
  float* bias[VQN_MAX_LAYERS];
  // transposed images for the reverse-mode gradient (built on first use): layer i as the GEMM  d in = d out . W_i^T,
  // i.e. N = inputs of layer i (incl. the skip concat), K = outputs of layer i; biasT = zeros
  bool has_T;
  int NpadT[VQN_MAX_LAYERS];
  int n_chunksT[VQN_MAX_LAYERS];
  uint8_t* wT[VQN_MAX_LAYERS];
  float* biasT[VQN_MAX_LAYERS];
};

// ---------------------------------------------------------------------------------------------
// weight packing: Keras kernel [in,out] -> per K-chunk swizzled images [Npad x 128 B] (tf32: hi image then lo image)
// K order = Keras row order (after a skip: y rows, then the x rows), each segment padded to whole chunks.
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ void tc_pack_body(const float* __restrict__ w, const float* __restrict__ b, int n_out, int Npad,
                                             int seg0_rows, int seg0_chunks, int seg1_rows, int seg1_chunks,
                                             uint8_t* __restrict__ pw, float* __restrict__ pb, int transpose_ld,
                                             const int block, const int n_blocks) {
  // transpose_ld > 0: the GEMM's N index runs over the Keras kernel's ROWS (n_out of them) and its K index over the
  // kernel's columns (seg0_rows of them, row length transpose_ld): the image of W^T for the backward pass
  constexpr int E = BF16 ? 64 : 32;
  const int n_chunks = seg0_chunks + seg1_chunks;
  const size_t plane = (size_t)Npad * 128;
  const size_t chunk_bytes = BF16 ? plane : 2 * plane;
  const long long total = (long long)n_chunks * Npad * E;
  for (long long i = block * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)n_blocks * blockDim.x) {
    int kk = (int)(i % E);
    int n = (int)((i / E) % Npad);
    int c = (int)(i / ((long long)E * Npad));
    int src;
    if (c < seg0_chunks) { int r = c * E + kk; src = r < seg0_rows ? r : -1; }
    else { int r = (c - seg0_chunks) * E + kk; src = r < seg1_rows ? seg0_rows + r : -1; }
    float v = (src >= 0 && n < n_out) ? (transpose_ld > 0 ? w[(size_t)n * transpose_ld + src] : w[(size_t)src * n_out + n]) : 0.f;
    uint8_t* base = pw + (size_t)c * chunk_bytes;
    if (BF16) {
      uint32_t off = tc::sw128_off(n, kk / 8) + (kk % 8) * 2;
      *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16_rn(v);
    } else {
      // plane H: tf32 hi (32 fp32 words per row); plane C: 64 bf16 per row = [bf16(w_hi) x 32 | bf16(w_lo) x 32],
      // the operands of the two correction products a_lo.w_hi and a_hi.w_lo (see store_chunk32)
      uint32_t off = tc::sw128_off(n, kk / 4) + (kk % 4) * 4;
      float hi = tc::tf32_rna(v);
      *reinterpret_cast<float*>(base + off) = hi;
      uint8_t* pc = base + plane;
      *reinterpret_cast<__nv_bfloat16*>(pc + tc::sw128_off(n, kk / 8) + (kk % 8) * 2) = __float2bfloat16_rn(hi);
      *reinterpret_cast<__nv_bfloat16*>(pc + tc::sw128_off(n, 4 + kk / 8) + (kk % 8) * 2) = __float2bfloat16_rn(v - hi);
    }
  }
  for (int c = block * blockDim.x + threadIdx.x; c < Npad; c += n_blocks * blockDim.x)
    pb[c] = (b && c < n_out) ? b[c] : 0.f;
}

template <bool BF16>
__global__ void tc_pack_kernel(const float* __restrict__ w, const float* __restrict__ b, int n_out, int Npad,
                               int seg0_rows, int seg0_chunks, int seg1_rows, int seg1_chunks,
                               uint8_t* __restrict__ pw, float* __restrict__ pb, int transpose_ld) {
  tc_pack_body<BF16>(w, b, n_out, Npad, seg0_rows, seg0_chunks, seg1_rows, seg1_chunks, pw, pb, transpose_ld,
                     (int)blockIdx.x, (int)gridDim.x);
}

// every layer of several networks in ONE launch (the training step refreshes the images of its eight networks after each
// optimizer step: 25 launches of ~4 us each otherwise); TC_PACK_BLOCKS blocks per layer
#define TC_PACK_JOBS 48
#define TC_PACK_BLOCKS 16
struct TcPackJob { const float* w; const float* b; uint8_t* pw; float* pb; int n_out, Npad, seg0_rows, seg0_chunks, seg1_rows, seg1_chunks, transpose_ld, pad_; };
struct TcPackBatch { int count; TcPackJob j[TC_PACK_JOBS]; };
template <bool BF16>
__global__ void tc_pack_batched_kernel(const __grid_constant__ TcPackBatch pb) {
  const TcPackJob& j = pb.j[blockIdx.x / TC_PACK_BLOCKS];
  tc_pack_body<BF16>(j.w, j.b, j.n_out, j.Npad, j.seg0_rows, j.seg0_chunks, j.seg1_rows, j.seg1_chunks, j.pw, j.pb,
                     j.transpose_ld, (int)(blockIdx.x % TC_PACK_BLOCKS), TC_PACK_BLOCKS);
}

static int tc_pack_fill(vqn_net* net, int p, cudaStream_t s) {
  TcPack* tp = net->tc_pack[p];
  const vqn_net_desc& d = net->desc;
  const bool bf16 = (p == 1);
  const int E = bf16 ? 64 : 32;
  for (int i = 0; i < d.n_layers; ++i) {
    bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
    int seg0_rows = (i == 0) ? d.in_dim : d.widths[i - 1];
    int seg1_rows = after_skip ? d.in_dim : 0;
    int c0 = vqn_round_up(seg0_rows, E) / E, c1 = vqn_round_up(seg1_rows, E) / E;
    if (bf16) tc_pack_kernel<true><<<128, 256, 0, s>>>(d.w[i], d.b[i], d.widths[i], tp->Npad[i], seg0_rows, c0,
                                                      seg1_rows, c1, tp->w[i], tp->bias[i], 0);
    else tc_pack_kernel<false><<<128, 256, 0, s>>>(d.w[i], d.b[i], d.widths[i], tp->Npad[i], seg0_rows, c0,
                                                   seg1_rows, c1, tp->w[i], tp->bias[i], 0);
    net->ctx->launches.fetch_add(1);
    VQN_CUDA(cudaGetLastError());
    if (tp->has_T && tp->wT[i]) {
      const int in_rows = seg0_rows + seg1_rows;             // Keras kernel rows (y rows, then the x rows after a skip)
      const int kc = vqn_round_up(d.widths[i], E) / E;
      if (bf16) tc_pack_kernel<true><<<128, 256, 0, s>>>(d.w[i], nullptr, in_rows, tp->NpadT[i], d.widths[i], kc, 0, 0,
                                                        tp->wT[i], tp->biasT[i], d.widths[i]);
      else tc_pack_kernel<false><<<128, 256, 0, s>>>(d.w[i], nullptr, in_rows, tp->NpadT[i], d.widths[i], kc, 0, 0,
                                                     tp->wT[i], tp->biasT[i], d.widths[i]);
      net->ctx->launches.fetch_add(1);
      VQN_CUDA(cudaGetLastError());
    }
  }
  return VQN_OK;
}

// reverse-mode gradient: allocate + fill the transposed images of an existing pack
static int tc_packT_ensure(vqn_net* net, int precision, cudaStream_t s) {
  const int p = precision == VQN_PREC_BF16 ? 1 : 0;
  TcPack* tp = net->tc_pack[p];
  if (!tp || tp->has_T) return VQN_OK;
  const vqn_net_desc& d = net->desc;
  const int E = p == 1 ? 64 : 32;
  for (int i = 0; i < d.n_layers; ++i) {
    const bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
    const int in_rows = ((i == 0) ? d.in_dim : d.widths[i - 1]) + (after_skip ? d.in_dim : 0);
    tp->NpadT[i] = vqn_round_up(in_rows, i > 0 ? E : 16);   // layer 0's input gradient is only contracted, not chained
    tp->n_chunksT[i] = vqn_round_up(d.widths[i], E) / E;
    if (tp->NpadT[i] > TC_NPAD_MAX) {                       // no transposed image: such a layer cannot be an MMA layer of a
      tp->NpadT[i] = 0; tp->n_chunksT[i] = 0;               // backward chain (vqn_sdf_forward / vqn_net_backward_train check)
      tp->wT[i] = nullptr; tp->biasT[i] = nullptr;
      continue;
    }
    const size_t chunk_bytes = (size_t)tp->NpadT[i] * 128 * (p == 1 ? 1 : 2);
    VQN_CUDA(cudaMalloc(&tp->wT[i], chunk_bytes * tp->n_chunksT[i]));
    VQN_CUDA(cudaMalloc(&tp->biasT[i], sizeof(float) * tp->NpadT[i]));
  }
  tp->has_T = true;
  return tc_pack_fill(net, p, s);
}

static int tc_pack_get(vqn_net* net, int precision, cudaStream_t s, TcPack** out) {
  const int p = precision == VQN_PREC_BF16 ? 1 : 0;
  if (net->tc_pack[p]) { *out = net->tc_pack[p]; return VQN_OK; }
  const vqn_net_desc& d = net->desc;
  const int E = p == 1 ? 64 : 32;
  TcPack* tp = new TcPack();
  memset(tp, 0, sizeof(*tp));
  tp->n_layers = d.n_layers;
  for (int i = 0; i < d.n_layers; ++i) {
    bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
    int seg0_rows = (i == 0) ? d.in_dim : d.widths[i - 1];
    int seg1_rows = after_skip ? d.in_dim : 0;
    // hidden layers are drained in whole K-chunks of the next layer: pad their accumulators to a chunk multiple (the
    // pad columns have zero weights and meet zero-filled weight rows of the next layer, e.g. NeuS' 217-wide lin3)
    tp->Npad[i] = vqn_round_up(d.widths[i], i + 1 < d.n_layers ? E : 16);
    tp->n_chunks[i] = vqn_round_up(seg0_rows, E) / E + vqn_round_up(seg1_rows, E) / E;
    size_t chunk_bytes = (size_t)tp->Npad[i] * 128 * (p == 1 ? 1 : 2);
    VQN_CUDA(cudaMalloc(&tp->w[i], chunk_bytes * tp->n_chunks[i]));
    VQN_CUDA(cudaMalloc(&tp->bias[i], sizeof(float) * tp->Npad[i]));
  }
  net->tc_pack[p] = tp;
  int rc = tc_pack_fill(net, p, s);
  if (rc != VQN_OK) return rc;
  *out = tp;
  return VQN_OK;
}

// called from vqn_net_create (nothing to do: packs are lazy) and vqn_net_repack (refresh existing images)
int vqn_tc_pack_create(vqn_net* net, cudaStream_t s) {
  for (int p = 0; p < 2; ++p)
    if (net->tc_pack[p]) { int rc = tc_pack_fill(net, p, s); if (rc != VQN_OK) return rc; }
  return VQN_OK;
}

void vqn_tc_pack_destroy(vqn_net* net) {
  for (int p = 0; p < 2; ++p) {
    TcPack* tp = net->tc_pack[p];
    if (!tp) continue;
    for (int i = 0; i < tp->n_layers; ++i) {
      cudaFree(tp->w[i]); cudaFree(tp->bias[i]);
      if (tp->has_T) { cudaFree(tp->wT[i]); cudaFree(tp->biasT[i]); }
    }
    delete tp;
    net->tc_pack[p] = nullptr;
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool BF16>
struct TcCfg {
  static constexpr int E = BF16 ? 64 : 32;                 // K elements per chunk
  static constexpr int PLANES = BF16 ? 1 : 2;
  // Ring depths.  A slot = one 32-K chunk of the 128-row tile (tf32x3: tf32 plane + correction plane = 32 KB).  A slot is
  // busy from the producers' first store until the MMAs reading it complete, so a ring of S slots sustains one chunk per
  // (store time + MMA time + hand-off latency) / S: with S = 2 that was ~1250 cycles against 745 cycles of tensor time on
  // the N = 128 layers.  The third A slot is paid for by the WEIGHT ring: its entries are 32 KB -- both planes of a chunk
  // of a layer with Npad <= 128, ONE plane of a wider layer (the tf32 MMAs of a chunk need only plane H, the correction
  // MMAs only plane C; each entry is released by its own tcgen05.commit) -- instead of two 64 KB slots.
  static constexpr int SA = BF16 ? 4 : 3;                  // A ring stages
  static constexpr int SW = BF16 ? 4 : 3;                  // W ring entries
  static constexpr int G = 2;                              // epilogue warp-groups (8 warps each: a thread owns 16 columns of a row)
  static constexpr int THREADS = 32 * (8 * G + 2);         // + MMA warp + weight-producer warp
  static constexpr uint32_t A_PLANE = TC_M * 128;          // 16 KB
  static constexpr uint32_t A_SLOT = A_PLANE * PLANES;
  static constexpr uint32_t W_SLOT = (uint32_t)TC_NPAD_MAX * 128;    // 32 KB
  static constexpr size_t SMEM = (size_t)SA * A_SLOT + (size_t)SW * W_SLOT + 1024;
};
// A operand from tensor memory (tf32x3 plain chain): two A slots of 64 columns at the top of TMEM -- per 32-K chunk
// [tf32 hi x 32 | bf16 pairs of a_lo x 16 | bf16 pairs of a_hi x 16], written by the producers with tcgen05.st and read by
// TS-mode MMAs.  The kernel is bound by the 128 B/clk shared-memory port (UMMA operand reads + A-chunk stores + weight
// copies: 128 KB per chunk of an N = 128 layer = 1024 cycles against 512 cycles of tensor time); a chunk that goes through
// TMEM takes its 32 KB of stores and 32 KB of operand reads off that port.  The host planner (tc_plan_ts) enables it per
// segment where the accumulators leave columns [TC_TS_COL, 512) free.
#define TC_TS_COL 384
#define TC_TS_SLOTS 2
#define TC_BIAS_FLOATS 7168
#define TC_STASH_SLOTS 8
// act' of the reverse-mode gradient is stashed as 16-bit fixed point (act' of relu / softplus / sigmoid lies in [0, 1]: step
// 1.5e-5, ~4e-5 on the gradient, inside its 1e-4 budget): 0.5 MB per CTA, 74 MB for the 148 CTAs -- under the 126 MB L2,
// where the fp32 stash (148 MB) was streamed through HBM (15 GB of DRAM traffic per 1 M points against 1.1 GB algorithmic)
#define TC_STASH_FLOATS (TC_STASH_SLOTS * 16 * TC_M * 8)      // 16 values = 32 bytes = 8 float slots; 0.5 MB per CTA

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// torch.nn.Softplus(beta=100) (fields.py:72): log1p(exp(100 x)) / 100 (x itself beyond the threshold 100 x > 20, where
// the formula below differs by < 2.1e-11).  lg2.approx is accurate to 2^-22 ABSOLUTE on (0.5, 2): 1.7e-9 after the /100.
__device__ __forceinline__ float softplus100(float t) {
  const float bt = 100.0f * t;
  const float e = __expf(-fabsf(bt));
  return (fmaxf(bt, 0.0f) + __logf(1.0f + e)) * 0.01f;
}

// Activation of a jet tile: lane 4p holds the value row of point p, lanes 4p+1..3 its tangent rows.
//   value row:   act(v + b)          tangent rows:   v * act'(pre-activation of the value row)
__device__ __forceinline__ void bias_act32_jet(float (&v)[16], const float* __restrict__ bias_s, int act, int lane) {
  const int src = lane & ~3;
  const bool is_val = (lane & 3) == 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float pre = __shfl_sync(0xffffffffu, v[j], src) + bias_s[j];
    float val, der;
    if (act == VQN_ACT_SOFTPLUS100) {
      const float bt = 100.0f * pre;
      const float e = __expf(-fabsf(bt));
      const float s = __fdividef(1.0f, 1.0f + e);
      val = (fmaxf(bt, 0.0f) + __logf(1.0f + e)) * 0.01f;
      der = bt >= 0.0f ? s : e * s;
    } else if (act == VQN_ACT_RELU) {
      val = fmaxf(pre, 0.0f); der = pre > 0.0f ? 1.0f : 0.0f;
    } else if (act == VQN_ACT_SIGMOID) {
      val = __fdividef(1.0f, 1.0f + __expf(-pre)); der = val * (1.0f - val);
    } else {
      val = pre; der = 1.0f;
    }
    v[j] = is_val ? val : v[j] * der;
  }
}

// Reverse mode: value AND act' of 16 pre-activations; act' goes to the stash (4 x 16-byte stores, a warp writes 2 KB)
__device__ __forceinline__ void bias_act32_stash(float (&v)[16], const float* __restrict__ bias_s, int act,
                                                 float* __restrict__ stash16) {
  float d[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float pre = v[j] + bias_s[j];
    if (act == VQN_ACT_SOFTPLUS100) {
      const float bt = 100.0f * pre;
      const float e = __expf(-fabsf(bt));
      const float s = __fdividef(1.0f, 1.0f + e);
      v[j] = (fmaxf(bt, 0.0f) + __logf(1.0f + e)) * 0.01f;
      d[j] = bt >= 0.0f ? s : e * s;
    } else if (act == VQN_ACT_RELU) {
      v[j] = fmaxf(pre, 0.0f); d[j] = pre > 0.0f ? 1.0f : 0.0f;
    } else if (act == VQN_ACT_SIGMOID) {
      v[j] = __fdividef(1.0f, 1.0f + __expf(-pre)); d[j] = v[j] * (1.0f - v[j]);
    } else {
      v[j] = pre; d[j] = 1.0f;
    }
  }
  // d in [0, 1] -> round(65535 d) through the 2^23 magic add (the integer lands in the low mantissa bits), two per word
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t lo = __float_as_uint(fmaf(d[2 * j], 65535.0f, 8388608.0f));
    const uint32_t hi = __float_as_uint(fmaf(d[2 * j + 1], 65535.0f, 8388608.0f));
    w[j] = __byte_perm(lo, hi, 0x5410);
  }
  // L2 evict_last: the 74 MB of stash of the 148 CTAs are re-read within the same tile and rewritten by the next one; the
  // output streams (1 KB of feature per point) must not push them out to HBM
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(stash16), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "l"(pol) : "memory");
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(stash16 + 4), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "l"(pol) : "memory");
}
__device__ __forceinline__ void stash_load16(const float* __restrict__ stash16, float (&d)[16]) {
  const uint4 a = __ldcg(reinterpret_cast<const uint4*>(stash16)), b = __ldcg(reinterpret_cast<const uint4*>(stash16) + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const float k = 1.0f / 65535.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    d[2 * j] = (__uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7610)) - 8388608.0f) * k;
    d[2 * j + 1] = (__uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7632)) - 8388608.0f) * k;
  }
}
// sum_k g[k] * d embed_{e0+k} / d x_j  for the 16 embedding columns e0 .. e0+15 (embedder.py: [x, sin(x f), cos(x f), ...])
__device__ __forceinline__ void embed_jac_contract(const float (&g)[16], int e0, int n_freqs, const float (&x)[3],
                                                   float& a0, float& a1, float& a2) {
  const int d = 3 + 6 * n_freqs;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int k = e0 + j;
    if (k < 0 || k >= d) continue;
    int c; float w;
    if (k < 3) { c = k; w = 1.0f; }
    else {
      const int f = (k - 3) / 6, q = (k - 3) % 6;
      c = q % 3;
      const float fr = exp2f((float)f);
      float sn, cs;
      sincosf(x[c] * fr, &sn, &cs);
      w = q < 3 ? fr * cs : -fr * sn;
    }
    const float t = g[j] * w;
    a0 += c == 0 ? t : 0.0f; a1 += c == 1 ? t : 0.0f; a2 += c == 2 ? t : 0.0f;
  }
}

template <int ACT>
__device__ __forceinline__ void bias_act32(float (&v)[16], const float* __restrict__ bias_s) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j);
    float t0 = v[j] + b4.x, t1 = v[j + 1] + b4.y, t2 = v[j + 2] + b4.z, t3 = v[j + 3] + b4.w;
    if (ACT == VQN_ACT_RELU) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); t2 = fmaxf(t2, 0.f); t3 = fmaxf(t3, 0.f); }
    if (ACT == VQN_ACT_SIGMOID) { t0 = fast_sigmoid(t0); t1 = fast_sigmoid(t1); t2 = fast_sigmoid(t2); t3 = fast_sigmoid(t3); }
    if (ACT == VQN_ACT_SOFTPLUS100) { t0 = softplus100(t0); t1 = softplus100(t1); t2 = softplus100(t2); t3 = softplus100(t3); }
    v[j] = t0; v[j + 1] = t1; v[j + 2] = t2; v[j + 3] = t3;
  }
}
__device__ __forceinline__ void bias_act32_dyn(float (&v)[16], const float* __restrict__ bias_s, int act) {
  if (act == VQN_ACT_RELU) bias_act32<VQN_ACT_RELU>(v, bias_s);
  else if (act == VQN_ACT_SIGMOID) bias_act32<VQN_ACT_SIGMOID>(v, bias_s);
  else if (act == VQN_ACT_SOFTPLUS100) bias_act32<VQN_ACT_SOFTPLUS100>(v, bias_s);
  else bias_act32<VQN_ACT_NONE>(v, bias_s);
}
__device__ __forceinline__ float act1_dyn(float t, int act) {
  return act == VQN_ACT_RELU ? fmaxf(t, 0.f)
       : act == VQN_ACT_SIGMOID ? fast_sigmoid(t) : act == VQN_ACT_SOFTPLUS100 ? softplus100(t) : t;
}

// store 16 consecutive K values (columns j0 .. j0+15 of the chunk, j0 % 16 == 0) of row r into the chunk slot
template <bool BF16>
__device__ __forceinline__ void store_chunk32(uint32_t slot, int r, int j0, const float (&v)[16]) {
  if (BF16) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
      __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
      tc::sts128(slot + tc::sw128_off(r, j0 / 8 + q), u);
    }
  } else {
    // plane H (tf32 hi) at slot, plane C at slot + 16 KB: [bf16(a_lo) x 32 | bf16(a_hi) x 32] per row.
    // j0 is 0 or 16, so the eight 16-byte chunks of this thread are two swizzled bases XOR small constants
    // ((c + k) ^ rx == (c ^ rx) ^ k for k below c's alignment): one LOP3 per store instead of xor + shift + add
    const uint32_t row = slot + (uint32_t)r * 128u;
    const uint32_t rx = (uint32_t)(r & 7);
    const uint32_t th = row + ((((uint32_t)j0 >> 2) ^ rx) << 4);                    // tf32 chunk j0 / 4
    const uint32_t tb = row + TC_M * 128 + ((((uint32_t)j0 >> 3) ^ rx) << 4);       // bf16 chunk j0 / 8 of the lo half
#pragma unroll
    for (int qq = 0; qq < 2; ++qq) {            // 8 K values per step
      float h[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { h[i] = tc::tf32_rna_fast(v[8 * qq + i]); l[i] = v[8 * qq + i] - h[i]; }
      tc::sts128(th ^ (uint32_t)(32 * qq), make_float4(h[0], h[1], h[2], h[3]));
      tc::sts128(th ^ (uint32_t)(32 * qq + 16), make_float4(h[4], h[5], h[6], h[7]));
      uint4 ul, uh;
      __nv_bfloat162 t0 = __floats2bfloat162_rn(l[0], l[1]), t1 = __floats2bfloat162_rn(l[2], l[3]);
      __nv_bfloat162 t2 = __floats2bfloat162_rn(l[4], l[5]), t3 = __floats2bfloat162_rn(l[6], l[7]);
      ul.x = *reinterpret_cast<uint32_t*>(&t0); ul.y = *reinterpret_cast<uint32_t*>(&t1);
      ul.z = *reinterpret_cast<uint32_t*>(&t2); ul.w = *reinterpret_cast<uint32_t*>(&t3);
      t0 = __floats2bfloat162_rn(h[0], h[1]); t1 = __floats2bfloat162_rn(h[2], h[3]);
      t2 = __floats2bfloat162_rn(h[4], h[5]); t3 = __floats2bfloat162_rn(h[6], h[7]);
      uh.x = *reinterpret_cast<uint32_t*>(&t0); uh.y = *reinterpret_cast<uint32_t*>(&t1);
      uh.z = *reinterpret_cast<uint32_t*>(&t2); uh.w = *reinterpret_cast<uint32_t*>(&t3);
      tc::sts128(tb ^ (uint32_t)(16 * qq), ul);                      // bf16 chunk j0 / 8 + qq of the lo half
      tc::sts128(tb ^ (uint32_t)(16 * qq + 64), uh);                 // ... and of the hi half (chunk + 4)
    }
  }
}

// one scalar K value (column `col` of the chunk) of row r
template <bool BF16>
__device__ __forceinline__ void store_chunk1(uint32_t slot, int r, int col, float v) {
  if (BF16) {
    tc::sts16(slot + tc::sw128_off(r, col / 8) + (col % 8) * 2, __float2bfloat16_rn(v));
  } else {
    const float h = tc::tf32_rna(v);
    tc::sts32(slot + tc::sw128_off(r, col / 4) + (col % 4) * 4, h);
    const uint32_t pc = slot + TC_M * 128;
    tc::sts16(pc + tc::sw128_off(r, col / 8) + (col % 8) * 2, __float2bfloat16_rn(v - h));
    tc::sts16(pc + tc::sw128_off(r, 4 + col / 8) + (col % 8) * 2, __float2bfloat16_rn(h));
  }
}

// Embedder.__call__ (embedder.py:35-47) for the chunk covering embedding columns [col0, col0 + E):
// [x, sin(x f0), cos(x f0), sin(x f1), ...], f_k = 2^k; written straight into the swizzled slot.
// jc = 0: the embedding itself; jc = 1..3 (tangent rows of a jet tile): its derivative with respect to x[jc-1], i.e.
// [e_m, f cos(x_m f) e_m, -f sin(x_m f) e_m, ...] with m = jc - 1.
template <bool BF16>
__device__ __forceinline__ void embed_chunk(uint32_t slot, int r, const float (&x)[3], int col0, int n_freqs, int jc) {
  constexpr int E = BF16 ? 64 : 32;
  const int d = 3 + 6 * n_freqs;
  if (col0 == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) store_chunk1<BF16>(slot, r, k, jc == 0 ? x[k] : (jc - 1 == k ? 1.0f : 0.0f));
  }
  int f_lo = col0 <= 3 ? 0 : (col0 - 3 - 5 + 5) / 6;       // first frequency with a column >= col0
  if (f_lo > 0 && 3 + 6 * (f_lo - 1) + 5 >= col0) f_lo -= 1;
  int f_hi = (col0 + E - 1 - 3) / 6;
  if (f_hi > n_freqs - 1) f_hi = n_freqs - 1;
  // sincosf (Cody-Waite reduction, ~40 instructions per call) only for every other octave; the odd octaves use ONE
  // double-angle step from the previous one (sin 2a = 2 s c, cos 2a = 1 - 2 s^2): <= ~4e-7 absolute error,
  // far inside the 1e-4 budget, for half the transcendental work of the embedding.
  float sn[3], cs[3];
  bool have_prev = false;
  for (int f = f_lo; f <= f_hi; ++f) {
    if ((f & 1) && have_prev) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float s2 = 2.0f * sn[k] * cs[k], c2 = fmaf(-2.0f * sn[k], sn[k], 1.0f);
        sn[k] = s2; cs[k] = c2;
      }
      have_prev = false;
    } else {
      const float sc = exp2f((float)f);
#pragma unroll
      for (int k = 0; k < 3; ++k) sincosf(x[k] * sc, &sn[k], &cs[k]);
      have_prev = true;
    }
    const int base = 3 + 6 * f - col0;                       // chunk-relative column of sin(x0 f)
    const float fr = exp2f((float)f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float vs = jc == 0 ? sn[k] : (jc - 1 == k ? fr * cs[k] : 0.0f);
      const float vc = jc == 0 ? cs[k] : (jc - 1 == k ? -fr * sn[k] : 0.0f);
      if (base + k >= 0 && base + k < E) store_chunk1<BF16>(slot, r, base + k, vs);
      if (base + 3 + k >= 0 && base + 3 + k < E) store_chunk1<BF16>(slot, r, base + 3 + k, vc);
    }
  }
  for (int col = (d > col0 ? d : col0); col < col0 + E; ++col) store_chunk1<BF16>(slot, r, col - col0, 0.f);
}

// named barrier of one 256-thread epilogue group (ids 1, 2; id 0 is __syncthreads)
__device__ __forceinline__ void group_bar(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

// MODE 0: plain chain (decomposition stage, SDF value-only); 1: jets (4 rows per point); 2: reverse-mode gradient;
// 3: plain chain that also stores every hidden layer's output (training forward)
template <bool BF16, int MODE, bool TS>
__global__ void __launch_bounds__(TcCfg<BF16>::THREADS, 1) mlp_tc_kernel(const __grid_constant__ TcProgram pg) {
  using C = TcCfg<BF16>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[4], a_empty[4], w_full[4], w_empty[4], acc_full, drain_done;
  __shared__ __align__(8) uint64_t t_full[TC_TS_SLOTS], t_empty[TC_TS_SLOTS];       // TMEM A slots (TS)
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[TC_BIAS_FLOATS];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + (size_t)C::SA * C::A_SLOT;
  const uint32_t a_ring_s = tc::smem_u32(a_ring);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MMA_WARP = 8 * C::G, W_WARP = 8 * C::G + 1;

#ifdef VQN_TC_TRACE
  // whole-kernel stamps of CTA 0 (slot [3][19] of the MMA-thread table): entry, prologue done, tile loops done, exit
  long long* ktr = (pg.trace && blockIdx.x == 0 && tid == 0) ? pg.trace + (3 * TC_MAX_LAYERS + 19) * 4 : nullptr;
  if (ktr) ktr[0] = clock64();
#endif
  if (warp == MMA_WARP) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&a_full[i], 256); tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < TC_TS_SLOTS; ++i) { tc::mbar_init(&t_full[i], 256); tc::mbar_init(&t_empty[i], 1); }
    tc::mbar_init(&acc_full, 1);
    tc::mbar_init(&drain_done, 256 * C::G);
    tc::mbar_fence_init();
  }
  // stage every layer's bias in shared memory (layer l at bias_off[l]; host guarantees the total fits)
  for (int l = 0; l < pg.n_layers; ++l) {
    for (int i = tid; i < pg.layers[l].Npad; i += C::THREADS) bias_s[pg.layers[l].bias_off + i] = pg.layers[l].bias[i];
    if (pg.layers[l].skip_w)          // skip kernel rows as float4 {w[k][0..3]} (zero padded)
      for (int i = tid; i < pg.g_dim * 4; i += C::THREADS) {
        const int k = i >> 2, o = i & 3;
        bias_s[pg.layers[l].skip_woff + i] = o < pg.layers[l].skip_n ? pg.layers[l].skip_w[k * pg.layers[l].skip_n + o] : 0.f;
      }
    if (pg.layers[l].tg_w)            // training backward: rows of the narrow last layer as float4 {w[c][0..2]} (y part, x part)
      for (int i = tid; i < (pg.layers[l].tg_rows + pg.layers[l].fin_skip_n) * 4; i += C::THREADS) {
        const int k = i >> 2, o = i & 3;
        const int row = k < pg.layers[l].tg_rows ? k : pg.layers[l].fin_skip_row0 + (k - pg.layers[l].tg_rows);
        bias_s[pg.layers[l].tg_woff + i] = o < pg.layers[l].tg_n ? pg.layers[l].tg_w[row * pg.layers[l].tg_n + o] : 0.f;
      }
    if (pg.layers[l].tail_w)          // tail kernel rows as float4 {w[k][0..3]} (zero padded)
      for (int i = tid; i < pg.layers[l].N * 4; i += C::THREADS) {
        const int k = i >> 2, o = i & 3;
        bias_s[pg.layers[l].tail_woff + i] = o < pg.layers[l].tail_n ? pg.layers[l].tail_w[k * pg.layers[l].tail_n + o] : 0.f;
      }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
#ifdef VQN_TC_TRACE
  if (ktr) ktr[1] = clock64();
#endif

  long long n = pg.n_dev ? (long long)*pg.n_dev : pg.n;
  if (n > pg.n) n = pg.n;
  constexpr bool jet = MODE == 1;                        // compile-time: the decomposition-stage kernels carry no jet code
  constexpr bool rev = MODE == 2;
  constexpr bool trn = MODE == 3;
  constexpr bool bwd = MODE == 4;
  constexpr int tile_pts = jet ? TC_M / 4 : TC_M;        // points per tile
  const long long n_tiles = (n + tile_pts - 1) / tile_pts;
  const int L = pg.n_layers;

  if (warp < 8 * C::G) {
    // ============== epilogue warp-groups: A-chunk producers + accumulator drain ==============
    // group g (8 warps) produces the chunks with (global chunk index % G == g); all groups walk the same sequence.
    // Warps w and w + 4 of a group share a TMEM lane quarter; `half` selects which 16 of every 32 columns a
    // thread handles, which halves the serial latency of a chunk (load, activation, split, 8 instead of 16 stores).
    const int grp = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int tg = tid & 255;                            // thread index inside the group
    const int r = 32 * (warp & 3) + lane;                // TMEM lane == point row of the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t ga = 0;                                     // global A-chunk counter (chunk -> group)
    uint32_t gs = 0, gt = 0;                             // chunks handed over in shared memory / in tensor memory so far
    uint32_t gl = 0;                                     // global layer counter (acc_full phase)
#ifdef VQN_TC_TRACE
    int ptrace_n = 0;
#endif
    float sp0 = 0.f, sp1 = 0.f, sp2 = 0.f, sp3 = 0.f;   // folded-skip partial sums of this thread (see TcLayer::skip_w)
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long pi = tile * tile_pts + (jet ? (r >> 2) : r);   // compact point index
      const int jc = jet ? (r & 3) : 0;                  // jet component of this row (0 = value)
      const bool valid = pi < n;
      const bool valid_out = valid && jc == 0;           // rows whose network outputs are stored
      float x[3] = {0.f, 0.f, 0.f};
      if (pg.pts && valid) {
        long long row = pg.row_idx ? (long long)pg.row_idx[pi] : pi;
        x[0] = pg.pts[row * 3]; x[1] = pg.pts[row * 3 + 1]; x[2] = pg.pts[row * 3 + 2];
      }
      for (int l = 0; l < L; ++l) {
        const int nseg = pg.layers[l].nseg;
        if (rev && l > 0 && pg.layers[l - 1].egrad_col0 >= 0 && !pg.layers[l - 1].grad_final) {
          // skip half of the skip_in layer's input gradient: d out / d embedding, contracted with the embedding Jacobian
          // right away (3 partial sums in registers) -- the columns are not an operand of any later layer
          const TcLayer& pl = pg.layers[l - 1];
          tc::mbar_wait(&acc_full, (gl - 1) & 1);
          tc::fence_after_sync();
          const int t4 = grp * 2 + half;
          const int c16 = (pl.egrad_col0 & ~15) + 16 * t4;
          if (c16 < pl.Npad) {                                   // warp-uniform
            float v[16];
            tc::tmem_ld16(lane_addr + (uint32_t)(pl.tmem_col + c16), v);
            embed_jac_contract(v, c16 - pl.egrad_col0, pg.n_freqs, x, sp0, sp1, sp2);
          }
        }
        for (int sg = 0; sg < nseg; ++sg) {
          const int st = pg.layers[l].seg_type[sg];
          const int nch = pg.layers[l].seg_chunks[sg];
          const int first = pg.layers[l].seg_first_chunk[sg];
          const int pact = l > 0 ? pg.layers[l - 1].act : 0;
          const float* pbias = bias_s + (l > 0 ? pg.layers[l - 1].bias_off : 0);
          const uint32_t pacc = lane_addr + (uint32_t)(l > 0 ? pg.layers[l - 1].tmem_col : 0);
          bool acc_ready = false;
#ifdef TC_TS_NO_PROD
          const bool ts = false;
#else
          const bool ts = TS && pg.layers[l].seg_ts[sg];
#endif
          // this group's chunks of the segment: c = first, first + G, ... (chunk ga0 + c belongs to group (ga0 + c) % G)
          const uint32_t ga0 = ga, gs0 = gs, gt0 = gt;
          ga += (uint32_t)nch;
          if (ts) gt += (uint32_t)nch; else gs += (uint32_t)nch;
          for (int c = (int)(((uint32_t)grp + C::G - ga0 % C::G) % C::G); c < nch; c += C::G) {
            const uint32_t gsm = gs0 + (uint32_t)c, gtm = gt0 + (uint32_t)c;  // this chunk's position in its ring
            if ((st == SRC_DRAIN || st == SRC_GRAD || st == SRC_BWD) && !acc_ready) {
              tc::mbar_wait(&acc_full, (gl - 1) & 1);    // previous layer's accumulator is complete
              tc::fence_after_sync();
              acc_ready = true;
            }
            const int slot = gsm % C::SA;
            const uint32_t a_par = ((gsm / C::SA) & 1) ^ 1;                     // parity of the slot's "empty" phase
            const uint32_t dst = a_ring_s + (uint32_t)slot * C::A_SLOT;          // shared-space address of the slot
            const int sc = first + c;                     // chunk index inside the source
#ifdef VQN_TC_TRACE
            long long* ptr_ = nullptr;                    // diagnostic stamps of one producer thread (tile 1 of CTA 0)
            if (pg.trace && blockIdx.x == 0 && tid == 0 && tile == (long long)gridDim.x && ptrace_n < 96)
              ptr_ = pg.trace + 4 * TC_MAX_LAYERS * 4 + 8 * (ptrace_n++);
            if (ptr_) { ptr_[0] = clock64(); ptr_[6] = l * 1000 + sg * 100 + c; }
#endif
#ifdef TC_EXP_NO_PROD
            if (sc >= 0) {
              tc::mbar_wait(&a_empty[slot], a_par);
            } else
#endif
            if (st == SRC_EMBED) {
              tc::mbar_wait(&a_empty[slot], a_par);
              if (half == 0) embed_chunk<BF16>(dst, r, x, sc * C::E, pg.n_freqs, jc);
            } else if (!BF16 && st == SRC_GLOBAL && !pg.g_local) {
              // Latent chunk [128 rows x 32 columns] of the row-major source: loaded COALESCED (a warp reads 4 rows x
              // 128 B per instruction; a thread-per-row load touches 32 lines per instruction and costs 8x the L1/smem
              // data-pipe wavefronts, which is the port this kernel is bound by), staged through the plane-C half of the
              // group's own slot with the 128-B swizzle (conflict-free both ways), then re-read row-wise.
              float4 ldv[4];
              const int col_base = sc * C::E;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int f = tg + 256 * i, rr = f >> 3, ch = f & 7;
                const long long prow = tile * TC_M + rr;
                ldv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (prow < n && col_base + 4 * ch < pg.g_dim) {
                  // per-CTA scratch: written earlier in this kernel by other threads -> L2-coherent load, no L1 line
                  ldv[i] = __ldg(reinterpret_cast<const float4*>(pg.gsrc + prow * pg.g_dim + col_base + 4 * ch));
                }
              }
              TC_STAMP(ptr_, 1);
              tc::mbar_wait(&a_empty[slot], a_par);
              TC_STAMP(ptr_, 2);
              const uint32_t stage = dst + C::A_PLANE;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int f = tg + 256 * i, rr = f >> 3, ch = f & 7;
                tc::sts128(stage + rr * 128 + ((ch ^ (rr & 7)) << 4), ldv[i]);
              }
              group_bar(grp);
              float v[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 t = tc::lds128(stage + r * 128 + (((4 * half + q) ^ (r & 7)) << 4));
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
              }
              group_bar(grp);                    // every row has been read before plane C is overwritten
              if (pg.layers[l].skip_w) {         // folded skip connection: x . W_x for this thread's 16 values
                const float4* sw = reinterpret_cast<const float4*>(bias_s + pg.layers[l].skip_woff) + col_base + 16 * half;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float4 w4 = sw[j];
                  sp0 = fmaf(v[j], w4.x, sp0); sp1 = fmaf(v[j], w4.y, sp1); sp2 = fmaf(v[j], w4.z, sp2); sp3 = fmaf(v[j], w4.w, sp3);
                }
              }
              store_chunk32<BF16>(dst, r, 16 * half, v);
            } else {
#pragma unroll
              for (int h = 0; h < C::E / 32; ++h) {
                float v[16];
                const int col0 = sc * C::E + 32 * h + 16 * half;   // first source column of this thread's 16-wide piece
                if (st == SRC_DRAIN) {
                  tc::tmem_ld16(pacc + (uint32_t)col0, v);
                  if (jet) bias_act32_jet(v, pbias + col0, pact, lane);
                  else if (rev && pg.layers[l - 1].stash_w >= 0 && !pg.layers[l - 1].tail_w)
                    bias_act32_stash(v, pbias + col0, pact,
                                     pg.stash + (size_t)blockIdx.x * TC_STASH_FLOATS +
                                         ((size_t)(pg.layers[l - 1].stash_w * 16 + (col0 >> 4)) * TC_M + r) * 8);
                  else bias_act32_dyn(v, pbias + col0, pact);
                  if (trn && pg.layers[l - 1].save && valid) {      // the previous layer's output, kept for the backward pass
                    const TcLayer& pl = pg.layers[l - 1];
                    float* sv = pl.save + (size_t)pi * pl.save_ld + col0;
                    if (col0 + 16 <= pl.N) {
                      tc::stg_row16(sv, v);
                    } else {
#pragma unroll
                      for (int j = 0; j < 16; ++j) if (col0 + j < pl.N) sv[j] = v[j];
                    }
                  }
                } else if (bwd && (st == SRC_BWD || st == SRC_TAILGRAD)) {
                  // dz of forward layer i = (gradient w.r.t. its output) * act'(its saved output); stored for the weight gradients
                  const TcLayer& ly = pg.layers[l];
                  if (st == SRC_BWD) tc::tmem_ld16(pacc + (uint32_t)col0, v);
                  else {
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
                    if (valid) {
                      const float* dzr = ly.tg_dz + (size_t)pi * ly.tg_dz_ld;
                      d0 = dzr[0]; d1 = ly.tg_n > 1 ? dzr[1] : 0.f; d2 = ly.tg_n > 2 ? dzr[2] : 0.f;
                    }
                    const float4* tw = reinterpret_cast<const float4*>(bias_s + ly.tg_woff) + col0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) { const float4 w4 = tw[j]; v[j] = fmaf(d0, w4.x, fmaf(d1, w4.y, d2 * w4.z)); }
                  }
                  if (valid && col0 < ly.bw_n) {
                    float y[16];
                    const float* yr = ly.bw_y + (size_t)pi * ly.bw_y_ld + col0;
                    float* dr = ly.bw_dz + (size_t)pi * ly.bw_dz_ld + col0;
                    if (col0 + 16 <= ly.bw_n) {
                      tc::ldg_row16(yr, y);
                    } else {
#pragma unroll
                      for (int j = 0; j < 16; ++j) y[j] = col0 + j < ly.bw_n ? __ldg(yr + j) : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const float dact = ly.bw_act == VQN_ACT_RELU ? (y[j] > 0.f ? 1.f : 0.f)
                                       : ly.bw_act == VQN_ACT_SIGMOID ? y[j] * (1.f - y[j]) : 1.f;
                      v[j] = col0 + j < ly.bw_n ? v[j] * dact : 0.f;
                    }
                    if (col0 + 16 <= ly.bw_n) {
                      tc::stg_row16(dr, v);
                    } else {
#pragma unroll
                      for (int j = 0; j < 16; ++j) if (col0 + j < ly.bw_n) dr[j] = v[j];
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                  }
                } else if (rev && (st == SRC_GRAD || st == SRC_GRADINIT)) {
                  float dv[16];
                  stash_load16(pg.stash + (size_t)blockIdx.x * TC_STASH_FLOATS +
                                   ((size_t)(pg.layers[l].grad_stash * 16 + (col0 >> 4)) * TC_M + r) * 8, dv);
                  if (st == SRC_GRAD) {
                    tc::tmem_ld16(pacc + (uint32_t)col0, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] *= dv[j];
                  } else {
                    const float4* tw = reinterpret_cast<const float4*>(bias_s + pg.layers[l].grad_tail_woff) + col0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = dv[j] * tw[j].x;
                  }
                } else {                                  // SRC_GLOBAL: this thread's own 16 values of its latent row
                  if (pg.g_local && col0 < pg.g_dim) {
                    // per-CTA scratch tile, ROW-INTERLEAVED ([col / 4][row][4]): a warp reads 32 rows x 16 B = 512 contiguous
                    // bytes per instruction -- coalesced without staging; written earlier in this kernel by other threads of
                    // the CTA -> L2-coherent loads, no L1 line
                    const float4* src = reinterpret_cast<const float4*>(pg.gsrc + (size_t)blockIdx.x * TC_M * pg.g_dim) +
                                        (size_t)(col0 >> 2) * TC_M + r;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float4 t = __ldcg(src + (size_t)j * TC_M);
                      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
                    }
                  } else if (valid && col0 < pg.g_dim) {
                    tc::ldg_row16(pg.gsrc + (size_t)pi * pg.g_dim + col0, v);
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                  }
                  if (pg.layers[l].skip_w) {     // folded skip connection (see the tf32 path)
                    const float4* sw = reinterpret_cast<const float4*>(bias_s + pg.layers[l].skip_woff) + col0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const float4 w4 = sw[j];
                      sp0 = fmaf(v[j], w4.x, sp0); sp1 = fmaf(v[j], w4.y, sp1); sp2 = fmaf(v[j], w4.z, sp2); sp3 = fmaf(v[j], w4.w, sp3);
                    }
                  }
                }
                // the values are ready in registers BEFORE the slot is claimed: the TMEM / global load latency
                // and the activation math overlap the MMAs that are still reading the slot's previous chunk
                if (h == 0) TC_STAMP(ptr_, 1);
                if (TS && ts) {
                  // hand-over in tensor memory: this thread's 16 K-values of its row -> 16 + 8 + 8 columns of the slot
                  const int tslot = gtm % TC_TS_SLOTS;
                  tc::mbar_wait(&t_empty[tslot], ((gtm / TC_TS_SLOTS) & 1) ^ 1);
                  tc::fence_after_sync();
                  TC_STAMP(ptr_, 2);
                  const uint32_t tdst = lane_addr + (uint32_t)(TC_TS_COL + 64 * tslot);
#pragma unroll
                  for (int qq = 0; qq < 2; ++qq) {
                    uint32_t hw[8], lw[4], cw[4];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                      const float h0 = tc::tf32_rna_fast(v[8 * qq + i]), h1 = tc::tf32_rna_fast(v[8 * qq + i + 1]);
                      hw[i] = __float_as_uint(h0); hw[i + 1] = __float_as_uint(h1);
                      __nv_bfloat162 pl = __floats2bfloat162_rn(v[8 * qq + i] - h0, v[8 * qq + i + 1] - h1);
                      __nv_bfloat162 ph = __floats2bfloat162_rn(h0, h1);
                      lw[i >> 1] = *reinterpret_cast<uint32_t*>(&pl); cw[i >> 1] = *reinterpret_cast<uint32_t*>(&ph);
                    }
                    tc::tmem_st8(tdst + (uint32_t)(16 * half + 8 * qq), hw);
                    tc::tmem_st4(tdst + (uint32_t)(32 + 8 * half + 4 * qq), lw);
                    tc::tmem_st4(tdst + (uint32_t)(48 + 8 * half + 4 * qq), cw);
                  }
                  tc::tmem_st_wait();
                  TC_STAMP(ptr_, 3);
                  tc::fence_before_sync();
                  TC_STAMP(ptr_, 4);
                  tc::mbar_arrive(&t_full[tslot]);
                  TC_STAMP(ptr_, 5);
                  continue;
                }
                if (h == 0) tc::mbar_wait(&a_empty[slot], a_par);
                if (h == 0) TC_STAMP(ptr_, 2);
                store_chunk32<BF16>(dst, r, 32 * h + 16 * half, v);
              }
              if (TS && ts) continue;
            }
            TC_STAMP(ptr_, 3);
            tc::fence_proxy_async();       // generic-proxy stores -> visible to the UMMA (async proxy)
            tc::fence_before_sync();       // order the tcgen05.ld's above before the hand-off
            TC_STAMP(ptr_, 4);
            tc::mbar_arrive(&a_full[slot]);
            TC_STAMP(ptr_, 5);
          }
        }
        ++gl;                               // layer l's chunks are all queued
        if (pg.layers[l].tail_w) {
          // narrow tail layer on the CUDA cores: out[o] = act(sum_k relu(acc[k] + b[k]) W_y[k][o] + x . W_x[o] + bias[o])
          const TcLayer& ly = pg.layers[l];
          tc::mbar_wait(&acc_full, (gl - 1) & 1);
          tc::fence_after_sync();
          const float* lb = bias_s + ly.bias_off;
          const float4* tw = reinterpret_cast<const float4*>(bias_s + ly.tail_woff);
          float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
          if (ly.tail_add_skip) { p0 = sp0; p1 = sp1; p2 = sp2; p3 = sp3; sp0 = sp1 = sp2 = sp3 = 0.f; }
          for (int cb = grp; cb * 32 < ly.N; cb += C::G) {
            const int c16 = cb * 32 + 16 * half;
            if (c16 < ly.N) {
              float v[16];
              tc::tmem_ld16(lane_addr + (uint32_t)(ly.tmem_col + c16), v);
              if (jet) bias_act32_jet(v, lb + c16, ly.act, lane);
              else if (rev && ly.stash_w >= 0)
                bias_act32_stash(v, lb + c16, ly.act, pg.stash + (size_t)blockIdx.x * TC_STASH_FLOATS +
                                                          ((size_t)(ly.stash_w * 16 + (c16 >> 4)) * TC_M + r) * 8);
              else bias_act32_dyn(v, lb + c16, ly.act);
              if (trn && ly.save && valid) {
                tc::stg_row16(ly.save + (size_t)pi * ly.save_ld + c16, v);
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 w4 = tw[c16 + j];
                p0 = fmaf(v[j], w4.x, p0); p1 = fmaf(v[j], w4.y, p1); p2 = fmaf(v[j], w4.z, p2); p3 = fmaf(v[j], w4.w, p3);
              }
            }
          }
          // the four partial sums of a row (2 groups x 2 column halves) meet in the (free) first A slot
          tc::sts128(a_ring_s + (uint32_t)(r * 4 + grp * 2 + half) * 16u, make_float4(p0, p1, p2, p3));
          tc::fence_before_sync();
          if (rev) __threadfence_block();      // the stash written above (and by the forward drains) is read by other threads
          asm volatile("bar.sync 3, %0;" ::"r"(256 * C::G) : "memory");
          if (grp == 0 && half == 0) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 2 * C::G; ++q) {
              const float4 t4 = tc::lds128(a_ring_s + (uint32_t)(r * 4 + q) * 16u);
              o[0] += t4.x; o[1] += t4.y; o[2] += t4.z; o[3] += t4.w;
            }
            float* go = pg.outs[ly.tail_out_slot];
            const int gs = pg.out_stride[ly.tail_out_slot];
            bool bad = false;
            if (jc != 0) {
              // tangent row of a jet tile (linear tail, tail_n == 1): d out / d x[jc-1]
              bad = !isfinite(o[0]);
              if (valid && pg.jet_grad) pg.jet_grad[pi * 3 + (jc - 1)] = o[0];
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (q < ly.tail_n) {
                  float t = act1_dyn(o[q] + ly.tail_b[q], ly.tail_act);
                  t = t * ly.tail_scale + ly.tail_bias;
                  bad |= !isfinite(t);
                  if (valid && go) go[pi * gs + q] = t;
                }
              }
            }
            if (valid && bad) atomicOr(pg.nonfinite, 1);
          }
          tc::fence_before_sync();
          asm volatile("bar.sync 3, %0;" ::"r"(256 * C::G) : "memory");   // scratch is free again
        }
        if (rev && pg.layers[l].grad_final) {
          // last backward layer: its accumulator is d out / d embedding; contract with the embedding Jacobian, add the
          // skip partials, combine the (up to) four partial gradients of a row through the free first A slot
          const TcLayer& ly = pg.layers[l];
          tc::mbar_wait(&acc_full, (gl - 1) & 1);
          tc::fence_after_sync();
          const int t4 = grp * 2 + half;
          const int c16 = (ly.egrad_col0 & ~15) + 16 * t4;
          if (c16 < ly.Npad) {                                   // warp-uniform
            float v[16];
            tc::tmem_ld16(lane_addr + (uint32_t)(ly.tmem_col + c16), v);
            embed_jac_contract(v, c16 - ly.egrad_col0, pg.n_freqs, x, sp0, sp1, sp2);
          }
          tc::sts128(a_ring_s + (uint32_t)(r * 4 + t4) * 16u, make_float4(sp0, sp1, sp2, 0.f));
          sp0 = sp1 = sp2 = 0.f;
          tc::fence_before_sync();
          asm volatile("bar.sync 3, %0;" ::"r"(256 * C::G) : "memory");
          if (t4 == 0) {
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 t = tc::lds128(a_ring_s + (uint32_t)(r * 4 + q) * 16u); g0 += t.x; g1 += t.y; g2 += t.z; }
            if (valid) {
              if (!isfinite(g0) || !isfinite(g1) || !isfinite(g2)) atomicOr(pg.nonfinite, 1);
              pg.jet_grad[pi * 3] = g0; pg.jet_grad[pi * 3 + 1] = g1; pg.jet_grad[pi * 3 + 2] = g2;
            }
          }
          tc::fence_before_sync();
          asm volatile("bar.sync 3, %0;" ::"r"(256 * C::G) : "memory");   // scratch is free again
        }
        if (bwd && pg.layers[l].out_slot >= 0) {
          // end of a backward chain: the accumulator is the gradient w.r.t. the chain's input (or, with fin_y, a last dz)
          const TcLayer& ly = pg.layers[l];
          tc::mbar_wait(&acc_full, (gl - 1) & 1);
          tc::fence_after_sync();
          float* go = pg.outs[ly.out_slot];
          const int gs = pg.out_stride[ly.out_slot];
          float d0 = 0.f, d1 = 0.f, d2 = 0.f;
          if (ly.fin_skip_n > 0 && valid) {
            const float* dzr = ly.tg_dz + (size_t)pi * ly.tg_dz_ld;
            d0 = dzr[0]; d1 = ly.tg_n > 1 ? dzr[1] : 0.f; d2 = ly.tg_n > 2 ? dzr[2] : 0.f;
          }
          for (int cb = grp; cb * 32 < ly.N; cb += C::G) {
            const int c16 = cb * 32 + 16 * half;
            if (c16 < ly.N) {                                   // warp-uniform
              float v[16];
              tc::tmem_ld16(lane_addr + (uint32_t)(ly.tmem_col + c16), v);
              if (valid) {
                if (ly.fin_skip_n > 0) {
                  const float4* sw = reinterpret_cast<const float4*>(bias_s + ly.tg_woff) + ly.tg_rows + c16;
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c16 + j < ly.fin_skip_n) { const float4 w4 = sw[j]; v[j] += fmaf(d0, w4.x, fmaf(d1, w4.y, d2 * w4.z)); }
                }
                // 16-byte accesses whenever the piece is whole and the rows are 16-byte aligned: a thread owns a row, so every
                // warp-level access touches 32 different sectors -- scalar loads / stores / atomics were 4x the L1 wavefronts
                // (23 us of a 54 us launch at 8192 rows)
                const bool whole = c16 + 16 <= ly.N;
                if (ly.fin_y) {
                  const float* yr = ly.fin_y + (size_t)pi * ly.fin_y_ld + c16;
                  float y[16];
                  if (whole && (ly.fin_y_ld & 3) == 0) {
                    tc::ldg_row16(yr, y);
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = c16 + j < ly.N ? __ldg(yr + j) : 0.f;
                  }
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    v[j] *= ly.fin_act == VQN_ACT_RELU ? (y[j] > 0.f ? 1.f : 0.f)
                          : ly.fin_act == VQN_ACT_SIGMOID ? y[j] * (1.f - y[j]) : 1.f;
                }
                float* dst = go + (size_t)pi * gs + c16;
                bool bad = false;
#pragma unroll
                for (int j = 0; j < 16; ++j) if (c16 + j < ly.N) bad |= !isfinite(v[j]);
                if (whole && (gs & 3) == 0 && (reinterpret_cast<uintptr_t>(go) & 15) == 0 && ly.fin_mode == 0) {
                  tc::stg_row16(dst, v);
                } else if (whole && (gs & 3) == 0 && (reinterpret_cast<uintptr_t>(go) & 15) == 0) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    float4* d4 = reinterpret_cast<float4*>(dst) + q;
                    if (ly.fin_mode == 2) {
                      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d4), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                                   "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
                    } else if (ly.fin_mode == 1) {
                      const float4 o = *d4;
                      *d4 = make_float4(o.x + v[4 * q], o.y + v[4 * q + 1], o.z + v[4 * q + 2], o.w + v[4 * q + 3]);
                    } else {
                      *d4 = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    if (c16 + j < ly.N) {
                      if (ly.fin_mode == 2) atomicAdd(dst + j, v[j]);
                      else if (ly.fin_mode == 1) dst[j] += v[j];
                      else dst[j] = v[j];
                    }
                  }
                }
                if (bad) atomicOr(pg.nonfinite, 1);
              }
            }
          }
          tc::fence_before_sync();
          tc::mbar_arrive(&drain_done);     // the MMA thread may now reuse this TMEM region
        } else
        if (pg.layers[l].out_slot >= 0) {
          // final layer of a network: drain its accumulator to global memory (column blocks split over groups)
          const TcLayer& ly = pg.layers[l];
          tc::mbar_wait(&acc_full, (gl - 1) & 1);
          tc::fence_after_sync();
          const bool local = pg.g_local && ly.out_slot == pg.local_slot;     // per-CTA scratch tile (wide path only)
          float* go = pg.outs[ly.out_slot];
          const int gs = pg.out_stride[ly.out_slot];
          const float* lb = bias_s + ly.bias_off;
          const bool wide = (ly.N % 32 == 0) && (gs % 4 == 0);   // full 32-column blocks of 16-byte aligned rows
          for (int cb = grp; cb * 32 < ly.N; cb += C::G) {
            const int c16 = cb * 32 + 16 * half;           // this thread's 16 columns of the 32-column block
            if (c16 < ly.N) {
              float v[16];
              tc::tmem_ld16(lane_addr + (uint32_t)(ly.tmem_col + c16), v);
              if (c16 + 16 <= ly.Npad) bias_act32_dyn(v, lb + c16, ly.act);
              else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float b = (c16 + j < ly.Npad) ? lb[c16 + j] : 0.f;
                  v[j] = act1_dyn(v[j] + b, ly.act);
                }
              }
              bool bad = false;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                v[j] = v[j] * ly.post_scale + ly.post_bias;
                bad |= (c16 + j < ly.N) && !isfinite(v[j]);
              }
              if (valid_out && bad) atomicOr(pg.nonfinite, 1);
              if (local) {
                // per-CTA latent scratch, row-interleaved (see the SRC_GLOBAL producers): 512 contiguous bytes per warp store;
                // every row is written (rows beyond n carry the finite values of x = 0), so no stale data is ever re-read
                float4* dstz = reinterpret_cast<float4*>(go + (size_t)blockIdx.x * TC_M * gs) + (size_t)(c16 >> 2) * TC_M + r;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  __stcg(dstz + (size_t)q * TC_M, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
              }
              if (local && !pg.out_dup) {
                // nothing else to store
              } else if (wide) {
                // stage the 32-column block in the group's own (free) A slot, swizzled, then store it coalesced:
                // a warp writes 4 rows x 128 B per instruction instead of 16 B into each of 32 rows
                const uint32_t stage = a_ring_s + (uint32_t)grp * C::A_SLOT;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  tc::sts128(stage + r * 128 + (((4 * half + q) ^ (r & 7)) << 4),
                             make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
              } else if (valid_out) {
                if (c16 + 16 <= ly.N && (gs & 3) == 0) {
                  float4* o = reinterpret_cast<float4*>(go + pi * gs + c16);
#pragma unroll
                  for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c16 + j < ly.N) go[pi * gs + c16 + j] = v[j];
                }
              }
            }
            if (wide && (!local || pg.out_dup)) {
              const uint32_t stage = a_ring_s + (uint32_t)grp * C::A_SLOT;
              float* gout = local ? pg.out_dup : go;        // (the caller's row-major copy of the latent / the output)
              group_bar(grp);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int f = tg + 256 * i, rr = f >> 3, ch = f & 7;
                const long long prow = tile * tile_pts + (jet ? (rr >> 2) : rr);
                if (prow < n && (!jet || (rr & 3) == 0)) {
                  const float4 val = tc::lds128(stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
                  *reinterpret_cast<float4*>(gout + (size_t)prow * gs + cb * 32 + 4 * ch) = val;
                }
              }
              group_bar(grp);
            }
          }
          tc::fence_before_sync();
          tc::mbar_arrive(&drain_done);     // the MMA thread may now reuse this TMEM region
          if (local) {
            // the heads' producers (all 512 threads) re-read this tile from the scratch: CTA-wide visibility
            __threadfence_block();
            asm volatile("bar.sync 3, %0;" ::"r"(256 * C::G) : "memory");
          } else if (wide && C::SA % C::G != 0) {
            // slots are not owned by one group: the other group's next chunk could land in this group's staging slot
            asm volatile("bar.sync 3, %0;" ::"r"(256 * C::G) : "memory");
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer (whole warp walks the loop, one elected lane issues) ===========================
    // Every operand below derives from kernel parameters, blockIdx and loop counters, so it lives in UNIFORM registers
    // and a UTCHMMA / UTCBAR is issued straight from them (see tc::mma_ss_e); the only per-thread input, the TMEM base
    // read from shared memory, is made uniform by a broadcast.
    {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a_ring_u = __shfl_sync(0xffffffffu, a_ring_s, 0);
      const uint32_t w_ring_u = a_ring_u + (uint32_t)C::SA * C::A_SLOT;
      const uint32_t bar_a_full = tc::smem_u32(&a_full[0]), bar_a_empty = tc::smem_u32(&a_empty[0]);
      const uint32_t bar_w_full = tc::smem_u32(&w_full[0]), bar_w_empty = tc::smem_u32(&w_empty[0]);
      const uint32_t bar_t_full = tc::smem_u32(&t_full[0]), bar_t_empty = tc::smem_u32(&t_empty[0]);
      uint32_t ga = 0, gs = 0, gt = 0, gw = 0, gd = 0;
      bool pending_drain = false;
      int pend_lo = 0, pend_hi = 0;         // TMEM columns of the accumulator still being drained to global
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          const TcLayer& ly = pg.layers[l];
          const uint32_t idesc = tc::make_idesc(BF16 ? tc::FMT_BF16 : tc::FMT_TF32, TC_M, ly.Npad);
          const uint32_t idesc_c = tc::make_idesc(tc::FMT_BF16, TC_M, ly.Npad);        // correction plane (kind::f16)
          const uint32_t d_tmem = tmem_u + (uint32_t)ly.tmem_col;
          const int nch = ly.seg_chunks[0] + (ly.nseg > 1 ? ly.seg_chunks[1] : 0);
          const uint32_t w_plane = (uint32_t)ly.Npad * 128;
#ifdef VQN_TC_TRACE
          long long* tr = (pg.trace && blockIdx.x == 0 && lane == 0 && tile < 4 * (long long)gridDim.x)
                              ? pg.trace + ((tile / gridDim.x) * TC_MAX_LAYERS + l) * 4 : nullptr;
#endif
          TC_STAMP(tr, 0);
          if (pending_drain) {
            // a final layer is being drained to global: wait before overwriting ITS columns (other columns may go on)
            const bool hit = (ly.tmem_col < pend_hi && ly.tmem_col + ly.Npad > pend_lo) || ly.out_slot >= 0;
            if (hit) {
              tc::mbar_wait_u(tc::smem_u32(&drain_done), gd & 1);
              ++gd; pending_drain = false;
            }
          }
          uint32_t acc = 0;
          const bool wsplit = !BF16 && ly.Npad > 128;     // the two planes of a chunk are separate ring entries
          for (int c = 0; c < nch; ++c, ++ga) {
            // A side: a shared-memory slot (descriptors) or, TS, a tensor-memory slot (column addresses)
            const bool ts = TS && ly.seg_ts[c < ly.seg_chunks[0] ? 0 : 1];
            uint32_t a_rel;                               // barrier that hands the slot back to the producers
            uint32_t a_addr;                              // smem address / TMEM column address of the slot
            if (c == 0) TC_STAMP(tr, 1);
            if (ts) {
              const uint32_t st_ = gt % TC_TS_SLOTS;
              tc::mbar_wait_u(bar_t_full + 8u * st_, (gt / TC_TS_SLOTS) & 1);
              a_rel = bar_t_empty + 8u * st_; a_addr = tmem_u + (uint32_t)TC_TS_COL + 64u * st_;
              ++gt;
            } else {
              const uint32_t sa = gs % C::SA;
              tc::mbar_wait_u(bar_a_full + 8u * sa, (gs / C::SA) & 1);
              a_rel = bar_a_empty + 8u * sa; a_addr = a_ring_u + sa * C::A_SLOT;
              ++gs;
            }
            uint32_t sw = gw % C::SW;
            tc::mbar_wait_u(bar_w_full + 8u * sw, (gw / C::SW) & 1);
            tc::fence_after_sync();
            if (c == 0) TC_STAMP(tr, 2);
            uint32_t w_addr = w_ring_u + sw * C::W_SLOT;
            if (BF16) {
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                tc::mma_ss_e<false>(d_tmem, tc::make_desc_sw128(a_addr + 32 * s), tc::make_desc_sw128(w_addr + 32 * s), idesc, acc);
                acc = 1;
              }
            } else {
              // leading products: four kind::tf32 MMAs on plane H
              if (ts) {
#pragma unroll
                for (int s = 0; s < 4; ++s) { tc::mma_ts_e<true>(d_tmem, a_addr + 8u * s, tc::make_desc_sw128(w_addr + 32 * s), idesc, acc); acc = 1; }
              } else {
#pragma unroll
                for (int s = 0; s < 4; ++s) { tc::mma_ss_e<true>(d_tmem, tc::make_desc_sw128(a_addr + 32 * s), tc::make_desc_sw128(w_addr + 32 * s), idesc, acc); acc = 1; }
              }
              if (wsplit) {
                // wide layer: plane C is the next ring entry; plane H's entry is released by its own commit
                tc::mma_commit_e(bar_w_empty + 8u * sw);
                ++gw;
                sw = gw % C::SW;
                tc::mbar_wait_u(bar_w_full + 8u * sw, (gw / C::SW) & 1);
                tc::fence_after_sync();
                w_addr = w_ring_u + sw * C::W_SLOT;
              } else {
                w_addr += w_plane;
              }
              // both correction products in ONE bf16 MMA of K = 16 per step: [a_lo | a_hi] . [w_hi ; w_lo] (plane C)
              if (ts) {
#pragma unroll
                for (int s = 0; s < 4; ++s) tc::mma_ts_e<false>(d_tmem, a_addr + 32u + 8u * s, tc::make_desc_sw128(w_addr + 32 * s), idesc_c, 1);
              } else {
#pragma unroll
                for (int s = 0; s < 4; ++s) tc::mma_ss_e<false>(d_tmem, tc::make_desc_sw128(a_addr + C::A_PLANE + 32 * s), tc::make_desc_sw128(w_addr + 32 * s), idesc_c, 1);
              }
            }
            tc::mma_commit_e(a_rel);              // slots are free once these MMAs have read them
            tc::mma_commit_e(bar_w_empty + 8u * sw);
            ++gw;
          }
          tc::mma_commit_e(tc::smem_u32(&acc_full));   // layer complete -> epilogue warps may drain it
          TC_STAMP(tr, 3);
          if (ly.out_slot >= 0) { pending_drain = true; pend_lo = ly.tmem_col; pend_hi = ly.tmem_col + ly.Npad; }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== weight producer (one thread, bulk copies) ===========================
    if (lane == 0) {
      uint32_t gw = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          const TcLayer& ly = pg.layers[l];
          const int nch = ly.seg_chunks[0] + (ly.nseg > 1 ? ly.seg_chunks[1] : 0);
          const uint32_t bytes = (uint32_t)ly.Npad * 128 * C::PLANES;        // chunk image: plane H, then plane C
          const int parts = (!BF16 && ly.Npad > 128) ? 2 : 1;                  // wide layer: one ring entry per plane
          const uint32_t part_bytes = bytes / parts;
          for (int c = 0; c < nch; ++c) {
            for (int pt = 0; pt < parts; ++pt, ++gw) {
              const int sw = gw % C::SW;
              tc::mbar_wait(&w_empty[sw], ((gw / C::SW) & 1) ^ 1);
#ifdef TC_EXPERIMENT_NO_W
              if (gw >= (uint32_t)C::SW) { tc::mbar_arrive(&w_full[sw]); continue; }
#endif
              tc::mbar_expect_tx(&w_full[sw], part_bytes);
              tc::bulk_g2s(w_ring + (size_t)sw * C::W_SLOT, ly.w + (size_t)c * bytes + (size_t)pt * part_bytes, part_bytes,
                           &w_full[sw]);
            }
          }
        }
      }
    }
    __syncwarp();
  }
#ifdef VQN_TC_TRACE
  if (ktr) ktr[2] = clock64();       // thread 0 (a producer of group 0) has finished its last drain
#endif
  tc::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tc::tmem_dealloc(tmem_base, 512);
#ifdef VQN_TC_TRACE
  if (ktr) ktr[3] = clock64();
#endif
}

// ---------------------------------------------------------------------------------------------
// host-side program assembly
// ---------------------------------------------------------------------------------------------
struct TcBuilder {
  TcProgram pg;
  bool ok;
  int E;
  TcBuilder(int precision) { memset(&pg, 0, sizeof(pg)); ok = true; E = precision == VQN_PREC_BF16 ? 64 : 32; }
};

// Append one network.  first_src: where its input x comes from (SRC_EMBED / SRC_GLOBAL / SRC_DRAIN = output of the
// previous appended layer).  out_slot: global output slot of the last layer (-1: it feeds the next network).
static bool tc_append_net(TcBuilder& B, vqn_net* net, TcPack* tp, int first_src, int out_slot, float post_scale,
                          float post_bias, bool fold_narrow_tail = false) {
  const vqn_net_desc& d = net->desc;
  const int first_layer = B.pg.n_layers;
  for (int i = 0; i < d.n_layers; ++i) {
    if (B.pg.n_layers >= TC_MAX_LAYERS) return false;
    TcLayer& ly = B.pg.layers[B.pg.n_layers];
    memset(&ly, 0, sizeof(ly));
    ly.w = tp->w[i]; ly.bias = tp->bias[i];
    ly.N = d.widths[i]; ly.Npad = tp->Npad[i]; ly.act = d.acts[i];
    ly.out_slot = -1; ly.post_scale = 1.f; ly.post_bias = 0.f;
    ly.tmem_col = (B.pg.n_layers & 1) * 256; ly.skip_w = nullptr; ly.skip_n = 0;
    ly.tail_w = nullptr; ly.tail_b = nullptr; ly.tail_n = 0; ly.tail_out_slot = -1; ly.tail_add_skip = 0;
    ly.stash_w = -1; ly.grad_stash = -1; ly.grad_tail_woff = 0; ly.egrad_col0 = -1; ly.grad_final = 0;
    ly.save = nullptr; ly.save_ld = 0;
    const bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
    const int seg0_rows = (i == 0) ? d.in_dim : d.widths[i - 1];
    ly.nseg = 1;
    ly.seg_type[0] = (i == 0) ? first_src : SRC_DRAIN;
    ly.seg_chunks[0] = vqn_round_up(seg0_rows, B.E) / B.E;
    ly.seg_first_chunk[0] = 0;
    if (ly.seg_type[0] == SRC_DRAIN && B.pg.n_layers == 0) return false;
    if (after_skip) {
      if (first_src == SRC_DRAIN) return false;   // x must be regenerable (embedding / global rows)
      const int xc = vqn_round_up(d.in_dim, B.E) / B.E;
      if (fold_narrow_tail && first_src == SRC_GLOBAL && i == d.n_layers - 1 && d.widths[i] <= 4 && i >= 2) {
        // Narrow last layer after the skip: y . W_y runs as the TAIL of the previous layer and x . W_x is accumulated
        // by the producers of the first layer (TcLayer::skip_w); no TcLayer is appended for it.
        TcLayer& l0 = B.pg.layers[first_layer];
        TcLayer& prev = B.pg.layers[B.pg.n_layers - 1];
        l0.skip_w = d.w[i] + (size_t)d.widths[i - 1] * d.widths[i];      // Keras rows: y first, then x (mlp.py:47-48)
        l0.skip_n = d.widths[i];
        prev.tail_w = d.w[i]; prev.tail_b = d.b[i]; prev.tail_n = d.widths[i]; prev.tail_act = d.acts[i];
        prev.tail_out_slot = out_slot; prev.tail_add_skip = 1;
        prev.tail_scale = post_scale; prev.tail_bias = post_bias;
        continue;
      } else {
        ly.nseg = 2;
        ly.seg_type[1] = first_src;
        ly.seg_chunks[1] = xc;
        ly.seg_first_chunk[1] = 0;
      }
    }
    if (i == d.n_layers - 1) { ly.out_slot = out_slot; ly.post_scale = post_scale; ly.post_bias = post_bias; }
    if (ly.Npad > TC_NPAD_MAX) return false;
    B.pg.n_layers++;
  }
  return true;
}

// Accumulator columns of a chain of layers (each draining the previous one) chosen so that as many layers as possible
// stay below TC_TS_COL: walk BACKWARDS from the last layer (a wide final layer may take the top of TMEM), giving every
// layer the lowest 128-aligned region that is disjoint from its consumer's.  Falls back to the ping-pong default.
static void tc_plan_columns(TcProgram& pg) {
  const int L = pg.n_layers;
  int col[TC_MAX_LAYERS];
  int next_lo = -1, next_hi = -1;
  for (int l = L - 1; l >= 0; --l) {
    const int N = pg.layers[l].Npad;
    int pick = -1;
    for (int c = 0; c + N <= 512 && pick < 0; c += 128) {
      if (l < L - 1 && c + N > TC_TS_COL) break;                        // only the last layer may reach into the slots
      if (next_lo >= 0 && c < next_hi && c + N > next_lo) continue;     // overlaps its consumer's accumulator
      pick = c;
    }
    if (l == L - 1 && N > 128) pick = 512 - N >= 0 && (512 - N) % 128 == 0 ? 512 - N : pick;   // keep [0, 384) for the others
    if (pick < 0) return;                                               // no plan: keep the defaults
    col[l] = pick; next_lo = pick; next_hi = pick + N;
  }
  for (int l = 0; l < L; ++l) pg.layers[l].tmem_col = col[l];
}

// diagnostic knob (not part of include/vqnerf_b200.h): device buffer of >= 4 * TC_MAX_LAYERS * 4 + 8 * 96 int64 (= 1088) that receives the
// MMA-thread time stamps of the next launches (NULL: off)
static long long* g_tc_trace = nullptr;
extern "C" void vqn_debug_tc_trace(long long* dev_buf) { g_tc_trace = dev_buf; }

// Which segments hand their A chunks over in tensor memory (TC_TS_COL).  A DRAIN segment, or a GLOBAL segment reading the
// per-CTA latent scratch, of a layer whose accumulator AND whose source accumulator lie below TC_TS_COL.  A layer whose
// accumulator reaches into the slot columns must (a) start with a DRAIN segment -- its first MMA is then issued after every
// earlier MMA (TS reads included) has completed -- and (b) be a final layer drained behind a CTA-wide barrier (the latent
// scratch, or a wide output: see the end of the final drain), so that no producer writes a slot while another group still
// reads those columns.  Anything else: no TS in this program.
static bool tc_plan_ts(TcProgram& pg) {
  static int env = -1;
  // opt-in (VQN_TC_TS=1): measured 2.72 ms against 2.63 ms for the shared-memory hand-over on mlp_main -- once the MMA
  // issue path was fixed the kernel turned out to be bound by the producers' instruction count, not by the smem port,
  // and two TMEM slots are a shallower ring than three shared-memory slots
  if (env < 0) { const char* e = getenv("VQN_TC_TS"); env = e ? atoi(e) : 0; }
  for (int l = 0; l < pg.n_layers; ++l) pg.layers[l].seg_ts[0] = pg.layers[l].seg_ts[1] = 0;
  if (!env) return false;
  for (int l = 0; l < pg.n_layers; ++l) {
    const TcLayer& ly = pg.layers[l];
    if (ly.tmem_col + ly.Npad <= TC_TS_COL) continue;
    const bool wide_out = ly.out_slot >= 0 && (ly.N % 32 == 0) && (pg.out_stride[ly.out_slot] % 4 == 0);
    if (!(ly.seg_type[0] == SRC_DRAIN && wide_out)) return false;
  }
  bool any = false;
  for (int l = 0; l < pg.n_layers; ++l) {
    TcLayer& ly = pg.layers[l];
    if (ly.tmem_col + ly.Npad > TC_TS_COL) continue;
    for (int sg = 0; sg < ly.nseg; ++sg) {
      const int st = ly.seg_type[sg];
      if (st == SRC_DRAIN) {
        if (l == 0) continue;
        const TcLayer& pl = pg.layers[l - 1];
        if (pl.tmem_col + pl.Npad > TC_TS_COL) continue;
        ly.seg_ts[sg] = 1; any = true;
      } else if (st == SRC_GLOBAL && pg.g_local && env != 3) {
        ly.seg_ts[sg] = 1; any = true;
      }
    }
  }
  if (env == 2) {            // experiment: the TS instantiation with every chunk in shared memory
    for (int l = 0; l < pg.n_layers; ++l) pg.layers[l].seg_ts[0] = pg.layers[l].seg_ts[1] = 0;
    return true;
  }
  if (env >= 4) {            // experiment: TS only for layers of width <= 128 (4) / only for wider layers (5)
    any = false;
    for (int l = 0; l < pg.n_layers; ++l)
      for (int sg = 0; sg < 2; ++sg) {
        if ((env == 4) != (pg.layers[l].Npad <= 128)) pg.layers[l].seg_ts[sg] = 0;
        any = any || pg.layers[l].seg_ts[sg];
      }
  }
  return any;
}

static int tc_launch(vqn_ctx* ctx, TcProgram& pg, int precision, cudaStream_t s) {
  pg.nonfinite = ctx->nonfinite_flag;
  pg.trace = g_tc_trace;
  int boff = 0;
  for (int l = 0; l < pg.n_layers; ++l) { pg.layers[l].bias_off = boff; boff += vqn_round_up(pg.layers[l].Npad, 32); }
  for (int l = 0; l < pg.n_layers; ++l) {
    if (pg.layers[l].tail_w) { pg.layers[l].tail_woff = boff; boff += 4 * pg.layers[l].N; }
    if (pg.layers[l].skip_w) { pg.layers[l].skip_woff = boff; boff += 4 * pg.g_dim; }
    if (pg.layers[l].tg_w) { pg.layers[l].tg_woff = boff; boff += 4 * (pg.layers[l].tg_rows + pg.layers[l].fin_skip_n); }
  }
  if (boff > TC_BIAS_FLOATS) { vqn_set_error("tensor-core MLP: bias table exceeds %d floats", TC_BIAS_FLOATS); return VQN_ERR_UNSUPPORTED; }
  if (pg.reverse) {
    int tw = -1;
    for (int l = 0; l < pg.n_layers; ++l) if (pg.layers[l].tail_w) tw = pg.layers[l].tail_woff;
    for (int l = 0; l < pg.n_layers; ++l)
      if (pg.layers[l].seg_type[0] == SRC_GRADINIT) {
        if (tw < 0) { vqn_set_error("tensor-core MLP: reverse mode needs a tail layer"); return VQN_ERR_UNSUPPORTED; }
        pg.layers[l].grad_tail_woff = tw;
      }
  }
  if (pg.pts && 3 + 6 * pg.n_freqs > 64) { vqn_set_error("tensor-core MLP: embedding wider than 64"); return VQN_ERR_UNSUPPORTED; }
  if (pg.jet) {
    for (int l = 0; l < pg.n_layers; ++l)
      for (int sg = 0; sg < pg.layers[l].nseg; ++sg)
        if (pg.layers[l].seg_type[sg] == SRC_GLOBAL) { vqn_set_error("tensor-core MLP: jet mode has no GLOBAL source"); return VQN_ERR_UNSUPPORTED; }
  }
  const int tile_pts = pg.jet ? TC_M / 4 : TC_M;
  long long tiles = (pg.n + tile_pts - 1) / tile_pts;
  int blocks = (int)(tiles < (long long)ctx->sm_count ? tiles : (long long)ctx->sm_count);
#define TC_LAUNCH(BF, JT, TSV)                                                                                         \
  do {                                                                                                                 \
    size_t smem = TcCfg<BF>::SMEM;                                                                                     \
    VQN_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<BF, JT, TSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mlp_tc_kernel<BF, JT, TSV><<<blocks, TcCfg<BF>::THREADS, smem, s>>>(pg);                                           \
  } while (0)
  const int mode = pg.jet ? 1 : (pg.reverse ? 2 : (pg.train ? 3 : (pg.train_bwd ? 4 : 0)));
  if (mode == 0 && precision != VQN_PREC_BF16 && tc_plan_ts(pg)) {
    TC_LAUNCH(false, 0, true);
    VQN_LAUNCHED(ctx);
    return VQN_OK;
  }
  if (precision == VQN_PREC_BF16 && mode == 4) { vqn_set_error("training backward chain: tf32x3 only"); return VQN_ERR_UNSUPPORTED; }
  if (precision == VQN_PREC_BF16) {
    if (mode == 1) TC_LAUNCH(true, 1, false); else if (mode == 2) TC_LAUNCH(true, 2, false); else if (mode == 3) TC_LAUNCH(true, 3, false); else TC_LAUNCH(true, 0, false);
  } else {
    if (mode == 1) TC_LAUNCH(false, 1, false); else if (mode == 2) TC_LAUNCH(false, 2, false); else if (mode == 3) TC_LAUNCH(false, 3, false);
    else if (mode == 4) TC_LAUNCH(false, 4, false); else TC_LAUNCH(false, 0, false);
  }
#undef TC_LAUNCH
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

#define TC_UNSUPPORTED(msg) do { vqn_set_error(msg); return VQN_ERR_UNSUPPORTED; } while (0)

int vqn_tc_net_forward(vqn_net* net, const float* x, int64_t n, float* y, int precision, cudaStream_t s) {
  if (net->in_dim % 4 != 0) TC_UNSUPPORTED("tensor-core net_forward needs in_dim % 4 == 0 (16-byte row loads)");
  TcPack* tp;
  int rc = tc_pack_get(net, precision, s, &tp);
  if (rc != VQN_OK) return rc;
  TcBuilder B(precision);
  B.pg.gsrc = x; B.pg.g_dim = net->in_dim; B.pg.n = n;
  B.pg.outs[0] = y; B.pg.out_stride[0] = net->desc.widths[net->n_layers - 1];
  if (!tc_append_net(B, net, tp, SRC_GLOBAL, 0, 1.f, 0.f)) TC_UNSUPPORTED("net_forward: network does not fit");
  return tc_launch(net->ctx, B.pg, precision, s);
}

int vqn_tc_pred_enc_at(vqn_ctx* ctx, vqn_net* fe, vqn_net* bn, int n_freqs, const float* pts, const int32_t* row_idx,
                       const int32_t* n_dev, int64_t n, float* z, int precision, cudaStream_t s) {
  TcPack *t0, *t1;
  int rc = tc_pack_get(fe, precision, s, &t0);
  if (rc != VQN_OK) return rc;
  rc = tc_pack_get(bn, precision, s, &t1);
  if (rc != VQN_OK) return rc;
  TcBuilder B(precision);
  B.pg.pts = pts; B.pg.row_idx = row_idx; B.pg.n_dev = n_dev; B.pg.n = n; B.pg.n_freqs = n_freqs;
  B.pg.outs[0] = z; B.pg.out_stride[0] = bn->desc.widths[bn->n_layers - 1];
  bool ok = tc_append_net(B, fe, t0, SRC_EMBED, -1, 1.f, 0.f) && tc_append_net(B, bn, t1, SRC_DRAIN, 0, 1.f, 0.f);
  if (!ok) TC_UNSUPPORTED("pred_enc_at: program does not fit the tensor-core kernel");
  tc_plan_columns(B.pg);
  return tc_launch(ctx, B.pg, precision, s);
}

int vqn_tc_pred_heads(vqn_ctx* ctx, vqn_net* diff, vqn_net* spec, vqn_net* rough, const float* z,
                      const int32_t* n_dev, int64_t n, float slope, float bias, float* d, float* sp, float* r,
                      int precision, cudaStream_t s) {
  vqn_net* nets[3] = {diff, spec, rough};
  float* outs[3] = {d, sp, r};
  TcBuilder B(precision);
  B.pg.gsrc = z; B.pg.n_dev = n_dev; B.pg.n = n;
  for (int h = 0; h < 3; ++h) {
    if (!nets[h]) continue;
    if (B.pg.g_dim && B.pg.g_dim != nets[h]->in_dim) TC_UNSUPPORTED("pred_heads: heads disagree on z_dim");
    B.pg.g_dim = nets[h]->in_dim;
    TcPack* tp;
    int rc = tc_pack_get(nets[h], precision, s, &tp);
    if (rc != VQN_OK) return rc;
    B.pg.outs[h] = outs[h]; B.pg.out_stride[h] = nets[h]->desc.widths[nets[h]->n_layers - 1];
    // [256, 128, out <= 4] head: layer 0 at TMEM columns [0,256), layer 1 at [256,384); the narrow last layer is
    // folded into the producers of layer 0 (x half) and the tail of layer 1 (y half)
    const vqn_net_desc& hd = nets[h]->desc;
    const bool plan = hd.n_layers == 3 && hd.skip_at == 1 && hd.widths[0] <= 256 && hd.widths[1] <= 128 &&
                      hd.widths[2] <= 4;
    const int l0 = B.pg.n_layers;
    if (!tc_append_net(B, nets[h], tp, SRC_GLOBAL, h, h == 0 ? slope : 1.f, h == 0 ? bias : 0.f, plan))
      TC_UNSUPPORTED("pred_heads: program does not fit the tensor-core kernel");
    if (plan && B.pg.n_layers - l0 == 2) { B.pg.layers[l0].tmem_col = 0; B.pg.layers[l0 + 1].tmem_col = 256; }
  }
  if (B.pg.g_dim % 4 != 0) TC_UNSUPPORTED("pred_heads: z_dim % 4 != 0");
  return tc_launch(ctx, B.pg, precision, s);
}

// encoder + bottleneck + three heads in ONE launch.  The latent tile of a CTA ([128, z_dim] fp32) goes through a per-CTA
// scratch (sm_count x 128 KB, L2-resident) between the bottleneck's final drain and the heads' layer-0 producers: it
// never travels to HBM (the two-launch form wrote and re-read 1 KB per point: 1.3 GB per 800 x 800 view) unless the
// caller asks for z (z_out != NULL: a second, streaming store).
int vqn_tc_mlp_main(vqn_ctx* ctx, vqn_net* fe, vqn_net* bn, vqn_net* diff, vqn_net* spec, vqn_net* rough, int n_freqs,
                    const float* pts, const int32_t* row_idx, const int32_t* n_dev, int64_t n, float slope, float bias,
                    float* z_out, float* d, float* sp, float* r, int precision, cudaStream_t s) {
  const int z_dim = bn->desc.widths[bn->n_layers - 1];
  vqn_net* nets[3] = {diff, spec, rough};
  float* outs[3] = {d, sp, r};
  bool fusable = z_dim % 32 == 0 && z_dim <= TC_NPAD_MAX && fe->n_layers + bn->n_layers + 6 <= TC_MAX_LAYERS;
  for (int h = 0; h < 3; ++h) {
    const vqn_net_desc& hd = nets[h]->desc;
    fusable = fusable && nets[h]->in_dim == z_dim && hd.n_layers == 3 && hd.skip_at == 1 && hd.widths[0] <= 256 &&
              hd.widths[1] <= 128 && hd.widths[2] <= 4;
  }
  if (!fusable) {                                 // generic shapes: two launches, the latent staged in z_out
    if (!z_out) TC_UNSUPPORTED("mlp_main (tensor-core modes): z_out buffer is required for this network shape");
    int rc = vqn_tc_pred_enc_at(ctx, fe, bn, n_freqs, pts, row_idx, n_dev, n, z_out, precision, s);
    if (rc != VQN_OK) return rc;
    return vqn_tc_pred_heads(ctx, diff, spec, rough, z_out, n_dev, n, slope, bias, d, sp, r, precision, s);
  }
  float* zscratch = static_cast<float*>(
      vqn_stream_scratch(ctx, VQN_SCRATCH_Z, s, sizeof(float) * (size_t)ctx->sm_count * TC_M * TC_NPAD_MAX));
  if (!zscratch) return VQN_ERR_CUDA;
  TcPack *t0, *t1;
  int rc = tc_pack_get(fe, precision, s, &t0);
  if (rc != VQN_OK) return rc;
  rc = tc_pack_get(bn, precision, s, &t1);
  if (rc != VQN_OK) return rc;
  TcBuilder B(precision);
  B.pg.pts = pts; B.pg.row_idx = row_idx; B.pg.n_dev = n_dev; B.pg.n = n; B.pg.n_freqs = n_freqs;
  B.pg.gsrc = zscratch; B.pg.g_dim = z_dim; B.pg.g_local = 1; B.pg.local_slot = 3; B.pg.out_dup = z_out;
  B.pg.outs[3] = zscratch; B.pg.out_stride[3] = z_dim;
  bool ok = tc_append_net(B, fe, t0, SRC_EMBED, -1, 1.f, 0.f) && tc_append_net(B, bn, t1, SRC_DRAIN, 3, 1.f, 0.f);
  if (ok) tc_plan_columns(B.pg);
  for (int h = 0; ok && h < 3; ++h) {
    TcPack* tp;
    rc = tc_pack_get(nets[h], precision, s, &tp);
    if (rc != VQN_OK) return rc;
    B.pg.outs[h] = outs[h]; B.pg.out_stride[h] = nets[h]->desc.widths[nets[h]->n_layers - 1];
    const int l0 = B.pg.n_layers;
    ok = tc_append_net(B, nets[h], tp, SRC_GLOBAL, h, h == 0 ? slope : 1.f, h == 0 ? bias : 0.f, true);
    if (ok && B.pg.n_layers - l0 == 2) { B.pg.layers[l0].tmem_col = 0; B.pg.layers[l0 + 1].tmem_col = 256; }
    else ok = false;
  }
  if (!ok) TC_UNSUPPORTED("mlp_main: program does not fit the tensor-core kernel");
  return tc_launch(ctx, B.pg, precision, s);
}

// SDFNetwork (geo/NeuS-ours2/models/fields.py:74-112): trunk (lin0 .. lin{L-2}, softplus) -> narrow tail = output 0 (sdf,
// on the CUDA cores while the last hidden accumulator is drained) [-> feature layer = outputs 1.. (one more MMA layer)].
// Gradient (grad_out): grad_mode 0 = jets (value, d/dx, d/dy, d/dz rows through every layer: 4 row-passes per point,
// nothing stored); grad_mode 1 = reverse mode (forward on value rows with act' stashed per CTA in an L2-resident scratch,
// then the transposed layers back to the embedding and a contraction with its Jacobian: 2 row-passes per point).
extern "C" int vqn_sdf_forward(vqn_ctx* ctx, vqn_net* trunk, const float* w_sdf, const float* b_sdf, vqn_net* feat,
                               int n_freqs, const float* pts, int64_t n, float* sdf, float* feat_out,
                               int64_t feat_stride, float* grad_out, int grad_mode, int precision, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && trunk && w_sdf && b_sdf && pts && sdf && n >= 0, "sdf_forward: null argument");
  VQN_CHECK_ARG(precision == VQN_PREC_TF32X3 || precision == VQN_PREC_BF16,
                "sdf_forward: precision must be tf32x3 or bf16 (tensor-core kernel)");
  VQN_CHECK_ARG(trunk->in_dim == 3 + 6 * n_freqs, "sdf_forward: trunk in_dim != 3 + 6*n_freqs");
  VQN_CHECK_ARG(grad_mode == 0 || grad_mode == 1, "sdf_forward: grad_mode must be 0 (jets) or 1 (reverse)");
  VQN_CHECK_ARG(!feat_out || (feat && feat->n_layers == 1 && feat->in_dim == vqn_net_out_dim(trunk) &&
                              feat_stride >= vqn_net_out_dim(feat) && feat_stride < (1 << 30)),
                "sdf_forward: feature layer / feat_stride mismatch");
  if (n == 0) return VQN_OK;
  cudaStream_t s = vqn_cs(stream);
  const bool reverse = grad_out && grad_mode == 1;
  TcPack *t0 = nullptr, *t1 = nullptr;
  int rc = tc_pack_get(trunk, precision, s, &t0);
  if (rc != VQN_OK) return rc;
  if (feat_out) { rc = tc_pack_get(feat, precision, s, &t1); if (rc != VQN_OK) return rc; }
  TcBuilder B(precision);
  B.pg.pts = pts; B.pg.n = n; B.pg.n_freqs = n_freqs;
  B.pg.jet = (grad_out && !reverse) ? 1 : 0; B.pg.jet_grad = grad_out;
  B.pg.outs[0] = feat_out; B.pg.out_stride[0] = (int)feat_stride;
  B.pg.outs[1] = sdf; B.pg.out_stride[1] = 1;
  if (!tc_append_net(B, trunk, t0, SRC_EMBED, -1, 1.f, 0.f)) TC_UNSUPPORTED("sdf_forward: trunk does not fit the tensor-core kernel");
  const int nt = trunk->n_layers;
  TcLayer& last = B.pg.layers[B.pg.n_layers - 1];
  last.tail_w = w_sdf; last.tail_b = b_sdf; last.tail_n = 1; last.tail_act = VQN_ACT_NONE; last.tail_out_slot = 1;
  last.tail_add_skip = 0; last.tail_scale = 1.f; last.tail_bias = 0.f;
  if (feat_out && !tc_append_net(B, feat, t1, SRC_DRAIN, 0, 1.f, 0.f))
    TC_UNSUPPORTED("sdf_forward: feature layer does not fit the tensor-core kernel");
  if (reverse) {
    if (nt > TC_STASH_SLOTS || B.pg.n_layers + nt > TC_MAX_LAYERS) TC_UNSUPPORTED("sdf_forward: too many layers for the reverse-mode gradient");
    rc = tc_packT_ensure(trunk, precision, s);
    if (rc != VQN_OK) return rc;
    for (int i = 0; i < nt; ++i) if (!t0->wT[i]) TC_UNSUPPORTED("sdf_forward: reverse-mode gradient needs layers with at most 256 inputs");
    float* stash = static_cast<float*>(
        vqn_stream_scratch(ctx, VQN_SCRATCH_STASH, s, sizeof(float) * (size_t)ctx->sm_count * TC_STASH_FLOATS));
    if (!stash) return VQN_ERR_CUDA;
    B.pg.reverse = 1; B.pg.stash = stash;
    const vqn_net_desc& d = trunk->desc;
    for (int i = 0; i < nt; ++i) B.pg.layers[i].stash_w = i;          // act' of every hidden layer
    for (int i = nt - 1; i >= 0; --i) {
      TcLayer& ly = B.pg.layers[B.pg.n_layers];
      memset(&ly, 0, sizeof(ly));
      const bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
      ly.w = t0->wT[i]; ly.bias = t0->biasT[i];
      ly.N = ((i == 0) ? d.in_dim : d.widths[i - 1]) + (after_skip ? d.in_dim : 0);
      ly.Npad = t0->NpadT[i]; ly.act = VQN_ACT_NONE;
      ly.nseg = 1; ly.seg_type[0] = (i == nt - 1) ? SRC_GRADINIT : SRC_GRAD;
      ly.seg_chunks[0] = t0->n_chunksT[i]; ly.seg_first_chunk[0] = 0;
      ly.out_slot = -1; ly.post_scale = 1.f; ly.post_bias = 0.f;
      ly.tmem_col = (B.pg.n_layers & 1) * 256;
      ly.tail_out_slot = -1;
      ly.stash_w = -1; ly.grad_stash = i;
      ly.egrad_col0 = after_skip ? d.widths[i - 1] : (i == 0 ? 0 : -1);
      ly.grad_final = (i == 0);
      B.pg.n_layers++;
    }
  }
  return tc_launch(ctx, B.pg, precision, s);
}

// Training forward of ONE mlp.Network (networks/mlp.py:39-50 under tf.GradientTape, train_nfr.py:562-576) as a single
// launch of the fused kernel: every layer's output is stored (y[i], leading dimension ldy[i]: the concat buffers of the
// skip connections) because the backward pass needs them; x [n, ldx] is the network input (ldx % 4 == 0, columns beyond
// in_dim zero).  The weight images are refreshed with vqn_net_repack_tc after every optimizer step.
extern "C" int vqn_net_forward_train(vqn_ctx* ctx, vqn_net* net, const float* x, int64_t ldx, int64_t n, float* const* y,
                                     const int64_t* ldy, float out_scale, float out_bias, int precision,
                                     vqn_stream stream) {
  VQN_CHECK_ARG(ctx && net && x && y && ldy && n >= 0, "net_forward_train: null argument");
  VQN_CHECK_ARG(precision == VQN_PREC_TF32X3 || precision == VQN_PREC_BF16, "net_forward_train: tensor-core precisions only");
  VQN_CHECK_ARG(ldx % 4 == 0 && ldx >= net->in_dim && ldx < (1 << 20), "net_forward_train: ldx must be a multiple of 4 >= in_dim");
  const int L = net->n_layers;
  for (int i = 0; i < L; ++i)
    VQN_CHECK_ARG(y[i] && ldy[i] >= net->desc.widths[i] && ldy[i] % 4 == 0 && ldy[i] < (1 << 20) &&
                      (reinterpret_cast<uintptr_t>(y[i]) & 15) == 0, "net_forward_train: bad activation buffer");
  if (n == 0) return VQN_OK;
  cudaStream_t s = vqn_cs(stream);
  TcPack* tp;
  int rc = tc_pack_get(net, precision, s, &tp);
  if (rc != VQN_OK) return rc;
  TcBuilder B(precision);
  B.pg.gsrc = x; B.pg.g_dim = (int)ldx; B.pg.n = n; B.pg.train = 1;
  B.pg.outs[0] = y[L - 1]; B.pg.out_stride[0] = (int)ldy[L - 1];
  if (!tc_append_net(B, net, tp, SRC_GLOBAL, 0, out_scale, out_bias, true))
    TC_UNSUPPORTED("net_forward_train: network does not fit the tensor-core kernel");
  for (int j = 0; j < B.pg.n_layers; ++j) {
    const bool is_output = (j == L - 1);                     // (absent when the narrow last layer runs as a tail)
    if (!is_output) { B.pg.layers[j].save = y[j]; B.pg.layers[j].save_ld = (int)ldy[j]; }
  }
  return tc_launch(ctx, B.pg, precision, s);
}

// vqn_net_repack_tc for several networks in ONE launch (packs that do not exist yet are built first, the usual way)
extern "C" int vqn_nets_repack_tc(vqn_net* const* nets, int count, int precision, vqn_stream stream) {
  VQN_CHECK_ARG(nets && count >= 1 && (precision == VQN_PREC_TF32X3 || precision == VQN_PREC_BF16), "nets_repack_tc args");
  const int p = precision == VQN_PREC_BF16 ? 1 : 0;
  const int E = p == 1 ? 64 : 32;
  cudaStream_t s = vqn_cs(stream);
  TcPackBatch batch;
  batch.count = 0;
  vqn_ctx* ctx = nullptr;
  auto flush = [&]() -> int {
    if (batch.count == 0) return VQN_OK;
    if (p == 1) tc_pack_batched_kernel<true><<<batch.count * TC_PACK_BLOCKS, 256, 0, s>>>(batch);
    else tc_pack_batched_kernel<false><<<batch.count * TC_PACK_BLOCKS, 256, 0, s>>>(batch);
    VQN_CUDA(cudaGetLastError());
    ctx->launches.fetch_add(1);
    batch.count = 0;
    return VQN_OK;
  };
  for (int q = 0; q < count; ++q) {
    vqn_net* net = nets[q];
    VQN_CHECK_ARG(net, "nets_repack_tc: null network");
    ctx = net->ctx;
    if (!net->tc_pack[p]) { TcPack* tp; int rc = tc_pack_get(net, precision, s, &tp); if (rc != VQN_OK) return rc; continue; }
    TcPack* tp = net->tc_pack[p];
    const vqn_net_desc& d = net->desc;
    for (int i = 0; i < d.n_layers; ++i) {
      const bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
      const int seg0_rows = (i == 0) ? d.in_dim : d.widths[i - 1];
      const int seg1_rows = after_skip ? d.in_dim : 0;
      for (int t = 0; t < ((tp->has_T && tp->wT[i]) ? 2 : 1); ++t) {
        if (batch.count == TC_PACK_JOBS) { int rc = flush(); if (rc != VQN_OK) return rc; }
        TcPackJob& j = batch.j[batch.count++];
        if (t == 0) {
          j = {d.w[i], d.b[i], tp->w[i], tp->bias[i], d.widths[i], tp->Npad[i], seg0_rows, vqn_round_up(seg0_rows, E) / E,
               seg1_rows, vqn_round_up(seg1_rows, E) / E, 0, 0};
        } else {
          j = {d.w[i], nullptr, tp->wT[i], tp->biasT[i], seg0_rows + seg1_rows, tp->NpadT[i], d.widths[i],
               vqn_round_up(d.widths[i], E) / E, 0, 0, d.widths[i], 0};
        }
      }
    }
  }
  return flush();
}

// refresh (or build) only the tensor-core weight images of `precision` from the caller's current weights
extern "C" int vqn_net_repack_tc(vqn_net* net, int precision, vqn_stream stream) {
  VQN_CHECK_ARG(net && (precision == VQN_PREC_TF32X3 || precision == VQN_PREC_BF16), "net_repack_tc args");
  const int p = precision == VQN_PREC_BF16 ? 1 : 0;
  if (!net->tc_pack[p]) { TcPack* tp; return tc_pack_get(net, precision, vqn_cs(stream), &tp); }   // builds and fills
  return tc_pack_fill(net, p, vqn_cs(stream));
}

// Training backward of ONE mlp.Network as a single launch (the backward-data half of tf.GradientTape over networks/mlp.py:39-50):
// given dz_last = d loss / d (pre-activation of the last layer) [n, lddz_last] (vqn_act_backward) and the outputs y[i] saved by
// vqn_net_forward_train, the chain  dz_{i-1} = (dz_i . W_i^T) * act'(y_{i-1})  runs through the transposed weight images of the
// fused kernel; every dz_i is stored to dz[i] (ld lddz[i]; the weight gradients read them) and the gradient w.r.t. the network
// input is written to d_input (ld ld_din; din_mode 0 store, 1 add, 2 atomic add; NULL: not needed -- then the chain stops at
// dz[0]).  Supported shapes: plain chains, a skip concat whose x half needs no gradient (d_input == NULL), and the head shape
// [w0, w1, out <= 3] with skip_at == 1, whose narrow last layer is evaluated backwards on the CUDA cores (its y half feeds dz[1],
// its x half is added to d_input in the final drain).
extern "C" int vqn_net_backward_train(vqn_ctx* ctx, vqn_net* net, const float* dz_last, int64_t lddz_last, int64_t n,
                                      const float* const* y, const int64_t* ldy, float* const* dz, const int64_t* lddz,
                                      float* d_input, int64_t ld_din, int din_mode, const float* din_y,
                                      int64_t ld_din_y, int din_act, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && net && dz_last && y && ldy && dz && lddz && n >= 0, "net_backward_train: null argument");
  VQN_CHECK_ARG(lddz_last % 4 == 0 && lddz_last < (1 << 20), "net_backward_train: lddz_last must be a multiple of 4");
  VQN_CHECK_ARG(din_mode >= 0 && din_mode <= 2, "net_backward_train: din_mode");
  if (n == 0) return VQN_OK;
  cudaStream_t s = vqn_cs(stream);
  const vqn_net_desc& d = net->desc;
  const int L = d.n_layers;
  TcPack* tp;
  int rc = tc_pack_get(net, VQN_PREC_TF32X3, s, &tp);
  if (rc != VQN_OK) return rc;
  rc = tc_packT_ensure(net, VQN_PREC_TF32X3, s);
  if (rc != VQN_OK) return rc;
  const bool head = L == 3 && d.skip_at == 1 && d.widths[2] <= 3;
  if (d.skip_at >= 0 && !head && d_input) TC_UNSUPPORTED("net_backward_train: input gradient through a skip concat (general shape)");
  if (head && !d_input) TC_UNSUPPORTED("net_backward_train: the head shape needs d_input");
  TcBuilder B(VQN_PREC_TF32X3);
  B.pg.n = n; B.pg.train_bwd = 1;
  // MMA layers: forward layer i backwards, i = top .. bottom, where bottom = 1 when no input gradient is needed
  const int top = head ? 1 : L - 1, bottom = d_input ? 0 : 1;
  if (top < bottom) TC_UNSUPPORTED("net_backward_train: nothing to do on the tensor cores for this shape");
  if (!head) { B.pg.gsrc = dz_last; B.pg.g_dim = (int)lddz_last; }
  for (int i = top; i >= bottom; --i) {
    if (B.pg.n_layers >= TC_MAX_LAYERS || !tp->wT[i]) TC_UNSUPPORTED("net_backward_train: layer does not fit the tensor-core kernel");
    TcLayer& ly = B.pg.layers[B.pg.n_layers];
    memset(&ly, 0, sizeof(ly));
    const bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
    ly.w = tp->wT[i]; ly.bias = tp->biasT[i];
    ly.N = ((i == 0) ? d.in_dim : d.widths[i - 1]) + (after_skip ? d.in_dim : 0);
    ly.Npad = tp->NpadT[i]; ly.act = VQN_ACT_NONE;
    ly.nseg = 1; ly.seg_chunks[0] = tp->n_chunksT[i]; ly.seg_first_chunk[0] = 0;
    ly.out_slot = -1; ly.post_scale = 1.f; ly.post_bias = 0.f;
    ly.tmem_col = (B.pg.n_layers & 1) * 256;
    ly.tail_out_slot = -1; ly.stash_w = -1; ly.grad_stash = -1; ly.egrad_col0 = -1;
    if (i == top && !head) ly.seg_type[0] = SRC_GLOBAL;          // dz of the last layer: the caller's buffer
    else {
      ly.seg_type[0] = (i == top && head) ? SRC_TAILGRAD : SRC_BWD;
      ly.bw_y = y[i]; ly.bw_y_ld = (int)ldy[i]; ly.bw_act = d.acts[i];
      ly.bw_dz = dz[i]; ly.bw_dz_ld = (int)lddz[i]; ly.bw_n = d.widths[i];
      VQN_CHECK_ARG(y[i] && dz[i] && ldy[i] % 4 == 0 && lddz[i] % 4 == 0 && ldy[i] >= d.widths[i] && lddz[i] >= d.widths[i],
                    "net_backward_train: activation / dz buffer");
    }
    if (head) {                                                 // both MMA layers may need the narrow layer's table / dz
      ly.tg_dz = dz_last; ly.tg_dz_ld = (int)lddz_last; ly.tg_n = d.widths[2]; ly.tg_w = d.w[2];
      ly.tg_rows = d.widths[1]; ly.fin_skip_row0 = d.widths[1];
      if (i != top) { ly.tg_rows = 0; ly.fin_skip_n = d.in_dim; }    // the last MMA layer: only the x half, for the final drain
    }
    B.pg.n_layers++;
  }
  TcLayer& last = B.pg.layers[B.pg.n_layers - 1];
  last.out_slot = 0;
  if (d_input) {
    VQN_CHECK_ARG(ld_din >= d.in_dim && ld_din < (1 << 20), "net_backward_train: ld_din");
    B.pg.outs[0] = d_input; B.pg.out_stride[0] = (int)ld_din; last.fin_mode = din_mode;
    last.N = d.in_dim;
    if (din_y) {                     // the input is another layer's activated output: end in that layer's dz
      if (head || d.skip_at >= 0) TC_UNSUPPORTED("net_backward_train: din_y with a skip concat");
      VQN_CHECK_ARG(ld_din_y >= d.in_dim && ld_din_y % 4 == 0 && ld_din_y < (1 << 20), "net_backward_train: ld_din_y");
      last.fin_y = din_y; last.fin_y_ld = (int)ld_din_y; last.fin_act = din_act;
    }
  } else {
    // the chain ends in dz[0] = (gradient w.r.t. y[0]) * act'(y[0])
    VQN_CHECK_ARG(y[0] && dz[0], "net_backward_train: y[0] / dz[0]");
    B.pg.outs[0] = dz[0]; B.pg.out_stride[0] = (int)lddz[0]; last.fin_mode = 0;
    last.fin_y = y[0]; last.fin_y_ld = (int)ldy[0]; last.fin_act = d.acts[0];
    last.N = d.widths[0];
  }
  return tc_launch(ctx, B.pg, VQN_PREC_TF32X3, s);
}
