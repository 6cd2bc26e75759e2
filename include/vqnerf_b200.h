/*
 * vqnerf_b200.h -- C ABI of libvqnerf_b200.so
 *
 * Drop-in boundary for the VQ-NeRF decomposition-stage per-surface-point shading path
 * (reference: decomp/nerfvq_nfr3, NeRFactor-derived) and the NeuS geo-stage per-ray scan
 * (reference: geo/NeuS-ours2/models/renderer.py), implemented as hand-written sm_100a CUDA.
 *
 * The reference has no FFI layer: its "operator API" is the Python method surface of
 * nerfactor/models/vq_nfr.py::Model and nerfactor/networks/*.  Every entry point below
 * cites the reference method (file:line, relative to /root/reference/) it replaces; the
 * Python mirror in vqnerf_release_b200/nerfactor/ binds them with ctypes (INTEGRATION.md
 * shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; all tensor pointers are DEVICE pointers to row-major contiguous
 *     fp32 unless stated; indices are int64 (TF argmax default);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous w.r.t. the host
 *     and never synchronises (the reference's .numpy() sync at vq_layers.py:318 is NOT
 *     reproduced);
 *   - every function returns a vqn_status; nothing aborts; vqn_status_str() names a code and
 *     vqn_last_error() returns the thread-local detail string of the last failure;
 *   - no hidden global state: learned state (weights, EMA hidden/average/counter, codebook)
 *     lives in caller-owned buffers so the caller's checkpointing keeps working.
 */
#ifndef VQNERF_B200_H_
#define VQNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQN_ABI_VERSION 1
#define VQN_MAX_LAYERS 8
#define VQN_NUM_LIGHTS 512 /* 16 x 32 probe, vq_nfr.ini:51 */

typedef enum vqn_status {
  VQN_OK = 0,
  VQN_ERR_INVALID_ARG = 1,   /* bad shape / null pointer / unsupported size  (Python: ValueError)         */
  VQN_ERR_CUDA = 2,          /* CUDA runtime error (detail in vqn_last_error)  (Python: RuntimeError)      */
  VQN_ERR_NONFINITE = 3,     /* check_numerics failure (vq_nfr.py:783,802,815,827,731) (InvalidArgument)  */
  VQN_ERR_UNSUPPORTED = 4,   /* feature not built / wrong GPU architecture                                */
  VQN_ERR_ZERO_NORM = 5      /* assert_greater on direction norms (shape.py:107-109,116-118)              */
} vqn_status;

typedef enum vqn_act {
  VQN_ACT_NONE = 0, VQN_ACT_RELU = 1, VQN_ACT_SIGMOID = 2,
  VQN_ACT_SOFTPLUS100 = 3 /* torch.nn.Softplus(beta=100), geo/NeuS-ours2/models/fields.py:72 (tensor-core modes only) */
} vqn_act;

/* arithmetic mode of the MLP kernels */
typedef enum vqn_precision {
  VQN_PREC_FP32 = 0,   /* fp32 FFMA on CUDA cores; parity mode (1e-4 rel)                     */
  VQN_PREC_BF16 = 1,   /* tcgen05 kind::f16 bf16 operands, fp32 accumulate in TMEM (1e-2 rel) */
  VQN_PREC_TF32X3 = 2  /* tcgen05 kind::tf32 hi/lo split, fp32 accumulate (1e-4 rel)          */
} vqn_precision;

typedef struct vqn_ctx vqn_ctx; /* one per device; re-entrant across streams */
typedef struct vqn_net vqn_net; /* device-side packed copy of one mlp.Network */
typedef void* vqn_stream;       /* cudaStream_t */

int vqn_abi_version(void);
const char* vqn_status_str(int status);
const char* vqn_last_error(void);

int vqn_ctx_create(int device, vqn_ctx** out);
int vqn_ctx_destroy(vqn_ctx* ctx);
/* number of kernels this library has launched on this ctx since creation (bench.py gpu_launches) */
int64_t vqn_ctx_launch_count(const vqn_ctx* ctx);
/* check_numerics flag: device-side sticky flag set by kernels that saw NaN/Inf in an output that the
 * reference guards with tf.debugging.check_numerics; read (with a sync of `stream`) and cleared here. */
int vqn_ctx_check_numerics(vqn_ctx* ctx, vqn_stream stream);

/* ---- workspace / buffer size queries (element counts) ---------------------------------------------------------- */
/* float64 entries of the VQ statistics vector: K counts + sum((q-x)^2) + rows + dw[Z,K] */
int64_t vqn_vq_stats_size(int z_dim, int k);
/* int32 entries of the compaction workspace of vqn_compact_mask for n rows */
int64_t vqn_compact_workspace_size(int64_t n);
/* entries of flag_ws (float) = nb_ws = valid_ws (int32) of vqn_sample_pairs for an H x W view; its compact_ws needs
 * vqn_compact_workspace_size of that count */
int64_t vqn_sample_pairs_workspace_size(int h, int w);

/* ---- light probe geometry: brdf/renderer.py:184-219 gen_light_xyz (host, float64) ------------- */
int vqn_gen_light_xyz(int envmap_h, int envmap_w, double envmap_radius, double* xyz_out /*[h,w,3]*/,
                      double* areas_out /*[h,w]*/);

/* ---- networks -------------------------------------------------------------------------------- */
/* nerfactor/networks/mlp.py:24-50 Network(widths, act, skip_at): Keras Dense kernels [in,out], bias
 * [out]; after layer `skip_at` the activation becomes concat(y, x_input).  skip_at = -1: none. */
typedef struct vqn_net_desc {
  int32_t n_layers;
  int32_t in_dim;
  int32_t skip_at;
  int32_t reserved;
  int32_t widths[VQN_MAX_LAYERS];
  int32_t acts[VQN_MAX_LAYERS];
  const float* w[VQN_MAX_LAYERS]; /* device, [in_i, widths[i]] */
  const float* b[VQN_MAX_LAYERS]; /* device, [widths[i]]       */
} vqn_net_desc;

/* Pack (pad / transpose / split for the tensor-core modes) the weights into library-owned device
 * memory.  vqn_net_repack re-reads the caller's weight buffers after an optimizer step. */
int vqn_net_create(vqn_ctx* ctx, const vqn_net_desc* desc, vqn_net** out, vqn_stream stream);
int vqn_net_repack(vqn_net* net, const vqn_net_desc* desc, vqn_stream stream);
int vqn_net_destroy(vqn_net* net);
int vqn_net_out_dim(const vqn_net* net);

/* mlp.Network.__call__ (mlp.py:39-50): y[n,out] = net(x[n,in]) */
int vqn_net_forward(vqn_net* net, const float* x, int64_t n, float* y, int precision, vqn_stream stream);

/* Embedder.__call__ (networks/embedder.py:23-47, kwargs of models/shape.py:82-89):
 * out[n, 3 + 6*n_freqs] = [x, sin(x 2^0), cos(x 2^0), ..., sin(x 2^(F-1)), cos(x 2^(F-1))] */
int vqn_embed(vqn_ctx* ctx, const float* x, int64_t n, int n_freqs, float* out, vqn_stream stream);
/* ... into rows of leading dimension ld_out >= 3 + 6 n_freqs (columns beyond the embedding are left untouched) */
int vqn_embed_ld(vqn_ctx* ctx, const float* x, int64_t n, int n_freqs, float* out, int64_t ld_out, vqn_stream stream);

/* Model._pred_enc_at (models/vq_nfr.py:771-784): z[n,256] = bottleneck(fine_enc(embed(pts))).
 * row_idx (optional, int32 [n]) gathers pts rows (mask compaction, vq_nfr.py:283-291); z is compact.
 * n_dev (optional device int32 scalar) overrides n (<= n) so that no host sync follows the compaction. */
int vqn_pred_enc_at(vqn_ctx* ctx, vqn_net* fine_enc, vqn_net* bottleneck, int n_freqs, const float* pts,
                    const int32_t* row_idx, const int32_t* n_dev, int64_t n, float* z_out, int precision,
                    vqn_stream stream);

/* Model._pred_diff_at / _pred_spec_at / _pred_rough_at (models/vq_nfr.py:786-828) fused over one read
 * of z[n,256]; any head may be NULL.  diff_out = slope*sigmoid(.)+bias (albedo_slope/bias). */
int vqn_pred_heads(vqn_ctx* ctx, vqn_net* diff, vqn_net* spec, vqn_net* rough, const float* z,
                   const int32_t* n_dev, int64_t n, float albedo_slope, float albedo_bias, float* diff_out /*[n,3]*/,
                   float* spec_out /*[n,out_dim(spec)]*/, float* rough_out /*[n,1]*/, int precision,
                   vqn_stream stream);

/* _pred_enc_at + _pred_diff_at + _pred_spec_at + _pred_rough_at of the main branch in one launch, as
 * fast_render chains them (models/vq_nfr.py:321,329-331): the latent stays on chip; z_out [n,256] is optional. */
int vqn_mlp_main(vqn_ctx* ctx, vqn_net* fine_enc, vqn_net* bottleneck, vqn_net* diff, vqn_net* spec, vqn_net* rough,
                 int n_freqs, const float* pts, const int32_t* row_idx, const int32_t* n_dev, int64_t n,
                 float albedo_slope, float albedo_bias, float* z_out, float* diff_out, float* spec_out,
                 float* rough_out, int precision, vqn_stream stream);

/* ---- vector quantiser ------------------------------------------------------------------------ */
/* Model.get_codebook (models/vq_nfr.py:761-769): out[Z,K] = l2_normalize(clip(raw,0,1), axis=0) */
int vqn_get_codebook(vqn_ctx* ctx, const float* raw, int z_dim, int k, float* out, vqn_stream stream);

/* mathutil.safe_l2_normalize(x, axis=1) (util/math.py:63-64) on rows of x[n,d] */
int vqn_l2_normalize_rows(vqn_ctx* ctx, const float* x, int64_t n, int d, float* out, vqn_stream stream);

/* VectorQuantizerEMA.__call__ forward half (networks/vq_layers.py:277-302,327-330).
 *   inputs   [n,Z] (Z == 256, base 32-byte aligned)   codebook [Z,K] (already normalised, 1 <= K <= 1024)
 *   sel_mask [K] or NULL: 1 = codeword selectable, 0 = dropped (the `roll >= thres` mask, :284-290);
 *   normalize_inputs != 0 fuses the caller's l2_normalize(z_enc, axis=1) (vq_nfr.py:575).
 * Outputs (each may be NULL): indices int64 [n] (0-based; callers add 1, vq_nfr.py:578), quantize [n,Z]
 * (= x + (q - x), the STE forward value, :327), distances [n,K], z_norm_out [n,Z],
 * stats (float64 [K+2+Z*K]): [0,K) one-hot counts, [K] sum((q-x)^2), [K+1] rows seen,
 * [K+2,..) dw[Z,K] = x^T one_hot (:308); stats are ACCUMULATED (zero them first) so that the
 * multi-GPU path can all-reduce them before vqn_vq_ema_update. */
int vqn_vq_assign(vqn_ctx* ctx, const float* inputs, int64_t n, int z_dim, const float* codebook, int k,
                  const float* sel_mask, int normalize_inputs, int64_t* indices, float* quantize,
                  float* distances, float* z_norm_out, double* stats, int want_dw, vqn_stream stream);

/* EMA half (vq_layers.py:304-321 + dm-sonnet 2.0.0 moving_averages.ExponentialMovingAverage).
 * state layout (caller-owned, float32 unless noted):
 *   cs_hidden[K], cs_average[K], dw_hidden[Z,K], dw_average[Z,K], counter (int64[2]: cs, dw)
 * Reads `stats` as written by vqn_vq_assign (after any cross-GPU all-reduce), writes
 * update[Z,K] (= 'update' of the returned dict), loss[1] = commitment*e_latent, perplexity[1]. */
int vqn_vq_ema_update(vqn_ctx* ctx, const double* stats, int z_dim, int k, const float* codebook,
                      float decay, float epsilon, float commitment_cost, int is_training,
                      float* cs_hidden, float* cs_average, float* dw_hidden, float* dw_average,
                      int64_t* counters, float* update, float* loss, float* perplexity,
                      vqn_stream stream);

/* ---- shading ---------------------------------------------------------------------------------- */
enum { VQN_LVIS_F32 = 0, VQN_LVIS_F16 = 1, VQN_LVIS_U8 = 2 };
typedef struct vqn_shade_args {
  /* per point, compacted rows i in [0,n); when row_idx != NULL the geometry inputs and outputs are
   * full-length arrays addressed through row_idx[i] (mask compaction/expansion, vq_nfr.py:283-291,
   * 347-370) while albedo/spec/rough are compact [n,*]. */
  const float* xyz;     /* [*,3] surface points                                   */
  const float* rayo;    /* [*,3] camera locations (_calc_vdir, shape.py:112-119)  */
  const float* normal;  /* [*,3] (un-corrected; _normal_correct is fused)         */
  const float* lvis;    /* [*,512] or NULL (non-'nerf' data, vq_nfr.py:707); element type: lvis_format */
  const float* albedo;  /* [n,3] */
  const float* spec;    /* [n,3] (f0) */
  const float* rough;   /* [n,1] */
  const int32_t* row_idx; /* [n] or NULL */
  const int32_t* n_dev;   /* optional device scalar overriding n (no host sync after compaction) */
  int64_t n;
  /* probes */
  const float* lxyz;    /* [512,3] fp32 light positions (gen_light_xyz)           */
  const float* lareas;  /* [512]                                                  */
  const float* lights;  /* [1+P,512,3]: probe 0 = model light (or novel_probes[dst_env])         */
  int32_t n_probes;     /* 1 + P                                                  */
  int32_t clip_light0;  /* apply clip(.,0,inf) to probe 0 (the `light` property, vq_nfr.py:759) */
  int32_t to_srgb;      /* data_type == 'nerf': linear2srgb on outputs (vq_nfr.py:350-356) */
  int32_t use_gamma;    /* data_type != 'nerf': rgb = (rgb*gamma_bias)^gamma_index (:715-716) */
  float gamma_bias, gamma_index;
  /* outputs (may be NULL) */
  float* rgb;           /* [*, n_probes, 3]  clip(.,0,1) (+sRGB)                  */
  float* rgb_diff;      /* [*,3] probe 0 only, diffuse lobe (vq_nfr.py:605-610)   */
  float* rgb_spec;      /* [*,3] probe 0 only, glossy lobe                        */
  float* normal_out;    /* [*,3] corrected normal                                 */
  /* Fused image gather (multi-GPU relighting): when n_peers > 0 the kernel also stores every shaded row
   * into the n_peers image buffers peer_rgb[p] ([n_global, n_probes, 3], P2P-mapped device pointers of ALL ranks
   * incl. this one, e.g. torch symmetric memory) at global row peer_row0 + row, so that the single gather of the
   * pixel-sharded render travels over NVLink from inside the shading kernel instead of a separate collective.
   * The caller synchronises the ranks afterwards (a barrier on the stream).  Large un-split batches only. */
  float* peer_rgb[8];
  int32_t n_peers;
  /* element type of lvis: VQN_LVIS_F32 (0, the reference's lvis.npy), VQN_LVIS_F16 or VQN_LVIS_U8 (v = q / 255) -- compact
   * opt-in formats of the host-buffer path (the fp32 rows are 98 % of a view's H2D bytes); large un-split batches only */
  int32_t lvis_format;
  int64_t peer_row0;
  /* 1: rgb is the raw, un-clipped integral (no gamma, no clip, no sRGB): the training step of non-'nerf' data applies
   * gamma with device-resident parameters (vqn_gamma_forward) so that a captured graph follows them; warp-per-point
   * kernel only (small batches) */
  int32_t no_clip;
  int32_t reserved0;
} vqn_shade_args;

/* _calc_ldir + _calc_vdir + _normal_correct + _eval_brdf_at (util/microfacet.py:9-89) + _render
 * (models/vq_nfr.py:694-733) fused; nothing of shape [N,512,3] is materialised. */
int vqn_shade(vqn_ctx* ctx, const vqn_shade_args* args, vqn_stream stream);

/* Fine-grained, materialising variants kept for API parity (debug sizes only):
 * _eval_brdf_at -> brdf, brdf_glossy, brdf_diffuse [n,512,3] from explicit directions. */
int vqn_eval_brdf(vqn_ctx* ctx, const float* pts2l /*[n,512,3]*/, const float* pts2c /*[n,3]*/,
                  const float* normal, const float* albedo, const float* spec, const float* rough,
                  int64_t n, float* brdf, float* brdf_glossy, float* brdf_diffuse, vqn_stream stream);
/* _render(brdf, l, n, light_vis) for one probe: rgb[n,3] (linear, clipped) */
int vqn_render(vqn_ctx* ctx, const float* brdf, const float* l, const float* normal, const float* lvis,
               const float* lareas, const float* light /*[512,3]*/, int64_t n, int use_gamma,
               float gamma_bias, float gamma_index, float* rgb, vqn_stream stream);

/* spec = ks*basecolor, albedo = (1-ks)*basecolor (models/vq_nfr.py:330-331,590-591) and fast_render's
 * opt_scale (:333-336, opt_scale[3] device or NULL); all [n,3] compact, ks [n,1]; outputs may be NULL. */
int vqn_material_combine(vqn_ctx* ctx, const float* basecolor, const float* ks, const float* opt_scale,
                         const int32_t* n_dev, int64_t n, float* albedo, float* spec, float* albedo_scaled,
                         float* spec_scaled, vqn_stream stream);

/* Fused image gather, background rows: vqn_shade stores only the live rows of a shard into the peers' image buffers
 * (vqn_shade_args.peer_rgb); this stores zeros for the shard's rows with alpha[row * alpha_stride] <= 0 at global row
 * peer_row0 + row ([.., width] floats per row) in each of the n_peers buffers (HOST array of P2P-mapped device
 * pointers), as scatter_nd leaves them (models/vq_nfr.py:347-370). */
int vqn_peer_clear_background(vqn_ctx* ctx, const float* alpha, int alpha_stride, int64_t n_local, int64_t peer_row0,
                              int width, float* const* peer_ptrs, int n_peers, vqn_stream stream);

/* Frame hand-shake of the fused gather on ONE destination rank (instead of a barrier over all ranks): a sender publishes
 * "frame f finished" into the destination's flags and waits until frame f-1 was consumed; the destination acknowledges
 * frame f-1 and waits for every sender's frame f.  counter: this rank's frame count (device int32, starts at 0);
 * local_flags / peer_flags[r]: world+1 int32 of symmetric memory on this rank / on rank r (HOST array of P2P pointers). */
int vqn_peer_frame_sync(vqn_ctx* ctx, int32_t* counter, const int32_t* local_flags, int32_t* const* peer_flags, int world,
                        int rank, int dst, vqn_stream stream);

/* fast_render(edit_mask=, edit_material=) (models/vq_nfr.py:258-260 `_update_material`, :293-295, :324-330): compact rows
 * whose edit_mask[row_idx[i] * mask_stride] > 0 get albedo := diff3, spec := spec3, rough := rough1 (HOST pointers; NULL =
 * the reference's "update[0] < 0: leave alone"), in place; the opt_scale'd copies the shading kernel reads are refreshed
 * (:333-336).  edit_mask is the full-length [n_total, mask_stride] tensor of the reference's signature. */
int vqn_material_edit(vqn_ctx* ctx, const float* edit_mask, int mask_stride, const int32_t* row_idx,
                      const int32_t* n_dev, int64_t n, const float* diff3, const float* spec3, const float* rough1,
                      const float* opt_scale, float* albedo, float* spec, float* rough, float* albedo_scaled,
                      float* spec_scaled, vqn_stream stream);

/* util/img.py:142-186 */
int vqn_linear2srgb(vqn_ctx* ctx, const float* x, int64_t count, float* out, vqn_stream stream);
int vqn_srgb2linear(vqn_ctx* ctx, const float* x, int64_t count, float* out, vqn_stream stream);

/* mask = alpha[:,0] > 0 ; ind = where(mask) (vq_nfr.py:283,345): row_idx[int32, n_total] and the
 * device count n_active[1]; no host synchronisation. */
int vqn_compact_mask(vqn_ctx* ctx, const float* alpha, int64_t n_total, int32_t* row_idx,
                     int32_t* n_active, int32_t* workspace /* >= n_total/1024 + 1 ints, or NULL: ctx scratch
                     (then only one compaction may be in flight per ctx) */, vqn_stream stream);

/* scatter_nd(ind, value, (n_total,c)) (vq_nfr.py:347-370): out must be pre-zeroed full length */
int vqn_scatter_rows(vqn_ctx* ctx, const float* compact, const int32_t* row_idx, const int32_t* n_dev,
                     int64_t n_max, int c, float* out, vqn_stream stream);
/* up to 4 scatters that share row_idx / n_dev in ONE launch: outs[q][row_idx[r], :widths[q]] = compact[q][r, :] (HOST arrays) */
int vqn_scatter_rows_multi(vqn_ctx* ctx, const float* const* compact, const int32_t* widths, int count,
                           const int32_t* row_idx, const int32_t* n_dev, int64_t n_max, float* const* outs, vqn_stream s);

/* ---- training step (BASELINE config #4): Model.call(mode='train') + compute_loss under tf.GradientTape and
 * the Adam(amsgrad) update (models/vq_nfr.py:534-692, 876-986; train_nfr.py:121-139, 562-576) -------------- */
/* Keras Dense forward with the output kept for the backward pass (networks/mlp.py:44-46):
 * Y[m,n] (leading dim ldy) = out_scale * act(X[m,k] (leading dim ldx) . W[k,n] + b[n]) + out_bias.  Leading
 * dimensions let a layer read/write a column slice of the concat buffer of a skip connection (mlp.py:47-48). */
int vqn_dense_forward(vqn_ctx* ctx, const float* x, int64_t ldx, const float* w, const float* b, float* y,
                      int64_t ldy, int64_t m, int k, int n, int act, float out_scale, float out_bias,
                      vqn_stream stream);
/* The same forward for a WHOLE network in one launch of the fused tensor-core kernel (precision tf32x3 / bf16): x [n, ldx]
 * (ldx % 4 == 0, columns >= in_dim zero), y[i] / ldy[i] = output buffer and leading dimension of layer i (the concat
 * buffers of mlp.py:47-48; the caller copies the x half of a concat).  vqn_net_repack_tc refreshes the pre-split weight
 * images of one precision from the caller's current weights (after every optimizer step). */
int vqn_net_forward_train(vqn_ctx* ctx, vqn_net* net, const float* x, int64_t ldx, int64_t n, float* const* y,
                          const int64_t* ldy, float out_scale, float out_bias, int precision, vqn_stream stream);
int vqn_net_repack_tc(vqn_net* net, int precision, vqn_stream stream);
/* Backward-data of ONE network as a single launch of the fused tensor-core kernel (tf32x3): dz_last = d loss / d
 * (pre-activation of the last layer) [n, lddz_last]; y[i] / ldy[i] = the outputs saved by vqn_net_forward_train; every
 * dz_i = (dz_{i+1} . W_{i+1}^T) * act'(y_i) is stored to dz[i] (ld lddz[i]) for the weight gradients; the gradient w.r.t. the
 * network input goes to d_input (ld ld_din; din_mode 0 store / 1 add / 2 atomic add) or, with d_input == NULL, the chain
 * ends in dz[0].  Shapes: plain chains; a skip concat whose x half needs no gradient; the head shape [w0, w1, out <= 3] with
 * skip_at == 1 (d_input required).  din_y (optional; plain chains): the network's input is the activated output din_y
 * [n, ld_din_y] of another layer (activation din_act) -- d_input is multiplied by act'(din_y) and is then that layer's dz.
 * Weight images are refreshed by vqn_net_repack_tc / vqn_nets_repack_tc. */
int vqn_net_backward_train(vqn_ctx* ctx, vqn_net* net, const float* dz_last, int64_t lddz_last, int64_t n,
                           const float* const* y, const int64_t* ldy, float* const* dz, const int64_t* lddz,
                           float* d_input, int64_t ld_din, int din_mode, const float* din_y, int64_t ld_din_y,
                           int din_act, vqn_stream stream);
/* the same for `count` networks in ONE launch (the training step refreshes all of its networks after the optimizer step) */
int vqn_nets_repack_tc(vqn_net* const* nets, int count, int precision, vqn_stream stream);
/* dX[m,k] (+)= (dZ[m,n] . W[k,n]^T) * act_prev'(Yprev[m,k]); act_prev' is taken from the stored activation
 * (relu: y > 0, sigmoid: y (1 - y)); accumulate != 0 adds into dX (several consumers of one tensor). */
int vqn_dense_backward_data(vqn_ctx* ctx, const float* dz, int64_t lddz, const float* w, float* dx, int64_t lddx,
                            const float* yprev, int64_t ldyp, int act_prev, int accumulate, int64_t m, int k, int n,
                            vqn_stream stream);
/* dW[k,n] += X[m,k]^T . dZ[m,n];  db[n] += colsum(dZ)  (gradient buffers are accumulated: zero them per step) */
int vqn_dense_backward_weights(vqn_ctx* ctx, const float* x, int64_t ldx, const float* dz, int64_t lddz, float* dw,
                               float* db, int64_t m, int k, int n, vqn_stream stream);
/* Learnable tone scaling of non-'nerf' data in the TRAINING step (models/vq_nfr.py:715-718, 736-745):
 *   out = clip((lin * gpar[0]) ^ clip(gpar[1], 0, 5), 0, 1),  gpar = [_gamma_bias, _gamma_index] in DEVICE memory;
 * backward: d_lin = d_out * d out / d lin (both clips pass the gradient: clip_by_value_preserve_gradient) and
 * d_gpar[0..1] += the sums over all `count` elements (atomic).  lin >= 0 (a light integral). */
int vqn_gamma_forward(vqn_ctx* ctx, const float* lin, const float* gpar, float* out, int64_t count, vqn_stream stream);
int vqn_gamma_backward(vqn_ctx* ctx, const float* lin, const float* gpar, const float* d_out, float* d_lin, float* d_gpar,
                       int64_t count, vqn_stream stream);

/* Batched backward GEMMs: up to 32 INDEPENDENT problems of one kind in ONE launch (the same-level layers of the head
 * networks; every weight-gradient GEMM of the step).  Fields per problem:
 *   backward-data:    a = dZ [m,n] (ld lda), w = W [k,n] row-major, out = dX [m,k] (ld ldo), yprev/ldy/act_prev as in
 *                     vqn_dense_backward_data, accumulate 0 = store, 1 = add, 2 = add atomically (shared target);
 *   backward-weights: a = X [m,k] (ld lda), w = dZ [m,n] (ld ldw), out = dW [k,n] (+=, atomics), colsum = db [n] or NULL. */
typedef struct vqn_dense_problem {
  const float* a; int64_t lda;
  const float* w; int64_t ldw;
  float* out; int64_t ldo;
  const float* yprev; int64_t ldy;
  float* colsum;
  int64_t m; int32_t k, n, act_prev, accumulate;
} vqn_dense_problem;
int vqn_dense_backward_data_batched(vqn_ctx* ctx, const vqn_dense_problem* problems, int count, vqn_stream stream);
int vqn_dense_backward_weights_batched(vqn_ctx* ctx, const vqn_dense_problem* problems, int count, vqn_stream stream);

/* dZ = scale * dY * act'(Y) through a net's LAST activation; Y is stored as out_scale * act(.) + out_bias */
int vqn_act_backward(vqn_ctx* ctx, const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t m, int n,
                     int act, float scale, float out_scale, float out_bias, float* dz, int64_t lddz,
                     vqn_stream stream);
/* vqn_act_backward for up to 8 networks (the six heads of a training step) in ONE launch */
typedef struct vqn_act_job {
  const float* dy; const float* y; float* dz;
  int64_t lddy, ldy, lddz, m;
  int32_t n, act;
  float scale, out_scale, out_bias, reserved;
} vqn_act_job;
int vqn_act_backward_batched(vqn_ctx* ctx, const vqn_act_job* jobs, int count, vqn_stream stream);
/* dst[:, 0:w] (leading dim ldd) = src[:, 0:w] (leading dim lds): the x half of concat(y, x) (mlp.py:47-48) */
int vqn_copy_cols(vqn_ctx* ctx, const float* src, int64_t lds, float* dst, int64_t ldd, int64_t m, int w,
                  vqn_stream stream);
/* up to 16 column-block copies (vqn_copy_cols) in ONE launch */
typedef struct vqn_copy_job { const float* src; float* dst; int64_t lds, ldd, m; int32_t w, reserved; } vqn_copy_job;
int vqn_copy_cols_batched(vqn_ctx* ctx, const vqn_copy_job* jobs, int count, vqn_stream stream);
/* Backward of vqn_shade for probe 0 (linear rgb; clip_by_value_preserve_gradient has identity gradient,
 * vq_nfr.py:718,759): d_rgb [n,3] compact -> d_albedo [n,3], d_spec [n,3], d_rough [n,1] compact and
 * d_light [512,3] (ACCUMULATED over points and calls; gradient w.r.t. the raw _light variable). */
int vqn_shade_backward(vqn_ctx* ctx, const float* xyz, const float* rayo, const float* normal, const float* lvis,
                       const int32_t* row_idx, int64_t n, const float* albedo, const float* spec, const float* rough,
                       const float* lxyz, const float* lareas, const float* light, int clip_light0,
                       const float* d_rgb, float* d_albedo, float* d_spec, float* d_rough, float* d_light,
                       vqn_stream stream);
/* compute_loss(mode='train') (vq_nfr.py:923-986) per example, without the two broadcast scalars (vqloss,
 * sim_smooth), and the gradients of sum(loss) * inv_global_bs (tf.nn.compute_average_loss, train_nfr.py:571)
 * w.r.t. rgb, vq_rgb, z_vq (pair smoothness) and spec (lambert).  Rows are (pixel, neighbour) pairs: n even.
 * sums[6] (optional, accumulated): rgb, vqrgb, chromaticity, chr_smooth, lambert, total. */
int vqn_loss_train(vqn_ctx* ctx, const float* gtc, const float* rgb, const float* vqrgb, const float* z_vq,
                   const float* spec, const float* rough, int64_t n, int z_dim, int data_is_nerf,
                   float combine_weight, float chromaticity_weight, float mat_sloss_weight, float lambert_weight,
                   float chr_alpha, float chr_thres, float inv_global_bs, float* loss_rows, float* d_rgb,
                   float* d_vqrgb, float* d_z, float* d_spec, float* sums, vqn_stream stream);
/* Backward of l2_normalize + VectorQuantizerEMA (vq_nfr.py:575-577, vq_layers.py:302,321,327): straight-through
 * d z_norm = d z_vq, plus the commitment term commit_coef * (z_norm - codebook[:,idx]) with
 * commit_coef = vq_loss_weight * commitment_cost * 2 / (global_bs * Z), through x * rsqrt(max(|x|^2, 1e-6)). */
int vqn_vq_backward(vqn_ctx* ctx, const float* z_enc, const int64_t* indices, const float* codebook, int k,
                    const float* d_zvq, float commit_coef, int64_t n, int accumulate, float* d_zenc,
                    vqn_stream stream);
/* vqn_vq_backward that also forms the dz of the layer that produced z_enc (its activated output):
 * dz_out[n, lddz] = d_zenc * act'(z_enc)  (nfr_unit.py:122: the bottleneck ends in a sigmoid); dz_out may be NULL. */
int vqn_vq_backward_act(vqn_ctx* ctx, const float* z_enc, const int64_t* indices, const float* codebook, int k,
                        const float* d_zvq, float commit_coef, int64_t n, int accumulate, float* d_zenc, int act,
                        float* dz_out, int64_t lddz, vqn_stream stream);
/* memset(ptrs[i], 0, bytes[i]) for up to 8 device buffers in ONE launch (4-byte aligned, sizes multiples of 4):
 * the accumulators of a training step (gradients, VQ statistics, d_z) */
int vqn_zero_batched(vqn_ctx* ctx, void* const* ptrs, const int64_t* bytes, int count, vqn_stream stream);
/* stats32 = (float) stats64 (the VQ statistics' slot of the all-reduce buffer) and rows_slot[0] += rows (optional) */
int vqn_train_pack_stats(vqn_ctx* ctx, const double* stats64, float* stats32, int64_t count, float* rows_slot, float rows,
                         vqn_stream stream);
/* spec = ks*basecolor, albedo = (1-ks)*basecolor backward (vq_nfr.py:590-591); d_spec_extra (optional) is added
 * to d_spec (the lambert term). */
int vqn_material_combine_backward(vqn_ctx* ctx, const float* basecolor, const float* ks, const float* d_albedo,
                                  const float* d_spec, const float* d_spec_extra, int64_t n, float* d_basecolor,
                                  float* d_ks, vqn_stream stream);
/* sim_smooth (vq_nfr.py:958-972) on get_codebook(raw): loss_out[1] = -log(min_{i!=j} |c_i - c_j|) (unweighted),
 * d_raw[Z,K] (+)= grad_scale * d loss / d raw. */
int vqn_codebook_sim_loss(vqn_ctx* ctx, const float* raw_codebook, int z_dim, int k, float grad_scale,
                          float* loss_out, float* d_raw, int accumulate, vqn_stream stream);
/* tf.keras.optimizers.Adam(amsgrad=True) dense update on a flat parameter buffer (train_nfr.py:121-139);
 * lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) is computed by the caller; lr_t_dev (optional device scalar)
 * overrides lr_t so that a CUDA-graph replay of the step can change the rate without re-capturing. */
int vqn_adam_amsgrad(vqn_ctx* ctx, float* param, const float* grad, float* m, float* v, float* vhat, int64_t count,
                     float lr_t, const float* lr_t_dev, float beta1, float beta2, float epsilon, vqn_stream stream);
/* VQ statistics (float64) <-> the fp32 tail of the flat all-reduce buffer */
int vqn_cast_f64_f32(vqn_ctx* ctx, const double* src, float* dst, int64_t count, vqn_stream stream);
int vqn_cast_f32_f64(vqn_ctx* ctx, const float* src, double* dst, int64_t count, vqn_stream stream);
/* The scalar tail of a training step (train_nfr.py:571; vq_nfr.py:971-986) in one launch:
 * out[4] = { sums[5]*inv_gbs + r*(vq_w*vq_loss[0] + sim_w*sim_loss[0]), vq_w*vq_loss[0], sim_w*sim_loss[0], r },
 * r = sums[6]*inv_gbs (share of active rows: weight of the two broadcast scalars).  sim_loss may be NULL. */
int vqn_train_scalars(vqn_ctx* ctx, const float* sums, const float* vq_loss, const float* sim_loss, float inv_gbs,
                      float vq_w, float sim_w, float* out, vqn_stream stream);

/* ---- training-batch assembler: outer_sample, nerfactor/train_nfr.py:380-467 ------------------- */
/* Index half (:401-448): one random 8-neighbour per interior pixel, pairs whose alphas both exceed alpha_thres,
 * bs draws with replacement -> rows int32 [2*bs] = [p1, p1_n, p2, p2_n, ...] (linear pixel indices of the H x W view;
 * -1 when no pair is valid), n_valid[1].  Randomness: splitmix64 counter hash of (seed, stream, index) -- TF's RNG
 * is not reproducible outside TF.  Workspaces: flag_ws float / nb_ws, valid_ws int32 of (H-2)(W-2) entries,
 * compact_ws int32 of (H-2)(W-2)/1024 + 2. */
int vqn_sample_pairs(vqn_ctx* ctx, const float* alpha, int h, int w, int use_alpha_thres, float alpha_thres, int bs,
                     uint64_t seed, float* flag_ws, int32_t* nb_ws, int32_t* valid_ws, int32_t* compact_ws,
                     int32_t* n_valid, int32_t* rows, vqn_stream stream);
/* tf.gather_nd(tensor, select_ind) (:450-465): out[n_out,c] = src[rows[o],:] */
int vqn_gather_rows(vqn_ctx* ctx, const float* src, const int32_t* rows, int64_t n_out, int c, float* out,
                    vqn_stream stream);

/* ---- NeuS geo stage (secondary path) -------------------------------------------------------- */
/* NeuSRenderer.up_sample (geo/NeuS-ours2/models/renderer.py:131-175) incl. sample_pdf(det=True)
 * (:39-69): z_samples[B,n_importance] from z_vals[B,S], sdf[B,S]. One warp per ray. */
int vqn_neus_up_sample(vqn_ctx* ctx, const float* rays_o, const float* rays_d, const float* z_vals,
                       const float* sdf, int64_t n_rays, int n_samples, float r_limit,
                       int n_importance, float inv_s, float* z_samples, vqn_stream stream);

/* NeuSRenderer.cat_z_vals sort/merge half (:177-191): merges new_z[B,I] into z_vals[B,S] (sorted),
 * carrying sdf along (new_sdf may be NULL when last=True): z_out[B,S+I], sdf_out[B,S+I]. */
int vqn_neus_cat_z_vals(vqn_ctx* ctx, const float* z_vals, const float* new_z, const float* sdf,
                        const float* new_sdf, int64_t n_rays, int n_samples, int n_importance,
                        float* z_out, float* sdf_out, vqn_stream stream);

/* vqn_neus_up_sample that also writes the positions of the new samples, pts_out[B,n_importance,3] = o + d * z_sample
 * (renderer.py:184, the input of the next SDF call); pts_out may be NULL. */
int vqn_neus_up_sample_pts(vqn_ctx* ctx, const float* rays_o, const float* rays_d, const float* z_vals,
                           const float* sdf, int64_t n_rays, int n_samples, float r_limit,
                           int n_importance, float inv_s, float* z_samples, float* pts_out, vqn_stream stream);

/* One hierarchical-sampling step of NeuSRenderer.render (renderer.py:343-366) in ONE launch: cat_z_vals of step i
 * (z_vals[B,S] + new_z[B,I] with sdf[B,S] + new_sdf[B,I] -> z_out / sdf_out [B,S+I], may both be NULL), then up_sample of
 * step i+1 on the merged row (n_importance new samples at inv_s -> new_z_out[B,n_importance] and their positions
 * pts_out[B,n_importance,3], either may be NULL), and with final_merge != 0 the last cat_z_vals (last=True: no SDF) and
 * render_core's mid points: z_final[B,S+I+n_importance], mid_pts / mid_dirs [B,S+I+n_importance,3] (:203-216).
 * Bit-identical to vqn_neus_cat_z_vals + vqn_neus_up_sample (+ vqn_neus_cat_z_vals + vqn_neus_mid_points). */
typedef struct vqn_neus_step_args {
  const float* rays_o; const float* rays_d; const float* z_vals; const float* new_z; const float* sdf; const float* new_sdf;
  int64_t n_rays; int32_t n_samples; int32_t n_new; int32_t n_importance; int32_t final_merge;
  float r_limit, inv_s, sample_dist, reserved;
  float* z_out; float* sdf_out; float* new_z_out; float* pts_out; float* z_final; float* mid_pts; float* mid_dirs;
} vqn_neus_step_args;
int vqn_neus_scan_step(vqn_ctx* ctx, const vqn_neus_step_args* args, vqn_stream stream);

/* NeuSRenderer.render_core compositing half (:229-282), n_outside == 0:
 *   inputs per sample: sdf[B,S], gradients[B,S,3], sampled_color[B,S,3], z_vals[B,S]; per ray rays_o,
 *   rays_d; scalars inv_s (already clipped to [1e-6,1e6]), cos_anneal_ratio, sample_dist, radius,
 *   background_rgb (3 floats or NULL).
 *   outputs (may be NULL): color[B,3], weights[B,S], surf[B,3], depth[B,1], cdf[B,S] (= prev_cdf),
 *   inside_sphere[B,S], mid_z_vals[B,S], dists[B,S], weight_sum[B,1], weight_max[B,1],
 *   grad_err_sums[2] (float64: sum(relax*err), sum(relax)) accumulated for gradient_error. */
typedef struct vqn_neus_composite_args {
  const float* rays_o; const float* rays_d; const float* z_vals; const float* sdf;
  const float* gradients; const float* sampled_color;
  int64_t n_rays; int32_t n_samples; int32_t reserved;
  float inv_s, cos_anneal_ratio, sample_dist, radius;
  const float* background_rgb;
  float* color; float* weights; float* surf; float* depth; float* cdf; float* inside_sphere;
  float* mid_z_vals; float* dists; float* weight_sum; float* weight_max; double* grad_err_sums;
} vqn_neus_composite_args;
int vqn_neus_composite(vqn_ctx* ctx, const vqn_neus_composite_args* args, vqn_stream stream);

/* mid-point sample positions of render_core (:205-216): mid_z = z + dists/2 with the last dist =
 * sample_dist; pts[B,S,3] = o + d*mid_z, dirs[B,S,3] = d.  (inputs of the SDF / colour networks) */
int vqn_neus_mid_points(vqn_ctx* ctx, const float* rays_o, const float* rays_d, const float* z_vals,
                        int64_t n_rays, int n_samples, float sample_dist, float* pts, float* dirs,
                        vqn_stream stream);

/* SDFNetwork.forward / .sdf / .gradient (geo/NeuS-ours2/models/fields.py:74-112) in ONE launch of the fused
 * tensor-core MLP kernel.  `trunk` = lin0 .. lin{L-2} packed as one network [in_dim = 3 + 6*n_freqs, skip_at = the layer
 * BEFORE the skip_in layer, activations VQN_ACT_SOFTPLUS100] with the effective weights: weight_norm folded
 * (w = g v/|v|), Keras layout [in,out], and the 1/sqrt(2) of the skip concat (fields.py:82) folded into the skip_in
 * layer.  The last Linear (d_hidden -> 1 + d_feature) is passed split: w_sdf[d_hidden] / b_sdf[1] = its output 0,
 * `feat` = a one-layer network holding outputs 1.. [NULL when feat_out is NULL].  scale == 1 (every shipped conf).
 *   sdf[n]  <- forward(x)[:, :1];  feat_out[n, feat_stride >= d_feature] <- forward(x)[:, 1:]  (optional)
 *   grad_out[n,3] <- d sdf / d x (optional), inside the same launch (autograd's second pass, fields.py:98-110, does not
 *   exist): grad_mode 0 propagates (value, d/dx, d/dy, d/dz) jets -- four rows of the 128-row MMA tile per point, nothing
 *   stored; grad_mode 1 runs the forward pass on value rows with act' of every hidden layer stashed in a per-CTA L2-resident
 *   scratch, then the transposed layers back to the embedding (2 row-passes per point instead of 4).
 * precision: VQN_PREC_TF32X3 or VQN_PREC_BF16. */
int vqn_sdf_forward(vqn_ctx* ctx, vqn_net* trunk, const float* w_sdf, const float* b_sdf, vqn_net* feat,
                    int n_freqs, const float* pts, int64_t n, float* sdf, float* feat_out, int64_t feat_stride,
                    float* grad_out, int grad_mode, int precision, vqn_stream stream);

/* RenderingNetwork input (fields.py:147-156, mode 'idr'): writes [points(3), embed(view_dirs)(3+6*multires_view),
 * normals(3), zero pad] into columns [col_off, col_off + width) of rows[n, row_stride]; the feature vector is
 * written in place by vqn_sdf_forward (feat_out = rows), so the 289-wide concat is never copied. */
int vqn_neus_color_input(vqn_ctx* ctx, const float* pts, const float* dirs, const float* normals, int64_t n,
                         int multires_view, float* rows, int64_t row_stride, int col_off, int width,
                         vqn_stream stream);

/* Light-visibility extraction, Runner.compute_vis (geo/NeuS-ours2/gen_geo.py:182-257) + intersect_circle (:346-357):
 * ray set-up for the pairs (surface point p, light l0 + j), pair index m = p * n_chunk + j:
 *   rays_o[m] = surf[p]; rays_d[m] = normalize(lxyz[l] - surf[p]); front[m] = (rays_d . normal[p] > 0) as 1.0 / 0.0;
 *   far[m] = larger root of |o + t d| = radius; near[m] = min(0.1, far / 2).
 * The caller compacts the front-lit pairs (vqn_compact_mask on `front`), renders them (NeuSRenderer.render) and
 * scatters 1 - weight_sum back with vqn_neus_lvis_scatter; back-lit entries of lvis[n_pts, n_lights] stay 0. */
int vqn_neus_light_rays(vqn_ctx* ctx, const float* surf, const float* normal, const float* lxyz, int64_t n_pts,
                        int l0, int n_chunk, float radius, float* rays_o, float* rays_d, float* near_out,
                        float* far_out, float* front, vqn_stream stream);
int vqn_neus_lvis_scatter(vqn_ctx* ctx, const float* weight_sum, const int32_t* row_idx, const int32_t* n_dev,
                          int64_t n_max, int l0, int n_chunk, int n_lights, float* lvis, vqn_stream stream);

/* ---- measurement helper (not a reference interface) -------------------------------------------- */
/* FP32-FMA peak of this GPU in TFLOP/s (mode 0: FFMA, mode 1: packed fma.rn.f32x2); synchronises. */
int vqn_microbench_fma(vqn_ctx* ctx, int mode, int iters, double* tflops_out);
/* Single-tile tcgen05 GEMM self-test of the tensor-core primitives: d[128,n] = a[128,k] . b[n,k]^T.
 * mode 0: kind::tf32 (inputs truncated to tf32), 1: kind::f16 bf16 operands, 2: 3xTF32 split (fp32 parity). */
int vqn_tc_selftest(vqn_ctx* ctx, int mode, int n, int k, const float* a, const float* b, float* d,
                    vqn_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* VQNERF_B200_H_ */
