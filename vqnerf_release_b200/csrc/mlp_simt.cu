// Fused small-MLP kernel, fp32 parity mode (VQN_PREC_FP32): FFMA on CUDA cores.
//
// Reference: networks/mlp.py:24-50 (Dense stack + skip concat), networks/embedder.py:23-47,
// models/vq_nfr.py:771-828 (_pred_enc_at, _pred_diff_at, _pred_spec_at, _pred_rough_at),
// models/nfr_unit.py:110-129 (net shapes).
//
// One CTA owns a tile of 64 points and runs a whole "program" (a chain of Dense layers from one or
// several mlp.Networks) without leaving the SM: activations live TRANSPOSED in two shared-memory buffers
// P[256][68] and Q[384][68] (row = feature, column = point, stride 68 floats so that both the float4
// row reads of the GEMM and the column-wise input transposition are (nearly) conflict free); weights
// are streamed from L2 in [32 x 128] chunks through a 2-stage cp.async ring that prefetches across
// layer boundaries.  A thread computes a 4 (points) x 8 (outputs) register tile: per k it issues
// 3 LDS.128 for 32 FFMA, so the kernel is FFMA-issue bound.  Skip connections never copy data: the
// packed weight rows of the layer after a skip are permuted to [x_pad ; y] and the layer schedule is
// chosen so that y lands next to the still-resident x (see build_net_steps).
// Narrow output layers (N <= 8: the 3/1-wide sigmoid heads) run as a per-(point,output) dot product.
#include <string.h>

#include "net.cuh"

#define MLP_TM 64
#define MLP_LD 68
#define MLP_KC 32
#define MLP_NB 128
#define MLP_THREADS 256
#define MLP_P_ROWS 256
#define MLP_Q_ROWS 384
#define MLP_MAX_STEPS 24
#define MLP_MAX_OUTS 8

enum { STEP_WIDE = 0, STEP_NARROW = 1, STEP_COPY = 2 };

struct MlpStep {
  const float* w;  // packed [K][Npad]
  const float* b;  // [Npad]
  int kind;
  int K, N, Npad;
  int in_buf, in_off, out_buf, out_off;  // buffer 0 = P, 1 = Q
  int act;
  int out_slot;  // >= 0: also/only written to global outs[out_slot] as [point][N]
  float post_scale, post_bias;
};

struct MlpProgram {
  int n_steps;
  int input_mode;  // 0: embed(xyz) -> rows [0,64) of Q (and P when emb_both); 1: z rows -> Q[0,in_dim)
  int n_freqs;
  int in_dim;      // input_mode 1: row length of `in`
  int in_pad;      // rows zero-filled up to in_pad
  const float* in;
  const int* row_idx;
  const int* n_dev;
  long long n;
  float* outs[MLP_MAX_OUTS];
  int out_stride[MLP_MAX_OUTS];
  int* nonfinite;
  MlpStep steps[MLP_MAX_STEPS];
};

static inline int round_up(int x, int m) { return vqn_round_up(x, m); }

// ---------------------------------------------------------------------------------------------
// weight packing: Keras [in,out] -> [K][Npad], K = rows in the kernel's smem order ([x_pad ; y] after a skip)
// ---------------------------------------------------------------------------------------------
__global__ void pack_dense_kernel(const float* __restrict__ w, const float* __restrict__ b, int in_rows,
                                  int n_out, int K, int Npad, int y_rows, int x_rows, int x_pad,
                                  float* __restrict__ pw, float* __restrict__ pb) {
  // destination row r: if x_rows > 0 (layer after a skip): r < x_pad -> source row y_rows + r (if r < x_rows)
  //                                                        r >= x_pad -> source row r - x_pad
  //                    else: source row r (if r < in_rows)
  int total = K * Npad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int r = i / Npad, c = i % Npad;
    int src = -1;
    if (x_rows > 0) {
      if (r < x_pad) src = r < x_rows ? y_rows + r : -1;
      else src = (r - x_pad) < y_rows ? (r - x_pad) : -1;
    } else {
      src = r < in_rows ? r : -1;
    }
    pw[i] = (src >= 0 && c < n_out) ? w[(size_t)src * n_out + c] : 0.f;
  }
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < Npad; c += gridDim.x * blockDim.x)
    pb[c] = c < n_out ? b[c] : 0.f;
}

static int net_layout(vqn_net* net) {
  const vqn_net_desc& d = net->desc;
  net->n_layers = d.n_layers;
  net->in_dim = d.in_dim;
  net->in_pad = round_up(d.in_dim, MLP_KC);
  int prev = d.in_dim;
  for (int i = 0; i < d.n_layers; ++i) {
    bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
    int k;
    if (i == 0) k = net->in_pad;
    else if (after_skip) k = net->in_pad + round_up(d.widths[i - 1], MLP_KC);
    else k = round_up(prev, MLP_KC);
    net->K[i] = k;
    net->N[i] = d.widths[i];
    bool last = (i == d.n_layers - 1);
    net->Npad[i] = (last && d.widths[i] <= 8) ? 8 : round_up(d.widths[i], MLP_NB);
    prev = d.widths[i];
  }
  return VQN_OK;
}

int vqn_tc_pack_create(vqn_net* net, cudaStream_t s);   // mlp_tc.cu
void vqn_tc_pack_destroy(vqn_net* net);

static int net_pack(vqn_net* net, cudaStream_t s) {
  const vqn_net_desc& d = net->desc;
  for (int i = 0; i < d.n_layers; ++i) {
    bool after_skip = (d.skip_at >= 0 && i == d.skip_at + 1);
    int in_rows = (i == 0) ? d.in_dim : (after_skip ? d.widths[i - 1] + d.in_dim : d.widths[i - 1]);
    int y_rows = after_skip ? d.widths[i - 1] : 0;
    int x_rows = after_skip ? d.in_dim : 0;
    pack_dense_kernel<<<64, 256, 0, s>>>(d.w[i], d.b[i], in_rows, d.widths[i], net->K[i], net->Npad[i], y_rows,
                                         x_rows, net->in_pad, net->packed_w[i], net->packed_b[i]);
    net->ctx->launches.fetch_add(1);
    VQN_CUDA(cudaGetLastError());
  }
  return VQN_OK;
}

extern "C" int vqn_net_create(vqn_ctx* ctx, const vqn_net_desc* desc, vqn_net** out, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && desc && out, "net_create: null");
  VQN_CHECK_ARG(desc->n_layers >= 1 && desc->n_layers <= VQN_MAX_LAYERS, "net_create: 1..8 layers");
  VQN_CHECK_ARG(desc->in_dim >= 1 && desc->in_dim <= 512, "net_create: in_dim must be <= 512");
  VQN_CHECK_ARG(desc->skip_at < desc->n_layers - 1, "net_create: skip_at must precede the last layer");
  // The FFMA kernel (VQN_PREC_FP32) handles the decomposition-stage shapes; anything else (NeuS: softplus, a 217-wide
  // layer, the 292-wide colour input) runs in the tensor-core modes only.
  bool simt_ok = desc->in_dim <= 256;
  for (int i = 0; i < desc->n_layers; ++i) {
    VQN_CHECK_ARG(desc->w[i] && desc->b[i], "net_create: null weight pointer");
    VQN_CHECK_ARG(desc->widths[i] >= 1 && desc->widths[i] <= 256, "net_create: widths must be <= 256");
    VQN_CHECK_ARG(desc->acts[i] >= 0 && desc->acts[i] <= 3, "net_create: bad activation");
    if (desc->acts[i] == VQN_ACT_SOFTPLUS100) simt_ok = false;
    if (i < desc->n_layers - 1 && desc->widths[i] % 32 != 0) simt_ok = false;
  }
  if (desc->skip_at >= 0 && round_up(desc->in_dim, MLP_KC) + desc->widths[desc->skip_at] > MLP_Q_ROWS) simt_ok = false;
  vqn_net* net = new vqn_net();
  net->ctx = ctx;
  net->desc = *desc;
  net->simt_ok = simt_ok;
  net->tc_pack[0] = nullptr; net->tc_pack[1] = nullptr;
  net_layout(net);
  cudaStream_t s = vqn_cs(stream);
  for (int i = 0; i < desc->n_layers; ++i) {
    net->packed_w[i] = nullptr; net->packed_b[i] = nullptr;
  }
  for (int i = 0; i < desc->n_layers; ++i) {
    VQN_CUDA(cudaMalloc(&net->packed_w[i], sizeof(float) * net->K[i] * net->Npad[i]));
    VQN_CUDA(cudaMalloc(&net->packed_b[i], sizeof(float) * net->Npad[i]));
  }
  int rc = net_pack(net, s);
  if (rc != VQN_OK) return rc;
  rc = vqn_tc_pack_create(net, s);
  if (rc != VQN_OK) return rc;
  *out = net;
  return VQN_OK;
}

extern "C" int vqn_net_repack(vqn_net* net, const vqn_net_desc* desc, vqn_stream stream) {
  VQN_CHECK_ARG(net && desc, "net_repack: null");
  VQN_CHECK_ARG(desc->n_layers == net->desc.n_layers && desc->in_dim == net->desc.in_dim &&
                    desc->skip_at == net->desc.skip_at, "net_repack: architecture changed");
  for (int i = 0; i < desc->n_layers; ++i)
    VQN_CHECK_ARG(desc->widths[i] == net->desc.widths[i] && desc->w[i] && desc->b[i], "net_repack: shape changed");
  net->desc = *desc;
  int rc = net_pack(net, vqn_cs(stream));
  if (rc != VQN_OK) return rc;
  return vqn_tc_pack_create(net, vqn_cs(stream));
}

extern "C" int vqn_net_destroy(vqn_net* net) {
  if (!net) return VQN_OK;
  for (int i = 0; i < net->n_layers; ++i) { cudaFree(net->packed_w[i]); cudaFree(net->packed_b[i]); }
  vqn_tc_pack_destroy(net);
  delete net;
  return VQN_OK;
}

extern "C" int vqn_net_out_dim(const vqn_net* net) { return net ? net->desc.widths[net->n_layers - 1] : 0; }

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
struct ChunkIt { int step, nb, kc; };

__device__ __forceinline__ bool chunk_valid(const MlpProgram& pg, const ChunkIt& it) { return it.step < pg.n_steps; }

__device__ __forceinline__ void chunk_first(const MlpProgram& pg, ChunkIt& it) {
  it.step = 0; it.nb = 0; it.kc = 0;
  while (it.step < pg.n_steps && pg.steps[it.step].kind != STEP_WIDE) it.step++;
}
__device__ __forceinline__ void chunk_next(const MlpProgram& pg, ChunkIt& it) {
  const MlpStep& st = pg.steps[it.step];
  if (++it.kc * MLP_KC >= st.K) {
    it.kc = 0;
    if (++it.nb * MLP_NB >= st.Npad) {
      it.nb = 0;
      do { it.step++; } while (it.step < pg.n_steps && pg.steps[it.step].kind != STEP_WIDE);
    }
  }
}
__device__ __forceinline__ void chunk_issue(const MlpProgram& pg, const ChunkIt& it, float* ws, int tid) {
  const MlpStep& st = pg.steps[it.step];
  const float* src = st.w + (size_t)(it.kc * MLP_KC) * st.Npad + it.nb * MLP_NB;
#pragma unroll
  for (int i = 0; i < (MLP_KC * MLP_NB / 4) / MLP_THREADS; ++i) {
    int idx = tid + i * MLP_THREADS;
    int r = idx >> 5, c4 = idx & 31;
    cp_async16(ws + r * MLP_NB + c4 * 4, src + (size_t)r * st.Npad + c4 * 4);
  }
}

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_simt_kernel(const __grid_constant__ MlpProgram pg) {
  extern __shared__ __align__(16) float sm[];
  float* bufP = sm;                                   // [256][68]
  float* bufQ = bufP + MLP_P_ROWS * MLP_LD;           // [384][68]
  float* wst = bufQ + MLP_Q_ROWS * MLP_LD;            // [2][32][128]
  float* xin = wst + 2 * MLP_KC * MLP_NB;             // [64][4] tile inputs (xyz)
  const int tid = threadIdx.x;
  const int mg = tid & 15, ng = tid >> 4;             // 16 point-groups x 16 output-groups
  const int m0 = mg * 4;

  long long n = pg.n_dev ? (long long)*pg.n_dev : pg.n;
  if (n > pg.n) n = pg.n;
  const long long n_tiles = (n + MLP_TM - 1) / MLP_TM;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * MLP_TM;
    __syncthreads();  // previous tile fully consumed
    // ---------------- input stage ----------------
    ChunkIt it; chunk_first(pg, it);
    int stage = 0;
    if (chunk_valid(pg, it)) { chunk_issue(pg, it, wst, tid); }
    cp_async_commit();
    if (pg.input_mode == 0) {
      if (tid < MLP_TM * 3) {
        int m = tid / 3, c = tid % 3;
        long long i = base + m;
        float v = 0.f;
        if (i < n) { long long row = pg.row_idx ? (long long)pg.row_idx[i] : i; v = pg.in[row * 3 + c]; }
        xin[m * 4 + c] = v;
      }
      __syncthreads();
      // Embedder: rows [0,63) of Q; row 63.. in_pad zero
      const int d = 3 + 6 * pg.n_freqs;
      for (int idx = tid; idx < pg.in_pad * MLP_TM; idx += MLP_THREADS) {
        int r = idx / MLP_TM, m = idx % MLP_TM;
        float v = 0.f;
        if (r < 3) v = xin[m * 4 + r];
        else if (r < d) {
          int q = r - 3, f = q / 6, w = q % 6;
          float a = xin[m * 4 + (w % 3)] * exp2f((float)f);
          v = w < 3 ? sinf(a) : cosf(a);
        }
        bufQ[r * MLP_LD + m] = v;
      }
    } else {
      // z rows [n,in_dim] -> Q[k][m] (transposed); lanes walk k so global reads are coalesced
      const int dpad = pg.in_pad;
      for (int idx = tid; idx < dpad * MLP_TM; idx += MLP_THREADS) {
        int k = idx % dpad, m = idx / dpad;
        long long i = base + m;
        float v = 0.f;
        if (i < n && k < pg.in_dim) {
          long long row = pg.row_idx ? (long long)pg.row_idx[i] : i;
          v = pg.in[row * pg.in_dim + k];
        }
        bufQ[k * MLP_LD + m] = v;
      }
    }
    // (first barrier of the step loop orders these writes before any read)

    // ---------------- layer steps ----------------
    for (int s = 0; s < pg.n_steps; ++s) {
      const MlpStep& st = pg.steps[s];
      const float* inb = (st.in_buf ? bufQ : bufP) + st.in_off * MLP_LD;
      float* outb = (st.out_buf ? bufQ : bufP) + st.out_off * MLP_LD;
      if (st.kind == STEP_WIDE) {
        for (int nb = 0; nb * MLP_NB < st.Npad; ++nb) {
          float acc[4][8];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
          for (int kc = 0; kc * MLP_KC < st.K; ++kc) {
            ChunkIt nx = it; chunk_next(pg, nx);
            const bool has_next = chunk_valid(pg, nx);
            if (has_next) chunk_issue(pg, nx, wst + (stage ^ 1) * MLP_KC * MLP_NB, tid);
            cp_async_commit();
            cp_async_wait<1>();          // chunk `it` (older group) has landed
            __syncthreads();
            const float* wsc = wst + stage * MLP_KC * MLP_NB + ng * 8;
            const float* ac = inb + (kc * MLP_KC) * MLP_LD + m0;
#pragma unroll 8
            for (int k = 0; k < MLP_KC; ++k) {
              const float4 av = *reinterpret_cast<const float4*>(ac + k * MLP_LD);
              const float4 w0 = *reinterpret_cast<const float4*>(wsc + k * MLP_NB);
              const float4 w1 = *reinterpret_cast<const float4*>(wsc + k * MLP_NB + 4);
              const float a4[4] = {av.x, av.y, av.z, av.w};
              const float w8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a4[i], w8[j], acc[i][j]);
            }
            __syncthreads();             // stage may be overwritten by the next prefetch
            it = nx; stage ^= 1;
          }
          // epilogue: bias + activation -> smem (transposed) and/or global
          const int n0 = nb * MLP_NB + ng * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int nn = n0 + j;
            const float bj = st.b[nn];
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = vqn_apply_act(acc[i][j] + bj, st.act);
            if (nn < st.N) {
              if (st.out_buf >= 0)
                *reinterpret_cast<float4*>(outb + nn * MLP_LD + m0) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i][j] = v[i];
            }
          }
          if (st.out_slot >= 0) {
            float* go = pg.outs[st.out_slot];
            const int gs = pg.out_stride[st.out_slot];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              long long pi = base + m0 + i;
              if (pi < n && n0 + 7 < st.N) {
                float4 o0 = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                float4 o1 = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
                bool bad = false;
#pragma unroll
                for (int j = 0; j < 8; ++j) bad |= !isfinite(acc[i][j]);
                if (bad) atomicOr(pg.nonfinite, 1);
                *reinterpret_cast<float4*>(go + pi * gs + n0) = o0;
                *reinterpret_cast<float4*>(go + pi * gs + n0 + 4) = o1;
              } else if (pi < n) {
                for (int j = 0; j < 8; ++j)
                  if (n0 + j < st.N) go[pi * gs + n0 + j] = acc[i][j];
              }
            }
          }
        }
      } else if (st.kind == STEP_NARROW) {
        __syncthreads();                 // inputs written by the previous epilogue
        const int m = tid & 63, jj = tid >> 6;
        float a0 = 0.f, a1 = 0.f;
        const bool h0 = jj < st.N, h1 = jj + 4 < st.N;
        if (h0) {
          const float* wp = st.w;
          for (int k = 0; k < st.K; ++k) {
            float av = inb[k * MLP_LD + m];
            a0 = fmaf(av, __ldg(wp + k * 8 + jj), a0);
            if (h1) a1 = fmaf(av, __ldg(wp + k * 8 + jj + 4), a1);
          }
          long long pi = base + m;
          if (pi < n && st.out_slot >= 0) {
            float* go = pg.outs[st.out_slot];
            const int gs = pg.out_stride[st.out_slot];
            float v0 = vqn_apply_act(a0 + st.b[jj], st.act) * st.post_scale + st.post_bias;
            if (!isfinite(v0)) atomicOr(pg.nonfinite, 1);
            go[pi * gs + jj] = v0;
            if (h1) {
              float v1 = vqn_apply_act(a1 + st.b[jj + 4], st.act) * st.post_scale + st.post_bias;
              if (!isfinite(v1)) atomicOr(pg.nonfinite, 1);
              go[pi * gs + jj + 4] = v1;
            }
          }
        }
      } else {  // STEP_COPY: rows [in_off, in_off+K) of in_buf -> out_buf rows [out_off, ...)
        __syncthreads();
        for (int idx = tid; idx < st.K * MLP_TM; idx += MLP_THREADS) {
          int r = idx / MLP_TM, m = idx % MLP_TM;
          outb[r * MLP_LD + m] = inb[r * MLP_LD + m];
        }
        __syncthreads();
      }
    }
    cp_async_wait<0>();
  }
}

// ---------------------------------------------------------------------------------------------
// host-side program assembly
// ---------------------------------------------------------------------------------------------
struct ProgBuilder {
  MlpProgram pg;
  int cur_buf, cur_off, cur_rows;   // where the current activation sits
  bool ok;
  ProgBuilder() { memset(&pg, 0, sizeof(pg)); ok = true; cur_buf = 1; cur_off = 0; cur_rows = 0; }
  MlpStep* add() {
    if (pg.n_steps >= MLP_MAX_STEPS) { ok = false; return nullptr; }
    MlpStep* s = &pg.steps[pg.n_steps++];
    memset(s, 0, sizeof(*s));
    s->out_slot = -1; s->post_scale = 1.f; s->post_bias = 0.f; s->out_buf = -1;
    return s;
  }
};

// Append one mlp.Network whose input x currently sits at (cur_buf, cur_off) with in_pad rows.
// out_slot_last: global output slot of the last layer (-1: keep in smem); out_slot_any: additionally write
// the last WIDE layer to global.
static bool build_net_steps(ProgBuilder& B, const vqn_net* net, int out_slot_last, float post_scale,
                            float post_bias) {
  const vqn_net_desc& d = net->desc;
  const int L = d.n_layers;
  const int skip = d.skip_at;
  if (!net->simt_ok) return false;     // tensor-core-only network (see vqn_net_create)
  // a net with a skip needs its input x resident in Q at offset 0 (x stays there until the concat)
  if (skip >= 0 && !(B.cur_buf == 1 && B.cur_off == 0)) {
    MlpStep* c = B.add(); if (!c) return false;
    c->kind = STEP_COPY; c->K = net->in_pad; c->in_buf = B.cur_buf; c->in_off = B.cur_off;
    c->out_buf = 1; c->out_off = 0;
    B.cur_buf = 1; B.cur_off = 0;
  }
  const int xpad = net->in_pad;
  int in_buf = B.cur_buf, in_off = B.cur_off;
  for (int i = 0; i < L; ++i) {
    MlpStep* s = B.add(); if (!s) return false;
    const bool last = (i == L - 1);
    s->w = net->packed_w[i]; s->b = net->packed_b[i];
    s->K = net->K[i]; s->N = net->N[i]; s->Npad = net->Npad[i];
    s->act = d.acts[i];
    s->in_buf = in_buf; s->in_off = in_off;
    s->kind = (net->Npad[i] == 8) ? STEP_NARROW : STEP_WIDE;
    if (last) {
      s->out_slot = out_slot_last;
      s->post_scale = post_scale; s->post_bias = post_bias;
    }
    if (s->kind == STEP_NARROW) { s->out_buf = -1; break; }
    // choose the output buffer: before/at the skip layer, layer `skip` must land in Q right after x_pad
    int ob, oo;
    if (skip >= 0 && i <= skip) {
      ob = ((skip - i) % 2 == 0) ? 1 : 0;
      oo = ob == 1 ? xpad : 0;
    } else {
      ob = in_buf ^ 1; oo = 0;
    }
    if (ob == 0 && s->N > MLP_P_ROWS) return false;
    if (ob == 1 && oo + s->N > MLP_Q_ROWS) return false;
    // a layer may not write rows it (or a concurrent n-block) still reads
    if (ob == in_buf && !(oo >= in_off + s->K || oo + s->N <= in_off)) return false;
    s->out_buf = ob; s->out_off = oo;
    if (last && out_slot_last < 0) { /* stays in smem */ }
    if (skip >= 0 && i == skip) { in_buf = 1; in_off = 0; }       // next layer reads [x_pad ; y]
    else { in_buf = ob; in_off = oo; }
    B.cur_buf = ob; B.cur_off = oo; B.cur_rows = s->N;
  }
  return true;
}

static int launch_program(vqn_ctx* ctx, MlpProgram& pg, cudaStream_t s) {
  pg.nonfinite = ctx->nonfinite_flag;
  size_t smem = sizeof(float) * ((MLP_P_ROWS + MLP_Q_ROWS) * MLP_LD + 2 * MLP_KC * MLP_NB + MLP_TM * 4);
  VQN_CHECK_ARG((int)smem <= ctx->max_smem_optin, "mlp: shared memory budget exceeded");
  VQN_CUDA(cudaFuncSetAttribute(mlp_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long tiles = (pg.n + MLP_TM - 1) / MLP_TM;
  int blocks = (int)(tiles < (long long)ctx->sm_count ? tiles : (long long)ctx->sm_count);
  mlp_simt_kernel<<<blocks, MLP_THREADS, smem, s>>>(pg);
  VQN_LAUNCHED(ctx);
  return VQN_OK;
}

int vqn_tc_net_forward(vqn_net* net, const float* x, int64_t n, float* y, int precision, cudaStream_t s);
int vqn_tc_pred_enc_at(vqn_ctx* ctx, vqn_net* fe, vqn_net* bn, int n_freqs, const float* pts, const int32_t* row_idx,
                       const int32_t* n_dev, int64_t n, float* z, int precision, cudaStream_t s);
int vqn_tc_pred_heads(vqn_ctx* ctx, vqn_net* diff, vqn_net* spec, vqn_net* rough, const float* z,
                      const int32_t* n_dev, int64_t n, float slope, float bias, float* d, float* sp, float* r,
                      int precision, cudaStream_t s);

extern "C" int vqn_net_forward(vqn_net* net, const float* x, int64_t n, float* y, int precision, vqn_stream stream) {
  VQN_CHECK_ARG(net && x && y && n >= 0, "net_forward args");
  if (n == 0) return VQN_OK;
  if (precision != VQN_PREC_FP32) return vqn_tc_net_forward(net, x, n, y, precision, vqn_cs(stream));
  ProgBuilder B;
  B.pg.input_mode = 1; B.pg.in = x; B.pg.in_dim = net->in_dim; B.pg.in_pad = net->in_pad; B.pg.n = n;
  B.cur_buf = 1; B.cur_off = 0;
  B.pg.outs[0] = y; B.pg.out_stride[0] = net->desc.widths[net->n_layers - 1];
  if (!build_net_steps(B, net, 0, 1.f, 0.f) || !B.ok) {
    vqn_set_error("net_forward: network does not fit the fused-kernel buffers");
    return VQN_ERR_UNSUPPORTED;
  }
  return launch_program(net->ctx, B.pg, vqn_cs(stream));
}

extern "C" int vqn_pred_enc_at(vqn_ctx* ctx, vqn_net* fine_enc, vqn_net* bottleneck, int n_freqs, const float* pts,
                               const int32_t* row_idx, const int32_t* n_dev, int64_t n, float* z_out, int precision,
                               vqn_stream stream) {
  VQN_CHECK_ARG(ctx && fine_enc && bottleneck && pts && z_out && n >= 0, "pred_enc_at args");
  VQN_CHECK_ARG(fine_enc->in_dim == 3 + 6 * n_freqs, "pred_enc_at: fine_enc in_dim != 3 + 6*n_freqs");
  VQN_CHECK_ARG(bottleneck->in_dim == vqn_net_out_dim(fine_enc), "pred_enc_at: bottleneck in_dim mismatch");
  if (n == 0) return VQN_OK;
  if (precision != VQN_PREC_FP32)
    return vqn_tc_pred_enc_at(ctx, fine_enc, bottleneck, n_freqs, pts, row_idx, n_dev, n, z_out, precision,
                              vqn_cs(stream));
  ProgBuilder B;
  B.pg.input_mode = 0; B.pg.in = pts; B.pg.n_freqs = n_freqs; B.pg.in_pad = fine_enc->in_pad; B.pg.n = n;
  B.pg.row_idx = row_idx; B.pg.n_dev = n_dev;
  B.pg.outs[0] = z_out; B.pg.out_stride[0] = vqn_net_out_dim(bottleneck);
  bool ok = build_net_steps(B, fine_enc, -1, 1.f, 0.f) && build_net_steps(B, bottleneck, 0, 1.f, 0.f);
  if (!ok || !B.ok) { vqn_set_error("pred_enc_at: program does not fit"); return VQN_ERR_UNSUPPORTED; }
  return launch_program(ctx, B.pg, vqn_cs(stream));
}

extern "C" int vqn_pred_heads(vqn_ctx* ctx, vqn_net* diff, vqn_net* spec, vqn_net* rough, const float* z,
                              const int32_t* n_dev, int64_t n, float albedo_slope, float albedo_bias, float* diff_out, float* spec_out,
                              float* rough_out, int precision, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && z && n >= 0, "pred_heads args");
  VQN_CHECK_ARG((diff == nullptr) == (diff_out == nullptr) && (spec == nullptr) == (spec_out == nullptr) &&
                    (rough == nullptr) == (rough_out == nullptr), "pred_heads: net/output mismatch");
  VQN_CHECK_ARG(diff || spec || rough, "pred_heads: no head given");
  if (n == 0) return VQN_OK;
  if (precision != VQN_PREC_FP32)
    return vqn_tc_pred_heads(ctx, diff, spec, rough, z, n_dev, n, albedo_slope, albedo_bias, diff_out, spec_out, rough_out,
                             precision, vqn_cs(stream));
  vqn_net* nets[3] = {diff, spec, rough};
  float* outs[3] = {diff_out, spec_out, rough_out};
  ProgBuilder B;
  int in_dim = 0;
  for (int h = 0; h < 3; ++h) if (nets[h]) { in_dim = nets[h]->in_dim; break; }
  B.pg.input_mode = 1; B.pg.in = z; B.pg.in_dim = in_dim; B.pg.in_pad = round_up(in_dim, MLP_KC); B.pg.n = n;
  B.pg.n_dev = n_dev;
  for (int h = 0; h < 3; ++h) {
    if (!nets[h]) continue;
    VQN_CHECK_ARG(nets[h]->in_dim == in_dim, "pred_heads: heads disagree on z_dim");
    B.cur_buf = 1; B.cur_off = 0;   // every head restarts from z, which stays resident in Q[0,in_pad)
    B.pg.outs[h] = outs[h]; B.pg.out_stride[h] = vqn_net_out_dim(nets[h]);
    bool ok = build_net_steps(B, nets[h], h, h == 0 ? albedo_slope : 1.f, h == 0 ? albedo_bias : 0.f);
    if (!ok || !B.ok) { vqn_set_error("pred_heads: program does not fit"); return VQN_ERR_UNSUPPORTED; }
    // the head must not have clobbered z: it may only write Q rows >= in_pad
    for (int s = 0; s < B.pg.n_steps; ++s)
      if (B.pg.steps[s].kind == STEP_WIDE && B.pg.steps[s].out_buf == 1 && B.pg.steps[s].out_off < B.pg.in_pad) {
        vqn_set_error("pred_heads: head would overwrite the shared latent (needs skip_at == n_layers-2)");
        return VQN_ERR_UNSUPPORTED;
      }
  }
  return launch_program(ctx, B.pg, vqn_cs(stream));
}

// _pred_enc_at + the three main heads in ONE launch: z stays in shared memory (Q[0,256)) between the bottleneck
// and the heads, so the 1 KB/point latent never round-trips through HBM unless z_out is requested
// (fast_render, models/vq_nfr.py:321-331).
int vqn_tc_mlp_main(vqn_ctx* ctx, vqn_net* fe, vqn_net* bn, vqn_net* diff, vqn_net* spec, vqn_net* rough, int n_freqs,
                    const float* pts, const int32_t* row_idx, const int32_t* n_dev, int64_t n, float slope, float bias,
                    float* z_out, float* d, float* sp, float* r, int precision, cudaStream_t s);

extern "C" int vqn_mlp_main(vqn_ctx* ctx, vqn_net* fine_enc, vqn_net* bottleneck, vqn_net* diff, vqn_net* spec,
                            vqn_net* rough, int n_freqs, const float* pts, const int32_t* row_idx,
                            const int32_t* n_dev, int64_t n, float albedo_slope, float albedo_bias, float* z_out,
                            float* diff_out, float* spec_out, float* rough_out, int precision, vqn_stream stream) {
  VQN_CHECK_ARG(ctx && fine_enc && bottleneck && diff && spec && rough && pts && diff_out && spec_out && rough_out &&
                    n >= 0, "mlp_main args");
  VQN_CHECK_ARG(fine_enc->in_dim == 3 + 6 * n_freqs, "mlp_main: fine_enc in_dim != 3 + 6*n_freqs");
  VQN_CHECK_ARG(bottleneck->in_dim == vqn_net_out_dim(fine_enc), "mlp_main: bottleneck in_dim mismatch");
  if (n == 0) return VQN_OK;
  if (precision != VQN_PREC_FP32)
    return vqn_tc_mlp_main(ctx, fine_enc, bottleneck, diff, spec, rough, n_freqs, pts, row_idx, n_dev, n, albedo_slope,
                           albedo_bias, z_out, diff_out, spec_out, rough_out, precision, vqn_cs(stream));
  ProgBuilder B;
  B.pg.input_mode = 0; B.pg.in = pts; B.pg.n_freqs = n_freqs; B.pg.in_pad = fine_enc->in_pad; B.pg.n = n;
  B.pg.row_idx = row_idx; B.pg.n_dev = n_dev;
  const int z_dim = vqn_net_out_dim(bottleneck);
  B.pg.outs[3] = z_out; B.pg.out_stride[3] = z_dim;
  bool ok = build_net_steps(B, fine_enc, -1, 1.f, 0.f) && build_net_steps(B, bottleneck, z_out ? 3 : -1, 1.f, 0.f);
  ok = ok && B.cur_buf == 1 && B.cur_off == 0;      // z must sit in Q[0, z_dim) for the heads
  const int first_head_step = B.pg.n_steps;
  vqn_net* nets[3] = {diff, spec, rough};
  float* outs[3] = {diff_out, spec_out, rough_out};
  for (int h = 0; ok && h < 3; ++h) {
    ok = ok && nets[h]->in_dim == z_dim;
    B.cur_buf = 1; B.cur_off = 0;
    B.pg.outs[h] = outs[h]; B.pg.out_stride[h] = vqn_net_out_dim(nets[h]);
    ok = ok && build_net_steps(B, nets[h], h, h == 0 ? albedo_slope : 1.f, h == 0 ? albedo_bias : 0.f);
  }
  for (int s = first_head_step; ok && s < B.pg.n_steps; ++s)
    if (B.pg.steps[s].kind == STEP_WIDE && B.pg.steps[s].out_buf == 1 && B.pg.steps[s].out_off < z_dim) ok = false;
  if (!ok || !B.ok) { vqn_set_error("mlp_main: program does not fit the fused kernel"); return VQN_ERR_UNSUPPORTED; }
  return launch_program(ctx, B.pg, vqn_cs(stream));
}
