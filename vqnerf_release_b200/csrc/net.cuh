// vqn_net: library-owned packed copies of one mlp.Network (nerfactor/networks/mlp.py:24-50).
#pragma once
#include "common.cuh"

struct TcPack;  // tensor-core weight images (mlp_tc.cu)

struct vqn_net {
  vqn_ctx* ctx;
  vqn_net_desc desc;       // host copy (device weight pointers are the caller's)
  int n_layers;
  int in_dim, in_pad;
  // fp32 SIMT packing (mlp_simt.cu): [K][Npad] row-major, rows in smem order ([x_pad ; y] after a skip)
  int K[VQN_MAX_LAYERS], N[VQN_MAX_LAYERS], Npad[VQN_MAX_LAYERS];
  float* packed_w[VQN_MAX_LAYERS];
  float* packed_b[VQN_MAX_LAYERS];
  bool simt_ok;            // false: only the tensor-core modes can run this network (NeuS widths / softplus)
  TcPack* tc_pack[2];      // [0] = tf32 hi/lo images, [1] = bf16 images; built lazily on first use
};

static inline int vqn_round_up(int x, int m) { return (x + m - 1) / m * m; }
