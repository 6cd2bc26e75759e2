// Tensor-core (tcgen05 / TMEM) MLP modes: VQN_PREC_BF16 and VQN_PREC_TF32X3.
// Round-1 status: packing hooks only; the entry points report VQN_ERR_UNSUPPORTED so that callers fail
// loudly instead of silently falling back to another precision.
#include "common.cuh"

struct vqn_net;
int vqn_tc_pack_create(vqn_net*, cudaStream_t) { return VQN_OK; }
void vqn_tc_pack_destroy(vqn_net*) {}

int vqn_tc_net_forward(vqn_net*, const float*, int64_t, float*, int, cudaStream_t) {
  vqn_set_error("tensor-core MLP modes are not built in this revision; use VQN_PREC_FP32");
  return VQN_ERR_UNSUPPORTED;
}
int vqn_tc_pred_enc_at(vqn_ctx*, vqn_net*, vqn_net*, int, const float*, const int32_t*, const int32_t*, int64_t, float*, int,
                       cudaStream_t) {
  vqn_set_error("tensor-core MLP modes are not built in this revision; use VQN_PREC_FP32");
  return VQN_ERR_UNSUPPORTED;
}
int vqn_tc_pred_heads(vqn_ctx*, vqn_net*, vqn_net*, vqn_net*, const float*, const int32_t*, int64_t, float, float, float*, float*,
                      float*, int, cudaStream_t) {
  vqn_set_error("tensor-core MLP modes are not built in this revision; use VQN_PREC_FP32");
  return VQN_ERR_UNSUPPORTED;
}
int vqn_tc_mlp_main(vqn_ctx*, vqn_net*, vqn_net*, vqn_net*, vqn_net*, vqn_net*, int, const float*, const int32_t*,
                    const int32_t*, int64_t, float, float, float*, float*, float*, float*, int, cudaStream_t) {
  vqn_set_error("tensor-core MLP modes are not built in this revision; use VQN_PREC_FP32");
  return VQN_ERR_UNSUPPORTED;
}
