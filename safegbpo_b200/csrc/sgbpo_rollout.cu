// C ABI of the fused rollout (include/sgbpo.h); the kernels live in sgbpo_rollout_impl.cuh and are
// instantiated per environment in sgbpo_rollout_env*.cu.
#include "sgbpo_rollout_impl.cuh"

namespace sgbpo {
int launch_rows_tc(int64_t M, const sgbpo_mlp* net, const void* x, void* out, cudaStream_t st);   // sgbpo_mlp_tc.cu
#define ROLLOUT_DISPATCH(FN, ...)                                                       \
  switch (env_id) {                                                                     \
    case ENV_BALANCE_PENDULUM: return FN<ENV_BALANCE_PENDULUM>(__VA_ARGS__);            \
    case ENV_BALANCE_CARTPOLE: return FN<ENV_BALANCE_CARTPOLE>(__VA_ARGS__);            \
    case ENV_SWINGUP_CARTPOLE: return FN<ENV_SWINGUP_CARTPOLE>(__VA_ARGS__);            \
    case ENV_BALANCE_QUADROTOR: return FN<ENV_BALANCE_QUADROTOR>(__VA_ARGS__);          \
    default: return fail2(SGBPO_E_UNSUPPORTED, "fused rollout: environment not instantiated"); \
  }
}  // namespace sgbpo

using namespace sgbpo;

extern "C" {

int sgbpo_rollout_fwd(int env_id, int dtype, const sgbpo_consts* consts, const sgbpo_mlp* policy,
                      const sgbpo_rollout* r, void* stream) {
  if (!consts || !policy || !r) return fail2(SGBPO_E_ARG, "sgbpo_rollout_fwd: null argument");
  if (r->N <= 0 || r->H <= 0) return SGBPO_OK;
  if (!r->eps || !r->noise || !r->states || !r->obs || !r->actions || !r->exec_actions || !r->rewards || !r->flags)
    return fail2(SGBPO_E_ARG, "sgbpo_rollout_fwd: null buffer");
  {
    const int n_stash = (r->stash_h1 != nullptr) + (r->stash_e1 != nullptr) + (r->stash_e2 != nullptr) + (r->stash_stats != nullptr);
    if (n_stash != 0 && n_stash != 4) return fail2(SGBPO_E_ARG, "sgbpo_rollout_fwd: pass all four stash buffers or none");
    if (n_stash == 4 && policy->n_hidden != 2) return fail2(SGBPO_E_UNSUPPORTED, "activation stash: policies with two hidden layers");
  }
  cudaStream_t st = (cudaStream_t)stream;
  ROLLOUT_DISPATCH(rollout_fwd_env, dtype, policy->width, consts, policy, r, st)
}

int sgbpo_rollout_bwd(int env_id, int dtype, const sgbpo_consts* consts, const sgbpo_mlp* policy,
                      const sgbpo_rollout* r, const sgbpo_rollout_grads* g, void* stream) {
  if (!consts || !policy || !r || !g) return fail2(SGBPO_E_ARG, "sgbpo_rollout_bwd: null argument");
  if (r->N <= 0 || r->H <= 0) return SGBPO_OK;
  if (!g->grw || !g->dz2_out || !g->g_w_in || !g->g_w_out || !g->g_b_hid || !g->g_gamma || !g->g_beta ||
      !g->g_b_out || !g->g_log_std)
    return fail2(SGBPO_E_ARG, "sgbpo_rollout_bwd: null buffer");
  if (!r->stash_h1 || !r->stash_e1 || !r->stash_e2 || !r->stash_stats)
    return fail2(SGBPO_E_ARG, "sgbpo_rollout_bwd: the forward pass must have been run with the activation stash buffers");
  cudaStream_t st = (cudaStream_t)stream;
  ROLLOUT_DISPATCH(rollout_bwd_env, dtype, policy->width, consts, policy, r, g, st)
}

int sgbpo_mlp_rows(int dtype, int64_t M, const sgbpo_mlp* net, const void* x, void* out, void* stream) {
  if (!net || (M > 0 && (!x || !out))) return fail2(SGBPO_E_ARG, "sgbpo_mlp_rows: null argument");
  if (M <= 0) return SGBPO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SGBPO_F32 && !(getenv("SGBPO_ROWS_TC") && getenv("SGBPO_ROWS_TC")[0] == '0')) {
    // fp32, width 128, 2-3 hidden layers: tcgen05 tensor-core kernel (csrc/sgbpo_mlp_tc.cu); anything else: FMA tiles
    const int rc = launch_rows_tc(M, net, x, out, st);
    if (rc != SGBPO_E_UNSUPPORTED) return rc;
  }
  switch (net->width) {
    case 32: return dtype ? launch_rows<double, 32>(M, net, x, out, st) : launch_rows<float, 32>(M, net, x, out, st);
    case 64: return dtype ? launch_rows<double, 64>(M, net, x, out, st) : launch_rows<float, 64>(M, net, x, out, st);
    case 128: return dtype ? launch_rows<double, 128>(M, net, x, out, st) : launch_rows<float, 128>(M, net, x, out, st);
    default: return fail2(SGBPO_E_UNSUPPORTED, "sgbpo_mlp_rows supports hidden widths 32, 64, 128");
  }
}

int sgbpo_mlp_rows_vjp(int dtype, int64_t M, const sgbpo_mlp* net, const void* x, const void* row_weight, void* gx, void* stream) {
  if (!net || (M > 0 && (!x || !row_weight || !gx))) return fail2(SGBPO_E_ARG, "sgbpo_mlp_rows_vjp: null argument");
  if (M <= 0) return SGBPO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (net->width) {
    case 32: return dtype ? launch_vjp<double, 32>(M, net, x, row_weight, gx, st) : launch_vjp<float, 32>(M, net, x, row_weight, gx, st);
    case 64: return dtype ? launch_vjp<double, 64>(M, net, x, row_weight, gx, st) : launch_vjp<float, 64>(M, net, x, row_weight, gx, st);
    case 128: return dtype ? launch_vjp<double, 128>(M, net, x, row_weight, gx, st) : launch_vjp<float, 128>(M, net, x, row_weight, gx, st);
    default: return fail2(SGBPO_E_UNSUPPORTED, "sgbpo_mlp_rows_vjp supports hidden widths 32, 64, 128");
  }
}

}  // extern "C"
