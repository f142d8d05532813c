// C ABI of the fused rollout (include/sgbpo.h); the kernels live in sgbpo_rollout_impl.cuh and are
// instantiated per environment in sgbpo_rollout_env*.cu.
#include "sgbpo_rollout_impl.cuh"
#include "sgbpo_rollout_tc.cuh"

namespace sgbpo {
namespace wm {      // sgbpo_rollout_wm.cuh, instantiated in sgbpo_rollout_wm_env*.cu
template <int ENV> int rollout_wm_fwd_env(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st);
}
int launch_rows_tc(int64_t M, const sgbpo_mlp* net, const void* x, void* out, cudaStream_t st);   // sgbpo_mlp_tc.cu
#define ROLLOUT_DISPATCH(FN, ...)                                                       \
  switch (env_id) {                                                                     \
    case ENV_BALANCE_PENDULUM: return FN<ENV_BALANCE_PENDULUM>(__VA_ARGS__);            \
    case ENV_BALANCE_CARTPOLE: return FN<ENV_BALANCE_CARTPOLE>(__VA_ARGS__);            \
    case ENV_SWINGUP_CARTPOLE: return FN<ENV_SWINGUP_CARTPOLE>(__VA_ARGS__);            \
    case ENV_BALANCE_QUADROTOR: return FN<ENV_BALANCE_QUADROTOR>(__VA_ARGS__);          \
    default: return fail2(SGBPO_E_UNSUPPORTED, "fused rollout: environment not instantiated"); \
  }
#define ROLLOUT_TC_DISPATCH(FN, ...)                                                    \
  switch (env_id) {                                                                     \
    case ENV_BALANCE_PENDULUM: return tcr::FN<ENV_BALANCE_PENDULUM>(__VA_ARGS__);       \
    case ENV_BALANCE_CARTPOLE: return tcr::FN<ENV_BALANCE_CARTPOLE>(__VA_ARGS__);       \
    case ENV_SWINGUP_CARTPOLE: return tcr::FN<ENV_SWINGUP_CARTPOLE>(__VA_ARGS__);       \
    case ENV_BALANCE_QUADROTOR: return tcr::FN<ENV_BALANCE_QUADROTOR>(__VA_ARGS__);     \
    case ENV_MANAGE_HOUSEHOLD: return tcr::FN<ENV_MANAGE_HOUSEHOLD>(__VA_ARGS__);       \
    case ENV_NAVIGATE_SEEKER: return tcr::FN<ENV_NAVIGATE_SEEKER>(__VA_ARGS__);         \
    default: return fail2(SGBPO_E_UNSUPPORTED, "tensor-core rollout: unknown environment"); \
  }
static int tc_fwd_dispatch(int env_id, const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st) {
  ROLLOUT_TC_DISPATCH(rollout_tc_fwd_env, c, pol, r, st)
}
static int wm_fwd_dispatch(int env_id, const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st) {
  ROLLOUT_DISPATCH(wm::rollout_wm_fwd_env, c, pol, r, st)
}
static int tc_jac_dispatch(int env_id, const sgbpo_consts* c, const sgbpo_rollout* r, cudaStream_t st) {
  ROLLOUT_TC_DISPATCH(rollout_jac_env, c, r, st)
}
static int tc_bwd_dispatch(int env_id, const sgbpo_mlp* pol, const sgbpo_rollout* r, const sgbpo_rollout_grads* g, cudaStream_t st) {
  ROLLOUT_TC_DISPATCH(rollout_tc_bwd_env, pol, r, g, st)
}
static int tc_pol_jac_stride(int env_id) {
  switch (env_id) {
    case ENV_BALANCE_PENDULUM: return tcr::pol_jac_stride_of<ENV_BALANCE_PENDULUM>();
    case ENV_BALANCE_CARTPOLE: return tcr::pol_jac_stride_of<ENV_BALANCE_CARTPOLE>();
    case ENV_SWINGUP_CARTPOLE: return tcr::pol_jac_stride_of<ENV_SWINGUP_CARTPOLE>();
    case ENV_BALANCE_QUADROTOR: return tcr::pol_jac_stride_of<ENV_BALANCE_QUADROTOR>();
    case ENV_MANAGE_HOUSEHOLD: return tcr::pol_jac_stride_of<ENV_MANAGE_HOUSEHOLD>();
    case ENV_NAVIGATE_SEEKER: return tcr::pol_jac_stride_of<ENV_NAVIGATE_SEEKER>();
    default: return -1;
  }
}
static int tc_jac_stride(int env_id) {
  switch (env_id) {
    case ENV_BALANCE_PENDULUM: return tcr::JacShape<ENV_BALANCE_PENDULUM>::STRIDE;
    case ENV_BALANCE_CARTPOLE: return tcr::JacShape<ENV_BALANCE_CARTPOLE>::STRIDE;
    case ENV_SWINGUP_CARTPOLE: return tcr::JacShape<ENV_SWINGUP_CARTPOLE>::STRIDE;
    case ENV_BALANCE_QUADROTOR: return tcr::JacShape<ENV_BALANCE_QUADROTOR>::STRIDE;
    case ENV_MANAGE_HOUSEHOLD: return tcr::JacShape<ENV_MANAGE_HOUSEHOLD>::STRIDE;
    case ENV_NAVIGATE_SEEKER: return tcr::JacShape<ENV_NAVIGATE_SEEKER>::STRIDE;
    default: return -1;
  }
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
  const int fwd_engine = (r->engine & 4) ? 2 : (r->engine & 1);
  if ((r->engine & 5) == 5) return fail2(SGBPO_E_ARG, "sgbpo_rollout_fwd: engine bits 0 and 2 are exclusive");
  if (fwd_engine == 0) {
    const int n_stash = (r->stash_h1 != nullptr) + (r->stash_e1 != nullptr) + (r->stash_e2 != nullptr) + (r->stash_stats != nullptr);
    if (n_stash != 0 && n_stash != 4 && !(n_stash == 3 && !r->stash_h1))
      return fail2(SGBPO_E_ARG, "sgbpo_rollout_fwd: pass the stash buffers (h1 optional) or none");
    if (n_stash != 0 && policy->n_hidden != 2) return fail2(SGBPO_E_UNSUPPORTED, "activation stash: policies with two hidden layers");
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (r->engine < 0 || r->engine > 7) return fail2(SGBPO_E_ARG, "sgbpo_rollout_fwd: unknown engine");
  if (fwd_engine == 2) {
    if (dtype != SGBPO_F32) return fail2(SGBPO_E_UNSUPPORTED, "warp-MMA rollout: float32 only");
    return wm_fwd_dispatch(env_id, consts, policy, r, st);
  }
  if (fwd_engine == 1) {
    if (dtype != SGBPO_F32) return fail2(SGBPO_E_UNSUPPORTED, "tensor-core rollout: float32 only");
    return tc_fwd_dispatch(env_id, consts, policy, r, st);
  }
  ROLLOUT_DISPATCH(rollout_fwd_env, dtype, policy->width, consts, policy, r, st)
}

int sgbpo_rollout_jac(int env_id, int dtype, const sgbpo_consts* consts, const sgbpo_rollout* r, void* stream) {
  if (!consts || !r) return fail2(SGBPO_E_ARG, "sgbpo_rollout_jac: null argument");
  if (dtype != SGBPO_F32) return fail2(SGBPO_E_UNSUPPORTED, "sgbpo_rollout_jac: float32 only");
  if (r->N <= 0 || r->H <= 0) return SGBPO_OK;
  if (!r->noise || !r->states || !r->actions || !r->jac) return fail2(SGBPO_E_ARG, "sgbpo_rollout_jac: null buffer");
  return tc_jac_dispatch(env_id, consts, r, (cudaStream_t)stream);
}

int sgbpo_rollout_jac_stride(int env_id) { return tc_jac_stride(env_id); }
int sgbpo_rollout_pol_jac_stride(int env_id) { return tc_pol_jac_stride(env_id); }

int sgbpo_rollout_engine_supported(int env_id, int dtype, const sgbpo_mlp* policy) {
  if (!policy || dtype != SGBPO_F32 || tc_jac_stride(env_id) < 0) return 0;
  tcr::PolicyTc p;
  return tcr::tc_policy(policy, p) ? 1 : 0;
}

int sgbpo_rollout_bwd(int env_id, int dtype, const sgbpo_consts* consts, const sgbpo_mlp* policy,
                      const sgbpo_rollout* r, const sgbpo_rollout_grads* g, void* stream) {
  if (!consts || !policy || !r || !g) return fail2(SGBPO_E_ARG, "sgbpo_rollout_bwd: null argument");
  if (r->N <= 0 || r->H <= 0) return SGBPO_OK;
  if ((r->engine >> 1) & 1) {
    if (dtype != SGBPO_F32) return fail2(SGBPO_E_UNSUPPORTED, "tensor-core rollout: float32 only");
    if (!g->grw || !g->g_w_hid || !g->g_w_in || !g->g_w_out || !g->g_b_hid || !g->g_gamma || !g->g_beta || !g->g_b_out || !g->g_log_std)
      return fail2(SGBPO_E_ARG, "sgbpo_rollout_bwd: null buffer");
    if (!r->stash_e1 || !r->stash_e2 || !r->stash_stats || !r->jac)
      return fail2(SGBPO_E_ARG, "sgbpo_rollout_bwd (reverse engine 1): forward stash and sgbpo_rollout_jac output are required");
    return tc_bwd_dispatch(env_id, policy, r, g, (cudaStream_t)stream);
  }
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
