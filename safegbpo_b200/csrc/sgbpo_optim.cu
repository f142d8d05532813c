// Policy update tail of SHAC.update_policy (reference: src/learning_algorithms/shac.py:148-149):
//   torch.nn.utils.clip_grad_norm_(policy.parameters(), 1.0); policy_optim.step()   (torch.optim.Adam)
// as ONE launch over all parameter tensors instead of ~30 small foreach kernels: global gradient norm,
// clip coefficient max_norm / (norm + 1e-6) clamped to 1, then the Adam update (no weight decay, no
// amsgrad) written into torch's own optimizer state tensors (exp_avg, exp_avg_sq, step), so
// `policy_optim.state_dict()` stays what torch would have produced.
#include <cuda_runtime.h>
#include <cstdio>

#include "../../include/sgbpo.h"

namespace sgbpo {
extern thread_local char g_err[512];

constexpr int OPT_THREADS = 1024;

template <class T> struct OptArgs {
  int n;
  T* param[SGBPO_OPT_MAX_TENSORS];
  T* grad[SGBPO_OPT_MAX_TENSORS];
  T* exp_avg[SGBPO_OPT_MAX_TENSORS];
  T* exp_avg_sq[SGBPO_OPT_MAX_TENSORS];
  void* step[SGBPO_OPT_MAX_TENSORS];
  int64_t numel[SGBPO_OPT_MAX_TENSORS];
};

template <class T> __device__ __forceinline__ T block_sum(T v, T* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T t = lane < (OPT_THREADS / 32) ? scratch[lane] : T(0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;      // every thread holds the block total
}

// single CTA: the policies on this path have 1e4..1e6 parameters, far below what would need a grid-wide reduction
template <class T, class STEP_T>
__global__ void __launch_bounds__(OPT_THREADS)
clip_adam_kernel(const __grid_constant__ OptArgs<T> A, double lr, double beta1, double beta2, double eps, double max_norm,
                 T* __restrict__ total_norm_out) {
  __shared__ T scratch[OPT_THREADS / 32];
  // pass 1: global L2 norm of all gradients (clip_grad_norm_: norm of the per-tensor norms == sqrt of the total sum of squares)
  T acc = T(0);
  for (int i = 0; i < A.n; ++i) {
    const T* g = A.grad[i];
    for (int64_t j = threadIdx.x; j < A.numel[i]; j += OPT_THREADS) acc += g[j] * g[j];
  }
  const T total_norm = sqrt(block_sum<T>(acc, scratch));
  T coef = T(max_norm) / (total_norm + T(1e-6));
  coef = coef < T(1) ? coef : T(1);
  if (threadIdx.x == 0 && total_norm_out) *total_norm_out = total_norm;
  // pass 2: scale the gradients in place (clip_grad_norm_ does) and apply Adam
  for (int i = 0; i < A.n; ++i) {
    STEP_T* sp = reinterpret_cast<STEP_T*>(A.step[i]);
    const double step = (double)(*sp) + 1.0;
    const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
    const T step_size = T(lr / bc1), bc2_sqrt = T(sqrt(bc2));
    T* p = A.param[i]; T* g = A.grad[i]; T* m = A.exp_avg[i]; T* v = A.exp_avg_sq[i];
    for (int64_t j = threadIdx.x; j < A.numel[i]; j += OPT_THREADS) {
      const T gj = g[j] * coef;
      g[j] = gj;
      const T mj = m[j] + (gj - m[j]) * T(1.0 - beta1);              // exp_avg.lerp_(grad, 1 - beta1)
      const T vj = v[j] * T(beta2) + gj * gj * T(1.0 - beta2);        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      m[j] = mj;
      v[j] = vj;
      p[j] = p[j] - step_size * (mj / (sqrt(vj) / bc2_sqrt + T(eps)));
    }
    __syncthreads();                 // all reads of *sp above happen before thread 0 bumps it
    if (threadIdx.x == 0) *sp = (STEP_T)step;
  }
}

static int opt_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

template <class T>
static int launch_clip_adam(const sgbpo_optim* o, int step_dtype, void* total_norm_out, cudaStream_t st) {
  OptArgs<T> a;
  a.n = o->n_tensors;
  for (int i = 0; i < o->n_tensors; ++i) {
    if (!o->param[i] || !o->grad[i] || !o->exp_avg[i] || !o->exp_avg_sq[i] || !o->step[i] || o->numel[i] < 0)
      return opt_fail(SGBPO_E_ARG, "sgbpo_clip_adam: null tensor pointer");
    a.param[i] = (T*)o->param[i]; a.grad[i] = (T*)o->grad[i]; a.exp_avg[i] = (T*)o->exp_avg[i];
    a.exp_avg_sq[i] = (T*)o->exp_avg_sq[i]; a.step[i] = o->step[i]; a.numel[i] = o->numel[i];
  }
  if (step_dtype == SGBPO_F32)
    clip_adam_kernel<T, float><<<1, OPT_THREADS, 0, st>>>(a, o->lr, o->beta1, o->beta2, o->eps, o->max_norm, (T*)total_norm_out);
  else
    clip_adam_kernel<T, double><<<1, OPT_THREADS, 0, st>>>(a, o->lr, o->beta1, o->beta2, o->eps, o->max_norm, (T*)total_norm_out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "clip_adam_kernel: %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  return SGBPO_OK;
}
}  // namespace sgbpo

using namespace sgbpo;

extern "C" int sgbpo_clip_adam(int dtype, int step_dtype, const sgbpo_optim* o, void* total_norm_out, void* stream) {
  if (!o) return opt_fail(SGBPO_E_ARG, "sgbpo_clip_adam: null argument");
  if (o->n_tensors < 0 || o->n_tensors > SGBPO_OPT_MAX_TENSORS) return opt_fail(SGBPO_E_ARG, "sgbpo_clip_adam: too many tensors");
  if (o->n_tensors == 0) return SGBPO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == SGBPO_F64 ? launch_clip_adam<double>(o, step_dtype, total_norm_out, st)
                            : launch_clip_adam<float>(o, step_dtype, total_norm_out, st);
}
