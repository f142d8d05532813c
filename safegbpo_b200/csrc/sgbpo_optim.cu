// Policy update tail of SHAC.update_policy (reference: src/learning_algorithms/shac.py:148-149):
//   torch.nn.utils.clip_grad_norm_(policy.parameters(), 1.0); policy_optim.step()   (torch.optim.Adam)
// as TWO small launches over all parameter tensors instead of ~30 foreach kernels: global gradient norm,
// clip coefficient max_norm / (norm + 1e-6) clamped to 1, then the Adam update (no weight decay, no
// amsgrad) written into torch's own optimizer state tensors (exp_avg, exp_avg_sq, step), so
// `policy_optim.state_dict()` stays what torch would have produced.
#include <cuda_runtime.h>
#include <cstdio>

#include <mutex>
#include <unordered_map>

#include "../../include/sgbpo.h"
#include "sgbpo_slots.cuh"

namespace sgbpo {
extern thread_local char g_err[512];

int stream_slot(cudaStream_t st) {
  static std::mutex mu;
  static std::unordered_map<cudaStream_t, int> slots;
  std::lock_guard<std::mutex> lock(mu);
  auto it = slots.find(st);
  if (it != slots.end()) return it->second;
  const int slot = (int)(slots.size() % N_SLOTS);
  slots.emplace(st, slot);
  return slot;
}

constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_CTAS = 296;
// scratch of the two-kernel update, one slot per calling stream (sgbpo_slots.cuh)
__device__ unsigned int g_optim_ticket[N_SLOTS];
__device__ double g_optim_partial[N_SLOTS][OPT_MAX_CTAS];
__device__ double g_optim_scalars[N_SLOTS][2 + 2 * SGBPO_OPT_MAX_TENSORS];      // [0] clip coefficient, then step_size / sqrt(bc2) per tensor

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

// kernel 1: sum of squares of all gradients (grid-strided partials, summed in a fixed order by whichever CTA
// finishes last -> deterministic), clip coefficient, Adam bias corrections, step counters += 1
template <class T, class STEP_T>
__global__ void __launch_bounds__(OPT_THREADS)
grad_norm_kernel(const __grid_constant__ OptArgs<T> A, double lr, double beta1, double beta2, double max_norm,
                 T* __restrict__ total_norm_out, int slot) {
  __shared__ T scratch[OPT_THREADS / 32];
  __shared__ bool s_last;
  const int64_t stride = (int64_t)gridDim.x * OPT_THREADS;
  T acc = T(0);
  for (int i = 0; i < A.n; ++i) {
    const T* __restrict__ g = A.grad[i];
    for (int64_t j = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; j < A.numel[i]; j += stride) acc += g[j] * g[j];
  }
  const T part = block_sum<T>(acc, scratch);
  if (threadIdx.x == 0) {
    g_optim_partial[slot][blockIdx.x] = (double)part;
    __threadfence();
    s_last = atomicInc(&g_optim_ticket[slot], gridDim.x - 1) == gridDim.x - 1;      // wraps to 0 for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    T total = T(0);
    for (unsigned b = 0; b < gridDim.x; ++b) total += (T)(*(volatile double*)&g_optim_partial[slot][b]);
    const T total_norm = sqrt(total);
    T coef = T(max_norm) / (total_norm + T(1e-6));     // clip_grad_norm_: max_norm / (norm + 1e-6), clamped to 1
    g_optim_scalars[slot][0] = (double)((coef < T(1) || coef != coef) ? coef : T(1));      // torch.clamp(max=1) keeps a NaN norm NaN
    if (total_norm_out) *total_norm_out = total_norm;
  }
  if (threadIdx.x < A.n) {
    STEP_T* sp = reinterpret_cast<STEP_T*>(A.step[threadIdx.x]);
    const double step = (double)(*sp) + 1.0;
    *sp = (STEP_T)step;
    g_optim_scalars[slot][2 + 2 * threadIdx.x] = lr / (1.0 - pow(beta1, step));
    g_optim_scalars[slot][3 + 2 * threadIdx.x] = sqrt(1.0 - pow(beta2, step));
  }
}

// kernel 2: scale the gradients in place (clip_grad_norm_ does) and apply Adam
template <class T>
__global__ void __launch_bounds__(OPT_THREADS)
adam_apply_kernel(const __grid_constant__ OptArgs<T> A, double beta1, double beta2, double eps, int slot) {
  const T coef = (T)g_optim_scalars[slot][0];
  const T one_m_b1 = T(1.0 - beta1), b2 = T(beta2), one_m_b2 = T(1.0 - beta2), eps_t = T(eps);
  const int64_t stride = (int64_t)gridDim.x * OPT_THREADS;
  for (int i = 0; i < A.n; ++i) {
    const T step_size = (T)g_optim_scalars[slot][2 + 2 * i], bc2_sqrt = (T)g_optim_scalars[slot][3 + 2 * i];
    T* __restrict__ p = A.param[i]; T* __restrict__ g = A.grad[i]; T* __restrict__ m = A.exp_avg[i]; T* __restrict__ v = A.exp_avg_sq[i];
    for (int64_t j = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; j < A.numel[i]; j += stride) {
      const T gj = g[j] * coef;
      const T mj = m[j] + (gj - m[j]) * one_m_b1;                    // exp_avg.lerp_(grad, 1 - beta1)
      const T vj = v[j] * b2 + gj * gj * one_m_b2;                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      g[j] = gj;
      m[j] = mj;
      v[j] = vj;
      p[j] = p[j] - step_size * (mj / (sqrt(vj) / bc2_sqrt + eps_t));
    }
  }
}

// small models (the policies of this path: ~18 k parameters): ONE CTA does both phases -- sum of squares, clip coefficient and bias
// corrections, then the update -- instead of two launches with a grid-wide ticket in between
constexpr int OPT_SMALL_THREADS = 1024, OPT_SMALL_MAX = 64 * 1024;
template <class T, class STEP_T>
__global__ void __launch_bounds__(OPT_SMALL_THREADS)
clip_adam_small_kernel(const __grid_constant__ OptArgs<T> A, double lr, double beta1, double beta2, double eps, double max_norm,
                       T* __restrict__ total_norm_out) {
  __shared__ T scratch[OPT_SMALL_THREADS / 32];
  __shared__ T s_coef, s_step[SGBPO_OPT_MAX_TENSORS], s_bc2[SGBPO_OPT_MAX_TENSORS];
  T acc = T(0);
  for (int i = 0; i < A.n; ++i) {
    const T* __restrict__ g = A.grad[i];
    for (int64_t j = threadIdx.x; j < A.numel[i]; j += OPT_SMALL_THREADS) acc += g[j] * g[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    T t = scratch[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      const T total_norm = sqrt(t);
      const T coef = T(max_norm) / (total_norm + T(1e-6));
      s_coef = (coef < T(1) || coef != coef) ? coef : T(1);
      if (total_norm_out) *total_norm_out = total_norm;
    }
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + A.n) {
    const int i = threadIdx.x - 32;
    STEP_T* sp = reinterpret_cast<STEP_T*>(A.step[i]);
    const double step = (double)(*sp) + 1.0;
    *sp = (STEP_T)step;
    s_step[i] = (T)(lr / (1.0 - pow(beta1, step)));
    s_bc2[i] = (T)sqrt(1.0 - pow(beta2, step));
  }
  __syncthreads();
  const T coef = s_coef, one_m_b1 = T(1.0 - beta1), b2 = T(beta2), one_m_b2 = T(1.0 - beta2), eps_t = T(eps);
  for (int i = 0; i < A.n; ++i) {
    const T step_size = s_step[i], bc2_sqrt = s_bc2[i];
    T* __restrict__ p = A.param[i]; T* __restrict__ g = A.grad[i]; T* __restrict__ m = A.exp_avg[i]; T* __restrict__ v = A.exp_avg_sq[i];
    for (int64_t j = threadIdx.x; j < A.numel[i]; j += OPT_SMALL_THREADS) {
      const T gj = g[j] * coef;
      const T mj = m[j] + (gj - m[j]) * one_m_b1;
      const T vj = v[j] * b2 + gj * gj * one_m_b2;
      g[j] = gj;
      m[j] = mj;
      v[j] = vj;
      p[j] = p[j] - step_size * (mj / (sqrt(vj) / bc2_sqrt + eps_t));
    }
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
  int64_t total = 0;
  for (int i = 0; i < o->n_tensors; ++i) {
    if (!o->param[i] || !o->grad[i] || !o->exp_avg[i] || !o->exp_avg_sq[i] || !o->step[i] || o->numel[i] < 0)
      return opt_fail(SGBPO_E_ARG, "sgbpo_clip_adam: null tensor pointer");
    a.param[i] = (T*)o->param[i]; a.grad[i] = (T*)o->grad[i]; a.exp_avg[i] = (T*)o->exp_avg[i];
    a.exp_avg_sq[i] = (T*)o->exp_avg_sq[i]; a.step[i] = o->step[i]; a.numel[i] = o->numel[i];
    total += o->numel[i];
  }
  if (total <= OPT_SMALL_MAX && !(getenv("SGBPO_OPT_SMALL") && getenv("SGBPO_OPT_SMALL")[0] == '0')) {
    if (step_dtype == SGBPO_F32)
      clip_adam_small_kernel<T, float><<<1, OPT_SMALL_THREADS, 0, st>>>(a, o->lr, o->beta1, o->beta2, o->eps, o->max_norm, (T*)total_norm_out);
    else
      clip_adam_small_kernel<T, double><<<1, OPT_SMALL_THREADS, 0, st>>>(a, o->lr, o->beta1, o->beta2, o->eps, o->max_norm, (T*)total_norm_out);
    const cudaError_t e1 = cudaGetLastError();
    if (e1 != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "clip_adam_small_kernel: %s", cudaGetErrorString(e1));
      return SGBPO_E_CUDA;
    }
    return SGBPO_OK;
  }
  int64_t grid = (total + 4 * OPT_THREADS - 1) / (4 * OPT_THREADS);
  grid = grid < 1 ? 1 : (grid > OPT_MAX_CTAS ? OPT_MAX_CTAS : grid);
  const int slot = stream_slot(st);
  if (step_dtype == SGBPO_F32)
    grad_norm_kernel<T, float><<<(int)grid, OPT_THREADS, 0, st>>>(a, o->lr, o->beta1, o->beta2, o->max_norm, (T*)total_norm_out, slot);
  else
    grad_norm_kernel<T, double><<<(int)grid, OPT_THREADS, 0, st>>>(a, o->lr, o->beta1, o->beta2, o->max_norm, (T*)total_norm_out, slot);
  adam_apply_kernel<T><<<(int)grid, OPT_THREADS, 0, st>>>(a, o->beta1, o->beta2, o->eps, slot);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "clip_adam kernels: %s", cudaGetErrorString(e));
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
