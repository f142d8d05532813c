// TD-lambda value targets of SHAC's critic (reference: src/learning_algorithms/shac.py:187-241,
// `calculate_estimated_values`): per environment a backward scan over the buffer rows t = H-1 .. 0 with three
// running quantities (td coefficient, discounted average of the bootstrapped returns, terminal return).  The
// reference walks the rows with masked scatter updates (~12 small torch launches per row); here one thread owns
// one environment, the three quantities stay in registers and every row is read once (coalesced across
// environments: buffers are time-major (T, N)).  Rows that are terminal produce no target (flagged in `valid`).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/sgbpo.h"

namespace sgbpo {
extern thread_local char g_err[512];

template <class T>
__global__ void __launch_bounds__(256)
td_lambda_kernel(int64_t N, int H, const T* __restrict__ rewards, const T* __restrict__ values,
                 const uint8_t* __restrict__ terminals, T gamma, T lam, T* __restrict__ est, uint8_t* __restrict__ valid) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  T td = T(1), avg = T(0), term_ret = T(0);
  bool term_next = terminals[(int64_t)H * N + e] != 0;
  for (int t = H - 1; t >= 0; --t) {
    const bool term_t = terminals[(int64_t)t * N + e] != 0;
    const T r = rewards[(int64_t)t * N + e], v_next = values[(int64_t)(t + 1) * N + e];
    if (!term_t) {
      if (term_next) {                       // shac.py:219-223: first estimate below a terminal row
        td = T(1); avg = T(0);
        term_ret = r + gamma * v_next;
      } else {                               // shac.py:225-232
        td = td * lam;
        avg = avg * (lam * gamma);
        avg = avg + gamma * v_next;
        avg = avg + ((T(1) - td) / (T(1) - lam)) * r;
        term_ret = r + gamma * term_ret;
      }
      est[(int64_t)t * N + e] = (T(1) - lam) * avg + td * term_ret;     // shac.py:235-236
      valid[(int64_t)t * N + e] = 1;
    } else {
      est[(int64_t)t * N + e] = T(0);
      valid[(int64_t)t * N + e] = 0;
    }
    term_next = term_t;
  }
}
}  // namespace sgbpo

using namespace sgbpo;

extern "C" int sgbpo_td_lambda(int dtype, int64_t N, int H, const void* rewards, const void* values, const uint8_t* terminals,
                               double gamma, double td_weight, void* est, uint8_t* valid, void* stream) {
  if (N <= 0 || H <= 0) return SGBPO_OK;
  if (!rewards || !values || !terminals || !est || !valid) {
    snprintf(g_err, sizeof(g_err), "sgbpo_td_lambda: null buffer");
    return SGBPO_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)((N + 255) / 256);
  if (dtype == SGBPO_F64)
    td_lambda_kernel<double><<<grid, 256, 0, st>>>(N, H, (const double*)rewards, (const double*)values, terminals, gamma,
                                                   td_weight, (double*)est, valid);
  else
    td_lambda_kernel<float><<<grid, 256, 0, st>>>(N, H, (const float*)rewards, (const float*)values, terminals, (float)gamma,
                                                  (float)td_weight, (float*)est, valid);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "td_lambda_kernel: %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  return SGBPO_OK;
}
