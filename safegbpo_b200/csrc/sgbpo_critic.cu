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
#include "sgbpo_slots.cuh"

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

// ---------------------------------------------------------------------------------------------------------------
// Critic fit (shac.py:154-184): the hidden layers are Linear -> ELU -> LayerNorm on batches of ~65 k rows.  ATen runs
// bias add, ELU, LayerNorm and their backward as separate memory-bound passes (and its LayerNorm weight-gradient
// kernel takes 0.5 ms for this shape); here ONE pass per direction: a warp owns a row at a time (lane l = columns
// l, l + 32, ...: every access is one coalesced 128-byte segment; row statistics by shuffles), the column sums of the backward (d gamma, d beta, d bias) are
// kept per lane in registers over the warp's rows and flushed with one atomicAdd per column per warp.
// ---------------------------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T wsum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float elu_neg(float x) { return expf(x) - 1.0f; }
__device__ __forceinline__ double elu_neg(double x) { return exp(x) - 1.0; }
__device__ __forceinline__ float rs(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rs(double x) { return 1.0 / sqrt(x); }

template <class T, int W>
__global__ void __launch_bounds__(256)
bias_elu_ln_fwd_kernel(int64_t M, const T* __restrict__ z, const T* __restrict__ bias, const T* __restrict__ gamma,
                       const T* __restrict__ beta, T* __restrict__ h, T* __restrict__ e_out, T* __restrict__ stats) {
  constexpr int CPL = W / 32;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  T b[CPL], g[CPL], be[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { b[j] = bias[lane + 32 * j]; g[j] = gamma[lane + 32 * j]; be[j] = beta[lane + 32 * j]; }
  for (int64_t row = warp; row < M; row += nwarps) {
    T e[CPL], s = T(0);
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const T v = z[row * W + lane + 32 * j] + b[j];
      e[j] = v > T(0) ? v : elu_neg(v);
      s += e[j];
    }
    const T mean = wsum(s) / T(W);
    T q = T(0);
#pragma unroll
    for (int j = 0; j < CPL; ++j) { const T d = e[j] - mean; q += d * d; }
    const T rstd = rs(wsum(q) / T(W) + T(1e-5));
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      e_out[row * W + lane + 32 * j] = e[j];
      h[row * W + lane + 32 * j] = ((e[j] - mean) * rstd) * g[j] + be[j];
    }
    if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
  }
}

template <class T, int W>
__global__ void __launch_bounds__(256)
bias_elu_ln_bwd_kernel(int64_t M, const T* __restrict__ gy, const T* __restrict__ e_in, const T* __restrict__ stats,
                       const T* __restrict__ gamma, T* __restrict__ dz, T* __restrict__ dgamma, T* __restrict__ dbeta,
                       T* __restrict__ dbias) {
  constexpr int CPL = W / 32;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  T g[CPL], ag[CPL], ab[CPL], az[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { g[j] = gamma[lane + 32 * j]; ag[j] = ab[j] = az[j] = T(0); }
  for (int64_t row = warp; row < M; row += nwarps) {
    const T mean = stats[row * 2], rstd = stats[row * 2 + 1];
    T e[CPL], xh[CPL], dx[CPL], s1 = T(0), s2 = T(0);
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const T gyj = gy[row * W + lane + 32 * j];
      e[j] = e_in[row * W + lane + 32 * j];
      xh[j] = (e[j] - mean) * rstd;
      ag[j] += gyj * xh[j];
      ab[j] += gyj;
      dx[j] = gyj * g[j];
      s1 += dx[j];
      s2 += dx[j] * xh[j];
    }
    const T m1 = wsum(s1) / T(W), m2 = wsum(s2) / T(W);
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const T de = rstd * (dx[j] - m1 - xh[j] * m2);
      const T d = de * (e[j] > T(0) ? T(1) : e[j] + T(1));          // dELU/dz from the ELU output
      dz[row * W + lane + 32 * j] = d;
      az[j] += d;
    }
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    atomicAdd(dgamma + lane + 32 * j, ag[j]);
    atomicAdd(dbeta + lane + 32 * j, ab[j]);
    atomicAdd(dbias + lane + 32 * j, az[j]);
  }
}

template <class T, int W>
static int launch_eln(bool fwd, int64_t M, const void* a0, const void* a1, const void* a2, const void* a3, void* o0, void* o1,
                      void* o2, void* o3, cudaStream_t st) {
  int64_t blocks = (M + 7) / 8;                       // 8 warps per block, one row per warp per pass
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (fwd)
    bias_elu_ln_fwd_kernel<T, W><<<(int)blocks, 256, 0, st>>>(M, (const T*)a0, (const T*)a1, (const T*)a2, (const T*)a3, (T*)o0,
                                                              (T*)o1, (T*)o2);
  else
    bias_elu_ln_bwd_kernel<T, W><<<(int)blocks, 256, 0, st>>>(M, (const T*)a0, (const T*)a1, (const T*)a2, (const T*)a3, (T*)o0,
                                                              (T*)o1, (T*)o2, (T*)o3);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "bias_elu_ln kernel: %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  return SGBPO_OK;
}
template <class T>
static int dispatch_eln(bool fwd, int W, int64_t M, const void* a0, const void* a1, const void* a2, const void* a3, void* o0,
                        void* o1, void* o2, void* o3, cudaStream_t st) {
  switch (W) {
    case 32: return launch_eln<T, 32>(fwd, M, a0, a1, a2, a3, o0, o1, o2, o3, st);
    case 64: return launch_eln<T, 64>(fwd, M, a0, a1, a2, a3, o0, o1, o2, o3, st);
    case 128: return launch_eln<T, 128>(fwd, M, a0, a1, a2, a3, o0, o1, o2, o3, st);
    default:
      snprintf(g_err, sizeof(g_err), "bias_elu_ln: widths 32, 64, 128");
      return SGBPO_E_UNSUPPORTED;
  }
}

// ---- episode scalars (fused_rollout.py graph bodies): one launch each instead of ~10 / ~7 library launches ----------
// Grid-wide deterministic sums: every CTA reduces its grid-stride share in double (shared-memory tree), stores the partial
// in a device-global slot, and the last CTA to finish (ticket counter) adds the partials in slot order and writes the
// outputs.  The scratch is one slot per calling stream (sgbpo_slots.cuh).
constexpr int EP_CTAS = 128, EP_THREADS = 256, EP_VALS = 3;
__device__ double g_ep_part[2 * N_SLOTS][EP_VALS][EP_CTAS];
__device__ unsigned int g_ep_ticket[2 * N_SLOTS];

template <int NV>
__device__ inline bool grid_sum(double (&v)[NV], int which) {      // which = 2 * slot + kernel
  __shared__ double sh[NV][EP_THREADS];
  __shared__ bool last;
#pragma unroll
  for (int q = 0; q < NV; ++q) sh[q][threadIdx.x] = v[q];
  __syncthreads();
  for (int s = EP_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int q = 0; q < NV; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) g_ep_part[which][q][blockIdx.x] = sh[q][0];
    __threadfence();
    last = atomicAdd(&g_ep_ticket[which], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  // the last CTA sums the per-CTA partials: every thread fetches one (independent loads, fixed tree order -> deterministic)
#pragma unroll
  for (int q = 0; q < NV; ++q)
    sh[q][threadIdx.x] = threadIdx.x < gridDim.x ? ((volatile double*)g_ep_part[which][q])[threadIdx.x] : 0.0;
  __syncthreads();
  for (int s = EP_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int q = 0; q < NV; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = sh[q][0];
    g_ep_ticket[which] = 0;
  }
  return threadIdx.x == 0;
}

template <class T>
__global__ void episode_stats_kernel(int64_t n_rew, const T* __restrict__ rewards, int64_t n_fl, const int32_t* __restrict__ flags,
                                     int interv_mask, int bad_mask, double* scal, int64_t* n_interv, int32_t* bad, int slot) {
  double v[3] = {0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = first; i < n_rew; i += stride) v[0] += (double)rewards[i];
  for (int64_t i = first; i < n_fl; i += stride) {
    const int32_t f = flags[i];
    v[1] += (f & interv_mask) ? 1.0 : 0.0;
    v[2] += (f & bad_mask) ? 1.0 : 0.0;
  }
  if (!grid_sum<3>(v, 2 * slot)) return;
  scal[0] = v[0];
  scal[1] = v[2] > 0.0 ? 1.0 : 0.0;
  if (n_interv) *n_interv = (int64_t)v[1];
  if (bad) *bad = v[2] > 0.0 ? 1 : 0;
}

// shac.py:131-142: loss = - sum_t disc_t * sum_n (terminal row t ? values[t][n] : rewards[t][n]) / norm
template <class T>
__global__ void shac_loss_kernel(int rows, int64_t N, const T* __restrict__ disc, const int32_t* __restrict__ term_int,
                                 const T* __restrict__ values, const T* __restrict__ rewards, const T* __restrict__ norm, T* loss, int slot) {
  double v[1] = {0.0};
  for (int t = blockIdx.x; t < rows; t += gridDim.x) {      // one CTA per buffer row: no division per element, rows of weight 0
    const double d = (double)disc[t];                        // are skipped whole, the loads of a row are independent
    if (d == 0.0) continue;
    const T* __restrict__ src = (term_int[t] != 0 ? values : rewards) + (int64_t)t * N;
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    int64_t e = threadIdx.x;
    for (; e + 3 * (int64_t)blockDim.x < N; e += 4 * (int64_t)blockDim.x) {
      const T x0 = src[e], x1 = src[e + blockDim.x], x2 = src[e + 2 * blockDim.x], x3 = src[e + 3 * blockDim.x];
      a[0] += (double)x0; a[1] += (double)x1; a[2] += (double)x2; a[3] += (double)x3;
    }
    for (; e < N; e += blockDim.x) a[0] += (double)src[e];
    v[0] += d * ((a[0] + a[1]) + (a[2] + a[3]));
  }
  if (!grid_sum<1>(v, 2 * slot + 1)) return;
  *loss = (T)(-v[0] / (double)norm[0]);
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

extern "C" int sgbpo_bias_elu_ln_fwd(int dtype, int64_t M, int width, const void* z, const void* bias, const void* gamma,
                                     const void* beta, void* h, void* e_out, void* stats, void* stream) {
  if (M <= 0) return SGBPO_OK;
  if (!z || !bias || !gamma || !beta || !h || !e_out || !stats) {
    snprintf(g_err, sizeof(g_err), "sgbpo_bias_elu_ln_fwd: null buffer");
    return SGBPO_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == SGBPO_F64 ? dispatch_eln<double>(true, width, M, z, bias, gamma, beta, h, e_out, stats, nullptr, st)
                            : dispatch_eln<float>(true, width, M, z, bias, gamma, beta, h, e_out, stats, nullptr, st);
}

extern "C" int sgbpo_bias_elu_ln_bwd(int dtype, int64_t M, int width, const void* gy, const void* e, const void* stats,
                                     const void* gamma, void* dz, void* dgamma, void* dbeta, void* dbias, void* stream) {
  if (M <= 0) return SGBPO_OK;
  if (!gy || !e || !stats || !gamma || !dz || !dgamma || !dbeta || !dbias) {
    snprintf(g_err, sizeof(g_err), "sgbpo_bias_elu_ln_bwd: null buffer");
    return SGBPO_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == SGBPO_F64 ? dispatch_eln<double>(false, width, M, gy, e, stats, gamma, dz, dgamma, dbeta, dbias, st)
                            : dispatch_eln<float>(false, width, M, gy, e, stats, gamma, dz, dgamma, dbeta, dbias, st);
}

extern "C" int sgbpo_episode_stats(int dtype, int64_t n_rewards, const void* rewards, int64_t n_flags, const int32_t* flags,
                                   int interv_mask, int bad_mask, double* scal, int64_t* n_interv, int32_t* bad, void* stream) {
  if (!rewards || !scal || (n_flags > 0 && !flags)) {
    snprintf(g_err, sizeof(g_err), "sgbpo_episode_stats: null buffer");
    return SGBPO_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SGBPO_F64)
    episode_stats_kernel<double><<<EP_CTAS, EP_THREADS, 0, st>>>(n_rewards, (const double*)rewards, n_flags, flags, interv_mask, bad_mask, scal, n_interv, bad, stream_slot(st));
  else
    episode_stats_kernel<float><<<EP_CTAS, EP_THREADS, 0, st>>>(n_rewards, (const float*)rewards, n_flags, flags, interv_mask, bad_mask, scal, n_interv, bad, stream_slot(st));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "episode_stats_kernel: %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  return SGBPO_OK;
}

extern "C" int sgbpo_shac_loss(int dtype, int rows, int64_t N, const void* disc, const int32_t* term_int, const void* values,
                               const void* rewards, const void* norm, void* loss, void* stream) {
  if (!disc || !term_int || !values || !rewards || !norm || !loss) {
    snprintf(g_err, sizeof(g_err), "sgbpo_shac_loss: null buffer");
    return SGBPO_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SGBPO_F64)
    shac_loss_kernel<double><<<EP_CTAS, EP_THREADS, 0, st>>>(rows, N, (const double*)disc, term_int, (const double*)values, (const double*)rewards,
                                                 (const double*)norm, (double*)loss, stream_slot(st));
  else
    shac_loss_kernel<float><<<EP_CTAS, EP_THREADS, 0, st>>>(rows, N, (const float*)disc, term_int, (const float*)values, (const float*)rewards,
                                                (const float*)norm, (float*)loss, stream_slot(st));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "shac_loss_kernel: %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  return SGBPO_OK;
}
