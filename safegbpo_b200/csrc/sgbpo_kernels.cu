// Per-step kernels + C ABI (include/sgbpo.h).  One thread per environment: n <= 6 states, m <= 2
// actions, all per-env matrices (B, G_s^-1 rows, duals) live in registers; global traffic is exactly
// the (N, n)/(N, m)/(N, o) rows, read/written as contiguous per-warp spans.  sm_100a only.
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

#include "../../include/sgbpo.h"
#include "sgbpo_step.cuh"
#include "sgbpo_host.cuh"

namespace sgbpo {

thread_local char g_err[512] = "";
static int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return SGBPO_E_CUDA;
}

template <int ROW, class T> __device__ __forceinline__ void load_row(const T* __restrict__ base, int64_t i, T* dst) {
#pragma unroll
  for (int q = 0; q < ROW; ++q) dst[q] = base[i * ROW + q];
}

// ---- forward ------------------------------------------------------------------------------------
template <int ENV, class T>
__global__ void __launch_bounds__(128)
step_fwd_kernel(int64_t N, const __grid_constant__ SafeConsts<T> k, const __grid_constant__ StepShared<T> sh,
                const T* __restrict__ s, const T* __restrict__ a, const T* __restrict__ w,
                const T* __restrict__ aux, const T* __restrict__ dirs, const T* __restrict__ reset_state,
                T* __restrict__ s_next, T* __restrict__ a_exec, T* __restrict__ obs, T* __restrict__ reward,
                int32_t* __restrict__ flags, int lanes) {
  using TASK = Task<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX;
  // `lanes` environments per warp (32 = one per thread; fewer for the long per-thread solves, see block_for)
  const int lane = threadIdx.x & 31;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = gw * lanes + lane; lane < lanes && i < N; i += nw * lanes) {
    T sv[NS], av[NA], wv[NS], auxv[NAUX > 0 ? NAUX : 1], dv[NA * 2 * NA], rs[NS];
    load_row<NS>(s, i, sv);
    load_row<NA>(a, i, av);
    load_row<NS>(w, i, wv);
    if (NAUX > 0) load_row<(NAUX > 0 ? NAUX : 1)>(aux, i, auxv);
    if (dirs) load_row<NA * 2 * NA>(dirs, i, dv);
    if (reset_state) load_row<NS>(reset_state, i, rs);
    T sn[NS], ae[NA], ob[NO], rew;
    int fl;
    safeguarded_step<TASK, T, T>(sv, av, wv, auxv, dirs ? dv : nullptr, &sh, k, reset_state != nullptr, rs, sn, ae,
                                 ob, rew, fl);
#pragma unroll
    for (int q = 0; q < NS; ++q) s_next[i * NS + q] = sn[q];
#pragma unroll
    for (int q = 0; q < NA; ++q) a_exec[i * NA + q] = ae[q];
#pragma unroll
    for (int q = 0; q < NO; ++q) obs[i * NO + q] = ob[q];
    reward[i] = rew;
    flags[i] = fl;
  }
}

// ---- backward: recompute with duals seeded on (s, a), contract with the cotangents -----------------
template <int ENV, class T>
__global__ void __launch_bounds__(128)
step_bwd_kernel(int64_t N, const __grid_constant__ SafeConsts<T> k, const __grid_constant__ StepShared<T> sh,
                const T* __restrict__ s, const T* __restrict__ a, const T* __restrict__ w,
                const T* __restrict__ aux, const T* __restrict__ dirs, const T* __restrict__ reset_state,
                const T* __restrict__ g_s_next, const T* __restrict__ g_a_exec, const T* __restrict__ g_obs,
                const T* __restrict__ g_reward, T* __restrict__ gs, T* __restrict__ ga, int lanes) {
  using TASK = Task<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX, K = NS + NA;
  using D = Dual<T, K>;
  // `lanes` environments per warp (32 = one per thread; fewer for the long per-thread solves, see block_for)
  const int lane = threadIdx.x & 31;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = gw * lanes + lane; lane < lanes && i < N; i += nw * lanes) {
    T sv[NS], av[NA], wv[NS], auxv[NAUX > 0 ? NAUX : 1], dv[NA * 2 * NA], rs[NS];
    load_row<NS>(s, i, sv);
    load_row<NA>(a, i, av);
    load_row<NS>(w, i, wv);
    if (NAUX > 0) load_row<(NAUX > 0 ? NAUX : 1)>(aux, i, auxv);
    if (dirs) load_row<NA * 2 * NA>(dirs, i, dv);
    if (reset_state) load_row<NS>(reset_state, i, rs);
    D sd[NS], ad[NA], sn[NS], ae[NA], ob[NO], rew;
#pragma unroll
    for (int q = 0; q < NS; ++q) sd[q] = D::seed(sv[q], q);
#pragma unroll
    for (int q = 0; q < NA; ++q) ad[q] = D::seed(av[q], NS + q);
    int fl;
    safeguarded_step<TASK, D, T>(sd, ad, wv, auxv, dirs ? dv : nullptr, &sh, k, reset_state != nullptr, rs, sn, ae,
                                 ob, rew, fl);
    T acc[K];
#pragma unroll
    for (int d = 0; d < K; ++d) acc[d] = T(0);
    if (g_s_next) {
#pragma unroll
      for (int q = 0; q < NS; ++q) {
        const T g = g_s_next[i * NS + q];
#pragma unroll
        for (int d = 0; d < K; ++d) acc[d] += g * sn[q].d[d];
      }
    }
    if (g_a_exec) {
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const T g = g_a_exec[i * NA + q];
#pragma unroll
        for (int d = 0; d < K; ++d) acc[d] += g * ae[q].d[d];
      }
    }
    if (g_obs) {
#pragma unroll
      for (int q = 0; q < NO; ++q) {
        const T g = g_obs[i * NO + q];
#pragma unroll
        for (int d = 0; d < K; ++d) acc[d] += g * ob[q].d[d];
      }
    }
    if (g_reward) {
      const T g = g_reward[i];
#pragma unroll
      for (int d = 0; d < K; ++d) acc[d] += g * rew.d[d];
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) gs[i * NS + q] = acc[q];
#pragma unroll
    for (int q = 0; q < NA; ++q) ga[i * NA + q] = acc[NS + q];
  }
}

template <int ENV, class T>
__global__ void __launch_bounds__(128)
linearise_kernel(int64_t N, const __grid_constant__ SafeConsts<T> k, const T* __restrict__ s, T* __restrict__ c,
                 T* __restrict__ A, T* __restrict__ B, T* __restrict__ E) {
  using TASK = Task<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    T sv[NS], cv[NS], Av[NS * NS], Bv[NS * NA], Ev[NS * NS];
    load_row<NS>(s, i, sv);
    linearise_full<TASK, T, T>(sv, k.noise_center, cv, Av, Bv, Ev);
#pragma unroll
    for (int q = 0; q < NS; ++q) c[i * NS + q] = cv[q];
#pragma unroll
    for (int q = 0; q < NS * NS; ++q) { A[i * NS * NS + q] = Av[q]; E[i * NS * NS + q] = Ev[q]; }
#pragma unroll
    for (int q = 0; q < NS * NA; ++q) B[i * NS * NA + q] = Bv[q];
  }
}

template <class T>
__global__ void __launch_bounds__(128)
seeker_action_set_kernel(int64_t N, const __grid_constant__ SafeConsts<T> k, const T* __restrict__ s,
                         const T* __restrict__ aux, T* __restrict__ gen, int32_t* __restrict__ status) {
  constexpr int NAUX = Task<ENV_NAVIGATE_SEEKER>::NAUX;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    T sv[2], av[NAUX], U[2][2], len[2];
    load_row<2>(s, i, sv);
    load_row<NAUX>(aux, i, av);
    const bool ok = seeker_action_set<T>(sv, av, k, U, len);
    gen[i * 4 + 0] = U[0][0] * len[0]; gen[i * 4 + 1] = U[0][1] * len[1];
    gen[i * 4 + 2] = U[1][0] * len[0]; gen[i * 4 + 3] = U[1][1] * len[1];
    status[i] = ok ? 0 : 1;
  }
}

template <class T>
__global__ void __launch_bounds__(128)
navquad_state_set_kernel(int64_t N, const __grid_constant__ SafeConsts<T> k, const T* __restrict__ s, const T* __restrict__ aux,
                         T* __restrict__ center, T* __restrict__ gen, int32_t* __restrict__ status) {
  using TASK = Task<ENV_NAVIGATE_QUADROTOR>;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    T sv[6], av[NAVQUAD_NAUX], cf[6], B[6][MAXM];
    load_row<6>(s, i, sv);
    load_row<NAVQUAD_NAUX>(aux, i, av);
    linearise_cB<TASK, T, T>(sv, k.noise_center, cf, B);
    StateDyn<T> sd;
    navquad_safe_state_set<T>(cf, B, av, k, sd);
#pragma unroll
    for (int q = 0; q < 6; ++q) center[i * 6 + q] = sd.center[q];
#pragma unroll
    for (int q = 0; q < 36; ++q) gen[i * 36 + q] = sd.gen[q];
    status[i] = sd.degenerate ? 1 : 0;
  }
}

static int env_lanes() {          // SGBPO_STEP_LANES: environments per warp of the per-step kernels (tuning)
  const char* e = getenv("SGBPO_STEP_LANES");
  const int v = e ? atoi(e) : 0;
  return (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) ? v : 0;
}

static int grid_for(int64_t N, int block) {
  int64_t g = (N + block - 1) / block;
  const int64_t cap = 148 * 16;      // grid-stride above 16 CTAs per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Threads per CTA of the per-step kernels.  The lifted program (sgbpo_lifted.cuh) is a long serial solve per thread:
// one warp per CTA spreads a few thousand environments over all SMs instead of packing them onto 32 of them.
static int block_for(int64_t N, const sgbpo_consts* c) {
  return (c->state_model == SM_LIFTED && N <= 148 * 32 * 16) ? 32 : 128;
}
// Environments per warp (SGBPO_STEP_LANES, tuning only).  Measured for the lifted solve at 4096 environments: 32 / 16 / 8 /
// 4 / 2 lanes per warp = 20.6 / 42.3 / 53.7 / 74.7 / 127.8 ms per forward step -- time follows the number of warps, so
// the dense mapping stays the default everywhere.
static int lanes_for(const sgbpo_consts* c) {
  const int forced = env_lanes();
  if (forced > 0) return forced;
  (void)c;
  return 32;
}

template <int ENV, class T> struct Launch {
  static int fwd(int64_t N, const sgbpo_consts* c, const sgbpo_step_shared* shp, const void* s, const void* a,
                 const void* w, const void* aux, const void* dirs, const void* reset_state, void* s_next,
                 void* a_exec, void* obs, void* reward, int32_t* flags, cudaStream_t st) {
    const SafeConsts<T> k = convert_consts<T>(*c);
    const StepShared<T> sh = convert_shared<T>(shp);
    const int block = block_for(N, c), lanes = lanes_for(c);
    step_fwd_kernel<ENV, T><<<grid_for(N * (32 / lanes), block), block, 0, st>>>(
        N, k, sh, (const T*)s, (const T*)a, (const T*)w, (const T*)aux, (const T*)dirs, (const T*)reset_state,
        (T*)s_next, (T*)a_exec, (T*)obs, (T*)reward, flags, lanes);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SGBPO_OK : cuda_fail(e, "step_fwd_kernel");
  }
  static int bwd(int64_t N, const sgbpo_consts* c, const sgbpo_step_shared* shp, const void* s, const void* a,
                 const void* w, const void* aux, const void* dirs, const void* reset_state, const void* g_s_next,
                 const void* g_a_exec, const void* g_obs, const void* g_reward, void* gs, void* ga, cudaStream_t st) {
    const SafeConsts<T> k = convert_consts<T>(*c);
    const StepShared<T> sh = convert_shared<T>(shp);
    const int block = block_for(N, c), lanes = lanes_for(c);
    step_bwd_kernel<ENV, T><<<grid_for(N * (32 / lanes), block), block, 0, st>>>(
        N, k, sh, (const T*)s, (const T*)a, (const T*)w, (const T*)aux, (const T*)dirs, (const T*)reset_state,
        (const T*)g_s_next, (const T*)g_a_exec, (const T*)g_obs, (const T*)g_reward, (T*)gs, (T*)ga, lanes);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SGBPO_OK : cuda_fail(e, "step_bwd_kernel");
  }
  static int lin(int64_t N, const sgbpo_consts* c, const void* s, void* cc, void* A, void* B, void* E, cudaStream_t st) {
    const SafeConsts<T> k = convert_consts<T>(*c);
    linearise_kernel<ENV, T><<<grid_for(N, 128), 128, 0, st>>>(N, k, (const T*)s, (T*)cc, (T*)A, (T*)B, (T*)E);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SGBPO_OK : cuda_fail(e, "linearise_kernel");
  }
};

#define SGBPO_DISPATCH(CALL)                                                                     \
  switch (env_id) {                                                                              \
    case ENV_BALANCE_PENDULUM: return dtype ? Launch<ENV_BALANCE_PENDULUM, double>::CALL : Launch<ENV_BALANCE_PENDULUM, float>::CALL; \
    case ENV_BALANCE_CARTPOLE: return dtype ? Launch<ENV_BALANCE_CARTPOLE, double>::CALL : Launch<ENV_BALANCE_CARTPOLE, float>::CALL; \
    case ENV_SWINGUP_CARTPOLE: return dtype ? Launch<ENV_SWINGUP_CARTPOLE, double>::CALL : Launch<ENV_SWINGUP_CARTPOLE, float>::CALL; \
    case ENV_BALANCE_QUADROTOR: return dtype ? Launch<ENV_BALANCE_QUADROTOR, double>::CALL : Launch<ENV_BALANCE_QUADROTOR, float>::CALL; \
    case ENV_MANAGE_HOUSEHOLD: return dtype ? Launch<ENV_MANAGE_HOUSEHOLD, double>::CALL : Launch<ENV_MANAGE_HOUSEHOLD, float>::CALL; \
    case ENV_NAVIGATE_SEEKER: return dtype ? Launch<ENV_NAVIGATE_SEEKER, double>::CALL : Launch<ENV_NAVIGATE_SEEKER, float>::CALL; \
    case ENV_NAVIGATE_QUADROTOR: return dtype ? Launch<ENV_NAVIGATE_QUADROTOR, double>::CALL : Launch<ENV_NAVIGATE_QUADROTOR, float>::CALL; \
    default: return fail(SGBPO_E_ARG, "unknown env_id");                                         \
  }

}  // namespace sgbpo

using namespace sgbpo;

extern "C" {

const char* sgbpo_version(void) { return "safegbpo-b200 0.1 (sm_100a)"; }
const char* sgbpo_last_error(void) { return g_err; }

int sgbpo_env_dims(int env_id, int* n, int* m, int* o, int* naux) {
#define DIMS(E) case E: *n = Task<E>::NS; *m = Task<E>::NA; *o = Task<E>::NO; *naux = Task<E>::NAUX; return SGBPO_OK;
  switch (env_id) {
    DIMS(ENV_BALANCE_PENDULUM) DIMS(ENV_BALANCE_CARTPOLE) DIMS(ENV_SWINGUP_CARTPOLE)
    DIMS(ENV_BALANCE_QUADROTOR) DIMS(ENV_MANAGE_HOUSEHOLD) DIMS(ENV_NAVIGATE_SEEKER) DIMS(ENV_NAVIGATE_QUADROTOR)
    default: return fail(SGBPO_E_ARG, "unknown env_id");
  }
#undef DIMS
}

int sgbpo_step_fwd(int env_id, int dtype, int64_t N, const sgbpo_consts* consts, const sgbpo_step_shared* shared,
                   const void* s, const void* a, const void* w, const void* aux, const void* dirs,
                   const void* reset_state, void* s_next, void* a_exec, void* obs, void* reward,
                   int32_t* flags, void* stream) {
  if (N == 0) return SGBPO_OK;
  if (!consts || !s || !a || !w || !s_next || !a_exec || !obs || !reward || !flags || N < 0)
    return fail(SGBPO_E_ARG, "sgbpo_step_fwd: null pointer or negative N");
  cudaStream_t st = (cudaStream_t)stream;
  SGBPO_DISPATCH(fwd(N, consts, shared, s, a, w, aux, dirs, reset_state, s_next, a_exec, obs, reward, flags, st))
}

int sgbpo_step_bwd(int env_id, int dtype, int64_t N, const sgbpo_consts* consts, const sgbpo_step_shared* shared,
                   const void* s, const void* a, const void* w, const void* aux, const void* dirs,
                   const void* reset_state, const void* g_s_next, const void* g_a_exec, const void* g_obs,
                   const void* g_reward, void* gs, void* ga, void* stream) {
  if (N == 0) return SGBPO_OK;
  if (!consts || !s || !a || !w || !gs || !ga || N < 0) return fail(SGBPO_E_ARG, "sgbpo_step_bwd: null pointer or negative N");
  cudaStream_t st = (cudaStream_t)stream;
  SGBPO_DISPATCH(bwd(N, consts, shared, s, a, w, aux, dirs, reset_state, g_s_next, g_a_exec, g_obs, g_reward, gs, ga, st))
}

int sgbpo_seeker_action_set(int dtype, int64_t N, const sgbpo_consts* consts, const void* s, const void* aux,
                            void* generator, int32_t* status, void* stream) {
  if (N == 0) return SGBPO_OK;
  if (!consts || !s || !aux || !generator || !status || N < 0) return fail(SGBPO_E_ARG, "sgbpo_seeker_action_set: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype) {
    seeker_action_set_kernel<double><<<grid_for(N, 128), 128, 0, st>>>(N, convert_consts<double>(*consts), (const double*)s,
                                                                      (const double*)aux, (double*)generator, status);
  } else {
    seeker_action_set_kernel<float><<<grid_for(N, 128), 128, 0, st>>>(N, convert_consts<float>(*consts), (const float*)s,
                                                                     (const float*)aux, (float*)generator, status);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail(e, "seeker_action_set_kernel");
}

int sgbpo_navquad_safe_state_set(int dtype, int64_t N, const sgbpo_consts* consts, const void* s, const void* aux,
                                 void* center, void* generator, int32_t* status, void* stream) {
  if (N == 0) return SGBPO_OK;
  if (!consts || !s || !aux || !center || !generator || !status || N < 0) return fail(SGBPO_E_ARG, "sgbpo_navquad_safe_state_set: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype) {
    navquad_state_set_kernel<double><<<grid_for(N, 128), 128, 0, st>>>(N, convert_consts<double>(*consts), (const double*)s,
                                                                      (const double*)aux, (double*)center, (double*)generator, status);
  } else {
    navquad_state_set_kernel<float><<<grid_for(N, 128), 128, 0, st>>>(N, convert_consts<float>(*consts), (const float*)s,
                                                                     (const float*)aux, (float*)center, (float*)generator, status);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail(e, "navquad_state_set_kernel");
}

int sgbpo_linearise(int env_id, int dtype, int64_t N, const sgbpo_consts* consts, const void* s, void* c, void* A,
                    void* B, void* E, void* stream) {
  if (N == 0) return SGBPO_OK;
  if (!consts || !s || !c || !A || !B || !E || N < 0) return fail(SGBPO_E_ARG, "sgbpo_linearise: null pointer or negative N");
  cudaStream_t st = (cudaStream_t)stream;
  SGBPO_DISPATCH(lin(N, consts, s, c, A, B, E, st))
}

}  // extern "C"
