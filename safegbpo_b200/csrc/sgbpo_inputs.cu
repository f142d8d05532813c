// Random inputs of one SHAC episode in ONE launch (rng_order = "batched"): the policy draws eps ~ N(0, 1) (policy.py:84-85),
// the noise samples w = c + h (2u - 1) (box.py:73-74, simulator.py:181), the ray-mask directions (ray_mask.py:184-185) and the
// reset states of the reset-all events inside the horizon (simulator.py:73 + the task's reset distribution, a zonotope
// c + G (2u - 1): rci_env / swingup_cartpole.py:62-63).  Replaces ~20 library launches per episode (normal_, uniform_,
// addcmul, the sample() arithmetic of every reset) with a counter-based generator (Philox4x32-10, the generator family torch
// itself uses on CUDA): element (t, environment) of episode e always gets the same numbers, whatever the launch shape, so the
// eager launches and the CUDA-graph replay of an episode are bit-identical.  The distributions are the reference's; the
// stream is not torch's (rng_order = "reference" keeps torch's own draws in the reference's call order for parity runs).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/sgbpo.h"

namespace sgbpo {
extern thread_local char g_err[512];

struct Philox {
  uint32_t key[2];
  __device__ __forceinline__ static void round(uint32_t (&c)[4], const uint32_t (&k)[2]) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
    uint32_t c[4] = {c0, c1, c2, c3}, k[2] = {key[0], key[1]};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, k);
      k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
};

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {        // 53 random bits in (0, 1)
  return ((double)(((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ float u01f(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// a stream of uniforms for one (element, purpose): counter = (element lo, element hi, call index | purpose << 24, episode)
template <class T> struct Uniforms {
  Philox ph;
  uint32_t e_lo, e_hi, purpose, episode, call, have;
  uint32_t buf[4];
  __device__ Uniforms(uint64_t seed, uint64_t element, uint32_t purpose_, uint32_t episode_)
      : e_lo((uint32_t)element), e_hi((uint32_t)(element >> 32)), purpose(purpose_), episode(episode_), call(0), have(0) {
    ph.key[0] = (uint32_t)seed; ph.key[1] = (uint32_t)(seed >> 32);
  }
  __device__ uint32_t bits() {
    if (have == 0) { ph(e_lo, e_hi, call | (purpose << 24), episode, buf); ++call; have = 4; }
    return buf[4 - have--];
  }
  __device__ T next();
};
template <> __device__ inline float Uniforms<float>::next() { return u01f(bits()); }
template <> __device__ inline double Uniforms<double>::next() { const uint32_t a = bits(), b = bits(); return u01(a, b); }

struct InputsArgs {
  int64_t N;
  int H, n, m, n_resets, n_reset_gens, want_dirs;
  uint64_t seed;
  const int64_t* episode;          // device: episode counter (changes between graph replays without re-capture)
  void* eps; void* noise; void* dirs; void* reset_states;
  double noise_lo[SGBPO_MAXN], noise_span[SGBPO_MAXN];
  double reset_center[SGBPO_MAXN], reset_gen[SGBPO_MAXN][SGBPO_MAX_RESET_GENS];
};

template <class T>
__global__ void __launch_bounds__(256)
episode_inputs_kernel(const __grid_constant__ InputsArgs A) {
  const uint32_t episode = (uint32_t)(*A.episode);
  const int64_t rows = (int64_t)A.H * A.N, reset_rows = (int64_t)A.n_resets * A.N;
  const int n = A.n, m = A.m;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < rows; idx += (int64_t)gridDim.x * blockDim.x) {
    {
      Uniforms<T> u(A.seed, (uint64_t)idx, 1u, episode);                // policy draws: Box-Muller
      T* out = (T*)A.eps + idx * m;
      for (int q = 0; q < m; q += 2) {
        const T u1 = u.next(), u2 = u.next();
        const T r = sqrt(T(-2) * log(u1)), ang = T(6.283185307179586476925286766559) * u2;
        out[q] = r * cos(ang);
        if (q + 1 < m) out[q + 1] = r * sin(ang);
      }
    }
    {
      Uniforms<T> u(A.seed, (uint64_t)idx, 2u, episode);                // noise samples
      T* out = (T*)A.noise + idx * n;
      for (int i = 0; i < n; ++i) out[i] = T(A.noise_lo[i]) + T(A.noise_span[i]) * u.next();
    }
    if (A.want_dirs) {                                                   // ray-mask directions: 2m unit columns of dimension m
      Uniforms<T> u(A.seed, (uint64_t)idx, 3u, episode);
      T* out = (T*)A.dirs + idx * (m * 2 * m);
      T d[SGBPO_MAXM][2 * SGBPO_MAXM];
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < 2 * m; ++j) d[i][j] = T(2) * u.next() - T(1);
      for (int j = 0; j < 2 * m; ++j) {
        T s = T(0);
        for (int i = 0; i < m; ++i) s += d[i][j] * d[i][j];
        const T inv = T(1) / sqrt(s);
        for (int i = 0; i < m; ++i) out[i * 2 * m + j] = d[i][j] * inv;
      }
    }
    if (idx < reset_rows) {                                              // reset states: c + G (2u - 1)
      Uniforms<T> u(A.seed, (uint64_t)idx, 4u, episode);
      T* out = (T*)A.reset_states + idx * n;
      T s[SGBPO_MAXN];
      for (int i = 0; i < n; ++i) s[i] = T(A.reset_center[i]);
      for (int j = 0; j < A.n_reset_gens; ++j) {
        const T c = T(2) * u.next() - T(1);
        for (int i = 0; i < n; ++i) s[i] += T(A.reset_gen[i][j]) * c;
      }
      for (int i = 0; i < n; ++i) out[i] = s[i];
    }
  }
}
}  // namespace sgbpo

using namespace sgbpo;

extern "C" int sgbpo_episode_inputs(int dtype, const sgbpo_episode_inputs_args* p, void* stream) {
  if (!p || !p->episode || !p->eps || !p->noise || (p->want_dirs && !p->dirs) || (p->n_resets > 0 && !p->reset_states)) {
    snprintf(g_err, sizeof(g_err), "sgbpo_episode_inputs: null argument");
    return SGBPO_E_ARG;
  }
  if (p->n < 1 || p->n > SGBPO_MAXN || p->m < 1 || p->m > SGBPO_MAXM || p->n_reset_gens < 0 || p->n_reset_gens > SGBPO_MAX_RESET_GENS ||
      p->n_resets < 0 || p->n_resets > p->H) {
    snprintf(g_err, sizeof(g_err), "sgbpo_episode_inputs: dimensions out of range");
    return SGBPO_E_ARG;
  }
  if (p->N <= 0 || p->H <= 0) return SGBPO_OK;
  InputsArgs a;
  a.N = p->N; a.H = p->H; a.n = p->n; a.m = p->m; a.n_resets = p->n_resets; a.n_reset_gens = p->n_reset_gens; a.want_dirs = p->want_dirs;
  a.seed = p->seed; a.episode = (const int64_t*)p->episode;
  a.eps = p->eps; a.noise = p->noise; a.dirs = p->dirs; a.reset_states = p->reset_states;
  for (int i = 0; i < SGBPO_MAXN; ++i) {
    a.noise_lo[i] = p->noise_lo[i]; a.noise_span[i] = p->noise_span[i]; a.reset_center[i] = p->reset_center[i];
    for (int j = 0; j < SGBPO_MAX_RESET_GENS; ++j) a.reset_gen[i][j] = p->reset_gen[i][j];
  }
  const int64_t rows = (int64_t)p->H * p->N;
  int64_t grid = (rows + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SGBPO_F64) episode_inputs_kernel<double><<<(int)grid, 256, 0, st>>>(a);
  else episode_inputs_kernel<float><<<(int)grid, 256, 0, st>>>(a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "episode_inputs_kernel: %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  return SGBPO_OK;
}
