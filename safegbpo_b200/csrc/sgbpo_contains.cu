// Containment verdicts on the device (reference: src/sets/zonotope.py:143-239 -- the verdict programs P9 -- and
// src/sets/box.py:90-102).  The reference decides `p in Z(c, G)` by  min |gamma|_inf s.t. G gamma = p - c  <= 1  and
// `Z(c_i, G_i) in Z(c, G)` by  min |[Gamma, beta]|_inf s.t. G Gamma = G_i, G beta = c - c_i  <= 1, one host LP per batch element.
// Both feasible regions are polytopes in the offset  d = p - c  (resp. c_i - c):  {d : h_k . d <= 1, k < K}.  For the point
// program the facets h_k are those of the zonotope itself and  max_k h_k . d  IS the LP value (the gauge of the zonotope, LP
// duality); for the zonotope program with fixed inner generators they are the facets of the LP's feasible region in d
// (safegbpo_b200/consts.py::polytope_facets, enumerated once on the host), so  max_k h_k . d <= 1  <=>  LP value <= 1.
// With per-environment inner generators the kernel adds the support term  sum_j |h_k . g_j|  (exact containment in the outer
// polytope; equal to the LP value when the outer zonotope is a parallelotope / box, never larger otherwise).
// One thread per environment, facets staged through shared memory in tiles; the arithmetic is done in the caller's dtype.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/sgbpo.h"

namespace sgbpo {
extern thread_local char g_err[512];

constexpr int FACET_TILE = 256;

template <class T>
__global__ void __launch_bounds__(128)
contains_kernel(int64_t N, int n, int K, int q, int center_per_env, const T* __restrict__ facets, const T* __restrict__ outer_center,
                const T* __restrict__ points, const T* __restrict__ inner_gen, T tol, T* __restrict__ value, int32_t* __restrict__ verdict) {
  __shared__ T tile[FACET_TILE * SGBPO_MAXN];
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool live = i < N;
  T d[SGBPO_MAXN];
  for (int a = 0; a < n; ++a)
    d[a] = live ? points[i * n + a] - outer_center[(center_per_env ? i * n : 0) + a] : T(0);
  T best = -INFINITY;
  for (int k0 = 0; k0 < K; k0 += FACET_TILE) {
    const int kn = min(FACET_TILE, K - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < kn * n; e += blockDim.x) tile[e] = facets[(int64_t)k0 * n + e];
    __syncthreads();
    if (!live) continue;
    for (int k = 0; k < kn; ++k) {
      const T* h = tile + k * n;
      T v = T(0);
      for (int a = 0; a < n; ++a) v = fma(h[a], d[a], v);
      if (inner_gen) {
        const T* G = inner_gen + i * (int64_t)n * q;
        for (int j = 0; j < q; ++j) {
          T s = T(0);
          for (int a = 0; a < n; ++a) s = fma(h[a], G[a * q + j], s);
          v += fabs(s);
        }
      }
      best = (v != v || best != best) ? T(NAN) : fmax(best, v);        // a non-finite input poisons the value: verdict 0
    }
  }
  if (live) {
    if (value) value[i] = best;
    if (verdict) verdict[i] = (best <= T(1) + tol) ? 1 : 0;          // NaN -> 0
  }
}
}  // namespace sgbpo

using namespace sgbpo;

extern "C" int sgbpo_contains(int dtype, const sgbpo_contains_args* p, void* stream) {
  if (!p || !p->facets || !p->outer_center || !p->points || (!p->value && !p->verdict)) {
    snprintf(g_err, sizeof(g_err), "sgbpo_contains: null argument");
    return SGBPO_E_ARG;
  }
  if (p->n < 1 || p->n > SGBPO_MAXN || p->K < 1 || p->q < 0 || p->N < 0 || (p->inner_gen && p->q < 1)) {
    snprintf(g_err, sizeof(g_err), "sgbpo_contains: dimensions out of range (n = %d, K = %d, q = %d)", p->n, p->K, p->q);
    return SGBPO_E_ARG;
  }
  if (p->N == 0) return SGBPO_OK;
  const int grid = (int)((p->N + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SGBPO_F64)
    contains_kernel<double><<<grid, 128, 0, st>>>(p->N, p->n, p->K, p->q, p->center_per_env, (const double*)p->facets,
                                                  (const double*)p->outer_center, (const double*)p->points,
                                                  (const double*)p->inner_gen, p->tol, (double*)p->value, p->verdict);
  else
    contains_kernel<float><<<grid, 128, 0, st>>>(p->N, p->n, p->K, p->q, p->center_per_env, (const float*)p->facets,
                                                 (const float*)p->outer_center, (const float*)p->points,
                                                 (const float*)p->inner_gen, (float)p->tol, (float*)p->value, p->verdict);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "contains_kernel: %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  return SGBPO_OK;
}
