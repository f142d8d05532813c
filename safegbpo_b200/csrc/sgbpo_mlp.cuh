// Warp-autonomous MLP tiles for the fused rollout (policy.py:50-89, value_function.py:38-66 of the
// reference: Linear -> ELU -> LayerNorm per hidden layer, then Linear).
//
// One warp owns R rows (environments; R = 8, 4 or 2 is a template parameter chosen at launch so that every
// SM sub-partition has several warps to interleave -- few rows per warp for small batches, 8 for large
// ones).  Lane l owns output columns {l, l+32, ...} (CPL = W/32 per lane) of every hidden layer, i.e. an
// R x CPL register tile; the R input activations of one k are one broadcast vector read from the warp's
// private shared-memory tile laid out [k][R]; the weights are
// shared by the CTA in shared memory as [k][LD] with LD = W + 1, so the forward read (lanes along
// columns) AND the transposed read of the data-gradient GEMM (lanes along k) are both conflict-free.
// LayerNorm row reductions are warp shuffles.  No __syncthreads inside the time loop.
//
// fp32 (or fp64) FMA on purpose: the parity bar is 1e-5 relative on actions/states/gradients through a
// 64-step unstable rollout; single-pass TF32/BF16 tensor-core MMA (~1e-3) does not meet it.
#pragma once
#include "sgbpo_dual.cuh"

namespace sgbpo {

constexpr int MLP_R_MAX = 8;      // rows per warp: template parameter R in {2, 4, 8}
constexpr int MAX_HIDDEN = 4;

template <class T> struct MlpParams {
  int in_dim, width, n_hidden, out_dim;
  const T* weight[MAX_HIDDEN + 1];   // torch layout [out][in]; index n_hidden = output layer
  const T* bias[MAX_HIDDEN + 1];
  const T* ln_weight[MAX_HIDDEN];
  const T* ln_bias[MAX_HIDDEN];
  const T* log_std;                  // [out_dim] (policy only, else nullptr)
};

template <class T, int W> struct NetSmem {
  static constexpr int LD = W + 1;
  T* w_in; T* w_hid; T* w_out; T* b_hid; T* gamma; T* beta; T* b_out;
  int in_dim, n_hidden, out_dim;
  __host__ __device__ static size_t floats(int in_dim, int n_hidden, int out_dim) {
    return (size_t)in_dim * LD + (size_t)(n_hidden - 1) * W * LD + (size_t)W * out_dim + 3 * (size_t)n_hidden * W + out_dim;
  }
  // carve `base` and copy the parameters in (whole CTA cooperates); returns the next free address
  __device__ T* load(T* base, const MlpParams<T>& p) {
    in_dim = p.in_dim; n_hidden = p.n_hidden; out_dim = p.out_dim;
    w_in = base; base += in_dim * LD;
    w_hid = base; base += (n_hidden - 1) * W * LD;
    w_out = base; base += W * out_dim;
    b_hid = base; base += n_hidden * W;
    gamma = base; base += n_hidden * W;
    beta = base; base += n_hidden * W;
    b_out = base; base += out_dim;
    // global reads walk torch's [out][in] layout contiguously; the transpose into [k][LD] happens on the
    // shared-memory side, where a stride of LD = W + 1 words is conflict-free
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < in_dim * W; i += nt) w_in[(i % in_dim) * LD + (i / in_dim)] = p.weight[0][i];
    for (int l = 0; l + 1 < n_hidden; ++l)
      for (int i = tid; i < W * W; i += nt) w_hid[l * W * LD + (i % W) * LD + (i / W)] = p.weight[l + 1][i];
    for (int i = tid; i < W * out_dim; i += nt) w_out[(i % W) * out_dim + (i / W)] = p.weight[n_hidden][i];
    for (int l = 0; l < n_hidden; ++l)
      for (int i = tid; i < W; i += nt) {
        b_hid[l * W + i] = p.bias[l][i];
        gamma[l * W + i] = p.ln_weight[l][i];
        beta[l * W + i] = p.ln_bias[l][i];
      }
    for (int i = tid; i < out_dim; i += nt) b_out[i] = p.bias[n_hidden][i];
    return base;
  }
};

template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ELU negative branch exactly as ATen evaluates it: exp(x) - 1 (not expm1); LayerNorm's rsqrt likewise
__device__ __forceinline__ float sg_elu_neg(float x) { return expf(x) - 1.0f; }
__device__ __forceinline__ double sg_elu_neg(double x) { return exp(x) - 1.0; }
__device__ __forceinline__ float sg_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double sg_rsqrt(double x) { return 1.0 / sqrt(x); }

// all-reduce of R independent values: the R shuffle chains are interleaved so their latencies overlap
template <class T, int R> __device__ __forceinline__ void warp_sum_rows(T (&v)[R]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T t[R];
#pragma unroll
    for (int r = 0; r < R; ++r) t[r] = __shfl_xor_sync(0xffffffffu, v[r], o);
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] += t[r];
  }
}

// R contiguous row values of one k: vector loads (the warp tiles are 32-byte aligned, R*sizeof(T) >= 8)
template <int R> __device__ __forceinline__ void load_rows(const float* __restrict__ p, float (&a)[R]) {
  if constexpr (R == 8) {
    const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
    a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w; a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
  } else if constexpr (R == 4) {
    const float4 v0 = *reinterpret_cast<const float4*>(p);
    a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
  } else {
    static_assert(R == 2, "rows per warp: 2, 4 or 8");
    const float2 v0 = *reinterpret_cast<const float2*>(p);
    a[0] = v0.x; a[1] = v0.y;
  }
}
template <int R> __device__ __forceinline__ void load_rows(const double* __restrict__ p, double (&a)[R]) {
#pragma unroll
  for (int i = 0; i < R / 2; ++i) {
    const double2 v = *reinterpret_cast<const double2*>(p + 2 * i);
    a[2 * i] = v.x; a[2 * i + 1] = v.y;
  }
}

// acc[r][j] = bias + sum_k inT[k][r] * Wsm[k][lane + 32 j]
template <class T, int W, int R>
__device__ __forceinline__ void tile_gemm(const T* __restrict__ inT, int K, const T* __restrict__ Wsm,
                                          const T* __restrict__ bias, int lane, T (&acc)[R][W / 32]) {
  constexpr int CPL = W / 32, LD = W + 1;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const T b = bias ? bias[lane + 32 * j] : T(0);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r][j] = b;
  }
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    T a[R], w[CPL];
    load_rows<R>(inT + k * R, a);
#pragma unroll
    for (int j = 0; j < CPL; ++j) w[j] = Wsm[k * LD + lane + 32 * j];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[r][j] = fma(a[r], w[j], acc[r][j]);
  }
}

// transposed product for the data gradient: out[r][j] = sum_c gT[c][r] * Wsm[lane + 32 j][c]
template <class T, int W, int R>
__device__ __forceinline__ void tile_gemm_t(const T* __restrict__ gT, const T* __restrict__ Wsm, int lane,
                                            T (&acc)[R][W / 32]) {
  constexpr int CPL = W / 32, LD = W + 1;
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r][j] = T(0);
#pragma unroll 4
  for (int c = 0; c < W; ++c) {
    T a[R], w[CPL];
    load_rows<R>(gT + c * R, a);
#pragma unroll
    for (int j = 0; j < CPL; ++j) w[j] = Wsm[(lane + 32 * j) * LD + c];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[r][j] = fma(a[r], w[j], acc[r][j]);
  }
}

// ELU + LayerNorm (eps 1e-5, biased variance) on the register tile.  z is overwritten by h.
// When SAVE, the ELU output e (the LayerNorm input), the row mean and the reciprocal standard deviation are
// kept for the backward pass: xhat = (e - mean) * rstd and dELU/dz = (e > 0 ? 1 : e + 1) are re-derived
// from e with the forward's own operations (bitwise the same values), so e is the only tile stashed.
template <class T, int W, int R, bool SAVE>
__device__ __forceinline__ void elu_layernorm(T (&z)[R][W / 32], const T* __restrict__ gamma,
                                              const T* __restrict__ beta, int lane, T (&ev)[R][W / 32],
                                              T (&mean_out)[R], T (&rstd)[R]) {
  constexpr int CPL = W / 32;
  T g[CPL], b[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { g[j] = gamma[lane + 32 * j]; b[j] = beta[lane + 32 * j]; }
  T sum[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    T s = T(0);
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const T v = z[r][j];
      const T en = sg_elu_neg(v < T(0) ? v : T(0));     // branch-free: exp() is evaluated for every element
      const T e = v > T(0) ? v : en;
      if (SAVE) ev[r][j] = e;
      z[r][j] = e;
      s += e;
    }
    sum[r] = s;
  }
  warp_sum_rows<T, R>(sum);
  T sq[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const T mean = sum[r] / T(W);
    if (SAVE) mean_out[r] = mean;
    T q = T(0);
#pragma unroll
    for (int j = 0; j < CPL; ++j) { z[r][j] -= mean; q += z[r][j] * z[r][j]; }
    sq[r] = q;
  }
  warp_sum_rows<T, R>(sq);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const T rs = sg_rsqrt(sq[r] / T(W) + T(1e-5));
    if (SAVE) rstd[r] = rs;
#pragma unroll
    for (int j = 0; j < CPL; ++j) z[r][j] = (z[r][j] * rs) * g[j] + b[j];
  }
}

// backward-side inverse of the stash: xhat and dELU/dz from the stored ELU output
template <class T, int W, int R>
__device__ __forceinline__ void unstash(const T (&ev)[R][W / 32], const T (&mean)[R], const T (&rstd)[R],
                                        T (&xhat)[R][W / 32], T (&egrad)[R][W / 32]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < W / 32; ++j) {
      const T e = ev[r][j];
      xhat[r][j] = (e - mean[r]) * rstd[r];
      egrad[r][j] = e > T(0) ? T(1) : e + T(1);
    }
}

// register tile <-> warp-private shared tile [c][R]: the R row values of one column are contiguous (vector access)
template <int R> __device__ __forceinline__ void store_rows(float* __restrict__ p, const float (&a)[R]) {
  if constexpr (R == 8) {
    *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(a[4], a[5], a[6], a[7]);
  } else if constexpr (R == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
  } else {
    *reinterpret_cast<float2*>(p) = make_float2(a[0], a[1]);
  }
}
template <int R> __device__ __forceinline__ void store_rows(double* __restrict__ p, const double (&a)[R]) {
#pragma unroll
  for (int i = 0; i < R / 2; ++i) *reinterpret_cast<double2*>(p + 2 * i) = make_double2(a[2 * i], a[2 * i + 1]);
}
template <class T, int W, int R>
__device__ __forceinline__ void store_tile(T* __restrict__ dstT, int lane, const T (&h)[R][W / 32]) {
#pragma unroll
  for (int j = 0; j < W / 32; ++j) {
    T col[R];
#pragma unroll
    for (int r = 0; r < R; ++r) col[r] = h[r][j];
    store_rows<R>(dstT + (lane + 32 * j) * R, col);
  }
}
template <class T, int W, int R>
__device__ __forceinline__ void load_tile(const T* __restrict__ srcT, int lane, T (&h)[R][W / 32]) {
#pragma unroll
  for (int j = 0; j < W / 32; ++j) {
    T col[R];
    load_rows<R>(srcT + (lane + 32 * j) * R, col);
#pragma unroll
    for (int r = 0; r < R; ++r) h[r][j] = col[r];
  }
}

// out[r][q] = b_out[q] + sum_c h[r][c] w_out[c][q]; every lane receives every (r, q)
template <class T, int W, int R, int OUT>
__device__ __forceinline__ void out_layer(const T (&h)[R][W / 32], const T* __restrict__ w_out,
                                          const T* __restrict__ b_out, int lane, T (&mu)[R][OUT]) {
#pragma unroll
  for (int q = 0; q < OUT; ++q) {
    T w[W / 32];
#pragma unroll
    for (int j = 0; j < W / 32; ++j) w[j] = w_out[(lane + 32 * j) * OUT + q];
    T p[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      p[r] = T(0);
#pragma unroll
      for (int j = 0; j < W / 32; ++j) p[r] = fma(h[r][j], w[j], p[r]);
    }
    warp_sum_rows<T, R>(p);
#pragma unroll
    for (int r = 0; r < R; ++r) mu[r][q] = p[r] + b_out[q];
  }
}

// whole-network forward for R rows whose inputs sit in inT ([in_dim][R]); bufA/bufB: warp-private
// ping-pong tiles [W][R].  Returns the last hidden activations in h.
template <class T, int W, int R>
__device__ __forceinline__ void net_forward(const NetSmem<T, W>& net, const T* inT, T* bufA, T* bufB, int lane,
                                            T (&h)[R][W / 32]) {
  T dummy_e[R][W / 32], dummy_m[R], dummy_r[R];
  tile_gemm<T, W, R>(inT, net.in_dim, net.w_in, net.b_hid, lane, h);
  elu_layernorm<T, W, R, false>(h, net.gamma, net.beta, lane, dummy_e, dummy_m, dummy_r);
  T* cur = bufA;
  T* nxt = bufB;
  for (int l = 1; l < net.n_hidden; ++l) {
    store_tile<T, W, R>(cur, lane, h);
    __syncwarp();
    tile_gemm<T, W, R>(cur, W, net.w_hid + (size_t)(l - 1) * W * (W + 1), net.b_hid + l * W, lane, h);
    elu_layernorm<T, W, R, false>(h, net.gamma + l * W, net.beta + l * W, lane, dummy_e, dummy_m, dummy_r);
    T* t = cur; cur = nxt; nxt = t;
  }
}

}  // namespace sgbpo
