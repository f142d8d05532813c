// Whole-horizon fused rollout on the 5th-generation tensor cores (tcgen05 + TMEM): the float32, width-128,
// two-hidden-layer instantiation of SHAC.collect_trajectories + the reverse pass of update_policy
// (reference: src/learning_algorithms/shac.py:96-151, components/policy.py:50-89).
//
// One CTA (512 threads) owns a tile of 128 environments for the whole horizon; tiles are dealt to persistent CTAs.
// Thread (row r = tid % 128, part p = tid / 128) owns environment r of the tile and the 32 hidden columns [32 p, 32 p + 32):
//   * row r is TMEM lane r, so `tcgen05.ld 32x32b` hands every thread ITS row segment of an accumulator; bias + ELU +
//     LayerNorm are thread-local up to an exchange of the four partial row sums through shared memory, synchronised with
//     a named barrier over the four warps that share a TMEM lane quadrant (not the whole CTA);
//   * the hidden-hidden GEMM (128 envs x 128 x 128) is 24 `tcgen05.mma.kind::f16` (fp16 hi/lo split, fp32 accumulate in
//     TMEM) issued by one thread; the A operand is written by the previous layer's epilogue straight from registers;
//   * the environment step (sgbpo_step.cuh) runs on the part-0 thread of each row with the state in registers.
//
//   rollout_tc_fwd_kernel   t = 0..H-1: policy forward -> tanh-Gaussian action -> safeguarded step; stashes the two ELU
//                           outputs + LayerNorm statistics for the reverse pass.
//   rollout_jac_kernel      one thread per (t, environment): the dual-number Jacobian of the safeguarded step w.r.t.
//                           (s_t, a_t).  It depends only on stored (s_t, a_t, w_t), so it leaves the sequential chain
//                           and runs at full occupancy over all H x N steps.
//   rollout_tc_bwd_kernel   t = H-1..0: contracts the stored Jacobian with the carried cotangents, policy backward with
//                           dh1 = dz2 W_hid and dW_hid += dz2^T h1 on the tensor cores.  Both read the SAME stored
//                           operands through K-major / MN-major descriptor views (sgbpo_tc.cuh); dW_hid accumulates in
//                           TMEM over the whole horizon under a power-of-two scale that keeps dz2 inside fp16 range.
#pragma once
#include <cuda_runtime.h>
#include <cmath>

#include "sgbpo_rollout_impl.cuh"
#include "sgbpo_tc.cuh"

namespace sgbpo {
namespace tcr {
using namespace tc;

constexpr int NSPLIT = 4;                 // threads per row
constexpr int NT = ROWS * NSPLIT;         // 512 threads per CTA
constexpr int CW = W / NSPLIT;            // 32 hidden columns per thread
constexpr int MAX_IN = 32;                // policy input width (observation + constant columns)

struct PolicyTc {                         // device pointers to the torch parameters (float32, torch layouts [out][in])
  int in_dim, out_dim;
  const float* w_in;  const float* w_hid;  const float* w_out;
  const float* b1;  const float* b2;  const float* gamma1;  const float* gamma2;  const float* beta1;  const float* beta2;
  const float* b_out;  const float* log_std;
};

struct TcExtra {
  float* jac;                // [H][N][jac_stride]  step Jacobians (rollout_jac_kernel -> rollout_tc_bwd_kernel)
  const float* reset_aux;    // [R][N][NAUX] aux in effect after reset r, or null (then reset_states[r][:, :NAUX])
  float* g_w_hid;            // [W][W] Linear2.weight.grad, accumulated with atomics (caller zero-fills)
};

template <int ENV> struct JacShape {
  using TASK = Task<ENV>;
  static constexpr int KD = TASK::NS + TASK::NA;                       // seed directions (s_t, a_t)
  static constexpr int ROWS_J = TASK::NS + TASK::NO + 1;               // outputs: s_{t+1}, obs_{t+1}, reward_t
  static constexpr int NE = ROWS_J * KD;
  static constexpr int STRIDE = (NE + 3) & ~3;                         // float4-aligned per-(t, env) record
};

__device__ __forceinline__ int quad_barrier_id(int warp) { return 1 + (warp & 3); }

// sum over the NSPLIT threads that share a row (warps q, q+4, q+8, q+12 of lane quadrant q): shared-memory exchange
__device__ __forceinline__ float row_sum4(float v, float* red, int part, int row, int bar_id) {
  red[part * ROWS + row] = v;
  bar_sync(bar_id, 128);
  return (red[row] + red[ROWS + row]) + (red[2 * ROWS + row] + red[3 * ROWS + row]);
}

// LayerNorm (eps 1e-5, biased variance, two-pass like ATen) of the row whose CW columns this thread holds; e is replaced
// by gamma * xhat + beta, the statistics are returned for the stash
__device__ __forceinline__ void normalise4(float (&e)[CW], float sum, const float* __restrict__ gamma, const float* __restrict__ beta,
                                           float* red_a, float* red_b, int part, int row, int bar_id, float& mean_out, float& rstd_out) {
  const float mean = row_sum4(sum, red_a, part, row, bar_id) / float(W);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CW; ++c) { e[c] -= mean; q += e[c] * e[c]; }
  const float rs = rsqrtf(row_sum4(q, red_b, part, row, bar_id) / float(W) + 1e-5f);
#pragma unroll
  for (int c4 = 0; c4 < CW / 4; ++c4) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + part * CW + 4 * c4);
    const float4 b = *reinterpret_cast<const float4*>(beta + part * CW + 4 * c4);
    e[4 * c4 + 0] = (e[4 * c4 + 0] * rs) * g.x + b.x;
    e[4 * c4 + 1] = (e[4 * c4 + 1] * rs) * g.y + b.y;
    e[4 * c4 + 2] = (e[4 * c4 + 2] * rs) * g.z + b.z;
    e[4 * c4 + 3] = (e[4 * c4 + 3] * rs) * g.w + b.w;
  }
  mean_out = mean;
  rstd_out = rs;
}

__device__ __forceinline__ void store_cols(float* __restrict__ dst, const float (&e)[CW]) {      // 32 contiguous floats, 16-byte aligned
#pragma unroll
  for (int c4 = 0; c4 < CW / 4; ++c4)
    *reinterpret_cast<float4*>(dst + 4 * c4) = make_float4(e[4 * c4], e[4 * c4 + 1], e[4 * c4 + 2], e[4 * c4 + 3]);
}
__device__ __forceinline__ void load_cols(const float* __restrict__ src, float (&e)[CW]) {
#pragma unroll
  for (int c4 = 0; c4 < CW / 4; ++c4) {
    const float4 v = *reinterpret_cast<const float4*>(src + 4 * c4);
    e[4 * c4] = v.x; e[4 * c4 + 1] = v.y; e[4 * c4 + 2] = v.z; e[4 * c4 + 3] = v.w;
  }
}
__device__ __forceinline__ void split_store_cols(unsigned char* hi, unsigned char* lo, int part, int row, const float (&e)[CW]) {
#pragma unroll
  for (int kc = 0; kc < CW / 8; ++kc) {
    const float h8[8] = {e[8 * kc], e[8 * kc + 1], e[8 * kc + 2], e[8 * kc + 3], e[8 * kc + 4], e[8 * kc + 5], e[8 * kc + 6], e[8 * kc + 7]};
    split_store(hi, lo, part * (CW / 8) + kc, row, h8);
  }
}
// Linear.weight [out = n][in = k] (row-major) -> fp16 hi / lo operand parts in the core-matrix layout (row = n, column = k)
__device__ __forceinline__ void load_weight_parts(unsigned char* hi, unsigned char* lo, const float* __restrict__ w, int tid) {
  for (int i = tid; i < W * KCH; i += NT) {
    const int n = i / KCH, kc = i % KCH;
    const float4 v0 = *reinterpret_cast<const float4*>(w + (size_t)n * W + kc * 8);
    const float4 v1 = *reinterpret_cast<const float4*>(w + (size_t)n * W + kc * 8 + 4);
    const float h[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    split_store(hi, lo, kc, n, h);
  }
}

// ======================================================================================================================
// forward
// ======================================================================================================================
struct FwdSmem {
  unsigned char *b_hi, *b_lo, *a_hi, *a_lo;
  float *w_in, *b1, *b2, *g1, *g2, *be1, *be2, *w_out, *misc, *obsS, *red_a, *red_b, *muS;
  uint64_t* bar;
  uint32_t* tmem_slot;
  __host__ __device__ static size_t bytes(int in_dim, int out_dim) {
    return 4 * (size_t)PART_BYTES + sizeof(float) * ((size_t)in_dim * W + 6 * W + (size_t)out_dim * W + 8 + (size_t)in_dim * ROWS +
                                                       2 * NSPLIT * ROWS + (size_t)out_dim * NSPLIT * ROWS) + 32;
  }
  __device__ void carve(unsigned char* smem, int in_dim, int out_dim) {
    b_hi = smem; b_lo = b_hi + PART_BYTES; a_hi = b_lo + PART_BYTES; a_lo = a_hi + PART_BYTES;
    float* f = reinterpret_cast<float*>(a_lo + PART_BYTES);
    w_in = f; f += in_dim * W;                 // [k][W]
    b1 = f; f += W; b2 = f; f += W; g1 = f; f += W; g2 = f; f += W; be1 = f; f += W; be2 = f; f += W;
    w_out = f; f += out_dim * W;               // [q][W]
    misc = f; f += 8;                          // b_out[0..2), sigma[0..2)
    obsS = f; f += in_dim * ROWS;              // [k][ROWS] policy input of the tile
    red_a = f; f += NSPLIT * ROWS;
    red_b = f; f += NSPLIT * ROWS;
    muS = f; f += out_dim * NSPLIT * ROWS;     // [q][part][row]
    bar = reinterpret_cast<uint64_t*>(f);
    tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  }
};

template <int ENV>
__global__ void __launch_bounds__(NT, 1)
rollout_tc_fwd_kernel(const __grid_constant__ RolloutArgs<float> A, const __grid_constant__ PolicyTc P, const __grid_constant__ TcExtra X,
                      const __grid_constant__ SafeConsts<float> k) {
  using TASK = Task<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX;
  constexpr uint32_t TMEM_COLS = 128;
  extern __shared__ __align__(1024) unsigned char smem[];
  FwdSmem S;
  const int IN = P.in_dim;
  S.carve(smem, IN, NA);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int trow = tid % ROWS, part = tid / ROWS, qbar = quad_barrier_id(warp);
  const bool owner = part == 0;

  // ---- one-time setup ---------------------------------------------------------------------------------------------
  for (int i = tid; i < IN * W; i += NT) S.w_in[(i % IN) * W + (i / IN)] = P.w_in[i];
  for (int i = tid; i < W; i += NT) {
    S.b1[i] = P.b1[i]; S.b2[i] = P.b2[i]; S.g1[i] = P.gamma1[i]; S.g2[i] = P.gamma2[i]; S.be1[i] = P.beta1[i]; S.be2[i] = P.beta2[i];
  }
  for (int i = tid; i < NA * W; i += NT) S.w_out[i] = P.w_out[i];
  if (tid < NA) { S.misc[tid] = P.b_out[tid]; S.misc[2 + tid] = expf(P.log_std[tid]); }
  load_weight_parts(S.b_hi, S.b_lo, P.w_hid, tid);
  if (tid == 0) {
    mbar_init(smem_u32(S.bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(smem_u32(S.tmem_slot));
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *S.tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(part * CW);
  const uint32_t bar_addr = smem_u32(S.bar);
  const uint32_t a_hi_addr = smem_u32(S.a_hi), a_lo_addr = smem_u32(S.a_lo), b_hi_addr = smem_u32(S.b_hi), b_lo_addr = smem_u32(S.b_lo);
  uint32_t phase = 0;
  const int64_t N = A.N;
  const int64_t n_tiles = (N + ROWS - 1) / ROWS;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t env = tile * ROWS + trow;
    const bool valid = env < N;
    float s[NS], aux[NAUX > 0 ? NAUX : 1];
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = 0.f;
    if (owner) {
#pragma unroll
      for (int i = 0; i < NS; ++i) s[i] = valid ? A.states[env * NS + i] : 0.f;
      if (NAUX > 0) {
#pragma unroll
        for (int i = 0; i < NAUX; ++i) aux[i] = valid ? A.aux0[env * NAUX + i] : 0.f;
      }
      for (int i = 0; i < IN; ++i) S.obsS[i * ROWS + trow] = valid ? A.obs[env * IN + i] : 0.f;
    }
    bar_sync(qbar, 128);

    for (int t = 0; t < A.H; ++t) {
      // this step's streaming inputs (owner threads), issued before the MLP so their latency hides behind it
      float epsv[NA], wv[NS], dv[NA * 2 * NA], rs[NS];
      const int ridx = A.reset_idx ? A.reset_idx[t] : -1;
      if (owner) {
#pragma unroll
        for (int q = 0; q < NA; ++q) epsv[q] = valid ? A.eps[((int64_t)t * N + env) * NA + q] : 0.f;
#pragma unroll
        for (int i = 0; i < NS; ++i) wv[i] = valid ? A.noise[((int64_t)t * N + env) * NS + i] : 0.f;
        if (A.dirs && valid) {
#pragma unroll
          for (int i = 0; i < NA * 2 * NA; ++i) dv[i] = A.dirs[((int64_t)t * N + env) * (NA * 2 * NA) + i];
        }
        if (ridx >= 0 && valid) {
#pragma unroll
          for (int i = 0; i < NS; ++i) rs[i] = A.reset_states[((int64_t)ridx * N + env) * NS + i];
        }
      }
      const size_t stash_off = ((size_t)t * N + (valid ? env : 0)) * W + part * CW;
      // ---- layer 1 on the FMA pipe (K = policy input width) ------------------------------------------------------------
      float e[CW];
#pragma unroll
      for (int c4 = 0; c4 < CW / 4; ++c4) {
        const float4 b4 = *reinterpret_cast<const float4*>(S.b1 + part * CW + 4 * c4);
        e[4 * c4 + 0] = b4.x; e[4 * c4 + 1] = b4.y; e[4 * c4 + 2] = b4.z; e[4 * c4 + 3] = b4.w;
      }
      for (int kk = 0; kk < IN; ++kk) {
        const float xk = S.obsS[kk * ROWS + trow];
        const float* wk = S.w_in + kk * W + part * CW;
#pragma unroll
        for (int c4 = 0; c4 < CW / 4; ++c4) {
          const float4 w = *reinterpret_cast<const float4*>(wk + 4 * c4);
          e[4 * c4 + 0] = fmaf(xk, w.x, e[4 * c4 + 0]); e[4 * c4 + 1] = fmaf(xk, w.y, e[4 * c4 + 1]);
          e[4 * c4 + 2] = fmaf(xk, w.z, e[4 * c4 + 2]); e[4 * c4 + 3] = fmaf(xk, w.w, e[4 * c4 + 3]);
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < CW; ++c) { e[c] = elu1(e[c]); sum += e[c]; }
      if (A.stash_e1 && valid) store_cols(A.stash_e1 + stash_off, e);
      float st[4];
      normalise4(e, sum, S.g1, S.be1, S.red_a, S.red_b, part, trow, qbar, st[0], st[1]);
      // ---- layer 2 on the tensor cores -----------------------------------------------------------------------------------
      split_store_cols(S.a_hi, S.a_lo, part, trow, e);
      fence_async_smem();            // generic-proxy stores -> async-proxy reads
      tc_fence_before();             // orders the previous tcgen05.ld before the next MMA
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        mma_split3<false, false>(tmem_base, a_hi_addr, a_lo_addr, b_hi_addr, b_lo_addr, 0);
        mma_commit(bar_addr);
      }
      mbar_wait(bar_addr, phase);
      phase ^= 1;
      tc_fence_after();
      {
        float z[32];
        tmem_ld32(tmem_row, z);
        sum = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < CW / 4; ++c4) {
          const float4 b4 = *reinterpret_cast<const float4*>(S.b2 + part * CW + 4 * c4);
          e[4 * c4 + 0] = elu1(z[4 * c4 + 0] + b4.x); e[4 * c4 + 1] = elu1(z[4 * c4 + 1] + b4.y);
          e[4 * c4 + 2] = elu1(z[4 * c4 + 2] + b4.z); e[4 * c4 + 3] = elu1(z[4 * c4 + 3] + b4.w);
          sum += (e[4 * c4 + 0] + e[4 * c4 + 1]) + (e[4 * c4 + 2] + e[4 * c4 + 3]);
        }
      }
      if (A.stash_e2 && valid) store_cols(A.stash_e2 + stash_off, e);
      normalise4(e, sum, S.g2, S.be2, S.red_a, S.red_b, part, trow, qbar, st[2], st[3]);
      // ---- output layer: partial dot products, exchanged over the row's four threads -----------------------------------
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        float o = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < CW / 4; ++c4) {
          const float4 w = *reinterpret_cast<const float4*>(S.w_out + q * W + part * CW + 4 * c4);
          o = fmaf(e[4 * c4 + 0], w.x, o); o = fmaf(e[4 * c4 + 1], w.y, o);
          o = fmaf(e[4 * c4 + 2], w.z, o); o = fmaf(e[4 * c4 + 3], w.w, o);
        }
        S.muS[(q * NSPLIT + part) * ROWS + trow] = o;
      }
      bar_sync(qbar, 128);
      // ---- environment step on the owner threads ---------------------------------------------------------------------------
      if (owner) {
        float a[NA];
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          const float* m = S.muS + q * NSPLIT * ROWS + trow;
          const float mu = ((m[0] + m[ROWS]) + (m[2 * ROWS] + m[3 * ROWS])) + S.misc[q];
          a[q] = valid ? tanhf(mu + S.misc[2 + q] * epsv[q]) : 0.f;
        }
        if (NAUX > 0 && ridx >= 0 && valid) {        // goal (and obstacles) in effect after the reset-all of this step
#pragma unroll
          for (int i = 0; i < NAUX; ++i)
            aux[i] = X.reset_aux ? X.reset_aux[((int64_t)ridx * N + env) * NAUX + i] : (i < NS ? rs[i] : 0.f);
        }
        float sn[NS], ae[NA], ob[NO], rew;
        int fl;
        const StepShared<float>* sh = A.shared ? A.shared + t : nullptr;
        StepShared<float> empty;
        if (!sh) { memset(&empty, 0, sizeof(empty)); sh = &empty; }
        safeguarded_step<TASK, float, float>(s, a, wv, aux, A.dirs ? dv : nullptr, sh, k, ridx >= 0, rs, sn, ae, ob, rew, fl);
#pragma unroll
        for (int i = 0; i < NS; ++i) s[i] = sn[i];
#pragma unroll
        for (int i = 0; i < NO; ++i) S.obsS[i * ROWS + trow] = ob[i];
        if (NAUX > 2) {                               // constant observation columns (obstacles) follow the kernel's NO
          for (int i = NO; i < IN; ++i) S.obsS[i * ROWS + trow] = aux[2 + (i - NO)];
        }
        if (valid) {
#pragma unroll
          for (int i = 0; i < NS; ++i) A.states[((int64_t)(t + 1) * N + env) * NS + i] = sn[i];
          for (int i = 0; i < IN; ++i) A.obs[((int64_t)(t + 1) * N + env) * IN + i] = S.obsS[i * ROWS + trow];
#pragma unroll
          for (int q = 0; q < NA; ++q) {
            A.actions[((int64_t)t * N + env) * NA + q] = a[q];
            A.exec_actions[((int64_t)t * N + env) * NA + q] = ae[q];
          }
          A.rewards[(int64_t)t * N + env] = rew;
          A.flags[(int64_t)t * N + env] = fl;
          if (A.stash_stats) *reinterpret_cast<float4*>(A.stash_stats + ((int64_t)t * N + env) * 4) = make_float4(st[0], st[1], st[2], st[3]);
        }
      }
      bar_sync(qbar, 128);          // the next policy input is visible to the row's other threads
    }
  }
  // ---- teardown ---------------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ======================================================================================================================
// step Jacobians: one thread per (t, environment)
// ======================================================================================================================
template <int ENV>
__global__ void __launch_bounds__(128)
rollout_jac_kernel(const __grid_constant__ RolloutArgs<float> A, const __grid_constant__ TcExtra X, const __grid_constant__ SafeConsts<float> k) {
  using TASK = Task<ENV>;
  using JS = JacShape<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX, KD = JS::KD;
  using D = Dual<float, KD>;
  const int64_t N = A.N, total = N * A.H;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(idx / N);
    const int64_t env = idx - (int64_t)t * N;
    float sv[NS], av[NA], wv[NS], dv[NA * 2 * NA], rs[NS], aux[NAUX > 0 ? NAUX : 1];
    const int ridx = A.reset_idx ? A.reset_idx[t] : -1;
#pragma unroll
    for (int i = 0; i < NS; ++i) { sv[i] = A.states[idx * NS + i]; wv[i] = A.noise[idx * NS + i]; rs[i] = 0.f; }
#pragma unroll
    for (int q = 0; q < NA; ++q) av[q] = A.actions[idx * NA + q];
    if (A.dirs) {
#pragma unroll
      for (int i = 0; i < NA * 2 * NA; ++i) dv[i] = A.dirs[idx * (NA * 2 * NA) + i];
    }
    if (ridx >= 0) {
#pragma unroll
      for (int i = 0; i < NS; ++i) rs[i] = A.reset_states[((int64_t)ridx * N + env) * NS + i];
    }
    if (NAUX > 0) {
      const int src = A.aux_src ? A.aux_src[t] : -1;
#pragma unroll
      for (int i = 0; i < NAUX; ++i) {
        if (src < 0) aux[i] = A.aux0[env * NAUX + i];
        else if (X.reset_aux) aux[i] = X.reset_aux[((int64_t)src * N + env) * NAUX + i];
        else aux[i] = i < NS ? A.reset_states[((int64_t)src * N + env) * NS + i] : 0.f;
      }
    }
    D sd[NS], ad[NA], sn[NS], ae[NA], ob[NO], rew;
#pragma unroll
    for (int q = 0; q < NS; ++q) sd[q] = D::seed(sv[q], q);
#pragma unroll
    for (int q = 0; q < NA; ++q) ad[q] = D::seed(av[q], NS + q);
    int fl;
    const StepShared<float>* sh = A.shared ? A.shared + t : nullptr;
    StepShared<float> empty;
    if (!sh) { memset(&empty, 0, sizeof(empty)); sh = &empty; }
    safeguarded_step<TASK, D, float>(sd, ad, wv, aux, A.dirs ? dv : nullptr, sh, k, ridx >= 0, rs, sn, ae, ob, rew, fl);
    float rec[JS::STRIDE];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int d = 0; d < KD; ++d) rec[i * KD + d] = sn[i].d[d];
#pragma unroll
    for (int i = 0; i < NO; ++i)
#pragma unroll
      for (int d = 0; d < KD; ++d) rec[(NS + i) * KD + d] = ob[i].d[d];
#pragma unroll
    for (int d = 0; d < KD; ++d) rec[(NS + NO) * KD + d] = rew.d[d];
#pragma unroll
    for (int i = JS::NE; i < JS::STRIDE; ++i) rec[i] = 0.f;
    float4* out = reinterpret_cast<float4*>(X.jac + idx * JS::STRIDE);
#pragma unroll
    for (int i = 0; i < JS::STRIDE / 4; ++i) out[i] = make_float4(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], rec[4 * i + 3]);
  }
}

// ======================================================================================================================
// backward
// ======================================================================================================================
struct BwdSmem {
  unsigned char *w_hi, *w_lo, *a_hi, *a_lo, *h_hi, *h_lo;
  float *w_in, *g1, *g2, *be1, *w_out, *misc, *obsS, *dmuS, *red_a, *red_b, *db2S;
  uint32_t* maxS;
  uint64_t* bar;
  uint32_t* tmem_slot;
  __host__ __device__ static int nop(int no) { return (no + 3) & ~3; }
  __host__ __device__ static size_t bytes(int in_dim, int no, int out_dim) {
    return 6 * (size_t)PART_BYTES + sizeof(float) * ((size_t)in_dim * W + 3 * W + (size_t)out_dim * W + 8 + (size_t)ROWS * nop(no) +
                                                       (size_t)out_dim * ROWS + 4 * NSPLIT * ROWS + W) + 8 + 32;
  }
  __device__ void carve(unsigned char* smem, int in_dim, int no, int out_dim) {
    w_hi = smem; w_lo = w_hi + PART_BYTES; a_hi = w_lo + PART_BYTES; a_lo = a_hi + PART_BYTES; h_hi = a_lo + PART_BYTES; h_lo = h_hi + PART_BYTES;
    float* f = reinterpret_cast<float*>(h_lo + PART_BYTES);
    w_in = f; f += in_dim * W;                 // [i][W]
    g1 = f; f += W; g2 = f; f += W; be1 = f; f += W;
    w_out = f; f += out_dim * W;               // [q][W]
    misc = f; f += 8;                          // sigma[0..2)
    obsS = f; f += ROWS * nop(no);             // [row][NOP] observation of the step being differentiated
    dmuS = f; f += out_dim * ROWS;             // [q][row]
    red_a = f; f += 2 * NSPLIT * ROWS;         // float2 per (part, row)
    red_b = f; f += 2 * NSPLIT * ROWS;
    db2S = f; f += W;
    maxS = reinterpret_cast<uint32_t*>(f);     // two alternating slots
    bar = reinterpret_cast<uint64_t*>(maxS + 2);
    tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  }
};

// pair of row sums over the NSPLIT threads of a row
__device__ __forceinline__ float2 row_sum4_pair(float a, float b, float* red, int part, int row, int bar_id) {
  reinterpret_cast<float2*>(red)[part * ROWS + row] = make_float2(a, b);
  bar_sync(bar_id, 128);
  const float2* r = reinterpret_cast<const float2*>(red);
  const float2 v0 = r[row], v1 = r[ROWS + row], v2 = r[2 * ROWS + row], v3 = r[3 * ROWS + row];
  return make_float2((v0.x + v1.x) + (v2.x + v3.x), (v0.y + v1.y) + (v2.y + v3.y));
}

// Column sums over the 32 rows of a warp through a warp-private 32 x 32 scratch tile (XOR-swizzled: conflict-free both ways).
// write: lane = row stores its CW values; read: lane = column.
__device__ __forceinline__ void scratch_put(float* scratch, int lane, const float (&v)[CW]) {
#pragma unroll
  for (int c = 0; c < CW; ++c) scratch[c * 32 + (lane ^ c)] = v[c];
}
__device__ __forceinline__ float scratch_get(const float* scratch, int lane, int r) { return scratch[lane * 32 + (r ^ lane)]; }

template <int ENV>
__global__ void __launch_bounds__(NT, 1)
rollout_tc_bwd_kernel(const __grid_constant__ RolloutArgs<float> A, const __grid_constant__ BackwardArgs<float> G,
                      const __grid_constant__ PolicyTc P, const __grid_constant__ TcExtra X) {
  using TASK = Task<ENV>;
  using JS = JacShape<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, KD = JS::KD;
  constexpr int NOP = (NO + 3) & ~3;
  constexpr uint32_t TMEM_COLS = 256;            // [0, 128): dh1 of the step, [128, 256): dW_hid accumulated over the horizon
  extern __shared__ __align__(1024) unsigned char smem[];
  BwdSmem S;
  const int IN = P.in_dim;
  S.carve(smem, IN, NO, NA);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int trow = tid % ROWS, part = tid / ROWS, qbar = quad_barrier_id(warp);
  const bool owner = part == 0;
  float* scratch = reinterpret_cast<float*>(S.a_hi) + warp * 1024;     // warp-private 4 KB inside the (idle) dz2 operand buffer
  float* exch = reinterpret_cast<float*>(S.a_hi);                      // [part][row][NOP] exchange of the input-gradient partials

  // ---- one-time setup ---------------------------------------------------------------------------------------------
  for (int i = tid; i < IN * W; i += NT) S.w_in[(i % IN) * W + (i / IN)] = P.w_in[i];
  for (int i = tid; i < W; i += NT) { S.g1[i] = P.gamma1[i]; S.g2[i] = P.gamma2[i]; S.be1[i] = P.beta1[i]; S.db2S[i] = 0.f; }
  for (int i = tid; i < NA * W; i += NT) S.w_out[i] = P.w_out[i];
  if (tid < NA) S.misc[tid] = expf(P.log_std[tid]);
  if (tid < 2) S.maxS[tid] = 0u;
  load_weight_parts(S.w_hi, S.w_lo, P.w_hid, tid);
  if (tid == 0) {
    mbar_init(smem_u32(S.bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(smem_u32(S.tmem_slot));
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *S.tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(part * CW);
  const uint32_t bar_addr = smem_u32(S.bar);
  const uint32_t a_hi_addr = smem_u32(S.a_hi), a_lo_addr = smem_u32(S.a_lo), w_hi_addr = smem_u32(S.w_hi), w_lo_addr = smem_u32(S.w_lo);
  const uint32_t h_hi_addr = smem_u32(S.h_hi), h_lo_addr = smem_u32(S.h_lo);
  uint32_t phase = 0;
  const int64_t N = A.N;
  const int64_t n_tiles = (N + ROWS - 1) / ROWS;

  // per-lane column accumulators (column part * 32 + lane, summed over this warp's 32 rows, all steps and tiles)
  float acc_b2 = 0.f, acc_b1 = 0.f, acc_g1 = 0.f, acc_S[NA], acc_win[NO], acc_B[NA], acc_lstd[NA];
#pragma unroll
  for (int q = 0; q < NA; ++q) { acc_S[q] = 0.f; acc_B[q] = 0.f; acc_lstd[q] = 0.f; }
#pragma unroll
  for (int i = 0; i < NO; ++i) acc_win[i] = 0.f;
  uint32_t it = 0;              // steps processed by this CTA (selects the alternating max slot)
  float scale = 1.f;            // power of two: the dz2 operand holds scale * dz2, the TMEM accumulator scale * dW_hid
  bool acc_live = false;        // the dW_hid accumulator holds data (CTA-uniform)

  // flush of the TMEM dW_hid accumulator into global memory (thread = output row c_out, its 32 input columns)
  auto flush_dw = [&](float inv_scale) {
    tc_fence_after();
    float z[32];
    tmem_ld32(tmem_row + 128, z);
    float* dst = X.g_w_hid + (size_t)trow * W + part * CW;
#pragma unroll
    for (int i = 0; i < 32; ++i) atomicAdd(dst + i, z[i] * inv_scale);
    tc_fence_before();
  };

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t env = tile * ROWS + trow;
    const bool valid = env < N;
    float gs[NS], gobs[NO];
#pragma unroll
    for (int i = 0; i < NS; ++i) gs[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NO; ++i) gobs[i] = 0.f;

    for (int t = A.H - 1; t >= 0; --t) {
      const size_t row_off = (size_t)t * N + (valid ? env : 0);
      // ---- loads of this step (latency overlaps the Jacobian contraction) -----------------------------------------------
      float e2[CW], e1[CW];
      if (valid) { load_cols(A.stash_e2 + row_off * W + part * CW, e2); load_cols(A.stash_e1 + row_off * W + part * CW, e1); }
      else {
#pragma unroll
        for (int c = 0; c < CW; ++c) { e2[c] = 0.f; e1[c] = 0.f; }
      }
      float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) st = *reinterpret_cast<const float4*>(A.stash_stats + row_off * 4);
      // ---- owner: contract the step Jacobian with the carried cotangents ------------------------------------------------
      if (owner) {
        float act[NA], epsv[NA], v[KD];
#pragma unroll
        for (int d = 0; d < KD; ++d) v[d] = 0.f;
        if (valid) {
          float rec[JS::STRIDE];
          const float4* src = reinterpret_cast<const float4*>(X.jac + row_off * JS::STRIDE);
#pragma unroll
          for (int i = 0; i < JS::STRIDE / 4; ++i) { const float4 x4 = src[i]; rec[4 * i] = x4.x; rec[4 * i + 1] = x4.y; rec[4 * i + 2] = x4.z; rec[4 * i + 3] = x4.w; }
          float go[NO];
#pragma unroll
          for (int i = 0; i < NO; ++i) go[i] = gobs[i];
          const int kt = G.term_of_row ? G.term_of_row[t + 1] : -1;
          if (kt >= 0) {
#pragma unroll
            for (int i = 0; i < NO; ++i) go[i] += G.gobs_term[((int64_t)kt * N + env) * IN + i];
          }
          const float gr = G.grw[t];
#pragma unroll
          for (int d = 0; d < KD; ++d) {
            float x = gr * rec[(NS + NO) * KD + d];
#pragma unroll
            for (int i = 0; i < NS; ++i) x = fmaf(gs[i], rec[i * KD + d], x);
#pragma unroll
            for (int i = 0; i < NO; ++i) x = fmaf(go[i], rec[(NS + i) * KD + d], x);
            v[d] = x;
          }
#pragma unroll
          for (int q = 0; q < NA; ++q) { act[q] = A.actions[row_off * NA + q]; epsv[q] = A.eps[row_off * NA + q]; }
#pragma unroll
          for (int i = 0; i < NO; ++i) S.obsS[trow * NOP + i] = A.obs[row_off * IN + i];
        } else {
#pragma unroll
          for (int q = 0; q < NA; ++q) { act[q] = 0.f; epsv[q] = 0.f; }
#pragma unroll
          for (int i = 0; i < NO; ++i) S.obsS[trow * NOP + i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) gs[i] = v[i];
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          const float ga_pre = v[NS + q] * (1.f - act[q] * act[q]);      // d loss / d (mu + sigma eps)
          S.dmuS[q * ROWS + trow] = ga_pre;
          acc_lstd[q] += ga_pre * epsv[q] * S.misc[q];
        }
      }
      bar_sync(qbar, 128);
      float dmu[NA];
#pragma unroll
      for (int q = 0; q < NA; ++q) dmu[q] = S.dmuS[q * ROWS + trow];
      const float mean1 = st.x, rstd1 = st.y, mean2 = st.z, rstd2 = st.w;
      // ---- h1 (the B operand of dW_hid) re-derived from the stash with the forward's own operations ---------------------
      {
        float h1[CW];
#pragma unroll
        for (int c4 = 0; c4 < CW / 4; ++c4) {
          const float4 g = *reinterpret_cast<const float4*>(S.g1 + part * CW + 4 * c4);
          const float4 b = *reinterpret_cast<const float4*>(S.be1 + part * CW + 4 * c4);
          h1[4 * c4 + 0] = ((e1[4 * c4 + 0] - mean1) * rstd1) * g.x + b.x;
          h1[4 * c4 + 1] = ((e1[4 * c4 + 1] - mean1) * rstd1) * g.y + b.y;
          h1[4 * c4 + 2] = ((e1[4 * c4 + 2] - mean1) * rstd1) * g.z + b.z;
          h1[4 * c4 + 3] = ((e1[4 * c4 + 3] - mean1) * rstd1) * g.w + b.w;
        }
        if (!valid) {
#pragma unroll
          for (int c = 0; c < CW; ++c) h1[c] = 0.f;
        }
        split_store_cols(S.h_hi, S.h_lo, part, trow, h1);
      }
      // ---- output layer + LayerNorm 2 + ELU 2 backward -> dz2 ---------------------------------------------------------------
      float gc[CW];           // gradient tile of this thread
      float xs[CW];           // dmu_q * xhat2 products for the column sums S_q (q = 0 first)
      {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          float dh = 0.f;
#pragma unroll
          for (int q = 0; q < NA; ++q) dh = fmaf(dmu[q], S.w_out[q * W + part * CW + c], dh);
          const float xh = (e2[c] - mean2) * rstd2;
          const float dx = dh * S.g2[part * CW + c];
          gc[c] = dx;
          s1 += dx;
          s2 = fmaf(dx, xh, s2);
        }
        const float2 tot = row_sum4_pair(s1, s2, S.red_a, part, trow, qbar);
        const float m1 = tot.x / float(W), m2 = tot.y / float(W);
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const float xh = (e2[c] - mean2) * rstd2;
          const float de = rstd2 * (gc[c] - m1 - xh * m2);
          gc[c] = de * (e2[c] > 0.f ? 1.f : e2[c] + 1.f);
        }
      }
      // ---- power-of-two scale that keeps dz2 inside fp16 range (CTA-uniform) -------------------------------------------------
      {
        float mx = 0.f;
#pragma unroll
        for (int c = 0; c < CW; ++c) mx = fmaxf(mx, fabsf(gc[c]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        uint32_t* slot = S.maxS + (it & 1);
        if (lane == 0) atomicMax(slot, __float_as_uint(mx));
        tc_fence_before();
        __syncthreads();
        const float m = __uint_as_float(*slot);
        if (tid == 0) S.maxS[(it + 1) & 1] = 0u;         // the other slot is next used two barriers from now
        ++it;
        const float prod = m * scale;
        if (m > 0.f && !(prod >= 0.25f && prod <= 8192.f)) {
          if (acc_live) flush_dw(1.f / scale);
          acc_live = false;
          int ex;
          frexpf(m, &ex);                                 // m = f * 2^ex, f in [0.5, 1)
          scale = (m - m == 0.f) ? ldexpf(1.f, 7 - ex) : 1.f;      // m * scale in [64, 128)
        }
      }
      {
        float sc[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) sc[c] = gc[c] * scale;
        split_store_cols(S.a_hi, S.a_lo, part, trow, sc);
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        // dh1[env][c_in] = sum_c_out dz2[env][c_out] W_hid[c_out][c_in]: A K-major, B = the forward weight operand viewed MN-major
        mma_split3<false, true>(tmem_base, a_hi_addr, a_lo_addr, w_hi_addr, w_lo_addr, 0);
        // dW_hid[c_out][c_in] += sum_env dz2[env][c_out] h1[env][c_in]: both operands viewed MN-major
        mma_split3<true, true>(tmem_base + 128, a_hi_addr, a_lo_addr, h_hi_addr, h_lo_addr, acc_live ? 1u : 0u);
        mma_commit(bar_addr);
      }
      acc_live = true;
      // column sums that do not need dh1 run while the MMAs execute: S_q (needs xhat2, dmu) cannot use the scratch yet (it
      // aliases the dz2 operand), so they are formed after the wait
      mbar_wait(bar_addr, phase);
      phase ^= 1;
      tc_fence_after();
      // ---- db2 and S_q column sums (scratch is free now: both MMAs have read the dz2 operand) --------------------------------
      scratch_put(scratch, lane, gc);
      __syncwarp();
      {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) a += scratch_get(scratch, lane, r);
        acc_b2 += a;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < NA; ++q) {
#pragma unroll
        for (int c = 0; c < CW; ++c) xs[c] = dmu[q] * ((e2[c] - mean2) * rstd2);
        scratch_put(scratch, lane, xs);
        __syncwarp();
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) { a += scratch_get(scratch, lane, r); b += S.dmuS[q * ROWS + (warp & 3) * 32 + r]; }
        acc_S[q] += a;
        acc_B[q] += b;
        __syncwarp();
      }
      // ---- dh1 from TMEM, LayerNorm 1 + ELU 1 backward -> dz1 ----------------------------------------------------------------
      {
        float z[32];
        tmem_ld32(tmem_row, z);
        const float inv = 1.f / scale;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const float dh = z[c] * inv;
          const float xh = (e1[c] - mean1) * rstd1;
          xs[c] = dh * xh;                               // d gamma1 contribution
          const float dx = dh * S.g1[part * CW + c];
          gc[c] = dx;
          s1 += dx;
          s2 = fmaf(dx, xh, s2);
        }
        const float2 tot = row_sum4_pair(s1, s2, S.red_b, part, trow, qbar);
        const float m1 = tot.x / float(W), m2 = tot.y / float(W);
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const float xh = (e1[c] - mean1) * rstd1;
          const float de = rstd1 * (gc[c] - m1 - xh * m2);
          gc[c] = de * (e1[c] > 0.f ? 1.f : e1[c] + 1.f);     // dz1
        }
      }
      // ---- d gamma1, db1, dW_in column sums ------------------------------------------------------------------------------------
      scratch_put(scratch, lane, xs);
      __syncwarp();
      {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) a += scratch_get(scratch, lane, r);
        acc_g1 += a;
      }
      __syncwarp();
      scratch_put(scratch, lane, gc);
      __syncwarp();
      {
        float a = 0.f;
        const float* ob = S.obsS + (size_t)((warp & 3) * 32) * NOP;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const float x = scratch_get(scratch, lane, r);
          a += x;
          float orow[NOP];
#pragma unroll
          for (int i4 = 0; i4 < NOP / 4; ++i4) {
            const float4 o4 = *reinterpret_cast<const float4*>(ob + r * NOP + 4 * i4);
            orow[4 * i4] = o4.x; orow[4 * i4 + 1] = o4.y; orow[4 * i4 + 2] = o4.z; orow[4 * i4 + 3] = o4.w;
          }
#pragma unroll
          for (int i = 0; i < NO; ++i) acc_win[i] = fmaf(x, orow[i], acc_win[i]);
        }
        acc_b1 += a;
      }
      // ---- input layer: d obs_t = dz1 W_in (partials over this thread's columns, summed by the owner) --------------------------
      bar_sync(qbar, 128);          // every warp of the quadrant is done with its scratch before the exchange overwrites the region
      {
        float p[NO];
#pragma unroll
        for (int i = 0; i < NO; ++i) {
          float a = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < CW / 4; ++c4) {
            const float4 w = *reinterpret_cast<const float4*>(S.w_in + i * W + part * CW + 4 * c4);
            a = fmaf(gc[4 * c4 + 0], w.x, a); a = fmaf(gc[4 * c4 + 1], w.y, a);
            a = fmaf(gc[4 * c4 + 2], w.z, a); a = fmaf(gc[4 * c4 + 3], w.w, a);
          }
          p[i] = a;
        }
        // the exchange region [part][row][NOP] overlaps OTHER quadrants' scratch tiles, so the whole CTA must have left the
        // column-sum phase
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NO; ++i) exch[((size_t)part * ROWS + trow) * NOP + i] = p[i];
        bar_sync(qbar, 128);
        if (owner) {
#pragma unroll
          for (int i = 0; i < NO; ++i) {
            float a = 0.f;
#pragma unroll
            for (int pp = 0; pp < NSPLIT; ++pp) a += exch[((size_t)pp * ROWS + trow) * NOP + i];
            gobs[i] = valid ? a : 0.f;
          }
        }
      }
      // the next step's stores into the dz2 operand / observation tile wait for every reader of this step
      __syncthreads();
    }
  }
  // ---- flush ------------------------------------------------------------------------------------------------------------------
  if (acc_live) flush_dw(1.f / scale);
  {
    const int c = part * CW + lane;
    atomicAdd(G.g_b_hid + c, acc_b1);
    atomicAdd(G.g_b_hid + W + c, acc_b2);
    atomicAdd(G.g_gamma + c, acc_g1);
    atomicAdd(S.db2S + c, acc_b2);
    float dg2 = 0.f, db2 = 0.f;
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const float wo = S.w_out[q * W + c];
      atomicAdd(G.g_w_out + q * W + c, P.gamma2[c] * acc_S[q] + P.beta2[c] * acc_B[q]);
      dg2 = fmaf(wo, acc_S[q], dg2);
      db2 = fmaf(wo, acc_B[q], db2);
    }
    atomicAdd(G.g_gamma + W + c, dg2);
    atomicAdd(G.g_beta + W + c, db2);
#pragma unroll
    for (int i = 0; i < NO; ++i) atomicAdd(G.g_w_in + (size_t)c * IN + i, acc_win[i]);
    if (owner) {
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        float l = acc_lstd[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        if (lane == 0) { atomicAdd(G.g_log_std + q, l); atomicAdd(G.g_b_out + q, acc_B[q]); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  // d beta1 = sum_rows dh1 = W_hid^T (sum_rows dz2): one matrix-vector product per CTA
  if (tid < W) {
    float a = 0.f;
    for (int kk = 0; kk < W; ++kk) a = fmaf(P.w_hid[(size_t)kk * W + tid], S.db2S[kk], a);
    atomicAdd(G.g_beta + tid, a);
  }
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ======================================================================================================================
// launchers
// ======================================================================================================================
inline bool tc_policy(const sgbpo_mlp* pol, PolicyTc& p) {
  if (pol->width != W || pol->n_hidden != 2 || pol->in_dim < 1 || pol->in_dim > MAX_IN || pol->out_dim < 1 || pol->out_dim > 2 || !pol->log_std)
    return false;
  p.in_dim = pol->in_dim; p.out_dim = pol->out_dim;
  p.w_in = (const float*)pol->weight[0]; p.w_hid = (const float*)pol->weight[1]; p.w_out = (const float*)pol->weight[2];
  p.b1 = (const float*)pol->bias[0]; p.b2 = (const float*)pol->bias[1]; p.b_out = (const float*)pol->bias[2];
  p.gamma1 = (const float*)pol->ln_weight[0]; p.gamma2 = (const float*)pol->ln_weight[1];
  p.beta1 = (const float*)pol->ln_bias[0]; p.beta2 = (const float*)pol->ln_bias[1];
  p.log_std = (const float*)pol->log_std;
  return true;
}
inline TcExtra tc_extra(const sgbpo_rollout* r, const sgbpo_rollout_grads* g) {
  TcExtra x;
  x.jac = (float*)r->jac;
  x.reset_aux = (const float*)r->reset_aux;
  x.g_w_hid = g ? (float*)g->g_w_hid : nullptr;
  return x;
}
inline int tc_grid(int64_t N) {
  const int64_t tiles = (N + ROWS - 1) / ROWS;
  const int sms = sm_count();
  const int64_t waves = (tiles + sms - 1) / sms;
  return (int)((tiles + waves - 1) / waves);          // even split of the tiles over persistent CTAs
}

template <int ENV> int rollout_tc_fwd(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st) {
  using TASK = Task<ENV>;
  PolicyTc p;
  if (!tc_policy(pol, p) || pol->out_dim != TASK::NA) return fail2(SGBPO_E_UNSUPPORTED, "tensor-core rollout: float32 policy [in -> 128 -> 128 -> m]");
  if (pol->in_dim < TASK::NO || (pol->in_dim > TASK::NO && TASK::NAUX <= 2)) return fail2(SGBPO_E_ARG, "policy input width does not match the environment");
  const size_t smem = FwdSmem::bytes(pol->in_dim, TASK::NA);
  if (smem > SMEM_CAP) return fail2(SGBPO_E_UNSUPPORTED, "tensor-core rollout: policy input too wide for shared memory");
  auto kern = rollout_tc_fwd_kernel<ENV>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail2(e, "cudaFuncSetAttribute(rollout_tc_fwd)");
  kern<<<tc_grid(r->N), NT, smem, st>>>(convert_rollout<float>(*r), p, tc_extra(r, nullptr), convert_consts<float>(*c));
  e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "rollout_tc_fwd_kernel");
}

template <int ENV> int rollout_jac(const sgbpo_consts* c, const sgbpo_rollout* r, cudaStream_t st) {
  if (!r->jac) return fail2(SGBPO_E_ARG, "sgbpo_rollout_jac: null Jacobian buffer");
  const int64_t total = r->N * (int64_t)r->H;
  int64_t grid = (total + 127) / 128;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  rollout_jac_kernel<ENV><<<(int)grid, 128, 0, st>>>(convert_rollout<float>(*r), tc_extra(r, nullptr), convert_consts<float>(*c));
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "rollout_jac_kernel");
}

template <int ENV> int rollout_tc_bwd(const sgbpo_mlp* pol, const sgbpo_rollout* r, const sgbpo_rollout_grads* g, cudaStream_t st) {
  using TASK = Task<ENV>;
  PolicyTc p;
  if (!tc_policy(pol, p) || pol->out_dim != TASK::NA) return fail2(SGBPO_E_UNSUPPORTED, "tensor-core rollout: float32 policy [in -> 128 -> 128 -> m]");
  if (!r->jac || !g->g_w_hid) return fail2(SGBPO_E_ARG, "tensor-core reverse pass: jac and g_w_hid buffers are required");
  const size_t smem = BwdSmem::bytes(pol->in_dim, TASK::NO, TASK::NA);
  if (smem > SMEM_CAP) return fail2(SGBPO_E_UNSUPPORTED, "tensor-core reverse pass: policy input too wide for shared memory");
  auto kern = rollout_tc_bwd_kernel<ENV>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail2(e, "cudaFuncSetAttribute(rollout_tc_bwd)");
  kern<<<tc_grid(r->N), NT, smem, st>>>(convert_rollout<float>(*r), convert_backward<float>(*g), p, tc_extra(r, g));
  e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "rollout_tc_bwd_kernel");
}

// one translation unit per environment: SGBPO_INSTANTIATE_ROLLOUT_TC(E) in sgbpo_rollout_tc_env<E>.cu
template <int ENV> int rollout_tc_fwd_env(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st);
template <int ENV> int rollout_jac_env(const sgbpo_consts* c, const sgbpo_rollout* r, cudaStream_t st);
template <int ENV> int rollout_tc_bwd_env(const sgbpo_mlp* pol, const sgbpo_rollout* r, const sgbpo_rollout_grads* g, cudaStream_t st);

#define SGBPO_INSTANTIATE_ROLLOUT_TC(E)                                                                                      \
  template <> int rollout_tc_fwd_env<E>(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st) { \
    return rollout_tc_fwd<E>(c, pol, r, st);                                                                                 \
  }                                                                                                                          \
  template <> int rollout_jac_env<E>(const sgbpo_consts* c, const sgbpo_rollout* r, cudaStream_t st) {                       \
    return rollout_jac<E>(c, r, st);                                                                                         \
  }                                                                                                                          \
  template <> int rollout_tc_bwd_env<E>(const sgbpo_mlp* pol, const sgbpo_rollout* r, const sgbpo_rollout_grads* g,           \
                                        cudaStream_t st) {                                                                   \
    return rollout_tc_bwd<E>(pol, r, g, st);                                                                                 \
  }

}  // namespace tcr
}  // namespace sgbpo
