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
//   policy_rows_bwd_kernel  the policy backward over ALL H x N rows as 128-row tiles (no time loop): dh1 = dz2 W_hid and
//                           dW_hid += dz2^T h1 on the tensor cores, both reading the SAME stored operands through
//                           K-major / MN-major descriptor views (sgbpo_tc.cuh); dW_hid accumulates in TMEM over all tiles of
//                           a CTA under a power-of-two scale that keeps dz2 inside fp16 range.  Run twice: Jacobian mode
//                           (d mu / d obs per row) before, gradient mode (parameter gradients) after
//   rollout_scan_kernel     the only sequential part of the reverse pass: a linear recursion of (n + o + m)-vectors per
//                           environment over t = H-1..0 (see the section comment below).
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
  float* jac;                // [H][N][jac_stride]  step Jacobians (rollout_jac_kernel -> rollout_scan_kernel)
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

// this thread's 32 columns of a stash row: contiguous (row-major stash) or eight 16-byte chunks 2 KB apart (tiled stash, where
// the 32 rows of a warp make each chunk store one contiguous 512-byte request)
__device__ __forceinline__ void store_cols_stash(float* __restrict__ base, bool tiled, int64_t row, int part, const float (&e)[CW]) {
#pragma unroll
  for (int c4 = 0; c4 < CW / 4; ++c4)
    *reinterpret_cast<float4*>(base + sgbpo::stash_index<W>(tiled, row, part * CW + 4 * c4)) =
        make_float4(e[4 * c4], e[4 * c4 + 1], e[4 * c4 + 2], e[4 * c4 + 3]);
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
    int ridx_next = A.reset_idx ? A.reset_idx[0] : -1;

    for (int t = 0; t < A.H; ++t) {
      // this step's streaming inputs (owner threads), issued before the MLP so their latency hides behind it
      float epsv[NA], wv[NS], dv[NA * 2 * NA], rs[NS];
      const int ridx = ridx_next;
      ridx_next = (A.reset_idx && t + 1 < A.H) ? A.reset_idx[t + 1] : -1;
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
      const int64_t stash_row = (int64_t)t * N + (valid ? env : 0);
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
      for (int c = 0; c < CW; ++c) { e[c] = elu1_fast(e[c]); sum += e[c]; }
      if (A.stash_e1 && valid) store_cols_stash(A.stash_e1, A.stash_tiled != 0, stash_row, part, e);
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
          e[4 * c4 + 0] = elu1_fast(z[4 * c4 + 0] + b4.x); e[4 * c4 + 1] = elu1_fast(z[4 * c4 + 1] + b4.y);
          e[4 * c4 + 2] = elu1_fast(z[4 * c4 + 2] + b4.z); e[4 * c4 + 3] = elu1_fast(z[4 * c4 + 3] + b4.w);
          sum += (e[4 * c4 + 0] + e[4 * c4 + 1]) + (e[4 * c4 + 2] + e[4 * c4 + 3]);
        }
      }
      if (A.stash_e2 && valid) store_cols_stash(A.stash_e2, A.stash_tiled != 0, stash_row, part, e);
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
// reverse pass.  The reverse-mode recursion of the rollout is LINEAR in the cotangents, and its only sequential part is
//     (gs_{t+1}, gobs_{t+1})  ->  v = J_t^T (gs_{t+1}, gobs_{t+1} + terminal term, grw_t)  ->  gs_t = v[:n],  dmu_t = v[n:] (1 - a_t^2)
//                             ->  gobs_t = M_t^T dmu_t
// with J_t the step Jacobian (rollout_jac_kernel) and M_t = d mu_t / d obs_t the policy's input-output Jacobian.  Both
// depend only on stored forward data, so they are evaluated for ALL H x N rows in parallel (policy_rows_bwd_kernel in
// Jacobian mode: tensor cores over 128-row tiles), the recursion shrinks to a scan of (n + o + m)-vectors per environment
// (rollout_scan_kernel), and the parameter gradients become ONE batched pass over the H x N rows with the now-known
// cotangents dmu_t (policy_rows_bwd_kernel in gradient mode: dW_hid accumulates in TMEM over all rows of a CTA).
// ======================================================================================================================
// NSP = threads per row (4 or 8: 512 or 1024 threads per CTA, 32 or 16 hidden columns per thread).  The kernel is a chain of
// short dependent phases separated by barriers, so the 8-way split (twice the warps, half the registers per thread) hides
// more latency; the 4-way split remains for wide policy inputs whose exchange buffers would not fit.
template <int NSP> struct BwdSmem {
  unsigned char *w_hi, *w_lo, *a_hi, *a_lo, *h_hi, *h_lo;
  float *w_in, *g1, *g2, *be1, *w_out, *obsS, *dmuS, *red_a, *red_b, *db2S;
  uint32_t* maxS;
  uint64_t* bar;
  uint64_t* tma_bar;
  uint32_t* tmem_slot;
  float* stage;              // (Jacobian mode) the NEXT tile of the e2 stash, landed by ONE 64 KB TMA bulk copy while this tile computes
                             // (the stash is stored in 128-row tiles of 16-byte column chunks, sgbpo::stash_index: a tile is contiguous in
                             // HBM and, staged as it is, a warp's 16-byte reads of one chunk are 512 contiguous bytes -> conflict-free)
  __host__ __device__ static int nop(int no) { return (no + 3) & ~3; }
  __host__ __device__ static size_t bytes(int in_dim, int no, int out_dim, bool grad) {
    return (grad ? 6 : 4) * (size_t)PART_BYTES + sizeof(float) * ((size_t)in_dim * W + 3 * W + (size_t)out_dim * W + (size_t)ROWS * nop(no) +
                                                                     (size_t)out_dim * ROWS + 2 * NSP * ROWS + W) + 8 + 48 +
           (grad ? 0 : 128 + sizeof(float) * (size_t)ROWS * W);
  }
  __device__ void carve(unsigned char* smem, int in_dim, int no, int out_dim, bool grad) {
    w_hi = smem; w_lo = w_hi + PART_BYTES; a_hi = w_lo + PART_BYTES; a_lo = a_hi + PART_BYTES;
    h_hi = a_lo + PART_BYTES; h_lo = h_hi + PART_BYTES;            // (gradient mode only)
    float* f = reinterpret_cast<float*>(grad ? h_lo + PART_BYTES : h_hi);
    w_in = f; f += in_dim * W;                 // [i][W]
    g1 = f; f += W; g2 = f; f += W; be1 = f; f += W;
    w_out = f; f += out_dim * W;               // [q][W]
    obsS = f; f += ROWS * nop(no);             // [row][NOP] policy input of the tile's rows (gradient mode)
    dmuS = f; f += out_dim * ROWS;             // [q][row]
    red_a = f; f += 2 * NSP * ROWS;            // float2 per (part, row); the two row-sum exchanges of a pass are separated by
    red_b = red_a;                             // CTA barriers, so they share the buffer
    db2S = f; f += W;
    maxS = reinterpret_cast<uint32_t*>(f);     // two alternating slots
    bar = reinterpret_cast<uint64_t*>(maxS + 2);
    tma_bar = bar + 1;
    tmem_slot = reinterpret_cast<uint32_t*>(tma_bar + 1);
    stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 127) & ~uintptr_t(127));  // bulk-copy destination
  }
};

// pair of row sums over the NSP threads of a row
template <int NSP>
__device__ __forceinline__ float2 row_sum_pair(float a, float b, float* red, int part, int row, int bar_id) {
  reinterpret_cast<float2*>(red)[part * ROWS + row] = make_float2(a, b);
  bar_sync(bar_id, 32 * NSP);
  const float2* r = reinterpret_cast<const float2*>(red);
  float2 t = make_float2(0.f, 0.f);
#pragma unroll
  for (int p = 0; p < NSP; ++p) { const float2 v = r[p * ROWS + row]; t.x += v.x; t.y += v.y; }
  return t;
}

// this thread's C columns [part * C, +C) of row `trow` of a stash tile (global or staged in shared memory): chunk stride 2 KB
template <int C> __device__ __forceinline__ void load_cols_n(const float* __restrict__ tile, int trow, int part, float (&e)[C]) {
#pragma unroll
  for (int c4 = 0; c4 < C / 4; ++c4) {
    const float4 v = *reinterpret_cast<const float4*>(tile + ((size_t)(part * (C / 4) + c4) * ROWS + trow) * 4);
    e[4 * c4] = v.x; e[4 * c4 + 1] = v.y; e[4 * c4 + 2] = v.z; e[4 * c4 + 3] = v.w;
  }
}
template <int C> __device__ __forceinline__ void split_store_cols_n(unsigned char* hi, unsigned char* lo, int col0, int row, const float (&e)[C]) {
#pragma unroll
  for (int kc = 0; kc < C / 8; ++kc) {
    const float h8[8] = {e[8 * kc], e[8 * kc + 1], e[8 * kc + 2], e[8 * kc + 3], e[8 * kc + 4], e[8 * kc + 5], e[8 * kc + 6], e[8 * kc + 7]};
    split_store(hi, lo, col0 / 8 + kc, row, h8);
  }
}
template <int C> __device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float (&v)[C]);
template <> __device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Column sums over the 32 rows of a warp through a warp-private C x 32 scratch tile, XOR-swizzled so that both the write
// (lane = row stores its C values) and the read (lane = column lane % C, rows [(lane / C) * C, +C)) are conflict-free.
// With C = 16 the two half-warps sum disjoint row halves of the same 16 columns; the partial sums are only ever ADDED
// (register accumulators flushed with atomics), so they are never combined explicitly.
template <int C> __device__ __forceinline__ void scratch_put_n(float* scratch, int lane, const float (&v)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) scratch[c * 32 + (lane ^ c)] = v[c];
}
template <int C> __device__ __forceinline__ float scratch_get_n(const float* scratch, int lane, int j) {
  const int col = lane % C, row = (lane / C) * C + j;
  return scratch[col * 32 + (row ^ col)];
}
template <int C> __device__ __forceinline__ float scratch_colsum_n(const float* scratch, int lane) {      // four independent chains
  float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < C; ++j) a[j & 3] += scratch_get_n<C>(scratch, lane, j);
  return (a[0] + a[1]) + (a[2] + a[3]);
}

struct RowsBwdArgs {
  int64_t M;                 // rows = H * N, row = t * N + env
  const float* e1;  const float* e2;  const float* stats;      // forward stash [M][W], [M][W], [M][4]
  const float* obs;          // [M (+ N)][in_dim] policy inputs (rollout buffer rows 0..H-1)       (gradient mode)
  const float* dmu;          // [M][out_dim] d loss / d mu                                         (gradient mode)
  float* pol_jac;            // [M][out_dim][NOP] d mu_q / d obs                                   (Jacobian mode)
  float* g_w_hid;  float* g_w_in;  float* g_w_out;  float* g_b_hid;  float* g_gamma;  float* g_beta;     // accumulated (atomics)
};

// GRAD = false: policy input-output Jacobian of every row (cotangent e_q per output q).  GRAD = true: parameter gradients of
// the batched policy for the row cotangents dmu.  NA = policy outputs, NO = differentiated inputs (the first NO of in_dim).
// NIN >= in_dim: compile-time bound on the policy input width (NavigateSeeker appends constant obstacle columns to the NO
// differentiated observations; they carry no input gradient but they do enter Linear1.weight.grad).
template <int NA, int NO, int NIN, bool GRAD, int NSP>
__global__ void __launch_bounds__(ROWS * NSP, 1)
policy_rows_bwd_kernel(const __grid_constant__ RowsBwdArgs R, const __grid_constant__ PolicyTc P) {
  constexpr int NOP = (NO + 3) & ~3, NIP = (NIN + 3) & ~3, C = W / NSP, NTK = ROWS * NSP;
  constexpr uint32_t TMEM_COLS = GRAD ? 256 : 128;       // [0, 128): dh1 of the tile; [128, 256): dW_hid over all tiles of this CTA
  extern __shared__ __align__(1024) unsigned char smem[];
  BwdSmem<NSP> S;
  const int IN = P.in_dim;
  S.carve(smem, IN, NIN, NA, GRAD);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int trow = tid % ROWS, part = tid / ROWS, qbar = quad_barrier_id(warp);
  const bool owner = part == 0;
  float* scratch = reinterpret_cast<float*>(S.a_hi) + warp * (C * 32);   // warp-private tile inside the (idle) dz2 operand buffer
  float* exch = reinterpret_cast<float*>(S.a_hi);                        // [part][row][NOP] exchange of the input-gradient partials
  static_assert(!GRAD || true, "");
  // this lane's column and row subset in the column-sum phases
  const int cs_col = part * C + lane % C, cs_row0 = (warp & 3) * 32 + (lane / C) * C;

  // ---- one-time setup ---------------------------------------------------------------------------------------------
  for (int i = tid; i < IN * W; i += NTK) S.w_in[(i % IN) * W + (i / IN)] = P.w_in[i];
  for (int i = tid; i < W; i += NTK) { S.g1[i] = P.gamma1[i]; S.g2[i] = P.gamma2[i]; S.be1[i] = P.beta1[i]; S.db2S[i] = 0.f; }
  for (int i = tid; i < NA * W; i += NTK) S.w_out[i] = P.w_out[i];
  if (tid < 2) S.maxS[tid] = 0u;
  for (int i = tid; i < W * KCH; i += NTK) {
    const int n = i / KCH, kc = i % KCH;
    const float4 v0 = *reinterpret_cast<const float4*>(P.w_hid + (size_t)n * W + kc * 8);
    const float4 v1 = *reinterpret_cast<const float4*>(P.w_hid + (size_t)n * W + kc * 8 + 4);
    const float h[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    split_store(S.w_hi, S.w_lo, kc, n, h);
  }
  if (tid == 0) {
    mbar_init(smem_u32(S.bar), 1);
    mbar_init(smem_u32(S.tma_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(smem_u32(S.tmem_slot));
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *S.tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(part * C);
  const uint32_t bar_addr = smem_u32(S.bar), tma_bar_addr = smem_u32(S.tma_bar);
  const uint32_t a_hi_addr = smem_u32(S.a_hi), a_lo_addr = smem_u32(S.a_lo), w_hi_addr = smem_u32(S.w_hi), w_lo_addr = smem_u32(S.w_lo);
  const uint32_t h_hi_addr = smem_u32(S.h_hi), h_lo_addr = smem_u32(S.h_lo);
  uint32_t phase = 0, tma_phase = 0;
  const int64_t n_tiles = (R.M + ROWS - 1) / ROWS;
  // TMA staging (Jacobian mode): thread 0 arms the transaction barrier and issues one bulk copy of the whole 64 KB e2 tile (the
  // stash buffers are padded to whole tiles).  Issued for tile i + 1 as soon as every thread holds tile i's e2 in registers, so
  // the copy runs under tile i's arithmetic and MMA.
  auto stage_tile = [&](int64_t tile) {
    if (tid != 0) return;
    mbar_expect_tx(tma_bar_addr, (uint32_t)(ROWS * W * 4));
    bulk_g2s(smem_u32(S.stage), R.e2 + (size_t)tile * ROWS * W, ROWS * W * 4, tma_bar_addr);
  };
  if (!GRAD && (int64_t)blockIdx.x < n_tiles) stage_tile(blockIdx.x);

  // per-lane column accumulators (column cs_col, summed over this lane's row subset and all tiles of this CTA)
  float acc_b2 = 0.f, acc_b1 = 0.f, acc_g1 = 0.f, acc_S[NA], acc_win[NIN], acc_B[NA];
#pragma unroll
  for (int q = 0; q < NA; ++q) { acc_S[q] = 0.f; acc_B[q] = 0.f; }
#pragma unroll
  for (int i = 0; i < NIN; ++i) acc_win[i] = 0.f;
  uint32_t it = 0;              // (tile, cotangent) passes processed by this CTA (selects the alternating max slot)
  float scale = 1.f;            // power of two: the dz2 operand holds scale * dz2, the TMEM accumulator scale * dW_hid
  bool acc_live = false;        // the dW_hid accumulator holds data (CTA-uniform)

  auto flush_dw = [&](float inv_scale) {      // TMEM dW_hid accumulator -> global (thread = output row c_out, its C input columns)
    tc_fence_after();
    float z[C];
    tmem_ld_cols<C>(tmem_row + 128, z);
    float* dst = R.g_w_hid + (size_t)trow * W + part * C;
#pragma unroll
    for (int i = 0; i < C; i += 4)          // 16-byte vector reductions: a quarter of the atomic requests of every CTA's flush
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(z[i] * inv_scale), "f"(z[i + 1] * inv_scale),
                   "f"(z[i + 2] * inv_scale), "f"(z[i + 3] * inv_scale) : "memory");
    tc_fence_before();
  };

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * ROWS + trow;
    const bool valid = row < R.M;
    if (tid == 0 && tile + gridDim.x < n_tiles) {      // the next tile's stash blocks (64 KB contiguous each) on their way into L2
      bulk_prefetch_l2(R.e1 + (size_t)(tile + gridDim.x) * ROWS * W, ROWS * W * 4);
      if (GRAD) bulk_prefetch_l2(R.e2 + (size_t)(tile + gridDim.x) * ROWS * W, ROWS * W * 4);
    }
    const size_t row_off = valid ? (size_t)row : 0;
    float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) st = *reinterpret_cast<const float4*>(R.stats + row_off * 4);
    const float mean1 = st.x, rstd1 = st.y, mean2 = st.z, rstd2 = st.w;
    float e2[C];
#pragma unroll
    for (int c = 0; c < C; ++c) e2[c] = 0.f;
    if (GRAD) {
      if (valid) load_cols_n<C>(R.e2 + (size_t)tile * ROWS * W, trow, part, e2);
    } else {
      mbar_wait(tma_bar_addr, tma_phase);            // this tile's rows have landed in the stage
      tma_phase ^= 1;
      if (valid) load_cols_n<C>(S.stage, trow, part, e2);
      fence_async_smem();                            // generic-proxy reads of the stage before the async-proxy refill
      __syncthreads();
      if (tile + gridDim.x < n_tiles) stage_tile(tile + gridDim.x);
    }
    if (GRAD) {
      if (owner) {
#pragma unroll
        for (int i = 0; i < NIN; ++i) S.obsS[trow * NIP + i] = (valid && i < IN) ? R.obs[row_off * IN + i] : 0.f;
#pragma unroll
        for (int q = 0; q < NA; ++q) S.dmuS[q * ROWS + trow] = valid ? R.dmu[row_off * NA + q] : 0.f;
      }
      // h1 (the B operand of dW_hid) re-derived from the stash with the forward's own operations
      float h1[C];
#pragma unroll
      for (int c = 0; c < C; ++c) h1[c] = 0.f;
      if (valid) load_cols_n<C>(R.e1 + (size_t)tile * ROWS * W, trow, part, h1);
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 g = *reinterpret_cast<const float4*>(S.g1 + part * C + 4 * c4);
        const float4 b = *reinterpret_cast<const float4*>(S.be1 + part * C + 4 * c4);
        h1[4 * c4 + 0] = valid ? ((h1[4 * c4 + 0] - mean1) * rstd1) * g.x + b.x : 0.f;
        h1[4 * c4 + 1] = valid ? ((h1[4 * c4 + 1] - mean1) * rstd1) * g.y + b.y : 0.f;
        h1[4 * c4 + 2] = valid ? ((h1[4 * c4 + 2] - mean1) * rstd1) * g.z + b.z : 0.f;
        h1[4 * c4 + 3] = valid ? ((h1[4 * c4 + 3] - mean1) * rstd1) * g.w + b.w : 0.f;
      }
      split_store_cols_n<C>(S.h_hi, S.h_lo, part * C, trow, h1);
      bar_sync(qbar, 32 * NSP);
    }
    constexpr int PASSES = GRAD ? 1 : NA;
#pragma unroll 1
    for (int pass = 0; pass < PASSES; ++pass) {
      float dmu[NA];
#pragma unroll
      for (int q = 0; q < NA; ++q) dmu[q] = GRAD ? S.dmuS[q * ROWS + trow] : (q == pass ? 1.f : 0.f);
      if (GRAD) {
        // ---- S_q[c] = sum_rows dmu_q xhat2[c] column sums.  The scratch aliases the dz2 operand, which is idle here: the
        //      previous tile's MMAs have completed and this tile's dz2 is stored after the next CTA barrier
#pragma unroll
        for (int q = 0; q < NA; ++q) {
#pragma unroll
          for (int c = 0; c < C; ++c) scratch[c * 32 + (lane ^ c)] = dmu[q] * ((e2[c] - mean2) * rstd2);
          __syncwarp();
          float b = 0.f;
#pragma unroll
          for (int r4 = 0; r4 < C / 4; ++r4) {
            const float4 d4 = *reinterpret_cast<const float4*>(S.dmuS + q * ROWS + cs_row0 + 4 * r4);
            b += (d4.x + d4.y) + (d4.z + d4.w);
          }
          acc_S[q] += scratch_colsum_n<C>(scratch, lane);
          acc_B[q] += b;
          __syncwarp();
        }
      }
      // ---- output layer + LayerNorm 2 + ELU 2 backward -> dz2 -----------------------------------------------------------
      float gc[C];
      {
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
          const float4 g4 = *reinterpret_cast<const float4*>(S.g2 + part * C + 4 * c4);
          float dh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int q = 0; q < NA; ++q) {
            const float4 w4 = *reinterpret_cast<const float4*>(S.w_out + q * W + part * C + 4 * c4);
            dh[0] = fmaf(dmu[q], w4.x, dh[0]); dh[1] = fmaf(dmu[q], w4.y, dh[1]);
            dh[2] = fmaf(dmu[q], w4.z, dh[2]); dh[3] = fmaf(dmu[q], w4.w, dh[3]);
          }
          const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xh = (e2[4 * c4 + j] - mean2) * rstd2;
            const float dx = dh[j] * gg[j];
            gc[4 * c4 + j] = dx;
            s1[j] += dx;
            s2[j] = fmaf(dx, xh, s2[j]);
          }
        }
        const float2 tot = row_sum_pair<NSP>((s1[0] + s1[1]) + (s1[2] + s1[3]), (s2[0] + s2[1]) + (s2[2] + s2[3]), S.red_a, part, trow, qbar);
        const float m1 = tot.x / float(W), m2 = tot.y / float(W);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float xh = (e2[c] - mean2) * rstd2;
          const float de = rstd2 * (gc[c] - m1 - xh * m2);
          gc[c] = valid ? de * (e2[c] > 0.f ? 1.f : e2[c] + 1.f) : 0.f;
        }
      }
      if (GRAD) {                // db2 column sums (the scratch is still idle)
        scratch_put_n<C>(scratch, lane, gc);
        __syncwarp();
        acc_b2 += scratch_colsum_n<C>(scratch, lane);
      }
      // ---- power-of-two scale that keeps dz2 inside fp16 range (CTA-uniform) ---------------------------------------------
      {
        float mx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mx[j] = fmaxf(mx[j], fabsf(gc[4 * c4 + j]));
        }
        float mxx = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        uint32_t* slot = S.maxS + (it & 1);
        if (lane == 0) atomicMax(slot, __float_as_uint(mxx));
        __syncthreads();           // (also: every warp is done reading its scratch tile before dz2 is stored over it)
        const float m = __uint_as_float(*slot);
        if (tid == 0) S.maxS[(it + 1) & 1] = 0u;         // the other slot is next used two barriers from now
        ++it;
        const float prod = m * scale;
        if (m > 0.f && !(prod >= 0.25f && prod <= 8192.f)) {
          if (GRAD && acc_live) flush_dw(1.f / scale);
          acc_live = false;
          int ex;
          frexpf(m, &ex);                                 // m = f * 2^ex, f in [0.5, 1)
          scale = (m - m == 0.f) ? ldexpf(1.f, 7 - ex) : 1.f;      // m * scale in [64, 128)
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) gc[c] *= scale;
      split_store_cols_n<C>(S.a_hi, S.a_lo, part * C, trow, gc);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        // dh1[row][c_in] = sum_c_out dz2[row][c_out] W_hid[c_out][c_in]: A K-major, B = the forward weight operand viewed MN-major
        mma_split3<false, true>(tmem_base, a_hi_addr, a_lo_addr, w_hi_addr, w_lo_addr, 0);
        if (GRAD)   // dW_hid[c_out][c_in] += sum_rows dz2[row][c_out] h1[row][c_in]: both operands viewed MN-major
          mma_split3<true, true>(tmem_base + 128, a_hi_addr, a_lo_addr, h_hi_addr, h_lo_addr, acc_live ? 1u : 0u);
        mma_commit(bar_addr);
      }
      if (GRAD) acc_live = true;
      float e1[C];              // (loaded while the MMAs execute)
#pragma unroll
      for (int c = 0; c < C; ++c) e1[c] = 0.f;
      if (valid) load_cols_n<C>(R.e1 + (size_t)tile * ROWS * W, trow, part, e1);
      mbar_wait(bar_addr, phase);
      phase ^= 1;
      tc_fence_after();
      // ---- dh1 from TMEM, LayerNorm 1 + ELU 1 backward -> dz1 (in place) ---------------------------------------------------
      tmem_ld_cols<C>(tmem_row, gc);
      {
        const float inv = 1.f / scale;
        if (GRAD) {              // d gamma1[c] = sum_rows dh1 xhat1 (the scratch is free again: both MMAs have read the dz2 operand)
#pragma unroll
          for (int c = 0; c < C; ++c) scratch[c * 32 + (lane ^ c)] = (gc[c] * inv) * ((e1[c] - mean1) * rstd1);
          __syncwarp();
          acc_g1 += scratch_colsum_n<C>(scratch, lane);
          __syncwarp();
        }
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
          const float4 g4 = *reinterpret_cast<const float4*>(S.g1 + part * C + 4 * c4);
          const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xh = (e1[4 * c4 + j] - mean1) * rstd1;
            const float dx = (gc[4 * c4 + j] * inv) * gg[j];
            gc[4 * c4 + j] = dx;
            s1[j] += dx;
            s2[j] = fmaf(dx, xh, s2[j]);
          }
        }
        const float2 tot = row_sum_pair<NSP>((s1[0] + s1[1]) + (s1[2] + s1[3]), (s2[0] + s2[1]) + (s2[2] + s2[3]), S.red_b, part, trow, qbar);
        const float m1 = tot.x / float(W), m2 = tot.y / float(W);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float xh = (e1[c] - mean1) * rstd1;
          const float de = rstd1 * (gc[c] - m1 - xh * m2);
          gc[c] = valid ? de * (e1[c] > 0.f ? 1.f : e1[c] + 1.f) : 0.f;     // dz1
        }
      }
      if (GRAD) {
        // ---- db1, dW_in column sums ------------------------------------------------------------------------------------------
        scratch_put_n<C>(scratch, lane, gc);
        __syncwarp();
        {
          float a[2] = {0.f, 0.f};
          const float* ob = S.obsS + (size_t)cs_row0 * NIP;
#pragma unroll 4
          for (int j = 0; j < C; ++j) {
            const float x = scratch_get_n<C>(scratch, lane, j);
            a[j & 1] += x;
            float orow[NIP];
#pragma unroll
            for (int i4 = 0; i4 < NIP / 4; ++i4) {
              const float4 o4 = *reinterpret_cast<const float4*>(ob + j * NIP + 4 * i4);
              orow[4 * i4] = o4.x; orow[4 * i4 + 1] = o4.y; orow[4 * i4 + 2] = o4.z; orow[4 * i4 + 3] = o4.w;
            }
#pragma unroll
            for (int i = 0; i < NIN; ++i) acc_win[i] = fmaf(x, orow[i], acc_win[i]);
          }
          acc_b1 += a[0] + a[1];
        }
      } else {
        // ---- input layer: d mu_q / d obs = dz1 W_in (partials over this thread's columns, summed by the owner) ----------------
        float p[NO];
#pragma unroll
        for (int i = 0; i < NO; ++i) {
          float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 w = *reinterpret_cast<const float4*>(S.w_in + i * W + part * C + 4 * c4);
            a[0] = fmaf(gc[4 * c4 + 0], w.x, a[0]); a[1] = fmaf(gc[4 * c4 + 1], w.y, a[1]);
            a[2] = fmaf(gc[4 * c4 + 2], w.z, a[2]); a[3] = fmaf(gc[4 * c4 + 3], w.w, a[3]);
          }
          p[i] = (a[0] + a[1]) + (a[2] + a[3]);
        }
        // (the exchange region aliases the dz2 operand: the MMA that read it has completed)
#pragma unroll
        for (int i = 0; i < NO; ++i) exch[((size_t)part * ROWS + trow) * NOP + i] = p[i];
        bar_sync(qbar, 32 * NSP);
        if (owner && valid) {
          float* dst = R.pol_jac + ((size_t)row * NA + pass) * NOP;
#pragma unroll
          for (int i = 0; i < NOP; ++i) {
            float a = 0.f;
            if (i < NO) {
#pragma unroll
              for (int pp = 0; pp < NSP; ++pp) a += exch[((size_t)pp * ROWS + trow) * NOP + i];
            }
            dst[i] = a;
          }
        }
      }
      // the next pass / tile stores into the dz2 operand, the observation tile and the exchange region
      __syncthreads();
    }
  }
  // ---- flush ------------------------------------------------------------------------------------------------------------------
  if (GRAD) {
    if (acc_live) flush_dw(1.f / scale);
    const int c = cs_col;
    atomicAdd(R.g_b_hid + c, acc_b1);
    atomicAdd(R.g_b_hid + W + c, acc_b2);
    atomicAdd(R.g_gamma + c, acc_g1);
    atomicAdd(S.db2S + c, acc_b2);
    float dg2 = 0.f, db2 = 0.f;
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const float wo = S.w_out[q * W + c];
      atomicAdd(R.g_w_out + q * W + c, P.gamma2[c] * acc_S[q] + P.beta2[c] * acc_B[q]);
      dg2 = fmaf(wo, acc_S[q], dg2);
      db2 = fmaf(wo, acc_B[q], db2);
    }
    atomicAdd(R.g_gamma + W + c, dg2);
    atomicAdd(R.g_beta + W + c, db2);
#pragma unroll
    for (int i = 0; i < NIN; ++i)
      if (i < IN) atomicAdd(R.g_w_in + (size_t)c * IN + i, acc_win[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (GRAD && tid < W) {      // d beta1 = sum_rows dh1 = W_hid^T (sum_rows dz2): one matrix-vector product per CTA
    float a = 0.f;
    for (int kk = 0; kk < W; ++kk) a = fmaf(P.w_hid[(size_t)kk * W + tid], S.db2S[kk], a);
    atomicAdd(R.g_beta + tid, a);
  }
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---- the sequential part: one thread per environment, a scan of (n + o + m)-vectors over t = H-1 .. 0 ------------------------
struct ScanArgs {
  int64_t N;
  int H, in_dim;
  const float* jac;          // [H][N][STRIDE]
  const float* pol_jac;      // [H][N][NA][NOP]
  const float* actions;      // [H][N][NA]   a_t = tanh(mu + sigma eps)
  const float* eps;          // [H][N][NA]
  const float* grw;          // [H]
  const float* gobs_term;    // [K][N][in_dim] or null
  const int* term_of_row;    // [H+2] or null
  const float* log_std;      // [NA]
  float* dmu;                // [H][N][NA]   out: d loss / d mu_t
  float* g_b_out;  float* g_log_std;       // [NA] accumulated (atomics)
};

constexpr int SCAN_STAGES = 4;
template <int ENV> struct ScanRing {       // words per thread and stage: the step's Jacobian row + policy Jacobian, pitch = 4 (mod 32)
  static constexpr int NOP = (Task<ENV>::NO + 3) & ~3;
  static constexpr int WORDS = JacShape<ENV>::STRIDE + Task<ENV>::NA * NOP + 4 * ((2 * Task<ENV>::NA + 3) / 4);   // + actions, eps
  static constexpr int PITCH = WORDS + ((4 - WORDS % 32) + 32) % 32;
  static size_t bytes(int block) { return (size_t)SCAN_STAGES * block * PITCH * 4; }
};

template <int ENV>
__global__ void __launch_bounds__(128)
rollout_scan_kernel(const __grid_constant__ ScanArgs A) {
  using TASK = Task<ENV>;
  using JS = JacShape<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, KD = JS::KD, NOP = (NO + 3) & ~3;
  const int64_t env = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool valid = env < A.N;
  const int64_t e = valid ? env : 0;
  float gs[NS], gobs[NO], sb[NA], sl[NA], sigma[NA];
#pragma unroll
  for (int i = 0; i < NS; ++i) gs[i] = 0.f;
#pragma unroll
  for (int i = 0; i < NO; ++i) gobs[i] = 0.f;
#pragma unroll
  for (int q = 0; q < NA; ++q) { sb[q] = 0.f; sl[q] = 0.f; sigma[q] = expf(A.log_std[q]); }
  // The rows of the recursion do not depend on it: they stream through a SCAN_STAGES-deep ring in shared memory (cp.async), so the
  // loads of step t - SCAN_STAGES are in flight while step t computes -- one step of arithmetic (~150 instructions) does not cover
  // an L2 round trip, four do.  The 32 rows of a warp are contiguous in HBM, so the lanes copy them COOPERATIVELY (lane l takes the
  // 16-byte chunks l, l + 32, ... of the block: 512 contiguous bytes per instruction instead of 32 different lines) into the
  // per-thread slots; a __syncwarp orders the copies against the owner's reads.  The slot pitch is 4 (mod 32) words: a quarter
  // warp's 16-byte reads touch all 32 banks once.
  constexpr int REC4 = JS::STRIDE / 4, PJ4 = NA * NOP / 4, P4 = ScanRing<ENV>::PITCH / 4;
  extern __shared__ float4 scan_ring[];
  float4* mine = scan_ring + (size_t)threadIdx.x * P4;
  const size_t stage_stride = (size_t)blockDim.x * P4;
  const int lane = threadIdx.x & 31;
  const int64_t env0 = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31);          // first environment of this warp
  const uint32_t warp_ring = smem_u32(scan_ring + (size_t)(threadIdx.x & ~31) * P4);
  auto issue = [&](int t, int stage) {
    if (t >= 0) {
      const size_t row = (size_t)t * A.N + e;
      const float4* src = reinterpret_cast<const float4*>(A.jac + ((size_t)t * A.N + env0) * JS::STRIDE);
      const float4* ps = reinterpret_cast<const float4*>(A.pol_jac + ((size_t)t * A.N + env0) * (NA * NOP));
      const uint32_t wdst = warp_ring + (uint32_t)(stage * stage_stride * 16);
      const uint32_t dst = smem_u32(mine + stage * stage_stride);
#pragma unroll
      for (int i = 0; i < REC4; ++i) {
        const int c = i * 32 + lane, th = c / REC4, off = c % REC4;
        if (env0 + th < A.N)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wdst + (uint32_t)(th * P4 + off) * 16), "l"(src + c) : "memory");
      }
#pragma unroll
      for (int i = 0; i < PJ4; ++i) {
        const int c = i * 32 + lane, th = c / PJ4, off = c % PJ4;
        if (env0 + th < A.N)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wdst + (uint32_t)(th * P4 + REC4 + off) * 16), "l"(ps + c) : "memory");
      }
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 16 * (REC4 + PJ4) + 4 * q), "l"(A.actions + row * NA + q) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 16 * (REC4 + PJ4) + 4 * (NA + q)), "l"(A.eps + row * NA + q) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");        // (an empty group keeps the wait count uniform)
  };
  // (the per-step scalars shared by all environments -- grw[t], term_of_row[t + 1] -- are L1 hits after the first touch of their
  // line; the terminal-value gradient rows are read where a terminal row occurs, once or twice per horizon)
  struct Small { float act[NA], eps[NA], term[NO], gr; };
  auto fetch_small = [&](int t, Small& in) {
    const int kt = A.term_of_row ? A.term_of_row[t + 1] : -1;
#pragma unroll
    for (int i = 0; i < NO; ++i) in.term[i] = kt >= 0 ? A.gobs_term[((size_t)kt * A.N + e) * A.in_dim + i] : 0.f;
    in.gr = A.grw[t];
  };
#pragma unroll
  for (int sidx = 0; sidx < SCAN_STAGES; ++sidx) issue(A.H - 1 - sidx, sidx);
  Small cur, nxt;
  fetch_small(A.H - 1, cur);
  int stage = 0;
  for (int t = A.H - 1; t >= 0; --t) {
    if (t > 0) fetch_small(t - 1, nxt);
    asm volatile("cp.async.wait_group %0;" ::"n"(SCAN_STAGES - 1) : "memory");
    __syncwarp();                            // every lane's share of this stage has landed
    const size_t row = (size_t)t * A.N + e;
    float rec[JS::STRIDE], pj[NA * NOP];
    const float4* in4 = mine + stage * stage_stride;
#pragma unroll
    for (int i = 0; i < REC4; ++i) { const float4 v = in4[i]; rec[4 * i] = v.x; rec[4 * i + 1] = v.y; rec[4 * i + 2] = v.z; rec[4 * i + 3] = v.w; }
#pragma unroll
    for (int i = 0; i < PJ4; ++i) { const float4 v = in4[REC4 + i]; pj[4 * i] = v.x; pj[4 * i + 1] = v.y; pj[4 * i + 2] = v.z; pj[4 * i + 3] = v.w; }
    {
      const float* sc = reinterpret_cast<const float*>(in4 + REC4 + PJ4);
#pragma unroll
      for (int q = 0; q < NA; ++q) { cur.act[q] = sc[q]; cur.eps[q] = sc[NA + q]; }
    }
    __syncwarp();                            // every lane holds its row in registers before any lane refills the slots
    issue(t - SCAN_STAGES, stage);
    stage = stage + 1 == SCAN_STAGES ? 0 : stage + 1;
    float go[NO];
#pragma unroll
    for (int i = 0; i < NO; ++i) go[i] = gobs[i] + cur.term[i];
    float v[KD];
#pragma unroll
    for (int d = 0; d < KD; ++d) {
      float x = cur.gr * rec[(NS + NO) * KD + d];
#pragma unroll
      for (int i = 0; i < NS; ++i) x = fmaf(gs[i], rec[i * KD + d], x);
#pragma unroll
      for (int i = 0; i < NO; ++i) x = fmaf(go[i], rec[(NS + i) * KD + d], x);
      v[d] = x;
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) gs[i] = v[i];
#pragma unroll
    for (int i = 0; i < NO; ++i) gobs[i] = 0.f;
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const float ga_pre = valid ? v[NS + q] * (1.f - cur.act[q] * cur.act[q]) : 0.f;      // d loss / d (mu + sigma eps)
      if (valid) A.dmu[row * NA + q] = ga_pre;
      sb[q] += ga_pre;
      sl[q] += ga_pre * cur.eps[q] * sigma[q];
#pragma unroll
      for (int i = 0; i < NO; ++i) gobs[i] = fmaf(pj[q * NOP + i], ga_pre, gobs[i]);
    }
    cur = nxt;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int q = 0; q < NA; ++q) {
    float b = sb[q], l = sl[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { b += __shfl_xor_sync(0xffffffffu, b, o); l += __shfl_xor_sync(0xffffffffu, l, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(A.g_b_out + q, b); atomicAdd(A.g_log_std + q, l); }
  }
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

template <int NA, int NO, int NIN, bool GRAD, int NSP>
int launch_rows_bwd(const RowsBwdArgs& R, const PolicyTc& p, int in_dim, cudaStream_t st) {
  if (in_dim > NIN) return fail2(SGBPO_E_ARG, "tensor-core reverse pass: policy input wider than the environment's observation");
  const size_t smem = BwdSmem<NSP>::bytes(in_dim, NIN, NA, GRAD);
  if (smem > SMEM_CAP) return fail2(SGBPO_E_UNSUPPORTED, "tensor-core reverse pass: policy input too wide for shared memory");
  auto kern = policy_rows_bwd_kernel<NA, NO, NIN, GRAD, NSP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail2(e, "cudaFuncSetAttribute(policy_rows_bwd)");
  kern<<<tc_grid(R.M), ROWS * NSP, smem, st>>>(R, p);
  e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "policy_rows_bwd_kernel");
}

template <int ENV> int rollout_tc_bwd(const sgbpo_mlp* pol, const sgbpo_rollout* r, const sgbpo_rollout_grads* g, cudaStream_t st) {
  using TASK = Task<ENV>;
  using JS = JacShape<ENV>;
  constexpr int NA = TASK::NA, NO = TASK::NO, NOP = (NO + 3) & ~3;
  constexpr int NIN = TASK::NAUX > 2 ? NO + (TASK::NAUX - 2) : NO;       // constant observation columns appended from aux (NavigateSeeker)
  PolicyTc p;
  if (!tc_policy(pol, p) || pol->out_dim != NA) return fail2(SGBPO_E_UNSUPPORTED, "tensor-core reverse pass: float32 policy [in -> 128 -> 128 -> m]");
  if (!r->jac || !r->pol_jac || !r->dmu || !g->g_w_hid) return fail2(SGBPO_E_ARG, "tensor-core reverse pass: jac, pol_jac, dmu and g_w_hid buffers are required");
  if (reinterpret_cast<uintptr_t>(g->g_w_hid) % 16) return fail2(SGBPO_E_ARG, "tensor-core reverse pass: g_w_hid must be 16-byte aligned");
  RowsBwdArgs R;
  R.M = r->N * (int64_t)r->H;
  R.e1 = (const float*)r->stash_e1; R.e2 = (const float*)r->stash_e2; R.stats = (const float*)r->stash_stats;
  R.obs = (const float*)r->obs; R.dmu = (const float*)r->dmu; R.pol_jac = (float*)r->pol_jac;
  R.g_w_hid = (float*)g->g_w_hid; R.g_w_in = (float*)g->g_w_in; R.g_w_out = (float*)g->g_w_out; R.g_b_hid = (float*)g->g_b_hid;
  R.g_gamma = (float*)g->g_gamma; R.g_beta = (float*)g->g_beta;
  // threads per row: 8 unless the Jacobian-mode exchange [8][128][NOP] would overflow the 64 KB operand buffer it aliases
  const int forced = env_int("SGBPO_ROWS_BWD_SPLIT");
  const bool split8 = forced == 8;      // (measured: the 4-way split is faster at 4096 and 65536 environments, profiles/r2_engines.md)
  // 1. policy input-output Jacobians of all H x N rows
  {
    const int rc = split8 ? launch_rows_bwd<NA, NO, NIN, false, 8>(R, p, pol->in_dim, st) : launch_rows_bwd<NA, NO, NIN, false, 4>(R, p, pol->in_dim, st);
    if (rc != SGBPO_OK) return rc;
  }
  // 2. the sequential recursion: a scan per environment
  {
    ScanArgs S;
    S.N = r->N; S.H = r->H; S.in_dim = pol->in_dim;
    S.jac = (const float*)r->jac; S.pol_jac = (const float*)r->pol_jac; S.actions = (const float*)r->actions; S.eps = (const float*)r->eps;
    S.grw = (const float*)g->grw; S.gobs_term = (const float*)g->gobs_term; S.term_of_row = g->term_of_row;
    S.log_std = p.log_std; S.dmu = (float*)r->dmu; S.g_b_out = (float*)g->g_b_out; S.g_log_std = (float*)g->g_log_std;
    int block = r->N >= 148 * 128 ? 128 : 32;
    while (block > 32 && ScanRing<ENV>::bytes(block) > 200 * 1024) block >>= 1;
    const size_t ring = ScanRing<ENV>::bytes(block);
    if (ring > 48 * 1024) {
      const cudaError_t ea = cudaFuncSetAttribute(rollout_scan_kernel<ENV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
      if (ea != cudaSuccess) return cuda_fail2(ea, "cudaFuncSetAttribute(rollout_scan)");
    }
    rollout_scan_kernel<ENV><<<(int)((r->N + block - 1) / block), block, ring, st>>>(S);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail2(e, "rollout_scan_kernel");
  }
  // 3. parameter gradients of the batched policy for the row cotangents
  {
    const int rc = split8 ? launch_rows_bwd<NA, NO, NIN, true, 8>(R, p, pol->in_dim, st) : launch_rows_bwd<NA, NO, NIN, true, 4>(R, p, pol->in_dim, st);
    if (rc != SGBPO_OK) return rc;
  }
  (void)JS::STRIDE; (void)NOP;
  return SGBPO_OK;
}

template <int ENV> int pol_jac_stride_of() { return Task<ENV>::NA * ((Task<ENV>::NO + 3) & ~3); }

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
