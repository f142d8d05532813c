// Whole-horizon fused rollout: SHAC.collect_trajectories + the reverse pass of update_policy
// (reference: src/learning_algorithms/shac.py:96-151) for N environments and H steps in THREE launches:
//
//   rollout_fwd_kernel   persistent over t = 0..H-1.  A warp owns 8 environments for the whole horizon:
//                        policy MLP (weights in shared memory, fp32/fp64 FMA register tiles) -> tanh-Gaussian
//                        action -> fused safeguarded step (sgbpo_step.cuh) on lanes 0..7 -> the next
//                        observation goes straight back into the warp's shared tile.  State never leaves
//                        registers between steps; HBM sees only the rollout-buffer rows (coalesced 8-env spans).
//   mlp_rows_kernel      target critic over all (H x N) observation rows, batched (not on the sequential chain).
//   rollout_bwd_kernel   persistent over t = H-1..0: recompute the policy forward, dual-number Jacobian of the
//                        step contracted with the carried cotangents, policy data-gradient; small parameter
//                        gradients accumulate in registers, the hidden-hidden operands (h1, dz2) are streamed
//                        out for one split-K GEMM.
//
// Environments are independent, so there is no inter-CTA synchronisation anywhere; the reset-all of Q6
// is uniform over environments and is resolved on the host into a per-step schedule.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/sgbpo.h"
#include "sgbpo_mlp.cuh"
#include "sgbpo_step.cuh"
#include "sgbpo_host.cuh"

namespace sgbpo {

extern thread_local char g_err[512];

template <class T> struct RolloutArgs {
  int64_t N;
  int H;
  // inputs
  const T* eps;           // [H][N][m]   standard normal draws of the policy
  const T* noise;         // [H][N][n]   noise samples w_t (centre included; household: T_out already substituted)
  const T* dirs;          // [H][N][m][2m] or null
  const T* aux0;          // [N][naux] or null
  const T* reset_states;  // [R][N][n] or null
  const int* reset_idx;   // [H]: -1, or the index r of the reset state that REPLACES s_{t+1}
  const int* aux_src;     // [H]: -1 -> aux0, else r -> reset_states[r][:, 0:naux] is the aux in effect after step t
  const StepShared<T>* shared;   // [H] or null (household)
  // rollout buffers
  T* states;              // [H+1][N][n]   row 0 = s_0 (input), rows 1..H written
  T* obs;                 // [H+2][N][o]   row 0 = obs_0 (input), rows 1..H written
  T* actions;             // [H+1][N][m]   policy actions a_t, rows 0..H-1 written
  T* exec_actions;        // [H][N][m]     executed (safeguarded) actions
  T* rewards;             // [H+2][N]      rows 0..H-1 written
  int32_t* flags;         // [H][N]
  // activation stash (fwd writes, bwd reads): [H][N][W] x 3 and [H][N][4]
  T* stash_h1; T* stash_e1; T* stash_e2; T* stash_stats;
  int stash_tiled;        // e1 / e2 in the 128-row tile layout the parallel reverse pass reads (stash_index)
};

// Position of (row = t * N + env, col) in an activation stash.  Row-major [row][W], or -- for the parallel reverse pass, whose
// threads own one row and 32 consecutive columns -- 128-row tiles of 16-byte column chunks, [row / 128][col / 4][row % 128][col % 4]:
// consecutive rows of a chunk are contiguous, so a warp (32 rows, one chunk) reads 512 contiguous bytes per request.
template <int W>
__device__ __forceinline__ size_t stash_index(bool tiled, int64_t row, int col) {
  return tiled ? ((((size_t)(row >> 7) * (W / 4) + (size_t)(col >> 2)) << 7) + (size_t)(row & 127)) * 4 + (size_t)(col & 3)
               : (size_t)row * W + col;
}

template <class T> struct BackwardArgs {
  const T* grw;           // [H]   d loss / d reward[t]  (0 on rows replaced by the critic value)
  const T* gobs_term;     // [K][N][o] gradient of the terminal-value terms w.r.t. obs rows, or null
  const int* term_of_row; // [H+2]: -1 or k
  T* dz2_out;             // [H][N][W]   layer-2 pre-activation grads: dW_hid = dz2^T stash_h1
  // parameter gradients (accumulated with atomics, caller zero-fills)
  T* g_w_in;              // [W][o]  (torch layout of Linear.weight)
  T* g_w_out;             // [m][W]
  T* g_b_hid;             // [2][W]
  T* g_gamma;             // [2][W]
  T* g_beta;              // [2][W]
  T* g_b_out;             // [m]
  T* g_log_std;           // [m]
};

// register tile <-> global [t][env][W]: per row, the warp touches W contiguous values (128-byte segments)
template <class T, int W, int R>
__device__ __forceinline__ void stream_tile(T* __restrict__ dst, int t, int64_t N, int64_t env0, int lane,
                                            const T (&v)[R][W / 32], bool tiled = false) {
  // (row, lane + 32 j) -> stash_index, with the lane and j parts hoisted: one add per store
  const size_t lane_off = tiled ? (size_t)(lane >> 2) * 512 + (lane & 3) : (size_t)lane, j_stride = tiled ? 8 * 512 : 32;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (env0 + r < N) {
      const int64_t row = (int64_t)t * N + env0 + r;
      T* p = dst + (tiled ? ((size_t)(row >> 7) * (W / 4) * 512 + (size_t)(row & 127) * 4) : (size_t)row * W) + lane_off;
#pragma unroll
      for (int j = 0; j < W / 32; ++j) p[j * j_stride] = v[r][j];
    }
  }
}
template <class T, int W, int R>
__device__ __forceinline__ void fetch_tile(const T* __restrict__ src, int t, int64_t N, int64_t env0, int lane,
                                           T (&v)[R][W / 32]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool ok = env0 + r < N;
#pragma unroll
    for (int j = 0; j < W / 32; ++j) v[r][j] = ok ? src[((int64_t)t * N + env0 + r) * W + lane + 32 * j] : T(0);
  }
}

// launch bounds per rows-per-warp R: fewer rows -> smaller register tiles -> more resident warps
template <int R> struct RolloutShape {
  static constexpr int MAX_WARPS = R == 8 ? 8 : (R == 4 ? 12 : 16);      // backward kernel: 255 / 168 / 128 registers per thread
};
// forward kernel: <= 80 registers per thread in fp32 (16 warps fit), fp64 tiles need the full register budget
template <class T, int R> struct FwdShape { static constexpr int MAX_WARPS = sizeof(T) == 8 ? RolloutShape<R>::MAX_WARPS : 16; };

// ---------------------------------------------------------------------------------------------------------
template <int ENV, class T, int W, int R>
__global__ void __launch_bounds__(FwdShape<T, R>::MAX_WARPS * 32)
rollout_fwd_kernel(const __grid_constant__ RolloutArgs<T> A, const __grid_constant__ MlpParams<T> P,
                   const __grid_constant__ SafeConsts<T> k) {
  using TASK = Task<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX, CPL = W / 32;
  extern __shared__ __align__(32) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  NetSmem<T, W> net;                       // same carve in every thread: pointers stay in registers
  T* after = net.load(smem, P);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // warp-private tiles (32-byte aligned): obsT [NO][R], bufA/bufB [W][R]
  size_t off = (size_t)(after - smem);
  off = (off + 7) & ~size_t(7);
  const size_t per_warp = ((size_t)(NO + 2 * W) * R + 7) & ~size_t(7);
  T* obsT = smem + off + warp * per_warp;
  T* bufA = obsT + NO * R;
  T* bufB = bufA + W * R;
  const int64_t env0 = ((int64_t)blockIdx.x * nwarps + warp) * R;       // first env of this warp
  const int64_t env = env0 + lane;
  const bool owner = lane < R;
  const bool valid = owner && env < A.N;
  const int64_t N = A.N;

  T s[NS], aux[NAUX > 0 ? NAUX : 1];
  T sigma[NA];
#pragma unroll
  for (int q = 0; q < NA; ++q) sigma[q] = sg_exp(P.log_std[q]);
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = valid ? A.states[env * NS + i] : T(0);
  if (NAUX > 0) {
#pragma unroll
    for (int i = 0; i < NAUX; ++i) aux[i] = valid ? A.aux0[env * NAUX + i] : T(0);
  }
  if (owner) {
#pragma unroll
    for (int i = 0; i < NO; ++i) obsT[i * R + lane] = valid ? A.obs[env * NO + i] : T(0);
  }
  __syncwarp();

  for (int t = 0; t < A.H; ++t) {
    // issue this step's streaming loads before the MLP so their latency hides behind it
    T epsv[NA], wv[NS], dv[NA * 2 * NA], rs[NS];
    const int ridx = A.reset_idx ? A.reset_idx[t] : -1;
    if (valid) {
#pragma unroll
      for (int q = 0; q < NA; ++q) epsv[q] = A.eps[((int64_t)t * N + env) * NA + q];
#pragma unroll
      for (int i = 0; i < NS; ++i) wv[i] = A.noise[((int64_t)t * N + env) * NS + i];
      if (A.dirs) {
#pragma unroll
        for (int i = 0; i < NA * 2 * NA; ++i) dv[i] = A.dirs[((int64_t)t * N + env) * (NA * 2 * NA) + i];
      }
      if (ridx >= 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) rs[i] = A.reset_states[((int64_t)ridx * N + env) * NS + i];
      }
    }
    // ---- policy forward -------------------------------------------------------------------------
    T h[R][CPL];
    if (A.stash_e1) {
      // two hidden layers, activations kept for the reverse pass (ELU outputs, LayerNorm statistics, h1)
      T ev[R][CPL], m1[R], r1[R], m2[R], r2[R];
      tile_gemm<T, W, R>(obsT, NO, net.w_in, net.b_hid, lane, h);
      elu_layernorm<T, W, R, true>(h, net.gamma, net.beta, lane, ev, m1, r1);
      store_tile<T, W, R>(bufA, lane, h);
      stream_tile<T, W, R>(A.stash_e1, t, N, env0, lane, ev, A.stash_tiled != 0);
      if (A.stash_h1) stream_tile<T, W, R>(A.stash_h1, t, N, env0, lane, h);
      __syncwarp();
      tile_gemm<T, W, R>(bufA, W, net.w_hid, net.b_hid + W, lane, h);
      elu_layernorm<T, W, R, true>(h, net.gamma + W, net.beta + W, lane, ev, m2, r2);
      stream_tile<T, W, R>(A.stash_e2, t, N, env0, lane, ev, A.stash_tiled != 0);
      if (valid) {
        T st[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (lane == r) { st[0] = m1[r]; st[1] = r1[r]; st[2] = m2[r]; st[3] = r2[r]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) A.stash_stats[((int64_t)t * N + env) * 4 + i] = st[i];
      }
    } else {
      net_forward<T, W, R>(net, obsT, bufA, bufB, lane, h);
    }
    T mu[R][NA];
    out_layer<T, W, R, NA>(h, net.w_out, net.b_out, lane, mu);
    // ---- environment step on the owner lanes ---------------------------------------------------
    __syncwarp();
    if (owner) {
      T a[NA];
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        T m = T(0);
#pragma unroll
        for (int r = 0; r < R; ++r) m = (lane == r) ? mu[r][q] : m;
        a[q] = valid ? sg_tanh(m + sigma[q] * epsv[q]) : T(0);
      }
      if (!valid) {
#pragma unroll
        for (int i = 0; i < NS; ++i) wv[i] = T(0);
      }
      if (NAUX > 0 && ridx >= 0 && valid) {     // quadrotor / seeker: goal = reset position
#pragma unroll
        for (int i = 0; i < NAUX; ++i) aux[i] = rs[i];
      }
      T sn[NS], ae[NA], ob[NO], rew;
      int fl;
      const StepShared<T>* sh = A.shared ? A.shared + t : nullptr;
      StepShared<T> empty;
      if (!sh) { memset(&empty, 0, sizeof(empty)); sh = &empty; }
      safeguarded_step<TASK, T, T>(s, a, wv, aux, A.dirs ? dv : nullptr, sh, k, ridx >= 0, rs, sn, ae, ob, rew, fl);
#pragma unroll
      for (int i = 0; i < NS; ++i) s[i] = sn[i];
#pragma unroll
      for (int i = 0; i < NO; ++i) obsT[i * R + lane] = ob[i];
      if (valid) {
#pragma unroll
        for (int i = 0; i < NS; ++i) A.states[((int64_t)(t + 1) * N + env) * NS + i] = sn[i];
#pragma unroll
        for (int i = 0; i < NO; ++i) A.obs[((int64_t)(t + 1) * N + env) * NO + i] = ob[i];
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          A.actions[((int64_t)t * N + env) * NA + q] = a[q];
          A.exec_actions[((int64_t)t * N + env) * NA + q] = ae[q];
        }
        A.rewards[(int64_t)t * N + env] = rew;
        A.flags[(int64_t)t * N + env] = fl;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// batched MLP forward over M rows: out[row] = net(x[row]) (critic: out_dim = 1).
template <class T, int W, int R>
__global__ void __launch_bounds__(R == 8 ? 256 : 512)
mlp_rows_kernel(int64_t M, int in_dim, const T* __restrict__ x, T* __restrict__ out,
                const __grid_constant__ MlpParams<T> P) {
  constexpr int CPL = W / 32;
  extern __shared__ __align__(32) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  NetSmem<T, W> net;                       // same carve in every thread: pointers stay in registers
  T* after = net.load(smem, P);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  size_t off = (size_t)(after - smem);
  off = (off + 7) & ~size_t(7);
  const size_t per_warp = ((size_t)(in_dim + 2 * W) * R + 7) & ~size_t(7);
  T* inT = smem + off + warp * per_warp;
  T* bufA = inT + in_dim * R;
  T* bufB = bufA + W * R;
  const int64_t n_blocks = (M + R - 1) / R;
  for (int64_t blk = (int64_t)blockIdx.x * nwarps + warp; blk < n_blocks; blk += (int64_t)gridDim.x * nwarps) {
    const int64_t row0 = blk * R;
    const int total = in_dim * R;
    for (int i = lane; i < total; i += 32) {        // contiguous R-row span, coalesced
      const int r = i / in_dim, c = i % in_dim;
      inT[c * R + r] = (row0 + r < M) ? x[row0 * in_dim + i] : T(0);
    }
    __syncwarp();
    T h[R][CPL];
    net_forward<T, W, R>(net, inT, bufA, bufB, lane, h);
    T v[R][1];
    out_layer<T, W, R, 1>(h, net.w_out, net.b_out, lane, v);
    if (lane < R && row0 + lane < M) {
      T mine = T(0);
#pragma unroll
      for (int r = 0; r < R; ++r) mine = (lane == r) ? v[r][0] : mine;
      out[row0 + lane] = mine;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
template <int ENV, class T, int W, int R>
__global__ void __launch_bounds__(RolloutShape<R>::MAX_WARPS * 32)
rollout_bwd_kernel(const __grid_constant__ RolloutArgs<T> A, const __grid_constant__ BackwardArgs<T> G,
                   const __grid_constant__ MlpParams<T> P, const __grid_constant__ SafeConsts<T> k) {
  using TASK = Task<ENV>;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX, CPL = W / 32, KD = NS + NA;
  // The step Jacobian is forward-mode: KD = n + m seed directions.  They are spread over the warp instead of
  // being carried by one lane: lane = r + R * g evaluates the step of environment r with the DPL directions
  // [g * DPL, (g + 1) * DPL) seeded, so the serial dual-number chain is (1 + DPL) wide instead of (1 + KD).
  constexpr int GROUPS = 32 / R, DPL = (KD + GROUPS - 1) / GROUPS, GUSED = (KD + DPL - 1) / DPL;
  using D = Dual<T, DPL>;
  extern __shared__ __align__(32) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  NetSmem<T, W> net;                       // same carve in every thread: pointers stay in registers
  T* after = net.load(smem, P);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  size_t off = (size_t)(after - smem);
  off = (off + 7) & ~size_t(7);
  // warp-private: obsT [NO][R] and g2T [W][R] (operand of the transposed data-gradient GEMM); the forward
  // activations come from the stash the forward kernel left in HBM, nothing is recomputed
  const size_t per_warp = ((size_t)(NO + W) * R + 7) & ~size_t(7);
  T* obsT = smem + off + warp * per_warp;
  T* g2T = obsT + NO * R;
  const int64_t N = A.N;
  const int64_t env0 = ((int64_t)blockIdx.x * nwarps + warp) * R;      // first env of this warp
  const int row_of_lane = lane % R, group = lane / R;                   // environment and direction group of this lane
  const int64_t env = env0 + row_of_lane;
  const bool owner = lane < R;
  const bool stepper = group < GUSED;                                   // lanes that evaluate a slice of the step Jacobian
  const bool valid = env < N;

  T sigma[NA];
#pragma unroll
  for (int q = 0; q < NA; ++q) sigma[q] = sg_exp(P.log_std[q]);
  const T* Whid = net.w_hid;                 // single hidden-hidden layer (n_hidden == 2)
  // carried cotangents of this lane's environment (replicated over its direction groups)
  T gs[NS], gobs[NO];
#pragma unroll
  for (int i = 0; i < NS; ++i) gs[i] = T(0);
#pragma unroll
  for (int i = 0; i < NO; ++i) gobs[i] = T(0);
  // per-lane parameter-gradient accumulators
  T a_b1[CPL], a_g1[CPL], a_be1[CPL], a_b2[CPL], a_g2[CPL], a_be2[CPL], a_wout[CPL][NA], a_win[NO][CPL];
  T a_bout[NA], a_lstd[NA];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    a_b1[j] = a_g1[j] = a_be1[j] = a_b2[j] = a_g2[j] = a_be2[j] = T(0);
#pragma unroll
    for (int q = 0; q < NA; ++q) a_wout[j][q] = T(0);
#pragma unroll
    for (int i = 0; i < NO; ++i) a_win[i][j] = T(0);
  }
#pragma unroll
  for (int q = 0; q < NA; ++q) a_bout[q] = a_lstd[q] = T(0);

  for (int t = A.H - 1; t >= 0; --t) {
    // ---- loads (every lane of a direction group reads its environment's step inputs) ---------------------
    T sv[NS], epsv[NA], wv[NS], dv[NA * 2 * NA], rs[NS], aux[NAUX > 0 ? NAUX : 1];
    const int ridx = A.reset_idx ? A.reset_idx[t] : -1;
#pragma unroll
    for (int i = 0; i < NS; ++i) { sv[i] = T(0); wv[i] = T(0); rs[i] = T(0); }
#pragma unroll
    for (int q = 0; q < NA; ++q) epsv[q] = T(0);
    if (valid && stepper) {
#pragma unroll
      for (int i = 0; i < NS; ++i) sv[i] = A.states[((int64_t)t * N + env) * NS + i];
#pragma unroll
      for (int q = 0; q < NA; ++q) epsv[q] = A.eps[((int64_t)t * N + env) * NA + q];
#pragma unroll
      for (int i = 0; i < NS; ++i) wv[i] = A.noise[((int64_t)t * N + env) * NS + i];
      if (A.dirs) {
#pragma unroll
        for (int i = 0; i < NA * 2 * NA; ++i) dv[i] = A.dirs[((int64_t)t * N + env) * (NA * 2 * NA) + i];
      }
      if (ridx >= 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) rs[i] = A.reset_states[((int64_t)ridx * N + env) * NS + i];
      }
      if (NAUX > 0) {
        const int src = A.aux_src ? A.aux_src[t] : -1;
#pragma unroll
        for (int i = 0; i < NAUX; ++i)
          aux[i] = src < 0 ? A.aux0[env * NAUX + i] : A.reset_states[((int64_t)src * N + env) * NS + i];
      }
    }
    if (owner) {
#pragma unroll
      for (int i = 0; i < NO; ++i) obsT[i * R + lane] = valid ? A.obs[((int64_t)t * N + env) * NO + i] : T(0);
    }
    __syncwarp();
    // ---- forward activations of this step from the stash (loads issued now, consumed after the step) ---------
    T mean1[R], rstd1[R], mean2[R], rstd2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = env0 + r < N;
      const T* st = A.stash_stats + ((int64_t)t * N + env0 + r) * 4;
      mean1[r] = ok ? st[0] : T(0); rstd1[r] = ok ? st[1] : T(0);
      mean2[r] = ok ? st[2] : T(0); rstd2[r] = ok ? st[3] : T(0);
    }
    T e2v[R][CPL];
    fetch_tile<T, W, R>(A.stash_e2, t, N, env0, lane, e2v);
    // ---- step Jacobian (duals) contracted with the carried cotangents, on the owner lanes -----------------
    T ga_pre[NA];          // d loss / d (mu + sigma eps)   (after the tanh)
#pragma unroll
    for (int q = 0; q < NA; ++q) ga_pre[q] = T(0);
    T act[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) act[q] = T(0);
    T part[DPL];           // this lane's slice of J^T cotangent: directions group * DPL + j
#pragma unroll
    for (int j = 0; j < DPL; ++j) part[j] = T(0);
    if (stepper) {
#pragma unroll
      for (int q = 0; q < NA; ++q) act[q] = valid ? A.actions[((int64_t)t * N + env) * NA + q] : T(0);   // a_t = tanh(mu + sigma eps)
      D sd[NS], ad[NA], sn[NS], ae[NA], ob[NO], rew;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        sd[i] = D::lift(sv[i]);
#pragma unroll
        for (int j = 0; j < DPL; ++j) sd[i].d[j] = (group * DPL + j == i) ? T(1) : T(0);
      }
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        ad[q] = D::lift(act[q]);
#pragma unroll
        for (int j = 0; j < DPL; ++j) ad[q].d[j] = (group * DPL + j == NS + q) ? T(1) : T(0);
      }
      int fl;
      const StepShared<T>* sh = A.shared ? A.shared + t : nullptr;
      StepShared<T> empty;
      if (!sh) { memset(&empty, 0, sizeof(empty)); sh = &empty; }
      safeguarded_step<TASK, D, T>(sd, ad, wv, aux, A.dirs ? dv : nullptr, sh, k, ridx >= 0, rs, sn, ae, ob, rew, fl);
      // cotangent of obs_{t+1}: carried policy-input gradient + terminal-value term of row t+1
      T go[NO];
#pragma unroll
      for (int i = 0; i < NO; ++i) go[i] = gobs[i];
      const int kt = G.term_of_row ? G.term_of_row[t + 1] : -1;
      if (kt >= 0 && valid) {
#pragma unroll
        for (int i = 0; i < NO; ++i) go[i] += G.gobs_term[((int64_t)kt * N + env) * NO + i];
      }
      const T gr = G.grw[t];
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
        T v = gr * rew.d[j];
#pragma unroll
        for (int i = 0; i < NS; ++i) v += gs[i] * sn[i].d[j];
#pragma unroll
        for (int i = 0; i < NO; ++i) v += go[i] * ob[i].d[j];
        part[j] = valid ? v : T(0);
      }
    }
    // gather the KD slices: direction d lives in slot d % DPL of lane row + R * (d / DPL); every lane of the
    // environment ends up with the complete vector, so the carried cotangents stay replicated
    {
      T acc[KD];
#pragma unroll
      for (int d = 0; d < KD; ++d) acc[d] = __shfl_sync(0xffffffffu, part[d % DPL], row_of_lane + R * (d / DPL));
#pragma unroll
      for (int i = 0; i < NS; ++i) gs[i] = acc[i];
      if (owner) {
#pragma unroll
        for (int q = 0; q < NA; ++q) ga_pre[q] = acc[NS + q] * (T(1) - act[q] * act[q]);
      }
    }
    // broadcast d mu to the whole warp: dmu[r][q] from lane r
    T dmu[R][NA];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < NA; ++q) dmu[r][q] = __shfl_sync(0xffffffffu, ga_pre[q], r);
    if (owner) {
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        a_bout[q] += ga_pre[q];
        a_lstd[q] += ga_pre[q] * epsv[q] * sigma[q];
      }
    }
    // ---- policy backward ---------------------------------------------------------------------------------
    T gcarry[R][CPL];       // gradient flowing into the current layer's output
    {
      // output layer: dh2 = dmu w_out^T ; dW_out += h2^T dmu (h2 = xhat2 * gamma2 + beta2)
      T xh[R][CPL], eg[R][CPL];
      unstash<T, W, R>(e2v, mean2, rstd2, xh, eg);
      T wo[CPL][NA], g2[CPL], b2[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        g2[j] = net.gamma[W + lane + 32 * j];
        b2[j] = net.beta[W + lane + 32 * j];
#pragma unroll
        for (int q = 0; q < NA; ++q) wo[j][q] = net.w_out[(lane + 32 * j) * NA + q];
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          T dh = T(0);
          const T h2 = xh[r][j] * g2[j] + b2[j];
#pragma unroll
          for (int q = 0; q < NA; ++q) { dh = fma(dmu[r][q], wo[j][q], dh); a_wout[j][q] = fma(h2, dmu[r][q], a_wout[j][q]); }
          gcarry[r][j] = dh;
        }
      // LayerNorm 2 + ELU 2 backward -> dz2
      {
        T s1[R], s2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          s1[r] = T(0); s2[r] = T(0);
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            a_g2[j] = fma(gcarry[r][j], xh[r][j], a_g2[j]);
            a_be2[j] += gcarry[r][j];
            const T dx = gcarry[r][j] * g2[j];
            gcarry[r][j] = dx;
            s1[r] += dx;
            s2[r] = fma(dx, xh[r][j], s2[r]);
          }
        }
        warp_sum_rows<T, R>(s1);
        warp_sum_rows<T, R>(s2);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const T m1 = s1[r] / T(W), m2 = s2[r] / T(W);
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const T de = rstd2[r] * (gcarry[r][j] - m1 - xh[r][j] * m2);
            const T dz = de * eg[r][j];
            gcarry[r][j] = dz;
            a_b2[j] += dz;
          }
        }
      }
      store_tile<T, W, R>(g2T, lane, gcarry);
      __syncwarp();
      // stream dz2 out for the dW_hid GEMM
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (env0 + r < N) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) G.dz2_out[((int64_t)t * N + env0 + r) * W + lane + 32 * j] = gcarry[r][j];
        }
      }
      // dh1 = dz2 W_hid^T   (the layer-1 stash is fetched first so its latency hides behind the GEMM)
      T e1v[R][CPL];
      fetch_tile<T, W, R>(A.stash_e1, t, N, env0, lane, e1v);
      tile_gemm_t<T, W, R>(g2T, Whid, lane, gcarry);
      // LayerNorm 1 + ELU 1 backward -> dz1
      unstash<T, W, R>(e1v, mean1, rstd1, xh, eg);
      T g1[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) g1[j] = net.gamma[lane + 32 * j];
      {
        T s1[R], s2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          s1[r] = T(0); s2[r] = T(0);
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            a_g1[j] = fma(gcarry[r][j], xh[r][j], a_g1[j]);
            a_be1[j] += gcarry[r][j];
            const T dx = gcarry[r][j] * g1[j];
            gcarry[r][j] = dx;
            s1[r] += dx;
            s2[r] = fma(dx, xh[r][j], s2[r]);
          }
        }
        warp_sum_rows<T, R>(s1);
        warp_sum_rows<T, R>(s2);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const T m1 = s1[r] / T(W), m2 = s2[r] / T(W);
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const T de = rstd1[r] * (gcarry[r][j] - m1 - xh[r][j] * m2);
            const T dz = de * eg[r][j];
            gcarry[r][j] = dz;
            a_b1[j] += dz;
          }
        }
      }
    }
    // input layer: dW_in += obs^T dz1 ; dobs = dz1 W_in^T (reduced over columns across the warp)
    T gobs_new[NO];
#pragma unroll
    for (int i = 0; i < NO; ++i) {
      T wi[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) wi[j] = net.w_in[i * (W + 1) + lane + 32 * j];
      T p[R], obr[R];
      load_rows<R>(obsT + i * R, obr);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        p[r] = T(0);
#pragma unroll
        for (int j = 0; j < CPL; ++j) { p[r] = fma(gcarry[r][j], wi[j], p[r]); a_win[i][j] = fma(obr[r], gcarry[r][j], a_win[i][j]); }
      }
      warp_sum_rows<T, R>(p);
      T mine = T(0);
#pragma unroll
      for (int r = 0; r < R; ++r) mine = (row_of_lane == r) ? p[r] : mine;
      gobs_new[i] = mine;
    }
#pragma unroll
    for (int i = 0; i < NO; ++i) gobs[i] = valid ? gobs_new[i] : T(0);
    __syncwarp();
  }
  // ---- flush the per-lane parameter gradients ----------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    atomicAdd(G.g_b_hid + c, a_b1[j]);  atomicAdd(G.g_b_hid + W + c, a_b2[j]);
    atomicAdd(G.g_gamma + c, a_g1[j]);  atomicAdd(G.g_gamma + W + c, a_g2[j]);
    atomicAdd(G.g_beta + c, a_be1[j]);  atomicAdd(G.g_beta + W + c, a_be2[j]);
#pragma unroll
    for (int q = 0; q < NA; ++q) atomicAdd(G.g_w_out + q * W + c, a_wout[j][q]);
#pragma unroll
    for (int i = 0; i < NO; ++i) atomicAdd(G.g_w_in + c * NO + i, a_win[i][j]);
  }
  if (owner) {
#pragma unroll
    for (int q = 0; q < NA; ++q) { atomicAdd(G.g_b_out + q, a_bout[q]); atomicAdd(G.g_log_std + q, a_lstd[q]); }
  }
}

// ---------------------------------------------------------------------------------------------------------
// input gradient of a scalar-output MLP over M rows:  gx[row] = w[row] * d net(x[row]) / d x[row]
// (the critic's terminal-value terms of the SHAC loss, shac.py:137-147: d(sum_k weight_k V'(obs_k)) / d obs_k, which
// the reverse pass of the rollout needs as a cotangent of the terminal observation rows).  Forward with the ELU
// outputs stashed in warp-private shared memory, then the transposed pass; no parameter gradients.
template <class T, int W, int R>
__global__ void __launch_bounds__(R == 8 ? 256 : 512)
mlp_rows_vjp_kernel(int64_t M, const T* __restrict__ x, const T* __restrict__ wrow, T* __restrict__ gx,
                    const __grid_constant__ MlpParams<T> P) {
  constexpr int CPL = W / 32;
  extern __shared__ __align__(32) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  NetSmem<T, W> net;
  T* after = net.load(smem, P);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int in_dim = net.in_dim, nh = net.n_hidden;
  size_t off = (size_t)(after - smem);
  off = (off + 7) & ~size_t(7);
  // warp-private: inT [in_dim][R], act [W][R], stash e_l [nh][W][R]
  const size_t per_warp = ((size_t)(in_dim + (1 + nh) * W) * R + 7) & ~size_t(7);
  T* inT = smem + off + warp * per_warp;
  T* act = inT + in_dim * R;
  T* eS = act + W * R;
  const int64_t n_blocks = (M + R - 1) / R;
  for (int64_t blk = (int64_t)blockIdx.x * nwarps + warp; blk < n_blocks; blk += (int64_t)gridDim.x * nwarps) {
    const int64_t row0 = blk * R;
    const int total = in_dim * R;
    {
      // rows whose weight is zero contribute nothing (unused terminal-row slots of the static episode graph): the
      // whole tile is skipped when all of its rows are (warp-uniform test)
      bool any = false;
#pragma unroll
      for (int r = 0; r < R; ++r) any = any || (row0 + r < M && wrow[row0 + r] != T(0));
      if (!any) {
        for (int i = lane; i < total; i += 32)
          if (row0 * in_dim + i < M * in_dim) gx[row0 * in_dim + i] = T(0);
        continue;
      }
    }
    for (int i = lane; i < total; i += 32) {
      const int r = i / in_dim, c = i % in_dim;
      inT[c * R + r] = (row0 + r < M) ? x[row0 * in_dim + i] : T(0);
    }
    __syncwarp();
    // ---- forward, keeping the ELU outputs and the LayerNorm statistics ----------------------------------------
    T mean[MAX_HIDDEN][R], rstd[MAX_HIDDEN][R];
    T h[R][CPL], ev[R][CPL];
    tile_gemm<T, W, R>(inT, in_dim, net.w_in, net.b_hid, lane, h);
    elu_layernorm<T, W, R, true>(h, net.gamma, net.beta, lane, ev, mean[0], rstd[0]);
    store_tile<T, W, R>(eS, lane, ev);
#pragma unroll
    for (int l = 1; l < MAX_HIDDEN; ++l) {
      if (l < nh) {
        store_tile<T, W, R>(act, lane, h);
        __syncwarp();
        tile_gemm<T, W, R>(act, W, net.w_hid + (size_t)(l - 1) * W * (W + 1), net.b_hid + l * W, lane, h);
        elu_layernorm<T, W, R, true>(h, net.gamma + l * W, net.beta + l * W, lane, ev, mean[l], rstd[l]);
        store_tile<T, W, R>(eS + (size_t)l * W * R, lane, ev);
        __syncwarp();
      }
    }
    // ---- backward w.r.t. the input -----------------------------------------------------------------------------
    T wr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) wr[r] = (row0 + r < M) ? wrow[row0 + r] : T(0);
    T gc[R][CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const T wo = net.w_out[(lane + 32 * j) * net.out_dim];
#pragma unroll
      for (int r = 0; r < R; ++r) gc[r][j] = wr[r] * wo;
    }
#pragma unroll
    for (int l = MAX_HIDDEN - 1; l >= 0; --l) {
      if (l < nh) {
        T xh[R][CPL], eg[R][CPL];
        load_tile<T, W, R>(eS + (size_t)l * W * R, lane, eg);
        unstash<T, W, R>(eg, mean[l], rstd[l], xh, eg);
        T s1[R], s2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          s1[r] = T(0); s2[r] = T(0);
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const T dx = gc[r][j] * net.gamma[l * W + lane + 32 * j];
            gc[r][j] = dx;
            s1[r] += dx;
            s2[r] = fma(dx, xh[r][j], s2[r]);
          }
        }
        warp_sum_rows<T, R>(s1);
        warp_sum_rows<T, R>(s2);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const T m1 = s1[r] / T(W), m2 = s2[r] / T(W);
#pragma unroll
          for (int j = 0; j < CPL; ++j) gc[r][j] = (rstd[l][r] * (gc[r][j] - m1 - xh[r][j] * m2)) * eg[r][j];
        }
        if (l > 0) {
          store_tile<T, W, R>(act, lane, gc);
          __syncwarp();
          tile_gemm_t<T, W, R>(act, net.w_hid + (size_t)(l - 1) * W * (W + 1), lane, gc);
          __syncwarp();
        }
      }
    }
    // input layer: gx[r][i] = sum_c dz1[r][c] w_in[i][c]
    for (int i = 0; i < in_dim; ++i) {
      T p[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        p[r] = T(0);
#pragma unroll
        for (int j = 0; j < CPL; ++j) p[r] = fma(gc[r][j], net.w_in[i * (W + 1) + lane + 32 * j], p[r]);
      }
      warp_sum_rows<T, R>(p);
      if (lane < R && row0 + lane < M) {
        T mine = T(0);
#pragma unroll
        for (int r = 0; r < R; ++r) mine = (lane == r) ? p[r] : mine;
        gx[(row0 + lane) * in_dim + i] = mine;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// launch shapes.  Shared memory per CTA = network parameters (shared by the CTA's warps) + per-warp tiles.
constexpr size_t SMEM_CAP = 227 * 1024;

template <class T, int W> size_t net_bytes(int in_dim, int n_hidden, int out_dim) {
  size_t f = NetSmem<T, W>::floats(in_dim, n_hidden, out_dim);
  f = (f + 7) & ~size_t(7);
  return f * sizeof(T) + 32;
}
template <class T> size_t warp_bytes(int leading, int tiles, int W, int R) {      // (leading + tiles * W) * R, 32-byte granules
  size_t f = (size_t)(leading + tiles * W) * R;
  f = (f + 7) & ~size_t(7);
  return f * sizeof(T);
}

inline int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}
inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

struct Shape { int R, warps, grid; size_t smem; };

// Rows per warp and warps per CTA for `units` independent rows (environments).  The time loop is a serial
// dependency chain per warp, so throughput comes from interleaving warps on every SM sub-partition: prefer
// 8 rows per warp (most FMAs per shared-memory weight read) when that still gives every SM >= 6 warps,
// otherwise fewer rows per warp.  SGBPO_ROLLOUT_R / SGBPO_ROLLOUT_WARPS override (tuning, tests).
template <class T>
Shape pick_shape(int64_t units, size_t fixed_bytes, int leading, int tiles, int W, const int* max_warps_of_R,
                 bool rows_kernel = false) {
  const int forced_R = env_int(rows_kernel ? "SGBPO_ROWS_R" : "SGBPO_ROLLOUT_R");
  const int forced_w = env_int(rows_kernel ? "SGBPO_ROWS_WARPS" : "SGBPO_ROLLOUT_WARPS");
  const int enough = env_int("SGBPO_ROLLOUT_ENOUGH") > 0 ? env_int("SGBPO_ROLLOUT_ENOUGH") : 6;
  const int sms = sm_count();
  Shape best{0, 0, 0, 0};
  const int cand[3] = {8, 4, 2};
  for (int ci = 0; ci < 3; ++ci) {
    const int R = cand[ci];
    if (forced_R && R != forced_R) continue;
    const size_t wb = warp_bytes<T>(leading, tiles, W, R);
    if (fixed_bytes + wb > SMEM_CAP) continue;
    int wmax = (int)((SMEM_CAP - fixed_bytes) / wb);
    if (wmax > max_warps_of_R[ci]) wmax = max_warps_of_R[ci];
    if (wmax < 1) continue;
    const int64_t total = (units + R - 1) / R;
    int w = (int)((total + sms - 1) / sms);
    if (w < 1) w = 1;
    if (w > wmax) w = wmax;
    if (forced_w) w = forced_w < wmax ? forced_w : wmax;
    const Shape s{R, w, (int)((total + w - 1) / w), fixed_bytes + (size_t)w * wb};
    if (!best.R || w > best.warps) best = s;
    if (w >= enough) break;
  }
  return best;
}

template <int R> struct RowsShape { static constexpr int MAX_WARPS = R == 8 ? 8 : 16; };

static int cuda_fail2(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return SGBPO_E_CUDA;
}
static int fail2(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

template <class T> MlpParams<T> convert_mlp(const sgbpo_mlp& m) {
  MlpParams<T> p;
  p.in_dim = m.in_dim; p.width = m.width; p.n_hidden = m.n_hidden; p.out_dim = m.out_dim;
  for (int l = 0; l <= MAX_HIDDEN; ++l) { p.weight[l] = (const T*)m.weight[l]; p.bias[l] = (const T*)m.bias[l]; }
  for (int l = 0; l < MAX_HIDDEN; ++l) { p.ln_weight[l] = (const T*)m.ln_weight[l]; p.ln_bias[l] = (const T*)m.ln_bias[l]; }
  p.log_std = (const T*)m.log_std;
  return p;
}
template <class T> RolloutArgs<T> convert_rollout(const sgbpo_rollout& r) {
  RolloutArgs<T> a;
  a.N = r.N; a.H = r.H;
  a.eps = (const T*)r.eps; a.noise = (const T*)r.noise; a.dirs = (const T*)r.dirs; a.aux0 = (const T*)r.aux0;
  a.reset_states = (const T*)r.reset_states; a.reset_idx = r.reset_idx; a.aux_src = r.aux_src;
  a.shared = (const StepShared<T>*)r.shared;
  a.states = (T*)r.states; a.obs = (T*)r.obs; a.actions = (T*)r.actions; a.exec_actions = (T*)r.exec_actions;
  a.rewards = (T*)r.rewards; a.flags = r.flags;
  a.stash_h1 = (T*)r.stash_h1; a.stash_e1 = (T*)r.stash_e1; a.stash_e2 = (T*)r.stash_e2; a.stash_stats = (T*)r.stash_stats;
  a.stash_tiled = (r.engine & 2) ? 1 : 0;
  return a;
}
template <class T> BackwardArgs<T> convert_backward(const sgbpo_rollout_grads& g) {
  BackwardArgs<T> b;
  b.grw = (const T*)g.grw; b.gobs_term = (const T*)g.gobs_term; b.term_of_row = g.term_of_row;
  b.dz2_out = (T*)g.dz2_out;
  b.g_w_in = (T*)g.g_w_in; b.g_w_out = (T*)g.g_w_out; b.g_b_hid = (T*)g.g_b_hid; b.g_gamma = (T*)g.g_gamma;
  b.g_beta = (T*)g.g_beta; b.g_b_out = (T*)g.g_b_out; b.g_log_std = (T*)g.g_log_std;
  return b;
}

template <int ENV, class T, int W> struct RolloutLaunch {
  template <int R>
  static int fwd_R(const Shape& sh, const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st) {
    auto kern = rollout_fwd_kernel<ENV, T, W, R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    if (e != cudaSuccess) return cuda_fail2(e, "cudaFuncSetAttribute(rollout_fwd)");
    kern<<<sh.grid, sh.warps * 32, sh.smem, st>>>(convert_rollout<T>(*r), convert_mlp<T>(*pol), convert_consts<T>(*c));
    e = cudaGetLastError();
    return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "rollout_fwd_kernel");
  }
  template <int R>
  static int bwd_R(const Shape& sh, const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r,
                   const sgbpo_rollout_grads* g, cudaStream_t st) {
    auto kern = rollout_bwd_kernel<ENV, T, W, R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    if (e != cudaSuccess) return cuda_fail2(e, "cudaFuncSetAttribute(rollout_bwd)");
    kern<<<sh.grid, sh.warps * 32, sh.smem, st>>>(convert_rollout<T>(*r), convert_backward<T>(*g), convert_mlp<T>(*pol),
                                                  convert_consts<T>(*c));
    e = cudaGetLastError();
    return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "rollout_bwd_kernel");
  }
  static int fwd(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st) {
    using TASK = Task<ENV>;
    if (pol->in_dim != TASK::NO || pol->out_dim != TASK::NA) return fail2(SGBPO_E_ARG, "policy dims do not match the environment");
    if (pol->n_hidden < 1 || pol->n_hidden > MAX_HIDDEN) return fail2(SGBPO_E_UNSUPPORTED, "1..4 hidden layers");
    const int caps[3] = {FwdShape<T, 8>::MAX_WARPS, FwdShape<T, 4>::MAX_WARPS, FwdShape<T, 2>::MAX_WARPS};
    const Shape sh = pick_shape<T>(r->N, net_bytes<T, W>(pol->in_dim, pol->n_hidden, pol->out_dim), TASK::NO, 2, W, caps);
    if (!sh.R) return fail2(SGBPO_E_UNSUPPORTED, "policy does not fit in shared memory");
    switch (sh.R) {
      case 8: return fwd_R<8>(sh, c, pol, r, st);
      case 4: return fwd_R<4>(sh, c, pol, r, st);
      default: return fwd_R<2>(sh, c, pol, r, st);
    }
  }
  static int bwd(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, const sgbpo_rollout_grads* g,
                 cudaStream_t st) {
    using TASK = Task<ENV>;
    if (pol->n_hidden < 1 || pol->n_hidden > MAX_HIDDEN) return fail2(SGBPO_E_UNSUPPORTED, "1..4 hidden layers");
    if (pol->n_hidden != 2) return fail2(SGBPO_E_UNSUPPORTED, "fused backward supports policies with two hidden layers");
    const int caps[3] = {RolloutShape<8>::MAX_WARPS, RolloutShape<4>::MAX_WARPS, RolloutShape<2>::MAX_WARPS};
    const Shape sh = pick_shape<T>(r->N, net_bytes<T, W>(pol->in_dim, pol->n_hidden, pol->out_dim), TASK::NO, 1, W, caps);
    if (!sh.R) return fail2(SGBPO_E_UNSUPPORTED, "policy does not fit in shared memory (backward)");
    switch (sh.R) {
      case 8: return bwd_R<8>(sh, c, pol, r, g, st);
      case 4: return bwd_R<4>(sh, c, pol, r, g, st);
      default: return bwd_R<2>(sh, c, pol, r, g, st);
    }
  }
};

template <class T, int W, int R>
int launch_rows_R(const Shape& sh, int64_t M, const sgbpo_mlp* net, const void* x, void* out, cudaStream_t st) {
  auto kern = mlp_rows_kernel<T, W, R>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
  if (e != cudaSuccess) return cuda_fail2(e, "cudaFuncSetAttribute(mlp_rows)");
  int64_t blocks = sh.grid;
  const int64_t cap = (int64_t)sm_count() * (sh.smem > 110 * 1024 ? 1 : 2);     // persistent: warps stride over row blocks
  if (blocks > cap) blocks = cap;
  kern<<<(int)blocks, sh.warps * 32, sh.smem, st>>>(M, net->in_dim, (const T*)x, (T*)out, convert_mlp<T>(*net));
  e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "mlp_rows_kernel");
}

template <class T, int W> int launch_rows(int64_t M, const sgbpo_mlp* net, const void* x, void* out, cudaStream_t st) {
  if (net->n_hidden < 1 || net->n_hidden > MAX_HIDDEN) return fail2(SGBPO_E_UNSUPPORTED, "1..4 hidden layers");
  const int caps[3] = {RowsShape<8>::MAX_WARPS, RowsShape<4>::MAX_WARPS, 0};
  const Shape sh = pick_shape<T>(M, net_bytes<T, W>(net->in_dim, net->n_hidden, net->out_dim), net->in_dim, 2, W, caps, true);
  if (!sh.R) return fail2(SGBPO_E_UNSUPPORTED, "network does not fit in shared memory");
  return sh.R == 8 ? launch_rows_R<T, W, 8>(sh, M, net, x, out, st) : launch_rows_R<T, W, 4>(sh, M, net, x, out, st);
}


template <class T, int W, int R>
int launch_vjp_R(const Shape& sh, int64_t M, const sgbpo_mlp* net, const void* x, const void* w, void* gx, cudaStream_t st) {
  auto kern = mlp_rows_vjp_kernel<T, W, R>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
  if (e != cudaSuccess) return cuda_fail2(e, "cudaFuncSetAttribute(mlp_rows_vjp)");
  int64_t blocks = sh.grid;
  const int64_t cap = (int64_t)sm_count() * (sh.smem > 110 * 1024 ? 1 : 2);
  if (blocks > cap) blocks = cap;
  kern<<<(int)blocks, sh.warps * 32, sh.smem, st>>>(M, (const T*)x, (const T*)w, (T*)gx, convert_mlp<T>(*net));
  e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "mlp_rows_vjp_kernel");
}

template <class T, int W> int launch_vjp(int64_t M, const sgbpo_mlp* net, const void* x, const void* w, void* gx, cudaStream_t st) {
  if (net->n_hidden < 1 || net->n_hidden > MAX_HIDDEN) return fail2(SGBPO_E_UNSUPPORTED, "1..4 hidden layers");
  if (net->out_dim != 1) return fail2(SGBPO_E_UNSUPPORTED, "sgbpo_mlp_rows_vjp: scalar-output networks");
  const int caps[3] = {RowsShape<8>::MAX_WARPS, RowsShape<4>::MAX_WARPS, 0};
  const Shape sh = pick_shape<T>(M, net_bytes<T, W>(net->in_dim, net->n_hidden, net->out_dim), net->in_dim, 1 + net->n_hidden, W,
                                 caps, true);
  if (!sh.R) return fail2(SGBPO_E_UNSUPPORTED, "network does not fit in shared memory");
  return sh.R == 8 ? launch_vjp_R<T, W, 8>(sh, M, net, x, w, gx, st) : launch_vjp_R<T, W, 4>(sh, M, net, x, w, gx, st);
}

// one translation unit per (environment, dtype): SGBPO_INSTANTIATE_ROLLOUT_ENV_T(E, T, SUFFIX) in sgbpo_rollout_env<E>_<f|d>.cu
template <int ENV, class T> int rollout_fwd_env_t(int width, const sgbpo_consts* c, const sgbpo_mlp* pol,
                                                  const sgbpo_rollout* r, cudaStream_t st);
template <int ENV, class T> int rollout_bwd_env_t(int width, const sgbpo_consts* c, const sgbpo_mlp* pol,
                                                  const sgbpo_rollout* r, const sgbpo_rollout_grads* g, cudaStream_t st);

#define SGBPO_INSTANTIATE_ROLLOUT_ENV_T(E, T)                                                                  \
  template <> int rollout_fwd_env_t<E, T>(int width, const sgbpo_consts* c, const sgbpo_mlp* pol,                \
                                          const sgbpo_rollout* r, cudaStream_t st) {                              \
    switch (width) {                                                                                              \
      case 32: return RolloutLaunch<E, T, 32>::fwd(c, pol, r, st);                                                \
      case 64: return RolloutLaunch<E, T, 64>::fwd(c, pol, r, st);                                                \
      case 128: return RolloutLaunch<E, T, 128>::fwd(c, pol, r, st);                                              \
      default: return fail2(SGBPO_E_UNSUPPORTED, "fused rollout supports hidden widths 32, 64, 128");            \
    }                                                                                                             \
  }                                                                                                               \
  template <> int rollout_bwd_env_t<E, T>(int width, const sgbpo_consts* c, const sgbpo_mlp* pol,                \
                                          const sgbpo_rollout* r, const sgbpo_rollout_grads* g, cudaStream_t st) { \
    switch (width) {                                                                                              \
      case 32: return RolloutLaunch<E, T, 32>::bwd(c, pol, r, g, st);                                             \
      case 64: return RolloutLaunch<E, T, 64>::bwd(c, pol, r, g, st);                                             \
      case 128: return RolloutLaunch<E, T, 128>::bwd(c, pol, r, g, st);                                           \
      default: return fail2(SGBPO_E_UNSUPPORTED, "fused rollout supports hidden widths 32, 64, 128");            \
    }                                                                                                             \
  }

template <int ENV> inline int rollout_fwd_env(int dtype, int width, const sgbpo_consts* c, const sgbpo_mlp* pol,
                                              const sgbpo_rollout* r, cudaStream_t st) {
  return dtype ? rollout_fwd_env_t<ENV, double>(width, c, pol, r, st) : rollout_fwd_env_t<ENV, float>(width, c, pol, r, st);
}
template <int ENV> inline int rollout_bwd_env(int dtype, int width, const sgbpo_consts* c, const sgbpo_mlp* pol,
                                              const sgbpo_rollout* r, const sgbpo_rollout_grads* g, cudaStream_t st) {
  return dtype ? rollout_bwd_env_t<ENV, double>(width, c, pol, r, g, st) : rollout_bwd_env_t<ENV, float>(width, c, pol, r, g, st);
}

}  // namespace sgbpo
