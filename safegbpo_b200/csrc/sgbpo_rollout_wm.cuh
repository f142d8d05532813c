// Whole-horizon forward kernel, warp-MMA engine (forward engine 2, `engine` bit 2): the policy MLP of every step on warp-level
// tensor-core MMAs (mma.sync.m16n8k16, fp16 operands split hi + lo, fp32 accumulate) with NO block-wide synchronisation inside
// the time loop.  Reference path: collect_trajectories (shac.py:84-117) = policy forward (policy.py:50-89) + Safeguard.actions
// + Simulator.step per step, the same work as rollout_fwd_kernel (sgbpo_rollout_impl.cuh) and rollout_tc_fwd_kernel.
//
// Why a third forward engine.  At the headline size (4096 environments) one step is a latency chain: the FMA register-tile
// kernel spends ~5 100 warp-instructions per warp-step (2 000 of them FFMA for the 128 x 128 layer) at one issue every ~3.5
// cycles, and the tcgen05 kernel pays a shared-memory -> MMA -> commit -> mbarrier -> tcgen05.ld round trip plus CTA barriers per
// layer and needs 128 rows per CTA (32 CTAs at this size).  Here a warp owns 8 environments for the whole horizon:
//   * A operand = the 8 rows' activations, hi halves in MMA rows 0-7 and lo halves in rows 8-15, so ONE mma against B_hi gives
//     A_hi B_hi (rows 0-7) and A_lo B_hi (rows 8-15) and one against B_lo gives A_hi B_lo: two MMAs per (k-tile, n-tile) for the
//     three-product split (288 per step instead of 2 048 FFMA + 128 LDS.128 for 4 rows);
//   * the m16n8 accumulator layout of layer 1 IS the A-fragment layout of layer 2 (row = lane / 4, column pairs 2 (lane % 4)),
//     so activations never leave registers between the layers; LayerNorm row sums are two shuffles inside a quad;
//   * B fragments (W_hid, and W_in with the bias folded in as an extra k row against a constant-one input) are prepared once per
//     CTA in shared memory in fragment order: one conflict-free LDS.128 per (k-tile, n-tile) brings hi and lo of both registers;
//   * the environment step runs on the 8 owner lanes (lane % 4 == 0) of the warp, as in the FMA kernel.
// fp32 policies [in <= 15 -> 128 -> 128 -> m] (in = the task's observation width); other shapes use the other engines.
#pragma once
#include <cuda_fp16.h>

#include "sgbpo_rollout_impl.cuh"

namespace sgbpo {
namespace wm {

constexpr int W = 128, R = 8, WARPS = 4, NT = W / 8, KT = W / 16;

__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) -> packed fp16 pair hi and the packed residual lo:  x = hi.x + lo.x up to 2^-22 |x|
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 b = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - b.x, y - b.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float elu_fast(float x) {      // exp(x) - 1 on the negative side: one multiply + MUFU.EX2 (2 ulp)
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return x > 0.f ? x : e - 1.f;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

struct WmSmem {
  uint4* w2f;      // [KT][NT][32]  {hi b0, hi b1, lo b0, lo b1} of W_hid for (k-tile, n-tile, lane)
  uint4* w1f;      // [NT][32]      the same for the input layer, K padded to 16, k row `in_dim` = Linear1.bias
  float* b2; float* g1; float* be1; float* g2; float* be2;      // [W]
  float* w_out;    // [NA][W]
  float* b_out;    // [NA]
  __host__ __device__ static size_t bytes(int out_dim) {
    return sizeof(uint4) * ((size_t)KT * NT * 32 + (size_t)NT * 32) + sizeof(float) * (5 * W + (size_t)out_dim * W + out_dim);
  }
  __device__ void carve(unsigned char* base, int out_dim) {
    w2f = reinterpret_cast<uint4*>(base);
    w1f = w2f + KT * NT * 32;
    float* f = reinterpret_cast<float*>(w1f + NT * 32);
    b2 = f; f += W; g1 = f; f += W; be1 = f; f += W; g2 = f; f += W; be2 = f; f += W;
    w_out = f; f += out_dim * W;
    b_out = f;
  }
};

// ELU, stash, LayerNorm of the 32 values of this lane (row g, column pairs nt * 8 + 2 tig); z holds acc hi-rows + lo-rows summed.
// On return h = LayerNorm(ELU(z)), mean / rstd are the row statistics (identical on the four lanes of the quad).
__device__ __forceinline__ void elu_layernorm_frag(float (&h)[NT][2], const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   int tig, float* __restrict__ stash, size_t nt_stride, bool store,
                                                   float& mean, float& rstd) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    h[nt][0] = elu_fast(h[nt][0]);
    h[nt][1] = elu_fast(h[nt][1]);
    s[nt & 3] += h[nt][0] + h[nt][1];
    if (store) *reinterpret_cast<float2*>(stash + nt * nt_stride) = make_float2(h[nt][0], h[nt][1]);
  }
  mean = quad_sum((s[0] + s[1]) + (s[2] + s[3])) * (1.f / W);
  float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    h[nt][0] -= mean;
    h[nt][1] -= mean;
    q[nt & 3] = fmaf(h[nt][0], h[nt][0], q[nt & 3]);
    q[nt & 3] = fmaf(h[nt][1], h[nt][1], q[nt & 3]);
  }
  rstd = rsqrtf(quad_sum((q[0] + q[1]) + (q[2] + q[3])) * (1.f / W) + 1e-5f);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const float2 g = *reinterpret_cast<const float2*>(gamma + nt * 8 + 2 * tig);
    const float2 b = *reinterpret_cast<const float2*>(beta + nt * 8 + 2 * tig);
    h[nt][0] = (h[nt][0] * rstd) * g.x + b.x;
    h[nt][1] = (h[nt][1] * rstd) * g.y + b.y;
  }
}

template <int ENV>
__global__ void __launch_bounds__(WARPS * 32)
rollout_wm_fwd_kernel(const __grid_constant__ RolloutArgs<float> A, const __grid_constant__ MlpParams<float> P,
                      const __grid_constant__ SafeConsts<float> k) {
  using TASK = Task<ENV>;
  using T = float;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX;
  static_assert(NO < 16, "the input layer is one k-tile: observation width + the bias row must fit in 16");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WmSmem S;
  S.carve(smem_raw, NA);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;

  // ---- one-time: B fragments of both layers (hi / lo split), small vectors ------------------------------------------------------
  for (int idx = tid; idx < (KT + 1) * NT * 32; idx += WARPS * 32) {
    const int l = idx & 31, nt = (idx >> 5) % NT, kt = (idx >> 5) / NT;          // kt == KT: the input layer
    const int n = nt * 8 + (l >> 2), k0 = 2 * (l & 3);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k0 + (j & 1) + (j >> 1) * 8;                                 // b0: k0, k0 + 1;  b1: k0 + 8, k0 + 9
      if (kt < KT) v[j] = P.weight[1][(size_t)n * W + kt * 16 + kk];
      else v[j] = kk < NO ? P.weight[0][(size_t)n * NO + kk] : (kk == NO ? P.bias[0][n] : 0.f);
    }
    uint4 f;
    split_pair(v[0], v[1], f.x, f.z);
    split_pair(v[2], v[3], f.y, f.w);
    (kt < KT ? S.w2f + ((size_t)kt * NT + nt) * 32 : S.w1f + (size_t)nt * 32)[l] = f;
  }
  for (int i = tid; i < W; i += WARPS * 32) {
    S.b2[i] = P.bias[1][i];
    S.g1[i] = P.ln_weight[0][i]; S.be1[i] = P.ln_bias[0][i];
    S.g2[i] = P.ln_weight[1][i]; S.be2[i] = P.ln_bias[1][i];
  }
  for (int i = tid; i < NA * W; i += WARPS * 32) S.w_out[i] = P.weight[2][i];
  for (int i = tid; i < NA; i += WARPS * 32) S.b_out[i] = P.bias[2][i];
  __syncthreads();

  const int64_t N = A.N;
  const int64_t env = ((int64_t)blockIdx.x * WARPS + warp) * R + g;       // the row of this quad
  const bool owner = tig == 0;
  const bool row_ok = env < N, valid = owner && row_ok;
  const int src = lane & ~3;                                              // the quad's owner lane
  const bool tiled = A.stash_tiled != 0, stash = A.stash_e1 != nullptr;
  const size_t nt_stride = tiled ? 1024 : 8;                              // floats between the column pairs of consecutive n-tiles

  T s[NS], aux[NAUX > 0 ? NAUX : 1], ob[NO];
  T sigma[NA];
#pragma unroll
  for (int q = 0; q < NA; ++q) sigma[q] = sg_exp(P.log_std[q]);
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = valid ? A.states[env * NS + i] : T(0);
  if (NAUX > 0) {
#pragma unroll
    for (int i = 0; i < NAUX; ++i) aux[i] = valid ? A.aux0[env * NAUX + i] : T(0);
  }
#pragma unroll
  for (int i = 0; i < NO; ++i) ob[i] = valid ? A.obs[env * NO + i] : T(0);

  for (int t = 0; t < A.H; ++t) {
    // this step's streaming loads first: their latency hides behind the MLP
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
    // ---- input layer: A = [obs, 1, 0 ...] of row g (hi in rows 0-7, lo in rows 8-15), one k-tile ----------------------------------
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = i < NO ? __shfl_sync(0xffffffffu, ob[i < NO ? i : 0], src) : (i == NO ? 1.f : 0.f);
    uint32_t a[4];
    {
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;        // x[2 tig], x[2 tig + 1], x[8 + 2 tig], x[9 + 2 tig]
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (tig == j) { p0 = x[2 * j]; p1 = x[2 * j + 1]; p2 = x[8 + 2 * j]; p3 = x[9 + 2 * j]; }
      }
      split_pair(p0, p1, a[0], a[1]);
      split_pair(p2, p3, a[2], a[3]);
    }
    float h[NT][2];
    {
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        const uint4 b = S.w1f[nt * 32 + lane];
        hmma(acc[nt], a, b.x, b.y);
        hmma(acc[nt], a, b.z, b.w);
        h[nt][0] = acc[nt][0] + acc[nt][2];
        h[nt][1] = acc[nt][1] + acc[nt][3];
      }
    }
    const int64_t row = (int64_t)t * N + (row_ok ? env : 0);
    const size_t stash_off = stash_index<W>(tiled, row, 2 * tig);
    float st[4];
    elu_layernorm_frag(h, S.g1, S.be1, tig, stash ? A.stash_e1 + stash_off : nullptr, nt_stride, stash && row_ok, st[0], st[1]);
    // ---- hidden layer: the accumulator layout of layer 1 is the A-fragment layout of layer 2 --------------------------------------
    {
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float2 b = *reinterpret_cast<const float2*>(S.b2 + nt * 8 + 2 * tig);
        acc[nt][0] = b.x; acc[nt][1] = b.y; acc[nt][2] = 0.f; acc[nt][3] = 0.f;
      }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        split_pair(h[2 * kt][0], h[2 * kt][1], a[0], a[1]);
        split_pair(h[2 * kt + 1][0], h[2 * kt + 1][1], a[2], a[3]);
        const uint4* bf = S.w2f + (size_t)kt * NT * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint4 b = bf[nt * 32];
          hmma(acc[nt], a, b.x, b.y);
          hmma(acc[nt], a, b.z, b.w);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        h[nt][0] = acc[nt][0] + acc[nt][2];
        h[nt][1] = acc[nt][1] + acc[nt][3];
      }
    }
    elu_layernorm_frag(h, S.g2, S.be2, tig, stash ? A.stash_e2 + stash_off : nullptr, nt_stride, stash && row_ok, st[2], st[3]);
    if (stash && valid) *reinterpret_cast<float4*>(A.stash_stats + row * 4) = make_float4(st[0], st[1], st[2], st[3]);
    // ---- output layer ------------------------------------------------------------------------------------------------------------
    T mu[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float2 w = *reinterpret_cast<const float2*>(S.w_out + q * W + nt * 8 + 2 * tig);
        o[nt & 3] = fmaf(h[nt][0], w.x, o[nt & 3]);
        o[nt & 3] = fmaf(h[nt][1], w.y, o[nt & 3]);
      }
      mu[q] = quad_sum((o[0] + o[1]) + (o[2] + o[3])) + S.b_out[q];
    }
    // ---- environment step on the owner lanes -------------------------------------------------------------------------------------
    if (owner) {
      T act[NA];
#pragma unroll
      for (int q = 0; q < NA; ++q) act[q] = valid ? sg_tanh(mu[q] + sigma[q] * epsv[q]) : T(0);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < NS; ++i) wv[i] = T(0);
      }
      if (NAUX > 0 && ridx >= 0 && valid) {     // quadrotor: goal = reset position
#pragma unroll
        for (int i = 0; i < NAUX; ++i) aux[i] = rs[i];
      }
      T sn[NS], ae[NA], rew;
      int fl;
      const StepShared<T>* sh = A.shared ? A.shared + t : nullptr;
      StepShared<T> empty;
      if (!sh) { memset(&empty, 0, sizeof(empty)); sh = &empty; }
      safeguarded_step<TASK, T, T>(s, act, wv, aux, A.dirs ? dv : nullptr, sh, k, ridx >= 0, rs, sn, ae, ob, rew, fl);
#pragma unroll
      for (int i = 0; i < NS; ++i) s[i] = sn[i];
      if (valid) {
#pragma unroll
        for (int i = 0; i < NS; ++i) A.states[((int64_t)(t + 1) * N + env) * NS + i] = sn[i];
#pragma unroll
        for (int i = 0; i < NO; ++i) A.obs[((int64_t)(t + 1) * N + env) * NO + i] = ob[i];
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          A.actions[((int64_t)t * N + env) * NA + q] = act[q];
          A.exec_actions[((int64_t)t * N + env) * NA + q] = ae[q];
        }
        A.rewards[(int64_t)t * N + env] = rew;
        A.flags[(int64_t)t * N + env] = fl;
      }
    }
    __syncwarp();
  }
}

// ---- pair-split variant: TWO warps share 8 environments, each computes 64 of the 128 hidden columns ------------------------------
// Per step the two warps exchange (through a pair-private shared-memory block, under a 64-thread named barrier): the observations
// (owner -> both), the LayerNorm half statistics (mean_h, M2_h: combined exactly, var = (M2_0 + M2_1 + 32 (mean_0 - mean_1)^2) / 128),
// the A fragments of their 4 k-tiles of layer 2 and the partial output-layer dot products.  The MLP part of the per-step dependency
// chain halves (144 instead of 288 HMMA per warp, 16 instead of 32 elementwise values per lane); the environment step stays on the
// owner lanes of the pair's first warp.
constexpr int PAIRS = 4, NTH = NT / 2, KTH = KT / 2;       // 8 warps per CTA = 4 pairs = 32 environments

struct PairSmem {          // one per pair
  float obs[R][16];        // policy input rows (k-padded): [obs, 1, 0...]
  float2 stat[2][2][R];    // [layer][warp half][row] (mean_h, M2_h)
  float mu[2][R][4];       // [warp half][row][q] partial output-layer sums
  uint4 frag[2][KTH][32];  // [warp half][local k-tile][lane] A fragments (hi/lo stacked) of layer 2
};

__device__ __forceinline__ void pair_sync(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory"); }

// ELU + stash + half-row statistics of this lane's 16 values (row g, its warp's 64 columns); h is left centred on the HALF mean
__device__ __forceinline__ void elu_half_stats(float (&h)[NTH][2], float* __restrict__ stash, size_t nt_stride, bool store,
                                               float& mean_h, float& m2_h) {
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NTH; ++j) {
    h[j][0] = elu_fast(h[j][0]);
    h[j][1] = elu_fast(h[j][1]);
    s[j & 1] += h[j][0] + h[j][1];
    if (store) *reinterpret_cast<float2*>(stash + j * nt_stride) = make_float2(h[j][0], h[j][1]);
  }
  mean_h = quad_sum(s[0] + s[1]) * (2.f / W);
  float q[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NTH; ++j) {
    h[j][0] -= mean_h;
    h[j][1] -= mean_h;
    q[j & 1] = fmaf(h[j][0], h[j][0], q[j & 1]);
    q[j & 1] = fmaf(h[j][1], h[j][1], q[j & 1]);
  }
  m2_h = quad_sum(q[0] + q[1]);
}
// combine the two half statistics (exact pooled mean / variance) and apply LayerNorm to the half-centred values
__device__ __forceinline__ void layernorm_half(float (&h)[NTH][2], float mean_h, float m2_h, float2 other, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, int col0, float& mean, float& rstd) {
  const float d = mean_h - other.x;
  mean = 0.5f * (mean_h + other.x);
  rstd = rsqrtf((m2_h + other.y + (W / 4) * d * d) * (1.f / W) + 1e-5f);
  const float shift = mean_h - mean;
#pragma unroll
  for (int j = 0; j < NTH; ++j) {
    const float2 gm = *reinterpret_cast<const float2*>(gamma + col0 + j * 8);
    const float2 bt = *reinterpret_cast<const float2*>(beta + col0 + j * 8);
    h[j][0] = ((h[j][0] + shift) * rstd) * gm.x + bt.x;
    h[j][1] = ((h[j][1] + shift) * rstd) * gm.y + bt.y;
  }
}

template <int ENV>
__global__ void __launch_bounds__(PAIRS * 64, 1)
rollout_wm2_fwd_kernel(const __grid_constant__ RolloutArgs<float> A, const __grid_constant__ MlpParams<float> P,
                       const __grid_constant__ SafeConsts<float> k) {
  using TASK = Task<ENV>;
  using T = float;
  constexpr int NS = TASK::NS, NA = TASK::NA, NO = TASK::NO, NAUX = TASK::NAUX;
  static_assert(NO < 16 && NA <= 4, "one k-tile input layer, at most 4 policy outputs");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WmSmem S;
  S.carve(smem_raw, NA);
  PairSmem* PS = reinterpret_cast<PairSmem*>(smem_raw + ((WmSmem::bytes(NA) + 15) & ~size_t(15)));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int pair = warp >> 1, half = warp & 1;
  PairSmem& X = PS[pair];

  for (int idx = tid; idx < (KT + 1) * NT * 32; idx += PAIRS * 64) {
    const int l = idx & 31, nt = (idx >> 5) % NT, kt = (idx >> 5) / NT;          // kt == KT: the input layer
    const int n = nt * 8 + (l >> 2), k0 = 2 * (l & 3);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k0 + (j & 1) + (j >> 1) * 8;
      if (kt < KT) v[j] = P.weight[1][(size_t)n * W + kt * 16 + kk];
      else v[j] = kk < NO ? P.weight[0][(size_t)n * NO + kk] : (kk == NO ? P.bias[0][n] : 0.f);
    }
    uint4 f;
    split_pair(v[0], v[1], f.x, f.z);
    split_pair(v[2], v[3], f.y, f.w);
    (kt < KT ? S.w2f + ((size_t)kt * NT + nt) * 32 : S.w1f + (size_t)nt * 32)[l] = f;
  }
  for (int i = tid; i < W; i += PAIRS * 64) {
    S.b2[i] = P.bias[1][i];
    S.g1[i] = P.ln_weight[0][i]; S.be1[i] = P.ln_bias[0][i];
    S.g2[i] = P.ln_weight[1][i]; S.be2[i] = P.ln_bias[1][i];
  }
  for (int i = tid; i < NA * W; i += PAIRS * 64) S.w_out[i] = P.weight[2][i];
  for (int i = tid; i < NA; i += PAIRS * 64) S.b_out[i] = P.bias[2][i];

  const int64_t N = A.N;
  const int64_t env = ((int64_t)blockIdx.x * PAIRS + pair) * R + g;
  const bool owner = half == 0 && tig == 0;
  const bool row_ok = env < N, valid = owner && row_ok;
  const bool tiled = A.stash_tiled != 0, stash = A.stash_e1 != nullptr;
  const size_t nt_stride = tiled ? 1024 : 8;
  const int col0 = half * (W / 2) + 2 * tig;              // this lane's first column (n-tile 8 half, pair 2 tig)

  T s[NS], aux[NAUX > 0 ? NAUX : 1], ob[NO];
  T sigma[NA];
#pragma unroll
  for (int q = 0; q < NA; ++q) sigma[q] = sg_exp(P.log_std[q]);
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = valid ? A.states[env * NS + i] : T(0);
  if (NAUX > 0) {
#pragma unroll
    for (int i = 0; i < NAUX; ++i) aux[i] = valid ? A.aux0[env * NAUX + i] : T(0);
  }
#pragma unroll
  for (int i = 0; i < NO; ++i) ob[i] = valid ? A.obs[env * NO + i] : T(0);
  if (half == 0) {
    for (int i = lane; i < R * 16; i += 32) X.obs[i / 16][i % 16] = (i % 16) == NO ? 1.f : 0.f;
    __syncwarp();
    if (tig == 0) {
#pragma unroll
      for (int i = 0; i < NO; ++i) X.obs[g][i] = ob[i];
    }
  }
  __syncthreads();

  for (int t = 0; t < A.H; ++t) {
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
    pair_sync(pair);                                   // B0: the observations of this step are in X.obs
    // ---- input layer, this warp's 8 n-tiles -------------------------------------------------------------------------------------
    uint32_t a[4];
    {
      const float2 p01 = *reinterpret_cast<const float2*>(&X.obs[g][2 * tig]);
      const float2 p23 = *reinterpret_cast<const float2*>(&X.obs[g][8 + 2 * tig]);
      split_pair(p01.x, p01.y, a[0], a[1]);
      split_pair(p23.x, p23.y, a[2], a[3]);
    }
    float h[NTH][2];
#pragma unroll
    for (int j = 0; j < NTH; ++j) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint4 b = S.w1f[(half * NTH + j) * 32 + lane];
      hmma(acc, a, b.x, b.y);
      hmma(acc, a, b.z, b.w);
      h[j][0] = acc[0] + acc[2];
      h[j][1] = acc[1] + acc[3];
    }
    const int64_t row = (int64_t)t * N + (row_ok ? env : 0);
    const size_t stash_off = stash_index<W>(tiled, row, col0);
    float st[4], mean_h, m2_h;
    elu_half_stats(h, stash ? A.stash_e1 + stash_off : nullptr, nt_stride, stash && row_ok, mean_h, m2_h);
    if (tig == 0) X.stat[0][half][g] = make_float2(mean_h, m2_h);
    pair_sync(pair);                                   // B1
    layernorm_half(h, mean_h, m2_h, X.stat[0][half ^ 1][g], S.g1, S.be1, col0, st[0], st[1]);
    // ---- A fragments of this warp's 4 k-tiles, exchanged with the partner ---------------------------------------------------------
    uint4 own[KTH], oth[KTH];
#pragma unroll
    for (int j = 0; j < KTH; ++j) {
      split_pair(h[2 * j][0], h[2 * j][1], own[j].x, own[j].y);
      split_pair(h[2 * j + 1][0], h[2 * j + 1][1], own[j].z, own[j].w);
      X.frag[half][j][lane] = own[j];
    }
    pair_sync(pair);                                   // B2
#pragma unroll
    for (int j = 0; j < KTH; ++j) oth[j] = X.frag[half ^ 1][j][lane];
    {
      float acc[NTH][4];
#pragma unroll
      for (int j = 0; j < NTH; ++j) {
        const float2 b = *reinterpret_cast<const float2*>(S.b2 + col0 + j * 8);
        acc[j][0] = b.x; acc[j][1] = b.y; acc[j][2] = 0.f; acc[j][3] = 0.f;
      }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const uint4 f = ((kt / KTH) == 0) == (half == 0) ? own[kt % KTH] : oth[kt % KTH];
        a[0] = f.x; a[1] = f.y; a[2] = f.z; a[3] = f.w;
        const uint4* bf = S.w2f + ((size_t)kt * NT + half * NTH) * 32 + lane;
#pragma unroll
        for (int j = 0; j < NTH; ++j) {
          const uint4 b = bf[j * 32];
          hmma(acc[j], a, b.x, b.y);
          hmma(acc[j], a, b.z, b.w);
        }
      }
#pragma unroll
      for (int j = 0; j < NTH; ++j) {
        h[j][0] = acc[j][0] + acc[j][2];
        h[j][1] = acc[j][1] + acc[j][3];
      }
    }
    elu_half_stats(h, stash ? A.stash_e2 + stash_off : nullptr, nt_stride, stash && row_ok, mean_h, m2_h);
    if (tig == 0) X.stat[1][half][g] = make_float2(mean_h, m2_h);
    pair_sync(pair);                                   // B3
    layernorm_half(h, mean_h, m2_h, X.stat[1][half ^ 1][g], S.g2, S.be2, col0, st[2], st[3]);
    if (stash && valid) *reinterpret_cast<float4*>(A.stash_stats + row * 4) = make_float4(st[0], st[1], st[2], st[3]);
    // ---- output layer: partial sums over this warp's columns --------------------------------------------------------------------
    T mu[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      float o[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < NTH; ++j) {
        const float2 w = *reinterpret_cast<const float2*>(S.w_out + q * W + col0 + j * 8);
        o[j & 1] = fmaf(h[j][0], w.x, o[j & 1]);
        o[j & 1] = fmaf(h[j][1], w.y, o[j & 1]);
      }
      mu[q] = quad_sum(o[0] + o[1]);
      if (tig == 0) X.mu[half][g][q] = mu[q];
    }
    pair_sync(pair);                                   // B4
    // ---- environment step on the owner lanes of the pair's first warp -----------------------------------------------------------------
    if (owner) {
      T act[NA];
#pragma unroll
      for (int q = 0; q < NA; ++q) act[q] = valid ? sg_tanh((mu[q] + X.mu[1][g][q] + S.b_out[q]) + sigma[q] * epsv[q]) : T(0);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < NS; ++i) wv[i] = T(0);
      }
      if (NAUX > 0 && ridx >= 0 && valid) {
#pragma unroll
        for (int i = 0; i < NAUX; ++i) aux[i] = rs[i];
      }
      T sn[NS], ae[NA], rew;
      int fl;
      const StepShared<T>* sh = A.shared ? A.shared + t : nullptr;
      StepShared<T> empty;
      if (!sh) { memset(&empty, 0, sizeof(empty)); sh = &empty; }
      safeguarded_step<TASK, T, T>(s, act, wv, aux, A.dirs ? dv : nullptr, sh, k, ridx >= 0, rs, sn, ae, ob, rew, fl);
#pragma unroll
      for (int i = 0; i < NS; ++i) s[i] = sn[i];
#pragma unroll
      for (int i = 0; i < NO; ++i) X.obs[g][i] = ob[i];
      if (valid) {
#pragma unroll
        for (int i = 0; i < NS; ++i) A.states[((int64_t)(t + 1) * N + env) * NS + i] = sn[i];
#pragma unroll
        for (int i = 0; i < NO; ++i) A.obs[((int64_t)(t + 1) * N + env) * NO + i] = ob[i];
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          A.actions[((int64_t)t * N + env) * NA + q] = act[q];
          A.exec_actions[((int64_t)t * N + env) * NA + q] = ae[q];
        }
        A.rewards[(int64_t)t * N + env] = rew;
        A.flags[(int64_t)t * N + env] = fl;
      }
    }
  }
}

template <int ENV>
int rollout_wm_fwd(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st) {
  using TASK = Task<ENV>;
  if (pol->in_dim != TASK::NO || pol->out_dim != TASK::NA) return fail2(SGBPO_E_ARG, "policy dims do not match the environment");
  if (pol->width != W || pol->n_hidden != 2) return fail2(SGBPO_E_UNSUPPORTED, "warp-MMA forward: policies [in -> 128 -> 128 -> m]");
  const int n_stash = (r->stash_e1 != nullptr) + (r->stash_e2 != nullptr) + (r->stash_stats != nullptr);
  if (n_stash != 0 && n_stash != 3) return fail2(SGBPO_E_ARG, "warp-MMA forward: pass e1, e2 and stats stash buffers or none");
  if (r->stash_h1) return fail2(SGBPO_E_UNSUPPORTED, "warp-MMA forward does not write the h1 stash (sequential reverse pass)");
  // two warps per 8 environments while one CTA per SM holds the whole batch (the pair kernel keeps 162 registers per thread: one
  // CTA of 4 pairs per SM), one warp per 8 environments beyond that; SGBPO_WM_SPLIT=1|2 forces
  const char* split = getenv("SGBPO_WM_SPLIT");
  const bool pairs_fit = (r->N + PAIRS * R - 1) / (PAIRS * R) <= sm_count();
  if (split ? split[0] == '2' : pairs_fit) {
    auto kern2 = rollout_wm2_fwd_kernel<ENV>;
    const size_t smem2 = ((WmSmem::bytes(TASK::NA) + 15) & ~size_t(15)) + PAIRS * sizeof(PairSmem);
    cudaError_t e2 = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e2 != cudaSuccess) return cuda_fail2(e2, "cudaFuncSetAttribute(rollout_wm2_fwd)");
    const int64_t grid2 = (r->N + PAIRS * R - 1) / (PAIRS * R);
    kern2<<<(unsigned)grid2, PAIRS * 64, smem2, st>>>(convert_rollout<float>(*r), convert_mlp<float>(*pol), convert_consts<float>(*c));
    e2 = cudaGetLastError();
    return e2 == cudaSuccess ? SGBPO_OK : cuda_fail2(e2, "rollout_wm2_fwd_kernel");
  }
  auto kern = rollout_wm_fwd_kernel<ENV>;
  const size_t smem = WmSmem::bytes(TASK::NA);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail2(e, "cudaFuncSetAttribute(rollout_wm_fwd)");
  const int64_t grid = (r->N + WARPS * R - 1) / (WARPS * R);
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(convert_rollout<float>(*r), convert_mlp<float>(*pol), convert_consts<float>(*c));
  e = cudaGetLastError();
  return e == cudaSuccess ? SGBPO_OK : cuda_fail2(e, "rollout_wm_fwd_kernel");
}

template <int ENV> int rollout_wm_fwd_env(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st);

#define SGBPO_INSTANTIATE_ROLLOUT_WM(E) \
  template <> int rollout_wm_fwd_env<E>(const sgbpo_consts* c, const sgbpo_mlp* pol, const sgbpo_rollout* r, cudaStream_t st) { \
    return rollout_wm_fwd<E>(c, pol, r, st);                                                                               \
  }

}  // namespace wm
}  // namespace sgbpo
