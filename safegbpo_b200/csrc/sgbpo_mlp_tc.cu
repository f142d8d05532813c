// Batched MLP forward on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
// Target critic over all (H x N) observation rows of the rollout buffer (reference:
// src/learning_algorithms/shac.py:111 `value = target_value_function(obs)`, value_function.py:38-66:
// Linear -> ELU -> LayerNorm per hidden layer, then Linear).  This is the one GEMM-shaped, non-sequential
// piece of the path (262 144 rows x [o -> 128 -> 128 -> 128 -> 1] for BASELINE configs[1]).
//
// Mapping.  One CTA = one tile of 128 rows at a time; 128 x NSPLIT threads: thread t owns row t % 128 (= its TMEM
// lane) and 128 / NSPLIT of its columns (NSPLIT = 4: four warps per SM sub-partition for the issue-bound epilogue):
//   * accumulator D[128 rows x 128 hidden] lives in TMEM (128 columns of fp32); `tcgen05.ld 32x32b` hands every
//     thread ITS row segment, so bias + ELU + LayerNorm are thread-local up to one shared-memory exchange of the
//     NSPLIT partial row sums (no shuffles);
//   * A = the tile's activations, B = Linear.weight [out][in] (K-major as torch stores it), both in shared
//     memory in the no-swizzle K-major core-matrix layout [K/8][128 rows][8 halves] (16-byte units; descriptor
//     SBO = 128 B between 8-row groups, LBO = 2048 B between K chunks).  The epilogue of layer l writes the A
//     operand of layer l+1 straight from registers: consecutive threads store consecutive 16-byte units;
//   * fp32 accuracy from fp16 tensor-core passes: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 mantissa
//     bits for O(1) values, absolute floor 2^-25), D = A_lo B_hi + A_hi B_lo + A_hi B_hi accumulated in fp32
//     in TMEM: 3 x 8 `tcgen05.mma.kind::f16` (M128 N128 K16) per layer per tile, issued by one thread and
//     tracked with `tcgen05.commit` on an mbarrier.  Activations after LayerNorm are O(1) and weights are
//     bounded, far inside fp16 range; a value beyond 65504 would surface as a non-finite output.
//   * the first layer (K = observation size, 3..23) and the 128 -> 1 output layer stay on the FMA pipe.
//
// Shared memory: weights of up to two hidden-hidden layers (2 x 2 x 32 KB) + one A tile (2 x 32 KB) + the small
// parameter vectors: ~200 KB, one CTA per SM, persistent over tiles.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../include/sgbpo.h"
#include "sgbpo_tc.cuh"

namespace sgbpo {
extern thread_local char g_err[512];

namespace tc {

constexpr int MAX_HH = 2;            // hidden-hidden layers kept in shared memory
constexpr int MAX_IN = 64;
constexpr uint32_t TMEM_COLS = 128;

struct Params {
  int in_dim, n_hidden;
  const float* w_in;  const float* w_hh[MAX_HH];  const float* w_out;      // torch layouts [out][in]
  const float* b[MAX_HH + 1];  const float* gamma[MAX_HH + 1];  const float* beta[MAX_HH + 1];
  const float* b_out;
};

// sum over the NSPLIT threads that share a row (they sit in different warps): shared-memory exchange
template <int NSPLIT>
__device__ __forceinline__ float row_sum(float v, float* red, int part, int row) {
  if constexpr (NSPLIT == 1) return v;
  red[part * ROWS + row] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int p = 0; p < NSPLIT; ++p) t += red[p * ROWS + row];
  return t;
}

// LayerNorm (eps 1e-5, biased variance, two-pass) of the row whose CW columns [c0, c0 + CW) this thread holds
template <int NSPLIT>
__device__ __forceinline__ void normalise(float (&e)[W / NSPLIT], float sum, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float* red_a, float* red_b, int part, int row) {
  constexpr int CW = W / NSPLIT;
  const float mean = row_sum<NSPLIT>(sum, red_a, part, row) / float(W);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CW; ++c) { e[c] -= mean; q += e[c] * e[c]; }
  const float rs = rsqrtf(row_sum<NSPLIT>(q, red_b, part, row) / float(W) + 1e-5f);
#pragma unroll
  for (int c4 = 0; c4 < CW / 4; ++c4) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + part * CW + 4 * c4);
    const float4 b = *reinterpret_cast<const float4*>(beta + part * CW + 4 * c4);
    e[4 * c4 + 0] = (e[4 * c4 + 0] * rs) * g.x + b.x;
    e[4 * c4 + 1] = (e[4 * c4 + 1] * rs) * g.y + b.y;
    e[4 * c4 + 2] = (e[4 * c4 + 2] * rs) * g.z + b.z;
    e[4 * c4 + 3] = (e[4 * c4 + 3] * rs) * g.w + b.w;
  }
}

// NSPLIT threads per row: thread tid owns row tid % 128 (= its TMEM lane; warps w, w+4, ... share a lane quadrant)
// and the CW = 128 / NSPLIT columns [part * CW, (part + 1) * CW), part = tid / 128.  More threads per row = more
// warps per SM sub-partition to interleave in the (issue-bound) epilogue.
template <int NSPLIT>
__global__ void __launch_bounds__(ROWS * NSPLIT, 1)
mlp_rows_tc_kernel(int64_t M, const float* __restrict__ x, float* __restrict__ out, const __grid_constant__ Params P) {
  constexpr int CW = W / NSPLIT, NT = ROWS * NSPLIT;
  static_assert(CW % 32 == 0, "tcgen05.ld moves 32 columns at a time");
  extern __shared__ __align__(1024) unsigned char smem[];
  const int n_hh = P.n_hidden - 1;
  unsigned char* b_parts = smem;                                     // [n_hh][hi, lo][PART_BYTES]
  unsigned char* a_hi = smem + (size_t)n_hh * 2 * PART_BYTES;
  unsigned char* a_lo = a_hi + PART_BYTES;
  float* w_in = reinterpret_cast<float*>(a_lo + PART_BYTES);         // [in_dim][W]
  float* vec = w_in + P.in_dim * W;                                  // per hidden layer: bias, gamma, beta; then w_out, b_out
  float* w_out = vec + 3 * P.n_hidden * W;
  float* red_a = w_out + W + 4;                                      // [NSPLIT][ROWS] exchange buffers
  float* red_b = red_a + NSPLIT * ROWS;
  float* red_c = red_b + NSPLIT * ROWS;                              // (three buffers: consecutive exchanges never reuse
  uint64_t* bar = reinterpret_cast<uint64_t*>(red_c + NSPLIT * ROWS);  //  a buffer without a barrier in between)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int trow = tid % ROWS, part = tid / ROWS;

  // ---- one-time setup: parameters -> shared memory (weights split into fp16 hi/lo B operands) -------------
  for (int i = tid; i < P.in_dim * W; i += NT) w_in[(i % P.in_dim) * W + (i / P.in_dim)] = P.w_in[i];
  for (int l = 0; l < P.n_hidden; ++l)
    for (int i = tid; i < W; i += NT) {
      vec[(3 * l + 0) * W + i] = P.b[l][i];
      vec[(3 * l + 1) * W + i] = P.gamma[l][i];
      vec[(3 * l + 2) * W + i] = P.beta[l][i];
    }
  for (int i = tid; i < W; i += NT) w_out[i] = P.w_out[i];
  if (tid == 0) w_out[W] = P.b_out[0];
  for (int l = 0; l < n_hh; ++l) {
    unsigned char* hi = b_parts + (size_t)l * 2 * PART_BYTES;
    unsigned char* lo = hi + PART_BYTES;
    for (int i = tid; i < W * KCH; i += NT) {                        // (n, kc): 8 consecutive k of output row n
      const int n = i / KCH, kc = i % KCH;
      const float4 v0 = *reinterpret_cast<const float4*>(P.w_hh[l] + (size_t)n * W + kc * 8);
      const float4 v1 = *reinterpret_cast<const float4*>(P.w_hh[l] + (size_t)n * W + kc * 8 + 4);
      const float h[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      split_store(hi, lo, kc, n, h);
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // weight operands visible to the tensor-core proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // this warp's 32 lanes (quadrant warp % 4) and this thread's first column
  const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(part * CW);
  const uint32_t bar_addr = smem_u32(bar);
  const uint32_t a_hi_addr = smem_u32(a_hi), a_lo_addr = smem_u32(a_lo);
  uint32_t phase = 0;

  const int64_t n_tiles = (M + ROWS - 1) / ROWS;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * ROWS + trow;
    const bool live = row < M;
    float e[CW];
    float sum = 0.f;
    // ---- layer 1 on the FMA pipe: K = in_dim (runtime), CW accumulators in registers ---------------------------
#pragma unroll
    for (int c4 = 0; c4 < CW / 4; ++c4) {
      const float4 b4 = *reinterpret_cast<const float4*>(vec + part * CW + 4 * c4);
      e[4 * c4 + 0] = b4.x; e[4 * c4 + 1] = b4.y; e[4 * c4 + 2] = b4.z; e[4 * c4 + 3] = b4.w;
    }
    for (int k = 0; k < P.in_dim; ++k) {
      const float xk = live ? x[row * P.in_dim + k] : 0.f;
      const float* wk = w_in + k * W + part * CW;
#pragma unroll
      for (int c4 = 0; c4 < CW / 4; ++c4) {
        const float4 w = *reinterpret_cast<const float4*>(wk + 4 * c4);
        e[4 * c4 + 0] = fmaf(xk, w.x, e[4 * c4 + 0]); e[4 * c4 + 1] = fmaf(xk, w.y, e[4 * c4 + 1]);
        e[4 * c4 + 2] = fmaf(xk, w.z, e[4 * c4 + 2]); e[4 * c4 + 3] = fmaf(xk, w.w, e[4 * c4 + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < CW; ++c) { e[c] = elu1_fast(e[c]); sum += e[c]; }
    normalise<NSPLIT>(e, sum, vec + 1 * W, vec + 2 * W, red_a, red_b, part, trow);
    // ---- hidden-hidden layers on the tensor cores ---------------------------------------------------------------
    for (int l = 0; l < n_hh; ++l) {
#pragma unroll
      for (int kc = 0; kc < CW / 8; ++kc) {
        const float h8[8] = {e[8 * kc], e[8 * kc + 1], e[8 * kc + 2], e[8 * kc + 3], e[8 * kc + 4], e[8 * kc + 5], e[8 * kc + 6], e[8 * kc + 7]};
        split_store(a_hi, a_lo, part * (CW / 8) + kc, trow, h8);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy stores -> async proxy reads
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");    // orders the previous tcgen05.ld before the next MMA
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b_hi_addr = smem_u32(b_parts + (size_t)l * 2 * PART_BYTES), b_lo_addr = b_hi_addr + PART_BYTES;
        uint32_t acc = 0;
        // small cross terms first, the dominant hi x hi product last
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t aa = term == 0 ? a_lo_addr : a_hi_addr;
          const uint32_t bb = term == 1 ? b_lo_addr : b_hi_addr;
#pragma unroll
          for (int ks = 0; ks < W / 16; ++ks) {                              // K = 16 per instruction = 2 chunks
            mma_f16(tmem_base, make_desc(aa + ks * 2 * LBO), make_desc(bb + ks * 2 * LBO), acc);
            acc = 1;
          }
        }
        mma_commit(bar_addr);
      }
      mbar_wait(bar_addr, phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float* bias = vec + (3 * (l + 1) + 0) * W + part * CW;
      sum = 0.f;
#pragma unroll
      for (int q = 0; q < CW / 32; ++q) {
        float z[32];
        tmem_ld32(tmem_row + q * 32, z);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + q * 32 + 4 * i4);
          const float v0 = elu1_fast(z[4 * i4 + 0] + b4.x), v1 = elu1_fast(z[4 * i4 + 1] + b4.y);
          const float v2 = elu1_fast(z[4 * i4 + 2] + b4.z), v3 = elu1_fast(z[4 * i4 + 3] + b4.w);
          e[q * 32 + 4 * i4 + 0] = v0; e[q * 32 + 4 * i4 + 1] = v1; e[q * 32 + 4 * i4 + 2] = v2; e[q * 32 + 4 * i4 + 3] = v3;
          sum += (v0 + v1) + (v2 + v3);
        }
      }
      normalise<NSPLIT>(e, sum, vec + (3 * (l + 1) + 1) * W, vec + (3 * (l + 1) + 2) * W, red_a, red_b, part, trow);
    }
    // ---- output head (128 -> 1) -------------------------------------------------------------------------------
    float o = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < CW / 4; ++c4) {
      const float4 w = *reinterpret_cast<const float4*>(w_out + part * CW + 4 * c4);
      o = fmaf(e[4 * c4 + 0], w.x, o); o = fmaf(e[4 * c4 + 1], w.y, o);
      o = fmaf(e[4 * c4 + 2], w.z, o); o = fmaf(e[4 * c4 + 3], w.w, o);
    }
    o = row_sum<NSPLIT>(o, red_c, part, trow) + w_out[W];
    if (live && part == 0) out[row] = o;
  }
  // ---- teardown ---------------------------------------------------------------------------------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <int NSPLIT>
static int launch_tc(int64_t M, const sgbpo_mlp* net, const Params& p, const void* x, void* out, cudaStream_t st) {
  const size_t smem = (size_t)(net->n_hidden - 1) * 2 * PART_BYTES + 2 * PART_BYTES +
                      sizeof(float) * ((size_t)net->in_dim * W + 3 * (size_t)net->n_hidden * W + W + 4 + 3 * NSPLIT * ROWS) + 16 + 1024;
  if (smem > 227 * 1024) return SGBPO_E_UNSUPPORTED;
  auto kern = mlp_rows_tc_kernel<NSPLIT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(mlp_rows_tc): %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (M + ROWS - 1) / ROWS;
  // persistent CTAs, tiles dealt round-robin; an even split (e.g. 2048 tiles -> 128 CTAs x 16) beats 148 CTAs with a ragged tail
  int64_t waves = (tiles + sms - 1) / sms;
  const int grid = (int)((tiles + waves - 1) / waves);
  kern<<<grid, ROWS * NSPLIT, smem, st>>>(M, (const float*)x, (float*)out, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "mlp_rows_tc_kernel: %s", cudaGetErrorString(e));
    return SGBPO_E_CUDA;
  }
  return SGBPO_OK;
}


// ---- operand-view self-test: D = X Z^T (mode 0), X Z (mode 1: B as MN-major view), X^T Z (mode 2: both MN-major) ----
// One CTA, thread r owns row r of X and Z (128 x 128 fp32 each); tests/test_gpu_tc_mma.py compares with float64 torch.
__global__ void __launch_bounds__(ROWS, 1)
tc_selftest_kernel(int mode, const float* __restrict__ X, const float* __restrict__ Z, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* x_hi = smem;
  unsigned char* x_lo = x_hi + PART_BYTES;
  unsigned char* z_hi = x_lo + PART_BYTES;
  unsigned char* z_lo = z_hi + PART_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(z_lo + PART_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int kc = 0; kc < KCH; ++kc) {
    float h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = X[tid * W + kc * 8 + i];
    split_store(x_hi, x_lo, kc, tid, h);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = Z[tid * W + kc * 8 + i];
    split_store(z_hi, z_lo, kc, tid, h);
  }
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t xh = smem_u32(x_hi), xl = smem_u32(x_lo), zh = smem_u32(z_hi), zl = smem_u32(z_lo);
    if (mode == 0) mma_split3<false, false>(tmem_base, xh, xl, zh, zl, 0);
    else if (mode == 1) mma_split3<false, true>(tmem_base, xh, xl, zh, zl, 0);
    else mma_split3<true, true>(tmem_base, xh, xl, zh, zl, 0);
    mma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
#pragma unroll
  for (int q = 0; q < W / 32; ++q) {
    float z[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + q * 32, z);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[tid * W + q * 32 + i] = z[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace tc

// returns SGBPO_E_UNSUPPORTED (no error text) when the network is outside what the tensor-core kernel covers
int launch_rows_tc(int64_t M, const sgbpo_mlp* net, const void* x, void* out, cudaStream_t st) {
  using namespace tc;
  if (net->width != W || net->out_dim != 1 || net->n_hidden < 2 || net->n_hidden > MAX_HH + 1 || net->in_dim > MAX_IN ||
      net->in_dim < 1)
    return SGBPO_E_UNSUPPORTED;
  Params p;
  p.in_dim = net->in_dim; p.n_hidden = net->n_hidden;
  p.w_in = (const float*)net->weight[0];
  for (int l = 0; l < MAX_HH; ++l) p.w_hh[l] = l + 1 < net->n_hidden ? (const float*)net->weight[l + 1] : nullptr;
  p.w_out = (const float*)net->weight[net->n_hidden];
  p.b_out = (const float*)net->bias[net->n_hidden];
  for (int l = 0; l <= MAX_HH; ++l) {
    p.b[l] = l < net->n_hidden ? (const float*)net->bias[l] : nullptr;
    p.gamma[l] = l < net->n_hidden ? (const float*)net->ln_weight[l] : nullptr;
    p.beta[l] = l < net->n_hidden ? (const float*)net->ln_bias[l] : nullptr;
  }
  const char* v = getenv("SGBPO_ROWS_TC_SPLIT");
  const int split = v ? atoi(v) : 4;
  if (split == 1) return launch_tc<1>(M, net, p, x, out, st);
  if (split == 2) return launch_tc<2>(M, net, p, x, out, st);
  return launch_tc<4>(M, net, p, x, out, st);
}

}  // namespace sgbpo

extern "C" int sgbpo_tc_selftest(int mode, const void* x, const void* z, void* d, void* stream) {
  using namespace sgbpo;
  using namespace sgbpo::tc;
  if (!x || !z || !d || mode < 0 || mode > 2) { snprintf(g_err, sizeof(g_err), "sgbpo_tc_selftest: bad argument"); return SGBPO_E_ARG; }
  const size_t smem = 4 * (size_t)PART_BYTES + 64;
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    tc_selftest_kernel<<<1, ROWS, smem, (cudaStream_t)stream>>>(mode, (const float*)x, (const float*)z, (float*)d);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "tc_selftest_kernel: %s", cudaGetErrorString(e)); return SGBPO_E_CUDA; }
  return SGBPO_OK;
}
