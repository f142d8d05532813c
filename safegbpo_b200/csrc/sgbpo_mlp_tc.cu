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

namespace sgbpo {
extern thread_local char g_err[512];

namespace tc {

constexpr int ROWS = 128;            // rows per tile = threads per CTA = TMEM lanes
constexpr int W = 128;               // hidden width
constexpr int KCH = W / 8;           // 16-byte K chunks per row
constexpr int PART_BYTES = ROWS * W * 2;          // one fp16 part of a 128 x 128 operand: 32 KB
constexpr uint32_t SBO = 128, LBO = ROWS * 16;    // bytes: next 8-row group / next K chunk
constexpr int MAX_HH = 2;            // hidden-hidden layers kept in shared memory
constexpr int MAX_IN = 64;
constexpr uint32_t TMEM_COLS = 128;

struct Params {
  int in_dim, n_hidden;
  const float* w_in;  const float* w_hh[MAX_HH];  const float* w_out;      // torch layouts [out][in]
  const float* b[MAX_HH + 1];  const float* gamma[MAX_HH + 1];  const float* beta[MAX_HH + 1];
  const float* b_out;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(W >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();      // never hang the GPU: a lost arrival becomes a launch failure
  }
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float elu1(float v) {                 // ATen: exp(x) - 1 on the negative side
  const float en = expf(v < 0.f ? v : 0.f) - 1.0f;
  return v > 0.f ? v : en;
}
// 8 consecutive K values of one operand row -> fp16 hi / lo parts, one 16-byte unit each
__device__ __forceinline__ void split_store(unsigned char* a_hi, unsigned char* a_lo, int kc, int row, const float (&h)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h2 = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
    const float2 back = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(h[2 * i] - back.x, h[2 * i + 1] - back.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  const int unit = (kc * ROWS + row) * 16;
  *reinterpret_cast<uint4*>(a_hi + unit) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(a_lo + unit) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

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
    for (int c = 0; c < CW; ++c) { e[c] = elu1(e[c]); sum += e[c]; }
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
          const float v0 = elu1(z[4 * i4 + 0] + b4.x), v1 = elu1(z[4 * i4 + 1] + b4.y);
          const float v2 = elu1(z[4 * i4 + 2] + b4.z), v3 = elu1(z[4 * i4 + 3] + b4.w);
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
