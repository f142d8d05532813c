// tcgen05 / TMEM / mbarrier building blocks shared by the tensor-core kernels (sgbpo_mlp_tc.cu, sgbpo_rollout_tc.cuh).
//
// Operand tiles are 128 x 128 fp16 "parts" (32 KB) in the no-swizzle core-matrix layout
//     byte(r, c) = (c / 8) * 2048 + (r / 8) * 128 + (r % 8) * 16 + (c % 8) * 2          (r = row, c = column)
// i.e. 16-byte units of 8 consecutive columns; unit (c / 8, r) sits at (c / 8) * 128 + r.  One buffer serves two roles:
//   * K-major operand  X[mn = r][k = c]   (descriptor: K-group stride LBO = 2048, MN-group stride SBO = 128), and
//   * MN-major operand X^T[mn = c][k = r] (descriptor: K-group stride LBO = 128,  MN-group stride SBO = 2048),
// so activations stored once by their owner threads (thread = row) feed both  Y = X W^T  (contraction over columns) and
// G = X^T Z (contraction over rows, e.g. weight gradients) without a transposed copy.
//
// fp32 accuracy from fp16 passes: x = hi + lo, hi = fp16(x), lo = fp16(x - hi); a product uses the three MMAs
// lo*hi + hi*lo + hi*hi with fp32 accumulation in TMEM (relative error ~2^-22 while |x| >= 2^-3, absolute floor 2^-25).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace sgbpo {
namespace tc {

constexpr int ROWS = 128;            // rows per tile = TMEM lanes
constexpr int W = 128;               // hidden width
constexpr int KCH = W / 8;           // 16-byte chunks per row
constexpr int PART_BYTES = ROWS * W * 2;          // one fp16 part of a 128 x 128 operand: 32 KB
constexpr uint32_t SBO = 128, LBO = ROWS * 16;    // bytes: next 8-row group / next 8-column chunk

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor, version 1): both strides in 16-byte units
__device__ __forceinline__ uint64_t make_desc_raw(uint32_t addr, uint32_t k_group_stride, uint32_t mn_group_stride) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((k_group_stride >> 4) & 0x3FFF) << 16;      // "leading byte offset"
  d |= (uint64_t)((mn_group_stride >> 4) & 0x3FFF) << 32;     // "stride byte offset"
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) { return make_desc_raw(addr, LBO, SBO); }        // K-major view
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr) { return make_desc_raw(addr, SBO, LBO); }     // MN-major view
// byte advance of the start address per K = 16 MMA step
constexpr uint32_t KSTEP_K = 2 * LBO;     // K-major: two 8-column chunks
constexpr uint32_t KSTEP_MN = 2 * SBO;    // MN-major: two 8-row groups

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, M = 128; bit 15 / 16: A / B is MN-major
__host__ __device__ constexpr uint32_t idesc_f16(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
}
constexpr uint32_t IDESC = idesc_f16(W, false, false);

__device__ __forceinline__ void mma_f16_i(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  mma_f16_i(tmem_d, da, db, IDESC, accumulate);
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP): global -> shared, completion counted in bytes on an mbarrier -------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {     // 16-byte multiples
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// L2 prefetch of a contiguous block (TMA unit, no destination): turns the DRAM latency of the next tile's loads into L2 latency
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {      // 16-byte multiples
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// named barrier over `count` threads (ids 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <uint32_t COLS> __device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
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
// same with the hardware exponential (ex2.approx, 2 ulp): absolute error <= 1e-6 on the ELU output, a tenth of the 1e-5
// per-step bar, for a third of the instructions of the time loop's epilogues
__device__ __forceinline__ float elu1_fast(float v) {            // one multiply + MUFU.EX2 + add + select
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
  return v > 0.f ? v : e - 1.0f;
}

// 8 consecutive columns [8 kc, 8 kc + 8) of operand row `row` -> fp16 hi / lo parts, one 16-byte unit each
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

// D (+)= A B^T-style product of two split operands over K = 128: lo*hi + hi*lo + hi*hi, 24 instructions, issued by ONE thread.
// a_step / b_step: byte advance per K = 16 step of each operand (KSTEP_K or KSTEP_MN); mn views use make_desc_mn.
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void mma_split3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t acc) {
  constexpr uint32_t idesc = idesc_f16(W, A_MN, B_MN);
  constexpr uint32_t a_step = A_MN ? KSTEP_MN : KSTEP_K, b_step = B_MN ? KSTEP_MN : KSTEP_K;
#pragma unroll
  for (int term = 0; term < 3; ++term) {           // small cross terms first, the dominant hi x hi product last
    const uint32_t aa = term == 0 ? a_lo : a_hi;
    const uint32_t bb = term == 1 ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < W / 16; ++ks) {
      const uint64_t da = A_MN ? make_desc_mn(aa + ks * a_step) : make_desc(aa + ks * a_step);
      const uint64_t db = B_MN ? make_desc_mn(bb + ks * b_step) : make_desc(bb + ks * b_step);
      mma_f16_i(tmem_d, da, db, idesc, acc);
      acc = 1;
    }
  }
}

}  // namespace tc
}  // namespace sgbpo
