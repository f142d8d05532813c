// Microbenchmark: issue rate and latency of mma.sync.m16n8k16 (f16 in, f32 accumulate) on sm_100a, one warp per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o hmma_rate hmma_rate.cu ; run: ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void bench(float* out, long long* cycles, int iters) {
  float acc[NACC][4];
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x38003800u};
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) hmma(acc[i], a, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NACC> void run(int warps, float* out, long long* cyc) {
  const int iters = 2000;
  bench<NACC><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("warps/CTA %2d  independent accumulators %2d : %.2f cycles per HMMA per warp, %.2f HMMA/cycle/SM\n", warps, NACC,
         (double)c / (iters * NACC), (double)warps * iters * NACC / c);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16}) { run<1>(w, out, cyc); run<4>(w, out, cyc); run<16>(w, out, cyc); }
  return 0;
}
