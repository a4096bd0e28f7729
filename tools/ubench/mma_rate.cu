// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands in the K-major no-swizzle layout
// the conv kernels use) as a function of N, issued back to back by one thread.  nvcc -arch=sm_100a.
//   mode 0: descriptors recomputed per MMA (address arithmetic in the issue loop)
//   mode 1: 8 precomputed descriptor pairs in registers, loop unrolled by 8 (pure issue rate)
//   mode 2: like 1, descriptors streamed from shared memory (ld.shared.b64 x2 per MMA)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../svcc23_fastsvc_b200/csrc/conv_tc3.cuh"
using namespace fsvc;
__global__ void __launch_bounds__(128, 1) mma_rate(int N, int n_mma, int mode, long long* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t s_desc[16];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  const uint32_t strip = 272 * 16, a_base = smem_u32(sm), b_base = a_base + 100 * 1024;
  const uint32_t b_strip = (uint32_t)N * 16;
  if (tid < 8) {
    s_desc[2 * tid] = umma_desc(a_base + (uint32_t)tid * 2 * strip + (uint32_t)(tid % 3) * 48, strip, 128);
    s_desc[2 * tid + 1] = umma_desc(b_base + (uint32_t)(tid % 4) * 2 * b_strip, b_strip, 128);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t ao = (uint32_t)(i % 8) * 2 * strip + (uint32_t)(i % 3) * 48;
        const uint64_t A = umma_desc(a_base + ao, strip, 128), Bd = umma_desc(b_base + (uint32_t)(i % 4) * 2 * b_strip, b_strip, 128);
        umma_bf16(tmem, A, Bd, idesc, i > 0);
      }
    } else if (mode == 1) {
      uint64_t A[8], Bd[8];
      for (int j = 0; j < 8; ++j) { A[j] = s_desc[2 * j]; Bd[j] = s_desc[2 * j + 1]; }
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_bf16(tmem, A[j], Bd[j], idesc, 1u);
      }
    } else {
      const uint32_t dp = smem_u32(s_desc);
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint64_t A, Bd;
          asm volatile("ld.shared.b64 %0, [%1];" : "=l"(A) : "r"(dp + 16 * j));
          asm volatile("ld.shared.b64 %0, [%1];" : "=l"(Bd) : "r"(dp + 16 * j + 8));
          umma_bf16(tmem, A, Bd, idesc, 1u);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait2(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n = 4096;
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {16, 32, 64, 96, 128, 192, 256}) {
      mma_rate<<<1, 128, 200 * 1024>>>(N, n, mode, d);
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("mode=%d N=%3d: %.1f cycles/MMA (%s)\n", mode, N, (double)h / n, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
