// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands in the K-major no-swizzle layout
// the conv kernels use) as a function of N, issued back to back by one thread.  nvcc -arch=sm_100a.
//   mode 0: descriptors recomputed per MMA (address arithmetic in the issue loop)
//   mode 1: 8 precomputed descriptor pairs in registers, loop unrolled by 8 (pure issue rate)
//   mode 2: like 1, descriptors streamed from shared memory (ld.shared.b64 x2 per MMA)
//   mode 3: tiles of 12 MMAs (6 x [N, N/2]) each followed by tcgen05.commit + mbarrier wait (hand-off latency)
//   mode 4: mode 3 while the other 3 warps stream 128-bit shared-memory stores/loads (bandwidth contention)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../svcc23_fastsvc_b200/csrc/conv_tc3.cuh"
using namespace fsvc;
__global__ void __launch_bounds__(128, 1) mma_rate(int N, int n_mma, int mode, long long* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t s_desc[16];
  __shared__ int stop_flag;
  bool return_early = false;
  if (threadIdx.x == 0) stop_flag = 0;
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
    if (mode >= 3) {
    } else if (mode == 0) {
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
    if (mode >= 3) {
      uint64_t A[6], Bd[6];
      for (int j = 0; j < 6; ++j) { A[j] = s_desc[2 * j]; Bd[j] = s_desc[2 * j + 1]; }
      const uint32_t idesc_h = umma_idesc_bf16(128, N / 2);
      uint32_t ph = 0;
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 12) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          umma_bf16(tmem, A[j], Bd[j], idesc, j > 0);
          umma_bf16(tmem, A[j], Bd[j], idesc_h, 1u);
        }
        umma_commit(&bar);
        mbar_wait2(&bar, ph);
        ph ^= 1u;
      }
      out[0] = clock64() - t0;
      stop_flag = 1;
      return_early = true;
    }
    if (!return_early) {
      umma_commit(&bar);
      mbar_wait2(&bar, 0);
      long long t1 = clock64();
      out[0] = t1 - t0;
    }
  } else if (mode == 4 && warp > 0) {
    // contention: stream 128-bit shared-memory traffic until the MMA thread is done
    uint32_t addr = smem_u32(sm) + 150 * 1024 + (uint32_t)tid * 16u;
    uint4 v = make_uint4(tid, 1, 2, 3);
    while (*(volatile int*)&stop_flag == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr + j * 2048u), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr + j * 2048u + 1024u) : "memory");
      }
    }
    if (v.x == 0xdeadbeef) out[1] = v.y;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n = 4092;
  for (int mode = 1; mode < 5; ++mode)
    for (int N : {32, 64, 96, 128}) {
      mma_rate<<<1, 128, 200 * 1024>>>(N, n, mode, d);
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("mode=%d N=%3d: %.1f cycles/MMA  %.0f cycles per 12-MMA tile (%s)\n", mode, N, (double)h / n, 12.0 * h / n, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
