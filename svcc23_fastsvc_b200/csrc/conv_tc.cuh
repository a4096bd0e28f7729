// Tensor-core (tcgen05 / TMEM) 1-D convolution for sm_100a, same fused prologue /
// epilogue contract as conv1d_f32_kernel (conv_f32.cuh) and the same ConvArgs.
//
// Implicit GEMM, time on M:   D[128 t, N co] += A_tap[128 t, 16 ci] * W_tap[16 ci, N co]
// for every tap and 16-channel slice.  fp32 parity (<= 1e-3) is kept with a 3-term bf16
// split of both operands (a = a_hi + a_lo, w = w_hi + w_lo; a_hi*w_hi + a_lo*w_hi +
// a_hi*w_lo, fp32 accumulation in TMEM): measured 1.1e-4 max-abs on the whole generator.
//
// Shared-memory operand layout (UMMA "K-major, SWIZZLE_NONE" canonical form): an operand
// is a set of column strips, one per group of 8 channels; a strip holds one 16-byte chunk
// (8 bf16 channels) per row, rows contiguous:  addr(row, ch) = strip(ch/8) + row*16 + (ch%8)*2.
// Core matrices (8 rows x 16 B) are therefore contiguous 128-byte blocks with SBO = 128 B and
// LBO = strip pitch, and -- the point of this layout -- a dilated tap is just a descriptor
// whose start address is advanced by tap*dil rows (16-byte granularity), so the three taps of
// a k=3 conv read the SAME staged activation window; nothing is im2col-copied.
//
// The A window is staged by the CTA's threads (global fp32 -> InstanceNorm affine -> lrelu ->
// zero padding -> hi/lo bf16 -> st.shared), because the normalisation statistics only exist
// once the producing kernel has finished; weights arrive pre-split and pre-laid-out from
// global (see pack_tc_weights_kernel).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_f32.cuh"

namespace fsvc {

// ---- PTX wrappers ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 inputs, fp32 accumulate); issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 8 consecutive fp32 columns: thread i of the warp gets row (lane_base + i).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) LBO>>4 (between the two 8-channel halves of a K=16 slice) |
// [32,46) SBO>>4 (between 8-row groups) | [46,48) version=1 | [61,64) layout=0 (SWIZZLE_NONE).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9,
// 10-12 = 1), both K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- packed tensor-core weights ----------------------------------------------------------
struct TcW {
  const __nv_bfloat16* w = nullptr;  // [n_tile][ci_blk][hi|lo][tap][CIB/8][N_tile][8]
  int K = 0, CIB = 0, n_blk = 0, N_tile = 0, n_ntiles = 0, N_alloc = 0;
  size_t elems() const { return (size_t)n_ntiles * n_blk * 2 * K * CIB * N_tile; }
  size_t chunk_elems() const { return (size_t)2 * K * CIB * N_tile; }
};

// fp32 packed [C_in][K][C_out]  ->  TcW layout (hi/lo bf16, zero padded).
__global__ void pack_tc_weights_kernel(const float* __restrict__ src, int C_in, int C_out, int K, int CIB, int n_blk,
                                       int N_tile, int n_ntiles, __nv_bfloat16* __restrict__ dst) {
  const size_t half = (size_t)K * CIB * N_tile;  // one (n_tile, blk, split) chunk
  const size_t total = (size_t)n_ntiles * n_blk * half;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = r % 8; r /= 8;
    const int n = r % N_tile; r /= N_tile;
    const int g = r % (CIB / 8); r /= (CIB / 8);
    const int k = r % K; r /= K;
    const int blk = r % n_blk; r /= n_blk;
    const int nt = (int)r;
    const int ci = blk * CIB + g * 8 + e, co = nt * N_tile + n;
    float v = 0.f;
    if (ci < C_in && co < C_out) v = src[((size_t)ci * K + k) * C_out + co];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const size_t base = ((size_t)(nt * n_blk + blk) * 2) * half + (((size_t)k * (CIB / 8) + g) * N_tile + n) * 8 + e;
    dst[base] = hi;
    dst[base + half] = lo;
  }
}

constexpr int kTcThreads = 256;
constexpr int kTcM = 128;  // time steps per CTA tile (UMMA M)

struct TcArgs {
  ConvArgs c;
  const __nv_bfloat16* w;
  int CIB, n_blk, N_tile, N_alloc;
};

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// grid = (ceil(T_out/128), n_ntiles, B); block = 256; dynamic smem = see tc_smem_bytes().
template <int K>
__global__ void __launch_bounds__(kTcThreads) conv1d_tc_kernel(const TcArgs p) {
  const ConvArgs& a = p.c;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int halo = (K / 2) * a.dil;
  const int W = kTcM + 2 * halo;              // staged rows
  const int G = p.CIB / 8;                    // channel groups per block
  const uint32_t strip = (uint32_t)W * 16;    // bytes per A strip (one group, all rows)
  const uint32_t a_plane = (uint32_t)G * strip;   // hi plane, lo plane follows
  const uint32_t b_strip = (uint32_t)p.N_tile * 16;
  const uint32_t b_half = (uint32_t)K * G * b_strip;  // one split (hi or lo)
  uint8_t* sA = smem_raw;                      // [2][G][W][16 B]
  uint8_t* sB = sA + 2 * a_plane;              // [2][K][G][N_tile][16 B]
  float* s_pa = (float*)(sB + 2 * b_half);     // [CIB * n_blk] IN affine of the input channels
  float* s_pc = s_pa + p.CIB * p.n_blk;
  uint64_t* bar = (uint64_t*)(s_pc + p.CIB * p.n_blk);
  uint32_t* s_tmem = (uint32_t*)(bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, nt = blockIdx.y, b = blockIdx.z;
  const int t0 = tile * kTcM;

  if (warp == 0) tmem_alloc(s_tmem, (uint32_t)p.N_alloc);
  if (tid == 32) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  for (int c = tid; c < p.CIB * p.n_blk; c += kTcThreads) {
    float pa = 1.f, pc = 0.f;
    if (a.pre_a && c < a.C_in) {
      pa = __ldg(a.pre_a + b * a.C_in + c);
      pc = __ldg(a.pre_c + b * a.C_in + c);
    }
    s_pa[c] = pa;
    s_pc[c] = pc;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const float* in_b = a.in + (long long)b * a.in_bs;
  const uint32_t idesc = umma_idesc_bf16(kTcM, p.N_tile);
  const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
  uint32_t phase = 0;

  for (int blk = 0; blk < p.n_blk; ++blk) {
    if (blk > 0) {  // the MMAs of the previous block must have consumed sA / sB
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    // ---- stage B: contiguous pre-laid-out chunk ----
    {
      const uint4* src = (const uint4*)(p.w + (size_t)(nt * p.n_blk + blk) * 2 * K * p.CIB * p.N_tile);
      uint4* dst = (uint4*)sB;
      const int n16 = (int)(2 * b_half / 16);
      for (int i = tid; i < n16; i += kTcThreads) dst[i] = __ldg(src + i);
    }
    // ---- stage A: (row, group) items; prologue applied once per element ----
    const int ci0 = blk * p.CIB;
    for (int item = tid; item < W * G; item += kTcThreads) {
      const int g = item / W, r = item - g * W;
      const int u = t0 - halo + r;
      uint32_t hi[4], lo[4];
      if (u >= 0 && u < a.T_out && ci0 + g * 8 < a.C_in) {
        const int src_t = (u / a.up) * a.down;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = ci0 + g * 8 + e;
          float x = 0.f;
          if (c < a.C_in) {
            x = __ldg(in_b + (long long)c * a.in_cs + src_t);
            x = fmaf(x, s_pa[c], s_pc[c]);
            if (a.pre_lrelu) x = lrelu(x, a.slope);
          }
          v[e] = x;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * e] - __bfloat162float(h0));
          const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * e + 1] - __bfloat162float(h1));
          hi[e] = pack_bf16x2(h0, h1);
          lo[e] = pack_bf16x2(l0, l1);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) hi[e] = lo[e] = 0u;
      }
      uint8_t* dst = sA + (size_t)g * strip + (size_t)r * 16;
      *(uint4*)dst = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *(uint4*)(dst + a_plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < K; ++k) {
        for (int kc = 0; kc < p.CIB / 16; ++kc) {
          const uint32_t a_off = (uint32_t)(2 * kc) * strip + (uint32_t)(k * a.dil) * 16;
          const uint32_t b_off = (uint32_t)(k * G + 2 * kc) * b_strip;
          const uint64_t a_hi = umma_desc(sA_addr + a_off, strip, 128);
          const uint64_t a_lo = umma_desc(sA_addr + a_plane + a_off, strip, 128);
          const uint64_t b_hi = umma_desc(sB_addr + b_off, b_strip, 128);
          const uint64_t b_lo = umma_desc(sB_addr + b_half + b_off, b_strip, 128);
          const uint32_t first = (blk == 0 && k == 0 && kc == 0) ? 0u : 1u;
          umma_bf16(tmem, a_lo, b_hi, idesc, first);   // small terms first, then the dominant one
          umma_bf16(tmem, a_hi, b_lo, idesc, 1u);
          umma_bf16(tmem, a_hi, b_hi, idesc, 1u);
        }
      }
      umma_commit(bar);
    }
  }
  mbar_wait(bar, phase);
  tc_fence_after();

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4).. (time rows), column half w/4 ----
  const int q = warp & 3, half = warp >> 2;
  const int t = t0 + q * 32 + lane;
  const bool t_ok = t < a.T_out;
  const int seg = tile * 4 + q;  // 32-step statistics segment
  const int n_valid = min(32, a.T_out - (t0 + q * 32));
  const int ncol_half = p.N_tile / 2;
  for (int c0 = half * ncol_half; c0 < (half + 1) * ncol_half; c0 += 8) {
    float v[8];
    tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = nt * p.N_tile + c0 + j;
      if (co >= a.C_out) break;  // warp-uniform
      float x = v[j] + (a.bias ? __ldg(a.bias + co) : 0.f);
      if (t_ok) {
        if (a.res) x += __ldg(a.res + (long long)b * a.res_bs + (long long)co * a.res_cs + t);
        if (a.raw) a.raw[(long long)b * a.raw_bs + (long long)co * a.raw_cs + t] = x;
        if (a.post_lrelu) x = lrelu(x, a.slope);
        if (a.gamma) {
          const long long gi = (long long)b * a.gb_bs + (long long)co * a.gb_cs + t;
          x = fmaf(__ldg(a.gamma + gi), x, __ldg(a.beta + gi));
        }
        if (a.out) a.out[(long long)b * a.out_bs + (long long)co * a.out_cs + t] = x;
      } else {
        x = 0.f;
      }
      if (a.stats && n_valid > 0) {
        float s1 = x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        const float mean = s1 / (float)n_valid;
        const float d = t_ok ? x - mean : 0.f;
        float m2 = d * d;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
        if (lane == 0) a.stats[((long long)b * a.C_out + co) * a.n_tiles + seg] = make_float2(mean, m2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.N_alloc);
}

inline size_t tc_smem_bytes(int K, int dil, int CIB, int n_blk, int N_tile) {
  const int W = kTcM + 2 * (K / 2) * dil;
  return (size_t)2 * (CIB / 8) * W * 16 + (size_t)2 * K * (CIB / 8) * N_tile * 16 + (size_t)2 * CIB * n_blk * 4 + 16;
}

}  // namespace fsvc
