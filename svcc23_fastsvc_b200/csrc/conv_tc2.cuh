// Persistent, software-pipelined tcgen05 1-D convolution over channels-last activations.
//
// This is the building block of the tensor-core forward (plan_tc2.cuh).  Activations live in
// HBM as [B][T][C] fp32 ("NTC": one time step = one contiguous row of channels), which is the
// shape the tensor core wants for its M-side operand (a row of the A tile = one time step) and
// the shape the TMEM epilogue produces (one thread = one time step, its registers = channels),
// so every global access of the hot loop is a 128-bit access to a contiguous row segment.
//
//   D[128 t, N co] (TMEM, fp32) += A_tap[128 t, 16 ci] (smem, bf16 hi|lo) * W_tap[16 ci, N co]
//
// fp32 parity (<= 1e-3, BASELINE.json) is kept with the 3-term bf16 split of conv_tc.cuh.
//
// Work item = (utterance b, 128-step time tile, N tile).  A CTA loops over its items
// (static round-robin) with a 2-deep ring of A buffers and 2 TMEM accumulators:
//
//     stage A(i)  ->  issue MMA(i) (async, one thread)  ->  epilogue(i-1)  ->  stage A(i+1) ...
//
// so the tensor core works on tile i while all threads drain tile i-1 (bias, residual, FiLM
// affine, InstanceNorm partial statistics, stores).  Weights stay resident in shared memory
// for the CTA's lifetime when they fit; otherwise the (N tile, ci block) chunk is streamed
// next to the A block with cp.async.bulk (one thread, mbarrier completion).
//
// Reference semantics implemented by the fused prologue / epilogue: Conv1d1x3 / Conv2d1x3 /
// Conv1d1x1 (layers/upsample.py:76-106, layers/residual_block.py:41-48), Stretch2d / Squeeze2d
// index maps (layers/upsample.py:38-74), _feature_affine + LeakyReLU (fastsvc.py:115-140,
// 56-75), FastSVCDownsampleNet's first conv and 1x1 residual on the raw 1-channel signal
// (fastsvc.py:164-172, "gen" operands below).
#pragma once
#include "conv_tc.cuh"

namespace fsvc {

struct Tc2Args {
  // ---- A operand source -------------------------------------------------------------
  const float* in;     // NTC [B][T_in][in_ld] (channel offset folded into the pointer) | NCT [B][C_in][T_in]
  int in_ld;           // NTC row stride (floats)
  int T_in;            // stored time steps per utterance
  int C_in;
  int in_nct;          // 1: input is (B, C_in, T_in) time-fastest (the caller's PPG tensor)
  int up, down;        // source row of output-rate index u: (u / up) * down
  const float* pre_a;  // [B][C_in] InstanceNorm affine applied on load (or nullptr)
  const float* pre_c;
  // Alternative to pre_a/pre_c (conv_tc3 only): the producer's per-segment (mean, M2) partials [B][pre_nseg][C_in]
  // are merged by the consumer itself (no separate finalize launch): a = rstd, c = pre_e - mean * rstd.
  const float2* pre_stats;
  const float* pre_e;  // [B][C_in] projected speaker embedding, or nullptr
  int pre_nseg;        // 32-step segments per utterance (the last may be short: T_in rows in total)
  float pre_eps;
  int pre_lrelu;
  const float* gen_w;  // != nullptr: `in` is a 1-channel signal [B][T_in] and the C_in operand channels are
  const float* gen_b;  //   gen_b[c] + sum_k gen_w[k*C_in+c] * lrelu(x[u+k-1])   (first conv of a level-0 chain)
  // ---- weights ------------------------------------------------------------------------
  const __nv_bfloat16* w;  // packed [n_tile][ci_blk][hi|lo][tap][CIB/8][N_tile][8]
  int CIB, n_blk, N_tile, n_ntiles;
  int w_resident;
  const float* bias;
  int dil, C_out, T_out;
  // ---- epilogue -------------------------------------------------------------------------
  const float* res;  // NTC [B][T_out][res_ld] added before `raw`
  int res_ld;
  const float* gres_w;  // != nullptr: residual generated from a 1-channel signal: gres_w[c]*x[b][t] + gres_b[c]
  const float* gres_b;
  const float* gres_x;
  float* raw;  // value before activation / FiLM (skip tensors)
  int raw_ld;
  int post_lrelu;
  const float* gamma;  // FiLM: v = gamma*v + beta, both [B][T_out][gb_ld]
  const float* beta;
  int gb_ld;
  float* out;
  int out_ld;
  float2* stats;  // [B][n_seg][C_out] (mean, M2) of the stored value per 32-step segment, or nullptr
  int n_seg;
  float slope;
};

struct Tc2Batch {  // up to 2 independent problems of identical tiling in one launch (the two conditioning branches)
  Tc2Args p[2];
  int n_prob;
  int B, m_tiles;  // items per problem and N tile = B * m_tiles
};

// Blocked channels-last activation layout of the tensor-core forward: [B][ceil(T/32)][ld/4][32 steps][4 channels].
// 32 consecutive time steps of one 4-channel group are 512 contiguous bytes, so a warp whose lanes own
// consecutive time steps (the TMEM epilogue, the A-window staging) touches whole 128-byte lines with every
// 128-bit access -- a plain [T][C] row layout costs one line per lane there.  Offsets are in floats; a channel
// offset co (multiple of 4) is folded into a base pointer as (co / 4) * 128.
__host__ __device__ inline long long ntc_tp(int T) { return (long long)((T + 31) / 32) * 32; }
__host__ __device__ inline long long ntc_row(long long Tp, int ld, int b, int t) {
  return ((long long)b * Tp + (t & ~31)) * ld + (t & 31) * 4;
}
__host__ __device__ inline long long ntc_col(int co) { return (long long)(co >> 2) * 128 + (co & 3); }

constexpr int kTc2Threads = 256;
constexpr int kTc2M = 128;

// Wait for the phase with the given parity.  The suspend-time hint lets the hardware park the thread until
// the phase completes instead of spinning (spinning waiters steal issue slots from the MMA-issuing warp).
__device__ __forceinline__ void mbar_wait2(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running.  launch_dependents lets OUR successor start early
// (every CTA issues it at once: all grids here are single-wave, so nothing of this kernel is left to schedule);
// griddep_wait blocks until the predecessor grid has completed and its writes are visible -- every thread calls
// it before its first access to activations / workspace (weights and parameters are not produced by kernels).
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (UBLKCP), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(taddr)
               : "memory");
  v[0] = __uint_as_float(r0);
  v[1] = __uint_as_float(r1);
  v[2] = __uint_as_float(r2);
  v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void split_store(uint8_t* dst, uint32_t plane, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
    hi[e] = *reinterpret_cast<const uint32_t*>(&h);
    lo[e] = *reinterpret_cast<const uint32_t*>(&l);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst + plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// Stage one ci block of the A window of tile (b, t0) into sA: rows = time steps t0-halo .. t0+127+halo.
__device__ __forceinline__ void tc2_stage_a(const Tc2Args& a, int b, int t0, int blk, int halo, int W, uint8_t* sA,
                                            const float* s_pa, const float* s_pc, int tid) {
  const int Gb = a.CIB >> 3;
  const int ci0 = blk * a.CIB;
  const uint32_t strip = (uint32_t)W * 16u, plane = (uint32_t)Gb * strip;
  const int items = W * Gb;
  if (a.gen_w) {
    const float* x = a.in + (long long)b * a.T_in;
    for (int idx = tid; idx < items; idx += kTc2Threads) {
      const int g = idx / W, r = idx - g * W;
      const int u = t0 - halo + r, c = ci0 + g * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (u >= 0 && u < a.T_out && c < a.C_in) {
        const float x0 = u > 0 ? lrelu(__ldg(x + u - 1), a.slope) : 0.f;
        const float x1 = lrelu(__ldg(x + u), a.slope);
        const float x2 = u + 1 < a.T_out ? lrelu(__ldg(x + u + 1), a.slope) : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float* gw = a.gen_w + c + e;  // packed [tap][C_in]
          float y = __ldg(a.gen_b + c + e);
          y = fmaf(__ldg(gw), x0, y);
          y = fmaf(__ldg(gw + a.C_in), x1, y);
          y = fmaf(__ldg(gw + 2 * a.C_in), x2, y);
          v[e] = a.pre_lrelu ? lrelu(y, a.slope) : y;
        }
      }
      split_store(sA + (size_t)g * strip + (size_t)r * 16, plane, v);
    }
    return;
  }
  if (a.in_nct) {
    const float* in_b = a.in + (long long)b * a.C_in * a.T_in;
    for (int idx = tid; idx < items; idx += kTc2Threads) {
      const int g = idx / W, r = idx - g * W;
      const int u = t0 - halo + r, c = ci0 + g * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (u >= 0 && u < a.T_out && c < a.C_in) {
        const int src = (u / a.up) * a.down;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (c + e < a.C_in) {
            float y = __ldg(in_b + (long long)(c + e) * a.T_in + src);
            if (a.pre_a) y = fmaf(y, s_pa[c + e], s_pc[c + e]);
            v[e] = a.pre_lrelu ? lrelu(y, a.slope) : y;
          }
        }
      }
      split_store(sA + (size_t)g * strip + (size_t)r * 16, plane, v);
    }
    return;
  }
  // channels-last fast path: the loads of up to 4 items per thread are all issued before the first is consumed
  const float* in_b = a.in + (long long)b * a.T_in * a.in_ld;
  const uint32_t gb_magic = 0xFFFFFFFFu / (uint32_t)Gb + 1u;    // exact quotients for idx < 2^16
  const uint32_t up_magic = 0xFFFFFFFFu / (uint32_t)a.up + 1u;  // (unused when up == 1)
  constexpr int CH = 4;
  for (int base = tid; base < items; base += CH * kTc2Threads) {
    float4 q[CH][2];
    bool live[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int idx = base + j * kTc2Threads;
      live[j] = false;
      q[j][0] = q[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < items) {
        const int r = (int)__umulhi((uint32_t)idx, gb_magic), g = idx - r * Gb;
        const int u = t0 - halo + r, c = ci0 + g * 8;
        if (u >= 0 && u < a.T_out && c < a.C_in) {
          const int src = (a.up == 1 ? u : (int)__umulhi((uint32_t)u, up_magic)) * a.down;
          const float4* p = reinterpret_cast<const float4*>(in_b + (long long)src * a.in_ld + c);
          q[j][0] = __ldg(p);
          q[j][1] = __ldg(p + 1);
          live[j] = true;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int idx = base + j * kTc2Threads;
      if (idx < items) {
        const int r = (int)__umulhi((uint32_t)idx, gb_magic), g = idx - r * Gb;
        const int c = ci0 + g * 8;
        float v[8] = {q[j][0].x, q[j][0].y, q[j][0].z, q[j][0].w, q[j][1].x, q[j][1].y, q[j][1].z, q[j][1].w};
        if (live[j]) {
          if (a.pre_a) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], s_pa[c + e], s_pc[c + e]);
          }
          if (a.pre_lrelu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = lrelu(v[e], a.slope);
          }
        }
        split_store(sA + (size_t)g * strip + (size_t)r * 16, plane, v);
      }
    }
  }
}

// Sum over the 32 lanes of 16 per-lane values with 16 shuffles: lane L ends up with the total of value (L >> 1).
__device__ __forceinline__ float warp_reduce16(const float (&x)[16], int lane) {
  float y8[8], y4[4], y2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = b4 ? x[j] : x[j + 8], keep = b4 ? x[j + 8] : x[j];
    y8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = b3 ? y8[j] : y8[j + 4], keep = b3 ? y8[j + 4] : y8[j];
    y4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = b2 ? y4[j] : y4[j + 2], keep = b2 ? y4[j + 2] : y4[j];
    y2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? y2[0] : y2[1], keep = b1 ? y2[1] : y2[0];
  float y = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  y += __shfl_xor_sync(0xffffffffu, y, 1);
  return y;
}

// Epilogue of one finished tile: TMEM -> registers -> fused elementwise -> global (NTC rows).
// Warp w owns TMEM lanes 32*(w%4).. (= time rows) and half (w/4) of the tile's valid output channels.
// The residual / FiLM operands of the first 12-channel pass are fetched by tc2_epi_prefetch, which the
// kernel calls a whole staging phase earlier so their latency is hidden.
struct Tc2EpiPre {
  float4 rs[3], ga[3], be[3];
  float gx;
};
struct Tc2EpiPos {
  int t, cb, half, co_tile;
  long long row;
  bool ok;
};
__device__ __forceinline__ Tc2EpiPos tc2_epi_pos(const Tc2Args& a, int b, int t0, int nt, int warp, int lane) {
  Tc2EpiPos p;
  const int q = warp & 3, h = warp >> 2;
  p.t = t0 + q * 32 + lane;
  p.ok = p.t < a.T_out;
  p.co_tile = nt * a.N_tile;
  const int nvalid = min(a.N_tile, a.C_out - p.co_tile);
  p.half = nvalid >> 1;  // multiple of 4 (C_out % 8 == 0)
  p.cb = h * p.half;     // first column of this thread inside the N tile
  p.row = (long long)b * a.T_out + p.t;
  return p;
}
__device__ __forceinline__ void tc2_epi_load(const Tc2Args& a, const Tc2EpiPos& p, int c0, Tc2EpiPre& e) {
  const int n4 = min(3, (p.half - c0) >> 2);
  const int co = p.co_tile + p.cb + c0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j < n4 && p.ok) {
      if (a.res) e.rs[j] = __ldg(reinterpret_cast<const float4*>(a.res + p.row * a.res_ld + co) + j);
      if (a.gamma) {
        e.ga[j] = __ldg(reinterpret_cast<const float4*>(a.gamma + p.row * a.gb_ld + co) + j);
        e.be[j] = __ldg(reinterpret_cast<const float4*>(a.beta + p.row * a.gb_ld + co) + j);
      }
    }
  }
}
__device__ __forceinline__ void tc2_epi_prefetch(const Tc2Args& a, int b, int t0, int nt, int warp, int lane,
                                                 Tc2EpiPre& e) {
  const Tc2EpiPos p = tc2_epi_pos(a, b, t0, nt, warp, lane);
  e.gx = 0.f;
  if (a.gres_w && p.ok) e.gx = __ldg(a.gres_x + p.row);
  tc2_epi_load(a, p, 0, e);
}

__device__ __forceinline__ void tc2_epilogue(const Tc2Args& a, int b, int t0, int nt, uint32_t tmem_acc, int warp,
                                             int lane, Tc2EpiPre& pre) {
  const Tc2EpiPos p = tc2_epi_pos(a, b, t0, nt, warp, lane);
  const int q = warp & 3;
  const bool ok = p.ok;
  const int n_rows_seg = min(32, a.T_out - (t0 + q * 32));  // valid rows of this warp's segment (may be <= 0)
  const int seg = (t0 >> 5) + q;
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)p.cb;
  const float gx = pre.gx;

  for (int c0 = 0; c0 < p.half; c0 += 12) {
    const int n4 = min(3, (p.half - c0) >> 2);
    float v[16];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j < n4) tmem_ld4_nowait(taddr + (uint32_t)(c0 + 4 * j), v + 4 * j);
    const int co = p.co_tile + p.cb + c0;  // first global output channel of this pass
    if (c0 > 0) tc2_epi_load(a, p, c0, pre);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (j < n4) {
        const float rs[4] = {pre.rs[j].x, pre.rs[j].y, pre.rs[j].z, pre.rs[j].w};
        const float ga[4] = {pre.ga[j].x, pre.ga[j].y, pre.ga[j].z, pre.ga[j].w};
        const float be[4] = {pre.be[j].x, pre.be[j].y, pre.be[j].z, pre.be[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 4 * j + e;
          float x = v[i] + __ldg(a.bias + co + i);
          if (a.res) x += ok ? rs[e] : 0.f;
          if (a.gres_w) x += fmaf(__ldg(a.gres_w + co + i), gx, __ldg(a.gres_b + co + i));
          v[i] = x;
        }
        if (a.raw && ok)
          reinterpret_cast<float4*>(a.raw + p.row * a.raw_ld + co)[j] =
              make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 4 * j + e;
          float x = v[i];
          if (a.post_lrelu) x = lrelu(x, a.slope);
          if (a.gamma) x = ok ? fmaf(ga[e], x, be[e]) : 0.f;
          v[i] = x;
        }
        if (a.out && ok)
          reinterpret_cast<float4*>(a.out + p.row * a.out_ld + co)[j] =
              make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    if (a.stats && n_rows_seg > 0) {
      // per-channel (mean, M2) over the <= 32 rows of this warp.  One-pass sums of the deviations from
      // the segment's first sample (no cancellation), reduced across the lanes with a transposing
      // butterfly (44 shuffles per 12 channels); merged over segments in double by in_finalize2_kernel.
      const int nc = 4 * n4;
      float d1[16], d2[16];
      float my_piv = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        d1[i] = 0.f;
        d2[i] = 0.f;
        if (i < 12 && i < nc) {
          const float piv = __shfl_sync(0xffffffffu, v[i], 0);
          const float d = ok ? v[i] - piv : 0.f;
          d1[i] = d;
          d2[i] = d * d;
          if ((lane >> 1) == i) my_piv = piv;
        }
      }
      const float s1 = warp_reduce16(d1, lane), s2 = warp_reduce16(d2, lane);
      const float inv_n = 1.f / (float)n_rows_seg;
      const int ch = lane >> 1;
      if ((lane & 1) == 0 && ch < nc)
        a.stats[((long long)b * a.n_seg + seg) * a.C_out + co + ch] =
            make_float2(my_piv + s1 * inv_n, fmaxf(s2 - s1 * s1 * inv_n, 0.f));
    }
  }
}

struct Tc2Smem {  // byte offsets into dynamic shared memory (computed on the host and the device alike)
  uint32_t w_off, a_off[2], b_off[2], pa_off, bar_off, total;
  uint32_t a_bytes, b_blk_bytes;
};
__host__ __device__ inline Tc2Smem tc2_smem_layout(int K, int CIB, int n_blk, int N_tile, int resident, int W,
                                                   int C_in) {
  Tc2Smem s;
  const uint32_t Gb = CIB / 8;
  s.b_blk_bytes = 2u * K * Gb * N_tile * 16u;
  s.a_bytes = 2u * Gb * W * 16u;
  uint32_t off = 0;
  s.w_off = off;
  if (resident) off += s.b_blk_bytes * n_blk;
  for (int i = 0; i < 2; ++i) {
    s.a_off[i] = off;
    off += s.a_bytes;
  }
  for (int i = 0; i < 2; ++i) {
    s.b_off[i] = off;
    if (!resident) off += s.b_blk_bytes;
  }
  s.pa_off = off;
  off += 2u * ((C_in + 7) / 8 * 8) * 4u;
  off = (off + 15u) & ~15u;
  s.bar_off = off;
  off += 8 * 8 + 16;
  s.total = off;
  return s;
}

template <int K>
__device__ __forceinline__ void tc2_body(const Tc2Args& a, const Tc2Batch& pb, uint8_t* smem_raw) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // static work assignment: N tile, then (b, time tile) round-robin
  const int rest = blockIdx.x / pb.n_prob;
  const int nt = rest % a.n_ntiles;
  const int first = rest / a.n_ntiles;
  const int step = gridDim.x / (pb.n_prob * a.n_ntiles);
  const int n_m = pb.B * pb.m_tiles;

  const int halo = (K / 2) * a.dil;
  const int W = kTc2M + 2 * halo;
  const Tc2Smem L = tc2_smem_layout(K, a.CIB, a.n_blk, a.N_tile, a.w_resident, W, a.C_in);
  uint8_t* sW = smem_raw + L.w_off;
  float* s_pa = reinterpret_cast<float*>(smem_raw + L.pa_off);
  const int cpad = (a.C_in + 7) / 8 * 8;
  float* s_pc = s_pa + cpad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bar_off);
  uint64_t* a_empty = bars;      // [2] MMAs that read A/B ring slot s have completed
  uint64_t* acc_full = bars + 2; // [2] accumulator s holds a finished tile
  uint64_t* b_full = bars + 4;   // [2] streamed weight block landed in ring slot s
  uint64_t* w_full = bars + 6;   // resident weights landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 8);

  // accumulator column stride / allocation: power of two >= N_tile
  uint32_t acc_stride = 32;
  while (acc_stride < (uint32_t)a.N_tile) acc_stride <<= 1;
  if (warp == 0) tmem_alloc(s_tmem, 2 * acc_stride);
  if (tid == 32) {
    for (int i = 0; i < 7; ++i) mbar_init(bars + i, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const __nv_bfloat16* w_nt = a.w + (size_t)nt * a.n_blk * (L.b_blk_bytes / 2);
  if (a.w_resident && tid == 0) {
    const uint32_t total = L.b_blk_bytes * a.n_blk;
    mbar_expect_tx(w_full, total);
    for (uint32_t o = 0; o < total; o += 32768u)
      bulk_g2s(sW + o, reinterpret_cast<const uint8_t*>(w_nt) + o, min(32768u, total - o), w_full);
  }
  bool w_ready = !a.w_resident;

  const uint32_t idesc = umma_idesc_bf16(kTc2M, a.N_tile);
  const uint32_t Gb = a.CIB >> 3;
  const uint32_t strip = (uint32_t)W * 16u, a_plane = Gb * strip;
  const uint32_t b_strip = (uint32_t)a.N_tile * 16u, b_half = (uint32_t)K * Gb * b_strip;

  uint32_t ring_uses0 = 0, ring_uses1 = 0;  // completed fills of A/B ring slot 0 / 1
  uint32_t ring_pos = 0;
  int it = 0;
  int pb_b = 0, pb_t0 = 0;  // previous tile (epilogue pending)

  for (int m = first; m < n_m; m += step, ++it) {
    const int b = m / pb.m_tiles, t0 = (m - b * pb.m_tiles) * kTc2M;
    if (a.pre_a) {  // (every thread is past the previous tile's staging: it crossed that tile's barrier)
      for (int c = tid; c < cpad; c += kTc2Threads) {
        s_pa[c] = c < a.C_in ? __ldg(a.pre_a + (long long)b * a.C_in + c) : 1.f;
        s_pc[c] = c < a.C_in ? __ldg(a.pre_c + (long long)b * a.C_in + c) : 0.f;
      }
      __syncthreads();
    }
    const uint32_t acc = (uint32_t)(it & 1);
    for (int blk = 0; blk < a.n_blk; ++blk, ++ring_pos) {
      const uint32_t s = ring_pos & 1u;
      const uint32_t uses = s ? ring_uses1 : ring_uses0;
      if (uses > 0) mbar_wait2(a_empty + s, (uses - 1) & 1u);
      uint8_t* sA = smem_raw + L.a_off[0] + s * L.a_bytes;
      uint8_t* sB = a.w_resident ? sW + (size_t)blk * L.b_blk_bytes : smem_raw + L.b_off[0] + s * L.b_blk_bytes;
      if (!a.w_resident && tid == 0) {
        mbar_expect_tx(b_full + s, L.b_blk_bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(w_nt) + (size_t)blk * L.b_blk_bytes;
        for (uint32_t o = 0; o < L.b_blk_bytes; o += 32768u)
          bulk_g2s(sB + o, src + o, min(32768u, L.b_blk_bytes - o), b_full + s);
      }
      Tc2EpiPre pre;
      if (blk == 0 && it > 0) tc2_epi_prefetch(a, pb_b, pb_t0, nt, warp, lane, pre);
      tc2_stage_a(a, b, t0, blk, halo, W, sA, s_pa, s_pc, tid);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        if (!w_ready) {
          mbar_wait2(w_full, 0);
          w_ready = true;
        }
        if (!a.w_resident) mbar_wait2(b_full + s, uses & 1u);
        tc_fence_after();
        const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
        const uint32_t d_tmem = tmem + acc * acc_stride;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          for (uint32_t kc = 0; kc < (uint32_t)a.CIB / 16u; ++kc) {
            const uint32_t a_off = 2u * kc * strip + (uint32_t)(k * a.dil) * 16u;
            const uint32_t b_off = ((uint32_t)k * Gb + 2u * kc) * b_strip;
            const uint64_t a_hi = umma_desc(sA_addr + a_off, strip, 128);
            const uint64_t a_lo = umma_desc(sA_addr + a_plane + a_off, strip, 128);
            const uint64_t b_hi = umma_desc(sB_addr + b_off, b_strip, 128);
            const uint64_t b_lo = umma_desc(sB_addr + b_half + b_off, b_strip, 128);
            const uint32_t accum = (blk == 0 && k == 0 && kc == 0) ? 0u : 1u;
            umma_bf16(d_tmem, a_lo, b_hi, idesc, accum);  // small terms first, then the dominant one
            umma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
            umma_bf16(d_tmem, a_hi, b_hi, idesc, 1u);
          }
        }
        umma_commit(a_empty + s);
        if (blk == a.n_blk - 1) umma_commit(acc_full + acc);
      }
      if (s) ring_uses1++;
      else ring_uses0++;
      if (blk == 0 && it > 0) {
        // drain the previous tile while the tensor core works on this one
        const uint32_t pacc = (uint32_t)((it - 1) & 1);
        mbar_wait2(acc_full + pacc, (uint32_t)(((it - 1) >> 1) & 1));
        tc_fence_after();
        tc2_epilogue(a, pb_b, pb_t0, nt, tmem + pacc * acc_stride, warp, lane, pre);
        tc_fence_before();
      }
    }
    pb_b = b;
    pb_t0 = t0;
  }
  if (it > 0) {
    const uint32_t pacc = (uint32_t)((it - 1) & 1);
    mbar_wait2(acc_full + pacc, (uint32_t)(((it - 1) >> 1) & 1));
    tc_fence_after();
    Tc2EpiPre pre;
    tc2_epi_prefetch(a, pb_b, pb_t0, nt, warp, lane, pre);
    tc2_epilogue(a, pb_b, pb_t0, nt, tmem + pacc * acc_stride, warp, lane, pre);
    tc_fence_before();
  }
  if (!w_ready && tid == 0) mbar_wait2(w_full, 0);  // never exit with a bulk copy in flight
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 2 * acc_stride);
}

// grid = multiple of n_prob * n_ntiles; block = 256; dynamic smem = tc2_smem_layout().total.
// The two problems get their own copy of the body so that every argument is a constant-bank operand.
template <int K>
__global__ void __launch_bounds__(kTc2Threads, 2) conv_tc2_kernel(const __grid_constant__ Tc2Batch pb) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  if (pb.n_prob == 1 || (blockIdx.x & 1) == 0) tc2_body<K>(pb.p[0], pb, smem_raw);
  else tc2_body<K>(pb.p[1], pb, smem_raw);
}

// ---- small companions -------------------------------------------------------------------

// Merge the per-32-step (mean, M2) partials [B][n_seg][C] of one (b, c) (Chan et al., double, fixed order)
// and emit the affine the next conv applies on load: a = rstd, c = e - mean*rstd.
// InstanceNorm2d: biased variance over the whole time axis, eps inside the sqrt (fastsvc.py:76,138).
// One warp per (b, c): lane l merges segments l, l+32, ... (loads issued 8 at a time), then a 5-step
// butterfly merges the lanes.  grid = ceil(B*C / 8) blocks of 256 threads.
__device__ __forceinline__ double shfl_xor_d(double v, int o) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return __hiloint2double(hi, lo);
}
__global__ void __launch_bounds__(256) in_finalize2_kernel(const float2* __restrict__ stats, int n_seg, int T, int C,
                                                           int BC, const float* __restrict__ e, float eps,
                                                           float* __restrict__ out_a, float* __restrict__ out_c) {
  const int lane = threadIdx.x & 31;
  const int bc = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (bc >= BC) return;
  const int b = bc / C, c = bc - b * C;
  const float2* sp = stats + (long long)b * n_seg * C + c;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int s0 = lane; s0 < n_seg; s0 += 8 * 32) {
    float2 pv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sg = s0 + 32 * u;
      pv[u] = sg < n_seg ? __ldg(sp + (long long)sg * C) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sg = s0 + 32 * u;
      if (sg < n_seg) {
        const double nb = (double)min(32, T - sg * 32);
        const double d = (double)pv[u].x - mean, nn = n + nb;
        mean += d * nb / nn;
        m2 += (double)pv[u].y + d * d * n * nb / nn;
        n = nn;
      }
    }
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n2 = shfl_xor_d(n, o), mean2 = shfl_xor_d(mean, o), m22 = shfl_xor_d(m2, o);
    // merge (lower lane's partial, upper lane's partial) in that order on both lanes: bitwise symmetric
    const bool low = (lane & o) == 0;
    const double na = low ? n : n2, ma = low ? mean : mean2, qa = low ? m2 : m22;
    const double nb = low ? n2 : n, mb = low ? mean2 : mean, qb = low ? m22 : m2;
    const double nn = na + nb;
    if (nn > 0.0) {
      const double d = mb - ma;
      mean = ma + d * nb / nn;
      m2 = qa + qb + d * d * na * nb / nn;
    }
    n = nn;
  }
  if (lane == 0) {
    const double var = m2 / (double)T;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    out_a[bc] = (float)rstd;
    out_c[bc] = (float)((double)(e ? e[bc] : 0.f) - mean * rstd);
  }
}

// All stages' speaker projections in one launch: e_i[b][c] = bias_i[c] + W_i[c] . normalize(spk[b])
// (nn.Linear(F.normalize(spk_emb)), fastsvc.py:135-137).  grid = (B, n_stages, ceil(C_max/32)), block = 256:
// a warp owns 4 output channels and keeps all their loads in flight.
struct SpkProjArgs {
  const float* W[8];
  const float* bias[8];
  float* e[8];
  int C[8];
};
__global__ void __launch_bounds__(256) spk_project_all_kernel(const float* __restrict__ spk, int S, SpkProjArgs p) {
  __shared__ float red[8];
  __shared__ float inv_norm;
  const int b = blockIdx.x, st = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C[st];
  const int c0 = blockIdx.z * 32 + warp * 4;
  if (blockIdx.z * 32 >= C) return;
  const float* x = spk + (long long)b * S;
  float ss = 0.f;
  for (int j = tid; j < S; j += 256) ss = fmaf(x[j], x[j], ss);
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[i];
    inv_norm = 1.f / fmaxf(sqrtf(v), 1e-12f);
  }
  __syncthreads();
  const float inv = inv_norm;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = lane; j < S; j += 32) {
    const float xv = x[j] * inv;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c0 + i < C) acc[i] = fmaf(__ldg(p.W[st] + (long long)(c0 + i) * S + j), xv, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && c0 + i < C) p.e[st][(long long)b * C + c0 + i] = v + p.bias[st][c0 + i];
  }
}

// conv_last (Conv1d1x1, fastsvc.py:301,330) from blocked channels-last x to (B, C_out, T):
// one thread per time step.  w is the packed fp32 layout [C][C_out].
__global__ void __launch_bounds__(256) conv_last_ntc_kernel(const float* __restrict__ x, int C, int T, long long BT,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            int C_out, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= BT) return;
  const long long b = i / T;
  const int t = (int)(i - b * T);
  const float4* xp = reinterpret_cast<const float4*>(x + ntc_row(ntc_tp(T), C, (int)b, t));
  for (int co = 0; co < C_out; ++co) {
    float acc = __ldg(bias + co);
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = __ldg(xp + 32 * c4);
      const float* wp = w + (long long)(4 * c4) * C_out + co;
      acc = fmaf(v.x, __ldg(wp), acc);
      acc = fmaf(v.y, __ldg(wp + C_out), acc);
      acc = fmaf(v.z, __ldg(wp + 2 * C_out), acc);
      acc = fmaf(v.w, __ldg(wp + 3 * C_out), acc);
    }
    out[(b * C_out + co) * T + t] = acc;
  }
}

}  // namespace fsvc
