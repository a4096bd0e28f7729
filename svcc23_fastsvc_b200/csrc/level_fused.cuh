// Fused conditioning level for the full-rate (1-channel input) level of the FastSVC generator.
//
// One kernel computes, for BOTH conditioning branches (loudness, sine excitation) of level 0,
//
//   a1 = Conv3_d1(lrelu(x))          x: the raw 1-channel signal            fastsvc.py:170-172
//   a2 = Conv3_d2(lrelu(a1))                                                 :173-175
//   y  = Conv3_d4(lrelu(a2)) + Conv1x1(x)                                    :164-167, 176-178, 190-192
//   h  = lrelu(Conv3_d1(y))           FastSVCFiLMNet.conv                    :209, 229
//   [gamma | beta] = Conv3_d1([h_lft | h_sine])  merged conv_scale/conv_shift of both branches, summed
//                                                                            :210-218, 230-231, 127-130
//
// and writes only what later kernels read: gamma|beta [B][T][2C] and the decimated level output
// y[::s] [B][T/s][C] per branch (the next level's input, Squeeze2d: layers/upsample.py:64-74).
// Every intermediate activation stays on chip: a layer's fp32 result is read from TMEM by the
// worker warps, activated, split into bf16 hi|lo and written straight back to shared memory in the
// UMMA K-major canonical layout as the next layer's A operand (dilated taps = descriptor row shifts).
// Unfused, this level moves ~12 full-rate activation tensors through HBM; fused it moves one.
//
// Work item = (utterance, 238 output steps): 2 MMA M-tiles (256 rows) per layer, halo 9 per side.
// 12 worker warps (row quarter x channel group) + 1 MMA/weights warp, one CTA per SM; all conv
// weights of the level (101 KB of bf16 hi|lo at C=24) stay resident in shared memory.
#pragma once
#include "conv_tc3.cuh"

namespace fsvc {

constexpr int kLfWorkers = 384;            // 12 warps: (lane quarter q) x (channel group select 0..2)
constexpr int kLfThreads = kLfWorkers + 32;
constexpr int kLfValid = 238;              // output steps per work item
constexpr int kLfRows = 272;               // rows of an activation buffer (time t0-9 .. t0+262)
constexpr int kLfHalo = 9;

struct LevelFusedArgs {
  const float* sig[2];       // [B][T] raw 1-channel signals: loudness, sine
  const float* c1_w[2];      // fp32 packed [tap][C]
  const float* c1_b[2];
  const float* r1_w[2];      // fp32 [C]
  const float* r1_b[2];
  const __nv_bfloat16* w_c2[2];   // tc2-packed (CIB = Cpad, N_tile = Npad)
  const __nv_bfloat16* w_c4[2];
  const __nv_bfloat16* w_film[2];
  const __nv_bfloat16* w_out;     // merged film_out, tc2-packed (CIB = 2C, N_tile = round16(2C))
  const float* b_c2[2];
  const float* b_c4[2];
  const float* b_film[2];
  const float* b_out;        // [2C]
  float* y_dec[2];           // [B][T/dec][C] or nullptr
  float* gb;                 // [B][T][2C]
  int C, T, B, dec;
  int n_tiles;               // ceil(T / kLfValid)
  int Gp;                    // padded groups of a C->C conv (CIB/8)
  int N1;                    // N tile of the C->C convs
  int N2;                    // N tile of film_out
  float slope;
};

struct LevelFusedSmem {
  uint32_t off_w_c2[2], off_w_c4[2], off_w_film[2], off_w_out, off_buf[3], off_sig, off_par, off_bar, total;
  uint32_t w1_bytes, w2_bytes, buf_bytes;
};
__host__ __device__ inline LevelFusedSmem level_fused_smem(int C, int Gp, int N1, int N2) {
  LevelFusedSmem s;
  const uint32_t G2 = 2 * C / 8;
  s.w1_bytes = 2u * 3u * Gp * N1 * 16u;
  s.w2_bytes = 2u * 3u * G2 * N2 * 16u;
  s.buf_bytes = 2u * Gp * kLfRows * 16u;
  uint32_t off = 0;
  for (int br = 0; br < 2; ++br) {
    s.off_w_c2[br] = off; off += s.w1_bytes;
    s.off_w_c4[br] = off; off += s.w1_bytes;
    s.off_w_film[br] = off; off += s.w1_bytes;
  }
  s.off_w_out = off; off += s.w2_bytes;
  for (int i = 0; i < 3; ++i) { s.off_buf[i] = off; off += s.buf_bytes; }
  s.off_sig = off; off += 2u * kLfRows * 4u;
  s.off_par = off; off += (2u * (3u + 1u + 1u + 1u + 1u + 1u + 1u) * 32u + 64u) * 4u;  // per-branch small fp32 params
  off = (off + 15u) & ~15u;
  s.off_bar = off; off += 8 * 8 + 16;
  s.total = off;
  return s;
}

// small fp32 parameters in shared memory, per branch (each padded to 32 floats): c1 w tap0..2, c1 b, r1 w, r1 b,
// b_c2, b_c4, b_film; then b_out (64 floats)
enum { kLfC1W = 0, kLfC1B = 3, kLfR1W = 4, kLfR1B = 5, kLfBC2 = 6, kLfBC4 = 7, kLfBFilm = 8, kLfParPerBranch = 9 };

__global__ void __launch_bounds__(kLfThreads, 1) level0_fused_kernel(const __grid_constant__ LevelFusedArgs p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* const smem = smem_raw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, G = C >> 3, Gp = p.Gp;
  const LevelFusedSmem L = level_fused_smem(C, Gp, p.N1, p.N2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* bar_ready = bars;        // workers -> MMA: the layer input in shared memory is complete (count 384)
  uint64_t* bar_acc = bars + 1;      // [2] MMA -> workers: accumulator of M-tile 0 / 1 is complete
  uint64_t* bar_w = bars + 3;        // weights landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_sig = reinterpret_cast<float*>(smem + L.off_sig);   // [2][kLfRows]: index i <-> time t0 - 10 + i
  float* s_par = reinterpret_cast<float*>(smem + L.off_par);
  float* s_bout = s_par + 2 * kLfParPerBranch * 32;
  const uint32_t strip = kLfRows * 16u, plane = (uint32_t)Gp * strip;

  if (tid == 0) {
    mbar_init(bar_ready, kLfWorkers);
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc + 1, 1);
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(s_tmem, 128);
  // zero the activation buffers once: the K-padding groups are never written again and must stay finite
  for (uint32_t i = tid; i < 3u * L.buf_bytes / 16u; i += kLfThreads)
    reinterpret_cast<uint4*>(smem + L.off_buf[0])[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2 * kLfParPerBranch * 32 + 64; i += kLfThreads) {
    float v = 0.f;
    if (i < 2 * kLfParPerBranch * 32) {
      const int br = i / (kLfParPerBranch * 32), r = (i / 32) % kLfParPerBranch, c = i & 31;
      if (c < C) {
        if (r < 3) v = p.c1_w[br][r * C + c];
        else if (r == kLfC1B) v = p.c1_b[br][c];
        else if (r == kLfR1W) v = p.r1_w[br][c];
        else if (r == kLfR1B) v = p.r1_b[br][c];
        else if (r == kLfBC2) v = p.b_c2[br][c];
        else if (r == kLfBC4) v = p.b_c4[br][c];
        else v = p.b_film[br][c];
      }
    } else {
      const int c = i - 2 * kLfParPerBranch * 32;
      if (c < 2 * C) v = p.b_out[c];
    }
    s_par[i] = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const int n_items = p.B * p.n_tiles;

  if (warp == 12) {
    // ================= weights + MMA issuer =================
    if (lane == 0) {
      const uint32_t total = 6u * L.w1_bytes + L.w2_bytes;
      mbar_expect_tx(bar_w, total);
      for (int br = 0; br < 2; ++br) {
        bulk_g2s(smem + L.off_w_c2[br], p.w_c2[br], L.w1_bytes, bar_w);
        bulk_g2s(smem + L.off_w_c4[br], p.w_c4[br], L.w1_bytes, bar_w);
        bulk_g2s(smem + L.off_w_film[br], p.w_film[br], L.w1_bytes, bar_w);
      }
      bulk_g2s(smem + L.off_w_out, p.w_out, L.w2_bytes, bar_w);
      mbar_wait2(bar_w, 0);
      const uint32_t idesc1 = umma_idesc_bf16(128, p.N1), idesc2 = umma_idesc_bf16(128, p.N2);
      const uint32_t buf_addr[3] = {smem_u32(smem + L.off_buf[0]), smem_u32(smem + L.off_buf[1]),
                                    smem_u32(smem + L.off_buf[2])};
      uint32_t ready_phase = 0;
      // one C->C layer: A = buffer `src`, output window starts at time offset s (row s + 9), dilation d
      auto issue_layer = [&](uint32_t a_base, uint32_t w_addr, int s, int d) {
        mbar_wait2(bar_ready, ready_phase);
        ready_phase ^= 1u;
        tc_fence_after();
        const uint32_t b_strip = (uint32_t)p.N1 * 16u, b_half = 3u * Gp * b_strip;
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t d_tmem = tmem + (uint32_t)mt * 64u;
          for (int k = 0; k < 3; ++k) {
            const uint32_t row = (uint32_t)(s + kLfHalo + (k - 1) * d + 128 * mt);
            for (int kc = 0; kc < Gp / 2; ++kc) {
              const uint32_t a_off = 2u * kc * strip + row * 16u;
              const uint32_t b_off = ((uint32_t)k * Gp + 2u * kc) * b_strip;
              const uint64_t a_hi = umma_desc(a_base + a_off, strip, 128);
              const uint64_t a_lo = umma_desc(a_base + plane + a_off, strip, 128);
              const uint64_t b_hi = umma_desc(w_addr + b_off, b_strip, 128);
              const uint64_t b_lo = umma_desc(w_addr + b_half + b_off, b_strip, 128);
              const uint32_t accum = (k == 0 && kc == 0) ? 0u : 1u;
              umma_bf16(d_tmem, a_lo, b_hi, idesc1, accum);
              umma_bf16(d_tmem, a_hi, b_lo, idesc1, 1u);
              umma_bf16(d_tmem, a_hi, b_hi, idesc1, 1u);
            }
          }
          umma_commit(bar_acc + mt);
        }
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        for (int br = 0; br < 2; ++br) {
          const uint32_t hbuf = buf_addr[1 + br];
          issue_layer(buf_addr[0], smem_u32(smem + L.off_w_c2[br]), -6, 2);    // a1 (buf 0) -> a2 (buf 1+br)
          issue_layer(hbuf, smem_u32(smem + L.off_w_c4[br]), -2, 4);           // a2 -> y (buf 0)
          issue_layer(buf_addr[0], smem_u32(smem + L.off_w_film[br]), -1, 1);  // y -> h (buf 1+br)
        }
        // merged film_out over the virtual channel concat [h_lft (buf 1) | h_sine (buf 2)]
        mbar_wait2(bar_ready, ready_phase);
        ready_phase ^= 1u;
        tc_fence_after();
        const uint32_t G2 = 2u * G, b_strip = (uint32_t)p.N2 * 16u, b_half = 3u * G2 * b_strip;
        const uint32_t w_addr = smem_u32(smem + L.off_w_out);
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t d_tmem = tmem + (uint32_t)mt * 64u;
          for (int k = 0; k < 3; ++k) {
            const uint32_t row = (uint32_t)(0 + kLfHalo + (k - 1) + 128 * mt);
            for (uint32_t kc = 0; kc < G2 / 2; ++kc) {
              const uint32_t v0 = 2u * kc, v1 = v0 + 1;
              const uint32_t addr0 = buf_addr[1 + (v0 >= (uint32_t)G)] + (v0 % G) * strip + row * 16u;
              const uint32_t addr1 = buf_addr[1 + (v1 >= (uint32_t)G)] + (v1 % G) * strip + row * 16u;
              const uint32_t lbo = addr1 - addr0;
              const uint32_t b_off = ((uint32_t)k * G2 + v0) * b_strip;
              const uint64_t a_hi = umma_desc(addr0, lbo, 128);
              const uint64_t a_lo = umma_desc(addr0 + plane, lbo, 128);
              const uint64_t b_hi = umma_desc(w_addr + b_off, b_strip, 128);
              const uint64_t b_lo = umma_desc(w_addr + b_half + b_off, b_strip, 128);
              const uint32_t accum = (k == 0 && kc == 0) ? 0u : 1u;
              umma_bf16(d_tmem, a_lo, b_hi, idesc2, accum);
              umma_bf16(d_tmem, a_hi, b_lo, idesc2, 1u);
              umma_bf16(d_tmem, a_hi, b_hi, idesc2, 1u);
            }
          }
          umma_commit(bar_acc + mt);
        }
      }
    }
  } else {
    // ================= workers =================
    const int q = warp & 3, gsel = warp >> 2;  // TMEM lane quarter; channel group select (0..2)
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.n_tiles, t0 = (item - b * p.n_tiles) * kLfValid;
      // ---- raw signal windows of both branches ----
      for (int i = tid; i < 2 * kLfRows; i += kLfWorkers) {
        const int br = i >= kLfRows, j = i - br * kLfRows;
        const int t = t0 - 10 + j;
        s_sig[i] = (t >= 0 && t < p.T) ? __ldg(p.sig[br] + (long long)b * p.T + t) : 0.f;
      }
      named_bar_sync(1, kLfWorkers);
      for (int br = 0; br < 2; ++br) {
        const float* sg = s_sig + br * kLfRows;
        const float* par = s_par + br * kLfParPerBranch * 32;
        // ---- a1 = Conv3_d1(lrelu(x)) on the CUDA cores, stored as lrelu(a1): rows time t0-8 .. t0+247 ----
        {
          uint8_t* dst = smem + L.off_buf[0];
          for (int idx = tid; idx < 256 * G; idx += kLfWorkers) {
            const int g = idx >> 8, r = idx & 255;
            const int tau = r - 8, t = t0 + tau;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
            if (t >= 0 && t < p.T) {
              const float x0 = sg[tau + 9], x1 = sg[tau + 10], x2 = sg[tau + 11];
              const float l0 = fmaxf(x0, x0 * p.slope), l1 = fmaxf(x1, x1 * p.slope), l2 = fmaxf(x2, x2 * p.slope);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int c = g * 8 + e;
                float y = par[kLfC1B * 32 + c];
                y = fmaf(par[(kLfC1W + 0) * 32 + c], l0, y);
                y = fmaf(par[(kLfC1W + 1) * 32 + c], l1, y);
                y = fmaf(par[(kLfC1W + 2) * 32 + c], l2, y);
                v[e] = fmaxf(y, y * p.slope);
              }
            }
            split_store(dst + (uint32_t)g * strip + (uint32_t)(tau + kLfHalo) * 16u, plane, v);
          }
          fence_proxy_async();
          mbar_arrive(bar_ready);
        }
        // ---- three tensor-core layers: TMEM -> (bias, residual, activation, zero padding) -> bf16 hi|lo -> smem ----
        for (int layer = 0; layer < 3; ++layer) {
          const int s = layer == 0 ? -6 : (layer == 1 ? -2 : -1);
          uint8_t* dst = smem + (layer == 1 ? L.off_buf[0] : L.off_buf[1 + br]);
          const float* bias = par + (layer == 0 ? kLfBC2 : (layer == 1 ? kLfBC4 : kLfBFilm)) * 32;
          for (int mt = 0; mt < 2; ++mt) {
            mbar_wait2(bar_acc + mt, acc_phase);
            tc_fence_after();
            const int tau = s + 128 * mt + q * 32 + lane, t = t0 + tau;
            const bool in_seq = t >= 0 && t < p.T;
            for (int g = gsel; g < G; g += 3) {
              float v[8];
              tmem_ld8(tmem + (uint32_t)mt * 64u + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 8), v);
              if (in_seq) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += bias[g * 8 + e];
                if (layer == 1) {  // + Conv1x1(x), then the level output y (kept raw for the FiLM conv)
                  const float x = sg[tau + 10];
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] += fmaf(par[kLfR1W * 32 + g * 8 + e], x, par[kLfR1B * 32 + g * 8 + e]);
                  if (p.y_dec[br] && tau >= 0 && tau < kLfValid && t % p.dec == 0) {
                    float4* yp = reinterpret_cast<float4*>(p.y_dec[br] + ((long long)b * (p.T / p.dec) + t / p.dec) * C + g * 8);
                    yp[0] = make_float4(v[0], v[1], v[2], v[3]);
                    yp[1] = make_float4(v[4], v[5], v[6], v[7]);
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], v[e] * p.slope);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.f;  // zero padding of the next conv's input
              }
              split_store(dst + (uint32_t)g * strip + (uint32_t)(tau + kLfHalo) * 16u, plane, v);
            }
          }
          acc_phase ^= 1u;
          fence_proxy_async();
          tc_fence_before();
          // (h of the first branch is consumed only by film_out: its completion rides on the next arrive)
          if (!(layer == 2 && br == 0)) mbar_arrive(bar_ready);
        }
      }
      // ---- merged film_out: gamma | beta rows straight to HBM ----
      for (int mt = 0; mt < 2; ++mt) {
        mbar_wait2(bar_acc + mt, acc_phase);
        tc_fence_after();
        const int tau = 128 * mt + q * 32 + lane, t = t0 + tau;
        const bool ok = tau < kLfValid && t < p.T;
        for (int g = gsel; g < 2 * G; g += 3) {
          float v[8];
          tmem_ld8(tmem + (uint32_t)mt * 64u + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 8), v);
          if (ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += s_bout[g * 8 + e];
            float4* op = reinterpret_cast<float4*>(p.gb + ((long long)b * p.T + t) * (2 * C) + g * 8);
            op[0] = make_float4(v[0], v[1], v[2], v[3]);
            op[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
      acc_phase ^= 1u;
      tc_fence_before();
      // (the next item's first MMA is gated by bar_ready, which every worker arrives on only after this point)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    __syncwarp();
    tmem_dealloc(tmem, 128);
  }
}

}  // namespace fsvc
