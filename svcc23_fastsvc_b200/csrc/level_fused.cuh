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
//
// What bounds it: with N = C = 24 the tensor core is limited by shared-memory operand reads (a 4 KB A
// tile per 128x32x16 MMA), not by its math rate.  Two measures cut that traffic and hide the rest:
//   * the 3-term split  a_hi*w_hi + a_hi*w_lo + a_lo*w_hi  is issued as TWO MMAs per K chunk:
//     a_hi x [w_hi | w_lo] (N doubled, the epilogue adds the two column halves) and a_lo x w_hi;
//   * the two branches are independent chains, so the MMAs of one branch run while the workers drain
//     the other branch's accumulators (per-branch ready / accumulator mbarriers, 4 accumulators in TMEM).
// Channel padding (C = 24 -> K = 32) is a single shared all-zero strip addressed through the descriptor's
// leading-dimension offset, so the four activation buffers hold only real channels.
#pragma once
#include "conv_tc3.cuh"

namespace fsvc {

constexpr int kLfWorkers = 384;            // 12 warps: (lane quarter q) x (channel group select 0..2)
constexpr int kLfThreads = kLfWorkers + 32;
constexpr int kLfValid = 238;              // output steps per work item
constexpr int kLfRows = 272;               // rows of an activation buffer (time t0-9 .. t0+262)
constexpr int kLfHalo = 9;

// Weights of the fused level: fp32 packed [C_in][K][C_out]  ->  [tap][group g of 8 ci][hi: N rows | lo: N rows][8 ci]
// bf16 (zero padded): a B descriptor over 2N rows sees [w_hi | w_lo], one over N rows sees w_hi.  Written by
// weight_jobs_tc_kernel (tc_forward.cu, job kind 5).

struct LevelFusedArgs {
  const float* sig[2];       // [B][T] raw 1-channel signals: loudness, sine
  const float* c1_w[2];      // fp32 packed [tap][C]
  const float* c1_b[2];
  const float* r1_w[2];      // fp32 [C]
  const float* r1_b[2];
  const __nv_bfloat16* w_c2[2];   // pack_tc_nc format, G = Gp groups, N = N1
  const __nv_bfloat16* w_c4[2];
  const __nv_bfloat16* w_film[2];
  const __nv_bfloat16* w_out;     // merged film_out, pack_tc_nc format, G = 2C/8, N = N2
  const float* b_c2[2];
  const float* b_c4[2];
  const float* b_film[2];
  const float* b_out;        // [2C]
  float* y_dec[2];           // [B][T/dec][C] or nullptr
  float* gb;                 // [B][T][2C]
  int C, T, B, dec;          // B: utterances of THIS launch, starting at utterance b_off
  int b_off;
  int n_tiles;               // ceil(T / kLfValid)
  int Gp;                    // groups of a C->C conv incl. K padding (even)
  int N1;                    // N of the C->C convs (multiple of 16, <= 32)
  int N2;                    // N of film_out (multiple of 16, <= 64)
  float slope;
  // UMMA descriptor low words (start address relative to the dynamic shared-memory base | leading byte offset) of
  // every (layer, branch, M-tile, tap, K chunk) in issue order: a_hi, a_lo, [w_hi | w_lo].  They live in the kernel
  // parameters (constant bank) so the issuing warp reads them straight into uniform registers.
  uint32_t desc[3 * 96];
};
constexpr int kLfMaxDesc = 96;

struct LevelFusedSmem {
  uint32_t off_w_c2[2], off_w_c4[2], off_w_film[2], off_w_out, off_buf[4], off_zero, off_sig, off_par, off_desc, off_bar, total;
  uint32_t w1_bytes, w2_bytes, buf_bytes;
};
__host__ __device__ inline LevelFusedSmem level_fused_smem(int C, int Gp, int N1, int N2) {
  LevelFusedSmem s;
  const uint32_t G = C / 8, G2 = 2 * C / 8;
  s.w1_bytes = 2u * 3u * Gp * N1 * 16u;
  s.w2_bytes = 2u * 3u * G2 * N2 * 16u;
  s.buf_bytes = 2u * G * kLfRows * 16u;  // hi plane | lo plane, real channel groups only
  uint32_t off = 0;
  for (int br = 0; br < 2; ++br) {
    s.off_w_c2[br] = off; off += s.w1_bytes;
    s.off_w_c4[br] = off; off += s.w1_bytes;
    s.off_w_film[br] = off; off += s.w1_bytes;
  }
  s.off_w_out = off; off += s.w2_bytes;
  for (int i = 0; i < 4; ++i) { s.off_buf[i] = off; off += s.buf_bytes; }
  s.off_zero = off; off += kLfRows * 16u;  // the shared K-padding strip (must lie above every buffer)
  s.off_sig = off; off += 2u * kLfRows * 4u;
  s.off_par = off; off += (2u * 9u * 32u + 64u) * 4u;
  off = (off + 15u) & ~15u;
  s.off_desc = off; off += (6u * 2u * 3u * (Gp / 2) + 2u * 3u * (G2 / 2)) * 24u;  // precomputed UMMA descriptors
  off = (off + 15u) & ~15u;
  s.off_bar = off; off += 12 * 8 + 16;
  s.total = off;
  return s;
}

// Descriptor low words of chunk e (issue order), addresses relative to the shared-memory base.  Every (layer, branch,
// M-tile, tap, K chunk) uses the same shared-memory addresses for every work item.
__host__ __device__ inline uint32_t lf_desc_lo(uint32_t addr, uint32_t lbo) {
  return ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
}
__host__ inline int level_fused_fill_desc(LevelFusedArgs* p) {
  const int C = p->C, G = C >> 3, Gp = p->Gp;
  const LevelFusedSmem L = level_fused_smem(C, Gp, p->N1, p->N2);
  const uint32_t strip = kLfRows * 16u, plane = (uint32_t)G * strip, buf_bytes = L.buf_bytes;
  // K = 16 chunks of a C -> C conv: pairs of (tap, 8-channel group) columns.  A column pair need not be adjacent --
  // the descriptor's leading-dimension offset is the distance between its two columns, for A and for B -- so the
  // 3 taps x G groups are packed into ceil(3G / 2) chunks instead of 3 * ceil(G / 2) (C = 24: 5 instead of 6 MMA
  // pairs per tap set; the tensor pipe is what bounds this kernel): inside a tap groups pair up (g, g+1); with G odd
  // the last groups of taps 0 and 1 pair with each other and the last group of tap 2 with the shared zero strip.
  // Every pair's second column lies at a higher address than its first, in the activation buffer and in the weights.
  struct Col { int k, g; };   // g < 0: zero strip
  Col pairs[16][2];
  int np = 0;
  for (int k = 0; k < 3; ++k)
    for (int g = 0; g + 1 < G; g += 2) { pairs[np][0] = {k, g}; pairs[np][1] = {k, g + 1}; ++np; }
  if (G & 1) {
    pairs[np][0] = {0, G - 1}; pairs[np][1] = {1, G - 1}; ++np;
    pairs[np][0] = {2, G - 1}; pairs[np][1] = {2, -1}; ++np;
  }
  const uint32_t n1 = (uint32_t)np, G2 = 2u * (uint32_t)G, n2 = 3u * (G2 / 2);
  const uint32_t n_chain = 6u * 2u * n1, n_all = n_chain + 2u * n2;
  if (n_all > (uint32_t)kLfMaxDesc) return -1;
  const uint32_t buf0 = L.off_buf[0], zero_addr = L.off_zero, w_base = 0;
  uint32_t b_lbo = 0;
  for (uint32_t e = 0; e < n_all; ++e) {
    uint32_t a0, a1h, a1l, b_addr, b_grp;
    if (e < n_chain) {
      const uint32_t lb = e / (2u * n1), r = e - lb * 2u * n1;
      const int layer = (int)(lb >> 1), br = (int)(lb & 1u);
      const int mt = (int)(r / n1), kk = (int)(r - (uint32_t)mt * n1);
      const int s0 = layer == 0 ? -6 : (layer == 1 ? -2 : -1), d = layer == 0 ? 2 : (layer == 1 ? 4 : 1);
      const int src = layer == 1 ? 1 : 0;  // a1 / y live in X (0), a2 in Y (1)
      const uint32_t w_addr = w_base + (layer == 0 ? L.off_w_c2[br] : (layer == 1 ? L.off_w_c4[br] : L.off_w_film[br]));
      const Col c0 = pairs[kk][0], c1 = pairs[kk][1];
      auto rbytes = [&](int k) { return (uint32_t)(s0 + kLfHalo - d + k * d + 128 * mt) * 16u; };
      const uint32_t abase = buf0 + (uint32_t)(2 * br + src) * buf_bytes;
      a0 = abase + (uint32_t)c0.g * strip + rbytes(c0.k);
      // second 8-channel column: another (tap, group) of the same buffer, or the shared zero strip for the K padding
      a1h = c1.g >= 0 ? abase + (uint32_t)c1.g * strip + rbytes(c1.k) : zero_addr + rbytes(c0.k);
      a1l = c1.g >= 0 ? a1h + plane : a1h;
      b_grp = 2u * (uint32_t)p->N1 * 16u;
      b_addr = w_addr + ((uint32_t)c0.k * Gp + c0.g) * b_grp;
      // (against the zero strip any finite weights do: the next block of the packed layout, zero padding itself)
      b_lbo = c1.g >= 0 ? w_addr + ((uint32_t)c1.k * Gp + c1.g) * b_grp - b_addr : b_grp;
    } else {
      const uint32_t r = e - n_chain;
      const int mt = (int)(r / n2), kk = (int)(r - (uint32_t)mt * n2);
      const int k = kk / (int)(G2 / 2);
      const uint32_t kc = (uint32_t)kk - (uint32_t)k * (G2 / 2);
      const uint32_t rbytes = (uint32_t)(kLfHalo + (k - 1) + 128 * mt) * 16u;
      const uint32_t v0 = 2u * kc, v1 = v0 + 1;  // virtual channel groups of [h_lft (Y_0) | h_sine (Y_1)]
      a0 = buf0 + (v0 >= (uint32_t)G ? 3u : 1u) * buf_bytes + (v0 % G) * strip + rbytes;
      a1h = buf0 + (v1 >= (uint32_t)G ? 3u : 1u) * buf_bytes + (v1 % G) * strip + rbytes;
      a1l = a1h + plane;
      b_grp = 2u * (uint32_t)p->N2 * 16u;
      b_addr = w_base + L.off_w_out + ((uint32_t)k * G2 + v0) * b_grp;
      b_lbo = b_grp;
    }
    p->desc[3 * e + 0] = lf_desc_lo(a0, a1h - a0);                    // a_hi
    p->desc[3 * e + 1] = lf_desc_lo(a0 + plane, a1l - (a0 + plane));  // a_lo
    p->desc[3 * e + 2] = lf_desc_lo(b_addr, b_lbo);                   // [w_hi | w_lo]
  }
  return 0;
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
  uint64_t* bar_ready = bars;        // [2] workers -> MMA: branch br's next layer input is complete (count 384)
  uint64_t* bar_acc = bars + 2;      // [2][2] MMA -> workers: accumulator (branch, M-tile) is complete
  uint64_t* bar_w = bars + 6;        // weights landed
  uint64_t* bar_a1 = bars + 7;       // [2] workers -> MMA: branch br's first-conv output of the NEXT item is stored
  uint64_t* bar_fo = bars + 9;       // [2] MMA -> workers: film_out accumulator (M-tile) is complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 12);
  float* s_sig = reinterpret_cast<float*>(smem + L.off_sig);   // [2][kLfRows]: index i <-> time t0 - 10 + i
  float* s_par = reinterpret_cast<float*>(smem + L.off_par);
  float* s_bout = s_par + 2 * kLfParPerBranch * 32;
  const uint32_t strip = kLfRows * 16u, plane = (uint32_t)G * strip;
  const uint32_t buf_off0 = L.off_buf[0], buf_bytes = L.buf_bytes;
  // buffers: X_br = buffer 2*br (a1, then y), Y_br = buffer 2*br+1 (a2, then h)

  if (tid == 0) FSVC_TL(63, 0);
  griddep_launch_dependents();
  if (tid == 0) {
    mbar_init(bar_ready, kLfWorkers);
    mbar_init(bar_ready + 1, kLfWorkers);
    for (int i = 0; i < 4; ++i) mbar_init(bar_acc + i, 1);
    mbar_init(bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_a1 + i, kLfWorkers);
      mbar_init(bar_fo + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(s_tmem, 512);  // 4 chain accumulators (256 columns) + 2 film_out accumulators
  // zero the K-padding strip (never written again) and the buffers (stale rows must stay finite)
  for (uint32_t i = tid; i < (4u * buf_bytes + strip) / 16u; i += kLfThreads)
    reinterpret_cast<uint4*>(smem + buf_off0)[i] = make_uint4(0, 0, 0, 0);
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
    // The whole warp runs the warp-uniform control flow; descriptors come from the kernel parameters (constant bank
    // -> uniform registers) plus the shared-memory base; one elected lane issues the tcgen05 instructions.  (Issued
    // from inside `if (lane == 0)` with the descriptors read from shared memory, each chunk cost 21 instructions
    // -- 3 LDS.64, 6 vector->uniform moves, election -- and the issuing warp was busy for the whole kernel.)
    if (lane == 0) {
      const uint32_t total = 6u * L.w1_bytes + L.w2_bytes;
      mbar_expect_tx(bar_w, total);
      for (int br = 0; br < 2; ++br) {
        bulk_g2s(smem + L.off_w_c2[br], p.w_c2[br], L.w1_bytes, bar_w);
        bulk_g2s(smem + L.off_w_c4[br], p.w_c4[br], L.w1_bytes, bar_w);
        bulk_g2s(smem + L.off_w_film[br], p.w_film[br], L.w1_bytes, bar_w);
      }
      bulk_g2s(smem + L.off_w_out, p.w_out, L.w2_bytes, bar_w);
    }
    __syncwarp();
    mbar_wait2(bar_w, 0);
    {
      const bool leader = elect_one_sync();
      const uint32_t idesc_c = umma_idesc_bf16(128, 2 * p.N1), idesc_h = umma_idesc_bf16(128, p.N1);
      const uint32_t idesc_oc = umma_idesc_bf16(128, 2 * p.N2), idesc_oh = umma_idesc_bf16(128, p.N2);
      const uint32_t base4 = smem_u32(smem) >> 4;
      const uint64_t desc_hi = umma_desc(0, 0, 128) & 0xFFFFFFFF00000000ull;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t n1 = (3u * (uint32_t)G + 1u) / 2u, n2 = 3u * (uint32_t)G;  // chunks per M-tile (level_fused_fill_desc)
      uint32_t ready_phase0 = 0u, ready_phase1 = 0u, a1_phase = 0u;
      // one M-tile: a_hi x [w_hi | w_lo] -> columns [0, 2N), a_lo x w_hi -> [0, N)
      auto issue_mtile = [&](uint32_t d_tmem, uint32_t& di, uint32_t n_chunks, uint32_t idesc_wide, uint32_t idesc_half) {
        uint32_t accum = 0u;
        for (uint32_t i = 0; i < n_chunks; ++i, di += 3u) {
          const uint64_t A_hi = desc_hi | (uint64_t)(p.desc[di] + base4), A_lo = desc_hi | (uint64_t)(p.desc[di + 1u] + base4);
          const uint64_t Bd = desc_hi | (uint64_t)(p.desc[di + 2u] + base4);
          if (leader) {
            umma_bf16(d_tmem, A_hi, Bd, idesc_wide, accum);
            umma_bf16(d_tmem, A_lo, Bd, idesc_half, 1u);
          }
          accum = 1u;
        }
      };
      int tl_it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tl_it) {
        uint32_t di = 0u;
        for (int layer = 0; layer < 3; ++layer) {      // a1 (X) -> a2 (Y) -> y (X) -> h (Y), branches interleaved
          for (int br = 0; br < 2; ++br) {
            if (leader && tl_it == 3) FSVC_TL(63, 1 + (layer * 2 + br) * 3);
            if (layer == 0) {  // the first conv's output was stored during the previous item (own barrier)
              mbar_wait2(bar_a1 + br, a1_phase);
            } else if (br == 0) {
              mbar_wait2(bar_ready, ready_phase0);
              ready_phase0 ^= 1u;
            } else {
              mbar_wait2(bar_ready + 1, ready_phase1);
              ready_phase1 ^= 1u;
            }
            tc_fence_after();
            if (leader && tl_it == 3) FSVC_TL(63, 2 + (layer * 2 + br) * 3);
            for (int mt = 0; mt < 2; ++mt) {
              issue_mtile(tmem_u + (uint32_t)(2 * br + mt) * 64u, di, n1, idesc_c, idesc_h);
              if (leader) umma_commit(bar_acc + 2 * br + mt);
            }
            if (leader && tl_it == 3) FSVC_TL(63, 3 + (layer * 2 + br) * 3);
          }
        }
        // merged film_out over the virtual channel concat [h_lft (Y_0) | h_sine (Y_1)]
        if (leader && tl_it == 3) FSVC_TL(63, 19);
        mbar_wait2(bar_ready, ready_phase0);
        ready_phase0 ^= 1u;
        mbar_wait2(bar_ready + 1, ready_phase1);
        ready_phase1 ^= 1u;
        tc_fence_after();
        if (leader && tl_it == 3) FSVC_TL(63, 20);
        for (int mt = 0; mt < 2; ++mt) {
          issue_mtile(tmem_u + 256u + (uint32_t)mt * 128u, di, n2, idesc_oc, idesc_oh);
          if (leader) umma_commit(bar_fo + mt);
        }
        a1_phase ^= 1u;
        if (leader && tl_it == 3) FSVC_TL(63, 21);
      }
      __syncwarp();
    }
  } else {
    // ================= workers =================
    const int q = warp & 3, gsel = warp >> 2;  // TMEM lane quarter; channel group select (0..2)
    uint32_t acc_phase0 = 0u, acc_phase1 = 0u;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    griddep_wait();  // the outputs of this kernel may still be read by the previous forward's kernels
    // Software pipeline over items.  The first conv of an item (a1 = Conv3_d1(lrelu(x)) on the CUDA cores, stored as
    // lrelu(a1), rows t0-8 .. t0+247) is computed at the END of the previous item, after its last chain layer has been
    // drained (the X buffers are free then) and BEFORE its film_out accumulators are drained: the tensor core goes
    // from this item's film_out straight to the next item's second conv while the workers store gamma | beta.
    // film_out has its own TMEM columns and barriers for that.  The raw signal windows of both branches (index i <->
    // branch i / kLfRows, time t0 - 10 + i % kLfRows) are requested an item before they are staged.  Measured before:
    // 0.8 us signal load + 0.8 us first conv at the start and 1.8 us gamma | beta drain at the end of an 11.4 us item,
    // all with the tensor core idle.
    float sig_pre[2];
    uint32_t fo_phase = 0u;
    auto prefetch_signals = [&](int item) {
      const int bl = item / p.n_tiles, b = p.b_off + bl, t0 = (item - bl * p.n_tiles) * kLfValid;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = tid + k * kLfWorkers;
        const int br = i >= kLfRows, j = i - br * kLfRows;
        const int t = t0 - 10 + j;
        sig_pre[k] = (i < 2 * kLfRows && t >= 0 && t < p.T) ? __ldg(p.sig[br] + (long long)b * p.T + t) : 0.f;
      }
    };
    auto first_conv = [&](int item) {  // stage the item's signals, compute + store a1 of both branches, release the MMAs
      const int bl = item / p.n_tiles, t0 = (item - bl * p.n_tiles) * kLfValid;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (tid + k * kLfWorkers < 2 * kLfRows) s_sig[tid + k * kLfWorkers] = sig_pre[k];
      named_bar_sync(1, kLfWorkers);
      if (item + (int)gridDim.x < n_items) prefetch_signals(item + gridDim.x);
      for (int br = 0; br < 2; ++br) {
        const float* sg = s_sig + br * kLfRows;
        const float* par = s_par + br * kLfParPerBranch * 32;
        uint8_t* dst = smem + buf_off0 + (uint32_t)(2 * br) * buf_bytes;
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
            for (int e = 0; e < 8; e += 2) {  // two channels per packed instruction
              const int c = g * 8 + e;
              float y0 = par[kLfC1B * 32 + c], y1 = par[kLfC1B * 32 + c + 1];
              fma2_acc(y0, y1, par[(kLfC1W + 0) * 32 + c], par[(kLfC1W + 0) * 32 + c + 1], l0, l0);
              fma2_acc(y0, y1, par[(kLfC1W + 1) * 32 + c], par[(kLfC1W + 1) * 32 + c + 1], l1, l1);
              fma2_acc(y0, y1, par[(kLfC1W + 2) * 32 + c], par[(kLfC1W + 2) * 32 + c + 1], l2, l2);
              lrelu2(y0, y1, p.slope);
              v[e] = y0;
              v[e + 1] = y1;
            }
          }
          split_store(dst + (uint32_t)g * strip + (uint32_t)(tau + kLfHalo) * 16u, plane, v);
        }
        fence_proxy_async();
        mbar_arrive(bar_a1 + br);
      }
    };
    if ((int)blockIdx.x < n_items) {
      prefetch_signals(blockIdx.x);
      first_conv(blockIdx.x);
    }
    int tl_it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tl_it) {
      const int bl = item / p.n_tiles, b = p.b_off + bl, t0 = (item - bl * p.n_tiles) * kLfValid;
      if (tid == 0 && tl_it == 3) FSVC_TL(63, 24);
      // ---- three tensor-core layers per branch, branches interleaved:
      //      TMEM -> (hi + lo halves, bias, residual, activation, zero padding) -> bf16 hi|lo -> smem ----
      for (int layer = 0; layer < 3; ++layer) {
        const int s = layer == 0 ? -6 : (layer == 1 ? -2 : -1);
        for (int br = 0; br < 2; ++br) {
          const float* sg = s_sig + br * kLfRows;
          const float* par = s_par + br * kLfParPerBranch * 32;
          uint8_t* dst = smem + buf_off0 + (uint32_t)(2 * br + (layer == 1 ? 0 : 1)) * buf_bytes;
          const float* bias = par + (layer == 0 ? kLfBC2 : (layer == 1 ? kLfBC4 : kLfBFilm)) * 32;
          const uint32_t ph = br == 0 ? acc_phase0 : acc_phase1;
          if (tid == 0 && tl_it == 3) FSVC_TL(63, 28 + (layer * 2 + br) * 3);
          for (int mt = 0; mt < 2; ++mt) {
            mbar_wait2(bar_acc + 2 * br + mt, ph);
            if (tid == 0 && tl_it == 3 && mt == 0) FSVC_TL(63, 29 + (layer * 2 + br) * 3);
            tc_fence_after();
            const int tau = s + 128 * mt + q * 32 + lane, t = t0 + tau;
            const bool in_seq = t >= 0 && t < p.T;
            const uint32_t tbase = tmem + (uint32_t)(2 * br + mt) * 64u + lane_addr;
            for (int g = gsel; g < G; g += 3) {
              float v[8], w[8];
              tmem_ld8(tbase + (uint32_t)(g * 8), v);
              tmem_ld8(tbase + (uint32_t)(p.N1 + g * 8), w);
              if (in_seq) {
#pragma unroll
                for (int e = 0; e < 8; e += 2) {  // v += w + bias, two channels per packed instruction
                  add2(w[e], w[e + 1], bias[g * 8 + e], bias[g * 8 + e + 1]);
                  add2(v[e], v[e + 1], w[e], w[e + 1]);
                }
                if (layer == 1) {  // + Conv1x1(x), then the level output y (kept raw for the FiLM conv)
                  const float x = sg[tau + 10];
#pragma unroll
                  for (int e = 0; e < 8; e += 2) {
                    float r0 = par[kLfR1B * 32 + g * 8 + e], r1 = par[kLfR1B * 32 + g * 8 + e + 1];
                    fma2_acc(r0, r1, par[kLfR1W * 32 + g * 8 + e], par[kLfR1W * 32 + g * 8 + e + 1], x, x);
                    add2(v[e], v[e + 1], r0, r1);
                  }
                  if (p.y_dec[br] && tau >= 0 && tau < kLfValid && t % p.dec == 0) {
                    float4* yp = reinterpret_cast<float4*>(p.y_dec[br] + ntc_row(ntc_tp(p.T / p.dec), C, b, t / p.dec) + g * 256);
                    yp[0] = make_float4(v[0], v[1], v[2], v[3]);
                    yp[32] = make_float4(v[4], v[5], v[6], v[7]);
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 8; e += 2) lrelu2(v[e], v[e + 1], p.slope);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.f;  // zero padding of the next conv's input
              }
              split_store(dst + (uint32_t)g * strip + (uint32_t)(tau + kLfHalo) * 16u, plane, v);
            }
          }
          if (br == 0) acc_phase0 ^= 1u;
          else acc_phase1 ^= 1u;
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(bar_ready + br);
          if (tid == 0 && tl_it == 3) FSVC_TL(63, 30 + (layer * 2 + br) * 3);
        }
      }
      // (the residual of layer 1 was the last reader of this item's signals, the FiLM conv's MMAs the last of X)
      if (item + (int)gridDim.x < n_items) first_conv(item + gridDim.x);
      if (tid == 0 && tl_it == 3) FSVC_TL(63, 46);
      // ---- merged film_out: gamma | beta rows straight to HBM ----
      for (int mt = 0; mt < 2; ++mt) {
        mbar_wait2(bar_fo + mt, fo_phase);
        tc_fence_after();
        const int tau = 128 * mt + q * 32 + lane, t = t0 + tau;
        const bool ok = tau < kLfValid && t < p.T;
        const uint32_t tbase = tmem + 256u + (uint32_t)mt * 128u + lane_addr;
        for (int g = gsel; g < 2 * G; g += 3) {
          float v[8], w[8];
          tmem_ld8(tbase + (uint32_t)(g * 8), v);
          tmem_ld8(tbase + (uint32_t)(p.N2 + g * 8), w);
          if (ok) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              add2(w[e], w[e + 1], s_bout[g * 8 + e], s_bout[g * 8 + e + 1]);
              add2(v[e], v[e + 1], w[e], w[e + 1]);
            }
            float4* op = reinterpret_cast<float4*>(p.gb + ntc_row(ntc_tp(p.T), 2 * C, b, t) + g * 256);
            op[0] = make_float4(v[0], v[1], v[2], v[3]);
            op[32] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
      fo_phase ^= 1u;
      tc_fence_before();
      if (tid == 0 && tl_it == 3) FSVC_TL(63, 48);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) FSVC_TL(63, 40);
  if (warp == 12) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace fsvc
