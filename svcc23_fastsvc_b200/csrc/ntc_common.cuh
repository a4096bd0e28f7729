// Channels-last ("NTC") activation layout, the argument block of one fused conv, and the small companion kernels of
// the tensor-core forward.
//
// Reference semantics implemented by the fused prologue / epilogue (Tc2Args): Conv1d1x3 / Conv2d1x3 / Conv1d1x1
// (layers/upsample.py:76-106, layers/residual_block.py:41-48), Stretch2d / Squeeze2d index maps
// (layers/upsample.py:38-74), _feature_affine + LeakyReLU (fastsvc.py:115-140, 56-75), FastSVCDownsampleNet's first
// conv and 1x1 residual on the raw 1-channel signal (fastsvc.py:164-172, "gen" operands below).
#pragma once
#include "packed_f32.cuh"
#include "tc_prims.cuh"

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
  // conv_last folded into the epilogue (last stage only; one N tile, one sub-tile): out_last[b][o][t] (zeroed before
  // the launch) += sum_c w_last[c][o] * v[c] over this thread's channels; exactly two contributions per sample.
  const float* last_w;  // [C_out][last_co] fp32, or nullptr
  const float* last_b;  // [last_co]
  float* last_out;      // (B, last_co, T_out), time fastest
  int last_co;
  float slope;
  // ---- operand planes: bf16 hi | lo in the shared-memory operand layout, [B][G][Tp][8 channels] -----------------
  // A producer whose consumer's prologue is at most a LeakyReLU writes its output split and activated, one 16-byte
  // chunk per (step, 8-channel group) -- exactly a row of the consumer's K-major A tile -- and the consumer's row
  // loader bulk-copies an item's window per group and plane straight into the A ring: no transform role (conv_tc3
  // MODE 6).  Tp = 128 * tiles + 2 * kPlPad rows per (utterance, group): kPlPad zero rows in front of step 0, zeros
  // from step T to the end (written by the producer), so the conv's zero padding is part of the tensor.
  const uint4* in_pl;    // consumer: hi plane of the input; the lo plane starts in_pl_lo chunks later
  long long in_pl_lo;
  int in_pl_G, in_pl_Tp;  // groups per utterance in the tensor (the utterance stride is G * Tp chunks), rows per group
  uint4* out_pl;         // producer (plain epilogue): hi plane, group 0 of this conv's first output channel
  long long out_pl_lo;
  int out_pl_G, out_pl_Tp;
  int out_pl_lrelu;      // LeakyReLU before the split (the consumer's prologue, applied once here)
  // ---- decimated copy (plain epilogue): steps t with t % out_dec_r == 0 also go to out_dec[b][t / out_dec_r], the
  // next level's input at its own rate -- its first convs then read contiguous rows instead of every r-th one
  float* out_dec;        // NTC [B][out_dec_T][out_dec_ld], or nullptr
  int out_dec_ld, out_dec_T, out_dec_r;
};
constexpr int kPlPad = 8;

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

// LeakyReLU of a pair, slope in (0, 1): max(x, x * slope)
__device__ __forceinline__ void lrelu2(float& x0, float& x1, float slope) {
  float m0, m1;
  mul2(m0, m1, x0, x1, slope);
  x0 = fmaxf(x0, m0);
  x1 = fmaxf(x1, m1);
}
// x * a + c over 8 values, a | c as two float4 each
__device__ __forceinline__ void affine8(float (&v)[8], const float4& a0, const float4& a1, const float4& c0, const float4& c1) {
  fma2(v[0], v[1], a0.x, a0.y, c0.x, c0.y);
  fma2(v[2], v[3], a0.z, a0.w, c0.z, c0.w);
  fma2(v[4], v[5], a1.x, a1.y, c1.x, c1.y);
  fma2(v[6], v[7], a1.z, a1.w, c1.z, c1.w);
}
__device__ __forceinline__ void lrelu8(float (&v)[8], float slope) {
#pragma unroll
  for (int e = 0; e < 4; ++e) lrelu2(v[2 * e], v[2 * e + 1], slope);
}
// bf16 hi|lo split of a pair: hi = bf16(x) (round to nearest even), lo = bf16(x - hi); the subtraction is exact.
// (One conversion, a shift and a mask to widen hi again, one packed subtract, one conversion: 5 instructions per pair;
//  through __nv_bfloat162 the compiler unpacked and repacked the halves with four extra permutes.)
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  float d0, d1;
  sub2(d0, d1, x0, x1, __uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d1), "f"(d0));
}
__device__ __forceinline__ void split_store(uint8_t* dst, uint32_t plane, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_pair(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
  *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst + plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---- small companions -------------------------------------------------------------------

// Merge the per-32-step (mean, M2) partials [B][n_seg][C] of one (b, c) and emit the affine the next conv applies
// on load: a = rstd, c = e - mean*rstd.  InstanceNorm2d: biased variance over the whole time axis, eps inside the sqrt
// (fastsvc.py:76,138).  One warp per (b, c): lane l takes segments l, l+32, ... and accumulates, in double and in a
// fixed order, the moments of the segment means about the FIRST segment's mean (n, sum n*d, sum n*d^2, sum M2) -- no
// division in the loop (FP64 division is what made the Chan-style merge slow) -- then the lanes are added by a
// butterfly whose two operands are always taken in lane order: bitwise deterministic.
// grid = ceil(B*C / 8) blocks of 256 threads.
__device__ __forceinline__ double shfl_xor_d(double v, int o) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return __hiloint2double(hi, lo);
}
struct InFinalizeArgs {
  const float2* stats;
  int n_seg, T, C, BC;
  const float* e;
  float eps;
  float* out_a;
  float* out_c;
};
__global__ void __launch_bounds__(256) in_finalize2_kernel(const InFinalizeArgs p) {
  const int lane = threadIdx.x & 31;
  const int bc = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (bc >= p.BC) return;
  const int b = bc / p.C, c = bc - b * p.C;
  const float2* sp = p.stats + (long long)b * p.n_seg * p.C + c;
  const double piv = (double)__ldcg(sp).x;
  double sw = 0.0, s1 = 0.0, s2 = 0.0, qq = 0.0;
  for (int s0 = lane; s0 < p.n_seg; s0 += 8 * 32) {
    float2 pv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sg = s0 + 32 * u;
      pv[u] = sg < p.n_seg ? __ldcg(sp + (long long)sg * p.C) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sg = s0 + 32 * u;
      const double nb = sg < p.n_seg ? (double)min(32, p.T - sg * 32) : 0.0;
      const double d = (double)pv[u].x - piv;
      sw += nb;
      s1 = fma(nb, d, s1);
      s2 = fma(nb * d, d, s2);
      qq += (double)pv[u].y;
    }
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double w2 = shfl_xor_d(sw, o), a2 = shfl_xor_d(s1, o), b2 = shfl_xor_d(s2, o), q2 = shfl_xor_d(qq, o);
    const bool low = (lane & o) == 0;  // (lower lane) + (upper lane) on both lanes: the same bits everywhere
    sw = low ? sw + w2 : w2 + sw;
    s1 = low ? s1 + a2 : a2 + s1;
    s2 = low ? s2 + b2 : b2 + s2;
    qq = low ? qq + q2 : q2 + qq;
  }
  if (lane == 0) {
    const double dm = s1 / sw;
    const double var = fmax(qq + s2 - s1 * dm, 0.0) / sw;
    const double rstd = 1.0 / sqrt(var + (double)p.eps);
    p.out_a[bc] = (float)rstd;
    p.out_c[bc] = (float)((double)(p.e ? p.e[bc] : 0.f) - (piv + dm) * rstd);
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
