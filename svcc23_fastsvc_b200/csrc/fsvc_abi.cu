// libfsvc.so -- C ABI (include/fsvc.h) and host-side launch plan of the
// B200-native FastSVC generator forward.  sm_100a only; no CPU path.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fsvc.h"
#include "conv_f32.cuh"
#include "conv_tc.cuh"
#include "conv_tc2.cuh"
#include "conv_tc3.cuh"
#include "level_fused.cuh"
#include "excitation.cuh"

namespace fsvc {

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define FSVC_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) return fail(FSVC_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));   \
  } while (0)

struct ConvW {  // packed [C_in][K][C_out] + bias[C_out], device
  float* w = nullptr;
  float* b = nullptr;
  int C_in = 0, C_out = 0, K = 0;
  TcW tc;   // tensor-core copy of the same weights (tc.w == nullptr: conv not eligible)
  TcW tc2;  // copy tiled for the persistent channels-last kernel (conv_tc2.cuh)
  int tc2_resident = 0;
  int ctx_dil = 1, ctx_up = 1;  // how the generator uses this conv (dilation, upsampling factor of its input)
  int ctx_wide = 0;             // conditioning conv: prefer one wide N tile
  const __nv_bfloat16* wnc = nullptr;  // [tap][group][hi|lo rows][8] copy for the fused level kernel
  int nc_G = 0, nc_N = 0;
};

// Tensor-core tiling of a conv: N tiles of <= 128 output channels (multiple of 16), input-channel
// blocks of <= 64 (multiple of 16).  Tiny convs (C_in or C_out < 8) stay on the fp32 kernel.
static bool tc_plan(int C_in, int C_out, int K, TcW* t) {
  if (C_in < 8 || C_out < 8) return false;
  const int n16 = (C_out + 15) / 16 * 16;
  t->n_ntiles = (n16 + 127) / 128;
  t->N_tile = ((n16 + t->n_ntiles - 1) / t->n_ntiles + 15) / 16 * 16;
  t->N_alloc = 32;
  while (t->N_alloc < t->N_tile) t->N_alloc *= 2;
  const int c16 = (C_in + 15) / 16 * 16;
  t->n_blk = (c16 + 63) / 64;
  t->CIB = ((c16 + t->n_blk - 1) / t->n_blk + 15) / 16 * 16;
  t->K = K;
  return true;
}

// Tiling of a conv for the warp-specialised kernel (conv_tc3.cuh): N tiles of <= 128 output channels,
// ci blocks of <= 64 input channels, weights resident in shared memory when they fit next to the rings.
// It depends only on the conv
// (never on the batch size), so results are independent of how utterances are batched.
static bool tc2_plan(int C_in, int C_out, int K, ConvW* cw) {
  TcW* t = &cw->tc2;
  if (C_out % 8 != 0 || C_in % 8 != 0 || C_in < 8) return false;
  const int c16 = (C_in + 15) / 16 * 16;
  const int n16 = (C_out + 15) / 16 * 16;
  // conditioning convs (two branches per launch, or 2C outputs) take N tiles of up to 256 columns so the A window
  // is converted once; stage convs (one problem, few time tiles) keep <= 128 columns for twice the CTAs
  t->n_ntiles = cw->ctx_wide ? (n16 + 255) / 256 : (n16 + 127) / 128;
  t->N_tile = ((n16 + t->n_ntiles - 1) / t->n_ntiles + 15) / 16 * 16;
  t->N_alloc = 32;
  while (t->N_alloc < t->N_tile) t->N_alloc *= 2;
  t->K = K;
  Tc2Args a;
  memset(&a, 0, sizeof(a));
  a.C_in = C_in;
  a.C_out = C_out;
  a.N_tile = t->N_tile;
  a.n_ntiles = t->n_ntiles;
  a.dil = cw->ctx_dil;
  a.up = cw->ctx_up;
  a.down = 1;
  a.res = (const float*)1;  // plan for the widest epilogue (residual + FiLM operands)
  a.gamma = (const float*)1;
  // largest ci block whose A ring is at least double-buffered (resident weights first); else anything that fits
  int best_cib = 0, best_res = 0, fb_cib = 0, fb_res = 0;
  const int nat_blk = (c16 + 63) / 64;
  const int nat_cib = ((c16 + nat_blk - 1) / nat_blk + 15) / 16 * 16;
  for (int cib = nat_cib; cib >= 16 && !best_cib; cib -= 16) {
    for (int resident = 1; resident >= 0 && !best_cib; --resident) {
      a.CIB = cib;
      a.n_blk = (c16 + cib - 1) / cib;
      a.w_resident = resident;
      Tc3Cfg cfg;
      if (!tc3_plan_smem(a, K, &cfg)) continue;
      if (!fb_cib) {
        fb_cib = cib;
        fb_res = resident;
      }
      if (cfg.a_slots >= 2) {
        best_cib = cib;
        best_res = resident;
      }
    }
  }
  if (!best_cib) {
    best_cib = fb_cib;
    best_res = fb_res;
  }
  if (!best_cib) return false;
  t->CIB = best_cib;
  t->n_blk = (c16 + best_cib - 1) / best_cib;
  cw->tc2_resident = best_res;
  return true;
}

struct WeightInfo {
  std::string name;
  int64_t numel;
};

struct StageW {
  ConvW first, up, d3, d9, d27, res;
  float* emb_w = nullptr;  // PyTorch layout [C][S]
  float* emb_b = nullptr;
};
struct LevelW {
  ConvW r1[2], c1[2], c2[2], c4[2], film[2];  // [0] = lft branch, [1] = sine branch
  ConvW film_out;                             // merged: in [h_lft | h_sine] (2C) -> out [gamma | beta] (2C)
};

}  // namespace fsvc

using namespace fsvc;

struct fsvc_handle {
  fsvc_config cfg;
  int n = 0;
  int dscale[FSVC_MAX_STAGES];
  int lvl_c[FSVC_MAX_STAGES];
  int hop = 1;
  std::vector<WeightInfo> winfo;
  float* store = nullptr;
  size_t store_floats = 0;
  __nv_bfloat16* tc_store = nullptr;  // tensor-core (bf16 hi/lo) copies of the conv weights
  size_t tc_elems = 0;
  bool tc2_ok = false;                // every conv of the forward can run on conv_tc2_kernel
  bool l0_fused = false;              // level 0 runs as the fused kernel (level_fused.cuh)
  int num_sms = 148;
  std::vector<ConvW*> convs;          // every conv of the generator (for the tensor-core repack)
  StageW stage[FSVC_MAX_STAGES];
  LevelW level[FSVC_MAX_STAGES];
  ConvW last;
  bool weights_set = false;
  int launches = 0;
  int device = 0;
  // fsvc_forward_host only: the PPG upload runs on a side stream while the conditioning levels (which need only the
  // two signals) compute; the forward waits for `ppg_ready` right before it first reads the PPG tensor
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_ppg = nullptr;
  cudaEvent_t ppg_ready = nullptr;  // set for the duration of one fsvc_forward_host call
};

namespace fsvc {

// ---------------------------------------------------------------------------
// workspace bump allocator (256-byte aligned)
// ---------------------------------------------------------------------------
struct Arena {
  char* base;
  size_t off = 0, cap;
  Arena(void* p, size_t c) : base((char*)p), cap(c) {}
  template <typename T>
  T* get(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    T* p = (T*)(base ? base + off : nullptr);
    off += bytes;
    return p;
  }
  bool ok() const { return off <= cap; }
};

struct Profiler {  // per-launch CUDA-event timing for fsvc_forward_profile (never active in fsvc_forward)
  std::vector<cudaEvent_t> ev;
  std::vector<fsvc_kernel_record> rec;
  void mark(cudaStream_t s) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    ev.push_back(e);
  }
};

struct Ctx {
  cudaStream_t stream;
  int B;
  float slope, eps;
  int launches = 0;
  int err = 0;
  Profiler* prof = nullptr;
  const char* label = "";
  bool tc = false;  // use the tcgen05 kernels where a conv is eligible
  // profiling bookkeeping: called right after a kernel launch
  void launched(const char* kind, double flops, double bytes) {
    launches++;
    if (!prof) return;
    fsvc_kernel_record r;
    memset(&r, 0, sizeof(r));
    snprintf(r.label, sizeof(r.label), "%s%s%s", label, label[0] ? "." : "", kind);
    r.flops = flops;
    r.bytes = bytes;
    prof->rec.push_back(r);
    prof->mark(stream);
  }
};

// ---------------------------------------------------------------------------
// fp32 conv dispatch
// ---------------------------------------------------------------------------
template <int RC, int RT, int K>
static void launch_conv_t(const ConvArgs& a, int B, cudaStream_t s) {
  constexpr int CO_T = kConvWarps * RC, T_T = 32 * RT;
  const int halo = (K / 2) * a.dil;
  const size_t smem = (size_t)(kCiTile * (T_T + 2 * halo) + kCiTile * K * CO_T) * sizeof(float);
  dim3 grid((a.T_out + T_T - 1) / T_T, (a.C_out + CO_T - 1) / CO_T, B);
  conv1d_f32_kernel<RC, RT, K><<<grid, kConvThreads, smem, s>>>(a);
}

static int conv_rt(int T_out) { return T_out >= 1024 ? 8 : 4; }
static int conv_tile_len(int T_out) { return 32 * conv_rt(T_out); }

static int conv_rc(int C_out) {
  if (C_out <= 8) return 1;
  if (C_out <= 16) return 2;
  if (C_out <= 24) return 3;
  if (C_out <= 32) return 4;
  if (C_out % 48 == 0) return 6;
  if (C_out % 32 == 0) return 4;
  return 6;
}

static bool use_tc(const Ctx& c, const ConvW& w) { return c.tc && w.tc.w != nullptr; }

static void launch_conv_tc(Ctx& c, const ConvArgs& a, const TcW& t) {
  TcArgs p;
  p.c = a;
  p.w = t.w;
  p.CIB = t.CIB;
  p.n_blk = t.n_blk;
  p.N_tile = t.N_tile;
  p.N_alloc = t.N_alloc;
  const size_t smem = tc_smem_bytes(t.K, a.dil, t.CIB, t.n_blk, t.N_tile);
  dim3 grid((a.T_out + kTcM - 1) / kTcM, t.n_ntiles, c.B);
  if (t.K == 3) conv1d_tc_kernel<3><<<grid, kTcThreads, smem, c.stream>>>(p);
  else conv1d_tc_kernel<1><<<grid, kTcThreads, smem, c.stream>>>(p);
}

static void launch_conv(Ctx& c, const ConvArgs& a, int K, const char* name = "conv") {
  const ConvW* cw = (const ConvW*)a.host_w;
  if (cw && use_tc(c, *cw)) {
    launch_conv_tc(c, a, cw->tc);
  } else {
  const int rc = conv_rc(a.C_out), rt = conv_rt(a.T_out);
#define FSVC_CASE(RC_, RT_)                                                  \
  if (rc == RC_ && rt == RT_) {                                              \
    if (K == 3) launch_conv_t<RC_, RT_, 3>(a, c.B, c.stream);                \
    else launch_conv_t<RC_, RT_, 1>(a, c.B, c.stream);                       \
  }
  FSVC_CASE(1, 4) FSVC_CASE(1, 8) FSVC_CASE(2, 4) FSVC_CASE(2, 8) FSVC_CASE(3, 4) FSVC_CASE(3, 8)
  FSVC_CASE(4, 4) FSVC_CASE(4, 8) FSVC_CASE(6, 4) FSVC_CASE(6, 8)
#undef FSVC_CASE
  }
  // algorithmic work of this launch: 2*Cin*Cout*K*T flops; every operand tensor touched once
  const double BT = (double)c.B * a.T_out;
  const double flops = 2.0 * a.C_in * a.C_out * K * BT;
  double elems = (double)c.B * a.C_in * ((double)a.T_out / a.up);
  elems += BT * a.C_out * ((a.out ? 1 : 0) + (a.raw ? 1 : 0) + (a.res ? 1 : 0) + (a.gamma ? 2 : 0));
  c.launched(name, flops, 4.0 * (elems + (double)a.C_in * a.C_out * K));
}

// Convenience builder: dense (B, C, T) tensors.
static ConvArgs conv_args(const Ctx& c, const ConvW& w, const float* in, int T_in_stored, int T_out, int dil,
                          float* out) {
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in;
  a.in_cs = T_in_stored;
  a.in_bs = (long long)w.C_in * T_in_stored;
  a.C_in = w.C_in;
  a.up = 1;
  a.down = 1;
  a.w = w.w;
  a.bias = w.b;
  a.dil = dil;
  a.C_out = w.C_out;
  a.T_out = T_out;
  a.out = out;
  a.out_cs = T_out;
  a.out_bs = (long long)w.C_out * T_out;
  a.slope = c.slope;
  a.host_w = &w;
  return a;
}

// ---------------------------------------------------------------------------
// blocks (fp32 path)
// ---------------------------------------------------------------------------
// FastSVCDownsampleNet.forward (fastsvc.py:180-193): in (B,Cin,T_in) -> y (B,C,T_in/scale).
static void run_downsample(Ctx& c, const ConvW& r1, const ConvW& c1, const ConvW& c2, const ConvW& c4,
                           const float* in, int T_in, int scale, float* tmp_r, float* tmp_a, float* tmp_b,
                           float* y) {
  const int T = T_in / scale;
  // r = Squeeze(Conv1x1(x)) == Conv1x1(Squeeze(x)) (pointwise conv commutes with decimation)
  ConvArgs a = conv_args(c, r1, in, T_in, T, 1, tmp_r);
  a.down = scale;
  launch_conv(c, a, 1, "down_r1x1");
  a = conv_args(c, c1, in, T_in, T, 1, tmp_a);
  a.down = scale;
  a.pre_lrelu = 1;
  launch_conv(c, a, 3, "down_d1");
  a = conv_args(c, c2, tmp_a, T, T, 2, tmp_b);
  a.pre_lrelu = 1;
  launch_conv(c, a, 3, "down_d2");
  a = conv_args(c, c4, tmp_b, T, T, 4, y);
  a.pre_lrelu = 1;
  a.res = tmp_r;
  a.res_cs = T;
  a.res_bs = (long long)c4.C_out * T;
  launch_conv(c, a, 3, "down_d4");
}

// One FastSVCUpsampleNet.forward (fastsvc.py:80-140).  gamma/beta are the summed
// FiLM tensors (B,C,T) with batch stride gb_bs.  spk_e: (B,C) projected speaker
// embedding or nullptr (=> no InstanceNorm, fastsvc.py:134).
struct StageBufs {
  float *h0, *xr, *t1, *x_, *t2, *t3;
  float2* stats;
  float *pa, *pc;  // [B][C] affine of the pending InstanceNorm
};

static void run_stage(Ctx& c, const StageW& w, const float* x, int T_in, int r, const float* gamma,
                      const float* beta, long long gb_bs, const float* spk_e, const StageBufs& sb, float* out) {
  const int C = w.first.C_out, T = T_in * r;
  const bool norm = spk_e != nullptr;
  const int tile_len = use_tc(c, w.d3) ? 32 : conv_tile_len(T), n_tiles = (T + tile_len - 1) / tile_len;
  auto film = [&](ConvArgs& a) {
    a.gamma = gamma;
    a.beta = beta;
    a.gb_bs = gb_bs;
    a.gb_cs = T;
    if (norm) {
      a.stats = sb.stats;
      a.n_tiles = n_tiles;
    }
  };
  auto finalize = [&]() {
    if (!norm) return;
    const int BC = c.B * C;
    in_finalize_kernel<<<(BC + 127) / 128, 128, 0, c.stream>>>(sb.stats, n_tiles, tile_len, T, BC, spk_e, c.eps,
                                                               sb.pa, sb.pc);
    c.launched("in_finalize", 0.0, 8.0 * BC * n_tiles);
  };
  auto pre = [&](ConvArgs& a) {
    if (norm) {
      a.pre_a = sb.pa;
      a.pre_c = sb.pc;
    }
    a.pre_lrelu = 1;
  };
  // h0 = conv_first(x)                                                     fastsvc.py:93
  ConvArgs a = conv_args(c, w.first, x, T_in, T_in, 1, sb.h0);
  launch_conv(c, a, 3, "conv_first");
  // xr = residual_block(h0) = Conv3(repeat_r(h0))                         :94
  a = conv_args(c, w.res, sb.h0, T_in, T, 1, sb.xr);
  a.up = r;
  launch_conv(c, a, 3, "residual");
  // t1 = gamma * lrelu(Conv3(repeat_r(lrelu(h0)))) + beta                 :97-98 (+ IN stats)
  a = conv_args(c, w.up, sb.h0, T_in, T, 1, sb.t1);
  a.up = r;
  a.pre_lrelu = 1;
  a.post_lrelu = 1;
  film(a);
  launch_conv(c, a, 3, "up_film");
  finalize();
  // x_ = Conv3_d3(lrelu(IN(t1)+e)) + xr ; t2 = gamma*x_ + beta            :99-105
  a = conv_args(c, w.d3, sb.t1, T, T, 3, sb.t2);
  pre(a);
  a.res = sb.xr;
  a.res_cs = T;
  a.res_bs = (long long)C * T;
  a.raw = sb.x_;
  a.raw_cs = T;
  a.raw_bs = (long long)C * T;
  film(a);
  launch_conv(c, a, 3, "d3_film");
  finalize();
  // t3 = gamma * Conv3_d9(lrelu(IN(t2)+e)) + beta                         :106-107
  a = conv_args(c, w.d9, sb.t2, T, T, 9, sb.t3);
  pre(a);
  film(a);
  launch_conv(c, a, 3, "d9_film");
  finalize();
  // out = Conv3_d27(lrelu(IN(t3)+e)) + x_                                 :108-111
  a = conv_args(c, w.d27, sb.t3, T, T, 27, out);
  pre(a);
  a.res = sb.x_;
  a.res_cs = T;
  a.res_bs = (long long)C * T;
  launch_conv(c, a, 3, "d27_skip");
}

static StageBufs alloc_stage_bufs(Arena& ar, int B, int C, int T_in, int T) {
  StageBufs sb;
  const size_t n = (size_t)B * C * T;
  sb.h0 = ar.get<float>((size_t)B * C * T_in);
  sb.xr = ar.get<float>(n);
  sb.t1 = ar.get<float>(n);
  sb.x_ = ar.get<float>(n);
  sb.t2 = ar.get<float>(n);
  sb.t3 = sb.t1;  // t1 is dead once t2 exists
  const int n_tiles = (T + 31) / 32;  // finest statistics granularity of any kernel family
  sb.stats = ar.get<float2>((size_t)B * C * n_tiles);
  sb.pa = ar.get<float>((size_t)B * C);
  sb.pc = ar.get<float>((size_t)B * C);
  return sb;
}

// ---------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------
static void repack(cudaStream_t s, const float* src, int C_out, int C_in, int K, float* dst, int dst_cout,
                   int ci_off, int co_off) {
  const int n = C_out * C_in * K;
  repack_weight_kernel<<<(n + 255) / 256, 256, 0, s>>>(src, C_out, C_in, K, dst, dst_cout, ci_off, co_off);
}
static void bias_sum(cudaStream_t s, const float* a, const float* b, int n, float* dst, int off) {
  bias_sum_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, n, dst, off);
}

static void pack_tc(cudaStream_t s, const ConvW& cw, const TcW& t) {
  const size_t total = t.elems() / 2;
  const int blocks = (int)((total + 255) / 256 < 1024 ? (total + 255) / 256 : 1024);
  pack_tc_weights_kernel<<<blocks, 256, 0, s>>>(cw.w, cw.C_in, cw.C_out, cw.K, t.CIB, t.n_blk, t.N_tile, t.n_ntiles,
                                                (__nv_bfloat16*)t.w);
}

static int tc_setup_kernels() {
  const int max_smem = 227 * 1024;
  FSVC_CUDA(cudaFuncSetAttribute(conv1d_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  FSVC_CUDA(cudaFuncSetAttribute(conv1d_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
#define FSVC_ATTR(K_, NH_, SM_)                                                                                         \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, SM_, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, SM_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, SM_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, SM_, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, SM_, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, SM_, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem))
  FSVC_ATTR(3, 3, false);
  FSVC_ATTR(3, 0, false);
  FSVC_ATTR(1, 3, false);
  FSVC_ATTR(1, 0, false);
#undef FSVC_ATTR
  FSVC_CUDA(cudaFuncSetAttribute(level0_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  return FSVC_OK;
}

static bool mode_uses_tc(int mode) { return mode == FSVC_MODE_TC_BF16X3 || mode == FSVC_MODE_AUTO; }

struct WS {  // full-forward workspace layout
  float* y[2][FSVC_MAX_STAGES];  // conditioning level outputs per branch
  float* H[FSVC_MAX_STAGES];     // [B][2C][T_l] lrelu(film.conv(y)) of both branches
  float* GB[FSVC_MAX_STAGES];    // [B][2C][T_l] gamma | beta (summed over branches)
  float *tmp_r, *tmp_a, *tmp_b;
  float* e[FSVC_MAX_STAGES];     // projected speaker embedding per stage [B][C]
  StageBufs sb[FSVC_MAX_STAGES];
  float* xs[FSVC_MAX_STAGES];    // stage outputs
};

static size_t layout_ws(const fsvc_handle* h, int B, int frames, void* base, size_t cap, WS* ws) {
  Arena ar(base, cap);
  const int n = h->n;
  const int T = frames * h->hop;
  int T_l = T;
  size_t max_lvl = 0;
  for (int l = 0; l < n; ++l) {
    T_l /= h->dscale[l];
    const int C = h->lvl_c[l];
    const size_t ne = (size_t)B * C * T_l;
    max_lvl = ne > max_lvl ? ne : max_lvl;
    for (int br = 0; br < 2; ++br) ws->y[br][l] = ar.get<float>(ne);
    ws->H[l] = ar.get<float>(2 * ne);
    ws->GB[l] = ar.get<float>(2 * ne);
  }
  ws->tmp_r = ar.get<float>(max_lvl);
  ws->tmp_a = ar.get<float>(max_lvl);
  ws->tmp_b = ar.get<float>(max_lvl);
  int T_in = frames;
  for (int i = 0; i < n; ++i) {
    const int C = h->cfg.mid_channels[i], r = h->cfg.upsampling_scales[i];
    ws->e[i] = ar.get<float>((size_t)B * C);
    ws->sb[i] = alloc_stage_bufs(ar, B, C, T_in, T_in * r);
    ws->xs[i] = ar.get<float>((size_t)B * C * T_in * r);
    T_in *= r;
  }
  return ar.off;
}

static const char* const stage_label[FSVC_MAX_STAGES] = {"s0", "s1", "s2", "s3", "s4", "s5", "s6", "s7"};
static const char* const lvl_label[FSVC_MAX_STAGES] = {"l0", "l1", "l2", "l3", "l4", "l5", "l6", "l7"};
static const char* const lvl_lft_label[FSVC_MAX_STAGES] = {"l0.lft", "l1.lft", "l2.lft", "l3.lft",
                                                           "l4.lft", "l5.lft", "l6.lft", "l7.lft"};
static const char* const lvl_sine_label[FSVC_MAX_STAGES] = {"l0.sine", "l1.sine", "l2.sine", "l3.sine",
                                                            "l4.sine", "l5.sine", "l6.sine", "l7.sine"};

static int forward_fp32(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                        float* out, int B, int frames, void* workspace, size_t ws_bytes, cudaStream_t stream,
                        int mode, Profiler* prof = nullptr) {
  WS ws;
  const size_t need = layout_ws(h, B, frames, workspace, ws_bytes, &ws);
  if (need > ws_bytes) return fail(FSVC_E_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  Ctx c;
  c.stream = stream;
  c.B = B;
  c.slope = h->cfg.lrelu_slope;
  c.eps = h->cfg.in_eps;
  c.prof = prof;
  c.tc = mode_uses_tc(mode);
  if (h->ppg_ready) cudaStreamWaitEvent(stream, h->ppg_ready, 0);  // fsvc_forward_host: PPG upload on the side stream
  if (prof) prof->mark(stream);
  const int n = h->n;
  const int T = frames * h->hop;

  // conditioning chains, computed once (the reference recomputes them per stage, fastsvc.py:322-326)
  int T_prev = T, T_l = T;
  for (int l = 0; l < n; ++l) {
    T_l = T_prev / h->dscale[l];
    const LevelW& lw = h->level[l];
    const int C = h->lvl_c[l];
    for (int br = 0; br < 2; ++br) {
      c.label = br == 0 ? lvl_lft_label[l] : lvl_sine_label[l];
      const float* src = l == 0 ? (br == 0 ? lft : sine) : ws.y[br][l - 1];
      run_downsample(c, lw.r1[br], lw.c1[br], lw.c2[br], lw.c4[br], src, T_prev, h->dscale[l], ws.tmp_r, ws.tmp_a,
                     ws.tmp_b, ws.y[br][l]);
      // h = lrelu(film.conv(y)) into the branch's half of H                fastsvc.py:229
      ConvArgs a = conv_args(c, lw.film[br], ws.y[br][l], T_l, T_l, 1, ws.H[l] + (size_t)br * C * T_l);
      a.out_bs = 2LL * C * T_l;
      a.post_lrelu = 1;
      launch_conv(c, a, 3, "film_conv");
    }
    // [gamma | beta] = merged (conv_scale, conv_shift) of both branches    fastsvc.py:230-231, 129-130
    ConvArgs a = conv_args(c, lw.film_out, ws.H[l], T_l, T_l, 1, ws.GB[l]);
    c.label = lvl_label[l];
    launch_conv(c, a, 3, "film_out");
    T_prev = T_l;
  }

  // upsampling stages
  const float* x = ppg;
  int T_in = frames;
  for (int i = 0; i < n; ++i) {
    const int C = h->cfg.mid_channels[i], r = h->cfg.upsampling_scales[i];
    const int l = n - 1 - i, T_s = T_in * r;
    const float* e = nullptr;
    c.label = stage_label[i];
    if (spk) {
      spk_project_kernel<<<B, 256, 0, stream>>>(spk, h->cfg.spk_emb_size, h->stage[i].emb_w, h->stage[i].emb_b, C,
                                                ws.e[i]);
      c.launched("spk_project", 2.0 * B * C * h->cfg.spk_emb_size, 4.0 * C * h->cfg.spk_emb_size);
      e = ws.e[i];
    }
    run_stage(c, h->stage[i], x, T_in, r, ws.GB[l], ws.GB[l] + (size_t)C * T_s, 2LL * C * T_s, e, ws.sb[i],
              ws.xs[i]);
    x = ws.xs[i];
    T_in = T_s;
  }
  // conv_last (1x1)                                                         fastsvc.py:330
  ConvArgs a = conv_args(c, h->last, x, T, T, 1, out);
  c.label = "";
  launch_conv(c, a, 1, "conv_last");
  h->launches = c.launches;
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}


// ===========================================================================
// tensor-core forward over channels-last activations (conv_tc2.cuh)
// ===========================================================================
struct WS2 {
  float* y[2][FSVC_MAX_STAGES];  // [B][T_l][C_l] conditioning level outputs per branch
  float* H[FSVC_MAX_STAGES];     // [B][T_l][2C] lrelu(film.conv(y)) of both branches, side by side
  float* GB[FSVC_MAX_STAGES];    // [B][T_l][2C] gamma | beta (summed over branches)
  float *tr[2], *ta[2], *tb[2];  // per-branch temporaries of a level chain
  float* e[FSVC_MAX_STAGES];     // [B][C] projected speaker embedding per stage
  float *h0[FSVC_MAX_STAGES], *xr[FSVC_MAX_STAGES], *t1[FSVC_MAX_STAGES], *x_[FSVC_MAX_STAGES],
      *t2[FSVC_MAX_STAGES], *xs[FSVC_MAX_STAGES];
  float2* stats[2];              // [B][n_seg][C], ping-pong: a conv reads its producer's while writing its own
  float *pa, *pc;                // [B][C]
  float* xin;                    // [B][frames][in_channels] channels-last copy of the PPG input
  float* ydec[2];                // [B][T/s][C0] level-0 output decimated for level 1 (fused level kernel)
};

static size_t layout_ws2(const fsvc_handle* h, int B, int frames, void* base, size_t cap, WS2* ws) {
  Arena ar(base, cap);
  const int n = h->n;
  const int T = frames * h->hop;
  int T_l = T;
  size_t max_lvl = 0, max_stat = 0, max_bc = 0;
  // every activation is blocked channels-last (conv_tc2.cuh: ntc_row): utterances are padded to 32-step blocks
  for (int l = 0; l < n; ++l) {
    T_l /= h->dscale[l];
    const size_t ne = (size_t)B * h->lvl_c[l] * ntc_tp(T_l);
    max_lvl = ne > max_lvl ? ne : max_lvl;
    for (int br = 0; br < 2; ++br) ws->y[br][l] = ar.get<float>(ne);
    ws->H[l] = ar.get<float>(2 * ne);
    ws->GB[l] = ar.get<float>(2 * ne);
  }
  for (int br = 0; br < 2; ++br) {
    ws->tr[br] = ar.get<float>(max_lvl);
    ws->ta[br] = ar.get<float>(max_lvl);
    ws->tb[br] = ar.get<float>(max_lvl);
  }
  int T_in = frames;
  for (int i = 0; i < n; ++i) {
    const int C = h->cfg.mid_channels[i], r = h->cfg.upsampling_scales[i];
    const size_t ne = (size_t)B * C * ntc_tp(T_in * r);
    ws->e[i] = ar.get<float>((size_t)B * C);
    ws->h0[i] = ar.get<float>((size_t)B * C * ntc_tp(T_in));
    ws->xr[i] = ar.get<float>(ne);
    ws->t1[i] = ar.get<float>(ne);
    ws->x_[i] = ar.get<float>(ne);
    ws->t2[i] = ar.get<float>(ne);
    ws->xs[i] = ar.get<float>(ne);
    T_in *= r;
    const size_t st = (size_t)B * C * ((T_in + 31) / 32);
    max_stat = st > max_stat ? st : max_stat;
    max_bc = (size_t)B * C > max_bc ? (size_t)B * C : max_bc;
  }
  for (int i = 0; i < 2; ++i) ws->stats[i] = ar.get<float2>(max_stat);
  ws->pa = ar.get<float>(max_bc);
  ws->pc = ar.get<float>(max_bc);
  ws->xin = ar.get<float>((size_t)B * ntc_tp(frames) * h->cfg.in_channels);
  for (int br = 0; br < 2; ++br) ws->ydec[br] = ar.get<float>((size_t)B * ntc_tp(T) * h->lvl_c[0]);
  return ar.off;
}

// Fill the tiling / weight half of the arguments of one conv.
static Tc2Args tc2_args(const Ctx& c, const ConvW& w, const float* in, int in_ld, int T_in, int T_out, int dil,
                        float* out, int out_ld) {
  Tc2Args a;
  memset(&a, 0, sizeof(a));
  a.in = in;
  a.in_ld = in_ld;
  a.T_in = T_in;
  a.C_in = w.C_in;
  a.up = 1;
  a.down = 1;
  a.w = w.tc2.w;
  a.CIB = w.tc2.CIB;
  a.n_blk = w.tc2.n_blk;
  a.N_tile = w.tc2.N_tile;
  a.n_ntiles = w.tc2.n_ntiles;
  a.w_resident = w.tc2_resident;
  a.bias = w.b;
  a.dil = dil;
  a.C_out = w.C_out;
  a.T_out = T_out;
  a.out = out;
  a.out_ld = out_ld;
  a.slope = c.slope;
  return a;
}

// Launch with programmatic stream serialization: the kernel may begin (barrier / TMEM setup, weight streaming)
// while its predecessor drains; it orders itself with griddepcontrol.wait before touching activations.
template <typename Kern, typename Arg>
static void launch_pdl(Kern kern, dim3 grid, int block, size_t smem, cudaStream_t stream, const Arg& arg) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, arg);
}

// Launch 1 or 2 problems of identical shape and flags (the two conditioning branches) as one persistent grid
// of the warp-specialised kernel; problem 1 is expressed as pointer deltas against problem 0.
static void launch_tc2(Ctx& c, const fsvc_handle* h, int K, const Tc2Args* p, int n_prob, const char* name) {
  Tc3Launch L;
  memset(&L, 0, sizeof(L));
  L.a = p[0];
  if (n_prob == 2) {
    const Tc2Args &x = p[0], &y = p[1];
    L.d_in = y.in - x.in;
    L.d_w = y.w - x.w;
    L.d_bias = y.bias - x.bias;
    L.d_gen_w = x.gen_w ? y.gen_w - x.gen_w : 0;
    L.d_gen_b = x.gen_w ? y.gen_b - x.gen_b : 0;
    L.d_res = x.res ? y.res - x.res : 0;
    L.d_gres_w = x.gres_w ? y.gres_w - x.gres_w : 0;
    L.d_gres_b = x.gres_w ? y.gres_b - x.gres_b : 0;
    L.d_gres_x = x.gres_w ? y.gres_x - x.gres_x : 0;
    L.d_raw = x.raw ? y.raw - x.raw : 0;
    L.d_out = x.out ? y.out - x.out : 0;
    // everything that is not a per-problem pointer must agree
    if ((x.up > 1 && x.down > 1) || x.pre_lrelu != y.pre_lrelu || x.post_lrelu != y.post_lrelu || x.gamma != y.gamma || x.stats != y.stats ||
        x.pre_a != y.pre_a || x.up != y.up || x.down != y.down || x.dil != y.dil || x.C_in != y.C_in ||
        x.C_out != y.C_out || x.T_out != y.T_out || x.T_in != y.T_in || x.in_ld != y.in_ld ||
        x.out_ld != y.out_ld || x.res_ld != y.res_ld || (x.res == nullptr) != (y.res == nullptr) ||
        (x.gen_w == nullptr) != (y.gen_w == nullptr) || (x.gres_w == nullptr) != (y.gres_w == nullptr)) {
      c.err = 2;
      return;
    }
  }
  L.tl_slot = c.launches;
  Tc3Cfg& cfg = L.c;
  cfg.n_prob = n_prob;
  cfg.B = c.B;
  cfg.m_tiles = (p[0].T_out + kTc2M - 1) / kTc2M;
  const int groups = n_prob * p[0].n_ntiles;
  const int items = c.B * cfg.m_tiles;
  // two half-size CTAs per SM when the conv allows it and there is enough work to keep both pipelines busy
  // (measured: two 224-thread CTAs per SM are no faster than one 512-thread CTA on any layer -- kept as a template
  //  option of the kernel, not instantiated)
  const bool small = false;
  // transform variant: 3 / 4 = lean path (direct rows, one ci block, a warp's <= 4 / <= 6 tasks of an item in one chunk)
  int mode = p[0].gen_w ? 1 : (p[0].up > 1 ? 2 : 0);
  if (mode == 0 && p[0].down == 1 && p[0].n_blk == 1 && !getenv("FSVC_NO_LEAN")) {
    const int ntask = (p[0].CIB / 8) * ((kTc2M + 2 * (K / 2) * p[0].dil + 31) / 32);
    const int rounds = (ntask + 5) / 6;
    if (rounds <= 4 && tc3_plan_smem(p[0], K, &cfg, false, 4)) mode = 3;
    else if (rounds <= 6 && tc3_plan_smem(p[0], K, &cfg, false, 6) && cfg.a_slots >= 2) mode = 4;
  }
  if (mode == 2 && p[0].down == 1 && p[0].up <= 8 && p[0].n_blk == 1 && !getenv("FSVC_NO_LEAN")) {
    const int Wd = kTc2M + 2 * (K / 2) * p[0].dil;
    const int ntask = (p[0].CIB / 8) * (((Wd + p[0].up - 1) / p[0].up + 1 + 31) / 32);
    if ((ntask + 5) / 6 <= 4 && tc3_plan_smem(p[0], K, &cfg, false, 4)) mode = 5;
  }
  if (mode < 3 && !tc3_plan_smem(p[0], K, &cfg, false)) {
    c.err = 1;
    return;
  }
  int per_group = (small ? 2 : 1) * h->num_sms / groups;
  per_group = per_group < 1 ? 1 : per_group;
  per_group = per_group > items ? items : per_group;
  const dim3 grid(per_group * groups);
  if (c.prof && getenv("FSVC_DEBUG_PLAN"))
    fprintf(stderr, "plan %-4s %-12s Cin=%d Cout=%d K=%d dil=%d up=%d down=%d CIB=%d n_blk=%d N_tile=%d n_ntiles=%d resident=%d "
                    "a_slots=%d stg_depth=%d smem=%u grid=%u items=%d\n",
            c.label, name, p[0].C_in, p[0].C_out, K, p[0].dil, p[0].up, p[0].down, p[0].CIB, p[0].n_blk, p[0].N_tile,
            p[0].n_ntiles, p[0].w_resident, cfg.a_slots, cfg.stg_depth, cfg.total, grid.x, items);
  // compile-time epilogue width (12 channels per thread) when every sub-tile of every N tile is full
  const int per_thread = cfg.nsub / (small ? 1 : 2);
  const bool nh3 = per_thread == 12 && p[0].C_out % cfg.nsub == 0 && (p[0].n_ntiles == 1 || p[0].N_tile % cfg.nsub == 0);
  const int threads = small ? kTc3ThreadsSmall : kTc3Threads;
#define FSVC_TC3(K_, NH_, SM_)                                                                               \
  do {                                                                                                      \
    if (mode == 1) launch_pdl(conv_tc3_kernel<K_, NH_, SM_, 1>, grid, threads, cfg.total, c.stream, L);      \
    else if (mode == 2) launch_pdl(conv_tc3_kernel<K_, NH_, SM_, 2>, grid, threads, cfg.total, c.stream, L); \
    else if (mode == 3) launch_pdl(conv_tc3_kernel<K_, NH_, SM_, 3>, grid, threads, cfg.total, c.stream, L); \
    else if (mode == 4) launch_pdl(conv_tc3_kernel<K_, NH_, SM_, 4>, grid, threads, cfg.total, c.stream, L); \
    else if (mode == 5) launch_pdl(conv_tc3_kernel<K_, NH_, SM_, 5>, grid, threads, cfg.total, c.stream, L); \
    else launch_pdl(conv_tc3_kernel<K_, NH_, SM_, 0>, grid, threads, cfg.total, c.stream, L);                \
  } while (0)
  if (K == 3) {
    if (nh3) FSVC_TC3(3, 3, false); else FSVC_TC3(3, 0, false);
  } else {
    if (nh3) FSVC_TC3(1, 3, false); else FSVC_TC3(1, 0, false);
  }
#undef FSVC_TC3
  double flops = 0, elems = 0;
  for (int i = 0; i < n_prob; ++i) {
    const Tc2Args& a = p[i];
    const double BT = (double)c.B * a.T_out;
    flops += 2.0 * a.C_in * a.C_out * K * BT;
    elems += a.gen_w ? BT : (double)c.B * a.C_in * ((double)a.T_out / a.up);
    elems += BT * a.C_out * ((a.out ? 1 : 0) + (a.raw ? 1 : 0) + (a.res ? 1 : 0) + (a.gamma ? 2 : 0));
    elems += (double)a.C_in * a.C_out * K;
  }
  c.launched(name, flops, 4.0 * elems);
}

static int forward_tc2(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                       float* out, int B, int frames, void* workspace, size_t ws_bytes, cudaStream_t stream,
                       Profiler* prof = nullptr) {
  WS2 ws;
  const size_t need = layout_ws2(h, B, frames, workspace, ws_bytes, &ws);
  if (need > ws_bytes) return fail(FSVC_E_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  Ctx c;
  c.stream = stream;
  c.B = B;
  c.slope = h->cfg.lrelu_slope;
  c.eps = h->cfg.in_eps;
  c.prof = prof;
  if (prof) prof->mark(stream);
  const int n = h->n;
  const int T = frames * h->hop;
  const int S = h->cfg.spk_emb_size;

  if (spk) {  // every stage's emb_projector(normalize(spk)) in one launch                fastsvc.py:135-137
    SpkProjArgs sp;
    memset(&sp, 0, sizeof(sp));
    for (int i = 0; i < n; ++i) {
      sp.W[i] = h->stage[i].emb_w;
      sp.bias[i] = h->stage[i].emb_b;
      sp.e[i] = ws.e[i];
      sp.C[i] = h->cfg.mid_channels[i];
    }
    int c_max = 0;
    for (int i = 0; i < n; ++i) c_max = sp.C[i] > c_max ? sp.C[i] : c_max;
    spk_project_all_kernel<<<dim3(B, n, (c_max + 31) / 32), 256, 0, stream>>>(spk, S, sp);
    c.label = "";
    c.launched("spk_project", 0.0, 0.0);
  }

  // ---- conditioning chains, both branches per launch (fastsvc.py:180-193, 220-232) ----
  int T_prev = T, T_l = T;
  bool fused_l0 = false;
  for (int l = 0; l < n; ++l) {
    T_l = T_prev / h->dscale[l];
    const LevelW& lw = h->level[l];
    const int C = h->lvl_c[l];
    c.label = lvl_label[l];
    Tc2Args p[2];
    if (l == 0 && h->l0_fused) {
      const LevelFusedSmem LF =
          level_fused_smem(C, lw.c2[0].nc_G, lw.c2[0].nc_N, lw.film_out.nc_N);
      LevelFusedArgs fa;
      memset(&fa, 0, sizeof(fa));
      const int dec = n > 1 ? h->dscale[1] : 1;
      for (int br = 0; br < 2; ++br) {
        fa.sig[br] = br == 0 ? lft : sine;
        fa.c1_w[br] = lw.c1[br].w;
        fa.c1_b[br] = lw.c1[br].b;
        fa.r1_w[br] = lw.r1[br].w;
        fa.r1_b[br] = lw.r1[br].b;
        fa.w_c2[br] = lw.c2[br].wnc;
        fa.w_c4[br] = lw.c4[br].wnc;
        fa.w_film[br] = lw.film[br].wnc;
        fa.b_c2[br] = lw.c2[br].b;
        fa.b_c4[br] = lw.c4[br].b;
        fa.b_film[br] = lw.film[br].b;
        fa.y_dec[br] = (n > 1 && T_l % dec == 0) ? ws.ydec[br] : nullptr;
      }
      fa.w_out = lw.film_out.wnc;
      fa.b_out = lw.film_out.b;
      fa.gb = ws.GB[0];
      fa.C = C;
      fa.T = T_l;
      fa.B = B;
      fa.dec = dec;
      fa.n_tiles = (T_l + kLfValid - 1) / kLfValid;
      fa.Gp = lw.c2[0].nc_G;
      fa.N1 = lw.c2[0].nc_N;
      fa.N2 = lw.film_out.nc_N;
      fa.slope = c.slope;
      const int items = B * fa.n_tiles;
      const int grid = items < h->num_sms ? items : h->num_sms;
      if (level_fused_fill_desc(&fa) != 0) return fail(FSVC_E_INVALID, "internal: fused level descriptor table overflow");
      launch_pdl(level0_fused_kernel, dim3(grid), kLfThreads, LF.total, stream, fa);
      const double BT = (double)B * T_l;
      c.launched("fused_level", 2.0 * BT * C * (2.0 * (3 + 1 + 9.0 * C) + 2.0 * 9 * C + 12.0 * C),
                 4.0 * (2 * BT + 2 * BT * C / dec + 2 * BT * C));
      fused_l0 = true;
      T_prev = T_l;
      continue;
    }
    if (lw.c1[0].C_in == 1) {
      // 1-channel input: the first conv (and the 1x1 residual) are generated inside the consumers
      for (int br = 0; br < 2; ++br) {
        const float* sig = br == 0 ? lft : sine;
        p[br] = tc2_args(c, lw.c2[br], sig, 1, T_prev, T_l, 2, ws.tb[br], C);
        p[br].gen_w = lw.c1[br].w;
        p[br].gen_b = lw.c1[br].b;
        p[br].pre_lrelu = 1;
        p[br].down = 1;
      }
      launch_tc2(c, h, 3, p, 2, "down_d1+d2");
      for (int br = 0; br < 2; ++br) {
        const float* sig = br == 0 ? lft : sine;
        p[br] = tc2_args(c, lw.c4[br], ws.tb[br], C, T_l, T_l, 4, ws.y[br][l], C);
        p[br].pre_lrelu = 1;
        p[br].gres_w = lw.r1[br].w;
        p[br].gres_b = lw.r1[br].b;
        p[br].gres_x = sig;
      }
      launch_tc2(c, h, 3, p, 2, "down_d4+r");
    } else {
      const int Cp = h->lvl_c[l - 1];
      // the fused level-0 kernel hands over its output already decimated
      const bool dec_in = l == 1 && fused_l0;
      for (int br = 0; br < 2; ++br) {
        p[br] = tc2_args(c, lw.r1[br], dec_in ? ws.ydec[br] : ws.y[br][l - 1], Cp, dec_in ? T_l : T_prev, T_l, 1,
                         ws.tr[br], C);
        p[br].down = dec_in ? 1 : h->dscale[l];
      }
      launch_tc2(c, h, 1, p, 2, "down_r1x1");
      for (int br = 0; br < 2; ++br) {
        p[br] = tc2_args(c, lw.c1[br], dec_in ? ws.ydec[br] : ws.y[br][l - 1], Cp, dec_in ? T_l : T_prev, T_l, 1,
                         ws.ta[br], C);
        p[br].down = dec_in ? 1 : h->dscale[l];
        p[br].pre_lrelu = 1;
      }
      launch_tc2(c, h, 3, p, 2, "down_d1");
      for (int br = 0; br < 2; ++br) {
        p[br] = tc2_args(c, lw.c2[br], ws.ta[br], C, T_l, T_l, 2, ws.tb[br], C);
        p[br].pre_lrelu = 1;
      }
      launch_tc2(c, h, 3, p, 2, "down_d2");
      for (int br = 0; br < 2; ++br) {
        p[br] = tc2_args(c, lw.c4[br], ws.tb[br], C, T_l, T_l, 4, ws.y[br][l], C);
        p[br].pre_lrelu = 1;
        p[br].res = ws.tr[br];
        p[br].res_ld = C;
      }
      launch_tc2(c, h, 3, p, 2, "down_d4");
    }
    for (int br = 0; br < 2; ++br) {
      p[br] = tc2_args(c, lw.film[br], ws.y[br][l], C, T_l, T_l, 1, ws.H[l] + ntc_col(br * C), 2 * C);
      p[br].post_lrelu = 1;
    }
    launch_tc2(c, h, 3, p, 2, "film_conv");
    p[0] = tc2_args(c, lw.film_out, ws.H[l], 2 * C, T_l, T_l, 1, ws.GB[l], 2 * C);
    launch_tc2(c, h, 3, p, 1, "film_out");
    T_prev = T_l;
  }

  // ---- upsampling stages (fastsvc.py:80-140) ----
  {  // the caller's (B, C, T') PPG tensor -> channels-last
    if (h->ppg_ready) cudaStreamWaitEvent(stream, h->ppg_ready, 0);  // fsvc_forward_host: upload on the side stream
    const int Cin = h->cfg.in_channels;
    nct_to_ntc_kernel<<<dim3((frames + 31) / 32, (Cin + 31) / 32, B), 256, 0, stream>>>(ppg, Cin, frames, ws.xin);
    c.label = "";
    c.launched("ppg_to_ntc", 0.0, 8.0 * B * Cin * frames);
  }
  const float* x = ws.xin;
  int x_ld = h->cfg.in_channels, T_in = frames;
  for (int i = 0; i < n; ++i) {
    const StageW& w = h->stage[i];
    const int C = h->cfg.mid_channels[i], r = h->cfg.upsampling_scales[i];
    const int l = n - 1 - i, T_s = T_in * r, n_seg = (T_s + 31) / 32;
    const float* gamma = ws.GB[l];
    const float* beta = ws.GB[l] + ntc_col(C);
    const bool norm = spk != nullptr;
    c.label = stage_label[i];
    // InstanceNorm: a conv with `film` writes per-segment (mean, M2) partials of its output; the next conv merges
    // them itself while it loads (conv_tc3 transform role) -- no finalize launch in between.
    // Long utterances (many segments) keep the separate merge kernel: inside the consumer the merge would sit on
    // every CTA's critical path (measured: +26 us per conv at 500 segments vs a 12 us launch).
    const bool fold = n_seg <= 128;
    int st_w = 0;  // statistics buffer the next producer writes
    auto film = [&](Tc2Args& a) {
      a.gamma = gamma;
      a.beta = beta;
      a.gb_ld = 2 * C;
      if (norm) {
        a.stats = ws.stats[st_w];
        a.n_seg = n_seg;
        st_w ^= 1;
      }
    };
    auto finalize = [&]() {
      if (!norm || fold) return;
      in_finalize2_kernel<<<(B * C + 7) / 8, 256, 0, stream>>>(ws.stats[st_w ^ 1], n_seg, T_s, C, B * C, ws.e[i], c.eps,
                                                               ws.pa, ws.pc);
      c.launched("in_finalize", 0.0, 8.0 * B * C * n_seg);
    };
    auto pre = [&](Tc2Args& a) {
      if (norm && fold) {
        a.pre_stats = ws.stats[st_w ^ 1];  // written by the previous conv of this stage
        a.pre_nseg = n_seg;
        a.pre_e = ws.e[i];
        a.pre_eps = c.eps;
      } else if (norm) {
        a.pre_a = ws.pa;
        a.pre_c = ws.pc;
      }
      a.pre_lrelu = 1;
    };
    Tc2Args p[2];
    // h0 = conv_first(x)                                                     fastsvc.py:93
    p[0] = tc2_args(c, w.first, x, x_ld, T_in, T_in, 1, ws.h0[i], C);
    launch_tc2(c, h, 3, p, 1, "conv_first");
    // xr = Conv3(repeat_r(h0)) ; t1 = gamma*lrelu(Conv3(repeat_r(lrelu(h0)))) + beta   :94, :97-98
    p[0] = tc2_args(c, w.res, ws.h0[i], C, T_in, T_s, 1, ws.xr[i], C);
    p[0].up = r;
    launch_tc2(c, h, 3, p, 1, "residual");
    p[0] = tc2_args(c, w.up, ws.h0[i], C, T_in, T_s, 1, ws.t1[i], C);
    p[0].up = r;
    p[0].pre_lrelu = 1;
    p[0].post_lrelu = 1;
    film(p[0]);
    launch_tc2(c, h, 3, p, 1, "up_film");
    finalize();
    // x_ = Conv3_d3(lrelu(IN(t1)+e)) + xr ; t2 = gamma*x_ + beta            :99-105
    p[0] = tc2_args(c, w.d3, ws.t1[i], C, T_s, T_s, 3, ws.t2[i], C);
    pre(p[0]);
    p[0].res = ws.xr[i];
    p[0].res_ld = C;
    p[0].raw = ws.x_[i];
    p[0].raw_ld = C;
    film(p[0]);
    launch_tc2(c, h, 3, p, 1, "d3_film");
    finalize();
    // t3 = gamma * Conv3_d9(lrelu(IN(t2)+e)) + beta                         :106-107
    p[0] = tc2_args(c, w.d9, ws.t2[i], C, T_s, T_s, 9, ws.t1[i], C);
    pre(p[0]);
    film(p[0]);
    launch_tc2(c, h, 3, p, 1, "d9_film");
    finalize();
    // out = Conv3_d27(lrelu(IN(t3)+e)) + x_                                 :108-111
    p[0] = tc2_args(c, w.d27, ws.t1[i], C, T_s, T_s, 27, ws.xs[i], C);
    pre(p[0]);
    p[0].res = ws.x_[i];
    p[0].res_ld = C;
    launch_tc2(c, h, 3, p, 1, "d27_skip");
    x = ws.xs[i];
    x_ld = C;
    T_in = T_s;
  }
  // conv_last (1x1)                                                         fastsvc.py:330
  {
    const long long BT = (long long)B * T;
    const int C = h->cfg.mid_channels[n - 1];
    conv_last_ntc_kernel<<<(unsigned)((BT + 255) / 256), 256, 0, stream>>>(x, C, T, BT, h->last.w, h->last.b,
                                                                          h->cfg.out_channels, out);
    c.label = "";
    c.launched("conv_last", 2.0 * BT * C * h->cfg.out_channels, 4.0 * BT * (C + h->cfg.out_channels));
  }
  h->launches = c.launches;
  if (c.err) return fail(FSVC_E_INVALID, "internal: conv launch plan failed (code %d)", c.err);
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}

// AUTO / TC_BF16X3 run the channels-last tensor-core forward when the configuration allows it (every
// mid channel count a multiple of 8); otherwise TC_BF16X3 uses the per-layer tcgen05 kernel and AUTO fp32.
static bool use_tc2(const fsvc_handle* h, int mode) { return mode != FSVC_MODE_FP32 && h->tc2_ok; }

}  // namespace fsvc

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int fsvc_abi_version(void) { return FSVC_ABI_VERSION; }
const char* fsvc_last_error(void) { return g_err; }

int fsvc_create(const fsvc_config* cfg, fsvc_handle** out) {
  if (!cfg || !out) return fail(FSVC_E_INVALID, "null argument");
  *out = nullptr;
  const int n = cfg->num_stages;
  if (n < 1 || n > FSVC_MAX_STAGES) return fail(FSVC_E_INVALID, "num_stages must be in [1, %d]", FSVC_MAX_STAGES);
  if (cfg->in_channels < 1 || cfg->out_channels < 1) return fail(FSVC_E_INVALID, "bad channel count");
  for (int i = 0; i < n; ++i)
    if (cfg->mid_channels[i] < 1 || cfg->upsampling_scales[i] < 1)
      return fail(FSVC_E_INVALID, "mid_channels / upsampling_scales must be positive");
  if (cfg->use_spk_emb && cfg->spk_emb_size < 1) return fail(FSVC_E_INVALID, "bad spk_emb_size");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(FSVC_E_NODEVICE, "no CUDA device: libfsvc has no CPU fallback");
  }
  fsvc_handle* h = new fsvc_handle();
  h->cfg = *cfg;
  h->n = n;
  cudaGetDevice(&h->device);
  // downsampling scales: reverse, drop last, put 1 in front (fastsvc.py:270-272)
  h->hop = 1;
  for (int i = 0; i < n; ++i) h->hop *= cfg->upsampling_scales[i];
  h->dscale[0] = 1;
  for (int l = 1; l < n; ++l) h->dscale[l] = cfg->upsampling_scales[n - l];
  for (int l = 0; l < n; ++l) h->lvl_c[l] = cfg->mid_channels[n - 1 - l];

  // canonical weight list + packed store layout
  size_t off = 0;
  auto add_info = [&](const std::string& prefix, int64_t wn, int64_t bn) {
    h->winfo.push_back({prefix + ".weight", wn});
    h->winfo.push_back({prefix + ".bias", bn});
  };
  auto reserve = [&](size_t count) {
    size_t o = off;
    off += (count + 63) & ~(size_t)63;
    return o;
  };
  std::vector<std::pair<ConvW*, size_t>> fix;  // (conv, w offset) ; bias follows
  auto add_conv = [&](ConvW& cw, const std::string& prefix, int co, int ci, int K, bool info = true) {
    cw.C_in = ci;
    cw.C_out = co;
    cw.K = K;
    size_t wo = reserve((size_t)co * ci * K);
    size_t bo = reserve(co);
    fix.push_back({&cw, wo});
    cw.b = (float*)bo;  // patched below
    h->convs.push_back(&cw);
    if (info) add_info(prefix, (int64_t)co * ci * K, co);
  };
  std::vector<std::pair<float**, size_t>> fixp;
  int cin = cfg->in_channels;
  for (int i = 0; i < n; ++i) {
    const int C = cfg->mid_channels[i];
    const std::string p = "upsampling_nets." + std::to_string(i);
    StageW& s = h->stage[i];
    add_conv(s.first, p + ".conv_first", C, cin, 3);
    add_conv(s.up, p + ".upsample_block0.2", C, C, 3);
    add_conv(s.d3, p + ".conv_block1.1", C, C, 3);
    add_conv(s.d9, p + ".conv_block2.1", C, C, 3);
    add_conv(s.d27, p + ".conv_block3.1", C, C, 3);
    add_conv(s.res, p + ".residual_block.1", C, C, 3);
    s.up.ctx_up = s.res.ctx_up = cfg->upsampling_scales[i];
    s.d3.ctx_dil = 3;
    s.d9.ctx_dil = 9;
    s.d27.ctx_dil = 27;
    if (cfg->use_spk_emb) {
      fixp.push_back({&s.emb_w, reserve((size_t)C * cfg->spk_emb_size)});
      fixp.push_back({&s.emb_b, reserve(C)});
      add_info(p + ".emb_projector", (int64_t)C * cfg->spk_emb_size, C);
    }
    cin = C;
  }
  const char* dn[2] = {"downsampling_lft.", "downsampling_sine."};
  const char* fn[2] = {"film_lft.", "film_sine."};
  for (int br = 0; br < 2; ++br) {
    int ci = 1;
    for (int l = 0; l < n; ++l) {
      const int C = h->lvl_c[l];
      const std::string p = dn[br] + std::to_string(l);
      add_conv(h->level[l].r1[br], p + ".residual_block.0", C, ci, 1);
      add_conv(h->level[l].c1[br], p + ".downsample_block.2", C, ci, 3);
      add_conv(h->level[l].c2[br], p + ".downsample_block.4", C, C, 3);
      add_conv(h->level[l].c4[br], p + ".downsample_block.6", C, C, 3);
      h->level[l].c2[br].ctx_dil = 2;
      h->level[l].c4[br].ctx_dil = 4;
      for (ConvW* cw : {&h->level[l].r1[br], &h->level[l].c1[br], &h->level[l].c2[br], &h->level[l].c4[br]})
        cw->ctx_wide = 1;
      ci = C;
    }
  }
  for (int br = 0; br < 2; ++br)
    for (int l = 0; l < n; ++l) {
      const int C = h->lvl_c[l];
      const std::string p = fn[br] + std::to_string(l);
      add_conv(h->level[l].film[br], p + ".conv", C, C, 3);
      h->level[l].film[br].ctx_wide = 1;
      add_info(p + ".conv_scale", (int64_t)C * C * 3, C);
      add_info(p + ".conv_shift", (int64_t)C * C * 3, C);
    }
  for (int l = 0; l < n; ++l) {
    add_conv(h->level[l].film_out, "", 2 * h->lvl_c[l], 2 * h->lvl_c[l], 3, false);
    h->level[l].film_out.ctx_wide = 1;
  }
  add_conv(h->last, "conv_last", cfg->out_channels, cfg->mid_channels[n - 1], 1);

  h->store_floats = off;
  if (cudaMalloc((void**)&h->store, off * sizeof(float)) != cudaSuccess) {
    int rc = fail(FSVC_E_CUDA, "cudaMalloc(%zu) failed: %s", off * sizeof(float), cudaGetErrorString(cudaGetLastError()));
    delete h;
    return rc;
  }
  cudaMemset(h->store, 0, off * sizeof(float));
  for (auto& f : fix) {
    f.first->w = h->store + f.second;
    f.first->b = h->store + (size_t)f.first->b;
  }
  for (auto& f : fixp) *f.first = h->store + f.second;
  // tensor-core weight copies
  size_t tc_off = 0;
  for (ConvW* cw : h->convs)
    if (tc_plan(cw->C_in, cw->C_out, cw->K, &cw->tc)) {
      cw->tc.w = (const __nv_bfloat16*)tc_off;  // offset, patched below
      tc_off += (cw->tc.elems() + 127) & ~(size_t)127;
    }
  // copies tiled for the channels-last persistent kernel; the forward is eligible when every conv
  // except the 1-channel ones (level-0 first conv / 1x1 residual: generated in-kernel) and conv_last has one
  h->tc2_ok = true;
  for (ConvW* cw : h->convs) {
    const bool tiny = (cw->C_in == 1) || cw == &h->last;
    if (tiny) continue;
    if (tc2_plan(cw->C_in, cw->C_out, cw->K, cw)) {
      cw->tc2.w = (const __nv_bfloat16*)tc_off;
      tc_off += (cw->tc2.elems() + 127) & ~(size_t)127;
    } else {
      h->tc2_ok = false;
    }
  }
  for (int i = 0; i < n; ++i)
    if (cfg->mid_channels[i] % 8 != 0) h->tc2_ok = false;
  {  // fused level-0 kernel: every weight of the level in shared memory next to four activation buffers
    LevelW& lw = h->level[0];
    const int C = h->lvl_c[0];
    if (h->tc2_ok && C % 8 == 0 && C <= 32) {
      const int Gp = (C / 8 + 1) / 2 * 2, N1 = (C + 15) / 16 * 16, N2 = (2 * C + 15) / 16 * 16;
      if (level_fused_smem(C, Gp, N1, N2).total <= 227u * 1024u) {
        h->l0_fused = true;
        auto want = [&](ConvW& cw, int G, int N) {
          cw.nc_G = G;
          cw.nc_N = N;
          cw.wnc = (const __nv_bfloat16*)tc_off;
          tc_off += ((size_t)cw.K * G * 2 * N * 8 + 127) & ~(size_t)127;
        };
        for (int br = 0; br < 2; ++br) {
          want(lw.c2[br], Gp, N1);
          want(lw.c4[br], Gp, N1);
          want(lw.film[br], Gp, N1);
        }
        want(lw.film_out, 2 * C / 8, N2);
      }
    }
  }
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess) h->num_sms = prop.multiProcessorCount;
  }
  h->tc_elems = tc_off;
  if (tc_off && cudaMalloc((void**)&h->tc_store, tc_off * sizeof(__nv_bfloat16)) != cudaSuccess) {
    int rc = fail(FSVC_E_CUDA, "cudaMalloc(tc weights) failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(h->store);
    delete h;
    return rc;
  }
  for (ConvW* cw : h->convs) {
    if (cw->tc.n_ntiles) cw->tc.w = h->tc_store + (size_t)cw->tc.w;
    if (cw->tc2.n_ntiles) cw->tc2.w = h->tc_store + (size_t)cw->tc2.w;
    if (cw->nc_G) cw->wnc = h->tc_store + (size_t)cw->wnc;
  }
  if (int rc = tc_setup_kernels()) {
    cudaFree(h->store);
    cudaFree(h->tc_store);
    delete h;
    return rc;
  }
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_ppg, cudaEventDisableTiming) != cudaSuccess) {
    fsvc_destroy(h);
    return fail(FSVC_E_CUDA, "cannot create the upload stream / events");
  }
  *out = h;
  return FSVC_OK;
}

void fsvc_destroy(fsvc_handle* h) {
  if (!h) return;
  if (h->store) cudaFree(h->store);
  if (h->tc_store) cudaFree(h->tc_store);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_ppg) cudaEventDestroy(h->ev_ppg);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
}

int fsvc_num_weight_tensors(const fsvc_handle* h) { return h ? (int)h->winfo.size() : FSVC_E_INVALID; }

int fsvc_weight_tensor_info(const fsvc_handle* h, int index, char* name, int name_capacity, int64_t* numel) {
  if (!h || index < 0 || index >= (int)h->winfo.size()) return fail(FSVC_E_INVALID, "bad weight index %d", index);
  if (name && name_capacity > 0) snprintf(name, name_capacity, "%s", h->winfo[index].name.c_str());
  if (numel) *numel = h->winfo[index].numel;
  return FSVC_OK;
}

int fsvc_set_weights(fsvc_handle* h, const float* const* p, int n, void* stream_) {
  if (!h || !p) return fail(FSVC_E_INVALID, "null argument");
  if (n != (int)h->winfo.size()) return fail(FSVC_E_INVALID, "expected %d weight tensors, got %d", (int)h->winfo.size(), n);
  for (int i = 0; i < n; ++i)
    if (!p[i]) return fail(FSVC_E_INVALID, "weight tensor %d (%s) is null", i, h->winfo[i].name.c_str());
  cudaStream_t s = (cudaStream_t)stream_;
  int k = 0;
  auto put = [&](ConvW& cw) {
    repack(s, p[k], cw.C_out, cw.C_in, cw.K, cw.w, cw.C_out, 0, 0);
    bias_sum(s, p[k + 1], nullptr, cw.C_out, cw.b, 0);
    k += 2;
  };
  const int ns = h->n;
  for (int i = 0; i < ns; ++i) {
    StageW& st = h->stage[i];
    put(st.first);
    put(st.up);
    put(st.d3);
    put(st.d9);
    put(st.d27);
    put(st.res);
    if (h->cfg.use_spk_emb) {
      const int C = h->cfg.mid_channels[i];
      FSVC_CUDA(cudaMemcpyAsync(st.emb_w, p[k], (size_t)C * h->cfg.spk_emb_size * sizeof(float),
                                cudaMemcpyDeviceToDevice, s));
      FSVC_CUDA(cudaMemcpyAsync(st.emb_b, p[k + 1], (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, s));
      k += 2;
    }
  }
  for (int br = 0; br < 2; ++br)
    for (int l = 0; l < ns; ++l) {
      put(h->level[l].r1[br]);
      put(h->level[l].c1[br]);
      put(h->level[l].c2[br]);
      put(h->level[l].c4[br]);
    }
  for (int br = 0; br < 2; ++br)
    for (int l = 0; l < ns; ++l) {
      const int C = h->lvl_c[l];
      LevelW& lw = h->level[l];
      put(lw.film[br]);
      // conv_scale -> output cols [0,C), conv_shift -> [C,2C); branch br reads input rows [br*C, (br+1)*C)
      repack(s, p[k], C, C, 3, lw.film_out.w, 2 * C, br * C, 0);
      repack(s, p[k + 2], C, C, 3, lw.film_out.w, 2 * C, br * C, C);
      k += 4;
    }
  // merged FiLM biases: gamma bias = b_scale_lft + b_scale_sine, beta bias likewise
  {
    // index of film_lft.l.conv_scale.bias etc. in the canonical list
    auto find = [&](const std::string& name) {
      for (int i = 0; i < n; ++i)
        if (h->winfo[i].name == name) return i;
      return -1;
    };
    for (int l = 0; l < ns; ++l) {
      const int C = h->lvl_c[l];
      const std::string sl = std::to_string(l);
      const int a0 = find("film_lft." + sl + ".conv_scale.bias"), a1 = find("film_sine." + sl + ".conv_scale.bias");
      const int b0 = find("film_lft." + sl + ".conv_shift.bias"), b1 = find("film_sine." + sl + ".conv_shift.bias");
      bias_sum(s, p[a0], p[a1], C, h->level[l].film_out.b, 0);
      bias_sum(s, p[b0], p[b1], C, h->level[l].film_out.b, C);
    }
  }
  put(h->last);
  if (k != n) return fail(FSVC_E_STATE, "internal: consumed %d of %d weight tensors", k, n);
  for (ConvW* cw : h->convs)
  {
    if (cw->tc.w) pack_tc(s, *cw, cw->tc);
    if (cw->tc2.w) pack_tc(s, *cw, cw->tc2);
    if (cw->nc_G) {
      const size_t total = (size_t)cw->K * cw->nc_G * cw->nc_N * 8;
      pack_tc_nc_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
          cw->w, cw->C_in, cw->C_out, cw->K, cw->nc_G, cw->nc_N, (__nv_bfloat16*)cw->wnc);
    }
  }
  FSVC_CUDA(cudaGetLastError());
  h->weights_set = true;
  return FSVC_OK;
}

static int check_shape(const fsvc_handle* h, int B, int frames) {
  if (!h) return fail(FSVC_E_INVALID, "null handle");
  if (B < 1 || frames < 1) return fail(FSVC_E_INVALID, "B and frames must be >= 1 (got %d, %d)", B, frames);
  if (B > 65535) return fail(FSVC_E_INVALID, "B must be <= 65535");
  if ((long long)frames * h->hop > (1LL << 30)) return fail(FSVC_E_INVALID, "utterance too long");
  return FSVC_OK;
}

size_t fsvc_workspace_bytes(const fsvc_handle* h, int B, int frames, int mode) {
  if (check_shape(h, B, frames) != FSVC_OK) return 0;
  if (use_tc2(h, mode)) {
    WS2 ws2;
    return layout_ws2(h, B, frames, nullptr, 0, &ws2);
  }
  WS ws;
  return layout_ws(h, B, frames, nullptr, 0, &ws);
}

int fsvc_forward(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk, float* out,
                 int B, int frames, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  int rc = check_shape(h, B, frames);
  if (rc) return rc;
  if (!ppg || !sine || !lft || !out || !workspace) return fail(FSVC_E_INVALID, "null tensor pointer");
  if (!h->weights_set) return fail(FSVC_E_STATE, "fsvc_forward called before fsvc_set_weights");
  if (spk && !h->cfg.use_spk_emb)
    return fail(FSVC_E_INVALID, "spk given but the generator was built with use_spk_emb=0 (no emb_projector)");
  if (mode != FSVC_MODE_FP32 && mode != FSVC_MODE_TC_BF16X3 && mode != FSVC_MODE_AUTO)
    return fail(FSVC_E_INVALID, "unknown mode %d", mode);
  if (use_tc2(h, mode))
    return forward_tc2(h, ppg, sine, lft, spk, out, B, frames, workspace, workspace_bytes, (cudaStream_t)stream);
  return forward_fp32(h, ppg, sine, lft, spk, out, B, frames, workspace, workspace_bytes, (cudaStream_t)stream, mode);
}

int fsvc_forward_profile(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                         float* out, int B, int frames, void* workspace, size_t workspace_bytes, int mode,
                         void* stream, fsvc_kernel_record* records, int capacity, int* count) {
  int rc = check_shape(h, B, frames);
  if (rc) return rc;
  if (!ppg || !sine || !lft || !out || !workspace || !records || !count) return fail(FSVC_E_INVALID, "null pointer");
  if (!h->weights_set) return fail(FSVC_E_STATE, "fsvc_forward_profile called before fsvc_set_weights");
  if (spk && !h->cfg.use_spk_emb) return fail(FSVC_E_INVALID, "spk given but use_spk_emb=0");
  Profiler prof;
  if (use_tc2(h, mode))
    rc = forward_tc2(h, ppg, sine, lft, spk, out, B, frames, workspace, workspace_bytes, (cudaStream_t)stream, &prof);
  else
    rc = forward_fp32(h, ppg, sine, lft, spk, out, B, frames, workspace, workspace_bytes, (cudaStream_t)stream, mode,
                      &prof);
  if (rc == FSVC_OK && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
    rc = fail(FSVC_E_CUDA, "stream synchronize failed: %s", cudaGetErrorString(cudaGetLastError()));
  int n = 0;
  if (rc == FSVC_OK) {
    for (size_t i = 0; i < prof.rec.size() && n < capacity; ++i, ++n) {
      records[n] = prof.rec[i];
      cudaEventElapsedTime(&records[n].ms, prof.ev[i], prof.ev[i + 1]);
    }
  }
  for (cudaEvent_t e : prof.ev) cudaEventDestroy(e);
  *count = n;
  return rc;
}

size_t fsvc_host_io_bytes(const fsvc_handle* h, int B, int frames) {
  if (check_shape(h, B, frames) != FSVC_OK) return 0;
  const size_t T = (size_t)frames * h->hop;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return al((size_t)B * h->cfg.in_channels * frames * 4) + 2 * al((size_t)B * T * 4) +
         al((size_t)B * h->cfg.spk_emb_size * 4) + al((size_t)B * h->cfg.out_channels * T * 4);
}

int fsvc_forward_host(fsvc_handle* h, const float* ppg_host, const float* sine_host, const float* lft_host,
                      const float* spk_host, float* out_host, int B, int frames, void* workspace,
                      size_t workspace_bytes, int mode, void* stream_) {
  int rc = check_shape(h, B, frames);
  if (rc) return rc;
  if (!ppg_host || !sine_host || !lft_host || !out_host || !workspace) return fail(FSVC_E_INVALID, "null pointer");
  const size_t io = fsvc_host_io_bytes(h, B, frames);
  if (workspace_bytes < io) return fail(FSVC_E_WORKSPACE, "workspace smaller than the host-io staging area");
  cudaStream_t s = (cudaStream_t)stream_;
  const size_t T = (size_t)frames * h->hop;
  Arena ar(workspace, workspace_bytes);
  const size_t n_ppg = (size_t)B * h->cfg.in_channels * frames, n_sig = (size_t)B * T;
  const size_t n_spk = (size_t)B * h->cfg.spk_emb_size, n_out = (size_t)B * h->cfg.out_channels * T;
  float* d_ppg = ar.get<float>(n_ppg);
  float* d_sine = ar.get<float>(n_sig);
  float* d_lft = ar.get<float>(n_sig);
  float* d_spk = ar.get<float>(n_spk);
  float* d_out = ar.get<float>(n_out);
  // The two signals (and the speaker vectors) go first on the caller's stream: the conditioning levels need only them.
  // The PPG tensor follows on the side stream -- after the signals on the copy engine, concurrently with the levels --
  // and the forward waits for it where it first reads it (stage 0).  The side stream starts after everything already
  // queued on the caller's stream, so the staging buffers of a previous call are never overwritten early.
  FSVC_CUDA(cudaMemcpyAsync(d_sine, sine_host, n_sig * 4, cudaMemcpyHostToDevice, s));
  FSVC_CUDA(cudaMemcpyAsync(d_lft, lft_host, n_sig * 4, cudaMemcpyHostToDevice, s));
  if (spk_host) FSVC_CUDA(cudaMemcpyAsync(d_spk, spk_host, n_spk * 4, cudaMemcpyHostToDevice, s));
  FSVC_CUDA(cudaEventRecord(h->ev_fork, s));
  FSVC_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_fork, 0));
  FSVC_CUDA(cudaMemcpyAsync(d_ppg, ppg_host, n_ppg * 4, cudaMemcpyHostToDevice, h->copy_stream));
  FSVC_CUDA(cudaEventRecord(h->ev_ppg, h->copy_stream));
  h->ppg_ready = h->ev_ppg;
  rc = fsvc_forward(h, d_ppg, d_sine, d_lft, spk_host ? d_spk : nullptr, d_out, B, frames, (char*)workspace + ar.off,
                    workspace_bytes - ar.off, mode, stream_);
  h->ppg_ready = nullptr;
  if (rc) {
    cudaStreamWaitEvent(s, h->ev_ppg, 0);  // rejoin the side stream even when the forward was not enqueued
    return rc;
  }
  FSVC_CUDA(cudaMemcpyAsync(out_host, d_out, n_out * 4, cudaMemcpyDeviceToHost, s));
  return FSVC_OK;
}

// ---- block-level entry points ------------------------------------------------
static ConvW tmp_conv(Arena& ar, cudaStream_t s, const float* w, const float* b, int co, int ci, int K) {
  ConvW cw;
  cw.C_in = ci;
  cw.C_out = co;
  cw.K = K;
  cw.w = ar.get<float>((size_t)co * ci * K);
  cw.b = ar.get<float>(co);
  __nv_bfloat16* tw = nullptr;
  if (tc_plan(ci, co, K, &cw.tc)) tw = ar.get<__nv_bfloat16>(cw.tc.elems());
  if (ar.ok()) {
    repack(s, w, co, ci, K, cw.w, co, 0, 0);
    bias_sum(s, b, nullptr, co, cw.b, 0);
    if (tw) {
      cw.tc.w = tw;
      pack_tc(s, cw, cw.tc);
    }
  }
  return cw;
}

size_t fsvc_block_workspace_bytes(int B, int c_in, int c, int T_out) {
  if (B < 1 || c_in < 1 || c < 1 || T_out < 1) return 0;
  const size_t act = ((size_t)B * c * T_out * 4 + 255) & ~(size_t)255;
  const size_t cm = (size_t)((c_in > c ? c_in : c) + 64);
  const size_t wts = 8 * ((((size_t)c + 64) * cm * 3 * (4 + 4 * 2) + 255 + 1024) & ~(size_t)255);
  return 10 * act + wts + (size_t)B * c * 4 * 4 + (1 << 16) + (size_t)B * c * (T_out / 128 + 2) * 8;
}

int fsvc_downsample_forward(const float* x, float* out, const float* const* w, int B, int c_in, int c, int T,
                            int scale, float slope, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  if (!x || !out || !w || !workspace) return fail(FSVC_E_INVALID, "null pointer");
  if (B < 1 || c_in < 1 || c < 1 || T < 1 || scale < 1) return fail(FSVC_E_INVALID, "bad shape");
  if (T % scale) return fail(FSVC_E_INVALID, "T (%d) must be divisible by the downsampling scale (%d)", T, scale);
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(workspace, workspace_bytes);
  ConvW r1 = tmp_conv(ar, s, w[0], w[1], c, c_in, 1), c1 = tmp_conv(ar, s, w[2], w[3], c, c_in, 3);
  ConvW c2 = tmp_conv(ar, s, w[4], w[5], c, c, 3), c4 = tmp_conv(ar, s, w[6], w[7], c, c, 3);
  const size_t ne = (size_t)B * c * (T / scale);
  float *tr = ar.get<float>(ne), *ta = ar.get<float>(ne), *tb = ar.get<float>(ne);
  if (!ar.ok()) return fail(FSVC_E_WORKSPACE, "workspace too small: need %zu bytes", ar.off);
  Ctx ctx;
  ctx.stream = s;
  ctx.B = B;
  ctx.slope = slope;
  ctx.eps = 0.f;
  ctx.tc = mode_uses_tc(mode);
  if (tc_setup_kernels()) return FSVC_E_CUDA;
  run_downsample(ctx, r1, c1, c2, c4, x, T, scale, tr, ta, tb, out);
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}

int fsvc_film_forward(const float* x, float* scale, float* shift, const float* const* w, int B, int c, int T,
                      float slope, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  if (!x || !scale || !shift || !w || !workspace) return fail(FSVC_E_INVALID, "null pointer");
  if (B < 1 || c < 1 || T < 1) return fail(FSVC_E_INVALID, "bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(workspace, workspace_bytes);
  ConvW cv = tmp_conv(ar, s, w[0], w[1], c, c, 3), cs = tmp_conv(ar, s, w[2], w[3], c, c, 3);
  ConvW ch = tmp_conv(ar, s, w[4], w[5], c, c, 3);
  float* hbuf = ar.get<float>((size_t)B * c * T);
  if (!ar.ok()) return fail(FSVC_E_WORKSPACE, "workspace too small: need %zu bytes", ar.off);
  Ctx ctx;
  ctx.stream = s;
  ctx.B = B;
  ctx.slope = slope;
  ctx.eps = 0.f;
  ctx.tc = mode_uses_tc(mode);
  if (tc_setup_kernels()) return FSVC_E_CUDA;
  ConvArgs a = conv_args(ctx, cv, x, T, T, 1, hbuf);
  a.post_lrelu = 1;
  launch_conv(ctx, a, 3);
  a = conv_args(ctx, cs, hbuf, T, T, 1, scale);
  launch_conv(ctx, a, 3);
  a = conv_args(ctx, ch, hbuf, T, T, 1, shift);
  launch_conv(ctx, a, 3);
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}

namespace fsvc {
__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o,
                            size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}
}  // namespace fsvc

int fsvc_upsample_forward(const float* x, const float* s_scale, const float* s_shift, const float* l_scale,
                          const float* l_shift, const float* spk, float* out, const float* const* w, int B, int c_in,
                          int c, int T, int scale, int spk_emb_size, float slope, float eps, void* workspace,
                          size_t workspace_bytes, int mode, void* stream) {
  if (!x || !s_scale || !s_shift || !l_scale || !l_shift || !out || !w || !workspace)
    return fail(FSVC_E_INVALID, "null pointer");
  if (B < 1 || c_in < 1 || c < 1 || T < 1 || scale < 1) return fail(FSVC_E_INVALID, "bad shape");
  if (spk && (!w[12] || !w[13] || spk_emb_size < 1)) return fail(FSVC_E_INVALID, "spk given without emb_projector");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(workspace, workspace_bytes);
  StageW sw;
  sw.first = tmp_conv(ar, s, w[0], w[1], c, c_in, 3);
  sw.up = tmp_conv(ar, s, w[2], w[3], c, c, 3);
  sw.d3 = tmp_conv(ar, s, w[4], w[5], c, c, 3);
  sw.d9 = tmp_conv(ar, s, w[6], w[7], c, c, 3);
  sw.d27 = tmp_conv(ar, s, w[8], w[9], c, c, 3);
  sw.res = tmp_conv(ar, s, w[10], w[11], c, c, 3);
  const int To = T * scale;
  const size_t ne = (size_t)B * c * To;
  float *gamma = ar.get<float>(ne), *beta = ar.get<float>(ne);
  float* e = ar.get<float>((size_t)B * c);
  StageBufs sb = alloc_stage_bufs(ar, B, c, T, To);
  if (!ar.ok()) return fail(FSVC_E_WORKSPACE, "workspace too small: need %zu bytes", ar.off);
  Ctx ctx;
  ctx.stream = s;
  ctx.B = B;
  ctx.slope = slope;
  ctx.eps = eps;
  ctx.tc = mode_uses_tc(mode);
  if (tc_setup_kernels()) return FSVC_E_CUDA;
  add2_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>(s_scale, l_scale, gamma, ne);
  add2_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>(s_shift, l_shift, beta, ne);
  if (spk) spk_project_kernel<<<B, 256, 0, s>>>(spk, spk_emb_size, w[12], w[13], c, e);
  run_stage(ctx, sw, x, T, scale, gamma, beta, (long long)c * To, spk ? e : nullptr, sb, out);
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}

// ---- excitation (before the generator) and PCM-16 quantisation (after it) ---------------------------
int fsvc_sine_excitation(const float* f0, const float* noise, float* out, int B, int frames, int hop,
                         float sample_rate, float sine_amp, float noise_amp, void* stream) {
  if (!f0 || !out || B < 0 || frames < 0 || hop < 1 || hop > 4096 || !(sample_rate > 0.f))
    return fail(FSVC_E_INVALID, "fsvc_sine_excitation: bad arguments (B=%d frames=%d hop=%d)", B, frames, hop);
  if (noise_amp > 0.f && !noise) return fail(FSVC_E_INVALID, "fsvc_sine_excitation: noise_amp > 0 needs a noise buffer");
  if (B == 0 || frames == 0) return FSVC_OK;
  if (B > 65535) return fail(FSVC_E_INVALID, "fsvc_sine_excitation: at most 65535 utterances per call");
  ExcArgs a;
  a.f0 = f0;
  a.noise = noise_amp > 0.f ? noise : nullptr;
  a.out = out;
  a.B = B;
  a.frames = frames;
  a.hop = hop;
  a.sample_rate = sample_rate;
  a.sine_amp = sine_amp;
  a.noise_amp = noise_amp;
  sine_excitation_kernel<<<dim3((frames + kExcFrames - 1) / kExcFrames, B), kExcThreads, 0, (cudaStream_t)stream>>>(a);
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}

int fsvc_pcm16(const float* x, int16_t* y, long long n, void* stream) {
  if (n < 0 || (n > 0 && (!x || !y))) return fail(FSVC_E_INVALID, "fsvc_pcm16: bad arguments");
  if (n == 0) return FSVC_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pcm16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, n);
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}

#ifdef FSVC_TIMELINE
// debug builds only (tools/timeline.py): copy the event stamps of the last forward
int fsvc_debug_timeline(unsigned long long* out, int n) {
  FSVC_CUDA(cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * (size_t)(n < 64 * 8 * 64 ? n : 64 * 8 * 64)));
  return FSVC_OK;
}
#endif

int fsvc_last_launch_count(const fsvc_handle* h) { return h ? h->launches : FSVC_E_INVALID; }

}  // extern "C"
