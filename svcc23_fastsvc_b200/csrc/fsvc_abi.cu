// libfsvc.so -- C ABI (include/fsvc.h) and host-side launch plan of the
// B200-native FastSVC generator forward.  sm_100a only; no CPU path.
#include "fsvc_internal.h"
#include "conv_f32_launch.cuh"
#include "excitation.cuh"

namespace fsvc {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

const char* const stage_label[FSVC_MAX_STAGES] = {"s0", "s1", "s2", "s3", "s4", "s5", "s6", "s7"};
const char* const lvl_label[FSVC_MAX_STAGES] = {"l0", "l1", "l2", "l3", "l4", "l5", "l6", "l7"};
const char* const lvl_lft_label[FSVC_MAX_STAGES] = {"l0.lft", "l1.lft", "l2.lft", "l3.lft",
                                                    "l4.lft", "l5.lft", "l6.lft", "l7.lft"};
const char* const lvl_sine_label[FSVC_MAX_STAGES] = {"l0.sine", "l1.sine", "l2.sine", "l3.sine",
                                                     "l4.sine", "l5.sine", "l6.sine", "l7.sine"};

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
  const int tile_len = conv_tile_len(T), n_tiles = (T + tile_len - 1) / tile_len;
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
  if (h->ppg_ready) cudaStreamWaitEvent(stream, h->ppg_ready, 0);  // fsvc_forward_host: uploads on the copy stream (the
                                                                    // PPG tensor is the last of them)
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


// AUTO / TC_BF16X3 run the channels-last tensor-core forward when the configuration allows it (every
// mid channel count a multiple of 8); otherwise TC_BF16X3 uses the per-layer tcgen05 kernel and AUTO fp32.
static bool use_tc2(const fsvc_handle* h, int mode) { return mode != FSVC_MODE_FP32 && h->tc2_ok; }

}  // namespace fsvc

using namespace fsvc;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int fsvc_abi_version(void) { return FSVC_ABI_VERSION; }
const char* fsvc_last_error(void) { return g_err; }

static int build_weight_jobs(fsvc_handle* h);

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
    size_t to = reserve((size_t)co * ci * K);
    fix.push_back({&cw, wo});
    cw.b = (float*)bo;  // offsets, patched below
    cw.wT = (float*)to;
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
    f.first->wT = h->store + (size_t)f.first->wT;
  }
  for (auto& f : fixp) *f.first = h->store + f.second;
  // tensor-core weight copies (tc_forward.cu)
  const size_t tc_off = tc_plan_handle(h);
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
  tc_fix_pointers(h);
  if (int rc = build_weight_jobs(h)) {
    fsvc_destroy(h);
    return rc;
  }
  if (int rc = tc_setup_kernels()) {
    cudaFree(h->store);
    cudaFree(h->tc_store);
    delete h;
    return rc;
  }
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_side_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_side_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_out_half, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_out_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sig[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sig[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sig[2], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sig[3], cudaEventDisableTiming) != cudaSuccess ||
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
  if (h->jobs_a) cudaFree(h->jobs_a);
  if (h->jobs_t) cudaFree(h->jobs_t);
  if (h->jobs_tc) cudaFree(h->jobs_tc);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_ppg) cudaEventDestroy(h->ev_ppg);
  for (int i = 0; i < fsvc_handle::kSigParts; ++i)
    if (h->ev_sig[i]) cudaEventDestroy(h->ev_sig[i]);
  if (h->ev_out_half) cudaEventDestroy(h->ev_out_half);
  if (h->ev_out_done) cudaEventDestroy(h->ev_out_done);
  if (h->ev_side_fork) cudaEventDestroy(h->ev_side_fork);
  if (h->ev_side_join) cudaEventDestroy(h->ev_side_join);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  delete h;
}

int fsvc_num_weight_tensors(const fsvc_handle* h) { return h ? (int)h->winfo.size() : FSVC_E_INVALID; }

int fsvc_weight_tensor_info(const fsvc_handle* h, int index, char* name, int name_capacity, int64_t* numel) {
  if (!h || index < 0 || index >= (int)h->winfo.size()) return fail(FSVC_E_INVALID, "bad weight index %d", index);
  if (name && name_capacity > 0) snprintf(name, name_capacity, "%s", h->winfo[index].name.c_str());
  if (numel) *numel = h->winfo[index].numel;
  return FSVC_OK;
}

// Device job tables of fsvc_set_weights (built once per handle; see WJob in fsvc_internal.h).
static int build_weight_jobs(fsvc_handle* h) {
  const int n = (int)h->winfo.size();
  if (n > kMaxWeightTensors) return fail(FSVC_E_INVALID, "too many weight tensors (%d)", n);
  std::vector<WJob> ja, jt;
  auto find = [&](const std::string& name) {
    for (int i = 0; i < n; ++i)
      if (h->winfo[i].name == name) return i;
    return -1;
  };
  auto repack_job = [&](int src, int co, int ci, int K, float* dst, int dst_ld, int ci_off, int co_off) {
    WJob j;
    memset(&j, 0, sizeof(j));
    j.kind = 0; j.src_a = src; j.src_b = -1; j.C_out = co; j.C_in = ci; j.K = K;
    j.dst = dst; j.dst_ld = dst_ld; j.ci_off = ci_off; j.co_off = co_off; j.total = (long long)co * ci * K;
    ja.push_back(j);
  };
  auto bias_job = [&](int a, int b, int cnt, float* dst, int off) {
    WJob j;
    memset(&j, 0, sizeof(j));
    j.kind = 1; j.src_a = a; j.src_b = b; j.dst = dst; j.co_off = off; j.total = cnt;
    ja.push_back(j);
  };
  auto copy_job = [&](int a, long long cnt, float* dst) {
    WJob j;
    memset(&j, 0, sizeof(j));
    j.kind = 2; j.src_a = a; j.src_b = -1; j.dst = dst; j.total = cnt;
    ja.push_back(j);
  };
  int k = 0;
  auto put = [&](ConvW& cw) {
    repack_job(k, cw.C_out, cw.C_in, cw.K, cw.w, cw.C_out, 0, 0);
    bias_job(k + 1, -1, cw.C_out, cw.b, 0);
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
      copy_job(k, (long long)C * h->cfg.spk_emb_size, st.emb_w);
      copy_job(k + 1, C, st.emb_b);
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
      repack_job(k, C, C, 3, lw.film_out.w, 2 * C, br * C, 0);
      repack_job(k + 2, C, C, 3, lw.film_out.w, 2 * C, br * C, C);
      k += 4;
    }
  // merged FiLM biases: gamma bias = b_scale_lft + b_scale_sine, beta bias likewise
  for (int l = 0; l < ns; ++l) {
    const int C = h->lvl_c[l];
    const std::string sl = std::to_string(l);
    bias_job(find("film_lft." + sl + ".conv_scale.bias"), find("film_sine." + sl + ".conv_scale.bias"), C,
             h->level[l].film_out.b, 0);
    bias_job(find("film_lft." + sl + ".conv_shift.bias"), find("film_sine." + sl + ".conv_shift.bias"), C,
             h->level[l].film_out.b, C);
  }
  put(h->last);
  if (k != n) return fail(FSVC_E_STATE, "internal: consumed %d of %d weight tensors", k, n);
  for (ConvW* cw : h->convs) {  // the transposed (data-gradient) copy of every conv, for the native backward
    WJob j;
    memset(&j, 0, sizeof(j));
    j.kind = 3; j.C_out = cw->C_out; j.C_in = cw->C_in; j.K = cw->K; j.w = cw->w; j.dst = cw->wT;
    j.total = (long long)cw->C_out * cw->C_in * cw->K;
    jt.push_back(j);
  }
  h->n_jobs_a = (int)ja.size();
  h->n_jobs_t = (int)jt.size();
  FSVC_CUDA(cudaMalloc((void**)&h->jobs_a, ja.size() * sizeof(WJob)));
  FSVC_CUDA(cudaMalloc((void**)&h->jobs_t, jt.size() * sizeof(WJob)));
  FSVC_CUDA(cudaMemcpy(h->jobs_a, ja.data(), ja.size() * sizeof(WJob), cudaMemcpyHostToDevice));
  FSVC_CUDA(cudaMemcpy(h->jobs_t, jt.data(), jt.size() * sizeof(WJob), cudaMemcpyHostToDevice));
  return tc_build_jobs(h);
}

int fsvc_set_weights(fsvc_handle* h, const float* const* p, int n, void* stream_) {
  if (!h || !p) return fail(FSVC_E_INVALID, "null argument");
  if (n != (int)h->winfo.size()) return fail(FSVC_E_INVALID, "expected %d weight tensors, got %d", (int)h->winfo.size(), n);
  for (int i = 0; i < n; ++i)
    if (!p[i]) return fail(FSVC_E_INVALID, "weight tensor %d (%s) is null", i, h->winfo[i].name.c_str());
  cudaStream_t s = (cudaStream_t)stream_;
  WSrc src;
  memset(&src, 0, sizeof(src));
  for (int i = 0; i < n; ++i) src.p[i] = p[i];
  // three launches: caller tensors -> packed fp32 store; packed -> transposed copies; packed -> tensor-core layouts
  weight_jobs_kernel<<<dim3(16, h->n_jobs_a), 256, 0, s>>>(h->jobs_a, src);
  weight_jobs_kernel<<<dim3(16, h->n_jobs_t), 256, 0, s>>>(h->jobs_t, src);
  tc_pack_weights(h, s);
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
  if (use_tc2(h, mode)) return tc_workspace_bytes(h, B, frames);
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
  // All uploads go, in the order the forward needs them, on the copy stream (which starts after everything already
  // queued on the caller's stream, so the staging buffers of a previous call are never overwritten early): the two
  // signals in four batch parts, the speaker vectors, the PPG tensor.  The forward waits for each piece where it first
  // reads it: the full-rate conditioning level runs per part (a part computes while the next ones are on the wire), the
  // speaker projections and stage 0 wait for the rest.
  // (measured at config 2: 1.400 ms without either overlap, 1.359 with two signal halves, 1.350 with the early waveform
  //  half as well; on the final build 2 / 3 / 4 parts: 1.1505 / 1.147 / 1.141 ms -- the first kernel starts after 1 MB
  //  instead of 2 MB of upload)
  static const int want_parts = getenv("FSVC_SIG_PARTS") ? atoi(getenv("FSVC_SIG_PARTS")) : fsvc_handle::kSigParts;
  int n_parts = want_parts < 1 ? 1 : (want_parts > fsvc_handle::kSigParts ? fsvc_handle::kSigParts : want_parts);
  if (n_parts > B) n_parts = B;
  FSVC_CUDA(cudaEventRecord(h->ev_fork, s));
  FSVC_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_fork, 0));
  for (int k = 0; k <= n_parts; ++k) h->sig_bounds[k] = (int)((long long)k * B / n_parts);
  for (int k = 0; k < n_parts; ++k) {
    const size_t o = (size_t)h->sig_bounds[k] * T, n = (size_t)(h->sig_bounds[k + 1] - h->sig_bounds[k]) * T;
    FSVC_CUDA(cudaMemcpyAsync(d_sine + o, sine_host + o, n * 4, cudaMemcpyHostToDevice, h->copy_stream));
    FSVC_CUDA(cudaMemcpyAsync(d_lft + o, lft_host + o, n * 4, cudaMemcpyHostToDevice, h->copy_stream));
    FSVC_CUDA(cudaEventRecord(h->ev_sig[k], h->copy_stream));
  }
  if (spk_host) FSVC_CUDA(cudaMemcpyAsync(d_spk, spk_host, n_spk * 4, cudaMemcpyHostToDevice, h->copy_stream));
  FSVC_CUDA(cudaMemcpyAsync(d_ppg, ppg_host, n_ppg * 4, cudaMemcpyHostToDevice, h->copy_stream));
  FSVC_CUDA(cudaEventRecord(h->ev_ppg, h->copy_stream));
  h->ppg_ready = h->ev_ppg;
  for (int k = 0; k < n_parts; ++k) h->sig_ready[k] = h->ev_sig[k];
  h->sig_parts = n_parts;
  h->out_host = out_host;
  h->out_host_done = 0;
  rc = fsvc_forward(h, d_ppg, d_sine, d_lft, spk_host ? d_spk : nullptr, d_out, B, frames, (char*)workspace + ar.off,
                    workspace_bytes - ar.off, mode, stream_);
  h->ppg_ready = nullptr;
  for (int k = 0; k < fsvc_handle::kSigParts; ++k) h->sig_ready[k] = nullptr;
  h->sig_parts = 0;
  h->out_host = nullptr;
  const size_t done = h->out_host_done;  // floats the forward already sent back on the copy stream
  h->out_host_done = 0;
  if (rc) {
    cudaStreamWaitEvent(s, h->ev_ppg, 0);  // rejoin the copy stream even when the forward was not enqueued
    return rc;
  }
  FSVC_CUDA(cudaMemcpyAsync(out_host + done, d_out + done, (n_out - done) * 4, cudaMemcpyDeviceToHost, s));
  if (done) FSVC_CUDA(cudaStreamWaitEvent(s, h->ev_out_done, 0));  // the caller synchronises `s` only
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
  if (ar.ok()) {
    repack(s, w, co, ci, K, cw.w, co, 0, 0);
    bias_sum(s, b, nullptr, co, cw.b, 0);
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


int fsvc_last_launch_count(const fsvc_handle* h) { return h ? h->launches : FSVC_E_INVALID; }

}  // extern "C"
