// Training path of libfsvc.so: a forward that keeps every activation the backward needs, and the native backward of
// the FastSVC generator (all parameters; inputs need no gradient -- train_fastsvc.py:168,199-206).  fp32 throughout, in
// the reference's (B, C, T) layout; kernels: conv_f32.cuh (forward convs and data gradients), train_kernels.cuh.
//
// Gradients are returned for the EFFECTIVE conv weights in PyTorch layout, in the canonical order of
// fsvc_weight_tensor_info; the host chains them through weight norm (g * v / ||v||) with stock autograd.
#include "conv_f32_launch.cuh"
#include "train_kernels.cuh"

namespace fsvc {

struct Saved {  // written by fsvc_forward_train, read by fsvc_backward
  // conditioning level l, branch br (0 = lft, 1 = sine)
  float* a1[2][FSVC_MAX_STAGES];  // Conv3_d1(lrelu(dec(src)))           [B][C_l][T_l]
  float* a2[2][FSVC_MAX_STAGES];  // Conv3_d2(lrelu(a1))
  float* y[2][FSVC_MAX_STAGES];   // level output
  float* H[FSVC_MAX_STAGES];      // lrelu(film.conv(y)), both branches  [B][2C_l][T_l]
  float* GB[FSVC_MAX_STAGES];     // gamma | beta (summed)               [B][2C_l][T_l]
  // stage i
  float* e[FSVC_MAX_STAGES];      // projected speaker embedding [B][C]
  float* h0[FSVC_MAX_STAGES];     // conv_first(x)                       [B][C][T_in]
  float* p[FSVC_MAX_STAGES];      // up conv output before its LeakyReLU [B][C][T]
  float* t1[FSVC_MAX_STAGES];     // gamma*lrelu(p) + beta
  float* x_[FSVC_MAX_STAGES];     // Conv3_d3(A1) + xr
  float* t2[FSVC_MAX_STAGES];     // gamma*x_ + beta
  float* x2[FSVC_MAX_STAGES];     // Conv3_d9(A2)
  float* t3[FSVC_MAX_STAGES];     // gamma*x2 + beta
  float* xs[FSVC_MAX_STAGES];     // stage output
  float* pa[FSVC_MAX_STAGES][3];  // InstanceNorm affines of the three FA() applications [B][C]
  float* pc[FSVC_MAX_STAGES][3];
};

static size_t layout_saved(const fsvc_handle* h, int B, int frames, void* base, size_t cap, Saved* sv) {
  Arena ar(base, cap);
  const int n = h->n;
  int T_l = frames * h->hop;
  for (int l = 0; l < n; ++l) {
    T_l /= h->dscale[l];
    const size_t ne = (size_t)B * h->lvl_c[l] * T_l;
    for (int br = 0; br < 2; ++br) {
      sv->a1[br][l] = ar.get<float>(ne);
      sv->a2[br][l] = ar.get<float>(ne);
      sv->y[br][l] = ar.get<float>(ne);
    }
    sv->H[l] = ar.get<float>(2 * ne);
    sv->GB[l] = ar.get<float>(2 * ne);
  }
  int T_in = frames;
  for (int i = 0; i < n; ++i) {
    const int C = h->cfg.mid_channels[i], r = h->cfg.upsampling_scales[i];
    const size_t ne = (size_t)B * C * T_in * r;
    sv->e[i] = ar.get<float>((size_t)B * C);
    sv->h0[i] = ar.get<float>((size_t)B * C * T_in);
    sv->p[i] = ar.get<float>(ne);
    sv->t1[i] = ar.get<float>(ne);
    sv->x_[i] = ar.get<float>(ne);
    sv->t2[i] = ar.get<float>(ne);
    sv->x2[i] = ar.get<float>(ne);
    sv->t3[i] = ar.get<float>(ne);
    sv->xs[i] = ar.get<float>(ne);
    for (int k = 0; k < 3; ++k) {
      sv->pa[i][k] = ar.get<float>((size_t)B * C);
      sv->pc[i][k] = ar.get<float>((size_t)B * C);
    }
    T_in *= r;
  }
  return ar.off;
}

// largest activation of any level / stage, in elements (B * C * T)
static size_t max_act(const fsvc_handle* h, int B, int frames) {
  size_t m = 0;
  int T_l = frames * h->hop;
  for (int l = 0; l < h->n; ++l) {
    T_l /= h->dscale[l];
    const size_t ne = 2 * (size_t)B * h->lvl_c[l] * T_l;
    m = ne > m ? ne : m;
  }
  int T_in = frames;
  for (int i = 0; i < h->n; ++i) {
    T_in *= h->cfg.upsampling_scales[i];
    const size_t ne = (size_t)B * h->cfg.mid_channels[i] * T_in;
    m = ne > m ? ne : m;
  }
  const size_t in = (size_t)B * h->cfg.in_channels * frames;
  return in > m ? in : m;
}

// ---------------------------------------------------------------------------------------------------------------
// forward (keeps activations)
// ---------------------------------------------------------------------------------------------------------------
static int train_forward(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                         float* out, int B, int frames, Saved& sv, void* workspace, size_t ws_bytes,
                         cudaStream_t stream) {
  Arena ar(workspace, ws_bytes);
  const size_t ma = max_act(h, B, frames);
  float* tmp_r = ar.get<float>(ma);
  float* xr = ar.get<float>(ma);
  const int T = frames * h->hop;
  float2* stats = ar.get<float2>((size_t)ma / 32 + (size_t)B * 4096);
  if (!ar.ok()) return fail(FSVC_E_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ar.off, ws_bytes);
  Ctx c;
  c.stream = stream;
  c.B = B;
  c.slope = h->cfg.lrelu_slope;
  c.eps = h->cfg.in_eps;
  const int n = h->n;

  // conditioning chains (fastsvc.py:180-193, 220-232), computed once
  int T_prev = T, T_l = T;
  for (int l = 0; l < n; ++l) {
    const int s = h->dscale[l];
    T_l = T_prev / s;
    const LevelW& lw = h->level[l];
    const int C = h->lvl_c[l];
    for (int br = 0; br < 2; ++br) {
      const float* src = l == 0 ? (br == 0 ? lft : sine) : sv.y[br][l - 1];
      ConvArgs a = conv_args(c, lw.r1[br], src, T_prev, T_l, 1, tmp_r);
      a.down = s;
      launch_conv(c, a, 1);
      a = conv_args(c, lw.c1[br], src, T_prev, T_l, 1, sv.a1[br][l]);
      a.down = s;
      a.pre_lrelu = 1;
      launch_conv(c, a, 3);
      a = conv_args(c, lw.c2[br], sv.a1[br][l], T_l, T_l, 2, sv.a2[br][l]);
      a.pre_lrelu = 1;
      launch_conv(c, a, 3);
      a = conv_args(c, lw.c4[br], sv.a2[br][l], T_l, T_l, 4, sv.y[br][l]);
      a.pre_lrelu = 1;
      a.res = tmp_r;
      a.res_cs = T_l;
      a.res_bs = (long long)C * T_l;
      launch_conv(c, a, 3);
      a = conv_args(c, lw.film[br], sv.y[br][l], T_l, T_l, 1, sv.H[l] + (size_t)br * C * T_l);
      a.out_bs = 2LL * C * T_l;
      a.post_lrelu = 1;
      launch_conv(c, a, 3);
    }
    ConvArgs a = conv_args(c, lw.film_out, sv.H[l], T_l, T_l, 1, sv.GB[l]);
    launch_conv(c, a, 3);
    T_prev = T_l;
  }

  // upsampling stages (fastsvc.py:80-140)
  const float* x = ppg;
  int T_in = frames;
  for (int i = 0; i < n; ++i) {
    const StageW& w = h->stage[i];
    const int C = h->cfg.mid_channels[i], r = h->cfg.upsampling_scales[i];
    const int l = n - 1 - i, Ts = T_in * r;
    const float* gamma = sv.GB[l];
    const float* beta = sv.GB[l] + (size_t)C * Ts;
    const bool norm = spk != nullptr;
    if (norm) spk_project_kernel<<<B, 256, 0, stream>>>(spk, h->cfg.spk_emb_size, w.emb_w, w.emb_b, C, sv.e[i]);
    const int tile_len = conv_tile_len(Ts), n_tiles = (Ts + tile_len - 1) / tile_len;
    auto film = [&](ConvArgs& a) {
      a.gamma = gamma;
      a.beta = beta;
      a.gb_bs = 2LL * C * Ts;
      a.gb_cs = Ts;
      if (norm) {
        a.stats = stats;
        a.n_tiles = n_tiles;
      }
    };
    auto finalize = [&](int k) {
      if (!norm) return;
      const int BC = B * C;
      in_finalize_kernel<<<(BC + 127) / 128, 128, 0, stream>>>(stats, n_tiles, tile_len, Ts, BC, sv.e[i], c.eps,
                                                               sv.pa[i][k], sv.pc[i][k]);
    };
    auto pre = [&](ConvArgs& a, int k) {
      if (norm) {
        a.pre_a = sv.pa[i][k];
        a.pre_c = sv.pc[i][k];
      }
      a.pre_lrelu = 1;
    };
    const long long bs = (long long)C * Ts;
    ConvArgs a = conv_args(c, w.first, x, T_in, T_in, 1, sv.h0[i]);
    launch_conv(c, a, 3);
    a = conv_args(c, w.res, sv.h0[i], T_in, Ts, 1, xr);
    a.up = r;
    launch_conv(c, a, 3);
    a = conv_args(c, w.up, sv.h0[i], T_in, Ts, 1, sv.t1[i]);
    a.up = r;
    a.pre_lrelu = 1;
    a.raw = sv.p[i];
    a.raw_cs = Ts;
    a.raw_bs = bs;
    a.post_lrelu = 1;
    film(a);
    launch_conv(c, a, 3);
    finalize(0);
    a = conv_args(c, w.d3, sv.t1[i], Ts, Ts, 3, sv.t2[i]);
    pre(a, 0);
    a.res = xr;
    a.res_cs = Ts;
    a.res_bs = bs;
    a.raw = sv.x_[i];
    a.raw_cs = Ts;
    a.raw_bs = bs;
    film(a);
    launch_conv(c, a, 3);
    finalize(1);
    a = conv_args(c, w.d9, sv.t2[i], Ts, Ts, 9, sv.t3[i]);
    pre(a, 1);
    a.raw = sv.x2[i];
    a.raw_cs = Ts;
    a.raw_bs = bs;
    film(a);
    launch_conv(c, a, 3);
    finalize(2);
    a = conv_args(c, w.d27, sv.t3[i], Ts, Ts, 27, sv.xs[i]);
    pre(a, 2);
    a.res = sv.x_[i];
    a.res_cs = Ts;
    a.res_bs = bs;
    launch_conv(c, a, 3);
    x = sv.xs[i];
    T_in = Ts;
  }
  ConvArgs a = conv_args(c, h->last, x, T, T, 1, out);
  launch_conv(c, a, 1);
  h->launches = c.launches;
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
struct Bwd {
  fsvc_handle* h;
  Ctx c;
  Arena* ar;
  float* const* grads;  // canonical order (fsvc_weight_tensor_info)
  int gi(const std::string& name) const {
    for (size_t i = 0; i < h->winfo.size(); ++i)
      if (h->winfo[i].name == name) return (int)i;
    return -1;
  }
  ReduceBatch rb;
  int n_jobs = 0;
  int err = 0;

  void flush() {
    if (!n_jobs) return;
    int max_total = 0;
    for (int i = 0; i < n_jobs; ++i) {
      const int t = rb.j[i].nco * rb.j[i].nci * rb.j[i].K;
      max_total = t > max_total ? t : max_total;
    }
    int bx = (max_total + 255) / 256;
    bx = bx > 64 ? 64 : bx;
    wgrad_reduce_kernel<<<dim3(bx, n_jobs), 256, 0, c.stream>>>(rb);
    c.launches++;
    n_jobs = 0;
  }
  void job(const float* part, float* dst, int n_split, int rows_tot, int row_len, int K, int ci0, int nci, int co0,
           int nco) {
    if (n_jobs == kReduceJobs) flush();
    ReduceJob& j = rb.j[n_jobs++];
    j.part = part;
    j.dst = dst;
    j.n_split = n_split;
    j.rows_tot = rows_tot;
    j.row_len = row_len;
    j.K = K;
    j.ci0 = ci0;
    j.nci = nci;
    j.co0 = co0;
    j.nco = nco;
  }

  struct Part {
    float* w;
    float* b;
    int n_split;
  };
  // weight-gradient partials of one conv: operand rebuilt from x (index map / affine / lrelu), output gradient g
  Part wgrad(const ConvW& w, const float* x, int x_T, int up, int down, const float* pre_a, const float* pre_c,
             int pre_lrelu, const float* g, long long g_bs, int T, int dil) {
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x;
    a.x_cs = x_T;
    a.x_bs = (long long)w.C_in * x_T;
    a.C_in = w.C_in;
    a.up = up;
    a.down = down;
    a.pre_a = pre_a;
    a.pre_c = pre_c;
    a.pre_lrelu = pre_lrelu;
    a.g = g;
    a.g_bs = g_bs;
    a.g_cs = T;
    a.C_out = w.C_out;
    a.T = T;
    a.B = c.B;
    a.dil = dil;
    a.slope = c.slope;
    const int by = (w.C_out + kWgCo - 1) / kWgCo, bz = (w.C_in + kWgCi - 1) / kWgCi;
    const int n_chunks = c.B * ((T + kWgTT - 1) / kWgTT);
    int n_split = (3 * h->num_sms + by * bz - 1) / (by * bz);
    n_split = n_split > n_chunks ? n_chunks : n_split;
    n_split = n_split < 1 ? 1 : n_split;
    Part p;
    p.n_split = n_split;
    p.w = ar->get<float>((size_t)n_split * w.C_in * w.K * w.C_out);
    p.b = ar->get<float>((size_t)n_split * w.C_out);
    if (!ar->ok()) {
      err = 1;
      return p;
    }
    a.part_w = p.w;
    a.part_b = p.b;
    const size_t smem = wgrad_smem_bytes(w.K, dil);
    if (w.K == 3) conv_wgrad_kernel<3><<<dim3(n_split, by, bz), kWgThreads, smem, c.stream>>>(a);
    else conv_wgrad_kernel<1><<<dim3(n_split, by, bz), kWgThreads, smem, c.stream>>>(a);
    c.launches++;
    return p;
  }
  // a plain conv whose parameters are "<prefix>.weight" / "<prefix>.bias"
  void wgrad_to(const std::string& prefix, const ConvW& w, const Part& p) {
    const int iw = gi(prefix + ".weight"), ib = gi(prefix + ".bias");
    if (iw < 0 || ib < 0) {
      err = 2;
      return;
    }
    job(p.w, grads[iw], p.n_split, w.C_in * w.K, w.C_out, w.K, 0, w.C_in, 0, w.C_out);
    job(p.b, grads[ib], p.n_split, 1, w.C_out, 1, 0, 1, 0, w.C_out);
  }
  // data gradient: the transposed conv on conv1d_f32_kernel
  ConvArgs dgrad_args(const ConvW& w, const float* g, long long g_bs, int T, int dil, float* out) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.in = g;
    a.in_cs = T;
    a.in_bs = g_bs;
    a.C_in = w.C_out;
    a.up = 1;
    a.down = 1;
    a.mask_up = 1;
    a.mask_down = 1;
    a.w = w.wT;
    a.bias = nullptr;
    a.dil = dil;
    a.C_out = w.C_in;
    a.T_out = T;
    a.out = out;
    a.out_cs = T;
    a.out_bs = (long long)w.C_in * T;
    a.slope = c.slope;
    return a;
  }
};

static size_t backward_scratch(const fsvc_handle* h, int B, int frames) {
  // 6 activation-sized buffers + g(gamma|beta) of every level + weight-gradient partials
  const size_t ma = max_act(h, B, frames);
  size_t total = 8 * ((ma * 4 + 255) & ~(size_t)255);
  int T_l = frames * h->hop;
  for (int l = 0; l < h->n; ++l) {
    T_l /= h->dscale[l];
    total += (2 * (size_t)B * h->lvl_c[l] * T_l * 4 + 255) & ~(size_t)255;
  }
  for (int i = 0; i < h->n; ++i) total += ((size_t)B * h->cfg.mid_channels[i] * 4 + 255) & ~(size_t)255;
  // partials: n_split <= 3*SMs/(tiles) + 1 per conv, i.e. at most ~ (3*SMs*128 + C_in*K*C_out) floats per conv
  for (const ConvW* cw : h->convs) {
    const size_t by = (cw->C_out + kWgCo - 1) / kWgCo, bz = (cw->C_in + kWgCi - 1) / kWgCi;
    const size_t n_split = (3 * (size_t)h->num_sms + by * bz - 1) / (by * bz);
    total += ((n_split * cw->C_in * cw->K * cw->C_out * 4 + 255) & ~(size_t)255) +
             ((n_split * cw->C_out * 4 + 255) & ~(size_t)255);
  }
  return total + (1 << 16);
}

static int train_backward(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                          const float* g_out, int B, int frames, const Saved& sv, float* const* grads, void* workspace,
                          size_t ws_bytes, cudaStream_t stream) {
  Arena ar(workspace, ws_bytes);
  Bwd bw;
  bw.h = h;
  bw.ar = &ar;
  bw.grads = grads;
  Ctx& c = bw.c;
  c.stream = stream;
  c.B = B;
  c.slope = h->cfg.lrelu_slope;
  c.eps = h->cfg.in_eps;
  const int n = h->n;
  const int T = frames * h->hop;
  const size_t ma = max_act(h, B, frames);
  float* bufA = ar.get<float>(ma);
  float* bufB = ar.get<float>(ma);
  float* bufC = ar.get<float>(ma);
  float* bufG = ar.get<float>(ma);  // gradient w.r.t. the current stage's output
  float* bufE = ar.get<float>(ma);  // gradient w.r.t. h0
  float* bufS[2] = {ar.get<float>(ma / 2 + 64), ar.get<float>(ma / 2 + 64)};  // grad w.r.t. dec(y[l-1]) per branch
  float* gGB[FSVC_MAX_STAGES];
  float* ge[FSVC_MAX_STAGES];
  {
    int T_l = T;
    for (int l = 0; l < n; ++l) {
      T_l /= h->dscale[l];
      gGB[l] = ar.get<float>(2 * (size_t)B * h->lvl_c[l] * T_l);
    }
    for (int i = 0; i < n; ++i) ge[i] = ar.get<float>((size_t)B * h->cfg.mid_channels[i]);
  }
  if (!ar.ok()) return fail(FSVC_E_WORKSPACE, "workspace too small: need > %zu bytes, got %zu", ar.off, ws_bytes);
  const bool norm = spk != nullptr;
  auto nblk = [](long long nn) { return (unsigned)((nn + 255) / 256); };

  // ---- conv_last (fastsvc.py:330) ----
  {
    const int C = h->cfg.mid_channels[n - 1], Co = h->cfg.out_channels;
    Bwd::Part p = bw.wgrad(h->last, sv.xs[n - 1], T, 1, 1, nullptr, nullptr, 0, g_out, (long long)Co * T, T, 1);
    bw.wgrad_to("conv_last", h->last, p);
    ConvArgs a = bw.dgrad_args(h->last, g_out, (long long)Co * T, T, 1, bufG);
    launch_conv(c, a, 1);
    (void)C;
  }

  // ---- upsampling stages, last to first (fastsvc.py:80-140) ----
  int Ts = T;
  for (int i = n - 1; i >= 0; --i) {
    const StageW& w = h->stage[i];
    const int C = h->cfg.mid_channels[i], r = h->cfg.upsampling_scales[i];
    const int l = n - 1 - i, T_in = Ts / r;
    const long long bs = (long long)C * Ts;
    const std::string pfx = "upsampling_nets." + std::to_string(i);
    const float* x_in = i == 0 ? ppg : sv.xs[i - 1];
    const float* pa[3] = {norm ? sv.pa[i][0] : nullptr, norm ? sv.pa[i][1] : nullptr, norm ? sv.pa[i][2] : nullptr};
    const float* pc[3] = {norm ? sv.pc[i][0] : nullptr, norm ? sv.pc[i][1] : nullptr, norm ? sv.pc[i][2] : nullptr};
    auto film_bwd = [&](const float* gA, const float* t, int k, const float* v, int v_lrelu, const float* add,
                        float* g_v, int first) {
      FilmBwdArgs f;
      memset(&f, 0, sizeof(f));
      f.gA = gA;
      f.t = t;
      f.pa = pa[k];
      f.pc = pc[k];
      f.e = norm ? sv.e[i] : nullptr;
      f.gamma = sv.GB[l];
      f.gb_bs = 2LL * C * Ts;
      f.v = v;
      f.v_lrelu = v_lrelu;
      f.add = add;
      f.g_v = g_v;
      f.g_gamma = gGB[l];
      f.g_beta = gGB[l] + (size_t)C * Ts;
      f.gb_accum = !first;
      f.g_e = norm ? ge[i] : nullptr;
      f.ge_accum = !first;
      f.C = C;
      f.T = Ts;
      f.slope = c.slope;
      film_in_bwd_kernel<<<dim3(C, B), 256, 0, stream>>>(f);
      c.launches++;
    };
    // out = Conv3_d27(A3) + x_ ;  A3 = lrelu(FA(x2))                                  :107-111
    Bwd::Part p27 = bw.wgrad(w.d27, sv.t3[i], Ts, 1, 1, pa[2], pc[2], 1, bufG, bs, Ts, 27);
    bw.wgrad_to(pfx + ".conv_block3.1", w.d27, p27);
    ConvArgs a = bw.dgrad_args(w.d27, bufG, bs, Ts, 27, bufA);
    launch_conv(c, a, 3);
    film_bwd(bufA, sv.t3[i], 2, sv.x2[i], 0, nullptr, bufB, 1);  // bufB = g_x2
    // x2 = Conv3_d9(A2) ;  A2 = lrelu(FA(x_))                                         :105-106
    Bwd::Part p9 = bw.wgrad(w.d9, sv.t2[i], Ts, 1, 1, pa[1], pc[1], 1, bufB, bs, Ts, 9);
    bw.wgrad_to(pfx + ".conv_block2.1", w.d9, p9);
    a = bw.dgrad_args(w.d9, bufB, bs, Ts, 9, bufA);
    launch_conv(c, a, 3);
    film_bwd(bufA, sv.t2[i], 1, sv.x_[i], 0, bufG, bufC, 0);  // bufC = g_x_ = g_out + g_t2*gamma
    // x_ = Conv3_d3(A1) + xr ;  A1 = lrelu(FA(lrelu(p)))                               :97-102
    Bwd::Part p3 = bw.wgrad(w.d3, sv.t1[i], Ts, 1, 1, pa[0], pc[0], 1, bufC, bs, Ts, 3);
    bw.wgrad_to(pfx + ".conv_block1.1", w.d3, p3);
    a = bw.dgrad_args(w.d3, bufC, bs, Ts, 3, bufA);
    launch_conv(c, a, 3);
    film_bwd(bufA, sv.t1[i], 0, sv.p[i], 1, nullptr, bufB, 0);  // bufB = g_p
    // p = Conv3(repeat(lrelu(h0))) ;  xr = Conv3(repeat(h0))                           :94, :97
    Bwd::Part pu = bw.wgrad(w.up, sv.h0[i], T_in, r, 1, nullptr, nullptr, 1, bufB, bs, Ts, 1);
    bw.wgrad_to(pfx + ".upsample_block0.2", w.up, pu);
    Bwd::Part pr = bw.wgrad(w.res, sv.h0[i], T_in, r, 1, nullptr, nullptr, 0, bufC, bs, Ts, 1);
    bw.wgrad_to(pfx + ".residual_block.1", w.res, pr);
    a = bw.dgrad_args(w.up, bufB, bs, Ts, 1, bufA);
    a.mask = sv.h0[i];  // derivative of the lrelu in front of the repeat
    a.mask_bs = (long long)C * T_in;
    a.mask_cs = T_in;
    a.mask_up = r;
    launch_conv(c, a, 3);
    a = bw.dgrad_args(w.res, bufC, bs, Ts, 1, bufB);
    a.res = bufA;
    a.res_cs = Ts;
    a.res_bs = bs;
    launch_conv(c, a, 3);
    fold_repeat_kernel<<<nblk((long long)B * C * T_in), 256, 0, stream>>>(bufB, bufE, (long long)B * C * T_in, r);
    c.launches++;
    // h0 = conv_first(x)                                                               :93
    Bwd::Part pf = bw.wgrad(w.first, x_in, T_in, 1, 1, nullptr, nullptr, 0, bufE, (long long)C * T_in, T_in, 1);
    bw.wgrad_to(pfx + ".conv_first", w.first, pf);
    if (i > 0) {
      a = bw.dgrad_args(w.first, bufE, (long long)C * T_in, T_in, 1, bufG);
      launch_conv(c, a, 3);
    }
    if (norm) {  // emb_projector(F.normalize(spk))                                     :135-137
      const int iw = bw.gi(pfx + ".emb_projector.weight"), ib = bw.gi(pfx + ".emb_projector.bias");
      if (iw < 0 || ib < 0) return fail(FSVC_E_STATE, "internal: emb_projector gradient slot missing");
      spk_bwd_kernel<<<C, 256, B * sizeof(float), stream>>>(spk, h->cfg.spk_emb_size, B, ge[i], C, grads[iw], grads[ib]);
      c.launches++;
    } else if (h->cfg.use_spk_emb) {
      const int iw = bw.gi(pfx + ".emb_projector.weight"), ib = bw.gi(pfx + ".emb_projector.bias");
      FSVC_CUDA(cudaMemsetAsync(grads[iw], 0, (size_t)C * h->cfg.spk_emb_size * 4, stream));
      FSVC_CUDA(cudaMemsetAsync(grads[ib], 0, (size_t)C * 4, stream));
    }
    Ts = T_in;
  }

  // ---- conditioning levels, coarsest to finest (fastsvc.py:180-193, 220-232) ----
  const char* dn[2] = {"downsampling_lft.", "downsampling_sine."};
  const char* fn[2] = {"film_lft.", "film_sine."};
  int T_lv[FSVC_MAX_STAGES];
  {
    int t = T;
    for (int l = 0; l < n; ++l) {
      t /= h->dscale[l];
      T_lv[l] = t;
    }
  }
  for (int l = n - 1; l >= 0; --l) {
    const LevelW& lw = h->level[l];
    const int C = h->lvl_c[l], T_l = T_lv[l], s = h->dscale[l];
    const int T_prev = T_l * s;
    const std::string sl = std::to_string(l);
    // [gamma | beta] = film_out([h_lft | h_sine])
    Bwd::Part po = bw.wgrad(lw.film_out, sv.H[l], T_l, 1, 1, nullptr, nullptr, 0, gGB[l], 2LL * C * T_l, T_l, 1);
    for (int br = 0; br < 2; ++br) {
      const int is = bw.gi(fn[br] + sl + ".conv_scale.weight"), ih = bw.gi(fn[br] + sl + ".conv_shift.weight");
      const int ibs = bw.gi(fn[br] + sl + ".conv_scale.bias"), ibh = bw.gi(fn[br] + sl + ".conv_shift.bias");
      if (is < 0 || ih < 0 || ibs < 0 || ibh < 0) return fail(FSVC_E_STATE, "internal: FiLM gradient slot missing");
      bw.job(po.w, grads[is], po.n_split, 2 * C * 3, 2 * C, 3, br * C, C, 0, C);
      bw.job(po.w, grads[ih], po.n_split, 2 * C * 3, 2 * C, 3, br * C, C, C, C);
      bw.job(po.b, grads[ibs], po.n_split, 1, 2 * C, 1, 0, 1, 0, C);
      bw.job(po.b, grads[ibh], po.n_split, 1, 2 * C, 1, 0, 1, C, C);
    }
    ConvArgs a = bw.dgrad_args(lw.film_out, gGB[l], 2LL * C * T_l, T_l, 1, bufA);  // bufA = g wrt film.conv output
    a.mask = sv.H[l];
    a.mask_bs = 2LL * C * T_l;
    a.mask_cs = T_l;
    launch_conv(c, a, 3);
    for (int br = 0; br < 2; ++br) {
      const float* src = l == 0 ? (br == 0 ? lft : sine) : sv.y[br][l - 1];
      const int C_src = l == 0 ? 1 : h->lvl_c[l - 1];
      const float* gH = bufA + (size_t)br * C * T_l;
      const long long bsC = (long long)C * T_l;
      // h = lrelu(film.conv(y))
      Bwd::Part pfc = bw.wgrad(lw.film[br], sv.y[br][l], T_l, 1, 1, nullptr, nullptr, 0, gH, 2LL * C * T_l, T_l, 1);
      bw.wgrad_to(fn[br] + sl + ".conv", lw.film[br], pfc);
      a = bw.dgrad_args(lw.film[br], gH, 2LL * C * T_l, T_l, 1, bufB);  // bufB = g_y
      launch_conv(c, a, 3);
      if (l < n - 1) {  // + the decimated path into level l+1
        const long long nd = (long long)B * C * T_lv[l + 1];
        scatter_dec_kernel<<<nblk(nd), 256, 0, stream>>>(bufS[br], bufB, nd, h->dscale[l + 1]);
        c.launches++;
      }
      // y = Conv3_d4(lrelu(a2)) + Conv1x1(dec(src))
      Bwd::Part p4 = bw.wgrad(lw.c4[br], sv.a2[br][l], T_l, 1, 1, nullptr, nullptr, 1, bufB, bsC, T_l, 4);
      bw.wgrad_to(dn[br] + sl + ".downsample_block.6", lw.c4[br], p4);
      Bwd::Part pr1 = bw.wgrad(lw.r1[br], src, T_prev, 1, s, nullptr, nullptr, 0, bufB, bsC, T_l, 1);
      bw.wgrad_to(dn[br] + sl + ".residual_block.0", lw.r1[br], pr1);
      a = bw.dgrad_args(lw.c4[br], bufB, bsC, T_l, 4, bufC);  // bufC = g_a2
      a.mask = sv.a2[br][l];
      a.mask_bs = bsC;
      a.mask_cs = T_l;
      launch_conv(c, a, 3);
      Bwd::Part p2 = bw.wgrad(lw.c2[br], sv.a1[br][l], T_l, 1, 1, nullptr, nullptr, 1, bufC, bsC, T_l, 2);
      bw.wgrad_to(dn[br] + sl + ".downsample_block.4", lw.c2[br], p2);
      a = bw.dgrad_args(lw.c2[br], bufC, bsC, T_l, 2, bufE);  // bufE = g_a1
      a.mask = sv.a1[br][l];
      a.mask_bs = bsC;
      a.mask_cs = T_l;
      launch_conv(c, a, 3);
      Bwd::Part p1 = bw.wgrad(lw.c1[br], src, T_prev, 1, s, nullptr, nullptr, 1, bufE, bsC, T_l, 1);
      bw.wgrad_to(dn[br] + sl + ".downsample_block.2", lw.c1[br], p1);
      if (l > 0) {  // gradient w.r.t. dec(y[l-1]) of this branch, scattered into g_y[l-1] by the next iteration
        a = bw.dgrad_args(lw.r1[br], bufB, bsC, T_l, 1, bufC);
        launch_conv(c, a, 1);
        a = bw.dgrad_args(lw.c1[br], bufE, bsC, T_l, 1, bufS[br]);
        a.mask = src;
        a.mask_bs = (long long)C_src * T_prev;
        a.mask_cs = T_prev;
        a.mask_down = s;
        a.res = bufC;
        a.res_cs = T_l;
        a.res_bs = (long long)C_src * T_l;
        launch_conv(c, a, 3);
      }
    }
  }
  bw.flush();
  h->launches = c.launches;
  if (bw.err) return fail(bw.err == 1 ? FSVC_E_WORKSPACE : FSVC_E_STATE, "backward: %s",
                          bw.err == 1 ? "workspace too small for the weight-gradient partials" : "gradient slot missing");
  FSVC_CUDA(cudaGetLastError());
  return FSVC_OK;
}

}  // namespace fsvc

using namespace fsvc;

extern "C" {

size_t fsvc_train_saved_bytes(const fsvc_handle* h, int B, int frames) {
  if (!h || B < 1 || frames < 1) return 0;
  Saved sv;
  return layout_saved(h, B, frames, nullptr, 0, &sv);
}

size_t fsvc_train_workspace_bytes(const fsvc_handle* h, int B, int frames) {
  if (!h || B < 1 || frames < 1) return 0;
  const size_t ma = max_act(h, B, frames);
  const size_t fwd = 2 * ((ma * 4 + 255) & ~(size_t)255) + (((ma / 32 + (size_t)B * 4096) * 8 + 255) & ~(size_t)255);
  const size_t bwd = backward_scratch(h, B, frames);
  return (fwd > bwd ? fwd : bwd) + 4096;
}

int fsvc_forward_train(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                       float* out, int B, int frames, void* saved, size_t saved_bytes, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (!h) return fail(FSVC_E_INVALID, "null handle");
  if (B < 1 || frames < 1 || B > 65535) return fail(FSVC_E_INVALID, "bad B / frames (%d, %d)", B, frames);
  if (!ppg || !sine || !lft || !out || !saved || !workspace) return fail(FSVC_E_INVALID, "null tensor pointer");
  if (!h->weights_set) return fail(FSVC_E_STATE, "fsvc_forward_train called before fsvc_set_weights");
  if (spk && !h->cfg.use_spk_emb) return fail(FSVC_E_INVALID, "spk given but use_spk_emb=0");
  Saved sv;
  const size_t need = layout_saved(h, B, frames, saved, saved_bytes, &sv);
  if (need > saved_bytes) return fail(FSVC_E_WORKSPACE, "saved-activation buffer too small: need %zu, got %zu", need, saved_bytes);
  return train_forward(h, ppg, sine, lft, spk, out, B, frames, sv, workspace, workspace_bytes, (cudaStream_t)stream);
}

int fsvc_backward(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                  const float* grad_out, int B, int frames, const void* saved, size_t saved_bytes,
                  float* const* grad_ptrs, int n, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) return fail(FSVC_E_INVALID, "null handle");
  if (B < 1 || frames < 1 || B > 65535) return fail(FSVC_E_INVALID, "bad B / frames (%d, %d)", B, frames);
  if (!ppg || !sine || !lft || !grad_out || !saved || !grad_ptrs || !workspace) return fail(FSVC_E_INVALID, "null pointer");
  if (!h->weights_set) return fail(FSVC_E_STATE, "fsvc_backward called before fsvc_set_weights");
  if (n != (int)h->winfo.size()) return fail(FSVC_E_INVALID, "expected %d gradient tensors, got %d", (int)h->winfo.size(), n);
  for (int i = 0; i < n; ++i)
    if (!grad_ptrs[i]) return fail(FSVC_E_INVALID, "gradient tensor %d (%s) is null", i, h->winfo[i].name.c_str());
  if (spk && !h->cfg.use_spk_emb) return fail(FSVC_E_INVALID, "spk given but use_spk_emb=0");
  Saved sv;
  const size_t need = layout_saved(h, B, frames, const_cast<void*>(saved), saved_bytes, &sv);
  if (need > saved_bytes) return fail(FSVC_E_WORKSPACE, "saved-activation buffer too small: need %zu, got %zu", need, saved_bytes);
  return train_backward(h, ppg, sine, lft, spk, grad_out, B, frames, sv, grad_ptrs, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}

}  // extern "C"
