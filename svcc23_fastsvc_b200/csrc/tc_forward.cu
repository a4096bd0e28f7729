// Tensor-core forward of the FastSVC generator over channels-last activations: launch plan, workspace layout and the
// weight repacks of the tcgen05 kernels (conv_tc3.cuh, level_fused.cuh).  One translation unit of libfsvc.so.
#include "conv_tc3.cuh"
#include "level_fused.cuh"

namespace fsvc {

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
  int nat_cib = ((c16 + nat_blk - 1) / nat_blk + 15) / 16 * 16;
  // Multi-block layers: ci blocks of at most 32 channels.  With the row loader (conv_tc3.cuh) a block costs ~1 us to
  // convert whatever its width, and 32-channel blocks leave room for a double-buffered A ring and staging next to the
  // weight ring (measured at config 2: 48 -> 32 channels took the forward from 1.340 to 1.271 ms; 16 gave 1.331 ms).
  if (C_in > 64 && nat_cib > 32) nat_cib = 32;
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

int tc_setup_kernels() {
  const int max_smem = 227 * 1024;
#define FSVC_ATTR(K_, NH_)                                                                                        \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
  FSVC_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<K_, NH_, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem))
  FSVC_ATTR(3, 3);
  FSVC_ATTR(3, 0);
  FSVC_ATTR(1, 3);
  FSVC_ATTR(1, 0);
#undef FSVC_ATTR
  FSVC_CUDA(cudaFuncSetAttribute(level0_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
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
  float* ydec_l[2][FSVC_MAX_STAGES];  // [B][T_l/s][C_l] output of level l >= 1 decimated for level l + 1 (d4's epilogue)
};

// rows per (utterance, 8-channel group) of an operand-plane tensor: whole 128-step tiles + kPlPad zero rows at both ends
static inline int pl_rows(int T) { return (T + kTc2M - 1) / kTc2M * kTc2M + 2 * kPlPad; }

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
    // the chain temporaries and H may hold bf16 hi|lo operand planes instead (ntc_common.cuh): C * pl_rows floats per utterance
    const size_t ne_pl = (size_t)B * h->lvl_c[l] * pl_rows(T_l);
    max_lvl = ne > max_lvl ? ne : max_lvl;
    max_lvl = ne_pl > max_lvl ? ne_pl : max_lvl;
    for (int br = 0; br < 2; ++br) ws->y[br][l] = ar.get<float>(ne);
    for (int br = 0; br < 2; ++br)
      ws->ydec_l[br][l] = (l >= 1 && l + 1 < n) ? ar.get<float>((size_t)B * h->lvl_c[l] * ntc_tp(T_l / h->dscale[l + 1])) : nullptr;
    ws->H[l] = ar.get<float>(2 * (ne_pl > ne ? ne_pl : ne));
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
// `hetero`: the two problems share shapes and tiling but not prologue / epilogue (a stage's residual conv and its
// upsampling conv read the same h0): each gets its own argument block.
static void launch_tc2(Ctx& c, const fsvc_handle* h, int K, const Tc2Args* p, int n_prob, const char* name, int b_lo = 0,
                       int b_hi = -1, bool hetero = false) {
  if (b_hi < 0) b_hi = c.B;
  Tc3Launch L;
  memset(&L, 0, sizeof(L));
  L.a = p[0];
  if (n_prob == 2 && hetero) {
    const Tc2Args &x = p[0], &y = p[1];
    L.b = y;
    L.hetero = 1;
    if (x.up != y.up || x.down != y.down || x.dil != y.dil || x.C_in != y.C_in || x.C_out != y.C_out ||
        x.T_out != y.T_out || x.T_in != y.T_in || x.in_ld != y.in_ld || x.CIB != y.CIB || x.n_blk != y.n_blk ||
        x.N_tile != y.N_tile || x.n_ntiles != y.n_ntiles || x.w_resident != y.w_resident || x.gen_w || y.gen_w ||
        x.pre_stats || y.pre_stats || x.pre_a || y.pre_a || x.in_pl || y.in_pl || x.out_pl || y.out_pl || x.out_dec ||
        y.out_dec) {
      c.err = 2;
      return;
    }
  } else if (n_prob == 2) {
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
    L.d_in_pl = x.in_pl ? y.in_pl - x.in_pl : 0;
    L.d_out_pl = x.out_pl ? y.out_pl - x.out_pl : 0;
    L.d_out_dec = x.out_dec ? y.out_dec - x.out_dec : 0;
    // everything that is not a per-problem pointer must agree
    if ((x.up > 1 && x.down > 1) || x.pre_lrelu != y.pre_lrelu || x.post_lrelu != y.post_lrelu || x.gamma != y.gamma || x.stats != y.stats ||
        x.pre_a != y.pre_a || x.up != y.up || x.down != y.down || x.dil != y.dil || x.C_in != y.C_in ||
        x.C_out != y.C_out || x.T_out != y.T_out || x.T_in != y.T_in || x.in_ld != y.in_ld ||
        x.out_ld != y.out_ld || x.res_ld != y.res_ld || (x.res == nullptr) != (y.res == nullptr) ||
        (x.gen_w == nullptr) != (y.gen_w == nullptr) || (x.gres_w == nullptr) != (y.gres_w == nullptr) ||
        (x.in_pl == nullptr) != (y.in_pl == nullptr) || (x.out_pl == nullptr) != (y.out_pl == nullptr) ||
        x.in_pl_lo != y.in_pl_lo || x.out_pl_lo != y.out_pl_lo || x.out_pl_lrelu != y.out_pl_lrelu ||
        (x.out_dec == nullptr) != (y.out_dec == nullptr) || x.out_dec_r != y.out_dec_r) {
      c.err = 2;
      return;
    }
  }
  L.tl_slot = c.launches;
  Tc3Cfg& cfg = L.c;
  cfg.n_prob = n_prob;
  cfg.B = c.B;
  cfg.b_lo = b_lo;
  cfg.b_hi = b_hi;
  cfg.m_tiles = (p[0].T_out + kTc2M - 1) / kTc2M;
  const int groups = n_prob * p[0].n_ntiles;
  const int items = (b_hi - b_lo) * cfg.m_tiles;
  // transform variant: 3 / 4 = lean path (direct rows, one ci block, a warp's <= 4 / <= 6 tasks of an item in one chunk)
  int mode = p[0].gen_w ? 1 : (p[0].up > 1 ? 2 : 0);
  if (mode == 0 && p[0].down == 1 && !p[0].in_pl && !getenv("FSVC_NO_LEAN")) {
    const int ntask = (p[0].CIB / 8) * ((kTc2M + 2 * (K / 2) * p[0].dil + 31) / 32);
    const int rounds = (ntask + 5) / 6;
    const bool bulk = (K / 2) * p[0].dil <= 32;  // the window spans at most 6 aligned 32-row blocks
    if (bulk && rounds <= 4 && tc3_plan_smem(p[0], K, &cfg, 4, true)) mode = 3;
    else if (bulk && rounds <= 6 && tc3_plan_smem(p[0], K, &cfg, 6, true) && cfg.a_slots >= 2) mode = 4;
  }
  if (p[0].in_pl) {  // operand planes written by the producer: no transform role
    if (mode != 0 || p[0].down != 1 || (K / 2) * p[0].dil > kPlPad || p[0].C_in % 8 != 0 || p[0].C_in % p[0].CIB != 0 ||
        !tc3_plan_smem(p[0], K, &cfg, 0, false, true)) {
      c.err = 3;
      return;
    }
    mode = 6;
  }
  if (mode == 2 && p[0].down == 1 && p[0].up <= 8 && p[0].n_blk == 1 && !getenv("FSVC_NO_LEAN")) {
    const int Wd = kTc2M + 2 * (K / 2) * p[0].dil;
    const int ntask = (p[0].CIB / 8) * (((Wd + p[0].up - 1) / p[0].up + 1 + 31) / 32);
    if ((ntask + 5) / 6 <= 4 && tc3_plan_smem(p[0], K, &cfg, 4)) mode = 5;
  }
  if (mode < 3 && !tc3_plan_smem(p[0], K, &cfg)) {
    c.err = 1;
    return;
  }
  int per_group = h->num_sms / groups;
  per_group = per_group < 1 ? 1 : per_group;
  per_group = per_group > items ? items : per_group;
  const dim3 grid(per_group * groups);
  if (hetero && items >= 8 * per_group) {
    // many items per CTA: share the grid by cost.  An item with the FiLM affine + statistics epilogue costs ~1.45x a
    // plain one (1.6 vs 1.1 us per item in the last stage's timeline); measured against the equal split: 1.245 -> 1.229 ms
    const double w0 = p[0].gamma ? 1.45 : 1.0, w1 = p[1].gamma ? 1.45 : 1.0;
    const int total = 2 * per_group;  // CTAs per N tile over both problems
    int c0 = (int)(total * w0 / (w0 + w1) + 0.5);
    c0 = c0 < 1 ? 1 : (c0 > total - 1 ? total - 1 : c0);
    cfg.cta0 = c0;
  }
  if (c.prof && getenv("FSVC_DEBUG_PLAN"))
    fprintf(stderr, "plan %-4s %-12s Cin=%d Cout=%d K=%d dil=%d up=%d down=%d CIB=%d n_blk=%d N_tile=%d n_ntiles=%d resident=%d "
                    "a_slots=%d stg_depth=%d smem=%u grid=%u items=%d\n",
            c.label, name, p[0].C_in, p[0].C_out, K, p[0].dil, p[0].up, p[0].down, p[0].CIB, p[0].n_blk, p[0].N_tile,
            p[0].n_ntiles, p[0].w_resident, cfg.a_slots, cfg.stg_depth, cfg.total, grid.x, items);
  // compile-time epilogue width (12 channels per thread) when every sub-tile of every N tile is full
  const int per_thread = cfg.nsub / Tc3Shape::kEH;
  const bool nh3 = per_thread == 12 && p[0].C_out % cfg.nsub == 0 && (p[0].n_ntiles == 1 || p[0].N_tile % cfg.nsub == 0) &&
                   !p[0].gres_w;  // the generated residual lives in the run-time-width instantiation only
  const int threads = kTc3Threads;
#define FSVC_TC3(K_, NH_)                                                                              \
  do {                                                                                                      \
    if (mode == 1) launch_pdl(conv_tc3_kernel<K_, NH_, 1>, grid, threads, cfg.total, c.stream, L);      \
    else if (mode == 2) launch_pdl(conv_tc3_kernel<K_, NH_, 2>, grid, threads, cfg.total, c.stream, L); \
    else if (mode == 3) launch_pdl(conv_tc3_kernel<K_, NH_, 3>, grid, threads, cfg.total, c.stream, L); \
    else if (mode == 4) launch_pdl(conv_tc3_kernel<K_, NH_, 4>, grid, threads, cfg.total, c.stream, L); \
    else if (mode == 5) launch_pdl(conv_tc3_kernel<K_, NH_, 5>, grid, threads, cfg.total, c.stream, L); \
    else if (mode == 6) launch_pdl(conv_tc3_kernel<K_, NH_, 6>, grid, threads, cfg.total, c.stream, L); \
    else launch_pdl(conv_tc3_kernel<K_, NH_, 0>, grid, threads, cfg.total, c.stream, L);                \
  } while (0)
  if (K == 3) {
    if (nh3) FSVC_TC3(3, 3); else FSVC_TC3(3, 0);
  } else {
    if (nh3) FSVC_TC3(1, 3); else FSVC_TC3(1, 0);
  }
#undef FSVC_TC3
  double flops = 0, elems = 0;
  for (int i = 0; i < n_prob; ++i) {
    const Tc2Args& a = p[i];
    const double BT = (double)(b_hi - b_lo) * a.T_out;
    flops += 2.0 * a.C_in * a.C_out * K * BT;
    elems += a.gen_w ? BT : (double)(b_hi - b_lo) * a.C_in * ((double)a.T_out / a.up);
    elems += BT * a.C_out * ((a.out ? 1 : 0) + (a.out_pl ? 1 : 0) + (a.raw ? 1 : 0) + (a.res ? 1 : 0) + (a.gamma ? 2 : 0));
    if (a.out_dec) elems += BT * a.C_out / a.out_dec_r;
    if (a.last_w) {  // folded conv_last
      flops += 2.0 * BT * a.C_out * a.last_co;
      elems += BT * a.last_co + (double)a.C_out * a.last_co;
    }
    elems += (double)a.C_in * a.C_out * K;
  }
  c.launched(name, flops, 4.0 * elems);
}

int forward_tc2(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk, float* out,
                int B, int frames, void* workspace, size_t ws_bytes, cudaStream_t stream, Profiler* prof) {
  WS2 ws;
  // the kernels index (utterance, step) rows with 32 bits
  if ((long long)B * ntc_tp(frames * h->hop) >= (1ll << 31))
    return fail(FSVC_E_INVALID, "batch of %d x %d steps exceeds 2^31 rows", B, frames * h->hop);
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

  // conv_last folds into the last stage's final conv when that conv runs as one N tile / one sub-tile of two column
  // halves (any last-stage width that is a multiple of 8 up to 32): exactly two atomic adds per output sample
  const int C_last = h->cfg.mid_channels[n - 1];
  const bool fold_last = C_last <= 32 && h->stage[n - 1].d27.tc2.n_ntiles == 1 && !getenv("FSVC_NO_FOLD_LAST");
  // Small launches nothing on the conditioning chain depends on (speaker projections, PPG transpose, clears) run on a
  // forked stream and rejoin before stage 0; a profiling run keeps them on the caller's stream (per-launch events).
  cudaStream_t side = prof ? stream : h->side_stream;
  if (!prof) {
    cudaEventRecord(h->ev_side_fork, stream);
    cudaStreamWaitEvent(side, h->ev_side_fork, 0);
  }
  if (fold_last) cudaMemsetAsync(out, 0, (size_t)B * h->cfg.out_channels * T * sizeof(float), side);
  {  // the caller's (B, C, T') PPG tensor -> channels-last
    if (h->ppg_ready) cudaStreamWaitEvent(side, h->ppg_ready, 0);  // fsvc_forward_host: upload on the copy stream
    const int Cin = h->cfg.in_channels;
    nct_to_ntc_kernel<<<dim3((frames + 31) / 32, (Cin + 31) / 32, B), 256, 0, side>>>(ppg, Cin, frames, ws.xin);
    c.label = "";
    c.launched("ppg_to_ntc", 0.0, 8.0 * B * Cin * frames);
  }
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
    spk_project_all_kernel<<<dim3(B, n, (c_max + 31) / 32), 256, 0, side>>>(spk, S, sp);
    c.label = "";
    c.launched("spk_project", 0.0, 0.0);
  }
  if (!prof) cudaEventRecord(h->ev_side_join, side);

  // fsvc_forward_host: without the fused level-0 kernel (which takes the signals half by half) wait for both halves here
  if (!h->l0_fused && h->sig_parts > 0) cudaStreamWaitEvent(stream, h->sig_ready[h->sig_parts - 1], 0);
  // ---- conditioning chains, both branches per launch (fastsvc.py:180-193, 220-232) ----
  int T_prev = T, T_l = T;
  bool fused_l0 = false, dec_prev = false;
  for (int l = 0; l < n; ++l) {
    T_l = T_prev / h->dscale[l];
    const LevelW& lw = h->level[l];
    const int C = h->lvl_c[l];
    c.label = lvl_label[l];
    Tc2Args p[2];
    bool use_pl = false;
    const int pl_Tp = pl_rows(T_l);
    // this level's last conv also writes its output decimated for the next level
    const bool dec_out = l >= 1 && l + 1 < n && h->dscale[l + 1] > 1 && T_l % h->dscale[l + 1] == 0 && !getenv("FSVC_NO_DEC");
    // operand planes of this level: `buf` holds the hi plane [B][G][pl_Tp] chunks, then the lo plane
    auto planes_out = [&](Tc2Args& a, float* buf, int G, int g0, int lrelu) {
      a.out_pl = reinterpret_cast<uint4*>(buf) + (long long)g0 * pl_Tp;
      a.out_pl_lo = (long long)B * G * pl_Tp;
      a.out_pl_G = G;
      a.out_pl_Tp = pl_Tp;
      a.out_pl_lrelu = lrelu;
    };
    auto planes_in = [&](Tc2Args& a, const float* buf, int G) {
      a.in_pl = reinterpret_cast<const uint4*>(buf);
      a.in_pl_lo = (long long)B * G * pl_Tp;
      a.in_pl_G = G;
      a.in_pl_Tp = pl_Tp;
    };
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
      fa.dec = dec;
      fa.n_tiles = (T_l + kLfValid - 1) / kLfValid;
      fa.Gp = lw.c2[0].nc_G;
      fa.N1 = lw.c2[0].nc_N;
      fa.N2 = lw.film_out.nc_N;
      fa.slope = c.slope;
      if (level_fused_fill_desc(&fa) != 0) return fail(FSVC_E_INVALID, "internal: fused level descriptor table overflow");
      // fsvc_forward_host uploads the signals in batch parts: one launch per part, each behind its own event
      const int n_part = h->sig_parts > 0 ? h->sig_parts : 1;
      for (int part = 0; part < n_part; ++part) {
        fa.b_off = h->sig_parts > 0 ? h->sig_bounds[part] : 0;
        fa.B = h->sig_parts > 0 ? h->sig_bounds[part + 1] - h->sig_bounds[part] : B;
        if (h->sig_parts > 0) cudaStreamWaitEvent(stream, h->sig_ready[part], 0);
        const int items = fa.B * fa.n_tiles;
        const int grid = items < h->num_sms ? items : h->num_sms;
        launch_pdl(level0_fused_kernel, dim3(grid), kLfThreads, LF.total, stream, fa);
        if (part > 0) c.launches++;
      }
      const double BT = (double)B * T_l;
      // both branches: first conv (3) + 1x1 residual (1) + d2, d4, FiLM conv (9C) MACs per channel and step; merged
      // film_out: 2C -> 2C, k = 3 (12C per channel); SURVEY 8d: 17.9 GFLOP at B = 32, C = 24, T = 16000
      c.launched("fused_level", 2.0 * BT * C * (2.0 * (3 + 1 + 9.0 * C) + 12.0 * C),
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
      // the fused level-0 kernel / the previous level's last conv hand over their output already decimated
      const bool dec_in = (l == 1 && fused_l0) || (l >= 2 && dec_prev);
      const float* const ydec_in[2] = {l == 1 ? ws.ydec[0] : ws.ydec_l[0][l - 1], l == 1 ? ws.ydec[1] : ws.ydec_l[1][l - 1]};
      for (int br = 0; br < 2; ++br) {
        p[br] = tc2_args(c, lw.r1[br], dec_in ? ydec_in[br] : ws.y[br][l - 1], Cp, dec_in ? T_l : T_prev, T_l, 1,
                         ws.tr[br], C);
        p[br].down = dec_in ? 1 : h->dscale[l];
      }
      launch_tc2(c, h, 1, p, 2, "down_r1x1");
      // d1 -> d2 -> d4 and film_conv -> film_out hand their tensors over as bf16 hi|lo operand planes (the consumer's
      // prologue is only a LeakyReLU, applied by the producer): d2, d4 and film_out run without a transform role
      // Only where a CTA sees few items (the layer is a latency chain per item and the transform's turn is a third of
      // it: level 3 convs 25-37 -> 20-26 us, level 2's film_out 35 -> 29 us); with many items per CTA the layer moves
      // at the HBM rate either way (level 1: 147 MB in 23 us) and the producer's longer epilogue costs 4 us.
      const long long lvl_items = 2ll * B * ((T_l + kTc2M - 1) / kTc2M);
      use_pl = C % 8 == 0 && C % lw.c2[0].tc2.CIB == 0 && C % lw.c4[0].tc2.CIB == 0 && (2 * C) % lw.film_out.tc2.CIB == 0 &&
               lvl_items <= 6ll * h->num_sms && !getenv("FSVC_NO_PLANES");
      for (int br = 0; br < 2; ++br) {
        p[br] = tc2_args(c, lw.c1[br], dec_in ? ydec_in[br] : ws.y[br][l - 1], Cp, dec_in ? T_l : T_prev, T_l, 1,
                         use_pl ? nullptr : ws.ta[br], C);
        p[br].down = dec_in ? 1 : h->dscale[l];
        p[br].pre_lrelu = 1;
        if (use_pl) planes_out(p[br], ws.ta[br], C / 8, 0, 1);
      }
      launch_tc2(c, h, 3, p, 2, "down_d1");
      for (int br = 0; br < 2; ++br) {
        p[br] = tc2_args(c, lw.c2[br], ws.ta[br], C, T_l, T_l, 2, use_pl ? nullptr : ws.tb[br], C);
        p[br].pre_lrelu = 1;
        if (use_pl) {
          planes_in(p[br], ws.ta[br], C / 8);
          planes_out(p[br], ws.tb[br], C / 8, 0, 1);
        }
      }
      launch_tc2(c, h, 3, p, 2, "down_d2");
      for (int br = 0; br < 2; ++br) {
        p[br] = tc2_args(c, lw.c4[br], ws.tb[br], C, T_l, T_l, 4, ws.y[br][l], C);
        p[br].pre_lrelu = 1;
        p[br].res = ws.tr[br];
        p[br].res_ld = C;
        if (use_pl) planes_in(p[br], ws.tb[br], C / 8);
        if (dec_out) {
          p[br].out_dec = ws.ydec_l[br][l];
          p[br].out_dec_ld = C;
          p[br].out_dec_T = T_l / h->dscale[l + 1];
          p[br].out_dec_r = h->dscale[l + 1];
        }
      }
      launch_tc2(c, h, 3, p, 2, "down_d4");
      dec_prev = dec_out;
    }
    for (int br = 0; br < 2; ++br) {
      p[br] = tc2_args(c, lw.film[br], ws.y[br][l], C, T_l, T_l, 1, use_pl ? nullptr : ws.H[l] + ntc_col(br * C), 2 * C);
      p[br].post_lrelu = 1;
      if (use_pl) planes_out(p[br], ws.H[l], 2 * C / 8, br * (C / 8), 0);  // both branches side by side: 2C/8 groups
    }
    launch_tc2(c, h, 3, p, 2, "film_conv");
    p[0] = tc2_args(c, lw.film_out, ws.H[l], 2 * C, T_l, T_l, 1, ws.GB[l], 2 * C);
    if (use_pl) planes_in(p[0], ws.H[l], 2 * C / 8);
    launch_tc2(c, h, 3, p, 1, "film_out");
    T_prev = T_l;
  }

  // ---- upsampling stages (fastsvc.py:80-140) ----
  if (!prof) cudaStreamWaitEvent(stream, h->ev_side_join, 0);
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
    // Long utterances (many segments) keep the separate merge kernel: inside the consumer the merge sits on every
    // CTA's critical path and grows with the utterance, the launch does not.  At config 2's 500 segments the two
    // cost the same (A/B on one box: 1.0779 ms with three finalize launches, 1.0758 ms merged, each conv ~10 us longer);
    // the separate kernel is kept there so that a conv launch's time is the conv's.
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
      InFinalizeArgs fa;
      fa.stats = ws.stats[st_w ^ 1];
      fa.n_seg = n_seg;
      fa.T = T_s;
      fa.C = C;
      fa.BC = B * C;
      fa.e = ws.e[i];
      fa.eps = c.eps;
      fa.out_a = ws.pa;
      fa.out_c = ws.pc;
      // plain launch: with programmatic serialization on this small grid the forward was 75 us SLOWER (A/B on one
      // box: 1.338 vs 1.265 ms) -- the convs on either side lose their own overlap
      in_finalize2_kernel<<<(B * C + 7) / 8, 256, 0, stream>>>(fa);
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
    // (one launch: both read h0 and have the same shape; the CTAs are split between the two problems)
    p[0] = tc2_args(c, w.res, ws.h0[i], C, T_in, T_s, 1, ws.xr[i], C);
    p[0].up = r;
    p[1] = tc2_args(c, w.up, ws.h0[i], C, T_in, T_s, 1, ws.t1[i], C);
    p[1].up = r;
    p[1].pre_lrelu = 1;
    p[1].post_lrelu = 1;
    film(p[1]);
    if (w.res.tc2_resident == w.up.tc2_resident && !getenv("FSVC_NO_MERGE")) {
      launch_tc2(c, h, 3, p, 2, "residual+up_film", 0, -1, true);
    } else {
      launch_tc2(c, h, 3, p, 1, "residual");
      launch_tc2(c, h, 3, p + 1, 1, "up_film");
    }
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
    if (i == n - 1 && fold_last) {  // waveform = conv_last(out): written by this conv's epilogue      fastsvc.py:330
      p[0].out = nullptr;
      p[0].last_w = h->last.w;
      p[0].last_b = h->last.b;
      p[0].last_out = out;
      p[0].last_co = h->cfg.out_channels;
    }
    if (i == n - 1 && fold_last && h->out_host && B > 1 && !prof) {
      // fsvc_forward_host: the waveform of the first half of the batch travels back while the second half's last conv
      // runs (items are ordered by utterance, so an item range is an utterance range)
      const int B0 = (B + 1) / 2;
      const size_t half = (size_t)B0 * h->cfg.out_channels * T;
      launch_tc2(c, h, 3, p, 1, "d27_skip+last", 0, B0);
      cudaEventRecord(h->ev_out_half, stream);
      cudaStreamWaitEvent(h->copy_stream, h->ev_out_half, 0);
      cudaMemcpyAsync(h->out_host, out, half * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream);
      cudaEventRecord(h->ev_out_done, h->copy_stream);
      launch_tc2(c, h, 3, p, 1, "d27_skip+last", B0, B);
      h->out_host_done = half;
    } else {
      launch_tc2(c, h, 3, p, 1, i == n - 1 && fold_last ? "d27_skip+last" : "d27_skip");
    }
    x = ws.xs[i];
    x_ld = C;
    T_in = T_s;
  }
  // conv_last (1x1)                                                         fastsvc.py:330
  if (!fold_last) {
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


// ---------------------------------------------------------------------------
// handle-level planning (called from fsvc_create / fsvc_set_weights / fsvc_workspace_bytes)
// ---------------------------------------------------------------------------
size_t tc_plan_handle(fsvc_handle* h) {
  size_t tc_off = 0;
  const int n = h->n;
  // copies tiled for the channels-last persistent kernel; the forward is eligible when every conv
  // except the 1-channel ones (level-0 first conv / 1x1 residual: generated in-kernel) and conv_last has one
  h->tc2_ok = true;
  for (ConvW* cw : h->convs) {
    const bool tiny = (cw->C_in == 1) || cw == &h->last;
    if (tiny) continue;
    if (tc2_plan(cw->C_in, cw->C_out, cw->K, cw)) {
      cw->tc2.w = (const __nv_bfloat16*)tc_off;  // offset, patched by tc_fix_pointers
      tc_off += (cw->tc2.elems() + 127) & ~(size_t)127;
    } else {
      h->tc2_ok = false;
    }
  }
  for (int i = 0; i < n; ++i)
    if (h->cfg.mid_channels[i] % 8 != 0) h->tc2_ok = false;
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
  return tc_off;
}

void tc_fix_pointers(fsvc_handle* h) {
  for (ConvW* cw : h->convs) {
    if (cw->tc2.n_ntiles) cw->tc2.w = h->tc_store + (size_t)cw->tc2.w;
    if (cw->nc_G) cw->wnc = h->tc_store + (size_t)cw->wnc;
  }
}

// Tensor-core weight packs as one launch over a device job table (kinds 4 / 5 of WJob): fp32 packed [C_in][K][C_out] ->
// bf16 hi|lo in the tiled UMMA layouts of conv_tc3 (pack_tc_weights_kernel's format) and of the fused level kernel.
static __global__ void __launch_bounds__(256) weight_jobs_tc_kernel(const WJob* __restrict__ jobs) {
  const WJob j = jobs[blockIdx.y];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(j.dst);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    if (j.kind == 4) {
      const long long half = (long long)j.K * j.CIB * j.N_tile;
      const int e = r % 8; r /= 8;
      const int n = r % j.N_tile; r /= j.N_tile;
      const int g = r % (j.CIB / 8); r /= (j.CIB / 8);
      const int k = r % j.K; r /= j.K;
      const int blk = r % j.n_blk; r /= j.n_blk;
      const int nt = (int)r;
      const int ci = blk * j.CIB + g * 8 + e, co = nt * j.N_tile + n;
      float v = 0.f;
      if (ci < j.C_in && co < j.C_out) v = j.w[((long long)ci * j.K + k) * j.C_out + co];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      const long long base = ((long long)(nt * j.n_blk + blk) * 2) * half + (((long long)k * (j.CIB / 8) + g) * j.N_tile + n) * 8 + e;
      dst[base] = hi;
      dst[base + half] = lo;
    } else {
      const int e = r % 8; r /= 8;
      const int n = r % j.N; r /= j.N;
      const int g = r % j.G; r /= j.G;
      const int k = (int)r;
      const int ci = g * 8 + e;
      float v = 0.f;
      if (ci < j.C_in && n < j.C_out) v = j.w[((long long)ci * j.K + k) * j.C_out + n];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      const long long base = (((long long)k * j.G + g) * 2 * j.N + n) * 8 + e;
      dst[base] = hi;
      dst[base + (long long)j.N * 8] = lo;
    }
  }
}

int tc_build_jobs(fsvc_handle* h) {
  std::vector<WJob> jobs;
  for (ConvW* cw : h->convs) {
    if (cw->tc2.w) {
      const TcW& t = cw->tc2;
      WJob j;
      memset(&j, 0, sizeof(j));
      j.kind = 4; j.C_out = cw->C_out; j.C_in = cw->C_in; j.K = cw->K; j.w = cw->w; j.dst = (void*)t.w;
      j.CIB = t.CIB; j.n_blk = t.n_blk; j.N_tile = t.N_tile; j.n_ntiles = t.n_ntiles;
      j.total = (long long)(t.elems() / 2);
      jobs.push_back(j);
    }
    if (cw->nc_G) {
      WJob j;
      memset(&j, 0, sizeof(j));
      j.kind = 5; j.C_out = cw->C_out; j.C_in = cw->C_in; j.K = cw->K; j.w = cw->w; j.dst = (void*)cw->wnc;
      j.G = cw->nc_G; j.N = cw->nc_N;
      j.total = (long long)cw->K * cw->nc_G * cw->nc_N * 8;
      jobs.push_back(j);
    }
  }
  h->n_jobs_tc = (int)jobs.size();
  if (jobs.empty()) return FSVC_OK;
  FSVC_CUDA(cudaMalloc((void**)&h->jobs_tc, jobs.size() * sizeof(WJob)));
  FSVC_CUDA(cudaMemcpy(h->jobs_tc, jobs.data(), jobs.size() * sizeof(WJob), cudaMemcpyHostToDevice));
  return FSVC_OK;
}

void tc_pack_weights(fsvc_handle* h, cudaStream_t s) {
  if (h->n_jobs_tc) weight_jobs_tc_kernel<<<dim3(16, h->n_jobs_tc), 256, 0, s>>>(h->jobs_tc);
}

size_t tc_workspace_bytes(const fsvc_handle* h, int B, int frames) {
  WS2 ws2;
  return layout_ws2(h, B, frames, nullptr, 0, &ws2);
}

}  // namespace fsvc

#ifdef FSVC_TIMELINE
// debug builds only (tools/timeline.py): copy the event stamps of the last forward
extern "C" int fsvc_debug_timeline(unsigned long long* out, int n) {
  FSVC_CUDA(cudaMemcpyFromSymbol(out, fsvc::g_tl, sizeof(unsigned long long) * (size_t)(n < 64 * 8 * 64 ? n : 64 * 8 * 64)));
  return FSVC_OK;
}
#endif
