// Warp-specialised, software-pipelined tcgen05 1-D convolution over channels-last activations.
//
// Tc2Args (ntc_common.cuh) describes the fused prologue / epilogue of every FastSVC conv; the kernel is organised
// as a four-role pipeline so that no role ever waits on HBM with nothing else to do:
//
//   warp 0      WEIGHTS    cp.async.bulk (UBLKCP) of the packed bf16 hi|lo weights: once when they fit in
//               + ROWS     shared memory, else one (N tile, ci block) chunk per MMA block through a 2-slot ring.  The same
//                          thread is the ROW LOADER of the direct layers (MODE 3 / 4: the <= 6 aligned 32-row blocks of
//                          raw rows covering an item's window, bulk-copied into a staging ring; MODE 6: the producer's
//                          bf16 hi|lo operand planes straight into the A ring)
//   warp 1      MMA        one thread issues tcgen05.mma (bf16 3-term split, fp32 accumulation in TMEM,
//                          two accumulators so tile i+1 is computed while tile i is drained)
//   warps 2-7   TRANSFORM  staged raw rows -> InstanceNorm affine -> LeakyReLU -> zero padding -> bf16 hi|lo in the
//                          UMMA K-major canonical layout (taps = descriptor row shifts); packed fp32 arithmetic.
//                          MODE 0 / 2 (decimated / wide nearest-repeat inputs) fetch their rows themselves with
//                          per-lane cp.async one chunk ahead; MODE 1 generates the operand from a 1-channel signal;
//                          MODE 6 has no transform work at all
//   warps 8-15  EPILOGUE   TMEM -> registers (+bias, residual, FiLM affine) -> 128-bit row stores; the
//                          residual / gamma / beta rows of the NEXT sub-tile are requested before the
//                          current one is waited for; InstanceNorm partial statistics through a per-warp
//                          shared-memory transpose.  Two instantiations: plain (bias / residual / LeakyReLU / stores,
//                          optionally operand planes and a decimated copy for the consumer) and general
//
// The roles meet only through mbarriers (full/empty pairs per ring slot, tcgen05.commit on the tensor-core
// side).  One CTA per SM, one contiguous range of (utterance, 128-step tile) items per CTA.
// Activations are blocked channels-last fp32 ([B][T/32][C/4][32][4], ntc_common.cuh): every global access is a
// 128-bit access or a bulk copy of a contiguous run.
#pragma once
#include "ntc_common.cuh"

namespace fsvc {

// Role split of the 512-thread CTA (one CTA per SM): weights + row-loader warp, MMA warp, 6 transform warps, 8 epilogue
// warps.  (Measured and dropped: two 224-thread CTAs per SM -- no faster on any layer; 10 transform + 4 epilogue warps
// for the plain-epilogue layers -- slower on every conv with a residual.)
struct Tc3Shape {
  static constexpr int kMmaWarp = 1;
  static constexpr int kX0 = 2;                         // first transform warp
  static constexpr int kXW = 6;                         // transform warps
  static constexpr int kE0 = kX0 + kXW;                // first epilogue warp
  static constexpr int kEW = 8;                         // epilogue warps (4 lane quarters x kEH column halves)
  static constexpr int kEH = kEW / 4;
  static constexpr int kXT = 32 * kXW, kET = 32 * kEW;
  static constexpr int kThreads = 32 * (kE0 + kEW);
};
constexpr int kTc3Threads = Tc3Shape::kThreads;
constexpr int kTc3ChunkItems = 4;      // 16-byte-chunk items per transform thread and prefetch chunk

struct Tc3Cfg {
  int n_prob;  // 1 or 2 problems (blockIdx.x % n_prob); problem 1 = problem 0 + the pointer deltas below
  int B, m_tiles;
  int b_lo, b_hi;  // utterances [b_lo, b_hi) of the batch are processed by this launch (normally 0, B)
  int cta0;        // hetero launches: CTAs per N tile of problem 0 (problem 1 gets the rest); 0 = equal interleaved split
  int nsub;       // epilogue sub-tile width (channels, multiple of 8, <= 32)
  int a_slots;    // depth of the A ring (1..3)
  int scr_pitch;  // floats per row of the per-warp statistics scratch (12 or 20)
  uint32_t a_bytes, b_bytes;  // per ring slot
  uint32_t off_w, off_a, off_b, off_scr, off_pa, off_bias, off_fin, off_tab, off_stg, off_bar, total;
  uint32_t stg_bytes;  // one slot of the transform's raw-row staging ring (a chunk: kXW warps x CH tasks x 1 KB)
  int stg_depth;       // chunks in flight (ring of stg_depth + 1 slots)
  int cpw;             // tasks per transform warp and chunk (<= kTc3ChunkItems): smaller chunks = smaller staging slots
};

struct Tc3Launch {
  Tc2Args a;
  Tc2Args b;   // hetero launches: problem 1's own argument block (same shapes and tiling, different prologue /
  int hetero;  // epilogue flags and tensors), instead of pointer deltas against problem 0
  long long d_in, d_w, d_bias, d_gen_w, d_gen_b, d_res, d_gres_w, d_gres_b, d_gres_x, d_raw, d_out;  // elements
  long long d_in_pl, d_out_pl;  // 16-byte chunks
  long long d_out_dec;          // elements
  Tc3Cfg c;
  int tl_slot;  // launch index inside the forward (event timeline builds only)
};

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int4 lds128i(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// two predicated 128-bit read-only loads 512 bytes apart (the two 4-channel halves of an 8-channel group)
__device__ __forceinline__ void ldg2_pred(const float4* p, uint32_t ok, float4& x, float4& y) {
  asm volatile(
      "{\n"
      ".reg .pred pp;\n"
      "setp.ne.u32 pp, %9, 0;\n"
      "@pp ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%8];\n"
      "@pp ld.global.nc.v4.f32 {%4,%5,%6,%7}, [%8+512];\n"
      "}"
      : "+f"(x.x), "+f"(x.y), "+f"(x.z), "+f"(x.w), "+f"(y.x), "+f"(y.y), "+f"(y.z), "+f"(y.w)
      : "l"(p), "r"(ok));
}
// bf16 hi|lo split of 8 fp32 values, stored as two 16-byte chunks (32-bit shared-space addresses)
__device__ __forceinline__ void split_store_s(uint32_t addr, uint32_t plane, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_pair(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
  sts128(addr, make_uint4(hi[0], hi[1], hi[2], hi[3]));
  sts128(addr + plane, make_uint4(lo[0], lo[1], lo[2], lo[3]));
}

// 16-byte asynchronous global -> shared copy (LDGSTS); bytes = 0 writes zeros without reading
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// wait until at most `n` of this thread's cp.async groups are still pending (n is a run-time value, 0..8)
__device__ __forceinline__ void cp_async_wait_n(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    case 7: cp_async_wait<7>(); break;
    default: cp_async_wait<8>(); break;
  }
}
// bf16 hi|lo split of 8 fp32 values into two packed 16-byte chunks
__device__ __forceinline__ void split_bf16(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h4[4], l4[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_pair(v[2 * e], v[2 * e + 1], h4[e], l4[e]);
  hi = make_uint4(h4[0], h4[1], h4[2], h4[3]);
  lo = make_uint4(l4[0], l4[1], l4[2], l4[3]);
}
// same split, stored under predicates: value where (pv & 1), zeros where (pz & 1), nothing otherwise
__device__ __forceinline__ void split_store_p(uint32_t addr, uint32_t plane, const float (&v)[8], uint32_t pv, uint32_t pz) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_pair(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
  asm volatile(
      "{\n"
      ".reg .pred pv, pz;\n"
      ".reg .b32 zz;\n"
      "and.b32 zz, %10, 1;\n"
      "setp.ne.u32 pv, zz, 0;\n"
      "and.b32 zz, %11, 1;\n"
      "setp.ne.u32 pz, zz, 0;\n"
      "mov.b32 zz, 0;\n"
      "@pv st.shared.v4.b32 [%0], {%2,%3,%4,%5};\n"
      "@pv st.shared.v4.b32 [%1], {%6,%7,%8,%9};\n"
      "@pz st.shared.v4.b32 [%0], {zz,zz,zz,zz};\n"
      "@pz st.shared.v4.b32 [%1], {zz,zz,zz,zz};\n"
      "}" ::"r"(addr), "r"(addr + plane), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(lo[0]), "r"(lo[1]),
      "r"(lo[2]), "r"(lo[3]), "r"(pv), "r"(pz)
      : "memory");
}

// ---- optional event timeline (build with -DFSVC_TIMELINE; tools/timeline.py): globaltimer stamps of the first
// 8 CTAs of every conv_tc3 launch of a forward, to see where a short kernel's latency goes.  Compiled out otherwise.
#ifdef FSVC_TIMELINE
__device__ unsigned long long g_tl[64 * 8 * 64];
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FSVC_TL(slot, ev)                                                                          \
  do {                                                                                             \
    if (blockIdx.x < 8 && (slot) < 64) g_tl[(((slot) * 8) + blockIdx.x) * 64 + (ev)] = tl_now();   \
  } while (0)
#else
#define FSVC_TL(slot, ev) \
  do {                    \
  } while (0)
#endif

template <bool V>
struct BoolC {
  static constexpr bool value = V;
};

enum {  // mbarrier indices
  kBarBFull = 0,     // [2] streamed weight block landed
  kBarBEmpty = 2,    // [2] MMAs that read it completed
  kBarAFull = 4,     // [3] A block converted
  kBarAEmpty = 7,    // [3] MMAs that read it completed
  kBarAccFull = 10,  // [2] accumulator holds a finished tile
  kBarAccEmpty = 12, // [2] epilogue drained it
  kBarWFull = 14,    // resident weights landed
  kBarStgFull = 15,  // [4] bulk-copied raw rows of a staging slot landed (lean transform, MODE 3 / 4)
  kBarStgEmpty = 19, // [4] every transform warp has converted the slot's rows
  kBarCount = 23
};

// Shared-memory plan of one launch.  Returns false when even a 1-deep A ring does not fit.
// cpw_force = 4 / 6: lean transform (MODE 3 / 4 / 5).  bulk = true (MODE 3 / 4): a staging slot holds the 6 aligned
// 32-row blocks of the input that cover an item's window, filled by cp.async.bulk.
__host__ inline bool tc3_plan_smem(const Tc2Args& a, int K, Tc3Cfg* c, int cpw_force = 0, bool bulk = false,
                                   bool direct = false) {
  const int halo = (K / 2) * a.dil, W = kTc2M + 2 * halo;
  const uint32_t Gb = a.CIB / 8;
  c->a_bytes = 2u * Gb * W * 16u;
  c->b_bytes = 2u * K * Gb * a.N_tile * 16u;
  const int nvalid_max = a.C_out < a.N_tile ? a.C_out : a.N_tile;
  // epilogue sub-tile: a thread owns nsub / (column halves) channels, at most 16
  const int eh = Tc3Shape::kEH;
  int nsub = 8 * eh;
  for (int cand : {12 * eh, 16 * eh, 8 * eh, 4 * eh})
    if (nvalid_max % cand == 0) {
      nsub = cand;
      break;
    }
  c->nsub = nsub;
  c->scr_pitch = nsub / eh <= 12 ? 12 : 20;
  const uint32_t w_bytes = a.w_resident ? c->b_bytes * a.n_blk : 0u;
  const uint32_t scr_bytes = (uint32_t)Tc3Shape::kEW * (32u * c->scr_pitch + 16u) * 4u;
  const uint32_t budget = 227u * 1024u;
  const uint32_t pa_bytes = 3u * 2u * ((a.C_in + 7) / 8 * 8) * 4u;  // triple-buffered per-utterance affine
  // Preference: deep staging (loads in flight) first, then A ring depth.
  const uint32_t xw = (uint32_t)Tc3Shape::kXW;
  // {A ring slots, staging depth, tasks per warp and chunk}: deep staging first, then A ring depth; half-size chunks
  // (half the staging memory) before giving up the double-buffered A ring
  static const int pref[][3] = {{3, 3, 4}, {2, 3, 4}, {3, 2, 4}, {2, 2, 4}, {3, 1, 4}, {2, 1, 4},
                                {3, 1, 2}, {2, 1, 2}, {1, 1, 4}, {1, 1, 2}};
  for (const auto& pr : pref) {
    const int slots = pr[0];
    c->a_slots = slots;
    c->stg_depth = a.gen_w ? 0 : pr[1];
    c->cpw = cpw_force ? cpw_force : pr[2];
    if (cpw_force && pr[2] != 4) continue;
    c->stg_bytes = a.gen_w ? 0u : xw * (uint32_t)c->cpw * 1024u;
    if (bulk) c->stg_bytes = 6u * (uint32_t)(a.CIB >> 2) * 512u;
    if (direct) {  // MODE 6: the A ring is filled from operand planes, nothing is staged
      c->stg_bytes = 0u;
      c->stg_depth = 0;
    }
    uint32_t off = 0;
    c->off_w = off;
    off += w_bytes;
    c->off_a = off;
    off += slots * c->a_bytes;
    c->off_b = off;
    off += a.w_resident ? 0u : 2u * c->b_bytes;
    c->off_scr = off;
    off += scr_bytes;
    c->off_pa = off;
    off += pa_bytes;
    off = (off + 15u) & ~15u;
    c->off_bias = off;
    off += (uint32_t)((a.N_tile + 3) / 4) * 16u;  // this N tile's bias, staged once by the epilogue warps
    c->off_fin = off;
    off += a.pre_stats ? (uint32_t)Tc3Shape::kXT * 32u : 0u;  // statistics merge scratch: 4 doubles per thread
    c->off_tab = off;
    off += Gb * (uint32_t)((W + 31) / 32) * 16u;  // transform geometry table
    c->off_stg = off;
    off += c->stg_bytes * (uint32_t)(c->stg_depth + 1);
    c->off_bar = off;
    off += kBarCount * 8 + 16;
    c->total = off;
    if (off <= budget) return true;
  }
  return false;
}

// NH4: float4 units per epilogue thread and sub-tile, fixed at compile time (3 = the 24-channel sub-tile every
// conv of the YAML generator uses) or 0 = decided at run time (any multiple-of-8 channel count).
// GEN: the A operand is generated from a 1-channel signal (first conv of an unfused level-0 chain).
// MODE: 0 = table-driven transform (source row = window row * down), 1 = GEN, 2 = nearest-repeat input (up > 1,
// down == 1): every source row is converted once and stored to its `up` window rows.
template <int K, int NH4, int MODE>
__global__ void __launch_bounds__(Tc3Shape::kThreads, 1)
conv_tc3_kernel(const __grid_constant__ Tc3Launch L) {
  using SH = Tc3Shape;
  constexpr int kTc3XformThreads = SH::kXT, kTc3EpiThreads = SH::kET;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* const smem = smem_raw;
  const Tc3Cfg& c = L.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // problem / N tile / CTA index of this block.  Two problems normally interleave and share the grid equally; a hetero
  // launch may give them different shares (cta0 > 0: the first cta0 * n_ntiles blocks belong to problem 0) when one
  // problem's items cost more than the other's
  const int g_nt = L.a.n_ntiles;
  const int split0 = c.cta0 * g_nt;
  const int prob = c.cta0 > 0 ? ((int)blockIdx.x >= split0 ? 1 : 0) : (int)(blockIdx.x % c.n_prob);
  const Tc2Args& a = (L.hetero && prob == 1) ? L.b : L.a;
  const int rest = c.cta0 > 0 ? (prob ? (int)blockIdx.x - split0 : (int)blockIdx.x) : (int)(blockIdx.x / c.n_prob);
  const int nt = rest % g_nt;
  // items (utterance, 128-step tile) of this CTA: one contiguous range -- neighbouring tiles share their halo rows
  // in L2 and the utterance (hence the InstanceNorm affine in shared memory) changes at most a few times per CTA
  const int n_cta = c.cta0 > 0 ? (prob ? ((int)gridDim.x - split0) / g_nt : c.cta0) : (int)(gridDim.x / (c.n_prob * g_nt));
  const int cta = rest / g_nt;
  const int n_all = (c.b_hi - c.b_lo) * c.m_tiles, item_lo = c.b_lo * c.m_tiles;
  const int first = item_lo + (int)((long long)cta * n_all / n_cta);
  const int n_m = item_lo + (int)((long long)(cta + 1) * n_all / n_cta);
  constexpr int step = 1;

  const int halo = (K / 2) * a.dil;
  const int W = kTc2M + 2 * halo;
  constexpr bool GEN = MODE == 1;
  constexpr bool gen = GEN;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + c.off_bar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kBarCount);
  uint32_t acc_stride = 32;
  while (acc_stride < (uint32_t)a.N_tile) acc_stride <<= 1;

  if (tid == 0) FSVC_TL(L.tl_slot, 0);
  griddep_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bars + kBarBFull + i, 1);
      mbar_init(bars + kBarBEmpty + i, 1);
      mbar_init(bars + kBarAccFull + i, 1);
      mbar_init(bars + kBarAccEmpty + i, kTc3EpiThreads);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(bars + kBarAFull + i, MODE == 6 ? 1 : kTc3XformThreads);  // MODE 6: the loader's expect_tx arrival
      mbar_init(bars + kBarAEmpty + i, 1);
    }
    mbar_init(bars + kBarWFull, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bars + kBarStgFull + i, 1);
      mbar_init(bars + kBarStgEmpty + i, SH::kXW);
    }
    fence_barrier_init();
  }
  if (warp == SH::kMmaWarp) tmem_alloc(s_tmem, 2 * acc_stride);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (tid == 0) FSVC_TL(L.tl_slot, 1);

  const int nvalid = min(a.N_tile, a.C_out - nt * a.N_tile);  // valid output channels of this N tile
  const int n_sub = (nvalid + c.nsub - 1) / c.nsub;           // epilogue sub-tiles of this N tile
  const int co_tile = nt * a.N_tile;

  if (warp == 0 && lane == 0) {
    // =============================== WEIGHTS (+ ROW LOADER of the lean transform) ===============================
    const uint8_t* w_nt = reinterpret_cast<const uint8_t*>(a.w + prob * L.d_w) + (size_t)nt * a.n_blk * c.b_bytes;
    if (a.w_resident) {
      const uint32_t total = c.b_bytes * a.n_blk;
      mbar_expect_tx(bars + kBarWFull, total);
      for (uint32_t o = 0; o < total; o += 32768u)
        bulk_g2s(smem + c.off_w + o, w_nt + o, min(32768u, total - o), bars + kBarWFull);
    }
    uint32_t pbk = 0;
    auto stream_weights = [&](int blk) {  // one (N tile, ci block) chunk into the 2-slot ring
      const uint32_t slot = pbk & 1u, use = pbk >> 1;
      if (use > 0) mbar_wait2(bars + kBarBEmpty + slot, (use + 1) & 1u);
      mbar_expect_tx(bars + kBarBFull + slot, c.b_bytes);
      const uint8_t* src = w_nt + (size_t)blk * c.b_bytes;
      uint8_t* dst = smem + c.off_b + slot * c.b_bytes;
      for (uint32_t o = 0; o < c.b_bytes; o += 32768u)
        bulk_g2s(dst + o, src + o, min(32768u, c.b_bytes - o), bars + kBarBFull + slot);
      ++pbk;
    };
    if constexpr (MODE == 3 || MODE == 4) {
      // The blocked channels-last layout keeps the channels of 32 consecutive time steps contiguous, so the <= 6
      // aligned 32-row blocks covering an item's window [t0 - halo, t0 + 128 + halo) are, per ci block, <= 6 bulk
      // copies (TMA engine) into one staging slot, completed on the slot's mbarrier.  They are issued from this
      // thread: measured from inside the transform role, issuing an item's copies cost 0.7 us on its critical path.
      const float4* in4 = reinterpret_cast<const float4*>(a.in + prob * L.d_in);
      const long long Tp_in = ntc_tp(a.T_in), ld4 = a.in_ld >> 2;
      const uint32_t row_bytes = (uint32_t)(a.CIB >> 2) * 512u;  // one staged 32-row block of a ci block
      const int n_blk32 = (int)(Tp_in >> 5), ring = c.stg_depth + 1;
      int lb = first / c.m_tiles, ltile = first - lb * c.m_tiles;
      uint32_t slot = 0, use = 0;
      bool waited = false;
      for (int m = first; m < n_m; ++m) {
        const int r0 = ltile * (kTc2M / 32) - 1;  // first staged block (-1 at the start of an utterance: not copied)
        int nv = 0;
        for (int r = 0; r < 6; ++r) nv += (r0 + r >= 0 && r0 + r < n_blk32) ? 1 : 0;
        for (int blk = 0; blk < a.n_blk; ++blk) {
          if (!a.w_resident) stream_weights(blk);
          if (!waited) {
            griddep_wait();  // first access to the predecessor's output (the weights above do not depend on it)
            waited = true;
          }
          if (use > 0) mbar_wait2(bars + kBarStgEmpty + slot, (use + 1) & 1u);
          const uint32_t nbytes = (uint32_t)(min(a.CIB, a.C_in - blk * a.CIB) >> 2) * 512u;  // real channels of the block
          uint64_t* bar = bars + kBarStgFull + slot;
          mbar_expect_tx(bar, (uint32_t)nv * nbytes);
          uint8_t* dst = smem + c.off_stg + slot * c.stg_bytes;
          const float4* src = in4 + ((long long)lb * Tp_in + (long long)r0 * 32) * ld4 + (long long)blk * (a.CIB >> 2) * 32;
          for (int r = 0; r < 6; ++r)
            if (r0 + r >= 0 && r0 + r < n_blk32) bulk_g2s(dst + (uint32_t)r * row_bytes, src + (long long)r * 32 * ld4, nbytes, bar);
          if (++slot == (uint32_t)ring) {
            slot = 0;
            ++use;
          }
        }
        if (++ltile == c.m_tiles) {
          ltile = 0;
          ++lb;
        }
      }
    } else if constexpr (MODE == 6) {
      // Operand planes (ntc_common.cuh): the window of an item, per 8-channel group and plane, is one contiguous run of
      // W 16-byte rows in global memory AND in the A slot -- 2 * Gb bulk copies per (item, ci block) fill the slot,
      // completed on the slot's AFull barrier.  Rows outside the utterance are zeros in the tensor.
      const uint4* pl = a.in_pl + prob * L.d_in_pl;
      const uint32_t Gb = (uint32_t)a.CIB >> 3, strip = (uint32_t)W * 16u, plane = Gb * strip;
      int lb = first / c.m_tiles, ltile = first - lb * c.m_tiles;
      uint32_t aslot = 0, ause = 0;
      bool waited = false;
      for (int m = first; m < n_m; ++m) {
        const long long row0 = (long long)ltile * kTc2M - halo + kPlPad;
        for (int blk = 0; blk < a.n_blk; ++blk) {
          if (!a.w_resident) stream_weights(blk);
          if (!waited) {
            griddep_wait();  // first access to the predecessor's output
            waited = true;
          }
          if (ause > 0) mbar_wait2(bars + kBarAEmpty + aslot, (ause + 1) & 1u);
          const int g0 = blk * (int)Gb, ng = min((int)Gb, (a.C_in >> 3) - g0);
          uint64_t* bar = bars + kBarAFull + aslot;
          mbar_expect_tx(bar, 2u * (uint32_t)ng * strip);
          uint8_t* dst = smem + c.off_a + aslot * c.a_bytes;
          const uint4* src = pl + ((long long)lb * a.in_pl_G + g0) * a.in_pl_Tp + row0;
          for (int g = 0; g < ng; ++g) {
            bulk_g2s(dst + (uint32_t)g * strip, src + (long long)g * a.in_pl_Tp, strip, bar);
            bulk_g2s(dst + plane + (uint32_t)g * strip, src + (long long)g * a.in_pl_Tp + a.in_pl_lo, strip, bar);
          }
          if (tid == 0 && (a.n_blk == 1 ? m - first : blk) < 12 && (a.n_blk == 1 || m == first))
            FSVC_TL(L.tl_slot, 4 + (a.n_blk == 1 ? m - first : blk));
          if (++aslot == (uint32_t)c.a_slots) {
            aslot = 0;
            ++ause;
          }
        }
        if (++ltile == c.m_tiles) {
          ltile = 0;
          ++lb;
        }
      }
    } else if (!a.w_resident) {
      for (int m = first; m < n_m; m += step)
        for (int blk = 0; blk < a.n_blk; ++blk) stream_weights(blk);
    }
  }
  if (warp == SH::kMmaWarp) {
    // =============================== MMA ISSUER ===============================
    // The whole warp runs the (warp-uniform) control flow and descriptor arithmetic so that it stays in the uniform
    // datapath; one elected lane issues the tcgen05 instructions.  (With the loop inside `if (lane == 0)` every
    // descriptor went through per-MMA vector->uniform moves: ~460 instructions per tile, 3000 cycles, the bound of
    // every 24-channel layer.)
    {
      const bool leader = elect_one_sync();
      const uint32_t idesc = umma_idesc_bf16(kTc2M, a.N_tile);
      const uint32_t Gb = (uint32_t)a.CIB >> 3;
      const uint32_t strip = (uint32_t)W * 16u, a_plane = Gb * strip;
      const uint32_t b_strip = (uint32_t)a.N_tile * 16u, b_half = (uint32_t)K * Gb * b_strip;
      const uint32_t a_hiw = (uint32_t)(umma_desc(0, strip, 128) >> 32), b_hiw = (uint32_t)(umma_desc(0, b_strip, 128) >> 32);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t smem_u = smem_u32(smem);
      const uint32_t n_kc = (uint32_t)a.CIB >> 4;
      if (a.w_resident) mbar_wait2(bars + kBarWFull, 0);
      uint32_t aslot = 0, ause = 0, pbk = 0;
      int it = 0;
      for (int m = first; m < n_m; m += step, ++it) {
        const uint32_t acc = (uint32_t)it & 1u;
        if (it >= 2) mbar_wait2(bars + kBarAccEmpty + acc, (((uint32_t)it >> 1) + 1) & 1u);
        const uint32_t d_tmem = tmem_u + acc * acc_stride;
        for (int blk = 0; blk < a.n_blk; ++blk) {
          if (leader && it == 5 && blk == 0) FSVC_TL(L.tl_slot, 51);
          mbar_wait2(bars + kBarAFull + aslot, ause & 1u);
          if (leader && it == 0 && blk < 12) FSVC_TL(L.tl_slot, 20 + blk);
          if (leader && it == 5 && blk == 0) FSVC_TL(L.tl_slot, 52);
          uint32_t sB_addr;
          const uint32_t bslot = pbk & 1u;
          if (a.w_resident) {
            sB_addr = smem_u + c.off_w + (uint32_t)blk * c.b_bytes;
          } else {
            mbar_wait2(bars + kBarBFull + bslot, (pbk >> 1) & 1u);
            sB_addr = smem_u + c.off_b + bslot * c.b_bytes;
          }
          tc_fence_after();
          const uint32_t sA_addr = smem_u + c.off_a + aslot * c.a_bytes;
          // descriptors differ only in the 14-bit start-address field of the low word (addresses < 256 KB)
          const uint32_t a_w0 = (uint32_t)umma_desc(sA_addr, strip, 128), b_w0 = (uint32_t)umma_desc(sB_addr, b_strip, 128);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            uint32_t a_w = a_w0 + (uint32_t)(k * a.dil), b_w = b_w0 + (((uint32_t)k * Gb * b_strip) >> 4);
            for (uint32_t kc = 0; kc < n_kc; ++kc) {
              const uint64_t a_hi = ((uint64_t)a_hiw << 32) | a_w, a_lo = ((uint64_t)a_hiw << 32) | (a_w + (a_plane >> 4));
              const uint64_t b_hi = ((uint64_t)b_hiw << 32) | b_w, b_lo = ((uint64_t)b_hiw << 32) | (b_w + (b_half >> 4));
              const uint32_t accum = (blk == 0 && k == 0 && kc == 0) ? 0u : 1u;
              if (leader) {
                umma_bf16(d_tmem, a_lo, b_hi, idesc, accum);  // small terms first, then the dominant one
                umma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
                umma_bf16(d_tmem, a_hi, b_hi, idesc, 1u);
              }
              a_w += (2u * strip) >> 4;
              b_w += (2u * b_strip) >> 4;
            }
          }
          if (leader) {
            umma_commit(bars + kBarAEmpty + aslot);
            if (!a.w_resident) umma_commit(bars + kBarBEmpty + bslot);
            if (blk == a.n_blk - 1) umma_commit(bars + kBarAccFull + acc);
            if (it == 0 && blk == a.n_blk - 1) FSVC_TL(L.tl_slot, 36);
            if (it == 5 && blk == a.n_blk - 1) FSVC_TL(L.tl_slot, 53);
          }
          if (!a.w_resident) ++pbk;
          if (++aslot == (uint32_t)c.a_slots) {
            aslot = 0;
            ++ause;
          }
        }
      }
      __syncwarp();
    }
  } else if (warp >= SH::kX0 && warp < SH::kE0) {
   if constexpr (MODE == 6) {
    // no transform: the A ring is filled by the row loader from the producer's operand planes
   } else if constexpr (GEN) {
    // =============================== TRANSFORM ===============================
    const int tt = tid - 32 * SH::kX0;
    const uint32_t smem_base = smem_u32(smem);
    const float* in = a.in + prob * L.d_in;  // gen mode: 1-channel signal [B][T_in]
    const float* gen_w = gen ? a.gen_w + prob * L.d_gen_w : nullptr;
    const float* gen_b = gen ? a.gen_b + prob * L.d_gen_b : nullptr;
    float* s_pa_base = reinterpret_cast<float*>(smem + c.off_pa);
    const int cpad = (a.C_in + 7) / 8 * 8;
    const int Gb = a.CIB >> 3;  // groups of 8 channels per (padded) ci block
    const uint32_t strip = (uint32_t)W * 16u, plane = (uint32_t)Gb * strip;
    const uint32_t w_magic = 0xFFFFFFFFu / (uint32_t)W + 1u;      // exact quotients below 2^16
    const long long Tp_in = ntc_tp(a.T_in);
    const uint32_t up_magic = 0xFFFFFFFFu / (uint32_t)a.up + 1u;
    const int items = W * Gb;
    constexpr int CH = kTc3ChunkItems, CHUNK = CH * kTc3XformThreads;
    const int n_chunks = (items + CHUNK - 1) / CHUNK;

    struct Cursor {
      int m, blk, ch, it, b, tile;  // (b, tile) track m without a division per chunk
    };
    const int step_b = step / c.m_tiles, step_t = step - step_b * c.m_tiles;
    auto advance = [&](Cursor& q) {
      if (++q.ch == n_chunks) {
        q.ch = 0;
        if (++q.blk == a.n_blk) {
          q.blk = 0;
          q.m += step;
          q.b += step_b;
          q.tile += step_t;
          if (q.tile >= c.m_tiles) {
            q.tile -= c.m_tiles;
            ++q.b;
          }
          ++q.it;
        }
      }
    };
    // Affine buffers are indexed by the number of utterance changes so far (three buffers: a rewrite follows the
    // named barrier of the previous change, which every thread passes only after converting the tiles before it).
    int pa_b = -1;
    uint32_t pa_epoch = 0, cv_epoch = 0;  // epoch of the tile being loaded / converted
    int cv_b = -1;
    // issue the loads of one chunk (no use of the results)
    auto load_chunk = [&](const Cursor& q, float4 (&d)[CH][2], uint32_t& live) {
      live = 0;
      if (q.m >= n_m) return;
      const int b = q.b, t0 = q.tile * kTc2M;
      const int ci0 = q.blk * a.CIB;
      // InstanceNorm affine of this tile's utterance, written one chunk early.  Three buffers: the write for
      // tile i+3 follows the named barrier of tile i+1, which every thread reaches only after converting tile i.
      if (q.blk == 0 && q.ch == 0 && a.pre_a && b != pa_b) {
        pa_b = b;
        ++pa_epoch;
        float* s_pa = s_pa_base + (pa_epoch % 3) * 2 * cpad;
        for (int ch = tt; ch < cpad; ch += kTc3XformThreads) {
          s_pa[ch] = ch < a.C_in ? __ldg(a.pre_a + (long long)b * a.C_in + ch) : 1.f;
          s_pa[cpad + ch] = ch < a.C_in ? __ldg(a.pre_c + (long long)b * a.C_in + ch) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int idx = q.ch * CHUNK + j * kTc3XformThreads + tt;
        if (idx < items) {
          const int g = (int)__umulhi((uint32_t)idx, w_magic), r = idx - g * W;  // lanes <-> consecutive steps
          const int u = t0 - halo + r, ch = ci0 + g * 8;
          if (u >= 0 && u < a.T_out && ch < a.C_in) {
            live |= 1u << j;
            if (gen) {
              const float* x = in + (long long)b * a.T_in + u;
              d[j][0].x = u > 0 ? __ldg(x - 1) : 0.f;
              d[j][0].y = __ldg(x);
              d[j][0].z = u + 1 < a.T_out ? __ldg(x + 1) : 0.f;
            } else {
              const int src = (a.up == 1 ? u : (int)__umulhi((uint32_t)u, up_magic)) * a.down;
              const float4* p = reinterpret_cast<const float4*>(in + ntc_row(Tp_in, a.in_ld, b, src) + (ch >> 2) * 128);
              d[j][0] = __ldg(p);
              d[j][1] = __ldg(p + 32);
            }
          }
        }
      }
    };
    griddep_wait();  // first access to the predecessor's output
    uint32_t pa_pos = 0;
    auto convert_chunk = [&](const Cursor& q, const float4 (&d)[CH][2], uint32_t live) {
      const uint32_t aslot = pa_pos % (uint32_t)c.a_slots, ause = pa_pos / (uint32_t)c.a_slots;
      if (q.ch == 0) {
        if (ause > 0) mbar_wait2(bars + kBarAEmpty + aslot, (ause + 1) & 1u);
        if (q.blk == 0 && a.pre_a && q.b != cv_b) {  // new utterance: its affine was written one chunk ago
          cv_b = q.b;
          ++cv_epoch;
          named_bar_sync(1, kTc3XformThreads);
        }
      }
      const uint32_t sA = smem_base + c.off_a + aslot * c.a_bytes;
      const uint32_t s_pa = smem_base + c.off_pa + (uint32_t)((cv_epoch % 3) * 2 * cpad) * 4u;
      const uint32_t s_pc = s_pa + (uint32_t)cpad * 4u;
      const int ci0 = q.blk * a.CIB;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int idx = q.ch * CHUNK + j * kTc3XformThreads + tt;
        if (idx < items) {
          const int g = (int)__umulhi((uint32_t)idx, w_magic), r = idx - g * W;
          const int ch = ci0 + g * 8;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.f;
          if (live & (1u << j)) {
            if (gen) {
              const float x0 = fmaxf(d[j][0].x, d[j][0].x * a.slope), x1 = fmaxf(d[j][0].y, d[j][0].y * a.slope),
                          x2 = fmaxf(d[j][0].z, d[j][0].z * a.slope);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float* gw = gen_w + ch + e;  // packed [tap][C_in]
                float y = __ldg(gen_b + ch + e);
                y = fmaf(__ldg(gw), x0, y);
                y = fmaf(__ldg(gw + a.C_in), x1, y);
                y = fmaf(__ldg(gw + 2 * a.C_in), x2, y);
                v[e] = y;
              }
            } else {
              v[0] = d[j][0].x; v[1] = d[j][0].y; v[2] = d[j][0].z; v[3] = d[j][0].w;
              v[4] = d[j][1].x; v[5] = d[j][1].y; v[6] = d[j][1].z; v[7] = d[j][1].w;
              if (a.pre_a) {
                const float4 a0 = lds128f(s_pa + (uint32_t)ch * 4u), a1 = lds128f(s_pa + (uint32_t)ch * 4u + 16u);
                const float4 c0 = lds128f(s_pc + (uint32_t)ch * 4u), c1 = lds128f(s_pc + (uint32_t)ch * 4u + 16u);
                affine8(v, a0, a1, c0, c1);
              }
            }
            if (a.pre_lrelu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], v[e] * a.slope);  // slope in (0, 1)
            }
          }
          split_store_s(sA + (uint32_t)g * strip + (uint32_t)r * 16u, plane, v);
        }
      }
      if (q.ch == n_chunks - 1) {
        fence_proxy_async();
        mbar_arrive(bars + kBarAFull + aslot);
        ++pa_pos;
      }
    };
    // software pipeline over the chunk sequence: chunk k+1 is in flight while chunk k is converted
    Cursor cur{first, 0, 0, 0, first / c.m_tiles, first % c.m_tiles};
    float4 dc[CH][2], dn[CH][2];
    uint32_t live_c, live_n;
    load_chunk(cur, dc, live_c);
    while (cur.m < n_m) {
      Cursor nxt = cur;
      advance(nxt);
      load_chunk(nxt, dn, live_n);
      convert_chunk(cur, dc, live_c);
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        dc[j][0] = dn[j][0];
        dc[j][1] = dn[j][1];
      }
      live_c = live_n;
      cur = nxt;
    }
   } else {
    // =============================== TRANSFORM (warp tasks) ===============================
    // A warp task = (8-channel group g, 32-row segment of the A window): lane <-> row, so g and every address term
    // but the lane are warp-uniform, all global / shared accesses are 512 contiguous bytes per warp instruction, and
    // the body is branch-free (rows outside the utterance are stored as zeros by a predicated store) so the up to
    // four tasks of a chunk interleave in the instruction stream.
    const int tt = tid - 32 * SH::kX0, xw = warp - SH::kX0;
    const uint32_t smem_base = smem_u32(smem);
    const float4* in4 = reinterpret_cast<const float4*>(a.in + prob * L.d_in);
    float* s_pa_base = reinterpret_cast<float*>(smem + c.off_pa);
    const int cpad = (a.C_in + 7) / 8 * 8;
    const int Gb = a.CIB >> 3;  // groups of 8 channels per (padded) ci block
    const uint32_t strip = (uint32_t)W * 16u, plane = (uint32_t)Gb * strip;
    // MODE 2: segments of source rows (at most ceil(W / up) + 1 of them touch a window)
    const int nseg = (MODE == 2 || MODE == 5) ? ((W + a.up - 1) / a.up + 1 + 31) >> 5 : (W + 31) >> 5, ntask = Gb * nseg;
    const long long Tp_in = ntc_tp(a.T_in);
    const long long ld4 = a.in_ld >> 2;
    const uint32_t up_magic = 0xFFFFFFFFu / (uint32_t)a.up + 1u;
    // Tile-invariant geometry.  With up == down == 1 ("direct") the source row of window row r is t0 - halo + r and
    // t0 is a multiple of 32, so its address splits into (tile base) + (per-thread term of lane - halo) + (per-task
    // term), the last one read from a shared-memory table: {global offset, smem offset, rows of the segment inside
    // the window, (g << 16) | first row of the segment}.
    const bool direct = a.up == 1 && a.down == 1;
    const int hl = lane - halo;
    const long long thr_goff = (long long)(hl & ~31) * ld4 + (hl & 31);  // float4 units
    int4* tab = reinterpret_cast<int4*>(smem + c.off_tab);
    if constexpr (MODE == 0) {
      for (int k = tt; k < ntask; k += kTc3XformThreads) {
        const int g = k / nseg, seg = k - g * nseg;
        tab[k] = make_int4((int)(seg * 32 * ld4) + g * 64, (int)(g * strip) + seg * 512, W - seg * 32, (g << 16) | (seg * 32));
      }
      named_bar_sync(1, kTc3XformThreads);
    }
    const uint32_t s_tab = smem_u32(tab);
    // staging ring: slot = one chunk; per task 2 x 512 B (the two 4-channel halves), lane-private 16 B pieces
    const uint32_t stg0 = smem_base + c.off_stg + (uint32_t)(xw * c.cpw) * 1024u + (uint32_t)lane * 16u;
    constexpr int CH = kTc3ChunkItems;
    const int rounds = (ntask + SH::kXW - 1) / SH::kXW;
    const int cpw = c.cpw;  // tasks per warp and chunk actually used (<= CH, the unroll bound)
    const int n_chunks = (rounds + cpw - 1) / cpw;
    const bool has_aff = a.pre_a != nullptr || a.pre_stats != nullptr;
    // statistics merge: fin_cw channels at a time, fin_P threads (segment phases) per channel
    const int fin_cw = min(a.C_in, kTc3XformThreads), fin_P = kTc3XformThreads / fin_cw;

    struct Cursor {
      int m, blk, ch, it, b, tile;  // (b, tile) track m without a division per chunk
    };
    const int step_b = step / c.m_tiles, step_t = step - step_b * c.m_tiles;
    auto advance = [&](Cursor& q) {
      if (++q.ch == n_chunks) {
        q.ch = 0;
        if (++q.blk == a.n_blk) {
          q.blk = 0;
          q.m += step;
          q.b += step_b;
          q.tile += step_t;
          if (q.tile >= c.m_tiles) {
            q.tile -= c.m_tiles;
            ++q.b;
          }
          ++q.it;
        }
      }
    };
    // Affine buffers are indexed by the number of utterance changes so far (three buffers: a rewrite follows the
    // named barrier of the previous change, which every thread passes only after converting the tiles before it).
    int pa_b = -1;
    uint32_t pa_wbuf = 0;  // affine buffer of the tile being loaded (== utterance changes % 3)
    int cv_b = -1;
    auto update_affine = [&](const Cursor& q) {
      const int b = q.b;
      if (q.blk == 0 && q.ch == 0 && has_aff && b != pa_b) {
        pa_b = b;
        pa_wbuf = pa_wbuf == 2 ? 0 : pa_wbuf + 1;
        float* s_pa = s_pa_base + pa_wbuf * 2 * cpad;
        if (a.pre_stats) {
          // InstanceNorm statistics of utterance b from the producer's per-segment (mean, M2) partials, merged here
          // in double in a fixed order (fastsvc.py:76,138: biased variance over the time axis, eps inside the sqrt).
          // Thread (part, ch) accumulates weighted moments of the segment means about the first segment's mean over
          // every fin_P-th segment; thread ch then adds the parts.
          double* scr = reinterpret_cast<double*>(smem + c.off_fin);
          for (int c0 = 0; c0 < a.C_in; c0 += fin_cw) {
            const int cw = min(fin_cw, a.C_in - c0);
            const int part = tt / cw, ch = c0 + tt - part * cw;
            double sw = 0.0, s1 = 0.0, s2 = 0.0, qq = 0.0;
            if (part < fin_P) {
              const float2* sp = a.pre_stats + (long long)b * a.pre_nseg * a.C_in + ch;
              const float pivf = __ldg(sp).x;
              for (int s0 = part; s0 < a.pre_nseg; s0 += 8 * fin_P) {  // 8 loads in flight, fp32 inside a batch
                float2 pv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const int sg = s0 + u * fin_P;
                  pv[u] = sg < a.pre_nseg ? __ldg(sp + (long long)sg * a.C_in) : make_float2(pivf, 0.f);
                }
                float fw = 0.f, f1 = 0.f, f2 = 0.f, fq = 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const int sg = s0 + u * fin_P;
                  const float nb = sg < a.pre_nseg ? (float)min(32, a.T_in - sg * 32) : 0.f;
                  const float d = pv[u].x - pivf;
                  fw += nb;
                  f1 = fmaf(nb, d, f1);
                  f2 = fmaf(nb * d, d, f2);
                  fq += pv[u].y;
                }
                sw += (double)fw;
                s1 += (double)f1;
                s2 += (double)f2;
                qq += (double)fq;
              }
              double* o = scr + (size_t)(part * cw + (ch - c0)) * 4;
              o[0] = sw; o[1] = s1; o[2] = s2; o[3] = qq;
            }
            named_bar_sync(2, kTc3XformThreads);
            if (tt < cw) {
              const int chh = c0 + tt;
              sw = s1 = s2 = qq = 0.0;
              for (int pp = 0; pp < fin_P; ++pp) {
                const double* o = scr + (size_t)(pp * cw + tt) * 4;
                sw += o[0]; s1 += o[1]; s2 += o[2]; qq += o[3];
              }
              const double piv = (double)__ldg(a.pre_stats + (long long)b * a.pre_nseg * a.C_in + chh).x;
              const double mean = piv + s1 / sw;
              const double var = fmax(qq + s2 - s1 * s1 / sw, 0.0) / sw;
              const double rstd = 1.0 / sqrt(var + (double)a.pre_eps);
              const double e = a.pre_e ? (double)__ldg(a.pre_e + (long long)b * a.C_in + chh) : 0.0;
              s_pa[chh] = (float)rstd;
              s_pa[cpad + chh] = (float)(e - mean * rstd);
            }
            if (c0 + fin_cw < a.C_in) named_bar_sync(2, kTc3XformThreads);  // scratch reuse by the next channel slab
          }
          for (int ch = a.C_in + tt; ch < cpad; ch += kTc3XformThreads) {
            s_pa[ch] = 1.f;
            s_pa[cpad + ch] = 0.f;
          }
        } else {
          for (int ch = tt; ch < cpad; ch += kTc3XformThreads) {
            s_pa[ch] = ch < a.C_in ? __ldg(a.pre_a + (long long)b * a.C_in + ch) : 1.f;
            s_pa[cpad + ch] = ch < a.C_in ? __ldg(a.pre_c + (long long)b * a.C_in + ch) : 0.f;
          }
        }
      }
    };
   if constexpr (MODE == 5) {
    // ---- MODE 5: lean nearest-repeat transform (up > 1, down == 1, one ci block, <= 4 tasks per warp and item):
    // MODE 2's arithmetic with the tile-invariant part of the geometry in registers.  A task = (8-channel group,
    // 32 SOURCE rows); source row sr feeds window rows sr*up - u_lo .. + up - 1.
    constexpr int NT = 4;
    int rl[NT];
    uint32_t gstrip[NT], gcol[NT], aoff[NT], on[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int k = xw + j * SH::kXW;
      const int g = k / nseg, seg = k - g * nseg;
      on[j] = (k < ntask && g * 8 < a.C_in) ? 1u : 0u;
      rl[j] = seg * 32 + lane;
      gstrip[j] = (uint32_t)g * strip;
      gcol[j] = (uint32_t)g * 64u;
      aoff[j] = (uint32_t)g * 32u;
    }
    for (int sl = 0; sl < c.a_slots; ++sl)  // channel-padding groups: zero once in every A slot
      for (int idx = tt; idx < Gb * W; idx += kTc3XformThreads) {
        const int g = idx / W, rw = idx - g * W;
        if (g * 8 >= a.C_in) {
          const uint32_t d = smem_base + c.off_a + (uint32_t)sl * c.a_bytes + (uint32_t)g * strip + (uint32_t)rw * 16u;
          sts128(d, make_uint4(0u, 0u, 0u, 0u));
          sts128(d + plane, make_uint4(0u, 0u, 0u, 0u));
        }
      }
    const int depth = c.stg_depth, up = a.up;
    const uint32_t sA0 = smem_base + c.off_a, s_pa0 = smem_base + c.off_pa;
    const uint32_t pa_stride = (uint32_t)(2 * cpad) * 4u, pc_off = (uint32_t)cpad * 4u;
    griddep_wait();  // first access to the predecessor's output
    if (tt == 0) FSVC_TL(L.tl_slot, 2);
    int am = first, ab = first / c.m_tiles, atile = first - ab * c.m_tiles;
    int cm = first, cb = ab, ctile = atile, it = 0;
    uint32_t slot_i = 0, slot_c = 0, aslot = 0, ause = 0, pa_buf = 0;
    auto issue = [&]() {
      if (am < n_m) {
        const int u_lo = atile * kTc2M - halo;
        const int s_first = (int)__umulhi((uint32_t)max(u_lo, 0), up_magic);
        const int s_last = (int)__umulhi((uint32_t)min(u_lo + W - 1, a.T_out - 1), up_magic);
        const long long rowbase = (long long)ab * Tp_in;
        const uint32_t dst0 = stg0 + slot_i * c.stg_bytes;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (on[j]) {  // warp-uniform, loop-invariant
            const int sr = s_first + rl[j];
            const bool ok = sr <= s_last;
            const float4* p = ok ? in4 + (rowbase + (sr & ~31)) * ld4 + ((uint32_t)(sr & 31) + gcol[j]) : in4;
            cp_async16(dst0 + (uint32_t)j * 1024u, p, ok ? 16u : 0u);
            cp_async16(dst0 + (uint32_t)j * 1024u + 512u, p + (ok ? 32 : 0), ok ? 16u : 0u);
          }
        }
        ++am;
        if (++atile == c.m_tiles) {
          atile = 0;
          ++ab;
        }
      }
      cp_async_commit();
      slot_i = slot_i == (uint32_t)depth ? 0u : slot_i + 1u;
    };
    for (int i = 0; i < depth; ++i) issue();
    if (cm < n_m) update_affine(Cursor{cm, 0, 0, 0, cb, ctile});
    while (cm < n_m) {
      int nb = cb, ntile = ctile + 1;
      if (ntile == c.m_tiles) {
        ntile = 0;
        ++nb;
      }
      if (cm + 1 < n_m) update_affine(Cursor{cm + 1, 0, 0, it + 1, nb, ntile});
      if (depth == 3) cp_async_wait<2>();
      else if (depth == 2) cp_async_wait<1>();
      else cp_async_wait<0>();
      if (tt == 0 && it == 0) FSVC_TL(L.tl_slot, 3);
      if (ause > 0) mbar_wait2(bars + kBarAEmpty + aslot, (ause + 1) & 1u);
      if (has_aff && cb != cv_b) {  // new utterance: its affine was written one item ago
        cv_b = cb;
        pa_buf = pa_buf == 2 ? 0 : pa_buf + 1;
        named_bar_sync(1, kTc3XformThreads);
      }
      {
        const uint32_t sA = sA0 + aslot * c.a_bytes;
        const uint32_t s_pa = s_pa0 + pa_buf * pa_stride, s_pc = s_pa + pc_off;
        const uint32_t src0 = stg0 + slot_c * c.stg_bytes;
        const int u_lo = ctile * kTc2M - halo;
        const int s_first = (int)__umulhi((uint32_t)max(u_lo, 0), up_magic);
        const int s_last = (int)__umulhi((uint32_t)min(u_lo + W - 1, a.T_out - 1), up_magic);
        const uint32_t rw_end = (uint32_t)min(W, a.T_out - u_lo);
        if (u_lo < 0 || u_lo + W > a.T_out) {  // zero padding rows of the window (first / last tile of an utterance)
          for (int idx = tt; idx < Gb * W; idx += kTc3XformThreads) {
            const int g = idx / W, rw = idx - g * W, u = u_lo + rw;
            if (u < 0 || u >= a.T_out) {
              sts128(sA + (uint32_t)g * strip + (uint32_t)rw * 16u, make_uint4(0u, 0u, 0u, 0u));
              sts128(sA + plane + (uint32_t)g * strip + (uint32_t)rw * 16u, make_uint4(0u, 0u, 0u, 0u));
            }
          }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (on[j]) {
            const float4 dA = lds128f(src0 + (uint32_t)j * 1024u), dB = lds128f(src0 + (uint32_t)j * 1024u + 512u);
            float v[8] = {dA.x, dA.y, dA.z, dA.w, dB.x, dB.y, dB.z, dB.w};
            if (has_aff) {
              const float4 a0 = lds128f(s_pa + aoff[j]), a1 = lds128f(s_pa + aoff[j] + 16u);
              const float4 c0 = lds128f(s_pc + aoff[j]), c1 = lds128f(s_pc + aoff[j] + 16u);
              affine8(v, a0, a1, c0, c1);
            }
            if (a.pre_lrelu) lrelu8(v, a.slope);  // slope in (0, 1)
            uint4 hi, lo;
            split_bf16(v, hi, lo);
            const int sr = s_first + rl[j];
            const int rw0 = sr * up - u_lo;  // window row of the first step this source row feeds (may be < 0)
            const uint32_t dst = sA + gstrip[j];
            if (sr <= s_last) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t rw = (uint32_t)(rw0 + i);
                if (i < up && rw < rw_end) {  // unsigned compare: negative rows fail it too
                  sts128(dst + rw * 16u, hi);
                  sts128(dst + plane + rw * 16u, lo);
                }
              }
            }
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(bars + kBarAFull + aslot);
      if (tt == 0 && it < 12) FSVC_TL(L.tl_slot, 4 + it);
      if (++aslot == (uint32_t)c.a_slots) {
        aslot = 0;
        ++ause;
      }
      slot_c = slot_c == (uint32_t)depth ? 0u : slot_c + 1u;
      issue();
      ++cm;
      ++it;
      cb = nb;
      ctile = ntile;
    }
    cp_async_wait<0>();
   } else if constexpr (MODE >= 3) {
    // ---- MODE 3 / 4: the lean transform of the many-tile layers (up == down == 1, one ci block, every warp's tasks of
    // an item in ONE chunk of at most NT).  Profiling the table-driven loop showed 726 instructions per item and warp for
    // three tasks at 0.2 IPC -- the role that paces every such layer -- most of them geometry, predicates and branches.
    // Here everything tile-invariant (global / shared offsets, window membership, affine row) is computed once into
    // registers; per item a task costs its address add, one bounds compare, two copies / two loads, the arithmetic
    // and two predicated stores.
    constexpr int NT = MODE == 4 ? 6 : 4;
    // Raw rows arrive by bulk copy (row loader in warp 0): per (item, ci block) a staging slot holds the <= 6 aligned
    // 32-row blocks covering the window -- the transform warps issue no loads and compute no global addresses.
    // Everything below the block index is tile- and block-invariant and lives in registers.
    int rel[NT];
    uint32_t stoff[NT], soff[NT], aoff[NT], flags[NT];  // flags: bit 0 = task exists, bit 1 = lane's row is in the window
    int gch[NT];                                        // first channel of the task's group inside its ci block
    const uint32_t row_bytes = (uint32_t)(a.CIB >> 2) * 512u;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int k = xw + j * SH::kXW;
      const int g = k / nseg, seg = k - g * nseg;
      const bool on = k < ntask;
      rel[j] = seg * 32 + hl;
      const int w32 = 32 + seg * 32 + hl;           // window row of this lane, counted from the first staged block
      stoff[j] = (uint32_t)(w32 >> 5) * row_bytes + (uint32_t)(2 * g) * 512u + (uint32_t)(w32 & 31) * 16u;
      soff[j] = (uint32_t)g * strip + (uint32_t)seg * 512u + (uint32_t)lane * 16u;
      aoff[j] = (uint32_t)g * 32u;
      gch[j] = g * 8;
      flags[j] = (on ? 1u : 0u) | ((on && seg * 32 + lane < W) ? 2u : 0u);
    }
    const int depth = c.stg_depth;  // ring of depth + 1 staging slots, one full / empty mbarrier pair each
    const uint32_t sA0 = smem_base + c.off_a, s_pa0 = smem_base + c.off_pa;
    const uint32_t pa_stride = (uint32_t)(2 * cpad) * 4u, pc_off = (uint32_t)cpad * 4u;
    const uint32_t stg_base = smem_base + c.off_stg;
    if (tt == 0) FSVC_TL(L.tl_slot, 2);
    int cm = first, cb = first / c.m_tiles, ctile = first - cb * c.m_tiles, it = 0;
    uint32_t slot_c = 0, use_c = 0, aslot = 0, ause = 0, pa_buf = 0;
    if (has_aff) griddep_wait();  // the affine below reads the predecessor's statistics
    if (cm < n_m) update_affine(Cursor{cm, 0, 0, 0, cb, ctile});
    while (cm < n_m) {
      // affine of the next item's utterance, one item ahead (as in the other modes)
      int nb = cb, ntile = ctile + 1;
      if (ntile == c.m_tiles) {
        ntile = 0;
        ++nb;
      }
      if (tt == 0 && it == 5) FSVC_TL(L.tl_slot, 41);
      if (cm + 1 < n_m) update_affine(Cursor{cm + 1, 0, 0, it + 1, nb, ntile});
      if (has_aff && cb != cv_b) {  // new utterance: its affine was written one item ago
        cv_b = cb;
        pa_buf = pa_buf == 2 ? 0 : pa_buf + 1;
        named_bar_sync(1, kTc3XformThreads);
      }
      const int t0 = ctile * kTc2M;
      for (int blk = 0; blk < a.n_blk; ++blk) {
        mbar_wait2(bars + kBarStgFull + slot_c, use_c & 1u);
        if (tt == 0 && it == 0 && blk == 0) FSVC_TL(L.tl_slot, 3);
        if (tt == 0 && it == 5 && blk == 0) FSVC_TL(L.tl_slot, 42);
        if (ause > 0) mbar_wait2(bars + kBarAEmpty + aslot, (ause + 1) & 1u);
        if (tt == 0 && it == 5 && blk == 0) FSVC_TL(L.tl_slot, 43);
        {
          const uint32_t sA = sA0 + aslot * c.a_bytes;
          const uint32_t s_pa = s_pa0 + pa_buf * pa_stride + (uint32_t)(blk * a.CIB) * 4u, s_pc = s_pa + pc_off;
          const uint32_t src0 = stg_base + slot_c * c.stg_bytes;
          const int c_left = a.C_in - blk * a.CIB;  // real channels from this block on
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            if (flags[j] & 1u) {
              // rows outside the utterance and channel padding are zero AFTER the prologue (their staging bytes may
              // be stale: never used)
              const uint32_t in_strip = (flags[j] >> 1) & 1u;
              const uint32_t real = ((unsigned)(t0 + rel[j]) < (unsigned)a.T_out && gch[j] < c_left) ? 1u : 0u;
              float4 dA = make_float4(0.f, 0.f, 0.f, 0.f), dB = dA;
              if (in_strip & real) {
                dA = lds128f(src0 + stoff[j]);
                dB = lds128f(src0 + stoff[j] + 512u);
              }
              float v[8] = {dA.x, dA.y, dA.z, dA.w, dB.x, dB.y, dB.z, dB.w};
              if (has_aff) {
                const float4 a0 = lds128f(s_pa + aoff[j]), a1 = lds128f(s_pa + aoff[j] + 16u);
                const float4 c0 = lds128f(s_pc + aoff[j]), c1 = lds128f(s_pc + aoff[j] + 16u);
                affine8(v, a0, a1, c0, c1);
              }
              if (a.pre_lrelu) lrelu8(v, a.slope);  // slope in (0, 1)
              split_store_p(sA + soff[j], plane, v, in_strip & real, in_strip & ~real);
            }
          }
        }
        if (tt == 0 && it == 5 && blk == 0) FSVC_TL(L.tl_slot, 44);
        fence_proxy_async();
        mbar_arrive(bars + kBarAFull + aslot);
        if (tt == 0 && (a.n_blk == 1 ? it : blk) < 12 && (a.n_blk == 1 || it == 0)) FSVC_TL(L.tl_slot, 4 + (a.n_blk == 1 ? it : blk));
        if (++aslot == (uint32_t)c.a_slots) {
          aslot = 0;
          ++ause;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + kBarStgEmpty + slot_c);  // this warp is done with the slot's rows
        if (slot_c == (uint32_t)depth) {
          slot_c = 0;
          ++use_c;
        } else {
          ++slot_c;
        }
      }
      if (tt == 0 && it == 5) FSVC_TL(L.tl_slot, 47);
      ++cm;
      ++it;
      cb = nb;
      ctile = ntile;
    }
   } else if constexpr (MODE == 2) {
    // ---- MODE 2: nearest-repeat input.  Window row rw <-> output-rate step u = u_lo + rw reads source row u / up,
    // so a source row s feeds the window rows of steps s*up .. s*up+up-1: load + convert it once, store it `up`
    // times.  Rows of the window outside [0, T_out) (first / last tile of an utterance) are zero-filled first.
    const uint32_t seg_magic = 0xFFFFFFFFu / (uint32_t)nseg + 1u;  // exact quotients below 2^16
    auto tile_rows = [&](const Cursor& q, int& u_lo, int& s_first, int& s_last) {
      u_lo = q.tile * kTc2M - halo;
      s_first = (int)__umulhi((uint32_t)max(u_lo, 0), up_magic);
      s_last = (int)__umulhi((uint32_t)min(u_lo + W - 1, a.T_out - 1), up_magic);
    };
    auto issue_chunk = [&](const Cursor& q, uint32_t slot) {
      if (q.m < n_m) {
        int u_lo, s_first, s_last;
        tile_rows(q, u_lo, s_first, s_last);
        const int cg0 = q.blk * Gb;
        const long long rowbase = (long long)q.b * Tp_in;
        const int kbase = q.ch * cpw * SH::kXW + xw;
        const uint32_t dst0 = stg0 + slot * c.stg_bytes;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int k = kbase + j * SH::kXW;
          if (j < cpw && k < ntask) {  // warp-uniform
            const int g = nseg == 1 ? k : (int)__umulhi((uint32_t)k, seg_magic), seg = k - g * nseg;
            const int sr = s_first + seg * 32 + lane;
            const bool ok = sr <= s_last && (cg0 + g) * 8 < a.C_in;
            const float4* p = ok ? in4 + (rowbase + (sr & ~31)) * ld4 + ((sr & 31) + (cg0 + g) * 64) : in4;
            cp_async16(dst0 + (uint32_t)j * 1024u, p, ok ? 16u : 0u);
            cp_async16(dst0 + (uint32_t)j * 1024u + 512u, p + (ok ? 32 : 0), ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
    };
    griddep_wait();  // first access to the predecessor's output
    if (tt == 0) FSVC_TL(L.tl_slot, 2);
    uint32_t aslot = 0, ause = 0, pa_buf = 0;  // A ring position; affine buffer of the tile being converted
    const uint32_t s_pa0 = smem_base + c.off_pa;
    auto convert_chunk = [&](const Cursor& q, uint32_t slot) {
      const uint32_t src0 = stg0 + slot * c.stg_bytes;
      int u_lo, s_first, s_last;
      tile_rows(q, u_lo, s_first, s_last);
      if (q.ch == 0) {
        if (ause > 0) mbar_wait2(bars + kBarAEmpty + aslot, (ause + 1) & 1u);
        if (q.blk == 0 && has_aff && q.b != cv_b) {  // new utterance: its affine was written one chunk ago
          cv_b = q.b;
          pa_buf = pa_buf == 2 ? 0 : pa_buf + 1;
          named_bar_sync(1, kTc3XformThreads);
        }
        if (u_lo < 0 || u_lo + W > a.T_out) {  // zero padding rows of the window (both planes, every group)
          const uint32_t sz = smem_base + c.off_a + aslot * c.a_bytes;
          for (int idx = tt; idx < Gb * W; idx += kTc3XformThreads) {
            const int g = idx / W, rw = idx - g * W, u = u_lo + rw;
            if (u < 0 || u >= a.T_out) {
              sts128(sz + (uint32_t)g * strip + (uint32_t)rw * 16u, make_uint4(0u, 0u, 0u, 0u));
              sts128(sz + plane + (uint32_t)g * strip + (uint32_t)rw * 16u, make_uint4(0u, 0u, 0u, 0u));
            }
          }
        }
      }
      const uint32_t sA = smem_base + c.off_a + aslot * c.a_bytes;
      const uint32_t s_pa = s_pa0 + pa_buf * (uint32_t)(2 * cpad) * 4u + (uint32_t)(q.blk * Gb) * 32u;
      const uint32_t s_pc = s_pa + (uint32_t)cpad * 4u;
      const int cg0 = q.blk * Gb;
      const int kbase = q.ch * cpw * SH::kXW + xw;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int k = kbase + j * SH::kXW;
        if (j < cpw && k < ntask) {  // warp-uniform
          const int g = nseg == 1 ? k : (int)__umulhi((uint32_t)k, seg_magic), seg = k - g * nseg;
          const float4 dA = lds128f(src0 + (uint32_t)j * 1024u), dB = lds128f(src0 + (uint32_t)j * 1024u + 512u);
          float v[8] = {dA.x, dA.y, dA.z, dA.w, dB.x, dB.y, dB.z, dB.w};
          if ((cg0 + g) * 8 < a.C_in) {  // channel padding of the last ci block stays zero
            if (has_aff) {
              const uint32_t ca = (uint32_t)g * 32u;
              const float4 a0 = lds128f(s_pa + ca), a1 = lds128f(s_pa + ca + 16u);
              const float4 c0 = lds128f(s_pc + ca), c1 = lds128f(s_pc + ca + 16u);
              affine8(v, a0, a1, c0, c1);
            }
            if (a.pre_lrelu) lrelu8(v, a.slope);  // slope in (0, 1)
          }
          uint4 hi, lo;
          split_bf16(v, hi, lo);
          const int sr = s_first + seg * 32 + lane;
          const int rw0 = sr * a.up - u_lo;  // window row of the first step this source row feeds
          const int rw_end = min(W, a.T_out - u_lo);
          const uint32_t dst = sA + (uint32_t)g * strip;
          if (sr <= s_last) {
            for (int i = 0; i < a.up; ++i) {
              const int rw = rw0 + i;
              if (rw >= 0 && rw < rw_end) {
                sts128(dst + (uint32_t)rw * 16u, hi);
                sts128(dst + plane + (uint32_t)rw * 16u, lo);
              }
            }
          }
        }
      }
      if (q.ch == n_chunks - 1) {
        fence_proxy_async();
        mbar_arrive(bars + kBarAFull + aslot);
        if (tt == 0 && q.it == 0 && q.blk < 12) FSVC_TL(L.tl_slot, 4 + q.blk);
        if (++aslot == (uint32_t)c.a_slots) {
          aslot = 0;
          ++ause;
        }
      }
    };
    // Software pipeline over the chunk sequence: the raw rows of the next `stg_depth` chunks are in flight (cp.async
    // into a per-lane staging ring, so no registers are tied up and no barrier is needed: a lane reads back only what
    // it copied itself) while chunk i is converted.  The InstanceNorm affine is prepared one chunk ahead.
    {
      Cursor cur{first, 0, 0, 0, first / c.m_tiles, first % c.m_tiles};
      Cursor ahead = cur;
      const int depth = c.stg_depth;
      uint32_t slot_i = 0, slot_c = 0;
      for (int i = 0; i < depth; ++i) {
        issue_chunk(ahead, slot_i);
        slot_i = slot_i + 1 == (uint32_t)depth + 1 ? 0 : slot_i + 1;
        advance(ahead);
      }
      if (cur.m < n_m) update_affine(cur);
      while (cur.m < n_m) {
        issue_chunk(ahead, slot_i);
        slot_i = slot_i + 1 == (uint32_t)depth + 1 ? 0 : slot_i + 1;
        advance(ahead);
        Cursor nxt = cur;
        advance(nxt);
        if (nxt.m < n_m) update_affine(nxt);
        cp_async_wait_n(depth);
        if (tt == 0 && cur.it == 0 && cur.blk == 0 && cur.ch == 0) FSVC_TL(L.tl_slot, 3);
        convert_chunk(cur, slot_c);
        slot_c = slot_c + 1 == (uint32_t)depth + 1 ? 0 : slot_c + 1;
        cur = nxt;
      }
      cp_async_wait<0>();
    }
   } else {
    // issue the loads of one chunk (no use of the results); bit j of `live`: this lane's row of task j is real data
    auto issue_chunk = [&](const Cursor& q, uint32_t slot) {
      if (q.m < n_m) {
        const int t0 = q.tile * kTc2M, ut = t0 + hl;
        const int cg0 = q.blk * Gb;
        const long long rowbase = (long long)q.b * Tp_in;
        const float4* tb = in4 + (rowbase + t0) * ld4 + thr_goff + cg0 * 64;
        const int kbase = q.ch * cpw * SH::kXW + xw;
        const uint32_t dst0 = stg0 + slot * c.stg_bytes;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int k = kbase + j * SH::kXW;
          if (j < cpw && k < ntask) {  // warp-uniform
            const int4 e = lds128i(s_tab + (uint32_t)k * 16u);
            const int g = e.w >> 16, u = ut + (e.w & 0xffff);
            const bool ok = lane < e.z && (unsigned)u < (unsigned)a.T_out && (cg0 + g) * 8 < a.C_in;
            const float4* p = tb + e.x;
            if (!direct) {
              const int src = max(u, 0) * a.down;  // MODE 0: up == 1
              p = in4 + (rowbase + (src & ~31)) * ld4 + ((src & 31) + (cg0 + g) * 64);
            }
            if (!ok) p = in4;
            cp_async16(dst0 + (uint32_t)j * 1024u, p, ok ? 16u : 0u);
            cp_async16(dst0 + (uint32_t)j * 1024u + 512u, p + (ok ? 32 : 0), ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
    };
    griddep_wait();  // first access to the predecessor's output
    if (tt == 0) FSVC_TL(L.tl_slot, 2);
    uint32_t aslot = 0, ause = 0, pa_buf = 0;  // A ring position; affine buffer of the tile being converted
    const uint32_t sA0 = smem_base + c.off_a + (uint32_t)lane * 16u;
    const uint32_t s_pa0 = smem_base + c.off_pa;
    auto convert_chunk = [&](const Cursor& q, uint32_t slot) {
      const uint32_t src0 = stg0 + slot * c.stg_bytes;
      const int ut = q.tile * kTc2M + hl;
      if (q.ch == 0) {
        if (ause > 0) mbar_wait2(bars + kBarAEmpty + aslot, (ause + 1) & 1u);
        if (q.blk == 0 && has_aff && q.b != cv_b) {  // new utterance: its affine was written one chunk ago
          cv_b = q.b;
          pa_buf = pa_buf == 2 ? 0 : pa_buf + 1;  // == (number of utterance changes) % 3, as on the writing side
          named_bar_sync(1, kTc3XformThreads);
        }
      }
      const uint32_t sA = sA0 + aslot * c.a_bytes;
      const uint32_t s_pa = s_pa0 + pa_buf * (uint32_t)(2 * cpad) * 4u + (uint32_t)(q.blk * Gb) * 32u;
      const uint32_t s_pc = s_pa + (uint32_t)cpad * 4u;
      const int kbase = q.ch * cpw * SH::kXW + xw;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int k = kbase + j * SH::kXW;
        if (j < cpw && k < ntask) {  // warp-uniform
          const int4 e = lds128i(s_tab + (uint32_t)k * 16u);
          const float4 dA = lds128f(src0 + (uint32_t)j * 1024u), dB = lds128f(src0 + (uint32_t)j * 1024u + 512u);
          float v[8] = {dA.x, dA.y, dA.z, dA.w, dB.x, dB.y, dB.z, dB.w};
          if (has_aff) {
            const uint32_t ca = (uint32_t)(e.w >> 16) * 32u;
            const float4 a0 = lds128f(s_pa + ca), a1 = lds128f(s_pa + ca + 16u);
            const float4 c0 = lds128f(s_pc + ca), c1 = lds128f(s_pc + ca + 16u);
            affine8(v, a0, a1, c0, c1);
          }
          if (a.pre_lrelu) lrelu8(v, a.slope);  // slope in (0, 1)
          // rows outside the utterance and the channel padding of the last ci block are zero AFTER the prologue
          const uint32_t in_strip = lane < e.z ? 1u : 0u;
          const uint32_t real = ((unsigned)(ut + (e.w & 0xffff)) < (unsigned)a.T_out && (q.blk * Gb + (e.w >> 16)) * 8 < a.C_in) ? 1u : 0u;
          split_store_p(sA + (uint32_t)e.y, plane, v, in_strip & real, in_strip & ~real);
        }
      }
      if (q.ch == n_chunks - 1) {
        fence_proxy_async();
        mbar_arrive(bars + kBarAFull + aslot);
        if (tt == 0 && q.it == 0 && q.blk < 12) FSVC_TL(L.tl_slot, 4 + q.blk);
        if (++aslot == (uint32_t)c.a_slots) {
          aslot = 0;
          ++ause;
        }
      }
    };
    // Software pipeline over the chunk sequence: the raw rows of the next `stg_depth` chunks are in flight (cp.async
    // into a per-lane staging ring, so no registers are tied up and no barrier is needed: a lane reads back only what
    // it copied itself) while chunk i is converted.  The InstanceNorm affine is prepared one chunk ahead.
    {
      Cursor cur{first, 0, 0, 0, first / c.m_tiles, first % c.m_tiles};
      Cursor ahead = cur;
      const int depth = c.stg_depth;
      uint32_t slot_i = 0, slot_c = 0;
      for (int i = 0; i < depth; ++i) {
        issue_chunk(ahead, slot_i);
        slot_i = slot_i + 1 == (uint32_t)depth + 1 ? 0 : slot_i + 1;
        advance(ahead);
      }
      if (cur.m < n_m) update_affine(cur);
      while (cur.m < n_m) {
        issue_chunk(ahead, slot_i);
        slot_i = slot_i + 1 == (uint32_t)depth + 1 ? 0 : slot_i + 1;
        advance(ahead);
        Cursor nxt = cur;
        advance(nxt);
        if (nxt.m < n_m) update_affine(nxt);
        cp_async_wait_n(depth);
        if (tt == 0 && cur.it == 0 && cur.blk == 0 && cur.ch == 0) FSVC_TL(L.tl_slot, 3);
        convert_chunk(cur, slot_c);
        slot_c = slot_c + 1 == (uint32_t)depth + 1 ? 0 : slot_c + 1;
        cur = nxt;
      }
      cp_async_wait<0>();
    }
   }
   }
  } else if (warp >= SH::kE0) {
    // =============================== EPILOGUE ===============================
    const int ew = warp - SH::kE0;
    const int q = warp & 3, h = ew >> 2;
    // Two instantiations of the epilogue loop: PLAIN (bias, optional residual, optional LeakyReLU, one store: every
    // conditioning conv, conv_first, the stage residual conv) and the general one (FiLM affine, raw copy, statistics,
    // folded conv_last, generated residual).  In the wide layers a thread walks 4..8 sub-tiles per item on its own, at
    // roughly one instruction per 4 cycles: the ~150 predicated-off instructions of the unused options were most of a
    // plain sub-tile's ~0.5 us.
    auto epilogue = [&](auto plain_c) {
      constexpr bool PLAIN = decltype(plain_c)::value;
      const float* bias = a.bias + prob * L.d_bias;
      const float* res = a.res ? a.res + prob * L.d_res : nullptr;
      // generated 1-channel residual (un-fused level-0 chain): only the run-time-width instantiation (NH4 == 0) carries it
      const float* gres_w = (!PLAIN && NH4 == 0 && a.gres_w) ? a.gres_w + prob * L.d_gres_w : nullptr;
      const float* gres_b = gres_w ? a.gres_b + prob * L.d_gres_b : nullptr;
      const float* gres_x = gres_w ? a.gres_x + prob * L.d_gres_x : nullptr;
      float* raw = (!PLAIN && a.raw) ? a.raw + prob * L.d_raw : nullptr;
      uint4* out_pl = (PLAIN && a.out_pl) ? a.out_pl + prob * L.d_out_pl : nullptr;  // operand planes (ntc_common.cuh)
      float* out_dec = (PLAIN && a.out_dec) ? a.out_dec + prob * L.d_out_dec : nullptr;
      const long long Tp_dec = out_dec ? ntc_tp(a.out_dec_T) : 0;
      float* out = a.out ? a.out + prob * L.d_out : nullptr;
      const bool has_film = !PLAIN && a.gamma != nullptr, has_res = res != nullptr, has_stats = !PLAIN && a.stats != nullptr;
      const bool has_last = !PLAIN && a.last_w != nullptr;
      const uint32_t scr = smem_u32(smem + c.off_scr) + (uint32_t)(ew * (32 * c.scr_pitch + 16)) * 4u;
      const uint32_t scr_row = scr + (uint32_t)(lane * c.scr_pitch) * 4u + (lane >= 16 ? 64u : 0u);
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      const int step_b = step / c.m_tiles, step_t = step - step_b * c.m_tiles;
      const int rl = q * 32 + lane;  // row of this thread inside the tile
      const long long Tp_out = ntc_tp(a.T_out);

      // channels of this thread in sub-tile `sub`: nh (multiple of 4), starting at column sub*nsub + h*nh
      auto unit_nh = [&](int sub) { return NH4 ? 4 * NH4 : (min(c.nsub, nvalid - sub * c.nsub) / SH::kEH); };

      // operands of one (tile, sub-tile) unit for this thread: requested one unit ahead of their use
      float4 o_res[4], o_ga[4], o_be[4];
      float o_gx = 0.f;
      // Every tensor of the epilogue is blocked channels-last over the same (utterance, step) rows: one warp-uniform
      // row-block index per item, then one multiply-add per tensor (the per-tensor 64-bit row arithmetic was ~100 of the
      // ~520 instructions a warp spent per item).
      // (fits 32 bits: the library refuses batches of 2^31 or more padded steps, tc_forward.cu)
      const uint32_t Tp32 = (uint32_t)Tp_out;
      auto row_block = [&](int b, int tile) { return (unsigned long long)((uint32_t)b * Tp32 + (uint32_t)(tile * kTc2M + q * 32)); };
      auto load_ops = [&](int b, int tile, int sub) {
        const int t = tile * kTc2M + rl;
        if (t >= a.T_out) return;
        const int nh = unit_nh(sub);
        const int co = co_tile + sub * c.nsub + h * nh;
        const unsigned long long rb = row_block(b, tile);
        if (gres_w) o_gx = __ldg(gres_x + (long long)b * a.T_out + t);
        if (has_res) {
          const float4* p = reinterpret_cast<const float4*>(res + rb * (uint32_t)a.res_ld + (uint32_t)((co >> 2) * 128)) + lane;
  #pragma unroll
          for (int j = 0; j < 4; ++j)
            if (NH4 ? j < NH4 : 4 * j < nh) o_res[j] = __ldg(p + 32 * j);
        }
        if (has_film) {
          const unsigned long long ro = rb * (uint32_t)a.gb_ld + (uint32_t)((co >> 2) * 128 + lane * 4);
          const float4* pg = reinterpret_cast<const float4*>(a.gamma + ro);
          const float4* pb = reinterpret_cast<const float4*>(a.beta + ro);
  #pragma unroll
          for (int j = 0; j < 4; ++j)
            if (NH4 ? j < NH4 : 4 * j < nh) {
              o_ga[j] = __ldg(pg + 32 * j);
              o_be[j] = __ldg(pb + 32 * j);
            }
        }
      };
      // The N tile's bias goes to shared memory once (it is a weight: no need to wait for the predecessor).  Read from
      // global in every sub-tile it cost a cold L2 round trip per sub-tile: ~0.8 us x 8 sub-tiles in a 192-channel
      // layer whose CTAs process a single item.
      const uint32_t s_bias = smem_u32(smem + c.off_bias);
      {
        const int et = tid - SH::kE0 * 32;
        for (int i = et; i < nvalid; i += kTc3EpiThreads)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_bias + 4u * i), "f"(__ldg(bias + co_tile + i)) : "memory");
        named_bar_sync(3, kTc3EpiThreads);
      }
      griddep_wait();  // before the first operand load and the first store
      if (ew == 0 && lane == 0) FSVC_TL(L.tl_slot, 37);
      int b = first / c.m_tiles, tile = first - b * c.m_tiles;
      if (first < n_m) load_ops(b, tile, 0);
      int it = 0;
      for (int m = first; m < n_m; m += step, ++it) {
        // next item of this CTA (no division in the loop)
        int nb = b + step_b, ntile = tile + step_t;
        if (ntile >= c.m_tiles) {
          ntile -= c.m_tiles;
          ++nb;
        }
        const int t0 = tile * kTc2M;
        const uint32_t acc = (uint32_t)it & 1u;
        const int t = t0 + rl;
        const bool ok = t < a.T_out;
        const int n_rows_seg = min(32, a.T_out - (t0 + q * 32));
        const unsigned long long rb = row_block(b, tile);
        float* raw_row = raw ? raw + rb * (uint32_t)a.raw_ld + lane * 4 : nullptr;
        float* out_row = out ? out + rb * (uint32_t)a.out_ld + lane * 4 : nullptr;
        // decimated copy: this thread's step is kept when it is a multiple of r
        float* dec_row = nullptr;
        if (out_dec && ok) {
          const int td = t / a.out_dec_r;
          if (td * a.out_dec_r == t) dec_row = out_dec + ntc_row(Tp_dec, a.out_dec_ld, b, td);
        }
        // this thread's row in group 0 of the operand planes; pad_rows: 0, or -/+ kPlPad when the thread also owns a
        // zero row in front of the utterance (first tile, first 8 steps) / behind its last tile
        uint4* pl_row = out_pl ? out_pl + (long long)b * a.out_pl_G * a.out_pl_Tp + (kPlPad + t) : nullptr;
        const int pad_rows = !out_pl ? 0 : (tile == 0 && rl < kPlPad) ? -kPlPad
                                         : (tile == c.m_tiles - 1 && rl >= kTc2M - kPlPad) ? kPlPad : 0;
        float2* st_row = has_stats ? a.stats + ((long long)b * a.n_seg + (t0 >> 5) + q) * a.C_out : nullptr;
        if (ew == 0 && lane == 0 && it == 5) FSVC_TL(L.tl_slot, 48);
        mbar_wait2(bars + kBarAccFull + acc, ((uint32_t)it >> 1) & 1u);
        if (ew == 0 && lane == 0 && it == 0) FSVC_TL(L.tl_slot, 38);
        if (ew == 0 && lane == 0 && it == 5) FSVC_TL(L.tl_slot, 49);
        tc_fence_after();
        const uint32_t tacc = tmem + acc * acc_stride + lane_addr;
        for (int sub = 0; sub < n_sub; ++sub) {
          const int nh = unit_nh(sub);
          const int col = sub * c.nsub + h * nh;  // first column inside the N tile
          const int co = co_tile + col;
          const bool tl_u = ew == 0 && lane == 0 && it == 0 && sub == (n_sub > 1 ? 1 : 0);
          if (tl_u) FSVC_TL(L.tl_slot, 54);
          float v[16];
  #pragma unroll
          for (int j = 0; j < 4; ++j)
            if (NH4 ? j < NH4 : 4 * j < nh) tmem_ld4_nowait(tacc + (uint32_t)(col + 4 * j), v + 4 * j);
          float4 b4[4];
  #pragma unroll
          for (int j = 0; j < 4; ++j)
            if (NH4 ? j < NH4 : 4 * j < nh) b4[j] = lds128f(s_bias + 4u * (uint32_t)col + 16u * j);
          tmem_ld_wait();
          if (tl_u) FSVC_TL(L.tl_slot, 55);
          const float gx = o_gx;
  #pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (NH4 ? j < NH4 : 4 * j < nh) {
              float x[4] = {v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]};
              add2(x[0], x[1], b4[j].x, b4[j].y);
              add2(x[2], x[3], b4[j].z, b4[j].w);
              if (has_res && ok) {
                add2(x[0], x[1], o_res[j].x, o_res[j].y);
                add2(x[2], x[3], o_res[j].z, o_res[j].w);
              }
              if (!PLAIN && NH4 == 0 && gres_w) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(gres_w + co) + j);
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(gres_b + co) + j);
                x[0] += fmaf(w4.x, gx, c4.x); x[1] += fmaf(w4.y, gx, c4.y);
                x[2] += fmaf(w4.z, gx, c4.z); x[3] += fmaf(w4.w, gx, c4.w);
              }
              if (raw_row && ok) reinterpret_cast<float4*>(raw_row + (co >> 2) * 128)[32 * j] = make_float4(x[0], x[1], x[2], x[3]);
              if (a.post_lrelu) {
                lrelu2(x[0], x[1], a.slope);
                lrelu2(x[2], x[3], a.slope);
              }
              if (has_film) {
                if (ok) {
                  fma2(x[0], x[1], o_ga[j].x, o_ga[j].y, o_be[j].x, o_be[j].y);
                  fma2(x[2], x[3], o_ga[j].z, o_ga[j].w, o_be[j].z, o_be[j].w);
                } else {
                  x[0] = x[1] = x[2] = x[3] = 0.f;
                }
              }
              if (out_row && ok) reinterpret_cast<float4*>(out_row + (co >> 2) * 128)[32 * j] = make_float4(x[0], x[1], x[2], x[3]);
              if (PLAIN && dec_row) reinterpret_cast<float4*>(dec_row + (co >> 2) * 128)[32 * j] = make_float4(x[0], x[1], x[2], x[3]);
              if (PLAIN && pl_row) {
                // 4 channels = half of a (step, group) chunk in each plane; steps past the utterance's end are zeros
                float y[4] = {x[0], x[1], x[2], x[3]};
                if (!ok) {
                  y[0] = y[1] = y[2] = y[3] = 0.f;
                } else if (a.out_pl_lrelu) {
                  lrelu2(y[0], y[1], a.slope);
                  lrelu2(y[2], y[3], a.slope);
                }
                uint32_t h0, l0, h1, l1;
                split_pair(y[0], y[1], h0, l0);
                split_pair(y[2], y[3], h1, l1);
                const int cq = co + 4 * j;
                uint2* d = reinterpret_cast<uint2*>(pl_row + (long long)(cq >> 3) * a.out_pl_Tp) + ((cq >> 2) & 1);
                d[0] = make_uint2(h0, h1);
                d[2 * a.out_pl_lo] = make_uint2(l0, l1);
                if (pad_rows != 0) {
                  d[2 * pad_rows] = make_uint2(0u, 0u);
                  d[2 * (a.out_pl_lo + pad_rows)] = make_uint2(0u, 0u);
                }
              }
              v[4 * j] = x[0]; v[4 * j + 1] = x[1]; v[4 * j + 2] = x[2]; v[4 * j + 3] = x[3];
            }
          }
          if (tl_u) FSVC_TL(L.tl_slot, 56);
          if (has_last && ok) {
            // conv_last (Conv1d1x1, fastsvc.py:301,330) on the value just produced: this thread's channels' share of every
            // output channel; the other column half adds its share (two commutative adds onto zero: deterministic)
            for (int o = 0; o < a.last_co; ++o) {
              float p = (h == 0) ? __ldg(a.last_b + o) : 0.f;
  #pragma unroll
              for (int j = 0; j < 16; ++j)
                if (NH4 ? j < 4 * NH4 : j < nh) p = fmaf(v[j], __ldg(a.last_w + (long long)(co + j) * a.last_co + o), p);
              atomicAdd(a.last_out + ((long long)b * a.last_co + o) * a.T_out + t, p);
            }
          }
          // request the next unit's operands now; they land while this thread waits for the next accumulator
          if (sub + 1 < n_sub) load_ops(b, tile, sub + 1);
          else if (m + step < n_m) load_ops(nb, ntile, 0);
          if (tl_u) FSVC_TL(L.tl_slot, 57);
          if (has_stats && n_rows_seg > 0) {
            // (mean, M2) of the stored values over this warp's <= 32 rows, per channel: transpose through the warp's
            // smem scratch (rows 16..31 shifted by 16 floats so the two half-warps read disjoint banks), then lane
            // (half, ch) sums its 16 rows about the segment's first sample and the halves are added.
            // in_finalize2_kernel / the consumer's transform role merge the segments in double.
            __syncwarp();
  #pragma unroll
            for (int j = 0; j < 4; ++j)
              if (NH4 ? j < NH4 : 4 * j < nh)
                sts128(scr_row + 16u * j, make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                     __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
            __syncwarp();
            const int hh = lane >> 4, ch = lane & 15;
            const int cnt = max(0, min(16, n_rows_seg - 16 * hh));
            float piv = 0.f, s1 = 0.f, s2 = 0.f;
            if (ch < nh) {
              const uint32_t pitch_b = (uint32_t)c.scr_pitch * 4u;
              const uint32_t colp = scr + (uint32_t)ch * 4u + (uint32_t)hh * (16u * pitch_b + 64u);
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(piv) : "r"(scr + (uint32_t)ch * 4u));
              if (cnt == 16) {
                float xv[16];
  #pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv[i]) : "r"(colp + i * pitch_b));
  #pragma unroll
                // even and odd rows accumulate side by side (packed fp32), added at the end: a fixed order
                float s1b = 0.f, s2b = 0.f;
  #pragma unroll
                for (int i = 0; i < 16; i += 2) {
                  float d0, d1;
                  sub2(d0, d1, xv[i], xv[i + 1], piv, piv);
                  add2(s1, s1b, d0, d1);
                  fma2_acc(s2, s2b, d0, d1, d0, d1);
                }
                s1 += s1b;
                s2 += s2b;
              } else {
                for (int i = 0; i < cnt; ++i) {
                  float xi;
                  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xi) : "r"(colp + i * pitch_b));
                  const float dd = xi - piv;
                  s1 += dd;
                  s2 = fmaf(dd, dd, s2);
                }
              }
            }
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
            if (hh == 0 && ch < nh) {
              float rn = 0.03125f;
              if (n_rows_seg < 32) rn = __frcp_rn((float)n_rows_seg);  // ragged last segment of an utterance only
              const float dm = s1 * rn;
              st_row[co + ch] = make_float2(piv + dm, fmaxf(fmaf(-s1, dm, s2), 0.f));
            }
          }
          if (tl_u) FSVC_TL(L.tl_slot, 58);
        }
        if (ew == 0 && lane == 0 && it == 5) FSVC_TL(L.tl_slot, 50);
        tc_fence_before();
        mbar_arrive(bars + kBarAccEmpty + acc);
        if (ew == 0 && lane == 0 && it == 0) FSVC_TL(L.tl_slot, 39);
        b = nb;
        tile = ntile;
      }
    };
    if (!a.gamma && !a.stats && !a.raw && !a.last_w && !(NH4 == 0 && a.gres_w)) epilogue(BoolC<true>{});
    else epilogue(BoolC<false>{});
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) FSVC_TL(L.tl_slot, 40);
  if (warp == SH::kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem, 2 * acc_stride);
  }
}

// (B, C, T) -> blocked channels-last: the caller's PPG tensor into the layout the transform role reads.
__global__ void __launch_bounds__(256) nct_to_ntc_kernel(const float* __restrict__ x, int C, int T, float* __restrict__ y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < T) ? x[((long long)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < C) y[ntc_row(ntc_tp(T), C, b, t) + ntc_col(c)] = tile[tx][i];
  }
}

}  // namespace fsvc
