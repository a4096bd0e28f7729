// Host-side state shared by the translation units of libfsvc.so:
//   fsvc_abi.cu    C ABI (include/fsvc.h), fp32 forward, block-level entry points, excitation / PCM-16
//   tc_forward.cu  tensor-core forward (tcgen05 kernels, channels-last workspace, launch plan)
//   train.cu       training forward (keeps activations) and the native backward
// Nothing here is visible outside the library.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fsvc.h"

namespace fsvc {

int fail(int code, const char* fmt, ...);  // sets the thread-local fsvc_last_error() text, returns `code`

#define FSVC_CUDA(expr)                                                                               \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess) return ::fsvc::fail(FSVC_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

// packed tensor-core weights of one conv (tc_prims.cuh: pack_tc_weights_kernel)
struct TcW {
  const __nv_bfloat16* w = nullptr;  // [n_tile][ci_blk][hi|lo][tap][CIB/8][N_tile][8]
  int K = 0, CIB = 0, n_blk = 0, N_tile = 0, n_ntiles = 0, N_alloc = 0;
  size_t elems() const { return (size_t)n_ntiles * n_blk * 2 * K * CIB * N_tile; }
  size_t chunk_elems() const { return (size_t)2 * K * CIB * N_tile; }
};

struct ConvW {  // packed [C_in][K][C_out] + bias[C_out], device
  float* w = nullptr;
  float* b = nullptr;
  float* wT = nullptr;  // [C_out][K flipped][C_in]: the same conv transposed, i.e. its data-gradient conv (train.cu)
  int C_in = 0, C_out = 0, K = 0;
  TcW tc2;  // copy tiled for the persistent channels-last kernel (conv_tc3.cuh); tc2.w == nullptr: not eligible
  int tc2_resident = 0;
  int ctx_dil = 1, ctx_up = 1;  // how the generator uses this conv (dilation, upsampling factor of its input)
  int ctx_wide = 0;             // conditioning conv: prefer one wide N tile
  const __nv_bfloat16* wnc = nullptr;  // [tap][group][hi|lo rows][8] copy for the fused level kernel
  int nc_G = 0, nc_N = 0;
};

// One weight-preparation job.  fsvc_set_weights runs three launches over device-resident job tables built once per
// handle (instead of ~290 tiny launches): caller tensors -> packed fp32 store (repack / bias / copy), then packed fp32
// -> transposed copies (train.cu) and -> tensor-core bf16 hi|lo layouts (tc_forward.cu).
struct WJob {
  int kind;            // 0 repack, 1 bias (a [+ b]), 2 copy, 3 transpose, 4 pack_tc, 5 pack_tc_nc
  int src_a, src_b;    // indices into the caller's pointer list (kinds 0-2); -1 = none
  int C_out, C_in, K;
  int dst_ld, ci_off, co_off;  // repack destination geometry
  int CIB, n_blk, N_tile, n_ntiles, G, N;  // tensor-core layouts (kinds 4, 5)
  const float* w;      // packed fp32 source (kinds 3-5)
  void* dst;
  long long total;     // elements the job loops over
};
constexpr int kMaxWeightTensors = 352;  // 8 stages: 14 per stage + 2 * 8 * (8 + 6) + 2
struct WSrc {
  const float* p[kMaxWeightTensors];
};

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

struct fsvc_handle {
  fsvc_config cfg;
  int n = 0;
  int dscale[FSVC_MAX_STAGES];
  int lvl_c[FSVC_MAX_STAGES];
  int hop = 1;
  std::vector<fsvc::WeightInfo> winfo;
  float* store = nullptr;
  size_t store_floats = 0;
  __nv_bfloat16* tc_store = nullptr;  // tensor-core (bf16 hi/lo) copies of the conv weights
  size_t tc_elems = 0;
  bool tc2_ok = false;                // every conv of the forward can run on conv_tc3_kernel
  bool l0_fused = false;              // level 0 runs as the fused kernel (level_fused.cuh)
  int num_sms = 148;
  std::vector<fsvc::ConvW*> convs;    // every conv of the generator (for the repacks)
  fsvc::StageW stage[FSVC_MAX_STAGES];
  fsvc::LevelW level[FSVC_MAX_STAGES];
  fsvc::ConvW last;
  fsvc::WJob* jobs_a = nullptr;  // device job tables of fsvc_set_weights (phase A: from the caller's tensors;
  fsvc::WJob* jobs_t = nullptr;  // transposes; tensor-core packs)
  fsvc::WJob* jobs_tc = nullptr;
  int n_jobs_a = 0, n_jobs_t = 0, n_jobs_tc = 0;
  bool weights_set = false;
  int launches = 0;
  int device = 0;
  // fsvc_forward_host only: the PPG upload runs on a side stream while the conditioning levels (which need only the
  // two signals) compute; the forward waits for `ppg_ready` right before it first reads the PPG tensor
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_ppg = nullptr;
  cudaEvent_t ppg_ready = nullptr;  // set for the duration of one fsvc_forward_host call
  // fsvc_forward_host: the two signals are uploaded in up to kSigParts batch parts; the fused level-0 kernel is launched
  // per part, so a part computes while the next ones are still on the wire.  sig_ready[k] / sig_bounds / sig_parts are
  // set for the call (sig_parts == 0: no split; part k = utterances [sig_bounds[k], sig_bounds[k + 1])).
  static constexpr int kSigParts = 4;
  cudaEvent_t ev_sig[kSigParts] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t sig_ready[kSigParts] = {nullptr, nullptr, nullptr, nullptr};
  int sig_bounds[kSigParts + 1] = {0, 0, 0, 0, 0};
  int sig_parts = 0;
  // fsvc_forward_host: host destination of the waveform (set for the call); the tensor-core forward copies the first
  // half back itself, while the last conv of the second half runs, and reports how many floats it has taken care of
  float* out_host = nullptr;
  size_t out_host_done = 0;
  cudaEvent_t ev_out_half = nullptr, ev_out_done = nullptr;
  // tensor-core forward: small independent launches (speaker projections, PPG transpose) run on a forked stream
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_side_fork = nullptr, ev_side_join = nullptr;
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

extern const char* const stage_label[FSVC_MAX_STAGES];
extern const char* const lvl_label[FSVC_MAX_STAGES];
extern const char* const lvl_lft_label[FSVC_MAX_STAGES];
extern const char* const lvl_sine_label[FSVC_MAX_STAGES];

// ---- tc_forward.cu ----------------------------------------------------------------------------------
// Plan the tensor-core copies of every conv of `h` (sets tc2_ok / l0_fused and the offsets into tc_store, which the
// caller allocates with the returned element count and then fixes up with tc_fix_pointers).
size_t tc_plan_handle(fsvc_handle* h);
void tc_fix_pointers(fsvc_handle* h);
int tc_setup_kernels();
int tc_build_jobs(fsvc_handle* h);                     // device job table of the tensor-core packs (after tc_fix_pointers)
void tc_pack_weights(fsvc_handle* h, cudaStream_t s);  // after the fp32 packed store is filled
size_t tc_workspace_bytes(const fsvc_handle* h, int B, int frames);
int forward_tc2(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk, float* out,
                int B, int frames, void* workspace, size_t ws_bytes, cudaStream_t stream, Profiler* prof = nullptr);

}  // namespace fsvc
