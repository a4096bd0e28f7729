// Host-side launch helpers of conv1d_f32_kernel (conv_f32.cuh), shared by the fp32 forward (fsvc_abi.cu) and the
// training forward / backward (train.cu).
#pragma once
#include "conv_f32.cuh"
#include "fsvc_internal.h"

namespace fsvc {

// ---------------------------------------------------------------------------
// fp32 conv dispatch
// ---------------------------------------------------------------------------
template <int RC, int RT, int K>
static inline void launch_conv_t(const ConvArgs& a, int B, cudaStream_t s) {
  constexpr int CO_T = kConvWarps * RC, T_T = 32 * RT;
  const int halo = (K / 2) * a.dil;
  const size_t smem = (size_t)(kCiTile * (T_T + 2 * halo) + kCiTile * K * CO_T) * sizeof(float);
  dim3 grid((a.T_out + T_T - 1) / T_T, (a.C_out + CO_T - 1) / CO_T, B);
  conv1d_f32_kernel<RC, RT, K><<<grid, kConvThreads, smem, s>>>(a);
}

// Time steps per lane.  Long layers take 8 (256-step tiles); short ones shrink the tile until the launch has a few
// CTAs per SM -- at T <= 816 (the low-rate levels / stages of a training batch) 128-step tiles left 64 CTAs on 148
// SMs, one per SM, each running its 12 input-channel tiles back to back with nothing to hide the loads behind.
static inline int conv_rt(int T_out) { return T_out >= 1024 ? 8 : (T_out > 512 ? 4 : (T_out > 192 ? 2 : 1)); }
static inline int conv_tile_len(int T_out) { return 32 * conv_rt(T_out); }

static inline int conv_rc(int C_out) {
  if (C_out <= 8) return 1;
  if (C_out <= 16) return 2;
  if (C_out <= 24) return 3;
  if (C_out <= 32) return 4;
  if (C_out % 48 == 0) return 6;
  if (C_out % 32 == 0) return 4;
  return 6;
}

static inline void launch_conv(Ctx& c, const ConvArgs& a, int K, const char* name = "conv") {
  const int rc = conv_rc(a.C_out), rt = conv_rt(a.T_out);
#define FSVC_CASE(RC_, RT_)                                                  \
  if (rc == RC_ && rt == RT_) {                                              \
    if (K == 3) launch_conv_t<RC_, RT_, 3>(a, c.B, c.stream);                \
    else launch_conv_t<RC_, RT_, 1>(a, c.B, c.stream);                       \
  }
  FSVC_CASE(1, 1) FSVC_CASE(1, 2) FSVC_CASE(1, 4) FSVC_CASE(1, 8) FSVC_CASE(2, 1) FSVC_CASE(2, 2) FSVC_CASE(2, 4)
  FSVC_CASE(2, 8) FSVC_CASE(3, 1) FSVC_CASE(3, 2) FSVC_CASE(3, 4) FSVC_CASE(3, 8) FSVC_CASE(4, 1) FSVC_CASE(4, 2)
  FSVC_CASE(4, 4) FSVC_CASE(4, 8) FSVC_CASE(6, 1) FSVC_CASE(6, 2) FSVC_CASE(6, 4) FSVC_CASE(6, 8)
#undef FSVC_CASE
  // algorithmic work of this launch: 2*Cin*Cout*K*T flops; every operand tensor touched once
  const double BT = (double)c.B * a.T_out;
  const double flops = 2.0 * a.C_in * a.C_out * K * BT;
  double elems = (double)c.B * a.C_in * ((double)a.T_out / a.up);
  elems += BT * a.C_out * ((a.out ? 1 : 0) + (a.raw ? 1 : 0) + (a.res ? 1 : 0) + (a.gamma ? 2 : 0));
  c.launched(name, flops, 4.0 * (elems + (double)a.C_in * a.C_out * K));
}

// Convenience builder: dense (B, C, T) tensors.
static inline ConvArgs conv_args(const Ctx& c, const ConvW& w, const float* in, int T_in_stored, int T_out, int dil,
                          float* out) {
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in;
  a.in_cs = T_in_stored;
  a.in_bs = (long long)w.C_in * T_in_stored;
  a.C_in = w.C_in;
  a.up = 1;
  a.down = 1;
  a.mask_up = 1;
  a.mask_down = 1;
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

}  // namespace fsvc
