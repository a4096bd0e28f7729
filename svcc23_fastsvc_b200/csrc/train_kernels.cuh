// Kernels of the native generator backward (callers train_fastsvc.py:168,199-206).
//
// Everything is fp32 in the reference's (B, C, T) layout.  The data gradient of a conv is the same conv with the
// transposed, tap-reversed weights, so it runs on conv1d_f32_kernel (conv_f32.cuh, `mask` epilogue = derivative of the
// LeakyReLU in front of the forward conv).  This file holds what has no forward counterpart:
//
//   conv_wgrad_kernel    dW[co][ci][k] = sum_{b,t} g[b,co,t] * A[b,ci,t+(k-1)d],  db[co] = sum g, A = the forward conv's
//                        operand rebuilt from the saved tensor (index map, InstanceNorm affine, LeakyReLU); split over
//                        time into per-CTA partials
//   wgrad_reduce_kernel  fixed-order sum of the partials into PyTorch (C_out, C_in, K) tensors (sub-blocks of a merged
//                        conv go to different tensors), many jobs per launch
//   film_in_bwd_kernel   backward of lrelu(InstanceNorm(gamma*v + beta) + e): two fixed-order reductions per (b, c)
//                        row, then g_v, g_gamma, g_beta, g_e           (fastsvc.py:115-140)
//   fold_repeat_kernel   adjoint of the nearest-neighbour repeat (Stretch2d, layers/upsample.py:38-50)
//   scatter_dec_kernel   adjoint of the decimation (Squeeze2d, layers/upsample.py:64-74)
//   spk_bwd_kernel       backward of emb_projector(F.normalize(spk)) w.r.t. its weight and bias (fastsvc.py:135-137)
#pragma once
#include "conv_f32.cuh"

namespace fsvc {

// ---------------------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------------------
struct WgradArgs {
  // forward operand A[b][ci][u] = lrelu?(pre_a*x[b][ci][(u/up)*down] + pre_c), zero outside [0, T)
  const float* x;
  long long x_bs;
  int x_cs, C_in, up, down;
  const float* pre_a;
  const float* pre_c;
  int pre_lrelu;
  // output gradient [B][C_out][T]
  const float* g;
  long long g_bs;
  int g_cs, C_out, T, B;
  int dil;
  float slope;
  float* part_w;  // [n_split = gridDim.x][C_in][K][C_out]
  float* part_b;  // [n_split][C_out] or nullptr
};

constexpr int kWgThreads = 256;
constexpr int kWgCo = 16;   // output channels per CTA (4 warp columns x 4)
constexpr int kWgCi = 8;    // input channels per CTA (2 warp rows x 4)
constexpr int kWgTT = 256;  // time steps staged per chunk

static inline size_t wgrad_smem_bytes(int K, int dil) {
  return (size_t)(kWgCo * kWgTT + kWgCi * (kWgTT + 2 * (K / 2) * dil)) * sizeof(float);
}

// grid = (n_split, ceil(C_out/16), ceil(C_in/8)).  Lanes run along time (conflict-free shared-memory rows), every warp
// owns a 4 co x 4 ci x K register tile for the CTA's whole time range and is reduced across lanes once at the end.
template <int K>
__global__ void __launch_bounds__(kWgThreads) conv_wgrad_kernel(const WgradArgs a) {
  extern __shared__ float smem[];
  const int halo = (K / 2) * a.dil;
  const int W = kWgTT + 2 * halo;
  float* g_s = smem;                  // [kWgCo][kWgTT]
  float* a_s = smem + kWgCo * kWgTT;  // [kWgCi][W]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int co0 = blockIdx.y * kWgCo, ci0 = blockIdx.z * kWgCi;
  const int wco = (warp & 3) * 4, wci = (warp >> 2) * 4;
  float acc[4][4][K];
  float bsum[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    bsum[r] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < K; ++k) acc[r][q][k] = 0.f;
  }
  const int n_tt = (a.T + kWgTT - 1) / kWgTT;
  const int n_chunks = a.B * n_tt;
  const int c_begin = (int)((long long)blockIdx.x * n_chunks / gridDim.x);
  const int c_end = (int)((long long)(blockIdx.x + 1) * n_chunks / gridDim.x);
  for (int ch = c_begin; ch < c_end; ++ch) {
    const int b = ch / n_tt, t0 = (ch - b * n_tt) * kWgTT;
    __syncthreads();
    // staging, channel-interleaved: g_s[group of 4 co][step][4], a_s[group of 4 ci][window row][4] -- the compute
    // loop then fetches a warp's 4 output / 4 input channels of a step with ONE 128-bit load each (4 instead of 16 loads
    // per 32 steps and 48 MACs).  A warp stages one group: 4 coalesced global loads per step, one conflict-free
    // 128-bit store.
    for (int job = warp; job < 8; job += kWgThreads / 32) {  // 4 co groups x 2 halves of the chunk
      const int grp = job & 3, half = job >> 2;
      const float* rows[4];
      bool ok_r[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int c = co0 + grp * 4 + r;
        ok_r[r] = c < a.C_out;
        rows[r] = a.g + (long long)b * a.g_bs + (long long)(ok_r[r] ? c : 0) * a.g_cs + t0;
      }
      const int n_ok = min(kWgTT, a.T - t0);
      for (int p = half * (kWgTT / 2) + lane; p < (half + 1) * (kWgTT / 2); p += 32) {
        float4 v;
        v.x = (ok_r[0] && p < n_ok) ? __ldg(rows[0] + p) : 0.f;
        v.y = (ok_r[1] && p < n_ok) ? __ldg(rows[1] + p) : 0.f;
        v.z = (ok_r[2] && p < n_ok) ? __ldg(rows[2] + p) : 0.f;
        v.w = (ok_r[3] && p < n_ok) ? __ldg(rows[3] + p) : 0.f;
        reinterpret_cast<float4*>(g_s)[grp * kWgTT + p] = v;
      }
    }
    for (int job = warp; job < 8; job += kWgThreads / 32) {  // 2 ci groups x 4 quarters of the window
      const int grp = job & 1, quarter = job >> 1;
      const float* rows[4];
      bool ok_r[4];
      float pa[4], pc[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int c = ci0 + grp * 4 + r;
        ok_r[r] = c < a.C_in;
        rows[r] = a.x + (long long)b * a.x_bs + (long long)(ok_r[r] ? c : 0) * a.x_cs;
        pa[r] = (a.pre_a && ok_r[r]) ? __ldg(a.pre_a + b * a.C_in + c) : 1.f;
        pc[r] = (a.pre_a && ok_r[r]) ? __ldg(a.pre_c + b * a.C_in + c) : 0.f;
      }
      const int q_len = (W + 3) / 4;
      for (int p = quarter * q_len + lane; p < min(W, (quarter + 1) * q_len); p += 32) {
        const int u = t0 - halo + p;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (u >= 0 && u < a.T) {
          const int src = (a.up == 1 ? u : u / a.up) * a.down;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (ok_r[r]) {
              float x = __ldg(rows[r] + src);
              if (a.pre_a) x = fmaf(x, pa[r], pc[r]);
              if (a.pre_lrelu) x = lrelu(x, a.slope);
              v[r] = x;
            }
          }
        }
        reinterpret_cast<float4*>(a_s)[grp * W + p] = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    __syncthreads();
    const float4* g4 = reinterpret_cast<const float4*>(g_s) + (warp & 3) * kWgTT;
    const float4* a4 = reinterpret_cast<const float4*>(a_s) + (warp >> 2) * W;
#pragma unroll 2
    for (int i = 0; i < kWgTT / 32; ++i) {
      const int p = lane + 32 * i;
      const float4 g = g4[p];
      const float gv[4] = {g.x, g.y, g.z, g.w};
      float av[4][K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4 x = a4[p + k * a.dil];
        av[0][k] = x.x; av[1][k] = x.y; av[2][k] = x.z; av[3][k] = x.w;
      }
#pragma unroll
      for (int r = 0; r < 4; r += 2) {  // output-channel pairs per packed FMA (bit-identical to the scalar loop)
        add2(bsum[r], bsum[r + 1], gv[r], gv[r + 1]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int k = 0; k < K; ++k) fma2_acc(acc[r][q][k], acc[r + 1][q][k], gv[r], gv[r + 1], av[q][k], av[q][k]);
      }
    }
  }
  float* pw = a.part_w + (size_t)blockIdx.x * a.C_in * K * a.C_out;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float v = acc[r][q][k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int co = co0 + wco + r, ci = ci0 + wci + q;
        if (lane == 0 && co < a.C_out && ci < a.C_in) pw[((size_t)ci * K + k) * a.C_out + co] = v;
      }
    if (a.part_b && blockIdx.z == 0 && wci == 0) {
      float v = bsum[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int co = co0 + wco + r;
      if (lane == 0 && co < a.C_out) a.part_b[(size_t)blockIdx.x * a.C_out + co] = v;
    }
  }
}

// dst[(co * nci + ci) * K + k] = sum_s part[s][(ci0 + ci) * K + k][co0 + co]   (bias: K = 1, rows_tot = 1, nci = 1)
struct ReduceJob {
  const float* part;
  float* dst;
  int n_split, rows_tot, row_len, K;
  int ci0, nci, co0, nco;
};
constexpr int kReduceJobs = 64;
struct ReduceBatch {
  ReduceJob j[kReduceJobs];
};

// grid = (blocks per job, jobs)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ ReduceBatch rb) {
  const ReduceJob& j = rb.j[blockIdx.y];
  const int total = j.nco * j.nci * j.K;
  const size_t split_stride = (size_t)j.rows_tot * j.row_len;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx % j.nco, rem = idx / j.nco;
    const int k = rem % j.K, ci = rem / j.K;
    const float* p = j.part + ((size_t)(j.ci0 + ci) * j.K + k) * j.row_len + j.co0 + co;
    double s = 0.0;
    for (int sp = 0; sp < j.n_split; ++sp) s += (double)p[sp * split_stride];
    j.dst[((size_t)co * j.nci + ci) * j.K + k] = (float)s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// FiLM + InstanceNorm + LeakyReLU backward
//
// forward (fastsvc.py:115-140, 56-75):  t = gamma*v' + beta,  v' = v or lrelu(v);   z = pa*t + pc  (= IN(t) + e);
//                                       A = lrelu(z)  -> conv
// given gA = dL/dA:  g_z = gA * lrelu'(z);  xh = z - e;  S1 = sum_t g_z;  S2 = sum_t g_z*xh;
//                    g_t = pa * (g_z - S1/T - xh*S2/T)      (no norm: g_t = g_z)
//                    g_v = g_t*gamma [* lrelu'(v)] [+ add];  g_gamma (+)= g_t*v';  g_beta (+)= g_t;  g_e (+)= S1
// ---------------------------------------------------------------------------------------------------------------
struct FilmBwdArgs {
  const float* gA;  // [B][C][T]
  const float* t;   // [B][C][T]
  const float* pa;  // [B][C] or nullptr (no InstanceNorm: spk_emb was None)
  const float* pc;
  const float* e;  // [B][C] or nullptr
  const float* gamma;
  long long gb_bs;  // batch stride of gamma / g_gamma / g_beta (2*C*T)
  const float* v;   // [B][C][T]
  int v_lrelu;
  const float* add;  // [B][C][T] or nullptr
  float* g_v;        // [B][C][T]
  float* g_gamma;
  float* g_beta;
  int gb_accum;
  float* g_e;  // [B][C] or nullptr
  int ge_accum;
  int C, T;
  float slope;
};

__device__ __forceinline__ double block_sum_double(double v, double* red) {  // 256 threads, fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = __shfl_xor_sync(0xffffffffu, hi, o);
    v += __hiloint2double(hi, lo);
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// grid = (C, B), block = 256
__global__ void __launch_bounds__(256) film_in_bwd_kernel(const FilmBwdArgs a) {
  __shared__ double red[8];
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const long long row = ((long long)b * a.C + c) * a.T;
  const long long gbrow = (long long)b * a.gb_bs + (long long)c * a.T;
  const float* gA = a.gA + row;
  const float* tt = a.t + row;
  const bool norm = a.pa != nullptr;
  const float pa = norm ? a.pa[b * a.C + c] : 1.f, pc = norm ? a.pc[b * a.C + c] : 0.f;
  const float e = (norm && a.e) ? a.e[b * a.C + c] : 0.f;
  float m1 = 0.f, m2 = 0.f;
  if (norm) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = tid; i < a.T; i += 256) {
      const float z = fmaf(pa, tt[i], pc);
      const float gz = gA[i] * (z > 0.f ? 1.f : a.slope);
      s1 += gz;
      s2 = fmaf(gz, z - e, s2);
    }
    const double S1 = block_sum_double((double)s1, red);
    const double S2 = block_sum_double((double)s2, red);
    m1 = (float)(S1 / (double)a.T);
    m2 = (float)(S2 / (double)a.T);
    if (a.g_e && tid == 0) {
      float* ge = a.g_e + b * a.C + c;
      *ge = (a.ge_accum ? *ge : 0.f) + (float)S1;
    }
  }
  const float* gam = a.gamma + gbrow;
  const float* vv = a.v + row;
  float* gg = a.g_gamma + gbrow;
  float* gb = a.g_beta + gbrow;
  for (int i = tid; i < a.T; i += 256) {
    const float z = fmaf(pa, tt[i], pc);
    const float gz = gA[i] * (z > 0.f ? 1.f : a.slope);
    const float gt = norm ? pa * (gz - m1 - (z - e) * m2) : gz;
    const float vr = vv[i];
    const float vp = a.v_lrelu ? lrelu(vr, a.slope) : vr;
    float gv = gt * gam[i];
    if (a.v_lrelu) gv *= vr > 0.f ? 1.f : a.slope;
    if (a.add) gv += a.add[row + i];
    a.g_v[row + i] = gv;
    if (a.gb_accum) {
      gg[i] += gt * vp;
      gb[i] += gt;
    } else {
      gg[i] = gt * vp;
      gb[i] = gt;
    }
  }
}

// g_src[b][c][j] = sum_{i<r} g_rep[b][c][j*r + i]
__global__ void __launch_bounds__(256) fold_repeat_kernel(const float* __restrict__ g_rep, float* __restrict__ g_src,
                                                          long long n_src, int r) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_src) return;
  const float* p = g_rep + i * r;
  float s = 0.f;
  for (int k = 0; k < r; ++k) s += p[k];
  g_src[i] = s;
}

// g_full[b][c][j*s] += g_dec[b][c][j]   (rows of T_full = T_dec * s)
__global__ void __launch_bounds__(256) scatter_dec_kernel(const float* __restrict__ g_dec, float* __restrict__ g_full,
                                                          long long n_dec, int s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dec) return;
  g_full[i * s] += g_dec[i];
}

// e = W n + bias, n = spk / max(||spk||, 1e-12):  dW[c][j] = sum_b ge[b][c] * n[b][j];  db[c] = sum_b ge[b][c].
// grid = (C), block = 256; dynamic smem = B floats (1 / norm per utterance)
__global__ void __launch_bounds__(256) spk_bwd_kernel(const float* __restrict__ spk, int S, int B,
                                                      const float* __restrict__ ge, int C, float* __restrict__ dW,
                                                      float* __restrict__ db) {
  extern __shared__ float inv[];
  __shared__ float red[8];
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int b = 0; b < B; ++b) {
    float ss = 0.f;
    for (int j = tid; j < S; j += 256) ss = fmaf(spk[(long long)b * S + j], spk[(long long)b * S + j], ss);
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    __syncthreads();
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (tid == 0) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v += red[w];
      inv[b] = 1.f / fmaxf(sqrtf(v), 1e-12f);
    }
  }
  __syncthreads();
  for (int j = tid; j < S; j += 256) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(ge[b * C + c], spk[(long long)b * S + j] * inv[b], s);
    dW[(long long)c * S + j] = s;
  }
  if (tid == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += ge[b * C + c];
    db[c] = s;
  }
}

}  // namespace fsvc
