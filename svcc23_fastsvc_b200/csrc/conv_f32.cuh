// fp32 (FFMA) direct 1-D convolution with every FastSVC prologue/epilogue fused.
//
// One kernel covers every layer of the generator in FSVC_MODE_FP32 and is the
// generic path for channel counts the tensor-core kernels do not take:
//
//   out[b,co,t] = bias[co] + sum_{ci,k} W[co,ci,k] * in'[b,ci,t+(k-(K-1)/2)*dil]
//   in'[b,ci,u] = 0                                   if u outside [0,T_out)   (zero padding is applied
//               = lrelu?( a[b,ci]*in[b,ci,map(u)] + c[b,ci] )  otherwise        AFTER norm/activation)
//   map(u)      = (u / up) * down        nearest repeat (Stretch2d) or decimation (Squeeze2d, T % s == 0)
//   epilogue    : v *= lrelu'(mask)?;  v += res;  raw = v;  v = lrelu?(v);  v = gamma*v + beta;  out = v;
//                 per-(b,co,tile) (mean, M2) partials of the stored value for InstanceNorm.
//
// Reference semantics: Conv1d1x3/Conv2d1x3/Conv1d1x1 (layers/upsample.py:76-106,
// layers/residual_block.py:41-48), Stretch2d/Squeeze2d (layers/upsample.py:38-74),
// _feature_affine (fastsvc.py:115-140).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fsvc_internal.h"
#include "packed_f32.cuh"

namespace fsvc {

struct ConvArgs {
  // input
  const float* in;
  long long in_bs;  // batch stride (elements)
  int in_cs;        // channel stride (elements) == stored length of the input
  int C_in;
  int up, down;        // index map
  const float* pre_a;  // [B][C_in] or nullptr
  const float* pre_c;  // [B][C_in] or nullptr
  int pre_lrelu;
  // weights, packed [C_in][K][C_out]
  const float* w;
  const float* bias;  // [C_out]
  int dil;
  int C_out, T_out;
  // epilogue
  const float* res;
  long long res_bs;
  int res_cs;
  float* raw;
  long long raw_bs;
  int raw_cs;
  int post_lrelu;
  const float* gamma;
  const float* beta;
  long long gb_bs;
  int gb_cs;
  float* out;
  long long out_bs;
  int out_cs;
  float2* stats;  // [B][C_out][n_tiles] (mean, M2) or nullptr
  int n_tiles;
  // backward use (train.cu): multiply the conv result by lrelu'(m) BEFORE `res` is added, with
  // m = mask[b][co][(t / mask_up) * mask_down] (optionally mask_a[b][co] * m + mask_c[b][co]): the derivative of the
  // LeakyReLU that preceded the forward conv whose data gradient this launch computes
  const float* mask;
  long long mask_bs;
  int mask_cs, mask_up, mask_down;
  const float* mask_a;
  const float* mask_c;
  float slope;
  const void* host_w;  // host-side ConvW* (launch dispatch only; never dereferenced on the device)
};

constexpr int kConvThreads = 256;
constexpr int kConvWarps = kConvThreads / 32;
constexpr int kCiTile = 16;

__device__ __forceinline__ float lrelu(float v, float slope) { return v >= 0.f ? v : v * slope; }

// RC output channels per warp (CTA: 8*RC), RT time steps per lane (CTA: 32*RT), K taps.
template <int RC, int RT, int K>
__global__ void __launch_bounds__(kConvThreads) conv1d_f32_kernel(const ConvArgs a) {
  constexpr int CO_T = kConvWarps * RC;
  constexpr int T_T = 32 * RT;
  extern __shared__ float smem[];
  const int halo = (K / 2) * a.dil;
  const int W_in = T_T + 2 * halo;
  float* in_s = smem;                   // [kCiTile][W_in]
  float* w_s = smem + kCiTile * W_in;   // [kCiTile][K][CO_T]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int t0 = tile * T_T;
  const int co0 = blockIdx.y * CO_T;
  const int b = blockIdx.z;

  float acc[RC][RT];
#pragma unroll
  for (int i = 0; i < RC; ++i)
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[i][j] = 0.f;

  const float* in_b = a.in + (long long)b * a.in_bs;
  for (int ci0 = 0; ci0 < a.C_in; ci0 += kCiTile) {
    __syncthreads();
    // stage inputs with the prologue applied once per element: a warp per channel row, lanes along time (coalesced,
    // no index division; the affine of the row is loaded once)
    for (int ci = warp; ci < kCiTile; ci += kConvWarps) {
      const int c = ci0 + ci;
      const bool row_ok = c < a.C_in;
      const float* row = in_b + (long long)(row_ok ? c : 0) * a.in_cs;
      float pa = 1.f, pc = 0.f;
      if (a.pre_a && row_ok) {
        pa = __ldg(a.pre_a + b * a.C_in + c);
        pc = __ldg(a.pre_c + b * a.C_in + c);
      }
      float* dst = in_s + ci * W_in;
      for (int p = lane; p < W_in; p += 32) {
        const int u = t0 - halo + p;
        float v = 0.f;
        if (row_ok && u >= 0 && u < a.T_out) {
          v = __ldg(row + (a.up == 1 ? u : u / a.up) * a.down);
          if (a.pre_a) v = fmaf(v, pa, pc);
          if (a.pre_lrelu) v = lrelu(v, a.slope);
        }
        dst[p] = v;
      }
    }
    for (int idx = tid; idx < kCiTile * K * CO_T; idx += kConvThreads) {
      const int ci = idx / (K * CO_T), rem = idx - ci * (K * CO_T);
      const int k = rem / CO_T, c = rem - k * CO_T;
      const int cig = ci0 + ci, cog = co0 + c;
      w_s[idx] = (cig < a.C_in && cog < a.C_out) ? __ldg(a.w + ((long long)cig * K + k) * a.C_out + cog) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int ci = 0; ci < kCiTile; ++ci) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float wv[RC], xv[RT];
        const float* wp = w_s + (ci * K + k) * CO_T + warp * RC;
#pragma unroll
        for (int i = 0; i < RC; ++i) wv[i] = wp[i];
        const float* xp = in_s + ci * W_in + k * a.dil + lane;
#pragma unroll
        for (int j = 0; j < RT; ++j) xv[j] = xp[32 * j];
        // two accumulators per packed FMA (same products, same order: bit-identical to the scalar loop): channel pairs
        // when RC is even, time pairs otherwise
        if constexpr (RC % 2 == 0) {
#pragma unroll
          for (int i = 0; i < RC; i += 2)
#pragma unroll
            for (int j = 0; j < RT; ++j) fma2_acc(acc[i][j], acc[i + 1][j], wv[i], wv[i + 1], xv[j], xv[j]);
        } else if constexpr (RT % 2 == 0) {
#pragma unroll
          for (int i = 0; i < RC; ++i)
#pragma unroll
            for (int j = 0; j < RT; j += 2) fma2_acc(acc[i][j], acc[i][j + 1], wv[i], wv[i], xv[j], xv[j + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < RC; ++i)
#pragma unroll
            for (int j = 0; j < RT; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
      }
    }
  }

  // epilogue
  const int n_valid = min(T_T, a.T_out - t0);
#pragma unroll
  for (int i = 0; i < RC; ++i) {
    const int co = co0 + warp * RC + i;
    if (co >= a.C_out) break;  // warp-uniform
    const float bias = a.bias ? __ldg(a.bias + co) : 0.f;
    // row pointers and per-(b, co) constants once per output channel: the step loop only adds t
    const float* mask_row = a.mask ? a.mask + (long long)b * a.mask_bs + (long long)co * a.mask_cs : nullptr;
    const float ma = (a.mask && a.mask_a) ? __ldg(a.mask_a + b * a.C_out + co) : 1.f;
    const float mc = (a.mask && a.mask_a) ? __ldg(a.mask_c + b * a.C_out + co) : 0.f;
    const float* res_row = a.res ? a.res + (long long)b * a.res_bs + (long long)co * a.res_cs : nullptr;
    float* raw_row = a.raw ? a.raw + (long long)b * a.raw_bs + (long long)co * a.raw_cs : nullptr;
    const long long gb_off = (long long)b * a.gb_bs + (long long)co * a.gb_cs;
    float* out_row = a.out ? a.out + (long long)b * a.out_bs + (long long)co * a.out_cs : nullptr;
    float vals[RT];
    float s1 = 0.f;
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      const int t = t0 + lane + 32 * j;
      float v = acc[i][j] + bias;
      if (t < a.T_out) {
        if (mask_row) {
          float m = __ldg(mask_row + (a.mask_up == 1 ? t : t / a.mask_up) * a.mask_down);
          if (a.mask_a) m = fmaf(m, ma, mc);
          v *= m > 0.f ? 1.f : a.slope;   // torch leaky_relu backward: x > 0 ? g : g * slope
        }
        if (res_row) v += __ldg(res_row + t);
        if (raw_row) raw_row[t] = v;
        if (a.post_lrelu) v = lrelu(v, a.slope);
        if (a.gamma) v = fmaf(__ldg(a.gamma + gb_off + t), v, __ldg(a.beta + gb_off + t));
        if (out_row) out_row[t] = v;
        s1 += v;
      } else {
        v = 0.f;
      }
      vals[j] = v;
    }
    if (a.stats) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      const float mean = s1 / (float)n_valid;
      float m2 = 0.f;
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int t = t0 + lane + 32 * j;
        const float d = vals[j] - mean;
        if (t < a.T_out) m2 = fmaf(d, d, m2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      if (lane == 0) a.stats[((long long)b * a.C_out + co) * a.n_tiles + tile] = make_float2(mean, m2);
    }
  }
}

// Merge the per-tile (mean, M2) partials of one (b, c) row in a fixed order
// (Chan et al.), in double, and emit the per-channel affine that the next conv
// applies on load:  norm(x) + e  ==  x * a + c  with a = rstd, c = e - mean*rstd.
// InstanceNorm2d semantics: biased variance over the whole time axis, eps inside
// the sqrt (fastsvc.py:76,138).  e = emb_projector(normalize(spk)) (fastsvc.py:135-137).
static __global__ void in_finalize_kernel(const float2* __restrict__ stats, int n_tiles, int tile_len, int T, int BC,
                                   const float* __restrict__ e, float eps, float* __restrict__ out_a,
                                   float* __restrict__ out_c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC) return;
  const float2* p = stats + (long long)i * n_tiles;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int t = 0; t < n_tiles; ++t) {
    const float2 s = p[t];
    const double nb = (double)min(tile_len, T - t * tile_len);
    const double d = (double)s.x - mean;
    const double nn = n + nb;
    mean += d * nb / nn;
    m2 += (double)s.y + d * d * n * nb / nn;
    n = nn;
  }
  const double var = m2 / (double)T;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  out_a[i] = (float)rstd;
  out_c[i] = (float)((double)(e ? e[i] : 0.f) - mean * rstd);
}

// e[b][c] = bias[c] + sum_j W[c][j] * spk[b][j] / max(||spk[b]||_2, 1e-12)
// (nn.Linear(F.normalize(spk_emb)), fastsvc.py:135-137).  grid = (B), block = 256.
static __global__ void spk_project_kernel(const float* __restrict__ spk, int S, const float* __restrict__ W,
                                   const float* __restrict__ bias, int C, float* __restrict__ e) {
  __shared__ float red[32];
  __shared__ float inv_norm;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = spk + (long long)b * S;
  float ss = 0.f;
  for (int j = tid; j < S; j += blockDim.x) ss = fmaf(x[j], x[j], ss);
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) inv_norm = 1.f / fmaxf(sqrtf(v), 1e-12f);
  }
  __syncthreads();
  const float inv = inv_norm;
  for (int c = warp; c < C; c += (blockDim.x >> 5)) {
    const float* w = W + (long long)c * S;
    float acc = 0.f;
    for (int j = lane; j < S; j += 32) acc = fmaf(w[j], x[j] * inv, acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) e[(long long)b * C + c] = acc + bias[c];
  }
}

// PyTorch (Cout, Cin, K) -> packed [ci_off + ci][k][co_off + co] with row length dst_cout (block-level entry points,
// which get their weights per call; the generator's weights go through weight_jobs_kernel).
static __global__ void repack_weight_kernel(const float* __restrict__ src, int C_out, int C_in, int K,
                                            float* __restrict__ dst, int dst_cout, int ci_off, int co_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C_out * C_in * K) return;
  const int k = i % K, ci = (i / K) % C_in, co = i / (K * C_in);
  dst[((long long)(ci_off + ci) * K + k) * dst_cout + co_off + co] = src[i];
}

// dst[off + i] = a[i] (+ b[i])
static __global__ void bias_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, int n,
                                       float* __restrict__ dst, int off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[off + i] = a[i] + (b ? b[i] : 0.f);
}

// Weight preparation, batched: job blockIdx.y of the table, elements strided over blockIdx.x (fsvc_internal.h: WJob).
//   0 repack     PyTorch (Cout, Cin, K) -> packed [ci_off + ci][k][co_off + co] with row length dst_ld
//   1 bias       dst[co_off + i] = a[i] (+ b[i])
//   2 copy       dst[i] = a[i]
//   3 transpose  packed [C_in][K][C_out] -> packed [C_out][K][C_in] with the taps reversed: the conv whose output is
//                the gradient of the original conv's input:  g_in[u] = sum_{k,co} W[co,ci,k] * g_out[u - (k-(K-1)/2)*dil]
static __global__ void __launch_bounds__(256) weight_jobs_kernel(const WJob* __restrict__ jobs, const WSrc src) {
  const WJob j = jobs[blockIdx.y];
  float* dst = reinterpret_cast<float*>(j.dst);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.total; i += (long long)gridDim.x * blockDim.x) {
    if (j.kind == 0) {
      const int k = (int)(i % j.K), ci = (int)((i / j.K) % j.C_in), co = (int)(i / ((long long)j.K * j.C_in));
      dst[((long long)(j.ci_off + ci) * j.K + k) * j.dst_ld + j.co_off + co] = src.p[j.src_a][i];
    } else if (j.kind == 1) {
      dst[j.co_off + i] = src.p[j.src_a][i] + (j.src_b >= 0 ? src.p[j.src_b][i] : 0.f);
    } else if (j.kind == 2) {
      dst[i] = src.p[j.src_a][i];
    } else {
      const int co = (int)(i % j.C_out), k = (int)((i / j.C_out) % j.K), ci = (int)(i / ((long long)j.C_out * j.K));
      dst[((long long)co * j.K + (j.K - 1 - k)) * j.C_in + ci] = j.w[i];
    }
  }
}

}  // namespace fsvc
