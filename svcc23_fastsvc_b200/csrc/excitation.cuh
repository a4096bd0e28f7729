// Sine excitation (the step immediately before the generator) and PCM-16 quantisation (the step immediately
// after it in offline conversion), batched over utterances.
//
// Reference: SignalGenerator.sinusoid, harana/utils/features.py:178-197 --
//     vuv    = nearest_upsample(f0 > 0)
//     rad    = (nearest_upsample(f0) / sample_rate) % 1
//     sine   = vuv * sin(cumsum(rad) * 2 * pi) * sine_amp  [+ randn * (vuv*noise_amp + (1-vuv)*noise_amp/3)]
// The reference's CPU cumsum accumulates the fp32 terms in double and rounds each output to fp32
// (oracle/features_numpy.py); so does this kernel: frame-level prefix sums in double (a term repeated `hop` times is
// one multiply), sample-level value prefix + (i+1)*rad in double, rounded once.  Every later operation is the
// reference's fp32 operation in the reference's order (no FMA contraction), so results agree to the last bits of sinf.
// The Gaussian noise is an INPUT (the reference draws it with torch.randn inside the call; the host wrapper makes the
// same draw) -- identical bits on both sides.
//
// HBM-bound: per sample one fp32 noise read + one fp32 write (8 B), 4 B per frame of f0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fsvc {

constexpr int kExcThreads = 256;
constexpr int kExcFrames = 32;  // frames per CTA (one warp scans them)

struct ExcArgs {
  const float* f0;     // [B][frames]
  const float* noise;  // [B][frames*hop] or nullptr
  float* out;          // [B][frames*hop]
  int B, frames, hop;
  float sample_rate, sine_amp, noise_amp;
};

__global__ void __launch_bounds__(kExcThreads) sine_excitation_kernel(const ExcArgs p) {
  __shared__ double s_red[kExcThreads / 32];
  __shared__ double s_base;
  __shared__ double s_pref[kExcFrames];  // cumulative sum BEFORE frame f (exclusive), including s_base
  __shared__ float s_rad[kExcFrames];
  __shared__ float s_vuv[kExcFrames];
  const int b = blockIdx.y, fb = blockIdx.x * kExcFrames;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* f0 = p.f0 + (long long)b * p.frames;
  auto rad_of = [&](float f) {
    // (f / sample_rate) % 1 with torch.remainder semantics (sign of the divisor)
    float r = __fdiv_rn(f, p.sample_rate);
    float m = fmodf(r, 1.0f);
    if (m < 0.f) m += 1.0f;
    return m;
  };
  // 1. sum of hop*rad over all frames before this CTA's block (double; order is irrelevant at 2^-53)
  double part = 0.0;
  for (int f = tid; f < fb; f += kExcThreads) part += (double)p.hop * (double)rad_of(__ldg(f0 + f));
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __double2loint(part), hi = __double2hiint(part);
    lo = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = __shfl_xor_sync(0xffffffffu, hi, o);
    part += __hiloint2double(hi, lo);
  }
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < kExcThreads / 32; ++i) s += s_red[i];
    s_base = s;
  }
  __syncthreads();
  // 2. exclusive scan over this block's frames (warp 0)
  if (warp == 0) {
    const int f = fb + lane;
    const float fv = f < p.frames ? __ldg(f0 + f) : 0.f;
    const float r = rad_of(fv);
    double inc = (double)p.hop * (double)r;
    double incl = inc;
    for (int o = 1; o < 32; o <<= 1) {
      int lo = __double2loint(incl), hi = __double2hiint(incl);
      lo = __shfl_up_sync(0xffffffffu, lo, o);
      hi = __shfl_up_sync(0xffffffffu, hi, o);
      const double v = __hiloint2double(hi, lo);
      if (lane >= o) incl += v;
    }
    s_pref[lane] = s_base + (incl - inc);
    s_rad[lane] = r;
    s_vuv[lane] = fv > 0.f ? 1.f : 0.f;
  }
  __syncthreads();
  // 3. samples of the block: lanes <-> consecutive samples (coalesced)
  const int n_fr = min(kExcFrames, p.frames - fb);
  const long long T = (long long)p.frames * p.hop;
  const long long t_base = (long long)fb * p.hop;
  const float namp_v = p.noise_amp;                                   // vuv * noise_amp (vuv = 1)
  const float namp_u = __fdiv_rn(p.noise_amp, 3.0f);                  // ((1 - vuv) * noise_amp) / 3 (vuv = 0)
  const float two_pi_hi = 3.14159274101257324f;                       // float32(np.pi)
  for (int i = tid; i < n_fr * p.hop; i += kExcThreads) {
    const int fl = i / p.hop, k = i - fl * p.hop;
    const float vuv = s_vuv[fl];
    const float cum = (float)(s_pref[fl] + (double)(k + 1) * (double)s_rad[fl]);
    const float ph = __fmul_rn(__fmul_rn(cum, 2.0f), two_pi_hi);
    float v = __fmul_rn(__fmul_rn(vuv, sinf(ph)), p.sine_amp);
    if (p.noise) {
      const float na = vuv > 0.f ? namp_v : namp_u;
      v = __fadd_rn(v, __fmul_rn(__ldg(p.noise + (long long)b * T + t_base + i), na));
    }
    p.out[(long long)b * T + t_base + i] = v;
  }
}

// float waveform -> PCM-16 as soundfile.write(..., "PCM_16") stores it (decode_fastsvc.py:193-198).  python-soundfile
// opens every file with SFC_SET_CLIPPING on, so libsndfile (third-party, not in /root/reference; 1.0.x/1.2.x src/pcm.c)
// converts with f2les_clip_array: scaled = x * 2^31 (float), saturate at >= 0x7FFFFFFF / <= -2^31, lrintf, keep the
// top 16 bits -- i.e. floor(x * 32768) with clipping, NOT rint(x * 32767).  Restated here from the published
// algorithm; neither libsndfile nor soundfile is installed in this image, so this step is "parity unpinned".
__global__ void __launch_bounds__(256) pcm16_kernel(const float* __restrict__ x, int16_t* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = i; j < n; j += stride) {
    const float v = __fmul_rn(x[j], 2147483648.0f);
    int q;
    if (v >= 2147483648.0f) q = 0x7FFF;           // (float)0x7FFFFFFF == 2^31
    else if (v <= -2147483648.0f) q = -0x8000;
    else q = __float2int_rn(v) >> 16;             // arithmetic shift: the top 16 bits of the 32-bit sample
    y[j] = (int16_t)q;
  }
}

}  // namespace fsvc
