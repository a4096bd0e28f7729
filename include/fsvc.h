/*
 * fsvc.h -- C ABI of the B200-native FastSVC generator forward pass.
 *
 * One shared library (libfsvc.so, built by nvcc for sm_100a only) holds every
 * CUDA kernel of the hot path; this header is everything a host binds against.
 * Plain C: pointers and sizes, no torch / C++ types.
 *
 * The reference (lesterphillip/SVCC23_FastSVC, pure Python/PyTorch) has no FFI
 * of its own; the boundary it offers is the Python class
 * `harana.models.FastSVCGenerator` (harana/models/fastsvc.py:235-383), looked
 * up by name at harana/bin/train_fastsvc.py:700-713 and
 * harana/utils/utils.py:266-275.  Each entry point below states which part of
 * that class it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success or a negative FSVC_E_* code and never
 *     throws, aborts or exits; fsvc_last_error() gives a thread-local message;
 *   - all tensor pointers are DEVICE pointers to contiguous fp32 data in the
 *     reference's layouts ((B, C, T), time fastest) unless the name says host;
 *   - buffers are borrowed for the duration of the call only; the library never
 *     allocates, frees, or synchronises inside fsvc_forward (it is CUDA-graph
 *     capturable); all work is enqueued on the caller's stream;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - a handle belongs to one (process, device) and is not re-entrant.
 */
#ifndef FSVC_H_
#define FSVC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSVC_ABI_VERSION 1
#define FSVC_MAX_STAGES 8

/* error codes */
#define FSVC_OK 0
#define FSVC_E_INVALID (-1)   /* bad argument / unsupported shape */
#define FSVC_E_CUDA (-2)      /* a CUDA runtime call failed */
#define FSVC_E_STATE (-3)     /* e.g. forward before set_weights */
#define FSVC_E_WORKSPACE (-4) /* workspace too small */
#define FSVC_E_NODEVICE (-5)  /* no sm_100 device: there is NO CPU fallback */

/* precision modes of fsvc_forward */
#define FSVC_MODE_FP32 0    /* fp32 FFMA everywhere: parity mode */
#define FSVC_MODE_TC_BF16X3 1 /* tcgen05 tensor cores, 3-term bf16 split, fp32 accumulate (channel counts that are
                                 not multiples of 8 fall back to fp32 FFMA) */
#define FSVC_MODE_AUTO 2    /* fastest mode that meets the 1e-3 parity bar for this config */

typedef struct fsvc_handle fsvc_handle;

/* Constructor arguments of FastSVCGenerator (fastsvc.py:238-246). */
typedef struct fsvc_config {
  int32_t in_channels;                        /* 144 */
  int32_t num_stages;                         /* len(mid_channels) == len(upsampling_scales) */
  int32_t mid_channels[FSVC_MAX_STAGES];      /* [192, 96, 48, 24] */
  int32_t upsampling_scales[FSVC_MAX_STAGES]; /* [2, 4, 4, 5] */
  int32_t out_channels;                       /* 1 */
  int32_t spk_emb_size;                       /* 512 */
  int32_t use_spk_emb;                        /* 1: emb_projector weights exist (fastsvc.py:77-78) */
  float lrelu_slope;                          /* 0.2  (fastsvc.py:58) */
  float in_eps;                               /* 1e-5 (nn.InstanceNorm2d default, fastsvc.py:76) */
} fsvc_config;

int fsvc_abi_version(void);
const char* fsvc_last_error(void);

/* Replaces FastSVCGenerator.__init__ (fastsvc.py:238-303): validates the
 * configuration and allocates the library-owned packed weight store on the
 * current device.  Fails with FSVC_E_NODEVICE when no CUDA device is present. */
int fsvc_create(const fsvc_config* cfg, fsvc_handle** out);
void fsvc_destroy(fsvc_handle* h);

/* The weight tensors the library expects, in canonical order.  Names are the
 * reference's state_dict names after remove_weight_norm() (fastsvc.py:342-352;
 * SURVEY.md 3.4), e.g. "upsampling_nets.0.conv_first.weight"; with weight norm
 * applied the host passes the effective weight g*v/||v|| under the same name.
 * Layout of each tensor is PyTorch's ((Cout, Cin, k) / (Cout, Cin, 1, k) /
 * (Cout, S) / (Cout,)). */
int fsvc_num_weight_tensors(const fsvc_handle* h);
int fsvc_weight_tensor_info(const fsvc_handle* h, int index, char* name, int name_capacity, int64_t* numel);

/* Replaces load_state_dict()/the weight-norm pre-hooks: copies and repacks the
 * n = fsvc_num_weight_tensors() effective fp32 weights (device pointers, in
 * canonical order) into the library's own layouts.  Call again whenever a
 * parameter changes (optimizer step, load_state_dict, remove_weight_norm). */
int fsvc_set_weights(fsvc_handle* h, const float* const* dev_ptrs, int n, void* stream);

/* Bytes of scratch fsvc_forward needs for a batch of B utterances of
 * `frames` PPG frames each (T = frames * prod(upsampling_scales) samples). */
size_t fsvc_workspace_bytes(const fsvc_handle* h, int B, int frames, int mode);

/* Replaces FastSVCGenerator.forward(x, s, l, spk_emb) (fastsvc.py:305-332):
 *   ppg  (B, in_channels, frames)     linguistic features x
 *   sine (B, 1, T)                    sine excitation s
 *   lft  (B, 1, T)                    loudness l
 *   spk  (B, spk_emb_size) or NULL    NULL => FiLM affine only, no InstanceNorm /
 *                                     speaker add (fastsvc.py:134)
 *   out  (B, out_channels, T)
 * Every utterance of the batch has the same length (InstanceNorm statistics
 * span the whole utterance, so padding is not result-neutral). */
int fsvc_forward(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                 float* out, int B, int frames, void* workspace, size_t workspace_bytes, int mode, void* stream);

/* Same call with HOST buffers (pinned for true asynchrony): copies the inputs
 * host->device into the workspace, runs fsvc_forward, copies the waveform
 * device->host, all on `stream`; the caller synchronises the stream.  The
 * workspace must be fsvc_workspace_bytes() + fsvc_host_io_bytes() large. */
size_t fsvc_host_io_bytes(const fsvc_handle* h, int B, int frames);
int fsvc_forward_host(fsvc_handle* h, const float* ppg_host, const float* sine_host, const float* lft_host,
                      const float* spk_host, float* out_host, int B, int frames, void* workspace,
                      size_t workspace_bytes, int mode, void* stream);

/* Block-level entry points (the reference's sub-modules stay importable and
 * are used standalone, e.g. tacotron2.py:22,458-459 uses FastSVCFiLMNet).
 * Weights are passed per call as device pointers in PyTorch layout.  The blocks
 * compute in fp32 (FFMA) in every `mode`: the tensor-core kernels work on the
 * generator's channels-last workspace, which a standalone block does not have.
 *
 * fsvc_downsample_forward replaces FastSVCDownsampleNet.forward
 * (fastsvc.py:180-193): w[] = {residual_block.0 (w,b), downsample_block.2
 * (w,b), .4 (w,b), .6 (w,b)}; x (B,Cin,T) -> out (B,C,T/scale); T % scale == 0. */
int fsvc_downsample_forward(const float* x, float* out, const float* const* w, int B, int c_in, int c, int T,
                            int scale, float slope, void* workspace, size_t workspace_bytes, int mode, void* stream);
/* fsvc_film_forward replaces FastSVCFiLMNet.forward (fastsvc.py:220-232):
 * w[] = {conv (w,b), conv_scale (w,b), conv_shift (w,b)}; x (B,C,T) ->
 * scale (B,C,T), shift (B,C,T). */
int fsvc_film_forward(const float* x, float* scale, float* shift, const float* const* w, int B, int c, int T,
                      float slope, void* workspace, size_t workspace_bytes, int mode, void* stream);
/* fsvc_upsample_forward replaces FastSVCUpsampleNet.forward
 * (fastsvc.py:80-140): w[] = {conv_first, upsample_block0.2, conv_block1.1,
 * conv_block2.1, conv_block3.1, residual_block.1, emb_projector} (w,b each; the
 * last pair may be NULL when spk is NULL); x (B,Cin,T) -> out (B,C,T*scale);
 * FiLM inputs s_scale, s_shift, l_scale, l_shift are (B,C,T*scale). */
int fsvc_upsample_forward(const float* x, const float* s_scale, const float* s_shift, const float* l_scale,
                          const float* l_shift, const float* spk, float* out, const float* const* w, int B,
                          int c_in, int c, int T, int scale, int spk_emb_size, float slope, float eps,
                          void* workspace, size_t workspace_bytes, int mode, void* stream);
size_t fsvc_block_workspace_bytes(int B, int c_in, int c, int T_out);

/* Profiling variant of fsvc_forward (NOT graph-capturable: it creates CUDA
 * events and synchronises the stream): brackets every kernel launch with
 * events on `stream` and reports, per launch, a label ("s3.d27_skip",
 * "l0.sine.down_d2", ...), the measured device time and the launch's
 * ALGORITHMIC work (flops = 2*Cin*Cout*K*T*B; bytes = every operand
 * tensor touched exactly once).  Used by bench.py for the roofline object. */
typedef struct fsvc_kernel_record {
  char label[48];
  float ms;
  double flops;
  double bytes;
} fsvc_kernel_record;
int fsvc_forward_profile(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                         float* out, int B, int frames, void* workspace, size_t workspace_bytes, int mode,
                         void* stream, fsvc_kernel_record* records, int capacity, int* count);

/* Number of kernels the last fsvc_forward on this handle enqueued. */
/*
 * Sine excitation, batched: replaces SignalGenerator.sinusoid (harana/utils/features.py:178-197), the step that
 * feeds `sine` of fsvc_forward; called once per utterance at batch 1 by FastSVCGenerator.inference
 * (fastsvc.py:381) and decode_fastsvc.py:182-186.
 *   f0    [B][frames] Hz, 0 = unvoiced;  noise [B][frames*hop] standard normal draws (the reference makes them with
 *   torch.randn inside the call, features.py:194) or NULL when noise_amp == 0;  out [B][frames*hop].
 * The phase is the reference's CPU cumsum (double accumulation, fp32 outputs); results match it to <= 2e-6.
 */
int fsvc_sine_excitation(const float* f0, const float* noise, float* out, int B, int frames, int hop,
                         float sample_rate, float sine_amp, float noise_amp, void* stream);

/*
 * Waveform -> PCM-16 as soundfile.write(..., "PCM_16") stores it (decode_fastsvc.py:193-198): libsndfile's clipping
 * float->short conversion (soundfile always enables SFC_SET_CLIPPING): top 16 bits of lrintf(x * 2^31), saturating,
 * i.e. floor(x * 32768) clipped to [-32768, 32767]. x, y: device pointers to n samples.
 */
int fsvc_pcm16(const float* x, int16_t* y, long long n, void* stream);

int fsvc_last_launch_count(const fsvc_handle* h);

/*
 * Training: replaces autograd through FastSVCGenerator.forward -- `y_ = self.model["generator"](*x)`
 * followed by `gen_loss.backward()` (harana/bin/train_fastsvc.py:168, 199-206).
 *
 * fsvc_forward_train is fsvc_forward in fp32 that also keeps, in the caller-owned `saved` buffer
 * (fsvc_train_saved_bytes), every activation the backward needs.  fsvc_backward takes dL/d(out) and writes the
 * gradient of every EFFECTIVE weight tensor -- n = fsvc_num_weight_tensors() device pointers in the canonical order of
 * fsvc_weight_tensor_info, PyTorch layouts, overwritten -- computed by hand-written kernels (data gradients, weight
 * gradients, InstanceNorm / FiLM / LeakyReLU / repeat / decimation adjoints; no autograd, no cuDNN).  With weight norm
 * applied the host chains them through w = g * v / ||v||.  Inputs receive no gradient (the reference never asks for
 * one).  The weights in the library must be the ones of the forward call; `workspace` (fsvc_train_workspace_bytes) is
 * scratch for either call.  Same stream / capture rules as fsvc_forward.
 */
size_t fsvc_train_saved_bytes(const fsvc_handle* h, int B, int frames);
size_t fsvc_train_workspace_bytes(const fsvc_handle* h, int B, int frames);
int fsvc_forward_train(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                       float* out, int B, int frames, void* saved, size_t saved_bytes, void* workspace,
                       size_t workspace_bytes, void* stream);
int fsvc_backward(fsvc_handle* h, const float* ppg, const float* sine, const float* lft, const float* spk,
                  const float* grad_out, int B, int frames, const void* saved, size_t saved_bytes,
                  float* const* grad_ptrs, int n, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSVC_H_ */
