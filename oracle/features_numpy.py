"""CPU oracle (numpy) of the excitation path that feeds the generator.  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and the CPU legs of the benches -- never by the product path.

Restates, citing the reference lines it follows:
  * ``SignalGenerator.sinusoid / vuv_binary / random_noise``  harana/utils/features.py:163-216
  * ``F0Statistics.convert``                                  harana/utils/features.py:79-108
Pinned by tests/golden/sine_*.npz, generated from the reference itself (tests/golden/make_golden_features.py).

Arithmetic notes (verified against the reference on torch 2.11 CPU):
  * ``interpolate(x, T*hop)`` (nearest, integer ratio) is ``repeat(hop)``;
  * ``torch.cumsum`` on CPU accumulates fp32 inputs in DOUBLE and rounds every output to fp32
    (ATen cumsum_cpu_kernel uses at::acc_type<float, false> = double);
  * ``* 2 * np.pi`` is two fp32 multiplications (python scalars are cast to the tensor dtype).
"""
import numpy as np

F32 = np.float32


def sinusoid(f0, noise, sample_rate=16000, hop_size=160, sine_amp=0.1, noise_amp=0.003):
    """f0 (B, 1, T') fp32, noise (B, 1, T'*hop) fp32 standard normal or None -> (B, 1, T'*hop) fp32.

    features.py:178-197.  The reference draws ``noise`` with torch.randn inside the call; here it is an input so
    that CPU and GPU see identical bits."""
    f0 = np.asarray(f0, F32)
    vuv = np.repeat((f0 > 0).astype(F32), hop_size, axis=2)                      # :189
    rad = np.repeat(f0, hop_size, axis=2) / F32(sample_rate)                     # :190
    rad = np.mod(rad, F32(1.0)).astype(F32)
    cum = np.cumsum(rad.astype(np.float64), axis=2).astype(F32)                  # :191 (double accumulate)
    phase = (cum * F32(2.0)) * F32(np.pi)
    sine = (vuv * np.sin(phase).astype(F32)) * F32(sine_amp)
    if noise_amp > 0:                                                            # :192-195
        namp = vuv * F32(noise_amp) + ((F32(1.0) - vuv) * F32(noise_amp)) / F32(3.0)
        sine = sine + np.asarray(noise, F32) * namp
    return sine.astype(F32)


def vuv_binary(f0, hop_size=160):
    """features.py:199-216: (B, 1, T') -> (B, 1, T'*hop) of {0, 1}."""
    return np.repeat((np.asarray(f0, F32) > 0).astype(F32), hop_size, axis=2)


def f0_convert(f0, org_stats, trg_stats):
    """F0Statistics.convert, features.py:79-108: log-Gaussian normalised transformation of the voiced frames."""
    f0 = np.asarray(f0)
    out = np.zeros(len(f0))
    nz = f0 > 0
    out[nz] = np.exp((trg_stats[1] / org_stats[1]) * (np.log(f0[nz]) - org_stats[0]) + trg_stats[0])
    return out


def pcm16(x):
    """float waveform -> int16 as ``soundfile.write(..., "PCM_16")`` stores it (decode_fastsvc.py:193-198).

    soundfile (python-soundfile, absent from this image; the reference pins none) opens files with
    SFC_SET_CLIPPING enabled, so libsndfile converts with ``f2les_clip_array`` (src/pcm.c, published algorithm):
    ``scaled = x * 2^31`` in float; ``>= (float)0x7FFFFFFF`` -> 0x7FFF; ``<= -2^31`` -> -0x8000; otherwise the top 16
    bits of ``lrintf(scaled)``.  Parity unpinned: restated from the library's source, not checked against it here."""
    x = np.asarray(x, dtype=np.float32)
    v = x * np.float32(2147483648.0)
    q = np.rint(np.clip(v, -2147483648.0, 2147483520.0).astype(np.float64)).astype(np.int64) >> 16
    q = np.where(v >= np.float32(2147483648.0), 0x7FFF, np.where(v <= np.float32(-2147483648.0), -0x8000, q))
    return q.astype(np.int16)
