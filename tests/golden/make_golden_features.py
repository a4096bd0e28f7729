"""Golden vectors of the excitation path, generated FROM THE REFERENCE (harana/utils/features.py) on CPU.

    python tests/golden/make_golden_features.py        # build container only (needs /root/reference)

The reference draws its noise with torch.randn inside SignalGenerator.sinusoid; the same seeded draw is stored
next to the output so the fixtures do not depend on torch's RNG stream.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def main():
    for name in ("h5py", "librosa", "kaldiio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    sys.path.insert(0, REF)
    import torch
    from harana.utils.features import F0Statistics, SignalGenerator
    import harana.utils.features as feats
    assert feats.__file__.startswith(REF)

    cases = {
        # name: (B, frames, hop, sample_rate, sine_amp, noise_amp, seed)
        "sine_b2_f20": (2, 20, 160, 16000, 0.1, 0.003, 11),
        "sine_b3_f100": (3, 100, 160, 16000, 0.1, 0.003, 12),
        "sine_b1_f500": (1, 500, 160, 16000, 0.1, 0.003, 13),
        "sine_nonoise_b2_f33": (2, 33, 160, 16000, 0.1, 0.0, 14),
        "sine_hop64_24k_b2_f50": (2, 50, 64, 24000, 0.2, 0.01, 15),
    }
    for name, (B, frames, hop, sr, samp, namp, seed) in cases.items():
        rs = np.random.RandomState(seed)
        f0 = np.exp(np.log(220.0) + 0.3 * rs.randn(B, 1, frames)).astype(np.float32)
        f0[rs.rand(B, 1, frames) < 0.3] = 0.0          # unvoiced frames
        f0[:, :, frames // 2] = 0.0
        gen = SignalGenerator(sample_rate=sr, hop_size=hop, sine_amp=samp, noise_amp=namp, signal_types=["sine"])
        torch.manual_seed(seed)
        out = gen(torch.from_numpy(f0)).numpy()
        torch.manual_seed(seed)
        noise = torch.randn((B, 1, frames * hop)).numpy()   # the draw sinusoid() made (features.py:194)
        uv = SignalGenerator(sample_rate=sr, hop_size=hop, signal_types=["uv"])(torch.from_numpy(f0)).numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), f0=f0, noise=noise, out=out, uv=uv,
                            meta=np.array([B, frames, hop, sr, samp, namp], dtype=np.float64))
        print(name, out.shape, float(np.abs(out).max()))
    # F0 mean transformation (decode_fastsvc.py:176-179)
    rs = np.random.RandomState(5)
    f0 = np.exp(np.log(180.0) + 0.25 * rs.randn(300))
    f0[rs.rand(300) < 0.35] = 0.0
    src, trg = np.array([5.1, 1.0]), np.array([5.6, 1.0])
    cv = F0Statistics().convert(f0, src, trg)
    np.savez_compressed(os.path.join(HERE, "f0_convert.npz"), f0=f0, src=src, trg=trg, out=cv)
    print("f0_convert", cv.shape)


if __name__ == "__main__":
    main()
