"""Batched offline conversion: the reference's decode loop, harana/bin/decode_fastsvc.py:150-200,
re-planned for one B200 per process.

The reference converts one utterance at a time (batch 1): F0 mean transformation on the host (decode:176-179),
``model.inference`` = sine excitation + generator forward (fastsvc.py:364-383), D2H of the fp32 waveform and
``soundfile.write(..., "PCM_16")`` (decode:193-198).  Here:

  * utterances are sharded round-robin over ranks (``i % world == rank``; no data-path collective) and grouped into
    EQUAL-LENGTH batches -- InstanceNorm statistics run over each utterance's whole time axis, so padding a shorter
    utterance would change its output; equal-length batching keeps every waveform identical to batch-1 decoding;
  * host staging is pinned and double-buffered: while batch k computes, batch k+1 is packed (F0 conversion in
    numpy, as in the reference) and copied H2D on a side stream;
  * sine excitation (``fsvc_sine_excitation``), generator forward (``fsvc_forward``) and PCM-16 quantisation
    (``fsvc_pcm16``) run on the device; only int16 samples travel back (half the bytes of the reference's fp32 D2H).

File formats: the reference reads per-utterance HDF5 ({ppg (T',C), f0 (T',1), lft (T,1)}, preprocess_fastsvc.py:
270-292) and writes PCM-16 wav with soundfile; h5py / soundfile are not part of this image, so the driver takes
in-memory ``Utterance`` records and ``write_wav`` uses the stdlib ``wave`` module (same RIFF/PCM-16 payload).
"""
import os
import wave
from collections import OrderedDict, namedtuple

import numpy as np
import torch

from . import features
from .sharding import shard_utterances

Utterance = namedtuple("Utterance", ["utt_id", "ppg", "f0", "lft"])  # ppg (T', C), f0 (T',) or (T', 1), lft (T,) or (T, 1)


def plan_batches(frame_counts, max_batch, rank=0, world=1):
    """Batches (lists of utterance indices) for one rank: round-robin shard, then equal-length groups of at most
    ``max_batch`` in first-appearance order.  Pure host logic (tested on CPU, world_size > 1 included)."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    mine = shard_utterances(len(frame_counts), rank, world)
    groups = OrderedDict()
    for i in mine:
        groups.setdefault(int(frame_counts[i]), []).append(i)
    batches = []
    for _, idx in groups.items():
        for k in range(0, len(idx), max_batch):
            batches.append(idx[k:k + max_batch])
    return batches


def write_wav(path, pcm, sampling_rate):
    """int16 mono samples -> RIFF/WAVE PCM-16 (what soundfile.write(path, y, sr, "PCM_16") produces)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sampling_rate))
        w.writeframes(pcm.tobytes())


class _Staging:
    """One pinned host staging set + its device mirror + events (two of these are used alternately).

    Buffers are flat, grow-only and re-viewed per utterance length.  They are never freed while work is in flight:
    device memory released to the caching allocator could be handed to a tensor of the compute stream while the
    copy stream still writes it (the allocator only orders reuse within one stream), so growth synchronises first."""

    def __init__(self, max_batch, hop, in_channels, device):
        self.capacity, self.hop, self.cin, self.device = max_batch, hop, in_channels, device
        self.cap_frames = 0
        self.h2d_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.consumed = torch.cuda.Event()  # compute no longer reads the device mirrors

    def shape(self, frames):
        if frames > self.cap_frames:
            torch.cuda.synchronize(self.device)
            B, T = self.capacity, frames * self.hop
            self._ppg = torch.empty(B * self.cin * frames, dtype=torch.float32, pin_memory=True)
            self._f0 = torch.empty(B * frames, dtype=torch.float32, pin_memory=True)
            self._lft = torch.empty(B * T, dtype=torch.float32, pin_memory=True)
            self._pcm = torch.empty(B * T, dtype=torch.int16, pin_memory=True)
            self._d_ppg = torch.empty(B * self.cin * frames, dtype=torch.float32, device=self.device)
            self._d_f0 = torch.empty(B * frames, dtype=torch.float32, device=self.device)
            self._d_lft = torch.empty(B * T, dtype=torch.float32, device=self.device)
            self.cap_frames = frames
        B, T = self.capacity, frames * self.hop
        self.frames = frames
        self.ppg = self._ppg[:B * self.cin * frames].view(B, self.cin, frames)
        self.f0 = self._f0[:B * frames].view(B, 1, frames)
        self.lft = self._lft[:B * T].view(B, 1, T)
        self.pcm = self._pcm[:B * T].view(B, T)
        self.d_ppg = self._d_ppg[:B * self.cin * frames].view(B, self.cin, frames)
        self.d_f0 = self._d_f0[:B * frames].view(B, 1, frames)
        self.d_lft = self._d_lft[:B * T].view(B, 1, T)
        return self


class BatchConverter:
    """Offline conversion of many utterances with one generator on one GPU.

    Args mirror decode_fastsvc.py: ``generator`` (eval mode, weight norm removed, on a CUDA device),
    ``signal_generator`` (``harana.utils.features.SignalGenerator`` with signal_types ["sine"]),
    ``sampling_rate``; ``max_batch`` utterances per launch.
    """

    def __init__(self, generator, signal_generator, sampling_rate=16000, max_batch=32, writer_threads=0):
        self.writer_threads = writer_threads   # convert_to_dir: wav files written by a small thread pool (0 = inline)
        self.g = generator
        self.sg = signal_generator
        self.sampling_rate = sampling_rate
        self.max_batch = max_batch
        self.device = next(generator.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("BatchConverter needs the generator on a CUDA device (no CPU fallback)")
        if list(signal_generator.signal_types) != ["sine"]:
            raise ValueError("the FastSVC generator takes the 'sine' excitation only (fastsvc.yaml signal_generator)")
        if signal_generator.hop_size != generator.hop_size:
            raise ValueError(f"signal generator hop {signal_generator.hop_size} != generator hop {generator.hop_size}")
        self.f0stats = features.F0Statistics()
        self._stage = {}
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self.stats = dict(batches=0, utterances=0, samples=0, h2d_bytes=0, d2h_bytes=0)

    # ---- host side -------------------------------------------------------------------------------------------
    def _staging(self, frames, slot):
        if slot not in self._stage:
            self._stage[slot] = _Staging(self.max_batch, self.g.hop_size, self.g.in_channels, self.device)
        return self._stage[slot].shape(frames)

    def _pack(self, st, utts, src_stats, trg_stats):
        hop = self.g.hop_size
        for j, u in enumerate(utts):
            ppg = np.asarray(u.ppg, dtype=np.float32)
            frames = ppg.shape[0]
            f0 = np.asarray(u.f0, dtype=np.float64).reshape(-1)
            lft = np.asarray(u.lft, dtype=np.float32).reshape(-1)
            if ppg.shape[1] != self.g.in_channels or len(f0) != frames or len(lft) != frames * hop:
                raise ValueError(f"{u.utt_id}: ppg {ppg.shape}, f0 {f0.shape}, lft {lft.shape} do not describe "
                                 f"{frames} frames of hop {hop}")
            if src_stats is not None:  # mean transformation towards the target speaker (decode:176-179)
                f0 = self.f0stats.convert(f0, src_stats[u.utt_id] if isinstance(src_stats, dict) else src_stats,
                                          trg_stats)
            st.ppg[j].copy_(torch.from_numpy(ppg.T))
            st.f0[j, 0].copy_(torch.from_numpy(f0.astype(np.float32)))
            st.lft[j, 0].copy_(torch.from_numpy(lft))

    # ---- device side -----------------------------------------------------------------------------------------
    def _upload(self, st, n):
        cs = self._copy_stream
        cs.wait_event(st.consumed)
        with torch.cuda.stream(cs):
            st.d_ppg[:n].copy_(st.ppg[:n], non_blocking=True)
            st.d_f0[:n].copy_(st.f0[:n], non_blocking=True)
            st.d_lft[:n].copy_(st.lft[:n], non_blocking=True)
            st.h2d_done.record(cs)
        self.stats["h2d_bytes"] += 4 * n * (st.ppg[0].numel() + st.f0[0].numel() + st.lft[0].numel())

    def _compute(self, st, n, spk_emb):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(st.h2d_done)
        with torch.no_grad():
            sine = self.sg(st.d_f0[:n])
            y = self.g(st.d_ppg[:n], sine, st.d_lft[:n], spk_emb)
            pcm = features.pcm16(y.view(n, -1))
        st.consumed.record(cur)
        st.pcm[:n].copy_(pcm, non_blocking=True)
        st.d2h_done.record(cur)
        self.stats["d2h_bytes"] += 2 * pcm.numel()

    def convert(self, utterances, spk_emb=None, src_stats=None, trg_stats=None, sink=None, rank=0, world=1):
        """Convert this rank's share of ``utterances``; returns {utt_id: int16 ndarray} (or feeds ``sink(utt_id,
        pcm)``).  ``spk_emb``: (1, S) / (S,) target x-vector on any device; ``src_stats`` ([mean, std] or a dict
        per utt_id) and ``trg_stats``: log-F0 statistics for the mean transformation, or None to keep F0."""
        if (src_stats is None) != (trg_stats is None):
            raise ValueError("src_stats and trg_stats go together")
        if spk_emb is not None:
            spk_emb = torch.as_tensor(spk_emb, dtype=torch.float32).reshape(1, -1).to(self.device)
        frame_counts = [np.asarray(u.ppg).shape[0] for u in utterances]
        batches = plan_batches(frame_counts, self.max_batch, rank, world)
        results = {} if sink is None else None

        def drain(st, utts):
            st.d2h_done.synchronize()
            for j, u in enumerate(utts):
                pcm = st.pcm[j].numpy().copy()
                if sink is None:
                    results[u.utt_id] = pcm
                else:
                    sink(u.utt_id, pcm)

        # Per iteration k: enqueue the compute of batch k-1 (already uploaded), read back batch k-2 (its D2H precedes
        # that compute in the stream), then pack + upload batch k into the staging slot batch k-2 just released --
        # the host packs while the device computes.
        pending = None  # (staging, utterances): computed, D2H in flight
        staged = None   # (staging, utterances): packed, H2D in flight
        with torch.cuda.device(self.device):
            for k, idx in enumerate(batches + [None]):
                computed = None
                if staged is not None:
                    st2, utts2 = staged
                    self._compute(st2, len(utts2), spk_emb)
                    computed = staged
                    self.stats["batches"] += 1
                    self.stats["utterances"] += len(utts2)
                    self.stats["samples"] += len(utts2) * st2.frames * self.g.hop_size
                if pending is not None:
                    drain(*pending)
                pending, staged = computed, None
                if idx is not None:
                    utts = [utterances[i] for i in idx]
                    st = self._staging(frame_counts[idx[0]], k & 1)
                    self._pack(st, utts, src_stats, trg_stats)
                    self._upload(st, len(utts))
                    staged = (st, utts)
            if pending is not None:
                drain(*pending)
        return results

    def convert_to_dir(self, utterances, outdir, suffix="_gen", **kw):
        """decode_fastsvc.py:193-198: one ``{utt_id}{suffix}.wav`` (PCM-16) per utterance."""
        os.makedirs(outdir, exist_ok=True)

        def write(uid, pcm):
            write_wav(os.path.join(outdir, f"{uid}{suffix}.wav"), pcm, self.sampling_rate)

        if self.writer_threads <= 0:
            self.convert(utterances, sink=write, **kw)
            return
        # file writing releases the GIL: a few threads keep it off the packing / launching thread
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=self.writer_threads) as pool:
            futures = []
            self.convert(utterances, sink=lambda uid, pcm: futures.append(pool.submit(write, uid, pcm)), **kw)
            for f in futures:
                f.result()
