"""Batched offline conversion driver (reference harana/bin/decode_fastsvc.py:150-200).

CPU: the batching / sharding plan (incl. two gloo ranks).  GPU: every waveform equals batch-1 decoding through the
reference-shaped ``inference`` call, bit for bit, and equals the CPU oracle within the forward tolerance."""
import os
import socket
import wave

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svcc23_fastsvc_b200 import convert as cv
from svcc23_fastsvc_b200 import sharding


def test_plan_batches_partitions_equal_lengths():
    rs = np.random.RandomState(0)
    frames = rs.choice([51, 100, 100, 250, 500], size=37).tolist()
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            batches = cv.plan_batches(frames, 4, rank, world)
            mine = [i for b in batches for i in b]
            assert sorted(mine) == sharding.shard_utterances(len(frames), rank, world)
            for b in batches:
                assert 1 <= len(b) <= 4 and len({frames[i] for i in b}) == 1
            seen += mine
        assert sorted(seen) == list(range(len(frames)))
    assert cv.plan_batches([], 4) == []
    with pytest.raises(ValueError):
        cv.plan_batches([1], 0)


def test_write_wav_roundtrip(tmp_path):
    pcm = (np.arange(-500, 500) * 60).astype(np.int16)
    path = os.path.join(tmp_path, "a.wav")
    cv.write_wav(path, pcm, 16000)
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, 1000)
        assert np.array_equal(np.frombuffer(w.readframes(1000), dtype="<i2"), pcm)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _plan_worker(rank, world, port, frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batches = cv.plan_batches(frames, 3, rank, world)
        mine = [i for b in batches for i in b]
        outs = [np.full(4, i, dtype=np.int16) for i in mine]           # stands in for the converted PCM
        full = sharding.gather_in_order(outs, mine, len(frames))
        q.put((rank, batches, None if rank else [int(x[0]) for x in full]))
    finally:
        dist.destroy_process_group()


def test_two_rank_conversion_plan_gloo():
    frames = [100, 51, 100, 100, 51, 500, 100, 51, 100]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [[0, 2, 6], [8], [4]] and res[1][1] == [[1, 7], [3], [5]]
    assert res[0][2] == list(range(len(frames)))


def _make_utts(n, frame_choices, cfg, seed):
    from svcc23_fastsvc_b200 import synthetic as syn
    rs = np.random.RandomState(seed)
    utts = []
    for i in range(n):
        frames = int(frame_choices[i % len(frame_choices)])
        ppg, _, lft, _ = syn.make_inputs(1, frames, cfg, seed=seed + 17 * i)
        f0 = np.exp(np.log(200.0) + 0.2 * rs.randn(frames))
        f0[rs.rand(frames) < 0.3] = 0.0
        utts.append(cv.Utterance(f"spk{i % 2}_utt{i:03d}", ppg[0].T.copy(), f0[:, None], lft[0, 0][:, None]))
    return utts


@pytest.mark.gpu
def test_batched_conversion_equals_batch1_decoding(tmp_path):
    import harana.models as M
    from harana.utils.features import F0Statistics, SignalGenerator
    from oracle import fastsvc_numpy as onp
    from oracle import features_numpy as fo
    from svcc23_fastsvc_b200 import features, synthetic as syn

    dev = torch.device("cuda:0")
    cfg = dict(syn.YAML_CONFIG)
    params = syn.make_params(cfg, seed=0)
    g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
    g.remove_weight_norm()
    g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    g = g.eval().to(dev)
    # noise_amp = 0: the excitation is deterministic, so batched and batch-1 runs see the same signal
    sg = SignalGenerator(sample_rate=16000, hop_size=160, sine_amp=0.1, noise_amp=0.0, signal_types=["sine"])
    utts = _make_utts(11, [20, 33, 20, 20, 7], cfg, seed=3)
    rs = np.random.RandomState(9)
    trg_emb = rs.randn(1, 512).astype(np.float32)
    src_stats = {u.utt_id: np.array([5.2 + 0.1 * (i % 2), 1.0]) for i, u in enumerate(utts)}
    trg_stats = np.array([5.5, 1.0])

    conv = cv.BatchConverter(g, sg, sampling_rate=16000, max_batch=3)
    got = conv.convert(utts, spk_emb=trg_emb, src_stats=src_stats, trg_stats=trg_stats)
    assert set(got) == {u.utt_id for u in utts}
    assert conv.stats["utterances"] == len(utts) and conv.stats["batches"] == 5

    # two ranks produce the same waveforms for their shares
    for rank in range(2):
        part = cv.BatchConverter(g, sg, max_batch=3).convert(utts, spk_emb=trg_emb, src_stats=src_stats,
                                                             trg_stats=trg_stats, rank=rank, world=2)
        assert set(part) == {utts[i].utt_id for i in range(rank, len(utts), 2)}
        for k, v in part.items():
            assert np.array_equal(v, got[k])

    pad_fn = torch.nn.ReplicationPad1d(0)
    emb = torch.from_numpy(trg_emb).to(dev)
    for u in utts:
        # the reference's decode loop body (decode_fastsvc.py:168-198), batch 1
        f0 = F0Statistics().convert(np.squeeze(u.f0, 1), src_stats[u.utt_id], trg_stats)
        f0_t = torch.FloatTensor(np.expand_dims(f0, 1)).to(dev)
        with torch.no_grad():
            y = g.inference(torch.FloatTensor(u.ppg).to(dev), f0_t, torch.FloatTensor(u.lft).to(dev), sg, pad_fn,
                            emb).view(-1)
        want = features.pcm16(y).cpu().numpy()
        assert got[u.utt_id].dtype == np.int16 and len(got[u.utt_id]) == len(f0) * 160
        assert np.array_equal(got[u.utt_id], want), u.utt_id
        # CPU oracle of the whole chain: F0 conversion -> excitation -> generator -> PCM-16 (<= 1e-3 * 32767 ~ 33 LSB)
        f0_o = fo.f0_convert(np.squeeze(u.f0, 1), src_stats[u.utt_id], trg_stats).astype(np.float32)[None, None]
        s_o = fo.sinusoid(f0_o, None, 16000, 160, 0.1, 0.0)
        y_o = onp.generator_forward(params, u.ppg.T[None], s_o, u.lft.T[None], trg_emb)
        pcm_o = fo.pcm16(y_o.reshape(-1)).astype(np.float64)
        clipped = np.abs(y_o.reshape(-1)) >= 0.999
        assert np.abs(got[u.utt_id].astype(np.int64) - pcm_o)[~clipped].max() <= 34

    conv.convert_to_dir(utts[:2], str(tmp_path), suffix="_trg_gen", spk_emb=trg_emb)
    with wave.open(os.path.join(tmp_path, f"{utts[0].utt_id}_trg_gen.wav"), "rb") as w:
        assert w.getframerate() == 16000 and w.getnframes() == 20 * 160
