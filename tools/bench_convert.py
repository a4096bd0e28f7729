#!/usr/bin/env python
"""BASELINE config 5: batched offline conversion of N synthetic 5-second utterances, end-to-end wall clock.

    python tools/bench_convert.py [--utts 10000] [--frames 500] [--batch 32] [--ref-utts 4]
    torchrun --nproc-per-node N ... tools/bench_convert.py ...     # round-robin shards, no data-path collective

Our arm: svcc23_fastsvc_b200.convert.BatchConverter (host packing + F0 conversion, pinned double-buffered H2D, sine
excitation + generator + PCM-16 on the GPU, int16 D2H).  Reference arm (rank 0, same process): the reference's decode
pattern -- batch 1, sequential (decode_fastsvc.py:168-198) -- through the CPU oracle port on a subsample of
``--ref-utts`` utterances, extrapolated linearly and reported as such.  Utterance records are drawn from a pool of 64
distinct synthetic utterances (bounded host memory); every record is packed, copied and converted individually.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=10000)
ap.add_argument("--frames", type=int, default=500)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--ref-utts", type=int, default=4)
args = ap.parse_args()

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)

import harana.models as M
from harana.utils.features import SignalGenerator
from svcc23_fastsvc_b200 import convert as cv, sharding, synthetic as syn

cfg = dict(syn.YAML_CONFIG)
params = syn.make_params(cfg, seed=0)
g = M.FastSVCGenerator(**{k: (list(v) if isinstance(v, list) else v) for k, v in cfg.items()})
g.remove_weight_norm()
g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
g = g.eval().to(dev)
sg = SignalGenerator(sample_rate=16000, hop_size=160, sine_amp=0.1, noise_amp=0.003, signal_types=["sine"])

rs = np.random.RandomState(1)
pool = []
for i in range(64):
    ppg, _, lft, _ = syn.make_inputs(1, args.frames, cfg, seed=100 + i)
    f0 = np.exp(np.log(220.0) + 0.3 * rs.randn(args.frames))
    f0[rs.rand(args.frames) < 0.3] = 0.0
    pool.append((ppg[0].T.copy(), f0[:, None], lft[0, 0][:, None]))
utts = [cv.Utterance(f"spk{i % 4}_{i:06d}", *pool[i % 64]) for i in range(args.utts)]
emb = rs.randn(1, 512).astype(np.float32)
src, trg = np.array([5.3, 1.0]), np.array([5.6, 1.0])

conv = cv.BatchConverter(g, sg, sampling_rate=16000, max_batch=args.batch)
warm = utts[: 2 * args.batch * world]
conv.convert(warm, spk_emb=emb, src_stats=src, trg_stats=trg, sink=lambda u, p: None, rank=rank, world=world)
torch.cuda.synchronize()
sharding.barrier()
n_done = [0]
def sink(uid, pcm):
    n_done[0] += 1
t0 = time.perf_counter()
conv.stats = dict(batches=0, utterances=0, samples=0, h2d_bytes=0, d2h_bytes=0)
conv.convert(utts, spk_emb=emb, src_stats=src, trg_stats=trg, sink=sink, rank=rank, world=world)
torch.cuda.synchronize()
sec = time.perf_counter() - t0
sec_max = sharding.max_over_ranks(sec, device=dev if world > 1 else "cpu")
sharding.barrier()

if rank == 0:
    total_samples = args.utts * args.frames * 160
    line = {
        "metric": "batched offline conversion, end-to-end wall clock (BASELINE configs[4])", "unit": "s",
        "value": sec_max, "higher_is_better": False, "n_gpus": world, "utterances": args.utts,
        "seconds_of_audio": total_samples / 16000.0, "samples_per_s": total_samples / sec_max,
        "realtime_factor": sec_max / (total_samples / 16000.0), "batch": args.batch, "frames": args.frames,
        "rank0": dict(conv.stats, utterances_written=n_done[0]), "data": "synthetic",
        "pipeline": "host pack + F0 conversion -> pinned H2D (side stream) -> sine excitation + generator + PCM-16 (GPU) -> int16 D2H",
    }
    if args.ref_utts > 0:
        from oracle import fastsvc_torch as otorch, features_numpy as fo
        tp = {k: torch.from_numpy(v) for k, v in params.items()}
        t1 = time.perf_counter()
        for u in utts[: args.ref_utts]:      # decode_fastsvc.py:168-198 at batch 1 on the host cores
            f0 = fo.f0_convert(np.squeeze(u.f0, 1), src, trg).astype(np.float32)[None, None]
            noise = np.random.RandomState(0).randn(1, 1, args.frames * 160).astype(np.float32)
            s = fo.sinusoid(f0, noise, 16000, 160, 0.1, 0.003)
            with torch.no_grad():
                y = otorch.generator_forward(tp, torch.from_numpy(u.ppg.T[None].copy()), torch.from_numpy(s),
                                             torch.from_numpy(u.lft.T[None].copy()), torch.from_numpy(emb),
                                             recompute=True)
            np.clip(np.rint(y.numpy().reshape(-1) * 32767.0), -32768, 32767).astype(np.int16)
        ref_sec = (time.perf_counter() - t1) / args.ref_utts
        line["cpu_reference"] = {"kind": "port", "cores": torch.get_num_threads(), "seconds_per_utterance": ref_sec,
                                 "extrapolated_s": ref_sec * args.utts,
                                 "sample": f"{args.ref_utts} utterances at batch 1, extrapolated linearly to {args.utts}"}
        line["speedup_vs_cpu_reference_extrapolated"] = ref_sec * args.utts / sec_max
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
