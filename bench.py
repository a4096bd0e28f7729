#!/usr/bin/env python
"""Benchmarks of the B200-native FastSVC generator path (BASELINE.json metric and configs).

    python bench.py --gpus N --steps K --warmup W                       # configs[1]: generator forward, batch 32 x 1 s
    python bench.py --config train   --gpus N ...                       # configs[2]/[3]: GAN training step, batch 16/GPU
    python bench.py --config convert --gpus N [--utts 10000]            # configs[4]: batched offline conversion
    python bench.py --impl reference [--config ...] --gpus N ...        # the reference's CPU path (oracle port), rank 0

A "step" is one pass of the path over one batch of synthetic input: one generator forward (infer), one
``Trainer._train_step`` (train: generator fwd+bwd, STFT + adversarial losses, critic update), one batch of utterances
converted end to end (convert).  For N > 1 launch under torchrun (one rank per GPU): every rank works on its own
batch (weak scaling); inference / conversion shard by utterance with no data-path collective, training all-reduces the
gradient buckets over NCCL.  The step time is the max over ranks.  Prints ONE JSON line on rank 0.
"""

import argparse
import json
import os
import platform
import shutil
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (REPO, os.path.join(REPO, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "audio samples/sec (generator fwd, bs32, 16k-sample clips)"
UNIT = "samples/s"
B, FRAMES = 32, 100          # BASELINE configs[1]: batch 32, 1-second clips
TRAIN_B, TRAIN_FRAMES = 16, 51   # configs[2]/[3]: batch 16, 8192 -> 8160 samples (collater rounds to the hop, D7)
CONV_FRAMES = 500            # configs[4]: 5-second utterances
L2_FLUSH_BYTES = 256 << 20   # > 126 MB L2


def _peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


def _env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def _yaml_kwargs():
    from svcc23_fastsvc_b200 import synthetic as syn
    return {k: (list(v) if isinstance(v, list) else v) for k, v in syn.YAML_CONFIG.items()}


# =====================================================================================================================
# CPU arms: the reference's op sequence (oracle/fastsvc_torch.py, recompute=True) on the host cores
# =====================================================================================================================
def _cpu_forward(tp, ppg, sine, lft, spk, lo, hi):
    import torch
    from oracle import fastsvc_torch as otorch
    args = [torch.from_numpy(a[lo:hi]) for a in (ppg, sine, lft, spk)]
    with torch.no_grad():
        return otorch.generator_forward(tp, *args, recompute=True)


def _time_cpu(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def _use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    import torch
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def reference_infer(args):
    import torch
    from svcc23_fastsvc_b200 import synthetic as syn
    cores = _use_all_cores()
    params = syn.make_params(syn.YAML_CONFIG, seed=0)
    ppg, sine, lft, spk = syn.make_inputs(B, FRAMES, syn.YAML_CONFIG, seed=1234)
    tp = {k: torch.from_numpy(v) for k, v in params.items()}
    T = FRAMES * 160
    dt = _time_cpu(lambda: _cpu_forward(tp, ppg, sine, lft, spk, 0, B), args.steps, max(1, min(args.warmup, 2)))
    value = B * T / dt
    sample = (f"the whole configs[1] batch per step: {B} utterances x {T} samples, fp32, torch CPU ops in the "
              f"reference's order (recompute as downsampling_loop does), {cores} threads on {_cpu_model()}")
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: batch 32 x 16000-sample clips (100 PPG frames), YAML generator "
                               "in=144 mid=[192,96,48,24] scales=[2,4,4,5] spk=512"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "cpu_model": _cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def _cpu_baseline_legs(params, ppg, sine, lft, spk):
    """cpu_baseline of the infer line (rank 0, N=1): all cores at B=32, plus the two variants BASELINE.md section 4
    asks for -- OMP_NUM_THREADS=1 (what the recipe's path.sh sets) and B=1 x 10."""
    import torch
    tp = {k: torch.from_numpy(v) for k, v in params.items()}
    T = FRAMES * 160
    cores = _use_all_cores()
    dt = _time_cpu(lambda: _cpu_forward(tp, ppg, sine, lft, spk, 0, B), 3)
    base = {"value": B * T / dt, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": _cpu_model(),
            "sample": f"3 forwards of the whole batch ({B} utterances x {T} samples), fp32 torch CPU ops in the "
                      f"reference's order, host has {os.cpu_count()} cpus"}
    dt1 = _time_cpu(lambda: [_cpu_forward(tp, ppg, sine, lft, spk, i, i + 1) for i in range(10)], 1)
    b1 = {"value": 10 * T / dt1, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"10 forwards at batch 1 ({T} samples each): the reference's decode pattern"}
    torch.set_num_threads(1)
    nb = 4
    dt2 = _time_cpu(lambda: _cpu_forward(tp, ppg, sine, lft, spk, 0, nb), 1)
    omp1 = {"value": nb * T / dt2, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 forward of {nb} of the {B} utterances with ONE thread (egs/svcc23/fastsvc1/path.sh:13 "
                      "exports OMP_NUM_THREADS=1)"}
    _use_all_cores()
    return base, omp1, b1


# =====================================================================================================================
# shared GPU helpers
# =====================================================================================================================
class Gpu:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world, self.rank, self.local = _env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
        self.dev = torch.device("cuda", self.local)
        torch.cuda.set_device(self.dev)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        import torch
        from svcc23_fastsvc_b200 import sharding
        sharding.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, flush=True):
        """Mean device time of `steps` calls (CUDA events on the current stream, L2 flushed between iterations outside
        the event pair), max over ranks."""
        import torch
        from svcc23_fastsvc_b200 import sharding
        for _ in range(warmup):
            fn()
        self.barrier()
        evs = []
        for _ in range(steps):
            if flush:
                self.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        self.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        return sharding.max_over_ranks(ms, self.dev)

    def close(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.destroy_process_group()


def _generator(dev, weight_norm=False, seed=0):
    import torch
    import harana.models as M
    from svcc23_fastsvc_b200 import synthetic as syn
    params = syn.make_params(syn.YAML_CONFIG, seed=seed, weight_norm=weight_norm)
    g = M.FastSVCGenerator(**_yaml_kwargs())
    if not weight_norm:
        g.remove_weight_norm()
    g.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    return g.to(dev), params


# =====================================================================================================================
# configs[1]: generator forward
# =====================================================================================================================
def _roofline(g, devin, gpu, n_stages):
    """Per-kernel profile (CUDA events around every launch, on the launching stream) -> the roofline object."""
    peaks = _peaks()
    recs = None
    for _ in range(3):
        gpu.flush.zero_()
        recs = g.profile(*devin)
    total = sum(r["ms"] for r in recs)
    agg = {}
    for r in recs:
        a = agg.setdefault(r["label"], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
        a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["bytes"] += r["bytes"]; a["n"] += 1
    name, a = max(agg.items(), key=lambda kv: kv[1]["ms"])
    traffic, traffic_src = None, None
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(REPO, "profiles", fn)
        if os.path.exists(tpath):
            with open(tpath) as f:
                t = json.load(f).get(name)
            if t is not None:
                traffic, traffic_src = t, f"profiles/{fn} (ncu --set full capture of an earlier run, NOT measured in this run)"
                break
    gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9
    tfs = a["flops"] / (a["ms"] * 1e-3) / 1e12
    tensor_bound = "fused" in name
    if tensor_bound:
        roof = {"bound": "tensor", "achieved": tfs, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": tfs / peaks["bf16_tflops"], "traffic": traffic,
                "note": "ALGORITHMIC flops (2*Cin*Cout*K*T per conv, SURVEY 8d); the 3-term bf16 split issues 3x "
                        "that on the tensor pipe"}
    else:
        roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": gbs / peaks["hbm_gbs"], "traffic": traffic}
    last = f"s{n_stages - 1}"
    dil = {k: v for k, v in agg.items() if k.split(".")[0] == last and k.split(".")[1][:2] in ("d3", "d9", "d2")}
    all_dil = {k: v for k, v in agg.items() if k[0] == "s" and "." in k and k.split(".")[1][:2] in ("d3", "d9", "d2")}

    def frac(group):
        by, ms = sum(v["bytes"] for v in group.values()), sum(v["ms"] for v in group.values())
        return (by / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if ms > 0 else None

    roof.update({
        "kernel": name, "kernel_ms": a["ms"], "kernel_share_of_step": a["ms"] / total, "kernel_gbs": gbs,
        "kernel_tflops": tfs, "peak_source": peaks["source"], "traffic_source": traffic_src,
        "byte_model": "algorithmic bytes of the launch: every operand tensor touched once, fp32 (DESIGN.md section 5)",
        "whole_forward": {"ms_sum_of_kernels": total, "algorithmic_gflop": sum(r["flops"] for r in recs) / 1e9,
                          "tflops": sum(r["flops"] for r in recs) / (total * 1e-3) / 1e12,
                          "algorithmic_GB": sum(r["bytes"] for r in recs) / 1e9,
                          "GBps": sum(r["bytes"] for r in recs) / (total * 1e-3) / 1e9,
                          "frac_of_hbm_peak": sum(r["bytes"] for r in recs) / (total * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        # the layers BASELINE.json's 60 %-of-HBM target names
        "dilated_conv_stage": {
            "last_stage": [dict(kernel=k, ms=v["ms"], gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                                frac_of_hbm_peak=v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"])
                           for k, v in dil.items()],
            "last_stage_aggregate_frac": frac(dil),
            "aggregate_frac": frac(all_dil),      # bytes-weighted over the dilated convs of ALL stages
            "n_kernels": len(all_dil)},
        "top5": [dict(kernel=k, ms=v["ms"], share=v["ms"] / total, gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                      tflops=v["flops"] / (v["ms"] * 1e-3) / 1e12)
                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:5]]})
    return roof


def run_infer(args):
    import torch
    from svcc23_fastsvc_b200 import sharding, synthetic as syn
    gpu = Gpu()
    dev, rank, world = gpu.dev, gpu.rank, gpu.world
    T = FRAMES * 160
    g, params = _generator(dev)
    g = g.eval()
    g.precision = args.precision
    ppg, sine, lft, spk = syn.make_inputs(B, FRAMES, syn.YAML_CONFIG, seed=1234 + rank)
    host = [torch.from_numpy(a).pin_memory() for a in (ppg, sine, lft, spk)]
    devin = [t.to(dev) for t in host]
    out_host = torch.empty((B, 1, T), dtype=torch.float32).pin_memory()

    with torch.no_grad():
        sampler = ClockSampler(gpu.local)
        if rank == 0:
            sampler.start()
        ms = gpu.timed(lambda: g(*devin), args.steps, args.warmup)
        launches = g.last_launch_count()
        ms_e2e = gpu.timed(lambda: g.forward_host(*host, out=out_host), args.steps, max(3, args.warmup // 2))
        clocks = sampler.stop() if rank == 0 else None

        parity = roof = cpu_base = omp1 = b1 = eager = None
        if rank == 0 and not args.quick:
            # parity of what was just timed (2 utterances vs the CPU oracle port)
            tp = {k: torch.from_numpy(v) for k, v in params.items()}
            y = g(*devin)[:2].cpu()
            parity = float((y - _cpu_forward(tp, ppg, sine, lft, spk, 0, 2)).abs().max())
            roof = _roofline(g, devin, gpu, len(syn.YAML_CONFIG["mid_channels"]))
            if world == 1:
                cpu_base, omp1, b1 = _cpu_baseline_legs(params, ppg, sine, lft, spk)
            # the reference's op sequence through stock PyTorch eager on this GPU (denominator of the 20x target)
            if not args.no_eager and world == 1:
                from oracle import fastsvc_torch as otorch
                torch.backends.cudnn.benchmark = True     # train_fastsvc.py:617
                tpd = {k: torch.from_numpy(v).to(dev) for k, v in params.items()}
                ems = gpu.timed(lambda: otorch.generator_forward(tpd, *devin, recompute=True), 10, 5)
                eager = {"ms_per_step": ems, "value": B * T / (ems * 1e-3), "unit": UNIT, "ratio_ours": ems / ms,
                         "what": "reference op sequence, PyTorch eager CUDA fp32 (cuDNN), same inputs"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": sharding.aggregate_throughput(B * T, world, ms), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: batch 32/GPU x 16000-sample clips (100 PPG frames), YAML generator "
                                   "in=144 mid=[192,96,48,24] scales=[2,4,4,5] spk=512",
                       "precision_mode": g.precision, "l2": "flushed (256 MiB memset) between timed iterations",
                       "parallelism": f"utterance-sharded x{world}, no collectives"},
            "e2e": {"value": sharding.aggregate_throughput(B * T, world, ms_e2e), "unit": UNIT,
                    "h2d_bytes_per_step": sum(t.numel() * 4 for t in host), "d2h_bytes_per_step": out_host.numel() * 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches * args.steps, "launches_per_step": launches, "clocks": clocks,
            "roofline": roof, "cpu_baseline": cpu_base, "cpu_baseline_omp1": omp1, "cpu_baseline_b1": b1,
            "torch_eager_cuda": eager, "parity_max_abs_vs_oracle": parity,
        }
        print(json.dumps(line))
    gpu.close()


# =====================================================================================================================
# configs[2]/[3]: GAN training step (generator fwd/bwd native, critic + losses host PyTorch), data parallel
# =====================================================================================================================
TRAIN_METRIC = "training audio samples/sec (GAN step: generator fwd+bwd + HiFiGAN MSMPD critic, bs16/GPU, 8160-sample segments)"


def _train_objects(dev, generator):
    import torch
    import gan_host
    from svcc23_fastsvc_b200.training import GanTrainer
    torch.manual_seed(0)
    D = gan_host.MultiScaleMultiPeriodCritic().to(dev)
    stft = gan_host.MultiResolutionSTFTLoss(**gan_host.STFT_PARAMS).to(dev)
    opt_g = torch.optim.RAdam(generator.parameters(), lr=1e-3, eps=1e-6)      # fastsvc.yaml:86-105
    opt_d = torch.optim.RAdam(D.parameters(), lr=1e-3, eps=1e-6)
    return GanTrainer(generator, D, stft, gan_host.generator_adversarial_loss, gan_host.discriminator_adversarial_loss,
                      opt_g, opt_d, lambda_adv=2.5, generator_grad_norm=10.0, discriminator_grad_norm=1.0), D


def _train_batch(rank):
    import numpy as np
    from svcc23_fastsvc_b200 import synthetic as syn
    ppg, sine, lft, spk = syn.make_inputs(TRAIN_B, TRAIN_FRAMES, syn.YAML_CONFIG, seed=4321 + rank)
    y = (0.1 * np.random.RandomState(99 + rank).standard_normal(size=(TRAIN_B, 1, TRAIN_FRAMES * 160))).astype("float32")
    return ppg, sine, lft, spk, y


def run_train(args):
    import torch
    from svcc23_fastsvc_b200 import sharding
    gpu = Gpu()
    dev, rank, world = gpu.dev, gpu.rank, gpu.world
    T = TRAIN_FRAMES * 160
    torch.backends.cudnn.benchmark = True      # train_fastsvc.py:617 (critic + STFT run on cuDNN / cuFFT)
    g, params = _generator(dev, weight_norm=True)
    g = g.train()
    g.precision = args.precision
    trainer, D = _train_objects(dev, g)
    host = [torch.from_numpy(a).pin_memory() for a in _train_batch(rank)]
    devin = [t.to(dev) for t in host]
    staged = [torch.empty_like(t) for t in devin]
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    counts = {}

    def step_resident():
        trainer.step(tuple(devin[:4]), devin[4], adversarial=True)

    def step_e2e():
        for d, h in zip(staged, host):
            d.copy_(h, non_blocking=True)                                   # train_fastsvc.py:161-162
        logs = trainer.step(tuple(staged[:4]), staged[4], adversarial=True)
        loss_host.copy_(logs["generator_loss"].reshape(1), non_blocking=True)   # the .item() of :196

    sampler = ClockSampler(gpu.local)
    if rank == 0:
        sampler.start()
    ms = gpu.timed(step_resident, args.steps, args.warmup, flush=False)
    ms_e2e = gpu.timed(step_e2e, args.steps, max(3, args.warmup // 2), flush=False)
    trainer.finish_discriminator_step()
    clocks = sampler.stop() if rank == 0 else None

    # our kernels per step: training forward + backward, and the no-grad forward of the critic phase
    y_ = g(*devin[:4])
    counts["forward_train"] = g.last_launch_count()
    (y_ * devin[4]).sum().backward()
    counts["backward"] = g.last_launch_count()
    with torch.no_grad():
        g(*devin[:4])
    counts["forward_nograd"] = g.last_launch_count()
    launches = sum(counts.values())

    parts = eager = cpu_base = None
    if rank == 0:
        w = devin[4]

        def g_only():
            trainer.gb.zero()
            (g(*devin[:4]) * w).sum().backward()

        def d_only():
            with torch.no_grad():
                yy = g(*devin[:4])
            trainer.db.zero()
            real, fake = trainer.dis_adv_loss(D(yy), D(devin[4]))
            (real + fake).backward()

        parts = {"generator_fwd_bwd_native_ms": gpu.timed(g_only, 10, 3, flush=False) if world == 1 else None,
                 "critic_fwd_bwd_host_torch_ms": gpu.timed(d_only, 10, 3, flush=False) if world == 1 else None}
        if not args.no_eager and world == 1:
            from oracle import fastsvc_torch as otorch
            tpd = {k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in params.items()}

            def eager_g():
                for t in tpd.values():
                    t.grad = None
                (otorch.generator_forward(tpd, *devin[:4], recompute=True) * w).sum().backward()

            ems = gpu.timed(eager_g, 5, 2, flush=False)
            eager = {"generator_fwd_bwd_ms": ems, "ratio_ours": ems / parts["generator_fwd_bwd_native_ms"],
                     "what": "reference generator op sequence (weight norm applied), PyTorch eager CUDA autograd, "
                             "cuDNN fp32/TF32 defaults, same inputs"}
        if world == 1:
            cpu_base = _train_cpu_step(2, 1)

    if rank == 0:
        gbytes, dbytes = trainer.gb.numel * 4, trainer.db.numel * 4
        line = {
            "metric": TRAIN_METRIC, "value": sharding.aggregate_throughput(TRAIN_B * T, world, ms), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[{2 if world == 1 else 3}]: GAN training step, batch {TRAIN_B}/GPU x {T}-sample "
                                   "segments (51 frames), YAML generator with weight norm, HiFiGAN multi-scale + "
                                   "multi-period critic (70.7 M params, host PyTorch), MR-STFT (6 resolutions) + LSGAN "
                                   "losses, RAdam; Trainer._train_step order (train_fastsvc.py:157-235)",
                       "precision_mode": "fp32 training forward/backward; no-grad forward: " + g.precision,
                       "l2": "not flushed: a step's working set (1 GB of saved activations + 283 MB critic) exceeds L2",
                       "parallelism": f"data parallel x{world}: flat fp32 gradient buckets ({gbytes >> 20} MiB generator, "
                                      f"{dbytes >> 20} MiB critic), NCCL all-reduce, clip after the reduce; the critic's "
                                      "reduce overlaps the next generator phase"},
            "e2e": {"value": sharding.aggregate_throughput(TRAIN_B * T, world, ms_e2e), "unit": UNIT,
                    "h2d_bytes_per_step": sum(t.numel() * 4 for t in host), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches * args.steps, "launches_per_step": launches, "launches_breakdown": counts,
            "allreduce_bytes_per_step": (gbytes + dbytes) if world > 1 else 0,
            "clocks": clocks, "step_breakdown": parts, "torch_eager_cuda": eager, "cpu_baseline": cpu_base,
            "roofline": None,
        }
        print(json.dumps(line))
    gpu.close()


def _train_cpu_step(nb, reps):
    """The reference training step on the host cores (oracle port of the generator with torch autograd + the same
    host critic / losses), bounded sample of `nb` utterances."""
    import torch
    import gan_host
    from oracle import fastsvc_torch as otorch
    from svcc23_fastsvc_b200 import synthetic as syn
    cores = _use_all_cores()
    params = syn.make_params(syn.YAML_CONFIG, seed=0, weight_norm=True)
    tp = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in params.items()}
    ppg, sine, lft, spk, y = [torch.from_numpy(a[:nb]) for a in _train_batch(0)]
    torch.manual_seed(0)
    D = gan_host.MultiScaleMultiPeriodCritic()
    stft = gan_host.MultiResolutionSTFTLoss(**gan_host.STFT_PARAMS)
    opt_g = torch.optim.RAdam(list(tp.values()), lr=1e-3, eps=1e-6)
    opt_d = torch.optim.RAdam(D.parameters(), lr=1e-3, eps=1e-6)

    def step():
        y_ = otorch.generator_forward(tp, ppg, sine, lft, spk, recompute=True)
        sc, mag = stft(y_, y)
        loss = sc + mag + 2.5 * gan_host.generator_adversarial_loss(D(y_))
        opt_g.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(tp.values()), 10.0)
        opt_g.step()
        with torch.no_grad():
            y_ = otorch.generator_forward(tp, ppg, sine, lft, spk, recompute=True)
        real, fake = gan_host.discriminator_adversarial_loss(D(y_.detach()), D(y))
        opt_d.zero_grad()
        (real + fake).backward()
        torch.nn.utils.clip_grad_norm_(D.parameters(), 1.0)
        opt_d.step()

    dt = _time_cpu(step, reps, 1)
    T = TRAIN_FRAMES * 160
    return {"value": nb * T / dt, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": _cpu_model(),
            "ms_per_step": dt * 1e3,
            "sample": f"{reps} training step(s) on {nb} of the {TRAIN_B} utterances ({T} samples each): reference op "
                      "sequence with torch autograd + the same critic / losses / RAdam, fp32"}


def reference_train(args):
    nb = 2
    cb = _train_cpu_step(nb, max(1, min(args.steps, 3)))
    return {
        "impl": "reference", "metric": TRAIN_METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[2]: GAN training step (bounded sample: {nb} of {TRAIN_B} utterances/step, at most 3 steps)"},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# =====================================================================================================================
# configs[4]: batched offline conversion, end to end (host records -> wav files)
# =====================================================================================================================
CONV_METRIC = "offline conversion audio samples/sec, end to end (10k x 5 s utterances sharded round-robin, host records -> PCM-16 wav files)"


def _make_utterances(indices, frames):
    """Distinct synthetic utterance records (every index its own seeds), decode_fastsvc.py's per-utterance inputs."""
    import numpy as np
    from svcc23_fastsvc_b200 import convert as cv, synthetic as syn
    utts = []
    for i in indices:
        rs = np.random.RandomState(100000 + i)
        ppg = rs.standard_normal(size=(frames, 144)).astype(np.float32)
        f0 = syn.make_f0(rs, 1, frames)[0, 0].astype(np.float64)[:, None]
        nl = (frames * 160 + 63) // 64
        lft = np.repeat(np.clip(-4.0 + 2.0 * rs.standard_normal(size=nl), -11.5, 3.0).astype(np.float32), 64)[:frames * 160]
        utts.append(cv.Utterance(f"spk{i % 4}_{i:06d}", ppg, f0, lft[:, None]))
    return utts


def run_convert(args):
    import numpy as np
    import torch
    from harana.utils.features import SignalGenerator
    from svcc23_fastsvc_b200 import convert as cv, sharding
    gpu = Gpu()
    dev, rank, world = gpu.dev, gpu.rank, gpu.world
    frames, T = CONV_FRAMES, CONV_FRAMES * 160
    g, params = _generator(dev)
    g = g.eval()
    g.precision = args.precision
    sg = SignalGenerator(sample_rate=16000, hop_size=160, sine_amp=0.1, noise_amp=0.003, signal_types=["sine"])
    n_utts = args.utts
    mine = sharding.shard_utterances(n_utts, rank, world)               # round-robin: utt_id % world == rank
    utts = _make_utterances(mine, frames)                                # the rank's share of the "dataset", in host RAM
    emb = np.random.RandomState(7).randn(1, 512).astype(np.float32)
    src, trg = np.array([5.3, 1.0]), np.array([5.6, 1.0])
    outdir = tempfile.mkdtemp(prefix=f"fsvc_convert_r{rank}_", dir=args.outdir or None)
    conv = cv.BatchConverter(g, sg, sampling_rate=16000, max_batch=args.batch, writer_threads=args.writer_threads)
    warm = utts[: min(len(utts), 3 * args.batch)]
    conv.convert_to_dir(warm, outdir, spk_emb=emb, src_stats=src, trg_stats=trg)
    conv.stats = dict(batches=0, utterances=0, samples=0, h2d_bytes=0, d2h_bytes=0)
    launches0 = None
    sampler = ClockSampler(gpu.local)
    if rank == 0:
        sampler.start()
    gpu.barrier()
    t0 = time.perf_counter()
    conv.convert_to_dir(utts, outdir, spk_emb=emb, src_stats=src, trg_stats=trg)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    sec_max = sharding.max_over_ranks(sec, device=dev)
    gpu.barrier()
    clocks = sampler.stop() if rank == 0 else None
    files = [f for f in os.listdir(outdir) if f.endswith(".wav")]
    wav_bytes = sum(os.path.getsize(os.path.join(outdir, f)) for f in files)
    ok = len(files) == len(utts) and all(os.path.getsize(os.path.join(outdir, f)) == 44 + 2 * T for f in files[:50])
    launches = g.last_launch_count()
    shutil.rmtree(outdir, ignore_errors=True)
    total_samples = n_utts * T
    if rank == 0:
        cpu_base = _convert_cpu(params, args.ref_utts) if (world == 1 or args.ref_utts) and args.ref_utts > 0 else None
        batches = conv.stats["batches"]
        line = {
            "metric": CONV_METRIC, "value": total_samples / sec_max, "unit": UNIT, "n_gpus": world,
            "steps": batches, "warmup": 3, "ms_per_step": sec_max * 1e3 / max(1, batches), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[4]: {n_utts} distinct synthetic utterances x {T} samples (5 s), round-robin over "
                                   f"{world} GPU(s), equal-length batches of {args.batch}, F0 mean transformation + sine "
                                   "excitation + generator + PCM-16 on the GPU, one PCM-16 wav file per utterance",
                       "precision_mode": g.precision, "timing": "wall clock of the whole job, max over ranks (host packing, "
                       "PCIe, kernels, wav writing)", "parallelism": f"utterance-sharded x{world}, no collectives",
                       "writer_threads": args.writer_threads},
            "wall_clock_s": sec_max, "utterances": n_utts, "seconds_of_audio": total_samples / 16000.0,
            "realtime_factor": sec_max / (total_samples / 16000.0),
            "e2e": {"value": total_samples / sec_max, "unit": UNIT,
                    "h2d_bytes_per_step": conv.stats["h2d_bytes"] // max(1, batches),
                    "d2h_bytes_per_step": conv.stats["d2h_bytes"] // max(1, batches), "ms_per_step": sec_max * 1e3 / max(1, batches)},
            "gpu_launches": (launches + 2) * batches, "launches_per_step": launches + 2,
            "rank0": dict(conv.stats, wav_files=len(files), wav_bytes=wav_bytes, files_ok=bool(ok)),
            "clocks": clocks, "cpu_baseline": cpu_base, "roofline": None,
        }
        if cpu_base:
            line["speedup_vs_cpu_reference_extrapolated"] = cpu_base["extrapolated_s"] / sec_max
        print(json.dumps(line))
    gpu.close()


def _convert_cpu(params, n_ref):
    """The reference's decode pattern on the host cores: batch 1, sequential (decode_fastsvc.py:168-198)."""
    import numpy as np
    import torch
    from oracle import fastsvc_torch as otorch, features_numpy as fo
    cores = _use_all_cores()
    tp = {k: torch.from_numpy(v) for k, v in params.items()}
    frames, T = CONV_FRAMES, CONV_FRAMES * 160
    utts = _make_utterances(range(n_ref), frames)
    emb = torch.from_numpy(np.random.RandomState(7).randn(1, 512).astype(np.float32))
    src, trg = np.array([5.3, 1.0]), np.array([5.6, 1.0])
    outdir = tempfile.mkdtemp(prefix="fsvc_convert_ref_")
    from svcc23_fastsvc_b200 import convert as cv
    t1 = time.perf_counter()
    for u in utts:
        f0 = fo.f0_convert(np.squeeze(u.f0, 1), src, trg).astype(np.float32)[None, None]
        noise = np.random.RandomState(0).randn(1, 1, T).astype(np.float32)
        s = fo.sinusoid(f0, noise, 16000, 160, 0.1, 0.003)
        with torch.no_grad():
            y = otorch.generator_forward(tp, torch.from_numpy(u.ppg.T[None].copy()), torch.from_numpy(s),
                                         torch.from_numpy(u.lft.T[None].copy()), emb, recompute=True)
        cv.write_wav(os.path.join(outdir, u.utt_id + "_gen.wav"), fo.pcm16(y.numpy().reshape(-1)), 16000)
    per = (time.perf_counter() - t1) / max(1, n_ref)
    shutil.rmtree(outdir, ignore_errors=True)
    return {"value": T / per, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": _cpu_model(),
            "seconds_per_utterance": per, "extrapolated_s": per * 10000,
            "sample": f"{n_ref} distinct 5-second utterances at batch 1, sequential, wavs written (the reference's decode "
                      "loop); extrapolated_s = the same rate for 10000 utterances"}


def reference_convert(args):
    from svcc23_fastsvc_b200 import synthetic as syn
    params = syn.make_params(syn.YAML_CONFIG, seed=0)
    cb = _convert_cpu(params, max(8, args.ref_utts))
    return {
        "impl": "reference", "metric": CONV_METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_utterance"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: batch-1 sequential decode of 5-second utterances on the host cores (bounded sample)"},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="infer", choices=["infer", "train", "convert"])
    ap.add_argument("--precision", default=os.environ.get("FSVC_MODE", "auto"))
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager-CUDA comparison leg")
    ap.add_argument("--quick", action="store_true", help="infer: timing only (no parity / roofline / CPU legs); for A/B runs")
    ap.add_argument("--utts", type=int, default=10000, help="convert: utterances in the whole job")
    ap.add_argument("--batch", type=int, default=32, help="convert: utterances per launch")
    ap.add_argument("--ref-utts", type=int, default=100, help="convert: utterances of the CPU reference leg (0 = skip)")
    ap.add_argument("--writer-threads", type=int, default=4, help="convert: wav writer threads per rank")
    ap.add_argument("--outdir", default="", help="convert: parent directory of the (temporary) wav output")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return                      # rank 0 alone runs the CPU arm
        fn = {"infer": reference_infer, "train": reference_train, "convert": reference_convert}[args.config]
        print(json.dumps(fn(args)))
    else:
        {"infer": run_infer, "train": run_train, "convert": run_convert}[args.config](args)


if __name__ == "__main__":
    main()
