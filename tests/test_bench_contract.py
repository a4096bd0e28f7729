"""bench.py's reference arm runs on the host CPU (no GPU needed): one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

from conftest import REPO


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")     # what torchrun exports: the arm must still use every core
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=REPO)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("audio samples/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in d["config"]["workload"] and "bounded" not in d["config"]["workload"]   # the whole batch of 32
    assert d["cpu_baseline"]["cores"] == os.cpu_count() and d["cpu_baseline"]["cpu_model"]
    assert abs(d["value"] - 32 * 16000 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]


def test_reference_arm_train_and_convert_configs():
    for cfg, extra in (("train", []), ("convert", ["--ref-utts", "2"])):
        res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--config", cfg,
                              "--steps", "1", "--warmup", "1"] + extra, capture_output=True, text=True, timeout=900,
                             cwd=REPO)
        assert res.returncode == 0, res.stderr[-2000:]
        d = json.loads([l for l in res.stdout.splitlines() if l.strip()][-1])
        assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
        assert d["e2e"]["value"] == d["value"] and f"configs[{2 if cfg == 'train' else 4}]" in d["config"]["workload"]


def test_reference_arm_is_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=REPO)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_committed_bench_lines_carry_the_contract_keys():
    """The lines under profiles/ are what DESIGN.md quotes: each must be one JSON object with the contract's keys and
    numbers that follow from each other (value = samples per step / ms_per_step, roofline.frac = achieved / peak)."""
    import json
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
    d = json.loads(open(os.path.join(root, "r2_bench_final.json")).read())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    samples_per_step = 32 * 16000
    assert abs(d["value"] - samples_per_step / (d["ms_per_step"] * 1e-3)) <= 1e-3 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-6
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # forward / training: fixed work per GPU (weak); offline conversion: one fixed job of 10 000 utterances (strong)
    for name, n, scaling in (("r2_bench_2gpu.json", 2, "weak"), ("r2_bench_8gpu.json", 8, "weak"),
                             ("r2_train_final.json", 1, "weak"), ("r2_train_2gpu.json", 2, "weak"),
                             ("r2_train_8gpu.json", 8, "weak"), ("r2_convert_final.json", 1, "strong"),
                             ("r2_convert_8gpu_10000utts.json", 8, "strong")):
        x = json.loads(open(os.path.join(root, name)).read())
        assert x["n_gpus"] == n and x["value"] > 0 and x["scaling"] == scaling, name
    ref = json.loads(open(os.path.join(root, "r2_bench_reference_final.json")).read())
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
