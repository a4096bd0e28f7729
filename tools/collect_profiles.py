"""Copy the final measurement pass (tools/gpu_final.sh -> gpurun_out/) into profiles/ and regenerate the derived files:

    r2_ncu_summary.txt   tools/ncu_summary.py over `ncu --page raw --csv` of the two --set full captures
    r2_traffic.json      dram__bytes_read.sum + dram__bytes_write.sum per launch (read by bench.py: roofline.traffic)
    r2_sass_opcodes.txt  cuobjdump -sass opcode counts per kernel family of the in-tree libfsvc.so

Run here (no GPU needed): python tools/collect_profiles.py"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
COPY = {"r2_bench_final.json": "r2_bench_final.json", "r2_bench_reference_final.json": "r2_bench_reference_final.json",
        "r2_train_final.json": "r2_train_final.json", "r2_convert_final.json": "r2_convert_final.json",
        "r2_kernels_final.txt": "r2_kernels_events_final.txt", "r2_launches.csv": "r2_launches.csv",
        "r2_timeline.txt": "r2_timeline_conv_tc3.txt", "r2_tests_final.txt": "r2_tests_final.txt"}


def last_json_line(path):
    lines = [l for l in open(path).read().splitlines() if l.strip().startswith("{")]
    return lines[-1] + "\n" if lines else None


for src, dst in COPY.items():
    s = os.path.join(G, src)
    if not os.path.exists(s):
        print("missing", src)
        continue
    if src.endswith(".json"):
        line = last_json_line(s)
        if line is None:
            print("no JSON line in", src)
            continue
        open(os.path.join(P, dst), "w").write(line)
    elif src == "r2_kernels_final.txt":
        txt = open(s).read().splitlines()
        open(os.path.join(P, dst), "w").write("\n".join(l for l in txt if not l.startswith("plan ")) + "\n")
        plan = [l for l in txt if l.startswith("plan ")]
        if plan:
            open(os.path.join(P, "r2_launch_plan.txt"), "w").write("\n".join(plan) + "\n")
    else:
        shutil.copy(s, os.path.join(P, dst))
    print("copied", dst)


def raw_csv(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    path = "/tmp/" + os.path.basename(rep) + ".csv"
    open(path, "w").write(out)
    return path


reps = [os.path.join(G, n) for n in ("r2_prof_l0.ncu-rep", "r2_prof_s3.ncu-rep")]
if all(os.path.exists(r) for r in reps):
    csvs = [raw_csv(r) for r in reps]
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py")] + csvs, capture_output=True,
                          text=True).stdout
    head = ("# ncu --set full --clock-control none, round-2 final build (tools/gpu_final.sh, tools/ncu_summary.py over "
            "--page raw --csv); B=32 x 16000 samples\n# first file: level0_fused_kernel; second: the last stage's conv_tc3 "
            "launches in order: conv_first, residual+up_film, d3_film, d9_film, d27_skip+last\n")
    open(os.path.join(P, "r2_ncu_summary.txt"), "w").write(head + summ)
    # DRAM traffic per launch
    traffic = {}
    names = [["l0.fused_level"], ["s3.conv_first", "s3.residual+up_film", "s3.d3_film", "s3.d9_film", "s3.d27_skip+last"]]
    for path, labels in zip(csvs, names):
        rows = list(csv.reader(open(path)))
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        h, units = rows[hdr], rows[hdr + 1]
        ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for lab, r in zip(labels, rows[hdr + 2:]):
            traffic[lab] = int(float(r[ir].replace(",", "")) * mult[units[ir]] + float(r[iw].replace(",", "")) * mult[units[iw]])
    traffic["_source"] = ("ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                          "round-2 final build (profiles/r2_ncu_summary.txt)")
    json.dump(traffic, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
    print("ncu summary + traffic written")
else:
    print("ncu captures missing: summary / traffic unchanged")

# SASS opcode counts per kernel family
so = os.path.join(ROOT, "svcc23_fastsvc_b200", "libfsvc.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fam, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        key = next((k for k in ("conv_tc3_kernel", "level0_fused_kernel", "conv1d_f32_kernel", "conv_wgrad_kernel") if k in name),
                   "other")
        cur = fam.setdefault(key, Counter())
        cur["functions"] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur["instructions"] += 1
        cur[m.group(1)] += 1
keep = ("UTCHMMA", "LDTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "F2FP", "LDS", "STS", "LDG",
        "STG", "RED", "ATOMG", "BAR", "ELECT", "UTCBAR", "UTCATOMSWS", "FENCE", "ACQBULK")
with open(os.path.join(P, "r2_sass_opcodes.txt"), "w") as f:
    f.write("# cuobjdump -sass svcc23_fastsvc_b200/libfsvc.so (sm_100a), instruction counts by kernel family, round-2 final build\n"
            "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA engine, non-tensor), UTMALDG = "
            "cp.async.bulk.tensor,\n# LDGSTS = cp.async (per-lane), SYNCS = mbarrier ops, FFMA2 / FADD2 / FMUL2 = packed fp32 "
            "(two operations per instruction), RED / ATOMG = global reductions\n")
    for k, c in sorted(fam.items(), key=lambda kv: -kv[1]["instructions"]):
        f.write(f"{k}: instructions={c['instructions']}, functions={c['functions']}, " +
                ", ".join(f"{o}={c[o]}" for o in keep if c[o]) + "\n")
print("sass opcode counts written")
