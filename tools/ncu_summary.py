"""Summarise `ncu --page raw --csv` exports: one line per profiled launch with the metrics DESIGN.md quotes."""
import csv, sys
KEYS = [("gpu__time_duration.sum", "dur_us"), ("dram__bytes_read.sum", "rd_MB"), ("dram__bytes_write.sum", "wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__shared_mem_per_block_dynamic", "smem"), ("smsp__cycles_active.avg", "cyc"),
        ("sm__inst_executed.sum", "inst"), ("launch__occupancy_limit_shared_mem", "lim_smem"),
        ("launch__occupancy_limit_registers", "lim_reg")]
def conv(v, unit, name):
    try: x = float(v.replace(",", ""))
    except ValueError: return v
    if name in ("rd_MB", "wr_MB"):
        f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6); return f"{x*f:.1f}"
    if name == "dur_us":
        f = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1e-3); return f"{x*f:.1f}"
    return f"{x:.1f}" if x != int(x) else str(int(x))
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}
    print(path)
    print("  " + " ".join(f"{k[1]:>8s}" for k in KEYS))
    for r in rows[hdr + 2:]:
        if len(r) < len(names): continue
        print("  " + " ".join(f"{conv(r[idx[k[0]]], units[idx[k[0]]], k[1]) if k[0] in idx else '-':>8s}" for k in KEYS))
