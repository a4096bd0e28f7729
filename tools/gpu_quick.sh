#!/bin/bash
# quick GPU round: parity tests, timeline, per-kernel profile, instruction counts of the last stage, short benches
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
bash tools/gpu_timeline.sh
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc3 --csv --log-file gpurun_out/inst.csv python tools/kernel_profile.py --once > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/inst.csv')) if len(r)>10]
h=rows[0]; ni=h.index("Metric Name"); vi=h.index("Metric Value"); ki=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault(r[ki],{})[r[ni]]=float(r[vi].replace(",",""))
ids=sorted(d,key=int)
print("last 5 conv_tc3 launches (inst M, us):", [(round(d[i]["smsp__inst_executed.sum"]/1e6,2), round(d[i]["gpu__time_duration.sum"]/1e3,1)) for i in ids[-5:]])
print("total inst M", round(sum(x["smsp__inst_executed.sum"] for x in d.values())/1e6,1), "n", len(ids))
PY
for r in 1 2; do
timeout 200 python bench.py --steps 40 --warmup 8 --no-eager --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done
