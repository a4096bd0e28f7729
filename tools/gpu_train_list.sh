#!/bin/bash
# kernel-time breakdown of two native training steps (ncu launch list, serialised times)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_probe.py --once --no-eager > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r)>10]
h=rows[0]; ni=h.index("Kernel Name"); vi=h.index("Metric Value")
d=defaultdict(float); n=defaultdict(int)
for r in rows[1:]:
    k=r[ni].split("(")[0].replace("void ","")[:60]
    d[k]+=float(r[vi].replace(",","")); n[k]+=1
tot=sum(d.values())
for k in sorted(d,key=lambda k:-d[k])[:14]: print(f"{k:62s} {n[k]:5d} {d[k]/1e6:8.2f} ms {100*d[k]/tot:5.1f}%")
print("total ms", tot/1e6)
PY
