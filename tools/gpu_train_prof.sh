#!/bin/bash
# launch list of two native training steps + ncu --set full (source) of the heaviest fp32 conv / wgrad launches
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_probe.py --once --no-eager > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r)>10]
h=rows[0]; ni=h.index("Kernel Name"); vi=h.index("Metric Value"); ii=h.index("ID")
ks=[(float(r[vi].replace(",","")), r[ii], r[ni][:60]) for r in rows[1:] if "fsvc::" in r[ni]]
ks.sort(reverse=True)
for t,i,n in ks[:12]: print(i, round(t/1e3,1), n)
PY
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv1d_f32_kernel --launch-skip 200 --launch-count 40 -f -o gpurun_out/src_f32 python tools/train_probe.py --once --no-eager > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_wgrad_kernel --launch-skip 60 --launch-count 20 -f -o gpurun_out/src_wgrad python tools/train_probe.py --once --no-eager > /dev/null 2>&1
ls -la gpurun_out/src_f32.ncu-rep gpurun_out/src_wgrad.ncu-rep
