#!/bin/bash
# ncu --set full with source-level sampling of a range of conv_tc3 launches of the third forward of
# kernel_profile.py --once (38 per forward: l1 76-81, l2 82-87, l3 88-93, s0 94-98, s1 99-103, s2 104-108, s3 109-113)
# usage: bash tools/gpu_ncu_src.sh NAME SKIP COUNT [NAME SKIP COUNT ...]
while [ $# -ge 3 ]; do
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tc3 --launch-skip $2 --launch-count $3 -f -o gpurun_out/src_$1 python tools/kernel_profile.py --once > /dev/null 2>&1
  shift 3
done
ls -la gpurun_out/*.ncu-rep
