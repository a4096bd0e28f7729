#!/bin/bash
# In-kernel timelines of one forward on the GPU box (needs libfsvc_tl.so: python -m svcc23_fastsvc_b200.build --timeline)
FSVC_LIB=$PWD/svcc23_fastsvc_b200/libfsvc_tl.so timeout 200 python tools/timeline.py > gpurun_out/timeline.txt 2>&1
tail -5 gpurun_out/timeline.txt
timeout 120 env FSVC_DEBUG_PLAN=1 python tools/kernel_profile.py > gpurun_out/kp_now.txt 2>&1
grep sum_ms gpurun_out/kp_now.txt
