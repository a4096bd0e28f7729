#!/bin/bash
# Final measurement pass on the GPU box; everything lands in gpurun_out/ (tools/collect_profiles.py copies it to
# profiles/).  Needs libfsvc.so and, for the timeline, libfsvc_tl.so (python -m svcc23_fastsvc_b200.build [--timeline]).
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r2_tests_final.txt; cat $O/r2_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err; tail -c 400 $O/r2_bench_final.json
timeout 600 python bench.py --impl reference > $O/r2_bench_reference_final.json 2>/dev/null
timeout 900 python bench.py --config train > $O/r2_train_final.json 2>/dev/null
timeout 900 python bench.py --config convert > $O/r2_convert_final.json 2>/dev/null
timeout 200 env FSVC_DEBUG_PLAN=1 python tools/kernel_profile.py > $O/r2_kernels_final.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-eager --quick > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:level0_fused --launch-skip 2 --launch-count 1 -f -o $O/r2_prof_l0 python tools/kernel_profile.py --once > /dev/null 2>&1
# conv_tc3 launches of the third forward of kernel_profile.py --once: 38 per forward, the last stage = the last 5
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tc3 --launch-skip 109 --launch-count 5 -f -o $O/r2_prof_s3 python tools/kernel_profile.py --once > /dev/null 2>&1
if [ -f svcc23_fastsvc_b200/libfsvc_tl.so ]; then
  FSVC_LIB=$PWD/svcc23_fastsvc_b200/libfsvc_tl.so timeout 200 python tools/timeline.py > $O/r2_timeline.txt 2>&1
fi
ls -la $O/r2_*final* $O/r2_prof_*.ncu-rep $O/r2_launches.csv
