#!/bin/bash
# compute-sanitizer memcheck + racecheck over __graft_entry__.smoke() -> gpurun_out/r2_sanitizer_final.txt
O=gpurun_out/r2_sanitizer_final.txt
echo "# compute-sanitizer over __graft_entry__.smoke() (fp32 + auto forward, B=2 x 10 frames, and the excitation kernel); round-2 final build" > $O
if [ "$1" != "race" ]; then
echo "## memcheck" >> $O
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^=========     \|^$" | tail -12 >> $O
fi
echo "## racecheck" >> $O
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_full.txt 2>&1
grep -v "^=========     \|^$" gpurun_out/racecheck_full.txt | tail -12 >> $O
head -c 30000 gpurun_out/racecheck_full.txt > gpurun_out/racecheck_head.txt
tail -5 $O
