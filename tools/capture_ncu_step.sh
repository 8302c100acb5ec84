#!/bin/bash
# the two ncu --set full captures of the bench step that profiles/ncu_traffic.json is made from (one GPU, under gpurun)
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fast"
$B > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_filter_const -s 150 -c 2 -o gpurun_out/r2_step_filter -f $B > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_(shade|backward|narrow_queue)" -s 12 -c 4 -o gpurun_out/r2_step_rest -f $B > gpurun_out/ncu_f2.log 2>&1
tail -n 1 gpurun_out/plain.log | cut -c1-200
