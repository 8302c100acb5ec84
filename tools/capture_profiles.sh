#!/bin/bash
# round-2 profile set on ONE GPU (run under gpurun): bench line, launch list, ncu --set full captures; outputs in gpurun_out/
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fast"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_config_e_n1.json 2> gpurun_out/r2_bench_config_e_n1.err
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter_const -s 150 -c 2 -o gpurun_out/r2_step_filter $B > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_(shade|backward|narrow_queue)" -s 12 -c 4 -o gpurun_out/r2_step_rest $B > gpurun_out/ncu_f2.log 2>&1
python tools/run_config.py D 5 > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_intersect_batch -s 3 -c 1 -o gpurun_out/r2_isect_D python tools/run_config.py D 5 > gpurun_out/ncu_d.log 2>&1
python tools/run_config.py B 5 > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_intersect_batch -s 3 -c 1 -o gpurun_out/r2_isect_B python tools/run_config.py B 5 > gpurun_out/ncu_b.log 2>&1
python tools/bench_configs.py > gpurun_out/r2_bench_configs.jsonl 2>&1
python bench.py --workload config_d --steps 20 --warmup 3 > gpurun_out/r2_bench_config_d_n1.json 2> gpurun_out/r2_bench_config_d_n1.err
python tools/bench_along_ray.py > gpurun_out/r2_bench_along_ray.log 2>&1
tail -n 2 gpurun_out/plain.log; tail -n 2 gpurun_out/plain_d.log
