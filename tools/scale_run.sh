#!/bin/bash
# multi-GPU measurement set for N ranks: bench (config E, config D eager + graph), dist_batch_check, parity worker
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29701 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_scale_e_n$N.json 2> gpurun_out/r2_scale_e_n$N.err
$TR --master-port 29702 bench.py --gpus $N --workload config_d --steps 20 --warmup 3 > gpurun_out/r2_scale_d_n$N.json 2> gpurun_out/r2_scale_d_n$N.err
$TR --master-port 29703 bench.py --gpus $N --workload config_d --steps 20 --warmup 3 --graph > gpurun_out/r2_scale_d_graph_n$N.json 2> gpurun_out/r2_scale_d_graph_n$N.err
$TR --master-port 29704 tools/dist_batch_check.py > gpurun_out/r2_dist_batch_n$N.log 2>&1
timeout 300 $TR --master-port 29705 tests/dist_worker.py > gpurun_out/r2_dist_worker_n$N.log 2>&1
for f in gpurun_out/r2_scale_e_n$N gpurun_out/r2_scale_d_n$N gpurun_out/r2_scale_d_graph_n$N; do tail -c 400 $f.json; echo; tail -c 300 $f.err | grep -i "error\|Traceback" ; done
tail -1 gpurun_out/r2_dist_batch_n$N.log; grep "^{" gpurun_out/r2_dist_worker_n$N.log | tail -1
