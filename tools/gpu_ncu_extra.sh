#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k rcsb_k_ik -s 1 -c 1 -o gpurun_out/run_ik -f python tools/bench_ik.py 4096 > gpurun_out/ncu_ik.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k rcsb_k_run_fr3_pickup -s 6 -c 1 -o gpurun_out/run_c3 -f python tools/bench_c3.py 4096 4 > gpurun_out/ncu_c3.log 2>&1
tail -1 gpurun_out/ncu_ik.log; tail -1 gpurun_out/ncu_c3.log
