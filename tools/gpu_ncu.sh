#!/bin/bash
# one full ncu capture of the hot kernel (5th rcsb_k_run launch of a short bench run)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k ${KERNEL:-rcsb_k_run_fr3_reduced} -s ${SKIP:-4} -c 1 -o gpurun_out/run_full -f python bench.py --steps 4 --warmup 3 --cpu-seconds 0.2 --envs ${ENVS:-4096} > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
