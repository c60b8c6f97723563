#!/bin/bash
# One GPU round: parity tests, bench, stage timing, ncu launch list + one full capture of the hot kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/stage_timing.py 4096 > gpurun_out/stage_timing.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 --cpu-seconds 0.2 > gpurun_out/ncu_bench.log 2>&1
bash tools/gpu_ncu.sh
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/stage_timing.txt
