#!/bin/bash
# quick GPU check: parity tests + bench (+ optional stage timing)
mkdir -p gpurun_out
if [ "$SKIPTESTS" != "1" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
fi
timeout 900 python bench.py --steps ${STEPS:-30} --warmup 5 --cpu-seconds ${CPUSEC:-1} $BENCHARGS > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python tools/bench_summary.py gpurun_out/bench.json
if [ -f robot-control-stack_b200/csrc/librcsb_prof.so ] && [ "$1" == "stages" ]; then timeout 300 python tools/stage_timing.py 4096 | tee gpurun_out/stage_timing.txt; fi
