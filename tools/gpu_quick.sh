#!/bin/bash
# quick GPU check: parity tests + bench (+ optional stage timing)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps ${STEPS:-30} --warmup 5 --cpu-seconds ${CPUSEC:-1} > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').readline())
print('env-steps/s %.0f  kernel_ms %.3f  e2e %.0f  warps %s'%(d['value'],d['roofline']['kernel_ms'],d['e2e']['value'],d['config']['warps_per_cta']))
PY
if [ -f robot-control-stack_b200/csrc/librcsb_prof.so ] && [ "$1" == "stages" ]; then timeout 300 python tools/stage_timing.py 4096 | tee gpurun_out/stage_timing.txt; fi
