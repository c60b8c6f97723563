#!/bin/bash
# compute-sanitizer passes over tools/sanitize_run.py; summaries land in gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" gpurun_out/sanitize_$tool.log | tail -5
done
RCSB_WARPS=4 RCSB_BAR_GROUPS=2 SANITIZE_ROUNDS=900 timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python tools/sanitize_run.py > gpurun_out/sanitize_synccheck_rounds.log 2>&1
echo "== synccheck multi-round rc=$?"; grep -E "ERROR SUMMARY|done" gpurun_out/sanitize_synccheck_rounds.log | tail -5
