#!/bin/bash
# One GPU round: parity tests, bench, ncu launch list + full captures of the hot kernels, summarised ON THE BOX into
# gpurun_out/<tag>_* (copy them to profiles/ afterwards). Round tag = $1, default r02.
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc >> gpurun_out/${tag}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 12 --warmup 3 --cpu-seconds 0.2 --no-sweep > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k rcsb_k_run_fr3_reduced -s 9 -c 1 -o gpurun_out/run_full -f python bench.py --steps 6 --warmup 4 --cpu-seconds 0.2 --no-sweep > gpurun_out/ncu_full.log 2>&1
bash tools/export_profiles.sh gpurun_out/run_full.ncu-rep gpurun_out/${tag}_ncu 69632
timeout 900 ncu --set full --clock-control none --import-source on -k rcsb_k_run_fr3_pickup -s 6 -c 1 -o gpurun_out/run_c3 -f python tools/bench_c3.py 4096 4 > gpurun_out/ncu_c3.log 2>&1
bash tools/export_profiles.sh gpurun_out/run_c3.ncu-rep gpurun_out/${tag}_c3_ncu 69632
timeout 900 ncu --set full --clock-control none --import-source on -k rcsb_k_ik8 -s 1 -c 1 -o gpurun_out/run_ik8 -f python tools/bench_ik.py 4096 > gpurun_out/ncu_ik8.log 2>&1
bash tools/export_profiles.sh gpurun_out/run_ik8.ncu-rep gpurun_out/${tag}_ik8_ncu 4096
timeout 900 ncu --set full --clock-control none --import-source on -k rcsb_k_depth -s 3 -c 1 -o gpurun_out/run_depth -f python tools/bench_part.py depth 4096 > gpurun_out/ncu_depth.log 2>&1
bash tools/export_profiles.sh gpurun_out/run_depth.ncu-rep gpurun_out/${tag}_depth_ncu 4096
timeout 900 ncu --set full --clock-control none --import-source on -k rcsb_k_run_xarm7_tabletop -s 5 -c 1 -o gpurun_out/run_c4 -f python tools/bench_part.py c4 4096 > gpurun_out/ncu_c4.log 2>&1
bash tools/export_profiles.sh gpurun_out/run_c4.ncu-rep gpurun_out/${tag}_c4_ncu 69632
rm -f gpurun_out/*.ncu-rep
for l in 8 1; do echo lanes=$l; RCSB_IK_LANES=$l python tools/bench_ik.py 4096 16384 65536; done > gpurun_out/${tag}_ik_timing.txt 2>&1
rm -f gpurun_out/*.log.tmp
tail -3 gpurun_out/${tag}_pytest_gpu.log; python tools/bench_summary.py gpurun_out/${tag}_bench.json; cat gpurun_out/${tag}_ik_timing.txt; ls -la gpurun_out
