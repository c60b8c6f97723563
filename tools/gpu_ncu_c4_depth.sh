mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k rcsb_k_run_xarm7_tabletop -s 5 -c 1 -o gpurun_out/run_c4 -f python tools/bench_part.py c4 4096 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
bash tools/export_profiles.sh gpurun_out/run_c4.ncu-rep gpurun_out/r02_c4_ncu 69632
timeout 900 ncu --set full --clock-control none --import-source on -k rcsb_k_depth -s 3 -c 1 -o gpurun_out/run_depth -f python tools/bench_part.py depth 4096 > gpurun_out/ncu_depth.log 2>&1
tail -2 gpurun_out/ncu_depth.log
bash tools/export_profiles.sh gpurun_out/run_depth.ncu-rep gpurun_out/r02_depth_ncu 4096
rm -f gpurun_out/*.ncu-rep
