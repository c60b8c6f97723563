#!/bin/bash
# Copy the judged summaries of the last GPU round from gpurun_out/ (scratch) into profiles/ (tracked).
tag=${1:-r01}
out=profiles
mkdir -p $out
cp gpurun_out/bench.json $out/${tag}_bench.json
cp gpurun_out/launches.csv $out/${tag}_launches.csv
ncu -i gpurun_out/run_full.ncu-rep --page details > $out/${tag}_ncu_details.txt 2>/dev/null
ncu -i gpurun_out/run_full.ncu-rep --page raw --csv > /tmp/_raw.csv 2>/dev/null
ncu -i gpurun_out/run_full.ncu-rep --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
python - <<PY > $out/${tag}_ncu_key_metrics.txt
import csv
rows=list(csv.reader(open('/tmp/_raw.csv')))
d=dict(zip(rows[0],rows[-1]))
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','dram__bytes_read.sum','dram__bytes_write.sum',
 'smsp__sass_inst_executed_op_shared_ld.sum','smsp__sass_inst_executed_op_shared_st.sum','smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum',
 'smsp__sass_inst_executed_op_global_ld.sum','smsp__sass_inst_executed_op_global_st.sum','smsp__inst_executed_op_branch.sum','smsp__inst_executed_op_tma_ld.sum',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__warps_active.avg.per_cycle_active','launch__registers_per_thread','launch__block_size','launch__grid_size','launch__shared_mem_per_block_dynamic']
units=dict(zip(rows[0],rows[1])) if len(rows)>2 else {}
print("kernel:", d.get('Kernel Name'))
for k in keys:
    if k in d: print(f"{k:70s} {d[k]} {units.get(k,'')}")
print("-- warp stall cycles per issued instruction")
out=[]
for k in rows[0]:
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
        out.append((float(d[k].replace(',','')),k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
for v,k in sorted(out,reverse=True): print(f"  {k:24s} {v:.2f}")
PY
python tools/ncu_funcs.py /tmp/_src.csv HEAD > $out/${tag}_ncu_functions.txt 2>/dev/null
python - <<PY
import json,csv
rows=list(csv.reader(open('/tmp/_raw.csv')))
d=dict(zip(rows[0],rows[-1])); u=dict(zip(rows[0],rows[1]))
def val(k):
    v=float(d[k].replace(',','')); mult={'Mbyte':1e6,'Kbyte':1e3,'Gbyte':1e9,'byte':1}.get(u.get(k,''),1)
    return v*mult
t={"kernel":d.get('Kernel Name','').split('(')[0],"source":"profiles/${tag}_ncu_details.txt (ncu --set full, one launch, 4096 envs x 17 substeps)",
   "dram_bytes_read":val('dram__bytes_read.sum'),"dram_bytes_write":val('dram__bytes_write.sum'),
   "inst_executed":float(d['smsp__inst_executed.sum']),"ipc_active":float(d['sm__inst_executed.avg.per_cycle_active']),
   "fp64_pipe_pct":float(d['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']),"lsu_pipe_pct":float(d['sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']),
   "kernel_ms_under_ncu":float(d['gpu__time_duration.sum'])*({'us':1e-3,'ms':1,'ns':1e-6}.get(u.get('gpu__time_duration.sum','ms'),1)),
   "envs":4096,"substeps":17}
json.dump(t,open('$out/${tag}_traffic.json','w'),indent=1)
PY
ls -la $out
