#!/bin/bash
# Turn one ncu --set full capture into the small text summaries kept under profiles/ (run where the .ncu-rep is, e.g. on
# the GPU box right after the capture: the reports themselves are too large to travel back).
#   tools/export_profiles.sh <report.ncu-rep> <out prefix, e.g. gpurun_out/r02_ncu> <work units per launch for the per-unit column>
rep=$1; out=$2; units=${3:-69632}
ncu -i $rep --page details > ${out}_details.txt 2>/dev/null
ncu -i $rep --page raw --csv > /tmp/_raw.csv 2>/dev/null
ncu -i $rep --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
python - <<PY > ${out}_key_metrics.txt
import csv
rows=list(csv.reader(open('/tmp/_raw.csv')))
d=dict(zip(rows[0],rows[-1]))
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','smsp__thread_inst_executed_per_inst_executed.ratio',
 'dram__bytes_read.sum','dram__bytes_write.sum',
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
python tools/ncu_funcs.py /tmp/_src.csv HEAD $units > ${out}_functions.txt 2>/dev/null
python tools/ncu_lines.py /tmp/_src.csv samp 40 > ${out}_hot_lines.txt 2>/dev/null
python - <<PY
import json,csv
rows=list(csv.reader(open('/tmp/_raw.csv')))
d=dict(zip(rows[0],rows[-1])); u=dict(zip(rows[0],rows[1]))
def val(k):
    v=float(d[k].replace(',','')); mult={'Mbyte':1e6,'Kbyte':1e3,'Gbyte':1e9,'byte':1}.get(u.get(k,''),1)
    return v*mult
t={"kernel":d.get('Kernel Name','').split('(')[0],"source":"${out}_details.txt".replace("gpurun_out/","profiles/")+" (ncu --set full --clock-control none, one launch)",
   "dram_bytes_read":val('dram__bytes_read.sum'),"dram_bytes_write":val('dram__bytes_write.sum'),
   "inst_executed":float(d['smsp__inst_executed.sum']),"ipc_active":float(d['sm__inst_executed.avg.per_cycle_active']),
   "fp64_pipe_pct":float(d['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']),"lsu_pipe_pct":float(d['sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']),
   "kernel_ms_under_ncu":float(d['gpu__time_duration.sum'])*({'us':1e-3,'ms':1,'ns':1e-6}.get(u.get('gpu__time_duration.sum','ms'),1))}
json.dump(t,open("${out}_traffic.json",'w'),indent=1)
PY
rm -f $rep
