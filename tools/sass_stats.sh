#!/bin/bash
# opcode / memory-space histogram of rcsb_k_run in the built library
LIB=${1:-/root/repo/robot-control-stack_b200/csrc/librcsb.so}
cuobjdump -sass $LIB > /tmp/_s.sass
start=$(grep -n "Function : .*rcsb_k_run" /tmp/_s.sass | cut -d: -f1)
awk -v s=$start 'NR>s' /tmp/_s.sass | grep -E "^\s+/\*[0-9a-f]{4,}\*/" > /tmp/_s.ins
echo "instructions: $(wc -l < /tmp/_s.ins)"
awk '{for(i=2;i<=NF;i++) if($i !~ /^@/){print $i; break}}' /tmp/_s.ins | sed 's/;//; s/\..*//' | sort | uniq -c | sort -rn | head -${2:-24} | tr '\n' ' '; echo
grep "registers\|spill" $(dirname $LIB)/build.log | grep -A1 -B1 "rcsb_k_run" | tail -3
