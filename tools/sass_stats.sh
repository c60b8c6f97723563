#!/bin/bash
# opcode histogram per kernel of the built library
LIB=${1:-/root/repo/robot-control-stack_b200/csrc/librcsb.so}
cuobjdump -sass $LIB > /tmp/_s.sass
grep -n "Function :" /tmp/_s.sass | while IFS=: read ln rest; do echo "$ln $rest"; done > /tmp/_s.fn
total=$(wc -l < /tmp/_s.sass)
awk '{print $1}' /tmp/_s.fn > /tmp/_s.ln; echo $total >> /tmp/_s.ln
i=0
while read ln rest; do
  i=$((i+1)); end=$(sed -n "$((i+1))p" /tmp/_s.ln)
  name=$(echo $rest | sed 's/.*Function : //')
  sed -n "${ln},${end}p" /tmp/_s.sass | grep -E "^\s+/\*[0-9a-f]{4,}\*/" > /tmp/_s.ins
  echo "== $name: $(wc -l < /tmp/_s.ins) instructions"
  awk '{for(i=2;i<=NF;i++) if($i !~ /^@/){print $i; break}}' /tmp/_s.ins | sed 's/;//; s/\..*//' | sort | uniq -c | sort -rn | head -${2:-16} | tr '\n' ' '; echo
done < /tmp/_s.fn
