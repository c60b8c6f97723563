for l in 0x010 0x090 0x210 0x290 0x0d0 0x030 0x2d0 1; do
  v=$(RCSB_LOCKSTEP=$l python tools/bench_part.py c4 8192 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(int(d['value']))")
  echo "LOCKSTEP=$l  c4 $v"
done
