for l in 0x210 0x218 0x230 0x250 0x310 0x200; do
  v=$(RCSB_LOCKSTEP=$l python tools/bench_part.py c4 8192 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(int(d['value']))")
  echo "LOCKSTEP=$l  c4 $v"
done
for l in 0x010 0x210; do
  c3=$(RCSB_LOCKSTEP=$l python tools/bench_c3.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(int(d['env_steps_per_s']))")
  c2=$(RCSB_LOCKSTEP=$l python bench.py --steps 30 --warmup 5 --cpu-seconds 0.2 --no-sweep 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(int(d['value']))")
  c2b=$(RCSB_LOCKSTEP=$l python tools/bench_part.py c2 16384 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(int(d['value']))")
  echo "LOCKSTEP=$l  c3 $c3  c2 $c2  c2@16k $c2b"
done
