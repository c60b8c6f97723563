# barrier-mask sweep (bit i = CTA barrier before stage i of the physics step, bit 9 = at its end)
for l in 0x010 0x090 0x110 0x030 0x080 0x020 0x008 0x040; do
  c3=$(RCSB_LOCKSTEP=$l python tools/bench_c3.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(int(d['env_steps_per_s']))")
  c2=$(RCSB_LOCKSTEP=$l python bench.py --steps 30 --warmup 5 --cpu-seconds 0.2 --no-sweep 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(int(d['value']))")
  echo "LOCKSTEP=$l  c3 $c3  c2 $c2"
done
