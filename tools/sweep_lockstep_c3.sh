for l in 0 2 0x210 0x208; do
  echo "C3 LOCKSTEP=$l"
  RCSB_LOCKSTEP=$l python tools/bench_c3.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['env_steps_per_s'])"
done
