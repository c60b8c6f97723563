for g in 1 2 4 7; do
  echo "BAR_GROUPS=$g"
  RCSB_BAR_GROUPS=$g python bench.py --steps 30 --warmup 5 --cpu-seconds 0.2 --no-sweep 2>/dev/null | python tools/bench_summary.py /dev/stdin | grep headline | cut -c1-110
done
for g in 1 2 3; do
  echo "C3 BAR_GROUPS=$g"
  RCSB_BAR_GROUPS=$g python tools/bench_c3.py 2>&1 | tail -2
done
