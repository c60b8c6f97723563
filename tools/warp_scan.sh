for w in 1 2 4 6 8 12; do
  n=$((148*w*4))
  RCSB_WARPS=$w timeout 200 python bench.py --steps 12 --warmup 3 --envs $n --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('warps',d['config']['warps_per_cta'],'envs',d['config']['envs_per_gpu'],'env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done
