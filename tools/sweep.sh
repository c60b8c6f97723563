for v in fixed generic; do
RCSB_VARIANT=$v python bench.py --steps 20 --warmup 4 --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('variant $v warps',d['config']['warps_per_cta'],'env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done
