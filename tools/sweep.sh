for g in 1 2 4 7 14; do for m in 0x3ff 0x2a5; do
RCSB_BAR_GROUPS=$g RCSB_LOCKSTEP=$m python bench.py --steps 20 --warmup 4 --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('groups $g mask $m env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done; done
