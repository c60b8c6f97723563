for v in 0 2 4 6 10 16; do for m in 0x2a5 0x3ff; do
RCSB_COLL_VOTE=$v RCSB_LOCKSTEP=$m timeout 120 python bench.py --steps 20 --warmup 4 --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('vote $v mask $m env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done; done
