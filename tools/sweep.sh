for m in 0x220 0x020 0x010 0x040 0x080 0x100 0x008 0x120 0x030 0x060 0x0a0; do
RCSB_LOCKSTEP=$m python bench.py --steps 30 --warmup 4 --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('mask $m env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done
