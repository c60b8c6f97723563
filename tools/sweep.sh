for cfg in "1 28" "2 28" "0 28"; do set -- $cfg
RCSB_LOCKSTEP=$1 RCSB_WARPS=$2 python bench.py --steps 20 --warmup 4 --cpu-seconds 0.1 --envs ${ENVS:-4096} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('lockstep $1 warps',d['config']['warps_per_cta'],'env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done
for n in 16384 65536; do python bench.py --steps 10 --warmup 3 --cpu-seconds 0.1 --envs $n 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('envs $n env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'e2e %.0f'%d['e2e']['value'])"
done
