for cfg in "1 28" "2 28" "0 28" "1 24" "1 20" "2 20"; do set -- $cfg
RCSB_LOCKSTEP=$1 RCSB_WARPS=$2 python bench.py --steps 20 --warmup 4 --cpu-seconds 0.1 --envs ${ENVS:-4096} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('lockstep $1 warps',d['config']['warps_per_cta'],'env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done
ENVS=16384; for cfg in "1 28" "1 24"; do set -- $cfg
RCSB_LOCKSTEP=$1 RCSB_WARPS=$2 python bench.py --steps 10 --warmup 4 --cpu-seconds 0.1 --envs 16384 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('16k lockstep $1 warps',d['config']['warps_per_cta'],'env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done
