for lib in librcsb.so librcsb_w21.so librcsb_w28.so; do
RCSB_LIB_PATH=$PWD/robot-control-stack_b200/csrc/$lib python bench.py --steps 20 --warmup 4 --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$lib warps',d['config']['warps_per_cta'],'env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done
