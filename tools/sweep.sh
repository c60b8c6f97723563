for lib in librcsb.so librcsb_r24.so librcsb_r21.so; do for n in 4096 16384 65536; do
RCSB_LIB_PATH=$PWD/robot-control-stack_b200/csrc/$lib python bench.py --steps 10 --warmup 3 --cpu-seconds 0.1 --envs $n 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$lib envs $n warps',d['config']['warps_per_cta'],'env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])"
done; done
