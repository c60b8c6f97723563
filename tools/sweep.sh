for n in 16384 65536; do python bench.py --steps 10 --warmup 3 --cpu-seconds 0.1 --envs $n 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('envs $n env-steps/s %.0f'%d['value'],'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'e2e %.0f'%d['e2e']['value'])"
done
python tools/bench_c3.py 16384 20 | tail -1 | cut -c1-220
python tools/bench_ik.py 4096 65536
