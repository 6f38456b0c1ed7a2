cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/front_scaling_probe2.py 2>&1 | grep k5
timeout 300 python bench.py --pool batch --inflight 96 --rounds 3 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r19_err.log | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
d=json.loads(txt[-1])
print('batch96', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'lat', d['single_frame_latency_ms']['stage_ms'], 'stage', d['stage_ms'])
"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
