cd /root/repo
mkdir -p gpurun_out
for F in 48 64 96 128; do
  timeout 600 python bench.py --inflight $F --rounds 2 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('inflight', d['config']['frames_in_flight_per_gpu'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/frame', round(d['ms_per_frame'],3), 'stage', d['stage_ms'])
"
done
nproc
