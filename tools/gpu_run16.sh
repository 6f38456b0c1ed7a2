cd /root/repo
mkdir -p gpurun_out
for cfg in "batch 96 3 0" "batch 128 2 0" "batch 96 3 8" "streams 32 4 0"; do
  set -- $cfg
  timeout 600 python bench.py --pool $1 --inflight $2 --rounds $3 --workers $4 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r16_err.log | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
if not txt: print('no output'); sys.exit()
d=json.loads(txt[-1])
print('$cfg', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/frame', round(d['ms_per_frame'],3), 'launches', d['gpu_launches'], 'stage', d['stage_ms'])
"
  tail -2 gpurun_out/r16_err.log
done
