cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "merge_batch" --tb=short 2>&1 | cut -c1-600 | tail -8
for cfg in "batch 96 3 0" "batch 64 4 0" "batch 48 6 0"; do
  set -- $cfg
  timeout 300 python bench.py --pool $1 --inflight $2 --rounds $3 --workers $4 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r18_err.log | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
if not txt: print('$cfg no output'); sys.exit()
d=json.loads(txt[-1])
print('$cfg', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/frame', round(d['ms_per_frame'],3), 'stage', d['stage_ms'])
"
  tail -2 gpurun_out/r18_err.log
done
