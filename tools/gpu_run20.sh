cd /root/repo
mkdir -p gpurun_out
for cfg in "0 0" "48 0" "24 0" "24 32" "12 32"; do
  set -- $cfg
  timeout 300 python bench.py --pool batch --inflight 96 --rounds 3 --expand-ctas $1 --workers $2 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r20_err.log | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
if not txt: print('$cfg no output'); sys.exit()
d=json.loads(txt[-1])
print('ctas/workers $cfg', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'stage', d['stage_ms'])
"
  tail -1 gpurun_out/r20_err.log
done
