cd /root/repo
mkdir -p gpurun_out
for cfg in "96 3" "64 4"; do
  set -- $cfg
  timeout 300 python bench.py --inflight $1 --rounds $2 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r21_err.log | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
if not txt: print('$cfg no output'); sys.exit()
d=json.loads(txt[-1])
print('batch $cfg', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e'])
"
  tail -1 gpurun_out/r21_err.log
done
