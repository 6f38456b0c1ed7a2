# sweep-pool variants of the C2 bench (no CPU legs, no C3 leg): value / e2e / ms per frame per variant
cd /root/repo
for cfg in "batch 96 4 0" "batch 96 4 8" "pipeline 72 8 1" "pipeline 72 8 2" "pipeline 96 6 1" "pipeline 64 9 1"; do
  set -- $cfg
  timeout 300 python bench.py --no-cpu --no-c3 --steps 4 --warmup 3 --pool $1 --inflight $2 --rounds $3 --expand-cluster $4 2> gpurun_out/bs.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pool $1 inflight $2 rounds $3 cluster $4: value %.1f e2e %.1f ms/frame %.3f lat %.2f launches %d stage %s' % (d['value'], d['e2e']['value'], d['ms_per_frame'], d['single_frame_latency_ms']['median'], d['gpu_launches'], d['stage_ms']))
" || tail -3 gpurun_out/bs.err
done
