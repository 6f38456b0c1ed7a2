# sweep-pool variants of the C2 bench (no CPU legs, no C3 leg): value / e2e / ms per frame per variant
cd /root/repo
for cfg in "96 3 16" "96 3 32" "96 3 48" "93 3 32" "64 4 32"; do
  set -- $cfg
  timeout 300 python bench.py --no-cpu --no-c3 --steps 4 --warmup 3 --inflight $1 --rounds $2 --workers $3 2> gpurun_out/bs.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('inflight $1 rounds $2 workers $3: value %.1f e2e %.1f ms/frame %.3f lat %.2f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_frame'], d['single_frame_latency_ms']['median'], d['gpu_launches']))
" || tail -3 gpurun_out/bs.err
done
