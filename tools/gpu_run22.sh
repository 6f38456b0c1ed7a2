cd /root/repo
mkdir -p gpurun_out
P=fast-3d-pointcloud-segmentation_b200
for mb in 3 4 2; do
  cp $P/libf3ps_mb$mb.so $P/libf3ps.so
  echo "== min blocks $mb"
  timeout 200 python tools/front_scaling_probe2.py 2>&1 | grep "k5 threads=16\|k5 threads= 1"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r22_err.log | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
if not txt: print('no output'); sys.exit()
d=json.loads(txt[-1])
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'solo expand', d['single_frame_latency_ms']['stage_ms']['expand'])
"
done
