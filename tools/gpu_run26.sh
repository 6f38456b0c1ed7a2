cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_general_merge_kernel_hub_graph 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r26_err.log | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
if not txt: print('no output'); sys.exit()
d=json.loads(txt[-1])
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'solo', d['single_frame_latency_ms']['stage_ms'])
"
tail -1 gpurun_out/r26_err.log
