# C5 in slab mode on 1 GPU (world 1): sizes, stage times, does the merge stage produce merges
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --workload c5 --points 10000000 --steps 1 --warmup 1 2> gpurun_out/r10_c5_10m.err | tee gpurun_out/r10_c5_10m.json
tail -3 gpurun_out/r10_c5_10m.err
timeout 1200 python bench.py --workload c5 --points 50000000 --steps 1 --warmup 1 2> gpurun_out/r10_c5_50m.err | tee gpurun_out/r10_c5_50m.json
tail -3 gpurun_out/r10_c5_50m.err
nvidia-smi --query-gpu=memory.used --format=csv
