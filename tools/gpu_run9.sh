cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q --tb=line 2>&1 | cut -c1-1500 | tail -40 > gpurun_out/r9_slab.log; cat gpurun_out/r9_slab.log
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q --tb=line 2>&1 | cut -c1-1500 | tail -5
