set -x
cd /root/repo
timeout 300 python tools/gpu_parity_probe.py vga > gpurun_out/probe_vga.log 2>&1; echo "probe rc=$?"
tail -22 gpurun_out/probe_vga.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -2 gpurun_out/bench.log
