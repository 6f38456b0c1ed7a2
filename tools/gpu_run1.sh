set -x
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python tools/gpu_parity_probe.py vga > gpurun_out/probe_vga.log 2>&1; echo "probe rc=$?"
tail -8 gpurun_out/probe_vga.log
python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -2 gpurun_out/bench.log
