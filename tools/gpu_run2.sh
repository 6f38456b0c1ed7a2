set -x
cd /root/repo
timeout 300 python tools/gpu_parity_probe.py small > gpurun_out/probe_small.log 2>&1; echo "probe small rc=$?"
tail -22 gpurun_out/probe_small.log
timeout 300 python tools/gpu_parity_probe.py vga > gpurun_out/probe_vga.log 2>&1; echo "probe rc=$?"
tail -22 gpurun_out/probe_vga.log
timeout 300 python tools/gpu_parity_probe.py vga eq > gpurun_out/probe_vga_eq.log 2>&1; echo "probe eq rc=$?"
tail -16 gpurun_out/probe_vga_eq.log
