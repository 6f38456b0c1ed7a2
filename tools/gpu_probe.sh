cd /root/repo
timeout 300 python tools/gpu_parity_probe.py vga > gpurun_out/probe_vga.log 2>&1; echo "probe rc=$?"
grep -E "DIFF|MISMATCH|Error|error" gpurun_out/probe_vga.log | head
tail -6 gpurun_out/probe_vga.log
