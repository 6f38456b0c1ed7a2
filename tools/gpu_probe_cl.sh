cd /root/repo
F3PS_MERGE_KERNEL=3 timeout 120 python tools/gpu_parity_probe.py small > gpurun_out/probe_cl_small.log 2>&1; echo "probe small rc=$?"
grep -E "DIFF|MISMATCH|Error|error" gpurun_out/probe_cl_small.log | head
tail -4 gpurun_out/probe_cl_small.log
F3PS_MERGE_KERNEL=3 timeout 120 python tools/gpu_parity_probe.py vga > gpurun_out/probe_cl_vga.log 2>&1; echo "probe vga rc=$?"
grep -E "DIFF|MISMATCH|Error|error" gpurun_out/probe_cl_vga.log | head
tail -4 gpurun_out/probe_cl_vga.log
timeout 120 python tools/gpu_parity_probe.py vga > gpurun_out/probe_vga.log 2>&1; echo "probe 1cta rc=$?"
grep -E "DIFF|MISMATCH|Error|error" gpurun_out/probe_vga.log | head
tail -4 gpurun_out/probe_vga.log
