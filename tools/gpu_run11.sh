cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_facade.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r11_pytest.log; cat gpurun_out/r11_pytest.log
timeout 600 python tools/c5_merge_probe.py 2>&1 | tail -4 | tee gpurun_out/r11_probe.log
