cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r8_slab.log; cat gpurun_out/r8_slab.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_facade.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r8_pytest.log; cat gpurun_out/r8_pytest.log
