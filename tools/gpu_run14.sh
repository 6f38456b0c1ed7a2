cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_facade.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r14_pytest.log; cat gpurun_out/r14_pytest.log
timeout 600 python tools/stage_roofline.py c4 merge > gpurun_out/r14_stage_roofline_c4.md 2>&1; tail -4 gpurun_out/r14_stage_roofline_c4.md
timeout 600 python tools/c5_merge_probe.py 2>&1 | tail -3 | tee gpurun_out/r14_probe.log
