cd /root/repo
timeout 800 python -m pytest tests/test_gpu_facade.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/stage_roofline.py c4 merge > gpurun_out/stage_roofline_c4.md 2>&1; echo "c4 rc=$?"
cat gpurun_out/stage_roofline_c4.md | tail -22
