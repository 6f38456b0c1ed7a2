# final validation of the round: every GPU test, smoke, both bench arms, launch list of a C2 frame
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r23_pytest.log; cat gpurun_out/r23_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r23_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/r23_ref.err | tee gpurun_out/r23_ref.json | cut -c1-400
timeout 600 python bench.py 2>gpurun_out/r23_bench.err | tee gpurun_out/r23_bench.json | cut -c1-3000
tail -2 gpurun_out/r23_bench.err
