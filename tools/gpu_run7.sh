# Session-4 re-verification: GPU parity tests, smoke, bench (both arms), launch list of one C2 frame.
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r7_pytest.log; cat gpurun_out/r7_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r7_smoke.log
timeout 600 python bench.py 2>gpurun_out/r7_bench.err | tee gpurun_out/r7_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/r7_ref.err | tee gpurun_out/r7_ref.json
cat > /tmp/one.py <<'PY'
import sys
sys.path.insert(0, "fast-3d-pointcloud-segmentation_b200")
import f3ps
from f3ps import synth
pts = synth.make_frame(seed=20020)
g = f3ps.Segmenter(); g.set_vccs_params(); g.set_merge_params(color_mode=0, geom_mode=1, merge_mode=1)
g.set_input(pts); g.run(0.2)
g.set_input(pts); g.run(0.2)
print(g.counts().n_merges, g.stage_ms())
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r7.csv python /tmp/one.py > gpurun_out/r7_ncu_launch.log 2>&1; echo "launch list rc=$?"
tail -2 gpurun_out/r7_ncu_launch.log
